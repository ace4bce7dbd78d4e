/* tests/cpp/aux_class_smoke.cpp -- the C++ mirrors of AudioSDRpreProcessor / AudioIQgenerator compile against the C ABI and link
 * to libsdr_aux.so.  Without a GPU the constructors must fail loudly; with one, a block of a ramp goes through both. */
#include <cstdio>
#include <vector>
#include "../../include/SdrAux.hpp"

int main() {
  std::printf("%s\n", sdr_aux_version());
  try {
    sdr::PreProcessorBatch pp(8);
    pp.setI2SerrorCompensation(sdr::all, 1);
    pp.swapIQ(3, true);
    std::vector<int16_t> I(8 * 128), Q(8 * 128), oi(8 * 128), oq(8 * 128);
    for (int i = 0; i < 8 * 128; i++) { I[i] = (int16_t)(i % 128 + 1); Q[i] = (int16_t)(-(i % 128) - 1); }
    pp.process_host(I.data(), Q.data(), 128, oi.data(), oq.data(), 128, 1);
    /* channel 0: I delayed by one (first sample = savedSample = 0); channel 3: the same, then swapped */
    bool ok = oi[0] == 0 && oi[1] == 1 && oq[0] == -1 && oq[3 * 128] == 0 && oi[3 * 128] == -1 && pp.getI2SerrorCompensation(5) == 1 &&
              !pp.getAutoI2SerrorDetectionStatus(5);
    sdr::IQGeneratorBatch g(8);
    g.setGainBalance(std::vector<uint32_t>{1, 2}, 1.5f);
    std::vector<int16_t> X(8 * 256, 0), gi(8 * 256, 7), gq(8 * 256, 7);
    for (int c = 0; c < 8; c++) X[c * 256 + 5] = 16384;
    g.process_host(X.data(), 256, gi.data(), gq.data(), 256, 2);
    /* the impulse comes out of the I rail 128 samples later (within 1 LSB), x 1.5 on channel 1 */
    ok = ok && gi[5] == 0 && (gi[133] == 16384 || gi[133] == 16383) && gi[256 + 133] > 24500 && gi[256 + 133] < 24580 && gq[0] == 0;
    sdr::GrabberBatch gr(8);
    int16_t snap[512];
    ok = ok && !gr.newDataAvailable(2) && !gr.grab(2, snap);
    std::printf("GPU_OK %d\n", (int)ok);
    return ok ? 0 : 2;
  } catch (const std::runtime_error &e) {
    std::printf("NO_DEVICE %s\n", e.what());
    return 0;
  }
}
