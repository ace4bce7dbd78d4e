/* tests/cpp/host_class_smoke.cpp -- the C++ host class compiles against the C ABI and links to the product library.
 * On a box without a GPU the constructor must fail loudly (no CPU fallback); with a GPU it runs one block of silence. */
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../include/SdrBatch.hpp"

int main() {
  std::printf("%s\n", sdr_batch_version());
  try {
    sdr::SdrBatch sdr(8);
    float off = sdr.setDemodMode(3, SDR_USB);
    sdr.enableAGC();
    sdr.setAGCmode(sdr::all, SDR_AGC_MEDIUM);
    sdr.setAudioFilter(std::vector<uint32_t>{1, 2}, SDR_AUDIO_2700);
    sdr.setDemodMode(std::vector<uint32_t>{}, SDR_AM); /* an empty selection selects nothing (not "all channels") */
    std::vector<int16_t> I(8 * 128, 0), Q(8 * 128, 0), out(8 * 128, 1);
    sdr.process_host(I.data(), Q.data(), 128, SDR_FMT_I16, out.data(), 128, SDR_FMT_I16, 1);
    std::printf("GPU_OK offset=%.0f mode=%d out0=%d\n", off, (int)sdr.getDemodMode(3), (int)out[0]);
    return (off == 5390.f && sdr.getDemodMode(3) == SDR_USB && sdr.getDemodMode(0) == SDR_LSB && out[0] == 0) ? 0 : 2;
  } catch (const std::runtime_error &e) {
    std::printf("NO_DEVICE %s\n", e.what());
    return 0;
  }
}
