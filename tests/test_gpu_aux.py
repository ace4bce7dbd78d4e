"""GPU tier, SURVEY 8f rows 2 and 4: the batched pre-processor and I/Q generator (audiosdr_b200/libsdr_aux.so, through the C ABI
of include/sdr_aux.h) against the oracle (oracle/sdr_aux_oracle.c) and the committed golden outputs of the unmodified
reference.  Integer outputs: every comparison is exact."""
import os

import numpy as np
import pytest

import aux_signals as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def aux():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    from audiosdr_b200 import aux as m
    assert os.path.exists(m.lib_path()), "libsdr_aux.so missing: __graft_entry__.build() must run before the GPU tier"
    m.load_library()
    return m


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def run_pp(aux, I, Q, events, chunks=(7, 1, 30), host=False):
    """Stream the planes through PreProcessorBatch in ragged calls, applying setter events before their block."""
    import torch
    nch, ns = I.shape
    nb = ns // 128
    p = aux.PreProcessorBatch(nch)
    ev = sorted(events, key=lambda e: e[1])
    oi, oq = np.empty_like(I), np.empty_like(Q)
    if not host:
        dI, dQ = _dev(I), _dev(Q)
        dOi, dOq = torch.empty_like(dI), torch.empty_like(dQ)
    pos = ei = k = 0
    while pos < nb:
        while ei < len(ev) and ev[ei][1] <= pos:
            e = ev[ei]; ei += 1
            getattr(p, e[2])(*([e[0]] + list(e[3:])))
        nxt = ev[ei][1] if ei < len(ev) else nb
        sz = min(chunks[k % len(chunks)], nb - pos, max(nxt - pos, 1)); k += 1
        z = slice(pos * 128, (pos + sz) * 128)
        if host:
            a, b = np.empty((nch, sz * 128), np.int16), np.empty((nch, sz * 128), np.int16)
            p.process_host(I[:, z], Q[:, z], a, b, n_blocks=sz)
            oi[:, z], oq[:, z] = a, b
        else:
            p.process(dI[:, z], dQ[:, z], dOi[:, z], dOq[:, z], n_blocks=sz)
        pos += sz
    while ei < len(ev):
        e = ev[ei]; ei += 1
        getattr(p, e[2])(*([e[0]] + list(e[3:])))
    if not host:
        torch.cuda.synchronize()
        oi, oq = dOi.cpu().numpy(), dOq.cpu().numpy()
    st = np.array([[s.auto_detect, s.correction, s.failure_count, s.success_count, s.saved_sample, s.swap] for s in p.status()], np.int32)
    launches = p.launch_count
    p.close()
    return oi, oq, st, launches


def run_iq(aux, X, events, chunks=(5, 1, 64), host=False):
    import torch
    nch, ns = X.shape
    nb = ns // 128
    g = aux.IQGeneratorBatch(nch)
    ev = sorted(events, key=lambda e: e[1])
    oi, oq = np.empty_like(X), np.empty_like(X)
    if not host:
        dX = _dev(X)
        dOi, dOq = torch.empty_like(dX), torch.empty_like(dX)
    pos = ei = k = 0
    while pos < nb:
        while ei < len(ev) and ev[ei][1] <= pos:
            e = ev[ei]; ei += 1
            g.setGainBalance(e[0], e[3])
        nxt = ev[ei][1] if ei < len(ev) else nb
        sz = min(chunks[k % len(chunks)], nb - pos, max(nxt - pos, 1)); k += 1
        z = slice(pos * 128, (pos + sz) * 128)
        if host:
            a, b = np.empty((nch, sz * 128), np.int16), np.empty((nch, sz * 128), np.int16)
            g.process_host(X[:, z], a, b, n_blocks=sz)
            oi[:, z], oq[:, z] = a, b
        else:
            g.process(dX[:, z], dOi[:, z], dOq[:, z], n_blocks=sz)
        pos += sz
    if not host:
        torch.cuda.synchronize()
        oi, oq = dOi.cpu().numpy(), dOq.cpu().numpy()
    launches = g.launch_count
    g.close()
    return oi, oq, launches


# ---------------------------------------------------------------- I/Q generator
def test_generator_matches_golden(aux):
    g = np.load(os.path.join(GOLD, "aux_iq.npz"))
    nch, nb = int(g["n_channels"]), int(g["n_blocks"])
    X = S.iq_case(nch, nb)
    oi, oq, launches = run_iq(aux, X, S.iq_events(nch, nb))
    assert launches > 0
    assert np.array_equal(oi, g["I_out"]) and np.array_equal(oq, g["Q_out"])


@pytest.mark.parametrize("chunks", [(1,), (2, 3), (33,), (100,)])
def test_generator_matches_oracle_any_call_shape(aux, chunks):
    """1-block calls (history shorter than the filter), calls that straddle the 4096-sample segments, one big call."""
    from oracle import aux_lib as A
    X = S.iq_case(37, 100, seed=21)
    ev = S.iq_events(37, 100)
    want = A.run("iq", (X,), ev)
    oi, oq, _ = run_iq(aux, X, ev, chunks=chunks)
    assert np.array_equal(oi, want[0]) and np.array_equal(oq, want[1])


def test_generator_host_planes_and_wrap(aux):
    from oracle import aux_lib as A
    X = S.iq_case(21, 48, seed=4)
    ev = S.iq_events(21, 48) + [(0, 5, "setGainBalance", 40000.0), (7, 9, "setGainBalance", -1.5)]
    want = A.run("iq", (X,), ev)
    oi, oq, _ = run_iq(aux, X, ev, chunks=(16, 7), host=True)
    assert np.array_equal(oi, want[0]) and np.array_equal(oq, want[1])
    assert (np.abs(want[1][3].astype(int)) > 30000).any()  # the driven channel really reaches the int16 wrap region


def test_generator_many_channels_properties(aux):
    """4096 channels x 256 blocks (the bench shape): sampled channels against the oracle, every channel through
    size-independent properties: I is the input delayed by 128 samples within 1 LSB, and Q is odd in the input."""
    import torch
    from oracle import aux_lib as A
    nch, nb = 4096, 256
    base = S.iq_case(64, nb, seed=8)
    base = np.clip(base.astype(np.int32), -32767, 32767).astype(np.int16)   # symmetric range so that -X is exact
    X = np.tile(base, (nch // 64, 1))
    X[1::2] = -X[1::2]
    g = aux.IQGeneratorBatch(nch)
    dX = _dev(X); dOi, dOq = torch.empty_like(dX), torch.empty_like(dX)
    g.process(dX, dOi, dOq, n_blocks=nb)
    torch.cuda.synchronize()
    oi, oq = dOi.cpu().numpy(), dOq.cpu().numpy()
    pick = [0, 1, 63, 64, 2047, 4095]
    want = A.run("iq", (X[pick],), [])
    assert np.array_equal(oi[pick], want[0]) and np.array_equal(oq[pick], want[1])
    assert np.max(np.abs(oi[:, 128:].astype(np.int32) - X[:, :-128].astype(np.int32))) <= 1
    # rows 64 apart carry the same input: identical outputs wherever a CTA landed
    assert np.array_equal(oq[:64], oq[64:128]) and np.array_equal(oq[:64], oq[-64:]) and np.array_equal(oi[:64], oi[-64:])
    # truncation toward zero is odd-symmetric, so H(-x) == -H(x) exactly: odd rows carry -x of ... their own base row; check
    # against the oracle fed with +x
    wp = A.run("iq", ((-X[[1, 3]].astype(np.int32)).astype(np.int16),), [])
    assert np.array_equal(oq[[1, 3]], -wp[1]) and np.array_equal(oi[[1, 3]], -wp[0])
    g.close()


# ---------------------------------------------------------------- pre-processor
def test_preprocessor_matches_golden(aux):
    g = np.load(os.path.join(GOLD, "aux_pp.npz"))
    nch, nb = int(g["n_channels"]), int(g["n_blocks"])
    I, Q = S.pp_case(nch, nb)
    oi, oq, st, launches = run_pp(aux, I, Q, S.pp_events(nch, nb), chunks=(40, 1, 133))
    assert launches > 0
    assert np.array_equal(S.block_crcs(oi, oq), g["crc"])
    assert np.array_equal(oi[:, :32 * 128], g["I_head"]) and np.array_equal(oq[:, :32 * 128], g["Q_head"])
    assert np.array_equal(st, g["status"][:, :6])


@pytest.mark.parametrize("chunks", [(1,), (3, 1, 20), (300,)])
def test_preprocessor_matches_oracle_any_call_shape(aux, chunks):
    from oracle import aux_lib as A
    nch, nb = 45, 300
    I, Q = S.pp_case(nch, nb, seed=99)
    ev = S.pp_events(nch, nb) + [(3, 10, "setI2SerrorCompensation", 1), (3, 11, "setI2SerrorCompensation", -1), (3, 12, "swapIQ", 1),
                                 (2, 299, "stopAutoI2SerrorDetection"), (1, 300, "swapIQ", 1)]
    want = A.run("pp", (I, Q), ev)
    oi, oq, st, _ = run_pp(aux, I, Q, ev, chunks=chunks)
    assert np.array_equal(oi, want[0]) and np.array_equal(oq, want[1])
    assert np.array_equal(st, want[2][:, :6])


def test_preprocessor_feed_forward_only_and_host_planes(aux):
    """No detector anywhere: only the HBM-bound copy kernel runs.  All corrections x swap, host planes."""
    from oracle import aux_lib as A
    nch, nb = 24, 64
    I, Q = S.pp_case(nch, nb, seed=12)
    ev = []
    for c in range(nch):
        ev.append((c, 0, "setI2SerrorCompensation", (c % 3) - 1))
        ev.append((c, 0, "swapIQ", (c // 3) % 2))
        ev.append((c, 20 + c, "setI2SerrorCompensation", ((c + 1) % 3) - 1))
    want = A.run("pp", (I, Q), ev)
    for host in (False, True):
        oi, oq, st, _ = run_pp(aux, I, Q, ev, chunks=(9, 2), host=host)
        assert np.array_equal(oi, want[0]) and np.array_equal(oq, want[1])
        assert np.array_equal(st, want[2][:, :6])


def test_preprocessor_many_channels(aux):
    """4096 channels x 64 blocks with the detector running on every channel, then 64 more blocks after it was stopped on
    half of them: sampled channels against the oracle; identity on every untouched (correction 0, no swap) channel."""
    import torch
    from oracle import aux_lib as A
    nch, nb = 4096, 128
    bI, bQ = S.pp_case(64, nb, seed=31)
    I, Q = np.tile(bI, (nch // 64, 1)), np.tile(bQ, (nch // 64, 1))
    p = aux.PreProcessorBatch(nch)
    p.startAutoI2SerrorDetection()
    dI, dQ = _dev(I), _dev(Q)
    dOi, dOq = torch.empty_like(dI), torch.empty_like(dQ)
    h = nb // 2 * 128
    p.process(dI[:, :h], dQ[:, :h], dOi[:, :h], dOq[:, :h])
    half = np.arange(0, nch, 2, dtype=np.uint32)
    p.stopAutoI2SerrorDetection(half)
    p.process(dI[:, h:], dQ[:, h:], dOi[:, h:], dOq[:, h:])
    torch.cuda.synchronize()
    oi, oq = dOi.cpu().numpy(), dOq.cpu().numpy()
    pick = [0, 1, 2, 3, 9, 10, 65, 2050, 4094, 4095]
    ev = [(None, 0, "startAutoI2SerrorDetection")] + [(k, nb // 2, "stopAutoI2SerrorDetection") for k, c in enumerate(pick) if c % 2 == 0]
    want = A.run("pp", (I[pick], Q[pick]), ev)
    assert np.array_equal(oi[pick], want[0]) and np.array_equal(oq[pick], want[1])
    st = p.status(pick)
    assert [s.correction for s in st] == list(want[2][:, 1]) and [s.auto_detect for s in st] == list(want[2][:, 0])
    assert np.array_equal(oi[64:128], oi[:64]) and np.array_equal(oq[-64:], oq[:64])
    p.close()


def test_aux_error_behaviour(aux):
    import torch
    p = aux.PreProcessorBatch(8)
    with pytest.raises(aux.AuxError):
        p.setI2SerrorCompensation(0, 2)
    with pytest.raises(aux.AuxError):
        p.swapIQ(8, True)
    a = torch.zeros((8, 256), dtype=torch.int16, device="cuda:0")
    with pytest.raises(aux.AuxError):
        p.process(a, a, a, a)            # outputs alias the inputs
    b, c, d = torch.zeros_like(a), torch.zeros_like(a), torch.zeros_like(a)
    with pytest.raises(aux.AuxError):
        p.process(a[:, 1:129], b[:, 1:129], c[:, :128], d[:, :128], n_blocks=1)   # misaligned
    p.process(a, b, c, d)
    g = aux.IQGeneratorBatch(8)
    with pytest.raises(aux.AuxError):
        g.setGainBalance(9, 1.0)
    with pytest.raises(aux.AuxError):
        g.process(a, a, c)
    g.process(a, c, d)
    torch.cuda.synchronize()
    assert int(c.abs().max()) == 0 and int(d.abs().max()) == 0


def test_cpp_host_classes_run_on_the_gpu(aux, tmp_path):
    """include/SdrBatch.hpp and include/SdrAux.hpp driven from C++ on the device (the CPU tier only proves they link)."""
    import subprocess
    from audiosdr_b200 import build
    for src, lib, tag in (("host_class_smoke.cpp", build.build_library(), "sdr_batch"), ("aux_class_smoke.cpp", build.build_aux_library(), "sdr_aux")):
        exe = str(tmp_path / src.replace(".cpp", ""))
        subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", src), "-o", exe, "-L" + os.path.dirname(lib), "-l" + tag,
                        "-Wl,-rpath," + os.path.dirname(lib)], check=True)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "GPU_OK" in r.stdout and "NO_DEVICE" not in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("chunks", [(1,), (2,), (1, 2, 3), (5, 4)])
def test_grabber_matches_oracle(aux, chunks):
    """AudioGrabberComplex256: after every call, grab() must hand out what the reference's object would
    after the same number of update() calls -- including 'nothing yet' and the new-data flag."""
    import torch
    from oracle import aux_lib as A
    nch, nb = 37, 23
    I, Q = S.pp_case(nch, nb, seed=77)
    dI, dQ = _dev(I), _dev(Q)
    g = aux.GrabberBatch(nch)
    pos = k = 0
    while pos < nb:
        sz = min(chunks[k % len(chunks)], nb - pos); k += 1
        g.process(dI[:, pos * 128:(pos + sz) * 128], dQ[:, pos * 128:(pos + sz) * 128], n_blocks=sz)
        pos += sz
        want, flags = A.grab_run(I[:, :pos * 128], Q[:, :pos * 128])
        if k % 2 == 0 or pos == nb:   # grab on some calls only: the flag must survive the calls in between
            pick = [0, 5, nch - 1]
            fresh = [g.newDataAvailable(c) for c in pick]
            got = g.grab(pick)
            if flags[0, 1]:
                assert got is not None and np.array_equal(got.astype(np.int32), want[pick])
                assert not any(g.newDataAvailable(c) for c in pick)
            else:
                assert got is None and not any(fresh)
    allc = g.grab()
    want, _ = A.grab_run(I, Q)
    assert np.array_equal(allc.astype(np.int32), want)
    out = torch.zeros((nch, 512), dtype=torch.int16, device="cuda:0")
    assert g.L.sdr_grabber_grab_device(g.h, out.data_ptr(), None) == 1
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().astype(np.int32), want)
    g.close()


def test_grabber_spectrum_matches_oracle(aux):
    """Spectrum tap: power of the 256-point FFT of every channel's snapshot, bit for bit the oracle's
    operation network; nothing before the first snapshot; the new-data flags are not consumed."""
    import torch
    from oracle import aux_lib as A
    nch, nb = 70, 7
    I, Q = S.pp_case(nch, nb, seed=5)
    g = aux.GrabberBatch(nch)
    assert g.spectrum() is None
    g.process(_dev(I[:, :128]), _dev(Q[:, :128]), n_blocks=1)
    assert g.spectrum([0, 3]) is None                       # one block: no pair has completed yet
    g.process(_dev(I[:, 128:]), _dev(Q[:, 128:]), n_blocks=nb - 1)
    snap, _ = A.grab_run(I, Q)
    want = A.grab_spectrum(snap.astype(np.int16))
    got = g.spectrum()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    pick = [69, 0, 17]
    assert np.array_equal(g.spectrum(pick).view(np.uint32), want[pick].view(np.uint32))
    dp = torch.empty((nch, 256), dtype=torch.float32, device="cuda:0")
    assert g.spectrum_device(dp)
    torch.cuda.synchronize()
    assert np.array_equal(dp.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert g.newDataAvailable(0)                            # a spectrum is not a grab()
    g.close()
