"""GPU tier: the CUDA library (sm_100a kernels) through the C ABI against the oracle and the reference goldens.

Bar: BIT-EXACT float32 audio and int16 pcm on every mode (the build disables FMA contraction and follows the
reference's operation order), which is far inside north_star's tolerance (1e-4 of full scale for the full
chain, 1e-6 relative for the feed-forward mixer/FIR).  NaN samples (only the setter-fuzz cases produce them)
compare equal as NaN: x86 and the GPU pick different NaN payloads.
Nothing here reads /root/reference; the oracle is oracle/libsdr_oracle.so and tests/golden/*.npz.
"""
import os

import numpy as np
import pytest

import harness
import signals as S
from test_oracle import GOLDEN, load_golden

pytestmark = pytest.mark.gpu
TOL_FULL_CHAIN = 1e-4  # north_star: max abs error of full scale; we require 0


def assert_same(a, want):
    if a.dtype == np.float32:
        nan_a, nan_w = np.isnan(a), np.isnan(want)
        assert np.array_equal(nan_a, nan_w), "NaN pattern differs"
        a = np.where(nan_a, np.float32(0), a); want = np.where(nan_w, np.float32(0), want)
        fin = np.isfinite(a) & np.isfinite(want)  # infinities must match exactly (checked bitwise below)
        err = float(np.max(np.abs(a[fin].astype(np.float64) - want[fin].astype(np.float64)))) if fin.any() else 0.0
        assert err <= TOL_FULL_CHAIN, "max abs error %g exceeds the north_star tolerance" % err
    assert harness.bits_equal(a, want), harness.describe_mismatch(a, want)


@pytest.fixture(scope="module")
def dev(cuda_lib):
    import torch
    return torch.device("cuda:0")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_reference_golden(cuda_lib, path):
    I, Q, ev, audio, pcm, status = load_golden(path)
    a = harness.run_batch(cuda_lib, I, Q, ev, chunks=(5, 2, 9))
    assert_same(a, audio)
    p = harness.run_batch(cuda_lib, I, Q, ev, chunks=(8,), out_dtype=np.int16)
    assert np.array_equal(p, pcm)


@pytest.mark.parametrize("cfg,nch,nblk", [(1, 1, 345), (2, 96, 120), (3, 64, 120), (4, 140, 90), (5, 40, 120)])
def test_cuda_matches_oracle_per_config(cuda_lib, oracle, dev, cfg, nch, nblk):
    I, Q, ev = S.make(cfg, list(range(nch)), nblk)
    o = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)
    a, b = harness.run_batch(cuda_lib, I, Q, ev, chunks=(17, 1, 40), device=dev, return_batch=True)
    assert_same(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)
    p = harness.run_batch(cuda_lib, I, Q, ev, chunks=(64,), out_dtype=np.int16, device=dev)
    assert np.array_equal(p, o["pcm"])
    # host-buffer entry point in one call: internally split into overlapped H2D / kernel / D2H chunks
    hp = harness.run_batch(cuda_lib, I, Q, ev, chunks=(nblk,), out_dtype=np.int16, device=None)
    assert np.array_equal(hp, o["pcm"])


def test_cuda_ten_seconds_usb(cuda_lib, oracle, dev):
    """BASELINE config 1 at full length: 10 s = 3446 blocks, one USB channel, in one call."""
    I, Q, ev = S.make(1, [0], S.BLOCKS_10S)
    o = oracle.run(I, Q, ev)
    a = harness.run_batch(cuda_lib, I, Q, ev, chunks=(S.BLOCKS_10S,), device=dev)
    assert_same(a, o["audio"])


def test_cuda_feed_forward_stage_isolation(cuda_lib, oracle, dev):
    """Mixer + Hilbert + combine only (NB, AGC, ALS, audio filter off): north_star's 1e-6 norm-relative bar."""
    I, Q, _ = S.make(2, list(range(8)), 60)
    ev = []
    for c in range(8):
        ev += [(c, 0, "setDemodMode", [0, 1, 2, 3, 6, 1, 0, 6][c]), (c, 0, "disableNoiseBlanker"), (c, 0, "disableAGC"),
               (c, 0, "setOutputGain", 1.0), (c, 0, "setMute", 0)]
    o = oracle.run(I, Q, ev)
    a = harness.run_batch(cuda_lib, I, Q, ev, chunks=(60,), device=dev)
    peak = np.max(np.abs(o["audio"]), axis=1, keepdims=True)
    assert np.max(np.abs(a - o["audio"]) / peak) <= 1e-6
    assert_same(a, o["audio"])


def test_cuda_float32_planes_and_gains(cuda_lib, oracle, dev):
    I, Q, ev = S.make(2, list(range(40)), 30)
    ev += [(c, 0, "setInputGain", 0.7) for c in range(0, 40, 3)] + [(c, 0, "setIQgainBalance", 1.05) for c in range(1, 40, 3)]
    If = (I.astype(np.float32) / np.float32(32767.0)).astype(np.float32)
    Qf = (Q.astype(np.float32) / np.float32(32767.0)).astype(np.float32)
    o = oracle.run(If, Qf, ev, threads=4)
    a = harness.run_batch(cuda_lib, If, Qf, ev, chunks=(4, 9), device=dev)
    assert_same(a, o["audio"])
    oi = oracle.run(I, Q, ev, threads=4)       # int16 wire format with non-unit gains: the reference's exact boundary
    ai = harness.run_batch(cuda_lib, I, Q, ev, chunks=(30,), device=dev)
    assert_same(ai, oi["audio"])


@pytest.mark.parametrize("tile,ctas,merge", [(16, 2, 0), (8, 2, 0), (8, 3, 1), (16, 2, 1)])
@pytest.mark.parametrize("cfg,nch,nblk", [(1, 1, 60), (3, 70, 60), (5, 40, 60)])
def test_cuda_short_tile_plans(cuda_lib, oracle, dev, monkeypatch, cfg, nch, nblk, tile, ctas, merge):
    """The plans for buckets without blanker and ALS (16- / 8-sample tiles, several groups per SM, merged warps): same bits."""
    from test_emu_pipeline import set_plan
    set_plan(monkeypatch, tile, ctas, merge)
    I, Q, ev = S.make(cfg, list(range(nch)), nblk)
    o = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)
    a, b = harness.run_batch(cuda_lib, I, Q, ev, chunks=(17, 1, 40), device=dev, return_batch=True)
    assert_same(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)


def test_cuda_lone_mode_switch(cuda_lib, oracle, dev):
    """setDemodMode as the only setter at a block boundary (same-class and SSB<->AM/SAM class changes)."""
    from test_emu_pipeline import lone_mode_switch_case
    I, Q, ev = lone_mode_switch_case()
    o = oracle.run(I, Q, ev, threads=4)
    a, b = harness.run_batch(cuda_lib, I, Q, ev, chunks=(30,), device=dev, return_batch=True)
    assert_same(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)


def test_cuda_state_export_import_continues_bit_exact(cuda_lib, oracle, dev):
    """Checkpoint / migration: channels moved between handles (other slot, other ring phase) continue bit for bit."""
    from test_emu_pipeline import migration_case
    I, Q, ev = S.make(4, list(range(70)), 21)
    want = oracle.run(I, Q, ev, threads=4)["audio"]
    got = migration_case(cuda_lib, I, Q, ev, device=dev)
    assert_same(got, want)


def test_cuda_setter_fuzz(cuda_lib, oracle, dev):
    rng = np.random.default_rng(4321)
    I, Q, ev = S.make(4, list(range(64)), 48)
    ev += harness.fuzz_events(rng, 64, 48, 700)
    o = oracle.run(I, Q, ev, threads=4)
    a = harness.run_batch(cuda_lib, I, Q, ev, chunks=(5, 2, 9), device=dev)
    assert_same(a, o["audio"])
    p = harness.run_batch(cuda_lib, I, Q, ev, chunks=(7,), out_dtype=np.int16, device=dev)
    assert np.array_equal(p, o["pcm"])


@pytest.mark.parametrize("cfg,total,nblk", [(2, 4096, 24), (4, 16384, 24), (3, 65536, 16), (5, 262144, 12)])
def test_cuda_full_channel_count_sampled(cuda_lib, oracle, dev, cfg, total, nblk):
    """BASELINE channel counts (short duration): every channel runs on the GPU, a sampled subset on the oracle."""
    import torch
    import audiosdr_b200 as A
    picks = S.sample_channels(cfg, total, 64)
    I, Q, ev = S.make(cfg, picks, nblk)
    o = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)
    # the full batch: sampled channels carry their own signal, the rest repeat sampled rows (any valid input will do)
    idx = np.arange(total) % len(picks)
    pos = {c: i for i, c in enumerate(picks)}
    for c in picks:
        idx[c] = pos[c]
    b = A.SdrBatch(total, _lib=cuda_lib)
    per_pick = [[tuple(e[2:]) for e in S.channel_events(cfg, src, 0)] for src in picks]
    calls = []
    for c in range(total):
        calls += [(c,) + e for e in per_pick[idx[c]]]
    b.configure(calls)
    dI = torch.from_numpy(I).to(dev)[torch.from_numpy(idx).to(dev)]
    dQ = torch.from_numpy(Q).to(dev)[torch.from_numpy(idx).to(dev)]
    out = torch.empty((total, nblk * 128), dtype=torch.float32, device=dev)
    k = 5 * 128
    b.process(dI[:, :k], dQ[:, :k], out[:, :k], n_blocks=5)
    b.process(dI[:, k:].contiguous(), dQ[:, k:].contiguous(), out[:, k:], n_blocks=nblk - 5)
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    assert_same(res[picks], o["audio"])
    # size-independent property: identical inputs + identical configuration => identical outputs, wherever they sit
    assert harness.bits_equal(res, res[np.array(picks)][idx])


def test_cuda_state_carry_long_stream_wspr(cuda_lib, oracle, dev):
    """Config-5 style drift check, shortened: 30 s of WSPR in 23 ragged calls; last 10 s compared, plus NCO phase."""
    nblk = 3 * S.BLOCKS_10S
    I, Q, ev = S.make(5, [0, 1, 2, 262143], nblk)
    o = oracle.run(I, Q, ev, threads=4)
    a, b = harness.run_batch(cuda_lib, I, Q, ev, chunks=(449, 450, 451), device=dev, return_batch=True)
    assert_same(a[:, -S.BLOCKS_10S * 128:], o["audio"][:, -S.BLOCKS_10S * 128:])
    assert_same(a, o["audio"])
    ch = oracle.Channel()
    for e in [e for e in ev if e[0] == 0]:
        ch.apply(e[2], *e[3:])
    for k in range(40):
        ch.update(I[0, k * 128:(k + 1) * 128], Q[0, k * 128:(k + 1) * 128])
    b2 = harness.run_batch(cuda_lib, I[:1, :40 * 128], Q[:1, :40 * 128], [e for e in ev if e[0] == 0], chunks=(40,), device=dev,
                           return_batch=True)[1]
    assert np.float32(b2.peek_state(0, 80)) == np.float32(ch.phases()[0])  # W_PH_SSB


def test_cuda_errors_and_launch_count(cuda_lib, dev):
    import torch
    import audiosdr_b200 as A
    b = A.SdrBatch(5, _lib=cuda_lib)
    with pytest.raises(A.SdrError):
        b.setDemodMode(0, 7)
    x = torch.zeros((5, 256), dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    n0 = b.launch_count
    b.process(x, x, y)
    assert b.launch_count > n0
    with pytest.raises(A.SdrError):
        b.process(x[:, 1:129], x[:, 1:129], y[:, :128], n_blocks=1)  # misaligned rows
    torch.cuda.synchronize()
    assert float(y.abs().max()) == 0.0  # silence in, silence out (NB on: zeros through the delay line)


def test_cuda_batched_envelope_equals_ieee_divide(cuda_lib):
    """sqrt_hack_batch (branch-free division fast path) == sqrt_hack (IEEE divide) for 2^28 bit patterns of every exponent."""
    import ctypes as C
    cuda_lib.sdrk_selftest_envelope.argtypes = [C.c_uint, C.c_uint, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
    for first, step, n in ((0, 16, 1 << 28), (0x3A000000, 1, 1 << 27), (0x00000000, 1, 1 << 24), (0x7F000000, 1, 1 << 24)):
        bad = C.c_ulonglong(12345)
        assert cuda_lib.sdrk_selftest_envelope(first, step, n, C.byref(bad)) == 0
        assert bad.value == 0, (hex(first), step, n, bad.value)


def test_cuda_als_edge_parameters(cuda_lib, oracle, dev):
    """ALS tap counts / delays at the ends of their ranges (taps up to the last array slot, delay 0 -- the one case where
    a tile cannot pre-compute its successor's first FIR sum --, histories reaching the far end of the ring), parameters
    changed mid-stream, calls of 1 to 40 blocks."""
    from test_emu_pipeline import ALS_EDGE_PARAMS, als_edge_events
    nch = 4 * len(ALS_EDGE_PARAMS) + 3
    I, Q, ev = S.make(4, list(range(nch)), 60)
    ev = [e for e in ev if not e[2].startswith("setALSfilterParams")] + als_edge_events(nch)
    o = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)
    a = harness.run_batch(cuda_lib, I, Q, ev, chunks=(3, 1, 40, 2), device=dev)
    assert_same(a, o["audio"])


@pytest.mark.parametrize("case", ["config4", "edge", "sliced", "sam", "small", "small-single-ring", "config4-largest-plan"])
def test_cuda_split_als_bucket(cuda_lib, oracle, dev, monkeypatch, case):
    """ALS buckets as two launches (the chain up to the AGC into a scratch plane, then the one-warp ALS + output post-pass):
    forced here at small channel counts; the 16 384-channel case of test_cuda_full_channel_count_sampled takes it by itself."""
    from test_emu_pipeline import ALS_EDGE_PARAMS, ALS_SMALL_PARAMS, als_edge_events, als_small_events
    monkeypatch.setenv("SDR_ALS_SPLIT", "1")
    chunks = (17, 1, 40)
    if case == "config4":
        I, Q, ev = S.make(4, list(range(140)), 90); chunks = (17, 1, 72)
    elif case == "edge":
        nch = 4 * len(ALS_EDGE_PARAMS) + 3
        I, Q, ev = S.make(4, list(range(nch)), 60)
        ev = [e for e in ev if not e[2].startswith("setALSfilterParams")] + als_edge_events(nch)
        chunks = (3, 1, 40, 16)
    elif case.startswith("small"):  # the doubled input ring with every tap count / delay it admits; the same on a single ring
        if case == "small-single-ring":
            monkeypatch.setenv("SDR_ALS_NO_MIRROR", "1")
        nch = 6 * len(ALS_SMALL_PARAMS)
        I, Q, ev = S.make(4, list(range(nch)), 60)
        ev = [e for e in ev if not e[2].startswith("setALSfilterParams")] + als_small_events(nch)
        chunks = (3, 1, 40, 16)
    elif case == "config4-largest-plan":
        monkeypatch.setenv("SDR_ALS_FULL_ROWS", "1")
        I, Q, ev = S.make(4, list(range(140)), 90); chunks = (17, 1, 72)
    elif case == "sliced":  # 1 MB of scratch: slices of 16 to 21 blocks
        monkeypatch.setenv("SDR_ALS_SCRATCH_MB", "1")
        I, Q, ev = S.make(4, list(range(70)), 100); chunks = (100,)
    else:  # SAM + ALS, the chain on the merged 7-warp plan
        monkeypatch.setenv("SDR_TILE_ENV", "16"); monkeypatch.setenv("SDR_CTAS_PER_SM", "3"); monkeypatch.setenv("SDR_IN_DEPTH", "1")
        I, Q, ev = S.make(3, list(range(70)), 60); chunks = (17, 1, 42)
        ev += [(c, 0, "enableALSfilter") for c in range(70)] + [(c, 0, "setALSfilterNotch" if c % 2 else "setALSfilterPeak") for c in range(70)]
    o = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)
    a, b = harness.run_batch(cuda_lib, I, Q, ev, chunks=chunks, device=dev, return_batch=True)
    assert_same(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)
    p = harness.run_batch(cuda_lib, I, Q, ev, chunks=(sum(chunks),), out_dtype=np.int16, device=dev)
    assert np.array_equal(p, o["pcm"])


@pytest.mark.parametrize("split", ["0", "1"])
def test_cuda_every_bucket_kind_in_one_handle(cuda_lib, oracle, dev, monkeypatch, split):
    """12 buckets = 12 launches on 12 forked streams per call (USB / AM / SAM x blanker x ALS)."""
    from test_emu_pipeline import every_bucket_case
    monkeypatch.setenv("SDR_ALS_SPLIT", split)
    I, Q, ev = every_bucket_case(40)
    o = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)
    a, b = harness.run_batch(cuda_lib, I, Q, ev, chunks=(5, 35), device=dev, return_batch=True)
    assert_same(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)


def test_cuda_als_forms_alternate_between_calls(cuda_lib, oracle, dev, monkeypatch):
    """One-launch and two-launch form of an ALS bucket from call to call on one handle: same state words, same bits."""
    import torch
    import audiosdr_b200 as A
    I, Q, ev = S.make(4, list(range(100)), 60)
    want = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)["audio"]
    monkeypatch.setenv("SDR_MAP_SEARCH", "1")  # plan at every call
    h = A.SdrBatch(100, _lib=cuda_lib)
    h.configure([(e[0], e[2]) + tuple(e[3:]) for e in ev])
    dI, dQ = torch.from_numpy(I).to(dev), torch.from_numpy(Q).to(dev)
    out = torch.empty((100, 60 * 128), dtype=torch.float32, device=dev)
    b0 = 0
    for k, nb in enumerate((3, 25, 1, 4, 27)):
        monkeypatch.setenv("SDR_ALS_SPLIT", str(k & 1))
        sl = slice(b0 * 128, (b0 + nb) * 128)
        h.process(dI[:, sl].contiguous(), dQ[:, sl].contiguous(), out[:, sl], n_blocks=nb)
        b0 += nb
    torch.cuda.synchronize()
    assert_same(out.cpu().numpy(), want)


def test_cuda_inrange_divide_equals_ieee_divide(cuda_lib):
    """div_inrange (the SAM PLL's division: fast path without range check / slow-path branch) == IEEE divide for 2^30
    operand pairs covering every exponent pair of [2^-60, 2^60], both signs, powers of two and their neighbours."""
    import ctypes as C
    cuda_lib.sdrk_selftest_divide.argtypes = [C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
    for seed in (1, 0xD1B54A32D192ED03):
        bad = C.c_ulonglong(12345)
        assert cuda_lib.sdrk_selftest_divide(seed, 1 << 29, C.byref(bad)) == 0
        assert bad.value == 0, (seed, bad.value)


def test_cuda_ragged_shapes_and_pitches(cuda_lib, oracle, dev):
    """Edge cases: 1 and 33 channels (partly filled groups), one block per call, row pitch larger than the call,
    planes that are views into a bigger buffer, int16 in / float32 out and the reverse."""
    import torch
    import audiosdr_b200 as A
    for nch in (1, 33):
        I, Q, ev = S.make(4, list(range(5, 5 + nch)), 9)
        o = oracle.run(I, Q, ev, threads=2)
        b = A.SdrBatch(nch, _lib=cuda_lib)
        b.configure([(e[0], e[2]) + tuple(e[3:]) for e in ev])
        big_i = torch.zeros((nch, 9 * 128 + 256), dtype=torch.int16, device=dev)
        big_q = torch.zeros_like(big_i)
        big_i[:, 128:128 + 9 * 128] = torch.from_numpy(I).to(dev)
        big_q[:, 128:128 + 9 * 128] = torch.from_numpy(Q).to(dev)
        out = torch.full((nch, 9 * 128 + 64), 7.0, dtype=torch.float32, device=dev)
        for k in range(9):  # one block per call, views with offset 128 samples (256 B: still 16-byte aligned) and long pitch
            a = 128 + k * 128
            b.process(big_i[:, a:a + 128], big_q[:, a:a + 128], out[:, k * 128:(k + 1) * 128], n_blocks=1)
        torch.cuda.synchronize()
        res = out.cpu().numpy()
        assert_same(res[:, :9 * 128], o["audio"])
        assert np.all(res[:, 9 * 128:] == 7.0)  # nothing written past the call
    # float32 in / int16 out
    I, Q, ev = S.make(2, list(range(40)), 16)
    If = (I.astype(np.float32) / np.float32(32767.0)).astype(np.float32); Qf = (Q.astype(np.float32) / np.float32(32767.0)).astype(np.float32)
    o = oracle.run(If, Qf, ev, threads=4)
    p = harness.run_batch(cuda_lib, If, Qf, ev, chunks=(16,), out_dtype=np.int16, device=dev)
    assert np.array_equal(p, o["pcm"])


def test_cuda_large_single_call(cuda_lib, oracle, dev):
    """A long single call (BASELINE's 10 s = 3446 blocks) on a few mixed channels, and the same stream cut into
    1-block calls: identical bits (state carry through HBM at every call boundary)."""
    I, Q, ev = S.make(4, list(range(14)), 800)
    o = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)
    a = harness.run_batch(cuda_lib, I, Q, ev, chunks=(800,), device=dev)
    assert_same(a, o["audio"])
    b = harness.run_batch(cuda_lib, I[:, :60 * 128], Q[:, :60 * 128], ev, chunks=(1,), device=dev)
    assert_same(b, o["audio"][:, :60 * 128])


def test_cuda_host_stream_of_submitted_calls(cuda_lib, oracle, dev):
    """sdr_batch_submit_host / wait_host: a stream of calls queued back to back from pinned host planes (int16 in and out), ragged
    lengths (odd chunk counts flip the staging rotation), one setter between two submits, one synchronous call in the middle --
    the audio is the oracle's, bit for bit, once wait_host() has returned."""
    import torch
    import audiosdr_b200 as A
    nblk = 150
    I, Q, ev = S.make(4, list(range(70)), nblk)
    ev = [e for e in ev if e[1] == 0] + [(3, 64, "setDemodMode", 2), (40, 64, "setAGCthreshold", -30.0)]
    o = oracle.run(I, Q, ev, threads=os.cpu_count() or 1)
    hI, hQ = torch.from_numpy(I).pin_memory(), torch.from_numpy(Q).pin_memory()
    hO = torch.zeros((70, nblk * 128), dtype=torch.int16).pin_memory()
    nI, nQ, nO = hI.numpy(), hQ.numpy(), hO.numpy()
    b = A.SdrBatch(70, _lib=cuda_lib)
    b.configure([(None if e[0] == 0xFFFFFFFF else e[0], e[2]) + tuple(e[3:]) for e in ev if e[1] == 0])
    pos = 0
    for k, n in enumerate((17, 16, 31, 50, 1, 35)):
        if pos == 64:
            b.configure([(3, "setDemodMode", 2), (40, "setAGCthreshold", -30.0)])
        a, z = pos * 128, (pos + n) * 128
        if k == 4:
            b.process_host(nI[:, a:z], nQ[:, a:z], nO[:, a:z])  # synchronous call inside the stream
        else:
            t = b.submit_host(nI[:, a:z], nQ[:, a:z], nO[:, a:z])
            assert t == k + 1  # tickets count the host calls of the handle, synchronous ones included
        if k == 2:  # one call at a time: the first call's audio is there while the third may still be in flight
            b.wait_host(ticket=1)
            assert np.array_equal(nO[:, :17 * 128], o["pcm"][:, :17 * 128])
        pos += n
    assert pos == nblk
    b.wait_host(ticket=6)
    assert np.array_equal(nO, o["pcm"])
    b.wait_host()
    b.wait_host()  # idempotent
    with pytest.raises(A.SdrError):
        b.wait_host(ticket=7)  # no such call yet
    assert np.array_equal(nO, o["pcm"]), harness.describe_mismatch(nO.astype(np.float32), o["pcm"].astype(np.float32))
