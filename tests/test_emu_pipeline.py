"""CPU tier: pipeline LOGIC of the product sources, run through the host emulation scaffold (tests/emu).

The kernel's role bodies (audiosdr_b200/csrc/sdr_pipeline.cuh) and the whole host layer (sdr_host.cpp) are
compiled for the host and stepped warp by warp with the kernel's barrier structure; results must equal the
oracle bit for bit.  This proves delays, ring slots, state carry across calls, setter replay and grouping;
the GPU tier (test_gpu_parity.py) proves the CUDA build itself.
"""
import os

import numpy as np
import pytest

import harness
import signals as S
from test_oracle import GOLDEN, load_golden


SCHEDULES = ["lockstep", "lockstep-reversed", "producers", "consumers", "random:1", "random:2/late", "producers/late"]


def set_schedule(monkeypatch, sched):
    """The stages of a group only obey the hand-over rules of the launch's plan (sdr_lay.h): every order those rules allow
    must give the same bits.  tests/emu runs them in lock step (stage order inside a step either way), with producers as
    far ahead as the rules allow, with consumers first, and in random order."""
    sched, _, landing = sched.partition("/")
    monkeypatch.setenv("SDR_EMU_ASYNC", landing or "early")  # asynchronous copies land at the request, or as late as their wait
    monkeypatch.setenv("SDR_EMU_REVERSE", "1" if sched == "lockstep-reversed" else "0")
    monkeypatch.setenv("SDR_EMU_SCHED", "lockstep" if sched.startswith("lockstep") else sched)


@pytest.mark.parametrize("sched", SCHEDULES)
@pytest.mark.parametrize("cfg,nch,nblk", [(1, 1, 40), (2, 40, 30), (3, 6, 30), (4, 70, 24), (5, 5, 30)])
def test_emulated_pipeline_matches_oracle(oracle, emu_lib, monkeypatch, cfg, nch, nblk, sched):
    set_schedule(monkeypatch, sched)
    I, Q, ev = S.make(cfg, list(range(nch)), nblk)
    o = oracle.run(I, Q, ev, threads=4)
    a, b = harness.run_batch(emu_lib, I, Q, ev, chunks=(7, 1, 13), return_batch=True)
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)
    p = harness.run_batch(emu_lib, I, Q, ev, chunks=(3, 11), out_dtype=np.int16)
    assert np.array_equal(p, o["pcm"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_emulated_pipeline_matches_reference_golden(emu_lib, path):
    I, Q, ev, audio, pcm, status = load_golden(path)
    a = harness.run_batch(emu_lib, I, Q, ev, chunks=(5, 2, 9))
    assert harness.bits_equal(a, audio), harness.describe_mismatch(a, audio)


def test_emulated_float32_planes(oracle, emu_lib):
    I, Q, ev = S.make(2, list(range(33)), 20)
    ev += [(c, 0, "setInputGain", 0.7) for c in range(0, 33, 3)] + [(c, 0, "setIQgainBalance", 1.05) for c in range(1, 33, 3)]
    If = (I.astype(np.float32) / np.float32(32767.0)).astype(np.float32)
    Qf = (Q.astype(np.float32) / np.float32(32767.0)).astype(np.float32)
    o = oracle.run(If, Qf, ev, threads=4)
    a = harness.run_batch(emu_lib, If, Qf, ev, chunks=(4, 9))
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])


def test_emulated_setter_fuzz(oracle, emu_lib):
    rng = np.random.default_rng(1234)
    I, Q, ev = S.make(4, list(range(48)), 40)
    ev += harness.fuzz_events(rng, 48, 40, 500)
    o = oracle.run(I, Q, ev, threads=4)
    a = harness.run_batch(emu_lib, I, Q, ev, chunks=(5, 2, 9))
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])


def test_emulated_pll_general_path_alone(oracle, monkeypatch):
    """The SAM PLL evaluates the tracking case as straight-line code behind a warp vote and everything else through the
    reference's general control flow; both must give the oracle's bits.  The default run takes whichever applies per
    sample; here every vote is made to fail, so the general path computes every sample (a fresh library instance: the
    switch is read once)."""
    import ctypes, shutil, subprocess, tempfile
    from audiosdr_b200 import api
    monkeypatch.setenv("SDR_EMU_VOTE_FAILS", "1")
    emu_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
    subprocess.run(["make", "-s", "-C", emu_dir], check=True)
    with tempfile.TemporaryDirectory() as d:
        dst = os.path.join(d, "libsdr_emu_general.so")  # a copy under another name is a separate instance for the loader
        shutil.copy(os.path.join(emu_dir, "libsdr_emu.so"), dst)
        lib = api._bind(ctypes.CDLL(dst))
        I, Q, ev = S.make(3, list(range(8)), 30)
        o = oracle.run(I, Q, ev, threads=4)
        a = harness.run_batch(lib, I, Q, ev, chunks=(7, 23))
        assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])


ALS_EDGE_PARAMS = [(126, 0.5, 3), (125, 0.3, 0), (128, 0.25, 1), (124, 0.5, 5), (1, 0.5, 0), (4, 0.6, 1), (55, 0.5, 3), (97, 0.1, 19), (60, 0.5, 69)]


def als_edge_events(nch):
    """ALS tap counts / delays at the ends of their ranges (C:393-398): taps up to the last array slot, no delay at all
    (the one case where a tile cannot pre-compute its successor's first sum), a history reaching the far end of the ring."""
    ev = []
    for c in range(nch):
        m, lam, d = ALS_EDGE_PARAMS[c % len(ALS_EDGE_PARAMS)]
        ev += [(c, 0, "setALSfilterParams", m, lam, d), (c, 0, "enableALSfilter"), (c, 0, "setALSfilterAdaptive" if c % 4 else "setALSfilterStatic")]
        ev += [(c, 0, "setALSfilterNotch" if c % 2 else "setALSfilterPeak")]
        if c % 3 == 0:
            ev += [(c, 7, "setALSfilterParams", ALS_EDGE_PARAMS[(c + 1) % len(ALS_EDGE_PARAMS)][0], 0.4, c % 2)]
    return ev


def test_emulated_als_edge_parameters(oracle, emu_lib):
    nch = 2 * len(ALS_EDGE_PARAMS)
    I, Q, ev = S.make(4, list(range(nch)), 14)
    ev = [e for e in ev if not e[2].startswith("setALSfilterParams")] + als_edge_events(nch)
    o = oracle.run(I, Q, ev, threads=4)
    a = harness.run_batch(emu_lib, I, Q, ev, chunks=(3, 1, 6, 4))
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])


def test_argument_errors(emu_lib):
    import audiosdr_b200 as A
    b = A.SdrBatch(3, _lib=emu_lib)
    with pytest.raises(A.SdrError):
        b.setDemodMode(0, 9)
    with pytest.raises(A.SdrError):
        b.setAudioFilter(0, 42)
    with pytest.raises(A.SdrError):
        b.set([7], "setMute", 1)
    with pytest.raises(A.SdrError):
        b.setALSfilterParams(0, 128, 0.5, 3)  # reaches before the reference's ring
    I = np.zeros((3, 100), np.int16)  # not a whole block
    with pytest.raises(A.SdrError):
        b.process_host(I, I, n_blocks=0)


def test_channel_results_do_not_depend_on_batch_position(oracle, emu_lib):
    """Shard invariance: a channel gives the same bits whatever lane / group / batch hosts it."""
    I, Q, ev = S.make(4, list(range(50)), 12)
    full = harness.run_batch(emu_lib, I, Q, ev, chunks=(12,))
    perm = np.random.default_rng(3).permutation(50)[:17]
    sub_ev = []
    for new, old in enumerate(perm):
        sub_ev += [(new,) + tuple(e[1:]) for e in ev if e[0] == old]
    sub = harness.run_batch(emu_lib, I[perm], Q[perm], sub_ev, chunks=(5, 7))
    assert harness.bits_equal(sub, full[perm])


def lone_mode_switch_case():
    """A setDemodMode that is the ONLY setter at its block boundary (regression: it used to skip the device
    configuration upload): same-class switches (LSB->USB, USB->CW_USB, AM->SAM) and class changes (SSB<->AM/SAM)."""
    I, Q, ev = S.make(4, list(range(14)), 30)
    ev = [e for e in ev if e[1] == 0]
    switches = [(0, 5, 1), (1, 5, 3), (2, 6, 4), (3, 6, 0), (4, 7, 5), (5, 7, 6), (6, 8, 2), (7, 9, 5), (8, 9, 4), (9, 11, 1),
                (0, 13, 4), (2, 14, 5), (2, 17, 1), (10, 19, 6), (11, 21, 0), (12, 23, 4), (13, 25, 3)]
    ev += [(c, blk, "setDemodMode", m) for c, blk, m in switches]
    return I, Q, ev


def test_emulated_lone_mode_switch(oracle, emu_lib):
    I, Q, ev = lone_mode_switch_case()
    o = oracle.run(I, Q, ev, threads=4)
    a, b = harness.run_batch(emu_lib, I, Q, ev, chunks=(30,), return_batch=True)
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)


def test_every_hand_over_rule_is_needed(oracle, emu_lib, monkeypatch):
    """Mutation check of the test scaffold itself: leaving out ANY single hand-over rule of the plans (sdr_lay.h) must be
    noticed by at least one schedule -- as wrong bits (shared memory is poisoned, a ring slot read too early or
    overwritten too early changes the output), a violated lock-step assertion or a stall.  Otherwise a forgotten rule in
    the kernel could hide behind the schedules tried here."""
    import audiosdr_b200 as A
    cases = []
    for cfg, nch, nblk in [(2, 40, 30), (4, 70, 24)]:
        I, Q, ev = S.make(cfg, list(range(nch)), nblk)
        cases.append((I, Q, ev, oracle.run(I, Q, ev, threads=4)["audio"]))

    def noticed():
        for sched, landing in (("producers", "early"), ("consumers", "early"), ("random:1", "late"), ("producers", "late"), ("random:6", "early")):
            monkeypatch.setenv("SDR_EMU_SCHED", sched)
            monkeypatch.setenv("SDR_EMU_ASYNC", landing)  # asynchronous copies land at the request / as late as their wait
            for I, Q, ev, want in cases:
                try:
                    got = harness.run_batch(emu_lib, I, Q, ev, chunks=(7, 1, 13))
                except A.SdrError:
                    return True
                if not harness.bits_equal(got, want):
                    return True
        return False

    monkeypatch.setenv("SDR_EMU_DROP_RULE", "-1")
    assert not noticed()  # all rules in place: every schedule is exact
    missed = []
    for rule in range(30):
        monkeypatch.setenv("SDR_EMU_DROP_RULE", str(rule))
        if not noticed():
            missed.append(rule)
    # rule numbers beyond a plan's rule count drop nothing; the shortest plan of these cases (ENV class with blanker) has 21 rules
    assert [r for r in missed if r < 21] == [], "schedules did not notice the missing rule(s) %s" % missed


@pytest.mark.parametrize("plan", [None, (8, 3, 1), (16, 2, 0)])
def test_emulated_sam_driven_out_of_lock_and_back(oracle, emu_lib, monkeypatch, plan):
    """SAM channels pushed out of the lock window and back (envelope fallback toggling per block, C:130-143), lean and
    blanker-carrying ENV plans, default and short-tile plans, under an adversarial schedule."""
    set_schedule(monkeypatch, "random:4/late")
    if plan:
        set_plan(monkeypatch, *plan)
    nch, nblk = 6, 400
    I, Q, ev = S.sam_lock_unlock_case(nch, nblk)
    o = oracle.run(I, Q, ev, threads=4)
    col = harness.STATUS_FIELDS.index("sam_locked")
    seg = nblk // 5
    before = oracle.run(I[:, :(seg - 2) * 128], Q[:, :(seg - 2) * 128], ev, threads=4, want_pcm=False)["status"][:, col]
    during = oracle.run(I[:, :(2 * seg - 2) * 128], Q[:, :(2 * seg - 2) * 128], ev, threads=4, want_pcm=False)["status"][:, col]
    assert before.all() and not during.any()  # the case does what it says
    a, b = harness.run_batch(emu_lib, I, Q, ev, chunks=(37, 1, 90), return_batch=True)
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)


def migration_case(lib, I, Q, ev, device=None):
    """Runs channels for a while on handle A, moves them -- blobs of export_state -- into OTHER channel slots of handle B
    that has already processed a different number of blocks (the blanker's block-indexed ring slots turn), and continues
    there.  Returns the concatenated audio of the moved channels."""
    import audiosdr_b200 as A
    nch, ns = I.shape
    nblk = ns // 128
    cut = nblk // 2 + 1
    a = A.SdrBatch(nch, _lib=lib)
    a.configure([(None if e[0] == 0xFFFFFFFF else e[0], e[2]) + tuple(e[3:]) for e in ev if e[1] == 0])
    first = np.empty((nch, cut * 128), np.float32)
    if device is None:
        a.process_host(I[:, :cut * 128], Q[:, :cut * 128], first)
    else:
        import torch
        dI, dQ = torch.from_numpy(I).to(device), torch.from_numpy(Q).to(device)
        dO = torch.empty((nch, ns), dtype=torch.float32, device=device)
        a.process(dI[:, :cut * 128], dQ[:, :cut * 128], dO[:, :cut * 128], n_blocks=cut); torch.cuda.synchronize()
        first = dO[:, :cut * 128].cpu().numpy()
    blobs = a.export_state()
    b = A.SdrBatch(nch + 5, _lib=lib)
    z = np.zeros((nch + 5, 2 * 128), np.int16)
    b.process_host(z, z, np.empty((nch + 5, 2 * 128), np.float32))      # B is two blocks old: other ring-slot phase than A (cut blocks)
    perm = np.random.default_rng(5).permutation(nch + 5)[:nch]           # other slots, other lanes, other groups
    assert len(set(perm.tolist())) == nch
    b.import_state(perm, blobs)
    I2 = np.zeros((nch + 5, ns - cut * 128), np.int16); Q2 = np.zeros_like(I2)
    I2[perm] = I[:, cut * 128:]; Q2[perm] = Q[:, cut * 128:]
    second = np.empty((nch + 5, ns - cut * 128), np.float32)
    b.process_host(I2, Q2, second)
    return np.concatenate([first, second[perm]], 1)


def test_emulated_state_export_import_continues_bit_exact(oracle, emu_lib):
    I, Q, ev = S.make(4, list(range(44)), 15)
    want = oracle.run(I, Q, ev, threads=4)["audio"]
    got = migration_case(emu_lib, I, Q, ev)
    assert harness.bits_equal(got, want), harness.describe_mismatch(got, want)


SHORT_PLANS = [(16, 2, 0), (8, 2, 0), (16, 1, 0), (8, 3, 1), (16, 2, 1)]  # tile length, groups per SM, SAM-only buckets on the merged 7-warp plan


def set_plan(monkeypatch, tile, ctas, merge):
    monkeypatch.setenv("SDR_TILE_SSB", str(tile)); monkeypatch.setenv("SDR_TILE_ENV", str(tile)); monkeypatch.setenv("SDR_CTAS_PER_SM", str(ctas))
    monkeypatch.setenv("SDR_NO_MERGE", "0" if merge else "1")


@pytest.mark.parametrize("sched", ["lockstep", "lockstep-reversed", "consumers", "random:3/late"])
@pytest.mark.parametrize("tile,ctas,merge", SHORT_PLANS)
@pytest.mark.parametrize("cfg,nch,nblk", [(1, 1, 24), (3, 6, 30), (5, 5, 30)])
def test_emulated_short_tile_plans(oracle, emu_lib, monkeypatch, cfg, nch, nblk, tile, ctas, merge, sched):
    """Buckets without blanker and ALS can run 16- or 8-sample tiles in a fraction of the shared memory (several groups
    per SM; the product picks them for large ENV buckets, SAM-only ones with the light stages merged into 7 warps): same
    bits for every tile length, ring budget and warp program."""
    set_schedule(monkeypatch, sched)
    set_plan(monkeypatch, tile, ctas, merge)
    I, Q, ev = S.make(cfg, list(range(nch)), nblk)
    o = oracle.run(I, Q, ev, threads=4)
    a, b = harness.run_batch(emu_lib, I, Q, ev, chunks=(7, 1, 13), return_batch=True)
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)


ALS_SMALL_PARAMS = [(55, 0.5, 3), (86, 0.3, 1), (1, 0.5, 0), (4, 0.6, 1), (60, 0.5, 27), (85, 0.25, 0), (33, 0.5, 54)]


def als_small_events(nch):
    """ALS parameters within what the post-pass keeps in a quarter of an SM with its input ring doubled (M <= 86, M + delay <= 87)."""
    ev = []
    for c in range(nch):
        m, lam, d = ALS_SMALL_PARAMS[c % len(ALS_SMALL_PARAMS)]
        ev += [(c, 0, "setALSfilterParams", m, lam, d), (c, 0, "enableALSfilter"), (c, 0, "setALSfilterAdaptive" if c % 5 else "setALSfilterStatic")]
        ev += [(c, 0, "setALSfilterNotch" if c % 2 else "setALSfilterPeak")]
    return ev


@pytest.mark.parametrize("sched", ["lockstep", "lockstep-reversed", "consumers", "random:4/late"])
@pytest.mark.parametrize("case", ["config4", "edge", "sliced", "sam", "small", "small-single-ring", "config4-largest-plan"])
def test_emulated_split_als_bucket(oracle, emu_lib, monkeypatch, case, sched):
    """A bucket with the ALS filter and more groups than SMs runs as two launches -- the chain up to the AGC into a scratch
    plane, then the ALS + output post-pass (sdr_lay.h, lay_build_als): same bits, whatever plan the chain runs on, also when
    the scratch plane only holds a slice of the call."""
    set_schedule(monkeypatch, sched)
    monkeypatch.setenv("SDR_ALS_SPLIT", "1")
    chunks = (7, 1, 13)
    if case == "config4":
        I, Q, ev = S.make(4, list(range(70)), 24)
    elif case == "edge":
        nch = 2 * len(ALS_EDGE_PARAMS)
        I, Q, ev = S.make(4, list(range(nch)), 14)
        ev = [e for e in ev if not e[2].startswith("setALSfilterParams")] + als_edge_events(nch)
        chunks = (3, 1, 6, 4)
    elif case.startswith("small"):  # doubled input ring (no sweep wraps) with every kind of tap count / delay it admits; the same on a single ring
        if case == "small-single-ring":
            monkeypatch.setenv("SDR_ALS_NO_MIRROR", "1")
        nch = 3 * len(ALS_SMALL_PARAMS)
        I, Q, ev = S.make(4, list(range(nch)), 14)
        ev = [e for e in ev if not e[2].startswith("setALSfilterParams")] + als_small_events(nch)
        chunks = (3, 1, 6, 4)
    elif case == "config4-largest-plan":  # all 128 tap rows and the deepest ring although the channels use the defaults
        monkeypatch.setenv("SDR_ALS_FULL_ROWS", "1")
        I, Q, ev = S.make(4, list(range(70)), 24)
    elif case == "sliced":  # 3 groups: 48 KB of scratch per block, 1 MB holds 21 blocks -> the 40-block call runs as two slices
        monkeypatch.setenv("SDR_ALS_SCRATCH_MB", "1")
        I, Q, ev = S.make(4, list(range(20)), 40)
        chunks = (40,)
    else:  # SAM channels with ALS and nothing else: the chain runs on the merged 7-warp plan with 16-sample tiles
        monkeypatch.setenv("SDR_TILE_ENV", "16"); monkeypatch.setenv("SDR_CTAS_PER_SM", "3"); monkeypatch.setenv("SDR_IN_DEPTH", "1")
        I, Q, ev = S.make(3, list(range(6)), 30)
        ev += [(c, 0, "enableALSfilter") for c in range(6)] + [(c, 0, "setALSfilterNotch" if c % 2 else "setALSfilterPeak") for c in range(6)]
    o = oracle.run(I, Q, ev, threads=4)
    a, b = harness.run_batch(emu_lib, I, Q, ev, chunks=chunks, return_batch=True)
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)
    p = harness.run_batch(emu_lib, I, Q, ev, chunks=(3, 11), out_dtype=np.int16)
    assert np.array_equal(p, o["pcm"])


def every_bucket_case(nblk=12):
    """One handle with every kind of bucket at once: USB / AM / SAM channels x blanker on/off x ALS on/off = 12 launches on 12
    streams per call."""
    modes = [S.USB, S.AM, S.SAM]
    chans = [(m, nb, als) for m in modes for nb in (0, 1) for als in (0, 1)]
    src = {S.USB: 2, S.AM: 4, S.SAM: 3}
    I = np.empty((len(chans), nblk * 128), np.int16); Q = np.empty_like(I)
    ev = []
    for c, (m, nb, als) in enumerate(chans):
        cfg = src[m]
        pick = next(k for k in range(4096) if S.channel_mode(cfg, k) == m)
        i, q, _ = S.make(cfg, [pick], nblk)
        I[c], Q[c] = i[0], q[0]
        ev += [(c, 0, "setDemodMode", m), (c, 0, "setAGCmode", S.AGC_MEDIUM), (c, 0, "enableAudioFilter"), (c, 0, "setMute", 0)]
        ev += [(c, 0, "enableNoiseBlanker" if nb else "disableNoiseBlanker"), (c, 0, "setNoiseBlankerThresholdDb", 10.0)]
        ev += [(c, 0, "enableALSfilter" if als else "disableALSfilter")]
    return I, Q, ev


@pytest.mark.parametrize("split", ["0", "1"])
def test_emulated_every_bucket_kind_in_one_handle(oracle, emu_lib, monkeypatch, split):
    monkeypatch.setenv("SDR_ALS_SPLIT", split)
    I, Q, ev = every_bucket_case()
    o = oracle.run(I, Q, ev, threads=4)
    a, b = harness.run_batch(emu_lib, I, Q, ev, chunks=(5, 7), return_batch=True)
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])
    assert np.array_equal(harness.status_matrix(b), o["status"], equal_nan=True)


def test_emulated_als_forms_alternate_between_calls(oracle, emu_lib, monkeypatch):
    """The one-launch and the two-launch form of an ALS bucket keep the same state words: a handle may change form from call
    to call (the plan is made again whenever the grouping changes)."""
    import audiosdr_b200 as A
    I, Q, ev = S.make(4, list(range(40)), 20)
    want = oracle.run(I, Q, ev, threads=4)["audio"]
    monkeypatch.setenv("SDR_MAP_SEARCH", "1")  # plan at every call
    h = A.SdrBatch(40, _lib=emu_lib)
    h.configure([(e[0], e[2]) + tuple(e[3:]) for e in ev])
    got = np.empty((40, 20 * 128), np.float32)
    b0 = 0
    for k, nb in enumerate((3, 5, 1, 4, 7)):
        monkeypatch.setenv("SDR_ALS_SPLIT", str(k & 1))
        out = np.empty((40, nb * 128), np.float32)
        h.process_host(np.ascontiguousarray(I[:, b0 * 128:(b0 + nb) * 128]), np.ascontiguousarray(Q[:, b0 * 128:(b0 + nb) * 128]), out)
        got[:, b0 * 128:(b0 + nb) * 128] = out
        b0 += nb
    assert harness.bits_equal(got, want), harness.describe_mismatch(got, want)


CONTRACT_TOL = 1e-4  # north_star: max abs error of full scale for the full chain


@pytest.mark.parametrize("cfg,nch,nblk", [(1, 1, 200), (2, 24, 120), (3, 6, 120), (4, 35, 120), (5, 5, 120)])
def test_emulated_contracting_build_error_profile(oracle, emu_contract_lib, cfg, nch, nblk):
    """The opt-in contracting build (fused multiply-adds and history-first sums in the cascades, the Hilbert FIR and the NCO
    product) is not bit-exact.  Up to the AGC its output stays far inside north_star's 1e-4 of full scale (checked with the
    AGC off: a few 1e-5 at most).  Behind the AGC a 1-ulp difference of the level can move the index into the reference's 129-entry
    gain table (C:419, C:483-494) by one step, which -- on the table's steep first segment -- changes the gain of that one
    sample by up to ~0.6 %: isolated samples beyond 1e-4, which no non-bit-exact build can avoid.  That is
    why the exact build is the default and the product's parity claim; this test pins the size and rarity of the effect."""
    I, Q, ev = S.make(cfg, list(range(nch)), nblk)
    o = oracle.run(I, Q, ev, threads=4)
    a = harness.run_batch(emu_contract_lib, I, Q, ev, chunks=(7, 1, 13))
    err = np.abs(a.astype(np.float64) - o["audio"])
    assert float(np.mean(err > CONTRACT_TOL)) < 1e-4 and float(err.max()) < 5e-3, "max %g, share above tolerance %g" % (err.max(), np.mean(err > CONTRACT_TOL))
    assert float(np.sqrt(np.mean(err ** 2))) < 2e-6
    assert not harness.bits_equal(a, o["audio"])  # it really is the other arithmetic
    ev_noagc = ev + [(c, 0, "disableAGC") for c in range(nch)]
    o2 = oracle.run(I, Q, ev_noagc, threads=4)
    a2 = harness.run_batch(emu_contract_lib, I, Q, ev_noagc, chunks=(7, 1, 13))
    fin = np.isfinite(o2["audio"]).all(axis=1) & np.isfinite(a2).all(axis=1)  # (config 4: without AGC the adaptive ALS runs away, in the reference too)
    if fin.any():
        scale = max(1.0, float(np.max(np.abs(o2["audio"][fin]))))
        assert float(np.max(np.abs(a2[fin].astype(np.float64) - o2["audio"][fin]))) <= CONTRACT_TOL * scale


def test_emulated_host_stream_of_submitted_calls(oracle, emu_lib):
    """submit_host / wait_host: ragged calls queued back to back (the staging rotation carries over from call to call), a
    setter between two submits, a synchronous call in between: the oracle's bits."""
    import audiosdr_b200 as A
    nblk = 40
    I, Q, ev = S.make(4, list(range(37)), nblk)
    ev = [e for e in ev if e[1] == 0] + [(3, 20, "setDemodMode", 2), (9, 20, "setAGCthreshold", -30.0)]
    o = oracle.run(I, Q, ev, threads=4)
    out = np.zeros((37, nblk * 128), np.int16)
    b = A.SdrBatch(37, _lib=emu_lib)
    b.configure([(None if e[0] == 0xFFFFFFFF else e[0], e[2]) + tuple(e[3:]) for e in ev if e[1] == 0])
    pos = 0
    for k, n in enumerate((17, 3, 12, 1, 7)):
        if pos == 20:
            b.configure([(3, "setDemodMode", 2), (9, "setAGCthreshold", -30.0)])
        a, z = pos * 128, (pos + n) * 128
        if k == 3:
            b.process_host(I[:, a:z], Q[:, a:z], out[:, a:z])
        else:
            assert b.submit_host(I[:, a:z], Q[:, a:z], out[:, a:z]) == k + 1  # the call's ticket
        pos += n
    b.wait_host(ticket=2)
    b.wait_host()
    assert np.array_equal(out, o["pcm"])
    with pytest.raises(A.SdrError):
        b.wait_host(ticket=99)


def test_emulated_agc_parameter_sweep_collects_tables(oracle, emu_lib):
    """A setter sweep (one new AGC table per block on two channels, 150 of them) crosses the point where the host drops the
    tables nobody uses any more and renumbers the rest; setters on single channels take the incremental upload path."""
    nblk = 150
    I, Q, ev = S.make(4, list(range(40)), nblk)
    ev = [e for e in ev if e[1] == 0]
    for k in range(1, nblk):
        ev.append((2, k, "setAGCthreshold", -10.0 - 0.25 * k))
        if k % 2:
            ev.append((33, k, "setAGCslope", 2.0 + 0.05 * k))
        if k % 17 == 0:
            ev.append((5, k, "setOutputGain", 0.1 + 0.01 * k))
    o = oracle.run(I, Q, ev, threads=4)
    a, b = harness.run_batch(emu_lib, I, Q, ev, chunks=(1, 3, 2), return_batch=True)
    assert harness.bits_equal(a, o["audio"]), harness.describe_mismatch(a, o["audio"])
    # the getters still find every channel's own table after the renumbering: same table as a fresh handle set to the final values
    import audiosdr_b200 as A
    f = A.SdrBatch(40, _lib=emu_lib)
    f.configure([(None if e[0] == 0xFFFFFFFF else e[0], e[2]) + tuple(e[3:]) for e in ev if e[1] == 0])
    f.configure([(2, "setAGCthreshold", -10.0 - 0.25 * (nblk - 1)), (33, "setAGCslope", 2.0 + 0.05 * (nblk - 1))])
    for c in (0, 2, 33, 39):
        assert np.array_equal(b.getAGClookup(c), f.getAGClookup(c))
