"""CPU tier: the oracle (plain-C restatement, oracle/sdr_oracle.c) against the reference's own outputs.

* tests/golden/*.npz were produced by the UNMODIFIED reference compiled on the host (tools/gen_golden.py);
  the oracle must reproduce them bit for bit (float audio, int16 pcm, getters).
* When oracle/_ref/refsdr is present (it is built wherever /root/reference exists and travels with the
  repo snapshot), the oracle is also compared live on fresh inputs of every BASELINE config and on a
  random setter sequence.
"""
import glob
import os

import numpy as np
import pytest

import harness
import signals as S
from oracle import ref_client as rc

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("aux_"))


def load_golden(path):
    z = np.load(path)
    ev = [(int(r[0]), int(r[1]), int(r[2]), float(r[3]), float(r[4]), float(r[5])) for r in z["events"]]
    return z["I"], z["Q"], ev, z["audio"], z["pcm"], z["status"]


def test_golden_files_exist():
    assert len(GOLDEN) >= 7


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_golden(oracle, path):
    I, Q, ev, audio, pcm, status = load_golden(path)
    o = oracle.run(I, Q, ev, threads=4)
    assert harness.bits_equal(o["audio"], audio), harness.describe_mismatch(o["audio"], audio)
    assert np.array_equal(o["pcm"], pcm)
    assert np.array_equal(o["status"], status, equal_nan=True)


@pytest.mark.skipif(not rc.available(), reason="oracle/_ref/refsdr not built (needs /root/reference)")
@pytest.mark.parametrize("cfg,nch,nblk", [(1, 1, 200), (2, 24, 120), (3, 8, 120), (4, 42, 100), (5, 4, 150)])
def test_oracle_vs_live_reference(oracle, cfg, nch, nblk):
    I, Q, ev = S.make(cfg, list(range(100, 100 + nch)) if cfg != 1 else [0], nblk)
    r = rc.run(I, Q, ev)
    o = oracle.run(I, Q, ev, threads=4)
    assert harness.bits_equal(o["audio"], r["audio"]), harness.describe_mismatch(o["audio"], r["audio"])
    assert np.array_equal(o["pcm"], r["pcm"])
    assert np.array_equal(o["status"], r["status"], equal_nan=True)


@pytest.mark.skipif(not rc.available(), reason="oracle/_ref/refsdr not built (needs /root/reference)")
def test_oracle_vs_live_reference_setter_fuzz(oracle):
    rng = np.random.default_rng(99)
    I, Q, ev = S.make(4, list(range(32)), 48)
    ev += harness.fuzz_events(rng, 32, 48, 400)
    r = rc.run(I, Q, ev)
    o = oracle.run(I, Q, ev, threads=4)
    assert harness.bits_equal(o["audio"], r["audio"]), harness.describe_mismatch(o["audio"], r["audio"])
    assert np.array_equal(o["pcm"], r["pcm"])


def test_oracle_f32_boundary_equals_i16_at_unit_gain(oracle):
    """float32 planes carry q/32767; with input gain 1 that is exactly the reference's C:67-70 scaling."""
    I, Q, ev = S.make(2, list(range(6)), 40)
    a = oracle.run(I, Q, ev)
    If = (I.astype(np.float32) / np.float32(32767.0)).astype(np.float32)
    Qf = (Q.astype(np.float32) / np.float32(32767.0)).astype(np.float32)
    b = oracle.run(If, Qf, ev)
    assert harness.bits_equal(a["audio"], b["audio"])


def test_known_answers(oracle):
    """Physics-level sanity of the pinned oracle: USB two-tone comes out at 700/1900 Hz, SAM locks at carrier+df."""
    I, Q, ev = S.make(1, [0], 400)
    a = oracle.run(I, Q, ev)["audio"][0][-16384:]
    sp = np.abs(np.fft.rfft(a * np.hanning(a.size)))
    f = np.fft.rfftfreq(a.size, 1 / S.FS)
    peaks = sorted(set(np.round(f[np.argsort(sp)[-6:]] / 50.0) * 50))
    assert 700.0 in peaks and 1900.0 in peaks
    I, Q, ev = S.make(3, [0, 1, 2], 400)
    st = oracle.run(I, Q, ev)["status"]
    for row, c in enumerate([0, 1, 2]):
        df = 100.0 * ((S.chash(3, c, 4) >> 11) / float(1 << 53)) - 50.0
        assert st[row, 5] == 1.0 and abs(st[row, 4] - (6890.0 + df)) < 25.0
