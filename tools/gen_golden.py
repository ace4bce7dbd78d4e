#!/usr/bin/env python3
"""tools/gen_golden.py -- generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/refsdr).

The reference ships no golden vectors, so these fixtures are outputs of the reference itself, run in
the build container (where /root/reference exists) through oracle/ref_driver.cpp.  They travel with the
repo; the GPU box (no /root/reference) checks the oracle and the CUDA path against them.
Each fixture: I, Q (int16 [C,S]), events (float64 [E,6]: channel, block, opcode, a0, a1, a2),
audio (float32 [C,S]), pcm (int16 [C,S]), status (float32 [C,16]).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
import signals as S  # noqa: E402
from oracle import ref_client as rc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ev_array(ev):
    rows = []
    for e in ev:
        e = tuple(e) + (0.0,) * (6 - len(e))
        rows.append([e[0], e[1], rc.OPS[e[2]] if isinstance(e[2], str) else e[2], e[3], e[4], e[5]])
    return np.array(rows, np.float64).reshape(-1, 6)


def save(name, I, Q, ev):
    r = rc.run(I, Q, ev)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), I=I, Q=Q, events=ev_array(ev), audio=r["audio"], pcm=r["pcm"],
                        status=r["status"])
    print(name, I.shape, "events", len(ev), "peak", float(np.nanmax(np.abs(r["audio"]))))


def main():
    os.makedirs(OUT, exist_ok=True)
    # one fixture per BASELINE config (first channels, 24 blocks = 3072 samples each)
    for cfg, n in ((1, 1), (2, 12), (3, 6), (4, 35), (5, 3)):
        I, Q, ev = S.make(cfg, list(range(n)), 24)
        save("config%d" % cfg, I, Q, ev)
    # every mode x {default construction state}: the power-on behaviour (NB on, audio filter off, LSB ...)
    I, Q, ev = S.make(4, list(range(7)), 16)
    ev = [(c, 0, "setDemodMode", c) for c in range(7)]
    save("poweron_modes", I, Q, ev)
    # setter semantics: random setters at random block boundaries (state resets, order dependence)
    rng = np.random.default_rng(20261017)
    I, Q, ev = S.make(4, list(range(24)), 32)
    ev += harness.fuzz_events(rng, 24, 32, 260)
    save("setter_fuzz", I, Q, ev)


if __name__ == "__main__":
    main()
