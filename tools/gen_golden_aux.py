#!/usr/bin/env python3
"""tools/gen_golden_aux.py -- golden fixtures for the pre-processor and the I/Q generator, produced by the UNMODIFIED
reference compiled on the host (oracle/_ref/refaux; needs /root/reference at build time).  Inputs are regenerated from
tests/aux_signals.py by the tests; only the reference's OUTPUTS are stored (tests/golden/aux_*.npz)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import aux_signals as S
from oracle import aux_lib as A

assert A.ref_available(), "build oracle/_ref/refaux first (make -C oracle ref)"
out = os.path.join(ROOT, "tests", "golden")
nch, nb = 16, 1200
I, Q = S.pp_case(nch, nb)
o0, o1, st = A.ref_run("pp", (I, Q), S.pp_events(nch, nb))
# 1200 blocks are needed for the detector to reach its 1001 successes; the outputs are stored as one CRC-32 per
# (channel, block) over the block's I then Q bytes, plus the first 32 blocks verbatim
np.savez_compressed(os.path.join(out, "aux_pp.npz"), n_channels=nch, n_blocks=nb, crc=S.block_crcs(o0, o1), I_head=o0[:, :32 * 128],
                    Q_head=o1[:, :32 * 128], status=st)
nch, nb = 10, 40
X = S.iq_case(nch, nb)
o0, o1, _ = A.ref_run("iq", (X,), S.iq_events(nch, nb))
np.savez_compressed(os.path.join(out, "aux_iq.npz"), n_channels=nch, n_blocks=nb, I_out=o0, Q_out=o1)
for f in ("aux_pp.npz", "aux_iq.npz"):
    print(f, os.path.getsize(os.path.join(out, f)), "bytes")
