"""tools/sanitize_smoke.py -- tiny all-mode run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import harness, signals as S
from audiosdr_b200 import api
from oracle import oracle_lib
lib = api.load_library()
I, Q, ev = S.make(4, list(range(70)), 6)
want = oracle_lib.run(I, Q, ev, threads=4)
got = harness.run_batch(lib, I, Q, ev, chunks=(2, 4), device=torch.device("cuda:0"))
assert np.array_equal(got.view(np.uint32), want["audio"].view(np.uint32))
pcm = harness.run_batch(lib, I, Q, ev, chunks=(6,), out_dtype=np.int16, device=None)
assert np.array_equal(pcm, want["pcm"])
print("sanitize smoke ok")
# the plans large ENV buckets run on: 16-sample tiles with 11 warps, and the merged 7-warp SAM plan (three groups per SM)
I, Q, ev = S.make(3, list(range(40)), 6)
want = oracle_lib.run(I, Q, ev, threads=4)
for tile, ctas, merge in ((16, 2, 0), (16, 3, 1), (8, 3, 1)):
    os.environ["SDR_TILE_ENV"] = str(tile); os.environ["SDR_CTAS_PER_SM"] = str(ctas); os.environ["SDR_NO_MERGE"] = "0" if merge else "1"
    got = harness.run_batch(lib, I, Q, ev, chunks=(2, 4), device=torch.device("cuda:0"))
    assert np.array_equal(got.view(np.uint32), want["audio"].view(np.uint32)), (tile, ctas, merge)
print("sanitize smoke ok (short-tile plans)")
# the two-launch form of ALS buckets (chain into a scratch plane, one-warp ALS + output post-pass)
for k in ("SDR_TILE_ENV", "SDR_CTAS_PER_SM", "SDR_NO_MERGE"):
    os.environ.pop(k, None)
os.environ["SDR_ALS_SPLIT"] = "1"
I, Q, ev = S.make(4, list(range(70)), 6)
want = oracle_lib.run(I, Q, ev, threads=4)
got = harness.run_batch(lib, I, Q, ev, chunks=(2, 4), device=torch.device("cuda:0"))
assert np.array_equal(got.view(np.uint32), want["audio"].view(np.uint32))
pcm = harness.run_batch(lib, I, Q, ev, chunks=(6,), out_dtype=np.int16, device=None)
assert np.array_equal(pcm, want["pcm"])
print("sanitize smoke ok (split ALS buckets)")
