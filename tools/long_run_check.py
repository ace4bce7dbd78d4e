"""tools/long_run_check.py -- BASELINE config 5 drift check at FULL length: 120 s = 41 344 blocks of WSPR per channel.

A handful of sampled channels (first/last of the 262 144, one per shard) stream 120 s through the CUDA path in ragged
calls and through the oracle; the whole output (5 292 032 samples per channel) and the final NCO phase must match bit for
bit.  (The full-channel-count run is tests/test_gpu_parity.py::test_cuda_full_channel_count_sampled at short duration.)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import harness, signals as S
from audiosdr_b200 import api
from oracle import oracle_lib

chans = [0, 32767, 32768, 131071, 196608, 262143]
nblk = S.BLOCKS_120S
t0 = time.time()
I, Q, ev = S.make(5, chans, nblk)
t1 = time.time()
want = oracle_lib.run(I, Q, ev, threads=len(chans), want_pcm=False)["audio"]
t2 = time.time()
got, b = harness.run_batch(api.load_library(), I, Q, ev, chunks=(3446, 1000, 2999), device=torch.device("cuda:0"), return_batch=True)
t3 = time.time()
ok = np.array_equal(got.view(np.uint32), want.view(np.uint32))
last = slice(-S.BLOCKS_10S * 128, None)
err = float(np.max(np.abs(got[:, last].astype(np.float64) - want[:, last])))
print("generate %.0fs oracle %.0fs gpu %.0fs" % (t1 - t0, t2 - t1, t3 - t2))
print("config5 120 s x %d channels: bit_exact=%s max_abs_err(last 10 s)=%g samples/channel=%d" % (len(chans), ok, err, got.shape[1]))
sys.exit(0 if ok else 1)
