#!/usr/bin/env python3
"""tools/diag_skip.py -- diagnostics: time subsets of the pipeline's stages in isolation (SDR_DIAG_SKIP, profiling runs only).

For every subset of stages the bench workload is launched with all other stages idling at the step barrier; printed are the
cycles per step and the busy cycles per tile of the stages that ran.  Outputs are meaningless in these runs; the point is
what a stage (or the stages of one SM sub-partition) costs without the others competing for issue slots and caches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SDR_ROLE_PROFILE"] = "1"
import numpy as np, torch
import bench
from audiosdr_b200 import api

NAMES = ["in", "nb_scan", "if_i", "if_q", "nco", "hil0", "hil1", "hil2", "hil3", "aud", "agc", "als_out", "envl", "nb_out"]
nch, blocks = 4096, 64
dev = torch.device("cuda:0")
I16, Q16, calls = bench.synth_planes(dev, 0, nch, blocks * 128, 1234, 2)
If, Qf = I16.float() / 32767.0, Q16.float() / 32767.0
out = torch.empty((nch, blocks * 128), dtype=torch.float32, device=dev)
b = api.SdrBatch(nch)
b.configure(calls)
stream = torch.cuda.current_stream()
steps = blocks * 4 + 8


def raw():
    busy = np.zeros(28, np.uint64); total = np.zeros(2, np.uint64); groups = np.zeros(2, np.uint64)
    b._check(b.L.sdr_batch_get_role_profile(b.h, busy.ctypes.data, total.ctypes.data, groups.ctypes.data))
    return busy[:14].astype(np.float64), float(total[0]), float(groups[0])


def run(keep, label, extra=0):
    skip = (0x3FFF & ~sum(1 << w for w in keep)) | extra
    os.environ["SDR_DIAG_SKIP"] = "%X" % skip
    for _ in range(2):
        b.process(If, Qf, out, n_blocks=blocks, stream=stream)
    torch.cuda.synchronize()
    b0, t0, g0 = raw()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        b.process(If, Qf, out, n_blocks=blocks, stream=stream)
    e1.record(stream); torch.cuda.synchronize()
    b1, t1, g1 = raw()
    g = g1 - g0
    cyc = (t1 - t0) / g / steps
    busy = {NAMES[w]: round((b1[w] - b0[w]) / g / (blocks * 4)) for w in keep}
    print("%-34s %6.0f cycles/step  %7.3f ms/launch  busy cycles/tile %s" % (label, cyc, e0.elapsed_time(e1) / 3, busy), flush=True)


ALL = list(range(14))
os.environ.pop("SDR_ROLE_PROFILE_NB", None)
run(ALL, "all stages")
run([5], "one Hilbert only")
run([5, 6], "Hil+Hil same sub-partition")
run([2], "IF-I only")
run([2, 3], "IF-I + IF-Q same sub-partition")
run([2, 3, 9], "three cascades same sub-partition")
run([2, 3, 9, 12], "sub-partition 1 as placed: 3 cascades + ENVL")
run([0, 5, 6, 11], "sub-partition 0 as placed: IN Hil Hil OUT")
run([7, 8, 13], "sub-partition 2 as placed: Hil Hil NB-out")
run([1, 4, 10], "sub-partition 3 as placed: NB-scan NCO AGC")
os.environ["SDR_MAP_SSB"] = "32A90D41CB8765"   # one Hilbert + one cascade per sub-partition
run([5, 2], "balanced map: Hil + IF-I same sub-partition")
run([5, 2, 13, 11], "balanced map: Hil + IF-I + NB-out + OUT (one sub-partition)")
run([5, 6, 7, 8, 2, 3, 9], "balanced map: Hilbert x4 + cascades x3")
run(ALL, "balanced map: all stages")
