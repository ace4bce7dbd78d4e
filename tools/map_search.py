#!/usr/bin/env python3
"""tools/map_search.py [--class ssb|env] [--seconds S] -- search the physical-warp -> stage placement of the pipeline kernel.

The placement decides which SM sub-partition (warp id % 4) a stage shares with which others and its priority there (the
scheduler prefers the higher warp id).  The host reads SDR_MAP_SSB / SDR_MAP_ENV at every launch, so one process can time
many placements: hill climbing over pair swaps with random restarts, each candidate timed with CUDA events (best of 3
launches of the default bench workload).  Prints the best placements found; the winner goes into sdr_types.h."""
import argparse, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from audiosdr_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--cls", default="ssb")
ap.add_argument("--seconds", type=float, default=150.0)
ap.add_argument("--blocks", type=int, default=128)
ap.add_argument("--start", default="")
ap.add_argument("--als", action="store_true", help="enable the ALS filter on every channel (the placement of buckets with ALS)")
ap.add_argument("--nb", action="store_true", help="enable the noise blanker on every channel (ENV buckets with blanker: BASELINE config 4's)")
ap.add_argument("--mode", type=int, default=-1, help="setDemodMode(mode) on every channel after the config's own setters (4 = AM, 5 = SAM)")
ap.add_argument("--split", action="store_true", help="force the two-launch form of ALS buckets (SDR_ALS_SPLIT=1): with --config 4 the ENV chain launches of BASELINE config 4 (blanker on, ALS in the post-pass)")
ap.add_argument("--idle", default="", help="hex digits of stages that idle in this bucket (config 5: 1CD = blanker scan, envelope, blanker out): placements that differ only in them count as one")
ap.add_argument("--config", type=int, default=0, help="BASELINE config to take signals and setters from (default: 2 for ssb, 3 for env)")
args = ap.parse_args()
cfg_id = args.config or (2 if args.cls == "ssb" else 3)
nch = 4096 if args.cls not in ("envlean", "envmerged") else 16384   # enough groups for the plans that share an SM (sdr_host.cpp, plan_bucket)
if args.cls == "envlean":
    os.environ["SDR_NO_MERGE"] = "1"; os.environ["SDR_TILE_ENV"] = "16"; os.environ["SDR_CTAS_PER_SM"] = "2"
if args.split:
    os.environ["SDR_ALS_SPLIT"] = "1"
os.environ["SDR_MAP_SEARCH"] = "1"               # the host re-plans at every call, so that it reads the placement variable again
dev = torch.device("cuda:0")
I16, Q16, calls = bench.synth_planes(dev, 0, nch, args.blocks * 128, 1234, cfg_id)
If, Qf = I16.float() / 32767.0, Q16.float() / 32767.0
out = torch.empty((nch, args.blocks * 128), dtype=torch.float32, device=dev)
b = api.SdrBatch(nch)
b.configure(calls)
if args.als:
    b.enableALSfilter(None)
if args.mode >= 0:
    b.setDemodMode(None, args.mode)
if args.nb:
    b.enableNoiseBlanker(None)
var = {"ssb": "SDR_MAP_SSB", "env": "SDR_MAP_ENV", "envlean": "SDR_MAP_ENV_LEAN", "envmerged": "SDR_MAP_ENV_MERGED"}[args.cls]
NW = {"envlean": 11, "envmerged": 7}.get(args.cls, 14)
stream = torch.cuda.current_stream()

def pack(perm):
    if args.cls == "envmerged":   # a string of program numbers, warp 0 first
        return "".join(str(p) for p in perm)
    return "%X" % sum(s << (4 * w) for w, s in enumerate(perm))

def measure(perm, reps=3):
    os.environ[var] = pack(perm)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); b.process(If, Qf, out, n_blocks=args.blocks, stream=stream); e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

IDLE = set(int(c, 16) for c in args.idle)
def canon(perm):  # the four Hilbert warps (SSB stages 5..8) are interchangeable, and so are stages that idle
    return tuple(5 if (args.cls == "ssb" and 5 <= s <= 8) else (99 if s in IDLE else s) for s in perm)

if args.cls == "envmerged":
    start = [int(c) for c in args.start] if args.start else [1, 2, 0, 5, 4, 6, 3]
else:
    start = [int(c, 16) for c in reversed(args.start)] if args.start else ([0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11] if args.cls == "envlean" else [5, 6, 7, 8, 2, 3, 11, 4, 9, 12, 10, 1, 0, 13])
for _ in range(3):
    measure(start)
seen = {}
perms = {}
best_perm, best_t = list(start), measure(start, 5)
print("start %s %.3f ms" % (pack(start), best_t), flush=True)
t_end = time.time() + args.seconds
cur, cur_t = list(best_perm), best_t
stale = 0
while time.time() < t_end:
    i, j = random.sample(range(NW), 2)
    cand = list(cur); cand[i], cand[j] = cand[j], cand[i]
    key = canon(cand)
    if key in seen or canon(cand) == canon(cur):
        continue
    t = measure(cand)
    seen[key] = t
    perms[key] = list(cand)
    if t < cur_t * 0.999:
        t2 = measure(cand, 5)
        if t2 < cur_t:
            cur, cur_t, stale = cand, t2, 0
            if cur_t < best_t:
                best_perm, best_t = list(cur), cur_t
                print("best %s %.3f ms" % (pack(best_perm), best_t), flush=True)
            continue
    stale += 1
    if stale > 120:   # restart from a random shuffle of the best
        cur = list(best_perm)
        for _ in range(4):
            i, j = random.sample(range(NW), 2); cur[i], cur[j] = cur[j], cur[i]
        cur_t, stale = measure(cur, 5), 0
print("evals/s %.1f" % (len(seen) / args.seconds))
print("evaluated %d placements; best %s %.3f ms (%.0f Msps)" % (len(seen), pack(best_perm), best_t, nch * args.blocks * 128 / best_t / 1e3))
for key, t in sorted(seen.items(), key=lambda kv: kv[1])[:8]:
    print("  top %s first %.3f ms, re-measured %.3f ms" % (pack(perms[key]), t, measure(perms[key], 8)))
