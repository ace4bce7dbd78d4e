"""CPU tier: the N>1 path.  Channels shard across ranks with no data-path collective; the only
communication is the benchmark's counter gather.  Two gloo ranks each run their contiguous channel range
through the (emulated) pipeline and all_gather counters + outputs; rank 0 checks the union equals the
single-process result bit for bit."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import ctypes, os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import harness, signals as S
    from audiosdr_b200 import api
    from bench import shard_range, gather_counters
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = api._bind(ctypes.CDLL(os.path.join(%(root)r, "tests", "emu", "libsdr_emu.so")))
    total, nblk = 70, 10
    lo, hi = shard_range(total, rank, world)
    I, Q, ev = S.make(4, list(range(lo, hi)), nblk)
    out = harness.run_batch(lib, I, Q, ev, chunks=(4, 6))
    counters = gather_counters(float((hi - lo) * nblk * 128), 1.0 + rank, world, device=None)
    parts = [None] * world
    dist.all_gather_object(parts, out)
    if rank == 0:
        If, Qf, evf = S.make(4, list(range(total)), nblk)
        full = harness.run_batch(lib, If, Qf, evf, chunks=(10,))
        assert harness.bits_equal(np.concatenate(parts, 0), full)
        assert counters["samples"] == total * nblk * 128 and counters["max_ms"] == float(world)
        print("SHARD_OK")
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_shards(emu_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARD_OK" in r.stdout
