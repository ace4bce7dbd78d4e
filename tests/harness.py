"""tests/harness.py -- drive a batch (CUDA library or its host emulation) like the oracle is driven.

`run_batch` replays an event list (channel, block, setter, args...) through SdrBatch.configure at the
right block boundaries and streams the input in ragged chunks, so that every test also exercises
state carry across process() calls.
"""
import numpy as np

import audiosdr_b200 as A
from oracle import ref_client as rc

N_BLOCK = 128
STATUS_FIELDS = rc.STATUS_FIELDS


def run_batch(lib, I, Q, events, chunks=(7, 1, 13), out_dtype=np.float32, device=None, return_batch=False, contract=False):
    """I/Q: host arrays [C, S] (int16 or float32).  `device`: None -> process_host; a torch device -> process()."""
    nch, ns = I.shape
    b = A.SdrBatch(nch, _lib=lib, contract=contract)
    ev = sorted(events, key=lambda e: e[1])
    nb_total = ns // N_BLOCK
    outs, pos, ei, k = [], 0, 0, 0
    if device is not None:
        import torch
        dI, dQ = torch.from_numpy(np.ascontiguousarray(I)).to(device), torch.from_numpy(np.ascontiguousarray(Q)).to(device)
        dO = torch.empty((nch, ns), dtype=torch.float32 if out_dtype == np.float32 else torch.int16, device=device)
    while pos < nb_total:
        calls = []
        while ei < len(ev) and ev[ei][1] <= pos:
            e = ev[ei]
            calls.append((None if e[0] == 0xFFFFFFFF else e[0], e[2]) + tuple(e[3:]))
            ei += 1
        if calls:
            b.configure(calls)
        nxt = ev[ei][1] if ei < len(ev) else nb_total
        sz = min(chunks[k % len(chunks)], nb_total - pos, max(nxt - pos, 1))
        k += 1
        a, z = pos * N_BLOCK, (pos + sz) * N_BLOCK
        if device is None:
            out = np.empty((nch, sz * N_BLOCK), out_dtype)
            b.process_host(I[:, a:z], Q[:, a:z], out)
            outs.append(out)
        else:
            b.process(dI[:, a:z], dQ[:, a:z], dO[:, a:z], n_blocks=sz)
        pos += sz
    if device is not None:
        import torch
        torch.cuda.synchronize()
        res = dO.cpu().numpy()
    else:
        res = np.concatenate(outs, 1)
    return (res, b) if return_batch else res


def status_matrix(batch):
    rows = []
    for s in batch.status():
        rows.append([s.tuning_offset, s.mode, s.agc_active, s.nb_detected, s.sam_frequency, s.sam_locked, s.am_carrier,
                     s.bpf_lower, s.bpf_upper, s.muted, s.audio_filter, s.agc_enabled, s.nb_enabled, s.als_enabled,
                     s.agc_gain, s.nb_average])
    return np.array(rows, np.float32)


def bits_equal(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    return np.array_equal(a, b)


def describe_mismatch(a, b):
    bad = np.argwhere(a.view(np.uint32) != b.view(np.uint32)) if a.dtype == np.float32 else np.argwhere(a != b)
    if len(bad) == 0:
        return "identical"
    c, n = bad[0]
    return "%d mismatching samples in channels %s; first at ch %d sample %d (block %d): got %r want %r" % (
        len(bad), sorted(set(bad[:, 0].tolist()))[:16], c, n, n // N_BLOCK, a[c, n], b[c, n])


def fuzz_events(rng, n_channels, n_blocks, n_events, safe=True):
    """Random setter calls at random blocks.  `safe` keeps ALS parameters inside the reference's ring."""
    names = [n for n in rc.OPS if n != "oracle_identity_IF"]
    ev = []
    for _ in range(n_events):
        ch = int(rng.integers(0, n_channels)); blk = int(rng.integers(0, n_blocks)); op = names[int(rng.integers(0, len(names)))]
        a = [0.0, 0.0, 0.0]
        if op == "setDemodMode": a[0] = int(rng.integers(0, 7))
        elif op == "setAudioFilter": a[0] = int(rng.integers(0, 11))
        elif op == "setAGCmode": a[0] = int(rng.integers(0, 4))
        elif op == "setALSfilterParams": a = [int(rng.integers(1, 100)), float(rng.uniform(0.01, 0.6)), int(rng.integers(0, 20))]
        elif op in ("setInputGain", "setOutputGain"): a[0] = float(rng.uniform(0, 2))
        elif op == "setIQgainBalance": a[0] = float(rng.uniform(0.8, 1.2))
        elif op == "setAGCthreshold": a[0] = float(rng.uniform(-80, -20))
        elif op == "setAGCslope": a[0] = float(rng.uniform(0.05, 0.9))
        elif op == "setAGCkneeWidth": a[0] = float(rng.uniform(0.5, 10))
        elif op in ("setAGCattackTime", "setAGCreleaseTime", "setAGChangTime"): a[0] = float(rng.uniform(1, 800))
        elif op == "setAGCstaticGain": a[0] = float(rng.uniform(1, 20))
        elif op == "setNoiseBlankerThreshold": a[0] = float(rng.uniform(1.1, 5))
        elif op == "setNoiseBlankerThresholdDb": a[0] = float(rng.uniform(1, 20))
        elif op == "setMute": a[0] = int(rng.integers(0, 2))
        ev.append((ch, blk, op, *a))
    return ev
