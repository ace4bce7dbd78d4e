#!/usr/bin/env python3
"""bench_aux.py -- measurement of the two blocks either side of the receiver chain, same rules as
bench.py: W >= 3 warm-up steps, K timed steps bracketed by CUDA events on the launching stream, inputs larger than L2,
one JSON line per path.  NOT the headline metric (that is bench.py); these lines carry `"path"`.

  python bench_aux.py [--path iqgen|preproc_static|preproc_detect|all] [--steps K --warmup W] [--no-cpu-baseline]

  iqgen           AudioIQgenerator::update(), 4096 channels x 256 blocks per step.  192 unfused FP32 operations and 6 bytes
                  per sample: FP32-issue bound; roofline against the live FMUL+FADD issue microbenchmark and against HBM.
  preproc_static  AudioSDRpreProcessor::update() with the detector off (correction +1 on every channel): a shifted copy,
                  8 bytes per sample, HBM bound.
  preproc_detect  ... with the detector running on every channel (noise input: it never reaches a verdict): one 128-point
                  FFT per block and channel.
`e2e` = the same through *_process_host with pinned host planes.  `cpu_baseline` = the unmodified reference (oracle/_ref/refaux)
on all host cores, else the oracle port, on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
UNIT = "Msps"
NCH, NBLK = 4096, 256


def measured_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback"


def fp32_issue_rate():
    """instructions/s of the unfused FMUL+FADD microbenchmark shipped in libsdr_batch.so (the same one bench.py uses)"""
    from audiosdr_b200 import api
    lib = api.load_library()
    ips, ms = C.c_double(), C.c_float()
    lib.sdrk_fp32_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    if lib.sdrk_fp32_peak(1, 4096, C.byref(ips), C.byref(ms)) != 0:
        raise RuntimeError("fp32 microbenchmark failed")
    return ips.value


def cpu_rate(kind, planes, events, seconds):
    from oracle import aux_lib as A
    cores = os.cpu_count() or 1
    if A.ref_available():
        r = A.ref_bench(kind, planes, events, seconds=seconds, jobs=cores)
        return dict(value=r["sps_update_only"] / 1e6, unit=UNIT, cores=cores, kind="reference", wall_msps=r["sps_wall"] / 1e6,
                    sample="%d sampled channels x %d blocks streamed round-robin for %.0f s, one channel per worker" % (planes[0].shape[0], planes[0].shape[1] // 128, seconds))
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < seconds:
        A.run(kind, planes, events, threads=cores); n += planes[0].size
    return dict(value=n / (time.perf_counter() - t0) / 1e6, unit=UNIT, cores=cores, kind="port", sample="oracle port, %d channels" % planes[0].shape[0])


def time_steps(fn, steps, warmup, stream):
    import torch
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record(stream)
    for k in range(steps):
        fn()
        ev[k + 1].record(stream)
    torch.cuda.synchronize()
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    return ev[0].elapsed_time(ev[-1]), per


def bench_path(path, args):
    import torch
    import aux_signals as S
    from audiosdr_b200 import aux
    from oracle import aux_lib as A
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream()
    ns = NBLK * 128
    g = torch.Generator(device=dev); g.manual_seed(1234)
    noise = lambda: (torch.randn((NCH, ns), generator=g, device=dev) * 3000.0).round().clamp(-32768, 32767).to(torch.int16)
    hbm, hbm_src = measured_hbm()
    parity = None
    if path == "iqgen":
        X = noise(); oi, oq = torch.empty_like(X), torch.empty_like(X)
        h = aux.IQGeneratorBatch(NCH)
        fn = lambda: h.process(X, oi, oq, n_blocks=NBLK, stream=stream)
        bytes_per_sample, kernel, planes_in, planes_out = 6.0, "iq_generate_kernel", 1, 2
        cpu_planes, cpu_events, kind = (S.iq_case(16, 64),), [], "iq"
        host_fn = lambda a, o: h.process_host(a[0], o[0], o[1], n_blocks=NBLK)
        # parity probe on this very launch shape: 8 sampled channels against the oracle (first call: zero history)
        h2 = aux.IQGeneratorBatch(NCH); h2.process(X, oi, oq, n_blocks=NBLK); torch.cuda.synchronize()
        pick = [0, 1, 777, 2048, 4095]
        want = A.run("iq", (X[pick].cpu().numpy(),), [])
        parity = dict(channels=len(pick), samples=len(pick) * ns, bit_exact=bool(np.array_equal(oi[pick].cpu().numpy(), want[0]) and np.array_equal(oq[pick].cpu().numpy(), want[1])))
        h2.close()
    else:
        I, Q = noise(), noise(); oi, oq = torch.empty_like(I), torch.empty_like(Q)
        if path == "preproc_detect":
            # one click per block over a little noise: a flat spectrum, so no line is ever "strong" (PP.cpp:104) and the
            # detector keeps running for the whole measurement on the GPU and on the CPU baseline alike
            I, Q = (I.float() / 100.0).round().to(torch.int16), (Q.float() / 100.0).round().to(torch.int16)
            I[:, ::128] = 20000; Q[:, ::128] = 20000
        h = aux.PreProcessorBatch(NCH)
        if path == "preproc_static":
            h.setI2SerrorCompensation(None, 1)
            kernel, cpu_events = "pp_static_kernel", [(None, 0, "setI2SerrorCompensation", 1)]
        else:
            h.startAutoI2SerrorDetection()
            kernel, cpu_events = "pp_detect_kernel", [(None, 0, "startAutoI2SerrorDetection")]
        fn = lambda: h.process(I, Q, oi, oq, n_blocks=NBLK, stream=stream)
        bytes_per_sample, planes_in, planes_out = 8.0, 2, 2
        r = np.random.default_rng(5)
        cpu_planes = tuple(np.round(r.normal(0, 3000, (16, 64 * 128))).astype(np.int16) for _ in range(2))
        if path == "preproc_detect":
            cpu_planes = tuple(np.round(a / 100.0).astype(np.int16) for a in cpu_planes)
            for a in cpu_planes:
                a[:, ::128] = 20000
        kind = "pp"
        host_fn = lambda a, o: h.process_host(a[0], a[1], o[0], o[1], n_blocks=NBLK)
        h2 = aux.PreProcessorBatch(NCH)
        (h2.setI2SerrorCompensation(None, 1) if path == "preproc_static" else h2.startAutoI2SerrorDetection())
        h2.process(I, Q, oi, oq, n_blocks=NBLK); torch.cuda.synchronize()
        pick = [0, 1, 777, 2048, 4095]
        want = A.run("pp", (I[pick].cpu().numpy(), Q[pick].cpu().numpy()), cpu_events)
        parity = dict(channels=len(pick), samples=len(pick) * ns, bit_exact=bool(np.array_equal(oi[pick].cpu().numpy(), want[0]) and np.array_equal(oq[pick].cpu().numpy(), want[1])))
        h2.close()
    l0 = h.launch_count
    total_ms, per = time_steps(fn, args.steps, args.warmup, stream)
    launches = (h.launch_count - l0) * args.steps // (args.steps + max(args.warmup, 3))
    samples = float(NCH) * ns
    value = samples * args.steps / (total_ms * 1e-3) / 1e6
    launch_s = float(np.mean(per)) * 1e-3
    # end to end with pinned host planes
    hin = [torch.empty((NCH, ns), dtype=torch.int16).pin_memory() for _ in range(planes_in)]
    hout = [torch.empty((NCH, ns), dtype=torch.int16).pin_memory() for _ in range(planes_out)]
    for t in hin:
        t.copy_((torch.randn((NCH, ns)) * 3000.0).round().clamp(-32768, 32767).to(torch.int16))
    a, o = [t.numpy() for t in hin], [t.numpy() for t in hout]
    host_fn(a, o)
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        host_fn(a, o)
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    e2e = dict(value=samples / e2e_s / 1e6, unit=UNIT, h2d_bytes_per_step=int(planes_in * samples * 2), d2h_bytes_per_step=int(planes_out * samples * 2),
               steps=args.e2e_steps, api="sdr_%s_process_host" % ("iqgen" if path == "iqgen" else "preproc"))
    ach = bytes_per_sample * samples / launch_s / 1e9
    roofline = dict(bound="hbm", achieved=ach, peak=hbm, unit="GB/s", frac=ach / hbm, traffic=None, peak_source=hbm_src, kernel=kernel,
                    algorithmic_bytes_per_sample=bytes_per_sample)
    extra = {}
    if path == "iqgen":
        try:
            rate = fp32_issue_rate()
            extra["roofline_fp32_issue"] = dict(bound="fp32 issue", achieved=192.0 * samples / launch_s / 1e9, peak=rate / 1e9, unit="Ginstr/s",
                                                frac=192.0 * samples / launch_s / rate, algorithmic_instr_per_sample=192.0,
                                                peak_source="measured live: unfused FMUL+FADD microbenchmark (parity forbids FMA contraction)")
        except Exception as e:
            extra["roofline_fp32_issue"] = dict(error=str(e))
    cpu = None if args.no_cpu_baseline else cpu_rate(kind, cpu_planes, cpu_events, args.cpu_seconds)
    line = dict(path=path, metric="channel_samples_per_s", value=value, unit=UNIT, n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="i16/f32", data="synthetic",
                config=dict(workload="%s: %d channels x %d blocks per step, int16 planes in HBM" % (path, NCH, NBLK), channels_per_gpu=NCH, blocks_per_step=NBLK,
                            l2="inputs per step %d MB >> 126 MB L2, no flush needed" % (planes_in * samples * 2 / 1e6)),
                e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, parity=parity,
                per_launch_ms=dict(mean=float(np.mean(per)), min=float(np.min(per)), max=float(np.max(per))), **extra)
    print(json.dumps(line), flush=True)
    h.close()


def bench_spectrum(args):
    """Spectrum tap on the grabber's snapshots: one 256-point complex FFT + power per channel and launch."""
    import torch
    from audiosdr_b200 import aux
    from oracle import aux_lib as A
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream()
    nch = 65536
    g = torch.Generator(device=dev); g.manual_seed(99)
    I = (torch.randn((nch, 256), generator=g, device=dev) * 3000.0).round().clamp(-32768, 32767).to(torch.int16)
    Q = (torch.randn((nch, 256), generator=g, device=dev) * 3000.0).round().clamp(-32768, 32767).to(torch.int16)
    t = torch.arange(256, device=dev, dtype=torch.float32)
    I += (8000.0 * torch.cos(2 * np.pi * 37.0 / 256.0 * t)).round().to(torch.int16)[None, :]
    Q += (8000.0 * torch.sin(2 * np.pi * 37.0 / 256.0 * t)).round().to(torch.int16)[None, :]
    h = aux.GrabberBatch(nch)
    h.process(I, Q, n_blocks=2, stream=stream)
    power = torch.empty((nch, 256), dtype=torch.float32, device=dev)
    fn = lambda: h.spectrum_device(power, stream=stream)
    assert fn()
    torch.cuda.synchronize()
    pick = [0, 1, 777, 40000, nch - 1]
    snap = h.grab(pick)
    want = A.grab_spectrum(snap)
    got = power[pick].cpu().numpy()
    parity = dict(channels=len(pick), bins=int(want.size), bit_exact=bool(np.array_equal(got.view(np.uint32), want.view(np.uint32))),
                  peak_bin=int(np.argmax(got[0])))
    total_ms, per = time_steps(fn, args.steps, args.warmup, stream)
    launch_s = float(np.mean(per)) * 1e-3
    hbm, hbm_src = measured_hbm()
    ach = 2048.0 * nch / launch_s / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        sn = np.ascontiguousarray(np.tile(snap, (200, 1)))
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < min(args.cpu_seconds, 3.0):
            A.grab_spectrum(sn); reps += 1
        dt = time.perf_counter() - t0
        cpu = dict(value=reps * sn.shape[0] * 256 / dt / 1e6, unit=UNIT, cores=1, kind="port", sample="%d snapshots x %d passes through the oracle's FFT on one core" % (sn.shape[0], reps))
    line = dict(path="grabber_spectrum", metric="channel_samples_per_s", value=256.0 * nch / launch_s / 1e6, unit=UNIT, spectra_per_s=nch / launch_s,
                n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="grabber spectrum tap: %d channels, one 256-point complex FFT + power per channel and launch" % nch, channels_per_gpu=nch),
                gpu_launches=args.steps,
                roofline=dict(bound="hbm", achieved=ach, peak=hbm, unit="GB/s", frac=ach / hbm, traffic=None, peak_source=hbm_src, kernel="grab_spectrum_kernel",
                              algorithmic_bytes_per_spectrum=2048.0),
                cpu_baseline=cpu, parity=parity, per_launch_ms=dict(mean=float(np.mean(per)), min=float(np.min(per)), max=float(np.max(per))))
    print(json.dumps(line), flush=True)
    h.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", default="all")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=5.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    from audiosdr_b200 import build
    build.build_aux_library()
    for p in (["iqgen", "preproc_static", "preproc_detect", "grabber_spectrum"] if args.path == "all" else [args.path]):
        if p == "grabber_spectrum":
            bench_spectrum(args)
        else:
            bench_path(p, args)


if __name__ == "__main__":
    main()
