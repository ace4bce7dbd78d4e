#!/usr/bin/env python3
"""tools/build_variants.py name=-DFLAG[,-DFLAG...] ... -- builds of the receiver library with experiment switches, for A/B runs on the
GPU box: variants/<name>.so (git-ignored, travels with the snapshot).  Select one at run time with SDR_LIB=variants/<name>.so."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiosdr_b200 import build as B

os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = os.path.join(ROOT, "variants", name + ".so")
    cmd = [B._nvcc()] + B.NVCC_FLAGS + ["--threads", "4"] + [f for f in flags.split(",") if f]
    for src in B.SOURCES:
        cmd += ["-x", "cu", os.path.join(B.CSRC, src)]
    cmd += ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(name, "ok" if r.returncode == 0 else "FAILED\n" + r.stdout + r.stderr)
