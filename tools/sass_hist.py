#!/usr/bin/env python3
"""tools/sass_hist.py [lib.so ...] -- per-kernel SASS opcode histogram of the shipped libraries (cuobjdump -sass), written to
profiles/<tag>_sass_<lib>.txt: instruction count per kernel, the mnemonics that matter for this design (packed FP32 FFMA2/FADD2,
asynchronous copies LDGSTS, mbarrier SYNCS, bulk/tensor copies UBLKCP/UTMALDG/UTMASTG, barriers BAR) and the full histogram."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = os.environ.get("SASS_TAG", "r02")
libs = sys.argv[1:] or [os.path.join(ROOT, "audiosdr_b200", "libsdr_batch.so"), os.path.join(ROOT, "audiosdr_b200", "libsdr_aux.so")]
KEY = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FMUL", "FADD", "MUFU", "DFMA", "DADD", "DMUL", "F2F", "LDGSTS", "SYNCS", "UBLKCP", "UTMALDG", "UTMASTG",
       "BAR", "WARPSYNC", "LDS", "STS", "LDG", "STG", "LDC", "LDCU", "IMAD", "BRA"]
for lib in libs:
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    hist = collections.OrderedDict()
    fn = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1); hist[fn] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    out = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, os.path.basename(lib).replace(".so", "")))
    with open(out, "w") as f:
        f.write("# cuobjdump -sass %s : opcode histogram per kernel (tools/sass_hist.py)\n" % os.path.relpath(lib, ROOT))
        for fn, h in hist.items():
            f.write("\n== %s: %d instructions\n" % (fn, sum(h.values())))
            f.write("   key: " + "  ".join("%s=%d" % (k, h[k]) for k in KEY if h[k]) + "\n")
            f.write("   all: " + "  ".join("%s=%d" % kv for kv in h.most_common()) + "\n")
    print(out)
