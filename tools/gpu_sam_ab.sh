#!/bin/bash
# tools/gpu_sam_ab.sh -- A/B on the SAM diagnostic workload: every variants/*.so in place of the tree's library, then the tree's
# library with the default and with an alternative ENV stage placement ($1)
LIB=audiosdr_b200/libsdr_batch.so
cp $LIB /tmp/tree_lib.so
one() { timeout 200 python bench.py --workload 3 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SAM $1: %.0f Msps %s' % (d['value'], d['parity']['bit_exact']))"; }
for v in variants/*.so; do [ -e "$v" ] || continue; cp "$v" $LIB; touch $LIB; one "$(basename $v .so)"; done
cp /tmp/tree_lib.so $LIB; touch $LIB
one tree
if [ -n "${1:-}" ]; then SDR_MAP_ENV=$1 one "tree map $1"; fi
one tree-again
