#!/bin/bash
# tools/gpu_ab2.sh [env map] -- headline and SAM A/B of every variants/*.so against the tree's library
LIB=audiosdr_b200/libsdr_batch.so
cp $LIB /tmp/tree_lib.so
one() { for w in ${AB_WORKLOADS:-2 3}; do timeout 200 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB %-22s workload $w: %.0f Msps bit_exact=%s' % ('$1', d['value'], d['parity']['bit_exact']))"; done; }
for v in variants/*.so; do [ -e "$v" ] || continue; cp "$v" $LIB; touch $LIB; one "$(basename $v .so)"; done
cp /tmp/tree_lib.so $LIB; touch $LIB
one tree
if [ -n "${1:-}" ]; then export SDR_MAP_ENV=$1; one "tree+map"; fi
