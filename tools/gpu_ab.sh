#!/bin/bash
# tools/gpu_ab.sh [map-search seconds] -- A/B experiments under gpurun: every variants/*.so (builds of the receiver library made
# in the container) takes the place of audiosdr_b200/libsdr_batch.so in turn and runs the headline
# bench (which carries the bit-exact parity probe); the library built from the tree then gets a role profile and,
# optionally, a stage-placement search.  Scratch output in gpurun_out/ab_*.
set -u
mkdir -p gpurun_out
SEARCH=${1:-0}
LIB=audiosdr_b200/libsdr_batch.so
cp $LIB /tmp/tree_lib.so
one() { # name
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$1.json').read().strip().splitlines()[-1])
    print('AB %-24s %8.0f Msps  %.3f ms  bit_exact=%s  e2e %.0f' % ('$1', d['value'], d['per_launch_ms']['mean'], d['parity']['bit_exact'], d['e2e']['value']))
except Exception as e:
    print('AB $1 failed', e); print(open('gpurun_out/ab_$1.err').read()[-800:])
PY
}
for v in variants/*.so; do
  [ -e "$v" ] || continue
  cp "$v" $LIB; touch $LIB
  one "$(basename "$v" .so)"
done
cp /tmp/tree_lib.so $LIB; touch $LIB
one tree
SDR_ROLE_PROFILE_NB=1 timeout 600 python bench.py --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_roles.json 2>&1
python - <<PY
import json
try:
    lines=open('gpurun_out/ab_roles.json').read().strip().splitlines()
    d=json.loads(lines[-1]); rp=d['role_profile']['ssb']; cyc=rp.pop('cta_cycles_per_launch'); steps=d['config']['blocks_per_step']*4+8
    print('role profile: %.0f Msps, cycles/step %.0f, busy kcycles/tile:'%(d['value'],cyc/steps), {k:round(v*cyc/steps/1000,1) for k,v in rp.items()})
except Exception as e: print('role profile failed', e)
PY
if [ "$SEARCH" != "0" ]; then
  timeout $((SEARCH + 120)) python tools/map_search.py --seconds $SEARCH > gpurun_out/ab_map_search.log 2>&1
  tail -12 gpurun_out/ab_map_search.log
fi
