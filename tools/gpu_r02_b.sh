#!/bin/bash
# tools/gpu_r02_b.sh -- A/B on the box: sync mechanism / fixed tile length variants (variants/*.so via SDR_LIB), lean co-resident plans.
set -u
mkdir -p gpurun_out
TAG=${1:-r02b}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name: %.0f Msps  ms/step %.3f  parity %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact')))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-400:])
PY
}
BARGS=""
run w2_tree X=1
for v in r01 lockstep fixedT lockfixed sleep hint; do run w2_$v SDR_LIB=variants/$v.so; done
BARGS="--workload 5"
run w5_tree X=1
run w5_r01 SDR_LIB=variants/r01.so
run w5_T16x2 SDR_TILE_SSB=16 SDR_CTAS_PER_SM=2
run w5_T16x2_s0 SDR_TILE_SSB=16 SDR_CTAS_PER_SM=2 SDR_SLACK=0
run w5_T8x2 SDR_TILE_SSB=8 SDR_CTAS_PER_SM=2
run w5_lockstep SDR_LIB=variants/lockstep.so
BARGS="--workload 3"
run w3_tree X=1
run w3_T16x2 SDR_TILE_ENV=16 SDR_CTAS_PER_SM=2
run w3_T8x2 SDR_TILE_ENV=8 SDR_CTAS_PER_SM=2
run w3_lockstep SDR_LIB=variants/lockstep.so
