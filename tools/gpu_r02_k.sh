#!/bin/bash
# tools/gpu_r02_k.sh -- placement search for the buckets with ALS (the ALS + output stage is their slowest), then config 4 with the result.
set -u
mkdir -p gpurun_out
TAG=${1:-r02k}
echo "== placement search, SSB buckets with ALS (headline workload + ALS on every channel)"
timeout 400 python tools/map_search.py --cls ssb --als --seconds 170 --blocks 64 > gpurun_out/${TAG}_map_ssb_als.log 2>&1; tail -6 gpurun_out/${TAG}_map_ssb_als.log
BEST=$(grep -E "^evaluated" gpurun_out/${TAG}_map_ssb_als.log | sed 's/.*best \([0-9A-F]*\) .*/\1/')
echo "best SSB+ALS map: $BEST"
echo "== placement search, ENV buckets with ALS (SAM workload + blanker + ALS on every channel)"
timeout 400 python tools/map_search.py --cls env --als --seconds 120 --blocks 64 > gpurun_out/${TAG}_map_env_als.log 2>&1; tail -6 gpurun_out/${TAG}_map_env_als.log
BESTE=$(grep -E "^evaluated" gpurun_out/${TAG}_map_env_als.log | sed 's/.*best \([0-9A-F]*\) .*/\1/')
echo "best ENV+ALS map: $BESTE"
run() { # name, env..., (BARGS)
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name: %.0f Msps  ms/step %.3f  parity %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact')))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
BARGS="--workload 4"; run w4_default X=1; run w4_ssbmap SDR_MAP_SSB=$BEST; run w4_bothmaps SDR_MAP_SSB=$BEST SDR_MAP_ENV=$BESTE
BARGS="--variant als"; run als_default X=1; run als_map SDR_MAP_SSB=$BEST
