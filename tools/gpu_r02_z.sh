#!/bin/bash
# tools/gpu_r02_z.sh -- config 3 strong-scaled: the shard one rank gets at N = 2, 4, 8 (SDR_BENCH_WORLD) under the plan the host picks and under
# the alternatives (merged plan at 2 or 3 groups per SM, lean 11-warp plan at 2 per SM)
set -u
mkdir -p gpurun_out
TAG=${1:-r02z}
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-headline --workload 3 > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name: %d channels  %.0f Msps  ms/step %.3f  parity %s' % (d['config']['channels_per_gpu'], d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact')))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-400:])
PY
}
for n in 2 4 8 16; do
  run n${n}_default SDR_BENCH_WORLD=$n
  run n${n}_merged3 SDR_BENCH_WORLD=$n SDR_TILE_ENV=16 SDR_CTAS_PER_SM=3 SDR_IN_DEPTH=1
  run n${n}_merged2 SDR_BENCH_WORLD=$n SDR_TILE_ENV=16 SDR_CTAS_PER_SM=2 SDR_IN_DEPTH=1
  run n${n}_lean2 SDR_BENCH_WORLD=$n SDR_TILE_ENV=16 SDR_CTAS_PER_SM=2 SDR_NO_MERGE=1
done
