#!/bin/bash
# tools/gpu_r02_c.sh -- A/B on the box: per-tile-length kernels, grouped barriers, input prefetch depth, co-resident lean plans.
set -u
mkdir -p gpurun_out
TAG=${1:-r02c}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env SDR_DEBUG_PLAN=1 "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    plan=[l.strip() for l in open('gpurun_out/${TAG}_$name.err') if l.startswith('[sdr] launch')][:1]
    print('$name: %.0f Msps  ms/step %.3f  parity %s | %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact'), plan[0][14:] if plan else ''))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-400:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
BARGS=""
run w2_tree X=1
run w2_lockstep SDR_LIB=variants/lockstep.so
run w2_slack0 SDR_SLACK=0
BARGS="--workload 5"
run w5_T32 X=1
run w5_lockstep SDR_LIB=variants/lockstep.so
run w5_T16x1 SDR_TILE_SSB=16 SDR_CTAS_PER_SM=1
run w5_T16x2 SDR_TILE_SSB=16 SDR_CTAS_PER_SM=2
run w5_T8x2 SDR_TILE_SSB=8 SDR_CTAS_PER_SM=2
BARGS="--workload 3"
run w3_T32 X=1
run w3_lockstep SDR_LIB=variants/lockstep.so
run w3_T16x1 SDR_TILE_ENV=16 SDR_CTAS_PER_SM=1
run w3_T16x2 SDR_TILE_ENV=16 SDR_CTAS_PER_SM=2
run w3_T8x2 SDR_TILE_ENV=8 SDR_CTAS_PER_SM=2
echo "== role profile w3 T16x2"
SDR_TILE_ENV=16 SDR_CTAS_PER_SM=2 SDR_ROLE_PROFILE_NB=1 timeout 300 python bench.py --workload 3 --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_w3_roles.json 2>&1
tail -c 3000 gpurun_out/${TAG}_w3_roles.json | grep -o '"role_profile.*' | cut -c1-900; grep '^\[sdr\]' gpurun_out/${TAG}_w3_roles.json | cut -c1-500
echo "== role profile w5 T16x2"
SDR_TILE_SSB=16 SDR_CTAS_PER_SM=2 SDR_ROLE_PROFILE_NB=1 timeout 300 python bench.py --workload 5 --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_w5_roles.json 2>&1
tail -c 3000 gpurun_out/${TAG}_w5_roles.json | grep -o '"role_profile.*' | cut -c1-900; grep '^\[sdr\]' gpurun_out/${TAG}_w5_roles.json | cut -c1-500
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
