#!/bin/bash
# tools/gpu_r02_d.sh -- A/B on the box: fixed 32-sample plan in lock step vs r01, lean co-resident plans in both sync forms.
set -u
mkdir -p gpurun_out
TAG=${1:-r02d}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env SDR_DEBUG_PLAN=1 "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    plan=[l.strip() for l in open('gpurun_out/${TAG}_$name.err') if l.startswith('[sdr] launch')][:1]
    print('$name: %.0f Msps  ms/step %.3f  parity %s | %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact'), plan[0][14:110] if plan else ''))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-400:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
BARGS=""
run w2_tree X=1
run w2_r01 SDR_LIB=variants/r01.so
run w2_handover32 SDR_LIB=variants/handover32.so
run w2_runtimeplan SDR_LIB=variants/runtimeplan.so
BARGS="--workload 5"
run w5_tree X=1
run w5_r01 SDR_LIB=variants/r01.so
run w5_T16x2 SDR_TILE_SSB=16 SDR_CTAS_PER_SM=2
run w5_T16x2_lock SDR_TILE_SSB=16 SDR_CTAS_PER_SM=2 SDR_LIB=variants/leanlock.so
BARGS="--workload 3"
run w3_tree X=1
run w3_r01 SDR_LIB=variants/r01.so
run w3_T16x2 SDR_TILE_ENV=16 SDR_CTAS_PER_SM=2
run w3_T16x2_lock SDR_TILE_ENV=16 SDR_CTAS_PER_SM=2 SDR_LIB=variants/leanlock.so
run w3_T16x2_s0 SDR_TILE_ENV=16 SDR_CTAS_PER_SM=2 SDR_SLACK=0
run w3_T8x2_lock SDR_TILE_ENV=8 SDR_CTAS_PER_SM=2 SDR_LIB=variants/leanlock.so
run w3_T8x3_lock SDR_TILE_ENV=8 SDR_CTAS_PER_SM=3 SDR_SLACK=0 SDR_LIB=variants/leanlock.so
BARGS="--variant als"
run als_tree X=1
run als_r01 SDR_LIB=variants/r01.so
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
