#!/bin/bash
# tools/gpu_r02_f.sh -- A/B against the round-1 library after the fixed-plan changes, role profiles, placement search for the lean ENV plan.
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env SDR_DEBUG_PLAN=1 "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name: %.0f Msps  ms/step %.3f  parity %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact')))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
BARGS=""; run w2_tree X=1; run w2_r01 SDR_LIB=variants/r01.so
BARGS="--workload 5"; run w5_tree X=1; run w5_r01 SDR_LIB=variants/r01.so
BARGS="--workload 3"; run w3_tree X=1
for w in 2 5; do
echo "== role profile w$w"; SDR_ROLE_PROFILE_NB=1 timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 --only-headline > gpurun_out/${TAG}_w${w}_roles.json 2>&1
tail -c 3000 gpurun_out/${TAG}_w${w}_roles.json | grep -o '"role_profile.*' | cut -c1-700; grep '^\[sdr\]' gpurun_out/${TAG}_w${w}_roles.json | cut -c1-400
done
echo "== placement search, lean ENV plan (16-sample tiles, two groups per SM)"
timeout 400 python tools/map_search.py --cls envlean --seconds 150 --blocks 32 > gpurun_out/${TAG}_map_envlean.log 2>&1; tail -12 gpurun_out/${TAG}_map_envlean.log
