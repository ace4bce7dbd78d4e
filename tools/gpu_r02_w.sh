#!/bin/bash
# tools/gpu_r02_w.sh -- AGC in two sweeps (level recurrence, then the table look-ups four in flight) against the single-sweep form
# (variants/agc_old.so), placement re-searched with it, and the ENV-with-blanker placements (tools/gpu_r02_v.sh)
set -u
mkdir -p gpurun_out
TAG=${1:-r02w}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 5 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    e=d['e2e']
    print('$name: %.0f Msps  ms/step %.3f  parity %s e2e %.0f (sync calls %.0f, link %.0f) clocks %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact'), e['value'], e.get('per_call_sync',{}).get('value',0), e.get('link_bound',{}).get('value',0), d['clocks'].get('samples')))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
for rep in 1 2; do
for w in 2 5 3 4; do
  BARGS="--workload $w"; run w${w}_new_$rep X=1; run w${w}_old_$rep SDR_LIB=variants/agc_old.so
done
done
echo "== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
echo "== config 5 placement with the new AGC"
timeout 400 python tools/map_search.py --cls ssb --config 5 --seconds 170 --idle 1CD --start BC84627A3510D9 > gpurun_out/${TAG}_map_w5.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_w5.log | tail -9
echo "== config 2 placement with the new AGC"
timeout 400 python tools/map_search.py --cls ssb --config 2 --seconds 170 --start CBA435D8961720 > gpurun_out/${TAG}_map_w2.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_w2.log | tail -9
bash tools/gpu_r02_v.sh ${TAG}
