#!/bin/bash
# tools/gpu_r02_u.sh -- streamed host calls: e2e against the number of time chunks per call (SDR_HOST_CHUNKS); longer placement search for
# SSB buckets without the blanker
set -u
mkdir -p gpurun_out
TAG=${1:-r02u}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 8 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    e=d['e2e']
    print('$name: %.0f Msps  e2e %.0f (sync calls %.0f, link %.0f)' % (d['value'], e['value'], e.get('per_call_sync',{}).get('value',0), e.get('link_bound',{}).get('value',0)))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then tail -5 gpurun_out/${TAG}_smoke.log; exit 1; fi
for rep in 1 2; do
for c in 12 2 3 4 6 8 16 24; do
  BARGS="--workload 2"; run w2_chunks${c}_$rep SDR_HOST_CHUNKS=$c SDR_HOST_MIN_CHUNK=4
done
done
BARGS="--workload 5"; run w5_chunks4 SDR_HOST_CHUNKS=4; run w5_chunks12 SDR_HOST_CHUNKS=12
echo "== config 5 placement, longer"
timeout 500 python tools/map_search.py --cls ssb --config 5 --seconds 300 --idle 1CD --start BC84627A3510D9 > gpurun_out/${TAG}_map_w5.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_w5.log | tail -14
