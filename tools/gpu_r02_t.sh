#!/bin/bash
# tools/gpu_r02_t.sh -- new placements (config 2 re-searched, SSB buckets without blanker), the skewed cascade (variants/skew.so, -DSDR_CASCADE_SKEW)
# against the product, the streaming host calls (e2e), then the GPU tests.
set -u
mkdir -p gpurun_out
TAG=${1:-r02t}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 5 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    e=d['e2e']
    print('$name: %.0f Msps  ms/step %.3f  parity %s e2e %.0f (sync calls %.0f, link %.0f)' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact'), e['value'], e.get('per_call_sync',{}).get('value',0), e.get('link_bound',{}).get('value',0)))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
for rep in 1 2; do
for w in 2 5 3 4; do
  BARGS="--workload $w"; run w${w}_base_$rep X=1; run w${w}_skew_$rep SDR_LIB=variants/skew.so
done
done
echo "== pytest gpu (new + quick ones)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
echo "== pytest with the skewed cascade"; SDR_LIB=variants/skew.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "per_config or golden or short_tile or split_als or every_bucket" > gpurun_out/${TAG}_pytest_skew.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_skew.log
