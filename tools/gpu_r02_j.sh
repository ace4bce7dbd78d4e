#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02j}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
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
BARGS=""; run w2_a X=1; run w2_b X=1
BARGS="--workload 5"; run w5_a X=1; run w5_b X=1
BARGS="--workload 3"; run w3 X=1
BARGS="--workload 4"; run w4 X=1
echo "== spectrum debug"; timeout 120 python tools/spec_debug.py 2>&1 | tail -4
