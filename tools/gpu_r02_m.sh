#!/bin/bash
# tools/gpu_r02_m.sh -- tree after the bulk-copy switch: GPU tests, every workload, ALS taps ten per trip against five, ALS placement.
set -u
mkdir -p gpurun_out
TAG=${1:-r02m}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name: %.0f Msps  ms/step %.3f  parity %s e2e %.0f' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact'), d['e2e']['value']))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
for w in 2 3 5; do BARGS="--workload $w"; run w${w} X=1; done
BARGS="--workload 4"
run w4 X=1
run w4_tap5 SDR_LIB=variants/alstap5.so
run w4_oldmap SDR_MAP_SSB=0x3BADC548961720
run w4_tap5_oldmap SDR_LIB=variants/alstap5.so SDR_MAP_SSB=0x3BADC548961720
