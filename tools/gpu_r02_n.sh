#!/bin/bash
# tools/gpu_r02_n.sh -- the two-launch form of ALS buckets: GPU tests, sanitizers, config 4 against the one-launch form, the other workloads unchanged.
set -u
mkdir -p gpurun_out
TAG=${1:-r02n}
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
echo "== pytest gpu (split ALS)"; timeout 900 python -m pytest tests -m gpu -x -q -k "split or als or full_channel" > gpurun_out/${TAG}_pytest_als.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_als.log
BARGS="--workload 4"
run w4 SDR_DEBUG_PLAN=1
grep "sdr\]" gpurun_out/${TAG}_w4.err | head -8
run w4_nosplit SDR_ALS_SPLIT=0
run w4_splitall SDR_ALS_SPLIT=1
for w in 2 3 5; do BARGS="--workload $w"; run w${w} X=1; done
for t in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $t --kernel-name kernel_substring=sdr_ --print-limit 3 python tools/sanitize_smoke.py > gpurun_out/${TAG}_sanitize_$t.log 2>&1
  echo "$t rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_$t.log | head -1)"
done
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
