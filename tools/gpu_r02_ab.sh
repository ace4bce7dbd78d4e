#!/bin/bash
# tools/gpu_r02_ab.sh -- first on-box pass of the hand-over pipeline: smoke, GPU tests, then ring-slack A/B on the three workloads.
set -u
mkdir -p gpurun_out
TAG=${1:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
run() { # name, env..., -- bench args
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    print('$name: %.0f Msps  ms/step %.3f  parity %s  e2e %.0f' % (d['value'], d['ms_per_step'], d['parity'], d['e2e']['value']))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-400:])
PY
}
for s in 0 1 2 3; do BARGS="" run w2_slack$s SDR_SLACK=$s; done
for s in 0 2; do BARGS="--workload 3" run w3_slack$s SDR_SLACK=$s; BARGS="--workload 5" run w5_slack$s SDR_SLACK=$s; done
BARGS="--variant als" run als_slack2 SDR_SLACK=2
echo "== role profile"; SDR_ROLE_PROFILE_NB=1 timeout 300 python bench.py --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_roles.json 2>&1
tail -c 3000 gpurun_out/${TAG}_roles.json | grep -o '"role_profile.*' | cut -c1-1200; grep '^\[sdr\]' gpurun_out/${TAG}_roles.json | cut -c1-600
echo "== racecheck (all-mode smoke case)"
timeout 420 compute-sanitizer --tool racecheck --kernel-name kernel_substring=sdr_ --print-limit 5 python tools/sanitize_smoke.py > gpurun_out/${TAG}_sanitize_racecheck.log 2>&1
echo "racecheck rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_racecheck.log | head -1)"
