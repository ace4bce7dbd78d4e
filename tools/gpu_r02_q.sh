#!/bin/bash
# tools/gpu_r02_q.sh -- (ALS setting as launch parameters when the bucket has one) config 4 in the two-launch ALS form: launch list, density of the post-pass, full capture of the post-pass kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r02q}
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
echo "== pytest gpu (buckets)"; timeout 900 python -m pytest tests -m gpu -x -q -k "every_bucket or split" > gpurun_out/${TAG}_pytest_als.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_als.log
BARGS="--workload 4"
run w4 SDR_DEBUG_PLAN=1
grep "sdr\]" gpurun_out/${TAG}_w4.err | head -8
run w4_no_uniform SDR_ALS_NO_UNIFORM=1
echo "== ncu launch list, config 4"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sdr_ -c 60 --csv --log-file gpurun_out/${TAG}_launches_w4.csv \
    python bench.py --workload 4 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --only-headline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02q_launches_w4.csv')) if len(r) > 10 and r[0].isdigit()]
for r in rows[-16:]:
    print(r[4][:40], r[7], r[-1], r[-2])
PY
echo "== ncu full: ALS post-pass kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sdr_als_pass -s 3 -c 1 -f -o gpurun_out/${TAG}_als_pass \
    python bench.py --workload 4 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --only-headline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
