#!/bin/bash
# tools/gpu_check_r02.sh [tag] -- the on-box evidence sequence of round 2 (run under gpurun): smoke, GPU tests, both bench arms,
# ncu launch list + full captures (config 2, config 3 and ALS post-pass kernels), sanitizers, aux benches.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${TAG}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
echo "== bench"; ( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
for w in 2 3 5; do
  echo "== role profile, config $w"
  SDR_ROLE_PROFILE_NB=1 timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 --only-headline > gpurun_out/${TAG}_roles_w$w.json 2> gpurun_out/${TAG}_roles_w$w.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_roles_w$w.json').read().strip().splitlines()[-1]); print(round(d['value']), {k:{kk:round(vv,3) for kk,vv in v.items()} for k,v in (d['role_profile'] or {}).items()})
except Exception as e: print('failed', e)"
done
echo "== aux benches"; timeout 600 python bench_aux.py > gpurun_out/${TAG}_aux_bench.json 2> gpurun_out/${TAG}_aux_bench.err; cut -c1-260 gpurun_out/${TAG}_aux_bench.json
echo "== compute-sanitizer (all-mode smoke case)"
for t in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $t --kernel-name kernel_substring=sdr_ --print-limit 3 python tools/sanitize_smoke.py > gpurun_out/${TAG}_sanitize_$t.log 2>&1
  echo "$t rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_$t.log | head -1)"
done
echo "== ncu launch list (same command as the bench line, fewer steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sdr_ -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full: config 2 kernel (32-sample plan)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sdr_pipeline -s 3 -c 1 -f -o gpurun_out/${TAG}_pipeline_w2 \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --only-headline > gpurun_out/${TAG}_ncu_full_w2.log 2>&1; echo "ncu full rc=$?"
echo "== ncu full: config 3 kernel (merged SAM plan, three groups per SM)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sdr_pipeline -s 3 -c 1 -f -o gpurun_out/${TAG}_pipeline_w3 \
    python bench.py --workload 3 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --only-headline > gpurun_out/${TAG}_ncu_full_w3.log 2>&1; echo "ncu full rc=$?"
echo "== ncu launch list + full capture: config 4 (ALS buckets as two launches; the post-pass kernel)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sdr_ -c 60 --csv --log-file gpurun_out/${TAG}_launches_w4.csv \
    python bench.py --workload 4 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --only-headline > gpurun_out/${TAG}_ncu_launch_w4.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sdr_als_pass -s 3 -c 1 -f -o gpurun_out/${TAG}_als_pass \
    python bench.py --workload 4 --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --only-headline > gpurun_out/${TAG}_ncu_full_w4.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | grep ${TAG}_ | head -40
