#!/bin/bash
# tools/gpu_check.sh -- the standard on-box sequence (run under gpurun): smoke, GPU tests, bench, ncu evidence.
# Everything lands in gpurun_out/ ; each stage has its own timeout so a hang cannot eat the lease.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== role profile"; timeout 600 python bench.py --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_roles.json 2>&1; python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_roles.json').read().strip().splitlines()[-1]); print(d['value'], json.dumps(d['role_profile']))"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sdr_ -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --blocks-per-step 64 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sdr_pipeline -s 3 -c 1 -f -o gpurun_out/${TAG}_pipeline \
    python bench.py --steps 1 --warmup 3 --blocks-per-step 32 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
