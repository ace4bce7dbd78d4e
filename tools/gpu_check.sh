#!/bin/bash
# tools/gpu_check.sh -- the standard on-box sequence (run under gpurun): smoke, GPU tests, bench (both arms), diagnostics,
# ncu evidence.  Everything lands in gpurun_out/ ; each stage has its own timeout so a hang cannot eat the lease.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== role profile"; SDR_ROLE_PROFILE_NB=1 timeout 600 python bench.py --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_roles.json 2>&1
python - <<PY
import json
try:
    lines=open('gpurun_out/${TAG}_roles.json').read().strip().splitlines()
    d=json.loads(lines[-1]); rp=d['role_profile']['ssb']; cyc=rp.pop('cta_cycles_per_launch'); steps=d['config']['blocks_per_step']*4+8
    print('role profile: %.0f Msps, cycles/step %.0f, busy kcycles/tile:'%(d['value'],cyc/steps), {k:round(v*cyc/steps/1000,1) for k,v in rp.items()})
    print([l for l in lines if l.startswith('[sdr]')])
except Exception as e: print('role profile failed', e)
PY
for w in 3 5; do
  echo "== diagnostic workload $w (product kernel, then the profiling twin with per-stage busy fractions)"
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_diag_w$w.json 2> gpurun_out/${TAG}_diag_w$w.err
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --role-profile > gpurun_out/${TAG}_diag_w${w}_roles.json 2>> gpurun_out/${TAG}_diag_w$w.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_diag_w$w.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/${TAG}_diag_w${w}_roles.json').read().strip().splitlines()[-1])
    print('workload $w:', round(d['value']), 'Msps', d['parity'], '| twin', round(r['value']), {k:{kk:round(vv,3) for kk,vv in v.items()} for k,v in (r['role_profile'] or {}).items()})
except Exception as e: print('diag failed', e); print(open('gpurun_out/${TAG}_diag_w$w.err').read()[-500:])"
done
echo "== diagnostic: headline workload with the ALS filter on every channel"
timeout 600 python bench.py --variant als --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_diag_als.json 2> gpurun_out/${TAG}_diag_als.err
python -c "
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_diag_als.json').read().strip().splitlines()[-1]); print('als variant:', round(d['value']), 'Msps')
except Exception as e: print('als diag failed', e)"
echo "== compute-sanitizer (all-mode smoke case)"
for t in memcheck racecheck synccheck initcheck; do
  timeout 420 compute-sanitizer --tool $t --kernel-name kernel_substring=sdr_ --print-limit 3 python tools/sanitize_smoke.py > gpurun_out/${TAG}_sanitize_$t.log 2>&1
  echo "$t rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_sanitize_$t.log | head -1)"
done
echo "== 120 s WSPR drift check (BASELINE config 5, sampled channels, full length)"
timeout 900 python tools/long_run_check.py > gpurun_out/${TAG}_long_run.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_long_run.log
echo "== ncu launch list (same command as the bench line, fewer steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sdr_ -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (dominant kernel, default bench configuration)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sdr_pipeline -s 3 -c 1 -f -o gpurun_out/${TAG}_pipeline \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | grep ${TAG}
