#!/bin/bash
# tools/gpu_check_aux.sh <tag> -- GPU evidence for the widened rows (pre-processor, I/Q generator): tests, bench lines, ncu.
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "== pytest gpu (aux)"; timeout 900 python -m pytest tests/test_gpu_aux.py -x -q > gpurun_out/${TAG}_aux_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_aux_pytest_gpu.log
echo "== bench_aux"; timeout 900 python bench_aux.py --steps 10 --warmup 3 > gpurun_out/${TAG}_aux_bench.json 2> gpurun_out/${TAG}_aux_bench.err; echo "rc=$?"; tail -3 gpurun_out/${TAG}_aux_bench.err
python - <<PY
import json
for l in open('gpurun_out/${TAG}_aux_bench.json'):
    d=json.loads(l); print(d['path'], '%.0f Msps'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'hbm frac %.3f'%d['roofline']['frac'], 'fp32 issue', d.get('roofline_fp32_issue',{}).get('frac'), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['parity'])
PY
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'iq_|pp_' -c 200 --csv --log-file gpurun_out/${TAG}_aux_launches.csv python bench_aux.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_aux_ncu_launch.log 2>&1; echo "rc=$?"
for k in iq_generate_kernel pp_static_kernel pp_detect_kernel; do
  p=iqgen; [ $k = pp_static_kernel ] && p=preproc_static; [ $k = pp_detect_kernel ] && p=preproc_detect
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${TAG}_$k python bench_aux.py --path $p --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_aux.py -x -q -k "golden or error or feed_forward" > gpurun_out/${TAG}_aux_sanitize_memcheck.log 2>&1; echo "memcheck rc=$? $(grep 'ERROR SUMMARY' gpurun_out/${TAG}_aux_sanitize_memcheck.log | head -1)"
compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_aux.py -x -q -k "golden" > gpurun_out/${TAG}_aux_sanitize_racecheck.log 2>&1; echo "racecheck rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/${TAG}_aux_sanitize_racecheck.log | head -1)"
ls -la gpurun_out | grep ${TAG}_ | grep -i aux
