#!/bin/bash
# tools/gpu_w34.sh -- GPU tests, then the diagnostic workloads that are bound by one serial stage: SAM (BASELINE config 3, PLL) and
# the headline workload with the ALS filter switched on everywhere (LMS), each with the product kernel and with the profiling twin
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for args in "--workload 3" "--workload 3 --role-profile" "--variant als" "--variant als --role-profile"; do timeout 200 python bench.py $args --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>gpurun_out/w34.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$args', round(d['value']), d['parity'], {k:{a:round(b,3) for a,b in v.items()} for k,v in (d['role_profile'] or {}).items()})" || tail -3 gpurun_out/w34.err; done
