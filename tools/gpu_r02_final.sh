#!/bin/bash
# tools/gpu_r02_final.sh -- last check of the round's final build: smoke, the whole GPU test tier, the default bench line, the shard one of
# eight ranks gets of config 3 (lean plan for SAM buckets of at most two groups per SM)
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
echo "== bench"; ( time timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value %.0f e2e %.0f sync %.0f link %.0f frac %.3f clocks %s' % (d['value'], e['value'], e['per_call_sync']['value'], e['link_bound']['value'], e['link_frac'], d['clocks']['samples']))
print({k: round(v['value']) for k, v in d['workloads'].items()}, {k: round(v.get('value', 0)) for k, v in (d['aux_blocks'] or {}).items()})
PY
for n in 8 4; do
  SDR_BENCH_WORLD=$n timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-headline --workload 3 > gpurun_out/${TAG}_w3_n$n.json 2> gpurun_out/${TAG}_w3_n$n.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_w3_n$n.json').read().strip().splitlines()[-1]); print('config 3, shard of $n ranks: %d channels %.0f Msps parity %s' % (d['config']['channels_per_gpu'], d['value'], d['parity']['bit_exact']))"
done
