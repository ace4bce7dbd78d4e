#!/bin/bash
# tools/gpu_r02_2gpu.sh -- two ranks under torchrun: the bench line (all workloads, config 3 strong-scaled) and the reference arm.
set -u
mkdir -p gpurun_out
TAG=${1:-r02_2gpu}
echo "== bench, 2 ranks"; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | grep real; tail -4 gpurun_out/${TAG}_bench.err | cut -c1-300
python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/${TAG}_bench.json').read().strip().splitlines() if l.startswith('{')][-1]
    print('bench N=%d: value %.0f e2e %.0f link_frac %.3f parity %s status %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['e2e'].get('link_frac', 0), d['parity'], d['status']))
    for k,w in (d.get('workloads') or {}).items():
        print(' ', k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in w.items() if kk in ('value','ms_per_step','error','scaling','channels_this_rank','n_gpus')}, 'parity', w.get('parity'), 'status', w.get('status'))
    print('  sustained', d['sustained'] and d['sustained']['value'], 'contracting', d['contracting_build'] and d['contracting_build'].get('value'))
except Exception as e:
    print('bench FAILED', e)
PY
echo "== reference arm, 2 ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "rc=$?"; grep -c '^{' gpurun_out/${TAG}_bench_reference.json; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
