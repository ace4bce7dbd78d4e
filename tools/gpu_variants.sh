#!/bin/bash
# tools/gpu_variants.sh -- diagnostics under gpurun: role profile of the bench workload with stages switched off.
set -u
mkdir -p gpurun_out
TAG=${1:-var}
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
for v in "" nonb nonb,noagc nonb,noagc,noaud; do
  timeout 300 python bench.py --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 --variant "$v" > gpurun_out/${TAG}_v_${v//,/_}.json 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_v_${v//,/_}.json').read().strip().splitlines()[-1])
    rp=d['role_profile']['ssb']; cyc=rp.pop('cta_cycles_per_launch')
    steps=d['config']['blocks_per_step']*4+7
    import re
    for l in open('gpurun_out/${TAG}_v_${v//,/_}.json'):
        if l.startswith('[sdr]'): print('   ', l.strip())
    print('variant [%s] %.0f Msps  cycles/step %.0f  busy kcycles/tile:'%('$v',d['value'],cyc/steps), {k:round(v*cyc/steps/1000,1) for k,v in rp.items()})
except Exception as e:
    print('variant [$v] failed', e); print(open('gpurun_out/${TAG}_v_${v//,/_}.json').read()[-600:])
PY
done
