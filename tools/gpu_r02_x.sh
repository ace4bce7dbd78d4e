#!/bin/bash
# tools/gpu_r02_x.sh -- the default bench line (streamed e2e after its own warm-up, clock samples inside the timed region), then the
# placement of ENV buckets with the blanker searched on BASELINE config 4 itself (two-launch form forced at 4 096 channels)
set -u
mkdir -p gpurun_out
TAG=${1:-r02x}
echo "== bench"; ( time timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value %.0f e2e %.0f sync %.0f link %.0f frac %.3f clocks %s' % (d['value'], e['value'], e['per_call_sync']['value'], e['link_bound']['value'], e['link_frac'], d['clocks']))
print({k: round(v['value']) for k, v in d['workloads'].items()})
PY
echo "== config 4, ENV placement"
timeout 400 python tools/map_search.py --cls env --config 4 --split --blocks 64 --seconds 200 --start A0D459B1328C67 > gpurun_out/${TAG}_map_env_w4.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_env_w4.log | tail -14
