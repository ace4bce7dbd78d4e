#!/bin/bash
# tools/gpu_r02_v.sh -- placement search for ENV buckets WITH the blanker on the 32-sample plan (BASELINE config 4's AM and SAM buckets; the
# default ENV placement was measured without the blanker's three warps), then the 2-GPU bench pair
set -u
mkdir -p gpurun_out
TAG=${1:-r02v}
echo "== SAM + blanker"
timeout 400 python tools/map_search.py --cls env --config 3 --nb --mode 5 --seconds 170 --start A0D459B1328C67 > gpurun_out/${TAG}_map_sam_nb.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_sam_nb.log | tail -12
echo "== AM + blanker"
timeout 400 python tools/map_search.py --cls env --config 3 --nb --mode 4 --seconds 170 --start A0D459B1328C67 > gpurun_out/${TAG}_map_am_nb.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_am_nb.log | tail -12
