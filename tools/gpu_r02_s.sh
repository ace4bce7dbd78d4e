#!/bin/bash
# tools/gpu_r02_s.sh -- placement search for SSB buckets WITHOUT the blanker (BASELINE config 5: the blanker's three warps idle, so the
# placement measured on config 2 leaves two Hilbert warps + IN + OUT on one scheduler), and a re-search for config 2 on the class kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r02s}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then tail -5 gpurun_out/${TAG}_smoke.log; exit 1; fi
echo "== config 5 from the balanced start"
timeout 400 python tools/map_search.py --cls ssb --config 5 --seconds 200 --start 9C87654A320BD1 > gpurun_out/${TAG}_map_w5_a.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_w5_a.log | tail -14
echo "== config 5 from the default"
timeout 300 python tools/map_search.py --cls ssb --config 5 --seconds 100 --start 3BADC548961720 > gpurun_out/${TAG}_map_w5_b.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_w5_b.log | tail -12
echo "== config 2 from the default"
timeout 400 python tools/map_search.py --cls ssb --config 2 --seconds 150 --start 3BADC548961720 > gpurun_out/${TAG}_map_w2.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_w2.log | tail -12
