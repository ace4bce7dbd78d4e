#!/bin/bash
# tools/gpu_r02_long.sh -- longer placement searches (hill climbing with restarts) for the two SSB bucket kinds
set -u
mkdir -p gpurun_out
TAG=${1:-r02l2}
echo "== config 2"
timeout 900 python tools/map_search.py --cls ssb --config 2 --seconds 600 --start CBA435D8961720 > gpurun_out/${TAG}_map_w2.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_w2.log | tail -14
echo "== config 5"
timeout 700 python tools/map_search.py --cls ssb --config 5 --seconds 400 --idle 1CD --start BC84627A3510D9 > gpurun_out/${TAG}_map_w5.log 2>&1; grep -E "^start|^best|^evaluated|top" gpurun_out/${TAG}_map_w5.log | tail -14
