#!/bin/bash
# tools/gpu_chunks.sh -- e2e (host int16 planes through sdr_batch_process_host) against the number of time chunks per call
IFS=";" read -ra CFGS <<< "${CHUNK_CFGS:-12 16;16 16;24 8;32 8;48 4;64 4}"
for cfg in "${CFGS[@]}"; do set -- $cfg
  SDR_HOST_CHUNKS=$1 SDR_HOST_MIN_CHUNK=$2 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $1 (min $2 blocks): e2e %.0f Msps, link bound %.0f, frac %.3f' % (d['e2e']['value'], d['e2e']['link_bound']['value'], d['e2e']['link_frac']))"
done
