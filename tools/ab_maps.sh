#!/bin/bash
# tools/ab_maps.sh [lib.so] map... -- headline bench value per stage placement (SDR_MAP_SSB), optionally with another build of the library
LIB=audiosdr_b200/libsdr_batch.so
if [ -f "$1" ]; then cp $LIB /tmp/keep.so; cp "$1" $LIB; touch $LIB; shift; fi
for m in default "$@"; do
  if [ "$m" = default ]; then unset SDR_MAP_SSB; else export SDR_MAP_SSB=$m; fi
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MAP $m -> %.0f Msps bit_exact=%s' % (d['value'], d['parity']['bit_exact']))"
done
[ -f /tmp/keep.so ] && cp /tmp/keep.so $LIB && touch $LIB
