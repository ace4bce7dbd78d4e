#!/bin/bash
for m in default 5D18C3A06247B9 5D10C3A86247B9; do
  if [ "$m" = default ]; then unset SDR_MAP_SSB; else export SDR_MAP_SSB=$m; fi
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MAP $m -> %.0f Msps bit_exact=%s' % (d['value'], d['parity']['bit_exact']))"
done
