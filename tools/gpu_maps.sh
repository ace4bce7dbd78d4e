#!/bin/bash
# tools/gpu_maps.sh -- stage-placement experiments (SDR_MAP_SSB): headline bench value per candidate map
python -c "import __graft_entry__ as g; g.build()" || exit 1
for m in default "$@"; do
  if [ "$m" = default ]; then unset SDR_MAP_SSB; else export SDR_MAP_SSB=$m; fi
  v=$(python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.0f %s' % (d['value'], d['parity']['bit_exact']))")
  echo "MAP $m -> $v"
done
