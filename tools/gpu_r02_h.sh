#!/bin/bash
# tools/gpu_r02_h.sh -- SAM plan variants (tile 16 merged x3, input depth), placement search of the merged plan, spectrum tap, GPU tests.
set -u
mkdir -p gpurun_out
TAG=${1:-r02h}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env SDR_DEBUG_PLAN=1 "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    plan=[l.strip() for l in open('gpurun_out/${TAG}_$name.err') if l.startswith('[sdr] launch')][:1]
    print('$name: %.0f Msps  ms/step %.3f  parity %s | %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact'), ' ; '.join(p[14:150] for p in plan)))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
BARGS="--workload 3"
run w3_default X=1
run w3_T16x3m_d1 SDR_TILE_ENV=16 SDR_CTAS_PER_SM=3 SDR_IN_DEPTH=1
run w3_T8x3m_d2 SDR_TILE_ENV=8 SDR_CTAS_PER_SM=3 SDR_IN_DEPTH=2
run w3_T8x3m_d1 SDR_TILE_ENV=8 SDR_CTAS_PER_SM=3 SDR_IN_DEPTH=1
echo "== spectrum tap"; timeout 300 python bench_aux.py --path grabber_spectrum > gpurun_out/${TAG}_aux_spectrum.json 2> gpurun_out/${TAG}_aux_spectrum.err; cut -c1-700 gpurun_out/${TAG}_aux_spectrum.json; tail -2 gpurun_out/${TAG}_aux_spectrum.err
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
echo "== placement search, merged SAM plan"
timeout 400 python tools/map_search.py --cls envmerged --seconds 150 --blocks 32 > gpurun_out/${TAG}_map_envmerged.log 2>&1; tail -8 gpurun_out/${TAG}_map_envmerged.log
