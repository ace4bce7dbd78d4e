#!/bin/bash
# tools/gpu_r02_i.sh -- defaults after the SAM plan changes, contracting build, ramped host chunks, spectrum debug, GPU tests.
set -u
mkdir -p gpurun_out
TAG=${1:-r02i}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env SDR_DEBUG_PLAN=1 "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 3 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    plan=[l.strip() for l in open('gpurun_out/${TAG}_$name.err') if l.startswith('[sdr] launch')][:1]
    print('$name: %.0f Msps  ms/step %.3f  parity %s e2e %.0f (link %.0f) | %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact'), d['e2e']['value'], d['e2e']['link_bound']['value'], ' ; '.join(p[14:120] for p in plan)))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
BARGS="--workload 3"; run w3_default X=1
BARGS=""; run w2 X=1; run w2_noramp SDR_HOST_RAMP=0; run w2_ramp4 SDR_HOST_RAMP=4; run w2_ramp16 SDR_HOST_RAMP=16
BARGS="--contract"; run w2_contract X=1
BARGS="--contract --workload 5"; run w5_contract X=1
BARGS="--workload 5"; run w5 X=1
echo "== spectrum debug"; timeout 120 python tools/spec_debug.py 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
echo "== full default bench line"; ( time timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | grep real; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('bench: value %.0f e2e %.0f link_frac %.3f contracting %s' % (d['value'], d['e2e']['value'], d['e2e'].get('link_frac', 0), d.get('contracting_build')))
    for k,w in (d.get('workloads') or {}).items():
        print(' ', k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in w.items() if kk in ('value','ms_per_step','error')}, 'parity', (w.get('parity') or {}).get('bit_exact'), 'fp32 frac', (w.get('roofline_fp32') or {}).get('frac'))
except Exception as e:
    print('bench FAILED', e)
PY
