#!/bin/bash
# tools/gpu_r02_e.sh -- product defaults on the box: quick A/B against the round-1 library, the full default bench line, GPU tests.
set -u
mkdir -p gpurun_out
TAG=${1:-r02e}
run() { # name, env..., (BARGS)
  local name=$1; shift
  env SDR_DEBUG_PLAN=1 "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --only-headline ${BARGS:-} > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1])
    plan=[l.strip() for l in open('gpurun_out/${TAG}_$name.err') if l.startswith('[sdr] launch')][:2]
    print('$name: %.0f Msps  ms/step %.3f  parity %s | %s' % (d['value'], d['ms_per_step'], (d['parity'] or {}).get('bit_exact'), ' ; '.join(p[14:95] for p in plan)))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/${TAG}_$name.err').read()[-600:])
PY
}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
BARGS=""; run w2_tree X=1
BARGS="--workload 5"; run w5_tree X=1
BARGS="--workload 3"; run w3_tree X=1
BARGS="--workload 4"; run w4_tree X=1
echo "== full default bench line"; ( time timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | grep real; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print('bench: value %.0f e2e %.0f parity %s status %s sustained %s' % (d['value'], d['e2e']['value'], d['parity'], d['status'], d['sustained'] and {k:d['sustained'][k] for k in ('value','seconds')}))
    for k,w in (d.get('workloads') or {}).items():
        print(' ', k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in w.items() if kk in ('value','ms_per_step','parity','status','error','scaling')}, 'fp32 frac', (w.get('roofline_fp32') or {}).get('frac'))
    print('  fp32', d['roofline_fp32']['frac'], 'issue', d['roofline_fp32']['issue_frac'], 'affinity', d['host_affinity'])
except Exception as e:
    print('bench FAILED', e)
PY
echo "== role profile w3"; SDR_ROLE_PROFILE_NB=1 timeout 300 python bench.py --workload 3 --steps 5 --warmup 3 --role-profile --no-cpu-baseline --e2e-steps 1 --only-headline > gpurun_out/${TAG}_w3_roles.json 2>&1
tail -c 3000 gpurun_out/${TAG}_w3_roles.json | grep -o '"role_profile.*' | cut -c1-700; grep '^\[sdr\]' gpurun_out/${TAG}_w3_roles.json | cut -c1-400
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
