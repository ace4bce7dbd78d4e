#!/bin/bash
# tools/ab_variants.sh -- headline bench with features switched off (diagnostics: what each group of stages costs in the product kernel)
for v in "" nonb noagc noaud nonb,noagc nonb,noagc,noaud; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 --variant "$v" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VARIANT [%s] -> %.0f Msps  %.3f ms' % ('$v', d['value'], d['per_launch_ms']['mean']))"
done
