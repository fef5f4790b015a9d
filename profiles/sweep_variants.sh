#!/usr/bin/env bash
# tuning helper: time the collide kernel configurations (MFLBM_VARIANT = 100*even + odd, mflbm.cu launch_collide_default;
# on the benchmark workload (last swept in round 2: profiles/README.md, the defaults are the fastest).
#   VARIANTS="0 1 2 3 4 100" PRECS="f64" bash profiles/sweep_variants.sh
for prec in ${PRECS:-f64 f32}; do
  for v in ${VARIANTS:-0 1 2 3 4 100}; do
    MFLBM_VARIANT=$v python bench.py --prec $prec --steps 100 --warmup 6 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$prec variant $v', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'MLUPS  step frac', round(r['whole_step']['frac'],3), ' collide odd/even ms', round(r['odd']['ms'],4), round(r['even']['ms'],4), 'frac', round(r['odd']['frac'],3), round(r['even']['frac'],3))"
  done
done
