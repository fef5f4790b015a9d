#!/usr/bin/env bash
# tuning helper: time the collide kernel configurations (MFLBM_VARIANT = 100*even + odd, mflbm.cu launch_collide_default)
# on the benchmark workload.   VARIANTS="0 1 2" PRECS="f64" bash profiles/sweep_variants.sh
for prec in ${PRECS:-f64 f32}; do
  for v in ${VARIANTS:-0 1 2 3 100 200}; do
    MFLBM_VARIANT=$v python bench.py --prec $prec --steps 100 --warmup 6 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$prec variant $v', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'MLUPS frac', round(d['roofline']['frac'],3))"
  done
done
