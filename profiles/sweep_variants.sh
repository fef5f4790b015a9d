#!/usr/bin/env bash
# tuning helper: time the collide schedule variants (MFLBM_VARIANT, kernels_step.cuh) on the benchmark workload
for prec in f64 f32; do
  for v in 0 1 4 5 6 7 9; do
    MFLBM_VARIANT=$v python bench.py --prec $prec --steps 60 --warmup 6 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$prec variant $v', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'MLUPS frac', round(d['roofline']['frac'],3))"
  done
done
