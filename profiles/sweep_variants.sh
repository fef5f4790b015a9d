#!/usr/bin/env bash
# tuning helper: time the collide kernel variants (MFLBM_VARIANT, mflbm.cu phase_collide) on the benchmark workload
#   0 = pipelined kernels, default stage counts; 201/202 = other stage counts; 100 = plain one-thread-per-node kernel
for prec in ${PRECS:-f64 f32}; do
  for v in ${VARIANTS:-0 201 202 100}; do
    MFLBM_VARIANT=$v python bench.py --prec $prec --steps 100 --warmup 6 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$prec variant $v', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'MLUPS frac', round(d['roofline']['frac'],3))"
  done
done
