#!/usr/bin/env bash
# ncu --set full of the collide kernels (one odd + one even launch) for both precisions: bash profiles/ncu_collide.sh <tag>
TAG=${1:-x}
for p in ${PRECS:-f64 f32}; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 6 -c 2 -o gpurun_out/${TAG}_collide_$p -f \
    python bench.py --prec $p --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_$p.log 2>&1
done
