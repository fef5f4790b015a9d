#!/usr/bin/env bash
# GPU-box script for the opt-in paths written at the end of round 1 (unmeasured there: the GPU budget was spent):
#   MFLBM_ACTIVITY=1  gradient chain with the interface-activity map (csrc/kernels_activity.cuh)
#   MFLBM_LANES=1     outlet kernel on a second lane next to the inlet kernel
# usage (under gpurun): bash profiles/measure_optins.sh <tag> [notest]
# 1. the whole GPU suite through both options (parity, slabs, host driver, reference comparisons), activity tests included
# 2. bench lines: plain / activity / activity + lanes, both precisions, 500 timed steps
# 3. ncu launch list of the activity chain (per-kernel times of k_act_scan, k_act_dilate, k_normals_act, ...)
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
if [ "$2" != "notest" ]; then
  ( time MFLBM_ACTIVITY=1 MFLBM_LANES=1 MFLBM_TEST_ACTIVITY=1 timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest_optins.log 2>&1
  tail -5 $O/${TAG}_pytest_optins.log
fi
for p in f64 f32; do
  for cfg in "0 0" "1 0" "1 1"; do
    set -- $cfg
    MFLBM_LANES=$2 timeout 300 python bench.py --prec $p --activity $1 --steps 500 --warmup 50 --no-cpu-baseline --no-e2e \
        > $O/${TAG}_bench_${p}_act$1_lanes$2.json 2> $O/${TAG}_bench_${p}_act$1_lanes$2.err
    python - "$O/${TAG}_bench_${p}_act$1_lanes$2.json" "$p act=$1 lanes=$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]
print(sys.argv[2], round(d["value"]), "MLUPS", round(d["ms_per_step"], 4), "ms/step; step frac", round(r["whole_step"]["frac"], 3),
      "collide share", round(r["whole_step"]["collide_share_of_step"], 3))
PY
  done
  MFLBM_ACTIVITY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_${p}_act.csv \
      python bench.py --prec $p --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_launch_${p}_act.log 2>&1
  python profiles/summarize.py launches $O/${TAG}_launches_${p}_act.csv > $O/${TAG}_launches_${p}_act.txt 2>&1
  head -16 $O/${TAG}_launches_${p}_act.txt | cut -c1-140
done
# multi-GPU (separate gpurun --gpus N calls; N = 2, 4, 8), equal vs cost-balanced cuts, plain vs activity map:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus N \
#       --steps 500 --warmup 50 --no-e2e --no-cpu-baseline [--partition balanced] [--activity 1]
