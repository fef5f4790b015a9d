#!/usr/bin/env bash
# GPU-box capture of the final state of round 2, third session (one B200): full GPU suite, bench lines at the driver's protocol
# (reference arm with WITH_REFERENCE=1), ncu launch list (LAUNCH_PRECS="f64 f32" for both precisions).
# usage (under gpurun): bash profiles/capture_r03.sh <tag>
TAG=${1:-r03z}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 420 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_k20_f64.json 2> $O/bench_k20_f64.err
if [ -n "$WITH_REFERENCE" ]; then timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_k20_f64.json 2>> $O/bench_k20_f64.err; fi
for p in ${LAUNCH_PRECS:-f64}; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 160 --csv --log-file $O/launches_$p.csv \
      python bench.py --prec $p --steps 20 --warmup 100 --no-cpu-baseline --no-e2e --no-fp32 > $O/ncu_launch_$p.log 2>&1
  python profiles/summarize.py launches $O/launches_$p.csv > $O/launches_$p.txt
done
head -c 1500 $O/bench_k20_f64.json; echo; cat $O/launches_f64.txt
