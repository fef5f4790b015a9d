#!/usr/bin/env bash
# GPU-box capture of round 2 (one B200): bench lines at the driver's protocol (both arms), the BASELINE protocol, ncu launch
# lists, ncu --set full of the collide and chain kernels, reference block-shape sweep.
# usage (under gpurun): bash profiles/capture_r02.sh <tag>
TAG=${1:-r02x}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_k20_f64.json 2> $O/bench_k20_f64.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_k20_f64.json 2>> $O/bench_k20_f64.err
timeout 300 python bench.py --prec f32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_k20_f32.json 2> $O/bench_k20_f32.err
timeout 300 python bench.py --impl reference --prec f32 --steps 20 --warmup 5 > $O/bench_reference_k20_f32.json 2>> $O/bench_k20_f32.err
timeout 600 python bench.py --no-cpu-baseline > $O/bench_f64.json 2> $O/bench_f64.err
for b in 256,1,1 64,1,1 64,2,1 32,4,1 128,2,1 32,2,2; do
  timeout 300 python bench.py --impl reference --steps 100 --warmup 10 --no-e2e --ref-block $b > $O/bench_reference_block_${b//,/x}.json 2>> $O/refblock.err
done
for p in f64 f32; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 160 --csv --log-file $O/launches_$p.csv \
      python bench.py --prec $p --steps 20 --warmup 100 --no-cpu-baseline --no-e2e --no-fp32 > $O/ncu_launch_$p.log 2>&1
  python profiles/summarize.py launches $O/launches_$p.csv > $O/launches_$p.txt
done
for p in f64 f32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 100 -c 2 -o $O/collide_$p -f \
      python bench.py --prec $p --steps 20 --warmup 100 --no-cpu-baseline --no-e2e --no-fp32 > $O/ncu_full_$p.log 2>&1
  python profiles/summarize.py full $O/collide_$p.ncu-rep > $O/collide_${p}_full.txt 2>&1
  python profiles/stalls.py $O/collide_$p.ncu-rep 24 2>&1 | cut -c1-220 > $O/collide_${p}_stalls.txt
  ncu -i $O/collide_$p.ncu-rep --page details --print-units base 2>/dev/null | grep -E "k_collide|Throughput|Busy|Hit Rate|Executed Ipc|No Eligible|Eligible Warps|Active Warps|Registers Per|Dynamic Shared|Duration|Theoretical Occ|Achieved Occ" > $O/collide_${p}_details.txt
  rm -f $O/collide_$p.ncu-rep
done
cat $O/bench_k20_f64.json | head -c 600
