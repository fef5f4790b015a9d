#!/usr/bin/env bash
# GPU-box capture: tests, bench (both precisions, both arms), ncu launch list and full capture of the collide kernels.
# usage (under gpurun): bash profiles/capture.sh <tag> [notest] [nochain]
TAG=${1:-r01x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt
if [ "$2" != "notest" ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
  tail -5 $O/${TAG}_pytest.log
fi
for p in f64 f32; do
  timeout 600 python bench.py --prec $p > $O/${TAG}_bench_$p.json 2> $O/${TAG}_bench_$p.err
  timeout 600 python bench.py --impl reference --prec $p --steps 400 --warmup 20 > $O/${TAG}_bench_reference_$p.json 2>> $O/${TAG}_bench_$p.err
done
cat $O/${TAG}_bench_f64.json $O/${TAG}_bench_f32.json
for p in f64 f32; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_$p.csv \
      python bench.py --prec $p --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_launch_$p.log 2>&1
done
# full captures are summarised on the box (ncu is there too) and only the small text files travel back: gpurun merges at
# most 64 MiB of gpurun_out/, one .ncu-rep of these kernels is ~15 MB per launch
for p in f64 f32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 4 -c 2 -o $O/${TAG}_collide_$p -f \
      python bench.py --prec $p --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_full_$p.log 2>&1
  python profiles/summarize.py full $O/${TAG}_collide_$p.ncu-rep > $O/${TAG}_collide_${p}_full.txt 2>&1
  python profiles/stalls.py $O/${TAG}_collide_$p.ncu-rep 24 2>&1 | cut -c1-220 > $O/${TAG}_collide_${p}_stalls.txt
  ncu -i $O/${TAG}_collide_$p.ncu-rep --page details --print-units base 2>/dev/null | grep -E "k_collide|Throughput|Busy|Hit Rate|Executed Ipc|No Eligible|Eligible Warps|Active Warps|Registers Per|Dynamic Shared|Duration|Theoretical Occ|Achieved Occ" > $O/${TAG}_collide_${p}_details.txt
  rm -f $O/${TAG}_collide_$p.ncu-rep
done
if [ "$3" != "nochain" ]; then
  timeout 600 ncu --set full --clock-control none -k regex:'k_normals|k_extrap|k_alter|k_inlet|k_outlet' -s 12 -c 6 -o $O/${TAG}_chain_f64 -f \
      python bench.py --prec f64 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_chain_f64.log 2>&1
  python profiles/summarize.py full $O/${TAG}_chain_f64.ncu-rep > $O/${TAG}_chain_f64_full.txt 2>&1
  rm -f $O/${TAG}_chain_f64.ncu-rep
fi
