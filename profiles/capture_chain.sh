#!/usr/bin/env bash
# GPU-box capture of the gradient-chain kernels (one B200): ncu --set full of one launch of each, key metrics + stall table.
# usage (under gpurun): bash profiles/capture_chain.sh <tag> [f64|f32]
TAG=${1:-r03x}
P=${2:-f64}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_chain|k_act' -s 240 -c 4 -o $O/chain_$P -f \
    python bench.py --prec $P --steps 20 --warmup 100 --no-cpu-baseline --no-e2e --no-fp32 > $O/ncu_chain_$P.log 2>&1
python profiles/summarize.py full $O/chain_$P.ncu-rep > $O/chain_${P}_full.txt 2>&1
python profiles/stalls.py $O/chain_$P.ncu-rep 30 2>&1 | cut -c1-220 > $O/chain_${P}_stalls.txt
ncu -i $O/chain_$P.ncu-rep --page details --print-units base 2>/dev/null | grep -E "k_chain|k_act|Throughput|Busy|Hit Rate|Executed Ipc|No Eligible|Eligible Warps|Active Warps|Registers Per|Shared Memory|Duration|Theoretical Occ|Achieved Occ" > $O/chain_${P}_details.txt
rm -f $O/chain_$P.ncu-rep
head -c 3000 $O/chain_${P}_full.txt
