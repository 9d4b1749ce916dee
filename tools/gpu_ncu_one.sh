#!/bin/bash
# ncu --set full capture of ONE kernel (regex $1) of the GraphLayer fwd+bwd at workload $2 (default C5);
# raw + source pages exported to CSV on the box (the .ncu-rep stays in /tmp)
K=${1:-k_lin_bwd}; W=${2:-C5}; SCRIPT=${3:-tools/prof_gl.py}
mkdir -p gpurun_out
timeout 300 python $SCRIPT $W 3 > gpurun_out/one_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/one_plain.log; exit 1; }
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:$K" --launch-skip 2 -c 1 -o /tmp/prof_one python $SCRIPT $W 3 > gpurun_out/ncu_one.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_one.ncu-rep --page raw --csv > gpurun_out/prof_one_raw.csv 2>/dev/null
ncu -i /tmp/prof_one.ncu-rep --page source --csv > gpurun_out/prof_one_src.csv 2>/dev/null
ls -la gpurun_out | tail -4
