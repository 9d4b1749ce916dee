#!/bin/bash
# ncu capture of the tensor-core graph engine kernels during warm-started builds (CSV exported on the box)
mkdir -p gpurun_out
timeout 300 python tools/tc_warm.py > gpurun_out/tc_warm_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:k_(rescore|gram_tc)' --launch-skip 8 -c 4 -o /tmp/prof_tc python tools/tc_warm.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_raw.csv 2>/dev/null
ncu -i /tmp/prof_tc.ncu-rep --page source --csv --kernel-name regex:k_rescore > gpurun_out/prof_tc_rescore_src.csv 2>/dev/null
ls -la gpurun_out | tail -5
