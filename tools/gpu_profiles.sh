#!/bin/bash
# ncu evidence for profiles/: launch list of the train step + GraphLayer leg (C5), full sets of the top kernels
mkdir -p gpurun_out
timeout 600 python tools/prof_step.py C5 3 > gpurun_out/prof_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_C5.csv python tools/prof_step.py C5 3 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 2400 ncu --set full --clock-control none --import-source on -k 'regex:^k_(attn_fwd|attn_bwd|bwd1|bwd2|bwd3|fwd_out|fwd_stats2|gram_tc|rescore|lin_fwd|lin_bwd|moments|transpose_scalars)' --launch-skip 22 -c 15 -o gpurun_out/prof_C5 python tools/prof_step.py C5 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -n 3 gpurun_out/ncu_full.log
