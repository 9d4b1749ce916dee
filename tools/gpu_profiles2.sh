#!/bin/bash
# ncu captures exported to CSV on the box (the .ncu-rep files exceed the 64 MiB copy-back limit)
mkdir -p gpurun_out
for w in C5 C4; do
  timeout 300 python tools/prof_gl.py $w 3 > gpurun_out/gl_plain_$w.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none -k regex:k_ --launch-skip 29 -c 13 -o /tmp/prof_gl_$w python tools/prof_gl.py $w 3 > gpurun_out/ncu_gl_$w.log 2>&1
  echo "ncu gl $w rc=$?"
  ncu -i /tmp/prof_gl_$w.ncu-rep --page raw --csv > gpurun_out/prof_gl_${w}_raw.csv 2>/dev/null
done
timeout 600 python tools/prof_step.py C5 3 > gpurun_out/prof_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_C5.csv python tools/prof_step.py C5 3 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 2400 ncu --set full --clock-control none -k 'regex:^k_(attn_fwd|attn_bwd|bwd1|bwd2|bwd3|fwd_out|fwd_stats2|gram_tc|rescore|moments|transpose_scalars)' --launch-skip 22 -c 11 -o /tmp/prof_C5 python tools/prof_step.py C5 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/prof_C5.ncu-rep --page raw --csv > gpurun_out/prof_C5_raw.csv 2>/dev/null
ls -la gpurun_out
