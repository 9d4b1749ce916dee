#!/bin/bash
# Round-1 measurement artefacts (one gpurun call, 1 GPU): bench lines, ncu launch list of the bench command,
# ncu --set full of the GraphLayer fwd+bwd (C5, C4) and of one warm train step (C5).  CSV exported on the box.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py > $O/bench_default.log 2>&1; echo "bench default rc=$?"; tail -n 1 $O/bench_default.log > $O/r01_bench_C5.json
timeout 600 python bench.py --workload C4 > $O/bench_C4_full.log 2>&1; echo "bench C4 rc=$?"; tail -n 1 $O/bench_C4_full.log > $O/r01_bench_C4.json
for w in C1 C2 C3; do timeout 300 python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.log 2>&1; tail -n 1 $O/bench_$w.log > $O/r01_bench_$w.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; echo "bench reference rc=$?"; tail -n 1 $O/bench_ref.log > $O/r01_bench_reference.json
# launch list of the bench command itself (short run; exits 0 without ncu first)
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 $CMD > $O/bench_short.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r01_launches_C5.csv $CMD > $O/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
for w in C5 C4; do
  timeout 300 python tools/prof_gl.py $w 3 > $O/gl_plain_$w.log 2>&1 && \
  timeout 1500 ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/prof_gl_$w python tools/prof_gl.py $w 3 > $O/ncu_gl_$w.log 2>&1
  echo "ncu gl $w rc=$?"
  ncu -i /tmp/prof_gl_$w.ncu-rep --page raw --csv > $O/r01_ncu_${w}_graphlayer_raw.csv 2>/dev/null
done
timeout 600 python tools/prof_step.py C5 3 > $O/prof_plain.log 2>&1 && \
timeout 2400 ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/prof_C5 python tools/prof_step.py C5 3 > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/prof_C5.ncu-rep --page raw --csv > $O/r01_ncu_C5_trainstep_raw.csv 2>/dev/null
timeout 200 python tools/tc_check.py > $O/r01_tc_engine_check.txt 2>&1
timeout 100 python tools/h2d_bw.py > $O/r01_h2d_bw.txt 2>&1
ls -la $O | tail -30
