#!/bin/bash
# Round-2 measurement artefacts (one gpurun call, 1 GPU): bench lines, ncu launch list of the bench command,
# ncu --set full of the GraphLayer fwd+bwd (C5, C4), of one warm train step (C5) and of the scorer.  CSV exported on the box.
mkdir -p gpurun_out
O=gpurun_out
R=r02
timeout 900 python bench.py > $O/bench_default.log 2>&1; echo "bench default rc=$?"; tail -n 1 $O/bench_default.log > $O/${R}_bench_C5.json
for w in C4 C3 C2 C1; do timeout 600 python bench.py --workload $w > $O/bench_$w.log 2>&1; echo "bench $w rc=$?"; tail -n 1 $O/bench_$w.log > $O/${R}_bench_$w.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; echo "bench reference rc=$?"; tail -n 1 $O/bench_ref.log > $O/${R}_bench_reference.json
# launch list of the bench command itself (short run; exits 0 without ncu first)
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 $CMD > $O/bench_short.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${R}_launches_C5.csv $CMD > $O/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
for w in C5 C4; do
  timeout 300 python tools/prof_gl.py $w 3 > $O/gl_plain_$w.log 2>&1 && \
  timeout 1500 ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/prof_gl_$w python tools/prof_gl.py $w 3 > $O/ncu_gl_$w.log 2>&1
  echo "ncu gl $w rc=$?"
  ncu -i /tmp/prof_gl_$w.ncu-rep --page raw --csv > $O/${R}_ncu_${w}_graphlayer_raw.csv 2>/dev/null
done
timeout 600 python tools/prof_step.py C5 3 > $O/prof_plain.log 2>&1 && \
timeout 2400 ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/prof_C5 python tools/prof_step.py C5 3 > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/prof_C5.ncu-rep --page raw --csv > $O/${R}_ncu_C5_trainstep_raw.csv 2>/dev/null
timeout 300 python tools/score_time.py 4096 16384 > $O/${R}_score_time.json 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_score_sensor|k_delta_transpose" -c 2 -o /tmp/prof_sc python tools/score_time.py 4096 16384 > $O/ncu_sc.log 2>&1
ncu -i /tmp/prof_sc.ncu-rep --page raw --csv > $O/${R}_ncu_score_raw.csv 2>/dev/null
timeout 200 python tools/tc_check.py > $O/${R}_tc_engine_check.txt 2>&1
ls -la $O | tail -30
