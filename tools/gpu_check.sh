#!/bin/bash
# full GPU regression: tests, TC engine timing, C4/C5 bench with the per-kernel breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
timeout 200 python tools/tc_warm.py | grep "N=\|gram_tc\|rescore"
for w in ${WORKLOADS:-C4 C5}; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_$w.log 2>&1
  echo "bench $w rc=$?"
  tail -n 1 gpurun_out/bench_$w.log | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
print('  kernels', {a:round(b,3) for a,b in list(d['kernels_ms_per_step'].items())[:14]})
"
done
echo "pytest -m gpu: $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head
