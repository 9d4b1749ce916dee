#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 12 gpurun_out/pytest_gpu.log
( python tools/diag_precision.py 4096 16 128 32 2; python tools/diag_precision.py 4096 16 128 32 2 0; python tools/diag_precision.py 127 5 128 30 256; python tools/diag_precision.py 51 5 64 15 128 ) > gpurun_out/diag.log 2>&1
cat gpurun_out/diag.log
for w in C2 C4 C5; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "bench $w rc=$?"
  tail -n 1 gpurun_out/bench_$w.log | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
print('  kernels', {a:round(b,3) for a,b in list(d['kernels_ms_per_step'].items())[:14]})
r=d['roofline']; print('  roofline frac', round(r['frac'],4), 'fwd_ms', round(r['fwd_ms'],3), 'bwd_ms', round(r['bwd_ms'],3))
"
done
