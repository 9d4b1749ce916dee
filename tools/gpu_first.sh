#!/bin/bash
# first GPU round: smoke, parity tests (all, no -x), small + default bench, memcheck of smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/bench_C2.log 2>&1; echo "bench C2 rc=$?"
timeout 600 python bench.py --workload C4 --steps 10 --warmup 3 > gpurun_out/bench_C4.log 2>&1; echo "bench C4 rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_C5.log 2>&1; echo "bench C5 rc=$?"
tail -3 gpurun_out/bench_C2.log gpurun_out/bench_C4.log gpurun_out/bench_C5.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/memcheck.log
