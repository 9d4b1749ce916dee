#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/pytest_gpu.log
python - > gpurun_out/graph_time.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from gdn_b200 import ops
for N, D, K in ((127,128,30),(4096,128,32),(16384,128,64)):
    V = (torch.rand(N, D, device='cuda')*2-1)/D**0.5
    for _ in range(2): ops.graph_build(V, K, use_tensor_cores=0)
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): ops.graph_build(V, K, use_tensor_cores=0)
    b.record(); torch.cuda.synchronize()
    print(f"graph_build N={N} D={D} K={K}: {a.elapsed_time(b)/3:.3f} ms")
PY
cat gpurun_out/graph_time.log
for w in C1 C2 C3 C4 C5; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.log 2>&1; echo "bench $w rc=$?"
done
timeout 600 python tools/prof_step.py C5 3 > gpurun_out/prof_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C5.csv python tools/prof_step.py C5 3 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 2400 ncu --set full --clock-control none --import-source on -k 'regex:^k_(attn_fwd|attn_bwd|bwd1|bwd2|bwd3|fwd_out|fwd_stats2|gram_topk|lin_fwd|lin_bwd|moments|transpose_scalars)' --launch-skip 20 -c 14 -o gpurun_out/prof_C5 python tools/prof_step.py C5 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -n 5 gpurun_out/ncu_full.log
ls -la gpurun_out
