"""Eager train step with torch's fused Adam vs gdn_b200.optim.FlatAdam (row f-4).  python tools/flat_adam_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS
from gdn_b200.dp import WindowShardedTrainer
from gdn_b200.models.GDN import GDN
for name in ("C1", "C3", "C5"):
    wl = WORKLOADS[name]
    N, W, D, K, B = wl["N"], wl["W"], wl["D"], wl["K"], wl["B"]
    x, y = torch.rand(B, N, W, device="cuda"), torch.rand(B, N, device="cuda")
    res = {}
    for flat in (False, True):
        torch.manual_seed(5)
        m = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).cuda().train()
        tr = WindowShardedTrainer(m, lr=1e-3, flat_adam=flat)
        for _ in range(10):
            tr.step(x, y)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            tr.step(x, y)
        b.record()
        torch.cuda.synchronize()
        res[flat] = a.elapsed_time(b) / 50
    print(f"{name}: eager step {res[False]:.3f} ms (torch fused Adam)  {res[True]:.3f} ms (FlatAdam)")
