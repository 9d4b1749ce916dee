"""Scorer timing: python tools/score_time.py [T] [N]  -> one JSON line (CUDA events, per-kernel breakdown)."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import profile_collect
from gdn_b200 import _lib, ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
dev = torch.device("cuda", 0)
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(1)
gt = torch.rand(T, N, device=dev, generator=g)
pred = gt + 0.05 * torch.randn(T, N, device=dev, generator=g)
ms = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.score(pred, gt)
    b.record()
    torch.cuda.synchronize()
    if i >= 3:
        ms.append(a.elapsed_time(b))
lib.gdn_profile_enable(1)
for _ in range(3):
    ops.score(pred, gt)
torch.cuda.synchronize()
_, rows = profile_collect(lib)
lib.gdn_profile_enable(0)
t = statistics.mean(ms)
alg = 16 * T * N
print(json.dumps({"T": T, "N": N, "ms": round(t, 4), "algorithmic_GBps": round(alg / t / 1e6, 1), "frac_of_6552": round(alg / t / 1e6 / 6552.3, 4),
                  "kernels_ms": {k: round(tt / c, 4) for k, (c, tt) in rows.items()}}))
