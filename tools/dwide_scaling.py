"""Per-kernel times of the fused train step at D = 128 vs D = 64 (same N, W, K, B): do the D-wide passes take less than
half the time at half the channels per lane (i.e. would splitting a row's channels over two warps pay)?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import profile_collect
from gdn_b200 import _lib
from gdn_b200.dp import WindowShardedTrainer
from gdn_b200.models.GDN import GDN

lib = _lib.load()
dev = torch.device("cuda", 0)
N, W, K, B = 16384, 16, 64, 64
for D in (128, 64):
    torch.manual_seed(5)
    model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).to(dev).train()
    tr = WindowShardedTrainer(model, cuda_graph=False)
    x, y = torch.rand(B, N, W, device=dev), torch.rand(B, N, device=dev)
    for _ in range(4):
        tr.step(x, y)
    torch.cuda.synchronize()
    lib.gdn_profile_enable(1)
    for _ in range(3):
        tr.step(x, y)
    torch.cuda.synchronize()
    _, rows = profile_collect(lib)
    lib.gdn_profile_enable(0)
    print("D =", D, {k: round(t / c, 4) for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]) if k in
                     ("k_fwd_stats2", "k_fwd_out", "k_bwd1", "k_bwd2", "k_bwd3", "k_moments")})
    del model, tr
