"""GraphLayer forward/backward at the module boundary only (for ncu).  python tools/prof_gl.py C5 3"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from gdn_b200 import ops
from gdn_b200.models.graph_layer import GraphLayer

name = sys.argv[1] if len(sys.argv) > 1 else "C5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = WORKLOADS[name]
N, W, D, K, B = wl["N"], wl["W"], wl["D"], wl["K"], wl["B"]
torch.manual_seed(5)
dev = torch.device("cuda", 0)
layer = GraphLayer(W, D, heads=1, concat=False).to(dev)
V = ((torch.rand(N, D, device=dev) * 2 - 1) / D ** 0.5).requires_grad_(True)
x = torch.rand(B, N, W, device=dev)
gout = torch.rand(B * N, D, device=dev)
_, nbr = ops.graph_build(V.detach(), K)
torch.cuda.synchronize()
for i in range(iters):
    if i == iters - 1:                      # ncu --profile-from-start off: only the last fwd+bwd is captured
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    out = layer.forward_batched(x, nbr, V)
    out.backward(gout)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
