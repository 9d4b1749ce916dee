"""Small driver for ncu: a few train steps (fused path) plus GraphLayer fwd+bwd at the module
boundary on one BASELINE workload.  python tools/prof_step.py C5 3"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from gdn_b200 import ops
from gdn_b200.dp import WindowShardedTrainer
from gdn_b200.models.GDN import GDN

name = sys.argv[1] if len(sys.argv) > 1 else "C5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = WORKLOADS[name]
N, W, D, K, B = wl["N"], wl["W"], wl["D"], wl["K"], wl["B"]
torch.manual_seed(5)
dev = torch.device("cuda", 0)
model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).to(dev).train()
trainer = WindowShardedTrainer(model)
x, y = torch.rand(B, N, W, device=dev), torch.rand(B, N, device=dev)
for i in range(steps):
    loss = trainer.step(x, y)
layer = model.gnn_layers[0].gnn
_, nbr = ops.graph_build(model.embedding.weight, K)
Vp = model.embedding.weight.detach().clone().requires_grad_(True)
gout = torch.rand(B * N, D, device=dev)
for i in range(steps):
    out = layer.forward_batched(x, nbr, Vp)
    out.backward(gout)
torch.cuda.synchronize()
print("done", float(loss))
