"""Per-parameter gradient / prediction error of the CUDA path against the float64 oracle, with the
fp32 oracle's own error next to it.  python tools/diag_precision.py N W D K B [stressed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_b200.models.GDN import GDN
from oracle import gdn_oracle as go

N, W, D, K, B = [int(v) for v in sys.argv[1:6]]
stressed = len(sys.argv) <= 6 or sys.argv[6] != "0"
sd = go.init_state(N, D, W, seed=5, stressed=stressed)
g = torch.Generator().manual_seed(17)
x, y = torch.rand(B, N, W, generator=g), torch.rand(B, N, generator=g)
mask = go.dropout_mask(B, N, D, seed=3)
model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K)
model.load_state_dict(sd)
model = model.cuda().train()
model.set_dropout_mask(mask.cuda())
pred = model(x.cuda(), None)
loss = torch.nn.functional.mse_loss(pred, y.cuda())
loss.backward()
l64, p64, g64, aux = go.loss_and_grads(go.cast_state(sd, torch.float64), x.double(), y.double(), K, drop_mask=mask.double())
l32, p32, g32, _ = go.loss_and_grads({k: v.clone() for k, v in sd.items()}, x, y, K, drop_mask=mask)
same = torch.equal(model.learned_graph.cpu(), aux["learned_graph"])
print(f"shape N={N} W={W} D={D} K={K} B={B} stressed={stressed}  graph identical: {same}")
def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30), (a - b).norm().item() / max(b.norm().item(), 1e-30)
print(f"{'tensor':34s} {'ours max':>10s} {'ours L2':>10s} {'ref32 max':>10s} {'ref32 L2':>10s}")
print(f"{'pred':34s} %10.2e %10.2e %10.2e %10.2e" % (*rel(pred.detach().cpu(), p64), *rel(p32, p64)))
print(f"{'loss':34s} %10.2e %10s %10.2e" % (abs(loss.item() - l64.item()) / abs(l64.item()), '', abs(l32.item() - l64.item()) / abs(l64.item())))
for k, p in model.named_parameters():
    print(f"{k:34s} %10.2e %10.2e %10.2e %10.2e" % (*rel(p.grad.cpu(), g64[k]), *rel(g32[k], g64[k])))
