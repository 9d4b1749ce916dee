"""Precision diagnostic at a ragged shape: prediction / gradient errors against the float64 oracle, and how
close the nearest ReLU pre-activations sit to zero (run under GDN_NO_MMA=0 and =4 to compare engines)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import gdn_oracle as go
from gdn_b200.models.GDN import GDN
N, W, D, K, B = 64, 12, 64, 9, 96
sd = go.init_state(N, D, W, seed=21, stressed=True)
g = torch.Generator().manual_seed(5)
x, y = torch.rand(B, N, W, generator=g), torch.rand(B, N, generator=g)
mask = go.dropout_mask(B, N, D, seed=9)
model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K)
model.load_state_dict(sd)
model = model.cuda().train()
model.set_dropout_mask(mask.cuda())
pred = model(x.cuda(), None)
loss = torch.nn.functional.mse_loss(pred, y.cuda())
loss.backward()
sd64 = go.cast_state(sd, torch.float64)
l64, p64, g64, aux = go.loss_and_grads(sd64, x.double(), y.double(), K, drop_mask=mask.double())
l32, p32, g32, _ = go.loss_and_grads({k: v.clone() for k, v in sd.items()}, x, y, K, drop_mask=mask)
nw = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
print("NO_MMA", os.environ.get("GDN_NO_MMA", "0"), " pred err ours %.3e  ref32 %.3e" % (nw(pred.detach().cpu(), p64), nw(p32, p64)))
for k, p in model.named_parameters():
    print("  %-36s ours %.3e   ref32 %.3e" % (k, nw(p.grad.cpu().reshape(-1), g64[k].reshape(-1)), nw(g32[k].reshape(-1), g64[k].reshape(-1))))
torch.save({"pred": pred.detach().cpu(), **{k: p.grad.cpu() for k, p in model.named_parameters()}},
           "gpurun_out/mma_diag_%s.pt" % os.environ.get("GDN_NO_MMA", "0"))
