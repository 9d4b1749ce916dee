"""Long-accumulation check of the tensor-core lin backward: gradients at C4/C5 under GDN_NO_MMA=0 vs the FMA
kernels (GDN_NO_MMA=7), compared in a second invocation (mode 'cmp')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if sys.argv[1] == "cmp":
    for w in sys.argv[2:]:
        m = os.environ.get("CMP_MASK", "0")
        a = torch.load(f"gpurun_out/mma2_{w}_{m}.pt"); b = torch.load(f"gpurun_out/mma2_{w}_7.pt")
        print(w, "mask", m, "vs 7")
        for k in a:
            d = (a[k].double() - b[k].double()).abs()
            print("  %-34s max|diff|/max|ref| %.3e   L2 rel %.3e" % (k, (d.max() / b[k].double().abs().max()).item(),
                                                                     (d.norm() / b[k].double().norm()).item()))
    sys.exit()
from gdn_b200.models.GDN import GDN
w = sys.argv[1]
N, W, D, K, B = {"C4": (4096, 16, 128, 32, 64), "C5": (16384, 16, 128, 64, 64)}[w]
torch.manual_seed(5)
model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).cuda().train()
model.dp.p = 0.0
x = torch.rand(B, N, W, device="cuda"); y = torch.rand(B, N, device="cuda")
loss = torch.nn.functional.mse_loss(model(x, None), y)
loss.backward()
out = {"pred_loss": loss.detach().cpu().reshape(1)}
out.update({k: p.grad.cpu() for k, p in model.named_parameters()})
torch.save(out, f"gpurun_out/mma2_{w}_{os.environ.get('GDN_NO_MMA', '0')}.pt")
print(w, "NO_MMA", os.environ.get("GDN_NO_MMA", "0"), "loss", loss.item())
