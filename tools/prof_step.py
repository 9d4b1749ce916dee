"""Small driver for ncu: fused train steps on one BASELINE workload; the last one sits inside a
cudaProfilerStart/Stop range.  python tools/prof_step.py C5 3"""
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
warm = 8                                   # warm-started graph builds, settled allocator
for i in range(warm + steps):
    if i == warm + steps - 1:              # ncu --profile-from-start off: only the last train step is captured
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    loss = trainer.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(loss))
