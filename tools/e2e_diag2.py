"""e2e feed variants at C5 (pageable float64 host batches): threaded x deferred-loss matrix + where the time goes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from gdn_b200.data import LossReader, Prefetcher
from gdn_b200.dp import WindowShardedTrainer
from gdn_b200.models.GDN import GDN

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C5"]
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
N, W, D, K, B = wl["N"], wl["W"], wl["D"], wl["K"], wl["B"]
dev = torch.device("cuda", 0)
torch.manual_seed(5)
model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).to(dev).train()
trainer = WindowShardedTrainer(model, lr=1e-3)
hx = [torch.rand(B, N, W, dtype=torch.float64) for _ in range(4)]
hy = [torch.rand(B, N, dtype=torch.float64) for _ in range(4)]
print("torch threads", torch.get_num_threads(), "cores", os.cpu_count())


def batches(n):
    for i in range(n):
        yield hx[i % 4], hy[i % 4]


def run(n, threaded, deferred, stage_threads=None, depth=2):
    reader = LossReader(dev) if deferred else None
    out = []
    for bx, by in Prefetcher(batches(n), dev, skip=(), reuse_buffers=True, threaded=threaded, stage_threads=stage_threads,
                             depth=depth):
        loss = trainer.step(bx, by)
        if deferred:
            v = reader.push(loss)
        else:
            out.append(loss.item())
    if deferred:
        reader.flush()
    torch.cuda.synchronize()


for threaded in (False, True):
    for deferred in (False, True):
        run(4, threaded, deferred)
        t = time.perf_counter()
        run(STEPS, threaded, deferred)
        print(f"threaded={threaded} deferred={deferred}: {(time.perf_counter() - t) / STEPS * 1e3:.3f} ms/step")
for st in (1, 2, 4, 8):
    for depth in (2, 3):
        run(4, True, True, st, depth)
        t = time.perf_counter()
        run(STEPS, True, True, st, depth)
        print(f"threaded deferred stage_threads={st} depth={depth}: {(time.perf_counter() - t) / STEPS * 1e3:.3f} ms/step")
# resident batches for reference
xd, yd = hx[0].float().to(dev), hy[0].float().to(dev)
for _ in range(3):
    trainer.step(xd, yd)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(STEPS):
    trainer.step(xd, yd)
torch.cuda.synchronize()
print(f"resident: {(time.perf_counter() - t) / STEPS * 1e3:.3f} ms/step")
