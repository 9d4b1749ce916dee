"""Per-kernel breakdown of the end-to-end (host-fed) training loop at one workload."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_b200 import _lib
from gdn_b200.models.GDN import GDN
from gdn_b200.dp import WindowShardedTrainer
from gdn_b200.data import Prefetcher
lib = _lib.load()
N, W, D, K, B = (16384, 16, 128, 64, 64) if (len(sys.argv) < 2 or sys.argv[1] == "C5") else (4096, 16, 128, 32, 64)
dev = torch.device("cuda:0")
torch.manual_seed(5)
model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).to(dev)
model.train()
trainer = WindowShardedTrainer(model, lr=1e-3)
xs = [torch.rand(B, N, W, device=dev) for _ in range(4)]
ys = [torch.rand(B, N, device=dev) for _ in range(4)]
for i in range(13):
    trainer.step(xs[i % 4], ys[i % 4])
torch.cuda.synchronize()
hx = [x.cpu().pin_memory() for x in xs]
hy = [y.cpu().pin_memory() for y in ys]
def batches(n):
    for i in range(n):
        yield hx[i % 4], hy[i % 4]
def run(n):
    out = []
    for bx, by in Prefetcher(batches(n), dev, skip=(), reuse_buffers=True):
        out.append(trainer.step(bx, by).item())
    return out
run(3)
torch.cuda.synchronize()
lib.gdn_profile_enable(1)
t = time.perf_counter()
run(10)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10 * 1e3
buf = C.create_string_buffer(1 << 16)
lib.gdn_profile_collect(buf, len(buf))
lib.gdn_profile_enable(0)
print(f"e2e (profiled, serialising) {dt:.3f} ms/step")
rows = []
for ln in buf.value.decode().splitlines():
    nm, cnt, ms = ln.rsplit(" ", 2)
    rows.append((float(ms) / 10, nm, int(cnt)))
for ms, nm, cnt in sorted(rows, reverse=True)[:10]:
    print(f"   {nm:24s} {ms:.4f} ms/step  ({cnt} launches)")
t = time.perf_counter()
run(10)
torch.cuda.synchronize()
print(f"e2e (plain) {(time.perf_counter() - t) / 10 * 1e3:.3f} ms/step")
