"""Host-side timeline of the threaded feed at C5: where does a slow e2e run lose its overlap?
python tools/e2e_trace.py [reps] -- prints, per repetition, ms/step and the mean duration of each host phase."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from gdn_b200 import data as gdata
from gdn_b200.data import LossReader, Prefetcher
from gdn_b200.dp import WindowShardedTrainer
from gdn_b200.models.GDN import GDN

wl = WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "C5"]
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N, W, D, K, B = wl["N"], wl["W"], wl["D"], wl["K"], wl["B"]
dev = torch.device("cuda", 0)
torch.manual_seed(5)
model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).to(dev).train()
trainer = WindowShardedTrainer(model, lr=1e-3)
hx = [torch.rand(B, N, W, dtype=torch.float64) for _ in range(4)]
hy = [torch.rand(B, N, dtype=torch.float64) for _ in range(4)]
T = {}


def timed(cls, name):
    orig = getattr(cls, name)

    def wrap(self, *a, **k):
        t = time.perf_counter()
        try:
            return orig(self, *a, **k)
        finally:
            T.setdefault(name, []).append((time.perf_counter() - t) * 1e3)
    setattr(cls, name, wrap)


for n in ("_stage", "_host_copy", "_issue", "_to_device"):
    timed(Prefetcher, n)
timed(LossReader, "push")


def batches(n):
    for i in range(n):
        yield hx[i % 4], hy[i % 4]


def run(n):
    reader = LossReader(dev)
    gaps, t_prev = [], None
    step_ms = []
    for bx, by in Prefetcher(batches(n), dev, skip=(), reuse_buffers=True, threaded=True):
        t0 = time.perf_counter()
        if t_prev is not None:
            gaps.append((t0 - t_prev) * 1e3)           # time the consumer waited for the batch
        loss = trainer.step(bx, by)
        t1 = time.perf_counter()
        reader.push(loss)
        t_prev = time.perf_counter()
        step_ms.append((t1 - t0) * 1e3)
    reader.flush()
    torch.cuda.synchronize()
    return gaps, step_ms


run(6)
for rep in range(REPS):
    T.clear()
    t = time.perf_counter()
    gaps, step_ms = run(40)
    ms = (time.perf_counter() - t) / 40 * 1e3
    mean = lambda v: sum(v) / max(len(v), 1)
    print(f"rep {rep}: {ms:.2f} ms/step | consumer: wait-for-batch {mean(gaps):.2f} issue-step {mean(step_ms):.2f} push {mean(T.get('push', [0])):.2f}"
          f" | worker: _issue {mean(T.get('_issue', [0])):.2f} (x) _stage {mean(T['_stage'][0::2]):.2f} host_copy {mean(T['_host_copy'][0::2]):.2f}"
          f" to_device {mean(T['_to_device'][0::2]):.2f}  max host_copy {max(T['_host_copy']):.2f}", flush=True)
