"""Which feed breaks the trainer's CUDA-graph capture?  python tools/capture_repro.py <variant>"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from gdn_b200.data import LossReader, Prefetcher
from gdn_b200.dp import WindowShardedTrainer
from gdn_b200.models.GDN import GDN

variant = sys.argv[1]
wl = WORKLOADS["C1"]
N, W, D, K, B = wl["N"], wl["W"], wl["D"], wl["K"], wl["B"]
dev = torch.device("cuda", 0)
torch.manual_seed(5)
model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).to(dev).train()
trainer = WindowShardedTrainer(model, lr=1e-3)
hx = [torch.rand(B, N, W, dtype=torch.float64) for _ in range(4)]
hy = [torch.rand(B, N, dtype=torch.float64) for _ in range(4)]


def batches(n):
    for i in range(n):
        yield hx[i % 4], hy[i % 4]


if os.environ.get("GDN_DBG"):
    from gdn_b200 import ops as _ops
    import ctypes as _C, threading as _th
    def _dbg_stream():
        st = torch.cuda.current_stream()
        print("  _stream:", hex(st.cuda_stream), "capturing", torch.cuda.is_current_stream_capturing(), "thread", _th.current_thread().name, flush=True)
        return _C.c_void_p(st.cuda_stream)
    _ops._stream = _dbg_stream

try:
    if variant == "resident_item":
        x, y = hx[0].float().to(dev), hy[0].float().to(dev)
        for _ in range(5):
            trainer.step(x, y).item()
    elif variant == "resident_keep":
        x, y = hx[0].float().to(dev), hy[0].float().to(dev)
        for _ in range(5):
            loss = trainer.step(x, y)
    elif variant == "resident_pf_unused":
        pf = Prefetcher(batches(6), dev, skip=(), reuse_buffers=True)
        x, y = hx[0].float().to(dev), hy[0].float().to(dev)
        for _ in range(5):
            trainer.step(x, y)
    elif variant == "resident_pf_lockstep":
        x, y = hx[0].float().to(dev), hy[0].float().to(dev)
        for bx, by in Prefetcher(batches(6), dev, skip=(), reuse_buffers=True):
            trainer.step(x, y)
    elif variant == "prefetch_noitem":
        for bx, by in Prefetcher(batches(6), dev, skip=(), reuse_buffers=True):
            trainer.step(bx, by)
    elif variant == "prefetch_clone":
        for bx, by in Prefetcher(batches(6), dev, skip=(), reuse_buffers=True):
            trainer.step(bx.clone(), by.clone())
    elif variant == "resident":
        x, y = hx[0].float().to(dev), hy[0].float().to(dev)
        for _ in range(5):
            trainer.step(x, y)
    elif variant == "sidestream":
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            x, y = hx[0].float().to(dev), hy[0].float().to(dev)
        torch.cuda.current_stream().wait_stream(s)
        for _ in range(5):
            trainer.step(x, y)
    elif variant.startswith("prefetch"):
        threaded = "thr" in variant
        reuse = "noreuse" not in variant
        for bx, by in Prefetcher(batches(6), dev, skip=(), reuse_buffers=reuse, threaded=threaded, stage_threads=1):
            loss = trainer.step(bx, by)
            loss.item()
    elif variant == "prefetch_after_warm":
        x, y = hx[0].float().to(dev), hy[0].float().to(dev)
        for _ in range(5):
            trainer.step(x, y)
        for bx, by in Prefetcher(batches(6), dev, skip=(), reuse_buffers=True, threaded=True, stage_threads=1):
            trainer.step(bx, by).item()
    torch.cuda.synchronize()
    print(variant, "OK graphs:", len(trainer._graphs))
except Exception as e:
    tb = traceback.format_exc()
    print(variant, "FAILED:", str(e).splitlines()[0])
    print("\n".join(l for l in tb.splitlines() if "File" in l)[-1500:])
