"""GraphLayer fwd+bwd at the module boundary: CUDA-event times (L2 flushed between iterations) and the library's
per-kernel breakdown, one JSON line.   python tools/gl_time.py C5 [reps]     (env switches: see csrc/attention.cu)"""
import ctypes
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS, graphlayer_bytes, profile_collect
from gdn_b200 import _lib, ops
from gdn_b200.models.graph_layer import GraphLayer

name = sys.argv[1] if len(sys.argv) > 1 else "C5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wl = WORKLOADS[name]
N, W, D, K, B = wl["N"], wl["W"], wl["D"], wl["K"], wl["B"]
torch.manual_seed(5)
dev = torch.device("cuda", 0)
lib = _lib.load()
layer = GraphLayer(W, D, heads=1, concat=False).to(dev)
V = ((torch.rand(N, D, device=dev) * 2 - 1) / D ** 0.5).requires_grad_(True)
x = torch.rand(B, N, W, device=dev)
gout = torch.rand(B * N, D, device=dev)
_, nbr = ops.graph_build(V.detach(), K)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fwd, bwd = [], []
for i in range(reps + 3):
    flush.fill_(i & 0xFF)
    a, b_, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    out = layer.forward_batched(x, nbr, V)
    b_.record()
    out.backward(gout)
    c.record()
    torch.cuda.synchronize()
    if i >= 3:
        fwd.append(a.elapsed_time(b_))
        bwd.append(b_.elapsed_time(c))
    layer.zero_grad(set_to_none=True)
    V.grad = None
lib.gdn_profile_enable(1)
for i in range(3):
    out = layer.forward_batched(x, nbr, V)
    out.backward(gout)
    layer.zero_grad(set_to_none=True)
    V.grad = None
torch.cuda.synchronize()
_, rows = profile_collect(lib)
lib.gdn_profile_enable(0)
fb, bb = graphlayer_bytes(wl)
f, b = statistics.mean(fwd), statistics.mean(bwd)
peak = 6552.3
print(json.dumps({"workload": name, "env": {k: v for k, v in os.environ.items() if k.startswith("GDN_")},
                  "fwd_ms": round(f, 4), "bwd_ms": round(b, 4), "frac": round((fb + bb) / ((f + b) * 1e-3) / 1e9 / peak, 4),
                  "kernels_ms": {k: round(t / c, 4) for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])}}))
