"""Timing of the warm-started tensor-core graph build over a sequence of Adam-sized embedding steps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_b200 import _lib, ops
lib = _lib.load()
MARGIN = float(os.environ.get("TC_MARGIN", "0.03"))
torch.manual_seed(0)
for N, D, K in ((4096, 128, 32), (16384, 128, 64)):
    V = (torch.rand(N, D, device="cuda") * 2 - 1) / D ** 0.5
    kth = torch.full((N,), float("-inf"), device="cuda")
    ops.graph_build(V, K, use_tensor_cores=1, kth=kth)
    torch.cuda.synchronize()
    lib.gdn_profile_enable(1)
    for _ in range(5):
        V = V + 1e-3 * torch.sign(torch.randn_like(V))
        ops.graph_build(V, K, use_tensor_cores=1, kth=kth, margin=MARGIN)
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    lib.gdn_profile_collect(buf, len(buf))
    lib.gdn_profile_enable(0)
    print(f"N={N} K={K} warm-started (margin {MARGIN}), lr=1e-3 sign steps")
    for ln in buf.value.decode().splitlines():
        nm, cnt, ms = ln.rsplit(" ", 2)
        print(f"   {nm:24s} {float(ms) / int(cnt):.4f} ms")
