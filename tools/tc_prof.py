import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_b200 import _lib, ops
lib = _lib.load()
torch.manual_seed(0)
for N, D, K in ((4096, 128, 32), (16384, 128, 64)):
    V = (torch.rand(N, D, device="cuda") * 2 - 1) / D ** 0.5
    for _ in range(2):
        ops.graph_build(V, K, use_tensor_cores=1)
    torch.cuda.synchronize()
    lib.gdn_profile_enable(1)
    for _ in range(5):
        ops.graph_build(V, K, use_tensor_cores=1)
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    lib.gdn_profile_collect(buf, len(buf))
    lib.gdn_profile_enable(0)
    print(f"N={N} K={K}")
    for ln in buf.value.decode().splitlines():
        nm, cnt, ms = ln.rsplit(" ", 2)
        print(f"   {nm:24s} {float(ms) / int(cnt):.4f} ms")
