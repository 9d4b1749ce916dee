"""tcgen05 graph engine vs the exact fp32 engine: equality, flags, protocol errors, timing."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_b200 import _lib
from gdn_b200._lib import ptr

lib = _lib.load()

def build(V, K, tc):
    N, D = V.shape
    idx = torch.empty((N, K), dtype=torch.int64, device=V.device)
    nbr = torch.empty((N, K + 1), dtype=torch.int32, device=V.device)
    nb = lib.gdn_graph_build_ws_bytes(N, D, K)
    ws = torch.zeros(nb, dtype=torch.uint8, device=V.device)
    rc = lib.gdn_graph_build(ptr(V), N, D, K, ptr(idx), ptr(nbr), ptr(ws), ws.numel(), tc,
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    if rc != 0:
        print("   rc", rc, lib.gdn_last_error())
    return idx, nbr, ws

def flags_of(ws, N, D, K):
    al = lambda v: (v + 255) // 256 * 256
    nblk = (N + 63) // 64
    off = lib.gdn_graph_build_ws_bytes(N, D, K) - al((nblk + 68) * 4)     # flags + error word close the TC workspace
    w = ws[off:off + (nblk + 4) * 4].view(torch.int32).cpu()
    return int(w[:nblk].sum()), int(w[nblk])

def case(name, V, K):
    N, D = V.shape
    i0, n0, _ = build(V, K, 0)
    i1, n1, ws = build(V, K, 1)
    nflag, err = flags_of(ws, N, D, K)
    bad = (i0 != i1).any(dim=1)
    print(f"{name}: N={N} D={D} K={K}  rows differing {int(bad.sum())}  nbr equal {torch.equal(n0, n1)}  "
          f"flagged blocks {nflag}  protocol err {err}")
    if bad.any():
        r = int(bad.nonzero()[0])
        print("   row", r, "exact", i0[r, :8].tolist(), "tc", i1[r, :8].tolist())
    for tc in (0, 1):
        for _ in range(2):
            build(V, K, tc)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            idx = torch.empty((N, K), dtype=torch.int64, device=V.device)
            nbr = torch.empty((N, K + 1), dtype=torch.int32, device=V.device)
            lib.gdn_graph_build(ptr(V), N, D, K, ptr(idx), ptr(nbr), ptr(ws), ws.numel(), tc,
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        b.record()
        torch.cuda.synchronize()
        print(f"   engine {'tcgen05' if tc else 'fp32   '}: {a.elapsed_time(b) / 5:.3f} ms")

torch.manual_seed(0)
dev = "cuda"
kaiming = lambda N, D: (torch.rand(N, D, device=dev) * 2 - 1) / D ** 0.5
case("C4 shape", kaiming(4096, 128), 32)
case("ragged", kaiming(1500, 128), 17)
case("dim 64", kaiming(2048, 64), 64)
case("C5 shape", kaiming(16384, 128), 64)
V = kaiming(2048, 128)
V[100:140] = V[100]                 # 40 identical sensors: exact ties, more than the slack can hold
V[900] = V[5] * 3.0                 # same direction, different norm
case("duplicates", V, 16)
