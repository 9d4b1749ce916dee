import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_b200 import _lib
from gdn_b200._lib import ptr
lib = _lib.load()
torch.manual_seed(0)
N, D, K = 16384, 128, 64
V = (torch.rand(N, D, device="cuda") * 2 - 1) / D ** 0.5
idx = torch.empty((N, K), dtype=torch.int64, device="cuda"); nbr = torch.empty((N, K + 1), dtype=torch.int32, device="cuda")
ws = torch.zeros(lib.gdn_graph_build_ws_bytes(N, D, K), dtype=torch.uint8, device="cuda")
lib.gdn_graph_build(ptr(V), N, D, K, ptr(idx), ptr(nbr), ptr(ws), ws.numel(), 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
al = lambda v: (v + 255) // 256 * 256
C_ = 256
off = al(N * 4) + 2 * al(N * D * 2) + 2 * al(N * C_ * 4) + al(N * 4) + ((N + 63) // 64 + 4) * 4
o = ws[off:off + 48].view(torch.int64).cpu().tolist()
print("cycles: wait %d  tmem-load %d  filter(+compact) %d  compact %d   compactions(warp-level) %d  appends(lane 3) %d" % tuple(o))
