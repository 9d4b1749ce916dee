import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_b200 import ops
torch.manual_seed(0)
N, D, K = 16384, 128, 64
V = (torch.rand(N, D, device="cuda") * 2 - 1) / D ** 0.5
for _ in range(2):
    idx, nbr = ops.graph_build(V, K, use_tensor_cores=1)
torch.cuda.synchronize()
print("ok", int(idx.sum()))
