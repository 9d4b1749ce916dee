import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_b200 import ops
from gdn_b200.dp import graph_row_shard
N, D, K = 16384, 128, 64
torch.manual_seed(N + K)
V = ((torch.rand(N, D, device="cuda") * 2 - 1) / D ** 0.5)
full_idx, full_nbr = ops.graph_build(V, K, use_tensor_cores=1)
torch.cuda.synchronize()
for world in (2, 3, 8):
    chunk = graph_row_shard(N, 0, world)[2]
    idx = torch.full((world * chunk, K), -7, dtype=torch.int64, device="cuda")
    nbr = torch.full((world * chunk, K + 1), -7, dtype=torch.int32, device="cuda")
    kth = torch.full((N,), float("-inf"), device="cuda")
    for rnd in range(2):
        for r in range(world):
            r0, r1, _ = graph_row_shard(N, r, world)
            if r1 > r0:
                ops.graph_build(V, K, use_tensor_cores=1, kth=kth, rows=(r0, r1), out=(idx, nbr))
                torch.cuda.synchronize()
                print("ok world", world, "rnd", rnd, "rank", r, (r0, r1), flush=True)
        print("   equal", torch.equal(idx[:N], full_idx), torch.equal(nbr[:N], full_nbr), flush=True)
