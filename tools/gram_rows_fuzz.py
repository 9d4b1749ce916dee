"""Randomised equality sweep: k_gram_rows (graphs of <= 2048 sensors) against the 64x64 tile kernel
(GDN_GRAM_ROWS=0), full builds and row ranges, incl. duplicated / scaled / zero rows.  python tools/gram_rows_fuzz.py [cases]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gdn_b200 import ops

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(7)
bad = 0
for c in range(cases):
    N = rng.choice([2, 3, 17, 27, 31, 32, 33, 51, 64, 65, 100, 127, 128, 129, 255, 300, 511, 777, 1024, 1500, 2047, 2048])
    D = rng.choice([1, 3, 16, 33, 64, 100, 128, 200, 256])
    K = rng.randint(1, min(N, 256))
    torch.manual_seed(c)
    V = (torch.rand(N, D, device="cuda") * 2 - 1)
    if c % 3 == 0:
        V = torch.round(V * 4) / 4                      # coarse grid: many exact ties
    if N > 8:
        V[rng.randrange(N)] = V[rng.randrange(N)]
        V[rng.randrange(N)] = V[rng.randrange(N)] * 0.5
        if c % 4 == 0:
            V[rng.randrange(N)] = 0.0
    outs = []
    for env in ("0", None):
        if env is None:
            os.environ.pop("GDN_GRAM_ROWS", None)
        else:
            os.environ["GDN_GRAM_ROWS"] = env
        kth = torch.full((N,), 3.0, device="cuda")
        idx, nbr = ops.graph_build(V, K, use_tensor_cores=0, kth=kth)
        torch.cuda.synchronize()
        outs.append((idx, nbr, kth))
    same = all(torch.equal(a, b) for a, b in zip(outs[0][:2], outs[1][:2])) and \
        torch.equal(outs[0][2].view(torch.int32), outs[1][2].view(torch.int32))
    if N >= 256:
        r0 = 128 * rng.randrange(N // 128)
        r1 = min(N, r0 + 128 * rng.randint(1, 3))
        if r1 % 128 and r1 != N:
            r1 = N
        out = (torch.full((N, K), -9, dtype=torch.int64, device="cuda"), torch.full((N, K + 1), -9, dtype=torch.int32, device="cuda"))
        i2, n2 = ops.graph_build(V, K, use_tensor_cores=0, rows=(r0, r1), out=out)
        same = same and torch.equal(i2[r0:r1], outs[1][0][r0:r1]) and torch.equal(n2[r0:r1], outs[1][1][r0:r1]) \
            and bool((i2[:r0] == -9).all()) and bool((i2[r1:] == -9).all())
    if not same:
        bad += 1
        print("MISMATCH", N, D, K, "case", c)
print(f"{cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
