"""Helpers to load tests/golden/*.npz (written by oracle/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GDN_CASES = ["c1_msl", "c1_stress", "c2_swat", "c3_wadi", "w16_small", "w10_odd", "mlp2"]
GDN_CASES_L1 = [c for c in GDN_CASES if c != "mlp2"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def state_dict(rec, prefix="sd/", dtype=None):
    sd = {}
    for k, v in rec.items():
        if k.startswith(prefix):
            t = torch.from_numpy(np.array(v))
            if dtype is not None and t.is_floating_point():
                t = t.to(dtype)
            sd[k[len(prefix):]] = t
    return sd


def meta(rec):
    N, W, D, K, B, L, inter = [int(v) for v in rec["meta"]]
    return dict(N=N, W=W, D=D, K=K, B=B, L=L, inter=inter)


def normwise(a, b):
    """max|a-b| / max|b| (element-wise relative error is meaningless for
    forecasts that cross zero)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
