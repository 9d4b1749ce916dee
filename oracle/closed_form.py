"""Vectorised CPU twin of the algebra the CUDA kernels implement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Unlike oracle/gdn_oracle.py -- which follows
the reference's op sequence edge by edge -- this file follows the KERNELS' closed forms step by
step (per-node scalars, aggregation before the linear map, BatchNorm-1 statistics from the
moments of A, the five D-wide passes of the backward), so that
  * tests/test_closed_form_cpu.py can prove on the CPU, in float64, that those closed forms equal
    the reference's forward and autograd, and
  * precision questions about a kernel can be reproduced in float32 without a GPU.
"""
import torch

NEG = 0.2
EPS_BN = 1e-5
G = "gnn_layers.0.gnn."


def neighbour_table(idx):
    """[N, K] top-k indices -> [N, K+1] list after remove_self_loops/add_self_loops (-1 padded)."""
    N, K = idx.shape
    nbr = torch.full((N, K + 1), -1, dtype=torch.long)
    for i in range(N):
        row = [int(j) for j in idx[i] if int(j) != i] + [i]
        nbr[i, :len(row)] = torch.tensor(row)
    return nbr


def forward_backward(sd, x, y, idx, drop_mask, bn1_from_moments=True, centred_gsi=True):
    """One training forward + backward in the dtype of `x`.  Returns (pred, loss, grads)."""
    dt = x.dtype
    P = {k: v.to(dt) for k, v in sd.items() if v.is_floating_point()}
    B, N, W = x.shape
    V = P["embedding.weight"]
    D = V.shape[1]
    Wl = P[G + "lin.weight"]
    a_i, a_j = P[G + "att_i"].view(D), P[G + "att_j"].view(D)
    ae_i, ae_j = P[G + "att_em_i"].view(D), P[G + "att_em_j"].view(D)
    bias = P[G + "bias"]
    g1, be1 = P["gnn_layers.0.bn.weight"], P["gnn_layers.0.bn.bias"]
    g2, be2 = P["bn_outlayer_in.weight"], P["bn_outlayer_in.bias"]
    wo, bo = P["out_layer.mlp.0.weight"].view(D), P["out_layer.mlp.0.bias"].view(())
    n = B * N
    nbr = neighbour_table(idx)
    valid = nbr >= 0
    src = nbr.clamp(min=0)

    # ---- attention (csrc/attention.cu)
    u_i, u_j = Wl.T @ a_i, Wl.T @ a_j
    e_i, e_j = V @ ae_i, V @ ae_j
    s_i = x @ u_i + e_i                                  # [B, N]
    s_j = x @ u_j + e_j
    pre = s_i.unsqueeze(-1) + s_j[:, src]                # [B, N, Kp]
    lr = torch.where(pre > 0, pre, NEG * pre)
    lr = lr.masked_fill(~valid, float("-inf"))
    m = lr.max(dim=-1, keepdim=True)[0]
    p = (lr - m).exp()
    linv = 1.0 / (p.sum(-1, keepdim=True) + 1e-16)
    alpha = p * linv
    xs = x[:, src]                                        # [B, N, Kp, W]
    A = (alpha.unsqueeze(-1) * xs).sum(2)                 # [B, N, W]

    # ---- BN1 statistics (csrc/dwide.cu k_moments + k_fin_bn1) and the D-wide chain
    Af = A.reshape(n, W)
    if bn1_from_moments:
        m1 = Af.double().mean(0)
        m2 = (Af.double().T @ Af.double()) / n
        cov = m2 - torch.outer(m1, m1)
        mean1 = (Wl.double() @ m1 + bias.double())
        var1 = ((Wl.double() @ cov) * Wl.double()).sum(1).clamp(min=0)
        mean1, var1 = mean1.to(dt), var1.to(dt)
    else:
        Zf = Af @ Wl.T + bias
        mean1, var1 = Zf.mean(0), Zf.var(0, unbiased=False)
    istd1 = 1.0 / torch.sqrt(var1 + EPS_BN)
    zp = Af @ Wl.T                                        # z' (no bias)
    xh1 = zp * istd1 + (bias - mean1) * istd1
    y1 = g1 * xh1 + be1
    r1 = y1.clamp(min=0)
    Vrep = V.repeat(B, 1)
    pp = r1 * Vrep
    mean2 = pp.double().mean(0).to(dt)
    var2 = (pp.double() ** 2).mean(0).to(dt) - mean2 ** 2
    var2 = var2.clamp(min=0)
    istd2 = 1.0 / torch.sqrt(var2 + EPS_BN)
    xh2 = pp * istd2 - mean2 * istd2
    y2 = g2 * xh2 + be2
    h2 = y2.clamp(min=0)
    kf = drop_mask.reshape(n, D).to(dt)
    hm = h2 * kf
    pred = (hm @ wo + bo).view(B, N)
    loss = ((pred - y) ** 2).mean()

    # ---- backward
    gp = (2.0 * (pred - y) / n).reshape(n, 1)
    g_wo = (gp * hm).sum(0)
    g_bo = gp.sum()
    gy2 = torch.where(y2 > 0, gp * wo * kf, torch.zeros((), dtype=dt))
    g_g2, g_b2 = (gy2 * xh2).sum(0), gy2.sum(0)
    gpp = (g2 * istd2) * (gy2 - g_b2 / n - xh2 * (g_g2 / n))
    g_V = (gpp * r1).view(B, N, D).sum(0)
    gy1 = torch.where(y1 > 0, gpp * Vrep, torch.zeros((), dtype=dt))
    g_g1, g_b1 = (gy1 * xh1).sum(0), gy1.sum(0)
    gz = (g1 * istd1) * (gy1 - g_b1 / n - xh1 * (g_g1 / n))
    g_bias = gz.sum(0)
    g_Wl = gz.T @ Af
    gA = (gz @ Wl).view(B, N, W)
    # attention backward
    ga = (gA.unsqueeze(2) * xs).sum(-1)                   # [B, N, Kp]
    ga = ga.masked_fill(~valid, 0.0)
    dot = (alpha * ga).sum(-1, keepdim=True)
    gl = alpha * (ga - dot)
    slope = torch.where(pre > 0, torch.ones((), dtype=dt), torch.full((), NEG, dtype=dt))
    gpre = (gl * slope).masked_fill(~valid, 0.0)
    if centred_gsi:
        # sum_k g_l = 0 exactly, so sum_k slope_k g_l = sum_k (slope_k - c) g_l for any c: take the
        # slope of the majority sign as c and only the minority edges contribute
        pos = ((pre > 0) & valid)
        npos = pos.sum(-1, keepdim=True)
        nval = valid.sum(-1, keepdim=True)
        maj_pos = npos * 2 >= nval
        contrib_neg = ((NEG - 1.0) * gl).masked_fill(pos | ~valid, 0.0).sum(-1)
        contrib_pos = ((1.0 - NEG) * gl).masked_fill(~pos, 0.0).sum(-1)
        g_si = torch.where(maj_pos.squeeze(-1), contrib_neg, contrib_pos)
    else:
        g_si = gpre.sum(-1)
    g_sj = torch.zeros(B, N, dtype=dt)
    g_sj.index_add_(1, src.reshape(-1), gpre.reshape(B, -1))
    g_ui = (g_si.unsqueeze(-1) * x).sum((0, 1))
    g_uj = (g_sj.unsqueeze(-1) * x).sum((0, 1))
    g_ei, g_ej = g_si.sum(0), g_sj.sum(0)
    grads = {
        "embedding.weight": g_V + torch.outer(g_ei, ae_i) + torch.outer(g_ej, ae_j),
        G + "lin.weight": g_Wl + torch.outer(a_i, g_ui) + torch.outer(a_j, g_uj),
        G + "att_i": (Wl @ g_ui).view(1, 1, D), G + "att_j": (Wl @ g_uj).view(1, 1, D),
        G + "att_em_i": (V.T @ g_ei).view(1, 1, D), G + "att_em_j": (V.T @ g_ej).view(1, 1, D),
        G + "bias": g_bias,
        "gnn_layers.0.bn.weight": g_g1, "gnn_layers.0.bn.bias": g_b1,
        "bn_outlayer_in.weight": g_g2, "bn_outlayer_in.bias": g_b2,
        "out_layer.mlp.0.weight": g_wo.view(1, D), "out_layer.mlp.0.bias": g_bo.view(1),
    }
    return pred, loss, grads
