"""CPU restatement of the reference's GDN forward/backward hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.

Every function follows the reference's own op sequence (edge lists, [E, D] per-edge
temporaries, scatter softmax) rather than the closed forms the CUDA kernels use, so
that (i) it checks those closed forms independently and (ii) timing it on host cores
is a fair stand-in for the reference's ``run.sh cpu`` path (``cpu_baseline.kind`` =
"port").  It is pinned against the reference's real Python files by
``oracle/make_golden.py`` (tests/golden/*.npz, tests/test_oracle_golden.py).

Parameter dictionaries use the reference's ``state_dict`` key names
(models/GDN.py:82-120).
"""
import math

import torch
import torch.nn.functional as F

NEG_SLOPE = 0.2      # models/graph_layer.py:13
BN_EPS = 1e-5        # nn.BatchNorm1d default (models/GDN.py:67,96)
BN_MOMENTUM = 0.1
DROP_P = 0.2         # models/GDN.py:114

_G = "gnn_layers.0.gnn."


# --------------------------------------------------------------------------- params
def init_state(node_num, dim, input_dim, out_layer_num=1, out_layer_inter_dim=256,
               seed=5, dtype=torch.float32, stressed=False):
    """Fresh parameters + buffers with the reference's initialisers.

    Order and distributions follow models/GDN.py:95-119 and
    models/graph_layer.py:41-49 (glorot for lin/att_i/att_j, zeros for att_em_*/bias,
    kaiming-uniform(a=sqrt(5)) for the embedding).  ``stressed=True`` replaces the
    zero/identity tensors by random ones so that every term of the path is exercised
    (SURVEY.md §8d).  Not bit-identical to constructing the reference module (RNG
    consumption order differs); parity tests load identical tensors into both sides.
    """
    g = torch.Generator().manual_seed(seed)
    N, D, W = node_num, dim, input_dim

    def uni(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul_(bound).to(dtype)

    sd = {}
    # nn.init.kaiming_uniform_(embedding.weight, a=sqrt(5)): bound = sqrt(6/((1+a^2)*fan_in)) = 1/sqrt(D)
    sd["embedding.weight"] = uni((N, D), 1.0 / math.sqrt(D))
    for bn, width in (("bn_outlayer_in", D), ("gnn_layers.0.bn", D)):
        sd[bn + ".weight"] = torch.ones(width, dtype=dtype)
        sd[bn + ".bias"] = torch.zeros(width, dtype=dtype)
        sd[bn + ".running_mean"] = torch.zeros(width, dtype=dtype)
        sd[bn + ".running_var"] = torch.ones(width, dtype=dtype)
        sd[bn + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    sd[_G + "lin.weight"] = uni((D, W), math.sqrt(6.0 / (D + W)))
    for name in ("att_i", "att_j"):
        sd[_G + name] = uni((1, 1, D), math.sqrt(6.0 / (1 + D)))
    for name in ("att_em_i", "att_em_j"):
        sd[_G + name] = torch.zeros((1, 1, D), dtype=dtype)
    sd[_G + "bias"] = torch.zeros(D, dtype=dtype)
    # OutLayer (models/GDN.py:27-43): nn.Linear default init = U(+-1/sqrt(fan_in))
    fan = D
    pos = 0
    for layer in range(out_layer_num):
        last = layer == out_layer_num - 1
        width = 1 if last else out_layer_inter_dim
        sd[f"out_layer.mlp.{pos}.weight"] = uni((width, fan), 1.0 / math.sqrt(fan))
        sd[f"out_layer.mlp.{pos}.bias"] = uni((width,), 1.0 / math.sqrt(fan))
        pos += 1
        if not last:
            sd[f"out_layer.mlp.{pos}.weight"] = torch.ones(width, dtype=dtype)
            sd[f"out_layer.mlp.{pos}.bias"] = torch.zeros(width, dtype=dtype)
            sd[f"out_layer.mlp.{pos}.running_mean"] = torch.zeros(width, dtype=dtype)
            sd[f"out_layer.mlp.{pos}.running_var"] = torch.ones(width, dtype=dtype)
            sd[f"out_layer.mlp.{pos}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
            pos += 2  # BatchNorm1d, ReLU
            fan = width
    if stressed:
        for name in ("att_em_i", "att_em_j"):
            sd[_G + name] = (torch.randn((1, 1, D), generator=g, dtype=torch.float64) * 0.5).to(dtype)
        sd[_G + "bias"] = (torch.randn(D, generator=g, dtype=torch.float64) * 0.3).to(dtype)
        for key in list(sd):
            if key.endswith("running_mean"):
                sd[key] = (torch.randn(sd[key].shape, generator=g, dtype=torch.float64) * 0.2).to(dtype)
            elif key.endswith("running_var"):
                sd[key] = (torch.rand(sd[key].shape, generator=g, dtype=torch.float64) + 0.5).to(dtype)
            elif (".bn." in key or "bn_outlayer_in" in key or _is_mlp_bn(sd, key)) and key.endswith(".weight"):
                sd[key] = (torch.rand(sd[key].shape, generator=g, dtype=torch.float64) + 0.5).to(dtype)
            elif (".bn." in key or "bn_outlayer_in" in key or _is_mlp_bn(sd, key)) and key.endswith(".bias"):
                sd[key] = (torch.randn(sd[key].shape, generator=g, dtype=torch.float64) * 0.5).to(dtype)
    return sd


def _is_mlp_bn(sd, key):
    if not key.startswith("out_layer.mlp."):
        return False
    stem = key.rsplit(".", 1)[0]
    return (stem + ".running_mean") in sd


def cast_state(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}


def param_names(sd):
    """Keys that are nn.Parameters in the reference (everything but BN buffers)."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


# --------------------------------------------------------------------------- graph
def learned_graph(V, topk):
    """models/GDN.py:145-159: cosine Gram of the (detached) embedding, row-wise top-k."""
    w = V.detach().clone()
    gram = torch.matmul(w, w.T)
    nrm = w.norm(dim=-1)
    denom = torch.matmul(nrm.view(-1, 1), nrm.view(1, -1))
    cos = gram / denom
    idx = torch.topk(cos, topk, dim=-1)[1]
    return idx, cos


def batch_edges(idx, batch_num):
    """models/GDN.py:161-165 and :15-24: (src = top-k index, dst = row), replicated per
    window with node offset b*N."""
    N, K = idx.shape
    dst = torch.arange(N).unsqueeze(1).repeat(1, K).flatten().unsqueeze(0)
    src = idx.flatten().unsqueeze(0)
    one = torch.cat((src, dst), dim=0)
    reps = one.repeat(1, batch_num).contiguous()
    E = one.shape[1]
    for b in range(batch_num):
        reps[:, b * E:(b + 1) * E] += b * N
    return reps.long()


# --------------------------------------------------------------------------- GraphLayer
def _segment_softmax(logit, seg, num_seg):
    """PyG 1.5.0 utils.softmax (see oracle/pyg_shim.py)."""
    shape_tail = tuple(logit.shape[1:])
    ix = seg.view([-1] + [1] * (logit.dim() - 1)).expand_as(logit)
    top = logit.new_full((num_seg,) + shape_tail, float("-inf"))
    top = top.scatter_reduce(0, ix, logit.detach(), reduce="amax", include_self=True)
    ex = (logit - top.gather(0, ix)).exp()
    tot = torch.zeros((num_seg,) + shape_tail, dtype=logit.dtype).scatter_add(0, ix, ex)
    return ex / (tot.gather(0, ix) + 1e-16)


def graph_layer_forward(x, edge_index, embedding, lin_weight, att_i, att_j, att_em_i, att_em_j,
                        bias, heads=1, concat=False, negative_slope=NEG_SLOPE):
    """models/graph_layer.py:53-117 with PyG 1.5.0 propagate semantics.

    x [n, W]; edge_index [2, E] (row 0 = source j, row 1 = target i); embedding [n, D].
    Returns (out, (edge_index_with_self_loops, alpha [E', heads, 1])).
    """
    n = x.shape[0]
    C = lin_weight.shape[0] // heads
    xl = F.linear(x, lin_weight)                                   # :56
    keep = edge_index[0] != edge_index[1]                          # :61 remove_self_loops
    ei = edge_index[:, keep]
    loops = torch.arange(n, dtype=ei.dtype).unsqueeze(0).repeat(2, 1)
    ei = torch.cat([ei, loops], dim=1)                             # :62 add_self_loops
    src, dst = ei[0], ei[1]
    x_i = xl.index_select(0, dst).view(-1, heads, C)               # PyG __collect__
    x_j = xl.index_select(0, src).view(-1, heads, C)
    if embedding is not None:                                      # :91-96
        emb_i = embedding[dst].unsqueeze(1).repeat(1, heads, 1)
        emb_j = embedding[src].unsqueeze(1).repeat(1, heads, 1)
        key_i = torch.cat((x_i, emb_i), dim=-1)
        key_j = torch.cat((x_j, emb_j), dim=-1)
        cat_i = torch.cat((att_i, att_em_i), dim=-1)               # :100-101
        cat_j = torch.cat((att_j, att_em_j), dim=-1)
    else:  # the reference would fail here (key_i undefined); kept for completeness
        key_i, key_j, cat_i, cat_j = x_i, x_j, att_i, att_j
    logit = (key_i * cat_i).sum(-1) + (key_j * cat_j).sum(-1)      # :103
    logit = logit.view(-1, heads, 1)
    logit = F.leaky_relu(logit, negative_slope)                    # :109
    alpha = _segment_softmax(logit, dst, n)                        # :110
    msg = x_j * alpha.view(-1, heads, 1)                           # :117
    agg = torch.zeros((n, heads, C), dtype=msg.dtype).index_add(0, dst, msg)
    out = agg.view(-1, heads * C) if concat else agg.mean(dim=1)   # :68-71
    if bias is not None:
        out = out + bias                                           # :73-74
    return out, (ei, alpha)


# --------------------------------------------------------------------------- GDN
def dropout_mask(batch_num, node_num, dim, seed, p=DROP_P, dtype=torch.float32):
    """A dropout keep-mask [B, N, D] with values in {0, 1/(1-p)}.

    Drawn exactly as nn.Dropout draws it for the reference's activation layout
    (models/GDN.py:178-182: the input is the permuted view of a contiguous [B, D, N]
    buffer and torch fills the mask in memory order)."""
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    m = F.dropout(torch.ones(batch_num, dim, node_num).permute(0, 2, 1), p, True)
    torch.set_rng_state(state)
    return m.contiguous().to(dtype)


def gdn_forward(sd, data, topk, training=False, drop_mask=None, update_buffers=True, idx=None):
    """models/GDN.py:122-187.  ``sd`` holds parameters *and* BN buffers (buffers are
    updated in place in training mode, as nn.BatchNorm1d does).  ``drop_mask`` [B,N,D]
    in {0, 1/(1-p)} replaces nn.Dropout's RNG (required when training).
    ``idx`` [N, K] int64, optional: use this learned graph instead of computing it (the graph is
    built without gradient, models/GDN.py:145, so everything downstream is a function of it) --
    lets a test compare the float path on the graph under test when near-tied cosines make two
    exact-arithmetic builds differ in a few rows.
    Returns (pred [B, N], aux dict)."""
    x = data.clone().detach()
    B, N, W = x.shape
    x = x.view(-1, W).contiguous()
    V = sd["embedding.weight"]
    if idx is None:
        idx, _ = learned_graph(V, topk)                            # :143-159
    else:
        idx = torch.as_tensor(idx).long()
        if tuple(idx.shape) != (N, topk):
            raise ValueError(f"idx must be [N, K] = {(N, topk)}, got {tuple(idx.shape)}")
    edges = batch_edges(idx, B)                                    # :161-165
    emb_rep = V.repeat(B, 1)                                       # :146
    z, (ei, alpha) = graph_layer_forward(
        x, edges, emb_rep, sd[_G + "lin.weight"], sd[_G + "att_i"], sd[_G + "att_j"],
        sd[_G + "att_em_i"], sd[_G + "att_em_j"], sd[_G + "bias"], heads=1, concat=False)
    h = _bn(sd, "gnn_layers.0.bn", z, training, update_buffers)    # GNNLayer :77
    h = F.relu(h)                                                  # :79
    h = h.view(B, N, -1)                                           # :172
    out = torch.mul(h, V)                                          # :176
    out = out.permute(0, 2, 1)
    out = F.relu(_bn(sd, "bn_outlayer_in", out, training, update_buffers))  # :179
    out = out.permute(0, 2, 1)
    if training:                                                   # :182
        if drop_mask is None:
            raise ValueError("training-mode oracle needs an explicit dropout mask")
        out = out * drop_mask
    pos = 0
    while f"out_layer.mlp.{pos}.weight" in sd:                     # OutLayer :45-56
        key = f"out_layer.mlp.{pos}"
        if key + ".running_mean" in sd:
            out = _bn(sd, key, out.permute(0, 2, 1), training, update_buffers).permute(0, 2, 1)
            out = F.relu(out)
            pos += 2
        else:
            out = F.linear(out, sd[key + ".weight"], sd[key + ".bias"])
            pos += 1
    pred = out.view(-1, N)                                         # :184
    return pred, {"learned_graph": idx, "edge_index": ei, "alpha": alpha, "z": z}


def _bn(sd, key, x, training, update_buffers):
    rm, rv = sd[key + ".running_mean"], sd[key + ".running_var"]
    if training and not update_buffers:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, sd[key + ".weight"], sd[key + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if training and update_buffers:
        sd[key + ".num_batches_tracked"] += 1
    return y


def mse_loss(pred, y):
    """train.py:20-23."""
    return F.mse_loss(pred, y, reduction="mean")


def loss_and_grads(sd, data, y, topk, drop_mask=None, training=True, update_buffers=False, idx=None):
    """One forward + backward (train.py:69-72) -> (loss, pred, {param: grad})."""
    names = param_names(sd)
    work = dict(sd)
    leaves = {}
    for k in names:
        leaves[k] = sd[k].detach().clone().requires_grad_(True)
        work[k] = leaves[k]
    pred, aux = gdn_forward(work, data, topk, training=training, drop_mask=drop_mask,
                            update_buffers=update_buffers, idx=idx)
    if update_buffers:
        for k in sd:
            if k not in leaves:
                sd[k] = work[k]
    loss = mse_loss(pred, y)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    out = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}
    return loss.detach(), pred.detach(), out, aux


class OracleTrainer:
    """The reference's train step (train.py:31,68-73) on host cores: Adam(lr=1e-3),
    zero_grad -> forward -> mse -> backward -> step.  Dropout masks are drawn by
    torch's RNG like nn.Dropout does.  Used as the CPU baseline."""

    def __init__(self, sd, topk, lr=1e-3, weight_decay=0.0):
        self.sd = {k: v.clone() for k, v in sd.items()}
        self.topk = topk
        self.names = param_names(self.sd)
        for k in self.names:
            self.sd[k].requires_grad_(True)
        self.opt = torch.optim.Adam([self.sd[k] for k in self.names], lr=lr, weight_decay=weight_decay)

    def train_step(self, data, y):
        B, N, _ = data.shape
        D = self.sd["embedding.weight"].shape[1]
        self.opt.zero_grad()
        mask = F.dropout(torch.ones(B, D, N).permute(0, 2, 1), DROP_P, True)
        pred, _ = gdn_forward(self.sd, data, self.topk, training=True, drop_mask=mask)
        loss = mse_loss(pred, y)
        loss.backward()
        self.opt.step()
        return float(loss.item())

    @torch.no_grad()
    def eval_step(self, data, y):
        pred, _ = gdn_forward(self.sd, data, self.topk, training=False)
        return pred, float(mse_loss(pred, y).item())
