"""Generate tests/golden/*.npz from the reference's OWN Python files.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs /root/reference):

    python -m oracle.make_golden

Imports models/GDN.py, models/graph_layer.py and evaluate.py unmodified through the
PyG-1.5.0 stand-in (oracle/pyg_shim.py), runs them on seeded inputs and stores inputs +
outputs.  tests/test_oracle_golden.py then pins oracle/gdn_oracle.py and
oracle/scoring_oracle.py against these vectors, and the -m gpu tests pin the CUDA path.

What is stored per case (float32 unless noted):
  x [B,N,W], y [B,N], every state_dict tensor ("sd/<key>"), drop_mask [B,N,D],
  idx [N,K] int64 (learned_graph), pred_eval, alpha_eval [E,1,1], edge_index [2,E],
  pred_train, loss_train, grad/<param> (fp32 reference autograd),
  grad64/<param> + pred_train64 + loss_train64 (reference run in float64),
  after/<buffer> (BN running stats after the training forward).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import gdn_oracle as go            # noqa: E402
from oracle import pyg_shim                    # noqa: E402
from oracle import scoring_oracle as so        # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    #  name        N    W   D    K   B  L  inter stressed
    "c1_msl":    (27,   5,  64,  5, 32, 1, 128, False),
    "c1_stress": (27,   5,  64,  5, 32, 1, 128, True),
    "c2_swat":   (51,   5,  64, 15, 16, 1, 128, True),
    "c3_wadi":   (127,  5, 128, 30,  8, 1, 128, True),
    "w16_small": (96,  16, 128, 12,  4, 1, 128, True),
    "w10_odd":   (33,  10,  32,  7,  5, 1, 128, True),
    "mlp2":      (27,   5,  64,  5,  8, 2,  32, True),
}


def fc_edge_index(n):
    src, dst = [], []
    for i in range(n):
        for j in range(n):
            if i != j:
                src.append(j)
                dst.append(i)
    return torch.tensor([src, dst], dtype=torch.long)


def msl_windows(batch, win):
    import pandas as pd
    df = pd.read_csv("/root/reference/data/msl/train.csv", index_col=0)
    feats = [l.strip() for l in open("/root/reference/data/msl/list.txt")]
    data = torch.tensor(df[feats].values.T.copy()).double()          # [N, T]
    xs, ys = [], []
    for i in range(win, win + batch):                                 # TimeDataset.process
        xs.append(data[:, i - win:i])
        ys.append(data[:, i])
    return torch.stack(xs).float(), torch.stack(ys).float()


def build_reference_model(gdn_mod, N, W, D, K, L, inter, sd):
    model = gdn_mod.GDN([fc_edge_index(N)], N, dim=D, input_dim=W, out_layer_num=L,
                        out_layer_inter_dim=inter, topk=K)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model


def run_case(name, gdn_mod):
    N, W, D, K, B, L, inter, stressed = CASES[name]
    sd = go.init_state(N, D, W, out_layer_num=L, out_layer_inter_dim=inter, seed=5, stressed=stressed)
    g = torch.Generator().manual_seed(1234 + N)
    if name == "c1_msl":
        x, y = msl_windows(B, W)
    else:
        x = torch.rand(B, N, W, generator=g)
        y = torch.rand(B, N, generator=g)
    mask_seed = 77
    mask = go.dropout_mask(B, N, D, mask_seed)
    rec = {"x": x.numpy(), "y": y.numpy(), "drop_mask": mask.numpy(),
           "meta": np.array([N, W, D, K, B, L, inter], dtype=np.int64)}
    for k, v in sd.items():
        rec["sd/" + k] = v.numpy()

    # ---- reference, eval mode
    model = build_reference_model(gdn_mod, N, W, D, K, L, inter, sd)
    model.eval()
    with torch.no_grad():
        pred_eval = model(x, None)
    rec["pred_eval"] = pred_eval.numpy()
    rec["idx"] = model.learned_graph.numpy()
    rec["alpha_eval"] = model.gnn_layers[0].att_weight_1.detach().numpy()
    rec["edge_index"] = model.gnn_layers[0].edge_index_1.numpy()

    # ---- reference, training forward/backward (fp32)
    model = build_reference_model(gdn_mod, N, W, D, K, L, inter, sd)
    model.train()
    torch.manual_seed(mask_seed)             # nn.Dropout is the only RNG consumer in forward
    pred_train = model(x, None)
    loss = torch.nn.functional.mse_loss(pred_train, y, reduction="mean")
    loss.backward()
    rec["pred_train"] = pred_train.detach().numpy()
    rec["loss_train"] = np.array(loss.item(), dtype=np.float32)
    for k, p in model.named_parameters():
        rec["grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    for k, b in model.named_buffers():
        rec["after/" + k] = b.numpy()

    # ---- reference in float64 (gradient noise floor reference)
    model64 = build_reference_model(gdn_mod, N, W, D, K, L, inter, sd).double()
    model64.train()
    torch.manual_seed(mask_seed)
    pred64 = model64(x.double(), None)
    loss64 = torch.nn.functional.mse_loss(pred64, y.double(), reduction="mean")
    loss64.backward()
    rec["pred_train64"] = pred64.detach().numpy()
    rec["loss_train64"] = np.array(loss64.item(), dtype=np.float64)
    for k, p in model64.named_parameters():
        rec["grad64/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()

    # ---- cross-check the oracle right here (fails loudly if the restatement drifts)
    sd_o = {k: v.clone() for k, v in sd.items()}
    pe, aux = go.gdn_forward(sd_o, x, K, training=False)
    assert torch.equal(aux["learned_graph"], model.learned_graph), name
    err_eval = (pe - pred_eval).abs().max().item()
    l_o, pt, grads, _ = go.loss_and_grads(sd_o, x, y, K, drop_mask=mask, update_buffers=True)
    err_train = (pt - pred_train.detach()).abs().max().item()
    gerr = max((grads[k] - torch.from_numpy(rec["grad/" + k])).abs().max().item() for k in grads)
    berr = max((sd_o[k].double() - torch.from_numpy(rec["after/" + k]).double()).abs().max().item()
               for k in sd_o if "running" in k or "tracked" in k)
    print(f"{name:10s} oracle-vs-reference: eval {err_eval:.2e} train {err_train:.2e} "
          f"grad {gerr:.2e} buffers {berr:.2e} loss {abs(l_o.item() - loss.item()):.2e}")
    assert err_eval < 1e-5 and err_train < 1e-5 and gerr < 1e-5 and berr < 1e-5
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


def run_graph_layer_case(gl_mod):
    """GraphLayer at its own module boundary on a *general* edge list (ragged in-degrees,
    explicit self loops, duplicate edges, isolated targets), heads=1 concat=False as GDN
    uses it, plus heads=2 concat=True."""
    g = torch.Generator().manual_seed(99)
    n, W, D = 70, 6, 16
    E = 400
    src = torch.randint(0, n, (E,), generator=g)
    dst = torch.randint(0, n - 5, (E,), generator=g)          # last 5 nodes: self loop only
    src[:10] = dst[:10]                                        # explicit self loops get dropped
    ei = torch.stack([src, dst])
    x = torch.rand(n, W, generator=g)
    emb = torch.randn(n, D, generator=g) * 0.3
    rec = {"x": x.numpy(), "edge_index": ei.numpy(), "embedding": emb.numpy()}
    for heads, concat in ((1, False), (2, True)):
        torch.manual_seed(3)
        layer = gl_mod.GraphLayer(W, D, heads=heads, concat=concat)
        with torch.no_grad():
            layer.att_em_i.normal_(0, 0.5)
            layer.att_em_j.normal_(0, 0.5)
            layer.bias.normal_(0, 0.3)
        tag = f"h{heads}"
        for k, v in layer.state_dict().items():
            rec[f"{tag}/sd/{k}"] = v.numpy()
        emb_h = emb.clone().requires_grad_(True)
        out, (ei2, alpha) = layer(x, ei, emb_h, return_attention_weights=True)
        gout = torch.rand(out.shape, generator=g)
        out.backward(gout)
        rec[f"{tag}/out"] = out.detach().numpy()
        rec[f"{tag}/alpha"] = alpha.detach().numpy()
        rec[f"{tag}/edge_index_out"] = ei2.numpy()
        rec[f"{tag}/gout"] = gout.numpy()
        rec[f"{tag}/grad_embedding"] = emb_h.grad.numpy()
        for k, p in layer.named_parameters():
            rec[f"{tag}/grad/{k}"] = p.grad.numpy()
        # oracle cross-check
        o_out, (o_ei, o_alpha) = go.graph_layer_forward(
            x, ei, emb, layer.lin.weight.detach(), layer.att_i.detach(), layer.att_j.detach(),
            layer.att_em_i.detach(), layer.att_em_j.detach(), layer.bias.detach(),
            heads=heads, concat=concat)
        assert torch.equal(o_ei, ei2)
        e1 = (o_out - out.detach()).abs().max().item()
        e2 = (o_alpha - alpha.detach()).abs().max().item()
        print(f"graphlayer {tag}: oracle-vs-reference out {e1:.2e} alpha {e2:.2e}")
        assert e1 < 1e-6 and e2 < 1e-6
    np.savez_compressed(os.path.join(OUT, "graph_layer_general.npz"), **rec)


def run_scoring_case(ev_mod):
    """evaluate.get_full_err_scores on forecast-like data (fp32 values in nested lists,
    as test.py:73-75 hands them over)."""
    rng = np.random.default_rng(11)
    for name, T, N in (("score_small", 257, 9), ("score_even", 600, 5)):
        gt = rng.random((T, N)).astype(np.float32)
        pred = (gt + rng.normal(0, 0.05, (T, N)) * (1 + 5 * (rng.random((T, N)) > 0.97))).astype(np.float32)
        pred[:, 0] = gt[:, 0]                     # a perfect sensor: median = IQR = 0
        labels = (rng.random((T, N)) > 0.9).astype(np.float32)
        res = [pred.tolist(), gt.tolist(), labels.tolist()]
        scores, normals = ev_mod.get_full_err_scores(res, res)
        mine = so.full_err_scores(pred, gt, vectorised=False)
        mine_v = so.full_err_scores(pred, gt, vectorised=True)
        d1 = np.abs(mine - scores).max()
        d2 = np.abs(mine_v - scores).max()
        print(f"{name}: scoring oracle-vs-reference loop {d1:.2e} vec {d2:.2e}")
        assert d1 == 0.0 and d2 < 1e-12
        np.savez_compressed(os.path.join(OUT, name + ".npz"), pred=pred, gt=gt, scores=scores,
                            top1=np.max(scores, axis=0))


def main():
    os.makedirs(OUT, exist_ok=True)
    gdn_mod, gl_mod, ev_mod = pyg_shim.import_reference()
    torch.set_num_threads(1)      # deterministic accumulation order for the fixtures
    for name in CASES:
        run_case(name, gdn_mod)
    run_graph_layer_case(gl_mod)
    run_scoring_case(ev_mod)


if __name__ == "__main__":
    main()
