"""GPU parity of GraphLayer at its own module boundary: the window-shared fast path and the
arbitrary-edge-list path, forward and backward, against the oracle / reference vectors."""
import pytest
import torch

from golden_util import load, normwise, state_dict
from oracle import gdn_oracle as go

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _layer(W, D, heads, concat, sd):
    from gdn_b200.models.graph_layer import GraphLayer
    layer = GraphLayer(W, D, heads=heads, concat=concat)
    layer.load_state_dict(sd)
    return layer.cuda()


@pytest.mark.parametrize("heads", [1, 2])
def test_general_edge_list_matches_reference(heads):
    rec = load("graph_layer_general")
    tag = f"h{heads}"
    sd = state_dict(rec, prefix=tag + "/sd/")
    W, D = rec["x"].shape[1], rec["embedding"].shape[1]
    layer = _layer(W, D, heads, heads == 2, sd)
    x = torch.from_numpy(rec["x"]).cuda()
    emb = torch.from_numpy(rec["embedding"]).cuda().requires_grad_(True)
    ei = torch.from_numpy(rec["edge_index"]).cuda()
    out, (ei2, alpha) = layer(x, ei, emb, return_attention_weights=True)
    assert torch.equal(ei2.cpu(), torch.from_numpy(rec[tag + "/edge_index_out"]))
    assert normwise(out.detach().cpu(), rec[tag + "/out"]) < TOL
    assert normwise(alpha.cpu(), rec[tag + "/alpha"]) < TOL
    out.backward(torch.from_numpy(rec[tag + "/gout"]).cuda())
    assert normwise(emb.grad.cpu(), rec[tag + "/grad_embedding"]) < TOL
    for k, p in layer.named_parameters():
        assert normwise(p.grad.cpu(), rec[f"{tag}/grad/{k}"]) < TOL, k


@pytest.mark.parametrize("shape", [(27, 5, 64, 5, 32), (51, 5, 64, 15, 40), (96, 16, 128, 12, 7), (33, 10, 32, 7, 5),
                                   (300, 16, 128, 32, 64), (4096, 16, 128, 32, 8), (16384, 16, 128, 64, 1)],
                         ids=["C1", "C2-B40", "w16", "w10", "n300", "C4-B8", "C5-B1"])
def test_batched_layer_matches_oracle(shape):
    """forward_batched (the hot path) vs the edge-list oracle, incl. all parameter gradients and
    the embedding gradient; the oracle runs in float64 for the gradient reference."""
    from gdn_b200 import ops
    N, W, D, K, B = shape
    sd = go.init_state(N, D, W, seed=11, stressed=True)
    G = "gnn_layers.0.gnn."
    lsd = {k[len(G):]: v for k, v in sd.items() if k.startswith(G)}
    layer = _layer(W, D, 1, False, lsd)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(B, N, W, generator=g)
    gout = torch.randn(B * N, D, generator=g)
    V = sd["embedding.weight"].clone()
    Vc = V.cuda().requires_grad_(True)
    idx, nbr = ops.graph_build(Vc.detach(), K, use_tensor_cores=0 if N <= 300 else -1)
    out, alpha_ell = layer.forward_batched(x.cuda(), nbr, Vc, return_attention_weights=True)
    out.backward(gout.cuda())
    # oracle in float64 on the same graph (at the BASELINE scale shapes near-tied cosines may rank differently
    # in the two builds -- tests/test_gpu_parity.py checks the graph itself by the top-k protocol -- so the
    # layer is compared on the graph it was given)
    if N <= 300:
        idx_o, _ = go.learned_graph(V, K)
        assert torch.equal(idx.cpu(), idx_o)
    else:
        idx_o = idx.cpu()
    edges = go.batch_edges(idx_o, B)
    P = {k: v.double().requires_grad_(True) for k, v in lsd.items()}
    V64 = V.double().requires_grad_(True)
    o64, (ei, a64) = go.graph_layer_forward(x.view(-1, W).double(), edges, V64.repeat(B, 1), P["lin.weight"],
                                            P["att_i"], P["att_j"], P["att_em_i"], P["att_em_j"], P["bias"])
    o64.backward(gout.double())
    print(f"\nGraphLayer N={N} B={B}: out err {normwise(out.detach().cpu(), o64.detach()):.3e}, "
          f"g_V err {normwise(Vc.grad.cpu(), V64.grad):.3e}, "
          + ", ".join(f"{k} {normwise(p.grad.cpu(), P[k].grad):.3e}" for k, p in layer.named_parameters()))
    assert normwise(out.detach().cpu(), o64.detach()) < TOL
    ei2, alpha = ops.reference_edge_layout(nbr, alpha_ell, B)
    assert torch.equal(ei2.cpu(), ei)
    assert normwise(alpha.cpu(), a64.detach()) < TOL
    assert normwise(Vc.grad.cpu(), V64.grad) < TOL
    for k, p in layer.named_parameters():
        assert normwise(p.grad.cpu(), P[k].grad) < TOL, k


def test_gnnlayer_generic_forward_matches_oracle():
    """GNNLayer.forward(x, edge_index, embedding) -- the reference's own call at models/GDN.py:166."""
    from gdn_b200.models.GDN import GNNLayer
    N, W, D, K, B = 27, 5, 64, 5, 6
    sd = go.init_state(N, D, W, seed=3, stressed=True)
    G = "gnn_layers.0."
    layer = GNNLayer(W, D, inter_dim=2 * D, heads=1)
    layer.load_state_dict({k[len(G):]: v for k, v in sd.items() if k.startswith(G)})
    layer = layer.cuda().eval()
    x = torch.rand(B * N, W)
    V = sd["embedding.weight"]
    idx, _ = go.learned_graph(V, K)
    edges = go.batch_edges(idx, B)
    with torch.no_grad():
        out = layer(x.cuda(), edges.cuda(), embedding=V.repeat(B, 1).cuda(), node_num=B * N)
    z, (ei, alpha) = go.graph_layer_forward(x, edges, V.repeat(B, 1), sd[G + "gnn.lin.weight"], sd[G + "gnn.att_i"],
                                            sd[G + "gnn.att_j"], sd[G + "gnn.att_em_i"], sd[G + "gnn.att_em_j"],
                                            sd[G + "gnn.bias"])
    ref = torch.relu(torch.nn.functional.batch_norm(z, sd[G + "bn.running_mean"], sd[G + "bn.running_var"],
                                                    sd[G + "bn.weight"], sd[G + "bn.bias"], False, 0.1, 1e-5))
    assert normwise(out.cpu(), ref) < TOL
    assert torch.equal(layer.edge_index_1.cpu(), ei)
    assert normwise(layer.att_weight_1.cpu(), alpha) < TOL
