"""CPU: the C-ABI library loads and exports every symbol include/gdn_b200.h declares; the
host-side mirror keeps the reference's module surface (names, shapes, state_dict keys, error
behaviour).  No compute calls: there is no GPU in the CPU test environment."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gdn_b200 import _lib, build
    build.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gdn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gdn_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from gdn_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gdn_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_size_queries(lib):
    from gdn_b200._lib import Dims
    assert lib.gdn_version() == 100
    d = Dims(64, 4096, 16, 128, 32)
    ctx = lib.gdn_fused_ctx_bytes(ctypes.byref(d))
    ws = lib.gdn_fused_ws_bytes(ctypes.byref(d))
    n = 64 * 4096
    assert ctx >= n * 16 * 4 * 2            # xT and A at least
    assert ws >= n * 16 * 4
    assert lib.gdn_graphlayer_ctx_bytes(ctypes.byref(d)) <= ctx
    assert lib.gdn_score_ws_bytes(1000, 10) >= 1000 * 10 * 8
    bad = Dims(4, 10, 40, 64, 3)            # slide_win > 32 is rejected, loudly
    assert lib.gdn_fused_ctx_bytes(ctypes.byref(bad)) == 0
    assert b"slide_win" in lib.gdn_last_error()


def test_argument_errors_do_not_touch_the_gpu(lib):
    from gdn_b200._lib import Dims
    rc = lib.gdn_graph_build(None, 10, 8, 3, None, None, None, 0, 0, None)
    assert rc < 0 and b"NULL" in lib.gdn_last_error()
    rc = lib.gdn_score(None, None, 10, 3, None, None, None, None, 0, None)
    assert rc < 0


def test_module_surface_matches_reference():
    from gdn_b200.models.GDN import GDN, GNNLayer, OutLayer, get_batch_edge_index
    from gdn_b200.models.graph_layer import GraphLayer
    model = GDN([torch.zeros(2, 5, dtype=torch.long)], 27, dim=64, out_layer_inter_dim=128, input_dim=5,
                out_layer_num=1, topk=5)
    keys = list(model.state_dict().keys())
    want = {
        "embedding.weight": (27, 64),
        "bn_outlayer_in.weight": (64,), "bn_outlayer_in.bias": (64,),
        "bn_outlayer_in.running_mean": (64,), "bn_outlayer_in.running_var": (64,),
        "bn_outlayer_in.num_batches_tracked": (),
        "gnn_layers.0.gnn.att_i": (1, 1, 64), "gnn_layers.0.gnn.att_j": (1, 1, 64),
        "gnn_layers.0.gnn.att_em_i": (1, 1, 64), "gnn_layers.0.gnn.att_em_j": (1, 1, 64),
        "gnn_layers.0.gnn.bias": (64,), "gnn_layers.0.gnn.lin.weight": (64, 5),
        "gnn_layers.0.bn.weight": (64,), "gnn_layers.0.bn.bias": (64,),
        "gnn_layers.0.bn.running_mean": (64,), "gnn_layers.0.bn.running_var": (64,),
        "gnn_layers.0.bn.num_batches_tracked": (),
        "out_layer.mlp.0.weight": (1, 64), "out_layer.mlp.0.bias": (1,),
    }
    assert set(keys) == set(want)
    for k, shp in want.items():
        assert tuple(model.state_dict()[k].shape) == shp, k
    for attr in ("embedding", "bn_outlayer_in", "gnn_layers", "out_layer", "dp", "topk", "learned_graph",
                 "edge_index_sets", "cache_edge_index_sets", "init_params"):
        assert hasattr(model, attr), attr
    assert model.dp.p == 0.2 and model.learned_graph is None
    assert repr(model.gnn_layers[0].gnn) == "GraphLayer(5, 64, heads=1)"
    # reference initialisers: zeros for att_em_*/bias, glorot bound for lin, kaiming for the embedding
    gnn = model.gnn_layers[0].gnn
    assert float(gnn.att_em_i.abs().max()) == 0 and float(gnn.bias.abs().max()) == 0
    assert float(gnn.lin.weight.abs().max()) <= (6.0 / (64 + 5)) ** 0.5 + 1e-6
    assert float(model.embedding.weight.abs().max()) <= 1.0 / 8 + 1e-6
    mlp2 = OutLayer(64, 27, 2, inter_num=32)
    assert [type(m).__name__ for m in mlp2.mlp] == ["Linear", "BatchNorm1d", "ReLU", "Linear"]
    e = get_batch_edge_index(torch.tensor([[0, 1], [1, 2]]), 3, 10)
    assert e.tolist() == [[0, 1, 10, 11, 20, 21], [1, 2, 11, 12, 21, 22]]
    assert isinstance(GNNLayer(5, 64, inter_dim=128).gnn, GraphLayer)


def test_reference_checkpoint_round_trip(tmp_path):
    """state_dict written by the reference loads here and vice versa (train.py:94, main.py:120)."""
    from golden_util import load, state_dict
    from gdn_b200.models.GDN import GDN
    rec = load("c1_stress")
    sd = state_dict(rec)
    model = GDN([torch.zeros(2, 5, dtype=torch.long)], 27, dim=64, input_dim=5, topk=5)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    path = tmp_path / "best.pt"
    torch.save(model.state_dict(), path)
    back = torch.load(path)
    assert list(back.keys()) == list(model.state_dict().keys())
    for k, v in sd.items():
        assert torch.equal(back[k], v), k


def test_cpu_device_is_rejected_not_emulated():
    from gdn_b200.models.GDN import GDN
    model = GDN([torch.zeros(2, 5, dtype=torch.long)], 27, dim=64, input_dim=5, topk=5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.zeros(4, 27, 5), None)
    with pytest.raises(NotImplementedError):
        GDN([torch.zeros(2, 5, dtype=torch.long)] * 2, 27)


def test_same_seed_initialisation_matches_reference_order():
    """Constructing under the same torch seed consumes the RNG in the reference's order
    (models/GDN.py:95-119), so weights are identical to a reference model built with that seed."""
    import sys
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present on this box")
    from oracle import pyg_shim
    gdn_mod, _, _ = pyg_shim.import_reference()
    from gdn_b200.models.GDN import GDN
    torch.manual_seed(5)
    ref = gdn_mod.GDN([torch.zeros(2, 5, dtype=torch.long)], 27, dim=64, input_dim=5, topk=5)
    torch.manual_seed(5)
    ours = GDN([torch.zeros(2, 5, dtype=torch.long)], 27, dim=64, input_dim=5, topk=5)
    for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    for name in ("models", "models.GDN", "models.graph_layer"):
        sys.modules.pop(name, None)
