"""Stand-in for the six torch-geometric==1.5.0 symbols the reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch-geometric 1.5.0 and
torch-scatter are pinned by the reference (install.sh:1-5, README.md:8-11) but are
not installable offline, so the behaviour of the symbols used at
models/graph_layer.py:4-7 and models/GDN.py:8 is restated here from the library's
documented semantics:

* ``MessagePassing(aggr='add', flow='source_to_target', node_dim=0)``:
  ``propagate(edge_index, **kw)`` gathers ``<name>_j = kw[name][edge_index[0]]``
  (source) and ``<name>_i = kw[name][edge_index[1]]`` (target) for every ``*_i`` /
  ``*_j`` argument of ``message``; passes ``edge_index_i``, ``edge_index_j``,
  ``size_i``, ``size_j`` and the remaining keyword arguments through; scatter-adds the
  messages by ``edge_index[1]`` into ``size_i`` rows; ``update`` is the identity.
* ``remove_self_loops``: order-preserving mask ``row != col``.
* ``add_self_loops``: appends ``(r, r)`` for ``r < num_nodes`` at the end.
* ``softmax(src, index, num_nodes)``: ``exp(src - segment_max[index])`` divided by
  ``segment_sum[index] + 1e-16``.
* ``glorot(t)``: ``U(-s, s)``, ``s = sqrt(6 / (t.size(-2) + t.size(-1)))``; ``zeros``.

``install()`` registers fake ``torch_geometric`` modules (plus empty
``matplotlib``/``pytz`` stubs the reference imports but never uses on this path) in
``sys.modules`` so the reference's files import unmodified.
"""
import inspect
import math
import sys
import types

import torch


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    if edge_attr is not None:
        edge_attr = edge_attr[keep]
    return edge_index[:, keep], edge_attr


def add_self_loops(edge_index, edge_weight=None, fill_value=1, num_nodes=None):
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    loops = loops.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        extra = edge_weight.new_full((num_nodes,), fill_value)
        edge_weight = torch.cat([edge_weight, extra], dim=0)
    return torch.cat([edge_index, loops], dim=1), edge_weight


def _expand_index(index, like):
    shape = [index.numel()] + [1] * (like.dim() - 1)
    return index.view(shape).expand_as(like)


def softmax(src, index, num_nodes=None):
    if num_nodes is None:
        num_nodes = int(index.max()) + 1
    idx = _expand_index(index, src)
    seg_max = src.new_full((num_nodes,) + tuple(src.shape[1:]), float("-inf"))
    seg_max = seg_max.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - seg_max.gather(0, idx)).exp()
    seg_sum = torch.zeros((num_nodes,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    seg_sum = seg_sum.scatter_add(0, idx, out)
    return out / (seg_sum.gather(0, idx) + 1e-16)


def glorot(tensor):
    if tensor is not None:
        bound = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-bound, bound)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target" and node_dim == 0
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self._msg_args = [p for p in inspect.signature(self.message).parameters]

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        sizes = [None, None] if size is None else list(size)
        call = {}
        for name in self._msg_args:
            if name in ("edge_index_i", "edge_index_j", "size_i", "size_j"):
                continue
            if name.endswith("_i") or name.endswith("_j"):
                side = 1 if name.endswith("_i") else 0
                data = kwargs[name[:-2]]
                if isinstance(data, (tuple, list)):
                    data = data[side]
                if data is None:
                    call[name] = None
                    continue
                if sizes[side] is None:
                    sizes[side] = data.size(0)
                call[name] = data.index_select(0, dst if side == 1 else src)
            else:
                call[name] = kwargs[name]
        sizes[0] = sizes[1] if sizes[0] is None else sizes[0]
        sizes[1] = sizes[0] if sizes[1] is None else sizes[1]
        for name in self._msg_args:
            if name == "edge_index_i":
                call[name] = dst
            elif name == "edge_index_j":
                call[name] = src
            elif name == "size_i":
                call[name] = sizes[1]
            elif name == "size_j":
                call[name] = sizes[0]
        msg = self.message(**call)
        out = torch.zeros((sizes[1],) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        out = out.index_add(0, dst, msg)
        return self.update(out)

    def message(self, x_j):  # pragma: no cover - always overridden
        return x_j

    def update(self, aggr_out):
        return aggr_out


class _Unused(torch.nn.Module):
    def __init__(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("not on the GDN hot path")


def install():
    """Register the stand-in modules; idempotent."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_gdn_shim", False):
        return
    import pandas  # noqa: F401  (pandas probes pytz's version: import it before the stub)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    tg = mod("torch_geometric", _gdn_shim=True, __version__="1.5.0-shim")
    tg.nn = mod("torch_geometric.nn", GCNConv=_Unused, GATConv=_Unused, EdgeConv=_Unused)
    tg.nn.conv = mod("torch_geometric.nn.conv", MessagePassing=MessagePassing)
    tg.nn.inits = mod("torch_geometric.nn.inits", glorot=glorot, zeros=zeros)
    tg.utils = mod("torch_geometric.utils", remove_self_loops=remove_self_loops,
                   add_self_loops=add_self_loops, softmax=softmax)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = mod("matplotlib")
            mpl.pyplot = mod("matplotlib.pyplot")
    if "pytz" not in sys.modules:
        try:
            import pytz  # noqa: F401
        except Exception:
            mod("pytz", utc=None, timezone=lambda *_a, **_k: None)


def import_reference(root="/root/reference"):
    """Import the reference's own modules (build container only; the GPU box has no
    /root/reference).  Returns (GDN_module, graph_layer_module, evaluate_module)."""
    import importlib
    import os

    if not os.path.isdir(root):
        raise FileNotFoundError(root)
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    # site-packages' HuggingFace `datasets` shadows the reference's namespace package
    for name in ("models", "util", "evaluate", "test", "train"):
        m = sys.modules.get(name)
        if m is not None and not str(getattr(m, "__file__", getattr(m, "__path__", ""))).startswith(root):
            del sys.modules[name]
    gdn = importlib.import_module("models.GDN")
    gl = importlib.import_module("models.graph_layer")
    ev = importlib.import_module("evaluate")
    return gdn, gl, ev
