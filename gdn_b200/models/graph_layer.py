"""GraphLayer -- same constructor, parameters, state_dict keys and forward signature as the
reference's models/graph_layer.py:11-124, without torch-geometric.

Two execution paths, both hand-written CUDA behind the C ABI:
  * forward(x, edge_index, embedding, ...)   arbitrary edge list (the reference signature):
    CSR-by-target attention kernels (csrc/csr.cu), any `heads`;
  * forward_batched(x[B,N,W], nbr, V)        the window-shared top-k graph GDN builds:
    the fused lane<->window attention kernels (csrc/attention.cu + csrc/dwide.cu).
"""
import math

import torch
from torch.nn import Linear, Parameter

from gdn_b200 import ops


def _glorot_(t):
    """PyG 1.5.0 `inits.glorot`: U(-a, a) with a = sqrt(6 / (fan_in + fan_out)) over the last two dims."""
    if t is None:
        return
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-bound, bound)


_ATT_NAMES = ("att_i", "att_j", "att_em_i", "att_em_j")       # state_dict keys of models/graph_layer.py:31-34


class GraphLayer(torch.nn.Module):
    """Constructor arguments as models/graph_layer.py:12-13 (`inter_dim` is accepted and unused there too)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True,
                 negative_slope=0.2, dropout=0, bias=True, inter_dim=-1, **kwargs):
        super().__init__()
        if kwargs.get("aggr", "add") != "add" or kwargs.get("flow", "source_to_target") != "source_to_target":
            raise NotImplementedError("GraphLayer: only aggr='add', flow='source_to_target' (the reference's use)")
        self.aggr, self.flow, self.node_dim = "add", "source_to_target", 0
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.__alpha__ = None
        # parameter creation order = the reference's, so that a same-seed construction draws the same numbers
        self.lin = Linear(in_channels, heads * out_channels, bias=False)
        for name in _ATT_NAMES:
            setattr(self, name, Parameter(torch.empty(1, heads, out_channels)))
        if bias:
            self.bias = Parameter(torch.empty(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        """models/graph_layer.py:42-49: glorot for lin / att_i / att_j (in this order), zeros for the rest."""
        for t in (self.lin.weight, self.att_i, self.att_j):
            _glorot_(t)
        with torch.no_grad():
            for t in (self.att_em_i, self.att_em_j, self.bias):
                if t is not None:
                    t.zero_()

    # ------------------------------------------------------------------ reference signature
    def forward(self, x, edge_index, embedding, return_attention_weights=False):
        """x [n, W]; edge_index [2, E] long (row 0 source, row 1 target); embedding [n, D]."""
        if not torch.is_tensor(x):
            raise NotImplementedError("GraphLayer: bipartite (x_src, x_dst) input is not provided")
        if embedding is None:
            raise RuntimeError("GraphLayer: embedding is required (the reference fails without it too)")
        if self.dropout != 0 and self.training:
            raise NotImplementedError("GraphLayer: attention dropout > 0 is not provided (GDN uses 0)")
        n = x.size(0)
        edge_index = edge_index.long()
        keep = edge_index[0] != edge_index[1]                                   # remove_self_loops
        loops = torch.arange(n, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
        edge_index = torch.cat([edge_index[:, keep], loops], dim=1)            # add_self_loops
        out_h, alpha = ops.GraphLayerCSRFn.apply(
            x, embedding, self.lin.weight, self.att_i, self.att_j, self.att_em_i, self.att_em_j,
            edge_index, self.heads, self.negative_slope)
        if self.concat:
            out = out_h.reshape(-1, self.heads * self.out_channels)
        else:
            out = out_h.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        if return_attention_weights:
            return out, (edge_index, alpha.view(-1, self.heads, 1))
        return out

    # ------------------------------------------------------------------ window-shared graph
    def forward_batched(self, x, nbr, V, return_attention_weights=False):
        """x [B, N, W]; nbr [N, K+1] int32 from ops.graph_build; V [N, D] -> out [B*N, D]
        (and optionally the slot-major attention weights [B*N, K+1])."""
        if self.heads != 1 or self.concat:
            raise NotImplementedError("forward_batched: heads=1, concat=False (as GDN builds it)")
        out, alpha = ops.GraphLayerBatchedFn.apply(
            x, V, nbr, self.lin.weight, self.att_i, self.att_j, self.att_em_i, self.att_em_j, self.bias,
            bool(return_attention_weights))
        return (out, alpha) if return_attention_weights else out

    def __repr__(self):
        return f"{type(self).__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"
