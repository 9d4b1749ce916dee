"""gdn_b200 -- B200-native (sm_100a) implementation of the GDN forward/backward hot path.

Public surface mirrors the reference (SchlomoFeng/GDN):
    gdn_b200.models.GDN.GDN, GNNLayer, OutLayer, get_batch_edge_index
    gdn_b200.models.graph_layer.GraphLayer
    gdn_b200.evaluate.get_err_scores / get_full_err_scores
All arithmetic runs in hand-written CUDA kernels behind the C ABI in include/gdn_b200.h
(libgdn_b200.so, built by `python -m gdn_b200.build`).  There is no CPU fallback.
"""
__version__ = "0.1.0"
