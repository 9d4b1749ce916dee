"""Host-side plumbing between PyTorch tensors and libgdn_b200.so.

PyTorch is used for device memory (torch.empty), streams and autograd bookkeeping only;
every arithmetic step of the hot path runs in the hand-written sm_100a kernels behind the
C ABI (include/gdn_b200.h).  Nothing here falls back to the CPU or to PyTorch math.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BN, Dims, Dropout, HeadGrads, HeadParams, LayerGrads, LayerParams, check, ptr

DROP_P = 0.2  # models/GDN.py:114


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(
            f"gdn_b200: {name} lives on {t.device}; this implementation has no CPU path "
            "(the reference's `-device cpu` mode is not provided)")


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _blob(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def make_dims(B, N, W, D, K):
    return Dims(int(B), int(N), int(W), int(D), int(K))


# ----------------------------------------------------------------------------- graph
def graph_build(V, topk, use_tensor_cores=-1, want_idx=True, kth=None, margin=0.03, rows=None, out=None):
    """models/GDN.py:143-159 -> (learned_graph [N,K] int64, nbr [N,K+1] int32).

    kth: optional float32 [N] CUDA tensor, in/out: the K-th largest cosine of every row from the previous
    build (fill with -inf for "no hint"); warm-starts the tensor-core engine's admission threshold and
    receives this build's values.  The result never depends on the hint.
    rows = (row0, row1): build only those rows (row0 % 128 == 0, row1 % 128 == 0 or row1 == N) -- the
    row-sharded build of the data-parallel trainer; `out` = (idx, nbr) buffers with at least N rows to write into
    (rows outside the range are left as they are)."""
    _need_cuda(V, "embedding.weight")
    lib = _lib.load()
    Vc = _f32c(V)
    N, D = Vc.shape
    K = int(topk)
    if not 1 <= K <= N:
        raise RuntimeError(f"topk={K} must be in 1..node_num={N} (torch.topk would raise too)")
    if out is not None:
        idx, nbr = out                      # idx may be None: only the neighbour table is wanted
        bad_idx = idx is not None and (idx.dtype != torch.int64 or idx.shape[0] < N or tuple(idx.shape[1:]) != (K,)
                                       or not idx.is_contiguous() or idx.device != Vc.device)
        if bad_idx or nbr.dtype != torch.int32 or nbr.shape[0] < N or tuple(nbr.shape[1:]) != (K + 1,) \
                or not nbr.is_contiguous() or nbr.device != Vc.device:
            raise RuntimeError("out must be contiguous (int64 [>=N, K] or None, int32 [>=N, K+1]) on the embedding's device")
    else:
        idx = torch.empty((N, K), dtype=torch.int64, device=Vc.device) if want_idx else None
        nbr = torch.empty((N, K + 1), dtype=torch.int32, device=Vc.device)
    nb = lib.gdn_graph_build_ws_bytes(N, D, K)
    ws = _blob(nb, Vc.device)
    if kth is not None and (kth.dtype != torch.float32 or kth.device != Vc.device or kth.numel() != N
                            or not kth.is_contiguous()):
        raise RuntimeError("kth must be a contiguous float32 [N] tensor on the embedding's device")
    if rows is not None:
        r0, r1 = int(rows[0]), int(rows[1])
        check(lib.gdn_graph_build_rows(ptr(Vc), N, D, K, r0, r1, ptr(idx), ptr(nbr), ptr(ws), ws.numel(),
                                       int(use_tensor_cores), ptr(kth), float(margin), _stream()), "gdn_graph_build_rows")
    elif kth is None:
        check(lib.gdn_graph_build(ptr(Vc), N, D, K, ptr(idx), ptr(nbr), ptr(ws), ws.numel(), int(use_tensor_cores),
                                  _stream()), "gdn_graph_build")
    else:
        check(lib.gdn_graph_build_warm(ptr(Vc), N, D, K, ptr(idx), ptr(nbr), ptr(ws), ws.numel(),
                                       int(use_tensor_cores), ptr(kth), float(margin), _stream()),
              "gdn_graph_build_warm")
    return idx, nbr


def idx_from_nbr(nbr):
    """learned_graph [N, K] int64 from the neighbour table alone.  The table lists the row's non-self top-k entries
    in order, then the row itself; when the row was in its own top-k (practically always: cos = 1) the last slot holds
    -2 - (its position there), so the top-k order is recoverable.  Index plumbing (a gather over N*K ints) for the
    data-parallel graph exchange: ranks all-gather the int32 table only and rebuild the int64 one locally."""
    N, Kp = nbr.shape
    K = Kp - 1
    last = nbr[:, K].long()
    pos = torch.where(last <= -2, -2 - last, torch.full_like(last, K))          # K: self not in the top-k / at the end
    j = torch.arange(K, device=nbr.device).unsqueeze(0)
    p = pos.unsqueeze(1)
    col = torch.where(j < p, j, torch.where(j == p, torch.full_like(j, K - 1), j - 1))
    return torch.gather(nbr[:, :K].long(), 1, col.expand(N, K) if col.shape[0] == 1 else col)


# ----------------------------------------------------------------------------- GraphLayer, shared graph
def _layer_params(lin_w, att_i, att_j, att_em_i, att_em_j, bias):
    return LayerParams(lin_w.data_ptr(), att_i.data_ptr(), att_j.data_ptr(), att_em_i.data_ptr(),
                       att_em_j.data_ptr(), bias.data_ptr() if bias is not None else None)


class GraphLayerBatchedFn(torch.autograd.Function):
    """GraphLayer (heads=1, concat=False) over the window-shared top-k graph:
    x [B,N,W], V [N,D], nbr [N,K+1] -> out [B*N, D]  (SURVEY.md section 8 rows a3/a4)."""

    @staticmethod
    def forward(ctx, x, V, nbr, lin_w, att_i, att_j, att_em_i, att_em_j, bias, want_alpha):
        lib = _lib.load()
        _need_cuda(x, "x")
        xc, Vc = _f32c(x), _f32c(V)
        B, N, W = xc.shape
        D = Vc.shape[1]
        K = nbr.shape[1] - 1
        dims = make_dims(B, N, W, D, K)
        tens = [_f32c(t) for t in (lin_w, att_i, att_j, att_em_i, att_em_j)]
        bias_c = _f32c(bias) if bias is not None else torch.zeros(D, dtype=torch.float32, device=xc.device)
        lp = _layer_params(*tens, bias_c)
        nb_ctx = lib.gdn_graphlayer_ctx_bytes(C.byref(dims))
        if nb_ctx == 0:
            check(-1, "gdn_graphlayer_ctx_bytes")
        blob = _blob(nb_ctx, xc.device)
        out = torch.empty((B * N, D), dtype=torch.float32, device=xc.device)
        alpha = torch.empty((B * N, K + 1), dtype=torch.float32, device=xc.device) if want_alpha else None
        check(lib.gdn_graphlayer_fwd(C.byref(dims), ptr(xc), ptr(Vc), ptr(nbr), C.byref(lp), ptr(out), ptr(alpha),
                                     ptr(blob), None, 0, _stream()), "gdn_graphlayer_fwd")
        ctx.dims = (B, N, W, D, K)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(Vc, nbr, blob, *tens, bias_c)
        ctx.mark_non_differentiable(*([alpha] if alpha is not None else []))
        return (out, alpha) if want_alpha else (out, None)

    @staticmethod
    def backward(ctx, g_out, _g_alpha):
        lib = _lib.load()
        Vc, nbr, blob, lin_w, att_i, att_j, att_em_i, att_em_j, bias_c = ctx.saved_tensors
        B, N, W, D, K = ctx.dims
        dims = make_dims(B, N, W, D, K)
        dev = Vc.device
        g_out = _f32c(g_out)
        lp = _layer_params(lin_w, att_i, att_j, att_em_i, att_em_j, bias_c)
        g_lin = torch.empty_like(lin_w)
        g_ai, g_aj = torch.empty_like(att_i), torch.empty_like(att_j)
        g_aei, g_aej = torch.empty_like(att_em_i), torch.empty_like(att_em_j)
        g_bias = torch.empty(D, dtype=torch.float32, device=dev)
        g_V = torch.empty_like(Vc)
        lg = LayerGrads(g_lin.data_ptr(), g_ai.data_ptr(), g_aj.data_ptr(), g_aei.data_ptr(), g_aej.data_ptr(),
                        g_bias.data_ptr(), g_V.data_ptr())
        ws = _blob(lib.gdn_graphlayer_ws_bytes(C.byref(dims)), dev)
        check(lib.gdn_graphlayer_bwd(C.byref(dims), ptr(g_out), ptr(Vc), ptr(nbr), C.byref(lp), ptr(blob),
                                     C.byref(lg), ptr(ws), ws.numel(), _stream()), "gdn_graphlayer_bwd")
        return (None, g_V, None, g_lin, g_ai, g_aj, g_aei, g_aej, g_bias if ctx.has_bias else None, None)


# ----------------------------------------------------------------------------- fused GDN
_dropout_counter = None


def set_dropout_counter(counter):
    """Device-side int64 counter added to the Philox offset when the dropout kernel RUNS (not when it is
    launched): a captured CUDA graph that bumps the counter draws a fresh mask on every replay."""
    global _dropout_counter
    if counter is not None and (counter.dtype != torch.int64 or not counter.is_cuda):
        raise RuntimeError("dropout counter must be a CUDA int64 tensor")
    _dropout_counter = counter


class _DropState:
    """Philox (seed, offset) bookkeeping: the seed follows torch.manual_seed, the offset
    advances by the number of 4-wide counters a forward consumes."""
    seed = None
    offset = 0

    @classmethod
    def next(cls, n_elems):
        seed = torch.cuda.initial_seed() if torch.cuda.is_available() else torch.initial_seed()
        if seed != cls.seed:
            cls.seed, cls.offset = seed, 0
        off = cls.offset
        cls.offset += (n_elems + 3) // 4
        return seed & 0xFFFFFFFFFFFFFFFF, off


class BatchNormSync:
    """SyncBN for the fused path: the library calls back at the four points where a BatchNorm needs
    batch-wide sums; the callback all-reduces that small buffer of doubles (a slice of the call's workspace blob) over
    the process group on the current stream.  Rows per rank must be equal."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.ws = None
        self.error = None
        self._cb = _lib.SYNC_FN(self._allreduce)           # keep the thunk alive as long as this object
        self.struct = _lib.Sync(self.world, self._cb, None)

    def _allreduce(self, buf, count, _user, _stream):
        try:
            ws = self.ws
            off = int(buf) - ws.data_ptr()
            if off < 0 or off + 8 * count > ws.numel() or off % 8:
                raise RuntimeError("SyncBN buffer lies outside the workspace")
            self.dist.all_reduce(ws[off:off + 8 * count].view(torch.float64), group=self.group)
            return 0
        except BaseException as e:                          # never let an exception cross the C frame
            self.error = e
            return 1

    def pointer(self, ws):
        self.ws, self.error = ws, None
        return C.byref(self.struct)

    def check(self):
        self.ws = None
        if self.error is not None:
            e, self.error = self.error, None
            raise e


class FusedGDNFn(torch.autograd.Function):
    """Whole GDN forward/backward for out_layer_num == 1 (models/GDN.py:122-187)."""

    @staticmethod
    def forward(ctx, x, V, nbr, lin_w, att_i, att_j, att_em_i, att_em_j, bias,
                bn1_w, bn1_b, bn2_w, bn2_b, out_w, out_b, bn1_buf, bn2_buf, training, drop_mask, drop_p, sync=None):
        lib = _lib.load()
        _need_cuda(x, "data")
        xc, Vc = _f32c(x), _f32c(V)
        B, N, W = xc.shape
        D = Vc.shape[1]
        K = nbr.shape[1] - 1
        dims = make_dims(B, N, W, D, K)
        dev = xc.device
        lt = [_f32c(t) for t in (lin_w, att_i, att_j, att_em_i, att_em_j, bias)]
        ht = [_f32c(t) for t in (bn1_w, bn1_b, bn2_w, bn2_b, out_w, out_b)]
        lp = _layer_params(*lt)
        rm1, rv1, nbt1 = bn1_buf
        rm2, rv2, nbt2 = bn2_buf
        for buf in (rm1, rv1, rm2, rv2):
            if buf.dtype != torch.float32 or not buf.is_contiguous() or buf.device != dev:
                raise RuntimeError("BatchNorm running statistics must be contiguous float32 tensors on the batch's device")
        for buf in (nbt1, nbt2):
            if buf is not None and (buf.dtype != torch.int64 or buf.device != dev):
                raise RuntimeError("num_batches_tracked must be an int64 tensor on the batch's device")
        hp = HeadParams(BN(ht[0].data_ptr(), ht[1].data_ptr(), rm1.data_ptr(), rv1.data_ptr(),
                           nbt1.data_ptr() if nbt1 is not None else None),
                        BN(ht[2].data_ptr(), ht[3].data_ptr(), rm2.data_ptr(), rv2.data_ptr(),
                           nbt2.data_ptr() if nbt2 is not None else None),
                        ht[4].data_ptr(), ht[5].data_ptr())
        mask_c = None
        if training and drop_mask is not None:
            mask_c = _f32c(drop_mask)
            if tuple(mask_c.shape) != (B, N, D):
                raise RuntimeError(f"dropout mask must be [B,N,D]={B, N, D}, got {tuple(mask_c.shape)}")
        seed, offset = (0, 0)
        if training and drop_p > 0 and mask_c is None:
            seed, offset = _DropState.next(B * N * D)
        ctr = _dropout_counter if (_dropout_counter is not None and _dropout_counter.device == dev) else None
        dp = Dropout(mask_c.data_ptr() if mask_c is not None else None, seed, offset, float(drop_p),
                     ctr.data_ptr() if ctr is not None else None)
        nb_ctx = lib.gdn_fused_ctx_bytes(C.byref(dims))
        if nb_ctx == 0:
            check(-1, "gdn_fused_ctx_bytes")
        blob = _blob(nb_ctx, dev)
        ws = _blob(lib.gdn_fused_ws_bytes(C.byref(dims)), dev)
        pred = torch.empty((B, N), dtype=torch.float32, device=dev)
        use_sync = sync is not None and training and sync.world > 1
        rc = lib.gdn_fused_fwd_sync(C.byref(dims), ptr(xc), ptr(Vc), ptr(nbr), C.byref(lp), C.byref(hp), C.byref(dp),
                                    1 if training else 0, ptr(pred), ptr(blob), ptr(ws), ws.numel(),
                                    sync.pointer(ws) if use_sync else None, _stream())
        if use_sync:
            sync.check()
        check(rc, "gdn_fused_fwd")
        ctx.sync = sync if use_sync else None
        ctx.dims = (B, N, W, D, K)
        ctx.training = bool(training)
        ctx.drop = (seed, offset, float(drop_p))
        ctx.bufs = (rm1, rv1, rm2, rv2)
        ctx.save_for_backward(Vc, nbr, blob, *lt, *ht)
        ctx.mark_non_differentiable(blob)
        return pred, blob

    @staticmethod
    def backward(ctx, g_pred, _g_blob):
        if not ctx.training:
            raise NotImplementedError(
                "gdn_b200: backward through an eval-mode forward is not implemented "
                "(the reference only differentiates in training mode, train.py:61-72)")
        lib = _lib.load()
        saved = ctx.saved_tensors
        Vc, nbr, blob = saved[0], saved[1], saved[2]
        lt, ht = saved[3:9], saved[9:15]
        B, N, W, D, K = ctx.dims
        dims = make_dims(B, N, W, D, K)
        dev = Vc.device
        g_pred = _f32c(g_pred)
        lp = _layer_params(*lt)
        rm1, rv1, rm2, rv2 = ctx.bufs
        hp = HeadParams(BN(ht[0].data_ptr(), ht[1].data_ptr(), rm1.data_ptr(), rv1.data_ptr(), None),
                        BN(ht[2].data_ptr(), ht[3].data_ptr(), rm2.data_ptr(), rv2.data_ptr(), None),
                        ht[4].data_ptr(), ht[5].data_ptr())
        seed, offset, p = ctx.drop
        dp = Dropout(None, seed, offset, p, None)
        g_layer = [torch.empty_like(t) for t in lt]
        g_V = torch.empty_like(Vc)
        g_head = [torch.empty_like(t) for t in ht]
        lg = LayerGrads(*[t.data_ptr() for t in g_layer], g_V.data_ptr())
        hg = HeadGrads(*[t.data_ptr() for t in g_head])
        ws = _blob(lib.gdn_fused_ws_bytes(C.byref(dims)), dev)
        sync = ctx.sync
        rc = lib.gdn_fused_bwd_sync(C.byref(dims), ptr(g_pred), ptr(Vc), ptr(nbr), C.byref(lp), C.byref(hp),
                                    C.byref(dp), ptr(blob), C.byref(lg), C.byref(hg), ptr(ws), ws.numel(),
                                    sync.pointer(ws) if sync is not None else None, _stream())
        if sync is not None:
            sync.check()
        check(rc, "gdn_fused_bwd")
        return (None, g_V, None, *g_layer, *g_head, None, None, None, None, None, None)


def ctx_alpha(blob, nbr, B, N, W, D, K):
    """Attention weights [B*N, K+1] of the forward whose saved state is `blob`."""
    lib = _lib.load()
    dims = make_dims(B, N, W, D, K)
    alpha = torch.empty((B * N, K + 1), dtype=torch.float32, device=blob.device)
    check(lib.gdn_ctx_alpha(C.byref(dims), ptr(nbr), ptr(blob), ptr(alpha), _stream()), "gdn_ctx_alpha")
    return alpha


def reference_edge_layout(nbr, alpha_ell, B):
    """Re-order slot-major attention weights into the reference's edge order
    (models/graph_layer.py:61-63: all non-self edges in (b, i, k) order, then the B*N self
    loops) -> (edge_index [2,E] int64, alpha [E,1,1])."""
    N, Kp = nbr.shape
    dev = nbr.device
    rows = torch.arange(N, device=dev).unsqueeze(1).expand(N, Kp)
    valid = nbr >= 0
    is_self = valid & (nbr.long() == rows)
    # the self entry is the LAST valid slot; an earlier equal entry cannot exist (it was removed)
    nonself = valid & ~is_self
    offs = (torch.arange(B, device=dev) * N).view(B, 1, 1)
    src = (nbr.long().unsqueeze(0) + offs)
    dst = (rows.unsqueeze(0) + offs)
    ns = nonself.unsqueeze(0).expand(B, N, Kp)
    sf = is_self.unsqueeze(0).expand(B, N, Kp)
    a = alpha_ell.view(B, N, Kp)
    edge_index = torch.stack([torch.cat([src[ns], src[sf]]), torch.cat([dst[ns], dst[sf]])])
    alpha = torch.cat([a[ns], a[sf]]).view(-1, 1, 1)
    return edge_index, alpha


# ----------------------------------------------------------------------------- general CSR GraphLayer
class GraphLayerCSRFn(torch.autograd.Function):
    """Attention message pass of GraphLayer on an arbitrary (self-loop-fixed) edge list.
    Returns out_h [n,H,D] and alpha [E,H] in the given edge order."""

    @staticmethod
    def forward(ctx, x, emb, lin_w, att_i, att_j, att_em_i, att_em_j, edge_index, heads, slope):
        lib = _lib.load()
        _need_cuda(x, "x")
        xc, embc = _f32c(x), _f32c(emb)
        n, W = xc.shape
        H = int(heads)
        D = lin_w.shape[0] // H
        dev = xc.device
        src, dst = edge_index[0], edge_index[1]
        E = int(src.numel())
        order = torch.sort(dst, stable=True)[1]
        col = src[order].to(torch.int32).contiguous()
        counts = torch.bincount(dst, minlength=n)
        rowptr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        ps = [_f32c(t) for t in (lin_w, att_i, att_j, att_em_i, att_em_j)]
        xl = torch.empty((n, H * D), dtype=torch.float32, device=dev)
        s_i = torch.empty((n, H), dtype=torch.float32, device=dev)
        s_j = torch.empty((n, H), dtype=torch.float32, device=dev)
        out_h = torch.empty((n, H, D), dtype=torch.float32, device=dev)
        alpha_csr = torch.empty((max(E, 1), H), dtype=torch.float32, device=dev)
        check(lib.gdn_csr_fwd(n, W, D, H, E, ptr(rowptr), ptr(col), ptr(xc), ptr(embc), *[ptr(t) for t in ps],
                              ptr(xl), ptr(s_i), ptr(s_j), ptr(out_h), ptr(alpha_csr), float(slope), _stream()),
              "gdn_csr_fwd")
        alpha = torch.empty((E, H), dtype=torch.float32, device=dev)
        alpha[order] = alpha_csr[:E]
        ctx.shape = (n, W, D, H, E, float(slope))
        ctx.save_for_backward(xc, embc, *ps, rowptr, col, xl, s_i, s_j, alpha_csr)
        ctx.mark_non_differentiable(alpha)
        return out_h, alpha

    @staticmethod
    def backward(ctx, g_out_h, _g_alpha):
        lib = _lib.load()
        xc, embc, lin_w, att_i, att_j, att_em_i, att_em_j, rowptr, col, xl, s_i, s_j, alpha_csr = ctx.saved_tensors
        n, W, D, H, E, slope = ctx.shape
        dev = xc.device
        g_out_h = _f32c(g_out_h)
        g_xl = torch.empty((n, H * D), dtype=torch.float32, device=dev)
        g_si = torch.empty((n, H), dtype=torch.float32, device=dev)
        g_sj = torch.empty((n, H), dtype=torch.float32, device=dev)
        check(lib.gdn_csr_bwd(n, W, D, H, E, ptr(rowptr), ptr(col), ptr(xc), ptr(embc), ptr(lin_w), ptr(att_i),
                              ptr(att_j), ptr(att_em_i), ptr(att_em_j), ptr(xl), ptr(s_i), ptr(s_j), ptr(alpha_csr),
                              ptr(g_out_h), ptr(g_xl), ptr(g_si), ptr(g_sj), slope, _stream()), "gdn_csr_bwd")
        # dense contractions around the sparse part (plain matmuls)
        ai, aj = att_i.view(1, H, D), att_j.view(1, H, D)
        g_xl3 = g_xl.view(n, H, D) + g_si.unsqueeze(-1) * ai + g_sj.unsqueeze(-1) * aj
        g_flat = g_xl3.reshape(n, H * D)
        g_lin = g_flat.t().mm(xc)
        g_x = g_flat.mm(lin_w)
        xl3 = xl.view(n, H, D)
        g_ai = (g_si.unsqueeze(-1) * xl3).sum(0, keepdim=True)
        g_aj = (g_sj.unsqueeze(-1) * xl3).sum(0, keepdim=True)
        g_aei = g_si.t().mm(embc).view(1, H, D)
        g_aej = g_sj.t().mm(embc).view(1, H, D)
        g_emb = g_si.mm(att_em_i.view(H, D)) + g_sj.mm(att_em_j.view(H, D))
        return (g_x, g_emb, g_lin, g_ai.view_as(att_i), g_aj.view_as(att_j), g_aei.view_as(att_em_i),
                g_aej.view_as(att_em_j), None, None, None)


# ----------------------------------------------------------------------------- scoring
def score(pred, gt, want_scores=True, want_top1=True, want_stats=False):
    """evaluate.py:48-68 for every sensor at once.  pred, gt: [T, N] float32 CUDA tensors.
    Returns (scores [N,T] f64 | None, top1 [T] f64 | None, stats [N,2] f64 | None)."""
    lib = _lib.load()
    _need_cuda(pred, "pred")
    p, g = _f32c(pred), _f32c(gt)
    if p.dim() != 2 or p.shape != g.shape:
        raise RuntimeError(f"score: pred/gt must both be [T, N], got {tuple(p.shape)} / {tuple(g.shape)}")
    T, N = p.shape
    dev = p.device
    scores = torch.empty((N, T), dtype=torch.float64, device=dev) if want_scores else None
    top1 = torch.empty((T,), dtype=torch.float64, device=dev) if want_top1 else None
    stats = torch.empty((N, 2), dtype=torch.float64, device=dev) if want_stats else None
    ws = _blob(lib.gdn_score_ws_bytes(T, N), dev)
    check(lib.gdn_score(ptr(p), ptr(g), T, N, ptr(scores), ptr(top1), ptr(stats), ptr(ws), ws.numel(), _stream()),
          "gdn_score")
    return scores, top1, stats
