"""Window-sharded data parallelism.

The path shards by window: every window's forward/backward is independent given the
(replicated) parameters, and the learned graph depends only on the embedding, so it is
identical on every rank.  One process per GPU; the only data-path collective is ONE
all-reduce of a flat fp32 gradient buffer per step (embedding + weight gradients, 8.4 MB at
the largest config) over NCCL / NVLink.  BatchNorm statistics stay per rank (DDP semantics
without SyncBN): results equal the reference evaluated on each rank's shard with gradients
averaged.

The host logic below is backend-agnostic (it runs under gloo on CPU in tests/test_dp_cpu.py);
bench.py uses it with the nccl backend.
"""
import os

import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world_size):
    """Contiguous window shard [lo, hi) of rank `rank`; shards differ by at most one window."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(int(global_batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradAllReduce:
    """Flatten the gradients of `params` into one buffer, all-reduce(sum), scale by 1/world,
    and hand the result back as views (p.grad aliases the flat buffer: no copy back)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def __call__(self):
        world = self.world_size()
        if world == 1:
            return None
        grads = []
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads.append(p.grad.reshape(-1))
        flat = torch.cat(grads)                       # one kernel
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.mul_(1.0 / world)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = flat[off:off + n].view_as(p)
            off += n
        self.flat = flat
        return flat


def graph_row_shard(n_rows, rank, world, align=128):
    """Row range of `rank` in the row-sharded graph build (optional exchange step): equal chunks of
    `chunk` rows, `chunk` a multiple of `align` (the tensor-core engine's row block), the last ones possibly short
    or empty.  Returns (row0, row1, chunk); row0 == row1 means "nothing to build"."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    per = -(-n_rows // world)
    chunk = -(-per // align) * align
    r0 = min(n_rows, rank * chunk)
    return r0, min(n_rows, r0 + chunk), chunk


class WindowShardedTrainer:
    """The reference's train step (train.py:68-73: zero_grad, forward, mse, backward, Adam
    step) on this rank's window shard, with the flat gradient all-reduce before the step."""

    # batches up to this many input values are replayed from a CUDA graph (GDN_GRAPH_MAX_ELEMS overrides)
    GRAPH_MAX_ELEMS = int(os.environ.get("GDN_GRAPH_MAX_ELEMS", 1 << 25))

    def __init__(self, model, lr=1e-3, weight_decay=0.0, group=None, fused_adam=None, shard_graph=None, flat_adam=None,
                 nvls=None, cuda_graph=None, sync_bn=False):
        """flat_adam (default on CUDA): parameters, gradients and Adam moments live in flat buffers
        (`gdn_b200.optim.FlatAdam`): autograd accumulates straight into the buffer the all-reduce works on, and the
        1/world scaling is fused into the single Adam kernel -- no `torch.cat`, no `mul_`, no re-pointing of `.grad`
        per step.  flat_adam=False: torch.optim.Adam (fused on CUDA) + FlatGradAllReduce (the CPU / gloo path).
        nvls (default: when the group's GPUs expose NVLink multicast): the all-reduce, the update and the broadcast of
        the new parameters run as ONE kernel over NVSwitch multicast memory (`gdn_b200.optim.NvlsFlatAdam`).
        cuda_graph (default: single process, flat Adam): the BASELINE configs with 27-127 sensors are launch-bound
        (~25 launches for a few hundred microseconds of device work), so after two eager steps on a batch shape the
        whole step -- graph build, forward, MSE, backward, Adam -- is captured once and replayed for every further
        batch of that shape (other shapes, e.g. a short last batch, run eagerly on the same optimiser state)."""
        self.model = model
        self.group = group
        params = list(model.parameters())
        self.flat = None
        if flat_adam is None:
            flat_adam = bool(params) and all(p.is_cuda and p.dtype == torch.float32 for p in params)
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if flat_adam and nvls is None:
            nvls = world > 1 and os.environ.get("GDN_NVLS", "1") != "0"
        self.nvls = None
        if flat_adam and nvls and world > 1:
            from .optim import NvlsFlatAdam
            try:
                self.nvls = NvlsFlatAdam(params, lr=lr, weight_decay=weight_decay, group=group)
            except NvlsFlatAdam.Unavailable as e:          # no multicast on this box / torch build: NCCL path below
                self.nvls_unavailable = str(e)
        if self.nvls is not None:
            self.flat = self.nvls
            self.opt = self.nvls
            self.reduce = None
        elif flat_adam:
            from .optim import FlatAdam
            self.flat = FlatAdam(params, lr=lr, weight_decay=weight_decay)
            self.opt = self.flat
            self.reduce = None
        else:
            kw = {}
            if fused_adam is None:
                fused_adam = all(p.is_cuda for p in params)
            if fused_adam:
                kw["fused"] = True
            self.opt = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, **kw)
            self.reduce = FlatGradAllReduce(params, group)
        # sync_bn: BatchNorm statistics over the global batch (GDN.sync_batchnorm) instead of per-rank (DDP) statistics
        if sync_bn and world > 1 and hasattr(model, "sync_batchnorm"):
            model.sync_batchnorm(group)
        if cuda_graph is None:
            cuda_graph = os.environ.get("GDN_CUDA_GRAPH", "1") != "0"
        self.cuda_graph = bool(cuda_graph) and self.flat is not None and self.nvls is None and world == 1
        self._graphs, self._seen, self._drop_counter, self.graph_launches = {}, {}, None, {}
        # the graph depends on the (replicated) embedding only: every rank builds 1/world of its rows and the
        # neighbour tables are all-gathered (the one exchange step of the forward; off for a single process)
        if shard_graph is None:
            shard_graph = (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
                           and all(p.is_cuda for p in params) and hasattr(model, "shard_graph_build"))
        if shard_graph:
            model.shard_graph_build(dist.get_rank(group), dist.get_world_size(group), group)

    # ---------------------------------------------------------------- CUDA-graph replay of the step (fixed shapes)
    def _flat_body(self, x, y):
        self.flat.zero_grad()
        out = self.model(x, None)
        loss = torch.nn.functional.mse_loss(out, y, reduction="mean")
        loss.backward()
        self.flat.step(grad_scale=1.0)
        return loss

    def _capture_stream(self, dev):
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(device=dev)
        return self._cap_stream

    def _accumulators_follow(self, stream):
        from .graphed import accumulators_follow
        return accumulators_follow(self.flat.params, stream)

    def _capture(self, key, x, y):
        from . import ops
        dev = x.device
        if self._drop_counter is None:
            self._drop_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        xs, ys = torch.empty_like(x), torch.empty_like(y)
        graph = torch.cuda.CUDAGraph()
        was_training = self.model.training
        self.model.train()
        ops.set_dropout_counter(self._drop_counter)          # the Philox offset is read from a counter the graph bumps
        from . import _lib
        n0 = _lib.load().gdn_launch_count()
        try:
            with torch.cuda.graph(graph, stream=self._capture_stream(dev)):
                self._drop_counter.add_(1)
                loss = self._flat_body(xs, ys)
        finally:
            ops.set_dropout_counter(None)
            self.model.train(was_training)
        # detached: the captured autograd graph (and the gradient accumulators bound to the capture stream) is released
        self._graphs[key] = (graph, xs, ys, loss.detach())
        self.graph_launches[key] = int(_lib.load().gdn_launch_count() - n0)    # our kernels per replay

    def _graph_signature(self):
        """Everything a captured step bakes in besides the batch shape: a change drops the graphs (they are
        re-captured after two eager steps), e.g. a learning-rate schedule writing `trainer.flat.lr`."""
        m, f = self.model, self.flat
        return (f.lr, f.betas, f.eps, f.weight_decay, float(getattr(getattr(m, "dp", None), "p", 0.0)),
                getattr(m, "topk", None), getattr(m, "use_tensor_cores", None), getattr(m, "graph_margin", None),
                id(getattr(m, "_bn_sync", None)), id(getattr(m, "_dropout_mask", None)), id(getattr(m, "_graph_shard", None)))

    def _graph_step(self, x, y):
        sig = self._graph_signature()
        if sig != getattr(self, "_graph_sig", None):
            self._graphs, self._seen, self.graph_launches, self._graph_sig = {}, {}, {}, sig
        if getattr(self.model, "_dropout_mask", None) is not None:
            return None                                       # explicit masks (tests) are per call: eager
        key = (tuple(x.shape), tuple(y.shape), x.dtype, y.dtype, str(x.device))
        entry = self._graphs.get(key)
        if entry is None:
            seen = self._seen.get(key, 0)
            if seen < 2 or not self.model.training:           # lazy initialisations happen in eager steps
                self._seen[key] = seen + 1
                return None
            if not self._accumulators_follow(self._capture_stream(x.device)):
                self._blocked = getattr(self, "_blocked", 0) + 1
                if self._blocked == 1:
                    import warnings
                    warnings.warn("gdn_b200: the train step is not captured as a CUDA graph while a tensor with grad_fn "
                                  "from an earlier forward of this model is still referenced (its gradient accumulators "
                                  "are bound to another stream); steps stay eager")
                if self._blocked >= 8:
                    self.cuda_graph = False                   # stop probing
                return None
            self._capture(key, x, y)
            entry = self._graphs[key]
        graph, xs, ys, loss = entry
        xs.copy_(x, non_blocking=True)
        ys.copy_(y, non_blocking=True)
        graph.replay()
        # the replay moves embedding.weight without touching its version counter: drop any eval-mode graph cache
        self.model._graph_cache = None
        return loss

    def step(self, x, y):
        """One train step (train.py:68-73); returns the loss DETACHED (train.py only reads `loss.item()`): a caller
        that keeps the previous step's loss must not thereby keep its autograd graph -- and the gradient-accumulator
        nodes bound to the stream it ran on -- alive into the step that is captured as a CUDA graph."""
        if self.cuda_graph and x.is_cuda and y.is_cuda and x.dtype == torch.float32 and y.dtype == torch.float32 \
                and x.numel() <= self.GRAPH_MAX_ELEMS and x.is_contiguous() and y.is_contiguous() \
                and not torch.cuda.is_current_stream_capturing():
            loss = self._graph_step(x, y)
            if loss is not None:
                return loss
        if self.nvls is not None:
            self.nvls.zero_grad()
            out = self.model(x, None)
            loss = torch.nn.functional.mse_loss(out, y, reduction="mean")
            loss.backward()
            self.nvls.step()                       # all-reduce + Adam + parameter broadcast: one kernel
            return loss.detach()
        if self.flat is not None:
            self.flat.zero_grad()
            out = self.model(x, None)
            loss = torch.nn.functional.mse_loss(out, y, reduction="mean")
            loss.backward()
            world = 1
            if dist.is_available() and dist.is_initialized():
                world = dist.get_world_size(self.group)
                if world > 1:
                    dist.all_reduce(self.flat.grad_buffer, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.step(grad_scale=1.0 / world)
            return loss.detach()
        self.opt.zero_grad(set_to_none=True)
        out = self.model(x, None)
        loss = torch.nn.functional.mse_loss(out, y, reduction="mean")
        loss.backward()
        self.reduce()
        self.opt.step()
        return loss.detach()


# ------------------------------------------------------------------------------------------------ evaluation
def shard_loader_batches(n_batches, rank, world):
    """Contiguous range [lo, hi) of a loader's batches evaluated by `rank` (test.py:43: batches are independent in
    eval mode; contiguous so that every rank ends up with a contiguous range of ticks)."""
    return shard_bounds(n_batches, rank, world)


class LoaderSlice:
    """Batches [lo, hi) of any loader; a `gdn_b200.datasets.WindowLoader` is re-indexed instead of skipped through
    (its batches are gathered on the device: no reason to form the ones another rank evaluates)."""

    def __init__(self, inner, lo, hi):
        self.inner, self.lo, self.hi = inner, int(lo), int(hi)

    def __iter__(self):
        inner, lo, hi = self.inner, self.lo, self.hi
        if all(hasattr(inner, a) for a in ("dataset", "indices", "batch_size")) and not getattr(inner, "shuffle", False) \
                and hasattr(inner.dataset, "loader"):
            bs = inner.batch_size
            yield from inner.dataset.loader(bs, indices=inner.indices[lo * bs:hi * bs])
            return
        for k, b in enumerate(inner):
            if k >= hi:
                break
            if k >= lo:
                yield b

    def __len__(self):
        return self.hi - self.lo


def sharded_test(model, dataloader, group=None):
    """test.py:20-75 on `world` ranks ("Scoring: shards ... by window for the eval forward"): every rank runs
    the eval forward on a contiguous slice of the loader's batches.  Returns (avg_loss over ALL ticks, pred_local
    [T_r, N], gt_local [T_r, N], labels_local [T_r]) -- device tensors; the global average loss costs one all-reduce
    of two doubles.  Feed the local tensors to `sharded_scores`."""
    from .test import test as _test
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_loader_batches(len(dataloader), rank, world)

    dev = next(model.parameters()).device
    if hi > lo:
        loss, res = _test(model, LoaderSlice(dataloader, lo, hi))
        pred, gt, lab = res.device_tensors
        n_batches = hi - lo
    else:                                                       # more ranks than batches
        N = model.embedding.num_embeddings
        loss, n_batches = 0.0, 0
        pred = torch.empty((0, N), dtype=torch.float32, device=dev)
        gt, lab = torch.empty_like(pred), torch.empty((0, N), dtype=torch.float32, device=dev)
    acc = torch.tensor([loss * n_batches, float(n_batches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(acc, group=group)
    return float(acc[0].item() / max(acc[1].item(), 1.0)), pred, gt, lab[:, 0] if lab.dim() == 2 else lab


def sharded_scores(pred_local, gt_local, group=None, want_scores=True, score_fn=None, tick_counts=None):
    """evaluate.py:6-36 + 134-139 on `world` ranks ("Scoring shards by sensor ... final gather of [T]
    maxima").  Input: this rank's contiguous tick range [T_r, N] of predictions / ground truth (ranks in tick order,
    as `sharded_test` leaves them).  The scorer needs whole series per sensor, so the one exchange step is an
    all-to-all that turns the tick sharding into a sensor sharding ([T_r, N] -> [T, N_r], 8 T N / world bytes per
    rank); every rank then scores its N_r sensors, and the per-tick maximum over sensors is one all-reduce(MAX) of
    [T] doubles.  `tick_counts` (ticks held by every rank, in rank order) saves the size exchange and its host sync
    when the caller knows them (sharded_test's contiguous split does).
    Returns (scores_local [N_r, T] float64 or None, top1 [T] float64 on every rank, (n_lo, n_hi))."""
    from . import ops
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if score_fn is None:
        score_fn = lambda p, g: ops.score(p, g, want_scores=want_scores, want_top1=True)[:2]
    T_r, N = pred_local.shape
    dev = pred_local.device
    if world == 1:
        s, top1 = score_fn(pred_local.contiguous(), gt_local.contiguous())
        return s, top1, (0, N)
    if tick_counts is not None:
        t_counts = [int(c) for c in tick_counts]
        if len(t_counts) != world or t_counts[rank] != T_r:
            raise ValueError("tick_counts must list every rank's tick count (this rank's included)")
    else:
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = T_r
        dist.all_reduce(counts, group=group)
        t_counts = [int(c) for c in counts.tolist()]
    T = sum(t_counts)
    cols = [shard_bounds(N, q, world) for q in range(world)]
    n_lo, n_hi = cols[rank]
    n_me = n_hi - n_lo
    # send to rank q: x[my ticks, q's sensors]; receive from rank q: x[q's ticks, my sensors].  One exchange per
    # tensor: the receive buffer, filled in rank (= tick) order, IS the contiguous [T, n_me] the scorer reads
    in_split = [T_r * (b - a) for a, b in cols]
    out_split = [tq * n_me for tq in t_counts]
    mine = []
    for x in (pred_local, gt_local):
        send = torch.cat([x[:, a:b].reshape(-1) for a, b in cols])
        recv = torch.empty(T * n_me, dtype=x.dtype, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split, group=group)
        mine.append(recv.view(T, n_me))
    if n_me > 0:
        s, top1 = score_fn(mine[0], mine[1])
    else:
        s = torch.empty((0, T), dtype=torch.float64, device=dev)
        top1 = torch.full((T,), float("-inf"), dtype=torch.float64, device=dev)
    dist.all_reduce(top1, op=dist.ReduceOp.MAX, group=group)
    return s, top1, (n_lo, n_hi)
