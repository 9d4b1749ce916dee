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

    def __init__(self, model, lr=1e-3, weight_decay=0.0, group=None, fused_adam=None, shard_graph=None, flat_adam=False):
        """flat_adam=True: parameters, gradients and Adam moments live in flat buffers
        (`gdn_b200.optim.FlatAdam`); the all-reduce works on the gradient buffer in place and the 1/world scaling is
        fused into the single Adam kernel.  Default: torch.optim.Adam (fused on CUDA) + FlatGradAllReduce."""
        self.model = model
        self.group = group
        params = list(model.parameters())
        self.flat = None
        if flat_adam:
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
        # the graph depends on the (replicated) embedding only: every rank builds 1/world of its rows and the
        # neighbour tables are all-gathered (the one exchange step of the forward; off for a single process)
        if shard_graph is None:
            shard_graph = (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
                           and all(p.is_cuda for p in params) and hasattr(model, "shard_graph_build"))
        if shard_graph:
            model.shard_graph_build(dist.get_rank(group), dist.get_world_size(group), group)

    def step(self, x, y):
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
            return loss
        self.opt.zero_grad(set_to_none=True)
        out = self.model(x, None)
        loss = torch.nn.functional.mse_loss(out, y, reduction="mean")
        loss.backward()
        self.reduce()
        self.opt.step()
        return loss
