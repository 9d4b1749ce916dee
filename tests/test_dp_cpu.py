"""CPU, world_size 2, gloo: the host-side data-parallel logic (window sharding + the flat
gradient all-reduce).  The model here is a small torch module: the CUDA path itself is covered
by the -m gpu tests; what is checked is that averaging per-shard gradients through ONE flat
buffer reproduces the full-batch gradient when every shard is the same size, that p.grad
aliases the flat buffer afterwards, and that ranks stay bit-identical after the optimiser step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gdn_b200.dp import FlatGradAllReduce, WindowShardedTrainer, shard_bounds


def test_shard_bounds_cover_the_batch_without_overlap():
    for gb, world in ((512, 8), (64, 1), (10, 4), (3, 5)):
        spans = [shard_bounds(gb, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
            assert b == c and b - a >= d - c >= 0
    assert shard_bounds(512, 3, 8) == (192, 256)
    with pytest.raises(ValueError):
        shard_bounds(8, 8, 8)


class _Tiny(torch.nn.Module):
    """[B, N, W] -> [B, N]; two parameters of different shapes plus one that gets no gradient."""

    def __init__(self, N, W):
        super().__init__()
        self.w = torch.nn.Parameter(torch.linspace(-1, 1, W).repeat(N, 1))
        self.b = torch.nn.Parameter(torch.zeros(N))
        self.unused = torch.nn.Parameter(torch.ones(3))

    def forward(self, x, _edge_index=None):
        return (x * self.w).sum(-1) + self.b


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        N, W, GB = 6, 4, 8
        x, y = torch.rand(GB, N, W), torch.rand(GB, N)
        lo, hi = shard_bounds(GB, rank, world)
        model = _Tiny(N, W)
        trainer = WindowShardedTrainer(model, lr=1e-2, fused_adam=False)
        loss = trainer.step(x[lo:hi], y[lo:hi])
        flat = trainer.reduce.flat
        assert flat is not None and flat.numel() == N * W + N + 3
        # p.grad must alias the flat buffer (no copy back)
        assert model.w.grad.data_ptr() == flat.data_ptr()
        # full-batch reference on every rank
        ref = _Tiny(N, W)
        torch.nn.functional.mse_loss(ref(x), y).backward()
        assert torch.allclose(model.w.grad, ref.w.grad, atol=1e-7)
        assert torch.allclose(model.b.grad, ref.b.grad, atol=1e-7)
        assert float(model.unused.grad.abs().max()) == 0.0
        # ranks agree bit-for-bit after the step
        mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        assert torch.equal(both[0], both[1])
        ret[rank] = float(loss)
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world


def test_single_process_is_a_noop():
    model = _Tiny(3, 2)
    red = FlatGradAllReduce(model.parameters())
    assert red() is None and red.world_size() == 1


def test_graph_row_shard_covers_rows_in_aligned_chunks():
    from gdn_b200.dp import graph_row_shard
    for n in (1, 127, 128, 1024, 1500, 4096, 16384, 16385):
        for world in (1, 2, 3, 8):
            seen, chunk0 = 0, None
            for r in range(world):
                r0, r1, chunk = graph_row_shard(n, r, world)
                chunk0 = chunk if chunk0 is None else chunk0
                assert chunk == chunk0 and chunk % 128 == 0 and world * chunk >= n
                assert r0 == min(n, r * chunk) and r0 <= r1 <= n and r0 % 128 == 0 or r0 == n
                assert r1 == n or r1 % 128 == 0
                assert r0 == seen or r0 == n                     # contiguous, in rank order
                seen = max(seen, r1)
            assert seen == n
    with pytest.raises(ValueError):
        graph_row_shard(10, 2, 2)


def _cpu_graph_rows(V, topk, use_tensor_cores=-1, want_idx=True, kth=None, margin=0.03, rows=None, out=None):
    """Stand-in for ops.graph_build (CUDA only) with the same contract, in torch on the CPU: rows [r0, r1) of the
    cosine top-k and the neighbour table (top-k without self, then self) written into `out`."""
    N, K = V.shape[0], int(topk)
    Vn = V / V.norm(dim=1, keepdim=True)
    r0, r1 = rows if rows is not None else (0, N)
    idx = torch.topk(Vn[r0:r1] @ Vn.t(), K, dim=1).indices
    nbr = torch.full((r1 - r0, K + 1), -1, dtype=torch.int32)
    for a in range(r1 - r0):
        keep = [int(j) for j in idx[a] if int(j) != r0 + a] + [r0 + a]
        nbr[a, :len(keep)] = torch.tensor(keep, dtype=torch.int32)
    full_idx, full_nbr = out
    full_idx[r0:r1] = idx
    full_nbr[r0:r1] = nbr
    return full_idx, full_nbr


def _graph_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gdn_b200 import ops
        from gdn_b200.models.GDN import GDN
        ops.graph_build = _cpu_graph_rows                      # the kernels are covered by the -m gpu tests
        torch.manual_seed(3)
        N, K = 300, 7                                           # 300 rows -> chunks of 256: rank 1 builds only 44 rows
        model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=16, input_dim=4, topk=K).train()
        model.shard_graph_build(rank, world)
        idx, nbr = model.build_graph()
        assert idx.shape == (N, K) and nbr.shape == (N, K + 1)
        full = (torch.empty(N, K, dtype=torch.int64), torch.empty(N, K + 1, dtype=torch.int32))
        _cpu_graph_rows(model.embedding.weight.detach(), K, out=full)
        assert torch.equal(idx, full[0]) and torch.equal(nbr, full[1])
        model.eval()                                            # eval forwards build locally: no collective
        if rank == 0:
            called = []
            ops.graph_build = lambda *a, **k: called.append(k.get("rows")) or _cpu_graph_rows(
                a[0], a[1], out=(torch.empty(N, K, dtype=torch.int64), torch.empty(N, K + 1, dtype=torch.int32)))
            idx_e, _ = model.build_graph()
            assert called == [None] and torch.equal(idx_e, full[0])
        model.shard_graph_build(None, None)
        assert model._graph_shard is None
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_row_sharded_graph_build_world2_gloo():
    """The data-parallel graph exchange on CPU (gloo, world 2): each rank fills its aligned row range, the in-place
    all-gather assembles the padded tables, the slices equal a full build; an eval forward on one rank needs no peer."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_graph_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world
