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
        row = [int(j) for j in idx[a]]
        keep = [j for j in row if j != r0 + a] + [r0 + a]
        nbr[a, :len(keep)] = torch.tensor(keep, dtype=torch.int32)
        if r0 + a in row:                                       # include/gdn_b200.h: last slot = -2 - position of self
            nbr[a, K] = -2 - row.index(r0 + a)
    full_idx, full_nbr = out
    if full_idx is not None:
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


def test_idx_from_nbr_rebuilds_the_topk_order():
    """The data-parallel exchange ships the int32 neighbour table only; learned_graph is rebuilt from it.  Self first
    (the usual case), self in the middle, self last, self absent."""
    from gdn_b200 import ops
    N, K = 9, 4
    idx = torch.tensor([[0, 3, 5, 7], [2, 1, 4, 6], [8, 7, 6, 2], [1, 2, 4, 5], [4, 0, 1, 2],
                        [0, 1, 5, 2], [6, 1, 2, 3], [0, 7, 2, 3], [1, 2, 3, 8]], dtype=torch.int64)
    full = (torch.empty(N, K, dtype=torch.int64), torch.empty(N, K + 1, dtype=torch.int32))
    nbr = torch.full((N, K + 1), -1, dtype=torch.int32)
    for i in range(N):
        row = [int(j) for j in idx[i]]
        keep = [j for j in row if j != i] + [i]
        nbr[i, :len(keep)] = torch.tensor(keep, dtype=torch.int32)
        if i in row:
            nbr[i, K] = -2 - row.index(i)
    assert int(nbr[3, K]) == 3                                  # self absent: K non-self entries, then self
    assert torch.equal(ops.idx_from_nbr(nbr), idx)


def _score_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from gdn_b200.dp import shard_bounds, sharded_scores
        from oracle import scoring_oracle as so
        rng = np.random.default_rng(7)
        T, N = 53, 7                                            # neither divides by the world size
        gt = rng.random((T, N)).astype(np.float32)
        pred = (gt + rng.normal(0, 0.05, (T, N))).astype(np.float32)

        def cpu_scorer(p, g):                                   # stand-in for ops.score (CUDA only), same contract
            s = so.full_err_scores(p.numpy(), g.numpy())
            return torch.from_numpy(s), torch.from_numpy(s.max(axis=0))

        lo, hi = shard_bounds(T, rank, world)
        s, top1, (n_lo, n_hi) = sharded_scores(torch.from_numpy(pred[lo:hi]), torch.from_numpy(gt[lo:hi]),
                                               score_fn=cpu_scorer)
        ref = so.full_err_scores(pred, gt)
        assert (n_lo, n_hi) == shard_bounds(N, rank, world)
        assert np.array_equal(s.numpy(), ref[n_lo:n_hi])
        assert np.array_equal(top1.numpy(), ref.max(axis=0))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scoring_gloo(world):
    """SURVEY section 8e / row e-2 on CPU: ticks sharded over ranks -> all-to-all -> sensors sharded -> scores of the
    local sensors + all-reduce(MAX) of the per-tick maximum; equals the single-process scorer bit for bit."""
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_score_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world
