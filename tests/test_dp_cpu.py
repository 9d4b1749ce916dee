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
