"""torchrun check of the row-sharded graph build: every rank's assembled tables equal a local full build, and a
few data-parallel train steps keep ranks in lock-step.  torchrun --nproc-per-node 2 tools/dp_graph_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gdn_b200 import ops
from gdn_b200.dp import WindowShardedTrainer
from gdn_b200.models.GDN import GDN

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
for N, W, D, K, B in ((1500, 5, 64, 17, 8), (4096, 16, 128, 32, 16), (16384, 16, 128, 64, 4)):
    torch.manual_seed(5)
    model = GDN([torch.zeros(2, 1, dtype=torch.long)], N, dim=D, input_dim=W, topk=K).to(dev).train()
    trainer = WindowShardedTrainer(model, lr=1e-3)
    assert model._graph_shard is not None
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for step in range(4):
        x, y = torch.rand(B, N, W, device=dev, generator=g), torch.rand(B, N, device=dev, generator=g)
        ref_idx, ref_nbr = ops.graph_build(model.embedding.weight.detach(), K)     # full build of the current embedding
        trainer.step(x, y)
        assert torch.equal(model.learned_graph, ref_idx), (N, step, rank)
    w = model.embedding.weight.detach().clone()
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    assert all(torch.equal(ws[0], t) for t in ws), "ranks diverged"
    if rank == 0:
        print(f"N={N}: sharded graph == full graph on every step, ranks in lock-step")
dist.barrier()
dist.destroy_process_group()
