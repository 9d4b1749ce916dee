"""Under torchrun (one process per GPU): the data-parallel step's numerics on the real model -- summed gradients vs the
float64 oracle per shard, the NVSwitch-multicast step vs the NCCL step, SyncBN vs the oracle on the concatenated
batch (bench.dp_gradient_check).  Rank 0 prints one JSON line; exit code 1 if any rank fails.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import dp_gradient_check

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
out = dp_gradient_check(dev, rank, world)
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if out["pass_all_ranks"] else 1)
