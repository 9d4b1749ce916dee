"""Multi-GPU numerics of the data-parallel step on real GPUs: needs at least two devices, so it is
skipped on the single-GPU box the driver runs `-m gpu` on; `bench.py --gpus N` carries the same check (`dp_check`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_numerics_on_two_gpus():
    """Summed gradients / G == average of the float64 oracle's per-shard gradients; the NVSwitch-multicast optimiser
    kernel == NCCL all-reduce + flat Adam (and leaves bit-identical replicas); SyncBN == the oracle on the concatenated
    batch."""
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["grad_pass"] and out["syncbn_pass"] and out["pass_all_ranks"], out
    assert out.get("nvls_pass", True), out
