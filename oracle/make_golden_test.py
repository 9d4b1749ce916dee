"""Golden vectors for the test() loop: the reference's own test.py:20-75 driving the
reference GDN (through the PyG-1.5.0 stand-in) over the reference TimeDataset/DataLoader on CPU.
    python oracle/make_golden_test.py        (build container only: needs /root/reference)"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyg_shim  # noqa: E402

pyg_shim.install()
ref_gdn, _, _ = pyg_shim.import_reference()
REF = "/root/reference"


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


sys.path.insert(0, REF)
ref_test = load("ref_test", "test.py")
ref_ds = load("ref_TimeDataset", "datasets/TimeDataset.py")
from util.env import set_device  # noqa: E402

set_device("cpu")
torch.set_num_threads(1)
rng = np.random.default_rng(23)
N, T, W, K, D, B = 9, 80, 5, 4, 32, 16
raw = np.concatenate([rng.random((N, T)), (rng.random((1, T)) > 0.85).astype(np.float64)], 0).tolist()
ei = torch.zeros(2, 1, dtype=torch.long)
ds = ref_ds.TimeDataset(raw, ei, mode="test", config={"slide_win": W, "slide_stride": 1})
dl = torch.utils.data.DataLoader(ds, batch_size=B, shuffle=False, num_workers=0)
torch.manual_seed(4)
model = ref_gdn.GDN([ei], N, dim=D, input_dim=W, topk=K)
with torch.no_grad():                      # move the BatchNorm running statistics off their defaults
    for bn in (model.bn_outlayer_in, model.gnn_layers[0].bn):
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.5, 1.5)
avg_loss, (pred, gt, labels) = ref_test.test(model, dl)
out = {"raw": np.asarray(raw), "dims": np.asarray([N, T, W, K, D, B]), "avg_loss": np.asarray(avg_loss),
       "pred": np.asarray(pred, dtype=np.float32), "gt": np.asarray(gt, dtype=np.float32),
       "labels": np.asarray(labels, dtype=np.float32)}
for k, v in model.state_dict().items():
    out["sd." + k] = v.numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "test_loop_small.npz"), **out)
print("wrote tests/golden/test_loop_small.npz: avg_loss", avg_loss, "pred", out["pred"].shape)
