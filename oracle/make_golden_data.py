"""Golden vectors for the window dataset: runs the reference's own
datasets/TimeDataset.py (loaded by path: site-packages' HuggingFace `datasets` shadows the package name)
on a small synthetic series and stores its windows; also asserts the oracle restatement equals it.
    python oracle/make_golden_data.py        (build container only: needs /root/reference)"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import data_oracle as do  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_TimeDataset", "/root/reference/datasets/TimeDataset.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(11)
N, T = 7, 53
raw = np.concatenate([rng.random((N, T)), (rng.random((1, T)) > 0.8).astype(np.float64)], 0).tolist()
out = {"raw": np.asarray(raw)}
for mode, W, S in (("train", 5, 3), ("test", 5, 3), ("train", 1, 1), ("test", 16, 7)):
    ds = ref.TimeDataset(raw, torch.zeros(2, 4, dtype=torch.long), mode=mode, config={"slide_win": W, "slide_stride": S})
    tag = f"{mode}_w{W}_s{S}"
    out[tag + "_x"], out[tag + "_y"], out[tag + "_labels"] = ds.x.numpy(), ds.y.numpy(), ds.labels.numpy()
    x, y, lab = do.process(raw, W, S, mode)
    assert np.array_equal(x, ds.x.numpy()) and np.array_equal(y, ds.y.numpy()), tag
    assert np.array_equal(lab, ds.labels.numpy().astype(np.float64)), tag
    item = ds[2]
    out[tag + "_item2_x"], out[tag + "_item2_y"] = item[0].numpy(), item[1].numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "timedataset_small.npz"), **out)
print("wrote tests/golden/timedataset_small.npz", len(out), "arrays")
