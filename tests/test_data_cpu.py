"""The window-dataset oracle (oracle/data_oracle.py) against the vectors the reference's own
datasets/TimeDataset.py produced (tests/golden/timedataset_small.npz, made by oracle/make_golden_data.py)."""
import numpy as np
import pytest

from golden_util import load
from oracle import data_oracle as do

CASES = [("train", 5, 3), ("test", 5, 3), ("train", 1, 1), ("test", 16, 7)]


@pytest.mark.parametrize("mode,W,S", CASES)
def test_oracle_windows_equal_reference(mode, W, S):
    rec = load("timedataset_small")
    x, y, lab = do.process(rec["raw"].tolist(), W, S, mode)
    tag = f"{mode}_w{W}_s{S}"
    assert np.array_equal(x, rec[tag + "_x"]) and np.array_equal(y, rec[tag + "_y"])
    assert np.array_equal(lab, rec[tag + "_labels"].astype(np.float64))
    assert np.array_equal(x[2], rec[tag + "_item2_x"]) and np.array_equal(y[2], rec[tag + "_item2_y"])


def test_window_ends_and_float32_rounding():
    assert do.window_ends(10, 3, 4, "train").tolist() == [3, 7]
    assert do.window_ends(10, 3, 4, "test").tolist() == [3, 4, 5, 6, 7, 8, 9]
    assert do.window_ends(3, 3, 1, "test").tolist() == []
    # lists go through torch.tensor -> float32 before .double(); float64 arrays keep their bits
    raw = [[0.1, 0.2, 0.3], [0.0, 1.0, 0.0]]
    x, _, _ = do.process(raw, 1, 1, "test")
    assert x[0, 0, 0] == np.float64(np.float32(0.1))
    x64, _, _ = do.process(np.asarray(raw, dtype=np.float64), 1, 1, "test")
    assert x64[0, 0, 0] == 0.1


def test_mirror_refuses_cpu():
    from gdn_b200.datasets import TimeDataset
    with pytest.raises(RuntimeError, match="no CPU path"):
        TimeDataset([[0.0, 1.0], [0.0, 0.0]], None, config={"slide_win": 1, "slide_stride": 1}, device="cpu")


def test_test_loop_oracle_equals_reference():
    """oracle test_loop (test.py:20-75) fed with the reference's own per-batch predictions reproduces the
    reference's avg_loss and result lists (tests/golden/test_loop_small.npz, oracle/make_golden_test.py)."""
    rec = load("test_loop_small")
    B = int(rec["dims"][5])
    pred, gt, labels = rec["pred"], rec["gt"], rec["labels"]
    cuts = list(range(0, len(pred), B))
    avg, (p, g, l) = do.test_loop([pred[c:c + B] for c in cuts], [gt[c:c + B] for c in cuts], [labels[c:c + B, 0] for c in cuts])
    assert abs(avg - float(rec["avg_loss"])) <= 2e-7 * abs(float(rec["avg_loss"]))
    assert np.array_equal(p, pred) and np.array_equal(g, gt) and np.array_equal(l, labels)
    # ground truth and labels of the loop are the dataset's windows
    N, T, W = (int(v) for v in rec["dims"][:3])
    _, y, lab = do.process(rec["raw"].tolist(), W, 1, "test")
    assert np.array_equal(gt, y.astype(np.float32)) and np.array_equal(labels[:, 0], lab.astype(np.float32))


def test_host_mirrors_fail_loudly_without_cuda():
    """No CPU fallback anywhere on the widened rows either (f-2, f-3, f-4): the mirrors raise instead of computing."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    from gdn_b200.evaluate import get_best_performance_data
    from gdn_b200.optim import FlatAdam
    from gdn_b200.test import test as run_test
    from gdn_b200.util.data import eval_scores
    with pytest.raises(RuntimeError, match="no CPU path"):
        run_test(torch.nn.Linear(2, 2), [])
    with pytest.raises(RuntimeError, match="no CPU path"):
        eval_scores([0.1, 0.2], [0, 1], 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        get_best_performance_data([[0.1, 0.2]], [0, 1])
    with pytest.raises(RuntimeError, match="no CPU path"):
        FlatAdam([torch.nn.Parameter(torch.zeros(3))])


def test_test_result_is_indexable_like_the_reference_lists():
    import torch
    from gdn_b200.test import TestResult
    p, g, l = torch.rand(5, 3), torch.rand(5, 3), torch.zeros(5, 3)
    res = TestResult(p, g, l)
    assert len(res) == 3 and torch.equal(res[0], p) and res.device_tensors[1] is g
    assert np.asarray(res).shape == (3, 5, 3)                                  # main.py:get_score does np.array(test_result)
    assert np.array_equal(np.asarray(res)[2, :, 0], np.zeros(5))
