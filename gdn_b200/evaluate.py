"""Test-time scoring with the reference's function names and return types
(evaluate.py:6-36, 48-68; util/data.py:75-82), computed by the CUDA scorer
(csrc/scoring.cu) instead of per-sensor numpy loops."""
import numpy as np
import torch

from gdn_b200 import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("gdn_b200.evaluate needs a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _scores_for(pred, gt):
    """pred, gt: array-likes [T, N] -> scores [N, T] float64 (numpy)."""
    dev = _device()
    p = torch.as_tensor(np.asarray(pred, dtype=np.float32)).to(dev)
    g = torch.as_tensor(np.asarray(gt, dtype=np.float32)).to(dev)
    scores, _, _ = ops.score(p, g, want_scores=True, want_top1=False)
    return scores.cpu().numpy()


def get_err_scores(test_res, val_res):
    """evaluate.py:48-68: (predict, gt) of ONE sensor -> smoothed normalised error [T]."""
    test_predict, test_gt = test_res
    p = np.asarray(test_predict, dtype=np.float32).reshape(-1, 1)
    g = np.asarray(test_gt, dtype=np.float32).reshape(-1, 1)
    return _scores_for(p, g)[0]


def get_full_err_scores(test_result, val_result):
    """evaluate.py:6-36: [pred, gt, labels] nested lists ([3][T][N]) for the test and the
    validation run -> (all_scores [N, T], all_normals [N, Tv]) float64."""
    np_test = np.asarray(test_result, dtype=np.float32)
    np_val = np.asarray(val_result, dtype=np.float32)
    all_scores = _scores_for(np_test[0], np_test[1])
    all_normals = _scores_for(np_val[0], np_val[1])
    return all_scores, all_normals


def get_final_err_scores(test_result, val_result):
    """evaluate.py:39-44 as intended (the reference version passes a keyword
    get_full_err_scores does not accept): per-tick maximum over sensors."""
    dev = _device()
    np_test = np.asarray(test_result, dtype=np.float32)
    p = torch.as_tensor(np_test[0]).to(dev)
    g = torch.as_tensor(np_test[1]).to(dev)
    _, top1, _ = ops.score(p, g, want_scores=False, want_top1=True)
    return top1.cpu().numpy()
