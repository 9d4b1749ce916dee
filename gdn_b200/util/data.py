"""Mirror of the reference's util/data.py:28-51 (eval_scores): the 400-step rank-threshold F1 sweep, computed on
the device from one stable sort.  The reference calls sklearn.f1_score and list.index once
per step."""
import numpy as np
import torch

from .. import _lib
from .._lib import check, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("gdn_b200.util.data needs a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def sorted_ticks(scores, labels):
    """float64 scores [T], labels [T] -> (sorted_scores, labels_sorted) on the device: stable ascending sort, i.e.
    position r holds the tick of ordinal rank r + 1 (scipy.stats.rankdata(method='ordinal'), util/data.py:36)."""
    dev = _device()
    s = torch.as_tensor(np.asarray(scores, dtype=np.float64)).to(dev) if not torch.is_tensor(scores) else scores.to(dev, torch.float64)
    l = torch.as_tensor(np.asarray(labels, dtype=np.float32)).to(dev) if not torch.is_tensor(labels) else labels.to(dev, torch.float32)
    if s.numel() != l.numel():
        raise RuntimeError("scores and labels must have the same length")
    ss, order = torch.sort(s.reshape(-1), stable=True)
    return ss.contiguous(), l.reshape(-1)[order].contiguous()


def sweep(sorted_scores, labels_sorted, th_steps):
    """The sweep proper on sorted ticks -> (fmeas [S], thresholds [S]) float64 CUDA tensors."""
    lib = _lib.load()
    T = int(sorted_scores.numel())
    th_vals = np.array(range(th_steps)) * 1.0 / th_steps                       # util/data.py:39, same float64 products
    x = th_vals * T
    k_pred = np.floor(x).astype(np.int32)                                      # rank > x  <=>  sorted position >= floor(x)
    k_thr = np.array([int(v + 1) - 1 for v in x], dtype=np.int32)              # position of rank int(x + 1) (:47)
    dev = sorted_scores.device
    kp, kt = torch.from_numpy(k_pred).to(dev), torch.from_numpy(k_thr).to(dev)
    fmeas = torch.empty(th_steps, dtype=torch.float64, device=dev)
    thresholds = torch.empty(th_steps, dtype=torch.float64, device=dev)
    check(lib.gdn_f1_sweep(ptr(sorted_scores), ptr(labels_sorted), T, ptr(kp), ptr(kt), int(th_steps), ptr(fmeas),
                           ptr(thresholds), _stream()), "gdn_f1_sweep")
    return fmeas, thresholds


def eval_scores(scores, true_scores, th_steps, return_thresold=False):
    """util/data.py:28-51: F1 of (ordinal rank > i/th_steps * T) for i in range(th_steps); optionally the score at
    each step's rank.  Same return types as the reference: lists of Python floats."""
    scores = list(np.asarray(scores, dtype=np.float64).reshape(-1))
    pad = len(true_scores) - len(scores)
    if pad > 0:
        scores = [0.0] * pad + scores                                            # :29-33
    if len(scores) != len(true_scores) or len(scores) == 0:
        raise ValueError("eval_scores: scores must not be longer than the labels, and not empty")
    ss, ls = sorted_ticks(scores, true_scores)
    fmeas, thresholds = sweep(ss, ls, int(th_steps))
    fm = fmeas.cpu().tolist()
    if return_thresold:
        return fm, thresholds.cpu().tolist()
    return fm
