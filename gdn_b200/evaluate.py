"""Test-time scoring and summary metrics with the reference's function names and return types
(evaluate.py:6-36, 48-68, 75-158; util/data.py:28-51, 75-82), computed by the CUDA scorer (csrc/scoring.cu)
and the sweep kernels (csrc/metrics.cu) instead of per-sensor numpy loops and 400 sklearn calls."""
import numpy as np
import torch

from gdn_b200 import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("gdn_b200.evaluate needs a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _pred_gt(result):
    """[pred, gt, labels] ([3][T][N] nested lists / arrays, or a gdn_b200.test.TestResult whose device copies
    are used as they are: no host round trip) -> (pred, gt) float32 CUDA tensors [T, N]."""
    dev = _device()
    held = getattr(result, "device_tensors", None)
    if held is not None:
        return held[0].to(dev, torch.float32).contiguous(), held[1].to(dev, torch.float32).contiguous()
    arr = np.asarray([np.asarray(result[0], dtype=np.float32), np.asarray(result[1], dtype=np.float32)])
    return torch.as_tensor(arr[0]).to(dev), torch.as_tensor(arr[1]).to(dev)


def _scores_for(pred, gt):
    """pred, gt: array-likes or tensors [T, N] -> scores [N, T] float64 (numpy)."""
    dev = _device()
    p = pred if torch.is_tensor(pred) else torch.as_tensor(np.asarray(pred, dtype=np.float32))
    g = gt if torch.is_tensor(gt) else torch.as_tensor(np.asarray(gt, dtype=np.float32))
    scores, _, _ = ops.score(p.to(dev, torch.float32).contiguous(), g.to(dev, torch.float32).contiguous(),
                             want_scores=True, want_top1=False)
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
    all_scores = _scores_for(*_pred_gt(test_result))
    all_normals = _scores_for(*_pred_gt(val_result))
    return all_scores, all_normals


def get_final_err_scores(test_result, val_result):
    """evaluate.py:39-44 as intended (the reference version passes a keyword
    get_full_err_scores does not accept): per-tick maximum over sensors."""
    p, g = _pred_gt(test_result)
    _, top1, _ = ops.score(p, g, want_scores=False, want_top1=True)
    return top1.cpu().numpy()


# ----------------------------------------------------------------------------- summary metrics
def _topk_sum(total_err_scores, topk):
    """evaluate.py:105-111 / 133-138: per tick, the sum of the topk largest sensor scores -> float64 CUDA [T]."""
    dev = _device()
    a = total_err_scores if torch.is_tensor(total_err_scores) else torch.as_tensor(np.asarray(total_err_scores, dtype=np.float64))
    a = a.to(dev, torch.float64)
    if a.dim() != 2 or not 1 <= topk <= a.shape[0]:
        raise ValueError("total_err_scores must be [N, T] with 1 <= topk <= N")
    if topk == 1:
        return a.max(dim=0).values.contiguous()
    return torch.topk(a, topk, dim=0).values.flip(0).sum(dim=0).contiguous()      # ascending order of summation, as np.sum


def _labels_dev(gt_labels, T):
    l = torch.as_tensor(np.asarray(gt_labels, dtype=np.float32)).to(_device())
    if l.numel() != T:
        raise ValueError("gt_labels must have one entry per tick")
    return l.contiguous()


def _prf_auc(scores, labels, threshold):
    """precision, recall, f1 of (scores > threshold) and ROC-AUC of the scores, from device-side integer counts."""
    from gdn_b200 import _lib
    from gdn_b200._lib import check, ptr
    from gdn_b200.util.data import sorted_ticks
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    T = int(scores.numel())
    counts = torch.zeros(4, dtype=torch.int64, device=scores.device)
    check(lib.gdn_binary_counts(ptr(scores), ptr(labels), T, float(threshold), ptr(counts), st), "gdn_binary_counts")
    ss, ls = sorted_ticks(scores, labels)
    ranksum = torch.zeros(1, dtype=torch.float64, device=scores.device)
    npos = torch.zeros(1, dtype=torch.int64, device=scores.device)
    check(lib.gdn_auc_ranksum(ptr(ss), ptr(ls), T, ptr(ranksum), ptr(npos), st), "gdn_auc_ranksum")
    tp, fp, fn, _ = (int(v) for v in counts.cpu().tolist())
    P = int(npos.item())
    if P == 0 or P == T:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")   # sklearn's error
    pre = tp / (tp + fp) if tp + fp else 0.0
    rec = tp / (tp + fn) if tp + fn else 0.0
    f1 = 2.0 * tp / (2 * tp + fp + fn) if 2 * tp + fp + fn else 0.0
    auc = (float(ranksum.item()) - P * (P + 1) / 2.0) / (P * (T - P))
    return f1, pre, rec, auc


def get_f1_scores(total_err_scores, gt_labels, topk=1):
    """evaluate.py:75-99"""
    from gdn_b200.util.data import eval_scores
    return eval_scores(_topk_sum(total_err_scores, topk).cpu().numpy(), gt_labels, 400)


def get_best_performance_data(total_err_scores, gt_labels, topk=1):
    """evaluate.py:129-158 -> (best F1 of the 400-step sweep, precision, recall, ROC-AUC, threshold)."""
    from gdn_b200.util.data import sorted_ticks, sweep
    s = _topk_sum(total_err_scores, topk)
    l = _labels_dev(gt_labels, int(s.numel()))
    ss, ls = sorted_ticks(s, l)
    fmeas, thresholds = sweep(ss, ls, 400)
    fm = fmeas.cpu().tolist()
    th_i = fm.index(max(fm))                                                       # first maximum, as list.index (:145)
    thresold = thresholds[th_i].item()
    _, pre, rec, auc = _prf_auc(s, l, thresold)
    return max(fm), pre, rec, auc, thresold


def get_val_performance_data(total_err_scores, normal_scores, gt_labels, topk=1):
    """evaluate.py:101-127: threshold = the largest validation score."""
    s = _topk_sum(total_err_scores, topk)
    l = _labels_dev(gt_labels, int(s.numel()))
    ns = normal_scores if torch.is_tensor(normal_scores) else torch.as_tensor(np.asarray(normal_scores, dtype=np.float64))
    thresold = float(ns.max().item())
    f1, pre, rec, auc = _prf_auc(s, l, thresold)
    return f1, pre, rec, auc, thresold
