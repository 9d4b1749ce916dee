"""CPU restatement of the reference's threshold sweep and summary metrics (TEST INFRASTRUCTURE ONLY -- see
oracle/__init__.py).  Follows util/data.py:28-51 (eval_scores) and evaluate.py:101-158
(get_val_performance_data, get_best_performance_data) with sklearn's f1/precision/recall written out as the
integer-count ratios they are, and roc_auc_score as the tie-averaged rank statistic.  Pinned against the
reference functions themselves (sklearn 1.9, scipy 1.18 in the build container) by oracle/make_golden_metrics.py
-> tests/golden/metrics_small.npz."""
import numpy as np


def _ranks_ordinal(scores):
    """scipy.stats.rankdata(scores, method='ordinal') (util/data.py:36): ties ranked by position."""
    order = np.argsort(np.asarray(scores, dtype=np.float64), kind="stable")
    ranks = np.empty(len(order), dtype=np.int64)
    ranks[order] = np.arange(1, len(order) + 1)
    return ranks


def eval_scores(scores, true_scores, th_steps, return_thresold=False):
    """util/data.py:28-51"""
    scores = [0] * (len(true_scores) - len(scores)) + list(scores)            # :29-33
    T = len(scores)
    ranks = _ranks_ordinal(scores)
    lab = np.asarray(true_scores) != 0
    th_vals = np.array(range(th_steps)) * 1.0 / th_steps                     # :39
    fmeas, thresholds = [None] * th_steps, [None] * th_steps
    pos_at = np.empty(T + 1, dtype=np.int64)
    pos_at[ranks] = np.arange(T)
    for i in range(th_steps):
        cur_pred = ranks > th_vals[i] * T                                    # :43
        denom = int(lab.sum()) + int(cur_pred.sum())
        fmeas[i] = 2.0 * int((cur_pred & lab).sum()) / denom if denom else 0.0   # sklearn f1_score, zero_division -> 0
        thresholds[i] = scores[pos_at[int(th_vals[i] * T + 1)]]              # :47-48
    return (fmeas, thresholds) if return_thresold else fmeas


def topk_sum(total_err_scores, topk=1):
    """evaluate.py:105-111 / 133-138: per tick, the sum of the topk largest sensor scores."""
    a = np.asarray(total_err_scores, dtype=np.float64)
    return np.sort(a, axis=0)[-topk:].sum(axis=0)


def _prf(pred, lab):
    tp = int((pred & lab).sum()); fp = int((pred & ~lab).sum()); fn = int((~pred & lab).sum())
    pre = tp / (tp + fp) if tp + fp else 0.0
    rec = tp / (tp + fn) if tp + fn else 0.0
    f1 = 2.0 * tp / (2 * tp + fp + fn) if 2 * tp + fp + fn else 0.0
    return f1, pre, rec


def roc_auc(labels, scores):
    """sklearn.metrics.roc_auc_score for binary labels == Mann-Whitney U with tie-averaged ranks."""
    from scipy.stats import rankdata
    lab = np.asarray(labels) != 0
    P, N = int(lab.sum()), int((~lab).sum())
    if P == 0 or N == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    r = rankdata(np.asarray(scores, dtype=np.float64), method="average")
    return (r[lab].sum() - P * (P + 1) / 2.0) / (P * N)


def get_best_performance_data(total_err_scores, gt_labels, topk=1):
    """evaluate.py:129-158"""
    s = topk_sum(total_err_scores, topk)
    fmeas, thresolds = eval_scores(s, gt_labels, 400, return_thresold=True)
    th_i = fmeas.index(max(fmeas))
    thresold = thresolds[th_i]
    lab = np.asarray(gt_labels) != 0
    _, pre, rec = _prf(s > thresold, lab)
    return max(fmeas), pre, rec, roc_auc(gt_labels, s), thresold


def get_val_performance_data(total_err_scores, normal_scores, gt_labels, topk=1):
    """evaluate.py:101-127"""
    s = topk_sum(total_err_scores, topk)
    thresold = np.max(normal_scores)
    lab = np.asarray(gt_labels) != 0
    f1, pre, rec = _prf(s > thresold, lab)
    return f1, pre, rec, roc_auc(gt_labels, s), thresold
