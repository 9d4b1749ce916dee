"""CPU restatement of the reference's test-time scoring.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows evaluate.py:6-36 (get_full_err_scores), evaluate.py:48-68 (get_err_scores) and
util/data.py:75-82 (get_err_median_and_iqr).  scipy.stats.iqr(x) is
``np.percentile(x, 75) - np.percentile(x, 25)`` with numpy's default 'linear'
interpolation; numpy's lerp is ``a + (b-a)*t`` for t < 0.5 and ``b - (b-a)*(1-t)``
otherwise, which ``quantile_linear`` reproduces without calling numpy's percentile so
that the CUDA kernel can be checked against an explicit formula.
"""
import numpy as np

EPSILON = 1e-2      # evaluate.py:58
BEFORE_NUM = 3      # evaluate.py:63


def quantile_linear(sorted_col, q):
    """numpy 'linear' quantile of an ascending float64 vector (q in [0, 1])."""
    n = sorted_col.shape[0]
    pos = q * (n - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    t = pos - lo
    a, b = sorted_col[lo], sorted_col[hi]
    d = b - a
    return a + d * t if t < 0.5 else b - d * (1.0 - t)


def median_of_sorted(sorted_col):
    """np.median: mean of the two middle order statistics (np.mean of 2 = (a+b)/2)."""
    n = sorted_col.shape[0]
    if n % 2 == 1:
        return sorted_col[n // 2]
    return (sorted_col[n // 2 - 1] + sorted_col[n // 2]) / 2.0


def err_scores(pred_col, gt_col):
    """evaluate.py:48-68 for one sensor: robust-normalised, smoothed forecast error."""
    delta = np.abs(np.asarray(pred_col, dtype=np.float64) - np.asarray(gt_col, dtype=np.float64))
    srt = np.sort(delta)
    med = median_of_sorted(srt)
    iqr = quantile_linear(srt, 0.75) - quantile_linear(srt, 0.25)
    err = (delta - med) / (np.abs(iqr) + EPSILON)
    out = np.zeros_like(err)
    for t in range(BEFORE_NUM, err.shape[0]):
        acc = -0.0                       # np.mean -> pairwise sum, sequential below 8 terms
        for v in err[t - BEFORE_NUM:t + 1]:
            acc += v
        out[t] = acc / (BEFORE_NUM + 1)
    return out


def err_scores_vec(pred_col, gt_col):
    """Vectorised twin of err_scores (same arithmetic order) for large T."""
    delta = np.abs(np.asarray(pred_col, dtype=np.float64) - np.asarray(gt_col, dtype=np.float64))
    srt = np.sort(delta)
    med = median_of_sorted(srt)
    iqr = quantile_linear(srt, 0.75) - quantile_linear(srt, 0.25)
    err = (delta - med) / (np.abs(iqr) + EPSILON)
    out = np.zeros_like(err)
    if err.shape[0] > BEFORE_NUM:
        acc = ((err[:-3] + err[1:-2]) + err[2:-1]) + err[3:]
        out[BEFORE_NUM:] = acc / 4.0
    return out


def full_err_scores(pred, gt, vectorised=True):
    """evaluate.py:6-36 for one result set: pred, gt [T, N] -> scores [N, T] float64."""
    pred = np.asarray(pred)
    gt = np.asarray(gt)
    fn = err_scores_vec if vectorised else err_scores
    return np.stack([fn(pred[:, i], gt[:, i]) for i in range(pred.shape[1])], axis=0)


def top1_over_sensors(scores):
    """evaluate.py:134-139 with topk=1: the per-tick maximum over sensors."""
    return np.max(scores, axis=0)
