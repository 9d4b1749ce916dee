"""GPU parity of the threshold sweep / summary metrics (csrc/metrics.cu, gdn_b200/util/data.py, gdn_b200/evaluate.py)
against the reference's own sklearn-based vectors and the oracle.  F1 values and thresholds are bit-exact
(ratios of integer counts / selected scores); precision, recall exact; AUC to 1e-12."""
import numpy as np
import pytest

from golden_util import load
from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["smooth", "ties", "short"])
def test_sweep_and_summary_equal_reference(name):
    from gdn_b200.evaluate import get_best_performance_data, get_f1_scores, get_val_performance_data
    from gdn_b200.util.data import eval_scores
    rec = load("metrics_small")
    scores, labels = rec[name + "_scores"], rec[name + "_labels"].tolist()
    top = scores.max(axis=0)
    fm, th = eval_scores(top.tolist(), labels, 400, return_thresold=True)
    assert isinstance(fm, list) and np.array_equal(np.asarray(fm), rec[name + "_fmeas"])
    assert np.array_equal(np.asarray(th), rec[name + "_thresholds"])
    assert np.array_equal(np.asarray(eval_scores(top.tolist()[5:], labels, 50)), rec[name + "_padded50"])   # front padding
    assert np.array_equal(np.asarray(get_f1_scores(scores, labels)), rec[name + "_fmeas"])
    best = get_best_performance_data(scores, labels, topk=1)
    assert best[0] == rec[name + "_best"][0] and best[1] == rec[name + "_best"][1] and best[2] == rec[name + "_best"][2]
    assert abs(best[3] - rec[name + "_best"][3]) <= 1e-12 and best[4] == rec[name + "_best"][4]
    val = get_val_performance_data(scores, rec[name + "_normal"], labels, topk=1)
    assert np.allclose(val, rec[name + "_val"], rtol=1e-12, atol=0)
    assert np.allclose(get_best_performance_data(scores, labels, topk=2), rec[name + "_best_top2"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("T", [1, 2, 399, 44986, 1 << 20], ids=["T1", "T2", "T399", "swat", "1M"])
def test_sweep_matches_oracle_and_properties(T):
    from gdn_b200.util.data import eval_scores
    rng = np.random.default_rng(T)
    scores = np.round(rng.gamma(2.0, 1.0, T), 2 if T > 1000 else 6)                     # ties at scale
    labels = (rng.random(T) < 0.12).astype(np.float64)
    fm, th = eval_scores(scores, labels, 400, return_thresold=True)
    if T <= 50000:
        ofm, oth = mo.eval_scores(scores.tolist(), labels.tolist(), 400, return_thresold=True)
        assert fm == ofm and th == oth
    # size-independent properties: thresholds are sorted scores at the stepped ranks, hence non-decreasing;
    # step 0 predicts every tick anomalous: F1 = 2P / (P + T)
    assert all(a <= b for a, b in zip(th, th[1:])) and th[0] == scores.min()
    P = labels.sum()
    assert fm[0] == (2.0 * P / (P + T) if P + T else 0.0)
    assert all(0.0 <= f <= 1.0 for f in fm)


def test_summary_errors_and_end_to_end_with_scorer():
    from gdn_b200.evaluate import get_best_performance_data, get_full_err_scores
    with pytest.raises(ValueError, match="one class"):
        get_best_performance_data(np.ones((2, 5)), [0, 0, 0, 0, 0])
    with pytest.raises(ValueError):
        get_best_performance_data(np.ones((2, 5)), [0, 1])
    # scorer -> sweep chain equals the oracle chain on the same predictions
    rng = np.random.default_rng(5)
    T, N = 3000, 12
    gt = rng.random((T, N)).astype(np.float32)
    labels = (rng.random(T) < 0.1).astype(np.float64)
    pred = gt + rng.normal(0, 0.03, (T, N)).astype(np.float32)
    pred[labels == 1, 3] += 0.4
    s, _ = get_full_err_scores([pred, gt, np.zeros_like(gt)], [pred, gt, np.zeros_like(gt)])
    got = get_best_performance_data(s, labels.tolist())
    want = mo.get_best_performance_data(s, labels.tolist())
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2] and abs(got[3] - want[3]) <= 1e-12 and got[4] == want[4]
    assert got[0] > 0.3              # isolated single-tick anomalies are smeared by the 4-tap smoothing
