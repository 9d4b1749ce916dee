"""The metrics oracle (oracle/metrics_oracle.py) against what the reference's own util/data.py:eval_scores and
evaluate.py:get_best/val_performance_data produced with sklearn + scipy (tests/golden/metrics_small.npz)."""
import numpy as np
import pytest

from golden_util import load
from oracle import metrics_oracle as mo

CASES = ["smooth", "ties", "short"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_equals_reference(name):
    rec = load("metrics_small")
    scores, labels = rec[name + "_scores"], rec[name + "_labels"].tolist()
    top = scores.max(axis=0)
    fm, th = mo.eval_scores(top.tolist(), labels, 400, return_thresold=True)
    assert np.array_equal(np.asarray(fm), rec[name + "_fmeas"]) and np.array_equal(np.asarray(th), rec[name + "_thresholds"])
    assert np.array_equal(np.asarray(mo.eval_scores(top.tolist()[5:], labels, 50)), rec[name + "_padded50"])
    best = mo.get_best_performance_data(scores, labels, topk=1)
    val = mo.get_val_performance_data(scores, rec[name + "_normal"], labels, topk=1)
    assert np.allclose(best, rec[name + "_best"], rtol=1e-12, atol=0) and np.allclose(val, rec[name + "_val"], rtol=1e-12, atol=0)
    assert np.allclose(mo.get_best_performance_data(scores, labels, topk=2), rec[name + "_best_top2"], rtol=1e-12, atol=0)


def test_oracle_edge_cases():
    assert mo.eval_scores([1.0, 1.0, 1.0], [0, 0, 0], 4) == [0.0, 0.0, 0.0, 0.0]          # no positives: F1 = 0
    fm, th = mo.eval_scores([3.0, 1.0, 2.0], [1, 0, 0], 3, return_thresold=True)
    assert fm == [0.5, 2.0 / 3.0, 1.0] and th == [1.0, 2.0, 3.0]
    with pytest.raises(ValueError):
        mo.roc_auc([1, 1], [0.1, 0.2])
