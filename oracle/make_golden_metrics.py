"""Golden vectors for the threshold sweep / summary metrics: the reference's own
util/data.py:eval_scores and evaluate.py:get_best_performance_data / get_val_performance_data (sklearn, scipy).
    python oracle/make_golden_metrics.py        (build container only: needs /root/reference)"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics_oracle as mo, pyg_shim  # noqa: E402

pyg_shim.install()
_, _, ref_eval = pyg_shim.import_reference()
warnings.filterwarnings("ignore")
rng = np.random.default_rng(31)
out = {}
cases = {
    "smooth": (11, 1000, 0.1, False),      # N sensors, T ticks, anomaly rate, heavy ties
    "ties": (5, 777, 0.3, True),
    "short": (3, 37, 0.2, False),
}
for name, (N, T, rate, ties) in cases.items():
    scores = rng.gamma(2.0, 1.0, (N, T))
    if ties:
        scores = np.round(scores, 0)
    labels = (rng.random(T) < rate).astype(np.float64)
    labels[:3] = [0, 1, 0]
    scores[:, labels == 1] += 1.5
    normal = rng.gamma(2.0, 1.0, (N, 200))
    top = np.max(scores, axis=0)
    fm, th = ref_eval.eval_scores(top.tolist(), labels.tolist(), 400, return_thresold=True)
    ofm, oth = mo.eval_scores(top.tolist(), labels.tolist(), 400, return_thresold=True)
    assert fm == ofm and th == oth, name
    padded = ref_eval.eval_scores(top.tolist()[5:], labels.tolist(), 50)
    assert padded == mo.eval_scores(top.tolist()[5:], labels.tolist(), 50), name
    best = ref_eval.get_best_performance_data(scores, labels.tolist(), topk=1)
    obest = mo.get_best_performance_data(scores, labels.tolist(), topk=1)
    val = ref_eval.get_val_performance_data(scores, normal, labels.tolist(), topk=1)
    oval = mo.get_val_performance_data(scores, normal, labels.tolist(), topk=1)
    for a, b in zip(best + val, obest + oval):
        assert abs(a - b) <= 1e-12 * max(1.0, abs(a)), (name, best, obest, val, oval)
    best2 = ref_eval.get_best_performance_data(scores, labels.tolist(), topk=2)
    out.update({f"{name}_scores": scores, f"{name}_labels": labels, f"{name}_normal": normal,
                f"{name}_fmeas": np.asarray(fm), f"{name}_thresholds": np.asarray(th), f"{name}_padded50": np.asarray(padded),
                f"{name}_best": np.asarray(best, dtype=np.float64), f"{name}_val": np.asarray(val, dtype=np.float64),
                f"{name}_best_top2": np.asarray(best2, dtype=np.float64)})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics_small.npz"), **out)
print("wrote tests/golden/metrics_small.npz", {k: v.shape for k, v in out.items() if k.endswith("best")})
