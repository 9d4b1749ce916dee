"""GPU parity of the test-time scorer (csrc/scoring.cu) against the reference's vectors and the
numpy oracle.  float64 throughout; the bar is bit-exact."""
import numpy as np
import pytest
import torch

from golden_util import load
from oracle import scoring_oracle as so

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["score_small", "score_even"])
def test_scores_match_reference_bit_exact(name):
    from gdn_b200 import ops
    rec = load(name)
    s, top1, stats = ops.score(torch.from_numpy(rec["pred"]).cuda(), torch.from_numpy(rec["gt"]).cuda(),
                               want_stats=True)
    assert np.array_equal(s.cpu().numpy(), rec["scores"])
    assert np.array_equal(top1.cpu().numpy(), rec["top1"])


@pytest.mark.parametrize("T,N", [(1, 3), (3, 2), (4, 1), (5, 7), (2044, 27), (17275, 127), (44986, 51)],
                         ids=["T1", "T3", "T4", "T5", "msl", "wadi", "swat"])
def test_scores_match_oracle(T, N):
    from gdn_b200 import ops
    rng = np.random.default_rng(T * 131 + N)
    gt = rng.random((T, N)).astype(np.float32)
    pred = (gt + rng.normal(0, 0.05, (T, N)) * (1 + 8 * (rng.random((T, N)) > 0.98))).astype(np.float32)
    if N > 2:
        pred[:, 1] = gt[:, 1]                      # zero error everywhere: median = IQR = 0
        pred[:, 2] = gt[:, 2] + np.float32(0.25)   # constant error: ties in every order statistic
    s, top1, stats = ops.score(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), want_stats=True)
    ref = so.full_err_scores(pred, gt, vectorised=True)
    assert np.array_equal(s.cpu().numpy(), ref)
    assert np.array_equal(top1.cpu().numpy(), ref.max(axis=0))
    delta = np.abs(pred.astype(np.float64) - gt.astype(np.float64))
    assert np.array_equal(stats.cpu().numpy()[:, 0], np.median(delta, axis=0))


def test_reference_named_entry_points():
    """gdn_b200.evaluate keeps the reference's function names / nested-list inputs."""
    from gdn_b200 import evaluate as ev
    rec = load("score_small")
    labels = np.zeros_like(rec["gt"])
    res = [rec["pred"].tolist(), rec["gt"].tolist(), labels.tolist()]
    scores, normals = ev.get_full_err_scores(res, res)
    assert np.array_equal(scores, rec["scores"]) and np.array_equal(normals, rec["scores"])
    one = ev.get_err_scores((rec["pred"][:, 3].tolist(), rec["gt"][:, 3].tolist()), None)
    assert np.array_equal(one, rec["scores"][3])
    assert np.array_equal(ev.get_final_err_scores(res, res), rec["top1"])


def test_scoring_is_shift_invariant_at_scale():
    """Size-independent property on a 1M-tick series: adding the same constant to pred and gt
    leaves |pred-gt| (hence the scores) unchanged only up to fp32 rounding of the inputs, but
    swapping pred and gt leaves them bit-identical."""
    from gdn_b200 import ops
    T, N = 1_000_000, 4
    g = torch.Generator(device="cuda").manual_seed(3)
    gt = torch.rand(T, N, device="cuda", generator=g)
    pred = gt + 0.05 * torch.randn(T, N, device="cuda", generator=g)
    a, ta, _ = ops.score(pred, gt)
    b, tb, _ = ops.score(gt, pred)
    assert torch.equal(a, b) and torch.equal(ta, tb)
    assert torch.equal(ta, a.max(dim=0)[0])
    assert (a[:, :3] == 0).all()


def test_scores_bit_exact_over_a_wide_dynamic_range():
    """The scorer divides by multiplying with the correctly rounded reciprocal plus an exact-residual correction; this
    must equal numpy's IEEE division bit for bit -- also for errors spanning many binades, huge and tiny sensors
    (outside the safe exponent window the kernel divides), and quantised sensors with heavy ties."""
    from gdn_b200 import ops
    rng = np.random.default_rng(11)
    T, N = 3001, 12
    gt = rng.random((T, N)).astype(np.float32)
    err = rng.normal(0, 1, (T, N)) * np.exp(rng.uniform(-12, 2, (T, N)))
    pred = (gt + err).astype(np.float32)
    pred[:, 0] = gt[:, 0] + np.float32(1e30) * rng.random(T).astype(np.float32)      # huge errors
    pred[:, 1] = gt[:, 1] * np.float32(1.0 + 1e-7)                                    # last-bit errors
    pred[:, 2] = gt[:, 2] + np.round(rng.random(T) * 4).astype(np.float32) / 8        # 5 distinct error levels
    gt[:, 3] *= np.float32(1e-30); pred[:, 3] = gt[:, 3] * np.float32(1.5)            # tiny sensor
    s, top1, _ = ops.score(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda())
    ref = so.full_err_scores(pred, gt, vectorised=True)
    assert np.array_equal(s.cpu().numpy(), ref)
    assert np.array_equal(top1.cpu().numpy(), ref.max(axis=0))
