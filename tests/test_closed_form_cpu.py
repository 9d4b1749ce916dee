"""CPU: the closed forms the CUDA kernels implement (oracle/closed_form.py: per-node scalars,
aggregation before the linear map, BatchNorm-1 statistics from the moments of A, the five D-wide
backward passes) equal the reference's forward and autograd in float64 -- so any difference the
GPU tests see is rounding, not algebra."""
import pytest
import torch

from golden_util import load, meta, normwise, state_dict
from oracle import closed_form as cf


@pytest.mark.parametrize("name", ["c1_msl", "c1_stress", "c2_swat", "c3_wadi", "w16_small", "w10_odd"])
def test_closed_forms_equal_reference_float64(name):
    rec = load(name)
    m = meta(rec)
    sd = state_dict(rec, dtype=torch.float64)
    x, y = torch.from_numpy(rec["x"]).double(), torch.from_numpy(rec["y"]).double()
    idx = torch.from_numpy(rec["idx"])
    mask = torch.from_numpy(rec["drop_mask"]).double()
    for moments in (True, False):
        pred, loss, grads = cf.forward_backward(sd, x, y, idx, mask, bn1_from_moments=moments)
        assert normwise(pred, rec["pred_train64"]) < 1e-11
        assert abs(loss.item() - float(rec["loss_train64"])) < 1e-12 * abs(float(rec["loss_train64"]))
        for k, g in grads.items():
            ref = torch.from_numpy(rec["grad64/" + k])
            scale = max(ref.abs().max().item(), 1e-9)
            assert (g.reshape(ref.shape) - ref).abs().max().item() <= 1e-9 * scale + 1e-13, (k, moments)   # gnn.bias grad is analytically 0


def test_neighbour_table_is_the_self_loop_fix_up():
    idx = torch.tensor([[0, 2, 1], [2, 0, 3], [3, 1, 0], [3, 2, 1]])
    nbr = cf.neighbour_table(idx)
    assert nbr.tolist() == [[2, 1, 0, -1], [2, 0, 3, 1], [3, 1, 0, 2], [2, 1, 3, -1]]
