"""Mirror of the reference's evaluation loop (test.py:20-75).

The reference grows three `[t, N]` tensors with `torch.cat` every batch (O(T^2) copies), synchronises on
`loss.item()` every batch and finally converts everything to nested Python lists (`.tolist()` of 3*T*N
floats) that `evaluate.py` turns back into arrays.  Here predictions, ground truth and labels land in
preallocated device buffers `[T, N]`, the per-batch losses stay on the device, and ONE synchronisation at the
end fetches the average loss.  The result is still indexable like the reference's `[pred, gt, labels]` (CPU
float32 tensors, `np.array(result)` works), and carries the device copies for `gdn_b200.evaluate` so that
scoring needs no host round trip.
"""
import torch
import torch.nn.functional as F


class TestResult(list):
    """`[pred, gt, labels]`, each `[T, N]` float32 on the CPU (test.py:71-75 returns nested lists of the same
    numbers); `.device_tensors` holds the same three on the GPU."""
    __test__ = False                     # not a pytest class

    def __init__(self, pred, gt, labels):
        self.device_tensors = (pred, gt, labels)
        super().__init__([pred.cpu(), gt.cpu(), labels.cpu()])


def _count(dataloader):
    for attr in ("indices", "dataset"):
        obj = getattr(dataloader, attr, None)
        if obj is not None:
            try:
                return len(obj)
            except TypeError:
                pass
    return None


def test(model, dataloader):
    """test.py:20-75 -> (avg_loss, TestResult).  `dataloader` yields `(x, y, labels, edge_index)`: a
    `gdn_b200.datasets.WindowLoader` (batches already on the device) or the reference's DataLoader."""
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("gdn_b200.test runs on a CUDA device (no CPU path)")
    model.eval()
    total = _count(dataloader)
    pred_buf = gt_buf = lab_buf = None
    losses = []
    t = 0
    with torch.no_grad():
        for x, y, labels, edge_index in dataloader:
            x = x.to(dev, non_blocking=True).float()
            y = y.to(dev, non_blocking=True).float()
            labels = labels.to(dev, non_blocking=True).float()
            predicted = model(x, edge_index).float()
            losses.append(F.mse_loss(predicted, y))                     # stays on the device (test.py:50,66-67 sync here)
            b, n = predicted.shape
            if pred_buf is None:
                cap = total if total is not None else max(4 * b, 1024)
                pred_buf = torch.empty((cap, n), dtype=torch.float32, device=dev)
                gt_buf, lab_buf = torch.empty_like(pred_buf), torch.empty((cap,), dtype=torch.float32, device=dev)
            if t + b > pred_buf.shape[0]:                               # unknown length: grow geometrically
                cap = max(2 * pred_buf.shape[0], t + b)
                pred_buf = torch.cat([pred_buf[:t], pred_buf.new_empty((cap - t, n))])
                gt_buf = torch.cat([gt_buf[:t], gt_buf.new_empty((cap - t, n))])
                lab_buf = torch.cat([lab_buf[:t], lab_buf.new_empty((cap - t,))])
            pred_buf[t:t + b] = predicted
            gt_buf[t:t + b] = y
            lab_buf[t:t + b] = labels
            t += b
    if t == 0:
        raise RuntimeError("test(): the dataloader yielded no batch (the reference divides by zero here, test.py:75)")
    avg_loss = float(torch.stack(losses).double().mean().item())        # sum(float32 losses as Python floats) / len
    n = pred_buf.shape[1]
    labels_tn = lab_buf[:t].unsqueeze(1).expand(t, n).contiguous()      # test.py:53
    return avg_loss, TestResult(pred_buf[:t], gt_buf[:t], labels_tn)
