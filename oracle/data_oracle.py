"""CPU restatement of the reference's window dataset (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Follows datasets/TimeDataset.py:10-62: raw_data = N sensor rows + one label row; windows end at
i in range(slide_win, T, slide_stride) in 'train' mode and range(slide_win, T) otherwise (:44);
x_i = data[:, i-slide_win:i], y_i = data[:, i], label_i = labels[i] (:48-53).  Pinned against the
reference class itself by oracle/make_golden.py -> tests/golden/timedataset_*.npz.
"""
import numpy as np


def window_ends(total_time_len, slide_win, slide_stride, mode):
    """datasets/TimeDataset.py:44"""
    if mode == "train":
        return np.arange(slide_win, total_time_len, slide_stride, dtype=np.int64)
    return np.arange(slide_win, total_time_len, dtype=np.int64)


def process(raw_data, slide_win, slide_stride, mode):
    """datasets/TimeDataset.py:16-27, 33-62 -> (x [M, N, W], y [M, N], labels [M]) float64."""
    # torch.tensor(list of Python floats) is float32 (default dtype) BEFORE the .double() of :23-24: the
    # reference's "doubles" are float32-rounded values (main.py hands over lists: util/preprocess.py:33-45)
    raw = np.asarray(raw_data)
    raw = (raw if isinstance(raw_data, np.ndarray) and raw.dtype == np.float64 else raw.astype(np.float32)).astype(np.float64)
    data, labels = raw[:-1], raw[-1]
    ends = window_ends(data.shape[1], slide_win, slide_stride, mode)
    x = np.stack([data[:, e - slide_win:e] for e in ends]) if len(ends) else np.zeros((0, data.shape[0], slide_win))
    y = np.stack([data[:, e] for e in ends]) if len(ends) else np.zeros((0, data.shape[0]))
    return x, y, labels[ends].astype(np.float32).astype(np.float64)   # torch.Tensor(labels_arr) is float32 (:58)


def test_loop(predict_batches, y_batches, label_batches):
    """test.py:20-75 given the per-batch model outputs: avg_loss = mean over batches of the float32
    MSELoss(reduction='mean') values (:49-50, :66-67, :75), and the concatenated [pred, gt, labels] with the
    labels repeated per sensor (:53-63).  Inputs: lists of float32 arrays [b, N], [b, N], [b]."""
    losses = [float(np.mean((p.astype(np.float32) - y.astype(np.float32)) ** 2, dtype=np.float32))
              for p, y in zip(predict_batches, y_batches)]
    pred = np.concatenate(predict_batches).astype(np.float32)
    gt = np.concatenate(y_batches).astype(np.float32)
    labels = np.concatenate([np.repeat(l.astype(np.float32)[:, None], p.shape[1], axis=1)
                             for l, p in zip(label_batches, predict_batches)])
    return sum(losses) / len(losses), [pred, gt, labels]
