"""Device-resident mirror of the reference's window dataset (datasets/TimeDataset.py:10-78).

Same constructor and item semantics as the reference class, but nothing W-fold redundant is ever built:
the series `[N, T]` is uploaded once as float32 and a batch of windows is gathered on the GPU by one kernel
(`gdn_window_batch`) from B window-end indices.  `loader(...)` replaces `DataLoader(dataset, ...)` in
`main.py:84-85,142-146`: it yields `(x [B,N,W], y [B,N], labels [B], edge_index)` already on the device and in
float32, so `train.py:66` / `test.py:43` (`item.float().to(device)`) become no-ops and the per-step H2D
transfer -- including the unused fully-connected `edge_index` replicated per sample -- disappears.
"""
import numpy as np
import torch

from .. import _lib
from .._lib import check, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


class TimeDataset:
    def __init__(self, raw_data, edge_index, mode="train", config=None, device="cuda"):
        """raw_data: N sensor rows + one label row (datasets/TimeDataset.py:16-17); config: slide_win, slide_stride."""
        self.raw_data = raw_data
        self.config = config
        self.edge_index = edge_index
        self.mode = mode
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("gdn_b200.datasets.TimeDataset keeps the series on a CUDA device (no CPU path)")
        raw = np.asarray(raw_data, dtype=np.float64)
        if raw.ndim != 2 or raw.shape[0] < 2:
            raise RuntimeError("raw_data must hold at least one sensor row and the label row")
        self.slide_win, self.slide_stride = int(config["slide_win"]), int(config["slide_stride"])
        self.node_num, self.total_time_len = raw.shape[0] - 1, raw.shape[1]
        if self.slide_win < 1 or self.slide_stride < 1:
            raise RuntimeError("slide_win and slide_stride must be >= 1")
        # (datasets/TimeDataset.py:23-24 make doubles; train.py:66 casts every batch to float)
        self.series = torch.from_numpy(raw[:-1].astype(np.float32)).contiguous().to(dev)
        self.label_series = torch.from_numpy(raw[-1].astype(np.float32)).contiguous().to(dev)
        rng = (range(self.slide_win, self.total_time_len, self.slide_stride) if mode == "train"
               else range(self.slide_win, self.total_time_len))                     # datasets/TimeDataset.py:44
        self.win_end = torch.tensor(list(rng), dtype=torch.int32, device=dev)
        self.labels = self.label_series[self.win_end.long()].cpu()                   # :58, float32

    @classmethod
    def from_series(cls, series, labels, edge_index, mode="train", config=None):
        """Same dataset from a float32 CUDA tensor `series [N, T]` (and `labels [T]` or None) that is already
        resident -- no host round trip (synthetic generators, or a series another stage left on the device)."""
        if not (torch.is_tensor(series) and series.is_cuda and series.dtype == torch.float32 and series.dim() == 2):
            raise RuntimeError("series must be a float32 CUDA tensor [N, T] (gdn_b200 has no CPU path)")
        self = cls.__new__(cls)
        self.raw_data, self.config, self.edge_index, self.mode = None, config, edge_index, mode
        self.slide_win, self.slide_stride = int(config["slide_win"]), int(config["slide_stride"])
        if self.slide_win < 1 or self.slide_stride < 1:
            raise RuntimeError("slide_win and slide_stride must be >= 1")
        self.node_num, self.total_time_len = int(series.shape[0]), int(series.shape[1])
        self.series = series.contiguous()
        self.label_series = (torch.zeros(self.total_time_len, dtype=torch.float32, device=series.device) if labels is None
                             else labels.to(series.device, torch.float32).contiguous())
        step = self.slide_stride if mode == "train" else 1                          # datasets/TimeDataset.py:44
        self.win_end = torch.arange(self.slide_win, self.total_time_len, step, dtype=torch.int32, device=series.device)
        self.labels = self.label_series[self.win_end.long()].cpu()
        return self

    def __len__(self):
        return int(self.win_end.numel())

    def batch(self, idx, _validated=False):
        """Windows `idx` (sequence or int tensor of dataset positions) -> (x [B,N,W], y [B,N], labels [B]) on the device.
        (`_validated`: the loader has checked its index set once, so its batches skip the host-synchronising range check.)"""
        lib = _lib.load()
        dev = self.series.device
        idx = torch.as_tensor(idx, device=dev).long().reshape(-1)
        B = int(idx.numel())
        N, W = self.node_num, self.slide_win
        x = torch.empty((B, N, W), dtype=torch.float32, device=dev)
        y = torch.empty((B, N), dtype=torch.float32, device=dev)
        lab = torch.empty((B,), dtype=torch.float32, device=dev)
        if B == 0:
            return x, y, lab
        if not _validated and (int(idx.min()) < 0 or int(idx.max()) >= len(self)):
            raise IndexError("window index out of range")
        ends = self.win_end[idx].contiguous()
        err = self._err_flag()
        check(lib.gdn_window_batch(ptr(self.series), ptr(self.label_series), N, self.total_time_len, W, ptr(ends), B,
                                   ptr(x), ptr(y), ptr(lab), ptr(err), _stream()), "gdn_window_batch")
        return x, y, lab

    def _err_flag(self):
        """One persistent device flag per dataset (no memset per batch): the kernel stores b+1 there when window b
        of a batch lies outside the series (and hands zeros to the model); `check_errors` reads it."""
        flag = getattr(self, "_err", None)
        if flag is None or flag.device != self.series.device:
            flag = torch.zeros(1, dtype=torch.int32, device=self.series.device)
            self._err = flag
        return flag

    def check_errors(self):
        """Host-synchronising check of the kernel's error flag (the loader calls it once per epoch)."""
        flag = getattr(self, "_err", None)
        if flag is not None:
            bad = int(flag.item())
            if bad:
                flag.zero_()
                raise IndexError(f"gdn_window_batch: window {bad - 1} of a batch lies outside the series")

    def __getitem__(self, idx):
        """Reference item: (feature [N,W], y [N], label, edge_index) as doubles (datasets/TimeDataset.py:64-73).
        Kept for compatibility; the fast path is `loader`."""
        if idx < 0:
            idx += len(self)
        x, y, lab = self.batch([idx])
        return x[0].double().cpu(), y[0].double().cpu(), lab[0].double().cpu(), self.edge_index.long()

    def loader(self, batch_size, shuffle=False, indices=None, generator=None, drop_last=False):
        return WindowLoader(self, batch_size, shuffle=shuffle, indices=indices, generator=generator, drop_last=drop_last)


class WindowLoader:
    """What `DataLoader(dataset | Subset(dataset, indices), batch_size, shuffle)` is to the reference
    (main.py:84-85, 142-146), with batches formed on the device.  The permutation is drawn on the device too, so
    an epoch moves no data over PCIe."""

    def __init__(self, dataset, batch_size, shuffle=False, indices=None, generator=None, drop_last=False):
        self.dataset = dataset
        self.batch_size = int(batch_size)
        if self.batch_size < 1:
            raise RuntimeError("batch_size must be >= 1")
        self.shuffle = bool(shuffle)
        dev = dataset.series.device
        self.indices = (torch.arange(len(dataset), device=dev) if indices is None
                        else torch.as_tensor(indices, device=dev).long().reshape(-1))
        if self.indices.numel() and (int(self.indices.min()) < 0 or int(self.indices.max()) >= len(dataset)):
            raise IndexError("subset index out of range")
        self.generator = generator
        self.drop_last = bool(drop_last)

    def __len__(self):
        n = int(self.indices.numel())
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = self.indices
        if self.shuffle:
            perm = torch.randperm(order.numel(), device=order.device, generator=self.generator)
            order = order[perm]
        for k in range(len(self)):
            sel = order[k * self.batch_size:(k + 1) * self.batch_size]
            x, y, lab = self.dataset.batch(sel, _validated=True)
            yield x, y, lab, self.dataset.edge_index
        self.dataset.check_errors()
