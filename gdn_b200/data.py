"""Host -> device batch feed for the reference's train/test loops (train.py:63-66, test.py:43-44).

The reference moves every batch with a blocking `.to(device)` inside the step, so the copy and the
compute serialise (and it ships the unused fully-connected `edge_index` every step).  `Prefetcher`
wraps any iterable of `(x, y, ...)` CPU batches -- e.g. the reference's DataLoader over `TimeDataset`
-- and keeps ONE batch in flight on a copy stream: pinned staging buffers, `non_blocking` H2D copies,
an event per batch, and the consumer stream waits on that event only.  Tensors whose name position is
listed in `skip` (default: the 4th, `edge_index`) are passed through untouched.
"""
import torch


class Prefetcher:
    def __init__(self, batches, device, skip=(3,), dtype=torch.float32):
        self.batches = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Prefetcher feeds a CUDA device (gdn_b200 has no CPU path)")
        self.skip = set(skip)
        self.dtype = dtype
        self.stream = torch.cuda.Stream(device=self.device)
        self._pinned = {}

    def __len__(self):
        return len(self.batches)

    def _stage(self, slot, pos, t):
        """Pinned staging copy of a CPU tensor (re-used buffers, one set per slot)."""
        if t.is_cuda:
            return t
        if t.is_floating_point() and t.dtype != self.dtype:
            t = t.to(self.dtype)
        if t.is_pinned():
            return t
        key = (slot, pos, tuple(t.shape), t.dtype)
        buf = self._pinned.get(key)
        if buf is None:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[key] = buf
        buf.copy_(t)
        return buf

    def _issue(self, slot, batch):
        if torch.is_tensor(batch):
            batch = (batch,)
        out = []
        with torch.cuda.stream(self.stream):
            for pos, t in enumerate(batch):
                if pos in self.skip or not torch.is_tensor(t):
                    out.append(t)
                else:
                    out.append(self._stage(slot, pos, t).to(self.device, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self):
        it = iter(self.batches)
        slot = 0
        try:
            nxt = self._issue(slot, next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            slot ^= 1
            try:
                nxt = self._issue(slot, next(it))      # next batch's copy overlaps this batch's compute
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in cur:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(torch.cuda.current_stream(self.device))
            yield tuple(cur)
