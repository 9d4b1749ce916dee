"""Host -> device batch feed for the reference's train/test loops (train.py:63-66, test.py:43-44).

The reference moves every batch with a blocking `.to(device)` inside the step, so the copy and the
compute serialise (and it ships the unused fully-connected `edge_index` every step).  `Prefetcher`
wraps any iterable of `(x, y, ...)` CPU batches -- e.g. the reference's DataLoader over `TimeDataset`
-- and keeps ONE batch in flight on a copy stream: pinned staging buffers, `non_blocking` H2D copies,
an event per batch, and the consumer stream waits on that event only.  Tensors whose name position is
listed in `skip` (default: the 4th, `edge_index`) are passed through untouched.
"""
import queue
import threading

import torch


class LossReader:
    """train.py:76-77 reads `loss.item()` after every step: a device synchronisation per step, which serialises the
    host side of step k+1 (feed, launches) behind the device side of step k.  This reader keeps the read -- 4 bytes
    device -> pinned host every step -- but hands the value back one step late: `push(loss)` enqueues the copy of
    step k and returns the value of step k-1 (already on the host by then), `flush()` returns what is still in
    flight.  The sequence of values is exactly what `.item()` per step would have produced."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = int(depth)
        self.host = torch.zeros(self.depth, dtype=torch.float32, pin_memory=True)
        self.events = [None] * self.depth
        self.count = 0

    def _take(self, k):
        self.events[k % self.depth].synchronize()
        return float(self.host[k % self.depth])

    def push(self, loss):
        k = self.count
        out = self._take(k - self.depth + 1) if k >= self.depth - 1 else None     # the oldest value still in flight
        slot = k % self.depth
        self.host[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.events[slot] = ev
        self.count += 1
        return out

    def flush(self):
        """The values still in flight, oldest first; the reader is empty afterwards (ready for the next epoch)."""
        first = max(0, self.count - self.depth + 1)
        out = [self._take(k) for k in range(first, self.count)]
        self.events = [None] * self.depth
        self.count = 0
        return out


class Prefetcher:
    def __init__(self, batches, device, skip=(3,), dtype=torch.float32, reuse_buffers=False, threaded=False,
                 stage_threads=None, depth=2):
        """reuse_buffers=True copies into two persistent device buffers per tensor position (no allocator
        traffic); a yielded batch is then only valid until the NEXT-BUT-ONE batch is requested -- right
        for train.py's loop, wrong for a consumer that keeps aliases of its inputs (test.py:55-57 keeps the
        first `y`), hence off by default."""
        self.batches = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Prefetcher feeds a CUDA device (gdn_b200 has no CPU path)")
        if self.device.index is None:                   # "cuda": the worker thread needs the index to bind itself
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.skip = set(skip)
        self.dtype = dtype
        self.reuse = bool(reuse_buffers)
        # threaded=True: staging (dtype cast into pinned memory: the reference's loader yields pageable float64,
        # datasets/TimeDataset.py:64-73) and the H2D copy of the NEXT batches run on a worker thread, so they overlap
        # the consumer's device work even when the consumer blocks on `loss.item()` every step
        # threaded="auto": by the size of the first batch -- below AUTO_THREAD_BYTES of host data per batch the hand-off
        # between two Python threads costs more than the staging it hides (measured at 27..127 sensors: 0.12 vs 0.16 ms
        # per step, with occasional 0.4 ms convoys on the interpreter lock)
        self.threaded = threaded if threaded == "auto" else bool(threaded)
        # stage_threads > 1: a large staging copy (with its dtype cast) is split by rows over that many plain host
        # threads, each running its slice single-threaded -- independent of OMP_NUM_THREADS (torch.distributed.run
        # pins it to 1 per rank, which would leave one thread to convert ~140 MB per step at the largest config) and
        # without an OpenMP team whose idle workers spin on the cores the launching thread needs
        # float64 -> float32 (the reference's loader) goes through the library's own converter (gdn_stage_f64_to_f32:
        # persistent native thread pool, AVX-512 + non-temporal stores, no interpreter lock); None = up to 8 threads,
        # shared fairly between the ranks of one host; 1 = leave the copy to torch (its OpenMP team, if it has one)
        if stage_threads is None:
            import os
            local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
            stage_threads = max(2, min(8, (os.cpu_count() or 2) // local))
        self.stage_threads = max(1, int(stage_threads))
        self._pool = None
        self.depth = max(2, int(depth))      # batches in flight in the threaded mode (staging + device buffers per slot)
        self.stream = torch.cuda.Stream(device=self.device)
        self._pinned = {}
        self._devbuf = {}
        self._copied = {}        # slot -> event after the slot's last H2D copies (guards the pinned staging buffers)

    def __len__(self):
        return len(self.batches)

    def _stage(self, slot, pos, t):
        """Pinned staging copy of a CPU tensor (re-used buffers, one set per slot).  A floating-point tensor of another
        dtype (the reference's loader yields float64) is converted BY the copy into the pinned buffer: one pass over
        the data, no intermediate tensor."""
        if t.is_cuda:
            return t
        want = self.dtype if (t.is_floating_point() and t.dtype != self.dtype) else t.dtype
        if t.is_pinned() and want == t.dtype:
            return t
        key = (slot, pos, tuple(t.shape), want)
        buf = self._pinned.get(key)
        if buf is None:
            buf = torch.empty(t.shape, dtype=want, pin_memory=True)
            self._pinned[key] = buf
        else:
            # the slot's previous H2D copy out of this buffer is only stream-ordered; a consumer without a
            # per-step host sync lets the host run ahead, so wait for that copy HERE before overwriting its source
            ev = self._copied.get(slot)
            if ev is not None:
                ev.synchronize()
        self._host_copy(buf, t)                     # converts while copying
        return buf

    def _host_copy(self, dst, src):
        if self.stage_threads > 1 and src.dtype == torch.float64 and dst.dtype == torch.float32 and not src.is_cuda \
                and src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel():
            from . import _lib
            rc = _lib.load().gdn_stage_f64_to_f32(src.data_ptr(), dst.data_ptr(), src.numel(), self.stage_threads)
            if rc != 0:
                raise RuntimeError("gdn_stage_f64_to_f32 failed")
            return
        n = src.shape[0] if src.dim() > 0 else 0
        k = min(self.stage_threads, n)
        if k <= 1 or src.numel() < (1 << 20):
            dst.copy_(src)
            return
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.stage_threads, thread_name_prefix="gdn-stage",
                                            initializer=torch.set_num_threads, initargs=(1,))   # per-thread OpenMP setting
        bounds = [n * i // k for i in range(k + 1)]
        list(self._pool.map(lambda i: dst[bounds[i]:bounds[i + 1]].copy_(src[bounds[i]:bounds[i + 1]]), range(k)))

    def _to_device(self, slot, pos, h):
        if h.is_cuda:
            return h
        if not self.reuse:
            return h.to(self.device, non_blocking=True)
        key = (slot, pos, tuple(h.shape), h.dtype)
        buf = self._devbuf.get(key)
        if buf is None:
            buf = torch.empty(h.shape, dtype=h.dtype, device=self.device)
            self._devbuf[key] = buf
        buf.copy_(h, non_blocking=True)
        return buf

    def _issue(self, slot, batch, after=None):
        if torch.is_tensor(batch):
            batch = (batch,)
        out = []
        with torch.cuda.stream(self.stream):
            if after is not None:
                self.stream.wait_event(after)          # the consumer's work on this slot's previous batch
            for pos, t in enumerate(batch):
                if pos in self.skip or not torch.is_tensor(t):
                    out.append(t)
                else:
                    out.append(self._to_device(slot, pos, self._stage(slot, pos, t)))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._copied[slot] = ev
        return out, ev

    def _iter_threaded(self, batches):
        cur_stream = torch.cuda.current_stream(self.device)
        free_q, ready_q = queue.Queue(), queue.Queue()
        for slot in range(self.depth):
            free_q.put((slot, None))
        stop = threading.Event()

        def work():
            try:
                torch.cuda.set_device(self.device)
                for batch in batches:
                    while True:
                        try:
                            slot, after = free_q.get(timeout=0.2)
                            break
                        except queue.Empty:
                            if stop.is_set():
                                return
                    out, ev = self._issue(slot, batch, after)
                    ready_q.put((slot, out, ev))
                ready_q.put(None)
            except BaseException as e:                      # surfaces in the consumer
                ready_q.put(e)

        th = threading.Thread(target=work, name="gdn-prefetch", daemon=True)
        th.start()
        try:
            while True:
                item = ready_q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                slot, cur, ev = item
                cur_stream.wait_event(ev)
                if not self.reuse:
                    for t in cur:
                        if torch.is_tensor(t) and t.is_cuda:
                            t.record_stream(cur_stream)
                yield tuple(cur)
                done = None
                if self.reuse:
                    done = torch.cuda.Event()
                    done.record(cur_stream)
                free_q.put((slot, done))
        finally:
            stop.set()

    AUTO_THREAD_BYTES = 8 << 20

    @staticmethod
    def _host_bytes(batch):
        ts = (batch,) if torch.is_tensor(batch) else batch
        return sum(t.numel() * t.element_size() for t in ts if torch.is_tensor(t) and not t.is_cuda)

    def __iter__(self):
        batches = self.batches
        threaded = self.threaded
        if threaded == "auto":
            import itertools
            it0 = iter(batches)
            try:
                first = next(it0)
            except StopIteration:
                return
            threaded = self._host_bytes(first) >= self.AUTO_THREAD_BYTES
            batches = itertools.chain([first], it0)
        if threaded:
            yield from self._iter_threaded(batches)
            return
        it = iter(batches)
        cur_stream = torch.cuda.current_stream(self.device)
        done = [None, None]
        slot = 0
        try:
            nxt = self._issue(slot, next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            other = slot ^ 1
            try:
                nxt = self._issue(other, next(it), done[other])   # overlaps this batch's compute
            except StopIteration:
                nxt = None
            cur_stream.wait_event(ev)
            if not self.reuse:
                for t in cur:
                    if torch.is_tensor(t) and t.is_cuda:
                        t.record_stream(cur_stream)
            yield tuple(cur)
            if self.reuse:
                done[slot] = torch.cuda.Event()
                done[slot].record(cur_stream)
            slot = other
