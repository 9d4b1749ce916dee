"""Host staging bandwidth: pageable float64 [64,16384,16] -> pinned float32, on the main thread, on a worker thread,
and split by hand over a thread pool.  python tools/stage_bw.py"""
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

x = torch.rand(64, 16384, 16, dtype=torch.float64)
buf = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
print("cores", os.cpu_count(), "torch threads", torch.get_num_threads(), "OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"))


def timeit(fn, n=8):
    fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t) / n * 1e3


print("main thread copy_            %.2f ms" % timeit(lambda: buf.copy_(x)))
res = {}


def in_thread(setn):
    def run():
        if setn:
            torch.set_num_threads(setn)
        res["t"] = timeit(lambda: buf.copy_(x))
    th = threading.Thread(target=run)
    th.start()
    th.join()
    return res["t"]


print("worker thread copy_          %.2f ms" % in_thread(0))
print("worker thread, set_num_threads(%d) inside: %.2f ms" % (os.cpu_count(), in_thread(os.cpu_count())))
for k in (2, 4, 8, 16):
    pool = ThreadPoolExecutor(max_workers=k)
    n = x.shape[0]
    bounds = [n * i // k for i in range(k + 1)]
    def part(i):
        buf[bounds[i]:bounds[i + 1]].copy_(x[bounds[i]:bounds[i + 1]])
    print("thread pool of %2d chunks      %.2f ms" % (k, timeit(lambda: list(pool.map(part, range(k))))))
