"""Host staging converter of the feed (include/gdn_b200.h: gdn_stage_f64_to_f32): replaces
train.py:63-66 `x.float()`.  Pure host code -- runs without a GPU."""
import threading

import numpy as np
import pytest
import torch

from gdn_b200 import _lib


@pytest.mark.parametrize("threads", [1, 2, 5, 8, 64, 1000])
def test_stage_matches_float_cast_bit_for_bit(threads):
    lib = _lib.load()
    g = torch.Generator().manual_seed(threads)
    for n in (0, 1, 15, 16, 17, 65535, 65536, 65537, 1000003):
        src = (torch.rand(n + 3, generator=g, dtype=torch.float64) * 2e3 - 1e3)[3:]        # 8-byte aligned only
        dst = torch.full((n + 5,), 7.0, dtype=torch.float32)
        rc = lib.gdn_stage_f64_to_f32(src.data_ptr() if n else None, dst[1:].data_ptr() if n else None, n, threads)
        assert rc == 0
        assert torch.equal(dst[1:n + 1], src.float())
        assert dst[0] == 7.0 and (dst[n + 1:] == 7.0).all()                               # nothing outside [0, n)


def test_stage_special_values_round_like_the_reference_cast():
    lib = _lib.load()
    vals = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e300, -1e300, 1e-300, 3.4028235677973366e38,
                     1.0 + 2.0 ** -24, 1.0 + 2.0 ** -24 + 2.0 ** -50, 1.0 + 3 * 2.0 ** -24, 2.0 ** -149, 2.0 ** -150],
                    dtype=np.float64)
    src = torch.from_numpy(np.tile(vals, 8192))                     # long enough for the vector path and the pool
    dst = torch.empty(src.numel(), dtype=torch.float32)
    assert lib.gdn_stage_f64_to_f32(src.data_ptr(), dst.data_ptr(), src.numel(), 4) == 0
    want = src.float()
    assert torch.equal(dst.view(torch.int32), want.view(torch.int32))                    # bit patterns, NaN included


def test_stage_rejects_null_and_serialises_concurrent_callers():
    lib = _lib.load()
    assert lib.gdn_stage_f64_to_f32(None, None, 16, 2) == -1
    srcs = [torch.rand(1 << 20, dtype=torch.float64) for _ in range(4)]
    dsts = [torch.empty(1 << 20, dtype=torch.float32) for _ in range(4)]

    def run(k):
        for _ in range(5):
            assert lib.gdn_stage_f64_to_f32(srcs[k].data_ptr(), dsts[k].data_ptr(), srcs[k].numel(), 3 + k) == 0

    ths = [threading.Thread(target=run, args=(k,)) for k in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for s, d in zip(srcs, dsts):
        assert torch.equal(d, s.float())
