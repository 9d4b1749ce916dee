"""Flat-buffer Adam for the data-parallel train step (train.py:31, 73).

`torch.optim.Adam` semantics (amsgrad off, L2 weight decay) on four flat float32 buffers: the model's parameters
are re-pointed to views of one buffer, their `.grad`s to views of another, so autograd accumulates straight into the
buffer NCCL all-reduces in place, and ONE kernel (`gdn_adam_flat`) applies the 1/world scaling and the update --
instead of `torch.cat` -> all_reduce -> copy back into every `p.grad` -> a multi-tensor Adam."""
import torch

from . import _lib
from ._lib import check, ptr


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise RuntimeError("FlatAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError("FlatAdam needs float32 parameters on one CUDA device (gdn_b200 has no CPU path)")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.step_count = 0
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad_buffer = torch.zeros(n, dtype=torch.float32, device=dev)      # what the all-reduce works on
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)                      # the module keeps its Parameters
                p.grad = self.grad_buffer[off:off + k].view_as(p)               # autograd accumulates in place
                off += k

    def zero_grad(self):
        """One memset; the `.grad` views stay attached (never `set_to_none`)."""
        self.grad_buffer.zero_()
        off = 0
        for p in self.params:                                                   # re-attach if something detached a view
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad_buffer.data_ptr() + 4 * off:
                p.grad = self.grad_buffer[off:off + k].view_as(p)
            off += k

    def step(self, grad_scale=1.0):
        lib = _lib.load()
        self.step_count += 1
        check(lib.gdn_adam_flat(ptr(self.flat), ptr(self.grad_buffer), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                self.step_count, float(grad_scale), torch.cuda.current_stream().cuda_stream), "gdn_adam_flat")
