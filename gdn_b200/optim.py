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
        # the step count lives on the device: `step()` bumps it on the stream and the kernel derives the bias
        # corrections from it, so the same optimiser works eagerly and inside a captured CUDA graph
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
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

    @property
    def step_count(self):
        """Number of steps taken (reads the device counter: a host sync; graph replays count too)."""
        return int(self.step_dev.item())

    @step_count.setter
    def step_count(self, value):
        self.step_dev.fill_(int(value))

    def step(self, grad_scale=1.0):
        lib = _lib.load()
        self.step_dev.add_(1)
        check(lib.gdn_adam_flat_dev(ptr(self.flat), ptr(self.grad_buffer), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                    self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                    ptr(self.step_dev), float(grad_scale), torch.cuda.current_stream().cuda_stream),
              "gdn_adam_flat_dev")


class NvlsFlatAdam:
    """The data-parallel optimiser step over NVSwitch multicast memory (SURVEY §8 row f-4 as written): gradient
    all-reduce, Adam and the broadcast of the new parameters are ONE kernel (`gdn_nvls_adam`) instead of an NCCL
    all-reduce followed by an update on every rank.

    Plumbing (torch.distributed._symmetric_memory): the flat parameter and gradient buffers are symmetric allocations
    bound to multicast objects; the module's Parameters (and their `.grad`s) are views of them.  Rank r owns a 1/G
    slice: it pulls the switch-reduced gradient of that slice (`multimem.ld_reduce`), updates its slice of the Adam
    moments -- the optimiser state is sharded G-fold -- and stores the new parameters to every replica
    (`multimem.st`).  Two tiny cross-rank barriers (signal pads) bracket the kernel on the compute stream."""

    class Unavailable(RuntimeError):
        pass

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None):
        import torch.distributed as dist
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise RuntimeError("NvlsFlatAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError("NvlsFlatAdam needs float32 parameters on one CUDA device (gdn_b200 has no CPU path)")
        if not (dist.is_available() and dist.is_initialized()):
            raise self.Unavailable("torch.distributed is not initialised")
        try:
            import torch.distributed._symmetric_memory as symm
        except Exception as e:                                   # pragma: no cover
            raise self.Unavailable(f"torch symmetric memory is not importable: {e}")
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.step_count = 0
        n = sum(p.numel() for p in self.params)
        unit = 4 * self.world                                     # every slice a whole number of 16-byte vectors
        self.n, self.n_pad = n, -(-n // unit) * unit
        self.slice = self.n_pad // self.world
        self.lo = self.rank * self.slice
        try:
            if hasattr(symm, "enable_symm_mem_for_group"):      # needed by older torch builds, a no-op (that warns) on newer
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    try:
                        symm.enable_symm_mem_for_group(group.group_name)
                    except Exception:
                        pass
            self.flat = symm.empty(self.n_pad, dtype=torch.float32, device=dev)
            self.grad_buffer = symm.empty(self.n_pad, dtype=torch.float32, device=dev)
            self.h_param = symm.rendezvous(self.flat, group)
            self.h_grad = symm.rendezvous(self.grad_buffer, group)
            self.p_mc, self.g_mc = int(self.h_param.multicast_ptr), int(self.h_grad.multicast_ptr)
        except Exception as e:
            raise self.Unavailable(f"symmetric memory rendezvous failed: {type(e).__name__}: {e}")
        if self.p_mc == 0 or self.g_mc == 0:
            raise self.Unavailable("this box / driver exposes no NVLink multicast (multicast_ptr == 0)")
        self.exp_avg = torch.zeros(self.slice, dtype=torch.float32, device=dev)       # this rank's slice only
        self.exp_avg_sq = torch.zeros(self.slice, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            self.flat.zero_()
            self.grad_buffer.zero_()
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.grad_buffer[off:off + k].view_as(p)
                off += k
        torch.cuda.synchronize(dev)
        self.h_param.barrier(channel=0)

    def zero_grad(self):
        self.grad_buffer.zero_()
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad_buffer.data_ptr() + 4 * off:
                p.grad = self.grad_buffer[off:off + k].view_as(p)
            off += k

    def step(self, grad_scale=None):
        lib = _lib.load()
        self.step_count += 1
        scale = 1.0 / self.world if grad_scale is None else float(grad_scale)
        self.h_grad.barrier(channel=0)                            # every rank's backward has written its gradients
        check(lib.gdn_nvls_adam(ptr(self.flat), self.p_mc, self.g_mc, ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                self.lo, self.slice, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                self.step_count, scale, torch.cuda.current_stream().cuda_stream), "gdn_nvls_adam")
        self.h_param.barrier(channel=1)                           # every slice of the new parameters has landed
