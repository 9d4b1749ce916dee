"""The reference's train step (train.py:68-73) captured ONCE into a CUDA graph and replayed.

The small BASELINE configs (27-127 sensors) are launch-bound: ~30 kernel launches plus the optimizer
for 0.3-0.7 ms of device work.  Every kernel of the path is enqueued on the caller's stream without
host synchronisation and all scratch comes from the torch allocator, so the whole step --
graph build, forward, MSE, backward, fused Adam -- is capturable.  Dropout stays fresh across replays
because the Philox offset is read from a device counter the graph itself increments
(`gdn_dropout.offset_dev`).
"""
import torch

from gdn_b200 import ops


def accumulators_follow(params, stream):
    """True when the gradient accumulator of every parameter would run on `stream` if a forward were made on it now.
    An accumulator node is created on first use and bound to the stream current at that moment; it lives as long as
    any autograd graph references it.  One that an older graph -- the previous step's loss, an output of the caller's
    own forward, still referenced -- keeps alive stays bound to the (default) stream of that forward, and a backward
    inside a capture would then touch the legacy stream: CUDA refuses the capture (cudaErrorStreamCaptureImplicit),
    and a refused capture leaves torch's allocator and generator in capture state.  So a capture is only attempted
    when this probe -- a zero-gradient backward through views of the parameters, with a pre-hook reading the stream
    each accumulator runs on -- comes back clean."""
    params = [p for p in params if p.requires_grad]
    seen = []
    cur = torch.cuda.current_stream(stream.device)
    stream.wait_stream(cur)
    with torch.cuda.stream(stream):
        views = [p.view_as(p) for p in params]
        hooks = [v.grad_fn.next_functions[0][0].register_prehook(
            lambda grads: seen.append(torch.cuda.current_stream().cuda_stream)) for v in views]
        try:
            torch.autograd.backward(views, [torch.zeros_like(p) for p in params])
        finally:
            for h in hooks:
                h.remove()
    cur.wait_stream(stream)
    del views
    return len(seen) == len(params) and all(sid == stream.cuda_stream for sid in seen)


class GraphedTrainStep:
    def __init__(self, model, batch_shape, lr=1e-3, weight_decay=0.0, warmup=3):
        B, N, W = batch_shape
        params = list(model.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device")
        self.model = model
        self.x = torch.zeros(B, N, W, device=dev)
        self.y = torch.zeros(B, N, device=dev)
        self.opt = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, fused=True, capturable=True)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.loss = None
        model.train()
        # warm-up on a side stream (allocator pools, lazy optimizer state, kernel attributes), then put
        # parameters, BatchNorm buffers and optimizer state back: the warm-up must not train
        saved = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ops.set_dropout_counter(self.counter)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self.opt.zero_grad(set_to_none=True)
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(saved[k])
            for st in self.opt.state.values():
                for t in st.values():
                    if torch.is_tensor(t):
                        t.zero_()
            self.counter.zero_()
        if not accumulators_follow(params, side):
            ops.set_dropout_counter(None)
            raise RuntimeError("GraphedTrainStep: a tensor with grad_fn from an earlier forward of this model is still "
                               "referenced (its gradient accumulators are bound to another stream); release it before "
                               "capturing the train step")
        self.opt.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            loss = self._body()
        self.loss = loss.detach()        # static storage of the graph; the captured autograd graph itself is released
        del loss
        ops.set_dropout_counter(None)

    def _body(self):
        self.counter.add_(1)
        out = self.model(self.x, None)
        loss = torch.nn.functional.mse_loss(out, self.y, reduction="mean")
        loss.backward()
        self.opt.step()
        return loss

    def step(self, x, y):
        """Copy the batch into the graph's static buffers and replay; returns the (device) loss."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        # the replay moves embedding.weight on the device without touching its version counter: an eval-mode
        # graph cached before this step (GDN.build_graph keys on data_ptr + _version) is stale now
        self.model._graph_cache = None
        return self.loss
