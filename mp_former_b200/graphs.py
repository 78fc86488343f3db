"""CUDA-graph capture of one training step of the hot path (forward + backward).

The path launches ~2,600 kernels per step at the bench geometry; issued eagerly from Python that is ~200 ms of host
time per step, which bounds the step once the kernels themselves are faster than that.  ``GraphedStep`` records
forward, loss and backward ONCE into a ``torch.cuda.CUDAGraph`` (tensor maps, launch parameters and the allocator's
addresses are frozen at capture time) and replays it per step: inputs are copied into static buffers, gradients
appear in the parameters' static ``.grad`` tensors.

Requirements on the captured callable: static shapes, no host synchronisation, no pageable host<->device copies
(the decoder keeps the host-derived DN index tensors in ``dn_args`` for this reason, see
``MultiScaleMaskedTransformerDecoderMaskDN._dn_indices``).
"""
import torch

from . import _lib


class GraphedStep:
    def __init__(self, step_fn, example_inputs, params, warmup=3, flat_grads=True):
        """``step_fn(inputs: dict[str, Tensor]) -> loss`` (scalar tensor); ``params``: parameters whose ``.grad`` the
        step produces.  ``example_inputs`` fixes shapes / dtypes / device.

        ``flat_grads``: after every replay the parameters' ``.grad`` tensors are views of ONE flat buffer
        (``self.flat_grad``), so the data-parallel exchange is a single in-place all-reduce of that buffer with no
        flatten / unflatten copies around it (DistributedDataParallel's ``gradient_as_bucket_view``).  The captured
        step computes every gradient into its own static tensor (``.grad`` is None at capture time, so autograd's
        AccumulateGrad stores instead of adding: no zero-fill and no add launch per parameter, ~170 launches per step
        for this head) and ends with one concatenation into the flat buffer."""
        self.params = [p for p in params if p.requires_grad]
        self.static_in = {k: v.detach().clone() for k, v in example_inputs.items()}
        dev = next(iter(self.static_in.values())).device
        self.flat_grad = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up off the default stream, as capture requires
            for _ in range(max(1, warmup)):
                for p in self.params:
                    p.grad = None
                step_fn(self.static_in).backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.params:
            p.grad = None
        flat = bool(flat_grads and self.params and all(p.dtype == self.params[0].dtype for p in self.params))
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_loss = step_fn(self.static_in)
            self.static_loss.backward()
            if flat:
                self.flat_grad = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                                            for p in self.params])
        self.launches_per_replay = _lib.launch_count() - n0     # kernels of this library inside the graph
        if flat:
            off = 0
            for p in self.params:
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
                off += p.numel()

    def load_inputs(self, inputs, non_blocking=True):
        for k, v in inputs.items():
            self.static_in[k].copy_(v, non_blocking=non_blocking)

    def replay(self):
        """Runs the captured step on the current stream; returns the (static) loss tensor."""
        self.graph.replay()
        return self.static_loss

    def __call__(self, inputs=None):
        if inputs is not None:
            self.load_inputs(inputs)
        return self.replay()


def allreduce_gradients(params, world_size, flat=None):
    """Data-parallel gradient exchange of the path (its only collective): ONE flat all-reduce of all
    gradients, averaged over ranks like DistributedDataParallel.  Used after a graph replay, where DDP's autograd
    hooks do not run.  ``flat``: the buffer the gradients are views of (``GraphedStep.flat_grad``): reduced in place."""
    import torch.distributed as dist
    if world_size == 1:
        return
    if flat is not None:
        dist.all_reduce(flat)
        flat.div_(world_size)
        return
    grads = [p.grad for p in params if p.grad is not None]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat)
    flat.div_(world_size)
    torch._foreach_copy_(grads, list(torch._utils._unflatten_dense_tensors(flat, grads)))
