"""Data-parallel plumbing for the training step: batch sharding and ONE flat fp32 gradient
all-reduce per step over torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests).

The reference wraps the model in DistributedDataParallel with default 25 MB buckets
(main_utils.py:343-346: 21.4 M trainable fp32 parameters = 85.7 MB = 4 all-reduces per step) and shards
the batch with DistributedSampler (main_utils.py:229).  Here all gradients live in one contiguous buffer
(`param.grad` tensors are views into it, so there is no flatten / unflatten copy) and a step issues a single
all-reduce — on NVSwitch the cost is launch latency, not link count, so one 85.7 MB message beats four
25 MB ones.  Inference needs no collective at all: scenes are independent (every kernel indexes
blockIdx = scene), so rank r simply owns scenes [r*B, (r+1)*B).
"""
import weakref

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced partition of range(n_items): the first n_items % world ranks get one extra."""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def broadcast_parameters(module, src=0):
    """Rank `src`'s parameters and buffers to everyone (what the DDP constructor does; the reference passes
    broadcast_buffers=False for the per-step sync, main_utils.py:345, but still needs identical initial state)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


class FlatGradients:
    """One contiguous fp32 gradient buffer for `module`'s trainable parameters.

        fg = FlatGradients(model)
        loss.backward()              # autograd accumulates straight into views of fg.flat
        fg.all_reduce_mean()         # ONE collective; grads now hold the mean over ranks
        optimizer.step(); fg.zero()
    """

    def __init__(self, module, fused_weight_grads=True):
        """fused_weight_grads: let the backward kernels accumulate weight gradients of THESE parameters straight into
        these buffers from a side stream instead of returning them to autograd.  The parameters are tagged with a weak
        reference to this object (nothing process-wide: other models in the process are untouched, and the tag dies
        with the object); the side stream is joined at the end of every backward pass, so `.grad` is complete on the
        current stream after `loss.backward()` returns."""
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise RuntimeError("FlatGradients: module has no trainable parameters")
        dev = self.params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError("FlatGradients: all trainable parameters must be fp32 on one device")
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self.fused = bool(fused_weight_grads and dev.type == "cuda")
        if self.fused:
            ref = weakref.ref(self)
            for p in self.params:
                p._eda_fused_grad_owner = ref

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()

    def release(self):
        """Switches fused weight-gradient accumulation off for this bucket's parameters: afterwards the backward
        kernels return their gradients to autograd like any other op."""
        self.sync()
        self.fused = False
        for p in self.params:
            if getattr(p, "_eda_fused_grad_owner", None) is not None:
                del p._eda_fused_grad_owner

    def sync(self):
        """Gradients are complete on the current stream after this (joins the side-stream weight-gradient kernels)."""
        if self.flat.is_cuda:
            from . import attn_ops
            attn_ops.join_wgrad()

    def check_views(self):
        """True while every param.grad is still a view into the flat buffer (an optimizer's
        zero_grad(set_to_none=True) would break that: call fg.zero() instead)."""
        base = self.flat.data_ptr()
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != base + off * 4:
                return False
            off += p.numel()
        return True

    def all_reduce_mean(self, async_op=False):
        """Sum over ranks then divide by the world size.  Returns the work handle when async_op."""
        self.sync()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        world = dist.get_world_size()
        if not self.check_views():
            raise RuntimeError("FlatGradients: a param.grad no longer aliases the flat buffer")
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            return _Scaled(work, self.flat, 1.0 / world)
        self.flat.mul_(1.0 / world)
        return None


class _Scaled:
    def __init__(self, work, flat, scale):
        self.work, self.flat, self.scale = work, flat, scale

    def wait(self):
        self.work.wait()
        self.flat.mul_(self.scale)
