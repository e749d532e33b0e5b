"""Data-parallel plumbing for the training step: batch sharding and ONE flat fp32 gradient
all-reduce per step over torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests).

The reference wraps the model in DistributedDataParallel with default 25 MB buckets
(main_utils.py:343-346: 21.4 M trainable fp32 parameters = 85.7 MB = 4 all-reduces per step) and shards
the batch with DistributedSampler (main_utils.py:229).  Here all gradients live in one contiguous buffer
(`param.grad` tensors are views into it, so there is no flatten / unflatten copy) and a step issues a single
all-reduce — on NVSwitch the cost is launch latency, not link count, so one 85.7 MB message beats four
25 MB ones.  Inference needs no collective at all: scenes are independent (every kernel indexes
blockIdx = scene), so rank r simply owns scenes [r*B, (r+1)*B).

Overlap (what DistributedDataParallel's bucketing buys the reference, main_utils.py:343-346): `enable_overlap()`
splits the flat buffer into a few contiguous REGIONS in backward order (by default one per top-level child module:
decoder -> encoder -> backbone for the hot path).  Every parameter reports when its gradient has been written — the
fused weight-gradient kernels through `attn_ops.grads_written`, everything else through a post-accumulate-grad hook —
and the moment a region is complete its all-reduce (AVG: the 1/world scaling rides in the collective) is issued on a
communication stream that waits only for the streams that wrote it, while the backward pass continues underneath.
Whatever is still open at the end of the backward pass is flushed there; `finish()` joins the communication stream.
The whole mechanism is capturable, so GraphedTrainStep records the collectives inside the step's CUDA graph.

Synchronised BatchNorm (main_utils.py:335-338) lives in eda_b200/syncbn.py; `convert_sync_batchnorm` is re-exported here.
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
        ref = weakref.ref(self)
        if self.fused:
            for p in self.params:
                p._eda_fused_grad_owner = ref
        self._module = module
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._offsets = []
        off = 0
        for p in self.params:
            self._offsets.append(off)
            off += p.numel()
        self.regions = None      # [(first param index, end param index)] in backward order once overlap is enabled
        self._hooks = []

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()
        if self.regions is not None:
            self._begin_step()

    # ------------------------------------------------------------------------------------------
    # overlapped, bucketed all-reduce
    # ------------------------------------------------------------------------------------------
    def enable_overlap(self, groups=None, group=None):
        """Splits the bucket into regions and arms the per-parameter completion tracking.  `groups`: submodules in the
        order their backward passes FINISH (default: the top-level children that own parameters, reversed); the
        parameters of each must be contiguous in module.parameters() order.  No-op without an initialised process group
        of more than one rank.  Returns self."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self
        if groups is None:
            groups = [c for c in self._module.children() if any(p.requires_grad for p in c.parameters())][::-1]
        regions, seen = [], set()
        for g in groups:
            idx = sorted(self._index[id(p)] for p in g.parameters() if id(p) in self._index and id(p) not in seen)
            if not idx:
                continue
            if idx != list(range(idx[0], idx[-1] + 1)):
                raise RuntimeError("FlatGradients.enable_overlap: a group's parameters are not contiguous in the bucket")
            seen.update(id(self.params[i]) for i in idx)
            regions.append((idx[0], idx[-1] + 1))
        rest = [i for i, p in enumerate(self.params) if id(p) not in seen]
        if rest:  # parameters outside every group (e.g. owned by the root module itself): one more region, flushed last
            if rest != list(range(rest[0], rest[-1] + 1)):
                raise RuntimeError("FlatGradients.enable_overlap: ungrouped parameters are not contiguous")
            regions.append((rest[0], rest[-1] + 1))
        self.regions = regions
        self._region_of = {}
        for r, (a, b) in enumerate(regions):
            for i in range(a, b):
                self._region_of[i] = r
        self._group = group
        self._avg = dist.get_backend(group) == "nccl"
        self._comm = torch.cuda.Stream(device=self.flat.device) if self.flat.is_cuda else None
        ref = weakref.ref(self)
        for i, p in enumerate(self.params):
            def hook(param, i=i, ref=ref):
                me = ref()
                if me is not None:
                    me._mark(i)
            self._hooks.append(p.register_post_accumulate_grad_hook(hook))
        self._begin_step()
        return self

    def _begin_step(self):
        self._pending = [b - a for a, b in self.regions]
        self._done = [False] * len(self.params)
        self._launched = [False] * len(self.regions)
        self._works = []
        self._flush_queued = False
        self._expected = {}   # parameter index -> fused backward contributions still to come (counted in the forward)
        self._streams = [set() for _ in self.regions]  # streams on which a region's gradient kernels were queued

    def _mark(self, i):
        """Parameter i's gradient has been written (its kernels are queued on the current or the side stream)."""
        if self.regions is None or self._done[i]:
            return
        self._done[i] = True
        r = self._region_of[i]
        self._pending[r] -= 1
        if self._comm is not None:
            self._streams[r].add(torch.cuda.current_stream(self.flat.device))
        if not self._flush_queued:
            try:  # regions still open at the end of this backward pass are flushed there
                torch.autograd.Variable._execution_engine.queue_callback(self._flush)
                self._flush_queued = True
            except RuntimeError:
                pass
        if self._pending[r] == 0:
            self._launch(r)

    def expect(self, p):
        i = self._index.get(id(p))
        if i is not None:
            self._expected[i] = self._expected.get(i, 0) + 1

    def written(self, p):
        i = self._index.get(id(p))
        if i is None:
            return
        left = self._expected.get(i, 1) - 1
        self._expected[i] = left
        if left <= 0:
            self._mark(i)

    def _launch(self, r):
        if self._launched[r]:
            return
        self._launched[r] = True
        a, b = self.regions[r]
        lo = self._offsets[a]
        hi = self._offsets[b - 1] + self.params[b - 1].numel()
        view = self.flat[lo:hi]
        world = dist.get_world_size(self._group)
        if self._comm is None:  # CPU bucket (gloo in the tests): synchronous
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self._group)
            view.mul_(1.0 / world)
            return
        dev = self.flat.device
        # the collective must follow every stream that carries gradient kernels of this region: the streams the
        # completion reports came from (main stream, the encoder's language-branch stream), the current one, and the
        # side stream of the fused weight-gradient kernels
        waits = set(self._streams[r])
        waits.add(torch.cuda.current_stream(dev))
        if self.fused:
            from . import attn_ops
            side = attn_ops._wgrad_streams.get(dev.index if dev.index is not None else torch.cuda.current_device())
            if side is not None:
                waits.add(side)
        for st in waits:
            self._comm.wait_stream(st)
        with torch.cuda.stream(self._comm):
            if self._avg:
                work = dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self._group, async_op=True)
            else:
                work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self._group, async_op=True)
            self._works.append((work, None if self._avg else (view, 1.0 / world)))

    def _flush(self):
        self._flush_queued = False
        if self.regions is None:
            return
        for r in range(len(self.regions)):
            self._launch(r)

    def finish(self):
        """The current stream waits for every region's all-reduce (after flushing regions that are still open)."""
        if self.regions is None:
            return
        self._flush()
        if self._comm is not None:
            with torch.cuda.stream(self._comm):
                for work, post in self._works:
                    work.wait()
                    if post is not None:
                        post[0].mul_(post[1])
            torch.cuda.current_stream(self.flat.device).wait_stream(self._comm)
        self._works = []

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
        if self.regions is not None:  # overlapped mode: the regions were (or are now) reduced on the comm stream
            self.finish()
            return None
        world = dist.get_world_size()
        if not self.check_views():
            raise RuntimeError("FlatGradients: a param.grad no longer aliases the flat buffer")
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            return _Scaled(work, self.flat, 1.0 / world)
        self.flat.mul_(1.0 / world)
        return None


def convert_sync_batchnorm(module, process_group=None):
    """SyncBatchNorm semantics for every BatchNorm of `module` (see eda_b200/syncbn.py)."""
    from . import syncbn

    return syncbn.convert_sync_batchnorm(module, process_group)


class _Scaled:
    def __init__(self, work, flat, scale):
        self.work, self.flat, self.scale = work, flat, scale

    def wait(self):
        self.work.wait()
        self.flat.mul_(self.scale)
