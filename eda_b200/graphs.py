"""CUDA-graph capture of a forward pass built from the kernels of this package.

The hot path is ~25 (SA backbone stage pair) to ~160 (3 encoder + 6 decoder layers) short kernel
launches per forward; issued one by one from Python they are bound by host launch latency, not by the
GPU.  `GraphedCallable` records one invocation — every kernel of libeda_b200.so launches on the
current stream handed in through the C ABI, and the side-stream FPS chain of
`backbone_module.fps_chain` forks from / joins the capturing stream, so the whole thing is
capturable — and afterwards replays it with a single `cudaGraphLaunch`.

Inference only: packed attention weights are cached across calls (attn_ops.pack_weight), so a graph
recorded before an optimizer step would keep using the old packed copy.  Re-capture after changing
weights.  Shapes are fixed at capture time (static input buffers, like any CUDA graph).
"""
import ctypes

import torch

from . import _lib


_capture_streams = {}


def _capture_stream(device):
    """ONE warm-up / capture side stream per device for every graph this module records: autograd's AccumulateGrad nodes
    remember the stream they first ran on, and a later capture (another length bucket, a re-capture) that warmed up on
    a different stream would trip over that."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _capture_streams.get(key)
    if st is None:
        st = _capture_streams[key] = torch.cuda.Stream(device=device)
    return st


class GraphedCallable:
    """graphed = GraphedCallable(fn, example_inputs); out = graphed(*inputs)

    `fn` takes tensors positionally and returns a tensor or a (nested) tuple/list/dict of tensors.
    Inputs are copied into static buffers before every replay; the returned tensors are the graph's
    static outputs (valid until the next call — clone them if they must survive it)."""

    def __init__(self, fn, example_inputs, warmup=3):
        if not example_inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedCallable needs CUDA tensor inputs")
        self.device = example_inputs[0].device
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):  # fills caches (packed weights, smem attributes) outside the capture
                fn(*self.static_inputs)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_outputs = fn(*self.static_inputs)

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_inputs):
            raise RuntimeError("GraphedCallable: wrong number of inputs")
        for dst, src in zip(self.static_inputs, inputs):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise RuntimeError("GraphedCallable: input shape/dtype differs from the captured one")
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs


class GraphedTrainStep:
    """One training step (zero grads -> forward -> loss -> backward) recorded as ONE CUDA graph.

        fg = ddp.FlatGradients(model)
        step = GraphedTrainStep(model, loss_fn, example_inputs, fg)
        loss = step(*inputs)          # cudaGraphLaunch; gradients are in fg.flat (views: param.grad)
        fg.all_reduce_mean(); optimizer.step()

    Eagerly the step is ~900 kernel launches of 5-50 us each issued from Python through ctypes / autograd and is bound
    by that host work; replayed, only the device time remains.  Weight packing is captured too — one batched launch at the
    start of the step refreshes persistent packed copies of every weight (attn_ops.PackRegistry) — so replays after an
    optimizer step use the updated parameters; BatchNorm running statistics and `num_batches_tracked` are updated by captured device ops.
    Warm-up runs real steps on the example inputs; BatchNorm buffers (running statistics, num_batches_tracked) and the
    dropout epoch are snapshotted before and restored after it, so the first user step starts from the state the model
    had when it was handed in.  BatchNorm `momentum` is read on the host while recording and is FROZEN into the graph
    (a BNMomentumScheduler or momentum=None needs a re-capture).
    Restrictions: fixed shapes; parameters, gradients (FlatGradients) and inputs must stay at their addresses.  Dropout
    (the reference's default 0.1) is supported: the host-drawn seeds are frozen into the graph, and a device epoch word
    that the graph increments once per replay is added to them by the kernels (`self.dropout_epoch`)."""

    def __init__(self, model, loss_fn, example_inputs, flat_grads, warmup=3):
        from . import attn_ops

        if not example_inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedTrainStep needs CUDA tensor inputs")
        # Dropout: the per-call seeds are drawn on the host while the step is recorded and are frozen into the graph.
        # A device "epoch" word that the graph itself increments at the start of every replay is added to every seed by
        # the kernels (their `dropout_epoch` argument), so each replay draws fresh masks — the same ones in its forward and backward.
        has_dropout = False
        for m in model.modules():
            p = getattr(m, "p", None) if isinstance(m, torch.nn.Dropout) else getattr(m, "dropout", None) \
                if isinstance(m, torch.nn.MultiheadAttention) else None
            has_dropout |= bool(m.training and isinstance(p, float) and p > 0)
        if not flat_grads.check_views():
            raise RuntimeError("GraphedTrainStep: param.grad must alias the FlatGradients buffer")
        self.device = example_inputs[0].device
        self.static_inputs = [t.clone() for t in example_inputs]
        self.flat_grads = flat_grads
        self.dropout_epoch = torch.zeros(1, dtype=torch.int32, device=self.device) if has_dropout else None
        epoch = self.dropout_epoch

        # all packed weights (forward and transposed-for-backward) live in persistent buffers that one kernel launch
        # refreshes at the start of every step, instead of ~370 pack launches spread over the step
        self.pack_registry = attn_ops.PackRegistry(self.device)
        registry = self.pack_registry
        # FlatGradients.enable_overlap(): the bucketed gradient all-reduce is issued region by region DURING the
        # backward pass on a communication stream and is therefore part of the captured step
        self.overlapped_allreduce = flat_grads.regions is not None
        overlapped = self.overlapped_allreduce

        def step():
            if epoch is not None:
                epoch.add_(1)
            registry.refresh()
            flat_grads.zero()
            loss = loss_fn(model(*self.static_inputs))
            loss.backward()
            flat_grads.sync()  # side-stream weight-gradient kernels rejoin the (capturing) stream
            if overlapped:
                flat_grads.finish()  # ... and so does the communication stream: gradients are averaged over the ranks
            return loss.detach()

        registry.attach(model)  # per-model context: only THIS model's modules consult the registry
        registry.active = True
        # the epoch word is a per-model context too: the modules of this model hand it to every dropout-applying kernel
        # they launch (a plain argument of the C ABI: the library keeps no dropout state)
        for m in model.modules():
            if epoch is not None:
                m.__dict__["_eda_dropout_epoch"] = epoch
            else:
                m.__dict__.pop("_eda_dropout_epoch", None)
        buffers = [(b, b.detach().clone()) for b in model.buffers()]
        try:
            side = _capture_stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    step()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.loss = step()
            with torch.no_grad():  # undo the warm-up's side effects (capture itself executes nothing)
                for b, saved in buffers:
                    b.copy_(saved)
                if epoch is not None:
                    epoch.zero_()
                flat_grads.zero()
        finally:
            registry.active = False  # eager calls on the model go back to the version-keyed caches
            for m in model.modules():  # eager launches after this are unaffected; the graph keeps the baked pointer
                m.__dict__.pop("_eda_dropout_epoch", None)

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_inputs):
            raise RuntimeError("GraphedTrainStep: wrong number of inputs")
        for dst, src in zip(self.static_inputs, inputs):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise RuntimeError("GraphedTrainStep: input shape/dtype differs from the captured one")
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss


class BucketedTrainStep:
    """GraphedTrainStep for batches whose sequence length varies from step to step.

    The reference pads every text batch to ITS longest sentence (`padding="longest"`, models/bdetr.py:170), so the token
    dimension L changes per batch while CUDA graphs need fixed shapes.  This wrapper keeps one captured step per LENGTH
    BUCKET (L rounded up to a multiple of `multiple`, default 16: 32, 48, 64, 80, ...), captured lazily on first use, and
    pads the listed inputs up to the bucket before every replay:

        step = BucketedTrainStep(model, loss_fn, fg, pad={2: (1, 0.0), 3: (1, True)})   # input 2: text (B, L, E) padded
        loss = step(*inputs)                                                           # with zeros, input 3: its
                                                                                       # key-padding mask padded with True

    `pad` maps input index -> (dimension that carries L, fill value).  Padded positions must be marked as padding in the
    mask the model already takes (fill value True above), so they are ignored as attention keys; what the model outputs
    AT padded positions is garbage by construction (as for the reference's own padded tokens) and `loss_fn` must not
    read it — the reference's losses mask by the same attention mask.  All buckets share the model, the FlatGradients
    bucket and therefore the parameters; BatchNorm buffers and the dropout epoch are protected from the warm-up steps of
    a late capture exactly as in GraphedTrainStep."""

    def __init__(self, model, loss_fn, flat_grads, pad, multiple=16, max_buckets=8):
        self.model, self.loss_fn, self.flat_grads = model, loss_fn, flat_grads
        self.pad = dict(pad)
        self.multiple = int(multiple)
        self.max_buckets = int(max_buckets)
        self.steps = {}  # bucket length -> GraphedTrainStep

    def bucket_of(self, length):
        return ((int(length) + self.multiple - 1) // self.multiple) * self.multiple

    def _padded(self, inputs, bucket):
        out = list(inputs)
        for i, (dim, fill) in self.pad.items():
            t = out[i]
            cur = t.size(dim)
            if cur == bucket:
                continue
            shape = list(t.shape)
            shape[dim] = bucket
            p = torch.full(shape, fill, dtype=t.dtype, device=t.device)
            p.narrow(dim, 0, cur).copy_(t)
            out[i] = p
        return out

    def __call__(self, *inputs):
        lengths = {inputs[i].size(dim) for i, (dim, _) in self.pad.items()}
        if len(lengths) != 1:
            raise RuntimeError("BucketedTrainStep: the padded inputs disagree on the sequence length")
        bucket = self.bucket_of(lengths.pop())
        padded = self._padded(inputs, bucket)
        step = self.steps.get(bucket)
        if step is None:
            if len(self.steps) >= self.max_buckets:
                raise RuntimeError(f"BucketedTrainStep: more than {self.max_buckets} length buckets in use")
            step = self.steps[bucket] = GraphedTrainStep(self.model, self.loss_fn, padded, self.flat_grads)
        return step(*padded)
