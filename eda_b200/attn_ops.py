"""Host side of the attention-layer kernels (eda_linear_forward, eda_attention_forward): weight
packing cache, launch wrappers, and the autograd boundary of the three fused blocks the layers in
encoder_decoder_layers.py are made of:

    mha_block  : LayerNorm(residual + MHA(q_in [+ q_pos], k_in [+ k_pos], v_in, key_padding_mask))
    ffn_block  : LayerNorm(x + W2 relu(W1 x + b1) + b2)
    linear     : act(x W^T + b)

Forward = hand-written CUDA only (3 launches per mha_block, 2 per ffn_block); train-mode dropout (attention
probabilities, block outputs, FFN hidden) is applied inside those kernels from a counter-based hash and the
backward pass regenerates the same masks from the same hash.

Backward = hand-written CUDA as well (csrc/attn_bwd.cu, csrc/grad_ops.cu): LayerNorm backward, the attention core's
flash-style backward from the saved log-sum-exp, activation gradients through the forward tcgen05 GEMM with the
transposed weight, weight gradients by the split-row wgrad kernel — ~14 launches per attention block where autograd
through the reference's math path issues ~45.  EDA_BACKWARD=torch selects the older recompute-with-torch-ops backward
(kept as a cross-check for the tests).
There is no CPU path: CPU tensors raise RuntimeError like the rest of the package.
"""
import os
import ctypes
import math

import torch
import torch.nn.functional as F

from . import _lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev):
    # torch.cuda.current_stream() builds a Python Stream object through several pure-Python helpers (~5 us); the raw
    # handle is all the C ABI needs
    idx = dev.index
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(idx if idx is not None else torch.cuda.current_device()))


def _require_cuda(t, name="input"):
    if not t.is_cuda:
        raise RuntimeError(f"eda_b200: {name} must be a CUDA tensor (CPU not supported)")


# ---------------------------------------------------------------------------------------------------
# packed weights
# ---------------------------------------------------------------------------------------------------
# Packed-weight caches are keyed on the parameter's (data_ptr, _version): right for eager execution, wrong inside a
# CUDA graph of a TRAINING step (the graph must re-pack after every optimizer step).  graphs.GraphedTrainStep attaches
# a PackRegistry to the modules of ITS model and activates it while it warms up and captures; nothing here is
# process-wide, so two models (or two graphs) in one process do not interfere.
PACK_CACHE = True  # debugging switch only (False: re-pack on every call); never toggled by the library itself


class PackRegistry:
    """Persistent packed copies of a model's weights that ONE kernel launch refreshes (eda_linear_pack_batch).

    Used by graphs.GraphedTrainStep: while a registry is attached to the model and active, pack_weight / pack_weight_t
    calls that carry a cache key return the registry's buffer for that (module, name) — created and registered on
    first sight — without launching anything; `refresh()` at the start of every step re-packs all of them from the
    current parameter values.  Parameters must keep their storage (as in any CUDA-graph training loop)."""

    def __init__(self, device):
        self.device = device
        self.active = False  # True while the owning GraphedTrainStep warms up / captures
        self.entries = {}   # (id(owner), name, transposed) -> packed tensor
        self.owners = []    # keep the modules alive so ids stay unique
        self.rows = []      # descriptor rows: w ptr, dst ptr, stride_n, stride_k, N | K << 32, Kpad
        self.descs = None
        self.max_elements = 0

    def get(self, owner, name, transposed, W):
        key = (id(owner), name, transposed)
        hit = self.entries.get(key)
        if hit is not None:
            return hit
        lib = _lib.load()
        if transposed:   # packs W^T: N' = W.size(1), K' = W.size(0), element (n, k) = W[k, n]
            N, K, sn, sk = W.size(1), W.size(0), W.stride(1), W.stride(0)
        else:
            N, K, sn, sk = W.size(0), W.size(1), W.stride(0), W.stride(1)
        n = lib.eda_linear_packed_floats(N, K)
        if n == 0:
            raise RuntimeError(f"eda_b200.linear: unsupported weight shape ({N},{K})")
        packed = torch.empty(n, dtype=torch.float32, device=W.device)
        with torch.cuda.device(W.device):  # valid from the start; later steps rely on refresh()
            rc = lib.eda_linear_pack_strided(_p(W), int(sn), int(sk), N, K, _p(packed), _stream(W.device))
        _lib.check(rc, "linear_pack_strided")
        kpad = (K + 7) & ~7
        self.rows.append([W.data_ptr(), packed.data_ptr(), int(sn), int(sk), N | (K << 32), kpad])
        self.max_elements = max(self.max_elements, N * kpad)
        self.entries[key] = packed
        self.owners.append(owner)
        self.descs = None  # rebuilt by the next refresh()
        return packed

    def refresh(self):
        """Re-packs every registered weight on the current stream (one launch)."""
        if not self.rows:
            return
        if self.descs is None:
            self.descs = torch.tensor(self.rows, dtype=torch.int64).to(self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().eda_linear_pack_batch(_p(self.descs), len(self.rows), int(self.max_elements),
                                                   _stream(self.device))
        _lib.check(rc, "linear_pack_batch")


    def attach(self, model):
        """Makes this registry the one the modules of `model` consult (while it is `active`)."""
        for m in model.modules():
            m.__dict__["_eda_pack_registry"] = self

    def detach(self, model):
        for m in model.modules():
            if m.__dict__.get("_eda_pack_registry") is self:
                del m.__dict__["_eda_pack_registry"]


def active_registry(owner):
    """The PackRegistry attached to `owner` (a module) if it is currently recording, else None."""
    reg = getattr(owner, "__dict__", {}).get("_eda_pack_registry")
    return reg if (reg is not None and reg.active) else None


def _cache_of(owner):
    """The packed-weight cache lives ON the owning module, so it dies with it (a process-wide dict keyed by
    id() / data_ptr() could hand a new module the packed weights of a dead one whose memory it reuses)."""
    d = owner.__dict__.get("_eda_pack_cache")
    if d is None:
        d = {}
        owner.__dict__["_eda_pack_cache"] = d
    return d


def pack_weight(W, scale=None, cache_key=None):
    """W (N,K) f32 CUDA (any row-contiguous view) -> packed tensor for eda_linear_forward.  With
    `cache_key` = (owner module, name) the result is reused until the weight is modified in place
    (optimizer step, load_state_dict) or reallocated (.to())."""
    lib = _lib.load()
    N, K = W.shape
    cache = None
    reg = active_registry(cache_key[0]) if (cache_key is not None and scale is None) else None
    if reg is not None:
        return reg.get(cache_key[0], cache_key[1], False, W.detach())
    if cache_key is not None and scale is None and PACK_CACHE:
        cache = _cache_of(cache_key[0])
        tag = (W.data_ptr(), W._version, N, K, W.device)
        hit = cache.get(cache_key[1])
        if hit is not None and hit[0] == tag:
            return hit[1]
    Wc = W.detach()
    if not Wc.is_contiguous():
        Wc = Wc.contiguous()
    n = lib.eda_linear_packed_floats(N, K)
    if n == 0:
        raise RuntimeError(f"eda_b200.linear: unsupported weight shape ({N},{K})")
    packed = torch.empty(n, dtype=torch.float32, device=W.device)
    with torch.cuda.device(W.device):
        rc = lib.eda_linear_pack(_p(Wc), _p(scale), N, K, _p(packed), _stream(W.device))
    _lib.check(rc, "linear_pack")
    if cache is not None:
        cache[cache_key[1]] = (tag, packed)
    return packed


def new_seed():
    """A 32-bit dropout seed drawn from torch's default CPU generator (so torch.manual_seed makes runs repeatable)."""
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())


def epoch_of(module):
    """The dropout epoch word (device int32 tensor) a GraphedTrainStep attached to `module`'s model, or None: every
    dropout-applying kernel launched for this module adds its value to the host-drawn seed when it runs."""
    return getattr(module, "__dict__", {}).get("_eda_dropout_epoch")


def _drop(dropout):
    """(p, seed[, epoch tensor]) or None -> (p, seed, epoch pointer) for the C ABI."""
    if dropout is None:
        return 0.0, 0, None
    p, seed = dropout[0], dropout[1]
    epoch = dropout[2] if len(dropout) > 2 else None
    return float(p), int(seed), _p(epoch)


def dropout_mask(seed, p, rows, cols, a_mul, a_add, device, epoch=None):
    """(rows, cols) f32 keep-mask (1 / 0) that a forward kernel seeded `seed` (+ the current value of the `epoch` word)
    applied (see include/eda_b200.h)."""
    out = torch.empty((rows, cols), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rc = _lib.load().eda_dropout_mask(int(seed), _p(epoch), float(p), int(rows), int(cols), int(a_mul), int(a_add),
                                          _p(out), _stream(device))
    _lib.check(rc, "dropout_mask")
    return out


def linear_raw(problems, K, N, relu=False, ln=None, dropout=None):
    """problems: list (<= 3) of dicts x (R,K), w_packed, optional pos, bias, residual.  Returns [y (R,N)].
    dropout = (p, seed): zero each output element with probability p (after bias / ReLU, before the residual)."""
    lib = _lib.load()
    dev = problems[0]["x"].device
    arr = (_lib.LinearProblem * len(problems))()
    outs, keep = [], []
    for i, pr in enumerate(problems):
        x = pr["x"]
        _require_cuda(x)
        assert x.dtype == torch.float32 and x.is_contiguous() and x.size(-1) == K
        R = x.numel() // K
        into = pr.get("y_into")  # optional (R, N) column-block view of a wider row-major matrix: written in place
        if into is not None:
            assert into.shape == (R, N) and into.stride(1) == 1 and into.stride(0) % 4 == 0 and into.dtype == torch.float32
            y = into
            arr[i].y_row_stride = into.stride(0)
        else:
            y = torch.empty((R, N), dtype=torch.float32, device=dev)
        pos, res, bias = pr.get("pos"), pr.get("residual"), pr.get("bias")
        if pos is not None:
            assert pos.is_contiguous() and pos.numel() == x.numel()
        if res is not None:
            assert res.is_contiguous() and res.numel() == R * N
        if bias is not None:
            bias = bias.detach().contiguous()
        keep.append((x, pos, res, bias, pr["w_packed"]))
        arr[i].x, arr[i].pos, arr[i].w_packed = x.data_ptr(), (pos.data_ptr() if pos is not None else None), \
            pr["w_packed"].data_ptr()
        arr[i].bias = bias.data_ptr() if bias is not None else None
        arr[i].residual = res.data_ptr() if res is not None else None
        tb = int(pr.get("y_batch_rows", 0))
        if tb:
            # channel-major output (batch, N, ld): what eda_attention_forward wants for V
            ld = (tb + 3) & ~3
            y = torch.empty((R // tb, N, ld), dtype=torch.float32, device=dev)
            arr[i].y_batch_rows, arr[i].y_ld = tb, ld
        arr[i].y, arr[i].rows = y.data_ptr(), R
        arr[i].round_tf32 = 1 if pr.get("round_tf32") else 0
        pre = pr.get("pre_ln")
        if pre is not None:
            assert pre.is_contiguous() and pre.numel() == R * N and ln is not None
            arr[i].pre_ln = pre.data_ptr()
        outs.append(y)
    g = b = None
    eps = 0.0
    if ln is not None:
        g, b, eps = ln
        g = g.detach().contiguous() if g is not None else None
        b = b.detach().contiguous() if b is not None else None
    with torch.cuda.device(dev):
        dp, dseed, depoch = _drop(dropout)
        rc = lib.eda_linear_forward(ctypes.cast(arr, ctypes.c_void_p), len(problems), K, N, 1 if relu else 0, _p(g),
                                    _p(b), float(eps), 1 if ln is not None else 0, dp, dseed, depoch, _stream(dev))
    _lib.check(rc, "linear_forward")
    return outs


def attention_raw(q, k, vt, key_padding_mask, B, Nq, Nk, H, dropout=None, lse=None):
    """q (B*Nq,E), k (B*Nk,E) projected, vt (B,E,ld) channel-major projected values (ld >= Nk, ld % 4 == 0);
    mask (B,Nk) bool or None -> ctx (B*Nq,E)."""
    lib = _lib.load()
    E = q.size(-1)
    D = E // H
    assert vt.dim() == 3 and vt.size(0) == B and vt.size(1) == E and vt.is_contiguous()
    ldv = vt.size(2)
    ctx = torch.empty((B * Nq, E), dtype=torch.float32, device=q.device)
    m = None
    if key_padding_mask is not None:
        m = key_padding_mask
        if m.dtype != torch.bool:
            m = m != 0
        m = m.contiguous().view(torch.uint8)
        assert m.shape == (B, Nk)
    with torch.cuda.device(q.device):
        dp, dseed, depoch = _drop(dropout)
        rc = lib.eda_attention_forward_lse(_p(q), _p(k), _p(vt), ldv, _p(m), B, Nq, Nk, H, D, 1.0 / math.sqrt(D),
                                           dp, dseed, depoch, _p(ctx), _p(lse), _stream(q.device))
    _lib.check(rc, "attention_forward")
    return ctx



# ---------------------------------------------------------------------------------------------------
# backward building blocks (csrc/grad_ops.cu, csrc/attn_bwd.cu)
# ---------------------------------------------------------------------------------------------------
CUDA_BACKWARD_HEAD_DIMS = (32, 36)  # csrc/attn_bwd.cu: launch_attention_backward<D>


def use_cuda_backward():
    return os.environ.get("EDA_BACKWARD", "cuda") != "torch"


def pack_weight_t(W, cache_key=None):
    """Packs W^T for eda_linear_forward: with W (Nout, Kin) the weight of y = x W^T (any row-strided view), the
    result drives the activation-gradient GEMM dx = dy W (K = Nout, N = Kin)."""
    lib = _lib.load()
    Nout, Kin = W.shape
    assert W.stride(1) == 1
    cache = None
    reg = active_registry(cache_key[0]) if cache_key is not None else None
    if reg is not None:
        return reg.get(cache_key[0], cache_key[1], True, W.detach())
    if cache_key is not None and PACK_CACHE:
        cache = _cache_of(cache_key[0])
        tag = (W.data_ptr(), W._version, Nout, Kin, W.stride(0), W.device)
        hit = cache.get(cache_key[1])
        if hit is not None and hit[0] == tag:
            return hit[1]
    n = lib.eda_linear_packed_floats(Kin, Nout)
    if n == 0:
        raise RuntimeError(f"eda_b200.linear: unsupported transposed weight shape ({Kin},{Nout})")
    packed = torch.empty(n, dtype=torch.float32, device=W.device)
    with torch.cuda.device(W.device):
        rc = lib.eda_linear_pack_strided(_p(W.detach()), 1, int(W.stride(0)), Kin, Nout, _p(packed), _stream(W.device))
    _lib.check(rc, "linear_pack_strided")
    if cache is not None:
        cache[cache_key[1]] = (tag, packed)
    return packed


# Fused gradient accumulation (opt-in per PARAMETER: ddp.FlatGradients tags the parameters it owns): when a tagged
# parameter owns a gradient buffer (param.grad, a view of the flat all-reduce bucket), the weight-gradient kernels
# accumulate straight into it and the autograd Function returns None for that parameter — no per-block zeroed
# scratch, no AccumulateGrad add kernel.  Because nothing downstream in the backward pass reads those buffers, the wgrad
# launches then run on a side stream, off the critical path (most kernels of the attention backward occupy < 64 of the
# 148 SMs).  The side stream is joined automatically at the END of the backward pass that used it (an autograd engine
# callback), so after loss.backward() every .grad is complete on the current stream like with any other op — clipping,
# optimizer.step or hooks need no extra call; join_wgrad() remains for callers outside an autograd backward.
_wgrad_streams = {}
_wgrad_pending = set()  # devices whose side stream has work the current stream has not waited for yet
_join_queued = False


def _wgrad_side(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _wgrad_streams.get(key)
    if st is None:
        st = _wgrad_streams[key] = torch.cuda.Stream(device=dev)
    return st


_aux_pending = {}  # device key -> auxiliary streams (the encoder's language-branch stream) used since the last join


def note_aux_stream(dev, stream):
    """A forward pass put work on `stream` (encoder_decoder_layers.BRANCH_STREAMS): its backward kernels will run there
    too, and whoever joins the gradient streams must wait for it as well."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    _aux_pending.setdefault(key, set()).add(stream)


def _capturing(stream):
    with torch.cuda.stream(stream):
        return torch.cuda.is_current_stream_capturing()


def join_wgrad():
    """The current stream waits for every weight-gradient kernel issued on the side streams so far, and for the
    auxiliary branch streams that carried part of the backward pass."""
    global _join_queued
    _join_queued = False
    # only streams with un-joined work: waiting on an idle side stream would, during CUDA-graph capture, make the
    # capturing stream depend on a stream that is not part of the capture
    for key in list(_wgrad_pending):
        torch.cuda.current_stream(torch.device("cuda", key)).wait_stream(_wgrad_streams[key])
    _wgrad_pending.clear()
    for key, streams in list(_aux_pending.items()):
        cur = torch.cuda.current_stream(torch.device("cuda", key))
        cap = torch.cuda.is_current_stream_capturing()
        for st in streams:
            if cap and not _capturing(st):
                continue  # used by an earlier, un-captured forward: nothing of this capture runs there
            cur.wait_stream(st)
    _aux_pending.clear()


def mark_side_pending(dev):
    """Records that `dev`'s side stream carries un-joined gradient work and arranges the join at the end of the
    running backward pass (no-op outside one: the caller joins with join_wgrad / FlatGradients.sync)."""
    global _join_queued
    _wgrad_pending.add(dev.index if dev.index is not None else torch.cuda.current_device())
    if not _join_queued:
        try:
            torch.autograd.Variable._execution_engine.queue_callback(join_wgrad)
            _join_queued = True
        except RuntimeError:  # not inside a backward pass
            pass


def _bucket_of(p):
    ref = getattr(p, "_eda_fused_grad_owner", None) if p is not None else None
    owner = ref() if ref is not None else None
    return owner if (owner is not None and owner.regions is not None) else None


def grads_expected(params):
    """Forward-side half of the completion tracking below: one more fused backward will write into these parameters'
    gradient buffers (a module evaluated twice in one forward writes twice)."""
    for p in params:
        owner = _bucket_of(p)
        if owner is not None:
            owner.expect(p)


def grads_written(params):
    """Tells the owning FlatGradients bucket that a fused backward has queued its contribution to the gradients of
    `params` (fused accumulation bypasses autograd's AccumulateGrad, so its hooks never fire for them): drives the
    overlapped, bucketed all-reduce.  A parameter counts as complete once every expected contribution is in."""
    for p in params:
        owner = _bucket_of(p)
        if owner is not None:
            owner.written(p)


def fused_grad_enabled(p):
    """True when parameter `p` is owned by a live ddp.FlatGradients with fused accumulation switched on."""
    ref = getattr(p, "_eda_fused_grad_owner", None)
    owner = ref() if ref is not None else None
    return owner is not None and owner.fused


def _grad_buffers(params):
    """Per parameter: its existing gradient buffer when fused accumulation applies, else None."""
    out = []
    for p in params:
        g = getattr(p, "grad", None) if (p is not None and fused_grad_enabled(p)) else None
        out.append(g if (g is not None and g.is_contiguous() and g.dtype == torch.float32 and g.shape == p.shape) else None)
    return out


def wgrad_side(problems, N, K, sums=()):
    """eda_wgrad on the side stream, ordered after everything issued on the current stream so far.  Only for
    problems whose outputs are fused-accumulation buffers (see fused_grad_enabled).  `sums`: (problem index, a, b) triples —
    that problem's x is a + b, formed on the side stream too (one elementwise add instead of a second wgrad problem
    for "x + pos" inputs)."""
    dev = problems[0]["dy"].device
    cur, side = torch.cuda.current_stream(dev), _wgrad_side(dev)
    side.wait_stream(cur)
    mark_side_pending(dev)
    with torch.cuda.stream(side):
        for i, a, b in sums:
            problems[i]["x"] = a + b  # allocated and consumed on the side stream
        wgrad(problems, N, K)
    for i, a, b in sums:
        a.record_stream(side)
        b.record_stream(side)
    summed = {i for i, _, _ in sums}
    for i, pr in enumerate(problems):  # inputs allocated on `cur`: keep them alive until the side-stream kernel has run
        pr["dy"].record_stream(side)
        if i not in summed:
            pr["x"].record_stream(side)


def wgrad(problems, N, K):
    """problems: list (<= 6) of dicts dy (R,N), x (R,K), dw (N,K) view [row stride = dw.stride(0)], db (N) or None,
    optional x_scale / x_shift (K): x is consumed as relu(x * x_scale + x_shift).
    dw += dy^T x, db += column sums of dy (accumulating: the caller zero-initialises)."""
    lib = _lib.load()
    dev = problems[0]["dy"].device
    arr = (_lib.WgradProblem * len(problems))()
    keep = []
    for i, pr in enumerate(problems):
        dy, x, dw, db = pr["dy"], pr["x"], pr["dw"], pr.get("db")
        assert dy.is_contiguous() and x.is_contiguous() and dy.dtype == torch.float32 and x.dtype == torch.float32
        R = dy.numel() // N
        assert x.numel() == R * K and dw.shape == (N, K) and dw.stride(1) == 1
        arr[i].dy, arr[i].x, arr[i].dw = dy.data_ptr(), x.data_ptr(), dw.data_ptr()
        arr[i].db = db.data_ptr() if db is not None else None
        arr[i].rows, arr[i].ldy, arr[i].ldx, arr[i].ldw = R, N, K, dw.stride(0)
        xs, xh = pr.get("x_scale"), pr.get("x_shift")
        arr[i].x_scale = xs.data_ptr() if xs is not None else None
        arr[i].x_shift = xh.data_ptr() if xh is not None else None
        keep.append((dy, x, dw, db, xs, xh))
    with torch.cuda.device(dev):
        rc = lib.eda_wgrad(ctypes.cast(arr, ctypes.c_void_p), len(problems), N, K, _stream(dev))
    _lib.check(rc, "wgrad")


def layernorm_backward(dy, u, gamma, eps, dgamma, dbeta, dropout=None):
    """dy, u (R,N).  Returns (du, dproj): du = gradient of the LayerNorm input, dproj = du with the output-dropout
    mask of the producing GEMM re-applied (the same tensor as du without dropout)."""
    lib = _lib.load()
    N = u.size(-1)
    R = u.numel() // N
    du = torch.empty_like(u)
    dp, dseed, depoch = _drop(dropout)
    dproj = torch.empty_like(u) if dp > 0 else du
    with torch.cuda.device(u.device):
        rc = lib.eda_layernorm_backward(_p(dy), _p(u), _p(gamma.detach().contiguous()), float(eps), R, N, _p(du),
                                        _p(dproj) if dp > 0 else None, _p(dgamma), _p(dbeta), dp, dseed, depoch,
                                        _stream(u.device))
    _lib.check(rc, "layernorm_backward")
    return du, dproj


def rows_gemm(x, W, transpose=False, in_scale=None, in_shift=None, stats=None):
    """y = f(x) W^T (transpose=False, W (N,K)) or f(x) W (transpose=True, W (K,N): the activation gradient of a layer
    y = a W^T), f = relu(x * in_scale + in_shift) when given.  x (R,K) contiguous; W any 2-D strided view.
    stats: optional zeroed (2N,) float64 tensor that receives [column sums, column sums of squares] of y."""
    lib = _lib.load()
    R, K = x.shape
    if transpose:
        assert W.size(0) == K
        N, sn, sk = W.size(1), W.stride(1), W.stride(0)
    else:
        assert W.size(1) == K
        N, sn, sk = W.size(0), W.stride(0), W.stride(1)
    y = torch.empty((R, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.eda_rows_gemm_stats(_p(x), K, _p(in_scale), _p(in_shift), _p(W), int(sn), int(sk), R, K, N, _p(y), N,
                                     _p(stats), _stream(x.device))
    _lib.check(rc, "rows_gemm")
    return y


def relu_backward(dy, y, scale=1.0):
    out = torch.empty_like(dy)
    with torch.cuda.device(dy.device):
        rc = _lib.load().eda_relu_backward(_p(dy), _p(y), float(scale), dy.numel(), _p(out), _stream(dy.device))
    _lib.check(rc, "relu_backward")
    return out


def _transpose_last2(x):
    B, R, C = x.shape
    out = torch.empty((B, C, R), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().eda_transpose_last2(_p(x), B, R, C, _p(out), _stream(x.device))
    _lib.check(rc, "transpose_last2")
    return out


def _channel_major(x, B, N, E):
    """(B*N, E) -> (B, E, ld) with ld = N rounded up to 4 (padding columns stay uninitialised: never read)."""
    ld = (N + 3) & ~3
    if ld == N:
        return _transpose_last2(x.view(B, N, E)), ld
    out = torch.empty((B, E, ld), dtype=torch.float32, device=x.device)
    out[:, :, :N].copy_(x.view(B, N, E).transpose(1, 2))
    return out, ld


def attention_backward_raw(q, k, vt, dctx, c, lse, key_padding_mask, B, Nq, Nk, H, dropout=None, impl=None):
    """q (B*Nq,E), k (B*Nk,E), vt (B,E,ld) as attention_raw took them; dctx, c (B*Nq,E); lse (B,H,Nq).
    Returns dq (B*Nq,E), dk (B*Nk,E), dv (B*Nk,E).  impl: "mma" (warp-level mma.sync kernel, 3 CTAs per SM; the
    default because it is the faster one at head dim 36: 164 us vs 377 us at Nq = Nk = 1024, B = 8) or "tc" (tcgen05
    kernel with TMEM-resident score tiles, one CTA per SM, phases not yet overlapped); EDA_ATTN_BWD overrides."""
    lib = _lib.load()
    impl = impl or os.environ.get("EDA_ATTN_BWD", "mma")
    E = q.size(-1)
    D = E // H
    ld = vt.size(2)
    dq = torch.empty_like(q)
    dk = torch.empty_like(k)
    dv = torch.empty_like(k)
    delta = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device)
    m = None
    if key_padding_mask is not None:
        m = key_padding_mask
        if m.dtype != torch.bool:
            m = m != 0
        m = m.contiguous().view(torch.uint8)
    dp, dseed, depoch = _drop(dropout)
    if impl == "tc":
        v = _transpose_last2(vt)  # (B, ld, E): row-major values, rows >= Nk are padding and never read
        kt, ldk = _channel_major(k, B, Nk, E)
        qt, ldq = _channel_major(q, B, Nq, E)
        dot, _ = _channel_major(dctx, B, Nq, E)
        with torch.cuda.device(q.device):
            rc = lib.eda_attention_backward_tc(_p(q), _p(k), _p(v), ld * E, _p(kt), ldk, _p(qt), _p(dot), ldq, _p(dctx),
                                               _p(c), _p(lse), _p(m), B, Nq, Nk, H, D, 1.0 / math.sqrt(D), dp,
                                               dseed, depoch, _p(delta), _p(dq), _p(dk), _p(dv), _stream(q.device))
        _lib.check(rc, "attention_backward_tc")
        return dq, dk, dv
    v_nat = _transpose_last2(vt) if impl == "mma_natural" else None  # the ABI's row-major-values variant (tests)
    with torch.cuda.device(q.device):  # the warp-level kernel reads the channel-major values as they are
        rc = lib.eda_attention_backward(_p(q), _p(k), _p(v_nat), ld * E if v_nat is not None else 0,
                                        None if v_nat is not None else _p(vt), ld, _p(dctx), _p(c), _p(lse), _p(m), B, Nq, Nk, H, D,
                                        1.0 / math.sqrt(D), dp, dseed, depoch, _p(delta), _p(dq), _p(dk), _p(dv),
                                        _stream(q.device))
    _lib.check(rc, "attention_backward")
    return dq, dk, dv


# ---------------------------------------------------------------------------------------------------
# differentiable restatements (EDA_BACKWARD=torch cross-check only)
# ---------------------------------------------------------------------------------------------------
def _mha_torch(q_in, q_pos, k_in, k_pos, v_in, mask, in_w, in_b, out_w, out_b, H, attn_keep=None, attn_scale=1.0):
    E = q_in.size(-1)
    D = E // H
    B, Nq, _ = q_in.shape
    Nk = k_in.size(1)
    q = F.linear(q_in if q_pos is None else q_in + q_pos, in_w[:E], in_b[:E])
    k = F.linear(k_in if k_pos is None else k_in + k_pos, in_w[E:2 * E], in_b[E:2 * E])
    v = F.linear(v_in, in_w[2 * E:], in_b[2 * E:])
    q = q.view(B, Nq, H, D).transpose(1, 2) * (1.0 / math.sqrt(D))
    k = k.view(B, Nk, H, D).transpose(1, 2)
    v = v.view(B, Nk, H, D).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if mask is not None:
        s = s.masked_fill(mask.view(B, 1, 1, Nk), float("-inf"))
    pr = torch.softmax(s, dim=-1)
    if attn_keep is not None:
        pr = pr * (attn_keep.view(B, H, Nq, Nk) * attn_scale)
    ctx = (pr @ v).transpose(1, 2).reshape(B, Nq, E)
    return F.linear(ctx, out_w, out_b)


class _MHABlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, eps, key, mask, q_in, q_pos, k_in, k_pos, v_in, residual, in_w, in_b, out_w, out_b, ln_w, ln_b,
                drop=(0.0, 0, 0.0, 0, None)):
        E = q_in.size(-1)
        B, Nq, _ = q_in.shape
        Nk = k_in.size(1)
        for t in (q_in, k_in, v_in):
            _require_cuda(t)
        q_in, k_in, v_in = q_in.contiguous(), k_in.contiguous(), v_in.contiguous()
        qp = q_pos.contiguous() if q_pos is not None else None
        kp = k_pos.contiguous() if k_pos is not None else None
        wq = pack_weight(in_w[:E], cache_key=(key, "q"))
        wk = pack_weight(in_w[E:2 * E], cache_key=(key, "k"))
        wv = pack_weight(in_w[2 * E:], cache_key=(key, "v"))
        wo = pack_weight(out_w, cache_key=(key, "o"))
        ib = in_b.detach()
        q, k, v = linear_raw([
            dict(x=q_in, pos=qp, w_packed=wq, bias=ib[:E]),
            dict(x=k_in, pos=kp, w_packed=wk, bias=ib[E:2 * E], round_tf32=True),
            dict(x=v_in, w_packed=wv, bias=ib[2 * E:], y_batch_rows=Nk, round_tf32=True),
        ], E, E)
        p_attn, seed_attn, p_out, seed_out, epoch = drop
        # eda_attention_backward is instantiated for head dims 32 and 36 (the forward also takes 64): other sizes
        # differentiate through the torch restatement instead of failing at backward() time
        train = any(ctx.needs_input_grad) and use_cuda_backward() and (E // H) in CUDA_BACKWARD_HEAD_DIMS
        lse = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device) if train else None
        c = attention_raw(q, k, v, mask, B, Nq, Nk, H, dropout=(p_attn, seed_attn, epoch) if p_attn > 0 else None, lse=lse)
        res = residual.contiguous() if residual is not None else None
        ln = (ln_w, ln_b, eps) if ln_w is not None else None
        u = torch.empty((B * Nq, E), dtype=torch.float32, device=q.device) if (train and ln is not None) else None
        (y,) = linear_raw([dict(x=c, w_packed=wo, bias=out_b, residual=res, pre_ln=u)], E, E, ln=ln,
                          dropout=(p_out, seed_out, epoch) if p_out > 0 else None)
        if ln is None and res is not None:
            y = y + res.view(-1, E)
        ctx.save_for_backward(q_in, q_pos, k_in, k_pos, v_in, residual, in_w, in_b, out_w, out_b, ln_w, ln_b,
                              *((q, k, v, c, lse, u) if train else ()))
        ctx.meta = (H, eps, mask, drop)
        ctx.key = key
        ctx.cuda_bw = train
        ctx.gbufs = _grad_buffers((in_w, in_b, out_w, out_b, ln_w, ln_b)) if train else None
        if train:
            grads_expected((in_w, in_b, out_w, out_b, ln_w, ln_b))
        return y.view(B, Nq, E)

    @staticmethod
    def backward(ctx, grad):
        if ctx.cuda_bw:
            return _MHABlockFn._backward_cuda(ctx, grad)
        H, eps, mask, drop = ctx.meta
        p_attn, seed_attn, p_out, seed_out, epoch = drop
        saved = ctx.saved_tensors
        with torch.enable_grad():
            ts = [None if t is None else t.detach().requires_grad_(ctx.needs_input_grad[4 + i])
                  for i, t in enumerate(saved)]
            q_in, q_pos, k_in, k_pos, v_in, residual, in_w, in_b, out_w, out_b, ln_w, ln_b = ts
            B_, Nq_, E_ = q_in.shape
            attn_keep = None
            if p_attn > 0:  # the exact keep-mask the forward kernel applied
                attn_keep = dropout_mask(seed_attn, p_attn, B_ * H * Nq_, k_in.size(1), 1, 0, q_in.device, epoch)
            y = _mha_torch(q_in, q_pos, k_in, k_pos, v_in, mask, in_w, in_b, out_w, out_b, H, attn_keep,
                           1.0 / (1.0 - p_attn))
            if p_out > 0:
                y = y * (dropout_mask(seed_out, p_out, B_ * Nq_, E_, 3, 0, q_in.device, epoch).view_as(y) * (1.0 / (1.0 - p_out)))
            if residual is not None:
                y = residual + y
            if ln_w is not None:
                y = F.layer_norm(y, (y.size(-1),), ln_w, ln_b, eps)
            wanted = [t for t in ts if t is not None and t.requires_grad]
            grads = torch.autograd.grad(y, wanted, grad, allow_unused=True) if wanted else []
        gmap = {id(t): g for t, g in zip(wanted, grads)}
        return (None, None, None, None, *[gmap.get(id(t)) if t is not None else None for t in ts], None)



    @staticmethod
    def _backward_cuda(ctx, grad):
        H, eps, mask, drop = ctx.meta
        p_attn, seed_attn, p_out, seed_out, epoch = drop
        (q_in, q_pos, k_in, k_pos, v_in, residual, in_w, in_b, out_w, out_b, ln_w, ln_b, q, k, vt, c, lse, u) = \
            ctx.saved_tensors
        key = ctx.key
        B, Nq, E = q_in.shape
        Nk = k_in.size(1)
        dev = q_in.device
        grad = grad.contiguous().view(-1, E)
        has_ln = ln_w is not None
        # gradient buffers: the parameters' own (fused accumulation) or one zeroed scratch for the whole block
        g_in_w, g_in_b, g_out_w, g_out_b, g_ln_w, g_ln_b = ctx.gbufs
        fused = all(g is not None for g in (g_in_w, g_in_b, g_out_w, g_out_b)) and \
            (not has_ln or (g_ln_w is not None and g_ln_b is not None))
        if fused:
            d_in_w, d_in_b, d_out_w, d_out_b, d_ln_w, d_ln_b = g_in_w, g_in_b, g_out_w, g_out_b, g_ln_w, g_ln_b
        else:
            flat = torch.zeros(3 * E * E + 3 * E + E * E + E + 2 * E, dtype=torch.float32, device=dev)
            o = 0
            d_in_w = flat[o:o + 3 * E * E].view(3 * E, E); o += 3 * E * E
            d_in_b = flat[o:o + 3 * E]; o += 3 * E
            d_out_w = flat[o:o + E * E].view(E, E); o += E * E
            d_out_b = flat[o:o + E]; o += E
            d_ln_w = flat[o:o + E]; o += E
            d_ln_b = flat[o:o + E]
        run_wgrad = wgrad_side if fused else wgrad
        # 1. LayerNorm (+ output dropout)
        if has_ln:
            du, dproj = layernorm_backward(grad, u, ln_w, eps, d_ln_w, d_ln_b,
                                           dropout=(p_out, seed_out, epoch) if p_out > 0 else None)
        else:
            du = grad
            dproj = grad
            if p_out > 0:
                dproj = grad * (dropout_mask(seed_out, p_out, B * Nq, E, 3, 0, dev, epoch) * (1.0 / (1.0 - p_out)))
        # 2. out-projection: activation gradient (tcgen05 GEMM with the transposed weight) and weight gradient
        (dctx,) = linear_raw([dict(x=dproj, w_packed=pack_weight_t(out_w, cache_key=(key, "ot")))], E, E)
        run_wgrad([dict(dy=dproj, x=c, dw=d_out_w, db=d_out_b)], E, E)
        # 3. attention core
        dq, dk, dv = attention_backward_raw(q, k, vt, dctx, c, lse, mask, B, Nq, Nk, H,
                                            dropout=(p_attn, seed_attn, epoch) if p_attn > 0 else None)
        # 4. in-projections
        dq_in, dk_in, dv_in = linear_raw([
            dict(x=dq, w_packed=pack_weight_t(in_w[:E], cache_key=(key, "qt"))),
            dict(x=dk, w_packed=pack_weight_t(in_w[E:2 * E], cache_key=(key, "kt"))),
            dict(x=dv, w_packed=pack_weight_t(in_w[2 * E:], cache_key=(key, "vt"))),
        ], E, E)
        probs = [dict(dy=dq, x=q_in, dw=d_in_w[:E], db=d_in_b[:E]),
                 dict(dy=dk, x=k_in, dw=d_in_w[E:2 * E], db=d_in_b[E:2 * E]),
                 dict(dy=dv, x=v_in, dw=d_in_w[2 * E:], db=d_in_b[2 * E:])]
        if fused:
            # dW_q = dq^T (q_in + q_pos): the sum is formed on the side stream (one add) instead of a second product
            sums = [(0, q_in, q_pos.contiguous())] if q_pos is not None else []
            if k_pos is not None:
                sums.append((1, k_in, k_pos.contiguous()))
            wgrad_side(probs, E, E, sums=sums)
        else:
            if q_pos is not None:
                probs.append(dict(dy=dq, x=q_pos.contiguous(), dw=d_in_w[:E]))
            if k_pos is not None:
                probs.append(dict(dy=dk, x=k_pos.contiguous(), dw=d_in_w[E:2 * E]))
            wgrad(probs, E, E)
        if fused:
            grads_written((in_w, in_b, out_w, out_b, ln_w, ln_b))
        dq_in = dq_in.view(B, Nq, E)
        dk_in = dk_in.view(B, Nk, E)
        dv_in = dv_in.view(B, Nk, E)
        return (None, None, None, None, dq_in, dq_in if q_pos is not None else None, dk_in,
                dk_in if k_pos is not None else None, dv_in, du.view(B, Nq, E) if residual is not None else None,
                *((None,) * 6 if fused else (d_in_w, d_in_b, d_out_w, d_out_b, d_ln_w if has_ln else None,
                                             d_ln_b if has_ln else None)), None)


def mha_block(mha, q_in, k_in, v_in, q_pos=None, k_pos=None, key_padding_mask=None, residual=None, norm=None,
              out_dropout=None):
    """LayerNorm(residual + MHA(q_in + q_pos, k_in + k_pos, v_in)) with `mha` an nn.MultiheadAttention
    (parameters only) and `norm` an nn.LayerNorm (or None: no residual LayerNorm, plain attention output
    [+ residual]).  All activations batch-first (B, S, E)."""
    p_attn = float(mha.dropout) if mha.training else 0.0
    p_out = float(out_dropout.p) if (out_dropout is not None and out_dropout.training) else 0.0
    drop = (p_attn, new_seed() if p_attn > 0 else 0, p_out, new_seed() if p_out > 0 else 0, epoch_of(mha))
    return _MHABlockFn.apply(mha.num_heads, norm.eps if norm is not None else 0.0, mha, key_padding_mask, q_in,
                             q_pos, k_in, k_pos, v_in, residual, mha.in_proj_weight, mha.in_proj_bias,
                             mha.out_proj.weight, mha.out_proj.bias, norm.weight if norm is not None else None,
                             norm.bias if norm is not None else None, drop)


class _FFNBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eps, key, x, w1, b1, w2, b2, ln_w, ln_b, drop=(0.0, 0, 0.0, 0, None)):
        _require_cuda(x)
        shape = x.shape
        E, Fh = w1.size(1), w1.size(0)
        x2 = x.contiguous().view(-1, E)
        p1 = pack_weight(w1, cache_key=(key, "w1"))
        p2 = pack_weight(w2, cache_key=(key, "w2"))
        pa, sa, pb, sb, epoch = drop
        train = any(ctx.needs_input_grad) and use_cuda_backward()
        (hdn,) = linear_raw([dict(x=x2, w_packed=p1, bias=b1)], E, Fh, relu=True, dropout=(pa, sa, epoch) if pa > 0 else None)
        u = torch.empty_like(x2) if train else None
        (y,) = linear_raw([dict(x=hdn, w_packed=p2, bias=b2, residual=x2, pre_ln=u)], Fh, E, ln=(ln_w, ln_b, eps),
                          dropout=(pb, sb, epoch) if pb > 0 else None)
        ctx.save_for_backward(x, w1, b1, w2, b2, ln_w, ln_b, *((hdn, u) if train else ()))
        ctx.eps = eps
        ctx.drop = drop
        ctx.key = key
        ctx.cuda_bw = train
        ctx.gbufs = _grad_buffers((w1, b1, w2, b2, ln_w, ln_b)) if train else None
        if train:
            grads_expected((w1, b1, w2, b2, ln_w, ln_b))
        return y.view(shape)

    @staticmethod
    def _backward_cuda(ctx, grad):
        x, w1, b1, w2, b2, ln_w, ln_b, hdn, u = ctx.saved_tensors
        pa, sa, pb, sb, epoch = ctx.drop
        key = ctx.key
        E, Fh = w1.size(1), w1.size(0)
        dev = x.device
        grad = grad.contiguous().view(-1, E)
        fused = all(g is not None for g in ctx.gbufs)
        if fused:
            dw1, db1, dw2, db2, d_ln_w, d_ln_b = ctx.gbufs
        else:
            flat = torch.zeros(2 * E * Fh + Fh + E + 2 * E, dtype=torch.float32, device=dev)
            o = 0
            dw1 = flat[o:o + Fh * E].view(Fh, E); o += Fh * E
            dw2 = flat[o:o + E * Fh].view(E, Fh); o += E * Fh
            db1 = flat[o:o + Fh]; o += Fh
            db2 = flat[o:o + E]; o += E
            d_ln_w = flat[o:o + E]; o += E
            d_ln_b = flat[o:o + E]
        run_wgrad = wgrad_side if fused else wgrad
        du, dproj = layernorm_backward(grad, u, ln_w, ctx.eps, d_ln_w, d_ln_b, dropout=(pb, sb, epoch) if pb > 0 else None)
        (dh,) = linear_raw([dict(x=dproj, w_packed=pack_weight_t(w2, cache_key=(key, "w2t")))], E, Fh)
        # hdn is the saved post-ReLU, post-dropout activation: > 0 exactly where the unit was active and kept
        dz = relu_backward(dh, hdn, 1.0 / (1.0 - pa) if pa > 0 else 1.0)
        run_wgrad([dict(dy=dproj, x=hdn, dw=dw2, db=db2)], E, Fh)
        run_wgrad([dict(dy=dz, x=x.contiguous().view(-1, E), dw=dw1, db=db1)], Fh, E)
        (dx,) = linear_raw([dict(x=dz, w_packed=pack_weight_t(w1, cache_key=(key, "w1t")))], Fh, E)
        dx += du
        if fused:
            grads_written((w1, b1, w2, b2, ln_w, ln_b))
            return (None, None, dx.view(x.shape), None, None, None, None, None, None, None)
        return (None, None, dx.view(x.shape), dw1, db1, dw2, db2, d_ln_w, d_ln_b, None)

    @staticmethod
    def backward(ctx, grad):
        if ctx.cuda_bw:
            return _FFNBlockFn._backward_cuda(ctx, grad)
        with torch.enable_grad():
            ts = [t.detach().requires_grad_(ctx.needs_input_grad[2 + i]) for i, t in enumerate(ctx.saved_tensors)]
            x, w1, b1, w2, b2, ln_w, ln_b = ts
            pa, sa, pb, sb, epoch = ctx.drop
            hdn = F.relu(F.linear(x, w1, b1))
            R = hdn.numel() // hdn.size(-1)
            if pa > 0:
                hdn = hdn * (dropout_mask(sa, pa, R, hdn.size(-1), 3, 0, x.device, epoch).view_as(hdn) * (1.0 / (1.0 - pa)))
            o = F.linear(hdn, w2, b2)
            if pb > 0:
                o = o * (dropout_mask(sb, pb, R, o.size(-1), 3, 0, x.device, epoch).view_as(o) * (1.0 / (1.0 - pb)))
            y = F.layer_norm(x + o, (x.size(-1),), ln_w, ln_b, ctx.eps)
            wanted = [t for t in ts if t.requires_grad]
            grads = torch.autograd.grad(y, wanted, grad, allow_unused=True) if wanted else []
        gmap = {id(t): g for t, g in zip(wanted, grads)}
        return (None, None, *[gmap.get(id(t)) for t in ts], None)


def ffn_block(ffn, x, norm):
    """norm(x + ffn(x)) for ffn = Sequential(Linear, ReLU, Dropout, Linear, Dropout) (indices 0 and 3)."""
    l1, l2 = ffn[0], ffn[3]
    pa = float(ffn[2].p) if ffn[2].training else 0.0
    pb = float(ffn[4].p) if ffn[4].training else 0.0
    drop = (pa, new_seed() if pa > 0 else 0, pb, new_seed() if pb > 0 else 0, epoch_of(ffn))
    return _FFNBlockFn.apply(norm.eps, ffn, x, l1.weight, l1.bias, l2.weight, l2.bias, norm.weight, norm.bias, drop)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, relu, key, x, w, b, scale):
        _require_cuda(x)
        N, K = w.shape
        x2 = x.contiguous().view(-1, K)
        packed = pack_weight(w, scale=scale, cache_key=key)
        (y,) = linear_raw([dict(x=x2, w_packed=packed, bias=b)], K, N, relu=relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.relu = relu
        ctx.has_scale = scale is not None
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, grad):
        if ctx.has_scale:
            raise RuntimeError("eda_b200.linear: the scale-folded variant is inference-only")
        x, w, y = ctx.saved_tensors
        N, K = w.shape
        g = grad.contiguous().view(-1, N)
        if ctx.relu:
            g = relu_backward(g, y.view(-1, N), 1.0)
        gx = gw = gb = None
        if ctx.needs_input_grad[2]:  # dX = dY W on the forward tcgen05 GEMM with the transposed packed weight
            (gx,) = linear_raw([dict(x=g, w_packed=pack_weight_t(w))], N, K)
            gx = gx.view(x.shape)
        if ctx.needs_input_grad[3] or ctx.needs_input_grad[4]:
            gw = torch.zeros((N, K), dtype=torch.float32, device=g.device)
            gb = torch.zeros(N, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[4] else None
            wgrad([dict(dy=g, x=x.contiguous().view(-1, K), dw=gw, db=gb)], N, K)
        return None, None, gx, gw, gb, None


def linear(x, weight, bias=None, relu=False, scale=None, cache_key=None):
    """act(x W^T * scale + bias) on the tcgen05 linear kernel; x (..., K), weight (N, K)."""
    return _LinearFn.apply(relu, cache_key, x, weight, bias, scale)
