"""Row-major MLP blocks on this package's own kernels, forward AND backward: a chain of
    z = x W^T [+ b]  ->  [BatchNorm (batch or running statistics, optionally synchronised over ranks)]  ->  [ReLU]
over (rows, channels) matrices.  Two users, both of which ran on torch Conv / BatchNorm / matmul kernels (cuDNN, cuBLAS)
in round 1:

  * PointnetFPModule's SharedMLP (pointnet2/pointnet2_modules.py:393-416): [interpolated | skip] rows -> 2 x
    (1x1 conv, BatchNorm2d, ReLU), together with `fp_rows` (3-NN weights + interpolation + concat in one pass,
    eda_fp_gather_rows) and its backward (eda_fp_scatter_rows);
  * PositionEmbeddingLearned (models/encoder_decoder_layers.py:19-34): Conv1d -> BatchNorm1d -> ReLU -> Conv1d.

GEMMs: eda_linear_forward (tcgen05 kind::tf32) for the products and the activation gradients (transposed packed
weight), eda_wgrad / eda_wgrad_small for the weight gradients; BatchNorm: eda_col_stats -> eda_bn_finalize ->
eda_bn_relu_apply forward, eda_bn_relu_backward_stats / _apply backward (csrc/sa_bwd.cu) — the same kernels the
set-abstraction stage uses.  No torch arithmetic on the path; CPU tensors raise like everywhere else in the package.
"""
import ctypes

import torch

from . import _lib, attn_ops as ops, syncbn


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _chk(rc, what):
    _lib.check(rc, what)


class Layer:
    """One layer of the chain: `weight` (N, K[, 1[, 1]]) and optional `bias` parameters, optional BatchNorm module
    (`bn`), optional ReLU.  `owner` / `name` key the packed-weight cache."""

    def __init__(self, weight, bias, bn, relu, owner, name, dropout=None):
        self.weight, self.bias, self.bn, self.relu, self.owner, self.name = weight, bias, bn, relu, owner, name
        self.dropout = dropout  # nn.Dropout applied to this layer's output (after the ReLU), or None
        if bn is not None and not relu:
            raise RuntimeError("eda_b200.rows_mlp: BatchNorm without ReLU is not on the path")

    def drop_p(self):
        d = self.dropout
        return float(d.p) if (d is not None and d.training and d.p > 0) else 0.0


def _w2d(w):
    return w.reshape(w.size(0), -1)


def bn_scale_shift(dev, stats, count, bn, C, training_update, want_stats=False):
    """Folded BatchNorm terms of one layer: scale = gamma / sqrt(var + eps), shift = beta - mean * scale, from the
    batch sums `stats` = [sum z, sum z^2] (fp64, 2C; summed over the ranks first when the layer is synchronised) or,
    with stats None, from the running statistics.  training_update: also update running_mean / running_var /
    num_batches_tracked like nn.BatchNorm does in training mode."""
    lib = _lib.load()
    scale = torch.empty(C, dtype=torch.float32, device=dev)
    shift = torch.empty(C, dtype=torch.float32, device=dev)
    mean_invstd = torch.empty(2, C, dtype=torch.float32, device=dev) if want_stats else None
    momentum = bn.momentum
    if training_update:
        bn.num_batches_tracked += 1
        if momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked.item())
    red = syncbn.reducer_of(bn) if stats is not None else None
    mom = float(momentum if momentum is not None else 0.0)
    if red is not None:
        count = red.total_count(count)
        if red.peer is not None and stats.is_cuda and 2 * C <= red.peer.max_elems:
            # synchronised layer, NVLink peer path: exchange + ordered sum + finalisation in ONE launch
            pr = red.peer
            rc = lib.eda_bn_finalize_peer(_p(pr.ptrs), pr.world, pr.rank, pr.max_elems, _p(stats), float(count),
                                          _p(bn.weight.detach()), _p(bn.bias.detach()), float(bn.eps), mom,
                                          _p(bn.running_mean), _p(bn.running_var), 1 if training_update else 0, C,
                                          _p(scale), _p(shift), _p(mean_invstd[0]) if want_stats else None,
                                          _p(mean_invstd[1]) if want_stats else None, ops._stream(dev))
            _chk(rc, "bn_finalize_peer")
            return scale, shift, mean_invstd
        red.all_reduce_sum_(stats)
    rc = lib.eda_bn_finalize(_p(stats), float(count), _p(bn.weight.detach()), _p(bn.bias.detach()), float(bn.eps),
                             mom, _p(bn.running_mean),
                             _p(bn.running_var), 1 if training_update else 0, C, _p(scale), _p(shift),
                             _p(mean_invstd[0]) if want_stats else None, _p(mean_invstd[1]) if want_stats else None,
                             ops._stream(dev))
    _chk(rc, "bn_finalize")
    return scale, shift, mean_invstd


def bn_backward_reduce(bn, stats, count):
    """Backward counterpart: `stats` = [sum dy, sum dy*zhat] (fp32, 2C) of the local rows.  Returns (stats to apply,
    total count, local stats = the BatchNorm affine gradients).  Synchronised layers sum the statistics over ranks;
    the affine gradients stay local sums, as in torch's SyncBatchNorm."""
    red = syncbn.reducer_of(bn)
    if red is None:
        return stats, count, stats
    local = stats.clone()
    red.all_reduce_sum_(stats)
    return stats, red.total_count(count), local


def _linear(x, W, bias, relu, owner, name, scale=None, dropout=None):
    N, K = W.shape
    packed = ops.pack_weight(W, scale=scale, cache_key=(owner, name) if owner is not None else None)
    (y,) = ops.linear_raw([dict(x=x, w_packed=packed, bias=bias)], K, N, relu=relu, dropout=dropout)
    return y


def _padded(L, W, bias):
    """Output widths the tcgen05 GEMM does not take (not a multiple of 16: the 3-channel centre / size heads, the
    1-channel objectness logit) run zero-padded to the next multiple of 16; the extra columns are dropped again."""
    N, K = W.shape
    N16 = (N + 15) // 16 * 16
    Wp = torch.zeros((N16, K), dtype=torch.float32, device=W.device)
    Wp[:N] = W
    bp = None
    if bias is not None:
        bp = torch.zeros(N16, dtype=torch.float32, device=W.device)
        bp[:N] = bias
    return Wp, bp


def _dropout_apply(y, p, seed, epoch):
    out = torch.empty_like(y)
    with torch.cuda.device(y.device):
        rc = _lib.load().eda_dropout_apply(_p(y), int(seed), _p(epoch), float(p), y.size(0), y.size(1), 1, 0, _p(out),
                                           ops._stream(y.device))
    _chk(rc, "dropout_apply")
    return out


def _dgrad(dz, W, owner, name):
    """da (R, K) = dz W for W (N, K): eda_linear_forward on the transposed packed weight, in column blocks the kernel
    takes (<= 320 output columns, multiples of 16), each written straight into its columns of the result."""
    N, K = W.shape
    if K % 16 != 0:
        raise RuntimeError(f"eda_b200.rows_mlp: input gradient of a layer with {K} input channels is not supported")
    if K <= 320:
        Wt = ops.pack_weight_t(W, cache_key=(owner, f"{name}.t") if owner is not None else None)
        (o,) = ops.linear_raw([dict(x=dz, w_packed=Wt)], N, K)
        return o
    out = torch.empty((dz.size(0), K), dtype=torch.float32, device=dz.device)
    c0 = 0
    while c0 < K:
        c1 = min(c0 + 256, K)
        Wt = ops.pack_weight_t(W[:, c0:c1], cache_key=(owner, f"{name}.t{c0}") if owner is not None else None)
        ops.linear_raw([dict(x=dz, w_packed=Wt, y_into=out[:, c0:c1])], N, c1 - c0)
        c0 = c1
    return out


def _wgrad(dz, a_in, W, dW, db):
    """dW (N, K) += dz^T a_in, db += column sums of dz."""
    lib = _lib.load()
    N, K = W.shape
    R = dz.size(0)
    if K % 4 == 0 and N % 4 == 0:
        ops.wgrad([dict(dy=dz, x=a_in, dw=dW, db=db)], N, K)
    elif K <= 8:
        with torch.cuda.device(dz.device):
            _chk(lib.eda_wgrad_small(_p(dz), N, _p(a_in), K, R, N, K, _p(dW), dW.stride(0), _p(db),
                                     ops._stream(dz.device)), "wgrad_small")
    else:
        raise RuntimeError(f"eda_b200.rows_mlp: weight gradient for shape ({N},{K}) is not supported")


class _RowsMLPFn(torch.autograd.Function):
    """y (R, N_last) = chain(x (R, K_0)).  params: per layer weight, bias, bn.weight, bn.bias (None where absent)."""

    @staticmethod
    def forward(ctx, layers, x, *params):
        ops._require_cuda(x)
        lib = _lib.load()
        dev = x.device
        x = x.contiguous()
        R = x.size(0)
        need_grad = any(ctx.needs_input_grad)
        saved = []
        a = x
        with torch.cuda.device(dev):
            for li, L in enumerate(layers):
                W = _w2d(L.weight.detach())
                N = W.size(0)
                bias = L.bias.detach() if L.bias is not None else None
                bn = L.bn
                pdrop = L.drop_p()
                drop = (pdrop, ops.new_seed(), ops.epoch_of(L.owner)) if pdrop > 0 else None
                if bn is None:
                    if N % 16:
                        Wp, bp = _padded(L, W, bias)
                        y = _linear(a, Wp, bp, L.relu, None, None, dropout=drop)[:, :N].contiguous()
                    else:
                        y = _linear(a, W, bias, L.relu, L.owner, L.name, dropout=drop)
                    saved.append((a, None, y if L.relu else None, None, drop))
                    a = y
                    continue
                if not bn.training and not need_grad:
                    # inference: running statistics folded into the GEMM (scale into the packed weight, shift as bias)
                    packed, shift = _folded(L, W, bias, bn, dev)
                    (a,) = ops.linear_raw([dict(x=a, w_packed=packed, bias=shift)], W.size(1), N, relu=True, dropout=drop)
                    saved.append(None)
                    continue
                z = _linear(a, W, bias, False, L.owner, L.name)
                if bn.training:
                    stats = torch.zeros(2 * N, dtype=torch.float64, device=dev)
                    _chk(lib.eda_col_stats(_p(z), R, N, _p(stats), ops._stream(dev)), "col_stats")
                    scale, shift, mi = bn_scale_shift(dev, stats, float(R), bn, N, True, want_stats=True)
                else:
                    scale, shift, mi = bn_scale_shift(dev, None, 0.0, bn, N, False, want_stats=True)
                y = torch.empty_like(z)
                _chk(lib.eda_bn_relu_apply(_p(z), _p(scale), _p(shift), R, N, _p(y), ops._stream(dev)), "bn_relu_apply")
                if drop is not None:
                    y = _dropout_apply(y, *drop)  # post-dropout activation: > 0 exactly where the unit was active and kept
                saved.append((a, z, y if drop is not None else None, (scale, shift, mi, bool(bn.training)), drop))
                a = y
        ctx.layers = layers
        ctx.saved = saved if need_grad else None
        ctx.gbufs = ops._grad_buffers(params) if need_grad else None
        if need_grad:
            for li, L in enumerate(layers):  # mirrors the `fused` decision of the backward pass
                W = _w2d(L.weight)
                gw, gb = ctx.gbufs[4 * li], ctx.gbufs[4 * li + 1]
                if gw is not None and (L.bias is None or gb is not None) and W.size(1) % 4 == 0 and W.size(0) % 4 == 0:
                    ops.grads_expected((L.weight, L.bias))
        return a

    @staticmethod
    def backward(ctx, grad):
        lib = _lib.load()
        layers, saved = ctx.layers, ctx.saved
        ctx.saved = None
        dev = grad.device
        g = grad.contiguous()
        own = False  # whether `g` is a tensor this function may overwrite
        grads = [None] * (4 * len(layers))
        dx = None
        with torch.cuda.device(dev):
            for li in range(len(layers) - 1, -1, -1):
                L = layers[li]
                a_in, z, y, bnstate, drop = saved[li]
                W = _w2d(L.weight.detach())
                N, K = W.shape
                R = g.size(0)
                pad = bnstate is None and N % 16 != 0
                if pad:  # zero-padded output columns (see _padded): their gradient is zero
                    Wp, _ = _padded(L, W, None)
                    g16 = torch.zeros((R, Wp.size(0)), dtype=torch.float32, device=dev)
                    g16[:, :N] = g
                    g, own = g16, True
                if bnstate is not None:
                    scale, shift, mi, batch = bnstate
                    if drop is not None:  # gradient through dropout o ReLU from the saved post-dropout activation
                        g = ops.relu_backward(g, y, 1.0 / (1.0 - drop[0]))
                        own = True
                    if not own:
                        g = g.clone()
                    stats = torch.zeros(2 * N, dtype=torch.float32, device=dev)
                    _chk(lib.eda_bn_relu_backward_stats(_p(g), _p(z), _p(scale), _p(shift), _p(mi[0]), _p(mi[1]), R, N,
                                                        _p(stats), ops._stream(dev)), "bn_relu_backward_stats")
                    count, local = float(R), stats
                    if batch:
                        stats, count, local = bn_backward_reduce(L.bn, stats, float(R))
                    _chk(lib.eda_bn_relu_backward_apply(_p(g), _p(z), _p(scale), _p(shift), _p(mi[0]), _p(mi[1]),
                                                        _p(stats), count, 1 if batch else 0, R, N, ops._stream(dev)),
                         "bn_relu_backward_apply")
                    dz = g
                    grads[4 * li + 2], grads[4 * li + 3] = local[N:], local[:N]   # d gamma = sum dy zhat, d beta = sum dy
                elif L.relu:
                    dz = ops.relu_backward(g, y, 1.0 / (1.0 - drop[0]) if drop is not None else 1.0)
                else:
                    dz = g
                # weight / bias gradients: straight into the parameters' own buffers when a FlatGradients bucket owns
                # them (side stream, off the critical path), else into zeroed scratch returned to autograd
                gw, gb = ctx.gbufs[4 * li], ctx.gbufs[4 * li + 1]
                has_b = L.bias is not None
                wants_w, wants_b = ctx.needs_input_grad[2 + 4 * li], has_b and ctx.needs_input_grad[3 + 4 * li]
                if pad and (wants_w or wants_b):
                    dW16 = torch.zeros(tuple(Wp.shape), dtype=torch.float32, device=dev)
                    db16 = torch.zeros(Wp.size(0), dtype=torch.float32, device=dev) if has_b else None
                    _wgrad(dz, a_in, Wp, dW16, db16)
                    grads[4 * li] = dW16[:N].reshape(L.weight.shape)
                    grads[4 * li + 1] = db16[:N] if has_b else None
                elif wants_w or wants_b:
                    fused = gw is not None and (not has_b or gb is not None) and K % 4 == 0 and N % 4 == 0
                    if fused:
                        ops.wgrad_side([dict(dy=dz, x=a_in, dw=_w2d(gw), db=gb if has_b else None)], N, K)
                        ops.grads_written((L.weight, L.bias))
                    else:
                        dW = torch.zeros((N, K), dtype=torch.float32, device=dev)
                        db = torch.zeros(N, dtype=torch.float32, device=dev) if has_b else None
                        _wgrad(dz, a_in, W, dW, db)
                        grads[4 * li] = dW.view(L.weight.shape)
                        grads[4 * li + 1] = db
                if li > 0 or ctx.needs_input_grad[1]:
                    g = _dgrad(dz, Wp, None, None) if pad else _dgrad(dz, W, L.owner, L.name)
                    own = True
                    if li == 0:
                        dx = g
        return (None, dx, *grads)


def _folded(L, W, bias, bn, dev):
    """Eval-mode BatchNorm folded into the packed weight (scale) and a bias vector (shift), cached on the owning
    module until a parameter or buffer changes."""
    tensors = [L.weight, L.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
    key = tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors) + (str(dev),)
    cache = L.owner.__dict__.setdefault("_eda_folded_cache", {}) if L.owner is not None else None
    use_cache = cache is not None and ops.PACK_CACHE and ops.active_registry(L.owner) is None
    hit = cache.get(L.name) if use_cache else None
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    N = W.size(0)
    scale, shift, _ = bn_scale_shift(dev, None, 0.0, bn, N, False)
    if bias is not None:
        shift = shift + bias * scale  # (x W^T + b) * scale + shift; once per weight update (cached), not per call
    packed = ops.pack_weight(W, scale=scale)
    if use_cache:
        cache[L.name] = (key, packed, shift)
    return packed, shift


def rows_mlp(x, layers):
    """x (R, K0) f32 CUDA -> (R, N_last) through `layers` (list of Layer)."""
    params = []
    for L in layers:
        params += [L.weight, L.bias, L.bn.weight if L.bn is not None else None, L.bn.bias if L.bn is not None else None]
    return _RowsMLPFn.apply(layers, x, *params)


class _TransposeFn(torch.autograd.Function):
    """(B, R, C) contiguous -> (B, C, R) contiguous on eda_transpose_last2, both directions."""

    @staticmethod
    def forward(ctx, x):
        from .pointnet2 import fused

        return fused.transpose_last2(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        from .pointnet2 import fused

        return fused.transpose_last2(g.contiguous())


def transpose_last2(x):
    return _TransposeFn.apply(x)


class _FPRowsFn(torch.autograd.Function):
    """x0 (B*n, C2 + C1) = [three_interpolate(known_feats, idx, weight) | unknow_feats] in row-major layout, with the
    inverse-distance weights formed in the same pass (pointnet2_modules.py:393-410).  known_feats (B,C2,m),
    unknow_feats (B,C1,n) or None are the module's channel-major tensors; their point-major copies are reused when a
    previous fused stage left them behind."""

    @staticmethod
    def forward(ctx, dist2, idx, known_feats, unknow_feats):
        from .pointnet2 import fused

        lib = _lib.load()
        dev = known_feats.device
        B, C2, m = known_feats.shape
        n = idx.size(1)
        C1 = 0 if unknow_feats is None else unknow_feats.size(1)
        known_pm = fused.point_major(known_feats.detach())
        skip_pm = fused.point_major(unknow_feats.detach()) if unknow_feats is not None else None
        x0 = torch.empty((B * n, C2 + C1), dtype=torch.float32, device=dev)
        weight = torch.empty((B, n, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.eda_fp_gather_rows(_p(known_pm), known_pm.stride(1), _p(skip_pm),
                                        skip_pm.stride(1) if skip_pm is not None else 0, _p(idx), _p(dist2), B, n, m, C2,
                                        C1, _p(x0), _p(weight), ops._stream(dev))
        _chk(rc, "fp_gather_rows")
        ctx.save_for_backward(idx, weight)
        ctx.dims = (B, n, m, C2, C1)
        return x0

    @staticmethod
    def backward(ctx, g):
        from .pointnet2 import fused

        lib = _lib.load()
        idx, weight = ctx.saved_tensors
        B, n, m, C2, C1 = ctx.dims
        dev = g.device
        g = g.contiguous()
        dk = ds = None
        with torch.cuda.device(dev):
            if ctx.needs_input_grad[2]:
                dk_pm = torch.empty((B, m, C2), dtype=torch.float32, device=dev)
                rc = lib.eda_fp_scatter_rows(_p(g), C2 + C1, _p(idx), _p(weight), B, n, m, C2, _p(dk_pm), ops._stream(dev))
                _chk(rc, "fp_scatter_rows")
                dk = fused.transpose_last2(dk_pm)
            if C1 and ctx.needs_input_grad[3]:
                ds = torch.empty((B, C1, n), dtype=torch.float32, device=dev)
                # (B, n, C1) slice of the rows -> (B, C1, n): strided transpose on the copy kernel
                rc = lib.eda_transpose_strided(_p(g), C2, C2 + C1, B, n, C1, _p(ds), ops._stream(dev))
                _chk(rc, "transpose_strided")
        return None, None, dk, ds


def fp_rows(dist2, idx, known_feats, unknow_feats):
    return _FPRowsFn.apply(dist2, idx, known_feats, unknow_feats)
