"""Host side of the fused set-abstraction kernel (eda_sa_mlp_forward): parameter folding, the
train-mode BatchNorm statistics passes, layout bookkeeping and the autograd boundary.

Forward is entirely hand-written CUDA: inference and no-grad calls run the fused TMEM-chained kernel (pack ->
[stats pass -> finalize] x3 in train mode -> fused gather+MLP+max-pool); calls under autograd run the row-major
formulation (`sa_forward_rows`: gather -> rows GEMM -> column statistics -> ... -> pool) that keeps the pre-activations
for the backward pass.  Backward is hand-written CUDA too (csrc/sa_bwd.cu + the GEMM kernels): the grouped rows are
rebuilt row-major, the three layers recomputed on the tcgen05 GEMM, and max-pool / ReLU / BatchNorm backward, the
activation and weight gradients and the feature scatter run as one-pass kernels (see `_sa_backward_cuda`).
EDA_BACKWARD=torch selects the older path (unfused CUDA ops + autograd through cuDNN), kept as a cross-check.
"""
import ctypes
import os

import torch
import torch.nn.functional as F

from .. import _lib
from . import _ext


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev):
    idx = dev.index
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(idx if idx is not None else torch.cuda.current_device()))


def transpose_last2(x):
    """(B,R,C) contiguous -> (B,C,R) contiguous, on the current stream."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    B, R, C = x.shape
    out = torch.empty((B, C, R), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().eda_transpose_last2(_p(x), B, R, C, _p(out), _stream(x.device))
    _lib.check(rc, "transpose_last2")
    return out


def point_major(features):
    """(B,C,N) -> a (B,N,C)-shaped tensor whose rows are contiguous, reusing the copy a previous fused
    layer left behind when there is one."""
    cached = getattr(features, "_eda_point_major", None)
    if cached is not None:
        pm, f_version, pm_version = cached
        # an in-place edit of either tensor since the copy was attached (scaling, masking, relu_ ...) bumps its
        # `_version`: the copy is stale then and is rebuilt from `features`
        if pm.shape == (features.size(0), features.size(2), features.size(1)) and pm.device == features.device \
                and features._version == f_version and pm._version == pm_version:
            return pm
    t = features.transpose(1, 2)
    if t.is_contiguous():
        return t
    return transpose_last2(features.contiguous())


def attach_point_major(features, pm):
    """Remembers `pm` (B,N,C) as the point-major copy of `features` (B,C,N), valid while neither is modified."""
    features._eda_point_major = (pm, features._version, pm._version)


def fusable(C, widths, nsample):
    if len(widths) != 3 or nsample is None:
        return False
    if nsample < 16 or (nsample & (nsample - 1)) != 0:
        return False
    return _lib.load().eda_sa_mlp_packed_floats(int(C), *[int(w) for w in widths]) > 0


def _bn_scale_shift(lib, dev, stats, count, bn, conv_bias, C, training_update, want_stats=False):
    """scale/shift of one layer.  bn None: scale = None (1), shift = conv bias (or None).  want_stats: also returns
    the (2, C) [mean, invstd] the scale/shift were built from (BatchNorm backward needs them).  Batch statistics of a
    synchronised layer (SyncBatchNorm semantics, eda_b200/syncbn.py) are summed over the ranks before they are finalised."""
    if bn is None:
        r = None, (conv_bias.detach().contiguous() if conv_bias is not None else None)
        return r + (None,) if want_stats else r
    from .. import rows_mlp

    scale, shift, mean_invstd = rows_mlp.bn_scale_shift(dev, stats, count, bn, C, training_update, want_stats=want_stats)
    return (scale, shift, mean_invstd) if want_stats else (scale, shift)


def sa_forward_raw(xyz, new_xyz, feat_pm, idx, layers, radius, normalize_xyz, training, state=None):
    """Runs the fused kernels.  layers = [(conv, bn_or_None)] * 3.  Returns out (B,M,C3) point-major.
    state: optional list that receives, per layer, (scale, shift, [mean, invstd]) for the backward pass."""
    lib = _lib.load()
    dev = xyz.device
    B, N, _ = xyz.shape
    M, S = idx.size(1), idx.size(2)
    C = 0 if feat_pm is None else feat_pm.size(2)
    widths = [conv.out_channels for conv, _ in layers]
    nfl = lib.eda_sa_mlp_packed_floats(C, *widths)
    assert nfl > 0
    Ws = [conv.weight.detach().reshape(conv.out_channels, -1).contiguous() for conv, _ in layers]
    assert Ws[0].size(1) == C + 3 and Ws[1].size(1) == widths[0] and Ws[2].size(1) == widths[1]
    packed = torch.empty(nfl, dtype=torch.float32, device=dev)
    feat_stride = 0
    if feat_pm is not None:
        assert feat_pm.stride(2) == 1 and feat_pm.stride(0) == N * feat_pm.stride(1)
        feat_stride = feat_pm.stride(1)
    stream = _stream(dev)
    scales, shifts = [None] * 3, [None] * 3
    count = float(B) * M * S

    def pack(nlayers):
        rc = lib.eda_sa_mlp_pack(_p(Ws[0]), _p(Ws[1]), _p(Ws[2]), _p(scales[0]), _p(scales[1]), _p(scales[2]), C,
                                 *widths, nlayers, _p(packed), stream)
        _lib.check(rc, "sa_mlp_pack")

    def run(stats_layer, out, stats):
        rc = lib.eda_sa_mlp_forward(_p(xyz), _p(new_xyz), _p(feat_pm), feat_stride, _p(idx), _p(packed),
                                    _p(shifts[0]), _p(shifts[1]), _p(shifts[2]), B, N, M, S, C, *widths,
                                    float(radius), 1 if normalize_xyz else 0, stats_layer, _p(out), _p(stats), stream)
        _lib.check(rc, "sa_mlp_forward")

    if not (training and any(bn is not None for _, bn in layers)):
        # eval mode: folded + packed weights are cached on the owning module
        owner = layers[0][0]
        packed_c, shifts_c = eval_packed(owner, layers, C, widths, dev)
        packed = packed_c
        shifts = list(shifts_c)
        with torch.cuda.device(dev):
            out = torch.empty((B, M, widths[2]), dtype=torch.float32, device=dev)
            run(0, out, None)
            if state is not None:
                for l, (conv, bn) in enumerate(layers):
                    state.append(_bn_scale_shift(lib, dev, None, 0.0, bn, conv.bias, widths[l], False, want_stats=True))
        return out
    with torch.cuda.device(dev):
        for l, (conv, bn) in enumerate(layers):
            batch_stats = training and bn is not None
            if batch_stats:
                # train-mode BatchNorm: batch statistics of layer l's conv output, layers < l already final
                stats = torch.empty(2 * widths[l], dtype=torch.float64, device=dev)  # fp64 sums (see csrc/sa_mlp.cu)
                pack(l + 1)
                run(l + 1, None, stats)
                scales[l], shifts[l], mi = _bn_scale_shift(lib, dev, stats, count, bn, conv.bias, widths[l], True,
                                                           want_stats=True)
            else:
                scales[l], shifts[l], mi = _bn_scale_shift(lib, dev, None, 0.0, bn, conv.bias, widths[l], False,
                                                           want_stats=True)
            if state is not None:
                state.append((scales[l], shifts[l], mi))
        out = torch.empty((B, M, widths[2]), dtype=torch.float32, device=dev)
        pack(3)
        run(0, out, None)
    return out


def _w1_permuted(W1, C, K0pad):
    """Layer-1 weight (C1, 3 + C) in the gathered column order [features | xyz | pad] (reference: [xyz | features])."""
    W1p = torch.zeros((W1.size(0), K0pad), dtype=torch.float32, device=W1.device)
    W1p[:, :C] = W1[:, 3:]
    W1p[:, C:C + 3] = W1[:, :3]
    return W1p


def sa_forward_rows(xyz, new_xyz, feat_pm, idx, layers, radius, normalize_xyz, training, state):
    """Training forward of one SA stage in the row-major formulation the backward pass uses, KEEPING the
    pre-activations: x0 = gather, z_l = relu(bn(z_{l-1})) W_l^T (eda_rows_gemm with the BatchNorm + ReLU prologue),
    batch statistics by eda_col_stats -> eda_bn_finalize, out = max-pool(relu(bn(z3))) (eda_sa_pool_forward).
    Same arithmetic class as the fused kernel (tf32 operands, fp32 accumulation); what it buys is that the backward
    pass starts from (x0, z1, z2, z3) instead of recomputing them (1.5 ms per step at B = 8).
    Returns out (B,M,C3) point-major and the tuple (x0, z1, z2, z3); `state` receives (scale, shift, [mean, invstd])."""
    from .. import attn_ops as ops

    lib = _lib.load()
    dev = xyz.device
    stream = _stream(dev)
    B, N, _ = xyz.shape
    M, S = idx.size(1), idx.size(2)
    C = 0 if feat_pm is None else feat_pm.size(2)
    widths = [conv.out_channels for conv, _ in layers]
    K0pad = ((C + 3 + 15) // 16) * 16
    R = B * M * S
    feat_stride = 0 if feat_pm is None else feat_pm.stride(1)
    Ws = [conv.weight.detach().reshape(conv.out_channels, -1) for conv, _ in layers]
    Wl = [_w1_permuted(Ws[0], C, K0pad), Ws[1], Ws[2]]
    with torch.cuda.device(dev):
        x0 = torch.empty((R, K0pad), dtype=torch.float32, device=dev)
        rc = lib.eda_sa_gather_rows(_p(xyz), _p(new_xyz), _p(feat_pm), feat_stride, _p(idx), B, N, M, S, C, K0pad,
                                    float(radius), 1 if normalize_xyz else 0, _p(x0), stream)
        _lib.check(rc, "sa_gather_rows")
        zs, xin, sc, sh = [], x0, None, None
        for l, (conv, bn) in enumerate(layers):
            if training:
                # batch statistics of this layer's output are taken in the GEMM epilogue (no second pass over z)
                stats = torch.zeros(2 * widths[l], dtype=torch.float64, device=dev)
                z = ops.rows_gemm(xin, Wl[l], in_scale=sc, in_shift=sh, stats=stats)
                sc, sh, mi = _bn_scale_shift(lib, dev, stats, float(R), bn, conv.bias, widths[l], True, want_stats=True)
            else:
                z = ops.rows_gemm(xin, Wl[l], in_scale=sc, in_shift=sh)
                sc, sh, mi = _bn_scale_shift(lib, dev, None, 0.0, bn, conv.bias, widths[l], False, want_stats=True)
            state.append((sc, sh, mi))
            zs.append(z)
            xin = z
        out = torch.empty((B, M, widths[2]), dtype=torch.float32, device=dev)
        amax = torch.empty((B * M, widths[2]), dtype=torch.int32, device=dev)
        rc = lib.eda_sa_pool_forward(_p(zs[2]), _p(sc), _p(sh), B * M, S, widths[2], _p(out), _p(amax), stream)
        _lib.check(rc, "sa_pool_forward")
    return out, (x0, zs[0], zs[1], zs[2], amax)


def _composed(xyz, new_xyz, features, idx, params, has_bn, radius, normalize_xyz, use_batch_stats, running, eps):
    """Unfused, differentiable restatement (used for the backward pass only)."""
    from . import pointnet2_utils as pu

    grouped_xyz = pu.grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        grouped_xyz = grouped_xyz / radius
    x = grouped_xyz if features is None else torch.cat([grouped_xyz, pu.grouping_operation(features, idx)], dim=1)
    it = iter(params)
    for l in range(3):
        w = next(it)
        bias = next(it)
        x = F.conv2d(x, w, bias)
        if has_bn[l]:
            g, b = next(it), next(it)
            if use_batch_stats:
                x = F.batch_norm(x, None, None, g, b, True, 0.0, eps[l])
            else:
                x = F.batch_norm(x, running[l][0], running[l][1], g, b, False, 0.0, eps[l])
        x = F.relu(x)
    return F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)


class FusedSAFunction(torch.autograd.Function):
    """new_features (B,C3,M) = maxpool(MLP(group(xyz, features, idx)))."""

    @staticmethod
    def forward(ctx, module, xyz, new_xyz, features, idx, *params):
        layers = module.mlp_module.fusable_layers()
        feat_pm = None if features is None else point_major(features)
        training = module.training
        has_bn = [bn is not None for _, bn in layers]
        # BN buffers as they are BEFORE this forward (eval-mode backward needs the ones that were used)
        ctx.running = [(bn.running_mean.clone(), bn.running_var.clone()) if (bn is not None and not training) else None
                       for _, bn in layers]
        ctx.cuda_bw = (any(ctx.needs_input_grad) and all(has_bn) and os.environ.get("EDA_BACKWARD", "cuda") != "torch")
        ctx.bns = [bn for _, bn in layers]
        ctx.state = [] if ctx.cuda_bw else None
        if ctx.cuda_bw:
            from .. import attn_ops
            # conv weights of layers 2 / 3 whose gradient buffers exist: their wgrad kernels accumulate straight
            # into them from the side stream (attn_ops.fused_grad_enabled); layer 1 needs a column permutation, so it stays
            ctx.gbufs = attn_ops._grad_buffers((params[4], params[8], params[0]))
            attn_ops.grads_expected([p for p, g in zip((params[4], params[8], params[0]), ctx.gbufs) if g is not None])
        ctx.rows = None
        if ctx.cuda_bw and training and os.environ.get("EDA_SA_RECOMPUTE", "0") != "1":
            # training mode: keep the row-major pre-activations for the backward pass (EDA_SA_RECOMPUTE=1, and eval mode
            # under autograd: fused forward kernel, the backward recomputes them — 1.9 GB less activation memory at
            # B = 8, 1 ms more per step)
            out_pm, ctx.rows = sa_forward_rows(xyz, new_xyz.contiguous(), feat_pm, idx, layers, module.radius,
                                               module.normalize_xyz, training, ctx.state)
        else:
            out_pm = sa_forward_raw(xyz, new_xyz, feat_pm, idx, layers, module.radius, module.normalize_xyz, training,
                                    state=ctx.state)
        out = transpose_last2(out_pm)
        ctx.save_for_backward(xyz, new_xyz, features, idx, *params)
        ctx.meta = (has_bn, float(module.radius), bool(module.normalize_xyz), training,
                    [bn.eps if bn is not None else 0.0 for _, bn in layers])
        ctx.mark_non_differentiable(out_pm)
        return out, out_pm

    @staticmethod
    def backward(ctx, grad_out, _grad_pm=None):
        if ctx.cuda_bw:
            return _sa_backward_cuda(ctx, grad_out)
        xyz, new_xyz, features, idx, *params = ctx.saved_tensors
        has_bn, radius, normalize_xyz, training, eps = ctx.meta
        with torch.enable_grad():
            f = None if features is None else features.detach().requires_grad_(ctx.needs_input_grad[3])
            ps = [None if p is None else p.detach().requires_grad_(True) for p in params]
            out = _composed(xyz, new_xyz, f, idx, ps, has_bn, radius, normalize_xyz, training, ctx.running, eps)
            wanted = [t for t in ([f] + ps) if t is not None and t.requires_grad]
            grads = torch.autograd.grad(out, wanted, grad_out, allow_unused=True)
        gmap = {id(t): g for t, g in zip(wanted, grads)}
        gf = gmap.get(id(f)) if f is not None else None
        gps = [gmap.get(id(p)) if p is not None else None for p in ps]
        return (None, None, None, gf, None, *gps)


def _sa_backward_cuda(ctx, grad_out):
    """Backward of one fused SA stage on the CUDA kernels.  Rows R = B*M*S, row-major pre-activations z_l (R, C_l):
         recompute   x0 = gather;  z_l = relu(bn(z_{l-1})) W_l^T  — eda_rows_gemm, the previous layer's folded BatchNorm +
                     ReLU applied as its prologue, so the post-activation tensors never exist in HBM
         layer 3     arg-max over the S rows of a centre + ReLU gate -> dy3 (one row per centre and channel)
         BatchNorm   dz_l = scale_l (dy_l - mean(dy_l) - zhat_l mean(dy_l zhat_l)); d beta = sum dy, d gamma = sum dy zhat
         GEMMs       dW_l = dz_l^T relu(bn(z_{l-1})) (eda_wgrad, same prologue), da_{l-1} = dz_l W_l (eda_rows_gemm, W^T by strides)
         features    d features[idx] += dx0[:, :C]"""
    from .. import attn_ops as ops

    lib = _lib.load()
    xyz, new_xyz, features, idx, *params = ctx.saved_tensors
    has_bn, radius, normalize_xyz, training, eps = ctx.meta
    dev = xyz.device
    stream = _stream(dev)
    B, N, _ = xyz.shape
    M, S = idx.size(1), idx.size(2)
    C = 0 if features is None else features.size(1)
    # params: per layer conv.weight, conv.bias, bn.weight, bn.bias (every layer has a BatchNorm on this path)
    Ws = [params[4 * l].detach().reshape(params[4 * l].size(0), -1) for l in range(3)]
    widths = [w.size(0) for w in Ws]
    K0pad = ((C + 3 + 15) // 16) * 16
    R = B * M * S
    feat_pm = None if features is None else point_major(features.detach())
    feat_stride = 0 if feat_pm is None else feat_pm.stride(1)
    new_xyz = new_xyz.contiguous()
    f32 = dict(dtype=torch.float32, device=dev)

    def chk(rc, what):
        _lib.check(rc, what)

    with torch.cuda.device(dev):
        W1p = _w1_permuted(Ws[0], C, K0pad)
        Wl = [W1p, Ws[1], Ws[2]]
        Kin = [K0pad, widths[0], widths[1]]
        # input of layer l as (tensor, scale, shift): layer 1 reads x0 as it is, layers 2 / 3 read relu(bn(z_{l-1}))
        amax = None
        if ctx.rows is not None:
            # the training forward kept them (sa_forward_rows); z3 is overwritten in place below, so they serve once
            x0, z0, z1, z2, amax = ctx.rows
            ctx.rows = None
            z = [z0, z1, z2]
            src = [(x0, None, None), (z0, ctx.state[0][0], ctx.state[0][1]), (z1, ctx.state[1][0], ctx.state[1][1])]
        else:
            # ---- recompute the forward, row-major ------------------------------------------------------------
            x0 = torch.empty((R, K0pad), **f32)
            chk(lib.eda_sa_gather_rows(_p(xyz), _p(new_xyz), _p(feat_pm), feat_stride, _p(idx), B, N, M, S, C, K0pad,
                                       float(radius), 1 if normalize_xyz else 0, _p(x0), stream), "sa_gather_rows")
            z = [None] * 3
            src = [(x0, None, None)]
            for l in range(3):
                xin, sc, sh = src[l]
                z[l] = ops.rows_gemm(xin, Wl[l], in_scale=sc, in_shift=sh)
                if l < 2:
                    src.append((z[l], ctx.state[l][0], ctx.state[l][1]))
        # ---- layer 3: max-pool + ReLU + BatchNorm backward -----------------------------------------------------
        g_pm = transpose_last2(grad_out.contiguous())  # (B, M, C3)
        from .. import rows_mlp
        stats = [torch.zeros(2 * widths[l], **f32) for l in range(3)]
        local_stats = list(stats)  # per layer [sum dy, sum dy zhat] of THIS rank's rows = the BatchNorm affine gradients
        bns = ctx.bns
        dWl = [torch.zeros((widths[l], Kin[l]), **f32) for l in range(3)]
        scale, shift, mi = ctx.state[2]
        if amax is not None:  # saved by the training forward: only the reductions are left to do
            chk(lib.eda_sa_pool_backward_stats(_p(z[2]), _p(amax), _p(mi[0]), _p(mi[1]), _p(g_pm), B * M, S, widths[2],
                                               _p(stats[2]), stream), "sa_pool_backward_stats")
        else:
            amax = torch.empty((B * M, widths[2]), dtype=torch.int32, device=dev)
            chk(lib.eda_sa_pool_backward(_p(z[2]), _p(scale), _p(shift), _p(mi[0]), _p(mi[1]), _p(g_pm), B * M, S,
                                         widths[2], _p(amax), _p(stats[2]), stream), "sa_pool_backward")
        count = float(R)
        if training:  # synchronised BatchNorm: the reductions (and the row count) cover all ranks
            stats[2], count, local_stats[2] = rows_mlp.bn_backward_reduce(bns[2], stats[2], float(R))
        chk(lib.eda_sa_pool_backward_apply(_p(z[2]), _p(amax), _p(g_pm), _p(scale), _p(mi[0]), _p(mi[1]), _p(stats[2]),
                                           count, 1 if training else 0, B * M, S, widths[2], stream),
            "sa_pool_backward_apply")
        dz = z[2]  # overwritten in place
        z[2] = None
        fused_w = [None, ctx.gbufs[0], ctx.gbufs[1]]
        g1 = ctx.gbufs[2]  # layer 1: scratch + column permutation, but also off the critical path when a buffer exists
        for l in (2, 1, 0):
            xin, sc, sh = src[l]
            if l == 0 and g1 is not None:
                cur, side = torch.cuda.current_stream(dev), ops._wgrad_side(dev)
                side.wait_stream(cur)
                ops.mark_side_pending(dev)
                with torch.cuda.stream(side):
                    ops.wgrad([dict(dy=dz, x=xin, dw=dWl[0])], widths[0], Kin[0])
                    gv = g1.view(widths[0], C + 3)
                    gv[:, :3] += dWl[0][:, C:C + 3]
                    if C:
                        gv[:, 3:] += dWl[0][:, :C]
                for t in (dz, xin, dWl[0]):
                    t.record_stream(side)
            elif fused_w[l] is not None:
                ops.wgrad_side([dict(dy=dz, x=xin, dw=fused_w[l].view(widths[l], Kin[l]), x_scale=sc, x_shift=sh)],
                               widths[l], Kin[l])
            else:
                ops.wgrad([dict(dy=dz, x=xin, dw=dWl[l], x_scale=sc, x_shift=sh)], widths[l], Kin[l])
            if l == 0:
                break
            da = ops.rows_gemm(dz, Wl[l], transpose=True)
            scale, shift, mi = ctx.state[l - 1]
            chk(lib.eda_bn_relu_backward_stats(_p(da), _p(z[l - 1]), _p(scale), _p(shift), _p(mi[0]), _p(mi[1]), R,
                                               widths[l - 1], _p(stats[l - 1]), stream), "bn_relu_backward_stats")
            count = float(R)
            if training:
                stats[l - 1], count, local_stats[l - 1] = rows_mlp.bn_backward_reduce(bns[l - 1], stats[l - 1], float(R))
            chk(lib.eda_bn_relu_backward_apply(_p(da), _p(z[l - 1]), _p(scale), _p(shift), _p(mi[0]), _p(mi[1]),
                                               _p(stats[l - 1]), count, 1 if training else 0, R, widths[l - 1], stream),
                "bn_relu_backward_apply")
            dz = da
        gf = None
        if features is not None and ctx.needs_input_grad[3]:
            dx0 = ops.rows_gemm(dz, W1p, transpose=True)
            dfeat_pm = torch.zeros((B, N, C), **f32)
            chk(lib.eda_sa_scatter_rows(_p(dx0), _p(idx), B, N, M, S, C, K0pad, _p(dfeat_pm), stream), "sa_scatter_rows")
            gf = transpose_last2(dfeat_pm)
    ops.grads_written([p for p, g in zip((params[4], params[8], params[0]), ctx.gbufs) if g is not None])
    gps = []
    for l in range(3):
        w = params[4 * l]
        if l == 0:
            dW = None if g1 is not None else torch.cat([dWl[0][:, C:C + 3], dWl[0][:, :C]], dim=1)
        else:
            dW = None if fused_w[l] is not None else dWl[l]
        gps += [None if dW is None else dW.reshape(w.shape), None, local_stats[l][widths[l]:], local_stats[l][:widths[l]]]
    return (None, None, None, gf, None, *gps)


def sa_params(layers):
    """Flat parameter list in the order _composed consumes it: per layer conv.weight, conv.bias (or None),
    then bn.weight, bn.bias when the layer has a BatchNorm."""
    out = []
    for conv, bn in layers:
        out.append(conv.weight)
        out.append(conv.bias)
        if bn is not None:
            out.append(bn.weight)
            out.append(bn.bias)
    return out


# ---------------------------------------------------------------------------------------------------
# Pipelined inference path: ball query + fused MLP on the centres FPS has already produced
# ---------------------------------------------------------------------------------------------------
class ProgressCounter:
    """Device words that eda_furthest_point_sampling_progress increments (never reset) — ONE WORD PER MILESTONE, so
    "word j reached its previous total + B" means every scene has passed milestone j, however the scenes were
    scheduled — plus the host-side running totals, so every launch knows which absolute values it will reach."""
    _by_device = {}
    MAX_MARKS = 64      # milestones per launch
    RING = 4096         # words; consecutive launches take consecutive slices, so launches in flight at the same time
                        # (other streams, other models) never share a word

    def __init__(self, device):
        self.words = torch.zeros(self.RING, dtype=torch.int32, device=device)
        self.totals = [0] * self.RING
        self.next = 0

    def take(self, n):
        """First word index of a fresh slice of n words."""
        if self.next + n > self.RING:
            self.next = 0
        first = self.next
        self.next += n
        return first

    @classmethod
    def get(cls, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        pc = cls._by_device.get(key)
        if pc is None:
            pc = cls._by_device[key] = ProgressCounter(device)
        return pc


class PipelinedFPS:
    """Handle of an in-flight furthest-point sampling launched with progress milestones every `every` samples."""

    def __init__(self, inds, counter, first, targets, every, done_event):
        self.inds, self.counter, self.targets, self.every, self.done_event = inds, counter, targets, every, done_event
        self.first = first
        self.B, self.m = inds.shape
        self.nchunks = len(targets)

    def wait_chunk(self, stream, j):
        """Makes `stream` wait (on the device, no SM occupied) until centres [0, (j+1)*every) of EVERY scene exist."""
        word = ctypes.c_void_p(self.counter.words.data_ptr() + 4 * (self.first + j))
        rc = _lib.load().eda_stream_wait_value32(ctypes.c_void_p(stream.cuda_stream), word, int(self.targets[j]))
        _lib.check(rc, "stream_wait_value32")


def launch_pipelined_fps(xyz, npoint, every, stream, not_identity=None):
    """Launches FPS(xyz, npoint) on `stream` with progress milestones.  Returns a PipelinedFPS.  `not_identity`:
    optional device flags from _ext.fps_identity_flags (verified scenes publish all milestones at once)."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    counter = ProgressCounter.get(xyz.device)
    nchunks = (npoint + every - 1) // every
    if nchunks > ProgressCounter.MAX_MARKS:
        raise RuntimeError(f"eda_b200: at most {ProgressCounter.MAX_MARKS} FPS progress milestones per launch")
    first = counter.take(nchunks)
    with torch.cuda.stream(stream):
        inds = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        nbytes = lib.eda_fps_scratch_bytes(B, N, npoint)
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=xyz.device) if nbytes else None
        with torch.cuda.device(xyz.device):
            rc = lib.eda_furthest_point_sampling_ex(_p(xyz), B, N, npoint, _p(scratch), _p(inds),
                                                    ctypes.c_void_p(counter.words.data_ptr() + 4 * first),
                                                    int(every), _p(not_identity), ctypes.c_void_p(stream.cuda_stream))
        _lib.check(rc, "furthest_point_sampling_progress")
        targets = []
        for j in range(first, first + nchunks):  # every scene adds exactly 1 to each of its milestone words
            counter.totals[j] = (counter.totals[j] + B) & 0xFFFFFFFF
            t = counter.totals[j]
            targets.append(t - (1 << 32) if t >= (1 << 31) else t)
        done = torch.cuda.Event()
        done.record(stream)
    return PipelinedFPS(inds, counter, first, targets, every, done)


def eval_packed(module, layers, C, widths, dev):
    """Eval-mode (running-statistics) folded + packed weights and per-layer shifts of a fused SA module, cached on
    the module until a parameter or BatchNorm buffer changes (in-place update bumps `_version`; .to() / load
    reallocates)."""
    lib = _lib.load()
    tensors = []
    for conv, bn in layers:
        tensors += [conv.weight, conv.bias]
        if bn is not None:
            tensors += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
    key = tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors) + (C, str(dev))
    from .. import attn_ops
    # inside a GraphedTrainStep capture of this model the fold + pack must be recorded, not reused
    use_cache = attn_ops.PACK_CACHE and attn_ops.active_registry(module) is None
    hit = module.__dict__.get("_eda_sa_eval_cache") if use_cache else None
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    nfl = lib.eda_sa_mlp_packed_floats(C, *widths)
    Ws = [conv.weight.detach().reshape(conv.out_channels, -1).contiguous() for conv, _ in layers]
    scales, shifts = [None] * 3, [None] * 3
    with torch.cuda.device(dev):
        for l, (conv, bn) in enumerate(layers):
            scales[l], shifts[l] = _bn_scale_shift(lib, dev, None, 0.0, bn, conv.bias, widths[l], False)
        packed = torch.empty(nfl, dtype=torch.float32, device=dev)
        rc = lib.eda_sa_mlp_pack(_p(Ws[0]), _p(Ws[1]), _p(Ws[2]), _p(scales[0]), _p(scales[1]), _p(scales[2]), C, *widths,
                                 3, _p(packed), _stream(dev))
        _lib.check(rc, "sa_mlp_pack")
    module.__dict__["_eda_sa_eval_cache"] = (key, packed, shifts)
    return packed, shifts


def sa_forward_pipelined(module, xyz, features, fps):
    """Eval-mode fused set abstraction consuming `fps` chunk by chunk on the current stream: for every chunk of
    centres, wait for the sampler's milestone, ball-query it, run the fused group + MLP + max-pool on it.
    Results are identical to the unpipelined path (same kernels, same arithmetic).
    Returns new_xyz (B,M,3), new_features (B,C3,M), feat_pm (B,M,C3), inds (B,M)."""
    lib = _lib.load()
    dev = xyz.device
    layers = module.mlp_module.fusable_layers()
    assert layers is not None and len(layers) == 3 and not module.training
    B, N, _ = xyz.shape
    M, S = fps.m, module.nsample
    feat_pm = None if features is None else point_major(features)
    C = 0 if feat_pm is None else feat_pm.size(2)
    widths = [conv.out_channels for conv, _ in layers]
    nfl = lib.eda_sa_mlp_packed_floats(C, *widths)
    assert nfl > 0
    main = torch.cuda.current_stream(dev)
    stream = _stream(dev)
    packed, shifts = eval_packed(layers[0][0], layers, C, widths, dev)
    with torch.cuda.device(dev):
        idx = torch.empty((B, M, S), dtype=torch.int32, device=dev)
        out = torch.zeros((B, M, widths[2]), dtype=torch.float32, device=dev)
        feat_stride = 0 if feat_pm is None else feat_pm.stride(1)
        fps.inds.record_stream(main)
        for j in range(fps.nchunks):
            m0 = j * fps.every
            mc = min(fps.every, M - m0)
            fps.wait_chunk(main, j)
            rc = lib.eda_ball_query_range(_p(xyz), _p(fps.inds), B, N, M, m0, mc, float(module.radius), S, _p(idx), stream)
            _lib.check(rc, "ball_query_range")
            rc = lib.eda_sa_mlp_forward_range(_p(xyz), _p(fps.inds), _p(feat_pm), feat_stride, _p(idx), _p(packed),
                                              _p(shifts[0]), _p(shifts[1]), _p(shifts[2]), B, N, M, m0, mc, S, C, *widths,
                                              float(module.radius), 1 if module.normalize_xyz else 0, _p(out), stream)
            _lib.check(rc, "sa_mlp_forward_range")
    main.wait_event(fps.done_event)
    new_xyz = torch.gather(xyz, 1, fps.inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    new_features = transpose_last2(out)
    attach_point_major(new_features, out)
    return new_xyz, new_features, out, fps.inds
