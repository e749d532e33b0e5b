"""BASELINE.json configs[3]: one TRAINING STEP (forward + backward) of the hot path at the metric's shapes — B = 8,
N = 50 000, L = 80, D = 132, K = 256 — against the reference CUDA build: the reference's own Python modules
(Pointnet2Backbone, BiEncoder, BiDecoderLayer; bytecode in oracle/_ref/pyref) + its own compiled `_ext` (oracle/_ref),
fp32-pinned, same parameters, same inputs, train-mode BatchNorm, dropout 0.

What is asserted, and why the gradient bounds are what they are:
  * loss equal to 2e-5 relative; outputs within the forward tolerances.
  * analytically-zero gradients (a bias in front of a train-mode BatchNorm; the key-projection bias of softmax attention)
    are ~0 in absolute terms instead of being compared noise against noise.
  * every parameter gradient within a bound tied to the reference's OWN tf32-vs-fp32 spread: the reference is run a
    second time with TF32 allowed (cudnn / cuBLAS), and our error against the fp32 run must stay within a small factor of
    that spread (tensor-core rounding flips max-pool arg-maxes and ReLU gates; that is a property of tf32, not of these
    kernels).
  * test_sa_gradients_match_fp32_when_gates_are_forced proves that attribution directly: with the max-pool selections
    and ReLU gates taken from our forward, an fp32 autograd evaluation of the same stage agrees with our backward kernels
    to tf32 rounding (<= 5e-3), while the free-running fp32 evaluation differs by the flip noise.
"""
import json
import os

import pytest
import torch

from eda_b200 import hotpath

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _ref_hotpath():
    from oracle import ref_hotpath, ref_loader, ref_model

    ext = ref_loader.load_reference_ext()
    if ext is None or ref_model.ref_dir() is None:
        pytest.skip("oracle/_ref (compiled reference _ext + reference bytecode) did not travel")
    return ref_hotpath.build(ext, dropout=0.0)


def _grads(model, inputs):
    for p in model.parameters():
        p.grad = None
    out = model(*inputs)
    loss = hotpath.quadratic_loss(out)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), [o.detach() for o in out], {n: p.grad.detach().clone() for n, p in model.named_parameters()}


def test_train_step_matches_reference_cuda_build():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    ours = hotpath.HotPath(dropout=0.0).cuda().train()
    ref = _ref_hotpath().cuda().train()
    ref.load_state_dict(ours.state_dict(), strict=True)
    inputs = [t.cuda() for t in hotpath.synthetic_inputs(8, 50000, 80, 132, 256, seed=100)]

    loss_r, out_r, g_r = _grads(ref, inputs)
    loss_o, out_o, g_o = _grads(ours, inputs)
    # the reference's own tf32 spread (same code, TF32 allowed in cuDNN / cuBLAS)
    ref.load_state_dict(ours.state_dict(), strict=False)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        loss_t, out_t, g_t = _grads(ref, inputs)
    finally:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    assert abs(loss_o - loss_r) <= 2e-5 * abs(loss_r), (loss_o, loss_r)
    # outputs: train-mode BatchNorm over random-init layers amplifies operand rounding along the 16-layer backbone, for
    # the reference just as much as for us (measured on B200: fp2_features max 0.071 / rms 7.0e-3 between the
    # reference's own tf32 and fp32 runs, 0.085 / 8.1e-3 for this package; the encoder alone on identical inputs agrees
    # to 1.4e-3) -> the bound is tied to the reference's own spread, with a floor at the eval-mode tolerances
    out_report = {}
    for name, a, b, t in zip(("query", "vis", "text"), out_o, out_r, out_t):
        d, dt = (a - b).abs(), (t - b).abs()
        mo, ro = d.max().item(), d.pow(2).mean().sqrt().item()
        mt, rt = dt.max().item(), dt.pow(2).mean().sqrt().item()
        out_report[name] = (mo, ro, mt, rt)

    # ---- analytically zero gradients: measured here, asserted ~0 below, excluded from the relative comparison ----
    zero = [n for n in g_r if n.endswith("position_embedding_head.0.bias")]      # bias in front of train-mode BatchNorm
    E = 288
    kbias = [n for n in g_r if n.endswith("in_proj_bias")]                        # softmax is shift-invariant per row
    zero_report = {}
    for n in zero:
        wn = n.replace(".0.bias", ".0.weight")
        zero_report[n] = (g_o[n].abs().max().item(), g_o[wn].abs().max().item(), g_r[n].abs().max().item())
    for n in kbias:
        zero_report[n + "[k]"] = (g_o[n][E:2 * E].abs().max().item(), g_o[n][:E].abs().max().item(),
                                  g_r[n][E:2 * E].abs().max().item())

    # ---- every other gradient, against the reference's own tf32-vs-fp32 spread ----
    rows = {}
    for n in g_r:
        if n in zero:
            continue
        a, b, t = g_o[n], g_r[n], g_t[n]
        if n in kbias:
            keep = torch.ones(3 * E, dtype=torch.bool, device=a.device)
            keep[E:2 * E] = False
            a, b, t = a[keep], b[keep], t[keep]
        rows[n] = (rel(a, b), rel(t, b))
    ours_err = sorted(v[0] for v in rows.values())
    tf32_err = sorted(v[1] for v in rows.values())
    med_o, med_t = ours_err[len(ours_err) // 2], tf32_err[len(tf32_err) // 2]
    worst = sorted(rows.items(), key=lambda kv: -kv[1][0])[:8]
    print(f"train step: loss ours {loss_o:.8f} ref {loss_r:.8f} ref-tf32 {loss_t:.8f}; grad rel err median ours {med_o:.2e} "
          f"(reference tf32 spread {med_t:.2e}); worst ours {worst}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "train_step_parity.json"), "w") as f:
            json.dump({"loss_ours": loss_o, "loss_ref_fp32": loss_r, "loss_ref_tf32": loss_t, "median_ours": med_o,
                       "outputs_max_rms_ours_then_ref_tf32": {k: list(v) for k, v in out_report.items()},
                       "median_ref_tf32": med_t, "per_parameter": {k: list(v) for k, v in rows.items()},
                       "analytically_zero_absmax_ours__scale__absmax_ref": {k: list(v) for k, v in zero_report.items()}},
                      f, indent=1)
    for name, (mo, ro, mt, rt) in out_report.items():
        assert mo <= max(2.5 * mt, 5e-2) and ro <= max(2.5 * rt, 5e-3), (name, mo, ro, mt, rt)
    assert len(zero) == 6
    for n, (z, scale, _zr) in zero_report.items():
        # |grad| of an analytically-zero entry relative to the live gradient next to it (the first conv's weight gradient
        # / the query-bias gradient of the same in-projection)
        assert z <= 1e-3 * max(scale, 1e-12), (n, z, scale)
    assert med_o <= 5e-3, med_o
    assert med_o <= 4 * max(med_t, 5e-4), (med_o, med_t)
    for n, (eo, et) in rows.items():
        # per tensor: within 4x the reference's own tf32 spread for that tensor, or 2e-2 — whichever is larger
        assert eo <= max(4 * et, 2e-2), (n, eo, et)


@pytest.mark.parametrize("shape", ["sa1_like", "sa2_like"])
def test_sa_gradients_match_fp32_when_gates_are_forced(shape):
    """DESIGN.md section 9's attribution, demonstrated: the only reason the fused SA stage's gradients differ from an
    fp32 evaluation by percents is that tf32 rounding resolves some max-pool arg-maxes / ReLU gates differently.  Take
    the gates and selections from OUR forward (the saved pre-activations and arg-max), evaluate everything else in fp32
    autograd, and the gradients agree to tf32 rounding."""
    from eda_b200 import synthetic
    from eda_b200.pointnet2 import fused, pointnet2_utils
    from eda_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if shape == "sa1_like":
        B, N, C, M, S, mlp, radius = 2, 20000, 3, 512, 64, [3, 64, 64, 128], 0.2
    else:
        B, N, C, M, S, mlp, radius = 2, 2048, 128, 256, 32, [128, 128, 128, 256], 0.4
    torch.manual_seed(1)
    m = PointnetSAModuleVotes(npoint=M, radius=radius, nsample=S, mlp=list(mlp), use_xyz=True, normalize_xyz=True).cuda().train()
    g = torch.Generator().manual_seed(2)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data.copy_(1 + 0.2 * torch.randn(mod.num_features, generator=g))
            mod.bias.data.copy_(0.2 * torch.randn(mod.num_features, generator=g))
    xyz = synthetic.point_clouds(B, N, "surface", channels=0).cuda().contiguous()
    feats = torch.randn(B, C, N, generator=g).cuda().requires_grad_(True)
    new_xyz, out, inds = m(xyz, feats)
    go = torch.randn(out.shape, generator=g).cuda()
    out.backward(go)
    torch.cuda.synchronize()
    ours = {n: p.grad.clone() for n, p in m.named_parameters()}
    ours["features"] = feats.grad.clone()

    # the same forward once more, keeping what the backward kernels started from (deterministic kernels: same values)
    layers = m.mlp_module.fusable_layers()
    idx = pointnet2_utils.ball_query(radius, S, xyz, new_xyz)
    state = []
    with torch.no_grad():
        out_pm, (x0, z1, z2, z3, amax) = fused.sa_forward_rows(xyz, new_xyz.contiguous(), fused.point_major(feats.detach()),
                                                               idx, layers, radius, True, True, state)
    assert torch.equal(fused.transpose_last2(out_pm), out.detach())
    K0pad = x0.size(1)
    R = x0.size(0)

    def evaluate(forced):
        ps = {n: p.detach().clone().requires_grad_(True) for n, p in m.named_parameters()}
        W = [ps[f"mlp_module.layer{l}.conv.weight"].reshape(mlp[l + 1], -1) for l in range(3)]
        W1p = torch.zeros(mlp[1], K0pad, device="cuda")
        W1p = torch.cat([W[0][:, 3:], W[0][:, :3], W1p[:, C + 3:]], 1)  # gathered column order [features | xyz | pad]
        x = x0.clone().requires_grad_(True)
        a, zs = x, [z1, z2, z3]
        y = None
        for l in range(3):
            z = a @ (W1p if l == 0 else W[l]).t()
            mean, var = z.mean(0), z.var(0, unbiased=False)
            y = (z - mean) * torch.rsqrt(var + 1e-5) * ps[f"mlp_module.layer{l}.bn.bn.weight"] + ps[f"mlp_module.layer{l}.bn.bn.bias"]
            if l < 2:
                if forced:
                    sc, sh, _ = state[l]
                    a = y * (zs[l] * sc + sh > 0)      # OUR gate
                else:
                    a = torch.relu(y)
        C3 = mlp[3]
        y3 = y.view(B * M, S, C3)
        if forced:
            sel = amax.long().clamp_min(0).unsqueeze(1)                      # OUR arg-max (first row attaining it)
            pooled = torch.gather(y3, 1, sel).squeeze(1) * (amax >= 0)       # -1: nothing positive -> ReLU kills it
        else:
            pooled = torch.relu(y3).max(1).values
        (pooled * fused.transpose_last2(go.contiguous()).reshape(B * M, C3)).sum().backward()
        got = {n: p.grad for n, p in ps.items() if p.grad is not None}
        dfeat = torch.zeros(B * N, C, device="cuda")
        flat_idx = (idx.long() + (torch.arange(B, device="cuda") * N).view(B, 1, 1)).reshape(-1)
        dfeat.index_add_(0, flat_idx, x.grad[:, :C])
        got["features"] = dfeat.view(B, N, C).transpose(1, 2)
        return got

    forced, free = evaluate(True), evaluate(False)
    report = {n: (rel(ours[n], forced[n]), rel(ours[n], free[n])) for n in forced}
    print(f"{shape}: rel err (gates forced, free-running fp32): {report}")
    for n, (ef, _) in report.items():
        assert ef <= 5e-3, (n, ef)           # tf32 operand rounding only
    # and the free-running comparison is what carries the flip noise (at least as large for the bulk of the tensors)
    assert sum(ef <= efree + 1e-6 for ef, efree in report.values()) >= len(report) - 1
