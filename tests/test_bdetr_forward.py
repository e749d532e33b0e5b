"""BASELINE.json configs[2]: the FULL BeaUTyDETR.forward through the reference's UNMODIFIED models/bdetr.py
(models/bdetr.py:208-339), once on this repo's modules (swapped in as INTEGRATION.md sections 2-3 describe) and once on
the all-reference stack (reference Python modules + the reference's own compiled `_ext`, oracle/_ref), same state dict,
same synthetic batch, eval mode, fp32-pinned reference.

Stage-wise parity (SURVEY.md section 7 risk "top-k flips"): backbone indices / coordinates exact, features within the
stated tf32 tolerance, KPS logits within tolerance and top-k sets overlapping; then the drop-in model is TEACHER-FORCED
with the reference's `query_points_sample_inds` and every decoder layer's output and every prediction head's output is
compared.

The reference files reach the GPU box as sourceless bytecode (oracle/build_ref.py -> oracle/_ref/pyref); RoBERTa is a
seeded random-init model of roberta-base width (no weights offline; the frozen text tower is outside the path and is
the SAME torch code in both models)."""
import os

import pytest
import torch

from oracle import ref_loader, ref_model

pytestmark = pytest.mark.gpu

TOL_FEAT_MAX, TOL_FEAT_RMS = 2e-2, 3e-3      # tf32 tensor-core contractions vs fp32, chained (relative to max |ref|)
TOL_LAYER_MAX, TOL_LAYER_RMS = 5e-2, 5e-3    # LayerNorm-ed decoder outputs after 6 chained layers


def _err(a, b):
    a, b = a.float(), b.float()
    scale = max(1.0, b.abs().max().item())
    d = (a - b).abs()
    return d.max().item() / scale, d.pow(2).mean().sqrt().item() / scale


def _models(roberta_layers=2):
    ext = ref_loader.load_reference_ext()
    if ext is None or ref_model.ref_dir() is None:
        pytest.skip("oracle/_ref (compiled reference _ext + reference bytecode) did not travel")
    ref = ref_model.build_bdetr(ref_model.load("reference", ext), roberta_layers=roberta_layers)
    eda = ref_model.build_bdetr(ref_model.load("eda"), roberta_layers=roberta_layers)
    assert type(ref.backbone_net).__module__ == "models.backbone_module"
    assert type(eda.backbone_net).__module__ == "eda_b200.backbone_module"
    assert type(eda.decoder[0]).__module__ == "eda_b200.encoder_decoder_layers"
    assert type(eda).__module__ == type(ref).__module__ == "models.bdetr"  # the same unmodified file drives both
    g = torch.Generator().manual_seed(5)
    for m in ref.modules():  # non-trivial BatchNorm state so eval-mode folding is exercised
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.6 + 0.8 * torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(1 + 0.1 * torch.randn(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    eda.load_state_dict(ref.state_dict(), strict=True)
    return ref.cuda().eval(), eda.cuda().eval()


def _to_cuda(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


def _hook_decoder(model, store):
    return [layer.register_forward_hook(lambda _m, _i, out, i=i: store.__setitem__(i, out.detach()))
            for i, layer in enumerate(model.decoder)]


def _teacher_force(model, sample_inds):
    """Replaces the top-k of models/bdetr.py:187-205 by the given indices (everything else of the method unchanged)."""
    def forced(xyz, features, end_points):
        logits = model.points_obj_cls(features)
        end_points['seeds_obj_cls_logits'] = logits
        x, f, s = model.gsample_module(xyz, features, sample_inds)
        end_points['query_points_xyz'] = x
        end_points['query_points_feature'] = f
        end_points['query_points_sample_inds'] = s
        return end_points
    model._generate_queries = forced


@pytest.mark.parametrize("B,N,L", [(2, 20000, 40), (8, 50000, 80)], ids=["B2_N20000", "configs2_B8_N50000_L80_K256"])
def test_bdetr_forward_matches_all_reference_model(B, N, L):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref, eda = _models()
    batch = _to_cuda(ref_model.synthetic_batch(B, N, L, seed=3))
    out_ref, out_eda = {}, {}
    hooks = _hook_decoder(ref, out_ref) + _hook_decoder(eda, out_eda)
    report = {}
    with torch.no_grad():
        ep_r = ref(batch)
        ep_free = eda(batch)
        # ---- backbone: index paths exact, features within tolerance ----
        for k in ("sa1_inds", "sa2_inds", "fp2_inds", "seed_inds"):
            assert torch.equal(ep_free[k], ep_r[k]), k
        for k in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "fp2_xyz"):
            assert torch.equal(ep_free[k], ep_r[k]), k
        for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
            mx, rms = _err(ep_free[k], ep_r[k])
            report[k] = (mx, rms)
            assert mx <= TOL_FEAT_MAX and rms <= TOL_FEAT_RMS, (k, mx, rms)
        # ---- text tower: the same torch code in both models ----
        assert torch.equal(ep_free["text_attention_mask"], ep_r["text_attention_mask"])
        torch.testing.assert_close(ep_free["text_feats"], ep_r["text_feats"], rtol=1e-5, atol=1e-5)
        assert ep_r["text_feats"].shape == (B, L, 288)
        # ---- cross-encoder outputs ----
        for k in ("seed_features", "text_memory"):
            mx, rms = _err(ep_free[k], ep_r[k])
            report[k] = (mx, rms)
            assert mx <= TOL_LAYER_MAX and rms <= TOL_LAYER_RMS, (k, mx, rms)
        # ---- KPS query generation: logits close, top-k sets overlap (near-ties at rank 256 may swap) ----
        mx, rms = _err(ep_free["seeds_obj_cls_logits"], ep_r["seeds_obj_cls_logits"])
        report["seeds_obj_cls_logits"] = (mx, rms)
        assert mx <= TOL_LAYER_MAX and rms <= TOL_LAYER_RMS
        overlap = []
        for b in range(B):
            a = set(ep_free["query_points_sample_inds"][b].tolist())
            r = set(ep_r["query_points_sample_inds"][b].tolist())
            overlap.append(len(a & r) / len(r))
        report["topk_overlap_min"] = min(overlap)
        assert min(overlap) >= 0.9, overlap
        assert ep_r["query_points_sample_inds"].shape == (B, 256)
        # ---- teacher-forced: every decoder layer and every head ----
        _teacher_force(eda, ep_r["query_points_sample_inds"])
        out_eda.clear()
        ep_e = eda(batch)
    for h in hooks:
        h.remove()
    assert torch.equal(ep_e["query_points_sample_inds"], ep_r["query_points_sample_inds"])
    assert torch.equal(ep_e["query_points_xyz"], ep_r["query_points_xyz"])
    mx, rms = _err(ep_e["query_points_feature"], ep_r["query_points_feature"])
    assert mx <= TOL_LAYER_MAX and rms <= TOL_LAYER_RMS
    assert len(out_ref) == len(out_eda) == 6
    for i in range(6):
        mx, rms = _err(out_eda[i], out_ref[i])
        report[f"decoder{i}"] = (mx, rms)
        assert mx <= TOL_LAYER_MAX and rms <= TOL_LAYER_RMS, (i, mx, rms)
    for prefix in ["proposal_"] + [f"{i}head_" for i in range(5)] + ["last_"]:
        for k in ("center", "pred_size", "sem_cls_scores", "proj_queries"):
            mx, rms = _err(ep_e[prefix + k], ep_r[prefix + k])
            report[prefix + k] = (mx, rms)
            assert mx <= TOL_LAYER_MAX and rms <= 2 * TOL_LAYER_RMS, (prefix + k, mx, rms)
    worst = sorted(report.items(), key=lambda kv: -(kv[1][1] if isinstance(kv[1], tuple) else 0))[:6]
    print(f"bdetr forward B={B} N={N} L={L}: top-k overlap {report['topk_overlap_min']:.3f}; worst (max, rms): {worst}")
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        import json
        with open(os.path.join(out_dir, f"bdetr_forward_B{B}_N{N}.json"), "w") as f:
            json.dump({k: list(v) if isinstance(v, tuple) else v for k, v in report.items()}, f, indent=1)
