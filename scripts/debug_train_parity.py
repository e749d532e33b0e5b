"""Stage-wise train-mode comparison ours vs the reference CUDA build (debugging aid for tests/test_train_step.py)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eda_b200 import hotpath
from oracle import ref_hotpath, ref_loader

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
ours = hotpath.HotPath(dropout=0.0).cuda().train()
ref = ref_hotpath.build(ref_loader.load_reference_ext(), dropout=0.0).cuda().train()
ref.load_state_dict(ours.state_dict(), strict=True)
inputs = [t.cuda() for t in hotpath.synthetic_inputs(8, 50000, 80, 132, 256, seed=100)]
res = {}
def err(a, b):
    d = (a.float() - b.float()).abs()
    return [d.max().item(), d.pow(2).mean().sqrt().item(), b.abs().max().item(), b.float().pow(2).mean().sqrt().item()]
with torch.no_grad():
    ep_o = ours.backbone(inputs[0])
    ep_r = ref.backbone(inputs[0], end_points={})
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        res[k] = err(ep_o[k], ep_r[k])
    for k in ("sa1_inds", "sa2_inds"):
        res[k + "_equal"] = bool(torch.equal(ep_o[k], ep_r[k]))
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    ref2 = ref_hotpath.build(ref_loader.load_reference_ext(), dropout=0.0).cuda().train()
    ref2.load_state_dict(ours.state_dict(), strict=False)
    ep_t = ref2.backbone(inputs[0], end_points={})
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        res["reftf32_" + k] = err(ep_t[k], ep_r[k])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    o = ours(*inputs); r = ref(*inputs)
    for n, a, b in zip(("query", "vis", "text"), o, r):
        res["out_" + n] = err(a, b)
    # encoder alone on identical inputs
    vis_r = ep_r["fp2_features"].transpose(1, 2).contiguous()
    vo, to = ours.encoder(vis_r, inputs[1], None, inputs[2], inputs[3], {}, detected_feats=inputs[4], detected_mask=inputs[5])
    vr, tr = ref.encoder(vis_feats=vis_r, pos_feats=inputs[1], padding_mask=None, text_feats=inputs[2], text_padding_mask=inputs[3], end_points={}, detected_feats=inputs[4], detected_mask=inputs[5])
    res["encoder_only_vis"] = err(vo, vr); res["encoder_only_text"] = err(to, tr)
    # where is the worst vis element?
    d = (o[1] - r[1]).abs()
    idx = torch.nonzero(d == d.max())[0].tolist()
    res["worst_vis_index"] = idx
    res["vis_err_quantiles"] = torch.quantile(d.flatten()[:: 7].float(), torch.tensor([0.5, 0.9, 0.99, 0.999, 0.9999], device="cuda")).tolist()
    dfp = (ep_o["fp2_features"] - ep_r["fp2_features"]).abs()
    res["fp2_err_quantiles"] = torch.quantile(dfp.flatten()[:: 7].float(), torch.tensor([0.5, 0.9, 0.99, 0.999, 0.9999], device="cuda")).tolist()
    res["fp2_worst_channel_errs"] = dfp.amax(dim=(0, 2)).topk(5).values.tolist()
print("columns: max abs err, rms err, max |ref|, rms ref")
print(json.dumps(res, indent=1))
