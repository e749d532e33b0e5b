"""The hot path assembled as one module: Pointnet2Backbone -> 3 x BiEncoderLayer -> 6 x BiDecoderLayer, i.e. the part
of BeaUTyDETR.forward (models/bdetr.py:208-339) that SURVEY.md section 8 puts in scope, with the tensors the
out-of-scope parts would provide (text features from the frozen RoBERTa tower, box features of the detector, query
proposals and their positional boxes) as inputs.  Used by the training-step measurements (bench.py `fwd_bwd`,
benchmarks/micro_train.py) with BASELINE.json configs[3] shapes: B scenes x N = 50 000 points, L = 80 text tokens,
D = 132 boxes, K = 256 queries, d_model 288, 8 heads, FFN 256.
"""
import torch
import torch.nn as nn

from . import encoder_decoder_layers as edl
from . import rows_mlp
from .backbone_module import Pointnet2Backbone

D_MODEL, HEADS, FFN = 288, 8, 256


class HotPath(nn.Module):
    def __init__(self, dropout=0.0, n_enc=3, n_dec=6):
        super().__init__()
        self.backbone = Pointnet2Backbone(input_feature_dim=3, width=1)
        self.encoder = edl.BiEncoder(edl.BiEncoderLayer(D_MODEL, dropout, "relu", HEADS, FFN, True, True, True), n_enc)
        self.decoder = nn.ModuleList(
            [edl.BiDecoderLayer(D_MODEL, HEADS, FFN, dropout, "relu", "loc_learned", True) for _ in range(n_dec)])

    def forward(self, pc, pos, text, text_mask, det, det_mask, query, qpos):
        ep = self.backbone(pc)
        # (B, 288, 1024) -> batch-first rows (B, 1024, 288) on the package's transpose kernel (models/bdetr.py:259 does
        # `.transpose(1, 2).contiguous()`, an ATen copy)
        vis = rows_mlp.transpose_last2(ep["fp2_features"]) if ep["fp2_features"].is_cuda \
            else ep["fp2_features"].transpose(1, 2).contiguous()
        v, t = self.encoder(vis, pos, None, text, text_mask, {}, detected_feats=det, detected_mask=det_mask)
        q = query
        for d in self.decoder:
            q = d(q, v, t, qpos, None, text_mask, detected_feats=det, detected_mask=det_mask)
        return q, v, t


def quadratic_loss(out):
    """Synthetic scalar objective (the reference's loss / matcher is out of scope): every output
    of the path receives a dense gradient."""
    q, v, t = out
    return q.pow(2).mean() + 0.1 * v.pow(2).mean() + 0.1 * t.pow(2).mean()


def ragged_mask(B, n, lo, g):
    """(B,n) bool, True = padded; row 0 keeps everything, the others keep a prefix of random length >= lo."""
    keep = torch.randint(lo, n + 1, (B,), generator=g)
    keep[0] = n
    return torch.arange(n)[None, :] >= keep[:, None]


def synthetic_inputs(B, N=50000, L=80, D=132, K=256, seed=100, vis_tokens=1024):
    """CPU tensors (pc, pos, text, text_mask, det, det_mask, query, qpos) of the configs[3] shapes."""
    from . import synthetic

    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    pc = synthetic.point_clouds(B, N, "surface", seed=synthetic.SEED + seed)
    text, det, query, pos = r(B, L, D_MODEL), r(B, D, D_MODEL), r(B, K, D_MODEL), 0.5 * r(B, vis_tokens, D_MODEL)
    text_mask = ragged_mask(B, L, 20, g)
    det_mask = ragged_mask(B, D, 20, g)
    qpos = torch.cat([4 * torch.rand(B, K, 3, generator=g) - 2, torch.rand(B, K, 3, generator=g) + .2], -1)
    return pc, pos, text, text_mask, det, det_mask, query, qpos
