"""The hot path (eda_b200/hotpath.py: Pointnet2Backbone -> 3 x BiEncoderLayer -> 6 x BiDecoderLayer) assembled from the
REFERENCE's own modules (TEST INFRASTRUCTURE: bench.py's reference legs and tests/ only).

Same constructor arguments as models/bdetr.py:60-135 uses (backbone width 1, d_model 288, 8 heads, FFN 256, butd,
self-attention in the encoder, `loc_learned` position embedding), same parameter names as eda_b200.hotpath.HotPath, so
state dicts are interchangeable.  With `_ext` = the reference's compiled CUDA extension this is the R-GPU baseline
(BASELINE.md section 2); with `_ext` = the C port it is the CPU baseline ("the reference's CPU-only PyTorch path":
reference Python modules unchanged + a CPU restatement of the nine native ops, which the reference does not have).
"""
import torch.nn as nn

from . import ref_model

D_MODEL, HEADS, FFN = 288, 8, 256


def build(ext, dropout=0.0, n_enc=3, n_dec=6):
    ref = ref_model.load("reference", ext)
    edl = ref.encoder_decoder_layers

    class RefHotPath(nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = ref.backbone_module.Pointnet2Backbone(input_feature_dim=3, width=1)
            self.encoder = edl.BiEncoder(edl.BiEncoderLayer(D_MODEL, dropout=dropout, activation="relu", n_heads=HEADS,
                                                            dim_feedforward=FFN, self_attend_lang=True,
                                                            self_attend_vis=True, use_butd_enc_attn=True), n_enc)
            self.decoder = nn.ModuleList(
                [edl.BiDecoderLayer(D_MODEL, n_heads=HEADS, dim_feedforward=FFN, dropout=dropout, activation="relu",
                                    self_position_embedding="loc_learned", butd=True) for _ in range(n_dec)])
            for m in self.modules():  # models/bdetr.py:341-345
                if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                    m.momentum = 0.1

        def forward(self, pc, pos, text, text_mask, det, det_mask, query, qpos):
            ep = self.backbone(pc, end_points={})
            vis = ep["fp2_features"].transpose(1, 2).contiguous()
            v, t = self.encoder(vis_feats=vis, pos_feats=pos, padding_mask=None, text_feats=text,
                                text_padding_mask=text_mask, end_points={}, detected_feats=det, detected_mask=det_mask)
            q = query
            for d in self.decoder:
                q = d(q, v, t, qpos, None, text_mask, detected_feats=det, detected_mask=det_mask)
            return q, v, t

    return RefHotPath()
