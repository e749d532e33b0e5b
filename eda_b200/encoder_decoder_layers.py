"""Cross-modal encoder / decoder layers with the class names, constructor and forward signatures
and state-dict keys of the reference's models/encoder_decoder_layers.py (PositionEmbeddingLearned
19-34, CrossAttentionLayer 37-124, TransformerEncoderLayerNoFFN 127-156,
PosTransformerEncoderLayerNoFFN 159-186, BiEncoderLayer 189-255, BiEncoder 259-285,
BiDecoderLayer 288-407), running on the sm_100a kernels of libeda_b200.so.

Submodules are the same torch parameter containers the reference builds (nn.MultiheadAttention,
nn.LayerNorm, nn.Linear inside nn.Sequential, nn.Conv1d / nn.BatchNorm1d) — so initialisation,
`state_dict()` keys and checkpoint loading are identical — but their `forward` is never called:
every attention block is  [q/k/v projection GEMM] -> [QK^T / masked softmax / PV kernel] ->
[out-projection GEMM + residual + LayerNorm epilogue]  (eda_b200/attn_ops.py), every FFN two GEMMs
with the ReLU and the residual + LayerNorm in their epilogues.  Activations stay batch-first
(B, S, E) throughout; the reference's seq-first transposes are views of the same storage and
disappear.
"""
from copy import deepcopy

from torch import nn

from . import attn_ops as ops


def _get_clones(module, N):
    return nn.ModuleList([deepcopy(module) for _ in range(N)])


# The language branch of an encoder layer (text self-attention, text <- vision cross-attention, its FFN: ~8 launches of
# 640-row problems that occupy a handful of SMs) is independent of the vision branch between the two points where the
# modalities exchange keys / values.  It runs on a second CUDA stream, so these latency-bound launches execute in the
# shadow of the vision branch instead of in front of it; autograd replays the same stream assignment in the backward
# pass, and inside a captured CUDA graph the two streams become parallel branches of the graph.  Results are identical
# (same kernels, same order within each branch).  BRANCH_STREAMS = False restores the single-stream order.
BRANCH_STREAMS = True
_branch_streams = {}


def _text_stream(device):
    import torch

    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _branch_streams.get(key)
    if st is None:
        st = _branch_streams[key] = torch.cuda.Stream(device=device)
    return st


class PositionEmbeddingLearned(nn.Module):
    """Absolute pos embedding, learned: Conv1d(C,F,1) -> BatchNorm1d -> ReLU -> Conv1d(F,F,1)."""

    def __init__(self, input_channel, num_pos_feats=288):
        super().__init__()
        self.position_embedding_head = nn.Sequential(
            nn.Conv1d(input_channel, num_pos_feats, kernel_size=1),
            nn.BatchNorm1d(num_pos_feats),
            nn.ReLU(inplace=True),
            nn.Conv1d(num_pos_feats, num_pos_feats, kernel_size=1))

    def forward_rows(self, xyz):
        """xyz (B, N, 3 or 6) -> (B, N, F), batch-first rows (what the attention kernels consume).  Both 1x1 convs are
        tcgen05 GEMMs over the B*N rows; train-mode BatchNorm1d takes its batch statistics (synchronised across ranks
        when the layer is a SyncBatchNorm) and its backward on this package's kernels (eda_b200/rows_mlp.py); in
        inference the running statistics are folded into the first GEMM."""
        from . import rows_mlp

        conv1, bn, _, conv2 = self.position_embedding_head
        layers = [rows_mlp.Layer(conv1.weight, conv1.bias, bn, True, conv1, "w"),
                  rows_mlp.Layer(conv2.weight, conv2.bias, None, False, conv2, "w")]
        B, N, C = xyz.shape
        return rows_mlp.rows_mlp(xyz.reshape(B * N, C), layers).view(B, N, -1)

    def forward(self, xyz):
        """Forward pass, xyz is (B, N, 3or6), output (B, F, N)."""
        return self.forward_rows(xyz).transpose(1, 2)


class CrossAttentionLayer(nn.Module):
    """Cross-attention between language and vision."""

    def __init__(self, d_model=256, dropout=0.1, n_heads=8, dim_feedforward=256, use_butd_enc_attn=False):
        super().__init__()
        self.use_butd_enc_attn = use_butd_enc_attn
        # language <- vision
        self.cross_lv = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout_lv = nn.Dropout(dropout)
        self.norm_lv = nn.LayerNorm(d_model)
        self.ffn_lv = nn.Sequential(
            nn.Linear(d_model, dim_feedforward), nn.ReLU(), nn.Dropout(dropout),
            nn.Linear(dim_feedforward, d_model), nn.Dropout(dropout))
        self.norm_lv2 = nn.LayerNorm(d_model)
        # vision <- language
        self.cross_vl = deepcopy(self.cross_lv)
        self.dropout_vl = nn.Dropout(dropout)
        self.norm_vl = nn.LayerNorm(d_model)
        self.ffn_vl = deepcopy(self.ffn_lv)
        self.norm_vl2 = nn.LayerNorm(d_model)
        if use_butd_enc_attn:
            self.cross_d = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
            self.dropout_d = nn.Dropout(dropout)
            self.norm_d = nn.LayerNorm(d_model)

    def forward(self, vis_feats, vis_key_padding_mask, text_feats, text_key_padding_mask, pos_feats,
                detected_feats=None, detected_mask=None):
        """Forward pass, vis/pos_feats (B, V, F), lang_feats (B, L, F)."""
        # language attends to vision (keys/values without pos), then its FFN
        text_new = ops.mha_block(self.cross_lv, text_feats, vis_feats, vis_feats,
                                 key_padding_mask=vis_key_padding_mask, residual=text_feats, norm=self.norm_lv,
                                 out_dropout=self.dropout_lv)
        text_new = ops.ffn_block(self.ffn_lv, text_new, self.norm_lv2)
        # vision (+pos on the query only) attends to the ORIGINAL language features
        vis_new = ops.mha_block(self.cross_vl, vis_feats, text_feats, text_feats, q_pos=pos_feats,
                                key_padding_mask=text_key_padding_mask, residual=vis_feats, norm=self.norm_vl,
                                out_dropout=self.dropout_vl)
        # vision attends to detected boxes
        if detected_feats is not None and self.use_butd_enc_attn:
            vis_new = ops.mha_block(self.cross_d, vis_new, detected_feats, detected_feats,
                                    key_padding_mask=detected_mask, residual=vis_new, norm=self.norm_d,
                                    out_dropout=self.dropout_d)
        vis_new = ops.ffn_block(self.ffn_vl, vis_new, self.norm_vl2)
        return vis_new, text_new


class TransformerEncoderLayerNoFFN(nn.Module):
    """TransformerEncoderLayer but without FFN (language self-attention)."""

    def __init__(self, d_model, nhead, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)

    def forward_rows(self, src, src_key_padding_mask=None):
        """src (B, S, F) batch-first."""
        return ops.mha_block(self.self_attn, src, src, src, key_padding_mask=src_key_padding_mask, residual=src,
                             norm=self.norm1, out_dropout=self.dropout1)

    def forward(self, src, src_mask=None, src_key_padding_mask=None):
        """src (S, B, F) seq-first like the reference; returns (S, B, F)."""
        if src_mask is not None:
            raise RuntimeError("eda_b200: attn_mask is not used anywhere on the reference path and is not supported")
        return self.forward_rows(src.transpose(0, 1), src_key_padding_mask).transpose(0, 1)


class PosTransformerEncoderLayerNoFFN(TransformerEncoderLayerNoFFN):
    """TransformerEncoderLayerNoFFN that adds pos_embed to query and key (vision self-attention)."""

    def __init__(self, d_model, nhead, dropout):
        super().__init__(d_model, nhead, dropout)

    def forward_rows(self, src, pos, src_key_padding_mask=None):
        return ops.mha_block(self.self_attn, src, src, src, q_pos=pos, k_pos=pos,
                             key_padding_mask=src_key_padding_mask, residual=src, norm=self.norm1,
                             out_dropout=self.dropout1)

    def forward(self, src, pos, src_mask=None, src_key_padding_mask=None):
        """src, pos (S, B, F) seq-first; returns (S, B, F)."""
        if src_mask is not None:
            raise RuntimeError("eda_b200: attn_mask is not used anywhere on the reference path and is not supported")
        return self.forward_rows(src.transpose(0, 1), pos.transpose(0, 1), src_key_padding_mask).transpose(0, 1)


class BiEncoderLayer(nn.Module):
    """Self->cross layer for both modalities."""

    def __init__(self, d_model=256, dropout=0.1, activation="relu", n_heads=8, dim_feedforward=256,
                 self_attend_lang=True, self_attend_vis=True, use_butd_enc_attn=False):
        super().__init__()
        self.self_attention_lang = (TransformerEncoderLayerNoFFN(d_model=d_model, nhead=n_heads, dropout=dropout)
                                    if self_attend_lang else None)
        self.self_attention_visual = (PosTransformerEncoderLayerNoFFN(d_model=d_model, nhead=n_heads, dropout=dropout)
                                      if self_attend_vis else None)
        self.cross_layer = CrossAttentionLayer(d_model, dropout, n_heads, dim_feedforward, use_butd_enc_attn)

    def forward(self, vis_feats, pos_feats, padding_mask, text_feats, text_padding_mask, end_points={},
                detected_feats=None, detected_mask=None):
        """Forward pass, feats (B, N, F), masks (B, N), diff N for V/L."""
        if BRANCH_STREAMS and vis_feats.is_cuda:
            return self._forward_two_streams(vis_feats, pos_feats, padding_mask, text_feats, text_padding_mask,
                                             detected_feats, detected_mask)
        if self.self_attention_visual is not None:
            vis_feats = self.self_attention_visual.forward_rows(vis_feats, pos_feats, padding_mask)
        if self.self_attention_lang is not None:
            text_feats = self.self_attention_lang.forward_rows(text_feats, text_padding_mask)
        return self.cross_layer(vis_feats=vis_feats, vis_key_padding_mask=padding_mask, text_feats=text_feats,
                                text_key_padding_mask=text_padding_mask, pos_feats=pos_feats,
                                detected_feats=detected_feats, detected_mask=detected_mask)

    def _forward_two_streams(self, vis_feats, pos_feats, padding_mask, text_feats, text_padding_mask, detected_feats,
                             detected_mask):
        """Same computation as `forward`, language branch on the side stream (see BRANCH_STREAMS)."""
        import torch

        dev = vis_feats.device
        main, side = torch.cuda.current_stream(dev), _text_stream(dev)
        if torch.is_grad_enabled():
            ops.note_aux_stream(dev, side)  # the backward pass will use it too: joined with the gradient streams
        cl = self.cross_layer
        side.wait_stream(main)  # everything produced so far (inputs of this layer) is visible to the side stream
        with torch.cuda.stream(side):
            if self.self_attention_lang is not None:
                text_sa = self.self_attention_lang.forward_rows(text_feats, text_padding_mask)
            else:
                text_sa = text_feats
        vis_sa = vis_feats
        if self.self_attention_visual is not None:
            vis_sa = self.self_attention_visual.forward_rows(vis_feats, pos_feats, padding_mask)
        # exchange point: each branch needs the other's self-attended features as keys / values
        side.wait_stream(main)
        main.wait_stream(side)
        text_sa.record_stream(main)
        vis_sa.record_stream(side)
        with torch.cuda.stream(side):
            text_new = ops.mha_block(cl.cross_lv, text_sa, vis_sa, vis_sa, key_padding_mask=padding_mask,
                                     residual=text_sa, norm=cl.norm_lv, out_dropout=cl.dropout_lv)
            text_new = ops.ffn_block(cl.ffn_lv, text_new, cl.norm_lv2)
        vis_new = ops.mha_block(cl.cross_vl, vis_sa, text_sa, text_sa, q_pos=pos_feats,
                                key_padding_mask=text_padding_mask, residual=vis_sa, norm=cl.norm_vl,
                                out_dropout=cl.dropout_vl)
        if detected_feats is not None and cl.use_butd_enc_attn:
            vis_new = ops.mha_block(cl.cross_d, vis_new, detected_feats, detected_feats, key_padding_mask=detected_mask,
                                    residual=vis_new, norm=cl.norm_d, out_dropout=cl.dropout_d)
        vis_new = ops.ffn_block(cl.ffn_vl, vis_new, cl.norm_vl2)
        # join: the caller (next layer, decoder, loss) may consume both on the current stream
        main.wait_stream(side)
        text_new.record_stream(main)
        return vis_new, text_new


class BiEncoder(nn.Module):
    """Encode jointly language and vision."""

    def __init__(self, bi_layer, num_layers):
        super().__init__()
        self.layers = _get_clones(bi_layer, num_layers)
        self.num_layers = num_layers

    def forward(self, vis_feats, pos_feats, padding_mask, text_feats, text_padding_mask, end_points={},
                detected_feats=None, detected_mask=None):
        """Forward pass, feats (B, N, F), masks (B, N), diff N for V/L."""
        for i, layer in enumerate(self.layers):
            vis_feats, text_feats = layer(vis_feats, pos_feats, padding_mask, text_feats, text_padding_mask,
                                          end_points, detected_feats=detected_feats, detected_mask=detected_mask)
            if 'lv_attention' in end_points:
                end_points['lv_attention%d' % i] = end_points['lv_attention']
        return vis_feats, text_feats


class BiDecoderLayer(nn.Module):
    """Self->cross_l->cross_d->cross_v layer for proposals."""

    def __init__(self, d_model, n_heads, dim_feedforward=2048, dropout=0.1, activation="relu",
                 self_position_embedding='loc_learned', butd=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.cross_l = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout_l = nn.Dropout(dropout)
        self.norm_l = nn.LayerNorm(d_model)
        if butd:
            self.cross_d = deepcopy(self.cross_l)
            self.dropout_d = nn.Dropout(dropout)
            self.norm_d = nn.LayerNorm(d_model)
        self.cross_v = deepcopy(self.cross_l)
        self.dropout_v = nn.Dropout(dropout)
        self.norm_v = nn.LayerNorm(d_model)
        self.ffn = nn.Sequential(
            nn.Linear(d_model, dim_feedforward), nn.ReLU(), nn.Dropout(dropout),
            nn.Linear(dim_feedforward, d_model), nn.Dropout(dropout))
        self.norm2 = nn.LayerNorm(d_model)
        if self_position_embedding == 'xyz_learned':
            self.self_posembed = PositionEmbeddingLearned(3, d_model)
        elif self_position_embedding == 'loc_learned':
            self.self_posembed = PositionEmbeddingLearned(6, d_model)
        else:
            self.self_posembed = None

    def forward(self, query, vis_feats, lang_feats, query_pos, padding_mask, text_key_padding_mask,
                detected_feats=None, detected_mask=None):
        """query (B,N,F), vis_feats (B,V,F), lang_feats (B,L,F), query_pos (B,N,3or6), padding_mask (B,N)
        for the queries, text_key_padding_mask (B,L) -> query (B,N,F)."""
        if self.self_posembed is not None:
            pos = self.self_posembed.forward_rows(query_pos)
        else:
            pos = None  # the reference adds an all-zero tensor
        query = ops.mha_block(self.self_attn, query, query, query, q_pos=pos, k_pos=pos,
                              key_padding_mask=padding_mask, residual=query, norm=self.norm1, out_dropout=self.dropout1)
        query = ops.mha_block(self.cross_l, query, lang_feats, lang_feats, q_pos=pos,
                              key_padding_mask=text_key_padding_mask, residual=query, norm=self.norm_l,
                              out_dropout=self.dropout_l)
        if detected_feats is not None:
            query = ops.mha_block(self.cross_d, query, detected_feats, detected_feats, q_pos=pos,
                                  key_padding_mask=detected_mask, residual=query, norm=self.norm_d,
                                  out_dropout=self.dropout_d)
        query = ops.mha_block(self.cross_v, query, vis_feats, vis_feats, q_pos=pos, key_padding_mask=None,
                              residual=query, norm=self.norm_v, out_dropout=self.dropout_v)
        query = ops.ffn_block(self.ffn, query, self.norm2)
        return query.contiguous()
