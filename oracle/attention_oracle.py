"""CPU restatement of the reference's cross-modal attention layers (TEST INFRASTRUCTURE).

Plain torch fp32 on CPU tensors, purely functional over a state dict (the reference's own key names),
with nn.MultiheadAttention spelled out as the maths its `need_weights=True` path performs
(third-party: torch.nn.functional.multi_head_attention_forward, torch 2.11 `functional.py:6607-6665`:
in-projection, q * 1/sqrt(head_dim), baddbmm with the -inf key-padding mask, softmax, bmm,
out-projection).  Follows, with file:line of the reference:

  position_embedding   models/encoder_decoder_layers.py:19-34   (Conv1d -> BatchNorm1d(eval) -> ReLU -> Conv1d)
  cross_attention      :75-124   (CrossAttentionLayer.forward)
  self_attention       :127-186  (TransformerEncoderLayerNoFFN / PosTransformerEncoderLayerNoFFN)
  bi_encoder_layer     :225-255  (BiEncoderLayer.forward)
  bi_encoder           :268-285
  bi_decoder_layer     :341-407  (BiDecoderLayer.forward)

Eval-mode semantics (dropout = identity, BatchNorm1d running statistics).  Pinned against the
reference's own module code imported from /root/reference by tests/golden/make_golden_attention.py
-> tests/golden/attn_*.npz.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this file.
"""
import math

import torch
import torch.nn.functional as F


def mha(sd, prefix, q_in, k_in, v_in, key_padding_mask=None, n_heads=8):
    """q_in (B,Nq,E), k_in, v_in (B,Nk,E) batch-first; mask (B,Nk) bool, True = ignore.  Returns (B,Nq,E)."""
    E = q_in.size(-1)
    D = E // n_heads
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    B, Nq, _ = q_in.shape
    Nk = k_in.size(1)
    q = F.linear(q_in, w[:E], b[:E]).view(B, Nq, n_heads, D).transpose(1, 2)
    k = F.linear(k_in, w[E:2 * E], b[E:2 * E]).view(B, Nk, n_heads, D).transpose(1, 2)
    v = F.linear(v_in, w[2 * E:], b[2 * E:]).view(B, Nk, n_heads, D).transpose(1, 2)
    s = (q * math.sqrt(1.0 / float(D))) @ k.transpose(-1, -2)
    if key_padding_mask is not None:
        s = s + torch.zeros(B, 1, 1, Nk, device=s.device).masked_fill(key_padding_mask.view(B, 1, 1, Nk), float("-inf"))
    ctx = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, Nq, E)
    return F.linear(ctx, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def layer_norm(sd, prefix, x, eps=1e-5):
    return F.layer_norm(x, (x.size(-1),), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def ffn(sd, prefix, x):
    """Sequential(Linear, ReLU, Dropout, Linear, Dropout) in eval mode (indices 0 and 3 carry parameters)."""
    h = F.relu(F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    return F.linear(h, sd[prefix + "3.weight"], sd[prefix + "3.bias"])


def position_embedding(sd, prefix, xyz, eps=1e-5):
    """xyz (B,N,C) -> (B,N,F) (the reference returns the (B,F,N) transpose of this)."""
    p = prefix + "position_embedding_head."
    h = F.linear(xyz, sd[p + "0.weight"].squeeze(-1), sd[p + "0.bias"])
    h = (h - sd[p + "1.running_mean"]) / torch.sqrt(sd[p + "1.running_var"] + eps) * sd[p + "1.weight"] + sd[p + "1.bias"]
    return F.linear(F.relu(h), sd[p + "3.weight"].squeeze(-1), sd[p + "3.bias"])


def cross_attention(sd, prefix, vis, vis_mask, text, text_mask, pos, detected=None, detected_mask=None, n_heads=8):
    """encoder_decoder_layers.py:75-124."""
    qv = vis + pos
    text2 = mha(sd, prefix + "cross_lv.", text, vis, vis, vis_mask, n_heads)
    t = layer_norm(sd, prefix + "norm_lv.", text + text2)
    t = layer_norm(sd, prefix + "norm_lv2.", t + ffn(sd, prefix + "ffn_lv.", t))
    vis2 = mha(sd, prefix + "cross_vl.", qv, text, text, text_mask, n_heads)
    v = layer_norm(sd, prefix + "norm_vl.", vis + vis2)
    if detected is not None and (prefix + "cross_d.in_proj_weight") in sd:
        vis2 = mha(sd, prefix + "cross_d.", v, detected, detected, detected_mask, n_heads)
        v = layer_norm(sd, prefix + "norm_d.", v + vis2)
    v = layer_norm(sd, prefix + "norm_vl2.", v + ffn(sd, prefix + "ffn_vl.", v))
    return v, t


def self_attention(sd, prefix, src, pos=None, mask=None, n_heads=8):
    """:149-156 (pos None) and :179-186 (pos added to query and key)."""
    qk = src if pos is None else src + pos
    src2 = mha(sd, prefix + "self_attn.", qk, qk, src, mask, n_heads)
    return layer_norm(sd, prefix + "norm1.", src + src2)


def bi_encoder_layer(sd, prefix, vis, pos, vis_mask, text, text_mask, detected=None, detected_mask=None, n_heads=8):
    """:225-255."""
    if (prefix + "self_attention_visual.self_attn.in_proj_weight") in sd:
        vis = self_attention(sd, prefix + "self_attention_visual.", vis, pos, vis_mask, n_heads)
    if (prefix + "self_attention_lang.self_attn.in_proj_weight") in sd:
        text = self_attention(sd, prefix + "self_attention_lang.", text, None, text_mask, n_heads)
    return cross_attention(sd, prefix + "cross_layer.", vis, vis_mask, text, text_mask, pos, detected, detected_mask,
                           n_heads)


def bi_encoder(sd, prefix, num_layers, vis, pos, vis_mask, text, text_mask, detected=None, detected_mask=None,
               n_heads=8):
    """:268-285."""
    for i in range(num_layers):
        vis, text = bi_encoder_layer(sd, f"{prefix}layers.{i}.", vis, pos, vis_mask, text, text_mask, detected,
                                     detected_mask, n_heads)
    return vis, text


def bi_decoder_layer(sd, prefix, query, vis, lang, query_pos, padding_mask, text_mask, detected=None,
                     detected_mask=None, n_heads=8):
    """:341-407."""
    if (prefix + "self_posembed.position_embedding_head.0.weight") in sd:
        pos = position_embedding(sd, prefix + "self_posembed.", query_pos)
    else:
        pos = torch.zeros_like(query)
    q2 = mha(sd, prefix + "self_attn.", query + pos, query + pos, query, padding_mask, n_heads)
    query = layer_norm(sd, prefix + "norm1.", query + q2)
    q2 = mha(sd, prefix + "cross_l.", query + pos, lang, lang, text_mask, n_heads)
    query = layer_norm(sd, prefix + "norm_l.", query + q2)
    if detected is not None:
        q2 = mha(sd, prefix + "cross_d.", query + pos, detected, detected, detected_mask, n_heads)
        query = layer_norm(sd, prefix + "norm_d.", query + q2)
    q2 = mha(sd, prefix + "cross_v.", query + pos, vis, vis, None, n_heads)
    query = layer_norm(sd, prefix + "norm_v.", query + q2)
    return layer_norm(sd, prefix + "norm2.", query + ffn(sd, prefix + "ffn.", query))
