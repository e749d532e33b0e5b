import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import attn_cases as ac
from eda_b200 import encoder_decoder_layers as edl, attn_ops as ops
from oracle import attention_oracle as ao

def stat(name, got, want):
    e = (got.cpu() - want).abs()
    rows = (e.view(-1, e.size(-1)).max(1).values > 1e-2).nonzero().flatten().tolist()
    print(f"{name}: max {e.max():.3e} rms {e.pow(2).mean().sqrt():.3e} ref_rms {want.pow(2).mean().sqrt():.3e} badrows {rows[:20]} n={len(rows)}")

m = edl.BiDecoderLayer(ac.E, ac.HEADS, ac.FF, 0.1, "relu", self_position_embedding="loc_learned", butd=True)
ac.fill_params(m, seed=100 + len("dec_layer")).eval()
sd = m.state_dict()
inp = ac.make_inputs("dec_layer")
ci = {k: v.cuda() for k, v in inp.items()}
mc = m.cuda()
with torch.no_grad():
    pos_o = ao.position_embedding(sd, "self_posembed.", inp["query_pos"])
    pos_c = mc.self_posembed.forward_rows(ci["query_pos"])
    stat("posembed", pos_c, pos_o)
    q = inp["query"]
    o1 = ao.layer_norm(sd, "norm1.", q + ao.mha(sd, "self_attn.", q + pos_o, q + pos_o, q))
    c1 = ops.mha_block(mc.self_attn, ci["query"], ci["query"], ci["query"], q_pos=pos_o.cuda(), k_pos=pos_o.cuda(), residual=ci["query"], norm=mc.norm1)
    stat("self_attn", c1, o1)
    o2 = ao.layer_norm(sd, "norm_l.", o1 + ao.mha(sd, "cross_l.", o1 + pos_o, inp["text"], inp["text"], inp["text_mask"]))
    c2 = ops.mha_block(mc.cross_l, o1.cuda(), ci["text"], ci["text"], q_pos=pos_o.cuda(), key_padding_mask=ci["text_mask"], residual=o1.cuda(), norm=mc.norm_l)
    stat("cross_l", c2, o2)
    o3 = ao.layer_norm(sd, "norm_d.", o2 + ao.mha(sd, "cross_d.", o2 + pos_o, inp["det"], inp["det"], inp["det_mask"]))
    c3 = ops.mha_block(mc.cross_d, o2.cuda(), ci["det"], ci["det"], q_pos=pos_o.cuda(), key_padding_mask=ci["det_mask"], residual=o2.cuda(), norm=mc.norm_d)
    stat("cross_d", c3, o3)
    o4 = ao.layer_norm(sd, "norm_v.", o3 + ao.mha(sd, "cross_v.", o3 + pos_o, inp["vis"], inp["vis"]))
    c4 = ops.mha_block(mc.cross_v, o3.cuda(), ci["vis"], ci["vis"], q_pos=pos_o.cuda(), residual=o3.cuda(), norm=mc.norm_v)
    stat("cross_v", c4, o4)
    o5 = ao.layer_norm(sd, "norm2.", o4 + ao.ffn(sd, "ffn.", o4))
    c5 = ops.ffn_block(mc.ffn, o4.cuda(), mc.norm2)
    stat("ffn", c5, o5)
    full = mc(ci["query"], ci["vis"], ci["text"], ci["query_pos"], None, ci["text_mask"], detected_feats=ci["det"], detected_mask=ci["det_mask"])
    stat("full", full, o5)
    for i in range(3):
        full2 = mc(ci["query"], ci["vis"], ci["text"], ci["query_pos"], None, ci["text_mask"], detected_feats=ci["det"], detected_mask=ci["det_mask"])
        print("rerun equal:", torch.equal(full, full2), (full - full2).abs().max().item())
