"""Attention-stack micro-benchmark on one B200 (BASELINE.json configs[2] shapes: B=8, V=1024 seeds, L=80 tokens,
D=132 boxes, K=256 queries, d=288, 8 heads): 3 x BiEncoderLayer + 6 x BiDecoderLayer forward, eval mode.

  ours      eda_b200.encoder_decoder_layers (tcgen05 linear + attention kernels)
  torch     the same maths as plain torch ops on the GPU (oracle/attention_oracle.py run on CUDA tensors:
            F.linear x3, matmul, softmax, matmul, F.linear, layer_norm — op for op what nn.MultiheadAttention's
            need_weights=True path launches for the reference), fp32 (allow_tf32 False, the reference's setting)
            and with TF32 allowed
Also times the individual kernels and prints achieved TFLOP/s (tensor work only).  JSON on stdout.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import attn_cases as ac  # noqa: E402

from eda_b200 import attn_ops as ops  # noqa: E402
from eda_b200 import encoder_decoder_layers as edl  # noqa: E402
from oracle import attention_oracle as ao  # noqa: E402  (benchmark baseline leg only)


def timeit(fn, warmup=5, iters=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return {"min_ms": ts[0], "med_ms": ts[len(ts) // 2]}


def main():
    B, V, L, D, K = 8, 1024, 80, 132, 256
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    vis, pos, text, det, query = r(B, V, ac.E), 0.5 * r(B, V, ac.E), r(B, L, ac.E), r(B, D, ac.E), r(B, K, ac.E)
    text_mask = ac.ragged_mask(B, L, 20, g).to(dev)
    det_mask = ac.ragged_mask(B, D, 20, g).to(dev)
    qpos = torch.cat([4 * torch.rand(B, K, 3, generator=g) - 2, torch.rand(B, K, 3, generator=g) + .2], -1).to(dev)

    enc = edl.BiEncoder(edl.BiEncoderLayer(ac.E, 0.1, "relu", ac.HEADS, ac.FF, True, True, True), 3)
    decs = [edl.BiDecoderLayer(ac.E, ac.HEADS, ac.FF, 0.1, "relu", "loc_learned", True) for _ in range(6)]
    ac.fill_params(enc, 1)
    for i, d in enumerate(decs):
        ac.fill_params(d, 10 + i)
    enc = enc.to(dev).eval()
    decs = [d.to(dev).eval() for d in decs]
    esd = enc.state_dict()
    dsds = [d.state_dict() for d in decs]

    def ours():
        with torch.no_grad():
            v, t = enc(vis, pos, None, text, text_mask, {}, detected_feats=det, detected_mask=det_mask)
            q = query
            for d in decs:
                q = d(q, v, t, qpos, None, text_mask, detected_feats=det, detected_mask=det_mask)
        return q

    def torch_path():
        with torch.no_grad():
            v, t = ao.bi_encoder(esd, "", 3, vis, pos, None, text, text_mask, det, det_mask)
            q = query
            for sd in dsds:
                q = ao.bi_decoder_layer(sd, "", q, v, t, qpos, None, text_mask, det, det_mask)
        return q

    res = {"B": B, "V": V, "L": L, "D": D, "K": K, "gpu": torch.cuda.get_device_name(0)}
    lib_launch0 = None
    from eda_b200 import _lib
    lib = _lib.load()
    ours()
    torch.cuda.synchronize()
    n0 = lib.eda_launch_count()
    ours()
    res["ours_launches_per_forward"] = int(lib.eda_launch_count() - n0)
    res["ours_enc3_dec6"] = timeit(ours)
    # the same forward recorded once and replayed as a CUDA graph (eda_b200/graphs.py): no host launch latency
    from eda_b200.graphs import GraphedCallable

    def ours_fn(vis_, pos_, text_, det_, query_, qpos_):
        v, t = enc(vis_, pos_, None, text_, text_mask, {}, detected_feats=det_, detected_mask=det_mask)
        q = query_
        for d in decs:
            q = d(q, v, t, qpos_, None, text_mask, detected_feats=det_, detected_mask=det_mask)
        return q

    try:
        graphed = GraphedCallable(ours_fn, [vis, pos, text, det, query, qpos])
        res["ours_graphed_equals_eager"] = bool(torch.equal(graphed(vis, pos, text, det, query, qpos), ours()))
        res["ours_enc3_dec6_graphed"] = timeit(lambda: graphed(vis, pos, text, det, query, qpos))
    except Exception as e:  # noqa: BLE001
        res["ours_enc3_dec6_graphed"] = repr(e)
    torch.backends.cuda.matmul.allow_tf32 = False
    a, b = ours(), torch_path()
    res["max_abs_diff_vs_torch_fp32"] = (a - b).abs().max().item()
    res["rms_diff_vs_torch_fp32"] = (a - b).pow(2).mean().sqrt().item()
    res["torch_fp32_enc3_dec6"] = timeit(torch_path)
    torch.backends.cuda.matmul.allow_tf32 = True
    res["torch_tf32_enc3_dec6"] = timeit(torch_path)
    torch.backends.cuda.matmul.allow_tf32 = False

    # individual kernels
    H, Dh = ac.HEADS, ac.E // ac.HEADS
    for name, Nq, Nk in (("attn_vis_self", V, V), ("attn_cross_v", K, V), ("attn_dec_self", K, K), ("attn_text_vis", L, V)):
        q, k, vt = r(B * Nq, ac.E), r(B * Nk, ac.E), r(B, ac.E, Nk)
        t = timeit(lambda: ops.attention_raw(q, k, vt, None, B, Nq, Nk, H))
        t["tflops"] = 4.0 * B * H * Nq * Nk * Dh / (t["min_ms"] * 1e-3) / 1e12
        res[name] = t
    W = r(ac.E, ac.E) / 17
    pw = ops.pack_weight(W)
    bias = r(ac.E)
    for name, R, nprob, ln in (("linear_8192x288x288", B * V, 1, False), ("linear_qkv_3x8192", B * V, 3, False),
                               ("linear_ln_8192", B * V, 1, True), ("linear_2048x288x288", B * K, 1, False)):
        x = r(R, ac.E)
        res_t = r(R, ac.E)
        probs = [dict(x=x, w_packed=pw, bias=bias, residual=res_t if ln else None) for _ in range(nprob)]
        t = timeit(lambda: ops.linear_raw(probs, ac.E, ac.E, ln=(bias, bias, 1e-5) if ln else None))
        t["tflops"] = 2.0 * nprob * R * ac.E * ac.E / (t["min_ms"] * 1e-3) / 1e12
        res[name] = t
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
