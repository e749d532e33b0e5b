"""Generates tests/golden/attn_*.npz by running the REFERENCE's own models/encoder_decoder_layers.py
(imported by path from /root/reference; pure PyTorch, runs on CPU here) in eval mode on the seeded
parameters / inputs of attn_cases.py.  Run in the build container:

    python tests/golden/make_golden_attention.py

The reference file is only imported, never copied.  fp32, torch CPU math path (no TF32)."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import attn_cases as ac  # noqa: E402

REF = "/root/reference/models/encoder_decoder_layers.py"


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_encoder_decoder_layers", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_ref()
    torch.set_num_threads(os.cpu_count() or 1)
    for name, (kind, B, V, L, D, K) in ac.CASES.items():
        inp = ac.make_inputs(name)
        if kind == "bi_encoder_layer":
            m = ref.BiEncoderLayer(ac.E, dropout=0.1, activation="relu", n_heads=ac.HEADS, dim_feedforward=ac.FF,
                                   self_attend_lang=True, self_attend_vis=True, use_butd_enc_attn=True)
        elif kind == "bi_encoder":
            layer = ref.BiEncoderLayer(ac.E, dropout=0.1, activation="relu", n_heads=ac.HEADS, dim_feedforward=ac.FF,
                                       self_attend_lang=True, self_attend_vis=True, use_butd_enc_attn=True)
            m = ref.BiEncoder(layer, 3)
        else:
            m = ref.BiDecoderLayer(ac.E, ac.HEADS, ac.FF, 0.1, "relu", self_position_embedding="loc_learned", butd=True)
        ac.fill_params(m, seed=100 + len(name)).eval()
        with torch.no_grad():
            if kind in ("bi_encoder_layer", "bi_encoder"):
                vis, text = m(inp["vis"], inp["pos"], None, inp["text"], inp["text_mask"], {},
                              detected_feats=inp["det"], detected_mask=inp["det_mask"])
                out = dict(vis=vis.numpy(), text=text.numpy())
            else:
                q = m(inp["query"], inp["vis"], inp["text"], inp["query_pos"], None, inp["text_mask"],
                      detected_feats=inp["det"], detected_mask=inp["det_mask"])
                out = dict(query=q.numpy())
        keys = np.array(sorted(m.state_dict().keys()))
        shapes = np.array([str(tuple(m.state_dict()[k].shape)) for k in keys])
        np.savez_compressed(os.path.join(HERE, f"attn_{name}.npz"), keys=keys, shapes=shapes, **out)
        print(name, {k: v.shape for k, v in out.items()}, len(keys), "params")
        if kind != "bi_encoder":
            grads = reference_gradients(m, kind, inp)
            np.savez_compressed(os.path.join(HERE, f"attn_grad_{name}.npz"), **grads)
            print(name, "gradients:", len(grads), "arrays")


def grad_weights(shape, seed):
    """The cotangent the gradient fixtures were taken with (tests regenerate it from the same seed)."""
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def reference_gradients(m, kind, inp):
    """Autograd through the REFERENCE module (eval mode: dropout off, BatchNorm1d on running statistics) for
    loss = sum_k <out_k, w_k>.  Stored: gradients of the activation inputs, every 1-D parameter gradient in full, and
    for every matrix-shaped parameter gradient its row sums, column sums and Frobenius norm (full matrices would be
    ~20 MB per case); plus one full matrix per case."""
    for prm in m.parameters():
        prm.grad = None
    leaves = {k: inp[k].clone().requires_grad_(True) for k in (("vis", "text") if kind == "bi_encoder_layer" else ("query",))}
    if kind == "bi_encoder_layer":
        vis, text = m(leaves["vis"], inp["pos"], None, leaves["text"], inp["text_mask"], {}, detected_feats=inp["det"],
                      detected_mask=inp["det_mask"])
        loss = (vis * grad_weights(vis.shape, 1234)).sum() + (text * grad_weights(text.shape, 1235)).sum()
    else:
        q = m(leaves["query"], inp["vis"], inp["text"], inp["query_pos"], None, inp["text_mask"],
              detected_feats=inp["det"], detected_mask=inp["det_mask"])
        loss = (q * grad_weights(q.shape, 1234)).sum()
    loss.backward()
    out = {f"input.{k}": v.grad.numpy() for k, v in leaves.items()}
    full = "cross_layer.cross_vl.in_proj_weight" if kind == "bi_encoder_layer" else "cross_v.in_proj_weight"
    for n, prm in m.named_parameters():
        if prm.grad is None:
            continue
        g = prm.grad.reshape(prm.grad.shape[0], -1) if prm.grad.dim() > 1 else prm.grad
        if g.dim() == 1 or n == full:
            out[f"param.{n}"] = prm.grad.numpy()
        else:
            out[f"rowsum.{n}"] = g.sum(1).numpy()
            out[f"colsum.{n}"] = g.sum(0).numpy()
            out[f"norm.{n}"] = np.array(g.norm().item(), dtype=np.float64)
    return out


if __name__ == "__main__":
    main()
