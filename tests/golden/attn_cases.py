"""Shared by make_golden_attention.py (runs the REFERENCE layers, in the build container) and
tests/test_attention.py (runs the oracle / the CUDA layers): deterministic parameters and inputs.
Parameters are filled in sorted state-dict key order from a seeded generator, independent of module
construction order, so the reference module and this repo's mirror get identical weights iff their
state-dict keys and shapes are identical (which the tests also assert)."""
import torch

E, HEADS, FF = 288, 8, 256

CASES = {
    # name: (kind, B, V, L, Dbox, K)
    "enc_layer": ("bi_encoder_layer", 2, 136, 19, 13, 0),
    "encoder3": ("bi_encoder", 2, 136, 19, 13, 0),
    "dec_layer": ("bi_decoder_layer", 2, 136, 19, 13, 70),
}


def fill_params(module, seed):
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    new = {}
    for key in sorted(sd.keys()):
        t = sd[key]
        if key.endswith("num_batches_tracked"):
            new[key] = torch.tensor(3, dtype=t.dtype)
        elif key.endswith("running_var"):
            new[key] = torch.rand(t.shape, generator=g) * 0.5 + 0.75
        elif key.endswith("running_mean"):
            new[key] = torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() == 1 and ("norm" in key or ".1.weight" in key) and key.endswith("weight"):
            new[key] = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif t.dim() == 1:
            new[key] = 0.05 * torch.randn(t.shape, generator=g)
        else:
            fan_in = t.shape[1] if t.dim() > 1 else t.shape[0]
            new[key] = torch.randn(t.shape, generator=g) / (fan_in ** 0.5)
    module.load_state_dict(new)
    return module


def ragged_mask(B, n, lo, g):
    """(B,n) bool, True = padded; row 0 keeps everything, others keep a prefix of random length >= lo."""
    keep = torch.randint(lo, n + 1, (B,), generator=g)
    keep[0] = n
    return torch.arange(n)[None, :] >= keep[:, None]


def make_inputs(name, seed=7):
    kind, B, V, L, D, K = CASES[name]
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    inp = dict(vis=r(B, V, E), pos=0.5 * r(B, V, E), text=r(B, L, E), text_mask=ragged_mask(B, L, 5, g),
               det=r(B, D, E), det_mask=ragged_mask(B, D, 3, g))
    if K:
        inp["query"] = r(B, K, E)
        inp["query_pos"] = torch.cat([4 * torch.rand(B, K, 3, generator=g) - 2, torch.rand(B, K, 3, generator=g) + 0.2], -1)
    return inp
