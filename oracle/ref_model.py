"""Loader for the reference's own PYTHON modules on the path (TEST INFRASTRUCTURE — only tests/ and bench.py's
reference legs import this).

Two "flavours" of the same, unmodified reference files can live side by side in one process:

  load("reference", ext)   models/bdetr.py + models/modules.py + models/backbone_module.py +
                           models/encoder_decoder_layers.py + pointnet2/*.py exactly as the reference ships them, with
                           `pointnet2._ext` bound to `ext`: the reference's compiled CUDA extension (oracle/_ref, the
                           R-GPU baseline) or the C port (oracle/pointnet2_oracle, the CPU baseline — the reference has
                           no CPU implementation of `_ext`, pointnet2/_ext_src/src/sampling.cpp:87)
  load("eda")              the SAME unmodified models/bdetr.py, with this repo's modules swapped in for pointnet2/*,
                           models/backbone_module.py, models/encoder_decoder_layers.py and (heads=True) models/modules.py
                           exactly as INTEGRATION.md sections 2-3 describe (the drop-in); heads=False keeps the
                           reference's own models/modules.py

Files come from /root/reference when it exists (build container), else from the sourceless bytecode
oracle/build_ref.py left in oracle/_ref/pyref (GPU box).  models/__init__.py is deliberately not executed (it pulls in
the evaluation / loss stack, which is outside the path).

The import is done under a scoped sys.modules / sys.path swap: nothing the flavours need stays registered under a
top-level name afterwards, so they cannot see each other (or the product package) by accident.
"""
import importlib
import os
import sys
import tempfile
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
_NAMES = ["pointnet2", "pointnet2._ext", "pointnet2.pointnet2_utils", "pointnet2_utils", "pointnet2_modules",
          "pytorch_utils", "models", "models.bdetr", "models.backbone_module", "models.modules",
          "models.encoder_decoder_layers"]
_cache = {}


_materialised = None


def ref_dir():
    """Directory holding models/ and pointnet2/ of the reference: the checkout itself in the build container, else a
    temporary directory into which the shipped bytecode (oracle/_ref/pyref/*/*.pyc.bin) is copied as importable .pyc."""
    global _materialised
    if os.path.isdir("/root/reference/models"):
        return "/root/reference"
    if _materialised is not None:
        return _materialised
    d = os.path.join(_HERE, "_ref", "pyref")
    if not os.path.isdir(os.path.join(d, "models")):
        return None
    import glob
    import shutil

    tmp = tempfile.mkdtemp(prefix="eda_pyref_")
    n = 0
    for sub in ("models", "pointnet2"):
        os.makedirs(os.path.join(tmp, sub))
        for src in glob.glob(os.path.join(d, sub, "*.pyc.bin")):
            shutil.copyfile(src, os.path.join(tmp, sub, os.path.basename(src)[:-4]))
            n += 1
    if n == 0:
        return None
    _materialised = tmp
    return tmp


def oracle_ext():
    """`pointnet2._ext` look-alike over the C port (CPU tensors)."""
    from . import pointnet2_oracle as orc

    ext = types.ModuleType("pointnet2._ext")
    for name in ("furthest_point_sampling", "gather_points", "gather_points_grad", "ball_query", "group_points",
                 "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad"):
        setattr(ext, name, getattr(orc, name))
    return ext


class Flavour:
    """Attribute bag: bdetr, modules, backbone_module, encoder_decoder_layers, pointnet2_modules, pointnet2_utils."""


def load(flavour, ext=None, heads=True):
    key = (flavour, id(ext), bool(heads))
    if key in _cache:
        return _cache[key]
    rd = ref_dir()
    if rd is None:
        raise RuntimeError("reference python modules unavailable (neither /root/reference nor oracle/_ref/pyref)")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    saved = {n: sys.modules.get(n) for n in _NAMES}
    saved_path = list(sys.path)
    for n in _NAMES:
        sys.modules.pop(n, None)
    try:
        models_pkg = types.ModuleType("models")
        models_pkg.__path__ = [os.path.join(rd, "models")]
        sys.modules["models"] = models_pkg
        if flavour == "reference":
            if ext is None:
                raise RuntimeError("the reference flavour needs an `_ext` module (compiled reference or C port)")
            pkg = types.ModuleType("pointnet2")
            pkg.__path__ = [os.path.join(rd, "pointnet2")]
            pkg._ext = ext
            sys.modules["pointnet2"] = pkg
            sys.modules["pointnet2._ext"] = ext
            sys.path.insert(0, os.path.join(rd, "pointnet2"))  # `import pytorch_utils`, `from pointnet2_modules import`
        elif flavour == "eda":
            import eda_b200.backbone_module
            import eda_b200.encoder_decoder_layers
            import eda_b200.pointnet2 as p2

            sys.modules["pointnet2"] = p2
            for name in ("pointnet2_utils", "pytorch_utils", "pointnet2_modules"):
                mod = importlib.import_module(f"eda_b200.pointnet2.{name}")
                sys.modules[name] = mod
                sys.modules[f"pointnet2.{name}"] = mod
            sys.modules["models.backbone_module"] = eda_b200.backbone_module
            sys.modules["models.encoder_decoder_layers"] = eda_b200.encoder_decoder_layers
            if heads:  # the query-generation / prediction-head modules as well
                import eda_b200.modules
                sys.modules["models.modules"] = eda_b200.modules
        else:
            raise ValueError(flavour)
        bdetr = importlib.import_module("models.bdetr")
        f = Flavour()
        f.name = flavour
        f.bdetr = bdetr
        f.modules = sys.modules["models.modules"]
        f.backbone_module = sys.modules["models.backbone_module"]
        f.encoder_decoder_layers = sys.modules["models.encoder_decoder_layers"]
        f.pointnet2_modules = sys.modules["pointnet2_modules"]
        f.pointnet2_utils = sys.modules.get("pointnet2_utils") or sys.modules["pointnet2.pointnet2_utils"]
        f.ref_dir = rd
    finally:
        for n in _NAMES:
            sys.modules.pop(n, None)
        for n, m in saved.items():
            if m is not None:
                sys.modules[n] = m
        sys.path[:] = saved_path  # the reference's backbone_module.py appends three directories as an import side effect
    _cache[key] = f
    return f


class FakeTokenizer:
    """Stands in for RobertaTokenizerFast (no tokenizer files offline; tokenisation is outside the path): a "sentence"
    is a string of space-separated token ids; pads with 1 (`<pad>`), like `batch_encode_plus(padding="longest")`."""

    def batch_encode_plus(self, texts, padding="longest", return_tensors="pt"):
        import torch
        from transformers import BatchEncoding

        rows = [[int(t) for t in s.split()] for s in texts]
        L = max(len(r) for r in rows)
        ids = torch.ones(len(rows), L, dtype=torch.long)
        mask = torch.zeros(len(rows), L, dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r)
            mask[i, :len(r)] = 1
        return BatchEncoding({"input_ids": ids, "attention_mask": mask})


def synthetic_text(B, L, seed, vocab=1000, min_len=20):
    """ids uniform in [3, vocab), <s> = 0 first, </s> = 2 last, lengths U[min_len, L] with row 0 at L."""
    import torch

    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(B):
        n = L if b == 0 else int(torch.randint(min_len, L + 1, (1,), generator=g))
        ids = [0] + torch.randint(3, vocab, (n - 2,), generator=g).tolist() + [2]
        out.append(" ".join(str(i) for i in ids))
    return out


def build_bdetr(flavour, roberta_layers=2, vocab=1000, seed=0, **kw):
    """BeaUTyDETR from the unmodified models/bdetr.py of the given flavour.  RoBERTa weights / tokenizer files are not
    available offline: `from_pretrained` is replaced, for the duration of the constructor, by a seeded random-init
    RobertaModel of roberta-base width (`roberta_layers` layers; the text tower is frozen and outside the path) and the
    FakeTokenizer; data/class_embeddings3d.npy (copied into an nn.Embedding) by a seeded random array."""
    import numpy as np
    import torch
    import transformers
    from transformers import RobertaConfig, RobertaModel

    def tiny_roberta(*a, **k):
        torch.manual_seed(seed + 17)
        cfg = RobertaConfig(vocab_size=vocab, hidden_size=768, num_hidden_layers=roberta_layers, num_attention_heads=12,
                            intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1)
        return RobertaModel(cfg)

    saved_m = transformers.RobertaModel.__dict__.get("from_pretrained")
    saved_t = transformers.RobertaTokenizerFast.__dict__.get("from_pretrained")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data"))
        rng = np.random.default_rng(seed + 5)
        np.save(os.path.join(tmp, "data", "class_embeddings3d.npy"), rng.standard_normal((485, 768), dtype=np.float32))
        transformers.RobertaModel.from_pretrained = classmethod(lambda cls, *a, **k: tiny_roberta())
        transformers.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: FakeTokenizer())
        try:
            os.chdir(tmp)
            torch.manual_seed(seed)
            args = dict(num_class=256, num_obj_class=485, input_feature_dim=3, num_queries=256, num_decoder_layers=6,
                        self_position_embedding="loc_learned", contrastive_align_loss=True, d_model=288, butd=True,
                        pointnet_ckpt=None, self_attend=True)
            args.update(kw)
            model = flavour.bdetr.BeaUTyDETR(**args)
        finally:
            os.chdir(cwd)
            for cls, old in ((transformers.RobertaModel, saved_m), (transformers.RobertaTokenizerFast, saved_t)):
                if old is not None:
                    cls.from_pretrained = old
                else:
                    del cls.from_pretrained
    return model


def synthetic_batch(B, N, L, seed=0, n_boxes=132):
    """The `inputs` dict BeaUTyDETR.forward reads (models/bdetr.py:208-250), synthetic per SURVEY.md 8d."""
    import torch

    from eda_b200 import synthetic

    g = torch.Generator().manual_seed(seed)
    pc = synthetic.point_clouds(B, N, "surface", seed=synthetic.SEED + seed)
    centres = torch.stack([8 * torch.rand(B, n_boxes, generator=g) - 4, 6 * torch.rand(B, n_boxes, generator=g) - 3,
                           3 * torch.rand(B, n_boxes, generator=g)], -1)
    sizes = 0.2 + 1.8 * torch.rand(B, n_boxes, 3, generator=g)
    nvalid = torch.randint(20, 61, (B,), generator=g)
    mask = torch.arange(n_boxes)[None, :] < nvalid[:, None]
    return {"point_clouds": pc, "text": synthetic_text(B, L, seed + 1),
            "det_boxes": torch.cat([centres, sizes], -1), "det_bbox_label_mask": mask,
            "det_class_ids": torch.randint(0, 485, (B, n_boxes), generator=g)}
