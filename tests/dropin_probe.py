"""Helper run in a subprocess by tests/test_dropin_bdetr.py (build container only: needs /root/reference).

    python tests/dropin_probe.py reference   # the reference's models/bdetr.py with its own modules
    python tests/dropin_probe.py eda         # the SAME unmodified models/bdetr.py with this repo's modules swapped in
                                             # exactly as INTEGRATION.md sections 2 and 3 describe

Prints one JSON object: state-dict keys -> shapes of BeaUTyDETR, plus the defining module of the swapped classes.
RoBERTa weights / tokenizer files are not available offline, so from_pretrained is replaced by a 1-layer random-init
RobertaModel of the right width and a dummy tokenizer (neither is on the hot path)."""
import importlib
import json
import os
import sys
import types

MODE = sys.argv[1]
REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

for name in ("termcolor", "ipdb"):  # imported by models/__init__.py -> ap_helper / utils, never used here
    if name not in sys.modules:
        try:
            importlib.import_module(name)
        except ImportError:
            stub = types.ModuleType(name)
            stub.colored = lambda s, *a, **k: s
            stub.set_trace = lambda *a, **k: None
            sys.modules[name] = stub

import torch  # noqa: E402
import transformers  # noqa: E402
from transformers import RobertaConfig, RobertaModel  # noqa: E402


def _tiny_roberta(*a, **k):
    torch.manual_seed(0)
    return RobertaModel(RobertaConfig(vocab_size=1000, hidden_size=768, num_hidden_layers=1, num_attention_heads=12,
                                      intermediate_size=256, max_position_embeddings=130, type_vocab_size=1))


transformers.RobertaModel.from_pretrained = classmethod(lambda cls, *a, **k: _tiny_roberta())
transformers.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: object())

sys.path.insert(0, ROOT)
if MODE == "eda":
    # the pointnet2 python package
    import eda_b200.pointnet2 as p2
    for name in ("pointnet2_utils", "pytorch_utils", "pointnet2_modules"):
        mod = importlib.import_module(f"eda_b200.pointnet2.{name}")
        sys.modules[name] = mod
        sys.modules[f"pointnet2.{name}"] = mod
    sys.modules["pointnet2"] = p2
    # the model-side files on the path
    import eda_b200.backbone_module
    import eda_b200.encoder_decoder_layers
    sys.modules["models.backbone_module"] = eda_b200.backbone_module
    sys.modules["models.encoder_decoder_layers"] = eda_b200.encoder_decoder_layers
    # the reference's backbone_module.py also extends sys.path as an import side effect (backbone_module.py:16-21);
    # utils/eval_det.py relies on it (`from metric_util import ...`), so the host does it when that file is replaced
    sys.path.append(os.path.join(REF, "utils"))
else:
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))  # the reference's own compiled pointnet2._ext

sys.path.insert(0, REF)
os.chdir(REF)  # models/bdetr.py loads data/class_embeddings3d.npy by relative path
import models.bdetr as bdetr  # noqa: E402  (unmodified reference file in both modes)

torch.manual_seed(0)
model = bdetr.BeaUTyDETR(num_class=256, num_obj_class=485, input_feature_dim=3, num_queries=256, num_decoder_layers=6,
                         self_position_embedding="loc_learned", contrastive_align_loss=True, d_model=288, butd=True,
                         pointnet_ckpt=None, self_attend=True)
sd = model.state_dict()
print(json.dumps({
    "keys": {k: list(v.shape) for k, v in sd.items()},
    "backbone_module": type(model.backbone_net).__module__,
    "sa_module": type(model.backbone_net.sa1).__module__,
    "encoder_layer_module": type(model.cross_encoder.layers[0]).__module__,
    "decoder_layer_module": type(model.decoder[0]).__module__,
    "bdetr_file": bdetr.__file__,
}))
