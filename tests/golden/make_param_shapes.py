"""Dump the parameter shape lists of the reference's Slow-R50 + MLP-head encoder
(the tensors K1 walks every step) — run in the authoring container only."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.load_reference()
from models.video_model_builder import ResNet  # noqa: E402

out = {}
for tag, over in {
    "slow_r50_moco_dim128": dict(CONTRASTIVE__TYPE="moco", MODEL__ARCH="slow", MODEL__NUM_CLASSES=128),
    "slow_r50_byol_dim256_pred2": dict(CONTRASTIVE__TYPE="byol", MODEL__ARCH="slow", MODEL__NUM_CLASSES=256,
                                       CONTRASTIVE__DIM=256, CONTRASTIVE__PREDICTOR_DEPTHS=[2]),
}.items():
    cfg = ref_shim.make_cfg(**over)
    net = ResNet(cfg)
    shapes = [[n, list(p.shape)] for n, p in net.named_parameters()]
    out[tag] = {"n_tensors": len(shapes), "n_params": sum(p.numel() for p in net.parameters()), "shapes": shapes}
    print(tag, out[tag]["n_tensors"], out[tag]["n_params"])
json.dump(out, open(os.path.join(HERE, "slow_r50_param_shapes.json"), "w"))
