"""Generates tests/golden/state_dict_contract.json from the UNMODIFIED reference
`ContrastiveModel` (models/contrastive.py:37-129): the names, shapes and dtypes of every state_dict
entry in each mode, with the stub backbone of ref_shim.  This is the checkpoint contract
(utils/misc.py:118-137,300-339) the drop-in module must keep.
Run:  python tests/golden/make_golden_statedict.py"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

rc = ref_shim.load_reference()
ref_shim.register_stub_backbone(rc)

CASES = {
    "moco": dict(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64),
    "moco_knn": dict(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64, CONTRASTIVE__KNN_ON=True,
                     CONTRASTIVE__LENGTH=10),
    "byol": dict(CONTRASTIVE__TYPE="byol", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64, CONTRASTIVE__PREDICTOR_DEPTHS=[1]),
    "swav": dict(CONTRASTIVE__TYPE="swav", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64),
    "swav_queue": dict(CONTRASTIVE__TYPE="swav", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64, CONTRASTIVE__SWAV_QEUE_LEN=32),
    "simclr": dict(CONTRASTIVE__TYPE="simclr", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64),
    "mem_1d": dict(CONTRASTIVE__TYPE="mem", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64, CONTRASTIVE__LENGTH=20,
                   CONTRASTIVE__MEM_TYPE="1d"),
    "mem_2d": dict(CONTRASTIVE__TYPE="mem", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64, CONTRASTIVE__LENGTH=20,
                   CONTRASTIVE__MEM_TYPE="2d"),
}


def main():
    out = {}
    for name, over in CASES.items():
        cfg = ref_shim.make_cfg(**over)
        torch.manual_seed(0)
        m = rc.ContrastiveModel(cfg)
        out[name] = {"cfg": {k: v for k, v in over.items()},
                     "state_dict": sorted([k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()),
                     "frozen": sorted(n for n, p in m.named_parameters() if not p.requires_grad)}
    with open(os.path.join(HERE, "state_dict_contract.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: len(v["state_dict"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
