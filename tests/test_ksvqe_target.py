"""The literal `KSVQE` key is NOT on the B200 path yet.  What exists is its pinned target:
golden vectors the REAL reference produced on seeded weights / inputs (tools/make_golden_ksvqe.py).  These checks keep
the fixture honest (structure the reference guarantees) and the product path loud about the gap."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_ksvqe_golden_has_the_reference_structure():
    g = np.load(os.path.join(GOLDEN, "ksvqe_t32_288.npz"))
    assert g["score"].shape == (1, 1) and np.isfinite(g["score"]).all() and np.isfinite(g["loss"])
    assert g["cls_attn"].shape == (4, 49)                    # 4 key frames x 7x7 CLIP patch tokens (cosine to CLS)
    assert np.abs(g["cls_attn"]).max() <= 1.0 + 1e-6
    region = g["region"][0]
    assert region.shape == (32,) and region.min() >= 0 and region.max() <= 8      # one of the 3x3 candidate regions
    # obtain_keyframes (KSVQE_model.py:1352-1376): frames share the region of their key-frame group
    for lo, hi in ((0, 7), (7, 15), (15, 23), (23, 32)):
        assert len(set(region[lo:hi].tolist())) == 1
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys_ksvqe.json")))["KSVQE"]
    assert keys["layers.0.blocks.0.attn.fragment_position_bias_table"][0] == [2535, 3]
    assert any(k.startswith("CLIP_tool.") for k in keys) and any(k.startswith("distortion_tool.") for k in keys)
    assert sum(int(np.prod(v[0])) for v in keys.values()) > 150e6          # ~158 M parameters + buffers


def test_ksvqe_dropin_has_the_reference_state_dict_and_refuses_cpu():
    import torch
    import models
    cfg = {"model": {"type": "KSVQE", "args": {"KSVQE": {
        "backbone": {"num_samples": 1, "sample_type": "topkpertubation", "CLIP_location": 8, "cls_use": True,
                     "tuning_stage": 2, "a1": 1.0, "a2": 1.0, "frozen_stages": -1},
        "head": {"in_channels": 768, "hidden_channels": 64}}}}}
    net = models.VQA_Network(cfg)
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys_ksvqe.json")))["KSVQE"]
    mine = {k: list(v.shape) for k, v in net.KSVQE_backbone.state_dict().items()}
    assert mine == {k: v[0] for k, v in keys.items()}                  # 759 tensors, names and shapes of the reference
    x = {"fragment": torch.zeros(1, 3, 32, 288, 288), "resize_video": torch.zeros(1, 3, 32, 112, 112),
         "dis_label": torch.zeros(1, dtype=torch.long)}
    with pytest.raises(RuntimeError, match="CUDA"):
        net.eval()(inputs=x, reduce_scores=True)
