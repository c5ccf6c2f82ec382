"""oracle/simplevqa.py must reproduce what the REAL reference (simpleVQA_model.resnet50 + simpleVQAHead, run by
tools/make_golden_extra.py) produced on the same seeded weights / inputs."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import simplevqa, synth

CASES = sorted(glob.glob(os.path.join(GOLDEN, "simplevqa_*.npz")))


def load_case(path):
    g = np.load(path)
    shape = tuple(int(v) for v in g["shape"])
    sd = synth.simplevqa_network_state_dict(int(g["wseed"]))
    x = synth.clip_input(shape, int(g["xseed"]))
    feat3d = synth.motion_features((shape[0], shape[2], 2304), int(g["xseed"]) + 1000)
    return g, sd, x, feat3d


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference_golden(path):
    g, sd, x, feat3d = load_case(path)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    feats, score = simplevqa.simplevqa_forward(x, feat3d, sd)
    assert feats.shape == g["feat"].shape and feats.shape[-1] == 9472
    assert float(np.abs(g["feat"][..., :7168]).mean()) > 0.5            # fixture is not degenerate
    np.testing.assert_allclose(feats.numpy(), g["feat"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(score.numpy(), g["score"], rtol=0, atol=2e-4)


def test_there_are_simplevqa_goldens():
    assert len(CASES) >= 4


def test_state_dict_keys_match_reference():
    with open(os.path.join(GOLDEN, "state_dict_keys_simplevqa.json")) as f:
        spec = json.load(f)
    ours = synth.resnet50_shapes()
    ref = {k: v for k, v in spec["ResNet"].items() if "num_batches_tracked" not in k}
    assert set(ours) == set(ref)
    for k, shape in ours.items():
        assert list(shape) == ref[k][0], k
    head = synth.simplevqa_head_shapes()
    assert {k: list(v) for k, v in head.items()} == {k: v[0] for k, v in spec["simpleVQAHead"].items()}


def test_dropin_module_tree_has_reference_keys():
    from models.backbones.simpleVQA_model import resnet50
    from models.head import simpleVQAHead
    with open(os.path.join(GOLDEN, "state_dict_keys_simplevqa.json")) as f:
        spec = json.load(f)
    m, h = resnet50(pretrained=False), simpleVQAHead(in_channels=9472, hidden_channels=128)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == {k: v[0] for k, v in spec["ResNet"].items()}
    assert {k: list(v.shape) for k, v in h.state_dict().items()} == {k: v[0] for k, v in spec["simpleVQAHead"].items()}


def test_simplevqa_refuses_cpu():
    from models.model import VQA_Network
    net = VQA_Network({"model": {"args": {"simpleVQA": {"backbone": None, "head": {"in_channels": 9472,
                                                                                  "hidden_channels": 128}}}}})
    net.eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        net({"simpleVQA": torch.zeros(1, 3, 2, 64, 64), "feat": torch.zeros(1, 2, 2304)})
