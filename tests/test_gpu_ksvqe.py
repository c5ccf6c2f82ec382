"""The literal `KSVQE` key (config/Kwai_KSVQE_test.yml: CLIP ViT-B/16 + adapters, QRS, CONTRIQUE, CDM, Swin3D-GRPB,
VQAHead) through the drop-in VQA_Network -> libkvq_b200.so, against golden vectors the REAL reference produced on the
same seeded weights / inputs (tools/make_golden_ksvqe.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from tools import synth

pytestmark = pytest.mark.gpu

CFG = {"model": {"type": "KSVQE", "args": {"KSVQE": {
    "backbone": {"num_samples": 1, "sample_type": "topkpertubation", "CLIP_location": 8, "cls_use": True,
                 "tuning_stage": 2, "a1": 1.0, "a2": 1.0, "frozen_stages": -1},
    "head": {"in_channels": 768, "hidden_channels": 64}}}}}


def _network(g):
    import models
    net = models.VQA_Network(CFG)
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys_ksvqe.json")))["KSVQE"]
    wseed = int(g["wseed"])
    sd = {}
    for k, v in keys.items():
        if not v[1].startswith("float"):
            continue
        t = synth.fill_like("KSVQE_backbone." + k, tuple(v[0]), wseed)
        if k in ("a1", "a2"):
            t = torch.ones(tuple(v[0]))
        sd["KSVQE_backbone." + k] = t
    for k, v in net.KSVQE_head.state_dict().items():
        sd["KSVQE_head." + k] = synth.fill_like("KSVQE_head." + k, tuple(v.shape), wseed)
    missing = net.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert all("num_batches_tracked" in k or "relative_position_index" in k for k in missing.missing_keys), missing.missing_keys
    return net.to("cuda:0").eval()


@pytest.mark.parametrize("name", ["ksvqe_t32_288", "ksvqe_b2_t32_288", "ksvqe_t96_288"])
def test_ksvqe_network_matches_the_reference_golden(name):
    """One 32-frame clip, a batch of two clips with different distortion labels, and one 96-frame item (what the
    reference's inferece_test feeds for this key)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, T = (int(g["B"]), int(g["T"])) if "B" in g else (1, 32)
    labels = g["labels"] if "labels" in g else np.zeros(1)
    net = _network(g)
    gen = torch.Generator().manual_seed(int(g["xseed"]))
    x = {"fragment": torch.randn((B, 3, T, 288, 288), generator=gen).cuda(),
         "resize_video": torch.randn((B, 3, T, 112, 112), generator=gen).cuda(),
         "dis_label": torch.from_numpy(np.asarray(labels)).long().cuda()}
    with torch.no_grad():
        (score, feats, loss) = net(inputs=x, reduce_scores=True, return_pooled_feats=True)
    feat = feats["KSVQE"].cpu()
    assert feat.shape == (B, 768, T // 2, 7, 7) and score.shape == (B, 1)
    ref_stats = g["feat_stats"]
    mine = np.array([feat.mean().item(), feat.abs().mean().item(), feat.abs().max().item(), feat.std().item()])
    ferr = (feat[0, :8, 0] - torch.from_numpy(g["feat_slice"])).abs().max().item()
    serr = np.abs(score.cpu().numpy().reshape(-1) - g["score"].reshape(-1)).max()
    print("ksvqe", name, "score", score.cpu().reshape(-1).tolist(), "ref", g["score"].reshape(-1).tolist(), "feat slice err",
          ferr, "stats", mine, ref_stats, "loss", float(loss), float(g["loss"]))
    assert serr < 1e-3, serr
    assert ferr < 3e-2 and np.abs(mine - ref_stats)[[0, 1, 3]].max() < 3e-3
    assert abs(float(loss) - float(g["loss"])) < 1e-2
    # the (scores, loss) contract of the KSVQE key (model.py:118-121, trainer.py:323-325)
    with torch.no_grad():
        out = net(inputs=x, reduce_scores=True)
    assert isinstance(out, tuple) and len(out) == 2 and torch.equal(out[0], score)


def test_trainer_flow_with_the_ksvqe_yaml(tmp_path):
    """config/Kwai_KSVQE_test.yml -> Trainer -> inferece(): three 32-frame clips per video stay concatenated (96 frames,
    one KSVQE clip) exactly as the reference's loop leaves them; output.txt has the reference's `name,score` lines."""
    import importlib.util
    import types
    import yaml
    from conftest import PKG
    spec = importlib.util.spec_from_file_location("kvq_trainer_ks", os.path.join(PKG, "trainer.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    with open(os.path.join(PKG, "config", "Kwai_KSVQE_test.yml")) as f:
        cfg = yaml.safe_load(f)
    cfg["data"]["val"]["args"]["num_videos"] = 2
    t = tr.Trainer(types.SimpleNamespace(gpu_id="0"), cfg)
    assert t.key_list == ["KSVQE"]
    res = t.inferece(str(tmp_path / "output.txt"))
    lines = (tmp_path / "output.txt").read_text().strip().split("\n")
    assert [l.split(",")[0] for l in lines] == ["synthetic_0000", "synthetic_0001"]
    assert all(np.isfinite(s) for _, s in res) and abs(res[0][1] - res[1][1]) > 0


def test_ksvqe_whole_step_cuda_graph_matches_eager():
    """The whole KSVQE step captured in one CUDA graph (the modulation is enqueued from the stage hook DURING capture)
    replays to the same scores / features as eager launches, also after the input buffers were refilled in place."""
    g = np.load(os.path.join(GOLDEN, "ksvqe_b2_t32_288.npz"))
    B, T = int(g["B"]), int(g["T"])
    net = _network(g)
    gen = torch.Generator().manual_seed(int(g["xseed"]))
    x = {"fragment": torch.randn((B, 3, T, 288, 288), generator=gen).cuda(),
         "resize_video": torch.randn((B, 3, T, 112, 112), generator=gen).cuda(),
         "dis_label": torch.from_numpy(np.asarray(g["labels"])).long().cuda()}
    with torch.no_grad():
        net.use_cuda_graph = False
        s0, f0, l0 = net(inputs=x, reduce_scores=True, return_pooled_feats=True)
        s0, f0, l0 = s0.clone(), f0["KSVQE"].clone(), l0.clone()
        net.use_cuda_graph = True
        s1, f1, l1 = net(inputs=x, reduce_scores=True, return_pooled_feats=True)      # capture + first replay
        torch.cuda.synchronize()
        assert torch.equal(s1, s0) and torch.equal(f1["KSVQE"], f0) and torch.allclose(l1, l0)
        # new clip contents in the SAME buffers: the replay must see them
        x["fragment"].copy_(torch.randn((B, 3, T, 288, 288), generator=gen))
        x["resize_video"].copy_(torch.randn((B, 3, T, 112, 112), generator=gen))
        s2, f2, _ = net(inputs=x, reduce_scores=True, return_pooled_feats=True)
        s2, f2 = s2.clone(), f2["KSVQE"].clone()
        net.use_cuda_graph = False
        s3, f3, _ = net(inputs=x, reduce_scores=True, return_pooled_feats=True)
        torch.cuda.synchronize()
    assert not torch.equal(s2, s0)
    assert torch.equal(s2, s3) and torch.equal(f2, f3["KSVQE"])
