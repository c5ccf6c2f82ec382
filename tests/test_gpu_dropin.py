"""GPU tests of the drop-in API surface: models.VQA_Network (reference constructor / forward contract), stand-alone
backbone + head modules, and the test.py / Trainer.inferece flow with the fragment kernel in front."""
import glob
import importlib.util
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, PKG

pytestmark = pytest.mark.gpu
CFG = {"model": {"type": "technical", "args": {"swin_tiny_grpb": {"head": {"in_channels": 768, "hidden_channels": 64}}}}}


def _golden_model(path, dev):
    import models
    from oracle import synth
    g = np.load(path)
    wseed = int(g["wseed"])
    sd = {"swin_tiny_grpb_backbone." + k: v for k, v in synth.synth_state_dict(synth.swin_shapes(), wseed).items()}
    sd.update({"swin_tiny_grpb_head." + k: v for k, v in synth.synth_state_dict(synth.vqa_head_shapes(), wseed).items()})
    m = models.VQA_Network(CFG)
    # reference checkpoints are saved from the DataParallel wrapper: keys carry `module.`
    msg = m.load_state_dict({k: v for k, v in sd.items()}, strict=False)
    assert not msg.unexpected_keys and all("relative_position_index" in k for k in msg.missing_keys)
    m = m.to(dev)
    m.eval()
    return g, m


def test_vqa_network_matches_reference_golden():
    from oracle import synth
    dev = torch.device("cuda:0")
    g, m = _golden_model(os.path.join(GOLDEN, "swin_t32_224.npz"), dev)
    x = synth.clip_input(tuple(int(v) for v in g["shape"]), int(g["xseed"])).to(dev)
    with torch.no_grad():
        s = m(inputs={"technical": x}, reduce_scores=True)
        s2, feats = m(inputs={"technical": x}, reduce_scores=False, return_pooled_feats=True)
    assert s.shape == (2, 1) and isinstance(s2, list) and torch.equal(s2[0], s)
    assert np.abs(s.cpu().numpy().reshape(-1) - g["score"].reshape(-1)).max() <= 1e-3
    f = feats["swin_tiny_grpb"]
    assert list(f.shape) == [int(v) for v in g["feat_shape"]]
    assert np.abs(f[:, ::16].cpu().numpy() - g["feat"]).max() < 5e-2

    # unfused module calls (backbone.forward then head.forward) agree with the fused call
    with torch.no_grad():
        feat = m.swin_tiny_grpb_backbone({"technical": x})
        s3 = m.swin_tiny_grpb_head(feat)
    assert torch.equal(feat, f)
    assert (s3 - s).abs().max().item() < 5e-4       # stand-alone head: fp16 fc_hid weights; fused call: hi/lo pair

    # changing a parameter invalidates the packed weights
    with torch.no_grad():
        m.swin_tiny_grpb_head.fc_last.bias += 1.0
        s4 = m(inputs={"technical": x}, reduce_scores=True)
    assert (s4 - s - 1.0).abs().max().item() < 1e-5


def test_trainer_inferece_writes_reference_format(tmp_path):
    """test.py flow: YAML-shaped config -> Trainer -> inferece() -> `video_name,score` lines; scores are checked
    against the CPU oracle fed with the reference-style fragment sampling of the same synthetic frames."""
    from oracle import swin3d, synth
    spec = importlib.util.spec_from_file_location("kvq_trainer", os.path.join(PKG, "trainer.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    cfg = {"num_workers": 0, "load_path": None, "model": CFG["model"],
           "data": {"val": {"type": "SyntheticFragmentDataset",
                            "args": {"num_videos": 2, "src_h": 100, "src_w": 90, "seed": 9,
                                     "sample_types": {"technical": {"fragments_h": 2, "fragments_w": 2, "fsize_h": 32,
                                                                    "fsize_w": 32, "aligned": 8, "clip_len": 16,
                                                                    "num_clips": 2}}}}}}
    sd = synth.swin_network_state_dict(21)
    ckpt = tmp_path / "ckpt.pth"
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, ckpt)     # DP-style checkpoint
    cfg["load_path"] = str(ckpt)
    t = tr.Trainer(types.SimpleNamespace(gpu_id="0"), cfg)
    out = tmp_path / "output.txt"
    res = t.inferece(str(out))
    lines = out.read_text().strip().split("\n")
    assert [l.split(",")[0] for l in lines] == ["synthetic_0000", "synthetic_0001"]

    mean = torch.tensor([123.675, 116.28, 103.53]).view(3, 1, 1, 1)
    std = torch.tensor([58.395, 57.12, 57.375]).view(3, 1, 1, 1)
    for i, line in enumerate(lines):
        item = t.val_dataset[i]
        video = item["frames"].permute(1, 0, 2, 3).float()           # [C,T,H,W] as in get_spatial_fragments
        T = video.shape[1]
        tgt = torch.zeros(3, T, 64, 64)
        hg = [min(100 // 2 * a, 100 - 32) for a in range(2)]
        wg = [min(90 // 2 * a, 90 - 32) for a in range(2)]
        for a in range(2):
            for b in range(2):
                for c in range(T // 8):
                    ho, wo = hg[a] + int(item["offsets"][0, a, b, c]), wg[b] + int(item["offsets"][1, a, b, c])
                    tgt[:, c * 8:(c + 1) * 8, a * 32:(a + 1) * 32, b * 32:(b + 1) * 32] = \
                        video[:, c * 8:(c + 1) * 8, ho:ho + 32, wo:wo + 32]
        tgt = (tgt - mean) / std
        clips = tr.split_clips(tgt[None], 2)
        ref = swin3d.vqa_network_swin(sd, clips).mean().item()
        got = float(line.split(",")[1])
        assert abs(got - ref) <= 1e-3, (i, got, ref)
        assert abs(res[i][1] - got) < 1e-6
    # labelled validation flow (trainer.py:251-296): correlations of the rescaled predictions with the labels
    cfg["data"]["val"]["args"]["num_videos"] = 4
    t4 = tr.Trainer(types.SimpleNamespace(gpu_id="0"), cfg)
    m = t4.inferece_val()
    assert set(m) == {"SRCC", "PLCC", "KRCC", "RMSE"} and all(v == v for v in m.values())      # finite numbers
    assert -1.0 <= m["SRCC"] <= 1.0 and -1.0 <= m["PLCC"] <= 1.0 and m["RMSE"] >= 0.0


def test_cuda_graph_replay_matches_eager():
    from oracle import synth
    dev = torch.device("cuda:0")
    g, m = _golden_model(os.path.join(GOLDEN, "swin_t16_64x64.npz"), dev)
    xa = synth.clip_input((3, 3, 16, 64, 64), 41).to(dev)
    xb = synth.clip_input((3, 3, 16, 64, 64), 42).to(dev)
    with torch.no_grad():
        m.use_cuda_graph = False
        ea, eb = m(inputs={"technical": xa}, reduce_scores=True).clone(), m(inputs={"technical": xb}, reduce_scores=True).clone()
        m.use_cuda_graph = True
        ga = m(inputs={"technical": xa}, reduce_scores=True).clone()       # capture + replay
        gb = m(inputs={"technical": xb}, reduce_scores=True).clone()       # second buffer -> second graph
        xa.copy_(xb)                                                       # same buffer, new contents -> replay
        ga2 = m(inputs={"technical": xa}, reduce_scores=True).clone()
    assert torch.equal(ga, ea) and torch.equal(gb, eb) and torch.equal(ga2, eb)
