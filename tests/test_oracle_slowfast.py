"""CPU checks of the SlowFast restatement (oracle/slowfast.py; PARITY UNPINNED: pytorchvideo is absent offline, the
reference holds no fixture for this path) and of the host-side pieces of the drop-in."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import slowfast as osf
from oracle import synth


def test_mac_count_matches_published_figure():
    """pytorchvideo's model zoo quotes 65.71 GFLOPs (multiply-accumulates) for slowfast_r50 8x8 at 32x256x256."""
    assert abs(osf.count_macs(32, 256, 256) / 1e9 - 65.71) < 0.01
    assert abs(2 * osf.count_macs(32, 224, 224) / 1e9 - 100.62) < 0.01


def test_per_stage_macs_and_parameter_count_match_the_published_model():
    """The strongest pins available offline for the absent pytorchvideo source (parity stays UNPINNED): the per-stage
    multiply-accumulate counts SURVEY.md Appendix B derives from create_slowfast's depth-50 defaults (stem 4.35 G, res2
    8.20 G, res3 12.75 G, res4 25.79 G, res5 14.62 G at 32x256x256) and the model zoo's parameter count for
    slowfast_r50 (34.57 M with the 400-class head; the trunk the reference keeps -- blocks 0..4, SlowFast_features.py:
    137-165 -- is that minus the 2304 x 400 + 400 projection)."""
    st = [m / 1e9 for m in osf.count_macs(32, 256, 256, per_stage=True)]
    for got, want in zip(st, (4.35, 8.20, 12.75, 25.79, 14.62)):
        assert abs(got - want) < 0.006, (st,)
    shapes = synth.slowfast_shapes()
    n_params = sum(int(np.prod(v)) for k, v in shapes.items() if not k.endswith(("running_mean", "running_var")))
    head = 2304 * 400 + 400
    assert abs((n_params + head) / 1e6 - 34.57) < 0.01, (n_params + head) / 1e6
    # hub layout: blocks.{0..4}.multipathway_blocks.{0,1}.*, fusion convs under multipathway_fusion (none after res5)
    fus = sorted({k.split(".multipathway_fusion")[0] for k in shapes if "multipathway_fusion" in k})
    assert len(fus) == 4


def test_slow_frame_indices():
    assert osf.slow_frame_indices(32).tolist() == [0, 4, 8, 13, 17, 22, 26, 31]
    x = torch.arange(2 * 3 * 32 * 4, dtype=torch.float32).view(2, 3, 32, 2, 2)
    slow, fast = osf.pack_pathway_output(x)
    assert fast is x and slow.shape == (2, 3, 8, 2, 2)
    assert torch.equal(slow[:, :, 3], x[:, :, 13])


def test_library_slow_frame_indices_match_torch_linspace():
    from kvq_b200 import lib
    L = lib.load()
    buf = (ctypes.c_int32 * 64)()
    for T in (4, 8, 16, 31, 32, 33, 48, 64, 96, 100, 128, 255):
        n = L.kvq_slow_frame_indices(T, 4, buf, 64)
        assert [buf[i] for i in range(n)] == torch.linspace(0, T - 1, T // 4).long().tolist(), T
    assert L.kvq_slow_frame_indices(3, 4, buf, 64) < 0            # T // 4 == 0: index_select of nothing
    assert L.kvq_slow_frame_indices(1024, 4, buf, 64) < 0


def test_dropin_module_tree_has_the_restated_state_dict():
    import SlowFast_features as sf
    m = sf.slowfast()
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    want = {k: tuple(v) for k, v in synth.slowfast_shapes().items()}
    assert sd == want
    assert isinstance(m.slow_avg_pool.slow_avg_pool, torch.nn.AvgPool3d)
    assert tuple(m.fast_avg_pool.fast_avg_pool.kernel_size) == (32, 7, 7)
    missing = m.load_state_dict(synth.slowfast_state_dict(3), strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.eval()([torch.zeros(1, 3, 8, 224, 224), torch.zeros(1, 3, 32, 224, 224)])


def test_oracle_shapes_and_head_pool_semantics():
    sd = synth.slowfast_state_dict(5)
    x = synth.slowfast_frames((1, 3, 8, 64, 64), 6)
    s, f = osf.trunk(*osf.pack_pathway_output(x), {k: v for k, v in sd.items()})
    assert s.shape == (1, 2048, 2, 2, 2) and f.shape == (1, 256, 8, 2, 2)
    # stride-1 average pooling followed by a global mean is a weighted mean with "windows covering me" weights
    t = torch.randn(1, 4, 3, 8, 8)
    ref = F.adaptive_avg_pool3d(F.avg_pool3d(t, (2, 7, 7), stride=1), 1).flatten()
    cover = lambda n, k: torch.tensor([min(p, n - k) - max(0, p - k + 1) + 1 for p in range(n)], dtype=torch.float32)
    w = cover(3, 2)[:, None, None] * cover(8, 7)[None, :, None] * cover(8, 7)[None, None, :]
    w = w / (2 * 2 * 2 * 2 * 7 * 7)
    assert torch.allclose((t * w).sum(dim=(2, 3, 4)).flatten(), ref, atol=1e-6)


def test_fp16_storage_emulation_is_close():
    sd = synth.slowfast_state_dict(7)
    x = synth.slowfast_frames((1, 3, 32, 224, 224), 8)
    inp = osf.pack_pathway_output(x)
    s, f = osf.slowfast_forward(inp, sd)
    s16, f16 = osf.slowfast_forward_fp16(inp, sd)
    assert s.shape == (1, 2048, 1, 1, 1) and f.shape == (1, 256, 1, 1, 1)
    for a, b in ((s, s16), (f, f16)):
        assert ((a - b).norm() / a.norm()).item() < 2e-3


def test_video_clip_schedule():
    import SlowFast_features as sf
    frames = torch.arange(100, dtype=torch.float32).view(100, 1, 1, 1).expand(100, 3, 2, 2)
    clips = sf.video_clips(frames, 30)
    assert len(clips) == 8 and all(c.shape == (32, 3, 2, 2) for c in clips)       # int(100/30) = 3 clips, padded to 8
    assert clips[1][0, 0, 0, 0] == 30 and clips[2][-1, 0, 0, 0] == 91
    assert torch.equal(clips[3], clips[2]) and torch.equal(clips[7], clips[2])
    clips = sf.video_clips(frames[:70], 30)                                        # tail clip repeats the last frame
    assert clips[1][9, 0, 0, 0] == 39 and clips[1][31, 0, 0, 0] == 61
    c = sf.video_clips(frames[:50], 30)[0]
    assert c[31, 0, 0, 0] == 31


def test_feature_files_round_trip(tmp_path):
    """SlowFast_features.py:196-197 writes what fusion_datasets.py:859-890 reads: [1,2048,1,1,1] / [1,256,1,1,1] per clip
    -> feat [8, 2304] (slow then fast)."""
    import SlowFast_features as sf
    import datasets
    s = torch.randn(8, 2048, 1, 1, 1)
    f = torch.randn(8, 256, 1, 1, 1)
    sf.save_clip_features(str(tmp_path / "vid"), s, f)
    assert np.load(tmp_path / "vid" / "feature_7_fast_feature.npy").shape == (1, 256, 1, 1, 1)
    feat = datasets.load_motion_features(str(tmp_path / "vid"))
    assert feat.shape == (8, 2304)
    assert torch.equal(feat[:, :2048], s.view(8, 2048)) and torch.equal(feat[:, 2048:], f.view(8, 256))
    assert datasets.load_motion_features(str(tmp_path / "vid"), "Fast").shape == (8, 256)
