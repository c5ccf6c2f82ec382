"""GPU parity of the SlowFast-R50 motion-feature path through the drop-in `SlowFast_features.slowfast` ->
libkvq_b200.so, against the CPU restatement oracle/slowfast.py (PARITY UNPINNED with respect to pytorchvideo: see
the oracle header).

Tolerance: the path stores weights / activations in fp16 with fp32 accumulation.  Pooled features are O(1..20);
oracle.slowfast_forward_fp16 reproduces the storage rounding on the CPU.  The kernel must be within rel-L2 3e-3 of the
fp32 oracle and within rel-L2 1e-3 / max-abs 2e-2 of the fp16 emulation (only the summation order differs)."""
import os

import numpy as np
import pytest
import torch

from oracle import slowfast as osf
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(seed):
    import SlowFast_features as sf
    m = sf.slowfast()
    sd = synth.slowfast_state_dict(seed)
    m.load_state_dict(sd, strict=False)
    return m.to(DEV).eval(), sd


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("B,T,H,W", [(1, 32, 224, 224), (2, 32, 224, 224), (1, 32, 256, 256), (1, 32, 224, 288)])
def test_slowfast_matches_oracle(B, T, H, W):
    """224^2 = SlowFast_features.py's default --resize; 256^2 = BASELINE config 3 (8x8 res5 map: overlapping pools);
    B = 2 makes the fast-stem patch matrix 1.2 GB, i.e. it exercises the L2-sized im2col chunking."""
    import SlowFast_features as sf
    m, sd = _model(11)
    x = synth.slowfast_frames((B, 3, T, H, W), 12)
    inputs = sf.pack_pathway_output(x, DEV)
    ref_in = osf.pack_pathway_output(x)
    assert torch.equal(inputs[0].cpu(), ref_in[0]) and torch.equal(inputs[1].cpu(), ref_in[1])
    s, f = m(inputs)
    assert s.shape == (B, 2048, 1, 1, 1) and f.shape == (B, 256, 1, 1, 1)
    rs, rf = osf.slowfast_forward(ref_in, sd)
    es, ef = osf.slowfast_forward_fp16(ref_in, sd)
    s, f = s.cpu(), f.cpu()
    assert _rel(s, rs) < 3e-3 and _rel(f, rf) < 3e-3, (_rel(s, rs), _rel(f, rf))
    assert _rel(s, es) < 1e-3 and _rel(f, ef) < 1e-3, (_rel(s, es), _rel(f, ef))
    assert (s - es).abs().max().item() < 2e-2 and (f - ef).abs().max().item() < 2e-2


def test_clips_are_independent_and_graph_replay_is_bit_stable():
    """Size-independent property: batching clips must not change any clip's features (no cross-clip op in eval)."""
    import SlowFast_features as sf
    m, _ = _model(13)
    x = synth.slowfast_frames((3, 3, 32, 224, 224), 14)
    inputs = sf.pack_pathway_output(x, DEV)
    s, f = m(inputs)
    for i in range(3):
        si, fi = m([inputs[0][i:i + 1].contiguous(), inputs[1][i:i + 1].contiguous()])
        assert (si - s[i:i + 1]).abs().max().item() < 1e-3 and (fi - f[i:i + 1]).abs().max().item() < 1e-3
    m.use_cuda_graph = True
    g1 = [t.clone() for t in m(inputs)]
    g2 = [t.clone() for t in m(inputs)]
    assert torch.equal(g1[0], g2[0]) and torch.equal(g1[1], g2[1])
    assert torch.equal(g1[0], s) and torch.equal(g1[1], f)


def test_bad_geometry_fails_loudly(tmp_path):
    import SlowFast_features as sf
    m, _ = _model(15)
    slow = torch.zeros(1, 3, 8, 224, 224, device=DEV)
    with pytest.raises(RuntimeError, match="slow frames"):
        m([slow, torch.zeros(1, 3, 24, 224, 224, device=DEV)])            # 24 fast frames fuse to 6 != 8
    with pytest.raises(RuntimeError, match="avg pool kernel"):
        m([torch.zeros(1, 3, 8, 160, 160, device=DEV), torch.zeros(1, 3, 32, 160, 160, device=DEV)])  # 5x5 map < 7x7
    # on-disk format consumed by datasets/fusion_datasets.py:883-890
    s, f = m([slow, torch.zeros(1, 3, 32, 224, 224, device=DEV)])
    sf.save_clip_features(str(tmp_path / "vid"), s, f, first_index=3)
    a = np.load(tmp_path / "vid" / "feature_3_slow_feature.npy")
    b = np.load(tmp_path / "vid" / "feature_3_fast_feature.npy")
    assert a.shape == (1, 2048, 1, 1, 1) and b.shape == (1, 256, 1, 1, 1) and a.dtype == np.float32
