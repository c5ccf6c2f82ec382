"""CPU-side checks: the C-ABI library loads and exports every symbol include/kvq_b200.h declares (no compute calls
without a GPU), the drop-in module tree has the reference's state_dict, host helpers behave, and the product path
refuses CPU tensors instead of falling back."""
import ctypes
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, PKG, ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "kvq_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kvq_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from kvq_b200 import lib
    assert os.path.exists(lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/kvq_b200.h but not exported"
    assert set(lib.PROTOTYPES) == set(names), set(lib.PROTOTYPES) ^ set(names)
    assert lib.load().kvq_build_info().decode().startswith("kvq_b200")


def test_host_only_entry_points():
    from kvq_b200 import lib
    L = lib.load()
    assert L.kvq_attn_table_len(8, 7, 7) == 2536 + 4336          # compact + bank-conflict-free fast layout
    assert L.kvq_attn_table_len(4, 4, 4) == 344
    cfg = lib.KvqSwinConfig()
    cfg.embed_dim, cfg.num_stages, cfg.head_hidden, cfg.ln_eps = 96, 4, 64, 1e-5
    for i, (d, h, f) in enumerate(zip((2, 2, 6, 2), (3, 6, 12, 24), (1, 1, 1, 0))):
        cfg.depths[i], cfg.num_heads[i], cfg.frag_bias[i] = d, h, f
    for i, w in enumerate((8, 7, 7)):
        cfg.window[i] = w
    assert L.kvq_swin3d_num_weights(ctypes.byref(cfg)) == 4 + 13 * 12 + 3 * 3 + 2 + 4
    ws1 = L.kvq_swin3d_workspace_bytes(ctypes.byref(cfg), 1, 32, 224, 224)
    ws8 = L.kvq_swin3d_workspace_bytes(ctypes.byref(cfg), 8, 32, 224, 224)
    assert 80e6 < ws1 < 160e6 and 7.5 * ws1 < ws8 < 8.5 * ws1
    assert L.kvq_window_rows(2, 16, 56, 56, lib.i3((8, 7, 7)), lib.i3((4, 3, 3))) == 2 * 128 * 392
    assert L.kvq_window_rows(1, 4, 10, 9, lib.i3((8, 7, 7)), lib.i3((0, 0, 0))) == 4 * 196   # clamped + padded
    # error reporting instead of crashing: bad config
    cfg.num_heads[0] = 5
    assert L.kvq_swin3d_workspace_bytes(ctypes.byref(cfg), 1, 32, 224, 224) == 0
    assert "heads" in lib.last_error()


def test_dropin_state_dict_matches_reference():
    import models
    m = models.VQA_Network({"model": {"args": {"swin_tiny_grpb": {"head": {"in_channels": 768, "hidden_channels": 64}}}}})
    spec = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    ref = {"swin_tiny_grpb_backbone." + k: v for k, v in spec["SwinTransformer3D"].items()}
    ref.update({"swin_tiny_grpb_head." + k: v for k, v in spec["VQAHead"].items()})
    sd = m.state_dict()
    assert set(sd) == set(ref)
    for k, (shape, dtype) in ref.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == "torch." + dtype, k
    # the derived buffer equals the reference's relative_position_index closed form (Appendix A)
    rpi = sd["swin_tiny_grpb_backbone.layers.0.blocks.0.attn.relative_position_index"]
    i, j = 391, 5
    di, hi, wi, dj, hj, wj = i // 49, (i // 7) % 7, i % 7, j // 49, (j // 7) % 7, j % 7
    assert int(rpi[i, j]) == (di - dj + 7) * 169 + (hi - hj + 6) * 13 + (wi - wj + 6)


def test_no_cpu_fallback():
    import models
    m = models.VQA_Network({"model": {"args": {"swin_tiny_grpb": {"head": {"in_channels": 768, "hidden_channels": 64}}}}})
    m.eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(inputs={"technical": torch.zeros(1, 3, 16, 64, 64)}, reduce_scores=True)


def test_trainer_helpers():
    import importlib
    spec = importlib.util.spec_from_file_location("kvq_trainer", os.path.join(PKG, "trainer.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    x = torch.arange(2 * 3 * 6 * 2 * 2, dtype=torch.float32).reshape(2, 3, 6, 2, 2)
    y = tr.split_clips(x, 3)
    assert y.shape == (6, 3, 2, 2, 2)
    assert torch.equal(y[1], x[0, :, 2:4]) and torch.equal(y[5], x[1, :, 4:6])       # clip c of video b = frames [c*L,(c+1)*L)
    assert tr.strip_module_prefix({"module.a.b": 1, "c": 2}) == {"a.b": 1, "c": 2}

    calls = []

    def fake_model(inputs, reduce_scores):
        calls.append(inputs["technical"].shape)
        return inputs["technical"].mean(dim=(1, 2, 3, 4)).reshape(-1, 1)

    data = {"technical": x[:1].clone(), "num_clips": {"technical": torch.tensor([3])}, "video_name": ["v"]}
    s = tr.score_video(fake_model, data, ["technical"])
    assert calls == [torch.Size([3, 3, 2, 2, 2])]
    assert abs(float(s) - float(x[:1].mean())) < 1e-5


def test_shard_bounds_cover_everything():
    from kvq_b200 import parallel
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
