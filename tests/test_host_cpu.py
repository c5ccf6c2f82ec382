"""CPU-side checks: the C-ABI library loads and exports every symbol include/kvq_b200.h declares (no compute calls
without a GPU), the drop-in module tree has the reference's state_dict, host helpers behave, and the product path
refuses CPU tensors instead of falling back."""
import ctypes
import json
import os
import re

import numpy as np

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
    assert L.kvq_attn_table_len(8, 7, 7) == 2536 + 4336 + 2 * 2366   # compact + conflict-free fast + paired float4 layouts
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


def test_conv_branch_host_entry_points():
    """Planning / validation entry points of the SlowFast and SimpleVQA paths run without a GPU and report errors
    through return codes + kvq_last_error_string (never exit / throw across the C boundary)."""
    from kvq_b200 import lib
    L = lib.load()
    sf = lib.KvqSlowFastConfig()
    for i, d in enumerate((3, 4, 6, 3)):
        sf.depths[i] = d
    sf.alpha = 4
    for i, (a, b) in enumerate(zip((8, 7, 7), (32, 7, 7))):
        sf.slow_pool[i], sf.fast_pool[i] = a, b
    # 3 stems/fusion pairs + per stage 2 pathways x (3 convs per block + branch1) pairs + fusions, plus one twin pair
    # for every fast-pathway convolution that reads 8 / 16 / 32 channels (include/kvq_b200.h)
    n = L.kvq_slowfast_num_weights(ctypes.byref(sf))
    base = 6 + sum(2 * 2 * (3 * d + 1) for d in (3, 4, 6, 3)) + 3 * 2
    twins = (2 * (3 + 4 + 6)          # conv_b and conv_c while inner < 64 (stages 0..2)
             + 2                      # branch1 + conv_a of stage 0 block 0 (8 channels in)
             + 2                      # conv_a of stage 0 blocks 1, 2 (32 channels in)
             + 2                      # branch1 + conv_a of stage 1 block 0 (32 channels in)
             + 2)                     # fusions after the stem (8) and after stage 0 (32)
    assert n == base + 2 * twins
    assert L.kvq_conv_image_kblocks(8, 9) == 2 and L.kvq_conv_image_kblocks(32, 9) == 5 and L.kvq_conv_image_kblocks(16, 3) == 1
    ws = L.kvq_slowfast_workspace_bytes(ctypes.byref(sf), 16, 8, 32, 256, 256)
    assert 1.5e9 < ws < 4e9                                         # a few GB of 180: activations of 16 clips + scratch
    assert L.kvq_slowfast_workspace_bytes(ctypes.byref(sf), 1, 8, 24, 224, 224) == 0
    assert "slow frames" in lib.last_error()                       # 24 fast frames fuse to 6 slow frames, not 8
    assert L.kvq_slowfast_workspace_bytes(ctypes.byref(sf), 1, 8, 32, 16, 16) == 0
    assert L.kvq_slowfast_forward(ctypes.byref(sf), None, n, None, None, 1, 8, 32, 224, 224, None, None, None, 0, None) < 0
    assert "NULL" in lib.last_error()
    rn = lib.KvqResNetConfig()
    for i, d in enumerate((3, 4, 6, 3)):
        rn.layers[i] = d
    rn.feat3d_dim, rn.head = 2304, 1
    assert L.kvq_resnet_feature_dim(ctypes.byref(rn)) == 9472      # simpleVQAHead in_channels (kwai_simpleVQA_test.yml)
    assert L.kvq_resnet_num_weights(ctypes.byref(rn)) == 2 + 2 * (3 * 16 + 4) + 2
    assert L.kvq_simplevqa_workspace_bytes(ctypes.byref(rn), 1, 8, 448, 448) > 0
    assert L.kvq_simplevqa_workspace_bytes(ctypes.byref(rn), 1, 8, 16, 16) == 0
    assert L.kvq_stem_weight_rows(8) == 16 and L.kvq_stem_weight_rows(64) == 64
    # implicit convolution: channel count must be a multiple of 64 (smaller widths go through the gather path)
    rc = L.kvq_conv_implicit_f16(1, 1, None, None, 0, 1, 64, 1, 1, 8, 8, 32, lib.i3((1, 3, 3)), lib.i3((1, 1, 1)),
                                 lib.i3((0, 1, 1)), 64, 0, 1, None)
    assert rc < 0 and "64" in lib.last_error()


def test_pack_conv_weight_layouts():
    """Host-side weight packing contracts of include/kvq_b200.h (tap-major / channel-minor, K padded to 64, row-folded
    twin, stem layout) -- checked on CPU tensors up to the fp16 cast, which needs the device."""
    import torch.nn.functional as F
    w = torch.randn(32, 8, 1, 1, 1)
    g = 8
    wf = torch.zeros(g * 32, 64)
    for j in range(g):
        wf[j * 32:(j + 1) * 32, j * 8:(j + 1) * 8] = w.reshape(32, 8)
    x = torch.randn(40, 8)                                  # 40 rows = 5 folded rows
    ref = x @ w.reshape(32, 8).t()
    got = (x.reshape(5, 64) @ wf.t()).reshape(40, 32)       # same bytes, read as [M/g, g*N]
    assert torch.allclose(ref, got, atol=1e-5)
    # tap-major / channel-minor K index of a (3,1,1) convolution
    wt = torch.randn(4, 16, 3, 1, 1)
    xx = torch.randn(1, 16, 5, 2, 2)
    ref = F.conv3d(xx, wt, padding=(1, 0, 0))
    w2 = wt.permute(0, 2, 3, 4, 1).reshape(4, -1)           # what ops.pack_conv_weight builds
    xp = F.pad(xx, (0, 0, 0, 0, 1, 1))
    cols = torch.stack([xp[:, :, t:t + 3] for t in range(5)], dim=2)          # [1,16,5,3,2,2]
    cols = cols.permute(0, 2, 4, 5, 3, 1).reshape(-1, 48)                       # rows (t,h,w), K = (dt, c)
    got = (cols @ w2.t()).reshape(1, 5, 2, 2, 4).permute(0, 4, 1, 2, 3)
    assert torch.allclose(ref, got, atol=1e-5)


def test_trainer_rescale_matches_reference_formula():
    """trainer.py:356-361."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("kvq_trainer_cpu", os.path.join(PKG, "trainer.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    pr, gt = np.array([0.2, 0.9, 0.4, 0.7]), np.array([1.0, 4.5, 2.0, 3.0])
    z = tr.Trainer.rescale(pr)
    assert abs(z.mean()) < 1e-12 and abs(z.std() - 1.0) < 1e-12
    r = tr.Trainer.rescale(pr, gt)
    assert abs(r.mean() - gt.mean()) < 1e-12 and abs(r.std() - gt.std()) < 1e-12


def test_view_host_entry_points():
    """View pipeline, host side only: the library's tap windows / weights equal the golden-pinned oracle
    bit for bit, the plan sizes the workspace for the crop window only, bad geometry is reported not crashed, and the
    drop-in UnifiedFrameSampler reproduces the reference's indices under the same numpy seed."""
    import numpy as np
    from kvq_b200 import lib
    from oracle import views as O
    L = lib.load()
    for n_in, n_out in [(1920, 112), (1080, 112), (1080, 520), (1920, 520), (64, 112), (7, 7), (5, 1), (1, 3),
                        (608, 224), (270, 224), (90, 70), (70, 90), (3840, 112), (2160, 448), (1030, 3)]:
        xmin, xsize, w = O.aa_weights(n_in, n_out)
        taps = L.kvq_resize_aa_taps(n_in, n_out)
        assert taps == w.shape[1] == O.aa_taps(n_in, n_out)
        a, b, c = np.zeros(n_out, np.int32), np.zeros(n_out, np.int32), np.zeros((n_out, taps), np.float32)
        assert L.kvq_resize_aa_weights(n_in, n_out, a.ctypes.data, b.ctypes.data, c.ctypes.data) == 0
        np.testing.assert_array_equal(a, xmin)
        np.testing.assert_array_equal(b, xsize)
        np.testing.assert_array_equal(c.view(np.uint32), w.view(np.uint32))
    assert L.kvq_resize_aa_taps(0, 4) < 0 and "resize_aa_taps" in lib.last_error()
    # workspace: tables + an L2-sized chunk of the float32 intermediate; a crop window shrinks rows and columns
    full = L.kvq_resize_view_workspace_bytes(1, 1, 1080, 1920, 520, 520, 0, 0, 0, 0)
    crop = L.kvq_resize_view_workspace_bytes(1, 1, 1080, 1920, 520, 520, 36, 36, 448, 448)
    assert 3 * 1080 * 520 * 4 <= full < 3 * 1080 * 520 * 4 + (1 << 20)
    assert crop < full * 0.8
    big = L.kvq_resize_view_workspace_bytes(8, 32, 1080, 1920, 112, 112, 0, 0, 0, 0)
    assert big < (34 << 20)                                        # chunked: never the whole batch's intermediate
    assert L.kvq_resize_view_workspace_bytes(1, 1, 16, 16, 8, 8, 4, 4, 8, 8) == 0 and "crop window" in lib.last_error()
    assert L.kvq_resize_view_workspace_bytes(1, 1, 4, 60000, 8, 8, 0, 0, 0, 0) == 0 and "row tile" in lib.last_error()

    from datasets.views import UnifiedFrameSampler, centre_crop_window
    g = np.load(os.path.join(GOLDEN, "views_frame_sampler.npz"))
    for i, (fsize_t, fragments_t, interval, num_clips, total, seed) in enumerate(g["cases"]):
        np.random.seed(int(seed))
        inds = UnifiedFrameSampler(int(fsize_t), int(fragments_t), int(interval), int(num_clips))(int(total))
        assert inds.dtype == np.int32
        np.testing.assert_array_equal(inds, g[f"inds_{i}"])
    assert centre_crop_window(520, 448) == (36, 448) and centre_crop_window(91, 45) == O.centre_crop_window(91, 45)


def test_view_functions_host_logic_with_oracle_backed_kernel(monkeypatch):
    """datasets/views.py (target sizes incl. arp, crop windows, train-phase draw order, [3,T,H,W] / [B,3,T,H,W] handling,
    the fused normalised variants) checked on CPU: ops.resize_view_u8 is replaced by a stand-in built on the golden-pinned
    oracle, so only the host logic is under test here (the kernel itself is tests/test_gpu_views.py)."""
    import glob
    import numpy as np
    from datasets import views as V
    from kvq_b200 import ops
    from oracle import views as O
    import views_cases

    def fake(frames, out_h, out_w, crop=None, mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, divisor=1.0, layout="BT3HW",
             want_u8=False, want_f32=True, workspace=None, antialias=True):
        x = frames.numpy()
        if layout == "BT3HW":
            x = x.transpose(0, 2, 1, 3, 4)
        r = O.resize_u8(x, out_h, out_w, antialias)
        if crop is not None:
            y, x0, h, w = crop
            r = r[..., y:y + h, x0:x0 + w]
        u8 = torch.from_numpy(np.ascontiguousarray(r)) if want_u8 else None
        f32 = torch.from_numpy(np.stack([O.normalise(c, mean, std, divisor) for c in r])) if want_f32 else None
        return u8, f32

    monkeypatch.setattr(ops, "resize_view_u8", fake)
    views_cases.check_geometry(V, "cpu")
    for path in sorted(glob.glob(os.path.join(GOLDEN, "views_resize*.npz"))):
        g = np.load(path)
        T, H, W = (int(v) for v in g["shape"])
        video = views_cases.golden_video(T, H, W, g["seed"])
        bt3hw = video.permute(1, 0, 2, 3).contiguous().unsqueeze(0)
        if str(g["kind"]) == "resize":
            out = V.get_resized_video(video.contiguous(), size_h=int(g["size_h"]), size_w=int(g["size_w"]))
            norm = V.resized_video_normalised(bt3hw, int(g["size_h"]), int(g["size_w"]))[0]
        else:
            out = V.get_resizecrop_video(video.contiguous(), resize=int(g["resize"]), crop=int(g["crop"]), phase="test")
            norm = V.resizecrop_video_normalised(bt3hw, int(g["resize"]), int(g["crop"]))[0]
        np.testing.assert_array_equal(out.numpy(), g["out"])
        np.testing.assert_array_equal(norm.numpy()[:, :, ::7, ::5], g["norm_sample"])
    with pytest.raises(NotImplementedError):
        V.get_resized_video(torch.zeros((3, 1, 8, 8), dtype=torch.uint8), random_crop=True)
    with pytest.raises(RuntimeError, match="uint8"):
        V.get_resized_video(torch.zeros((1, 8, 8), dtype=torch.uint8))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys, no
    CUDA needed; under torchrun only rank 0 prints.  The view-pipeline workload keeps the run to a few seconds."""
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "views", "--steps", "1",
           "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="0"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["steps"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1"))
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_float_reciprocal_division_is_exact_on_the_ranges_the_kernels_use():
    """fdiv_i (csrc/kvq_kernels.cuh): floor(n / d) from a float32 reciprocal estimate plus one correction step.  The row maps
    call it with quotients far below 2^21 and either n < 2^24 or a divisor >= 256; this emulates the float32 arithmetic
    and checks it against integer division on those ranges (rows up to 2^31 / rows-per-clip, token grids, windows)."""
    rng = np.random.default_rng(3)

    def fdiv(n, d):
        inv = np.float32(1.0) / d.astype(np.float32)
        q = (n.astype(np.float32) * inv).astype(np.int64)          # C++ float -> int conversion truncates
        r = n - q * d
        q = q + (r >= d) - (r < 0)
        return q

    cases = []
    # small divisors (window sizes, lanes) with n < 2^22
    cases.append((rng.integers(0, 1 << 22, 200000), rng.integers(1, 400, 200000)))
    # rows / rows-per-clip and tokens / tokens-per-clip: n up to 2^31 - 1, divisor >= 392
    cases.append((rng.integers(0, (1 << 31) - 1, 200000), rng.integers(392, 1 << 20, 200000)))
    # exact multiples and off-by-one neighbours (where a float estimate is most likely to land on the wrong side)
    d = rng.integers(1, 5000, 100000)
    k = rng.integers(0, 4000, 100000)
    for delta in (-1, 0, 1):
        n = np.maximum(k * d + delta, 0)
        cases.append((n, d))
    for n, d in cases:
        n = n.astype(np.int64)
        d = d.astype(np.int64)
        keep = (n // d) < (1 << 21)
        assert np.array_equal(fdiv(n[keep], d[keep]), n[keep] // d[keep])


@pytest.mark.parametrize("dims,window,shift", [
    ((16, 56, 56), (8, 7, 7), (0, 0, 0)), ((16, 56, 56), (8, 7, 7), (4, 3, 3)), ((16, 14, 14), (8, 7, 7), (4, 3, 3)),
    ((9, 25, 19), (8, 7, 7), (4, 3, 3)), ((4, 24, 32), (8, 7, 7), (4, 3, 3)), ((48, 7, 7), (8, 7, 7), (4, 3, 3)),
    ((8, 16, 16), (4, 4, 4), (2, 2, 2))])
def test_window_row_maps_match_the_oracle_tables(dims, window, shift):
    """kvq_window_row_map runs, on the host, the same inline row-map functions the kernels call (float-reciprocal divisions
    included): roll(-shift) + window_partition against the oracle's token table, the d-fastest order against its
    definition, and src_to_win_row as the inverse on grids without padding."""
    import ctypes as C
    from kvq_b200 import lib
    from oracle import swin3d
    L = lib.load()
    D, H, W = dims
    cw, cs = swin3d.clamp_window(dims, window, shift)
    Dp, Hp, Wp = [-(-n // w) * w for n, w in zip(dims, cw)]
    tabs = swin3d.token_tables((Dp, Hp, Wp), cw, cs)
    src = tabs["src"].numpy()                                   # [nW, N] indices into the PADDED grid
    pd, ph, pw = src // (Hp * Wp), (src // Wp) % Hp, src % Wp
    expect = np.where((pd < D) & (ph < H) & (pw < W), (pd * H + ph) * W + pw, -1).reshape(-1)
    rows = expect.size
    i3 = lambda v: (C.c_int32 * 3)(*v)
    for dfast in (0, 1):
        r2s = np.full(rows, -7, dtype=np.int32)
        s2r = np.full(D * H * W, -7, dtype=np.int32)
        n = L.kvq_window_row_map(D, H, W, i3(window), i3(shift), dfast, r2s.ctypes.data, s2r.ctypes.data)
        assert n == rows
        if dfast == 0:
            assert np.array_equal(r2s, expect)
        else:                                                   # row (h, w, d) of a window <-> row (d, h, w)
            wd, wh, ww = cw
            e = expect.reshape(-1, wd, wh * ww).transpose(0, 2, 1).reshape(-1)
            assert np.array_equal(r2s, e)
        if (Dp, Hp, Wp) == (D, H, W):
            assert np.array_equal(s2r[r2s], np.arange(rows))    # bijection, and src_to_win_row is its inverse
        else:
            assert (s2r == -7).all() and (r2s == -1).any()


def test_gelu_polynomial_of_the_kernels_is_within_its_stated_error():
    """gelu_erf_fast (csrc/kvq_common.cuh): the coefficients in the source are the ones tools/fit_gelu.py derives, and the
    float32 evaluation order of the kernel stays within 1e-6 of the exact erf GELU over [-8, 8] (stated: 9.2e-7)."""
    import importlib.util
    from scipy.special import erf
    spec = importlib.util.spec_from_file_location("fit_gelu", os.path.join(ROOT, "tools", "fit_gelu.py"))
    fg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fg)
    src = open(os.path.join(PKG, "csrc", "kvq_common.cuh")).read()
    body = src[src.index("float gelu_erf_fast(float x)"):]
    body = body[:body.index("}\n")]
    horner = [ln for ln in body.splitlines() if "q = fmaf(" in ln]
    consts = [float(v) for ln in horner for v in re.findall(r"(-?\d\.\d+(?:e[+-]\d+)?)f", ln)]
    # source order: c5, c4, c3, c2, c1 (Horner from the highest coefficient)
    coeffs = consts[::-1]
    assert len(coeffs) == 5
    fitted = [float(np.float32(v)) for v in fg.fit()]
    assert np.allclose(coeffs, fitted, rtol=2e-4, atol=1e-7), (coeffs, fitted)
    x = np.linspace(-8, 8, 200001)
    exact = 0.5 * x * (1 + erf(x / np.sqrt(2)))
    got = fg.gelu_kernel(x, np.array(coeffs)).astype(np.float64)
    assert np.abs(got - exact).max() < 1e-6
