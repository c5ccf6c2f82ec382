"""View pipeline through the C ABI (kvq_resize_view_u8) against the golden vectors the REAL reference
produced (tools/make_golden_views.py: get_resized_video / get_resizecrop_video + the datasets' normalisation lines) and
against oracle/views.py on seeded inputs.  The bar is BIT-EXACT: every uint8 of the resized view, every float32 of the
normalised view."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CASES = sorted(glob.glob(os.path.join(GOLDEN, "views_resize*.npz")))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden_frames(g):
    T, H, W = (int(v) for v in g["shape"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    return torch.randint(0, 256, (T, H, W, 3), generator=gen, dtype=torch.uint8)      # decoder order [T,H,W,3]


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_views_match_reference_golden(path):
    from datasets import views as V
    from kvq_b200 import ops
    g = np.load(path)
    thwc = _golden_frames(g)
    video = thwc.permute(3, 0, 1, 2).contiguous().cuda()                              # the reference's [3,T,H,W]
    if str(g["kind"]) == "resize":
        out = V.get_resized_video(video, size_h=int(g["size_h"]), size_w=int(g["size_w"]))
        bt3hw = thwc.permute(0, 3, 1, 2).contiguous().unsqueeze(0).cuda()             # [1,T,3,H,W]
        norm = V.resized_video_normalised(bt3hw, int(g["size_h"]), int(g["size_w"]))[0]
    else:
        out = V.get_resizecrop_video(video, resize=int(g["resize"]), crop=int(g["crop"]), phase="test")
        bt3hw = thwc.permute(0, 3, 1, 2).contiguous().unsqueeze(0).cuda()
        norm = V.resizecrop_video_normalised(bt3hw, int(g["resize"]), int(g["crop"]))[0]
    out, norm = out.cpu().numpy(), norm.cpu().numpy()
    assert out.dtype == np.uint8 and norm.dtype == np.float32
    np.testing.assert_array_equal(out, g["out"])
    assert _sha(out) == str(g["out_sha256"])
    np.testing.assert_array_equal(norm[:, :, ::7, ::5], g["norm_sample"])
    assert _sha(norm) == str(g["norm_sha256"])                                        # every float32
    # both outputs of one call agree with the two separate calls
    lo_n = V.centre_crop_window(int(g["resize"]), int(g["crop"])) if str(g["kind"]) != "resize" else None
    u8, f32 = ops.resize_view_u8(video.unsqueeze(0), out.shape[-2] if lo_n is None else int(g["resize"]),
                                 out.shape[-1] if lo_n is None else int(g["resize"]),
                                 crop=None if lo_n is None else (lo_n[0], lo_n[0], lo_n[1], lo_n[1]),
                                 mean=V.CLIP_MEAN if lo_n is None else V.IMAGENET_MEAN,
                                 std=V.CLIP_STD if lo_n is None else V.IMAGENET_STD,
                                 divisor=255.0 if lo_n is None else 1.0, layout="B3THW", want_u8=True)
    np.testing.assert_array_equal(u8[0].cpu().numpy(), out)
    np.testing.assert_array_equal(f32[0].cpu().numpy(), norm)


NOAA_CASES = sorted(glob.glob(os.path.join(GOLDEN, "views_noaa_*.npz")))


@pytest.mark.parametrize("path", NOAA_CASES, ids=[os.path.basename(p)[:-4] for p in NOAA_CASES])
def test_views_without_antialias_match_torchvision_golden(path):
    """antialias=False (kvq_resize_view_bilinear_u8): what torchvision < 0.17 -- the torch ~= 1.10 environment the
    reference pins -- computes for Resize on a uint8 tensor.  Goldens: the reference functions' lines run with
    Resize(..., antialias=False) (tools/make_golden_views.py).  Bit-exact, like the anti-aliased path."""
    from datasets import views as V
    g = np.load(path)
    thwc = _golden_frames(g)
    video = thwc.permute(3, 0, 1, 2).contiguous().cuda()
    bt3hw = thwc.permute(0, 3, 1, 2).contiguous().unsqueeze(0).cuda()
    if str(g["kind"]) == "resize":
        out = V.get_resized_video(video, size_h=int(g["size_h"]), size_w=int(g["size_w"]), antialias=False)
        norm = V.resized_video_normalised(bt3hw, int(g["size_h"]), int(g["size_w"]), antialias=False)[0]
        aa = V.get_resized_video(video, size_h=int(g["size_h"]), size_w=int(g["size_w"]))
    else:
        out = V.get_resizecrop_video(video, resize=int(g["resize"]), crop=int(g["crop"]), phase="test", antialias=False)
        norm = V.resizecrop_video_normalised(bt3hw, int(g["resize"]), int(g["crop"]), antialias=False)[0]
        aa = V.get_resizecrop_video(video, resize=int(g["resize"]), crop=int(g["crop"]), phase="test")
    out, norm = out.cpu().numpy(), norm.cpu().numpy()
    np.testing.assert_array_equal(out, g["out"])
    assert _sha(out) == str(g["out_sha256"])
    np.testing.assert_array_equal(norm[:, :, ::7, ::5], g["norm_sample"])
    assert _sha(norm) == str(g["norm_sha256"])
    T, H, W = (int(v) for v in g["shape"])
    if H > 2 * out.shape[-2]:                     # a real down-scale: the two filters must differ (the flag is live)
        assert not np.array_equal(out, aa.cpu().numpy())
    # module-level switch = per-call argument
    V.RESIZE_ANTIALIAS = False
    try:
        if str(g["kind"]) == "resize":
            again = V.get_resized_video(video, size_h=int(g["size_h"]), size_w=int(g["size_w"]))
        else:
            again = V.get_resizecrop_video(video, resize=int(g["resize"]), crop=int(g["crop"]), phase="test")
    finally:
        V.RESIZE_ANTIALIAS = True
    np.testing.assert_array_equal(again.cpu().numpy(), out)


@pytest.mark.parametrize("B,T,Hs,Ws,oh,ow,crop", [
    (2, 3, 135, 241, 112, 112, None),            # KSVQE resize_video geometry, odd source, unaligned row starts
    (1, 2, 540, 960, 112, 112, None),            # 8.6x / 4.8x down-scaling: tap windows of 19 / 11
    (1, 2, 67, 45, 80, 90, None),                # enlargement on both axes (support 1, 3 taps)
    (2, 2, 131, 77, 52, 52, (4, 6, 44, 40)),     # rectangular crop window off the centre
    (1, 1, 9, 1030, 5, 3, None),                 # rows wider than one CTA pass, 2-row plane, 687-tap windows
    (1, 1, 33, 17, 33, 17, None),                # identity size: exact copy
])
def test_views_match_oracle_seeded(B, T, Hs, Ws, oh, ow, crop):
    from kvq_b200 import ops
    from oracle import views as O
    gen = torch.Generator().manual_seed(1000 + Hs * 7 + Ws)
    frames = torch.randint(0, 256, (B, T, 3, Hs, Ws), generator=gen, dtype=torch.uint8)
    u8, f32 = ops.resize_view_u8(frames.cuda(), oh, ow, crop=crop, divisor=255.0, mean=ops.CLIP_MEAN, std=ops.CLIP_STD,
                                 want_u8=True)
    ref = O.resize_u8(frames.permute(0, 2, 1, 3, 4).numpy(), oh, ow)                   # [B,3,T,oh,ow]
    if crop is not None:
        y, x, h, w = crop
        ref = ref[..., y:y + h, x:x + w]
    np.testing.assert_array_equal(u8.cpu().numpy(), ref)
    refn = np.stack([O.normalise(r, O.CLIP_MEAN, O.CLIP_STD, 255.0) for r in ref])
    np.testing.assert_array_equal(f32.cpu().numpy(), refn)
    if (oh, ow) == (Hs, Ws):
        np.testing.assert_array_equal(u8.cpu().numpy(), frames.permute(0, 2, 1, 3, 4).numpy())


@pytest.mark.parametrize("variant", ["2", "3", "4"])
@pytest.mark.parametrize("B,T,Hs,Ws,oh,ow,crop", [
    (1, 2, 135, 240, 112, 112, None),            # odd row count: the last row pair of a tile is half empty
    (2, 1, 37, 64, 20, 24, (3, 2, 15, 20)),      # tiles of one, two and three rows; crop window
    (1, 1, 50, 1920, 9, 112, None),              # decoder-wide rows, 37-tap windows, 8-row tiles
    (1, 2, 30, 44, 45, 66, None),                # enlargement (3 taps: the fused tail only)
    (1, 3, 21, 75, 10, 15, None),                # odd width: every row starts at another byte phase; 11-tap windows
    (2, 2, 19, 17, 7, 1, None),                  # one output column: the window is the whole 17-byte row
])
def test_views_row_kernels_agree_with_oracle(variant, B, T, Hs, Ws, oh, ow, crop, monkeypatch):
    """All W-axis kernels (KVQ_VIEWS_VARIANT=2: float32 rows in shared memory; 3: paired rows + packed fp32x2, rows of
    whole words; 4: raw bytes in shared memory, four taps per shared load) give the oracle's bytes and floats;
    KVQ_VIEWS_ROWS changes the tile height only."""
    from kvq_b200 import ops
    from oracle import views as O
    monkeypatch.setenv("KVQ_VIEWS_VARIANT", variant)
    gen = torch.Generator().manual_seed(4000 + Hs + Ws)
    frames = torch.randint(0, 256, (B, T, 3, Hs, Ws), generator=gen, dtype=torch.uint8)
    ref = O.resize_u8(frames.permute(0, 2, 1, 3, 4).numpy(), oh, ow)
    if crop is not None:
        y, x, h, w = crop
        ref = ref[..., y:y + h, x:x + w]
    refn = np.stack([O.normalise(r, O.IMAGENET_MEAN, O.IMAGENET_STD) for r in ref])
    for rows in (None, "4", "12"):
        if rows is None:
            monkeypatch.delenv("KVQ_VIEWS_ROWS", raising=False)
        else:
            monkeypatch.setenv("KVQ_VIEWS_ROWS", rows)
        u8, f32 = ops.resize_view_u8(frames.cuda(), oh, ow, crop=crop, want_u8=True)
        np.testing.assert_array_equal(u8.cpu().numpy(), ref)
        np.testing.assert_array_equal(f32.cpu().numpy(), refn)


def test_views_full_size_properties():
    """KSVQE's resize_video view at a decoder-sized source (8 frames of 1080x1920 -> 112x112): size-independent checks --
    a constant frame stays constant (the weights of every window sum to 1 within rounding: |v - c| <= 1 is the bound, 0
    is what ATen gives), frames are independent of their neighbours in the batch, layouts agree, and a horizontal /
    vertical flip of the source flips the view (the filter is symmetric) up to the last-bit order of summation."""
    from kvq_b200 import ops
    T, Hs, Ws = 8, 1080, 1920
    gen = torch.Generator().manual_seed(77)
    frames = torch.randint(0, 256, (1, T, 3, Hs, Ws), generator=gen, dtype=torch.uint8).cuda()
    u8, f32 = ops.resize_view_u8(frames, 112, 112, divisor=255.0, mean=ops.CLIP_MEAN, std=ops.CLIP_STD, want_u8=True)
    assert tuple(u8.shape) == (1, 3, T, 112, 112) and tuple(f32.shape) == (1, 3, T, 112, 112)
    # normalisation is the stated function of the bytes
    m = torch.tensor(ops.CLIP_MEAN, device="cuda").view(1, 3, 1, 1, 1)
    s = torch.tensor(ops.CLIP_STD, device="cuda").view(1, 3, 1, 1, 1)
    assert torch.equal(f32, (u8.float() / torch.full((1,), 255.0, device="cuda") - m) / s)   # tensor divisor: true division
    # layout B3THW on the permuted frames gives the same bytes
    u8b, _ = ops.resize_view_u8(frames.permute(0, 2, 1, 3, 4).contiguous(), 112, 112, layout="B3THW", want_u8=True,
                                want_f32=False)
    assert torch.equal(u8, u8b)
    # a sub-batch gives the same frames
    u8c, _ = ops.resize_view_u8(frames[:, 2:5].contiguous(), 112, 112, want_u8=True, want_f32=False)
    assert torch.equal(u8[:, :, 2:5], u8c)
    # constants are preserved
    const = torch.full((1, 2, 3, Hs, Ws), 201, dtype=torch.uint8, device="cuda")
    uc, _ = ops.resize_view_u8(const, 112, 112, want_u8=True, want_f32=False)
    assert int(uc.min()) == 201 and int(uc.max()) == 201
    # noise averaged over a 17x10 footprint: mean close to 127.5, far smaller spread than the source
    assert abs(float(u8.float().mean()) - 127.5) < 1.0 and float(u8.float().std()) < 15.0
    # symmetric filter: flipping the source flips the view within one grey level
    uf, _ = ops.resize_view_u8(frames.flip(-1).contiguous(), 112, 112, want_u8=True, want_f32=False)
    assert int((uf.flip(-1).int() - u8.int()).abs().max()) <= 1


def test_views_error_paths():
    from kvq_b200 import ops
    frames = torch.zeros((1, 2, 3, 16, 16), dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="crop window"):
        ops.resize_view_u8(frames, 8, 8, crop=(4, 4, 8, 8))
    with pytest.raises(RuntimeError, match="uint8"):
        ops.resize_view_u8(frames.float(), 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.resize_view_u8(frames.cpu(), 8, 8)


def test_arp_and_train_crop_match_reference_golden():
    """get_resized_video(arp=True) and get_resizecrop_video(phase='train') of the REAL reference (views_geometry.npz)
    through the C ABI; the same body runs on CPU with an oracle-backed kernel (tests/test_host_cpu.py)."""
    from datasets import views as V
    import views_cases
    views_cases.check_geometry(V, "cuda")


def test_ksvqe_from_raw_frames_matches_cpu_views():
    """Decoded uint8 frames -> (fragment kernel + resize kernel on the device) -> literal KSVQE key, against the same
    network fed with the views the golden-pinned CPU restatements build (oracle/fragments.py, oracle/views.py): the
    resized view is bit-exact, the fragment view within one ulp of the normalisation, the score unchanged."""
    import importlib.util
    import models
    from conftest import PKG
    from datasets import SyntheticKSVQEDataset
    from oracle import fragments as ofr
    from oracle import views as O
    from tools import synth
    spec = importlib.util.spec_from_file_location("kvq_trainer_views", os.path.join(PKG, "trainer.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    cfg = {"model": {"type": "KSVQE", "args": {"KSVQE": {
        "backbone": {"num_samples": 1, "sample_type": "topkpertubation", "CLIP_location": 8, "cls_use": True,
                     "tuning_stage": 2, "a1": 1.0, "a2": 1.0, "frozen_stages": -1},
        "head": {"in_channels": 768, "hidden_channels": 64}}}}}
    net = models.VQA_Network(cfg)
    sd = {k: (torch.ones(tuple(v.shape)) if k.endswith((".a1", ".a2")) else synth.fill_like(k, tuple(v.shape), 61))
          for k, v in net.state_dict().items() if v.is_floating_point()}
    net.load_state_dict(sd, strict=False)
    net = net.to("cuda:0").eval()
    st = {"fragments_h": 9, "fragments_w": 9, "fsize_h": 32, "fsize_w": 32, "size_h": 112, "size_w": 112, "aligned": 8,
          "clip_len": 32, "num_clips": 1}
    ds = SyntheticKSVQEDataset({"num_videos": 1, "raw_frames": True, "src_h": 360, "src_w": 640,
                                "sample_types": {"technical": st}})
    item = ds[0]
    assert item["frames"].shape == (32, 3, 360, 640) and item["frames"].dtype == torch.uint8
    dev = torch.device("cuda:0")
    data1 = next(iter(torch.utils.data.DataLoader(ds, batch_size=1)))                 # as Trainer.inferece feeds it
    s1 = float(tr.score_video(net, data1, ["KSVQE"], dev))

    frames = item["frames"][None]                                                     # [1,T,3,H,W]
    frag = ofr.fragment_clip(frames, item["offsets"][None], 9, 9, 32, 8)              # f32 [1,3,T,288,288]
    rv = O.normalise(O.resized_video(frames[0].permute(1, 0, 2, 3).numpy(), 112, 112), O.CLIP_MEAN, O.CLIP_STD, 255.0)
    np.testing.assert_array_equal(data1["resize_video"][0].cpu().numpy(), rv)
    assert tuple(data1["fragment"].shape) == (1, 3, 32, 288, 288)
    np.testing.assert_allclose(data1["fragment"].cpu().numpy(), frag.numpy(), rtol=0, atol=1e-6)
    data2 = {"fragment": frag, "resize_video": torch.from_numpy(rv)[None], "dis_label": item["dis_label"][None],
             "num_clips": item["num_clips"], "video_name": item["video_name"]}
    s2 = float(tr.score_video(net, data2, ["KSVQE"], dev))
    print("ksvqe raw-frame score", s1, "cpu-view score", s2)
    assert np.isfinite(s1) and abs(s1 - s2) < 1e-4
