"""oracle/views.py against the golden vectors the REAL reference produced (tools/make_golden_views.py ran
datasets/fusion_datasets.get_resized_video / get_resizecrop_video / UnifiedFrameSampler and the datasets' normalisation
lines): uint8 views and normalised float32 views bit-exact, frame indices equal under the same numpy seed."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import views

# anti-aliased goldens (torchvision >= 0.17 tensor default) and the views_noaa_* twins (antialias=False: torchvision
# < 0.17, the torch ~= 1.10 environment the reference pins)
CASES = sorted(glob.glob(os.path.join(GOLDEN, "views_resize*.npz")) + glob.glob(os.path.join(GOLDEN, "views_noaa_*.npz")))


def golden_frames(g):
    """the generator's frames: [T,H,W,3] u8 from the seeded torch generator, viewed as [3,T,H,W]"""
    T, H, W = (int(v) for v in g["shape"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    return torch.randint(0, 256, (T, H, W, 3), generator=gen, dtype=torch.uint8).permute(3, 0, 1, 2)


def golden_view(g, video):
    aa = bool(int(g["antialias"])) if "antialias" in g else True
    if str(g["kind"]) == "resize":
        out = views.resized_video(video, size_h=int(g["size_h"]), size_w=int(g["size_w"]), antialias=aa)
        return out, views.normalise(out, views.CLIP_MEAN, views.CLIP_STD, 255.0)
    out = views.resizecrop_video(video, resize=int(g["resize"]), crop=int(g["crop"]), antialias=aa)
    return out, views.normalise(out, views.IMAGENET_MEAN, views.IMAGENET_STD)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_views_match_reference_golden(path):
    g = np.load(path)
    out, norm = golden_view(g, golden_frames(g).numpy())
    np.testing.assert_array_equal(out, g["out"])
    assert sha(out) == str(g["out_sha256"])
    assert sha(norm) == str(g["norm_sha256"])                       # every float32 of the normalised view
    np.testing.assert_array_equal(norm[:, :, ::7, ::5], g["norm_sample"])


def test_frame_sampler_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "views_frame_sampler.npz"))
    for i, (fsize_t, fragments_t, interval, num_clips, total, seed) in enumerate(g["cases"]):
        np.random.seed(int(seed))
        inds = views.frame_indices(int(total), int(fsize_t), int(fragments_t), int(interval), int(num_clips))
        assert inds.dtype == np.int32
        np.testing.assert_array_equal(inds, g[f"inds_{i}"])


def test_aa_weights_properties():
    for n_in, n_out in [(1920, 112), (1080, 520), (64, 112), (7, 7), (5, 1), (1, 3)]:
        xmin, xsize, w = views.aa_weights(n_in, n_out)
        assert (xmin >= 0).all() and (xmin + xsize <= n_in).all() and (xsize >= 1).all()
        assert (np.diff(xmin) >= 0).all() and (np.diff(xmin + xsize) >= 0).all()      # windows move monotonically
        np.testing.assert_allclose(w.sum(1), 1.0, atol=2e-6)
        assert w.shape[1] == views.aa_taps(n_in, n_out) and (xsize <= w.shape[1]).all()
    xmin, xsize, w = views.aa_weights(9, 9)                          # identity resize: one tap of weight 1 ... or 3 taps
    ident = views.interpolate_bilinear_aa(np.arange(81, dtype=np.float32).reshape(9, 9), 9, 9)
    np.testing.assert_array_equal(ident, np.arange(81, dtype=np.float32).reshape(9, 9))


def test_arp_and_train_crop_match_reference_golden():
    """get_resized_video(arp=True) (aspect-ratio-preserving target size, fusion_datasets.py:229-241) and
    get_resizecrop_video(phase='train') (random.randrange rows, then columns, :308-311) of the REAL reference."""
    import random
    g = np.load(os.path.join(GOLDEN, "views_geometry.npz"))
    for i, (T, H, W, sh, sw, seed) in enumerate(g["arp_cases"]):
        gen = torch.Generator().manual_seed(int(seed))
        video = torch.randint(0, 256, (int(T), int(H), int(W), 3), generator=gen, dtype=torch.uint8).permute(3, 0, 1, 2)
        oh, ow = views.resize_hw(int(sh), int(sw), int(H), int(W), arp=True)
        np.testing.assert_array_equal(views.resize_u8(video.numpy(), oh, ow), g[f"arp_out_{i}"])
    for i, (T, H, W, rs, cr, seed, rseed) in enumerate(g["train_cases"]):
        gen = torch.Generator().manual_seed(int(seed))
        video = torch.randint(0, 256, (int(T), int(H), int(W), 3), generator=gen, dtype=torch.uint8).permute(3, 0, 1, 2)
        random.seed(int(rseed))
        y, x = views.train_crop_window(int(rs), int(cr), random)
        assert [y, x] == g[f"train_yx_{i}"].tolist()
        np.testing.assert_array_equal(views.resize_u8(video.numpy(), int(rs), int(rs))[..., y:y + int(cr), x:x + int(cr)],
                                      g[f"train_out_{i}"])


def _byte_perm(lo, hi, sel):
    b = [(lo >> (8 * i)) & 255 for i in range(4)] + [(hi >> (8 * i)) & 255 for i in range(4)]
    return sum(b[(sel >> (4 * e)) & 7] << (8 * e) for e in range(4))


def _emulate_rows_bytes_kernel(tile, head, out_w):
    """Python walk of resize_rows_bytes_kernel (csrc/kvq_views.cu): raw bytes at a 16 B phase `head` in shared memory,
    aligned word pairs funnelled by PRMT, taps taken four at a time in three rounding regimes (product; 4*((n-1)//4)
    rounded product + rounded add; fused tail)."""
    F = np.float32
    rows, Ws = tile.shape
    xmin, xsize, wts = views.aa_weights(Ws, out_w)
    nbytes = rows * Ws
    nvec = (head + nbytes + 15) >> 4
    sb = np.zeros((nvec + 1) * 16, np.uint8)
    sb[head:head + nbytes] = tile.reshape(-1)
    sw = sb.view("<u4")
    fma = lambda t, v, w: F(np.float64(t) + np.float64(v) * np.float64(w))
    out = np.zeros((rows, out_w), F)
    for r in range(rows):
        for o in range(out_w):
            n, x0 = int(xsize[o]), int(xmin[o])
            bo = head + r * Ws + x0
            wi, sel = bo >> 2, 0x3210 + 0x1111 * (bo & 3)
            m = (n - 1) >> 2
            b4 = _byte_perm(int(sw[wi]), int(sw[wi + 1]), sel)
            byte = lambda e: F((b4 >> (8 * e)) & 255)
            acc = F(byte(0) * wts[o, 0])
            if m == 0:
                for e in range(1, 4):
                    if e < n:
                        acc = fma(acc, byte(e), wts[o, e])
            else:
                for e in range(1, 4):
                    acc = F(acc + F(byte(e) * wts[o, e]))
                for g in range(1, m):
                    b4 = _byte_perm(int(sw[wi + g]), int(sw[wi + g + 1]), sel)
                    for e in range(4):
                        acc = F(acc + F(byte(e) * wts[o, 4 * g + e]))
                b4 = _byte_perm(int(sw[wi + m]), int(sw[wi + m + 1]), sel)
                acc = F(acc + F(byte(0) * wts[o, 4 * m]))
                for e in range(1, 4):
                    if 4 * m + e < n:
                        acc = fma(acc, byte(e), wts[o, 4 * m + e])
            out[r, o] = acc
    return out


@pytest.mark.parametrize("rows,Ws,out_w,head", [(5, 75, 15, 3), (4, 64, 24, 0), (3, 17, 1, 9), (6, 44, 66, 15),
                                               (2, 240, 14, 7), (3, 100, 33, 1), (2, 37, 37, 2)])
def test_byte_funnel_tap_order_equals_oracle(rows, Ws, out_w, head):
    """The index / funnel / rounding-regime logic of the default W-axis kernel, walked in Python, gives the oracle's
    float32 bits for any byte phase of the tile and of the rows (the kernel itself: tests/test_gpu_views.py)."""
    rng = np.random.default_rng(rows * 1000 + Ws)
    tile = rng.integers(0, 256, (rows, Ws), dtype=np.uint8)
    ref = views.reduce_last_axis(tile.astype(np.float32), out_w)
    got = _emulate_rows_bytes_kernel(tile, head, out_w)
    np.testing.assert_array_equal(ref.view(np.uint32), got.view(np.uint32))
