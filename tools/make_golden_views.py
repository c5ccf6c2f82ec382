"""Golden vectors for the view pipeline: runs the REAL reference functions
datasets/fusion_datasets.{get_resized_video, get_resizecrop_video, UnifiedFrameSampler} and the datasets' normalisation
lines on seeded uint8 frames, in the authoring container.  Usage: python -m tools.make_golden_views

Also asserts, on every case, that oracle/views.py reproduces torch's float32 anti-aliased interpolate bit-exactly
BEFORE the uint8 rounding (the claim in that file's header)."""
import hashlib
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_import, views

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# (name, T, H, W, kind, kwargs, seed)
RESIZE_CASES = [
    ("views_resize_t3_270x480", 3, 270, 480, "resize", dict(size_h=112, size_w=112), 61),      # KSVQE resize_video, landscape
    ("views_resize_t2_1080x608", 2, 1080, 608, "resize", dict(size_h=112, size_w=112), 62),    # portrait short video, 21 / 13 taps
    ("views_resize_up_t3_64x96", 3, 64, 96, "resize", dict(size_h=112, size_w=112), 63),       # enlargement on both axes
    ("views_resize_mixed_t2_100x300", 2, 100, 300, "resize", dict(size_h=112, size_w=112), 64),  # H enlarged, W reduced
    ("views_resize_rect_t2_90x70", 2, 90, 70, "resize", dict(size_h=48, size_w=80), 65),       # size_h != size_w
    ("views_resizecrop_t2_360x640", 2, 360, 640, "resizecrop", dict(resize=520, crop=448), 66),  # kwai_simpleVQA_test.yml view
    ("views_resizecrop_odd_t2_90x70", 2, 90, 70, "resizecrop", dict(resize=65, crop=47), 67),  # odd resize / crop: floor halves
]

# (fsize_t, fragments_t, frame_interval, num_clips, total_frames, numpy seed)
SAMPLER_CASES = [
    (32, 3, 4, 1, 600, 71),       # Kwai_KSVQE_test.yml: UnifiedFrameSampler(clip_len, num_clips, frame_interval)
    (32, 3, 4, 1, 300, 72),       # cell of 100 frames <= span 128: no draw, indices wrap
    (1, 8, 10, 1, 250, 73),       # kwai_simpleVQA_test.yml: clip_len // t_frag = 1, t_frag = 8
    (4, 2, 2, 3, 97, 74),         # three clips, ragged total
    (32, 1, 2, 4, 40, 75),        # the 40-frame probe the datasets print at construction
]


# get_resized_video(arp=True): (T, H, W, size_h, size_w, seed);  get_resizecrop_video(phase='train'): (T, H, W, resize, crop,
# frame seed, `random` seed)
ARP_CASES = [(2, 60, 90, 32, 32, 81), (2, 90, 60, 32, 32, 82), (1, 50, 50, 24, 40, 83), (1, 1080 // 8, 608 // 8, 28, 28, 84)]
TRAIN_CROP_CASES = [(2, 70, 110, 40, 28, 91, 5), (1, 64, 64, 33, 20, 92, 6), (2, 40, 30, 48, 47, 93, 7)]


def seeded_frames(T, H, W, seed):
    """decord-like frames [T,H,W,3] u8 stacked and permuted exactly as spatial_temporal_view_decomposition does."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randint(0, 256, (T, H, W, 3), generator=g, dtype=torch.uint8)
    return torch.stack(list(frames), 0).permute(3, 0, 1, 2)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = ref_import.load_reference()
    fd = ref.fusion
    mean, std = torch.FloatTensor(views.IMAGENET_MEAN), torch.FloatTensor(views.IMAGENET_STD)
    cmean, cstd = torch.FloatTensor(views.CLIP_MEAN), torch.FloatTensor(views.CLIP_STD)
    for name, T, H, W, kind, kw, seed in RESIZE_CASES:
        video = seeded_frames(T, H, W, seed)
        if kind == "resize":
            out = fd.get_resized_video(video, **kw)
            # fusion_datasets.py:1021-1024
            norm = ((out.permute(1, 2, 3, 0) / 255.0 - cmean) / cstd).permute(3, 0, 1, 2)
            oh, ow = kw["size_h"], kw["size_w"]
            mine = views.resized_video(video.numpy(), **kw)
            mine_norm = views.normalise(mine, views.CLIP_MEAN, views.CLIP_STD, 255.0)
        else:
            out = fd.get_resizecrop_video(video, phase="test", **kw)
            norm = ((out.permute(1, 2, 3, 0) - mean) / std).permute(3, 0, 1, 2)     # fusion_datasets.py:902-905
            oh = ow = kw["resize"]
            mine = views.resizecrop_video(video.numpy(), **kw)
            mine_norm = views.normalise(mine, views.IMAGENET_MEAN, views.IMAGENET_STD)
        assert out.dtype == torch.uint8
        # the header claim of oracle/views.py: float32 interpolate reproduced bit for bit
        x = video.permute(1, 0, 2, 3).to(torch.float32)
        f_ref = F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False, antialias=True).numpy()
        f_mine = views.interpolate_bilinear_aa(x.numpy(), oh, ow)
        assert np.array_equal(f_ref, f_mine), (name, int((f_ref != f_mine).sum()))
        assert np.array_equal(mine, out.numpy()), name
        assert np.array_equal(mine_norm, norm.contiguous().numpy()), name
        o = out.contiguous().numpy()
        n = norm.contiguous().numpy()
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), out=o, out_sha256=np.array(sha(o)),
                            norm_sha256=np.array(sha(n)), norm_sample=n[:, :, ::7, ::5], kind=np.array(kind),
                            shape=np.array([T, H, W]), seed=seed, **{k: np.array(v) for k, v in kw.items()})
        print(name, tuple(out.shape), "mean", float(o.mean()), "float32 interpolate bit-exact, u8 exact, norm exact")
    # ---- the same views under torchvision < 0.17 semantics (tensors are NOT anti-aliased: what the torch ~= 1.10
    # environment pinned by the reference computes): the reference functions' own lines with Resize(..., antialias=False)
    import torchvision
    for name, T, H, W, kind, kw, seed in RESIZE_CASES:
        video = seeded_frames(T, H, W, seed)
        if kind == "resize":
            oh, ow = kw["size_h"], kw["size_w"]
            out = torchvision.transforms.Resize((oh, ow), antialias=False)(video.permute(1, 0, 2, 3)).permute(1, 0, 2, 3)
            norm = ((out.permute(1, 2, 3, 0) / 255.0 - cmean) / cstd).permute(3, 0, 1, 2)
            mine = views.resized_video(video.numpy(), antialias=False, **kw)
            mine_norm = views.normalise(mine, views.CLIP_MEAN, views.CLIP_STD, 255.0)
        else:
            oh = ow = kw["resize"]
            r = torchvision.transforms.Resize((oh, ow), antialias=False)(video.permute(1, 0, 2, 3)).permute(1, 0, 2, 3)
            lo, n = views.centre_crop_window(kw["resize"], kw["crop"])
            out = r[..., lo:lo + n, lo:lo + n]                                       # fusion_datasets.py:313-315
            norm = ((out.permute(1, 2, 3, 0) - mean) / std).permute(3, 0, 1, 2)
            mine = views.resizecrop_video(video.numpy(), antialias=False, **kw)
            mine_norm = views.normalise(mine, views.IMAGENET_MEAN, views.IMAGENET_STD)
        x = video.permute(1, 0, 2, 3).to(torch.float32)
        f_ref = F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False, antialias=False).numpy()
        f_mine = views.interpolate_bilinear(x.numpy(), oh, ow)
        assert np.array_equal(f_ref, f_mine), (name, int((f_ref != f_mine).sum()))
        assert np.array_equal(mine, out.numpy()), name
        assert np.array_equal(mine_norm, norm.contiguous().numpy()), name
        o, n_ = out.contiguous().numpy(), norm.contiguous().numpy()
        nm = name.replace("views_", "views_noaa_")
        np.savez_compressed(os.path.join(GOLD, nm + ".npz"), out=o, out_sha256=np.array(sha(o)),
                            norm_sha256=np.array(sha(n_)), norm_sample=n_[:, :, ::7, ::5], kind=np.array(kind),
                            shape=np.array([T, H, W]), seed=seed, antialias=np.array(0),
                            **{k: np.array(v) for k, v in kw.items()})
        print(nm, tuple(out.shape), "mean", float(o.mean()), "plain bilinear: float32 interpolate bit-exact, u8 exact, norm exact")
    import random
    geo = {"arp_cases": np.array(ARP_CASES), "train_cases": np.array(TRAIN_CROP_CASES)}
    for i, (T, H, W, sh, sw, seed) in enumerate(ARP_CASES):
        video = seeded_frames(T, H, W, seed)
        out = fd.get_resized_video(video, size_h=sh, size_w=sw, arp=True).contiguous().numpy()
        oh, ow = views.resize_hw(sh, sw, H, W, arp=True)
        assert out.shape[-2:] == (oh, ow), (out.shape, oh, ow)
        assert np.array_equal(out, views.resize_u8(video.numpy(), oh, ow))
        geo[f"arp_out_{i}"] = out
        print("arp", (T, H, W, sh, sw), "->", out.shape)
    for i, (T, H, W, rs, cr, seed, rseed) in enumerate(TRAIN_CROP_CASES):
        video = seeded_frames(T, H, W, seed)
        random.seed(rseed)
        out = fd.get_resizecrop_video(video, resize=rs, crop=cr, phase="train").contiguous().numpy()
        random.seed(rseed)
        y, x = views.train_crop_window(rs, cr, random)
        assert np.array_equal(out, views.resize_u8(video.numpy(), rs, rs)[..., y:y + cr, x:x + cr])
        geo[f"train_out_{i}"] = out
        geo[f"train_yx_{i}"] = np.array([y, x])
        print("train crop", (T, H, W, rs, cr), "window", (y, x), "->", out.shape)
    np.savez_compressed(os.path.join(GOLD, "views_geometry.npz"), **geo)
    rows = []
    for fsize_t, fragments_t, interval, num_clips, total, seed in SAMPLER_CASES:
        sampler = fd.UnifiedFrameSampler(fsize_t, fragments_t, interval, num_clips)
        np.random.seed(seed)
        inds = sampler(total, False)
        np.random.seed(seed)
        mine = views.frame_indices(total, fsize_t, fragments_t, interval, num_clips)
        assert inds.dtype == mine.dtype and np.array_equal(inds, mine), (fsize_t, fragments_t, interval, num_clips, total)
        rows.append(inds)
        print("sampler", (fsize_t, fragments_t, interval, num_clips, total), inds[:6], "...", len(inds))
    np.savez_compressed(os.path.join(GOLD, "views_frame_sampler.npz"), cases=np.array(SAMPLER_CASES),
                        **{f"inds_{i}": r for i, r in enumerate(rows)})


if __name__ == "__main__":
    main()
