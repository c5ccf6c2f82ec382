"""CPU restatement of the reference's view pipeline in front of the models.

TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline leg).  Pinned to the real reference by
tests/golden/views_*.npz: tools/make_golden_views.py imports datasets/fusion_datasets.py from /root/reference and runs
get_resized_video / get_resizecrop_video / UnifiedFrameSampler and the normalisation lines of the two datasets on
seeded uint8 frames; tests/test_oracle_views.py replays them (bit-exact, see below).

Reference walk (paths relative to the reference root):
  datasets/fusion_datasets.py:229-241   get_resize_function -> torchvision.transforms.Resize((size_h, size_w))
  datasets/fusion_datasets.py:244-252   get_resized_video    [3,T,H,W] u8 -> [3,T,size_h,size_w] u8
  datasets/fusion_datasets.py:299-316   get_resizecrop_video  Resize((resize, resize)) + centre crop (test phase)
  datasets/fusion_datasets.py:612-660   UnifiedFrameSampler   frame indices of the num_clips x fragments_t x fsize_t grid
  datasets/fusion_datasets.py:1017-1027 KVQ dataset: (v - mean) / std for the fragment view, (v / 255 - clip_mean) /
                                        clip_std for the resized view
  datasets/fusion_datasets.py:902-905   SimpleVQA dataset: (v - mean) / std

Third-party arithmetic: `torchvision.transforms.Resize` on a uint8 tensor (torchvision is unpinned in
requirements.txt:2; 0.26 installed here) = cast to float32, `torch.nn.functional.interpolate(mode='bilinear',
align_corners=False, antialias=True)`, `torch.round`, cast back (torchvision/transforms/_functional_tensor.py
`resize`).  The interpolate is ATen's separable anti-aliased kernel (aten/src/ATen/native/cpu/UpSampleKernel.cpp):
  * per output index i: scale = float(in) / out; support = scale if scale >= 1 else 1; center = scale * (i + 0.5);
    taps [xmin, xmin + xsize) with xmin = max(int(center - support + 0.5), 0), xmax = min(int(center + support + 0.5),
    in); triangle weights max(0, 1 - |(j + xmin - center + 0.5) / max(scale, 1)|), each DIVIDED by their float32 sum;
  * the W axis is reduced first, the H axis second, the intermediate stays float32;
  * each reduction is `t = v0 * w0; for j in 1..n-1: t += v_j * w_j` -- and the order of roundings is what the x86
    build of torch 2.11 executes (measured, tools/make_golden_views.py asserts it): the first ((n - 1) // 4) * 4
    terms of the loop are a rounded product followed by a rounded add (a 4-lane in-order vector reduction), the
    remaining (n - 1) % 4 terms are fused multiply-adds (the scalar tail, contracted by the compiler).
With that order this file reproduces `F.interpolate` BIT-EXACTLY (every float32, hence every rounded uint8).
"""
import numpy as np

F32 = np.float32

IMAGENET_MEAN = (123.675, 116.28, 103.53)
IMAGENET_STD = (58.395, 57.12, 57.375)
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def aa_taps(in_size, out_size):
    """ATen max_interp_size for the bilinear filter: ceil(support) * 2 + 1."""
    scale = F32(F32(in_size) / F32(out_size))
    support = scale if scale >= 1.0 else F32(1.0)
    return int(np.ceil(support)) * 2 + 1


def aa_weights(in_size, out_size):
    """-> (xmin int64 [out], xsize int64 [out], weights float32 [out, taps]); HelperInterpBase::
    _compute_indices_min_size_weights_aa with scalar_t = float (mixed float / double expression types kept)."""
    scale = F32(F32(in_size) / F32(out_size))
    support = scale if scale >= 1.0 else F32(1.0)
    taps = int(np.ceil(support)) * 2 + 1
    invscale = F32(1.0 / float(scale)) if scale >= 1.0 else F32(1.0)
    xmin = np.zeros(out_size, np.int64)
    xsize = np.zeros(out_size, np.int64)
    weights = np.zeros((out_size, taps), F32)
    for i in range(out_size):
        center = F32(float(scale) * (i + 0.5))
        lo = int(float(F32(center - support)) + 0.5)
        hi = int(float(F32(center + support)) + 0.5)
        x0 = max(lo, 0)
        n = min(max(min(hi, in_size) - x0, 0), taps)
        total = F32(0)
        for j in range(n):
            d = F32(F32(j + x0) - center)
            x = abs(F32((float(d) + 0.5) * float(invscale)))
            w = F32(1.0) - x if x < 1.0 else F32(0)
            weights[i, j] = w
            total = F32(total + w)
        if total != 0:
            weights[i, :n] = (weights[i, :n] / total).astype(F32)
        xmin[i], xsize[i] = x0, n
    return xmin, xsize, weights


def _fma(t, v, w):
    # float32 fused multiply-add through float64: the product of two float32 is exact in float64
    return (t.astype(np.float64) + v.astype(np.float64) * np.float64(w)).astype(F32)


def reduce_last_axis(x, out_size, lanes=4):
    """One separable pass over the last axis of float32 `x` in ATen's order of roundings (module docstring)."""
    xmin, xsize, weights = aa_weights(x.shape[-1], out_size)
    out = np.zeros(x.shape[:-1] + (out_size,), F32)
    for o in range(out_size):
        n = int(xsize[o])
        if n == 0:
            continue
        t = (x[..., xmin[o]] * weights[o, 0]).astype(F32)
        unfused = ((n - 1) // lanes) * lanes
        for k in range(n - 1):
            v, w = x[..., xmin[o] + k + 1], weights[o, k + 1]
            t = (t + (v * w).astype(F32)).astype(F32) if k < unfused else _fma(t, v, w)
        out[..., o] = t
    return out


def interpolate_bilinear_aa(x, out_h, out_w):
    """float32 [..., H, W] -> float32 [..., out_h, out_w] = F.interpolate(x, (out_h, out_w), mode='bilinear',
    align_corners=False, antialias=True) on CPU."""
    x = np.ascontiguousarray(x, dtype=F32)
    h = reduce_last_axis(x, out_w)                                   # W first
    v = reduce_last_axis(np.ascontiguousarray(np.swapaxes(h, -1, -2)), out_h)
    return np.ascontiguousarray(np.swapaxes(v, -1, -2))


def bilinear_taps(in_size, out_size):
    """ATen upsample_bilinear2d source index / lambdas (align_corners=False), float32 as the x86 build computes them:
    scale = float(in) / float(out); src = max(fma(scale, i + 0.5, -0.5), 0) -- the multiply-add is contracted, pinned
    against torch in tools/make_golden_views.py"""
    i = np.arange(out_size)
    scale = np.float32(in_size) / np.float32(out_size)
    src = (np.float64(scale) * (i.astype(np.float64) + 0.5) - 0.5).astype(np.float32)   # one rounding = fma
    src = np.maximum(src, np.float32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    l1 = np.clip(src - i0.astype(np.float32), 0, 1).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return i0, i1, l0, l1


def _fma_exact(a, b, c):
    return (a.astype(np.longdouble) * b.astype(np.longdouble) + c.astype(np.longdouble)).astype(np.float32)


def interpolate_bilinear(x, out_h, out_w):
    """float32 [..., H, W] -> [..., out_h, out_w]: F.interpolate(mode="bilinear", align_corners=False, antialias=False)
    on CPU, bit for bit: W axis inside, H axis outside, each fma(v0, l0, round(v1 * l1))."""
    x = np.asarray(x, dtype=np.float32)
    y0, y1, hy0, hy1 = bilinear_taps(x.shape[-2], out_h)
    x0, x1, wx0, wx1 = bilinear_taps(x.shape[-1], out_w)
    r = _fma_exact(x[..., x0], wx0, (x[..., x1] * wx1).astype(np.float32))
    t0, t1 = r[..., y0, :], r[..., y1, :]
    return _fma_exact(t0, hy0[:, None], (t1 * hy1[:, None]).astype(np.float32))


def resize_u8(video, out_h, out_w, antialias=True):
    if not antialias:
        f = interpolate_bilinear(np.asarray(video).astype(np.float32), out_h, out_w)
        return np.clip(np.rint(f), 0, 255).astype(np.uint8)
    """torchvision Resize((out_h, out_w)) on a uint8 array [..., H, W]: float32 interpolate, round half to even, cast."""
    return np.rint(interpolate_bilinear_aa(video.astype(F32), out_h, out_w)).astype(np.uint8)


def resized_video(video, size_h=224, size_w=224, antialias=True, **_):
    """get_resized_video (fusion_datasets.py:244-252), arp=False: Resize((size_h, size_w)) on the [T,3,H,W] permutation."""
    return resize_u8(video, size_h, size_w, antialias)


def resize_hw(size_h, size_w, src_h, src_w, arp=False):
    """get_resize_function (fusion_datasets.py:229-241) as get_resized_video calls it (:247-249): with arp the target
    ratio is src_h / src_w and the LONGER side is recomputed from the other one (int() truncation)."""
    ratio = src_h / src_w if arp else 1
    if ratio > 1:
        size_h = int(ratio * size_w)
    elif ratio < 1:
        size_w = int(size_h / ratio)
    return size_h, size_w


def train_crop_window(resize, crop, rng):
    """get_resizecrop_video, train phase (fusion_datasets.py:308-311): rows then columns, random.randrange(res - crop)."""
    y = rng.randrange(resize - crop)
    x = rng.randrange(resize - crop)
    return y, x


def centre_crop_window(resize, crop):
    """fusion_datasets.py:313-315: rows / columns [resize//2 - crop//2, resize//2 + crop//2)."""
    lo = resize // 2 - crop // 2
    return lo, resize // 2 + crop // 2 - lo


def resizecrop_video(video, resize=520, crop=448, antialias=True, **_):
    """fusion_datasets.py:299-316, test phase: Resize((resize, resize)) then the centre crop."""
    r = resize_u8(video, resize, resize, antialias)
    lo, n = centre_crop_window(resize, crop)
    return r[..., lo:lo + n, lo:lo + n]


def normalise(video_u8, mean, std, divisor=1.0):
    """[3,T,h,w] u8 -> float32 ((v / divisor) - mean[c]) / std[c]  (fusion_datasets.py:1017-1024, :902-905)."""
    v = video_u8.astype(F32)
    if divisor != 1.0:
        v = v / F32(divisor)
    m = np.asarray(mean, F32).reshape(3, 1, 1, 1)
    s = np.asarray(std, F32).reshape(3, 1, 1, 1)
    return ((v - m) / s).astype(F32)


def frame_indices(num_frames, fsize_t, fragments_t, frame_interval=1, num_clips=1, start_index=0, rng=None):
    """UnifiedFrameSampler.__call__ (fusion_datasets.py:612-660, drop_rate = 0): per clip, `fragments_t` temporal cells
    of `num_frames // fragments_t` frames; inside each cell `fsize_t` frames `frame_interval` apart starting at a
    uniform offset (one np.random.randint(0, tlength - fsize_t * frame_interval, size=fragments_t) draw per clip, only
    when the cell is longer than the span); indices wrap modulo num_frames.  rng=None uses numpy's global RNG like the
    reference."""
    rng = np.random if rng is None else rng
    tlength = num_frames // fragments_t
    span = fsize_t * frame_interval
    clips = []
    for _ in range(num_clips):
        cell0 = np.arange(fragments_t, dtype=np.int64) * tlength
        if tlength > span:
            off = rng.randint(0, tlength - span, size=fragments_t)
        else:
            off = np.zeros(fragments_t, np.int64)
        clips.append((cell0[:, None] + off[:, None] + np.arange(fsize_t)[None, :] * frame_interval).reshape(-1))
    return np.mod(np.concatenate(clips) + start_index, num_frames).astype(np.int32)
