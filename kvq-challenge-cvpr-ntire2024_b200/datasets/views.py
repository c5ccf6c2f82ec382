"""The reference's view functions in front of the models, on the B200 path.

Same names, argument meaning and results as datasets/fusion_datasets.py of the reference:
  get_resized_video(video, size_h, size_w)          :244-252  torchvision Resize((size_h, size_w)) on uint8 frames
  get_resizecrop_video(video, resize, crop, phase)  :299-316  Resize((resize, resize)) + centre crop (test phase)
  UnifiedFrameSampler                               :612-660  frame indices (host arithmetic, numpy's global RNG)
`video` is the reference's uint8 [3,T,H,W] CUDA tensor (or a batch [B,3,T,H,W]); the results equal the reference's byte
for byte (`kvq_resize_view_u8` keeps ATen's order of roundings).  `*_normalised` additionally fuse the datasets'
normalisation lines (:1017-1027, :902-905) and return the float32 model input directly.  There is no CPU fallback.

Which filter `Resize` applies to a uint8 tensor depends on the torchvision the reference runs under (requirements.txt pins
torch ~= 1.10 and leaves torchvision open): since torchvision 0.17 tensors are anti-aliased by default (`antialias=True`),
before that they were not.  Every function here takes `antialias` (default: the module-level RESIZE_ANTIALIAS = True, the
behaviour of a current install; set it to False, or pass the argument, to reproduce a checkpoint's pinned environment).
Both filters are bit-exact against goldens of torchvision's own output (tests/golden/views_*.npz)."""
import random

import numpy as np

from kvq_b200 import ops

IMAGENET_MEAN, IMAGENET_STD = ops.IMAGENET_MEAN, ops.IMAGENET_STD
CLIP_MEAN, CLIP_STD = ops.CLIP_MEAN, ops.CLIP_STD
RESIZE_ANTIALIAS = True      # torchvision >= 0.17 tensor default; False = torchvision < 0.17 (the torch 1.10 era)


def _aa(antialias):
    return RESIZE_ANTIALIAS if antialias is None else bool(antialias)


def _batched(video):
    if video.dim() == 4:
        return video.unsqueeze(0), True
    if video.dim() == 5:
        return video, False
    raise RuntimeError(f"view functions take uint8 [3,T,H,W] or [B,3,T,H,W] frames, got {tuple(video.shape)}")


def _resize_hw(size_h, size_w, src_h, src_w, arp):
    """get_resize_function (:229-241): with arp the longer side follows the source aspect ratio"""
    ratio = src_h / src_w if arp else 1
    if ratio > 1:
        size_h = int(ratio * size_w)
    elif ratio < 1:
        size_w = int(size_h / ratio)
    return size_h, size_w


def get_resized_video(video, size_h=224, size_w=224, random_crop=False, arp=False, antialias=None, **kwargs):
    if random_crop:
        raise NotImplementedError("get_resized_video(random_crop=True) is a training augmentation (RandomResizedCrop)")
    v, squeeze = _batched(video)
    size_h, size_w = _resize_hw(size_h, size_w, v.shape[-2], v.shape[-1], arp)
    out, _ = ops.resize_view_u8(v.contiguous(), size_h, size_w, layout="B3THW", want_u8=True, want_f32=False,
                                antialias=_aa(antialias))
    return out[0] if squeeze else out


def centre_crop_window(resize, crop):
    """:313-315: rows / columns [resize//2 - crop//2, resize//2 + crop//2) -> (first, count)"""
    lo = resize // 2 - crop // 2
    return lo, resize // 2 + crop // 2 - lo


def get_resizecrop_video(video, resize=520, crop=448, phase="train", antialias=None, **kwargs):
    v, squeeze = _batched(video)
    if phase == "train":   # same draws as the reference (:308-310): random.randrange for rows, then columns
        y, x, n = random.randrange(resize - crop), random.randrange(resize - crop), crop
    else:
        y, n = centre_crop_window(resize, crop)
        x = y
    out, _ = ops.resize_view_u8(v.contiguous(), resize, resize, crop=(y, x, n, n), layout="B3THW", want_u8=True,
                                want_f32=False, antialias=_aa(antialias))
    return out[0] if squeeze else out


def resized_video_normalised(frames, size_h=224, size_w=224, mean=CLIP_MEAN, std=CLIP_STD, divisor=255.0,
                             layout="BT3HW", antialias=None):
    """decoder-order frames u8 [B,T,3,H,W] -> the KSVQE `resize_video` input f32 [B,3,T,size_h,size_w]:
    get_resized_video + (v / 255 - clip_mean) / clip_std (:1021-1027) in one pass"""
    return ops.resize_view_u8(frames, size_h, size_w, mean=mean, std=std, divisor=divisor, layout=layout,
                              antialias=_aa(antialias))[1]


def resizecrop_video_normalised(frames, resize=520, crop=448, mean=IMAGENET_MEAN, std=IMAGENET_STD, layout="BT3HW",
                                antialias=None):
    """decoder-order frames u8 [B,T,3,H,W] -> the SimpleVQA input f32 [B,3,T,crop,crop]: get_resizecrop_video (test
    phase) + (v - mean) / std (:902-905) in one pass"""
    lo, n = centre_crop_window(resize, crop)
    return ops.resize_view_u8(frames, resize, resize, crop=(lo, lo, n, n), mean=mean, std=std, layout=layout,
                              antialias=_aa(antialias))[1]


class UnifiedFrameSampler:
    """:612-660.  `fragments_t` temporal cells of total // fragments_t frames; inside each cell fsize_t frames
    frame_interval apart, starting at a uniform offset drawn from numpy's global RNG (one draw of fragments_t offsets per
    clip, only when the cell is longer than the span); cells listed in random.sample(...) are dropped (drop_rate);
    indices wrap modulo the frame count."""

    def __init__(self, fsize_t, fragments_t, frame_interval=1, num_clips=1, drop_rate=0.0):
        self.fragments_t = fragments_t
        self.fsize_t = fsize_t
        self.size_t = fragments_t * fsize_t
        self.frame_interval = frame_interval
        self.num_clips = num_clips
        self.drop_rate = drop_rate

    def get_frame_indices(self, num_frames, train=False):
        cell = num_frames // self.fragments_t
        span = self.fsize_t * self.frame_interval
        starts = np.arange(self.fragments_t, dtype=np.int32) * np.int32(cell)
        if cell > span:
            starts = starts + np.random.randint(0, cell - span, size=self.fragments_t)
        dropped = set(random.sample(list(range(self.fragments_t)), int(self.fragments_t * self.drop_rate)))
        steps = np.arange(self.fsize_t) * self.frame_interval
        return np.concatenate([starts[i] + steps for i in range(self.fragments_t) if i not in dropped])

    def __call__(self, total_frames, train=False, start_index=0):
        inds = np.concatenate([self.get_frame_indices(total_frames) for _ in range(self.num_clips)])
        return np.mod(inds + start_index, total_frames).astype(np.int32)
