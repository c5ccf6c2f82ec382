"""CPU restatement of Grid Mini-patch Sampling + normalisation (the step in front of the Swin backbone).

TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline leg).  Pinned to the real reference by
tests/golden/fragments_*.npz (tools/make_golden_extra.py runs datasets/fusion_datasets.get_spatial_fragments under a
seeded global RNG; tests/test_oracle_fragments.py replays the same draws).

Reference walk (paths relative to the reference root):
  datasets/fusion_datasets.py:22-121    get_spatial_fragments (random=False branch: per-cell offsets, h then w)
  datasets/fusion_datasets.py:1017-1020 (v - mean) / std with ImageNet*255 statistics
"""
import torch
import torch.nn.functional as F

MEAN = (123.675, 116.28, 103.53)
STD = (58.395, 57.12, 57.375)


def draw_offsets(res_h, res_w, dur_t, fragments_h=7, fragments_w=7, fsize_h=32, fsize_w=32, aligned=32, generator=None):
    """fusion_datasets.py:73-89: rnd_h then rnd_w from torch.randint(hlength - fsize, (fh, fw, T // aligned)) (zeros
    when a cell is not larger than a fragment).  generator=None draws from the global RNG like the reference."""
    if dur_t == 1:
        aligned = 1
    n = (fragments_h, fragments_w, dur_t // aligned)
    hl, wl = res_h // fragments_h, res_w // fragments_w
    rnd_h = torch.randint(hl - fsize_h, n, generator=generator) if hl > fsize_h else torch.zeros(n, dtype=torch.int64)
    rnd_w = torch.randint(wl - fsize_w, n, generator=generator) if wl > fsize_w else torch.zeros(n, dtype=torch.int64)
    return rnd_h, rnd_w


def spatial_fragments(video, rnd_h, rnd_w, fragments_h=7, fragments_w=7, fsize_h=32, fsize_w=32, aligned=32):
    """video [C,T,H,W] (0..255 valued) -> [C,T,fh*fs,fw*fs]  (fusion_datasets.py:34-50, :60-69, :91-117)."""
    size_h, size_w = fragments_h * fsize_h, fragments_w * fsize_w
    if video.shape[1] == 1:
        aligned = 1
    dur_t, res_h, res_w = video.shape[-3:]
    ratio = min(res_h / size_h, res_w / size_w)
    if ratio < 1:                                                       # fallback_type == "upsample" (:43-50)
        ovideo = video
        video = F.interpolate(video / 255.0, scale_factor=1 / ratio, mode="bilinear")
        video = (video * 255.0).type_as(ovideo)
    assert dur_t % aligned == 0
    # NOTE (:64-69, :71): grids and cell lengths come from the ORIGINAL resolution even after the upsample
    hgrids = [min(res_h // fragments_h * i, res_h - fsize_h) for i in range(fragments_h)]
    wgrids = [min(res_w // fragments_w * i, res_w - fsize_w) for i in range(fragments_w)]
    out = torch.zeros(video.shape[:-2] + (size_h, size_w))
    for i, hs in enumerate(hgrids):
        for j, ws in enumerate(wgrids):
            for t in range(dur_t // aligned):
                ho, wo = hs + int(rnd_h[i][j][t]), ws + int(rnd_w[i][j][t])
                out[:, t * aligned:(t + 1) * aligned, i * fsize_h:(i + 1) * fsize_h, j * fsize_w:(j + 1) * fsize_w] = \
                    video[:, t * aligned:(t + 1) * aligned, ho:ho + fsize_h, wo:wo + fsize_w]
    return out


def normalise(clip):
    """[C,T,H,W] -> (v - mean) / std (fusion_datasets.py:1017-1020)."""
    mean = torch.tensor(MEAN).view(3, 1, 1, 1)
    std = torch.tensor(STD).view(3, 1, 1, 1)
    return (clip - mean) / std


def fragment_clip(frames_u8, offsets, fragments_h=7, fragments_w=7, fsize=32, aligned=8):
    """The C-ABI contract of kvq_fragment_gather_u8: frames u8 [B,T,3,H,W], offsets [B,2,fh,fw,T//aligned] (h then w)
    -> normalised f32 [B,3,T,fh*fs,fw*fs]."""
    out = []
    for b in range(frames_u8.shape[0]):
        video = frames_u8[b].permute(1, 0, 2, 3).float()
        out.append(normalise(spatial_fragments(video, offsets[b, 0], offsets[b, 1], fragments_h, fragments_w, fsize,
                                               fsize, aligned)))
    return torch.stack(out)
