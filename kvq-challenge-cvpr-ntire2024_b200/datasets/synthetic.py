"""Synthetic stand-in for ViewDecompositionDataset_KVQ (reference datasets/fusion_datasets.py:930-1050).

Each item mimics what the reference loader yields for the `technical` view: raw uint8-valued frames are sampled
into a fragment grid (Grid Mini-patch Sampling, :22-121) and normalised (:1017-1020); `num_clips` clips of
`clip_len` frames are concatenated along T.  With `device='cuda'` the gather + normalisation run in the sm_100a
fragment kernel; with the default `device=None` the item carries the raw frames and offsets and the Trainer runs the
kernel after the H2D copy (the DataLoader workers stay CPU-only)."""
import torch


class SyntheticFragmentDataset(torch.utils.data.Dataset):
    def __init__(self, opt, _unused=None):
        self.opt = opt
        st = opt["sample_types"]["technical"]
        self.fh, self.fw = st.get("fragments_h", 7), st.get("fragments_w", 7)
        self.fs = st.get("fsize_h", 32)
        self.aligned = st.get("aligned", 8)
        self.clip_len, self.num_clips = st.get("clip_len", 32), st.get("num_clips", 1)
        self.src_h, self.src_w = opt.get("src_h", 448), opt.get("src_w", 448)
        self.n = opt.get("num_videos", 4)
        self.seed = opt.get("seed", 5)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        T = self.clip_len * self.num_clips
        frames = torch.randint(0, 256, (T, 3, self.src_h, self.src_w), generator=g, dtype=torch.uint8)
        hl, wl = self.src_h // self.fh, self.src_w // self.fw
        nt = T // self.aligned
        # reference draw order: rnd_h then rnd_w (fusion_datasets.py:87-98)
        rnd_h = torch.randint(hl - self.fs, (self.fh, self.fw, nt), generator=g) if hl > self.fs else \
            torch.zeros(self.fh, self.fw, nt, dtype=torch.int64)
        rnd_w = torch.randint(wl - self.fs, (self.fh, self.fw, nt), generator=g) if wl > self.fs else \
            torch.zeros(self.fh, self.fw, nt, dtype=torch.int64)
        return {"frames": frames, "offsets": torch.stack([rnd_h, rnd_w]).int(),
                "num_clips": {"technical": self.num_clips}, "video_name": f"synthetic_{i:04d}",
                "label": torch.rand((), generator=g) * 4.0 + 1.0,        # stand-in MOS in [1, 5] (inferece_val)
                "fragment_opts": {"fragments_h": self.fh, "fragments_w": self.fw, "fsize": self.fs,
                                  "aligned": self.aligned}}


class SyntheticSimpleVQADataset(torch.utils.data.Dataset):
    """Stand-in for ViewDecompositionDataset_add_forSimpleVQA (datasets/fusion_datasets.py:780-930): `clip_len` frames
    resized / centre-cropped to `crop` and ImageNet-normalised (`simpleVQA` [3,T,crop,crop]) plus the pre-extracted
    SlowFast features of the video (`feat` [T, 2304], see datasets.load_motion_features)."""

    def __init__(self, opt, _unused=None):
        st = opt["sample_types"]["simpleVQA"]
        self.crop, self.clip_len, self.num_clips = st.get("crop", 448), st.get("clip_len", 8), st.get("num_clips", 1)
        self.n, self.seed = opt.get("num_videos", 4), opt.get("seed", 7)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        T = self.clip_len * self.num_clips
        return {"simpleVQA": torch.randn((3, T, self.crop, self.crop), generator=g),
                "feat": torch.randn((T, 2304), generator=g).abs(),
                "num_clips": {"simpleVQA": self.num_clips}, "video_name": f"synthetic_{i:04d}",
                "label": torch.rand((), generator=g) * 4.0 + 1.0}


class SyntheticKSVQEDataset(torch.utils.data.Dataset):
    """Stand-in for ViewDecompositionDataset_KVQ in the KSVQE configuration (datasets/fusion_datasets.py:930-1050,
    config/Kwai_KSVQE_test.yml): `fragment` = num_clips * clip_len frames of a fragments_h x fragments_w grid of fsize
    patches, ImageNet-normalised; `resize_video` = the same frames resized to size_h x size_w, CLIP-normalised;
    `dis_label` = the distortion class of the video.  (As in the reference, the clips of a video stay concatenated on
    the T axis: trainer.py never splits them for this key.)"""

    def __init__(self, opt, _unused=None):
        st = opt["sample_types"]["technical"]
        self.fh, self.fw, self.fs = st.get("fragments_h", 9), st.get("fragments_w", 9), st.get("fsize_h", 32)
        self.rh, self.rw = st.get("size_h", 112), st.get("size_w", 112)
        self.T = st.get("clip_len", 32) * st.get("num_clips", 3)
        self.num_clips = st.get("num_clips", 3)
        self.n, self.seed = opt.get("num_videos", 4), opt.get("seed", 11)
        self.aligned = st.get("aligned", 8)
        # raw_frames: the item carries decoded uint8 frames [T,3,src_h,src_w] and the fragment offsets; the Trainer
        # builds `fragment` and `resize_video` with the view kernels after the H2D copy
        self.raw = bool(opt.get("raw_frames", False))
        self.src_h, self.src_w = opt.get("src_h", 540), opt.get("src_w", 960)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        common = {"num_clips": {"technical": self.num_clips}, "video_name": f"synthetic_{i:04d}"}
        if self.raw:
            frames = torch.randint(0, 256, (self.T, 3, self.src_h, self.src_w), generator=g, dtype=torch.uint8)
            hl, wl, nt = self.src_h // self.fh, self.src_w // self.fw, self.T // self.aligned
            # reference draw order: rnd_h then rnd_w (fusion_datasets.py:87-98)
            # a grid cell exactly fsize wide leaves no room to jitter: zero offsets (fusion_datasets.py:88-98)
            rnd_h = torch.randint(hl - self.fs, (self.fh, self.fw, nt), generator=g) if hl > self.fs else \
                torch.zeros((self.fh, self.fw, nt), dtype=torch.int64)
            rnd_w = torch.randint(wl - self.fs, (self.fh, self.fw, nt), generator=g) if wl > self.fs else \
                torch.zeros((self.fh, self.fw, nt), dtype=torch.int64)
            return dict(common, frames=frames, offsets=torch.stack([rnd_h, rnd_w]).int(),
                        fragment_opts={"fragments_h": self.fh, "fragments_w": self.fw, "fsize": self.fs,
                                       "aligned": self.aligned},
                        resize_opts={"size_h": self.rh, "size_w": self.rw},
                        dis_label=torch.tensor(int(torch.randint(0, 5, (1,), generator=g))),
                        label=torch.rand((), generator=g) * 4.0 + 1.0)
        frag = torch.randn((3, self.T, self.fh * self.fs, self.fw * self.fs), generator=g)
        return dict(common, technical=frag, fragment=frag,
                    resize_video=torch.randn((3, self.T, self.rh, self.rw), generator=g),
                    dis_label=torch.tensor(int(torch.randint(0, 5, (1,), generator=g))),
                    label=torch.rand((), generator=g) * 4.0 + 1.0)
