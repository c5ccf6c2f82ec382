"""Drop-in for the reference SlowFast_features.py: `pack_pathway_output`, the `slowfast` module (same module tree and
state_dict names: `feature_extraction.<0..4>.` + pytorchvideo's slowfast_r50 names) and `main` writing
`feature_<i>_{slow,fast}_feature.npy` per clip.

The modules are parameter containers; the arithmetic (both pathways, lateral fusions, head pools) runs in
libkvq_b200.so (kvq_slowfast_forward).  There is no CPU fallback.  pytorchvideo is absent offline, so the pretrained
Kinetics weights the reference downloads (SlowFast_features.py:140) must come in through `load_state_dict`; parity of
the architecture restatement is unpinned (oracle/slowfast.py header)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from kvq_b200 import ops  # noqa: E402

DEPTHS = (3, 4, 6, 3)
ALPHA = 4


def pack_pathway_output(frames, device):
    """frames [B,3,T,H,W] -> [slow [B,3,T//4,H,W], fast [B,3,T,H,W]] on `device` (SlowFast_features.py:112-135).
    The slow pathway is gathered on the device (kvq_pack_pathway_slow_f32) after ONE host->device copy of the clip;
    the reference index_selects on the host and copies both pathways."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("kvq_b200: pack_pathway_output targets a CUDA device (no CPU fallback exists)")
    fast = frames.to(device, torch.float32, non_blocking=True).contiguous()
    with torch.cuda.device(device):
        slow = ops.pack_pathway_slow(fast, ALPHA)
    return [slow, fast]


def _conv(cin, cout, k, s=(1, 1, 1), p=(0, 0, 0)):
    return nn.Conv3d(cin, cout, k, s, p, bias=False)


class _Stem(nn.Module):                      # pytorchvideo ResNetBasicStem
    def __init__(self, cout, kt):
        super().__init__()
        self.conv = _conv(3, cout, (kt, 7, 7), (1, 2, 2), (kt // 2, 3, 3))
        self.norm = nn.BatchNorm3d(cout)
        self.activation = nn.ReLU()
        self.pool = nn.MaxPool3d((1, 3, 3), (1, 2, 2), (0, 1, 1))


class _Fuse(nn.Module):                      # pytorchvideo FuseFastToSlow
    def __init__(self, cf):
        super().__init__()
        self.conv_fast_to_slow = _conv(cf, 2 * cf, (7, 1, 1), (ALPHA, 1, 1), (3, 0, 0))
        self.norm = nn.BatchNorm3d(2 * cf)
        self.activation = nn.ReLU()


class _Bottleneck(nn.Module):                # pytorchvideo BottleneckBlock
    def __init__(self, cin, inner, ka, stride):
        super().__init__()
        self.conv_a = _conv(cin, inner, (ka, 1, 1), p=(ka // 2, 0, 0))
        self.norm_a = nn.BatchNorm3d(inner)
        self.act_a = nn.ReLU()
        self.conv_b = _conv(inner, inner, (1, 3, 3), (1, stride, stride), (0, 1, 1))
        self.norm_b = nn.BatchNorm3d(inner)
        self.act_b = nn.ReLU()
        self.conv_c = _conv(inner, 4 * inner, (1, 1, 1))
        self.norm_c = nn.BatchNorm3d(4 * inner)


class _ResBlock(nn.Module):                  # pytorchvideo ResBlock
    def __init__(self, cin, inner, ka, stride, first):
        super().__init__()
        if first:
            self.branch1_conv = _conv(cin, 4 * inner, (1, 1, 1), (1, stride, stride))
            self.branch1_norm = nn.BatchNorm3d(4 * inner)
        self.branch2 = _Bottleneck(cin, inner, ka, stride)
        self.activation = nn.ReLU()


class _ResStage(nn.Module):
    def __init__(self, cin, inner, ka, stride, depth):
        super().__init__()
        self.res_blocks = nn.ModuleList(
            [_ResBlock(cin if j == 0 else 4 * inner, inner, ka, stride if j == 0 else 1, j == 0) for j in range(depth)])


class _MultiPathWayWithFuse(nn.Module):
    def __init__(self, blocks, fusion):
        super().__init__()
        self.multipathway_blocks = nn.ModuleList(blocks)
        self.multipathway_fusion = fusion if fusion is not None else nn.Identity()


class slowfast(torch.nn.Module):
    """SlowFast_features.py:137-165.  forward([slow, fast]) -> (slow_feature [B,2048,1,1,1], fast_feature [B,256,1,1,1])."""

    def __init__(self):
        super().__init__()
        fe = [_MultiPathWayWithFuse([_Stem(64, 1), _Stem(8, 5)], _Fuse(8))]
        cs, cf = 80, 8
        for s, depth in enumerate(DEPTHS):
            inner_s, inner_f, stride = 64 << s, 8 << s, 1 if s == 0 else 2
            fe.append(_MultiPathWayWithFuse(
                [_ResStage(cs, inner_s, (1, 1, 3, 3)[s], stride, depth), _ResStage(cf, inner_f, 3, stride, depth)],
                _Fuse(4 * inner_f) if s < 3 else None))
            cs, cf = 4 * inner_s + (8 * inner_f if s < 3 else 0), 4 * inner_f
        self.feature_extraction = torch.nn.Sequential(*fe)
        self.slow_avg_pool = torch.nn.Sequential()
        self.fast_avg_pool = torch.nn.Sequential()
        self.adp_avg_pool = torch.nn.Sequential()
        self.slow_avg_pool.add_module("slow_avg_pool", nn.AvgPool3d((8, 7, 7), stride=(1, 1, 1)))
        self.fast_avg_pool.add_module("fast_avg_pool", nn.AvgPool3d((32, 7, 7), stride=(1, 1, 1)))
        self.adp_avg_pool.add_module("adp_avg_pool", nn.AdaptiveAvgPool3d(1))
        self.use_cuda_graph = None
        self._packed = None
        self._packed_key = None

    def packed(self):
        tensors = list(self.parameters()) + list(self.buffers())
        dev = tensors[0].device
        if dev.type != "cuda":
            raise RuntimeError("kvq_b200: slowfast runs on a CUDA device only (call .to('cuda')); no CPU fallback")
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._packed is None or self._packed_key != key:
            pool = lambda m: tuple(m.kernel_size) if isinstance(m.kernel_size, (tuple, list)) else (m.kernel_size,) * 3
            with torch.cuda.device(dev):
                self._packed = ops.SlowFastWeights(
                    dict(self.state_dict()), dev, depths=DEPTHS, alpha=ALPHA,
                    slow_pool=pool(self.slow_avg_pool.slow_avg_pool), fast_pool=pool(self.fast_avg_pool.fast_avg_pool),
                    eps=self.feature_extraction[0].multipathway_blocks[0].norm.eps)
            self._packed_key = key
        return self._packed

    def forward(self, x):
        if self.training:
            raise RuntimeError("kvq_b200: inference path only -- call model.eval() (BatchNorm is folded)")
        slow, fast = x
        if not (slow.is_cuda and fast.is_cuda):
            raise RuntimeError("kvq_b200: pathway tensors must be CUDA tensors -- this path has no CPU fallback")
        with torch.no_grad(), torch.cuda.device(slow.device):
            s, f = self.packed().forward(slow, fast, graph=self.use_cuda_graph)
        return s.view(-1, 2048, 1, 1, 1), f.view(-1, 256, 1, 1, 1)


def save_clip_features(folder, slow_feature, fast_feature, first_index=0):
    """On-disk format consumed by datasets/fusion_datasets.py:883-890: one pair of .npy per clip, shapes
    [1,2048,1,1,1] / [1,256,1,1,1] (SlowFast_features.py:196-197)."""
    os.makedirs(folder, exist_ok=True)
    s, f = slow_feature.to("cpu").numpy(), fast_feature.to("cpu").numpy()
    for i in range(s.shape[0]):
        np.save(os.path.join(folder, f"feature_{first_index + i}_slow_feature"), s[i:i + 1])
        np.save(os.path.join(folder, f"feature_{first_index + i}_fast_feature"), f[i:i + 1])


def video_clips(frames, frame_rate, clip_len=32, min_clips=8):
    """The reference's clip schedule (SlowFast_features.py:64-105): one 32-frame clip starting at each whole second,
    the tail padded by repeating the last frame, at least 8 clips (the last one repeated).  frames [L,3,H,W]."""
    L = frames.shape[0]
    n = 10 if frame_rate == 0 else int(L / frame_rate)
    clips = []
    for i in range(n):
        s = i * frame_rate
        if s + clip_len <= L:
            clips.append(frames[s:s + clip_len])
        else:
            c = torch.zeros((clip_len,) + tuple(frames.shape[1:]), dtype=frames.dtype)
            k = L - s
            c[:k] = frames[s:]
            c[k:] = c[k - 1]
            clips.append(c)
    while len(clips) < min_clips:
        clips.append(clips[n - 1])
    return clips


class VideoDataset_NR_SlowFast_feature(torch.utils.data.Dataset):
    """SlowFast_features.py:25-107 (decode with OpenCV on the host; out of the accelerated path)."""

    def __init__(self, args, transform, video_root, videos_csv):
        super().__init__()
        import csv
        self.resize, self.transform, self.args = args.resize, transform, args
        self.video_root, self.videos_csv = video_root, videos_csv
        with open(videos_csv, newline="") as f:
            rows = csv.reader(f)
            next(rows)
            self.video_infos = [row[0] for row in rows]

    def __len__(self):
        return len(self.video_infos)

    def __getitem__(self, index):
        import cv2
        from PIL import Image
        name = self.video_infos[index]
        cap = cv2.VideoCapture(os.path.join(self.video_root, name))
        length = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        rate = int(round(cap.get(cv2.CAP_PROP_FPS)))
        frames = torch.zeros([length, 3, self.resize, self.resize])
        read = 0
        for _ in range(length):
            ok, frame = cap.read()
            if ok:
                frames[read] = self.transform(Image.fromarray(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)))
                read += 1
        if 0 < read < length:
            frames[read:] = frames[read - 1]
        cap.release()
        return video_clips(frames, rate), name


def main(config, video_root, videos_dict):
    """SlowFast_features.py:167-197, with all clips of a video batched into one forward."""
    from torchvision import transforms
    device = torch.device("cuda")
    model = slowfast().to(device)
    if getattr(config, "weights", None):
        sd = torch.load(config.weights, map_location="cpu")
        model.load_state_dict(sd.get("state_dict", sd), strict=True)
    tf = transforms.Compose([transforms.Resize([config.resize, config.resize]), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.45, 0.45, 0.45], std=[0.225, 0.225, 0.225])])
    data = VideoDataset_NR_SlowFast_feature(config, tf, video_root, videos_dict)
    loader = torch.utils.data.DataLoader(data, batch_size=1, shuffle=False, num_workers=config.num_workers)
    folder = config.feature_save_folder + "/" + config.database + "/"
    model.eval()
    with torch.no_grad():
        for video, video_name in loader:
            video_name = video_name[0]
            print(video_name)
            clips = torch.cat(video, dim=0).permute(0, 2, 1, 3, 4)           # [n,3,32,H,W]
            slow_feature, fast_feature = model(pack_pathway_output(clips, device))
            save_clip_features(folder + video_name, slow_feature, fast_feature)


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--num_workers", type=int, default=6)
    parser.add_argument("--resize", type=int, default=224)
    parser.add_argument("--gpu_ids", type=list, default=None)
    parser.add_argument("--database", type=str, default="kvq")
    parser.add_argument("--video_root", type=str, default=None)
    parser.add_argument("--video_csv", type=str, default=None)
    parser.add_argument("--feature_save_folder", type=str, default="./feature/simpleVQA/")
    parser.add_argument("--weights", type=str, default=None,
                        help="state_dict of the slowfast module (the reference downloads it from the pytorchvideo hub)")
    config = parser.parse_args()
    main(config, config.video_root, config.video_csv)
