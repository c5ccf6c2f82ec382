"""Device time of the two Resize flavours on the KSVQE geometry (uint8 [B,32,3,1080,1920] -> 112x112 CLIP view):
kvq_resize_view_u8 (anti-aliased, torchvision >= 0.17) and kvq_resize_view_bilinear_u8 (plain bilinear, torchvision < 0.17).
CUDA events, 20 calls after 3 warm-ups; algorithmic bytes = source planes read once + float32 view written once.
    python tools/views_noaa_timing.py [clips]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
from kvq_b200 import ops  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    frames = torch.randint(0, 256, (B, 32, 3, 1080, 1920), generator=g, dtype=torch.uint8, device=dev)
    alg = B * 96 * (1080 * 1920 + 112 * 112 * 4)
    for aa in (True, False):
        def call():
            return ops.resize_view_u8(frames, 112, 112, mean=ops.CLIP_MEAN, std=ops.CLIP_STD, divisor=255.0, antialias=aa)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(json.dumps({"antialias": aa, "clips": B, "ms": round(ms, 4), "algorithmic_GBps": round(alg / ms / 1e6, 1),
                          "note": "the plain bilinear filter touches 2x2 source pixels per output: it reads ~12 % of the frame"}))


if __name__ == "__main__":
    main()
