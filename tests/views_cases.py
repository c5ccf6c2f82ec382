"""Shared body of the datasets/views.py checks against tests/golden/views_geometry.npz (the REAL reference's
get_resized_video(arp=True) and get_resizecrop_video(phase='train') outputs, tools/make_golden_views.py).  The CPU suite
runs it with an oracle-backed stand-in for ops.resize_view_u8 (host logic: sizes, crop windows, draw order, layouts); the
GPU suite runs it through the C ABI."""
import os
import random

import numpy as np
import torch

from conftest import GOLDEN


def golden_video(T, H, W, seed):
    gen = torch.Generator().manual_seed(int(seed))
    return torch.randint(0, 256, (int(T), int(H), int(W), 3), generator=gen, dtype=torch.uint8).permute(3, 0, 1, 2)


def check_geometry(V, device):
    g = np.load(os.path.join(GOLDEN, "views_geometry.npz"))
    for i, (T, H, W, sh, sw, seed) in enumerate(g["arp_cases"]):
        video = golden_video(T, H, W, seed).contiguous().to(device)
        out = V.get_resized_video(video, size_h=int(sh), size_w=int(sw), arp=True)
        assert out.dtype == torch.uint8 and tuple(out.shape) == g[f"arp_out_{i}"].shape
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"arp_out_{i}"])
        batched = V.get_resized_video(torch.stack([video, video.flip(1)]), size_h=int(sh), size_w=int(sw), arp=True)
        np.testing.assert_array_equal(batched[0].cpu().numpy(), g[f"arp_out_{i}"])      # [B,3,T,H,W] input
        np.testing.assert_array_equal(batched[1].flip(1).cpu().numpy(), g[f"arp_out_{i}"])
    for i, (T, H, W, rs, cr, seed, rseed) in enumerate(g["train_cases"]):
        video = golden_video(T, H, W, seed).contiguous().to(device)
        random.seed(int(rseed))
        out = V.get_resizecrop_video(video, resize=int(rs), crop=int(cr), phase="train")
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"train_out_{i}"])
        assert random.random() == _after_two_draws(int(rseed), int(rs) - int(cr))       # exactly two draws consumed


def _after_two_draws(seed, n):
    random.seed(seed)
    random.randrange(n)
    random.randrange(n)
    return random.random()
