"""Device time of kvq_resize_view_u8 on the KSVQE geometry (uint8 [B,32,3,1080,1920] -> 112x112 CLIP view) for both W-axis
kernels and several tile heights (KVQ_VIEWS_VARIANT / KVQ_VIEWS_ROWS), CUDA events, 20 calls after 3 warm-ups.
    python tools/views_timing.py [clips]      -> one JSON line per configuration"""
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
    need = ops._l.load().kvq_resize_view_workspace_bytes(B, 32, 1080, 1920, 112, 112, 0, 0, 0, 0)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    alg = B * 96 * (1080 * 1920 + 112 * 112 * 4)
    ref = None
    for variant in ("2", "3", "4"):
        for rows in ("4", "8", "16", "32"):
            os.environ["KVQ_VIEWS_VARIANT"], os.environ["KVQ_VIEWS_ROWS"] = variant, rows

            def call():
                return ops.resize_view_u8(frames, 112, 112, mean=ops.CLIP_MEAN, std=ops.CLIP_STD, divisor=255.0,
                                          workspace=ws)[1]
            for _ in range(3):
                out = call()
            if ref is None:
                ref = out.clone()
            same = bool(torch.equal(ref, out))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(json.dumps({"variant": int(variant), "rows_per_cta": int(rows), "ms_per_call": ms,
                              "algorithmic_GBps": alg / ms / 1e6, "clips": B, "identical_to_first": same}), flush=True)


if __name__ == "__main__":
    main()
