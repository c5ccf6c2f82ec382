"""One warm call, then one call of the view pipeline inside a cudaProfilerStart/Stop range (for ncu):
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:resize_ -c 4 \
        -o gpurun_out/views_full python tools/views_profile.py [clips]
KSVQE geometry: decoder-order uint8 frames [clips,32,3,1080,1920] -> 112x112 CLIP-normalised view + 9x9 fragments."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
from kvq_b200 import ops  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    frames = torch.randint(0, 256, (B, 32, 3, 1080, 1920), generator=g, dtype=torch.uint8, device=dev)
    offs = torch.stack([torch.randint(88, (B, 9, 9, 4), generator=g, device=dev),
                        torch.randint(181, (B, 9, 9, 4), generator=g, device=dev)], dim=1).to(torch.int32).contiguous()

    def step():
        for variant in ("4",):               # the default W-axis kernel (resize_rows_bytes_kernel)
            os.environ["KVQ_VIEWS_VARIANT"] = variant
            ops.resize_view_u8(frames, 112, 112, mean=ops.CLIP_MEAN, std=ops.CLIP_STD, divisor=255.0)
        ops.fragment_gather_u8(frames, offs, 9, 9, 32, 8)

    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
