"""Times the fused MLP kernel (kvq_mlp_fused) on the stage-0 / stage-1 token counts of a batch-8 32x224x224 clip.
Usage on the GPU box: [KVQ_LIB=tools/libkvq_var_x.so] python tools/mlp_bench.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
import torch  # noqa: E402
from kvq_b200 import lib, ops  # noqa: E402
if os.environ.get("KVQ_LIB"):          # A/B builds: point the loader at another libkvq_b200.so
    lib.LIB_PATH = os.environ["KVQ_LIB"]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    for tag, M, C in (("s0", 8 * 16 * 56 * 56, 96), ("s1", 8 * 16 * 28 * 28, 192)):
        a = torch.randn(M, C, generator=g).half().to(dev)
        w1 = (torch.randn(4 * C, C, generator=g) * C ** -0.5).half().to(dev)
        w2 = (torch.randn(C, 4 * C, generator=g) * (4 * C) ** -0.5).half().to(dev)
        b1 = torch.randn(4 * C, generator=g).to(dev) * 0.1
        b2 = torch.randn(C, generator=g).to(dev) * 0.1
        x = torch.randn(M, C, generator=g).to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        x0 = x.clone()
        ops.mlp_fused(a, w1, b1, w2, b2, x)
        torch.cuda.synchronize()
        ref = x0[:4096] + torch.nn.functional.gelu(a[:4096].float() @ w1.float().t() + b1).half().float() @ w2.float().t() + b2
        err = (x[:4096] - ref).abs().max().item()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.mlp_fused(a, w1, b1, w2, b2, x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(json.dumps({"stage": tag, "M": M, "C": C, "us_median": round(ts[len(ts) // 2], 1), "us_min": round(ts[0], 1),
                          "max_abs_err_vs_torch": err}))


if __name__ == "__main__":
    main()
