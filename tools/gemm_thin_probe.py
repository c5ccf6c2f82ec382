"""Probe: conv GEMM time vs K for a very tall problem (per-tile fixed cost)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
from kvq_b200 import ops

dev = "cuda:0"
M = 1 << 21
for N, nv in ((64, 64), (64, 32), (64, 8), (256, 256)):
    for K in (8, 24, 64, 72, 128, 256, 576):
        for resid in (False, True):
            if N == 256 and K > 128: continue
            a = torch.randn(M, K, device=dev).half()
            w = torch.randn(N, K, device=dev).half()
            b = torch.zeros(N, device=dev)
            r = torch.randn(M, nv, device=dev).half() if resid else None
            out = torch.empty(M, nv, device=dev, dtype=torch.float16)
            for _ in range(2):
                ops.conv_gemm_f16(a, w, b, r, relu=True, nvalid=0 if nv == N else nv, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv_gemm_f16(a, w, b, r, relu=True, nvalid=0 if nv == N else nv, out=out)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 5 * 1e3
            byts = M * (K + nv * (2 if resid else 1)) * 2
            print(f"N={N:3d} nvalid={nv:3d} K={K:3d} resid={int(resid)}: {us:7.1f} us  {us / (M / 128 / 148):6.3f} us/tile/SM  {byts / us / 1e6:6.2f} TB/s")
