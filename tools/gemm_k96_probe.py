"""Probe: is a K=96 GEMM (second 64-wide TMA box half out of bounds) slower than K=128 on the same rows?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
from kvq_b200 import ops
dev = "cuda:0"
M = 401408
for N in (96, 288, 384):
    for K in (64, 96, 128, 192):
        a = torch.randn(M, K, device=dev).half()
        w = torch.randn(N, K, device=dev).half()
        b = torch.zeros(N, device=dev)
        x = torch.randn(M, N, device=dev)
        for mode in ("store_f16", "resid_f32"):
            if mode == "resid_f32" and N != 96: continue
            f = (lambda: ops.linear_f16(a, w, b)) if mode == "store_f16" else (lambda: ops.linear_resid_f32(a, w, b, resid=x, out=x))
            for _ in range(2): f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): f()
            e1.record(); torch.cuda.synchronize()
            print(f"{mode} M={M} N={N:3d} K={K:3d}: {e0.elapsed_time(e1) / 5 * 1e3:7.1f} us")
