"""Run one GEMM shape a few times (for ncu): python tools/gemm_one.py M N K mode(store|gelu|resid|conv)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
from kvq_b200 import ops
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "store"
dev = "cuda:0"
a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev).half(); b = torch.zeros(N, device=dev)
x = torch.randn(M, N, device=dev) if mode == "resid" else None
out = torch.empty(M, N, device=dev, dtype=torch.float16)
for _ in range(3):
    if mode == "resid": ops.linear_resid_f32(a, w, b, resid=x, out=x)
    elif mode == "conv": ops.conv_gemm_f16(a, w, b, None, relu=True, out=out)
    else: ops.linear_f16(a, w, b, gelu=(mode == "gelu"))
torch.cuda.synchronize()
