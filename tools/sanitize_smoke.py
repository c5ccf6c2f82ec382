"""compute-sanitizer target: one small forward that exercises every kernel family of the default path (generic + third-
generation attention, patch-embed / LN1+qkv / proj+norm2 / fused-MLP kernels, the GEMM epilogues, LN / gather kernels,
fragment gather), with fp32 and fp16 clips.
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
import torch  # noqa: E402
from kvq_b200 import ops  # noqa: E402
from tools import synth  # noqa: E402

dev = torch.device("cuda:0")
sd = synth.swin_network_state_dict(3)
wts = ops.SwinWeights(sd, dev, prefix="swin_tiny_grpb_backbone.", head_prefix="swin_tiny_grpb_head.")
# 16 x 112 x 112: stage 0/1 run full (8,7,7) windows (fast kernels), stages 2/3 clamp the window (generic kernel)
x = synth.clip_input((1, 3, 16, 112, 112), 4).to(dev)
# default kernel paths (third-generation attention, patch_embed / ln_qkv / proj_ln / fused MLP kernels); the older paths are
# selected by environment, read once per process: run the tool again under KVQ_ATTN_VARIANT=5, KVQ_EMBED_EXPLICIT=1, ...
feat, score = wts.forward(x, want_feat=True)
x16 = x.half()
feat16, score16 = wts.forward(x16, want_feat=True)
torch.cuda.synchronize()
frames = torch.randint(0, 256, (1, 8, 3, 80, 72), dtype=torch.uint8, device=dev)
offs = torch.zeros((1, 2, 2, 2, 1), dtype=torch.int32, device=dev)
clips = ops.fragment_gather_u8(frames, offs, 2, 2, 32, 8)
torch.cuda.synchronize()
print("sanitize_smoke ok", float(score[0]), tuple(feat.shape), tuple(clips.shape))
