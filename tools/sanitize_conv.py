"""compute-sanitizer target for the convolution branches: one SlowFast clip (implicit-GEMM stems with the TMEM A ring,
64-channel and narrow-channel implicit convolutions, row-folded pointwise layers, TMA-store epilogue, pools) and a
small SimpleVQA forward.
    compute-sanitizer --tool memcheck  python tools/sanitize_conv.py
    compute-sanitizer --tool racecheck python tools/sanitize_conv.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
import torch  # noqa: E402
from kvq_b200 import ops  # noqa: E402
from tools import synth  # noqa: E402

dev = torch.device("cuda:0")
w = ops.SlowFastWeights(synth.slowfast_state_dict(1), dev)
fast = synth.slowfast_frames((1, 3, 32, 224, 224), 2).to(dev)
slow = ops.pack_pathway_slow(fast)
s, f = w.forward(slow, fast)
torch.cuda.synchronize()
sv = ops.SimpleVQAWeights(synth.simplevqa_network_state_dict(1), dev, prefix="simpleVQA_backbone.",
                          head_prefix="simpleVQA_head.")
x = synth.clip_input((1, 3, 2, 100, 76), 3).to(dev)
feats, score = sv.forward(x, synth.motion_features((1, 2, 2304), 4).to(dev))
torch.cuda.synchronize()
print("sanitize_conv ok", float(s.abs().mean()), float(f.abs().mean()), float(score[0]))
