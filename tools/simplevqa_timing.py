"""Device timing of the SimpleVQA path (kwai_simpleVQA_test.yml geometry: 8 frames of 448x448 per clip) with the
per-category CUDA-event breakdown.  Usage: python tools/simplevqa_timing.py [batch] [iters]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
from kvq_b200 import lib, ops  # noqa: E402
from tools import synth  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda:0")
    sd = synth.simplevqa_network_state_dict(1)
    w = ops.SimpleVQAWeights(sd, dev, prefix="simpleVQA_backbone.", head_prefix="simpleVQA_head.")
    x = synth.clip_input((B, 3, 8, 448, 448), 2).to(dev)
    f3 = synth.motion_features((B, 8, 2304), 3).to(dev)
    for graph in (False, True):
        for _ in range(3):
            w.forward(x, f3, graph=graph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            w.forward(x, f3, graph=graph)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"batch {B} graph={graph}: {ms:.3f} ms/step  {B / ms * 1e3:.1f} clips/s  {B * 8 / ms * 1e3:.0f} frames/s")
    L = lib.load()
    L.kvq_profile_enable(1)
    w.forward(x, f3, graph=False)
    n = L.kvq_profile_num_categories()
    ms = (ctypes.c_float * n)()
    cnt = (ctypes.c_int * n)()
    L.kvq_profile_collect(ms, cnt, n)
    L.kvq_profile_enable(0)
    tot = sum(ms)
    for i in range(n):
        if cnt[i]:
            print(f"  {L.kvq_profile_category_name(i).decode():18s} {cnt[i]:4d} launches {ms[i]:8.3f} ms  {100 * ms[i] / tot:5.1f}%")
    print(f"  total {tot:.3f} ms")


if __name__ == "__main__":
    main()
