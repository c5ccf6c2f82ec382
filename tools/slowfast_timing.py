"""Device timing of the SlowFast feature extractor (BASELINE config 3: batch 16 clips 32x3x256x256) with the
per-category CUDA-event breakdown.  Usage: python tools/slowfast_timing.py [batch] [size] [iters]"""
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda:0")
    w = ops.SlowFastWeights(synth.slowfast_state_dict(1), dev)
    fast = synth.slowfast_frames((B, 3, 32, R, R), 2).to(dev)
    slow = ops.pack_pathway_slow(fast)
    for graph in (False, True):
        for _ in range(2):
            w.forward(slow, fast, graph=graph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            w.forward(slow, fast, graph=graph)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"batch {B} @{R} graph={graph}: {ms:.3f} ms/step  {B / ms * 1e3:.1f} clips/s")
    L = lib.load()
    L.kvq_profile_enable(1)
    w.forward(slow, fast, graph=False)
    n = L.kvq_profile_num_categories()
    ms = (ctypes.c_float * n)()
    cnt = (ctypes.c_int * n)()
    L.kvq_profile_collect(ms, cnt, n)
    L.kvq_profile_enable(0)
    tot = sum(ms)
    for i in range(n):
        if cnt[i]:
            print(f"  {L.kvq_profile_category_name(i).decode():18s} {cnt[i]:4d} launches {ms[i]:8.3f} ms  {100 * ms[i] / tot:5.1f}%")
    print(f"  total {tot:.3f} ms   workspace {w._ws.numel() / 2**30:.2f} GiB")


if __name__ == "__main__":
    main()
