"""Debug: per-phase cycle breakdown of the fast attention kernel (needs tools/libkvq_b200_timing.so built with
-DKVQ_TIMING).  Usage on the GPU box: python tools/attn_timing.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
from kvq_b200 import lib  # noqa: E402

lib.LIB_PATH = os.path.join(ROOT, "tools", "libkvq_b200_timing.so")
import torch  # noqa: E402
from kvq_b200 import ops  # noqa: E402
from tools import synth  # noqa: E402

NAMES = ["table wait", "wait S(t>0)", "pass 1", "max barrier", "wait PV + O epilogue", "pass 2", "wait S(0) incl. load", "prologue/arrive"]


def main():
    dev = torch.device("cuda:0")
    L = lib.load()
    window = (8, 7, 7)
    for (B, D, H, W, C, heads, shift) in [(8, 16, 56, 56, 96, 3, (0, 0, 0)), (8, 16, 56, 56, 96, 3, (4, 3, 3)),
                                          (8, 16, 14, 14, 384, 12, (4, 3, 3))]:
        sd = synth.synth_state_dict({"attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
                                     "attn.relative_position_bias_table": (2535, heads),
                                     "attn.fragment_position_bias_table": (2535, heads)}, 5)
        rows = ops.window_rows(B, D, H, W, window, shift)
        xw = torch.randn(rows, C, device=dev).half()
        tab = ops.pack_bias_table(sd["attn.relative_position_bias_table"].to(dev),
                                  sd["attn.fragment_position_bias_table"].to(dev), window, heads)
        w = ops.cast_f16(sd["attn.qkv.weight"].to(dev))
        b = sd["attn.qkv.bias"].to(dev)
        buf = (ctypes.c_ulonglong * 16)()
        for it in range(3):
            L.kvq_debug_attn_timers(buf, 1)
            ops.window_attention(xw, w, b, tab, B, D, H, W, heads, window, shift)
            torch.cuda.synchronize()
        assert L.kvq_debug_attn_timers(buf, 1) == 1, "not a -DKVQ_TIMING build"
        warps = buf[8]
        units = B * (D // 8) * (H // 7) * (W // 7) * heads
        tot = sum(buf[k] for k in range(8))
        print(f"geometry D{D} H{H} W{W} C{C} shift{shift}: units {units}, softmax warps {warps}")
        for k in range(8):
            print(f"  {NAMES[k]:24s} {buf[k] / warps / (units / (warps / 8)):10.0f} cycles/unit  {100.0 * buf[k] / tot:5.1f}%")
        print(f"  total {tot / warps / (units / (warps / 8)):10.0f} cycles/unit (per-warp average)")


if __name__ == "__main__":
    main()
