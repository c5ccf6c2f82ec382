"""Times the fused window-attention kernel generations on the four stage geometries of a 32x224x224 clip (batch 8),
attention kernel only (the library's per-category CUDA-event profile separates it from the QKV GEMM).
Usage on the GPU box: python tools/attn_bench.py [variants...]   (default: 5 6)"""
import ctypes
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
from tools import synth  # noqa: E402

GEOMS = [("s0", 8, 16, 56, 56, 96, 3), ("s1", 8, 16, 28, 28, 192, 6), ("s2", 8, 16, 14, 14, 384, 12),
         ("s3", 8, 16, 7, 7, 768, 24)]
if os.environ.get("ATTN_BENCH_BATCHES"):     # e.g. "s2:8,16,32": the same stage geometry at several batch sizes
    _tag, _bs = os.environ["ATTN_BENCH_BATCHES"].split(":")
    _g = next(g for g in GEOMS if g[0] == _tag)
    GEOMS = [(f"{_tag}_b{b}", int(b)) + _g[2:] for b in _bs.split(",")]


def main():
    variants = [int(v) for v in sys.argv[1:]] or [5, 6]
    dev = torch.device("cuda:0")
    L = lib.load()
    window = (8, 7, 7)
    ncat = L.kvq_profile_num_categories()
    names = [L.kvq_profile_category_name(i).decode() for i in range(ncat)]
    out = []
    for (tag, B, D, H, W, C, heads) in GEOMS:
        sd = synth.synth_state_dict({"attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
                                     "attn.relative_position_bias_table": (2535, heads),
                                     "attn.fragment_position_bias_table": (2535, heads)}, 5)
        tab = ops.pack_bias_table(sd["attn.relative_position_bias_table"].to(dev),
                                  sd["attn.fragment_position_bias_table"].to(dev), window, heads)
        w = ops.cast_f16(sd["attn.qkv.weight"].to(dev))
        b = sd["attn.qkv.bias"].to(dev)
        for shift in [(0, 0, 0), (4, 3, 3)]:
            rows = ops.window_rows(B, D, H, W, window, shift)
            xw = torch.randn(rows, C, device=dev).half()
            ref = None
            for v in variants:
                for _ in range(3):
                    o = ops.window_attention(xw, w, b, tab, B, D, H, W, heads, window, shift, debug_variant=v)
                torch.cuda.synchronize()
                L.kvq_profile_enable(1)
                n = 10
                for _ in range(n):
                    o = ops.window_attention(xw, w, b, tab, B, D, H, W, heads, window, shift, debug_variant=v)
                torch.cuda.synchronize()
                ms = (ctypes.c_float * ncat)()
                cnt = (ctypes.c_int * ncat)()
                L.kvq_profile_collect(ms, cnt, ncat)
                L.kvq_profile_enable(0)
                t = {names[i]: ms[i] / max(cnt[i], 1) for i in range(ncat) if cnt[i] > 0}
                att = [val for key, val in t.items() if key.startswith("window_attn")][0]
                o32 = o.float()
                diff = None
                if ref is None:
                    ref = o32
                else:
                    diff = (o32 - ref).abs().max().item()
                units = B * (D // 8) * -(-H // 7) * -(-W // 7) * heads
                flops = units * 2 * 2 * 392 * 392 * 32
                rec = {"geom": tag, "shift": shift, "variant": v, "attn_us": round(att * 1e3, 1),
                       "tflops": round(flops / (att * 1e-3) / 1e12, 1), "finite": bool(torch.isfinite(o32).all()),
                       "maxdiff_vs_first": diff}
                print(json.dumps(rec), flush=True)
                out.append(rec)
    return out


if __name__ == "__main__":
    main()
