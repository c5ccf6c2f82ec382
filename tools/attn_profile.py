"""Three launches of the window-attention op on one stage geometry, for ncu captures.
Usage: python tools/attn_profile.py <variant> [stage 0-3] [shifted 0/1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
import torch  # noqa: E402
from kvq_b200 import lib, ops  # noqa: E402
if os.environ.get("KVQ_LIB"):
    lib.LIB_PATH = os.environ["KVQ_LIB"]
from tools import synth  # noqa: E402
from tools.attn_bench import GEOMS  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 6
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 0
shift = (4, 3, 3) if len(sys.argv) > 3 and int(sys.argv[3]) else (0, 0, 0)
tag, B, D, H, W, C, heads = GEOMS[stage]
dev = torch.device("cuda:0")
window = (8, 7, 7)
sd = synth.synth_state_dict({"attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
                             "attn.relative_position_bias_table": (2535, heads),
                             "attn.fragment_position_bias_table": (2535, heads)}, 5)
tab = ops.pack_bias_table(sd["attn.relative_position_bias_table"].to(dev),
                          sd["attn.fragment_position_bias_table"].to(dev), window, heads)
w = ops.cast_f16(sd["attn.qkv.weight"].to(dev))
b = sd["attn.qkv.bias"].to(dev)
xw = torch.randn(ops.window_rows(B, D, H, W, window, shift), C, device=dev).half()
for _ in range(3):
    o = ops.window_attention(xw, w, b, tab, B, D, H, W, heads, window, shift, debug_variant=variant)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
