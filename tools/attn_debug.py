"""Debug: compare an attention variant against the generic kernel, error by row tile / head."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kvq-challenge-cvpr-ntire2024_b200"))
import torch
from kvq_b200 import ops
from tools import synth
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda:0")
window = (8, 7, 7)
for (B, D, H, W, C, heads, shift) in [(1, 8, 7, 7, 96, 3, (0, 0, 0)), (2, 8, 14, 14, 96, 3, (0, 0, 0)), (2, 16, 14, 14, 96, 3, (4, 3, 3))]:
    sd = synth.synth_state_dict({"attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
                                 "attn.relative_position_bias_table": (2535, heads),
                                 "attn.fragment_position_bias_table": (2535, heads)}, 5)
    tab = ops.pack_bias_table(sd["attn.relative_position_bias_table"].to(dev), sd["attn.fragment_position_bias_table"].to(dev), window, heads)
    w = ops.cast_f16(sd["attn.qkv.weight"].to(dev)); b = sd["attn.qkv.bias"].to(dev)
    rows = ops.window_rows(B, D, H, W, window, shift)
    torch.manual_seed(1)
    xw = torch.randn(rows, C, device=dev).half()
    o_ref = ops.window_attention(xw, w, b, tab, B, D, H, W, heads, window, shift, debug_variant=2).float()
    o = ops.window_attention(xw, w, b, tab, B, D, H, W, heads, window, shift, debug_variant=variant).float()
    torch.cuda.synchronize()
    err = (o - o_ref).abs().reshape(-1, 392, heads, 32)
    print(f"geom B{B} D{D} H{H} W{W} shift{shift}: max err {err.max().item():.4g}, units {err.shape[0]}")
    qkv = xw.float() @ w.float().t() + b
    qq = qkv[:392, 0:32] * (32 ** -0.5) * 1.4426950408889634; kk_ = qkv[:392, C:C + 32]
    qq = qq.half().float(); kk_ = kk_.half().float()
    for row in (0, 1, 384, 385):
        toks0 = [0, 49, 98, 147]; toks6 = [48 + 49 * d for d in (4, 5, 6, 7)]
        print(f"  host S row {row}: " + " ".join(f"{float(qq[row] @ kk_[t]):.6f}" for t in toks0) + " | " + " ".join(f"{float(qq[row] @ kk_[t]):.6f}" for t in toks6))
    if H == 7 and W == 7:
        L2E = 1.4426950408889634
        rel = sd["attn.relative_position_bias_table"][:, 0].to(dev); frg = sd["attn.fragment_position_bias_table"][:, 0].to(dev)
        idx = torch.arange(392, device=dev); dd = idx // 49; hh = (idx % 49) // 7; ww = idx % 7
        for row in (0, 384):
            rpi = (dd[row] - dd + 7) * 169 + (hh[row] - hh + 6) * 13 + (ww[row] - ww + 6)
            fg = ((hh[row] - hh).abs() + (ww[row] - ww).abs()).float()
            v = qq[row] @ kk_.t() + (frg[rpi] + fg * (rel[rpi] - frg[rpi])) * L2E
            order = (hh * 56 + ww * 8 + dd).argsort()
            vc = v[order].reshape(7, 56)
            m0 = vc[0].max(); l = 0.0
            for c in range(7):
                l += float(torch.exp2(vc[c] - m0).sum())
                print(f"  host row {row} chunk {c}: max {float(vc[c].max()):.5f} m_ref {float(m0):.5f} l_run {l:.5f}")
    oo = o.reshape(-1, 392, heads, 32); rr = o_ref.reshape(-1, 392, heads, 32)
    print("  diff row 384 head 0:", [round(float(x), 4) for x in (oo[0, 384, 0] - rr[0, 384, 0])])
    print("  diff row 391 head 1:", [round(float(x), 4) for x in (oo[0, 391, 1] - rr[0, 391, 1])])
    print("  ref out row 0:", rr[0, 0, 0, :4].tolist(), " row 384:", rr[0, 384, 0, :4].tolist())
    print("  got out row 0:", oo[0, 0, 0, :4].tolist(), " row 384:", oo[0, 384, 0, :4].tolist())
    for t in range(384, 392):
        d = (rr[0, :, 0] - oo[0, t, 0][None]).abs().max(dim=1).values
        print(f"  tail row {t} head 0: closest ref row {int(d.argmin())} (dist {d.min().item():.3g}); own dist {d[t].item():.3g}")
    for u in range(min(err.shape[0], 1)):
        for hd in range(heads):
            e = err[u, :, hd]
            tiles = [e[0:128].max().item(), e[128:256].max().item(), e[256:384].max().item(), e[384:392].max().item()]
            bad = (e.max(dim=1).values > 5e-3).nonzero().flatten().tolist()
            print(f"  win {u} head {hd}: tile max " + " ".join(f"{t:.3g}" for t in tiles) + f"  bad rows {bad[:12]}{'...' if len(bad) > 12 else ''} ({len(bad)})")
