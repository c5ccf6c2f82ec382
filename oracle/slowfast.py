"""CPU fp32 restatement of the SlowFast-R50 motion-feature extractor of SlowFast_features.py.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / reference legs).

**PARITY UNPINNED.**  The arithmetic lives in a third-party dependency that is absent from /root/reference and from
this image: `pytorchvideo.models.hub.slowfast_r50` (version not pinned by the reference: it is not in
requirements.txt; SlowFast_features.py:21 is the only import, :140 and :148-152 the only call sites).  The reference
holds no test, fixture or golden vector for this path, so this file restates pytorchvideo's published
`create_slowfast` defaults for depth 50 and anchors on the reference call sites:

  SlowFast_features.py:112-135  pack_pathway_output: slow = frames[linspace(0, T-1, T//4).long()], fast = all frames
  SlowFast_features.py:137-152  slowfast.__init__: net.blocks[0..4] -> feature_extraction, blocks[5].pool[0|1]
                                (AvgPool3d((8,7,7)) / AvgPool3d((32,7,7)), stride 1) and blocks[6].output_pool
                                (AdaptiveAvgPool3d(1)); the 400-way projection is NOT used
  SlowFast_features.py:155-165  slowfast.forward -> (slow_feature [B,2048,1,1,1], fast_feature [B,256,1,1,1])

One structural sanity check is available offline and is asserted by tests/test_oracle_slowfast.py: the MAC count of
this restatement at 32x256x256 (65.71 G) equals the figure pytorchvideo publishes for slowfast_r50.

Restated architecture (state_dict names as `slowfast().state_dict()` would carry them, i.e. pytorchvideo's names
under `feature_extraction.<block>.`):
  block 0  multipathway_blocks.{0,1}: stems  conv (1,7,7)/s(1,2,2)/p(0,3,3) 3->64 | conv (5,7,7)/s(1,2,2)/p(2,3,3) 3->8,
           norm (BatchNorm3d eps 1e-5), ReLU, MaxPool3d((1,3,3), s(1,2,2), p(0,1,1))
           multipathway_fusion: conv_fast_to_slow (7,1,1)/s(4,1,1)/p(3,0,0) Cf->2Cf, norm, ReLU, cat([slow, fuse], 1)
  block s  (s = 1..4) multipathway_blocks.{0,1}.res_blocks.{j}: branch2.conv_a (ka,1,1)/p(ka//2,0,0) -> norm_a -> ReLU
           -> conv_b (1,3,3)/s(1,st,st)/p(0,1,1) -> norm_b -> ReLU -> conv_c 1x1x1 -> norm_c; first block of a stage:
           branch1_conv 1x1x1/s(1,st,st) + branch1_norm; sum -> ReLU.  ka slow = 1,1,3,3, fast = 3,3,3,3; st = 1,2,2,2;
           depths 3,4,6,3; slow inner 64..512 -> out 256..2048 (stage inputs 80/320/640/1280); fast inner 8..64 -> out
           32..256; fusion after blocks 1..3, none after block 4; every conv has bias=False.
"""
import torch
import torch.nn.functional as F

DEPTHS = (3, 4, 6, 3)
SPATIAL_STRIDES = (1, 2, 2, 2)
SLOW_KA = (1, 1, 3, 3)
FAST_KA = (3, 3, 3, 3)
BN_EPS = 1e-5
ALPHA = 4                      # slow pathway keeps T // ALPHA frames (SlowFast_features.py:129)
FUSE_KT = 7                    # pytorchvideo slowfast_fusion_conv_kernel_size = (7,1,1), stride (ALPHA,1,1)
SLOW_POOL = (8, 7, 7)          # blocks[5].pool[0]
FAST_POOL = (32, 7, 7)         # blocks[5].pool[1]


def slow_frame_indices(T):
    """SlowFast_features.py:127-131: torch.linspace(0, T-1, T//4).long() (= [0,4,8,13,17,22,26,31] for T = 32)."""
    return torch.linspace(0, T - 1, T // ALPHA).long()


def pack_pathway_output(frames):
    """frames [B,3,T,H,W] -> [slow [B,3,T//4,H,W], fast [B,3,T,H,W]]  (SlowFast_features.py:112-135, minus .to(device))."""
    return [torch.index_select(frames, 2, slow_frame_indices(frames.shape[2])), frames]


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        training=False, eps=BN_EPS)


def _stem(x, sd, p, kt):
    x = F.conv3d(x, sd[p + "conv.weight"], stride=(1, 2, 2), padding=(kt // 2, 3, 3))
    x = F.relu(_bn(x, sd, p + "norm."))
    return F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))


def _fuse(slow, fast, sd, p):
    f = F.conv3d(fast, sd[p + "conv_fast_to_slow.weight"], stride=(ALPHA, 1, 1), padding=(FUSE_KT // 2, 0, 0))
    return torch.cat([slow, F.relu(_bn(f, sd, p + "norm."))], dim=1)


def _res_block(x, sd, p, ka, stride, first):
    out = F.conv3d(x, sd[p + "branch2.conv_a.weight"], padding=(ka // 2, 0, 0))
    out = F.relu(_bn(out, sd, p + "branch2.norm_a."))
    out = F.conv3d(out, sd[p + "branch2.conv_b.weight"], stride=(1, stride, stride), padding=(0, 1, 1))
    out = F.relu(_bn(out, sd, p + "branch2.norm_b."))
    out = _bn(F.conv3d(out, sd[p + "branch2.conv_c.weight"]), sd, p + "branch2.norm_c.")
    idn = x
    if first:
        idn = _bn(F.conv3d(x, sd[p + "branch1_conv.weight"], stride=(1, stride, stride)), sd, p + "branch1_norm.")
    return F.relu(out + idn)


def trunk(slow, fast, sd, prefix="feature_extraction."):
    """feature_extraction = net.blocks[0..4]: [slow, fast] -> [slow [B,2048,Ts,h,w], fast [B,256,Tf,h,w]]."""
    p = prefix + "0."
    slow = _stem(slow, sd, p + "multipathway_blocks.0.", 1)
    fast = _stem(fast, sd, p + "multipathway_blocks.1.", 5)
    slow = _fuse(slow, fast, sd, p + "multipathway_fusion.")
    for s in range(4):
        p = f"{prefix}{s + 1}."
        for j in range(DEPTHS[s]):
            st = SPATIAL_STRIDES[s] if j == 0 else 1
            slow = _res_block(slow, sd, f"{p}multipathway_blocks.0.res_blocks.{j}.", SLOW_KA[s], st, j == 0)
            fast = _res_block(fast, sd, f"{p}multipathway_blocks.1.res_blocks.{j}.", FAST_KA[s], st, j == 0)
        if s < 3:
            slow = _fuse(slow, fast, sd, p + "multipathway_fusion.")
    return slow, fast


def head_pools(slow, fast):
    """blocks[5].pool[i] (AvgPool3d, stride 1, no padding) then blocks[6].output_pool (AdaptiveAvgPool3d(1)).  With a
    res5 map larger than the 7x7 kernel (8x8 at 256^2) this is the mean of overlapping window means, not a plain
    global mean (SURVEY.md section 8c 'pool caveat')."""
    s = F.adaptive_avg_pool3d(F.avg_pool3d(slow, SLOW_POOL, stride=1), 1)
    f = F.adaptive_avg_pool3d(F.avg_pool3d(fast, FAST_POOL, stride=1), 1)
    return s, f


def slowfast_forward(inputs, sd, prefix="feature_extraction."):
    """slowfast.forward (SlowFast_features.py:155-165): [slow, fast] -> (slow_feature, fast_feature)."""
    with torch.no_grad():
        sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
        s, f = trunk(inputs[0].float(), inputs[1].float(), sd, prefix)
        return head_pools(s, f)


# ---- fp16-storage emulation (same rounding points as the CUDA path: BN folded into fp16 weights, fp16 activations
# between layers, fp32 accumulation); separates "kernel wrong" from "fp16 storage moves the features by this much"
def _h(t):
    return t.half().float()


def _fold(sd, conv, bn):
    scale = sd[bn + "weight"] / torch.sqrt(sd[bn + "running_var"] + BN_EPS)
    return _h(sd[conv] * scale.view(-1, 1, 1, 1, 1)), sd[bn + "bias"] - sd[bn + "running_mean"] * scale


def _stem16(x, sd, p, kt):
    w, b = _fold(sd, p + "conv.weight", p + "norm.")
    x = _h(F.relu(F.conv3d(_h(x), w, b, stride=(1, 2, 2), padding=(kt // 2, 3, 3))))
    return F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))


def _fuse16(slow, fast, sd, p):
    w, b = _fold(sd, p + "conv_fast_to_slow.weight", p + "norm.")
    f = _h(F.relu(F.conv3d(fast, w, b, stride=(ALPHA, 1, 1), padding=(FUSE_KT // 2, 0, 0))))
    return torch.cat([slow, f], dim=1)


def _res_block16(x, sd, p, ka, stride, first):
    wa, ba = _fold(sd, p + "branch2.conv_a.weight", p + "branch2.norm_a.")
    wb, bb = _fold(sd, p + "branch2.conv_b.weight", p + "branch2.norm_b.")
    wc, bc = _fold(sd, p + "branch2.conv_c.weight", p + "branch2.norm_c.")
    out = _h(F.relu(F.conv3d(x, wa, ba, padding=(ka // 2, 0, 0))))
    out = _h(F.relu(F.conv3d(out, wb, bb, stride=(1, stride, stride), padding=(0, 1, 1))))
    idn = x
    if first:
        w1, b1 = _fold(sd, p + "branch1_conv.weight", p + "branch1_norm.")
        idn = _h(F.conv3d(x, w1, b1, stride=(1, stride, stride)))
    return _h(F.relu(F.conv3d(out, wc, bc) + idn))


def slowfast_forward_fp16(inputs, sd, prefix="feature_extraction."):
    with torch.no_grad():
        sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
        p = prefix + "0."
        slow = _stem16(inputs[0].float(), sd, p + "multipathway_blocks.0.", 1)
        fast = _stem16(inputs[1].float(), sd, p + "multipathway_blocks.1.", 5)
        slow = _fuse16(slow, fast, sd, p + "multipathway_fusion.")
        for s in range(4):
            p = f"{prefix}{s + 1}."
            for j in range(DEPTHS[s]):
                st = SPATIAL_STRIDES[s] if j == 0 else 1
                slow = _res_block16(slow, sd, f"{p}multipathway_blocks.0.res_blocks.{j}.", SLOW_KA[s], st, j == 0)
                fast = _res_block16(fast, sd, f"{p}multipathway_blocks.1.res_blocks.{j}.", FAST_KA[s], st, j == 0)
            if s < 3:
                slow = _fuse16(slow, fast, sd, p + "multipathway_fusion.")
        return head_pools(slow, fast)


def count_macs(T=32, H=256, W=256, per_stage=False):
    """Multiply-accumulates of the trunk per clip, from the conv shapes alone (no arithmetic).  per_stage=True returns
    [stem (+ its lateral), res2, res3, res4, res5] with each stage's trailing lateral conv counted in the stage."""
    def out(n, k, s, p):
        return (n + 2 * p - k) // s + 1

    macs = 0
    stages = []
    Ts, Tf = T // ALPHA, T
    h, w = out(H, 7, 2, 3), out(W, 7, 2, 3)
    macs += Ts * h * w * 64 * 3 * 49 + Tf * h * w * 8 * 3 * 5 * 49
    h, w = out(h, 3, 2, 1), out(w, 3, 2, 1)
    cs, cf = 64, 8
    macs += Ts * h * w * (2 * cf) * cf * FUSE_KT
    cs += 2 * cf
    stages.append(macs)
    for s in range(4):
        inner_s, inner_f = 64 << s, 8 << s
        for j in range(DEPTHS[s]):
            st = SPATIAL_STRIDES[s] if j == 0 else 1
            ho, wo = out(h, 3, st, 1), out(w, 3, st, 1)
            for (T_, cin, inner, ka) in ((Ts, cs, inner_s, SLOW_KA[s]), (Tf, cf, inner_f, FAST_KA[s])):
                macs += T_ * h * w * inner * cin * ka            # conv_a at the input resolution
                macs += T_ * ho * wo * inner * inner * 9         # conv_b
                macs += T_ * ho * wo * inner * 4 * inner         # conv_c
                if j == 0:
                    macs += T_ * ho * wo * inner * 4 * cin       # branch1
            cs, cf = inner_s * 4, inner_f * 4
            h, w = ho, wo
        if s < 3:
            macs += Ts * h * w * (2 * cf) * cf * FUSE_KT
            cs += 2 * cf
        stages.append(macs - sum(stages))
    return stages if per_stage else macs
