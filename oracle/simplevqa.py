"""CPU fp32 restatement of the SimpleVQA spatial branch: per-frame ResNet-50, multi-scale mean / unbiased-std pooling,
concatenation with the pre-extracted motion features, and simpleVQAHead.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).  Pinned to the real
reference by tests/golden/simplevqa_*.npz (tools/make_golden_extra.py runs the reference's
models/backbones/simpleVQA_model.resnet50 + models/head.simpleVQAHead on the same seeded weights).

Reference walk (paths relative to the reference root):
  models/backbones/simpleVQA_model.py:220-264  ResNet.forward   (frame loop folded into the batch, pools, cat)
  models/backbones/simpleVQA_model.py:106-126  Bottleneck.forward
  models/backbones/simpleVQA_model.py:18-20    global_std_pool2d  (torch.std over H*W, unbiased)
  models/head.py:10-31                         simpleVQAHead  (Linear -> Linear, mean over frames)
"""
import torch
import torch.nn.functional as F

LAYERS = (3, 4, 6, 3)
BN_EPS = 1e-5


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        training=False, eps=BN_EPS)


def bottleneck(x, sd, p, stride, has_down):
    """simpleVQA_model.py:106-126 (stride lives on the 3x3 conv, :98)."""
    out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1."))
    out = F.relu(_bn(F.conv2d(out, sd[p + "conv2.weight"], stride=stride, padding=1), sd, p + "bn2."))
    out = _bn(F.conv2d(out, sd[p + "conv3.weight"]), sd, p + "bn3.")
    idn = x
    if has_down:
        idn = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1.")
    return F.relu(out + idn)


def resnet_features(x, feat3d, sd, prefix="", layers=LAYERS):
    """x [B,3,T,H,W], feat3d [B,T,F] -> [B,T,7168+F]  (simpleVQA_model.py:220-264)."""
    B, _, T, H, W = x.shape
    x = x.permute(0, 2, 1, 3, 4).reshape(B * T, 3, H, W)                           # :225-231
    x = F.relu(_bn(F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3), sd, prefix + "bn1."))
    x = F.max_pool2d(x, 3, 2, 1)                                                    # :238
    pooled = []
    for s, depth in enumerate(layers):
        for j in range(depth):
            x = bottleneck(x, sd, f"{prefix}layer{s + 1}.{j}.", 2 if (j == 0 and s > 0) else 1, j == 0)
        if s >= 1:                                                                  # :242-251
            flat = x.flatten(2)
            pooled += [flat.mean(dim=2), flat.std(dim=2)]
    feats = torch.cat(pooled + [feat3d.reshape(B * T, -1).to(x.dtype)], dim=1)      # :252-256
    return feats.view(B, T, -1)


def simplevqa_head(feats, sd, prefix=""):
    """head.py:28-31: quality = Linear(9472,128) -> Linear(128,1); mean over the frame axis.  [B,T,F] -> [B,1]."""
    h = F.linear(feats, sd[prefix + "quality.0.weight"], sd[prefix + "quality.0.bias"])
    h = F.linear(h, sd[prefix + "quality.1.weight"], sd[prefix + "quality.1.bias"])
    return h.mean(dim=1)


def simplevqa_forward(x, feat3d, sd, backbone_prefix="simpleVQA_backbone.", head_prefix="simpleVQA_head."):
    """VQA_Network.forward for the simpleVQA key (models/model.py:93-121): returns (feats [B,T,F], score [B,1])."""
    with torch.no_grad():
        sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
        feats = resnet_features(x.float(), feat3d.float(), sd, backbone_prefix)
        return feats, simplevqa_head(feats, sd, head_prefix)


# ---- fp16-storage emulation: the same arithmetic with weights / activations rounded where the CUDA path rounds them
# (BatchNorm folded into fp16 weights, fp16 activations between layers, fp32 accumulation).  Used by the tests to
# separate "the kernel is wrong" from "fp16 storage moves the score by this much".
def _h(t):
    return t.half().float()


def _fold(sd, conv, bn):
    scale = sd[bn + "weight"] / torch.sqrt(sd[bn + "running_var"] + BN_EPS)
    return _h(sd[conv + "weight"] * scale.view(-1, 1, 1, 1)), sd[bn + "bias"] - sd[bn + "running_mean"] * scale


def resnet_features_fp16(x, feat3d, sd, prefix="", layers=LAYERS):
    B, _, T, H, W = x.shape
    x = _h(x.permute(0, 2, 1, 3, 4).reshape(B * T, 3, H, W))
    w, b = _fold(sd, prefix + "conv1.", prefix + "bn1.")
    x = _h(F.relu(F.conv2d(x, w, b, stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    pooled = []
    for s, depth in enumerate(layers):
        for j in range(depth):
            p = f"{prefix}layer{s + 1}.{j}."
            stride = 2 if (j == 0 and s > 0) else 1
            w1, b1 = _fold(sd, p + "conv1.", p + "bn1.")
            w2, b2 = _fold(sd, p + "conv2.", p + "bn2.")
            w3, b3 = _fold(sd, p + "conv3.", p + "bn3.")
            out = _h(F.relu(F.conv2d(x, w1, b1)))
            out = _h(F.relu(F.conv2d(out, w2, b2, stride=stride, padding=1)))
            idn = x
            if j == 0:
                wd, bd = _fold(sd, p + "downsample.0.", p + "downsample.1.")
                idn = _h(F.conv2d(x, wd, bd, stride=stride))
            x = _h(F.relu(F.conv2d(out, w3, b3) + idn))
        if s >= 1:
            flat = x.flatten(2)
            pooled += [flat.mean(dim=2), flat.std(dim=2)]
    return torch.cat(pooled + [feat3d.reshape(B * T, -1)], dim=1).view(B, T, -1)


def simplevqa_forward_fp16(x, feat3d, sd, backbone_prefix="simpleVQA_backbone.", head_prefix="simpleVQA_head."):
    with torch.no_grad():
        sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
        feats = resnet_features_fp16(x.float(), feat3d.float(), sd, backbone_prefix)
        return feats, simplevqa_head(feats, sd, head_prefix)
