"""Deterministic synthetic weights / inputs shared by the golden generator, the tests, the timing tools and bench.py.

TEST INFRASTRUCTURE ONLY.  There is no network for checkpoints, so every parity number is taken on
seeded random weights.  Each tensor is drawn from its own generator keyed by (seed, crc32(key)) so
the values do not depend on module construction order, and every class of parameter the
reference's default init leaves degenerate (zero fragment tables, zero biases, unit LayerNorm,
BN running stats (0,1) -- SURVEY.md section 0-8) is randomised.
"""
import zlib

import torch


def _gen(seed, key):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))


def fill_like(key, shape, seed, dtype=torch.float32):
    g = _gen(seed, key)
    n = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    leaf = key.rsplit(".", 1)[-1]
    if "position_bias_table" in key:
        t = 0.2 * n
    elif leaf == "running_var":
        t = 0.5 + torch.rand(tuple(shape), generator=g)
    elif leaf == "running_mean":
        t = 0.1 * n
    elif leaf == "num_batches_tracked":
        return torch.zeros(tuple(shape), dtype=torch.int64)
    elif leaf == "bias":
        t = 0.1 * n
    elif leaf == "weight" and len(shape) == 1:          # LayerNorm / BatchNorm scale
        t = 1.0 + 0.2 * n
    elif leaf == "weight" or (leaf.endswith("_weight") and len(shape) == 2):   # Linear / Conv / MHA in_proj_weight:
        # variance-preserving-ish (an O(0.1) in_proj makes the attention logits O(10): a chaotic, near one-hot softmax)
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        t = n * (1.0 / fan_in) ** 0.5
    else:
        t = 0.1 * n
    return t.to(dtype)


def synth_state_dict(shapes, seed):
    """shapes: {key: shape or (shape, 'int64-buffer-value')}.  Returns {key: tensor}."""
    return {k: fill_like(k, s, seed) for k, s in shapes.items()}


def swin_shapes(prefix="", embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24),
                window=(8, 7, 7), frag_biases=(True, True, True, False), patch=(2, 4, 4), in_chans=3):
    """Float parameter shapes of SwinTransformer3D (reference swin_backbone.py:760-842); the int64
    `relative_position_index` buffer is derived, not random, and is rebuilt by whoever needs it."""
    nb = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
    s = {prefix + "patch_embed.proj.weight": (embed_dim, in_chans) + tuple(patch),
         prefix + "patch_embed.proj.bias": (embed_dim,),
         prefix + "patch_embed.norm.weight": (embed_dim,),
         prefix + "patch_embed.norm.bias": (embed_dim,)}
    for i, depth in enumerate(depths):
        C = embed_dim * 2 ** i
        for j in range(depth):
            b = f"{prefix}layers.{i}.blocks.{j}."
            s[b + "norm1.weight"] = (C,)
            s[b + "norm1.bias"] = (C,)
            s[b + "attn.relative_position_bias_table"] = (nb, num_heads[i])
            if frag_biases[i]:
                s[b + "attn.fragment_position_bias_table"] = (nb, num_heads[i])
            s[b + "attn.qkv.weight"] = (3 * C, C)
            s[b + "attn.qkv.bias"] = (3 * C,)
            s[b + "attn.proj.weight"] = (C, C)
            s[b + "attn.proj.bias"] = (C,)
            s[b + "norm2.weight"] = (C,)
            s[b + "norm2.bias"] = (C,)
            s[b + "mlp.fc1.weight"] = (4 * C, C)
            s[b + "mlp.fc1.bias"] = (4 * C,)
            s[b + "mlp.fc2.weight"] = (C, 4 * C)
            s[b + "mlp.fc2.bias"] = (C,)
        if i < len(depths) - 1:
            b = f"{prefix}layers.{i}.downsample."
            s[b + "reduction.weight"] = (2 * C, 4 * C)
            s[b + "norm.weight"] = (4 * C,)
            s[b + "norm.bias"] = (4 * C,)
    Cf = embed_dim * 2 ** (len(depths) - 1)
    s[prefix + "norm.weight"] = (Cf,)
    s[prefix + "norm.bias"] = (Cf,)
    return s


def vqa_head_shapes(prefix="", in_channels=768, hidden=64):
    return {prefix + "fc_hid.weight": (hidden, in_channels, 1, 1, 1), prefix + "fc_hid.bias": (hidden,),
            prefix + "fc_last.weight": (1, hidden, 1, 1, 1), prefix + "fc_last.bias": (1,)}


def swin_network_state_dict(seed, key="swin_tiny_grpb", **kw):
    shapes = dict(swin_shapes(prefix=f"{key}_backbone.", **kw))
    shapes.update(vqa_head_shapes(prefix=f"{key}_head."))
    return synth_state_dict(shapes, seed)


def clip_input(shape, seed):
    """Synthetic normalised frames, [B,3,T,H,W] fp32 ~ N(0,1) (ImageNet-normalised scale)."""
    return torch.randn(tuple(shape), generator=torch.Generator().manual_seed(seed))


def resnet50_shapes(prefix="", layers=(3, 4, 6, 3), feat_dim=9472, hidden=128):
    """Float parameter / buffer shapes of the reference's per-frame ResNet-50 (simpleVQA_model.py:129-218), including
    its own unused `quality` regressor (:168) so that a strict load_state_dict works."""
    s = {prefix + "conv1.weight": (64, 3, 7, 7)}

    def bn(p, c):
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            s[p + leaf] = (c,)

    bn(prefix + "bn1.", 64)
    cin = 64
    for i, depth in enumerate(layers):
        planes = 64 << i
        for j in range(depth):
            b = f"{prefix}layer{i + 1}.{j}."
            s[b + "conv1.weight"] = (planes, cin, 1, 1)
            bn(b + "bn1.", planes)
            s[b + "conv2.weight"] = (planes, planes, 3, 3)
            bn(b + "bn2.", planes)
            s[b + "conv3.weight"] = (planes * 4, planes, 1, 1)
            bn(b + "bn3.", planes * 4)
            if j == 0:
                s[b + "downsample.0.weight"] = (planes * 4, cin, 1, 1)
                bn(b + "downsample.1.", planes * 4)
            cin = planes * 4
    s.update(simplevqa_head_shapes(prefix, feat_dim, hidden))
    return s


def simplevqa_head_shapes(prefix="", feat_dim=9472, hidden=128):
    return {prefix + "quality.0.weight": (hidden, feat_dim), prefix + "quality.0.bias": (hidden,),
            prefix + "quality.1.weight": (1, hidden), prefix + "quality.1.bias": (1,)}


def simplevqa_network_state_dict(seed, key="simpleVQA"):
    """Seeded weights of VQA_Network({'simpleVQA': ...}).  The generic 1/fan_in draw keeps the pooled features O(1)
    through the 16 residual blocks (measured |feat| mean ~1.5, max ~15, like ImageNet-trained features); a ReLU gain
    of sqrt(2) makes them grow past the fp16 range, which no trained checkpoint does."""
    shapes = dict(resnet50_shapes(prefix=f"{key}_backbone."))
    shapes.update(simplevqa_head_shapes(prefix=f"{key}_head."))
    return synth_state_dict(shapes, seed)


def motion_features(shape, seed):
    """Stand-in for the pre-extracted SlowFast features batch['feat'] [B,T,2304] (non-negative: they are pooled
    post-ReLU activations in the reference pipeline)."""
    return torch.randn(tuple(shape), generator=torch.Generator().manual_seed(seed)).abs()


def slowfast_shapes(prefix="feature_extraction."):
    """Float parameter / buffer shapes of slowfast().feature_extraction (SlowFast_features.py:137-152) = blocks[0..4]
    of pytorchvideo's slowfast_r50, under pytorchvideo's module names (oracle/slowfast.py header)."""
    s = {}

    def bn(p, c):
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            s[p + leaf] = (c,)

    p = prefix + "0."
    s[p + "multipathway_blocks.0.conv.weight"] = (64, 3, 1, 7, 7)
    bn(p + "multipathway_blocks.0.norm.", 64)
    s[p + "multipathway_blocks.1.conv.weight"] = (8, 3, 5, 7, 7)
    bn(p + "multipathway_blocks.1.norm.", 8)
    s[p + "multipathway_fusion.conv_fast_to_slow.weight"] = (16, 8, 7, 1, 1)
    bn(p + "multipathway_fusion.norm.", 16)
    cs, cf = 80, 8
    for i, depth in enumerate((3, 4, 6, 3)):
        p = f"{prefix}{i + 1}."
        for path, (cin, inner, ka) in enumerate(((cs, 64 << i, (1, 1, 3, 3)[i]), (cf, 8 << i, 3))):
            for j in range(depth):
                b = f"{p}multipathway_blocks.{path}.res_blocks.{j}."
                c_in = cin if j == 0 else inner * 4
                if j == 0:
                    s[b + "branch1_conv.weight"] = (inner * 4, c_in, 1, 1, 1)
                    bn(b + "branch1_norm.", inner * 4)
                s[b + "branch2.conv_a.weight"] = (inner, c_in, ka, 1, 1)
                bn(b + "branch2.norm_a.", inner)
                s[b + "branch2.conv_b.weight"] = (inner, inner, 1, 3, 3)
                bn(b + "branch2.norm_b.", inner)
                s[b + "branch2.conv_c.weight"] = (inner * 4, inner, 1, 1, 1)
                bn(b + "branch2.norm_c.", inner * 4)
        cs, cf = (64 << i) * 4, (8 << i) * 4
        if i < 3:
            s[p + "multipathway_fusion.conv_fast_to_slow.weight"] = (2 * cf, cf, 7, 1, 1)
            bn(p + "multipathway_fusion.norm.", 2 * cf)
            cs += 2 * cf
    return s


def slowfast_state_dict(seed, prefix="feature_extraction."):
    """Seeded weights of the SlowFast trunk (pretrained Kinetics weights are unavailable offline)."""
    return synth_state_dict(slowfast_shapes(prefix), seed)


def slowfast_frames(shape, seed):
    """Synthetic Kinetics-normalised frames [B,3,T,H,W] ((v - 0.45) / 0.225 of uniform [0,1] pixels,
    SlowFast_features.py:173-174)."""
    u = torch.rand(tuple(shape), generator=torch.Generator().manual_seed(seed))
    return (u - 0.45) / 0.225
