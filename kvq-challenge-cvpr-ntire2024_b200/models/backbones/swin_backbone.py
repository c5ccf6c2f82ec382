"""SwinTransformer3D (GRPB) with the reference's constructor, parameter names and forward signature
(reference models/backbones/swin_backbone.py:736-1085); the forward is one C-ABI call into libkvq_b200.so.

The nn.Modules below are PARAMETER CONTAINERS ONLY: they reproduce the reference state_dict
(`patch_embed.proj.weight`, `layers.{s}.blocks.{j}.attn.{qkv,proj}.{weight,bias}`,
`.attn.relative_position_bias_table`, `.attn.fragment_position_bias_table`, buffer `.attn.relative_position_index`,
`.norm{1,2}`, `.mlp.fc{1,2}`, `layers.{s}.downsample.{reduction,norm}`, `norm`) so reference checkpoints load
unchanged (optionally with the DataParallel `module.` prefix, see load_swin).  No torch op runs in forward and there
is no CPU / eager fallback: CPU tensors raise.
"""
import torch
import torch.nn as nn

from kvq_b200 import ops


def _relative_position_index(window_size):
    """Buffer of WindowAttention3D (:213-235): rpi[i,j] = (d_i-d_j+wd-1)*(2wh-1)(2ww-1) + (h_i-h_j+wh-1)*(2ww-1) + ..."""
    wd, wh, ww = window_size
    coords = torch.stack(torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij"))
    flat = coords.flatten(1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += wd - 1
    rel[:, :, 1] += wh - 1
    rel[:, :, 2] += ww - 1
    rel[:, :, 0] *= (2 * wh - 1) * (2 * ww - 1)
    rel[:, :, 1] *= 2 * ww - 1
    return rel.sum(-1)


class WindowAttention3D(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, frag_bias=False):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        n = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1)
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n, num_heads))
        if frag_bias:
            self.fragment_position_bias_table = nn.Parameter(torch.zeros(n, num_heads))   # zero-init, :202-210
        self.register_buffer("relative_position_index", _relative_position_index(window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class SwinTransformerBlock3D(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio, qkv_bias, frag_bias):
        super().__init__()
        self.shift_size = shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention3D(dim, window_size, num_heads, qkv_bias, frag_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchMerging(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, qkv_bias, downsample, frag_bias):
        super().__init__()
        shift = tuple(i // 2 for i in window_size)
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(dim, num_heads, window_size, (0, 0, 0) if i % 2 == 0 else shift, mlp_ratio,
                                   qkv_bias, frag_bias) for i in range(depth)])
        self.downsample = PatchMerging(dim) if downsample else None


class PatchEmbed3D(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim, norm):
        super().__init__()
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim) if norm else None


class SwinTransformer3D(nn.Module):
    """Same constructor surface as the reference (:760-782).  Arguments that only matter for training
    (drop rates, use_checkpoint, frozen_stages) are accepted and ignored; `pretrained` is loaded with load_swin
    semantics when it names an existing file and silently skipped otherwise (the reference raises at import)."""

    def __init__(self, pretrained=None, pretrained2d=False, patch_size=(2, 4, 4), in_chans=3, embed_dim=96,
                 depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=(8, 7, 7), mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm,
                 patch_norm=True, frozen_stages=-1, use_checkpoint=True,
                 jump_attention=(False, False, False, False), frag_biases=(True, True, True, False),
                 base_x_size=(32, 224, 224)):
        super().__init__()
        if tuple(patch_size) != (2, 4, 4) or in_chans != 3 or not patch_norm or not qkv_bias or qk_scale is not None \
                or mlp_ratio != 4.0 or any(jump_attention):
            raise NotImplementedError("kvq_b200 builds the configurations the reference's YAMLs use: patch (2,4,4), "
                                      "3 input channels, patch_norm, qkv_bias, mlp_ratio 4, no jump_attention")
        self.depths, self.num_heads = tuple(depths), tuple(num_heads)
        self.window_size, self.frag_biases = tuple(window_size), tuple(bool(f) for f in frag_biases)
        self.embed_dim, self.num_layers = embed_dim, len(depths)
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.base_x_size = base_x_size
        self.patch_embed = PatchEmbed3D(patch_size, in_chans, embed_dim, patch_norm)
        self.layers = nn.ModuleList([
            BasicLayer(int(embed_dim * 2 ** i), depths[i], num_heads[i], self.window_size, mlp_ratio, qkv_bias,
                       i < self.num_layers - 1, self.frag_biases[i]) for i in range(self.num_layers)])
        self.norm = nn.LayerNorm(self.num_features)
        self.apply(self._init_weights)
        self._packed = None
        self._packed_key = None
        if isinstance(pretrained, str):
            import os
            if os.path.exists(pretrained):
                self.load_swin(torch.load(pretrained, map_location="cpu"))

    @staticmethod
    def _init_weights(m):                                       # reference :1017-1024
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def load_swin(self, ckpt, strict=False):
        """load_swin (:933-1006): take `state_dict`, strip `backbone.` / `module.`, fork the relative table into the
        fragment table where the checkpoint has none, drop shape-mismatched keys."""
        sd = ckpt.get("state_dict", ckpt)
        own = self.state_dict()
        clean = {}
        for k, v in sd.items():
            for pre in ("module.", "backbone."):
                if k.startswith(pre):
                    k = k[len(pre):]
            clean[k] = v
        # second pass over the CLEANED keys: a fragment table the checkpoint really contains always wins over the fork
        # (reference :945-952 never overwrites an existing fragment_position_bias_table, whatever the key order)
        for k in [k for k in clean if "relative_position_bias_table" in k]:
            fk = k.replace("relative_position_bias_table", "fragment_position_bias_table")
            if fk in own and fk not in clean:
                clean[fk] = clean[k]
        out = {k: v for k, v in clean.items() if k in own and own[k].shape == v.shape}
        return self.load_state_dict(out, strict=strict)

    # ---- packed device weights, rebuilt only when a parameter changed ----
    def _state_key(self, extra=()):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(extra))

    def packed(self, head=None):
        dev = self.norm.weight.device
        if dev.type != "cuda":
            raise RuntimeError("kvq_b200: SwinTransformer3D runs on a CUDA device only (call .to('cuda')); "
                               "there is no CPU fallback")
        extra = list(head.parameters()) if head is not None else []
        key = (self._state_key(extra), id(head))
        if self._packed is None or self._packed_key != key:
            sd = {k: v for k, v in self.state_dict().items()}
            hp = None
            if head is not None:
                sd.update({"__head__." + k: v for k, v in head.state_dict().items()})
                hp = "__head__."
            with torch.cuda.device(dev):
                self._packed = ops.SwinWeights(sd, dev, prefix="", head_prefix=hp, embed_dim=self.embed_dim,
                                               depths=self.depths, num_heads=self.num_heads, window=self.window_size,
                                               frag_biases=self.frag_biases)
            self._packed_key = key
        return self._packed

    @staticmethod
    def _require_cuda(x):
        if not x.is_cuda:
            raise RuntimeError("kvq_b200: input clips must be CUDA tensors -- this path has no CPU fallback")

    def forward(self, batch, multi=False, layer=-1, adaptive_window_size=False):
        """batch['technical'] f32 [B,3,T,H,W] -> [B, 8C, T/2, H/32, W/32]  (:1044-1080).

        adaptive_window_size: every block partitions with window * input_size // base_x_size (get_adaptive_window_size
        :53-61) -- a C-ABI configuration field, same kernels.  layer > -1 / multi return the un-normalised stage
        outputs (feats[layer]) / all of them resized to the last grid and concatenated: the stage outputs are picked up
        through the C-ABI stage hook; `multi` then uses ATen's trilinear resize on them (off the hot path)."""
        x = batch["technical"] if isinstance(batch, dict) else batch
        self._require_cuda(x)
        with torch.cuda.device(x.device):
            wts = self.packed()
            if adaptive_window_size:
                rw = tuple((w * xi) // bx for w, xi, bx in zip(self.window_size, x.shape[2:], self.base_x_size))
                if min(rw) < 1 or any(r > w for r, w in zip(rw, self.window_size)):
                    raise RuntimeError(f"kvq_b200: adaptive window {rw} for input {tuple(x.shape[2:])} must lie within "
                                       f"1..{self.window_size} (the reference's bias-table slice fails outside it too)")
                wts.set_resized_window(rw)
            else:
                wts.set_resized_window(None)
            if not (multi or layer > -1):
                feat, _ = wts.forward(x, want_feat=True, want_score=False)
                return feat
            # feats[0] = patch embedding, feats[s + 1] = BasicLayer s (after its PatchMerging), channels-last tokens
            B, _, T, H, W = x.shape
            D, h, w = (T + 1) // 2, (H + 3) // 4, (W + 3) // 4
            grids = [(D, h, w)]
            for s_ in range(self.num_layers):
                if s_ < self.num_layers - 1:
                    h, w = (h + 1) // 2, (w + 1) // 2
                grids.append((D, h, w))
            feats = {}

            def hook(stage, tokens_ptr, nrows, channels, stream_ptr):
                idx = stage + 1
                if multi or idx == layer:
                    d_, h_, w_ = grids[idx]
                    t = ops.device_view_f32(tokens_ptr, (B, d_, h_, w_, channels), x.device)
                    feats[idx] = t.permute(0, 4, 1, 2, 3).contiguous()            # n c d h w, own storage
                return 0

            final, _ = wts.forward_hooked(x.float() if x.dtype != torch.float32 else x, hook, want_feat=True,
                                          want_score=False)
            if multi:
                import torch.nn.functional as F
                return torch.cat([F.interpolate(feats[i], size=final.shape[2:], mode="trilinear")
                                  for i in range(self.num_layers)], 1)
            if layer not in feats:
                raise IndexError(f"layer={layer}: SwinTransformer3D has feats[0..{self.num_layers}]")
            return feats[layer]

    def forward_with_head(self, x, head, want_feat=False, graph=None):
        """Fused backbone + VQAHead: one C-ABI call, score [B,1] (+ features when asked)."""
        self._require_cuda(x)
        with torch.cuda.device(x.device):
            feat, score = self.packed(head).forward(x, want_feat=want_feat, want_score=True, graph=graph)
        return feat, score.reshape(-1, 1)


def swin_3d_tiny(**kwargs):
    return SwinTransformer3D(depths=[2, 2, 6, 2], frag_biases=[0, 0, 0, 0], **kwargs)       # :1088-1090


def swin_3d_small(**kwargs):
    return SwinTransformer3D(depths=[2, 2, 18, 2], frag_biases=[0, 0, 0, 0], **kwargs)      # :1093-1095
