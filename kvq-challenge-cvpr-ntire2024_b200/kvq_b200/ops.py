"""torch-tensor front ends of the C-ABI operators.  PyTorch only supplies device memory and the stream."""
import ctypes
import os

import torch

from . import lib as _l


# kernels launched through CUDA-graph replays (kvq_launch_count() only sees eager launches and captures)
GRAPH_KERNEL_LAUNCHES = 0


def kernel_launches():
    """Total kernels of libkvq_b200.so launched by this process: eager launches + graph-replayed ones."""
    return int(_l.load().kvq_launch_count()) + GRAPH_KERNEL_LAUNCHES


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("kvq_b200 operators take CUDA tensors only (no CPU fallback exists)")
        if t is not None and not t.is_contiguous():
            raise RuntimeError("kvq_b200 operators take contiguous tensors")


def cast_f16(t):
    t = t.detach().float().contiguous()
    _need_cuda(t)
    out = torch.empty(t.shape, dtype=torch.float16, device=t.device)
    _l.check(_l.load().kvq_cast_f16(_p(t), _p(out), t.numel(), _stream()), "cast_f16")
    return out


def pack_split_f16(w2d):
    """fp32 [rows, K] -> fp16 [rows, 2*ceil64(K)] = [hi | lo]: hi + lo reproduces the fp32 weight to 2^-22."""
    w2d = w2d.detach().float().contiguous()
    _need_cuda(w2d)
    rows, K = w2d.shape
    out = torch.empty((rows, 2 * ((K + 63) // 64 * 64)), dtype=torch.float16, device=w2d.device)
    _l.check(_l.load().kvq_pack_split_f16(_p(w2d), _p(out), rows, K, _stream()), "pack_split_f16")
    return out


class _DeviceView:
    """A raw device pointer exposed through __cuda_array_interface__ so torch can alias it without a copy."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


def device_view_f32(ptr, shape, device):
    """fp32 tensor aliasing `ptr` (used inside stage hooks: valid only while the forward's workspace is alive)."""
    return torch.as_tensor(_DeviceView(ptr, shape), device=device)


def attn_table_len(window):
    return _l.load().kvq_attn_table_len(*[int(w) for w in window])


def pack_bias_table(rel, frag, window, heads):
    """[L, heads] fp32 tables -> packed [heads, tab_len, 2] fp32 ({frag, rel - frag} or {rel, 0})."""
    rel = rel.detach().float().contiguous()
    frag = frag.detach().float().contiguous() if frag is not None else None
    _need_cuda(rel, frag)
    out = torch.empty((heads, attn_table_len(window), 2), dtype=torch.float32, device=rel.device)
    _l.check(_l.load().kvq_pack_bias_table(_p(rel), _p(frag), _p(out), int(window[0]), int(window[1]), int(window[2]),
                                           heads, _stream()), "pack_bias_table")
    return out


def linear_f16(a, w, bias=None, gelu=False):
    _need_cuda(a, w, bias)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float16, device=a.device)
    _l.check(_l.load().kvq_linear_f16(_p(a), _p(w), _p(bias), _p(out), M, N, K, int(gelu), _stream()), "linear_f16")
    return out


def linear_resid_f32(a, w, bias=None, resid=None, out=None):
    _need_cuda(a, w, bias, resid, out)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    _l.check(_l.load().kvq_linear_resid_f32(_p(a), _p(w), _p(bias), _p(resid), _p(out), M, N, K, _stream()),
             "linear_resid_f32")
    return out


def mlp_fused(a, w1, b1, w2, b2, x):
    """In place: x[M,C] (f32) += fc2(gelu(fc1(a) + b1)) + b2, a f16 [M,C] (the LN2 output)."""
    _need_cuda(a, w1, b1, w2, b2, x)
    M, C = a.shape
    _l.check(_l.load().kvq_mlp_fused(_p(a), _p(w1), _p(b1), _p(w2), _p(b2), _p(x), M, C, _stream()), "mlp_fused")
    return x


def window_rows(B, D, H, W, window, shift):
    return int(_l.load().kvq_window_rows(B, D, H, W, _l.i3(window), _l.i3(shift)))


def ln_window(x, gamma, beta, window, shift, eps=1e-5):
    """x f32 [B,D,H,W,C] -> f16 [B*nW*N, C]: norm1 + roll(-shift) + window_partition."""
    _need_cuda(x, gamma, beta)
    B, D, H, W, C = x.shape
    rows = window_rows(B, D, H, W, window, shift)
    out = torch.empty((rows, C), dtype=torch.float16, device=x.device)
    _l.check(_l.load().kvq_ln_window(_p(x), _p(out), _p(gamma), _p(beta), eps, B, D, H, W, C, _l.i3(window),
                                     _l.i3(shift), _stream()), "ln_window")
    return out


def window_attention(xw, qkv_w, qkv_b, packed_table, B, D, H, W, heads, window, shift, debug_variant=0):
    """xw f16 [B*nW*N, C] (window order) -> attention output f16 [B*nW*N, C] (before proj)."""
    _need_cuda(xw, qkv_w, qkv_b, packed_table)
    C = xw.shape[1]
    lib = _l.load()
    nbytes = lib.kvq_window_attention_workspace_bytes(B, D, H, W, C, _l.i3(window), _l.i3(shift))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=xw.device)
    out = torch.empty_like(xw)
    _l.check(lib.kvq_window_attention(_p(xw), _p(qkv_w), _p(qkv_b), _p(packed_table), _p(out), B, D, H, W, C, heads,
                                      _l.i3(window), _l.i3(shift), _p(ws), nbytes, debug_variant, _stream()),
             "window_attention")
    return out


def vqa_head(feat, w1_f16, b1, w2, b2):
    """feat f32 [B,C,D,H,W] -> score f32 [B,1]  (VQAHead.forward, head.py:60-68; dropout is identity in eval)."""
    _need_cuda(feat, w1_f16, b1, w2, b2)
    if feat.dtype != torch.float32:
        raise RuntimeError("vqa_head: features must be float32")
    B, C = feat.shape[:2]
    tokens = feat[0, 0].numel()
    lib = _l.load()
    nbytes = lib.kvq_vqa_head_workspace_bytes(B, C, tokens)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=feat.device)
    out = torch.empty(B, dtype=torch.float32, device=feat.device)
    _l.check(lib.kvq_vqa_head(_p(feat), _p(w1_f16), _p(b1), _p(w2), _p(b2), _p(out), B, C, tokens, w1_f16.shape[0],
                              _p(ws), nbytes, _stream()), "vqa_head")
    return out.reshape(B, 1)


IMAGENET_MEAN = (123.675, 116.28, 103.53)
IMAGENET_STD = (58.395, 57.12, 57.375)


def fragment_gather_u8(frames, offsets, fragments_h=7, fragments_w=7, fsize=32, aligned=8, mean=IMAGENET_MEAN,
                       std=IMAGENET_STD, out=None):
    """frames u8 [B,T,3,Hs,Ws], offsets i32 [B,2,fh,fw,T//aligned] -> f32 [B,3,T,fh*fsize,fw*fsize] normalised."""
    _need_cuda(frames, offsets)
    if frames.dtype != torch.uint8 or offsets.dtype != torch.int32:
        raise RuntimeError("fragment_gather_u8: frames must be uint8 and offsets int32")
    B, T, _, Hs, Ws = frames.shape
    shape = (B, 3, T, fragments_h * fsize, fragments_w * fsize)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=frames.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise RuntimeError(f"fragment_gather_u8: out must be a contiguous float32 {shape} tensor")
    _l.check(_l.load().kvq_fragment_gather_u8(_p(frames), _p(offsets), _p(out), B, T, Hs, Ws, fragments_h, fragments_w,
                                              fsize, aligned, _l.f3(mean), _l.f3(std), _stream()), "fragment_gather_u8")
    return out


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_VIEW_LAYOUTS = {"BT3HW": 0, "B3THW": 1}


def resize_view_u8(frames, out_h, out_w, crop=None, mean=IMAGENET_MEAN, std=IMAGENET_STD, divisor=1.0, layout="BT3HW",
                   want_u8=False, want_f32=True, workspace=None, antialias=True):
    """torchvision Resize((out_h, out_w)) on uint8 frames (get_resized_video / get_resizecrop_video,
    fusion_datasets.py:244-252, :299-316) + optional crop window + normalisation (:1017-1027, :902-905), byte- and
    float-exact to the reference.  frames u8 [B,T,3,Hs,Ws] (layout "BT3HW") or [B,3,T,Hs,Ws] ("B3THW"); crop =
    (y, x, h, w) on the resized frame or None -> (u8 [B,3,T,h,w] or None, f32 [B,3,T,h,w] = ((v / divisor) - mean[c]) /
    std[c] or None).  antialias=True is torchvision's tensor default since 0.17 (ATen's anti-aliased filter); False is
    what the torch ~= 1.10 environment pinned by the reference computes (plain two-tap bilinear) -- both bit-exact."""
    _need_cuda(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 5:
        raise RuntimeError("resize_view_u8: frames must be a 5-D uint8 tensor")
    if layout not in _VIEW_LAYOUTS:
        raise RuntimeError(f"resize_view_u8: layout {layout!r} is not one of {sorted(_VIEW_LAYOUTS)}")
    if layout == "BT3HW":
        B, T, C, Hs, Ws = frames.shape
    else:
        B, C, T, Hs, Ws = frames.shape
    if C != 3:
        raise RuntimeError(f"resize_view_u8: {C} colour channels (3 expected) for layout {layout}")
    cy, cx, ch, cw = (0, 0, out_h, out_w) if crop is None else (int(v) for v in crop)
    lib = _l.load()
    if not antialias:
        out_u8 = torch.empty((B, 3, T, ch, cw), dtype=torch.uint8, device=frames.device) if want_u8 else None
        out_f32 = torch.empty((B, 3, T, ch, cw), dtype=torch.float32, device=frames.device) if want_f32 else None
        _l.check(lib.kvq_resize_view_bilinear_u8(_p(frames), _VIEW_LAYOUTS[layout], B, T, Hs, Ws, out_h, out_w, cy, cx, ch,
                                                 cw, float(divisor), _l.f3(mean), _l.f3(std), _p(out_u8), _p(out_f32),
                                                 _stream()), "resize_view_bilinear_u8")
        return out_u8, out_f32
    need = lib.kvq_resize_view_workspace_bytes(B, T, Hs, Ws, out_h, out_w, cy, cx, ch, cw)
    if need == 0:
        raise RuntimeError(f"kvq_b200.resize_view_u8: {_l.last_error()}")
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=frames.device)
    out_u8 = torch.empty((B, 3, T, ch, cw), dtype=torch.uint8, device=frames.device) if want_u8 else None
    out_f32 = torch.empty((B, 3, T, ch, cw), dtype=torch.float32, device=frames.device) if want_f32 else None
    _l.check(lib.kvq_resize_view_u8(_p(frames), _VIEW_LAYOUTS[layout], B, T, Hs, Ws, out_h, out_w, cy, cx, ch, cw,
                                    float(divisor), _l.f3(mean), _l.f3(std), _p(out_u8), _p(out_f32), _p(workspace),
                                    workspace.numel() * workspace.element_size(), _stream()), "resize_view_u8")
    return out_u8, out_f32


def qrs_select_gather(fragment, cls_attn, n_key=4, anchor=32, region_patches=7):
    """KSVQE quality-aware region selection (patchnet.py:461-550, eval): fragment f32 [B,3,T,H,W], cls_attn f32
    [B*n_key, L] -> (x_sel f32 [B,3,T,anchor*region_patches, anchor*region_patches], region i32 [B, n_key])."""
    _need_cuda(fragment, cls_attn)
    if fragment.dtype != torch.float32 or cls_attn.dtype != torch.float32:
        raise RuntimeError("qrs_select_gather takes float32 tensors")
    fragment, cls_attn = fragment.contiguous(), cls_attn.contiguous()
    B, _, T, H, W = fragment.shape
    S = anchor * region_patches
    out = torch.empty((B, 3, T, S, S), dtype=torch.float32, device=fragment.device)
    region = torch.empty((B, n_key), dtype=torch.int32, device=fragment.device)
    _l.check(_l.load().kvq_qrs_select_gather(_p(fragment), _p(cls_attn), _p(out), _p(region), B, T, H, W, n_key,
                                             cls_attn.shape[-1], anchor, region_patches, _stream()), "qrs_select_gather")
    return out, region


class SwinWeights:
    """Device-resident packed weights of SwinTransformer3D (+ VQAHead) in the order include/kvq_b200.h documents.

    `sd` maps the REFERENCE state_dict names (swin_backbone.py:760-842; head.py:33-58) to tensors; `prefix` /
    `head_prefix` select the sub-module (e.g. 'swin_tiny_grpb_backbone.' / 'swin_tiny_grpb_head.')."""

    def __init__(self, sd, device, prefix="", head_prefix=None, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window=(8, 7, 7), frag_biases=(True, True, True, False), eps=1e-5,
                 split_weights=7):
        self.device = torch.device(device)
        self.cfg = _l.KvqSwinConfig()
        self.cfg.embed_dim = embed_dim
        self.cfg.num_stages = len(depths)
        for i, (d, h, f) in enumerate(zip(depths, num_heads, frag_biases)):
            self.cfg.depths[i], self.cfg.num_heads[i], self.cfg.frag_bias[i] = int(d), int(h), int(bool(f))
        for i in range(3):
            self.cfg.window[i] = int(window[i])
        self.cfg.ln_eps = eps
        self.cfg.head_hidden = 0
        # fp16 hi/lo weight pairs for the few GEMMs whose weight rounding dominates the score error (patch-embed,
        # the three PatchMerging reductions, VQAHead.fc_hid): measured 7.2e-4 -> see DESIGN.md section 2
        self.cfg.split_weights = int(split_weights)
        self.depths, self.final_dim = tuple(depths), embed_dim * 2 ** (len(depths) - 1)

        def f32(k):
            return sd[k].detach().to(self.device, torch.float32).contiguous()

        def f16(k, shape=None, split=False):
            t = f32(k)
            t = t.reshape(shape) if shape is not None else t
            return pack_split_f16(t.reshape(t.shape[0], -1)) if split else cast_f16(t)

        ts = [f16(prefix + "patch_embed.proj.weight", (embed_dim, -1), split=bool(split_weights & 1)), f32(prefix + "patch_embed.proj.bias"),
              f32(prefix + "patch_embed.norm.weight"), f32(prefix + "patch_embed.norm.bias")]
        for s, depth in enumerate(depths):
            for j in range(depth):
                b = f"{prefix}layers.{s}.blocks.{j}."
                frag = f32(b + "attn.fragment_position_bias_table") if frag_biases[s] else None
                ts += [f32(b + "norm1.weight"), f32(b + "norm1.bias"), f16(b + "attn.qkv.weight"),
                       f32(b + "attn.qkv.bias"),
                       pack_bias_table(f32(b + "attn.relative_position_bias_table"), frag, window, num_heads[s]),
                       f16(b + "attn.proj.weight"), f32(b + "attn.proj.bias"), f32(b + "norm2.weight"),
                       f32(b + "norm2.bias"), f16(b + "mlp.fc1.weight"), f32(b + "mlp.fc1.bias"),
                       f16(b + "mlp.fc2.weight"), f32(b + "mlp.fc2.bias")]
        # per-stage downsample entries follow ALL blocks (header order): collect separately, then interleave
        merged = ts[:4]
        pos = 4
        for s, depth in enumerate(depths):
            merged += ts[pos:pos + 13 * depth]
            pos += 13 * depth
            if s < len(depths) - 1:
                b = f"{prefix}layers.{s}.downsample."
                merged += [f32(b + "norm.weight"), f32(b + "norm.bias"),
                           f16(b + "reduction.weight", split=bool(split_weights & 2))]
        merged += [f32(prefix + "norm.weight"), f32(prefix + "norm.bias")]
        if head_prefix is not None:
            hid = sd[head_prefix + "fc_hid.weight"].shape[0]
            self.cfg.head_hidden = int(hid)
            merged += [f16(head_prefix + "fc_hid.weight", (hid, -1), split=bool(split_weights & 4)), f32(head_prefix + "fc_hid.bias"),
                       f32(head_prefix + "fc_last.weight").reshape(-1).contiguous(),
                       f32(head_prefix + "fc_last.bias").reshape(-1).contiguous()]
        self.tensors = merged
        n = _l.load().kvq_swin3d_num_weights(ctypes.byref(self.cfg))
        if n != len(merged):
            raise RuntimeError(f"kvq_b200: weight table has {len(merged)} entries, library expects {n}: "
                               f"{_l.last_error()}")
        self.ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in merged])
        self._ws = None
        self._graphs = {}

    def workspace(self, B, T, H, W):
        need = _l.load().kvq_swin3d_workspace_bytes(ctypes.byref(self.cfg), B, T, H, W)
        if need == 0:
            raise RuntimeError(f"kvq_b200: cannot plan a [{B},3,{T},{H},{W}] forward: {_l.last_error()}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._graphs = {}                      # captured graphs point into the old workspace
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, need

    def feat_shape(self, B, T, H, W):
        D, h, w = (T + 1) // 2, (H + 3) // 4, (W + 3) // 4
        for _ in range(len(self.depths) - 1):
            h, w = (h + 1) // 2, (w + 1) // 2
        return (B, self.final_dim, D, h, w)

    def _launch(self, x, feat, score, ws):
        B, _, T, H, W = x.shape
        fn = _l.load().kvq_swin3d_forward_x16 if x.dtype == torch.float16 else _l.load().kvq_swin3d_forward
        rc = fn(ctypes.byref(self.cfg), self.ptrs, len(self.tensors), _p(x), B, T, H, W, _p(feat), _p(score), _p(ws),
                ws.numel(), _stream())
        _l.check(rc, "swin3d_forward")

    def set_resized_window(self, window):
        """adaptive_window_size: partition with `window` (<= the base window per dim) from the next forward on; None
        switches it off.  Captured graphs are per setting."""
        rw = (0, 0, 0) if window is None else tuple(int(v) for v in window)
        if tuple(self.cfg.resized_window) != rw:
            self.cfg.resized_window = (ctypes.c_int32 * 3)(*rw)
            self._graphs = {}
            self._ws = None

    def forward_hooked(self, x, stage_hook, want_feat=True, want_score=True):
        """Eager forward with `stage_hook(stage, tokens_ptr, rows, channels, stream_ptr) -> int` called after the patch
        embedding (stage -1) and after every stage; tokens_ptr is the raw device pointer of the fp32 [rows, channels]
        output, which work enqueued on the stream may modify in place (KSVQE's modulation).  Returns (feat or None,
        score or None)."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("kvq_b200: input clips must be float32 CUDA tensors (no CPU fallback exists)")
        x = x.contiguous()
        B, _, T, H, W = x.shape
        ws, _ = self.workspace(B, T, H, W)
        feat = torch.empty(self.feat_shape(B, T, H, W), dtype=torch.float32, device=x.device) if want_feat else None
        score = torch.empty(B, dtype=torch.float32, device=x.device) if want_score and self.cfg.head_hidden else None
        err = []

        def tramp(_arg, stage, tokens, rows, channels, stream):
            try:
                return int(stage_hook(stage, tokens, rows, channels, stream) or 0)
            except Exception as e:                      # never let an exception cross the C boundary
                err.append(e)
                return -1

        cb = _l.STAGE_HOOK(tramp)
        rc = _l.load().kvq_swin3d_forward_hooked(ctypes.byref(self.cfg), self.ptrs, len(self.tensors), _p(x), B, T, H, W,
                                                 _p(feat), _p(score), _p(ws), ws.numel(), _stream(), cb, None)
        if err:
            raise err[0]
        _l.check(rc, "swin3d_forward_hooked")
        return feat, score

    def forward(self, x, want_feat=False, want_score=True, score_out=None, graph=None):
        """x f32 (or f16: clips shipped at half the bytes, identical operands after the patch embedding's fp16
        rounding) [B,3,T,H,W] on this device -> (feat [B,Cf,D,h,w] or None, score [B] or None).

        graph=True (default: env KVQ_CUDA_GRAPH=1) replays the ~95 kernel launches of the forward from a CUDA graph
        captured per (input buffer, shape): the library enqueues everything on the current stream, allocates nothing
        and takes tensor maps by value, so the whole call is capturable.  Outputs of a graphed call live in buffers
        owned by the graph and are overwritten by the next replay for the same input buffer."""
        if not x.is_cuda or x.dtype not in (torch.float32, torch.float16):
            raise RuntimeError("kvq_b200: input clips must be float32 / float16 CUDA tensors (no CPU fallback exists)")
        x = x.contiguous()
        B, _, T, H, W = x.shape
        ws, need = self.workspace(B, T, H, W)
        has_score = want_score and self.cfg.head_hidden > 0
        if graph is None:
            graph = os.environ.get("KVQ_CUDA_GRAPH", "0") == "1"
        if graph and score_out is None:
            key = (x.data_ptr(), tuple(x.shape), x.dtype, bool(want_feat), has_score, ws.data_ptr())
            entry = self._graphs.get(key)
            if entry is None:
                feat = torch.empty(self.feat_shape(B, T, H, W), dtype=torch.float32, device=x.device) if want_feat else None
                score = torch.empty(B, dtype=torch.float32, device=x.device) if has_score else None
                self._launch(x, feat, score, ws)                       # warm-up: one-time function attributes
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = _l.load().kvq_launch_count()
                with torch.cuda.graph(g):
                    self._launch(x, feat, score, ws)
                nodes = int(_l.load().kvq_launch_count() - n0)          # kernels a replay launches
                if len(self._graphs) >= 8:                               # bounded cache
                    self._graphs.pop(next(iter(self._graphs)))
                entry = self._graphs[key] = (g, feat, score, nodes)
            entry[0].replay()
            global GRAPH_KERNEL_LAUNCHES
            GRAPH_KERNEL_LAUNCHES += entry[3]
            return entry[1], entry[2]
        feat = torch.empty(self.feat_shape(B, T, H, W), dtype=torch.float32, device=x.device) if want_feat else None
        score = None
        if has_score:
            score = score_out if score_out is not None else torch.empty(B, dtype=torch.float32, device=x.device)
        self._launch(x, feat, score, ws)
        return feat, score


# ---------------------------------------------------------------------------------------------------------------
# convolution building blocks + the SimpleVQA spatial branch
# ---------------------------------------------------------------------------------------------------------------
def _i3(v):
    return (ctypes.c_int32 * 3)(*[int(a) for a in v])


def pack_conv_weight(w, bn=None, eps=1e-5, pad_out=64, fold=False):
    """Conv weight [Cout,Cin,*k] (+ eval-mode BatchNorm (gamma, beta, mean, var)) -> (fp16 [Np, Kp] tap-major /
    channel-minor with the BN scale folded in, fp32 [Np] shift).  Np = Cout rounded up to `pad_out`, Kp to 64 (whole
    TMA boxes).  fold=True (pointwise conv with 8/16/32 input channels) also returns the row-folded twin
    (fp16 [g*Cout, 64] block-diagonal, fp32 [g*Cout]) documented in include/kvq_b200.h."""
    w = w.detach().float()
    cout = w.shape[0]
    if bn is not None:
        g, b, m, v = [t.detach().float() for t in bn]
        scale = g / torch.sqrt(v + eps)
        shift = b - m * scale
        w = w * scale.reshape(-1, *([1] * (w.dim() - 1)))
    else:
        shift = torch.zeros(cout, device=w.device)
    perm = (0,) + tuple(range(2, w.dim())) + (1,)
    w2 = w.permute(*perm).reshape(cout, -1)
    K = w2.shape[1]
    Kp, Np = (K + 63) // 64 * 64, (cout + pad_out - 1) // pad_out * pad_out
    wp = torch.zeros((Np, Kp), dtype=torch.float32, device=w.device)
    wp[:cout, :K] = w2
    sp = torch.zeros(Np, dtype=torch.float32, device=w.device)
    sp[:cout] = shift
    packed = (cast_f16(wp.contiguous()), sp.contiguous())
    if not fold:
        return packed
    if K not in (8, 16, 32) or (64 // K * cout) % 64 != 0:
        raise RuntimeError(f"kvq_b200: cannot row-fold a [{cout},{K}] pointwise weight")
    g = 64 // K
    wf = torch.zeros((g * cout, 64), dtype=torch.float32, device=w.device)
    for j in range(g):
        wf[j * cout:(j + 1) * cout, j * K:(j + 1) * K] = w2
    return packed + (cast_f16(wf.contiguous()), shift.repeat(g).contiguous())


def pack_conv_image(w, bn=None, eps=1e-5):
    """Conv weight [Cout<=64, C in {8,16,32}, kt, kh, kw] (+ eval-mode BatchNorm) -> (fp16 images
    [kblocks, 4096] = the shared-memory B tiles of kvq_conv_narrow_f16, fp32 [64] shift)."""
    w = w.detach().float()
    if w.dim() == 4:
        w = w.unsqueeze(2)
    cout, C = w.shape[0], w.shape[1]
    if C not in (8, 16, 32) or cout > 64:
        raise RuntimeError(f"kvq_b200: pack_conv_image needs C in (8,16,32) and Cout <= 64, got {tuple(w.shape)}")
    if bn is not None:
        g, b, m, v = [t.detach().float() for t in bn]
        scale = g / torch.sqrt(v + eps)
        shift = b - m * scale
        w = w * scale.reshape(-1, 1, 1, 1, 1)
    else:
        shift = torch.zeros(cout, device=w.device)
    taps = w.shape[2] * w.shape[3] * w.shape[4]
    tpb = 64 // C
    nkb = (taps + tpb - 1) // tpb
    w2 = torch.zeros((64, nkb * tpb, C), dtype=torch.float32, device=w.device)
    w2[:cout, :taps] = w.permute(0, 2, 3, 4, 1).reshape(cout, taps, C)           # tap-major (dt, dh, dw), channel-minor
    n = torch.arange(64, device=w.device).view(64, 1, 1)
    t = torch.arange(nkb * tpb, device=w.device).view(1, -1, 1)
    k = torch.arange(C, device=w.device).view(1, 1, C)
    sw = {8: torch.zeros_like(n), 16: (n >> 2) & 1, 32: (n >> 1) & 3}[C]
    off = (t % tpb) * (64 * C) + n * C + (((k // 8) ^ sw) * 8) + (k % 8)          # in halfs, inside the K block
    img = torch.zeros((nkb, 4096), dtype=torch.float32, device=w.device)
    img.view(-1)[((t // tpb) * 4096 + off).reshape(-1)] = w2.reshape(-1)
    sp = torch.zeros(64, dtype=torch.float32, device=w.device)
    sp[:cout] = shift
    return cast_f16(img.contiguous()), sp.contiguous()


def conv_narrow_f16(x, w_image, bias, kernel, stride, pad, cout, resid=None, relu=False):
    """x f16 [B,T,H,W,C] with C in (8,16,32) -> f16 [B*To*Ho*Wo, cout] through the narrow implicit GEMM."""
    _need_cuda(x, w_image, bias, resid)
    B, T, H, W, C = x.shape
    od = [conv_out_size(n, k, s, p) for n, k, s, p in zip((T, H, W), kernel, stride, pad)]
    out = torch.empty((B * od[0] * od[1] * od[2], cout), dtype=torch.float16, device=x.device)
    rc = _l.load().kvq_conv_narrow_f16(_p(x), _p(w_image), _p(bias), _p(resid),
                                       resid.stride(0) if resid is not None else 0, _p(out), out.stride(0), B, T, H, W, C,
                                       _i3(kernel), _i3(stride), _i3(pad), int(cout), int(bool(relu)), _stream())
    _l.check(rc, "conv_narrow_f16")
    return out, tuple(od)


def pack_stem_weight(w, bn=None, eps=1e-5):
    """Stem conv weight [Cout,3,kt,7,7] (+ eval-mode BatchNorm) -> (fp16 [kt*3, rows, 64] in the layout
    kvq_stem_conv_f16 documents, fp32 [rows] shift)."""
    w = w.detach().float()
    if w.dim() == 4:
        w = w.unsqueeze(2)
    cout, cin, kt, kh, kw = w.shape
    if cin != 3 or kh != 7 or kw != 7:
        raise RuntimeError(f"kvq_b200: stem weight {tuple(w.shape)} is not [Cout,3,kt,7,7]")
    if bn is not None:
        g, b, m, v = [t.detach().float() for t in bn]
        scale = g / torch.sqrt(v + eps)
        shift = b - m * scale
        w = w * scale.reshape(-1, 1, 1, 1, 1)
    else:
        shift = torch.zeros(cout, device=w.device)
    rows = _l.load().kvq_stem_weight_rows(cout)
    wp = torch.zeros((kt, 3, rows, 8, 8), dtype=torch.float32, device=w.device)
    wp[:, :, :cout, :7, :7] = w.permute(2, 1, 0, 3, 4)
    sp = torch.zeros(rows, dtype=torch.float32, device=w.device)
    sp[:cout] = shift
    return cast_f16(wp.reshape(kt * 3, rows, 64).contiguous()), sp.contiguous()


def stem_conv_f16(x, w_packed, shift, kt, cout):
    """x f32 [N,3,T,H,W] -> f16 [N,T,Hs,Ws,cout] = relu(conv(kt,7,7)/s(1,2,2)/p(kt//2,3,3) + shift)."""
    _need_cuda(x, w_packed, shift)
    x = x.contiguous()
    N, _, T, H, W = x.shape
    Hs, Ws = conv_out_size(H, 7, 2, 3), conv_out_size(W, 7, 2, 3)
    out = torch.empty((N, T, Hs, Ws, cout), dtype=torch.float16, device=x.device)
    _l.check(_l.load().kvq_stem_conv_f16(_p(x), _p(w_packed), _p(shift), _p(out), N, T, H, W, int(kt), int(cout),
                                         _stream()), "stem_conv_f16")
    return out


def conv_gemm_f16(a, w, bias=None, resid=None, relu=False, nvalid=0, out=None):
    """out[M, :nvalid] = act(a[M,K] @ w[N,K]^T + bias + resid); a / resid / out fp16 row-major (may be column slices)."""
    M, K = a.shape
    N = w.shape[0]
    nv = nvalid if nvalid else N
    if out is None:
        out = torch.empty((M, nv), dtype=torch.float16, device=a.device)
    for t in (a, out, resid):
        if t is not None and (not t.is_cuda or t.stride(1) != 1):
            raise RuntimeError("kvq_b200: conv_gemm_f16 takes row-major CUDA tensors (no CPU fallback exists)")
    _need_cuda(w, bias)
    rc = _l.load().kvq_conv_gemm_f16(_p(a), a.stride(0), _p(w), _p(bias), _p(resid),
                                     resid.stride(0) if resid is not None else 0, _p(out), out.stride(0), M, N, K,
                                     int(nvalid), int(bool(relu)), _stream())
    _l.check(rc, "conv_gemm_f16")
    return out


def conv_implicit_f16(x, w, bias, kernel, stride, pad, resid=None, relu=False, nvalid=0):
    """x f16 [B,T,H,W,C] (C % 64 == 0), w f16 [N, kt*kh*kw*C] -> f16 [B*To*Ho*Wo, nvalid or N] (implicit GEMM)."""
    _need_cuda(x, w, bias, resid)
    B, T, H, W, C = x.shape
    od = [conv_out_size(n, k, s, p) for n, k, s, p in zip((T, H, W), kernel, stride, pad)]
    N = w.shape[0]
    nv = nvalid if nvalid else N
    out = torch.empty((B * od[0] * od[1] * od[2], nv), dtype=torch.float16, device=x.device)
    rc = _l.load().kvq_conv_implicit_f16(_p(x), _p(w), _p(bias), _p(resid), resid.stride(0) if resid is not None else 0,
                                         _p(out), out.stride(0), B, T, H, W, C, _i3(kernel), _i3(stride), _i3(pad), N,
                                         int(nvalid), int(bool(relu)), _stream())
    _l.check(rc, "conv_implicit_f16")
    return out, tuple(od)


def conv_out_size(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def im2col_cl_f16(x, kernel, stride, pad, Kp=None):
    """x f16 [B,T,H,W,C] -> (patch matrix f16 [B*To*Ho*Wo, Kp], (To,Ho,Wo))."""
    _need_cuda(x)
    B, T, H, W, C = x.shape
    od = [conv_out_size(n, k, s, p) for n, k, s, p in zip((T, H, W), kernel, stride, pad)]
    K = kernel[0] * kernel[1] * kernel[2] * C
    Kp = Kp or (K + 7) // 8 * 8
    out = torch.empty((B * od[0] * od[1] * od[2], Kp), dtype=torch.float16, device=x.device)
    _l.check(_l.load().kvq_im2col_cl_f16(_p(x), _p(out), B, T, H, W, C, _i3(kernel), _i3(stride), _i3(pad), Kp,
                                         _stream()), "im2col_cl_f16")
    return out, tuple(od)


def im2col_stem_f32(x, kernel, stride, pad, Kp=None):
    """x f32 [N,3,T,H,W] -> (patch matrix f16 [N*To*Ho*Wo, Kp], (To,Ho,Wo))."""
    _need_cuda(x)
    N, C, T, H, W = x.shape
    if C != 3:
        raise RuntimeError("kvq_b200: im2col_stem_f32 is the 3-channel network-input gather")
    od = [conv_out_size(n, k, s, p) for n, k, s, p in zip((T, H, W), kernel, stride, pad)]
    K = kernel[0] * kernel[1] * kernel[2] * 3
    Kp = Kp or (K + 7) // 8 * 8
    out = torch.empty((N * od[0] * od[1] * od[2], Kp), dtype=torch.float16, device=x.device)
    _l.check(_l.load().kvq_im2col_stem_f32(_p(x), _p(out), N, T, H, W, _i3(kernel), _i3(stride), _i3(pad), Kp,
                                           _stream()), "im2col_stem_f32")
    return out, tuple(od)


def maxpool_hw_f16(x):
    """nn.MaxPool2d(3, 2, 1) on channels-last f16 [N,H,W,C]."""
    _need_cuda(x)
    N, H, W, C = x.shape
    out = torch.empty((N, conv_out_size(H, 3, 2, 1), conv_out_size(W, 3, 2, 1), C), dtype=torch.float16,
                      device=x.device)
    _l.check(_l.load().kvq_maxpool_hw_f16(_p(x), _p(out), N, H, W, C, _stream()), "maxpool_hw_f16")
    return out


def pool_stats_f16(x, weights=None, want_std=True):
    """x f16 [N,L,C] -> (mean f32 [N,C], unbiased std f32 [N,C] or None); weights f32 [L] = weighted sum instead."""
    _need_cuda(x, weights)
    N, L, C = x.shape
    mean = torch.empty((N, C), dtype=torch.float32, device=x.device)
    std = torch.empty((N, C), dtype=torch.float32, device=x.device) if want_std else None
    _l.check(_l.load().kvq_pool_stats_f16(_p(x), _p(weights), _p(mean), _p(std), N, L, C, C, _stream()),
             "pool_stats_f16")
    return mean, std


def rowdot_mean_f32(x, w, b, group):
    """score[g] = mean over `group` consecutive rows of (x[row] . w) + b; x f32 [rows,K], w f32 [K], b f32 [1]."""
    _need_cuda(x, w, b)
    rows, K = x.shape
    score = torch.empty(rows // group, dtype=torch.float32, device=x.device)
    _l.check(_l.load().kvq_rowdot_mean_f32(_p(x), _p(w), _p(b), _p(score), rows, K, group, _stream()),
             "rowdot_mean_f32")
    return score


def fold_simplevqa_head(sd, prefix):
    """simpleVQAHead.quality = Linear(F,128) -> Linear(128,1) with nothing in between (head.py:22-26): one affine map."""
    w0, b0 = sd[prefix + "quality.0.weight"].detach().double(), sd[prefix + "quality.0.bias"].detach().double()
    w1, b1 = sd[prefix + "quality.1.weight"].detach().double(), sd[prefix + "quality.1.bias"].detach().double()
    return (w1 @ w0).reshape(-1).float().contiguous(), (w1 @ b0 + b1).reshape(1).float().contiguous()


class SimpleVQAWeights:
    """Device-resident packed weights of the SimpleVQA ResNet-50 (+ simpleVQAHead) in the order include/kvq_b200.h
    documents.  `sd` maps the REFERENCE state_dict names (simpleVQA_model.py:129-218; head.py:10-31) to tensors."""

    def __init__(self, sd, device, prefix="", head_prefix=None, layers=(3, 4, 6, 3), feat3d_dim=2304, eps=1e-5):
        self.device = torch.device(device)
        self.cfg = _l.KvqResNetConfig()
        for i, n in enumerate(layers):
            self.cfg.layers[i] = int(n)
        self.cfg.feat3d_dim = int(feat3d_dim)
        self.cfg.head = int(head_prefix is not None)

        def t(k):
            return sd[k].detach().to(self.device, torch.float32)

        def conv_bn(conv, bn):
            return list(pack_conv_weight(t(conv + ".weight"), [t(bn + "." + leaf) for leaf in
                                                              ("weight", "bias", "running_mean", "running_var")], eps))

        def stem_bn(conv, bn):
            return list(pack_stem_weight(t(conv + ".weight"), [t(bn + "." + leaf) for leaf in
                                                              ("weight", "bias", "running_mean", "running_var")], eps))

        ts = stem_bn(prefix + "conv1", prefix + "bn1")
        for s, depth in enumerate(layers):
            for j in range(depth):
                b = f"{prefix}layer{s + 1}.{j}."
                ts += conv_bn(b + "conv1", b + "bn1") + conv_bn(b + "conv2", b + "bn2") + conv_bn(b + "conv3", b + "bn3")
                if j == 0:
                    ts += conv_bn(b + "downsample.0", b + "downsample.1")
        if head_prefix is not None:
            dsd = {k: v.to(self.device) for k, v in sd.items() if k.startswith(head_prefix)}
            ts += list(fold_simplevqa_head(dsd, head_prefix))
        self.tensors = ts
        n = _l.load().kvq_resnet_num_weights(ctypes.byref(self.cfg))
        if n != len(ts):
            raise RuntimeError(f"kvq_b200: ResNet weight table has {len(ts)} entries, library expects {n}")
        self.ptrs = (ctypes.c_void_p * n)(*[x.data_ptr() for x in ts])
        self.feature_dim = _l.load().kvq_resnet_feature_dim(ctypes.byref(self.cfg))
        self._ws = None
        self._graphs = {}

    def workspace(self, B, T, H, W):
        need = _l.load().kvq_simplevqa_workspace_bytes(ctypes.byref(self.cfg), B, T, H, W)
        if need == 0:
            raise RuntimeError(f"kvq_b200: cannot plan a [{B},3,{T},{H},{W}] SimpleVQA forward: {_l.last_error()}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._graphs = {}
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _launch(self, x, feat3d, feats, score, ws):
        B, _, T, H, W = x.shape
        rc = _l.load().kvq_simplevqa_forward(ctypes.byref(self.cfg), self.ptrs, len(self.tensors), _p(x), _p(feat3d),
                                             B, T, H, W, _p(feats), _p(score), _p(ws), ws.numel(), _stream())
        _l.check(rc, "simplevqa_forward")

    def forward(self, x, feat3d, graph=None):
        """x f32 [B,3,T,H,W], feat3d f32 [B,T,feat3d_dim] -> (feats f32 [B,T,feature_dim], score f32 [B] or None)."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("kvq_b200: input frames must be float32 CUDA tensors (no CPU fallback exists)")
        x = x.contiguous()
        B, _, T, H, W = x.shape
        if self.cfg.feat3d_dim > 0:
            feat3d = feat3d.to(x.device, torch.float32).contiguous()
            if tuple(feat3d.shape) != (B, T, self.cfg.feat3d_dim):
                raise RuntimeError(f"kvq_b200: batch['feat'] is {tuple(feat3d.shape)}, expected "
                                   f"{(B, T, self.cfg.feat3d_dim)}")
        else:
            feat3d = None
        ws = self.workspace(B, T, H, W)
        if graph is None:
            graph = os.environ.get("KVQ_CUDA_GRAPH", "0") == "1"

        def alloc():
            return (torch.empty((B, T, self.feature_dim), dtype=torch.float32, device=x.device),
                    torch.empty(B, dtype=torch.float32, device=x.device) if self.cfg.head else None)

        if graph:
            # the SlowFast features are copied into a graph-owned staging buffer, so a fresh batch['feat'] tensor per video
            # neither re-captures the graph nor pins the old tensor
            key = (x.data_ptr(), tuple(x.shape), ws.data_ptr())
            entry = self._graphs.get(key)
            if entry is None:
                feats, score = alloc()
                stage = feat3d.clone() if feat3d is not None else None
                self._launch(x, stage, feats, score, ws)
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = _l.load().kvq_launch_count()
                with torch.cuda.graph(g):
                    self._launch(x, stage, feats, score, ws)
                nodes = int(_l.load().kvq_launch_count() - n0)
                if len(self._graphs) >= 8:
                    self._graphs.pop(next(iter(self._graphs)))
                entry = self._graphs[key] = (g, feats, score, nodes, stage)
            if entry[4] is not None:
                entry[4].copy_(feat3d)
            entry[0].replay()
            global GRAPH_KERNEL_LAUNCHES
            GRAPH_KERNEL_LAUNCHES += entry[3]
            return entry[1], entry[2]
        feats, score = alloc()
        self._launch(x, feat3d, feats, score, ws)
        return feats, score


# ---------------------------------------------------------------------------------------------------------------
# SlowFast-R50 motion features (SlowFast_features.py:112-165)
# ---------------------------------------------------------------------------------------------------------------
def slow_frame_indices(T, alpha=4):
    """torch.linspace(0, T-1, T // alpha).long() (SlowFast_features.py:127-131), evaluated by the library."""
    buf = (ctypes.c_int32 * 64)()
    n = _l.load().kvq_slow_frame_indices(int(T), int(alpha), buf, 64)
    if n < 0:
        _l.check(n, "slow_frame_indices")
    return [int(buf[i]) for i in range(n)]


def pack_pathway_slow(frames, alpha=4, out=None):
    """frames f32 [B,3,T,H,W] (CUDA) -> slow pathway [B,3,T//alpha,H,W]."""
    _need_cuda(frames)
    frames = frames.contiguous()
    B, C, T, H, W = frames.shape
    if C != 3 or frames.dtype != torch.float32:
        raise RuntimeError("kvq_b200: pack_pathway_slow takes float32 [B,3,T,H,W] frames")
    if out is None:
        out = torch.empty((B, 3, T // alpha, H, W), dtype=torch.float32, device=frames.device)
    elif tuple(out.shape) != (B, 3, T // alpha, H, W) or out.dtype != torch.float32 or not out.is_contiguous():
        raise RuntimeError("pack_pathway_slow: out must be a contiguous float32 [B,3,T//alpha,H,W] tensor")
    _l.check(_l.load().kvq_pack_pathway_slow_f32(_p(frames), _p(out), B, T, H, W, int(alpha), _stream()),
             "pack_pathway_slow_f32")
    return out


class SlowFastWeights:
    """Device-resident packed weights of the SlowFast-R50 trunk in the order include/kvq_b200.h documents.  `sd` maps
    the state_dict names of the reference's `slowfast` module (`feature_extraction.<block>.` + pytorchvideo's names,
    oracle/slowfast.py) to tensors."""

    DEPTHS = (3, 4, 6, 3)

    def __init__(self, sd, device, prefix="feature_extraction.", depths=DEPTHS, alpha=4, slow_pool=(8, 7, 7),
                 fast_pool=(32, 7, 7), eps=1e-5):
        self.device = torch.device(device)
        self.cfg = _l.KvqSlowFastConfig()
        for i, n in enumerate(depths):
            self.cfg.depths[i] = int(n)
        self.cfg.alpha = int(alpha)
        for i in range(3):
            self.cfg.slow_pool[i] = int(slow_pool[i])
            self.cfg.fast_pool[i] = int(fast_pool[i])

        def t(k):
            return sd[k].detach().to(self.device, torch.float32)

        def conv_bn(conv, bn, fold=False, image=False):
            stats = [t(bn + "." + leaf) for leaf in ("weight", "bias", "running_mean", "running_var")]
            out = list(pack_conv_weight(t(conv + ".weight"), stats, eps, fold=fold))
            if image:       # narrow-channel implicit GEMM twin (include/kvq_b200.h)
                out += list(pack_conv_image(t(conv + ".weight"), stats, eps))
            return out

        def stem_bn(conv, bn):
            return list(pack_stem_weight(t(conv + ".weight"), [t(bn + "." + leaf) for leaf in
                                                              ("weight", "bias", "running_mean", "running_var")], eps))

        p = prefix + "0."
        ts = stem_bn(p + "multipathway_blocks.0.conv", p + "multipathway_blocks.0.norm")
        ts += stem_bn(p + "multipathway_blocks.1.conv", p + "multipathway_blocks.1.norm")
        ts += conv_bn(p + "multipathway_fusion.conv_fast_to_slow", p + "multipathway_fusion.norm", image=True)
        for s, depth in enumerate(depths):
            p = f"{prefix}{s + 1}."
            for path in (0, 1):
                for j in range(depth):
                    b = f"{p}multipathway_blocks.{path}.res_blocks.{j}."
                    inner = (8 << s) if path == 1 else (64 << s)
                    cin = inner * 4 if j > 0 else ((80 if s == 0 else (64 << (s - 1)) * 4 + (8 << (s - 1)) * 8)
                                                   if path == 0 else (8 if s == 0 else (8 << (s - 1)) * 4))
                    narrow_in = path == 1 and cin in (8, 16, 32)
                    if j == 0:      # stride 1 (stage 0): row-folded twin; stride 2: narrow-conv image
                        ts += conv_bn(b + "branch1_conv", b + "branch1_norm", fold=narrow_in and s == 0,
                                      image=narrow_in and s > 0)
                    ts += conv_bn(b + "branch2.conv_a", b + "branch2.norm_a", image=narrow_in)
                    ts += conv_bn(b + "branch2.conv_b", b + "branch2.norm_b", image=path == 1 and inner in (8, 16, 32))
                    ts += conv_bn(b + "branch2.conv_c", b + "branch2.norm_c", fold=inner in (8, 16, 32))
            if s < 3:
                ts += conv_bn(p + "multipathway_fusion.conv_fast_to_slow", p + "multipathway_fusion.norm",
                              image=(8 << s) * 4 in (8, 16, 32))
        self.tensors = ts
        n = _l.load().kvq_slowfast_num_weights(ctypes.byref(self.cfg))
        if n != len(ts):
            raise RuntimeError(f"kvq_b200: SlowFast weight table has {len(ts)} entries, library expects {n}")
        self.ptrs = (ctypes.c_void_p * n)(*[x.data_ptr() for x in ts])
        self.slow_dim, self.fast_dim = 2048, 256
        self._ws = None
        self._graphs = {}

    def workspace(self, B, Ts, Tf, H, W):
        need = _l.load().kvq_slowfast_workspace_bytes(ctypes.byref(self.cfg), B, Ts, Tf, H, W)
        if need == 0:
            raise RuntimeError(f"kvq_b200: cannot plan a SlowFast forward on slow [{B},3,{Ts},{H},{W}] / fast "
                               f"[{B},3,{Tf},{H},{W}]: {_l.last_error()}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._graphs = {}
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _launch(self, slow, fast, so, fo, ws):
        B, _, Ts, H, W = slow.shape
        rc = _l.load().kvq_slowfast_forward(ctypes.byref(self.cfg), self.ptrs, len(self.tensors), _p(slow), _p(fast),
                                            B, Ts, fast.shape[2], H, W, _p(so), _p(fo), _p(ws), ws.numel(), _stream())
        _l.check(rc, "slowfast_forward")

    def forward(self, slow, fast, graph=None):
        """slow f32 [B,3,Ts,H,W], fast f32 [B,3,Tf,H,W] -> (slow_feature f32 [B,2048], fast_feature f32 [B,256])."""
        for x in (slow, fast):
            if not x.is_cuda or x.dtype != torch.float32:
                raise RuntimeError("kvq_b200: SlowFast pathways must be float32 CUDA tensors (no CPU fallback exists)")
        slow, fast = slow.contiguous(), fast.contiguous()
        B, _, Ts, H, W = slow.shape
        if fast.shape[0] != B or tuple(fast.shape[3:]) != (H, W) or slow.shape[1] != 3 or fast.shape[1] != 3:
            raise RuntimeError(f"kvq_b200: pathway shapes {tuple(slow.shape)} / {tuple(fast.shape)} do not match")
        ws = self.workspace(B, Ts, fast.shape[2], H, W)
        if graph is None:
            graph = os.environ.get("KVQ_CUDA_GRAPH", "0") == "1"

        def alloc():
            return (torch.empty((B, self.slow_dim), dtype=torch.float32, device=slow.device),
                    torch.empty((B, self.fast_dim), dtype=torch.float32, device=slow.device))

        if graph:
            key = (slow.data_ptr(), fast.data_ptr(), tuple(slow.shape), tuple(fast.shape), ws.data_ptr())
            entry = self._graphs.get(key)
            if entry is None:
                so, fo = alloc()
                self._launch(slow, fast, so, fo, ws)
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = _l.load().kvq_launch_count()
                with torch.cuda.graph(g):
                    self._launch(slow, fast, so, fo, ws)
                nodes = int(_l.load().kvq_launch_count() - n0)
                if len(self._graphs) >= 8:
                    self._graphs.pop(next(iter(self._graphs)))
                entry = self._graphs[key] = (g, so, fo, nodes)
            entry[0].replay()
            global GRAPH_KERNEL_LAUNCHES
            GRAPH_KERNEL_LAUNCHES += entry[3]
            return entry[1], entry[2]
        so, fo = alloc()
        self._launch(slow, fast, so, fo, ws)
        return so, fo


# ---------------------------------------------------------------------------------------------------------------
# CONTRIQUE distortion encoder of KSVQE (KSVQE_model.py:1622-1665)
# ---------------------------------------------------------------------------------------------------------------
class ContriqueWeights:
    """Packed weights of CONTRIQUE_model (torchvision ResNet-50 trunk `encoder.<i>.` + `projector.`); `sd` maps the
    reference state_dict names (e.g. `distortion_tool.encoder.0.weight`) to tensors."""

    def __init__(self, sd, device, prefix="distortion_tool.", eps=1e-5):
        self.device = torch.device(device)

        def t(k):
            return sd[prefix + k].detach().to(self.device, torch.float32)

        def bn(p):
            return [t(p + "." + leaf) for leaf in ("weight", "bias", "running_mean", "running_var")]

        ts = list(pack_stem_weight(t("encoder.0.weight"), bn("encoder.1"), eps))
        for li, depth in enumerate((3, 4, 6, 3)):
            for j in range(depth):
                b = f"encoder.{4 + li}.{j}."
                for c in (1, 2, 3):
                    ts += list(pack_conv_weight(t(b + f"conv{c}.weight"), bn(b + f"bn{c}"), eps))
                if j == 0:
                    ts += list(pack_conv_weight(t(b + "downsample.0.weight"), bn(b + "downsample.1"), eps))
        ts += list(pack_conv_weight(t("projector.0.weight"), bn("projector.1"), eps))
        ts += list(pack_conv_weight(t("projector.3.weight"), bn("projector.4"), eps))
        self.tensors = ts
        n = _l.load().kvq_contrique_num_weights()
        if n != len(ts):
            raise RuntimeError(f"kvq_b200: CONTRIQUE weight table has {len(ts)} entries, library expects {n}")
        self.ptrs = (ctypes.c_void_p * n)(*[x.data_ptr() for x in ts])
        self._ws = None

    def forward(self, x, anchor=32, frame_step=2):
        """x f32 [B,3,T,H,W] (x_sel_ori) -> z f32 [B, T/frame_step, (H/anchor)*(W/anchor), 128]."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("kvq_b200: CONTRIQUE input must be a float32 CUDA tensor (no CPU fallback exists)")
        x = x.contiguous()
        B, _, T, H, W = x.shape
        need = _l.load().kvq_contrique_workspace_bytes(B, T, H, W, anchor, frame_step)
        if need == 0:
            raise RuntimeError(f"kvq_b200: cannot plan a CONTRIQUE forward on {tuple(x.shape)}: {_l.last_error()}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        z = torch.empty((B, T // frame_step, (H // anchor) * (W // anchor), 128), dtype=torch.float32, device=x.device)
        rc = _l.load().kvq_contrique_forward(self.ptrs, len(self.tensors), _p(x), B, T, H, W, anchor, frame_step, _p(z),
                                             _p(self._ws), self._ws.numel(), _stream())
        _l.check(rc, "contrique_forward")
        return z


# ---------------------------------------------------------------------------------------------------------------
# CLIP ViT-B/16 visual tower with CLS adapters (CLIP_backbone.py:156-202)
# ---------------------------------------------------------------------------------------------------------------
def resize_pos_embed(pos, src_grid, tgt_grid):
    """resize_pos_embed2d (CLIP_backbone.py:35-69): bicubic resize of the patch-grid part of positional_embedding
    [1 + s*s, C] to [1 + t*t, C]; weight preprocessing, done once at load time."""
    import torch.nn.functional as F
    if src_grid == tgt_grid:
        return pos
    cls, grid = pos[:1], pos[1:]
    grid = grid.t().reshape(1, -1, src_grid, src_grid)
    grid = F.interpolate(grid, size=(tgt_grid, tgt_grid), mode="bicubic", align_corners=False)
    return torch.cat([cls, grid.permute(0, 2, 3, 1).reshape(tgt_grid * tgt_grid, -1)], dim=0)


class ClipVisualWeights:
    """Packed weights of CLIP_extractor_addadapter_cls; `sd` maps the reference names (`CLIP_tool.visual.*`,
    `CLIP_tool.adapter_layer.*`) to tensors."""

    def __init__(self, sd, device, prefix="CLIP_tool.", image_size=112, adapter_from=8):
        self.device = torch.device(device)

        def t(k):
            return sd[prefix + k].detach().to(self.device, torch.float32).contiguous()

        v = "visual."
        conv = t(v + "conv1.weight")                                   # [768, 3, 16, 16]
        width, patch = conv.shape[0], conv.shape[-1]
        layers = 1 + max(int(k[len(prefix + v + "transformer.resblocks."):].split(".")[0]) for k in sd
                         if k.startswith(prefix + v + "transformer.resblocks."))
        self.cfg = _l.KvqClipConfig(width, width // 64, layers, patch, adapter_from)
        self.grid = image_size // patch
        pos = t(v + "positional_embedding")
        src = int(round((pos.shape[0] - 1) ** 0.5))
        ts = [cast_f16(conv.permute(0, 2, 3, 1).reshape(width, -1).contiguous()), t(v + "class_embedding"),
              resize_pos_embed(pos, src, self.grid).contiguous(), t(v + "ln_pre.weight"), t(v + "ln_pre.bias")]
        for i in range(layers):
            b = f"{v}transformer.resblocks.{i}."
            ts += [t(b + "ln_1.weight"), t(b + "ln_1.bias"), cast_f16(t(b + "attn.in_proj_weight")),
                   t(b + "attn.in_proj_bias"), cast_f16(t(b + "attn.out_proj.weight")), t(b + "attn.out_proj.bias"),
                   t(b + "ln_2.weight"), t(b + "ln_2.bias"), cast_f16(t(b + "mlp.c_fc.weight")), t(b + "mlp.c_fc.bias"),
                   cast_f16(t(b + "mlp.c_proj.weight")), t(b + "mlp.c_proj.bias")]
            if i >= adapter_from:
                a = f"adapter_layer.{i - adapter_from}."
                ts += [t(a + "0.weight"), t(a + "0.bias"), t(a + "2.weight"), t(a + "2.bias")]
        self.tensors = ts
        n = _l.load().kvq_clip_num_weights(ctypes.byref(self.cfg))
        if n != len(ts):
            raise RuntimeError(f"kvq_b200: CLIP weight table has {len(ts)} entries, library expects {n}")
        self.ptrs = (ctypes.c_void_p * n)(*[x.data_ptr() for x in ts])
        self._ws = None

    def forward(self, images):
        """images f32 [n,3,H,W] -> (cls_attn f32 [n, g*g], tokens f32 [n, 1 + g*g, 768])."""
        if not images.is_cuda or images.dtype != torch.float32:
            raise RuntimeError("kvq_b200: CLIP input must be a float32 CUDA tensor (no CPU fallback exists)")
        images = images.contiguous()
        n, _, H, W = images.shape
        if H // self.cfg.patch != self.grid:
            raise RuntimeError(f"kvq_b200: positional embedding was resized for a {self.grid}x{self.grid} grid")
        need = _l.load().kvq_clip_workspace_bytes(ctypes.byref(self.cfg), n, H, W)
        if need == 0:
            raise RuntimeError(f"kvq_b200: cannot plan a CLIP forward on {tuple(images.shape)}: {_l.last_error()}")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        L = 1 + self.grid * self.grid
        attn = torch.empty((n, L - 1), dtype=torch.float32, device=images.device)
        tokens = torch.empty((n, L, self.cfg.width), dtype=torch.float32, device=images.device)
        rc = _l.load().kvq_clip_visual_forward(ctypes.byref(self.cfg), self.ptrs, len(self.tensors), _p(images), n, H, W,
                                               _p(attn), _p(tokens), _p(self._ws), self._ws.numel(), _stream())
        _l.check(rc, "clip_visual_forward")
        return attn, tokens
