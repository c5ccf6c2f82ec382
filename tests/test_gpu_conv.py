"""GPU parity tests of the convolution building blocks (channels-last fp16) against fp32 torch on the same
fp16-rounded operands.  All calls go through libkvq_b200.so."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("M,N,K,resid,relu,nvalid", [
    (1000, 64, 152, False, True, 0),      # stem: K tail zero-filled by TMA
    (3000, 256, 64, True, True, 0),       # bottleneck conv3 + identity
    (777, 192, 576, False, False, 0),     # BLOCK_N = 192, no activation (downsample branch)
    (128 * 150 + 3, 64, 64, True, True, 0),   # more tiles than SMs
    (500, 64, 72, False, True, 8),        # padded output channels: only 8 valid columns are stored
    (640, 2048, 512, True, True, 0),      # many N blocks
    (128 * 160 + 5, 256, 64, True, True, 0),      # >= 1 tile per SM at BLOCK_N = 256
    (128 * 80, 512, 576, False, True, 0),         # BLOCK_N = 256, two N blocks, K tail
    (128 * 100 + 77, 256, 128, True, False, 0),   # BLOCK_N = 128
    (128 * 150, 384, 192, False, True, 0),        # BLOCK_N = 192
])
def test_conv_gemm(M, N, K, resid, relu, nvalid):
    from kvq_b200 import ops
    a = _rand((M, K), 1).half()
    w = (_rand((N, K), 2) / math.sqrt(K)).half()
    b = _rand((N,), 3, 0.1)
    nv = nvalid or N
    r = _rand((M, nv), 4).half() if resid else None
    ref = a.float() @ w.float().t() + b
    ref = ref[:, :nv]
    if resid:
        ref = ref + r.float()
    if relu:
        ref = F.relu(ref)
    out = torch.full((M, nv + 8), 7.0, dtype=torch.float16, device=DEV)       # column slice: guard columns stay 7
    ops.conv_gemm_f16(a.to(DEV), w.to(DEV), b.to(DEV), r.to(DEV) if resid else None, relu=relu, nvalid=nvalid,
                      out=out[:, :nv])
    got = out.float().cpu()
    assert (got[:, nv:] == 7.0).all()
    assert (got[:, :nv] - ref).abs().max().item() < 8e-3, (got[:, :nv] - ref).abs().max().item()


@pytest.mark.parametrize("B,T,H,W,C,kernel,stride,pad", [
    (2, 1, 14, 14, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),     # ResNet 3x3
    (3, 1, 15, 11, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1)),    # stride-2 3x3 on odd sizes
    (2, 1, 9, 13, 256, (1, 1, 1), (1, 2, 2), (0, 0, 0)),     # strided 1x1 gather (downsample)
    (1, 8, 6, 7, 16, (3, 1, 1), (1, 1, 1), (1, 0, 0)),       # SlowFast temporal conv
    (1, 16, 6, 6, 8, (5, 1, 1), (4, 1, 1), (2, 0, 0)),       # SlowFast fast->slow lateral (alpha = 4)
    (1, 4, 5, 5, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1)),        # full 3-D
])
def test_im2col_cl_matches_conv3d(B, T, H, W, C, kernel, stride, pad):
    """(im2col @ W^T) must equal F.conv3d with W permuted tap-major / channel-minor -- exercises gather + layout."""
    from kvq_b200 import ops
    x = _rand((B, C, T, H, W), 5).half()
    wt = _rand((24, C) + kernel, 6).half()
    ref = F.conv3d(x.float(), wt.float(), stride=stride, padding=pad)           # [B,24,To,Ho,Wo]
    xcl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    col, od = ops.im2col_cl_f16(xcl, kernel, stride, pad)
    assert od == tuple(ref.shape[2:])
    w2 = wt.permute(0, 2, 3, 4, 1).reshape(24, -1)
    got = (col.float().cpu()[:, :w2.shape[1]] @ w2.float().t()).view(B, *od, 24).permute(0, 4, 1, 2, 3)
    assert (got - ref).abs().max().item() < 1e-3 * math.sqrt(w2.shape[1])
    assert (col[:, w2.shape[1]:] == 0).all()


@pytest.mark.parametrize("N,T,H,W,kernel,stride,pad", [
    (3, 1, 33, 47, (1, 7, 7), (1, 2, 2), (0, 3, 3)),         # ResNet stem, per-frame (T folded into N by the caller)
    (1, 4, 32, 32, (1, 7, 7), (1, 2, 2), (0, 3, 3)),         # T axis of [B,3,T,H,W]
    (1, 8, 20, 20, (5, 7, 7), (1, 2, 2), (2, 3, 3)),         # SlowFast fast stem
])
def test_im2col_stem(N, T, H, W, kernel, stride, pad):
    from kvq_b200 import ops
    x = _rand((N, 3, T, H, W), 7)
    wt = _rand((16, 3) + kernel, 8).half()
    ref = F.conv3d(x.half().float(), wt.float(), stride=stride, padding=pad)
    col, od = ops.im2col_stem_f32(x.to(DEV), kernel, stride, pad)
    assert od == tuple(ref.shape[2:]) and col.shape[1] % 8 == 0
    w2 = wt.permute(0, 2, 3, 4, 1).reshape(16, -1)
    got = (col.float().cpu()[:, :w2.shape[1]] @ w2.float().t()).view(N, *od, 16).permute(0, 4, 1, 2, 3)
    assert (got - ref).abs().max().item() < 1e-3 * math.sqrt(w2.shape[1])


@pytest.mark.parametrize("N,H,W,C", [(2, 32, 32, 64), (3, 17, 23, 64), (1, 5, 4, 8)])
def test_maxpool(N, H, W, C):
    from kvq_b200 import ops
    x = _rand((N, C, H, W), 9).half()
    ref = F.max_pool2d(x.float(), 3, 2, 1)
    got = ops.maxpool_hw_f16(x.permute(0, 2, 3, 1).contiguous().to(DEV)).float().cpu().permute(0, 3, 1, 2)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("N,L,C", [(4, 49, 512), (3, 196, 2048), (2, 4, 1024), (1, 784, 264)])
def test_pool_stats(N, L, C):
    from kvq_b200 import ops
    x = (_rand((N, L, C), 10) + 0.5).relu().half()
    mean, std = ops.pool_stats_f16(x.to(DEV))
    assert (mean.cpu() - x.float().mean(dim=1)).abs().max().item() < 1e-5
    assert (std.cpu() - x.float().std(dim=1)).abs().max().item() < 1e-5
    wts = torch.rand(L, generator=torch.Generator().manual_seed(11))
    wmean, none = ops.pool_stats_f16(x.to(DEV), weights=wts.to(DEV), want_std=False)
    assert none is None
    assert (wmean.cpu() - (x.float() * wts.view(1, L, 1)).sum(dim=1)).abs().max().item() < 1e-4 * L ** 0.5


def test_rowdot_mean():
    from kvq_b200 import ops
    x, w, b = _rand((24, 9472), 12), _rand((9472,), 13, 0.01), _rand((1,), 14)
    ref = (x @ w + b).view(3, 8).mean(dim=1)
    got = ops.rowdot_mean_f32(x.to(DEV), w.to(DEV), b.to(DEV), group=8).cpu()
    assert (got - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("N,T,H,W,kt,cout", [
    (2, 3, 64, 64, 1, 64),        # ResNet / slow stem, whole tiles
    (1, 2, 37, 53, 1, 64),        # ragged frame: partial tiles in both directions
    (1, 8, 48, 80, 5, 8),         # fast stem: 5-frame halo ring, clip-boundary zero padding
    (2, 32, 40, 40, 5, 8),        # T split into segments across work items
    (1, 5, 224, 224, 5, 8),
])
def test_stem_conv_matches_conv3d(N, T, H, W, kt, cout):
    from kvq_b200 import ops
    x = _rand((N, 3, T, H, W), 21)
    w = _rand((cout, 3, kt, 7, 7), 22) / math.sqrt(3 * kt * 49)
    bn = [1.0 + 0.2 * _rand((cout,), 23), 0.1 * _rand((cout,), 24), 0.1 * _rand((cout,), 25),
          0.5 + torch.rand(cout, generator=torch.Generator().manual_seed(26))]
    wp, sp = ops.pack_stem_weight(w.to(DEV), [t.to(DEV) for t in bn])
    got = ops.stem_conv_f16(x.to(DEV), wp, sp, kt, cout).float().cpu()
    scale = bn[0] / torch.sqrt(bn[3] + 1e-5)
    wf = (w * scale.view(-1, 1, 1, 1, 1)).half().float()
    ref = F.relu(F.conv3d(x.half().float(), wf, bn[1] - bn[2] * scale, stride=(1, 2, 2), padding=(kt // 2, 3, 3)))
    ref = ref.permute(0, 2, 3, 4, 1)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 5e-3, (got - ref).abs().max().item()


@pytest.mark.parametrize("B,T,H,W,C,N,kernel,stride,pad,resid", [
    (2, 1, 16, 16, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),      # ResNet 3x3, whole tiles
    (3, 1, 14, 14, 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1), True),     # 14x14 map: padded tile grid, + residual
    (2, 1, 15, 11, 64, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1), False),      # stride 2 on odd sizes (TMA traversal stride)
    (2, 1, 9, 13, 256, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0), False),     # strided 1x1 (downsample branch)
    (1, 8, 7, 7, 128, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),       # SlowFast temporal conv_a, 7x7 map
    (2, 8, 8, 8, 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), True),
    (1, 32, 6, 6, 64, 128, (7, 1, 1), (4, 1, 1), (3, 0, 0), False),      # fast->slow lateral: temporal stride 4
    (1, 4, 9, 10, 64, 64, (3, 3, 3), (1, 2, 1), (1, 1, 1), False),       # full 3-D, mixed strides
    (16, 1, 56, 56, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),     # more tiles than SMs
])
def test_conv_implicit_matches_conv3d(B, T, H, W, C, N, kernel, stride, pad, resid):
    from kvq_b200 import ops
    x = _rand((B, C, T, H, W), 31).half()
    wt = (_rand((N, C) + kernel, 32) / math.sqrt(C * kernel[0] * kernel[1] * kernel[2])).half()
    bias = _rand((N,), 33, 0.1)
    ref = F.conv3d(x.float(), wt.float(), bias, stride=stride, padding=pad)          # [B,N,To,Ho,Wo]
    ref = ref.permute(0, 2, 3, 4, 1).reshape(-1, N)
    r = _rand(tuple(ref.shape), 34).half() if resid else None
    if resid:
        ref = ref + r.float()
    ref = F.relu(ref)
    xcl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    w2 = wt.permute(0, 2, 3, 4, 1).reshape(N, -1).contiguous().to(DEV)
    got, od = ops.conv_implicit_f16(xcl, w2, bias.to(DEV), kernel, stride, pad, resid=r.to(DEV) if resid else None,
                                    relu=True)
    assert got.shape == ref.shape
    err = (got.float().cpu() - ref).abs().max().item()
    assert err < 8e-3, err


@pytest.mark.parametrize("B,T,H,W,C,cout,kernel,stride,pad,resid", [
    (2, 4, 16, 16, 8, 8, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),       # fast res2 conv_b: 9 taps = 8 + 1 (partial K block)
    (1, 8, 8, 8, 8, 8, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),         # fast res2 conv_a block 0: 3 taps of 8
    (1, 8, 14, 14, 32, 8, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),      # conv_a on 32 channels: 64 B swizzle, padded tiles
    (2, 2, 15, 11, 16, 16, (1, 3, 3), (1, 2, 2), (0, 1, 1), False),     # res3 conv_b: 32 B swizzle, stride 2, odd sizes
    (2, 4, 12, 12, 32, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1), True),      # res4 conv_b + residual
    (1, 32, 6, 6, 8, 16, (7, 1, 1), (4, 1, 1), (3, 0, 0), False),       # lateral fusion after the stem
    (1, 16, 6, 7, 32, 64, (7, 1, 1), (4, 1, 1), (3, 0, 0), False),      # lateral fusion after res2 (64 outputs)
    (2, 2, 9, 13, 32, 64, (1, 1, 1), (1, 2, 2), (0, 0, 0), False),      # strided branch1 of fast res3
    (8, 32, 56, 56, 8, 8, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),      # many more tiles than SMs
])
def test_conv_narrow_matches_conv3d(B, T, H, W, C, cout, kernel, stride, pad, resid):
    from kvq_b200 import ops
    x = _rand((B, C, T, H, W), 41).half()
    wt = (_rand((cout, C) + kernel, 42) / math.sqrt(C * kernel[0] * kernel[1] * kernel[2])).half()
    bias = _rand((cout,), 43, 0.1)
    ref = F.conv3d(x.float(), wt.float(), bias, stride=stride, padding=pad)
    ref = ref.permute(0, 2, 3, 4, 1).reshape(-1, cout)
    r = _rand(tuple(ref.shape), 44).half() if resid else None
    if resid:
        ref = ref + r.float()
    ref = F.relu(ref)
    xcl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    img, sp = ops.pack_conv_image(wt.float().to(DEV))
    sp[:cout] = bias.to(DEV)
    got, od = ops.conv_narrow_f16(xcl, img, sp, kernel, stride, pad, cout, resid=r.to(DEV) if resid else None, relu=True)
    assert got.shape == ref.shape
    err = (got.float().cpu() - ref).abs().max().item()
    assert err < 8e-3, err
