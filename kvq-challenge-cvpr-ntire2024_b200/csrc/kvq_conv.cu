// Convolution-side row kernels for the SimpleVQA ResNet-50 spatial branch (models/backbones/simpleVQA_model.py:129-264)
// and the SlowFast motion trunk (SlowFast_features.py:26-77).  Activations are channels-last fp16 ([B,T,H,W,C]), so a
// 1x1 convolution IS the row GEMM of kvq_gemm.cu and a kxkxk convolution is an im2col gather (16-byte chunks, K index
// tap-major so a chunk never straddles two taps) followed by the same GEMM with the folded-BatchNorm / residual / ReLU
// epilogue (EPI_CONV_F16).  All kernels here are HBM-bound gathers / reductions: 16-byte accesses, grid-stride loops
// sized to the SM count.
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int CONV_THREADS = 256;

inline int grid_for(long long items, int threads, int per_sm = 8) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

struct ConvGeom {
  int B, T, H, W, C;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int To, Ho, Wo;
  int K, Kp;
};

// G lanes share one output row and stride over its 16-byte chunks (tap-major, channel-minor), so the row decode
// (three 32-bit divisions) is paid once per row and the (tap, channel) position advances by carries, not divisions.
// G is chosen by the host so that every lane moves several chunks per row: G*16 contiguous bytes per row per store.
template <int G>
__global__ void __launch_bounds__(CONV_THREADS)
im2col_cl_kernel(const __half* __restrict__ in, __half* __restrict__ out, const ConvGeom g, long long row0, int rows) {
  pdl_wait();
  const int kchunks = g.Kp >> 3, cpt = g.C >> 3, kvalid = g.K >> 3;
  const int lane_g = threadIdx.x % G;
  constexpr int RPB = CONV_THREADS / G;
  // (dt, dh, dw, c8) of this lane's first chunk: thread constants
  int tap0 = lane_g / cpt;
  const int c80 = lane_g - tap0 * cpt;
  const int dw0 = tap0 % g.kw; tap0 /= g.kw;
  const int dh0 = tap0 % g.kh;
  const int dt0 = tap0 / g.kh;
  for (int r = blockIdx.x * RPB + threadIdx.x / G; r < rows; r += gridDim.x * RPB) {
    const long long m = row0 + r;
    unsigned q = static_cast<unsigned>(m % (static_cast<long long>(g.To) * g.Ho * g.Wo));
    const int b = static_cast<int>(m / (static_cast<long long>(g.To) * g.Ho * g.Wo));
    const int wo = q % g.Wo; q /= g.Wo;
    const int ho = q % g.Ho;
    const int to = q / g.Ho;
    const int t0 = to * g.st - g.pt, h0 = ho * g.sh - g.ph, w0 = wo * g.sw - g.pw;
    const __half* src_b = in + static_cast<long long>(b) * g.T * g.H * g.W * g.C;
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<long long>(r) * g.Kp);
    int dt = dt0, dh = dh0, dw = dw0, c8 = c80;
    for (int j = lane_g; j < kchunks; j += G) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      const int t = t0 + dt, h = h0 + dh, w = w0 + dw;
      if (j < kvalid && t >= 0 && t < g.T && h >= 0 && h < g.H && w >= 0 && w < g.W)
        v = __ldg(reinterpret_cast<const uint4*>(src_b + ((static_cast<long long>(t) * g.H + h) * g.W + w) * g.C) + c8);
      dst[j] = v;
      c8 += G;
      while (c8 >= cpt) {
        c8 -= cpt;
        if (++dw == g.kw) { dw = 0; if (++dh == g.kh) { dh = 0; ++dt; } }
      }
    }
  }
  pdl_launch_dependents();
}

// stem: fp32 NCDHW input with 3 channels; one half2 (two consecutive k) per thread
__global__ void __launch_bounds__(CONV_THREADS)
im2col_stem_kernel(const float* __restrict__ in, __half* __restrict__ out, const ConvGeom g, long long row0,
                   long long pairs) {
  pdl_wait();
  const int kpairs = g.Kp >> 1;
  const long long plane = static_cast<long long>(g.T) * g.H * g.W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < pairs;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long ml = idx / kpairs;
    const long long m = row0 + ml;
    const int k0 = static_cast<int>(idx - ml * kpairs) << 1;
    long long r = m;
    const int wo = static_cast<int>(r % g.Wo); r /= g.Wo;
    const int ho = static_cast<int>(r % g.Ho); r /= g.Ho;
    const int to = static_cast<int>(r % g.To); r /= g.To;
    const int b = static_cast<int>(r);
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = k0 + e;
      if (k < g.K) {
        const int tap = k / 3, c = k - tap * 3;
        const int dw = tap % g.kw, dh = (tap / g.kw) % g.kh, dt = tap / (g.kw * g.kh);
        const int t = to * g.st - g.pt + dt, h = ho * g.sh - g.ph + dh, w = wo * g.sw - g.pw + dw;
        if (t >= 0 && t < g.T && h >= 0 && h < g.H && w >= 0 && w < g.W)
          v[e] = __ldg(in + (static_cast<long long>(b) * 3 + c) * plane + (static_cast<long long>(t) * g.H + h) * g.W + w);
      }
    }
    *reinterpret_cast<uint32_t*>(out + idx * 2) = pack_half2(v[0], v[1]);
  }
  pdl_launch_dependents();
}

__global__ void __launch_bounds__(CONV_THREADS)
maxpool_hw_kernel(const __half* __restrict__ in, __half* __restrict__ out, int N, int H, int W, int C, int Ho, int Wo,
                  int ldo, long long chunks) {
  pdl_wait();
  const int cch = C >> 3;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < chunks;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = idx;
    const int c = static_cast<int>(r % cch) << 3; r /= cch;
    const int wo = static_cast<int>(r % Wo); r /= Wo;
    const int ho = static_cast<int>(r % Ho); r /= Ho;
    const int n = static_cast<int>(r);
    __half2 best[4];
    bool any = false;
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int h = ho * 2 - 1 + dh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        const int w = wo * 2 - 1 + dw;
        if (w < 0 || w >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * H + h) * W + w) * C + c));
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
        if (!any) {
#pragma unroll
          for (int j = 0; j < 4; ++j) best[j] = hv[j];
          any = true;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) best[j] = __hmax2(best[j], hv[j]);
        }
      }
    }
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + ho) * Wo + wo) * ldo + c) =
        *reinterpret_cast<const uint4*>(best);
  }
  pdl_launch_dependents();
}

// block = (n, 64-channel slab): 8 channel groups of 8 channels x 32 partitions of the L axis (a warp reads four
// 128-byte row segments per load); two passes (mean, then centred sum of squares, the second one out of L2) so the
// unbiased std does not cancel in fp32.  64-channel slabs give N * C/64 blocks: 256+ for every pooled layer here
// (the 256-channel slabs of the first version left 64 blocks on 148 SMs).
constexpr int PS_CG = 8, PS_LP = CONV_THREADS / PS_CG;   // 8 channel groups x 32 L partitions
__global__ void __launch_bounds__(CONV_THREADS)
pool_stats_kernel(const __half* __restrict__ in, const float* __restrict__ weights, float* __restrict__ out_mean,
                  float* __restrict__ out_std, int L, int C, int ldo) {
  pdl_wait();
  __shared__ float red[PS_LP][PS_CG][8];
  __shared__ float mean_s[PS_CG][8];
  const int n = blockIdx.x;
  const int cg = threadIdx.x % PS_CG, lp = threadIdx.x / PS_CG;
  const int c = (blockIdx.y * PS_CG + cg) << 3;
  const bool ok = c < C;
  const __half* base = in + static_cast<long long>(n) * L * C + c;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ok) {
    for (int l = lp; l < L; l += PS_LP) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(l) * C));
      const __half2* hv = reinterpret_cast<const __half2*>(&v);
      const float wgt = weights != nullptr ? __ldg(weights + l) : 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hv[j]);
        acc[2 * j] = fmaf(f.x, wgt, acc[2 * j]);
        acc[2 * j + 1] = fmaf(f.y, wgt, acc[2 * j + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[lp][cg][j] = acc[j];
  __syncthreads();
  if (threadIdx.x < PS_CG * 8) {       // one thread per (channel group, channel)
    const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < PS_LP; ++q) s += red[q][g][j];
    s *= weights != nullptr ? 1.f : 1.f / static_cast<float>(L);
    mean_s[g][j] = s;
    const int cc = ((blockIdx.y * PS_CG + g) << 3) + j;
    if (cc < C) out_mean[static_cast<long long>(n) * ldo + cc] = s;
  }
  if (out_std == nullptr) { pdl_launch_dependents(); return; }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ok) {
    float mu[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mu[j] = mean_s[cg][j];
    for (int l = lp; l < L; l += PS_LP) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(l) * C));
      const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hv[j]);
        const float a = f.x - mu[2 * j], b = f.y - mu[2 * j + 1];
        acc[2 * j] = fmaf(a, a, acc[2 * j]);
        acc[2 * j + 1] = fmaf(b, b, acc[2 * j + 1]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) red[lp][cg][j] = acc[j];
  __syncthreads();
  if (threadIdx.x < PS_CG * 8) {
    const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < PS_LP; ++q) s += red[q][g][j];
    const int cc = ((blockIdx.y * PS_CG + g) << 3) + j;
    // torch.std default: unbiased (L - 1); L == 1 gives NaN there too
    if (cc < C) out_std[static_cast<long long>(n) * ldo + cc] = sqrtf(s / static_cast<float>(L - 1));
  }
  pdl_launch_dependents();
}

// block per group of `group` consecutive rows: score = mean_rows(dot(x_row, w)) + b
__global__ void __launch_bounds__(CONV_THREADS)
rowdot_mean_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ score, int K,
                   int group) {
  pdl_wait();
  __shared__ float red[CONV_THREADS / 32];
  const float* base = x + static_cast<long long>(blockIdx.x) * group * K;
  float acc = 0.f;
  const int k4 = K >> 2;
  for (int r = 0; r < group; ++r) {
    const float4* row = reinterpret_cast<const float4*>(base + static_cast<long long>(r) * K);
    for (int i = threadIdx.x; i < k4; i += blockDim.x) {
      const float4 a = __ldg(row + i);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + i);
      acc += (a.x * ww.x + a.y * ww.y) + (a.z * ww.z + a.w * ww.w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CONV_THREADS / 32; ++i) s += red[i];
    score[blockIdx.x] = s / static_cast<float>(group) + (b != nullptr ? __ldg(b) : 0.f);
  }
  pdl_launch_dependents();
}

// windows of an AvgPool(k, stride 1, no padding) over n positions that cover position p
__device__ __forceinline__ int pool_cover(int p, int n, int k) {
  const int lo = p - k + 1 > 0 ? p - k + 1 : 0, hi = p < n - k ? p : n - k;
  return hi - lo + 1;
}

__global__ void __launch_bounds__(CONV_THREADS)
pool_window_weights_kernel(float* __restrict__ out, int T, int H, int W, int kt, int kh, int kw) {
  pdl_wait();
  const int L = T * H * W;
  const float norm = 1.f / (static_cast<float>(T - kt + 1) * (H - kh + 1) * (W - kw + 1) * kt * kh * kw);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
    const int w = i % W, h = (i / W) % H, t = i / (W * H);
    out[i] = static_cast<float>(pool_cover(t, T, kt) * pool_cover(h, H, kh) * pool_cover(w, W, kw)) * norm;
  }
  pdl_launch_dependents();
}

struct FrameIdx { int v[64]; };

// out[bc, i, :] = in[bc, idx[i], :]; one float4 per thread (plane % 4 == 0)
__global__ void __launch_bounds__(CONV_THREADS)
select_frames_kernel(const float4* __restrict__ in, float4* __restrict__ out, int T, int n, long long plane4,
                     const FrameIdx idx, long long total4) {
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long frame = i / plane4, off = i - frame * plane4;
    const long long bc = frame / n;
    const int fi = static_cast<int>(frame - bc * n);
    out[i] = __ldg(in + (bc * T + idx.v[fi]) * plane4 + off);
  }
  pdl_launch_dependents();
}

int fill_geom(ConvGeom& g, int B, int T, int H, int W, int C, int kt, int kh, int kw, int st, int sh, int sw, int pt,
              int ph, int pw, int Kp) {
  KVQ_REQUIRE(B > 0 && T > 0 && H > 0 && W > 0 && C > 0 && kt > 0 && kh > 0 && kw > 0 && st > 0 && sh > 0 && sw > 0 &&
                  pt >= 0 && ph >= 0 && pw >= 0,
              KVQ_ERR_BAD_SHAPE, "im2col: bad geometry B=%d T=%d H=%d W=%d C=%d k=(%d,%d,%d) s=(%d,%d,%d)", B, T, H, W,
              C, kt, kh, kw, st, sh, sw);
  g = ConvGeom{B, T, H, W, C, kt, kh, kw, st, sh, sw, pt, ph, pw, 0, 0, 0, 0, 0};
  g.To = (T + 2 * pt - kt) / st + 1;
  g.Ho = (H + 2 * ph - kh) / sh + 1;
  g.Wo = (W + 2 * pw - kw) / sw + 1;
  g.K = kt * kh * kw * C;
  g.Kp = Kp;
  KVQ_REQUIRE(g.To > 0 && g.Ho > 0 && g.Wo > 0, KVQ_ERR_BAD_SHAPE, "im2col: empty output %dx%dx%d", g.To, g.Ho, g.Wo);
  KVQ_REQUIRE(Kp >= g.K && Kp % 8 == 0, KVQ_ERR_BAD_SHAPE, "im2col: Kp=%d must be a multiple of 8 and >= K=%d", Kp, g.K);
  return KVQ_OK;
}

}  // namespace

int launch_im2col_cl(const __half* in, __half* out, int B, int T, int H, int W, int C, int kt, int kh, int kw, int st,
                     int sh, int sw, int pt, int ph, int pw, int Kp, cudaStream_t stream, long long row0,
                     long long rows) {
  ConvGeom g;
  const int rc = fill_geom(g, B, T, H, W, C, kt, kh, kw, st, sh, sw, pt, ph, pw, Kp);
  if (rc != 0) return rc;
  KVQ_REQUIRE(C % 8 == 0, KVQ_ERR_MISALIGNED, "im2col: channels-last input needs C %% 8 == 0 (got %d)", C);
  const long long total = static_cast<long long>(B) * g.To * g.Ho * g.Wo;
  if (rows < 0) rows = total - row0;
  KVQ_REQUIRE(row0 >= 0 && rows > 0 && row0 + rows <= total, KVQ_ERR_BAD_SHAPE,
              "im2col: row range [%lld, +%lld) outside %lld output rows", row0, rows, total);
  KVQ_REQUIRE(rows < (1LL << 31), KVQ_ERR_BAD_SHAPE, "im2col: %lld rows in one launch", rows);
  const int kchunks = Kp / 8;
  const int G = kchunks >= 32 ? 8 : kchunks >= 16 ? 4 : kchunks >= 8 ? 2 : 1;
  const long long threads = rows * G;
  const int grid = grid_for(threads, CONV_THREADS);
  const int nrows = static_cast<int>(rows);
  count_launch();
  switch (G) {
    case 8: return launch_pdl(im2col_cl_kernel<8>, dim3(grid), dim3(CONV_THREADS), 0, stream, in, out, g, row0, nrows);
    case 4: return launch_pdl(im2col_cl_kernel<4>, dim3(grid), dim3(CONV_THREADS), 0, stream, in, out, g, row0, nrows);
    case 2: return launch_pdl(im2col_cl_kernel<2>, dim3(grid), dim3(CONV_THREADS), 0, stream, in, out, g, row0, nrows);
    default: return launch_pdl(im2col_cl_kernel<1>, dim3(grid), dim3(CONV_THREADS), 0, stream, in, out, g, row0, nrows);
  }
}

int launch_im2col_stem(const float* in, __half* out, int N, int T, int H, int W, int kt, int kh, int kw, int st, int sh,
                       int sw, int pt, int ph, int pw, int Kp, cudaStream_t stream, long long row0, long long rows) {
  ConvGeom g;
  const int rc = fill_geom(g, N, T, H, W, 3, kt, kh, kw, st, sh, sw, pt, ph, pw, Kp);
  if (rc != 0) return rc;
  const long long total = static_cast<long long>(N) * g.To * g.Ho * g.Wo;
  if (rows < 0) rows = total - row0;
  KVQ_REQUIRE(row0 >= 0 && rows > 0 && row0 + rows <= total, KVQ_ERR_BAD_SHAPE,
              "im2col: row range [%lld, +%lld) outside %lld output rows", row0, rows, total);
  const long long pairs = rows * (Kp / 2);
  count_launch();
  return launch_pdl(im2col_stem_kernel, dim3(grid_for(pairs, CONV_THREADS)), dim3(CONV_THREADS), 0, stream, in, out, g,
                    row0, pairs);
}

int launch_maxpool_hw(const __half* in, __half* out, int N, int H, int W, int C, cudaStream_t stream, int ldo) {
  if (ldo <= 0) ldo = C;
  KVQ_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && ldo >= C && ldo % 8 == 0, KVQ_ERR_BAD_SHAPE,
              "maxpool: bad shape N=%d H=%d W=%d C=%d ldo=%d (C, ldo %% 8 == 0)", N, H, W, C, ldo);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long chunks = static_cast<long long>(N) * Ho * Wo * (C / 8);
  count_launch();
  return launch_pdl(maxpool_hw_kernel, dim3(grid_for(chunks, CONV_THREADS)), dim3(CONV_THREADS), 0, stream, in, out, N,
                    H, W, C, Ho, Wo, ldo, chunks);
}

int launch_pool_stats(const __half* in, const float* weights, float* out_mean, float* out_std, int N, int L, int C,
                      int ldo, cudaStream_t stream) {
  KVQ_REQUIRE(N > 0 && L > 0 && C > 0 && C % 8 == 0 && ldo >= C, KVQ_ERR_BAD_SHAPE,
              "pool_stats: bad shape N=%d L=%d C=%d ldo=%d", N, L, C, ldo);
  count_launch();
  return launch_pdl(pool_stats_kernel, dim3(N, (C + 63) / 64), dim3(CONV_THREADS), 0, stream, in, weights, out_mean,
                    out_std, L, C, ldo);
}

int launch_rowdot_mean(const float* x, const float* w, const float* b, float* score, int rows, int K, int group,
                       cudaStream_t stream) {
  KVQ_REQUIRE(rows > 0 && K > 0 && K % 4 == 0 && group > 0 && rows % group == 0, KVQ_ERR_BAD_SHAPE,
              "rowdot_mean: rows=%d K=%d group=%d (K %% 4 == 0, rows %% group == 0)", rows, K, group);
  count_launch();
  return launch_pdl(rowdot_mean_kernel, dim3(rows / group), dim3(CONV_THREADS), 0, stream, x, w, b, score, K, group);
}

int launch_pool_window_weights(float* out, int T, int H, int W, int kt, int kh, int kw, cudaStream_t stream) {
  KVQ_REQUIRE(kt >= 1 && kh >= 1 && kw >= 1 && T >= kt && H >= kh && W >= kw, KVQ_ERR_BAD_SHAPE,
              "avg pool kernel (%d,%d,%d) larger than the %dx%dx%d map (AvgPool3d would raise)", kt, kh, kw, T, H, W);
  count_launch();
  return launch_pdl(pool_window_weights_kernel, dim3(grid_for(static_cast<long long>(T) * H * W, CONV_THREADS)),
                    dim3(CONV_THREADS), 0, stream, out, T, H, W, kt, kh, kw);
}

int launch_select_frames(const float* in, float* out, int BC, int T, long long plane, const int* idx, int n,
                         cudaStream_t stream) {
  KVQ_REQUIRE(BC > 0 && T > 0 && n > 0 && n <= 64 && plane > 0 && plane % 4 == 0, KVQ_ERR_BAD_SHAPE,
              "select_frames: BC=%d T=%d n=%d plane=%lld (n <= 64, H*W %% 4 == 0)", BC, T, n, plane);
  FrameIdx fi;
  for (int i = 0; i < n; ++i) {
    KVQ_REQUIRE(idx[i] >= 0 && idx[i] < T, KVQ_ERR_BAD_SHAPE, "select_frames: index %d outside [0,%d)", idx[i], T);
    fi.v[i] = idx[i];
  }
  const long long total4 = static_cast<long long>(BC) * n * (plane / 4);
  count_launch();
  return launch_pdl(select_frames_kernel, dim3(grid_for(total4, CONV_THREADS)), dim3(CONV_THREADS), 0, stream,
                    reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), T, n, plane / 4, fi, total4);
}

}  // namespace kvq
