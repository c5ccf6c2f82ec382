// HBM-bound row kernels around the GEMMs: LayerNorm prologues (with the cyclic-shift / window-partition /
// patch-merge gathers folded into the load), patch-embed im2col, head mean, fp32->fp16 weight packing and the
// Grid-Mini-patch fragment gather.  One warp per output row, float4 loads / 8-byte fp16 stores, fp32 statistics.
#include <algorithm>

#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int ROW_THREADS = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {   // sum over the LPR lanes that share a row
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm of a logical row made of `nseg` segments of C floats (1: plain / window gather, 4: PatchMerging concat);
// seg_ptr[s] == nullptr means zeros.  LPR lanes cooperate on one row (a warp covers 32/LPR rows, so narrow rows
// still keep >= 1.5 KB of loads in flight per warp); each lane owns MAXV float4 chunks: ch = sub + LPR*k.
// fp32 two-pass statistics (mean, then centred variance) like nn.LayerNorm.
template <int LPR, int MAXV>
__device__ __forceinline__ void ln_row(const float* const* seg_ptr, int nseg, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, __half* __restrict__ out,
                                       float* __restrict__ out_f32_cf, size_t cf_stride, int sub) {
  const int chunks_per_seg = C >> 2;
  const int chunks = chunks_per_seg * nseg;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int ch = sub + LPR * k;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch < chunks) {
      const int sg = ch / chunks_per_seg;
      const float* sp = seg_ptr[sg];
      if (sp != nullptr) v[k] = __ldg(reinterpret_cast<const float4*>(sp) + (ch - sg * chunks_per_seg));
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  const float inv_n = 1.0f / static_cast<float>(4 * chunks);
  const float mean = group_sum<LPR>(s) * inv_n;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int ch = sub + LPR * k;
    if (ch < chunks) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(group_sum<LPR>(q) * inv_n + eps);
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int ch = sub + LPR * k;
    if (ch < chunks) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + ch);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + ch);
      const float y0 = (v[k].x - mean) * rstd * g.x + b.x;
      const float y1 = (v[k].y - mean) * rstd * g.y + b.y;
      const float y2 = (v[k].z - mean) * rstd * g.z + b.z;
      const float y3 = (v[k].w - mean) * rstd * g.w + b.w;
      if (out != nullptr) {
        uint2 h;
        h.x = pack_half2(y0, y1);
        h.y = pack_half2(y2, y3);
        *reinterpret_cast<uint2*>(out + 4 * ch) = h;
      }
      if (out_f32_cf != nullptr) {
        out_f32_cf[static_cast<size_t>(4 * ch) * cf_stride] = y0;
        out_f32_cf[static_cast<size_t>(4 * ch + 1) * cf_stride] = y1;
        out_f32_cf[static_cast<size_t>(4 * ch + 2) * cf_stride] = y2;
        out_f32_cf[static_cast<size_t>(4 * ch + 3) * cf_stride] = y3;
      }
    }
  }
}

// All lanes of a warp must reach the shuffles: rows past the end are clamped to the last row and only skip stores.
template <int LPR, int MAXV>
__global__ void __launch_bounds__(ROW_THREADS)
ln_window_kernel(const float* __restrict__ x, __half* __restrict__ out, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, int rows, int C, WinGeom g) {
  constexpr int RPW = 32 / LPR;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  int row = (blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool live = row < rows;
  if (!live) row = rows - 1;
  const int rows_in = g.nW * g.N;
  const int b = fdiv_i(row, rows_in, g.r_rows);
  const int src = win_row_to_src(g, row - b * rows_in);
  __half* orow = live ? out + static_cast<size_t>(row) * C : nullptr;
  const float* sp = x + (static_cast<size_t>(b) * g.tokens + (src < 0 ? 0 : src)) * C;
  if (src < 0) {  // padded slot: zeros AFTER the norm (swin_backbone.py:416-424); still join the shuffles below
    if (live)
      for (int ch = sub; ch < (C >> 2); ch += LPR) *reinterpret_cast<uint2*>(orow + 4 * ch) = make_uint2(0u, 0u);
    orow = nullptr;
  }
  ln_row<LPR, MAXV>(&sp, 1, C, gamma, beta, eps, orow, nullptr, 0, sub);
}

template <int LPR, int MAXV>
__global__ void __launch_bounds__(ROW_THREADS)
ln_rows_kernel(const float* __restrict__ x, __half* __restrict__ out, float* __restrict__ feat_cf,
               const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int rows, int C,
               int tokens_per_clip) {
  constexpr int RPW = 32 / LPR;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  int row = (blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool live = row < rows;
  if (!live) row = rows - 1;
  const float* sp = x + static_cast<size_t>(row) * C;
  float* cf = nullptr;
  if (feat_cf != nullptr && live) {
    const int b = row / tokens_per_clip, t = row - b * tokens_per_clip;
    cf = feat_cf + static_cast<size_t>(b) * C * tokens_per_clip + t;  // [B, C, tokens]
  }
  ln_row<LPR, MAXV>(&sp, 1, C, gamma, beta, eps, (out != nullptr && live) ? out + static_cast<size_t>(row) * C : nullptr,
                    cf, static_cast<size_t>(tokens_per_clip), sub);
}

// LN1 + roll + window_partition for grids WITHOUT padding, walked in SOURCE order: the fp32 rows are read contiguously
// (the gather form above reads them in window order, i.e. 4C-byte pieces that are far apart: 73 vs 47 us for the same
// bytes at stage 0) and each fp16 output row -- whole 32 B sectors -- is scattered to its window-order slot.
template <int LPR, int MAXV>
__global__ void __launch_bounds__(ROW_THREADS)
ln_window_scatter_kernel(const float* __restrict__ x, __half* __restrict__ out, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, int rows, int C, WinGeom g) {
  constexpr int RPW = 32 / LPR;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  int row = (blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool live = row < rows;
  if (!live) row = rows - 1;
  const int b = fdiv_i(row, g.tokens, g.r_tokens);
  const int dst = b * (g.nW * g.N) + src_to_win_row(g, row - b * g.tokens);
  const float* sp = x + static_cast<size_t>(row) * C;
  ln_row<LPR, MAXV>(&sp, 1, C, gamma, beta, eps, live ? out + static_cast<size_t>(dst) * C : nullptr, nullptr, 0, sub);
}

// PatchMerging (:533-555): rows (b, d, h2, w2); channel order [x(2h,2w) | x(2h+1,2w) | x(2h,2w+1) | x(2h+1,2w+1)],
// odd H/W zero-padded BEFORE the norm.
template <int LPR, int MAXV>
__global__ void __launch_bounds__(ROW_THREADS)
ln_merge_kernel(const float* __restrict__ x, __half* __restrict__ out, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int rows, int D, int H, int W, int C) {
  constexpr int RPW = 32 / LPR;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  int row = (blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool live = row < rows;
  if (!live) row = rows - 1;
  const int H2 = (H + 1) >> 1, W2 = (W + 1) >> 1;
  const int t2 = fdiv_i(row, W2, 1.0f / W2), w2 = row - t2 * W2;      // (float-reciprocal divisions: kvq_kernels.cuh)
  const int bd = fdiv_i(t2, H2, 1.0f / H2), h2 = t2 - bd * H2;       // bd = b*D + d
  const float* seg[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int hh = 2 * h2 + (s & 1), ww = 2 * w2 + (s >> 1);
    seg[s] = (hh < H && ww < W) ? x + ((static_cast<size_t>(bd) * H + hh) * W + ww) * C : nullptr;
  }
  ln_row<LPR, MAXV>(seg, 4, C, gamma, beta, eps, live ? out + static_cast<size_t>(row) * 4 * C : nullptr, nullptr, 0,
                    sub);
}

// PatchEmbed3D im2col (:715-731).  One thread per (output token, c, kt, kh): copies the 4 kw taps (16 B in, 8 B out).
// Zero padding when T/H/W are not multiples of (2,4,4).
// TIn = float (the reference's input dtype) or __half (clips stored / shipped as fp16: the operand is rounded to fp16
// here anyway, so a host-side round-to-nearest gives bit-identical operands at half the H2D and read traffic).
template <typename TIn>
__global__ void __launch_bounds__(256)
patch_im2col_kernel(const TIn* __restrict__ x, __half* __restrict__ out, int B, int T, int H, int W, int D, int Hs,
                    int Ws, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // ws fastest so a warp reads 512 contiguous bytes of one input row; the 8 B stores merge in L2
  const int ws = static_cast<int>(idx % Ws);
  const int kq = static_cast<int>((idx / Ws) % 24);  // c*8 + kt*4 + kh
  const long long bdh = idx / (static_cast<long long>(Ws) * 24);
  const int hs = static_cast<int>(bdh % Hs);
  const int d = static_cast<int>((bdh / Hs) % D);
  const int b = static_cast<int>(bdh / (static_cast<long long>(Hs) * D));
  const int c = kq >> 3, kt = (kq >> 2) & 1, kh = kq & 3;
  const long long tok = ((static_cast<long long>(b) * D + d) * Hs + hs) * Ws + ws;
  const int t = 2 * d + kt, h = 4 * hs + kh, w0 = 4 * ws;
  uint2 h2 = make_uint2(0u, 0u);
  if (t < T && h < H) {
    const TIn* src = x + (((static_cast<size_t>(b) * 3 + c) * T + t) * H + h) * W + w0;
    if constexpr (sizeof(TIn) == 4) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (w0 + 3 < W && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(src));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (w0 + k < W) v[k] = __ldg(src + k);
      }
      h2.x = pack_half2(v[0], v[1]);
      h2.y = pack_half2(v[2], v[3]);
    } else {
      if (w0 + 3 < W && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
        h2 = __ldg(reinterpret_cast<const uint2*>(src));
      } else {
        unsigned short u[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (w0 + k < W) u[k] = __ldg(reinterpret_cast<const unsigned short*>(src) + k);
        h2.x = u[0] | (static_cast<uint32_t>(u[1]) << 16);
        h2.y = u[2] | (static_cast<uint32_t>(u[3]) << 16);
      }
    }
  }
  *reinterpret_cast<uint2*>(out + tok * 96 + c * 32 + kt * 16 + kh * 4) = h2;
}

__global__ void __launch_bounds__(256)
row_mean_kernel(const float* __restrict__ rowscore, float* __restrict__ score, int tokens) {
  __shared__ float part[8];
  pdl_launch_dependents();
  pdl_wait();
  const float* src = rowscore + static_cast<size_t>(blockIdx.x) * tokens;
  float s = 0.f;
  for (int i = threadIdx.x; i < tokens; i += blockDim.x) s += src[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i];
    score[blockIdx.x] = t / static_cast<float>(tokens);
  }
}

__global__ void __launch_bounds__(256)
cast_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}

// Grid Mini-patch Sampling (get_spatial_fragments, fusion_datasets.py:22-121) + (v - mean) / std (:1017-1020).
// One thread per 4 consecutive output pixels of one fragment row: 4 B in, 16 B out.
__global__ void __launch_bounds__(256)
fragment_gather_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ offsets, float* __restrict__ out,
                       int T, int Hs, int Ws, int fh, int fw, int fs, int aligned, float m0, float m1, float m2,
                       float is0, float is1, float is2, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int OW = fw * fs, OH = fh * fs, OW4 = OW >> 2;
  const int x4 = static_cast<int>(idx % OW4);
  const int y = static_cast<int>((idx / OW4) % OH);
  const int t = static_cast<int>((idx / (static_cast<long long>(OW4) * OH)) % T);
  const int c = static_cast<int>((idx / (static_cast<long long>(OW4) * OH * T)) % 3);
  const int b = static_cast<int>(idx / (static_cast<long long>(OW4) * OH * T * 3));
  const int x = x4 * 4;
  const int i = y / fs, dy = y - i * fs, j = x / fs, dx = x - j * fs;
  const int nt = T / aligned, tc = t / aligned;
  // hgrids / wgrids (:64-69): cell origin clamped so the patch stays inside the frame
  int hg = (Hs / fh) * i; if (hg > Hs - fs) hg = Hs - fs;
  int wg = (Ws / fw) * j; if (wg > Ws - fs) wg = Ws - fs;
  const int32_t* off = offsets + static_cast<size_t>(b) * 2 * fh * fw * nt;
  // offsets live on the device and cannot be validated by the launcher: clamp them so a bad offset tensor (negative,
  // beyond the cell, drawn for another source size) reads inside the frame instead of out of bounds
  int oh = off[(i * fw + j) * nt + tc];
  int ow = off[fh * fw * nt + (i * fw + j) * nt + tc];
  oh = oh < 0 ? 0 : (oh > Hs - fs - hg ? Hs - fs - hg : oh);
  ow = ow < 0 ? 0 : (ow > Ws - fs - wg ? Ws - fs - wg : ow);
  const uint8_t* src = frames + (((static_cast<size_t>(b) * T + t) * 3 + c) * Hs + (hg + oh + dy)) * Ws + (wg + ow + dx);
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
  const float istd = c == 0 ? is0 : (c == 1 ? is1 : is2);
  float4 v;
  v.x = (static_cast<float>(src[0]) - mean) * istd;
  v.y = (static_cast<float>(src[1]) - mean) * istd;
  v.z = (static_cast<float>(src[2]) - mean) * istd;
  v.w = (static_cast<float>(src[3]) - mean) * istd;
  *reinterpret_cast<float4*>(out + (((static_cast<size_t>(b) * 3 + c) * T + t) * OH + y) * OW + x) = v;
}

// Same sampling when the source is smaller than the fragment canvas (fallback_type == "upsample", fusion_datasets.py:
// 43-50): the frame is first enlarged with F.interpolate(video / 255, scale_factor = 1 / ratio, mode = "bilinear")
// (align_corners = False: src = (dst + 0.5) * rscale - 0.5 clamped at 0, rscale = float(1 / scale_factor)) and scaled
// back by 255; the cell grid and the offsets keep using the ORIGINAL resolution (:64-71).  The enlarged frame is never
// materialised: each output pixel interpolates its four source pixels.  One thread per output pixel (rare path).
__global__ void __launch_bounds__(256)
fragment_gather_upsample_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ offsets,
                                float* __restrict__ out, int T, int Hs, int Ws, int fh, int fw, int fs, int aligned,
                                float rscale, float m0, float m1, float m2, float is0, float is1, float is2,
                                long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int OW = fw * fs, OH = fh * fs;
  const int x = static_cast<int>(idx % OW);
  const int y = static_cast<int>((idx / OW) % OH);
  const int t = static_cast<int>((idx / (static_cast<long long>(OW) * OH)) % T);
  const int c = static_cast<int>((idx / (static_cast<long long>(OW) * OH * T)) % 3);
  const int b = static_cast<int>(idx / (static_cast<long long>(OW) * OH * T * 3));
  const int i = y / fs, dy = y - i * fs, j = x / fs, dx = x - j * fs;
  const int nt = T / aligned, tc = t / aligned;
  int hg = (Hs / fh) * i; if (hg > Hs - fs) hg = Hs - fs;
  int wg = (Ws / fw) * j; if (wg > Ws - fs) wg = Ws - fs;
  const int32_t* off = offsets + static_cast<size_t>(b) * 2 * fh * fw * nt;
  int oh = off[(i * fw + j) * nt + tc], ow = off[fh * fw * nt + (i * fw + j) * nt + tc];
  oh = oh < 0 ? 0 : oh;                                                       // (the source reads below are clamped)
  ow = ow < 0 ? 0 : ow;
  const int uy = hg + oh + dy;                                                // pixel of the enlarged frame
  const int ux = wg + ow + dx;
  float sy = rscale * (static_cast<float>(uy) + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
  float sx = rscale * (static_cast<float>(ux) + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
  int y0 = static_cast<int>(sy); if (y0 > Hs - 1) y0 = Hs - 1;
  int x0 = static_cast<int>(sx); if (x0 > Ws - 1) x0 = Ws - 1;
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
  const float ly = sy - static_cast<float>(y0), lx = sx - static_cast<float>(x0);
  const uint8_t* src = frames + ((static_cast<size_t>(b) * T + t) * 3 + c) * Hs * Ws;
  const float inv255 = 1.0f / 255.0f;   // (the reference divides; the difference is below 1 ulp of the product)
  const float a = __fdiv_rn(static_cast<float>(src[y0 * Ws + x0]), 255.0f), bq = __fdiv_rn(static_cast<float>(src[y0 * Ws + x1]), 255.0f);
  const float cq = __fdiv_rn(static_cast<float>(src[y1 * Ws + x0]), 255.0f), d = __fdiv_rn(static_cast<float>(src[y1 * Ws + x1]), 255.0f);
  (void)inv255;
  const float top = __fadd_rn(__fmul_rn(1.0f - lx, a), __fmul_rn(lx, bq));
  const float bot = __fadd_rn(__fmul_rn(1.0f - lx, cq), __fmul_rn(lx, d));
  const float v = __fmul_rn(__fadd_rn(__fmul_rn(1.0f - ly, top), __fmul_rn(ly, bot)), 255.0f);
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
  const float istd = c == 0 ? is0 : (c == 1 ? is1 : is2);
  out[(((static_cast<size_t>(b) * 3 + c) * T + t) * OH + y) * OW + x] = (v - mean) * istd;
}

// ---- QRS: quality-aware region selection of KSVQE (RegionNet_CLIP.forward eval branch, patchnet.py:461-550) ----
// score [B*n_key, L] = cos(CLS, patch) of each key frame on an l x l grid (l*l = L).  Per key frame: nearest-neighbour
// resize to the g x g fragment grid (F.interpolate(scale_factor = g/l, mode="nearest"): src = floor(dst * float(l/g))),
// mean over every ks x ks window (F.unfold, stride 1), arg-max (min_max_norm is monotone; HardTopK(1) keeps the
// first maximum) -> region[b, key] = ry * (g - ks + 1) + rx.  One thread per key frame; the map is at most 16 x 16.
__global__ void __launch_bounds__(128)
qrs_select_kernel(const float* __restrict__ score, int32_t* __restrict__ region, int n, int l, int g, int ks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* s = score + static_cast<size_t>(i) * l * l;
  const float rs = 1.0f / (static_cast<float>(g) / static_cast<float>(l));   // ATen: scale = 1 / scale_factor (fp32)
  const int nr = g - ks + 1;
  float best = -INFINITY;
  int arg = 0;
  for (int ry = 0; ry < nr; ++ry)
    for (int rx = 0; rx < nr; ++rx) {
      float acc = 0.f;
      for (int dy = 0; dy < ks; ++dy) {
        int sy = static_cast<int>(floorf(static_cast<float>(ry + dy) * rs)); if (sy > l - 1) sy = l - 1;
        for (int dx = 0; dx < ks; ++dx) {
          int sx = static_cast<int>(floorf(static_cast<float>(rx + dx) * rs)); if (sx > l - 1) sx = l - 1;
          acc += s[sy * l + sx];
        }
      }
      acc /= static_cast<float>(ks * ks);
      if (acc > best) { best = acc; arg = ry * nr + rx; }
    }
  region[i] = arg;
}

// x_sel[b, c, t] = fragment[b, c, t, a*ry : a*ry + a*ks, a*rx : a*rx + a*ks] with (ry, rx) = region of frame t's key-frame
// group (obtain_keyframes, KSVQE_model.py:1352-1376: groups change at t = T/4 - 1, T/2 - 1, 3T/4 - 1).  float4 per thread.
__global__ void __launch_bounds__(256)
qrs_gather_kernel(const float* __restrict__ frag, const int32_t* __restrict__ region, float* __restrict__ out, int T,
                  int H, int W, int n_key, int anchor, int ks, int nr, long long total4) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int S = anchor * ks, S4 = S >> 2;
  const int x4 = static_cast<int>(idx % S4);
  const int y = static_cast<int>((idx / S4) % S);
  const int t = static_cast<int>((idx / (static_cast<long long>(S4) * S)) % T);
  const long long bc = idx / (static_cast<long long>(S4) * S * T);
  const int b = static_cast<int>(bc / 3);
  int grp = (t >= T / 4 - 1) + (t >= T / 2 - 1) + (t >= T * 3 / 4 - 1);
  if (grp > n_key - 1) grp = n_key - 1;
  const int r = region[b * n_key + grp];
  const int ry = r / nr, rx = r - ry * nr;
  const float4 v = *reinterpret_cast<const float4*>(frag + ((bc * T + t) * H + (anchor * ry + y)) * W + anchor * rx + 4 * x4);
  *reinterpret_cast<float4*>(out + ((bc * T + t) * S + y) * S + 4 * x4) = v;
}

// CONTRIQUE_model.forward patch split (KSVQE_model.py:1643-1651) of every `step`-th frame:
// x [B,3,T,H,W] -> patches [N,3,1,a,a], N = B * T/step * (H/a) * (W/a) ordered (b, t, gy, gx); float4 per thread
__global__ void __launch_bounds__(256)
patch_split_kernel(const float* __restrict__ x, float* __restrict__ out, int T, int H, int W, int a, int step,
                   long long total4) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int a4 = a >> 2, gw = W / a, gh = H / a, Ts = T / step;
  long long r = idx;
  const int x4 = static_cast<int>(r % a4); r /= a4;
  const int py = static_cast<int>(r % a); r /= a;
  const int c = static_cast<int>(r % 3); r /= 3;
  const int gx = static_cast<int>(r % gw); r /= gw;
  const int gy = static_cast<int>(r % gh); r /= gh;
  const int t = static_cast<int>(r % Ts);
  const long long b = r / Ts;
  const float4 v = *reinterpret_cast<const float4*>(
      x + (((b * 3 + c) * T + static_cast<long long>(t) * step) * H + (gy * a + py)) * W + gx * a + 4 * x4);
  reinterpret_cast<float4*>(out)[idx] = v;
}

// F.normalize(h, dim=1) (:1657): rows of C halfs, one warp per row, fp32 norm, eps 1e-12
__global__ void __launch_bounds__(256)
row_l2norm_kernel(const __half* __restrict__ in, __half* __restrict__ out, int rows, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const __half2* src = reinterpret_cast<const __half2*>(in + static_cast<size_t>(row) * C);
  float ss = 0.f;
  for (int i = lane; i < (C >> 1); i += 32) {
    const float2 f = __half22float2(src[i]);
    ss += f.x * f.x + f.y * f.y;
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  __half2* dst = reinterpret_cast<__half2*>(out + static_cast<size_t>(row) * C);
  for (int i = lane; i < (C >> 1); i += 32) {
    const float2 f = __half22float2(src[i]);
    dst[i] = __floats2half2_rn(f.x * inv, f.y * inv);
  }
}

__global__ void __launch_bounds__(256)
f16_to_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __half2float(in[i]);
}

// channels-first fp32 [B, C, tokens] -> token rows fp16 [B*tokens, C] (input side of a stand-alone VQAHead)
__global__ void __launch_bounds__(256)
cf_to_rows_kernel(const float* __restrict__ in, __half* __restrict__ out, int C, int tokens) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, t = t0 + tx;
    tile[k][tx] = (c < C && t < tokens) ? in[(static_cast<size_t>(b) * C + c) * tokens + t] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int t = t0 + k, c = c0 + tx;
    if (t < tokens && c < C) out[(static_cast<size_t>(b) * tokens + t) * C + c] = __float2half_rn(tile[tx][k]);
  }
}

// (lanes per row, chunks per lane) for a row of `chunks` float4s
// fp32 [rows, K] -> fp16 [rows, 2*Kp]: hi at [0,K), lo = fp16(w - hi) at [Kp, Kp+K), zeros elsewhere
__global__ void __launch_bounds__(256)
pack_split_kernel(const float* __restrict__ in, __half* __restrict__ out, int rows, int K, int Kp) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * 2 * Kp) return;
  const int c = static_cast<int>(i % (2 * Kp));
  const long long r = i / (2 * Kp);
  const int k = c < Kp ? c : c - Kp;
  __half v = __float2half_rn(0.f);
  if (k < K) {
    const float w = in[r * K + k];
    const __half hi = __float2half_rn(w);
    v = c < Kp ? hi : __float2half_rn(w - __half2float(hi));
  }
  out[i] = v;
}

template <typename F>
int dispatch_ln(int chunks, F&& f) {
  if (chunks <= 24) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 3>{});
  if (chunks <= 48) return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 3>{});
  if (chunks <= 96) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 3>{});
  if (chunks <= 192) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 6>{});
  if (chunks <= 384) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 12>{});
  if (chunks <= 768) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 24>{});
  set_error("LayerNorm row of %d floats is wider than this build supports (3072)", chunks * 4);
  return KVQ_ERR_BAD_SHAPE;
}

}  // namespace

int launch_ln_window(const float* x, __half* out, const float* gamma, const float* beta, float eps, int B, int C,
                     const WinGeom& g, cudaStream_t stream) {
  KVQ_REQUIRE(C % 4 == 0, KVQ_ERR_BAD_SHAPE, "ln_window: C=%d not a multiple of 4", C);
  const long long rows = static_cast<long long>(B) * g.nW * g.N;
  KVQ_REQUIRE(rows < (1ll << 31), KVQ_ERR_BAD_SHAPE, "ln_window: %lld rows overflow int32", rows);
  if (g.Dp == g.D && g.Hp == g.H && g.Wp == g.W) {   // no padded slots: source-order walk, scattered output rows
    return dispatch_ln(C / 4, [&](auto lpr, auto mv) {
      constexpr int RPC = (ROW_THREADS / 32) * (32 / decltype(lpr)::value);
      const int grid = static_cast<int>((rows + RPC - 1) / RPC);
      count_launch();
      return launch_pdl(ln_window_scatter_kernel<decltype(lpr)::value, decltype(mv)::value>, dim3(grid), dim3(ROW_THREADS),
                        0, stream, x, out, gamma, beta, eps, static_cast<int>(rows), C, g);
    });
  }
  return dispatch_ln(C / 4, [&](auto lpr, auto mv) {
    constexpr int RPC = (ROW_THREADS / 32) * (32 / decltype(lpr)::value);
    const int grid = static_cast<int>((rows + RPC - 1) / RPC);
    count_launch();
    return launch_pdl(ln_window_kernel<decltype(lpr)::value, decltype(mv)::value>, dim3(grid), dim3(ROW_THREADS), 0,
                      stream, x, out, gamma, beta, eps, static_cast<int>(rows), C, g);
  });
}

int launch_ln_rows(const float* x, __half* out, float* feat_cf, const float* gamma, const float* beta, float eps,
                   int M, int C, int tokens_per_clip, cudaStream_t stream) {
  KVQ_REQUIRE(C % 4 == 0 && M > 0, KVQ_ERR_BAD_SHAPE, "ln_rows: M=%d C=%d", M, C);
  return dispatch_ln(C / 4, [&](auto lpr, auto mv) {
    constexpr int RPC = (ROW_THREADS / 32) * (32 / decltype(lpr)::value);
    const int grid = (M + RPC - 1) / RPC;
    count_launch();
    return launch_pdl(ln_rows_kernel<decltype(lpr)::value, decltype(mv)::value>, dim3(grid), dim3(ROW_THREADS), 0,
                      stream, x, out, feat_cf, gamma, beta, eps, M, C, tokens_per_clip);
  });
}

int launch_ln_merge(const float* x, __half* out, const float* gamma, const float* beta, float eps, int B, int D,
                    int H, int W, int C, cudaStream_t stream) {
  KVQ_REQUIRE(C % 4 == 0, KVQ_ERR_BAD_SHAPE, "ln_merge: C=%d not a multiple of 4", C);
  const long long rows = static_cast<long long>(B) * D * ((H + 1) / 2) * ((W + 1) / 2);
  KVQ_REQUIRE(rows < (1ll << 31), KVQ_ERR_BAD_SHAPE, "ln_merge: %lld rows overflow int32", rows);
  return dispatch_ln(C, [&](auto lpr, auto mv) {
    constexpr int RPC = (ROW_THREADS / 32) * (32 / decltype(lpr)::value);
    const int grid = static_cast<int>((rows + RPC - 1) / RPC);
    count_launch();
    return launch_pdl(ln_merge_kernel<decltype(lpr)::value, decltype(mv)::value>, dim3(grid), dim3(ROW_THREADS), 0,
                      stream, x, out, gamma, beta, eps, static_cast<int>(rows), D, H, W, C);
  });
}

int launch_patch_im2col(const void* x, int x_is_f16, __half* out, int B, int T, int H, int W, cudaStream_t stream) {
  const int D = (T + 1) / 2, Hs = (H + 3) / 4, Ws = (W + 3) / 4;
  const long long total = static_cast<long long>(B) * D * Hs * Ws * 24;
  const long long grid = (total + 255) / 256;
  KVQ_REQUIRE(grid < (1ll << 31), KVQ_ERR_BAD_SHAPE, "im2col: grid too large");
  count_launch();
  if (x_is_f16)
    return launch_pdl(patch_im2col_kernel<__half>, dim3(static_cast<unsigned>(grid)), dim3(256), 0, stream,
                      static_cast<const __half*>(x), out, B, T, H, W, D, Hs, Ws, total);
  return launch_pdl(patch_im2col_kernel<float>, dim3(static_cast<unsigned>(grid)), dim3(256), 0, stream,
                    static_cast<const float*>(x), out, B, T, H, W, D, Hs, Ws, total);
}

int launch_row_mean(const float* rowscore, float* score, int B, int tokens, cudaStream_t stream) {
  count_launch();
  return launch_pdl(row_mean_kernel, dim3(B), dim3(256), 0, stream, rowscore, score, tokens);
}

int launch_fragment_gather_u8(const uint8_t* frames, const int32_t* offsets, float* out, int B, int T, int Hs, int Ws,
                              int fh, int fw, int fs, int aligned, const float mean[3], const float stdv[3],
                              cudaStream_t stream) {
  KVQ_REQUIRE(B > 0 && T > 0 && fh > 0 && fw > 0 && fs > 0 && fs % 4 == 0, KVQ_ERR_BAD_SHAPE,
              "fragment_gather: bad geometry (fsize must be a multiple of 4)");
  KVQ_REQUIRE(aligned > 0 && T % aligned == 0, KVQ_ERR_BAD_SHAPE,
              "fragment_gather: T=%d is not a multiple of aligned=%d (fusion_datasets.py:59)", T, aligned);
  KVQ_REQUIRE(Hs >= fs && Ws >= fs, KVQ_ERR_BAD_SHAPE, "fragment_gather: %dx%d source is smaller than one %dx%d patch",
              Hs, Ws, fs, fs);
  if (Hs < fh * fs || Ws < fw * fs) {
    // fallback_type == "upsample" (:43-50): ratio and scale factor are Python floats (doubles) in the reference
    const double ratio = std::min(static_cast<double>(Hs) / (fh * fs), static_cast<double>(Ws) / (fw * fs));
    const double scale = 1.0 / ratio;
    const float rscale = static_cast<float>(1.0 / scale);
    const long long tot = static_cast<long long>(B) * 3 * T * (fh * fs) * (fw * fs);
    const long long g = (tot + 255) / 256;
    KVQ_REQUIRE(g < (1ll << 31), KVQ_ERR_BAD_SHAPE, "fragment_gather: grid too large");
    fragment_gather_upsample_kernel<<<static_cast<unsigned>(g), 256, 0, stream>>>(
        frames, offsets, out, T, Hs, Ws, fh, fw, fs, aligned, rscale, mean[0], mean[1], mean[2], 1.0f / stdv[0],
        1.0f / stdv[1], 1.0f / stdv[2], tot);
    count_launch();
    return check_cuda(cudaGetLastError(), "fragment_gather_upsample_kernel launch");
  }
  const long long total = static_cast<long long>(B) * 3 * T * (fh * fs) * (fw * fs / 4);
  const long long grid = (total + 255) / 256;
  KVQ_REQUIRE(grid < (1ll << 31), KVQ_ERR_BAD_SHAPE, "fragment_gather: grid too large");
  fragment_gather_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(
      frames, offsets, out, T, Hs, Ws, fh, fw, fs, aligned, mean[0], mean[1], mean[2], 1.0f / stdv[0], 1.0f / stdv[1],
      1.0f / stdv[2], total);
  count_launch();
  return check_cuda(cudaGetLastError(), "fragment_gather_kernel launch");
}

int launch_patch_split(const float* x, float* out, int B, int T, int H, int W, int a, int step, cudaStream_t stream) {
  KVQ_REQUIRE(B > 0 && T > 0 && step > 0 && T % step == 0 && a > 0 && a % 4 == 0 && H % a == 0 && W % a == 0,
              KVQ_ERR_BAD_SHAPE, "patch_split: x %dx3x%dx%dx%d, patch %d, frame step %d", B, T, H, W, a, step);
  const long long total4 = static_cast<long long>(B) * (T / step) * (H / a) * (W / a) * 3 * a * (a / 4);
  const long long grid = (total4 + 255) / 256;
  KVQ_REQUIRE(grid < (1ll << 31), KVQ_ERR_BAD_SHAPE, "patch_split: grid too large");
  patch_split_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(x, out, T, H, W, a, step, total4);
  count_launch();
  return check_cuda(cudaGetLastError(), "patch_split_kernel launch");
}

int launch_row_l2norm(const __half* in, __half* out, int rows, int C, cudaStream_t stream) {
  KVQ_REQUIRE(rows > 0 && C > 0 && C % 2 == 0, KVQ_ERR_BAD_SHAPE, "row_l2norm: rows=%d C=%d", rows, C);
  row_l2norm_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(in, out, rows, C);
  count_launch();
  return check_cuda(cudaGetLastError(), "row_l2norm_kernel launch");
}

int launch_f16_to_f32(const __half* in, float* out, size_t n, cudaStream_t stream) {
  f16_to_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(in, out, n);
  count_launch();
  return check_cuda(cudaGetLastError(), "f16_to_f32_kernel launch");
}

int launch_qrs_select_gather(const float* fragment, const float* score, float* out, int32_t* region, int B, int T, int H,
                             int W, int n_key, int L, int anchor, int ks, cudaStream_t stream) {
  int l = 1;
  while (l * l < L) ++l;
  KVQ_REQUIRE(fragment && score && out && region, KVQ_ERR_BAD_SHAPE, "qrs: NULL argument");
  KVQ_REQUIRE(B > 0 && T >= 4 && n_key == 4 && l * l == L && anchor > 0 && anchor % 4 == 0 && ks > 0, KVQ_ERR_BAD_SHAPE,
              "qrs: B=%d T=%d n_key=%d (4 key frames) L=%d (square) anchor=%d ks=%d", B, T, n_key, L, anchor, ks);
  KVQ_REQUIRE(H == W && H % anchor == 0 && H / anchor >= ks && W % 4 == 0, KVQ_ERR_BAD_SHAPE,
              "qrs: fragment %dx%d must be a square grid of %d-pixel patches with at least %d per side", H, W, anchor, ks);
  const int g = H / anchor, nr = g - ks + 1;
  qrs_select_kernel<<<(B * n_key + 127) / 128, 128, 0, stream>>>(score, region, B * n_key, l, g, ks);
  count_launch();
  KVQ_CUDA(cudaGetLastError());
  const long long total4 = static_cast<long long>(B) * 3 * T * (anchor * ks) * (anchor * ks / 4);
  const long long grid = (total4 + 255) / 256;
  KVQ_REQUIRE(grid < (1ll << 31), KVQ_ERR_BAD_SHAPE, "qrs: grid too large");
  qrs_gather_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(fragment, region, out, T, H, W, n_key, anchor, ks, nr,
                                                                   total4);
  count_launch();
  return check_cuda(cudaGetLastError(), "qrs_gather_kernel launch");
}

int launch_cf_to_rows(const float* in, __half* out, int B, int C, int tokens, cudaStream_t stream) {
  dim3 grid((tokens + 31) / 32, (C + 31) / 32, B);
  cf_to_rows_kernel<<<grid, 256, 0, stream>>>(in, out, C, tokens);
  count_launch();
  return check_cuda(cudaGetLastError(), "cf_to_rows_kernel launch");
}

int launch_pack_split(const float* in, __half* out, int rows, int K, cudaStream_t stream) {
  const int Kp = (K + 63) / 64 * 64;
  const long long n = static_cast<long long>(rows) * 2 * Kp;
  if (n == 0) return KVQ_OK;
  pack_split_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(in, out, rows, K, Kp);
  count_launch();
  return check_cuda(cudaGetLastError(), "pack_split_kernel launch");
}

int launch_cast_f16(const float* in, __half* out, size_t n, cudaStream_t stream) {
  if (n == 0) return KVQ_OK;
  cast_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(in, out, n);
  count_launch();
  return check_cuda(cudaGetLastError(), "cast_f16_kernel launch");
}

}  // namespace kvq
