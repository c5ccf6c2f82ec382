// Small kernels of KSVQE's cross-gating modulation (CDM, models/backbones/KSVQE_model.py:1436-1482): multi-head
// attention over short token sequences (crossattention1 :1553-1587, Attention :1508-1551), the SFT-style gating
// (Semantic_Transformation2 :817-835, Dist_Transformation3 :934-960, mix :1482) and tiny fp32 Linears on per-clip
// statistics.  The Linears on token rows are the tcgen05 GEMMs of kvq_gemm.cu; everything here is CUDA-core work on
// 784 tokens per clip.
#include "../../include/kvq_b200.h"
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

using namespace kvq;

namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// softmax(q k^T * scale) v for one (sequence, head); head_dim 64; Lq, Lkv <= 64.  Token t of sequence (o, i) is row
// o*outer + i*inner + t*tstride of q / k / v / out (row strides ldq / ldk / ldv / ldo halfs).
__global__ void __launch_bounds__(64)
mha_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, int ldk, const __half* __restrict__ v,
           int ldv, __half* __restrict__ out, int ldo, int n_inner, long long outer, long long inner, long long tstride,
           int Lq, int Lkv, float scale) {
  __shared__ float sk[64][65], sv[64][65];
  const int seq = blockIdx.x, h = blockIdx.y, t = threadIdx.x;
  const long long base = (seq / n_inner) * outer + (seq % n_inner) * inner;
  for (int i = threadIdx.x; i < Lkv * 64; i += blockDim.x) {
    const int r = i >> 6, d = i & 63;
    const long long row = base + r * tstride;
    sk[r][d] = __half2float(k[row * ldk + h * 64 + d]);
    sv[r][d] = __half2float(v[row * ldv + h * 64 + d]);
  }
  __syncthreads();
  if (t >= Lq) return;
  const long long row = base + t * tstride;
  float qr[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) qr[d] = __half2float(q[row * ldq + h * 64 + d]) * scale;
  float s[64];
  float m = -INFINITY;
  for (int j = 0; j < Lkv; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) a = fmaf(qr[d], sk[j][d], a);
    s[j] = a;
    m = fmaxf(m, a);
  }
  float l = 0.f;
  for (int j = 0; j < Lkv; ++j) { s[j] = __expf(s[j] - m); l += s[j]; }
  const float inv = 1.0f / l;
  __half* op = out + row * ldo + h * 64;
  for (int d0 = 0; d0 < 64; d0 += 8) {
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < Lkv; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(s[j], sv[j][d0 + e], o[e]);
    }
    uint4 w;
    w.x = pack_half2(o[0] * inv, o[1] * inv); w.y = pack_half2(o[2] * inv, o[3] * inv);
    w.z = pack_half2(o[4] * inv, o[5] * inv); w.w = pack_half2(o[6] * inv, o[7] * inv);
    *reinterpret_cast<uint4*>(op + d0) = w;
  }
}

// x[r, c] <- (a1 * (sigmoid(gd[b, c]) * x + bd[b, c]) + a2 * (gs_r * x + bs_r)) / 2   with, per token row r,
// gs_r = sigmoid(<es[r], wg> + bg), bs_r = <es[r], wb> + bb   (b = r / rows_per_clip).  One warp per row.
__global__ void __launch_bounds__(256)
cdm_mix_kernel(float* __restrict__ x, const __half* __restrict__ es, const float* __restrict__ wg,
               const float* __restrict__ bg, const float* __restrict__ wb, const float* __restrict__ bb,
               const float* __restrict__ gd, const float* __restrict__ bd, const float* __restrict__ a1,
               const float* __restrict__ a2, int rows, int C, int rows_per_clip) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const __half* e = es + static_cast<size_t>(r) * C;
  float dg = 0.f, db = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float ev = __half2float(e[c]);
    dg = fmaf(ev, wg[c], dg);
    db = fmaf(ev, wb[c], db);
  }
  dg = wsum(dg) + bg[0];
  db = wsum(db) + bb[0];
  const float gs = 1.0f / (1.0f + __expf(-dg));
  const float A1 = a1[0], A2 = a2[0];
  const int b = r / rows_per_clip;
  float* xr = x + static_cast<size_t>(r) * C;
  for (int c = lane; c < C; c += 32) {
    const float xv = xr[c];
    const float g = 1.0f / (1.0f + __expf(-gd[static_cast<size_t>(b) * C + c]));
    const float xd = g * xv + bd[static_cast<size_t>(b) * C + c];
    const float xs = gs * xv + db;
    xr[c] = 0.5f * (A1 * xd + A2 * xs);
  }
}

// out[m, n] = <x[m, :K], w[n, :K]> + b[n]   (fp32, tiny M): one warp per output element
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                    float* __restrict__ out, int M, int N, int K) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= M * N) return;
  const int m = o / N, n = o - m * N, lane = threadIdx.x & 31;
  float a = 0.f;
  for (int k = lane; k < K; k += 32) a = fmaf(x[static_cast<size_t>(m) * K + k], w[static_cast<size_t>(n) * K + k], a);
  a = wsum(a);
  if (lane == 0) out[o] = a + (b != nullptr ? b[n] : 0.f);
}

// out = wa * a (fp16) + wb * z (fp32), written as fp16 and / or fp32   (dist_token blend, :1426)
__global__ void __launch_bounds__(256)
blend_kernel(const __half* __restrict__ a, const float* __restrict__ z, float wa, float wb, __half* __restrict__ o16,
             float* __restrict__ o32, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = wa * __half2float(a[i]) + wb * z[i];
  if (o16 != nullptr) o16[i] = __float2half_rn(v);
  if (o32 != nullptr) o32[i] = v;
}

}  // namespace

extern "C" {

int kvq_mha_f16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                int n_outer, int n_inner, long long outer_stride, long long inner_stride, long long t_stride, int Lq,
                int Lkv, int heads, float scale, void* stream) {
  KVQ_REQUIRE(q && k && v && out, KVQ_ERR_BAD_SHAPE, "mha: NULL argument");
  KVQ_REQUIRE(n_outer > 0 && n_inner > 0 && Lq >= 1 && Lq <= 64 && Lkv >= 1 && Lkv <= 64 && heads >= 1, KVQ_ERR_BAD_SHAPE,
              "mha: n=%dx%d Lq=%d Lkv=%d heads=%d (sequences of at most 64 tokens, head_dim 64)", n_outer, n_inner, Lq,
              Lkv, heads);
  KVQ_REQUIRE(ldq >= heads * 64 && ldk >= heads * 64 && ldv >= heads * 64 && ldo >= heads * 64 && ldo % 8 == 0,
              KVQ_ERR_BAD_SHAPE, "mha: row strides %d/%d/%d/%d too small for %d heads of 64", ldq, ldk, ldv, ldo, heads);
  mha_kernel<<<dim3(n_outer * n_inner, heads), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(q), ldq, static_cast<const __half*>(k), ldk, static_cast<const __half*>(v), ldv,
      static_cast<__half*>(out), ldo, n_inner, outer_stride, inner_stride, t_stride, Lq, Lkv, scale);
  count_launch();
  return check_cuda(cudaGetLastError(), "mha_kernel launch");
}

int kvq_cdm_mix(float* x, const void* es_f16, const float* wg, const float* bg, const float* wb, const float* bb,
                const float* gd_pre, const float* bd, const float* a1, const float* a2, int rows, int C,
                int rows_per_clip, void* stream) {
  KVQ_REQUIRE(x && es_f16 && wg && bg && wb && bb && gd_pre && bd && a1 && a2, KVQ_ERR_BAD_SHAPE, "cdm_mix: NULL argument");
  KVQ_REQUIRE(rows > 0 && C > 0 && rows_per_clip > 0 && rows % rows_per_clip == 0, KVQ_ERR_BAD_SHAPE,
              "cdm_mix: rows=%d C=%d rows_per_clip=%d", rows, C, rows_per_clip);
  cdm_mix_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<const __half*>(es_f16), wg, bg, wb, bb, gd_pre, bd, a1, a2, rows, C, rows_per_clip);
  count_launch();
  return check_cuda(cudaGetLastError(), "cdm_mix_kernel launch");
}

int kvq_small_linear_f32(const float* x, const float* w, const float* b, float* out, int M, int N, int K, void* stream) {
  KVQ_REQUIRE(x && w && out && M > 0 && N > 0 && K > 0, KVQ_ERR_BAD_SHAPE, "small_linear: bad arguments");
  small_linear_kernel<<<(M * N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, b, out, M, N, K);
  count_launch();
  return check_cuda(cudaGetLastError(), "small_linear_kernel launch");
}

int kvq_blend_f16_f32(const void* a_f16, const float* z, float wa, float wb, void* out_f16, float* out_f32, size_t n,
                      void* stream) {
  KVQ_REQUIRE(a_f16 && z && (out_f16 || out_f32) && n > 0, KVQ_ERR_BAD_SHAPE, "blend: bad arguments");
  blend_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(a_f16), z, wa, wb, static_cast<__half*>(out_f16), out_f32, n);
  count_launch();
  return check_cuda(cudaGetLastError(), "blend_kernel launch");
}

}  // extern "C"
