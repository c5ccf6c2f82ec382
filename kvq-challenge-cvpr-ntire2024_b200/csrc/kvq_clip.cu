// CLIP ViT-B/16 visual tower with CLS adapters, as KSVQE uses it on its four 112x112 key frames
// (CLIP_extractor_addadapter_cls.forward, models/backbones/CLIP_backbone.py:156-202; ResidualAttentionBlock /
// QuickGELU / nn.MultiheadAttention from models/backbones/clip/model.py).  Third piece of the literal KSVQE key.
//
// Token rows are [n_img * L, width] (L = 1 + grid^2 = 50 tokens per image, CLS first), residual stream fp32.  Every
// Linear is the tcgen05 GEMM of kvq_gemm.cu (fp16 operands, fp32 accumulate; bias / QuickGELU / residual epilogues);
// the 50-token attention, the CLS adapters, the embedding assembly and the cosine map are small CUDA-core kernels
// (16 M of the tower's 1.1 G MACs per image).
#include <cmath>

#include "../../include/kvq_b200.h"
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

using namespace kvq;

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x[i, 0, :] = class_embedding + pos[0];  x[i, 1 + p, :] = patch[i, p, :] + pos[1 + p]   (:163-168)
__global__ void __launch_bounds__(256)
clip_embed_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                  float* __restrict__ x, int L, int C, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const long long row = idx / C;
  const int t = static_cast<int>(row % L);
  const long long i = row / L;
  const float v = t == 0 ? cls[c] : patch[(i * (L - 1) + (t - 1)) * C + c];
  x[idx] = v + pos[t * C + c];
}

// in-place fp32 LayerNorm over rows (ln_pre, :169): one warp per row, two-pass statistics
__global__ void __launch_bounds__(256)
ln_rows_f32_kernel(float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b, int rows, int C,
                   float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* r = x + static_cast<size_t>(row) * C;
  float s = 0.f;
  for (int i = lane; i < C; i += 32) s += r[i];
  const float mean = warp_sum_f(s) / static_cast<float>(C);
  float q = 0.f;
  for (int i = lane; i < C; i += 32) { const float d = r[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum_f(q) / static_cast<float>(C) + eps);
  for (int i = lane; i < C; i += 32) r[i] = (r[i] - mean) * rstd * g[i] + b[i];
}

// nn.MultiheadAttention core for one (image, head): softmax(q k^T / sqrt(hd)) v over L <= 64 tokens, hd = 64.
// qkv f16 [rows, 3*C] (q | k | v, head h at columns h*64); out f16 [rows, C].  Thread t < L owns query row t.
__global__ void __launch_bounds__(64)
clip_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int L, int C) {
  __shared__ float sk[64][65], sv[64][65];
  const int img = blockIdx.x, h = blockIdx.y, t = threadIdx.x;
  const size_t row0 = static_cast<size_t>(img) * L;
  for (int i = threadIdx.x; i < L * 64; i += blockDim.x) {
    const int r = i >> 6, d = i & 63;
    const __half* base = qkv + (row0 + r) * 3 * C + h * 64 + d;
    sk[r][d] = __half2float(base[C]);
    sv[r][d] = __half2float(base[2 * C]);
  }
  __syncthreads();
  if (t >= L) return;
  float q[64];
  const __half* qp = qkv + (row0 + t) * 3 * C + h * 64;
#pragma unroll
  for (int d = 0; d < 64; ++d) q[d] = __half2float(qp[d]) * 0.125f;          // head_dim^-0.5
  float s[64];
  float m = -INFINITY;
  for (int j = 0; j < L; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 64; ++d) a = fmaf(q[d], sk[j][d], a);
    s[j] = a;
    m = fmaxf(m, a);
  }
  float l = 0.f;
  for (int j = 0; j < L; ++j) { s[j] = __expf(s[j] - m); l += s[j]; }
  const float inv = 1.0f / l;
  __half* op = out + (row0 + t) * C + h * 64;
  for (int d0 = 0; d0 < 64; d0 += 8) {
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < L; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(s[j], sv[j][d0 + e], o[e]);
    }
    uint4 v;
    v.x = pack_half2(o[0] * inv, o[1] * inv); v.y = pack_half2(o[2] * inv, o[3] * inv);
    v.z = pack_half2(o[4] * inv, o[5] * inv); v.w = pack_half2(o[6] * inv, o[7] * inv);
    *reinterpret_cast<uint4*>(op + d0) = v;
  }
}

// cls <- 0.5 * relu(W2 relu(W1 cls + b1) + b2) + 0.5 * cls on the CLS row of every image (:184-191); fp32 weights
// W1 [C/4, C], W2 [C, C/4].  One block per image.
__global__ void __launch_bounds__(256)
clip_cls_adapter_kernel(float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                        const float* __restrict__ w2, const float* __restrict__ b2, int L, int C) {
  extern __shared__ float sh[];          // cls [C] | hidden [C/4]
  float* scls = sh;
  float* shid = sh + C;
  float* row = x + static_cast<size_t>(blockIdx.x) * L * C;
  const int H = C / 4, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < C; i += blockDim.x) scls[i] = row[i];
  __syncthreads();
  for (int j = warp; j < H; j += nw) {
    float a = 0.f;
    for (int i = lane; i < C; i += 32) a = fmaf(w1[static_cast<size_t>(j) * C + i], scls[i], a);
    a = warp_sum_f(a);
    if (lane == 0) shid[j] = fmaxf(a + b1[j], 0.f);
  }
  __syncthreads();
  for (int j = warp; j < C; j += nw) {
    float a = 0.f;
    for (int i = lane; i < H; i += 32) a = fmaf(w2[static_cast<size_t>(j) * H + i], shid[i], a);
    a = warp_sum_f(a);
    if (lane == 0) row[j] = 0.5f * fmaxf(a + b2[j], 0.f) + 0.5f * scls[j];
  }
}

// cls_attn[i, p] = cosine_similarity(x[i, 0], x[i, 1 + p]) (:198, eps 1e-8 on each norm); one warp per (i, p)
__global__ void __launch_bounds__(256)
clip_cosine_kernel(const float* __restrict__ x, float* __restrict__ out, int n_img, int L, int C) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_img * (L - 1)) return;
  const int i = w / (L - 1), p = w - i * (L - 1), lane = threadIdx.x & 31;
  const float* a = x + static_cast<size_t>(i) * L * C;
  const float* b = a + static_cast<size_t>(1 + p) * C;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  for (int k = lane; k < C; k += 32) { ab = fmaf(a[k], b[k], ab); aa = fmaf(a[k], a[k], aa); bb = fmaf(b[k], b[k], bb); }
  ab = warp_sum_f(ab); aa = warp_sum_f(aa); bb = warp_sum_f(bb);
  if (lane == 0) out[w] = ab / (fmaxf(sqrtf(aa), 1e-8f) * fmaxf(sqrtf(bb), 1e-8f));
}

struct ClipPlan {
  int grid, L, rows;
  size_t x_bytes, a16_bytes, big_bytes, total;
};

int make_clip_plan(const KvqClipConfig* cfg, int n_img, int H, int W, ClipPlan* pl) {
  KVQ_REQUIRE(cfg != nullptr, KVQ_ERR_BAD_SHAPE, "config is NULL");
  KVQ_REQUIRE(cfg->width == 768 && cfg->heads * 64 == cfg->width && cfg->patch == 16 && cfg->layers >= 1 &&
                  cfg->adapter_from >= 0 && cfg->adapter_from <= cfg->layers,
              KVQ_ERR_BAD_SHAPE, "clip: built for ViT-B/16 (width 768, 12 heads of 64, patch 16); got width=%d heads=%d "
              "patch=%d layers=%d adapter_from=%d", cfg->width, cfg->heads, cfg->patch, cfg->layers, cfg->adapter_from);
  KVQ_REQUIRE(n_img >= 1 && H == W && H % cfg->patch == 0 && H / cfg->patch >= 1 && (H / cfg->patch) * (W / cfg->patch) < 64,
              KVQ_ERR_BAD_SHAPE, "clip: %d images of %dx%d (square, multiple of 16, at most 63 patch tokens)", n_img, H, W);
  pl->grid = H / cfg->patch;
  pl->L = 1 + pl->grid * pl->grid;
  pl->rows = n_img * pl->L;
  const size_t slack = 256 * 3072 * 2;
  pl->x_bytes = align_up(static_cast<size_t>(pl->rows) * cfg->width * 4, 256);
  pl->a16_bytes = align_up(static_cast<size_t>(pl->rows) * cfg->width * 2 + slack, 256);
  pl->big_bytes = align_up(static_cast<size_t>(pl->rows) * 4 * cfg->width * 4 + slack, 256);   // patch fp32 / qkv / hidden
  pl->total = 2 * pl->a16_bytes + pl->big_bytes;
  return KVQ_OK;
}

int linear(int epi, const __half* a, const void* w, const void* b, void* out, const float* resid, int M, int N, int K,
           int quick, cudaStream_t st) {
  GemmParams gp{};
  gp.M = M; gp.N = N; gp.K = K;
  gp.bias = static_cast<const float*>(b);
  gp.out = out; gp.ldo = N;
  gp.resid = resid;
  gp.quick_gelu = quick;
  return launch_gemm(epi, a, K, static_cast<const __half*>(w), K, gp, st);
}

}  // namespace

extern "C" {

int kvq_clip_num_weights(const KvqClipConfig* cfg) {
  if (cfg == nullptr) return KVQ_ERR_BAD_SHAPE;
  return 5 + 12 * cfg->layers + 4 * (cfg->layers - cfg->adapter_from);
}

size_t kvq_clip_workspace_bytes(const KvqClipConfig* cfg, int n_img, int H, int W) {
  ClipPlan pl;
  if (make_clip_plan(cfg, n_img, H, W, &pl) != 0) return 0;
  return pl.total;
}

int kvq_clip_visual_forward(const KvqClipConfig* cfg, const void* const* weights, int num_weights, const float* images,
                            int n_img, int H, int W, float* cls_attn_out, float* tokens_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  ClipPlan pl;
  int rc = make_clip_plan(cfg, n_img, H, W, &pl);
  if (rc != 0) return rc;
  KVQ_REQUIRE(num_weights == kvq_clip_num_weights(cfg), KVQ_ERR_BAD_SHAPE, "clip: %d weight pointers, expected %d",
              num_weights, kvq_clip_num_weights(cfg));
  KVQ_REQUIRE(images && cls_attn_out && tokens_out && workspace && weights, KVQ_ERR_BAD_SHAPE, "clip: NULL argument");
  KVQ_REQUIRE(workspace_bytes >= pl.total, KVQ_ERR_WORKSPACE, "clip: workspace %zu < %zu bytes", workspace_bytes, pl.total);
  KVQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, KVQ_ERR_MISALIGNED, "workspace not 256 B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = cfg->width, L = pl.L, rows = pl.rows, P = cfg->patch;
  uint8_t* wp = static_cast<uint8_t*>(workspace);
  float* x = tokens_out;                                   // the residual stream IS the returned token tensor
  __half* a16 = reinterpret_cast<__half*>(wp); wp += pl.a16_bytes;
  __half* b16 = reinterpret_cast<__half*>(wp); wp += pl.a16_bytes;
  uint8_t* big = wp;
  int wi = 0;
  auto Wn = [&]() { return weights[wi++]; };

  // conv1 (16x16 / 16, no bias) as gather + GEMM -> fp32 patch embeddings (:158-162)
  const void* w_conv = Wn();
  const float* cls_emb = static_cast<const float*>(Wn());
  const float* pos = static_cast<const float*>(Wn());          // already resized to (grid x grid) + CLS at load time
  const float* lnpre_g = static_cast<const float*>(Wn());
  const float* lnpre_b = static_cast<const float*>(Wn());
  const int prow = n_img * (L - 1), Kc = 3 * P * P;
  rc = launch_im2col_stem(images, a16, n_img, 1, H, W, 1, P, P, 1, P, P, 0, 0, 0, Kc, st);
  if (rc != 0) return rc;
  float* patch = reinterpret_cast<float*>(big);
  rc = linear(EPI_RESID_F32, a16, w_conv, nullptr, patch, nullptr, prow, C, Kc, 0, st);
  if (rc != 0) return rc;
  {
    const long long total = static_cast<long long>(rows) * C;
    clip_embed_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(patch, cls_emb, pos, x, L, C, total);
    count_launch();
    ln_rows_f32_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, lnpre_g, lnpre_b, rows, C, 1e-5f);
    count_launch();
    KVQ_CUDA(cudaGetLastError());
  }
  for (int i = 0; i < cfg->layers; ++i) {
    const float* ln1g = static_cast<const float*>(Wn()); const float* ln1b = static_cast<const float*>(Wn());
    const void* w_in = Wn(); const void* b_in = Wn();
    const void* w_out = Wn(); const void* b_out = Wn();
    const float* ln2g = static_cast<const float*>(Wn()); const float* ln2b = static_cast<const float*>(Wn());
    const void* w_fc = Wn(); const void* b_fc = Wn();
    const void* w_pr = Wn(); const void* b_pr = Wn();
    // x = x + out_proj(attention(in_proj(ln_1(x))))   (ResidualAttentionBlock)
    rc = launch_ln_rows(x, a16, nullptr, ln1g, ln1b, 1e-5f, rows, C, L, st);
    if (rc != 0) return rc;
    __half* qkv = reinterpret_cast<__half*>(big);
    rc = linear(EPI_STORE_F16, a16, w_in, b_in, qkv, nullptr, rows, 3 * C, C, 0, st);
    if (rc != 0) return rc;
    clip_attn_kernel<<<dim3(n_img, cfg->heads), 64, 0, st>>>(qkv, b16, L, C);
    count_launch();
    KVQ_CUDA(cudaGetLastError());
    rc = linear(EPI_RESID_F32, b16, w_out, b_out, x, x, rows, C, C, 0, st);
    if (rc != 0) return rc;
    // x = x + c_proj(QuickGELU(c_fc(ln_2(x))))
    rc = launch_ln_rows(x, a16, nullptr, ln2g, ln2b, 1e-5f, rows, C, L, st);
    if (rc != 0) return rc;
    __half* hid = reinterpret_cast<__half*>(big);
    rc = linear(EPI_GELU_F16, a16, w_fc, b_fc, hid, nullptr, rows, 4 * C, C, 1, st);
    if (rc != 0) return rc;
    rc = linear(EPI_RESID_F32, hid, w_pr, b_pr, x, x, rows, C, 4 * C, 0, st);
    if (rc != 0) return rc;
    if (i >= cfg->adapter_from) {
      const float* aw1 = static_cast<const float*>(Wn()); const float* ab1 = static_cast<const float*>(Wn());
      const float* aw2 = static_cast<const float*>(Wn()); const float* ab2 = static_cast<const float*>(Wn());
      clip_cls_adapter_kernel<<<n_img, 256, (C + C / 4) * 4, st>>>(x, aw1, ab1, aw2, ab2, L, C);
      count_launch();
      KVQ_CUDA(cudaGetLastError());
    }
  }
  clip_cosine_kernel<<<(n_img * (L - 1) + 7) / 8, 256, 0, st>>>(x, cls_attn_out, n_img, L, C);
  count_launch();
  return check_cuda(cudaGetLastError(), "clip_cosine_kernel launch");
}

}  // extern "C"
