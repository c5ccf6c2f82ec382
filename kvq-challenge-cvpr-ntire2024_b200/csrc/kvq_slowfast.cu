// SlowFast-R50 motion-feature extractor (SlowFast_features.py:112-165; trunk = pytorchvideo slowfast_r50 blocks 0..4,
// restated in oracle/slowfast.py): host orchestration + C-ABI entry points.  Activations are channels-last fp16
// [B,T,H,W,C]; every convolution is (chunked im2col ->) the tcgen05 row GEMM with the folded-BatchNorm / residual /
// ReLU epilogue (EPI_CONV_F16).  The fast->slow lateral convolution writes straight into the extra channels of the
// slow pathway's stage output (row stride = slow channels + 2 * fast channels), so torch.cat never materialises.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "../../include/kvq_b200.h"
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

using namespace kvq;

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int conv_out(int n, int k, int s, int p) { return (n + 2 * p - k) / s + 1; }
inline int round64(int v) { return (v + 63) / 64 * 64; }

constexpr int SF_FUSE_KT = 7;          // conv_fast_to_slow kernel (7,1,1), stride (alpha,1,1), padding (3,0,0)
constexpr int SF_SLOW_KA[4] = {1, 1, 3, 3};
constexpr int SF_FAST_KA[4] = {3, 3, 3, 3};
constexpr size_t SF_SLACK = 256 * 4608 * 2;  // TMA tile overhang past the last row of a buffer

// im2col scratch per chunk (bounds the workspace; fewer, larger launches measured faster than L2-sized chunks)
size_t col_chunk_cap() {
  static size_t cap = [] {
    const char* e = std::getenv("KVQ_CONV_CHUNK_MB");
    long mb = e != nullptr ? std::atol(e) : 128;   // measured 16..128 MB at 16x32x256x256: 1692 / 1935 / 1990 / 1998 / 2022 clips/s
    if (mb < 4) mb = 4;
    return static_cast<size_t>(mb) << 20;
  }();
  return cap;
}

struct SfPlan {
  int Hs, Ws, Hp, Wp;   // after the stem conv / after the stem max pool
  int h5, w5;           // res5 map
  size_t col_bytes, slow_act, fast_act, pw_bytes, total;
};

int make_sf_plan(const KvqSlowFastConfig* cfg, int B, int Ts, int Tf, int H, int W, SfPlan* pl) {
  KVQ_REQUIRE(cfg != nullptr, KVQ_ERR_BAD_SHAPE, "config is NULL");
  KVQ_REQUIRE(B >= 1 && Ts >= 1 && Tf >= 1 && H >= 32 && W >= 32, KVQ_ERR_BAD_SHAPE,
              "slowfast: input B=%d Ts=%d Tf=%d %dx%d too small", B, Ts, Tf, H, W);
  KVQ_REQUIRE(cfg->alpha >= 1 && conv_out(Tf, SF_FUSE_KT, cfg->alpha, SF_FUSE_KT / 2) == Ts, KVQ_ERR_BAD_SHAPE,
              "slowfast: %d fast frames fuse to %d slow frames at alpha=%d, but the slow pathway has %d (torch.cat "
              "would raise)", Tf, conv_out(Tf, SF_FUSE_KT, std::max(cfg->alpha, 1), SF_FUSE_KT / 2), cfg->alpha, Ts);
  for (int s = 0; s < 4; ++s) KVQ_REQUIRE(cfg->depths[s] >= 1, KVQ_ERR_BAD_SHAPE, "depths[%d]=%d", s, cfg->depths[s]);
  pl->Hs = conv_out(H, 7, 2, 3); pl->Ws = conv_out(W, 7, 2, 3);
  pl->Hp = conv_out(pl->Hs, 3, 2, 1); pl->Wp = conv_out(pl->Ws, 3, 2, 1);
  const size_t ns = static_cast<size_t>(B) * Ts, nf = static_cast<size_t>(B) * Tf;
  KVQ_REQUIRE(nf * pl->Hs * pl->Ws < (1ull << 31), KVQ_ERR_BAD_SHAPE, "slowfast: %zu stem rows (int32 overflow)",
              nf * pl->Hs * pl->Ws);
  size_t sa = ns * pl->Hs * pl->Ws * 64 * 2, fa = nf * pl->Hs * pl->Ws * 8 * 2;
  size_t col = 0, max_row = 0;
  sa = std::max(sa, ns * pl->Hp * pl->Wp * 80 * 2);
  col = std::max(col, ns * pl->Hp * pl->Wp * static_cast<size_t>(round64(SF_FUSE_KT * 8)) * 2);
  int h = pl->Hp, w = pl->Wp, cs = 80, cf = 8;
  for (int s = 0; s < 4; ++s) {
    const int is = 64 << s, jf = 8 << s, stride = s == 0 ? 1 : 2;
    const int ho = conv_out(h, 3, stride, 1), wo = conv_out(w, 3, stride, 1);
    const size_t rs = ns * h * w, rso = ns * ho * wo, rf = nf * h * w, rfo = nf * ho * wo;
    const int cs_out = 4 * is + (s < 3 ? 8 * jf : 0);
    sa = std::max({sa, rs * is * 2, rso * static_cast<size_t>(cs_out) * 2});
    fa = std::max({fa, rf * jf * 2, rfo * static_cast<size_t>(4 * jf) * 2});
    // conv_a (ka,1,1) at the input resolution, conv_b (1,3,3) and the strided 1x1 gather at the output resolution
    const size_t ka_s = round64(SF_SLOW_KA[s] * std::max(cs, 4 * is));
    const size_t ka_f = round64(SF_FAST_KA[s] * std::max(cf, 4 * jf));
    if (SF_SLOW_KA[s] > 1) { col = std::max(col, rs * ka_s * 2); max_row = std::max(max_row, ka_s * 2); }
    col = std::max({col, rf * ka_f * 2, rso * static_cast<size_t>(9 * is) * 2,
                    rfo * static_cast<size_t>(round64(9 * jf)) * 2});
    max_row = std::max({max_row, ka_f * 2, static_cast<size_t>(9 * is) * 2});
    if (stride == 2) col = std::max({col, rso * static_cast<size_t>(round64(cs)) * 2,
                                     rfo * static_cast<size_t>(round64(cf)) * 2});
    if (s < 3) col = std::max(col, rso * static_cast<size_t>(round64(SF_FUSE_KT * 4 * jf)) * 2);
    cs = cs_out; cf = 4 * jf;
    h = ho; w = wo;
  }
  pl->h5 = h; pl->w5 = w;
  col = std::min(col, std::max(col_chunk_cap(), 128 * max_row));
  pl->col_bytes = align_up(col + SF_SLACK, 256);
  pl->slow_act = align_up(sa + SF_SLACK, 256);
  pl->fast_act = align_up(fa + SF_SLACK, 256);
  pl->pw_bytes = align_up(static_cast<size_t>(Ts + Tf) * h * w * 4, 256);
  pl->total = pl->col_bytes + 5 * pl->slow_act + 5 * pl->fast_act + pl->pw_bytes;
  return KVQ_OK;
}

struct ConvW {
  const __half* w;
  const float* b;
  const __half* wf = nullptr;   // twin for layers with 8 / 16 / 32 input channels: the row-folded block-diagonal weight
  const float* bf = nullptr;    // (pointwise, see fold_rows()) or the packed smem image of the narrow implicit GEMM
};

// A pointwise convolution with C < 64 input channels makes every 64-wide TMA box mostly out of bounds and every
// 128-row tile nearly empty (measured: a 2 M-row K = 8 GEMM takes 3x longer than K = 64).  g = 64 / C consecutive
// rows of the contiguous [M, C] activation ARE one row of a [M/g, 64] matrix; multiplying it by the block-diagonal
// weight [g*N, 64] (W on the diagonal blocks, packed once at load time) yields [M/g, g*N] = the same bytes as the
// [M, N] output.  The extra MMA work is on zeros and the tensor pipe is idle here anyway.
inline int fold_rows(int C) { return (C == 8 || C == 16 || C == 32) ? 64 / C : 1; }

struct SfCtx {
  __half* col;
  size_t col_bytes;
  cudaStream_t st;
};

int gemm_conv(const __half* A, int lda, const ConvW& cw, int Np, int nvalid, int K, const __half* resid, int ldr,
              __half* out, int ldo, long long M, bool relu, int stage, cudaStream_t st) {
  GemmParams gp{};
  gp.M = static_cast<int>(M); gp.N = Np; gp.K = K;
  gp.bias = cw.b;
  gp.out = out; gp.ldo = ldo;
  gp.resid_h = resid; gp.ldr = ldr;
  gp.relu = relu ? 1 : 0;
  gp.nvalid = nvalid == Np ? 0 : nvalid;
  ProfScope ps(PK_CONV_GEMM, stage, st);
  return launch_gemm(EPI_CONV_F16, A, lda, cw.w, round64(K), gp, st);   // packed weights: row stride round64(K)
}

// One convolution on a channels-last activation [B,T,H,W,C] (row stride C): 1x1x1 / stride 1 is the row GEMM itself;
// anything else gathers the patch matrix in L2-sized row chunks and runs the GEMM per chunk.
//   out[m, 0:cout] (row stride ldo) = act(conv + bias (+ resid[m, 0:cout])); weights [round64(cout), round64(K)]
int conv_op(const SfCtx& cx, const __half* in, int B, int T, int H, int W, int C, const int k[3], const int s[3],
            const int p[3], const ConvW& cw, int cout, const __half* resid, int ldr, __half* out, int ldo, bool relu,
            int stage) {
  const int Np = round64(cout);
  const int To = conv_out(T, k[0], s[0], p[0]), Ho = conv_out(H, k[1], s[1], p[1]), Wo = conv_out(W, k[2], s[2], p[2]);
  const long long M = static_cast<long long>(B) * To * Ho * Wo;
  const bool pointwise = k[0] * k[1] * k[2] == 1 && s[0] * s[1] * s[2] == 1;
  if (pointwise) {
    const int g = fold_rows(C);
    if (g > 1 && cw.wf != nullptr && M % g == 0 && ldo == cout && (resid == nullptr || ldr == cout) &&
        (g * cout) % 64 == 0) {
      ConvW f{cw.wf, cw.bf};
      return gemm_conv(in, 64, f, g * cout, g * cout, 64, resid, g * cout, out, g * cout, M / g, relu, stage, cx.st);
    }
    return gemm_conv(in, C, cw, Np, cout, C, resid, ldr, out, ldo, M, relu, stage, cx.st);
  }
  if (conv_implicit_supported(C, k[0], k[1], k[2], s[0], s[1], s[2])) {
    // C >= 64: the patch matrix never exists -- the GEMM's TMA producer walks (tap, channel block) over the activation
    GemmParams gp{};
    gp.N = Np;
    gp.bias = cw.b;
    gp.out = out; gp.ldo = ldo;
    gp.resid_h = resid; gp.ldr = ldr;
    gp.relu = relu ? 1 : 0;
    gp.nvalid = cout == Np ? 0 : cout;
    ProfScope ps(PK_CONV_GEMM, stage, cx.st);
    return launch_conv_implicit(in, B, T, H, W, C, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2], cw.w, gp,
                                cx.st);
  }
  if (cw.wf != nullptr && conv_narrow_supported(C, cout)) {
    // 8 / 16 / 32 input channels: implicit GEMM whose K blocks are 64 / C taps (kvq_gemm.cu, narrow mode)
    GemmParams gp{};
    gp.N = 64;
    gp.bias = cw.bf;
    gp.out = out; gp.ldo = ldo;
    gp.resid_h = resid; gp.ldr = ldr;
    gp.relu = relu ? 1 : 0;
    gp.nvalid = cout == 64 ? 0 : cout;
    ProfScope ps(PK_CONV_GEMM, stage, cx.st);
    return launch_conv_narrow(in, B, T, H, W, C, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2], cw.wf, gp, cx.st);
  }
  // gathered K is padded to whole 64-wide TMA boxes (weights are packed with the same zero columns)
  const int Kp = round64(k[0] * k[1] * k[2] * C);
  long long chunk = static_cast<long long>((cx.col_bytes - SF_SLACK) / (static_cast<size_t>(Kp) * 2));
  chunk = chunk / 128 * 128;
  KVQ_REQUIRE(chunk >= 128, KVQ_ERR_WORKSPACE, "slowfast: im2col scratch of %zu bytes cannot hold 128 rows of K=%d",
              cx.col_bytes, Kp);
  for (long long m0 = 0; m0 < M; m0 += chunk) {
    const long long rows = std::min(chunk, M - m0);
    int rc;
    {
      ProfScope ps(PK_CONV_IM2COL, stage, cx.st);
      rc = launch_im2col_cl(in, cx.col, B, T, H, W, C, k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2], Kp, cx.st,
                            m0, rows);
    }
    if (rc != 0) return rc;
    rc = gemm_conv(cx.col, Kp, cw, Np, cout, Kp, resid != nullptr ? resid + m0 * ldr : nullptr, ldr, out + m0 * ldo,
                   ldo, rows, relu, stage, cx.st);
    if (rc != 0) return rc;
  }
  return KVQ_OK;
}

// stem: fp32 NCDHW input -> conv (kt,7,7)/s(1,2,2) + BN + ReLU -> [B,T,Hs,Ws,cout] fp16 (implicit GEMM, kvq_stem.cu)
int stem_op(const SfCtx& cx, const float* x, int B, int T, int H, int W, int kt, const ConvW& cw, int cout, __half* out) {
  ProfScope ps(PK_CONV_STEM, 0, cx.st);
  return launch_stem_conv(x, cw.w, cw.b, out, B, T, H, W, kt, cout, cx.st);
}

struct Pathway {
  __half* act[5];
  int ci;   // index of the buffer holding the current activation
  int C;    // its channels (= row stride)
  int T;
};

// one res stage of one pathway (ResStage of pytorchvideo; oracle/slowfast.py:_res_block).  The last block writes its
// output with row stride `last_ld` (>= 4*inner) so that the lateral fusion can append channels in place.
int res_stage(const SfCtx& cx, Pathway& P, const void* const* weights, int& wi, int B, int h, int w, int inner, int ka,
              int depth, int stride, int last_ld, int stage) {
  const int one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
  const int ho = conv_out(h, 3, stride, 1), wo = conv_out(w, 3, stride, 1);
  const int cout = 4 * inner;
  auto next = [&](bool folded = false) {
    ConvW c{static_cast<const __half*>(weights[wi]), static_cast<const float*>(weights[wi + 1])};
    wi += 2;
    if (folded) {
      c.wf = static_cast<const __half*>(weights[wi]);
      c.bf = static_cast<const float*>(weights[wi + 1]);
      wi += 2;
    }
    return c;
  };
  for (int j = 0; j < depth; ++j) {
    const int sj = j == 0 ? stride : 1;
    const int hi = j == 0 ? h : ho, wj = j == 0 ? w : wo;
    __half* cur = P.act[P.ci];
    __half* t1 = P.act[(P.ci + 1) % 5];
    __half* t2 = P.act[(P.ci + 2) % 5];
    __half* idn = P.act[(P.ci + 3) % 5];
    __half* out = P.act[(P.ci + 4) % 5];
    const int ldo = j == depth - 1 ? last_ld : cout;
    const __half* resid = cur;
    int ldr = P.C;
    int rc;
    if (j == 0) {
      const ConvW c1 = next(fold_rows(P.C) > 1);   // stride 1: row-folded twin; stride 2: narrow-conv image
      const int s1[3] = {1, sj, sj};
      rc = conv_op(cx, cur, B, P.T, hi, wj, P.C, one, s1, zero, c1, cout, nullptr, 0, idn, cout, false, stage);
      if (rc != 0) return rc;
      resid = idn; ldr = cout;
    }
    const ConvW ca = next(ka > 1 && fold_rows(P.C) > 1), cb = next(fold_rows(inner) > 1), cc = next(fold_rows(inner) > 1);
    const int kA[3] = {ka, 1, 1}, pA[3] = {ka / 2, 0, 0};
    rc = conv_op(cx, cur, B, P.T, hi, wj, P.C, kA, one, pA, ca, inner, nullptr, 0, t1, inner, true, stage);
    if (rc != 0) return rc;
    const int kB[3] = {1, 3, 3}, sB[3] = {1, sj, sj}, pB[3] = {0, 1, 1};
    rc = conv_op(cx, t1, B, P.T, hi, wj, inner, kB, sB, pB, cb, inner, nullptr, 0, t2, inner, true, stage);
    if (rc != 0) return rc;
    rc = conv_op(cx, t2, B, P.T, ho, wo, inner, one, one, zero, cc, cout, resid, ldr, out, ldo, true, stage);
    if (rc != 0) return rc;
    P.ci = (P.ci + 4) % 5;
    P.C = ldo;   // the extra channels (if any) are filled by the fusion that follows
  }
  return KVQ_OK;
}

}  // namespace

extern "C" {

int kvq_slowfast_num_weights(const KvqSlowFastConfig* cfg) {
  if (cfg == nullptr) return KVQ_ERR_BAD_SHAPE;
  int n = 6;
  for (int s = 0; s < 4; ++s) {
    n += 2 * 2 * (3 * cfg->depths[s] + 1) + (s < 3 ? 2 : 0);
    // twins of the fast pathway's narrow layers (see the header): per block conv_b + conv_c while inner < 64, conv_a
    // while its input has < 64 channels, branch1 likewise, and the fusions that read < 64 channels
    const int inner = 8 << s, cin0 = s == 0 ? 8 : 4 * (8 << (s - 1));
    if (inner < 64) n += 2 * 2 * cfg->depths[s];
    if (cin0 < 64) n += 2 + 2;                                       // branch1 + conv_a of block 0
    if (4 * inner < 64) n += 2 * (cfg->depths[s] - 1);               // conv_a of the other blocks
    if (s < 3 && 4 * inner < 64) n += 2;                             // the stage's fusion
  }
  return n + 2;                                                      // block-0 fusion (8 channels)
}

size_t kvq_slowfast_workspace_bytes(const KvqSlowFastConfig* cfg, int B, int Ts, int Tf, int H, int W) {
  SfPlan pl;
  if (make_sf_plan(cfg, B, Ts, Tf, H, W, &pl) != 0) return 0;
  return pl.total;
}

int kvq_slow_frame_indices(int T, int alpha, int32_t* out, int cap) {
  KVQ_REQUIRE(T >= 1 && alpha >= 1 && out != nullptr, KVQ_ERR_BAD_SHAPE, "slow_frame_indices: T=%d alpha=%d", T, alpha);
  const int n = T / alpha;
  KVQ_REQUIRE(n >= 1 && n <= cap, KVQ_ERR_BAD_SHAPE, "slow_frame_indices: %d indices, room for %d", n, cap);
  // torch.linspace(0, T-1, n) in fp32 (step = (T-1)/(n-1); the upper half is evaluated from the end), then .long()
  const float step = n > 1 ? static_cast<float>(T - 1) / static_cast<float>(n - 1) : 0.f;
  for (int i = 0; i < n; ++i) {
    const float v = i < n / 2 ? step * static_cast<float>(i) : static_cast<float>(T - 1) - step * static_cast<float>(n - 1 - i);
    out[i] = n == 1 ? 0 : static_cast<int32_t>(v);   // steps == 1: torch returns [start]
  }
  return n;
}

int kvq_pack_pathway_slow_f32(const float* frames, float* slow_out, int B, int T, int H, int W, int alpha,
                              void* stream) {
  KVQ_REQUIRE(frames && slow_out, KVQ_ERR_BAD_SHAPE, "pack_pathway: NULL argument");
  int32_t idx[64];
  const int n = kvq_slow_frame_indices(T, alpha, idx, 64);
  if (n < 0) return n;
  return launch_select_frames(frames, slow_out, B * 3, T, static_cast<long long>(H) * W, idx, n,
                              static_cast<cudaStream_t>(stream));
}

int kvq_slowfast_forward(const KvqSlowFastConfig* cfg, const void* const* weights, int num_weights, const float* slow,
                         const float* fast, int B, int Ts, int Tf, int H, int W, float* slow_out, float* fast_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  SfPlan pl;
  int rc = make_sf_plan(cfg, B, Ts, Tf, H, W, &pl);
  if (rc != 0) return rc;
  KVQ_REQUIRE(num_weights == kvq_slowfast_num_weights(cfg), KVQ_ERR_BAD_SHAPE,
              "slowfast: %d weight pointers, expected %d", num_weights, kvq_slowfast_num_weights(cfg));
  KVQ_REQUIRE(slow && fast && slow_out && fast_out && workspace && weights, KVQ_ERR_BAD_SHAPE,
              "slowfast: NULL argument");
  KVQ_REQUIRE(workspace_bytes >= pl.total, KVQ_ERR_WORKSPACE, "slowfast: workspace %zu < %zu bytes", workspace_bytes,
              pl.total);
  KVQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, KVQ_ERR_MISALIGNED, "workspace not 256 B aligned");
  for (int i = 0; i < 3; ++i)
    KVQ_REQUIRE(cfg->slow_pool[i] >= 1 && cfg->fast_pool[i] >= 1, KVQ_ERR_BAD_SHAPE, "slowfast: pool kernel < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* wp = static_cast<uint8_t*>(workspace);
  SfCtx cx{reinterpret_cast<__half*>(wp), pl.col_bytes, st};
  wp += pl.col_bytes;
  Pathway S{}, F{};
  for (int i = 0; i < 5; ++i) { S.act[i] = reinterpret_cast<__half*>(wp); wp += pl.slow_act; }
  for (int i = 0; i < 5; ++i) { F.act[i] = reinterpret_cast<__half*>(wp); wp += pl.fast_act; }
  float* pw_slow = reinterpret_cast<float*>(wp);
  float* pw_fast = pw_slow + static_cast<size_t>(Ts) * pl.h5 * pl.w5;
  S.T = Ts; F.T = Tf;
  int wi = 0;
  auto next = [&](bool twin = false) {
    ConvW c{static_cast<const __half*>(weights[wi]), static_cast<const float*>(weights[wi + 1])};
    wi += 2;
    if (twin) {
      c.wf = static_cast<const __half*>(weights[wi]);
      c.bf = static_cast<const float*>(weights[wi + 1]);
      wi += 2;
    }
    return c;
  };
  const int one[3] = {1, 1, 1};

  // block 0: stems (conv + BN + ReLU, max pool (1,3,3)/s(1,2,2)/p(0,1,1)); the slow pool output leaves room for the
  // 16 fused channels
  const ConvW ws = next(), wf = next(), wfuse0 = next(true);   // (8-channel fusion input: narrow-conv image twin)
  rc = stem_op(cx, slow, B, Ts, H, W, 1, ws, 64, S.act[0]);
  if (rc != 0) return rc;
  {
    ProfScope ps(PK_CONV_POOL, 0, st);
    rc = launch_maxpool_hw(S.act[0], S.act[1], B * Ts, pl.Hs, pl.Ws, 64, st, 80);
  }
  if (rc != 0) return rc;
  rc = stem_op(cx, fast, B, Tf, H, W, 5, wf, 8, F.act[0]);
  if (rc != 0) return rc;
  {
    ProfScope ps(PK_CONV_POOL, 0, st);
    rc = launch_maxpool_hw(F.act[0], F.act[1], B * Tf, pl.Hs, pl.Ws, 8, st, 8);
  }
  if (rc != 0) return rc;
  S.ci = 1; S.C = 80;
  F.ci = 1; F.C = 8;
  const int kF[3] = {SF_FUSE_KT, 1, 1}, sF[3] = {cfg->alpha, 1, 1}, pF[3] = {SF_FUSE_KT / 2, 0, 0};
  // multipathway_fusion: relu(bn(conv_fast_to_slow(fast))) -> channels [64, 80) of the slow activation
  rc = conv_op(cx, F.act[F.ci], B, Tf, pl.Hp, pl.Wp, 8, kF, sF, pF, wfuse0, 16, nullptr, 0, S.act[S.ci] + 64, 80, true, 0);
  if (rc != 0) return rc;

  int h = pl.Hp, w = pl.Wp;
  for (int s = 0; s < 4; ++s) {
    const int is = 64 << s, jf = 8 << s, stride = s == 0 ? 1 : 2;
    const int slow_ld = 4 * is + (s < 3 ? 8 * jf : 0);
    rc = res_stage(cx, S, weights, wi, B, h, w, is, SF_SLOW_KA[s], cfg->depths[s], stride, slow_ld, s);
    if (rc != 0) return rc;
    rc = res_stage(cx, F, weights, wi, B, h, w, jf, SF_FAST_KA[s], cfg->depths[s], stride, 4 * jf, s);
    if (rc != 0) return rc;
    h = conv_out(h, 3, stride, 1); w = conv_out(w, 3, stride, 1);
    if (s < 3) {
      const ConvW wfu = next(fold_rows(4 * jf) > 1);
      rc = conv_op(cx, F.act[F.ci], B, Tf, h, w, 4 * jf, kF, sF, pF, wfu, 8 * jf, nullptr, 0, S.act[S.ci] + 4 * is,
                   slow_ld, true, s);
      if (rc != 0) return rc;
    }
  }
  (void)one;
  // blocks[5].pool[i] = AvgPool3d(kernel, stride 1) followed by blocks[6].output_pool = AdaptiveAvgPool3d(1)
  // (SlowFast_features.py:148-163) = one weighted mean over the res5 map
  {
    ProfScope ps(PK_CONV_POOL, 3, st);
    rc = launch_pool_window_weights(pw_slow, Ts, h, w, cfg->slow_pool[0], cfg->slow_pool[1], cfg->slow_pool[2], st);
    if (rc != 0) return rc;
    rc = launch_pool_window_weights(pw_fast, Tf, h, w, cfg->fast_pool[0], cfg->fast_pool[1], cfg->fast_pool[2], st);
    if (rc != 0) return rc;
    rc = launch_pool_stats(S.act[S.ci], pw_slow, slow_out, nullptr, B, Ts * h * w, S.C, S.C, st);
    if (rc != 0) return rc;
    rc = launch_pool_stats(F.act[F.ci], pw_fast, fast_out, nullptr, B, Tf * h * w, F.C, F.C, st);
    if (rc != 0) return rc;
  }
  return KVQ_OK;
}

}  // extern "C"
