// C-ABI entry points (include/kvq_b200.h) and the host-side orchestration of the Swin3D-GRPB + VQAHead forward.
// Everything is enqueued on the caller's stream; the library allocates nothing.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "../../include/kvq_b200.h"
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

using namespace kvq;

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct StageDims {
  int D, H, W, C, heads;
};

struct Plan {
  int D0, H0, W0;
  StageDims st[KVQ_MAX_STAGES];
  size_t x_bytes, x2_bytes, a16_bytes, img_bytes, hid_bytes, rs_bytes;
  size_t total;
};

int validate_cfg(const KvqSwinConfig* cfg) {
  KVQ_REQUIRE(cfg != nullptr, KVQ_ERR_BAD_SHAPE, "config is NULL");
  KVQ_REQUIRE(cfg->num_stages >= 1 && cfg->num_stages <= KVQ_MAX_STAGES, KVQ_ERR_BAD_SHAPE, "num_stages=%d",
              cfg->num_stages);
  KVQ_REQUIRE(cfg->embed_dim == 96, KVQ_ERR_BAD_SHAPE,
              "embed_dim=%d: the fused patch-embed LayerNorm epilogue is built for 96 channels", cfg->embed_dim);
  for (int s = 0; s < cfg->num_stages; ++s) {
    const int C = cfg->embed_dim << s;
    KVQ_REQUIRE(cfg->depths[s] >= 1, KVQ_ERR_BAD_SHAPE, "depths[%d]=%d", s, cfg->depths[s]);
    KVQ_REQUIRE(cfg->num_heads[s] * 32 == C, KVQ_ERR_BAD_SHAPE, "stage %d: heads=%d x 32 != C=%d", s,
                cfg->num_heads[s], C);
  }
  KVQ_REQUIRE(cfg->window[0] >= 1 && cfg->window[0] <= 8 && cfg->window[1] >= 1 && cfg->window[1] <= 7 &&
                  cfg->window[2] >= 1 && cfg->window[2] <= 7,
              KVQ_ERR_BAD_SHAPE, "window (%d,%d,%d) exceeds (8,7,7)", cfg->window[0], cfg->window[1], cfg->window[2]);
  KVQ_REQUIRE(cfg->head_hidden == 0 || cfg->head_hidden == 64, KVQ_ERR_BAD_SHAPE, "head_hidden=%d (0 or 64)",
              cfg->head_hidden);
  const int32_t* rw = cfg->resized_window;
  KVQ_REQUIRE((rw[0] == 0 && rw[1] == 0 && rw[2] == 0) ||
                  (rw[0] >= 1 && rw[0] <= cfg->window[0] && rw[1] >= 1 && rw[1] <= cfg->window[1] && rw[2] >= 1 &&
                   rw[2] <= cfg->window[2]),
              KVQ_ERR_BAD_SHAPE, "resized_window (%d,%d,%d) must be 0,0,0 or within window (%d,%d,%d)", rw[0], rw[1],
              rw[2], cfg->window[0], cfg->window[1], cfg->window[2]);
  return KVQ_OK;
}

int make_plan(const KvqSwinConfig* cfg, int B, int T, int H, int W, Plan* pl) {
  KVQ_REQUIRE(B >= 1 && T >= 1 && H >= 1 && W >= 1, KVQ_ERR_BAD_SHAPE, "empty input %dx3x%dx%dx%d", B, T, H, W);
  pl->D0 = (T + 1) / 2;
  pl->H0 = (H + 3) / 4;
  pl->W0 = (W + 3) / 4;
  int D = pl->D0, Hh = pl->H0, Ww = pl->W0;
  size_t max_a16 = 0, max_img = 0, max_hid = 0, max_x2 = 0;
  const int zero[3] = {0, 0, 0};
  for (int s = 0; s < cfg->num_stages; ++s) {
    const int C = cfg->embed_dim << s;
    pl->st[s] = {D, Hh, Ww, C, cfg->num_heads[s]};
    const WinGeom g = make_geom(D, Hh, Ww, cfg->resized_window[0] > 0 ? cfg->resized_window : cfg->window, zero);  // padded sizes do not depend on the shift
    const size_t rows_w = static_cast<size_t>(B) * g.nW * g.N;
    const size_t rows = static_cast<size_t>(B) * g.tokens;
    KVQ_REQUIRE(rows_w < (1ull << 31), KVQ_ERR_BAD_SHAPE, "stage %d has %zu window rows (int32 overflow)", s, rows_w);
    max_a16 = std::max(max_a16, rows_w * C * 2);
    max_img = std::max(max_img, static_cast<size_t>(B) * g.nW * cfg->num_heads[s] * ATT2_UNIT_BYTES);
    max_hid = std::max(max_hid, rows * 4 * C * 2);
    if (s + 1 < cfg->num_stages) {
      const int H2 = (Hh + 1) / 2, W2 = (Ww + 1) / 2;
      const size_t rows2 = static_cast<size_t>(B) * D * H2 * W2;
      max_a16 = std::max(max_a16, rows2 * 4 * C * 2);
      max_x2 = std::max(max_x2, rows2 * 2 * C * 4);
      Hh = H2;
      Ww = W2;
    }
  }
  const size_t rows0 = static_cast<size_t>(B) * pl->D0 * pl->H0 * pl->W0;
  pl->x_bytes = align_up(rows0 * cfg->embed_dim * 4, 256);
  pl->x2_bytes = align_up(std::max<size_t>(max_x2, 256), 256);
  pl->a16_bytes = align_up(std::max(max_a16, rows0 * 96 * 2) + 128 * 3072 * 2, 256);  // + slack for TMA tile overhang
  pl->img_bytes = align_up(max_img, 256);
  pl->hid_bytes = align_up(max_hid + 128 * 3072 * 2, 256);
  const StageDims& last = pl->st[cfg->num_stages - 1];
  pl->rs_bytes = align_up(static_cast<size_t>(B) * last.D * last.H * last.W * 4, 256);
  pl->total = pl->x_bytes + pl->x2_bytes + pl->a16_bytes + pl->img_bytes + pl->hid_bytes + pl->rs_bytes;
  return KVQ_OK;
}

int run_attention(const __half* xw, const __half* qkv_w, const float* qkv_b, const float* packed_tab, __half* out,
                  __half* img, int B, int C, int heads, const WinGeom& g, const int32_t base_win[3], int variant,
                  cudaStream_t st, int stage = 0, int rpi_geometric = 0, bool qkv_done = false) {
  const int rows_w = B * g.nW * g.N;
  GemmParams gp{};
  gp.M = rows_w; gp.N = 3 * C; gp.K = C;
  gp.bias = qkv_b;
  gp.geom = g;
  gp.img = img;
  gp.att_pitch = window_attn_pitch(g, base_win, variant);
  gp.C = C;
  gp.heads = heads;
  gp.qscale = attn_qscale();  // head_dim^-0.5 (:191), folded with log2(e) for the exp2 softmax
  int rc = 0;
  if (!qkv_done) {   // (the C = 96 stage of the fused forward writes the images from its LN1 + qkv kernel)
    ProfScope ps(PK_QKV_GEMM, stage, st);
    rc = launch_gemm(EPI_QKV_IMG, xw, C, qkv_w, C, gp, st);
  }
  if (rc != 0) return rc;
  AttnParams ap{};
  ap.img = img;
  ap.out = out;
  ap.packed_tab = packed_tab;
  ap.B = B; ap.C = C; ap.heads = heads;
  ap.shifted = (g.sd | g.sh | g.sw) != 0;
  ap.geom = g;
  ap.base_wd = base_win[0]; ap.base_wh = base_win[1]; ap.base_ww = base_win[2];
  ap.variant = variant;
  ap.rows_dfast = g.dfast;
  ap.rpi_geometric = rpi_geometric;
  ProfScope ps(PK_ATTN, stage, st);
  return launch_window_attn(ap, st);
}

}  // namespace

extern "C" {

int kvq_fragment_gather_u8(const uint8_t* frames, const int32_t* offsets, float* out, int B, int T, int Hs, int Ws,
                           int fragments_h, int fragments_w, int fsize, int aligned, const float mean[3],
                           const float std[3], void* stream) {
  KVQ_REQUIRE(frames && offsets && out && mean && std, KVQ_ERR_BAD_SHAPE, "fragment_gather: NULL argument");
  return launch_fragment_gather_u8(frames, offsets, out, B, T, Hs, Ws, fragments_h, fragments_w, fsize, aligned, mean,
                                   std, static_cast<cudaStream_t>(stream));
}

int kvq_qrs_select_gather(const float* fragment, const float* cls_attn, float* x_sel_out, int32_t* region_out, int B,
                          int T, int H, int W, int n_key, int L, int anchor, int region_patches, void* stream) {
  return launch_qrs_select_gather(fragment, cls_attn, x_sel_out, region_out, B, T, H, W, n_key, L, anchor, region_patches,
                                  static_cast<cudaStream_t>(stream));
}

size_t kvq_vqa_head_workspace_bytes(int B, int C, int tokens) {
  return align_up(static_cast<size_t>(B) * tokens * C * 2 + 128 * 3072 * 2, 256) +
         align_up(static_cast<size_t>(B) * tokens * 4, 256);
}

int kvq_vqa_head(const float* feat, const void* w1_f16, const float* b1, const float* w2, const float* b2,
                 float* score_out, int B, int C, int tokens, int hidden, void* workspace, size_t workspace_bytes,
                 void* stream) {
  KVQ_REQUIRE(feat && w1_f16 && b1 && w2 && b2 && score_out, KVQ_ERR_BAD_SHAPE, "vqa_head: NULL argument");
  KVQ_REQUIRE(hidden == 64 && C % 8 == 0 && B > 0 && tokens > 0, KVQ_ERR_BAD_SHAPE,
              "vqa_head: hidden=%d (64) C=%d (multiple of 8)", hidden, C);
  const size_t need = kvq_vqa_head_workspace_bytes(B, C, tokens);
  KVQ_REQUIRE(workspace != nullptr && workspace_bytes >= need, KVQ_ERR_WORKSPACE, "workspace %zu B < required %zu B",
              workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* rows = static_cast<__half*>(workspace);
  float* rowscore = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) +
                                             align_up(static_cast<size_t>(B) * tokens * C * 2 + 128 * 3072 * 2, 256));
  int rc = launch_cf_to_rows(feat, rows, B, C, tokens, st);
  if (rc != 0) return rc;
  GemmParams gp{};
  gp.M = B * tokens; gp.N = hidden; gp.K = C;
  gp.bias = b1; gp.w2 = w2; gp.b2ptr = b2; gp.rowscore = rowscore;
  rc = launch_gemm(EPI_HEAD, rows, C, static_cast<const __half*>(w1_f16), C, gp, st);
  if (rc != 0) return rc;
  return launch_row_mean(rowscore, score_out, B, tokens, st);
}

int kvq_debug_attn_timers(unsigned long long* out16, int reset) { return kvq::debug_attn_timers(out16, reset); }

long long kvq_launch_count(void) { return kvq::launch_count(); }

void kvq_profile_enable(int on) { kvq::prof_set(on != 0); }

int kvq_profile_num_categories(void) { return PK_COUNT * 4; }

const char* kvq_profile_category_name(int cat) {
  static const char* kinds[PK_COUNT] = {"embed_im2col", "embed_gemm", "ln_window", "qkv_gemm", "window_attn",
                                        "proj_gemm", "ln_rows", "fc1_gemm", "fc2_gemm", "merge_ln", "merge_gemm",
                                        "final_ln", "head", "fused_mlp", "conv_im2col", "conv_gemm", "conv_pool", "conv_stem"};
  static thread_local char buf[48];
  if (cat < 0 || cat >= PK_COUNT * 4) return "?";
  snprintf(buf, sizeof(buf), "%s.s%d", kinds[cat / 4], cat % 4);
  return buf;
}

int kvq_profile_collect(float* ms_per_category, int* launches_per_category, int num_categories) {
  return kvq::prof_collect(ms_per_category, launches_per_category, num_categories);
}

const char* kvq_last_error_string(void) { return kvq::last_error(); }

const char* kvq_build_info(void) { return "kvq_b200 sm_100a tcgen05/TMEM/TMA fp16-operand fp32-accumulate"; }

int kvq_swin3d_num_weights(const KvqSwinConfig* cfg) {
  if (validate_cfg(cfg) != 0) return KVQ_ERR_BAD_SHAPE;
  int n = 4;
  for (int s = 0; s < cfg->num_stages; ++s) n += 13 * cfg->depths[s];
  n += 3 * (cfg->num_stages - 1);
  n += 2;
  if (cfg->head_hidden > 0) n += 4;
  return n;
}

size_t kvq_swin3d_workspace_bytes(const KvqSwinConfig* cfg, int B, int T, int H, int W) {
  Plan pl;
  if (validate_cfg(cfg) != 0 || make_plan(cfg, B, T, H, W, &pl) != 0) return 0;
  return pl.total;
}

int kvq_swin3d_forward(const KvqSwinConfig* cfg, const void* const* weights, int num_weights, const float* x, int B,
                       int T, int H, int W, float* feat_out, float* score_out, void* workspace,
                       size_t workspace_bytes, void* stream) {
  return kvq_swin3d_forward_hooked(cfg, weights, num_weights, x, B, T, H, W, feat_out, score_out, workspace,
                                   workspace_bytes, stream, nullptr, nullptr);
}

static int swin3d_forward_impl(const KvqSwinConfig* cfg, const void* const* weights, int num_weights, const void* x,
                               int x_is_f16, int B, int T, int H, int W, float* feat_out, float* score_out,
                               void* workspace, size_t workspace_bytes, void* stream, kvq_stage_hook hook, void* hook_arg);

int kvq_swin3d_forward_hooked(const KvqSwinConfig* cfg, const void* const* weights, int num_weights, const float* x,
                              int B, int T, int H, int W, float* feat_out, float* score_out, void* workspace,
                              size_t workspace_bytes, void* stream, kvq_stage_hook hook, void* hook_arg) {
  return swin3d_forward_impl(cfg, weights, num_weights, x, 0, B, T, H, W, feat_out, score_out, workspace, workspace_bytes,
                             stream, hook, hook_arg);
}

int kvq_swin3d_forward_x16(const KvqSwinConfig* cfg, const void* const* weights, int num_weights, const void* x_f16,
                           int B, int T, int H, int W, float* feat_out, float* score_out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  return swin3d_forward_impl(cfg, weights, num_weights, x_f16, 1, B, T, H, W, feat_out, score_out, workspace,
                             workspace_bytes, stream, nullptr, nullptr);
}

static int swin3d_forward_impl(const KvqSwinConfig* cfg, const void* const* weights, int num_weights, const void* x,
                               int x_is_f16, int B, int T, int H, int W, float* feat_out, float* score_out,
                               void* workspace, size_t workspace_bytes, void* stream, kvq_stage_hook hook, void* hook_arg) {
  int rc = validate_cfg(cfg);
  if (rc != 0) return rc;
  Plan pl;
  rc = make_plan(cfg, B, T, H, W, &pl);
  if (rc != 0) return rc;
  KVQ_REQUIRE(num_weights == kvq_swin3d_num_weights(cfg), KVQ_ERR_BAD_SHAPE, "expected %d weight pointers, got %d",
              kvq_swin3d_num_weights(cfg), num_weights);
  for (int i = 0; i < num_weights; ++i)
    KVQ_REQUIRE(weights[i] != nullptr && (reinterpret_cast<uintptr_t>(weights[i]) & 15) == 0, KVQ_ERR_MISALIGNED,
                "weight %d is NULL or not 16 B aligned", i);
  KVQ_REQUIRE(workspace != nullptr && workspace_bytes >= pl.total, KVQ_ERR_WORKSPACE,
              "workspace %zu B < required %zu B", workspace_bytes, pl.total);
  KVQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              KVQ_ERR_MISALIGNED, "workspace must be 256 B aligned, x 16 B aligned");
  KVQ_REQUIRE(score_out == nullptr || cfg->head_hidden > 0, KVQ_ERR_BAD_SHAPE, "score requested but head_hidden == 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* xa = reinterpret_cast<float*>(ws); ws += pl.x_bytes;
  float* xb = reinterpret_cast<float*>(ws); ws += pl.x2_bytes;
  __half* a16 = reinterpret_cast<__half*>(ws); ws += pl.a16_bytes;
  __half* img = reinterpret_cast<__half*>(ws); ws += pl.img_bytes;
  __half* hid = reinterpret_cast<__half*>(ws); ws += pl.hid_bytes;
  float* rowscore = reinterpret_cast<float*>(ws);

  int wi = 0;
  auto WH = [&]() { return static_cast<const __half*>(weights[wi++]); };
  auto WF = [&]() { return static_cast<const float*>(weights[wi++]); };
  const float eps = cfg->ln_eps;

  // ---- PatchEmbed3D: one implicit-GEMM kernel (kvq_embed.cu); KVQ_EMBED_EXPLICIT=1 keeps the round-1 path
  //      im2col -> GEMM (K = 96) with bias + LayerNorm epilogue for comparison ----
  {
    static const bool explicit_embed = []() { const char* e = getenv("KVQ_EMBED_EXPLICIT"); return e && atoi(e) != 0; }();
    const __half* w = WH();
    const float* pe_bias = WF();
    const float* pe_gamma = WF();
    const float* pe_beta = WF();
    const int split = cfg->split_weights & 1;
    if (cfg->embed_dim == 96 && !explicit_embed) {
      ProfScope ps(PK_EMBED_GEMM, 0, st);
      rc = launch_patch_embed(x, x_is_f16, w, split ? 256 : 96, split, pe_bias, pe_gamma, pe_beta, eps, xa, B, T, H, W, st);
      if (rc != 0) return rc;
    } else {
      {
        ProfScope ps(PK_EMBED_IM2COL, 0, st);
        rc = launch_patch_im2col(x, x_is_f16, a16, B, T, H, W, st);
      }
      if (rc != 0) return rc;
      GemmParams gp{};
      gp.M = B * pl.D0 * pl.H0 * pl.W0; gp.N = 96; gp.K = 96;
      gp.bias = pe_bias; gp.gamma = pe_gamma; gp.beta = pe_beta; gp.eps = eps;
      gp.out = xa; gp.ldo = 96;
      gp.split_b = split;
      {
        ProfScope ps(PK_EMBED_GEMM, 0, st);
        rc = launch_gemm(EPI_LN_F32, a16, 96, w, gp.split_b ? 256 : 96, gp, st);
      }
      if (rc != 0) return rc;
    }
  }

  if (hook != nullptr) {   // feats[0] of SwinTransformer3D.forward (:1058): the patch-embedding output
    rc = hook(hook_arg, -1, xa, B * pl.D0 * pl.H0 * pl.W0, cfg->embed_dim, stream);
    KVQ_REQUIRE(rc == 0, rc < 0 ? rc : KVQ_ERR_BAD_SHAPE, "stage hook failed after the patch embedding (code %d)", rc);
  }
  float* xcur = xa;
  float* xnext = xb;
  const int shift_full[3] = {cfg->window[0] / 2, cfg->window[1] / 2, cfg->window[2] / 2};   // the BASE window's shift (:633)
  const bool adaptive = cfg->resized_window[0] > 0;
  const int shift_none[3] = {0, 0, 0};
  for (int s = 0; s < cfg->num_stages; ++s) {
    const StageDims& sd = pl.st[s];
    const int C = sd.C, M = B * sd.D * sd.H * sd.W;
    for (int j = 0; j < cfg->depths[s]; ++j) {
      WinGeom g = make_geom(sd.D, sd.H, sd.W, adaptive ? cfg->resized_window : cfg->window,
                            (j & 1) ? shift_full : shift_none);
      // rows of a window travel d-fastest between LN1 and proj when the third-generation attention kernel runs
      g.dfast = window_attn_pitch(g, cfg->window, 0) == ATT3_PITCH ? 1 : 0;
      const float* n1g = WF(); const float* n1b = WF();
      const __half* qkv_w = WH(); const float* qkv_b = WF(); const float* tab = WF();
      const __half* proj_w = WH(); const float* proj_b = WF();
      const float* n2g = WF(); const float* n2b = WF();
      const __half* fc1_w = WH(); const float* fc1_b = WF();
      const __half* fc2_w = WH(); const float* fc2_b = WF();

      // forward_part1 (:407-488).  C = 96 with full windows: LN1 + partition + qkv in one kernel (KVQ_LNQKV_SPLIT=1 keeps
      // the two-kernel form).  The kernel also exists for C = 192 / 384 with a TMA-streamed weight (KVQ_LNQKV_MAX_C=384),
      // correct but measured slower than ln_window + the 128 x 192-tile GEMM there (r02: 0.228 vs 0.141 ms at stage 1,
      // 0.608 vs 0.288 ms at stage 2: N = 96 MMAs and 12 KB weight slices), so it is used for the resident-weight case only
      static const bool lnqkv_split = []() { const char* e = getenv("KVQ_LNQKV_SPLIT"); return e && atoi(e) != 0; }();
      static const int lnqkv_max_c = []() { const char* e = getenv("KVQ_LNQKV_MAX_C"); return e ? atoi(e) : 96; }();
      const bool fused_qkv = ln_qkv_supported(C) && C <= lnqkv_max_c && g.dfast && g.N == 392 && sd.heads * 32 == C && !lnqkv_split;
      if (fused_qkv) {
        ProfScope ps(PK_QKV_GEMM, s, st);
        rc = launch_ln_qkv(xcur, n1g, n1b, eps, qkv_w, qkv_b, img, B, C, sd.heads, attn_qscale(), g, st);
      } else {
        ProfScope ps(PK_LN_WINDOW, s, st);
        rc = launch_ln_window(xcur, a16, n1g, n1b, eps, B, C, g, st);
      }
      if (rc != 0) return rc;
      rc = run_attention(a16, qkv_w, qkv_b, tab, a16, img, B, C, sd.heads, g, cfg->window, 0, st, s, adaptive ? 1 : 0,
                         fused_qkv);
      if (rc != 0) return rc;
      // proj + window_reverse + shortcut (:322-324, :472-488, :509) and norm2 (:490).  C = 96: one kernel that also writes
      // the fp16 norm2 output (into the free PatchMerging buffer: a16 is still being read as the proj operand);
      // KVQ_PROJLN_SPLIT=1 keeps proj GEMM + ln_rows.  The kernel exists for C = 192 too (KVQ_PROJLN_MAX_C=192) but measured
      // slower there than the two kernels (0.225 vs 0.142 ms per step at stage 1; stage 0: 0.230 vs 0.283 ms).
      static const bool projln_split = []() { const char* e = getenv("KVQ_PROJLN_SPLIT"); return e && atoi(e) != 0; }();
      const __half* mlp_in = a16;
      static const int projln_max_c = []() { const char* e = getenv("KVQ_PROJLN_MAX_C"); return e ? atoi(e) : 96; }();
      if (proj_ln_supported(C) && C <= projln_max_c && !projln_split) {
        __half* a2 = reinterpret_cast<__half*>(xnext);
        ProfScope ps(PK_PROJ_GEMM, s, st);
        rc = launch_proj_ln(a16, proj_w, proj_b, n2g, n2b, eps, xcur, a2, B, C, g, st);
        if (rc != 0) return rc;
        mlp_in = a2;
      } else {
        {
          GemmParams gp{};
          gp.M = B * g.nW * g.N; gp.N = C; gp.K = C;
          gp.bias = proj_b; gp.out = xcur; gp.ldo = C; gp.resid = xcur; gp.remap = 1; gp.geom = g;
          ProfScope ps(PK_PROJ_GEMM, s, st);
          rc = launch_gemm(EPI_RESID_F32, a16, C, proj_w, C, gp, st);
          if (rc != 0) return rc;
        }
        // forward_part2 (:490-491)
        {
          ProfScope ps(PK_LN_ROWS, s, st);
          rc = launch_ln_rows(xcur, a16, nullptr, n2g, n2b, eps, M, C, sd.D * sd.H * sd.W, st);
        }
        if (rc != 0) return rc;
      }
      static const bool fuse_mlp = []() { const char* e = getenv("KVQ_FUSED_MLP"); return e == nullptr || atoi(e) != 0; }();
      if (fuse_mlp && fused_mlp_supported(C)) {
        ProfScope ps(PK_FUSED_MLP, s, st);
        rc = launch_fused_mlp(mlp_in, fc1_w, fc1_b, fc2_w, fc2_b, xcur, M, C, st);
        if (rc != 0) return rc;
      } else {
        {
          GemmParams gp{};
          gp.M = M; gp.N = 4 * C; gp.K = C;
          gp.bias = fc1_b; gp.out = hid; gp.ldo = 4 * C;
          ProfScope ps(PK_FC1_GEMM, s, st);
          rc = launch_gemm(EPI_GELU_F16, mlp_in, C, fc1_w, C, gp, st);
          if (rc != 0) return rc;
        }
        {
          GemmParams gp{};
          gp.M = M; gp.N = C; gp.K = 4 * C;
          gp.bias = fc2_b; gp.out = xcur; gp.ldo = C; gp.resid = xcur;
          ProfScope ps(PK_FC2_GEMM, s, st);
          rc = launch_gemm(EPI_RESID_F32, hid, 4 * C, fc2_w, 4 * C, gp, st);
          if (rc != 0) return rc;
        }
      }
    }
    if (s + 1 < cfg->num_stages) {  // PatchMerging (:533-555)
      const float* ng = WF(); const float* nb = WF(); const __half* red_w = WH();
      {
        ProfScope ps(PK_MERGE_LN, s, st);
        rc = launch_ln_merge(xcur, a16, ng, nb, eps, B, sd.D, sd.H, sd.W, C, st);
      }
      if (rc != 0) return rc;
      ProfScope ps(PK_MERGE_GEMM, s, st);
      GemmParams gp{};
      gp.M = B * sd.D * ((sd.H + 1) / 2) * ((sd.W + 1) / 2); gp.N = 2 * C; gp.K = 4 * C;
      gp.out = xnext; gp.ldo = 2 * C;
      gp.split_b = (cfg->split_weights >> 1) & 1;
      rc = launch_gemm(EPI_RESID_F32, a16, 4 * C, red_w, gp.split_b ? 2 * ((4 * C + 63) / 64 * 64) : 4 * C, gp, st);
      if (rc != 0) return rc;
      std::swap(xcur, xnext);
      // after the first merge the big buffer is free: keep ping-ponging between the two (xa always fits)
    }
    if (hook != nullptr) {
      // BasicLayer output (blocks + downsample), channels-last fp32 tokens: KSVQE modulates it in place after stages
      // >= tuning_stage (KSVQE_model.py:1436-1482); the hook enqueues its work on the same stream
      const bool merged = s + 1 < cfg->num_stages;
      const int h2 = merged ? (sd.H + 1) / 2 : sd.H, w2 = merged ? (sd.W + 1) / 2 : sd.W;
      rc = hook(hook_arg, s, xcur, B * sd.D * h2 * w2, merged ? 2 * C : C, stream);
      KVQ_REQUIRE(rc == 0, rc < 0 ? rc : KVQ_ERR_BAD_SHAPE, "stage hook failed after stage %d (code %d)", s, rc);
    }
  }

  // ---- final norm (:1066-1080) [+ VQAHead (head.py:60-68)] ----
  {
    const StageDims& sd = pl.st[cfg->num_stages - 1];
    const int tokens = sd.D * sd.H * sd.W, M = B * tokens;
    const float* ng = WF(); const float* nb = WF();
    {
      ProfScope ps(PK_FINAL_LN, cfg->num_stages - 1, st);
      rc = launch_ln_rows(xcur, cfg->head_hidden > 0 ? a16 : nullptr, feat_out, ng, nb, eps, M, sd.C, tokens, st);
    }
    if (rc != 0) return rc;
    if (cfg->head_hidden > 0 && score_out != nullptr) {
      const __half* w1 = WH(); const float* b1 = WF(); const float* w2 = WF(); const float* b2 = WF();
      GemmParams gp{};
      gp.M = M; gp.N = cfg->head_hidden; gp.K = sd.C;
      gp.bias = b1; gp.w2 = w2; gp.b2ptr = b2; gp.rowscore = rowscore;
      gp.split_b = (cfg->split_weights >> 2) & 1;
      ProfScope ps(PK_HEAD, cfg->num_stages - 1, st);
      rc = launch_gemm(EPI_HEAD, a16, sd.C, w1, gp.split_b ? 2 * ((sd.C + 63) / 64 * 64) : sd.C, gp, st);
      if (rc != 0) return rc;
      rc = launch_row_mean(rowscore, score_out, B, tokens, st);
      if (rc != 0) return rc;
    }
  }
  return KVQ_OK;
}

int kvq_pack_split_f16(const float* in, void* out_f16, int rows, int K, void* stream) {
  KVQ_REQUIRE(in && out_f16 && rows > 0 && K > 0, KVQ_ERR_BAD_SHAPE, "pack_split_f16: bad arguments");
  return launch_pack_split(in, static_cast<__half*>(out_f16), rows, K, static_cast<cudaStream_t>(stream));
}

int kvq_cast_f16(const float* in, void* out_f16, size_t n, void* stream) {
  return launch_cast_f16(in, static_cast<__half*>(out_f16), n, static_cast<cudaStream_t>(stream));
}

int kvq_attn_table_len(int wd, int wh, int ww) { return attn_table_len(wd, wh, ww); }

int kvq_pack_bias_table(const float* rel, const float* frag, float* out, int wd, int wh, int ww, int heads,
                        void* stream) {
  KVQ_REQUIRE(rel != nullptr && out != nullptr && heads > 0, KVQ_ERR_BAD_SHAPE, "pack_bias_table: NULL argument");
  return launch_pack_bias(rel, frag, out, wd, wh, ww, heads, static_cast<cudaStream_t>(stream));
}

int kvq_linear_f16(const void* a_f16, const void* w_f16, const float* bias, void* out_f16, int M, int N, int K,
                   int gelu, void* stream) {
  GemmParams gp{};
  gp.M = M; gp.N = N; gp.K = K;
  gp.bias = bias; gp.out = out_f16; gp.ldo = N;
  return launch_gemm(gelu ? EPI_GELU_F16 : EPI_STORE_F16, static_cast<const __half*>(a_f16), K,
                     static_cast<const __half*>(w_f16), K, gp, static_cast<cudaStream_t>(stream));
}

int kvq_linear_resid_f32(const void* a_f16, const void* w_f16, const float* bias, const float* resid, float* out,
                         int M, int N, int K, void* stream) {
  GemmParams gp{};
  gp.M = M; gp.N = N; gp.K = K;
  gp.bias = bias; gp.out = out; gp.ldo = N; gp.resid = resid;
  return launch_gemm(EPI_RESID_F32, static_cast<const __half*>(a_f16), K, static_cast<const __half*>(w_f16), K, gp,
                     static_cast<cudaStream_t>(stream));
}

int kvq_conv_gemm_f16(const void* a_f16, int lda, const void* w_f16, const float* bias, const void* resid_f16, int ldr,
                      void* out_f16, int ldo, int M, int N, int K, int nvalid, int relu, void* stream) {
  GemmParams gp{};
  gp.M = M; gp.N = N; gp.K = K;
  gp.bias = bias; gp.out = out_f16; gp.ldo = ldo;
  gp.resid_h = static_cast<const __half*>(resid_f16); gp.ldr = ldr;
  gp.relu = relu; gp.nvalid = nvalid;
  return launch_gemm(EPI_CONV_F16, static_cast<const __half*>(a_f16), lda, static_cast<const __half*>(w_f16), K, gp,
                     static_cast<cudaStream_t>(stream));
}

int kvq_im2col_cl_f16(const void* in_f16, void* out_f16, int B, int T, int H, int W, int C, const int32_t kernel[3],
                      const int32_t stride[3], const int32_t pad[3], int Kp, void* stream) {
  KVQ_REQUIRE(in_f16 && out_f16 && kernel && stride && pad, KVQ_ERR_BAD_SHAPE, "im2col: NULL argument");
  return launch_im2col_cl(static_cast<const __half*>(in_f16), static_cast<__half*>(out_f16), B, T, H, W, C, kernel[0],
                          kernel[1], kernel[2], stride[0], stride[1], stride[2], pad[0], pad[1], pad[2], Kp,
                          static_cast<cudaStream_t>(stream));
}

int kvq_im2col_stem_f32(const float* in, void* out_f16, int N, int T, int H, int W, const int32_t kernel[3],
                        const int32_t stride[3], const int32_t pad[3], int Kp, void* stream) {
  KVQ_REQUIRE(in && out_f16 && kernel && stride && pad, KVQ_ERR_BAD_SHAPE, "im2col_stem: NULL argument");
  return launch_im2col_stem(in, static_cast<__half*>(out_f16), N, T, H, W, kernel[0], kernel[1], kernel[2], stride[0],
                            stride[1], stride[2], pad[0], pad[1], pad[2], Kp, static_cast<cudaStream_t>(stream));
}

int kvq_conv_implicit_f16(const void* in_f16, const void* w_f16, const float* bias, const void* resid_f16, int ldr,
                          void* out_f16, int ldo, int B, int T, int H, int W, int C, const int32_t kernel[3],
                          const int32_t stride[3], const int32_t pad[3], int N, int nvalid, int relu, void* stream) {
  KVQ_REQUIRE(in_f16 && w_f16 && out_f16 && kernel && stride && pad, KVQ_ERR_BAD_SHAPE, "conv_implicit: NULL argument");
  GemmParams gp{};
  gp.N = N;
  gp.bias = bias;
  gp.out = out_f16; gp.ldo = ldo;
  gp.resid_h = static_cast<const __half*>(resid_f16); gp.ldr = ldr;
  gp.relu = relu;
  gp.nvalid = nvalid;
  return launch_conv_implicit(static_cast<const __half*>(in_f16), B, T, H, W, C, kernel[0], kernel[1], kernel[2],
                              stride[0], stride[1], stride[2], pad[0], pad[1], pad[2],
                              static_cast<const __half*>(w_f16), gp, static_cast<cudaStream_t>(stream));
}

int kvq_conv_narrow_f16(const void* in_f16, const void* w_image, const float* bias, const void* resid_f16, int ldr,
                        void* out_f16, int ldo, int B, int T, int H, int W, int C, const int32_t kernel[3],
                        const int32_t stride[3], const int32_t pad[3], int nvalid, int relu, void* stream) {
  KVQ_REQUIRE(in_f16 && w_image && out_f16 && kernel && stride && pad, KVQ_ERR_BAD_SHAPE, "conv_narrow: NULL argument");
  GemmParams gp{};
  gp.N = 64;
  gp.bias = bias;
  gp.out = out_f16; gp.ldo = ldo;
  gp.resid_h = static_cast<const __half*>(resid_f16); gp.ldr = ldr;
  gp.relu = relu;
  gp.nvalid = nvalid == 64 ? 0 : nvalid;
  return launch_conv_narrow(static_cast<const __half*>(in_f16), B, T, H, W, C, kernel[0], kernel[1], kernel[2], stride[0],
                            stride[1], stride[2], pad[0], pad[1], pad[2], w_image, gp, static_cast<cudaStream_t>(stream));
}

int kvq_conv_image_kblocks(int C, int taps) {
  if (!(C == 8 || C == 16 || C == 32) || taps < 1) return KVQ_ERR_BAD_SHAPE;
  return conv_image_kblocks(C, taps);
}

int kvq_stem_conv_f16(const float* x, const void* w_packed_f16, const float* shift, void* out_f16, int N, int T, int H,
                      int W, int kt, int cout, void* stream) {
  return launch_stem_conv(x, static_cast<const __half*>(w_packed_f16), shift, static_cast<__half*>(out_f16), N, T, H, W,
                          kt, cout, static_cast<cudaStream_t>(stream));
}

int kvq_stem_weight_rows(int cout) { return stem_weight_rows(cout); }

int kvq_maxpool_hw_f16(const void* in_f16, void* out_f16, int N, int H, int W, int C, void* stream) {
  return launch_maxpool_hw(static_cast<const __half*>(in_f16), static_cast<__half*>(out_f16), N, H, W, C,
                           static_cast<cudaStream_t>(stream));
}

int kvq_pool_stats_f16(const void* in_f16, const float* weights, float* out_mean, float* out_std, int N, int L, int C,
                       int ldo, void* stream) {
  KVQ_REQUIRE(in_f16 && out_mean, KVQ_ERR_BAD_SHAPE, "pool_stats: NULL argument");
  return launch_pool_stats(static_cast<const __half*>(in_f16), weights, out_mean, out_std, N, L, C, ldo,
                           static_cast<cudaStream_t>(stream));
}

int kvq_rowdot_mean_f32(const float* x, const float* w, const float* b, float* score, int rows, int K, int group,
                        void* stream) {
  KVQ_REQUIRE(x && w && score, KVQ_ERR_BAD_SHAPE, "rowdot_mean: NULL argument");
  return launch_rowdot_mean(x, w, b, score, rows, K, group, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------------
// SimpleVQA spatial branch: torchvision-style ResNet-50 run per frame + multi-scale mean/std pooling
// (models/backbones/simpleVQA_model.py:220-264) + simpleVQAHead (models/head.py:10-31)
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct ResPlan {
  int Hs, Ws, Hp, Wp;         // after the stem conv / after the max pool
  size_t col_bytes, act_bytes;  // im2col scratch; one activation buffer (5 of them)
  size_t total;
};

inline int conv_out(int n, int k, int s, int p) { return (n + 2 * p - k) / s + 1; }

int make_res_plan(const KvqResNetConfig* cfg, int B, int T, int H, int W, ResPlan* pl) {
  KVQ_REQUIRE(cfg != nullptr, KVQ_ERR_BAD_SHAPE, "config is NULL");
  KVQ_REQUIRE(B >= 1 && T >= 1 && H >= 32 && W >= 32, KVQ_ERR_BAD_SHAPE, "resnet: input %dx3x%dx%dx%d too small", B, T,
              H, W);
  for (int s = 0; s < 4; ++s) KVQ_REQUIRE(cfg->layers[s] >= 1, KVQ_ERR_BAD_SHAPE, "layers[%d]=%d", s, cfg->layers[s]);
  KVQ_REQUIRE(cfg->feat3d_dim >= 0 && cfg->feat3d_dim % 4 == 0, KVQ_ERR_BAD_SHAPE, "feat3d_dim=%d (multiple of 4)",
              cfg->feat3d_dim);
  const size_t N = static_cast<size_t>(B) * T;
  pl->Hs = conv_out(H, 7, 2, 3); pl->Ws = conv_out(W, 7, 2, 3);
  pl->Hp = conv_out(pl->Hs, 3, 2, 1); pl->Wp = conv_out(pl->Ws, 3, 2, 1);
  size_t col = 0;
  size_t act = N * pl->Hs * pl->Ws * 64 * 2;
  int h = pl->Hp, w = pl->Wp;
  for (int s = 0; s < 4; ++s) {
    const int planes = 64 << s;
    const int stride = s == 0 ? 1 : 2;
    const int ho = conv_out(h, 3, stride, 1), wo = conv_out(w, 3, stride, 1);
    act = std::max(act, N * h * w * static_cast<size_t>(planes) * 2);        // conv1 output at the input resolution
    act = std::max(act, N * ho * wo * static_cast<size_t>(planes) * 4 * 2);  // block output
    col = std::max(col, N * ho * wo * static_cast<size_t>(planes) * 9 * 2);  // 3x3 im2col
    if (stride == 2) col = std::max(col, N * ho * wo * static_cast<size_t>(planes) * 2 * 2);  // strided 1x1 gather
    h = ho; w = wo;
  }
  KVQ_REQUIRE(N * pl->Hs * pl->Ws < (1ull << 31), KVQ_ERR_BAD_SHAPE, "resnet: %zu stem rows (int32 overflow)",
              N * pl->Hs * pl->Ws);
  const size_t slack = 256 * 4608 * 2;  // TMA tile overhang past the last row
  pl->col_bytes = align_up(col + slack, 256);
  pl->act_bytes = align_up(act + slack, 256);
  pl->total = pl->col_bytes + 5 * pl->act_bytes;
  return KVQ_OK;
}

int conv_gemm(const __half* A, int lda, const void* w, const void* b, const __half* resid, int ldr, __half* out, int M,
              int N, int K, bool relu, int stage, cudaStream_t st) {
  GemmParams gp{};
  gp.M = M; gp.N = N; gp.K = K;
  gp.bias = static_cast<const float*>(b);
  gp.out = out; gp.ldo = N;
  gp.resid_h = resid; gp.ldr = ldr;
  gp.relu = relu ? 1 : 0;
  ProfScope ps(PK_CONV_GEMM, stage, st);
  return launch_gemm(EPI_CONV_F16, A, lda, static_cast<const __half*>(w), K, gp, st);
}

// k x k spatial convolution (stride, pad) on [N,H,W,C] fp16: implicit GEMM when C % 64 == 0, else im2col + GEMM
int conv_spatial(const __half* in, int N, int H, int W, int C, int k, int stride, int pad, const void* w, const void* b,
                 __half* out, int cout, bool relu, __half* col, int stage, cudaStream_t st) {
  if (conv_implicit_supported(C, 1, k, k, 1, stride, stride)) {
    GemmParams gp{};
    gp.N = cout;
    gp.bias = static_cast<const float*>(b);
    gp.out = out; gp.ldo = cout;
    gp.relu = relu ? 1 : 0;
    ProfScope ps(PK_CONV_GEMM, stage, st);
    // the N frames are the T axis of one "clip": a 128-pixel tile may then span several frames of a small map (7x7 and
    // 14x14 maps would otherwise waste up to 2.6x of every tile), and a (1,k,k) kernel never mixes frames
    return launch_conv_implicit(in, 1, N, H, W, C, 1, k, k, 1, stride, stride, 0, pad, pad,
                                static_cast<const __half*>(w), gp, st);
  }
  const int Ho = conv_out(H, k, stride, pad), Wo = conv_out(W, k, stride, pad);
  const int Kp = (k * k * C + 63) / 64 * 64;   // packed weights carry the same zero columns
  int rc;
  {
    ProfScope ps(PK_CONV_IM2COL, stage, st);
    rc = launch_im2col_cl(in, col, N, 1, H, W, C, 1, k, k, 1, stride, stride, 0, pad, pad, Kp, st);
  }
  if (rc != 0) return rc;
  return conv_gemm(col, Kp, w, b, nullptr, 0, out, N * Ho * Wo, cout, Kp, relu, stage, st);
}

// layer1..layer4 of a torchvision-style Bottleneck ResNet (stride on the 3x3) on N channels-last maps, starting from
// act[1] = [N, h, w, 64] (after the stem max pool); `on_stage(s, cur, h*w, cout)` runs after each layer.
extern "C++" {
template <typename OnStage>
int resnet_layers(const int32_t layers[4], const void* const* weights, int& wi, __half* const act[5], __half* col, int N,
                  int h, int w, cudaStream_t st, OnStage on_stage, __half** final_act, int* final_h, int* final_w) {
  auto W_ = [&]() { return weights[wi++]; };
  __half* cur = act[1];
  int ci = 1;  // index of `cur`
  int cin = 64, rc;
  for (int s = 0; s < 4; ++s) {
    const int planes = 64 << s, cout = planes * 4;
    for (int j = 0; j < layers[s]; ++j) {
      const int stride = (j == 0 && s > 0) ? 2 : 1;
      const int ho = conv_out(h, 3, stride, 1), wo = conv_out(w, 3, stride, 1);
      const int M = N * h * w, Mo = N * ho * wo;
      __half* t1 = act[(ci + 1) % 5];
      __half* t2 = act[(ci + 2) % 5];
      __half* idn = act[(ci + 3) % 5];
      __half* out = act[(ci + 4) % 5];
      const void* w1 = W_(); const void* b1 = W_();
      const void* w2 = W_(); const void* b2 = W_();
      const void* w3 = W_(); const void* b3 = W_();
      // Bottleneck.forward (:106-126): 1x1 -> 3x3 (stride here) -> 1x1, + identity / downsample, ReLU
      rc = conv_gemm(cur, cin, w1, b1, nullptr, 0, t1, M, planes, cin, true, s, st);
      if (rc != 0) return rc;
      rc = conv_spatial(t1, N, h, w, planes, 3, stride, 1, w2, b2, t2, planes, true, col, s, st);
      if (rc != 0) return rc;
      const __half* resid = cur;
      if (j == 0) {
        const void* wd = W_(); const void* bd = W_();
        if (stride == 2) rc = conv_spatial(cur, N, h, w, cin, 1, 2, 0, wd, bd, idn, cout, false, col, s, st);
        else rc = conv_gemm(cur, cin, wd, bd, nullptr, 0, idn, Mo, cout, cin, false, s, st);
        if (rc != 0) return rc;
        resid = idn;
      }
      rc = conv_gemm(t2, planes, w3, b3, resid, cout, out, Mo, cout, planes, true, s, st);
      if (rc != 0) return rc;
      cur = out;
      ci = (ci + 4) % 5;
      cin = cout;
      h = ho; w = wo;
    }
    rc = on_stage(s, cur, h * w, cout);
    if (rc != 0) return rc;
  }
  if (final_act != nullptr) { *final_act = cur; *final_h = h; *final_w = w; }
  return KVQ_OK;
}
}  // extern "C++"

}  // namespace

int kvq_resnet_num_weights(const KvqResNetConfig* cfg) {
  if (cfg == nullptr) return KVQ_ERR_BAD_SHAPE;
  int n = 2;
  for (int s = 0; s < 4; ++s) n += 2 * (3 * cfg->layers[s] + 1);
  if (cfg->head) n += 2;
  return n;
}

int kvq_resnet_feature_dim(const KvqResNetConfig* cfg) {
  if (cfg == nullptr) return KVQ_ERR_BAD_SHAPE;
  return 2 * (512 + 1024 + 2048) + cfg->feat3d_dim;
}

size_t kvq_simplevqa_workspace_bytes(const KvqResNetConfig* cfg, int B, int T, int H, int W) {
  ResPlan pl;
  if (make_res_plan(cfg, B, T, H, W, &pl) != 0) return 0;
  return pl.total;
}

int kvq_simplevqa_forward(const KvqResNetConfig* cfg, const void* const* weights, int num_weights, const float* x,
                          const float* feat3d, int B, int T, int H, int W, float* feat_out, float* score_out,
                          void* workspace, size_t workspace_bytes, void* stream) {
  ResPlan pl;
  int rc = make_res_plan(cfg, B, T, H, W, &pl);
  if (rc != 0) return rc;
  KVQ_REQUIRE(num_weights == kvq_resnet_num_weights(cfg), KVQ_ERR_BAD_SHAPE, "resnet: %d weight pointers, expected %d",
              num_weights, kvq_resnet_num_weights(cfg));
  KVQ_REQUIRE(x && feat_out && workspace && weights, KVQ_ERR_BAD_SHAPE, "resnet: NULL argument");
  KVQ_REQUIRE(cfg->feat3d_dim == 0 || feat3d != nullptr, KVQ_ERR_BAD_SHAPE, "resnet: feat3d is NULL");
  KVQ_REQUIRE(!cfg->head || score_out != nullptr, KVQ_ERR_BAD_SHAPE, "resnet: score_out is NULL");
  KVQ_REQUIRE(workspace_bytes >= pl.total, KVQ_ERR_WORKSPACE, "resnet: workspace %zu < %zu bytes", workspace_bytes,
              pl.total);
  KVQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, KVQ_ERR_MISALIGNED, "workspace not 256 B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* wp = static_cast<uint8_t*>(workspace);
  __half* col = reinterpret_cast<__half*>(wp);
  __half* act[5];
  for (int i = 0; i < 5; ++i) act[i] = reinterpret_cast<__half*>(wp + pl.col_bytes + i * pl.act_bytes);
  const int N = B * T;
  const int fdim = kvq_resnet_feature_dim(cfg);
  int wi = 0;
  auto W_ = [&]() { return weights[wi++]; };

  // stem: conv 7x7/2 + BN + ReLU (:235-237), max pool 3x3/2 (:238); frames are the T axis of x (:225-231).
  // Implicit GEMM straight from the fp32 NCDHW input (kvq_stem.cu): no patch matrix
  {
    const void* w = W_(); const void* b = W_();
    ProfScope ps(PK_CONV_STEM, 0, st);
    rc = launch_stem_conv(x, static_cast<const __half*>(w), static_cast<const float*>(b), act[0], B, T, H, W, 1, 64, st);
  }
  if (rc != 0) return rc;
  {
    ProfScope ps(PK_CONV_POOL, 0, st);
    rc = launch_maxpool_hw(act[0], act[1], N, pl.Hs, pl.Ws, 64, st);
  }
  if (rc != 0) return rc;

  int col_off = 0;
  rc = resnet_layers(cfg->layers, weights, wi, act, col, N, pl.Hp, pl.Wp, st,
                     [&](int s, const __half* cur, int hw, int cout) -> int {
                       if (s < 1) return KVQ_OK;
                       // AdaptiveAvgPool2d(1) + global_std_pool2d of layer2/3/4 (:242-251), concatenated (:252)
                       ProfScope ps(PK_CONV_POOL, s, st);
                       const int r = launch_pool_stats(cur, nullptr, feat_out + col_off, feat_out + col_off + cout, N, hw,
                                                       cout, fdim, st);
                       col_off += 2 * cout;
                       return r;
                     },
                     nullptr, nullptr, nullptr);
  if (rc != 0) return rc;
  if (cfg->feat3d_dim > 0) {
    // torch.cat((x, x_3D_features), dim=1) (:256)
    KVQ_CUDA(cudaMemcpy2DAsync(feat_out + col_off, static_cast<size_t>(fdim) * 4, feat3d,
                               static_cast<size_t>(cfg->feat3d_dim) * 4, static_cast<size_t>(cfg->feat3d_dim) * 4, N,
                               cudaMemcpyDeviceToDevice, st));
  }
  if (cfg->head) {
    // simpleVQAHead (head.py:22-31): Linear(9472,128) -> Linear(128,1) has no activation in between, so it is the
    // single affine map w_eff = W2 W1, b_eff = W2 b1 + b2 (folded in fp32 at pack time); then the mean over frames
    const float* weff = static_cast<const float*>(W_());
    const float* beff = static_cast<const float*>(W_());
    ProfScope ps(PK_HEAD, 0, st);
    rc = launch_rowdot_mean(feat_out, weff, beff, score_out, N, fdim, T, st);
    if (rc != 0) return rc;
  }
  return KVQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// CONTRIQUE distortion encoder of KSVQE (KSVQE_model.py:1622-1665): 32x32 patches of every 2nd frame -> torchvision
// ResNet-50 trunk -> L2 normalise -> Linear + BN1d + ReLU -> Linear + BN1d
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct CqPlan {
  int N, gh, gw;
  size_t patch_bytes;
  ResPlan res;
  size_t total;
};
int make_cq_plan(int B, int T, int H, int W, int anchor, int step, CqPlan* pl) {
  KVQ_REQUIRE(B >= 1 && T >= 1 && step >= 1 && T % step == 0 && anchor >= 32 && H % anchor == 0 && W % anchor == 0,
              KVQ_ERR_BAD_SHAPE, "contrique: x %dx3x%dx%dx%d, patch %d, frame step %d", B, T, H, W, anchor, step);
  pl->gh = H / anchor; pl->gw = W / anchor;
  const long long n = static_cast<long long>(B) * (T / step) * pl->gh * pl->gw;
  KVQ_REQUIRE(n < (1 << 24), KVQ_ERR_BAD_SHAPE, "contrique: %lld patches", n);
  pl->N = static_cast<int>(n);
  KvqResNetConfig rc{{3, 4, 6, 3}, 0, 0};
  int r = make_res_plan(&rc, 1, pl->N, anchor, anchor, &pl->res);
  if (r != 0) return r;
  pl->patch_bytes = align_up(static_cast<size_t>(pl->N) * 3 * anchor * anchor * 4, 256);
  pl->total = pl->patch_bytes + pl->res.total;
  return KVQ_OK;
}
}  // namespace

int kvq_contrique_num_weights(void) { return 2 + 2 * (3 * 16 + 4) + 4; }

size_t kvq_contrique_workspace_bytes(int B, int T, int H, int W, int anchor, int frame_step) {
  CqPlan pl;
  if (make_cq_plan(B, T, H, W, anchor, frame_step, &pl) != 0) return 0;
  return pl.total;
}

int kvq_contrique_forward(const void* const* weights, int num_weights, const float* x, int B, int T, int H, int W,
                          int anchor, int frame_step, float* z_out, void* workspace, size_t workspace_bytes,
                          void* stream) {
  CqPlan pl;
  int rc = make_cq_plan(B, T, H, W, anchor, frame_step, &pl);
  if (rc != 0) return rc;
  KVQ_REQUIRE(num_weights == kvq_contrique_num_weights(), KVQ_ERR_BAD_SHAPE, "contrique: %d weight pointers, expected %d",
              num_weights, kvq_contrique_num_weights());
  KVQ_REQUIRE(x && z_out && workspace && weights, KVQ_ERR_BAD_SHAPE, "contrique: NULL argument");
  KVQ_REQUIRE(workspace_bytes >= pl.total, KVQ_ERR_WORKSPACE, "contrique: workspace %zu < %zu bytes", workspace_bytes,
              pl.total);
  KVQ_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, KVQ_ERR_MISALIGNED, "workspace not 256 B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* wp = static_cast<uint8_t*>(workspace);
  float* patches = reinterpret_cast<float*>(wp);
  wp += pl.patch_bytes;
  __half* col = reinterpret_cast<__half*>(wp);
  __half* act[5];
  for (int i = 0; i < 5; ++i) act[i] = reinterpret_cast<__half*>(wp + pl.res.col_bytes + i * pl.res.act_bytes);
  const int N = pl.N;
  int wi = 0;
  rc = launch_patch_split(x, patches, B, T, H, W, anchor, frame_step, st);
  if (rc != 0) return rc;
  {   // encoder[0..3]: conv1 + bn1 + relu + maxpool; the N patches are the T axis of one [1,3,N,a,a] "clip"
    const void* w = weights[wi++]; const void* b = weights[wi++];
    ProfScope ps(PK_CONV_STEM, 0, st);
    rc = launch_stem_conv(patches, static_cast<const __half*>(w), static_cast<const float*>(b), act[0], 1, N, anchor,
                          anchor, 1, 64, st);
  }
  if (rc != 0) return rc;
  rc = launch_maxpool_hw(act[0], act[1], N, pl.res.Hs, pl.res.Ws, 64, st);
  if (rc != 0) return rc;
  const int32_t layers[4] = {3, 4, 6, 3};
  __half* fin = nullptr;
  int fh = 0, fw = 0;
  rc = resnet_layers(layers, weights, wi, act, col, N, pl.res.Hp, pl.res.Wp, st,
                     [](int, const __half*, int, int) -> int { return KVQ_OK; }, &fin, &fh, &fw);
  if (rc != 0) return rc;
  KVQ_REQUIRE(fh == 1 && fw == 1, KVQ_ERR_BAD_SHAPE,
              "contrique: %dx%d patches leave a %dx%d map; h.view(-1, 2048) (:1655) needs 1x1 (32x32 patches)", anchor,
              anchor, fh, fw);
  // three free activation buffers: anything but `fin`
  __half* t[3];
  for (int i = 0, k = 0; i < 5 && k < 3; ++i)
    if (act[i] != fin) t[k++] = act[i];
  rc = launch_row_l2norm(fin, t[0], N, 2048, st);                                          // F.normalize (:1657)
  if (rc != 0) return rc;
  const void* w1 = weights[wi++]; const void* b1 = weights[wi++];
  const void* w2 = weights[wi++]; const void* b2 = weights[wi++];
  rc = conv_gemm(t[0], 2048, w1, b1, nullptr, 0, t[1], N, 2048, 2048, true, 3, st);         // Linear + BN1d + ReLU
  if (rc != 0) return rc;
  rc = conv_gemm(t[1], 2048, w2, b2, nullptr, 0, t[2], N, 128, 2048, false, 3, st);         // Linear + BN1d
  if (rc != 0) return rc;
  return launch_f16_to_f32(t[2], z_out, static_cast<size_t>(N) * 128, st);
}

int kvq_mlp_fused(const void* a_f16, const void* w1_f16, const float* b1, const void* w2_f16, const float* b2, float* x,
                  int M, int C, void* stream) {
  return launch_fused_mlp(static_cast<const __half*>(a_f16), static_cast<const __half*>(w1_f16), b1,
                          static_cast<const __half*>(w2_f16), b2, x, M, C, static_cast<cudaStream_t>(stream));
}

int64_t kvq_window_rows(int B, int D, int H, int W, const int32_t window[3], const int32_t shift[3]) {
  const WinGeom g = make_geom(D, H, W, window, shift);
  return static_cast<int64_t>(B) * g.nW * g.N;
}

int64_t kvq_window_row_map(int D, int H, int W, const int32_t window[3], const int32_t shift[3], int d_fastest,
                           int32_t* row_to_src, int32_t* src_to_row) {
  KVQ_REQUIRE(D > 0 && H > 0 && W > 0 && window && shift && row_to_src, KVQ_ERR_BAD_SHAPE, "window_row_map: bad arguments");
  WinGeom g = make_geom(D, H, W, window, shift);
  g.dfast = d_fastest ? 1 : 0;
  const int rows = g.nW * g.N;
  for (int r = 0; r < rows; ++r) row_to_src[r] = win_row_to_src(g, r);
  if (src_to_row != nullptr && g.Dp == g.D && g.Hp == g.H && g.Wp == g.W)
    for (int t = 0; t < g.tokens; ++t) src_to_row[t] = src_to_win_row(g, t);
  return rows;
}

int kvq_ln_window(const float* x, void* out_f16, const float* gamma, const float* beta, float eps, int B, int D,
                  int H, int W, int C, const int32_t window[3], const int32_t shift[3], void* stream) {
  const WinGeom g = make_geom(D, H, W, window, shift);
  return launch_ln_window(x, static_cast<__half*>(out_f16), gamma, beta, eps, B, C, g,
                          static_cast<cudaStream_t>(stream));
}

size_t kvq_window_attention_workspace_bytes(int B, int D, int H, int W, int C, const int32_t window[3],
                                            const int32_t shift[3]) {
  const WinGeom g = make_geom(D, H, W, window, shift);
  return static_cast<size_t>(B) * g.nW * (C / 32) * ATT2_UNIT_BYTES;
}

int kvq_window_attention(const void* xw_f16, const void* qkv_w_f16, const float* qkv_b, const float* packed_table,
                         void* out_f16, int B, int D, int H, int W, int C, int heads, const int32_t window[3],
                         const int32_t shift[3], void* workspace, size_t workspace_bytes, int debug_variant,
                         void* stream) {
  KVQ_REQUIRE(window[0] <= 8 && window[1] <= 7 && window[2] <= 7, KVQ_ERR_BAD_SHAPE, "window exceeds (8,7,7)");
  KVQ_REQUIRE(heads * 32 == C, KVQ_ERR_BAD_SHAPE, "heads=%d x 32 != C=%d", heads, C);
  const WinGeom g = make_geom(D, H, W, window, shift);
  const size_t need = static_cast<size_t>(B) * g.nW * heads * ATT2_UNIT_BYTES;
  KVQ_REQUIRE(workspace != nullptr && workspace_bytes >= need, KVQ_ERR_WORKSPACE, "workspace %zu B < required %zu B",
              workspace_bytes, need);
  return run_attention(static_cast<const __half*>(xw_f16), static_cast<const __half*>(qkv_w_f16), qkv_b,
                       packed_table, static_cast<__half*>(out_f16), static_cast<__half*>(workspace), B, C, heads, g,
                       window, debug_variant, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
