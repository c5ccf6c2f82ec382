// Internal (C++) interfaces between the host orchestration (kvq_swin.cu), the C-ABI shims (kvq_capi.cu)
// and the kernels.  Nothing here crosses the shared-library boundary; include/kvq_b200.h does.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kvq {

const char* last_error();
int num_sms();
void count_launch();
long long launch_count();
bool prof_enabled();
void prof_set(bool on);
void prof_mark(int cat, bool begin, cudaStream_t st);
int prof_collect(float* ms, int* launches, int ncat);

// timing categories: kind * 4 + stage
enum ProfKind : int {
  PK_EMBED_IM2COL = 0, PK_EMBED_GEMM, PK_LN_WINDOW, PK_QKV_GEMM, PK_ATTN, PK_PROJ_GEMM, PK_LN_ROWS, PK_FC1_GEMM,
  PK_FC2_GEMM, PK_MERGE_LN, PK_MERGE_GEMM, PK_FINAL_LN, PK_HEAD, PK_FUSED_MLP, PK_CONV_IM2COL, PK_CONV_GEMM, PK_CONV_POOL, PK_CONV_STEM,
  PK_COUNT
};
struct ProfScope {
  int cat;
  cudaStream_t st;
  ProfScope(int kind, int stage, cudaStream_t s) : cat(kind * 4 + stage), st(s) { prof_mark(cat, true, st); }
  ~ProfScope() { prof_mark(cat, false, st); }
};

// ----------------------------------------------------------------------------------------------
// window geometry of one Swin block (reference: swin_backbone.py get_window_size :145-158,
// window_partition :92-117, forward_part1 :407-488).  All token grids are [D,H,W] per clip.
// ----------------------------------------------------------------------------------------------
struct WinGeom {
  int D, H, W;        // token grid of the stage
  int Dp, Hp, Wp;     // padded up to window multiples
  int wd, wh, ww;     // clamped window
  int sd, sh, sw;     // clamped shift (0 on unshifted blocks)
  int nwd, nwh, nww;  // windows per dim
  int N;              // wd*wh*ww  tokens per window
  int SL;             // wh*ww     tokens per temporal slab of a window
  int nW;             // windows per clip
  int tokens;         // D*H*W
  int dfast;          // order of the N rows inside a window: 0 = (d,h,w) with w fastest (window_partition :92-117),
                      // 1 = (h,w,d) with d fastest -- the order the third-generation attention kernel keeps its keys
                      // in; used by the fused forward so LN output rows, Q rows, key slots and attention output rows
                      // all share one order (coalesced QKV-epilogue stores).  Only the order of intermediate rows
                      // changes: proj scatters back through the same map
  // reciprocals of the runtime divisors of the row maps (filled by make_geom): the maps run once per row inside
  // latency-bound producer / epilogue code, where a generic 32-bit division is a ~100-cycle dependent chain
  float r_N, r_wd, r_wh, r_ww, r_SL, r_nww, r_nhw, r_W, r_HW, r_rows, r_tokens;
};

// floor(n / d) with inv = 1.0f / d: the float estimate is off by at most one (corrected below) as long as the quotient
// stays below 2^21 and either n < 2^24 or d >= 256 (float(n) rounds n by up to 128 beyond 2^24).  Every call site
// divides rows / tokens (< 2^31) by rows or tokens per clip, or a window-sized index by a window dimension
// (tests/test_host_cpu.py emulates the float arithmetic on those ranges).
__host__ __device__ __forceinline__ int fdiv_i(int n, int d, float inv) {
  int q = static_cast<int>(static_cast<float>(n) * inv);
  const int r = n - q * d;
  q += (r >= d) ? 1 : 0;
  q -= (r < 0) ? 1 : 0;
  return q;
}

static inline WinGeom make_geom(int D, int H, int W, const int base_win[3], const int shift[3]) {
  WinGeom g;
  g.D = D; g.H = H; g.W = W;
  const int dims[3] = {D, H, W};
  int w[3], s[3];
  for (int i = 0; i < 3; ++i) {
    if (dims[i] <= base_win[i]) { w[i] = dims[i]; s[i] = 0; } else { w[i] = base_win[i]; s[i] = shift[i]; }
  }
  g.wd = w[0]; g.wh = w[1]; g.ww = w[2];
  g.sd = s[0]; g.sh = s[1]; g.sw = s[2];
  g.Dp = (D + g.wd - 1) / g.wd * g.wd;
  g.Hp = (H + g.wh - 1) / g.wh * g.wh;
  g.Wp = (W + g.ww - 1) / g.ww * g.ww;
  g.nwd = g.Dp / g.wd; g.nwh = g.Hp / g.wh; g.nww = g.Wp / g.ww;
  g.N = g.wd * g.wh * g.ww;
  g.SL = g.wh * g.ww;
  g.nW = g.nwd * g.nwh * g.nww;
  g.tokens = D * H * W;
  g.dfast = 0;
  // torch.roll semantics: modulo (an adaptive window keeps the base shift, which can exceed a tiny padded grid)
  g.sd %= g.Dp; g.sh %= g.Hp; g.sw %= g.Wp;
  g.r_N = 1.0f / g.N; g.r_wd = 1.0f / g.wd; g.r_wh = 1.0f / g.wh; g.r_ww = 1.0f / g.ww; g.r_SL = 1.0f / g.SL;
  g.r_nww = 1.0f / g.nww; g.r_nhw = 1.0f / (g.nwh * g.nww); g.r_W = 1.0f / W; g.r_HW = 1.0f / (H * W);
  g.r_rows = 1.0f / (g.nW * g.N); g.r_tokens = 1.0f / g.tokens;
  return g;
}

// Window-order row r in [0, nW*N) of one clip  ->  flat original token index, or -1 for a padded slot.
// shifted[p] = x[(p + s) mod size]  (torch.roll by -s, swin_backbone.py:430-435).
__host__ __device__ __forceinline__ int win_row_to_src(const WinGeom& g, int r) {
  const int win = fdiv_i(r, g.N, g.r_N), i = r - win * g.N;
  const int wdi = fdiv_i(win, g.nwh * g.nww, g.r_nhw);
  const int rem = win - wdi * (g.nwh * g.nww);
  const int whi = fdiv_i(rem, g.nww, g.r_nww), wwi = rem - whi * g.nww;
  int td, th, tw;
  if (g.dfast) {
    const int pos = fdiv_i(i, g.wd, g.r_wd);
    td = i - pos * g.wd;
    th = fdiv_i(pos, g.ww, g.r_ww);
    tw = pos - th * g.ww;
  } else {
    td = fdiv_i(i, g.SL, g.r_SL);
    const int pos = i - td * g.SL;
    th = fdiv_i(pos, g.ww, g.r_ww);
    tw = pos - th * g.ww;
  }
  // shifted[p] = x[(p + s) mod size]; make_geom keeps s < size, so one conditional subtraction is the modulo
  int od = wdi * g.wd + td + g.sd;
  int oh = whi * g.wh + th + g.sh;
  int ow = wwi * g.ww + tw + g.sw;
  od -= od >= g.Dp ? g.Dp : 0;
  oh -= oh >= g.Hp ? g.Hp : 0;
  ow -= ow >= g.Wp ? g.Wp : 0;
  if (od >= g.D || oh >= g.H || ow >= g.W) return -1;
  return (od * g.H + oh) * g.W + ow;
}

// Inverse of win_row_to_src for grids without padding (Dp == D, Hp == H, Wp == W: then the map is a bijection):
// flat original token index -> window-order row in [0, nW*N).
__host__ __device__ __forceinline__ int src_to_win_row(const WinGeom& g, int t) {
  const int od = fdiv_i(t, g.H * g.W, g.r_HW);
  const int rem = t - od * (g.H * g.W);
  const int oh = fdiv_i(rem, g.W, g.r_W), ow = rem - oh * g.W;
  int pd = od - g.sd, ph = oh - g.sh, pw = ow - g.sw;   // position in the rolled grid: shifted[p] = x[(p + s) mod size]
  pd += pd < 0 ? g.Dp : 0;
  ph += ph < 0 ? g.Hp : 0;
  pw += pw < 0 ? g.Wp : 0;
  const int wdi = fdiv_i(pd, g.wd, g.r_wd), td = pd - wdi * g.wd;
  const int whi = fdiv_i(ph, g.wh, g.r_wh), th = ph - whi * g.wh;
  const int wwi = fdiv_i(pw, g.ww, g.r_ww), tw = pw - wwi * g.ww;
  const int win = (wdi * g.nwh + whi) * g.nww + wwi;
  const int i = g.dfast ? (th * g.ww + tw) * g.wd + td : (td * g.wh + th) * g.ww + tw;
  return win * g.N + i;
}

// ----------------------------------------------------------------------------------------------
// attention operand images.  One (window, head) unit owns three 25 600 B images Q | K | V, each
// 400 rows x 32 halfs in the UMMA no-swizzle core-matrix order
//     byte(r, c) = (r/8)*512 + (c/8)*128 + (r%8)*16 + (c%8)*2
// Q rows are tokens in window order i; K/V rows are slab-padded key slots  col(i) = (i/SL)*50 + i%SL
// so that one temporal slab (<= 49 keys) starts every 50 columns of the logits tile.
// ----------------------------------------------------------------------------------------------
constexpr int ATT_ROWS = 400;
constexpr int ATT_SLAB = 50;
constexpr int ATT_HD = 32;
constexpr int ATT_IMG_BYTES = ATT_ROWS * ATT_HD * 2;  // 25 600
constexpr int ATT_UNIT_BYTES = 3 * ATT_IMG_BYTES;     // 76 800

// second-generation layout (kvq_attn2.cu): 52-slot key pitch so that four temporal slabs fill one 208-key MMA block;
// unit = Q (400 rows) | K (416 rows) | V (416 rows)
constexpr int ATT2_PITCH = 52;
constexpr int ATT2_KV_ROWS = 8 * ATT2_PITCH;               // 416
constexpr int ATT2_KV_BYTES = ATT2_KV_ROWS * ATT_HD * 2;   // 26 624
constexpr int ATT2_Q_BYTES = ATT_IMG_BYTES;                // 25 600
constexpr int ATT2_UNIT_BYTES = ATT2_Q_BYTES + 2 * ATT2_KV_BYTES;   // 78 848

// third-generation layout (kvq_attn3.cu): keys ordered (h, w, d) with d fastest, one 56-key chunk per window row h:
// slot = h*56 + w*8 + d, 392 keys + 8 zero rows; unit = Q (400 rows, natural order) | K (400 rows) | V (400 rows).
// Its bias table is the paired layout float4 {t0[e], t0[e-1], t1[e], t1[e-1]} at (e+6)*SD + (dh+6)*SH + (dw+6).
constexpr int ATT3_PITCH = 56;
constexpr int ATT3_KV_ROWS = 400;
constexpr int ATT3_KV_BYTES = ATT3_KV_ROWS * ATT_HD * 2;   // 25 600
constexpr int ATT3_UNIT_BYTES = ATT_IMG_BYTES + 2 * ATT3_KV_BYTES;   // 76 800
// compact strides: a quarter-warp of the softmax threads holds the 8 temporal positions of ONE (h,w) cell (rows are
// d-fastest), whose entries are SD apart: any odd SD is bank-conflict-free for LDS.128 (16 B entries)
constexpr int ATT3_SH = 13, ATT3_SD = 169;
constexpr int ATT3_PAIR_LEN = 14 * ATT3_SD;                          // 2366 float4 per head

__host__ __device__ __forceinline__ int att_img_offset(int r, int chunk) {  // byte offset of a 16 B chunk
  return (r >> 3) * 512 + chunk * 128 + (r & 7) * 16;
}

// ----------------------------------------------------------------------------------------------
// GEMM  D[M,N] = A[M,K] * B[N,K]^T  (fp16 operands, fp32 accumulate in TMEM) + fused epilogues
// ----------------------------------------------------------------------------------------------
enum GemmEpilogue : int {
  EPI_GELU_F16 = 0,  // out_h[m, n] = gelu(acc + bias[n])                         (Mlp.fc1 + act, :64-89)
  EPI_RESID_F32 = 1, // out_f[row(m), n] = resid[row(m), n] + acc + bias[n]       (proj / fc2 / reduction)
  EPI_QKV_IMG = 2,   // q*scale | k | v scattered into the attention operand images (WindowAttention3D :253-261)
  EPI_LN_F32 = 3,    // out_f[m, :] = LayerNorm(acc + bias) * gamma + beta, N == BLOCK_N  (PatchEmbed3D :715-733)
  EPI_HEAD = 4,      // rowscore[m] = sum_n gelu(acc + bias[n]) * w2[n] + b2      (VQAHead, head.py:60-68)
  EPI_STORE_F16 = 5, // out_h[m, n] = acc + bias[n]
  EPI_CONV_F16 = 6,  // out_h[m, n] = act(acc + bias[n] + resid_h[m, n]), n < nvalid   (conv + folded BN + ReLU)
};

struct GemmParams {
  int M, N, K;
  int split_b;          // B = [W_hi | W_lo] (fp16 pair summing to the fp32 weight), row stride 2*ceil64(K)
  const float* bias;    // [N] or nullptr
  void* out;            // fp16 / fp32 matrix, row stride ldo elements
  int ldo;
  const float* resid;   // EPI_RESID_F32: nullptr = no residual; may alias out
  const __half* resid_h; // EPI_CONV_F16: fp16 residual [M, ldr] or nullptr
  int ldr;
  int quick_gelu;       // EPI_GELU_F16: x * sigmoid(1.702 x) (CLIP's QuickGELU) instead of the exact erf GELU
  int relu;             // EPI_CONV_F16: apply ReLU
  int nvalid;           // EPI_CONV_F16: columns >= nvalid are padding (not stored); 0 = N
  // EPI_CONV_F16 implicit-GEMM mode (conv.on): A is the channels-last activation [B,T,H,W,C] behind a 5-D tensor map;
  // an M tile is a bt x bh x bw block of output pixels of one clip, K runs over (tap, 64-channel block)
  struct ConvImplicit {
    int on;
    int C, kt, kh, kw, st, sh, sw, pt, ph, pw;
    int To, Ho, Wo;        // output map
    int bt, bh, bw;        // tile (bt*bh*bw == 128)
    int nt, nh, nw;        // tiles per dim
    // narrow-channel mode (csub = C in {8, 16, 32}; 0 = 64-channel blocks): a K block is 64 / csub TAPS, each tap one
    // TMA sub-tile [128 pixels x csub] (no swizzle / 32 B / 64 B swizzle); the weights arrive as pre-packed smem
    // images, one 8 KB bulk copy per K block (N = 64 only)
    int csub, taps;
    const void* wimg;
  } conv;
  // EPI_RESID_F32 row remap (proj): rows are window-ordered; rows_in = nW*N per clip, rows_out = tokens per clip
  int remap;            // 0 = identity
  WinGeom geom;
  // EPI_QKV_IMG
  __half* img;          // unit-major operand images
  int att_pitch;        // key-slot pitch per temporal slab: ATT_SLAB (50) or ATT2_PITCH (52)
  int C;                // channels (N == 3C)
  int heads;
  float qscale;
  // EPI_LN_F32
  const float* gamma;
  const float* beta;
  float eps;
  // EPI_HEAD
  const float* w2;
  const float* b2ptr;   // device scalar
  float* rowscore;
};

int launch_gemm(int epi, const __half* A, int lda, const __half* B, int ldb, const GemmParams& p,
                cudaStream_t stream);
// implicit-GEMM convolution on a channels-last fp16 activation in [B,T,H,W,C] (C % 64 == 0): out[m, 0:nvalid] =
// act(conv(in) + bias (+ resid)); W fp16 [N, kt*kh*kw*C] tap-major / channel-minor.  p carries bias / out / ldo /
// resid_h / ldr / relu / nvalid / N; M, K and the tiling are filled in here.
bool conv_implicit_supported(int C, int kt, int kh, int kw, int st, int sh, int sw);
// narrow-channel implicit convolution (C = 8 / 16 / 32, at most 64 output channels): Wimg = kvq_pack_conv_image layout
bool conv_narrow_supported(int C, int cout);
int conv_image_kblocks(int C, int taps);   // 8 KB images per weight
int launch_conv_narrow(const __half* in, int B, int T, int H, int W, int C, int kt, int kh, int kw, int st, int sh,
                       int sw, int pt, int ph, int pw, const void* Wimg, GemmParams p, cudaStream_t stream);
int launch_conv_implicit(const __half* in, int B, int T, int H, int W, int C, int kt, int kh, int kw, int st, int sh,
                         int sw, int pt, int ph, int pw, const __half* Wt, GemmParams p, cudaStream_t stream);

// fused fc1 + GELU + fc2 + residual (hidden activation stays on chip); built for C = 96 / 192
bool fused_mlp_supported(int C);
int launch_fused_mlp(const __half* a, const __half* w1, const float* b1, const __half* w2, const float* b2, float* x,
                     int M, int C, cudaStream_t stream);

// ----------------------------------------------------------------------------------------------
// memory-bound row kernels
// ----------------------------------------------------------------------------------------------
// LN1 + cyclic shift + window partition: out[b*nW*N + r, :] = LN(x[b*tokens + src(r), :]) (zeros for padded slots)
int launch_ln_window(const float* x, __half* out, const float* gamma, const float* beta, float eps, int B, int C,
                     const WinGeom& g, cudaStream_t stream);
// plain LN over rows: out[m, :] = LN(x[m, :]);  optionally also writes the fp32 result transposed into
// feat[b, c, t] (channels-first, SwinTransformer3D.forward :1066-1080)
int launch_ln_rows(const float* x, __half* out, float* feat_cf, const float* gamma, const float* beta, float eps,
                   int M, int C, int tokens_per_clip, cudaStream_t stream);
// PatchMerging gather + LN(4C) (:533-555): out[(b,d,h2,w2), 4C]
int launch_ln_merge(const float* x, __half* out, const float* gamma, const float* beta, float eps, int B, int D,
                    int H, int W, int C, cudaStream_t stream);
// PatchEmbed3D im2col: x[B,3,T,H,W] fp32 (or fp16) -> A[B*D*Hs*Ws, 96] fp16, K index = c*32 + kt*16 + kh*4 + kw
int launch_patch_im2col(const void* x, int x_is_f16, __half* out, int B, int T, int H, int W, cudaStream_t stream);
// LN1 + window partition + qkv Linear + attention-image scatter in one kernel (kvq_lnqkv.cu), C = 96 / 192 / 384: x fp32
// [B*tokens, C] -> img (third-generation layout); needs g.dfast, the full (8,7,7) window and head_dim 32
bool ln_qkv_supported(int C);
int launch_ln_qkv(const float* x, const float* gamma, const float* beta, float eps, const __half* qkv_w,
                  const float* qkv_b, __half* img, int B, int C, int heads, float qscale, const WinGeom& g,
                  cudaStream_t stream);
// proj Linear + window_reverse + shortcut + norm2 in one kernel (kvq_projln.cu), C = 96 / 192: x += attn_out W^T + b scattered
// back to token order, a2 = LayerNorm(x) fp16 [tokens, C] (must not alias attn_out)
bool proj_ln_supported(int C);
int launch_proj_ln(const __half* attn_out, const __half* w, const float* bias, const float* gamma, const float* beta,
                   float eps, float* x, __half* a2, int B, int C, const WinGeom& g, cudaStream_t stream);
// PatchEmbed3D in one kernel (kvq_embed.cu): clip [B,3,T,H,W] fp32 / fp16 -> LayerNorm(conv + bias) fp32 [tokens, 96];
// w = patch_embed.proj.weight f16 [96, 96] (ldw 96) or its split pair [W_hi | W_lo] (ldw 256)
int launch_patch_embed(const void* x, int x_is_f16, const __half* w, int ldw, int split, const float* bias,
                       const float* gamma, const float* beta, float eps, float* out, int B, int T, int H, int W,
                       cudaStream_t stream);
// score[b] = mean_t rowscore[b*tokens + t]
int launch_row_mean(const float* rowscore, float* score, int B, int tokens, cudaStream_t stream);
// Grid mini-patch sampling + normalisation (datasets/fusion_datasets.py:22-121, :1017-1020)
int launch_fragment_gather_u8(const uint8_t* frames, const int32_t* offsets, float* out, int B, int T, int Hs, int Ws,
                              int fh, int fw, int fs, int aligned, const float mean[3], const float stdv[3],
                              cudaStream_t stream);
// KSVQE quality-aware region selection + gather (RegionNet_CLIP eval branch, patchnet.py:461-550)
int launch_qrs_select_gather(const float* fragment, const float* score, float* out, int32_t* region, int B, int T, int H,
                             int W, int n_key, int L, int anchor, int ks, cudaStream_t stream);
// CONTRIQUE pieces (KSVQE_model.py:1622-1665): patch split of every step-th frame, F.normalize over rows, fp16 -> fp32
int launch_patch_split(const float* x, float* out, int B, int T, int H, int W, int a, int step, cudaStream_t stream);
int launch_row_l2norm(const __half* in, __half* out, int rows, int C, cudaStream_t stream);
int launch_f16_to_f32(const __half* in, float* out, size_t n, cudaStream_t stream);
// fp32 [B, C, tokens] -> fp16 [B*tokens, C]
int launch_cf_to_rows(const float* in, __half* out, int B, int C, int tokens, cudaStream_t stream);
int launch_pack_split(const float* in, __half* out, int rows, int K, cudaStream_t stream);
// ---- convolution support (SimpleVQA ResNet-50 / SlowFast): channels-last fp16 activations ----
// in [B,T,H,W,C] -> A [B*To*Ho*Wo, Kp], K index = ((dt*kh + dh)*kw + dw)*C + c, zero padding, Kp >= kt*kh*kw*C
int launch_im2col_cl(const __half* in, __half* out, int B, int T, int H, int W, int C, int kt, int kh, int kw, int st,
                     int sh, int sw, int pt, int ph, int pw, int Kp, cudaStream_t stream, long long row0 = 0,
                     long long rows = -1);   // [row0, row0 + rows) of the output rows only (out = first row of the range)
// stem: in fp32 [N,3,T,H,W] (NCDHW) -> A [N*To*Ho*Wo, Kp], K index = ((dt*kh + dh)*kw + dw)*3 + c
int launch_im2col_stem(const float* in, __half* out, int N, int T, int H, int W, int kt, int kh, int kw, int st, int sh,
                       int sw, int pt, int ph, int pw, int Kp, cudaStream_t stream, long long row0 = 0,
                       long long rows = -1);
// implicit-GEMM stem: Conv3d(3 -> cout, (kt,7,7), stride (1,2,2), pad (kt/2,3,3)) + shift + ReLU on fp32 NCDHW input;
// w_packed fp16 [kt*3][stem_weight_rows(cout)][64] with k = dy*8 + dx per (frame tap, channel) block; out [N*T*Hs*Ws, cout]
int stem_weight_rows(int cout);
int launch_stem_conv(const float* x, const __half* w_packed, const float* shift, __half* out, int N, int T, int H, int W,
                     int kt, int cout, cudaStream_t stream);
// max pool (1,3,3) / stride (1,2,2) / pad (0,1,1) on [N,H,W,C] fp16
int launch_maxpool_hw(const __half* in, __half* out, int N, int H, int W, int C, cudaStream_t stream, int ldo = 0);
// weights of "AvgPool3d(kernel, stride 1) then global mean" over a [T,H,W] map: w[t,h,w] = (windows covering the
// position) / (number of windows * kernel volume); out fp32 [T*H*W]
int launch_pool_window_weights(float* out, int T, int H, int W, int kt, int kh, int kw, cudaStream_t stream);
// pack_pathway_output (SlowFast_features.py:112-135): slow[b,c,i] = frames[b,c,idx[i]], idx = linspace(0,T-1,T/4).long()
int launch_select_frames(const float* in, float* out, int BC, int T, long long plane, const int* idx, int n,
                         cudaStream_t stream);
// per (n, c): weighted mean over the HW (or THW) axis and optional unbiased std; in [N, L, C] fp16; weights fp32 [L]
// or nullptr (uniform); outputs fp32 with row stride ldo: mean at out_mean[n*ldo + c], std at out_std[n*ldo + c]
int launch_pool_stats(const __half* in, const float* weights, float* out_mean, float* out_std, int N, int L, int C,
                      int ldo, cudaStream_t stream);
// out[n] = dot(x[n, :K], w) + *b (b: device scalar or nullptr); then mean over groups of `group` consecutive rows -> score[n / group]
int launch_rowdot_mean(const float* x, const float* w, const float* b, float* score, int rows, int K, int group,
                       cudaStream_t stream);
// fp32 -> fp16 cast (weight packing)
int launch_cast_f16(const float* in, __half* out, size_t n, cudaStream_t stream);

// ----------------------------------------------------------------------------------------------
// fused 3-D window attention
// ----------------------------------------------------------------------------------------------
struct AttnParams {
  const __half* img;        // [units][Q|K|V] images
  __half* out;              // [B*nW*N, C] window-order rows
  const float* packed_tab;  // [heads][tab_len] float2 {t0, t1}: bias = t0 + fg*t1  (launch_pack_bias)
  int B, C, heads;
  int shifted;              // apply the SW-MSA region mask
  WinGeom geom;
  int base_wd, base_wh, base_ww;  // un-clamped window the bias tables are indexed with
  int variant;              // debug: 1 swaps LBO/SBO of the MN-major V descriptor
  int rows_dfast;           // third generation: geom.dfast rows in / out (fused forward); 0 = natural rows (stand-alone op)
  int rpi_geometric;        // adaptive_window_size: bias index from the token's own (d,h,w) in the resized window
                            // (relative_position_index.reshape(*base,*base)[:d,:h,:w,...], :264-271) instead of the
                            // flat [:N,:N] slice the reference takes for clamped windows
};
// entries per head of the packed table for a base window (L rounded up to even)
int attn_table_len(int bd, int bh, int bw);
// q pre-scale the attention kernels expect from the QKV epilogue: head_dim^-0.5 * log2(e)
float attn_qscale();
// rel/frag: [L, heads] fp32 (frag may be nullptr) -> out [heads][tab_len] float2
int launch_pack_bias(const float* rel, const float* frag, float* out, int bd, int bh, int bw, int heads,
                     cudaStream_t stream);
// debug builds (-DKVQ_TIMING) only: 1 + fills out16; production build returns 0
int debug_attn_timers(unsigned long long* out16, int reset);
int launch_window_attn(const AttnParams& p, cudaStream_t stream);
// which operand-image layout launch_window_attn will consume for this geometry: returns the key pitch (50 or 52)
int window_attn_pitch(const WinGeom& g, const int base_win[3], int variant);
int launch_window_attn2(const AttnParams& p, cudaStream_t stream);
int launch_window_attn3(const AttnParams& p, cudaStream_t stream);
// paired-table section of the packed bias buffer (third generation): out = float4 [heads][ATT3_PAIR_LEN]
int launch_pack_bias_pair(const float* rel, const float* frag, float* out, int heads, cudaStream_t stream);

}  // namespace kvq
