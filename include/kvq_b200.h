/*
 * kvq_b200.h -- C ABI of the B200 (sm_100a) hot path for the KVQ / KSVQE video-quality forward.
 *
 * The reference (lixinustc/KVQ-Challenge-CVPR-NTIRE2024) is pure PyTorch and has no FFI; the functions below are
 * the entry points a binding for the path  fragment frames -> SwinTransformer3D (GRPB) -> VQAHead -> score  calls
 * instead of the ATen/cuDNN ops behind the reference modules.  Each one cites the reference code it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (weights, activations, workspace); 16-byte aligned
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden synchronisation,
 *     no allocation, graph-capturable
 *   - return value 0 = ok, negative = error (see KVQ_ERR_*); kvq_last_error_string() describes the last failure
 *     on the calling thread
 *   - fp16 tensors are IEEE binary16; "packed" weights are produced once by the kvq_pack_* helpers
 */
#ifndef KVQ_B200_H_
#define KVQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KVQ_OK 0
#define KVQ_ERR_BAD_SHAPE (-1)
#define KVQ_ERR_MISALIGNED (-2)
#define KVQ_ERR_ARCH (-3)
#define KVQ_ERR_CUDA (-4)
#define KVQ_ERR_WORKSPACE (-5)
#define KVQ_ERR_DRIVER (-6)

#define KVQ_MAX_STAGES 4

/* Architecture of SwinTransformer3D (models/backbones/swin_backbone.py:760-842) + VQAHead (models/head.py:33-58). */
typedef struct KvqSwinConfig {
  int32_t embed_dim;                  /* 96 */
  int32_t num_stages;                 /* 4 */
  int32_t depths[KVQ_MAX_STAGES];     /* 2,2,6,2 */
  int32_t num_heads[KVQ_MAX_STAGES];  /* 3,6,12,24 (head_dim must be 32) */
  int32_t window[3];                  /* 8,7,7 */
  int32_t frag_bias[KVQ_MAX_STAGES];  /* 1,1,1,0: stage owns a fragment_position_bias_table (GRPB) */
  int32_t head_hidden;                /* 64; 0 = backbone only */
  float ln_eps;                       /* 1e-5 */
  int32_t split_weights;              /* bit 0 patch-embed, bit 1 PatchMerging reductions, bit 2 VQAHead fc_hid: the
                                         weight is passed as an fp16 pair [W_hi | W_lo] (kvq_pack_split_f16, row
                                         stride 2*ceil64(K)) so its rounding error drops from 2^-11 to 2^-22 */
  int32_t resized_window[3];          /* adaptive_window_size (swin_backbone.py:53-61, :1049-1055): every block
                                         partitions with this window (<= window per dim) instead of `window`, indexes
                                         the bias tables with the token's own (d,h,w) and keeps the base shift
                                         window/2.  0,0,0 = off */
} KvqSwinConfig;

/*
 * Weight pointer table for kvq_swin3d_forward, in this order (f16 = packed by kvq_cast_f16, f32 = as in the
 * reference state_dict):
 *   [0] patch_embed.proj.weight  f16 [C0, 96]   (flattened [C0,3,2,4,4])
 *   [1] patch_embed.proj.bias    f32 [C0]
 *   [2] patch_embed.norm.weight  f32 [C0]
 *   [3] patch_embed.norm.bias    f32 [C0]
 *   then for every stage s, for every block j of the stage, 13 entries:
 *     norm1.weight, norm1.bias (f32 [C]); attn.qkv.weight (f16 [3C, C]); attn.qkv.bias (f32 [3C]);
 *     packed bias table (f32 [heads][kvq_attn_table_len][2], from kvq_pack_bias_table);
 *     attn.proj.weight (f16 [C, C]); attn.proj.bias (f32 [C]); norm2.weight, norm2.bias (f32 [C]);
 *     mlp.fc1.weight (f16 [4C, C]); mlp.fc1.bias (f32 [4C]); mlp.fc2.weight (f16 [C, 4C]); mlp.fc2.bias (f32 [C])
 *   followed, for every stage but the last, by 3 entries:
 *     downsample.norm.weight, downsample.norm.bias (f32 [4C]); downsample.reduction.weight (f16 [2C, 4C])
 *   then  norm.weight, norm.bias (f32 [Cf])
 *   then (head_hidden > 0)  fc_hid.weight (f16 [hidden, Cf]); fc_hid.bias (f32 [hidden]);
 *                           fc_last.weight (f32 [hidden]); fc_last.bias (f32 [1])
 */
int kvq_swin3d_num_weights(const KvqSwinConfig* cfg);

/* Bytes of caller-owned scratch kvq_swin3d_forward needs for a [B,3,T,H,W] batch. */
size_t kvq_swin3d_workspace_bytes(const KvqSwinConfig* cfg, int B, int T, int H, int W);

/*
 * Whole hot path: replaces VQA_Network.forward for one Swin key (models/model.py:93-121) =
 * SwinTransformer3D.forward (swin_backbone.py:1044-1080) + VQAHead.forward (head.py:60-68).
 *   x         f32 [B,3,T,H,W]   (batch['technical'])
 *   feat_out  f32 [B,Cf,D,Hf,Wf] or NULL  (the backbone's return value)
 *   score_out f32 [B] or NULL             (head output, mean over D,H,W)
 */
int kvq_swin3d_forward(const KvqSwinConfig* cfg, const void* const* weights, int num_weights, const float* x, int B,
                       int T, int H, int W, float* feat_out, float* score_out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Same forward with a host callback after every stage (BasicLayer incl. its PatchMerging): `tokens` is the fp32
 * channels-last activation [rows, channels] (rows = B*D*h*w of the stage output) that the next stage reads, and may be
 * modified in place by work the hook enqueues on `stream` (KSVQE's cross-gating modulation after stages >= tuning_stage,
 * models/backbones/KSVQE_model.py:1436-1482).  The hook runs on the calling thread while the forward is being
 * enqueued (also under CUDA-graph capture); a non-zero return aborts the forward.  It is also called once with
 * stage = -1 on the patch-embedding output (feats[0] of SwinTransformer3D.forward :1058, for its `layer` / `multi`
 * outputs). */
typedef int (*kvq_stage_hook)(void* arg, int stage, float* tokens, int rows, int channels, void* stream);
int kvq_swin3d_forward_hooked(const KvqSwinConfig* cfg, const void* const* weights, int num_weights, const float* x,
                              int B, int T, int H, int W, float* feat_out, float* score_out, void* workspace,
                              size_t workspace_bytes, void* stream, kvq_stage_hook hook, void* hook_arg);

/* Same forward for clips already stored as IEEE fp16 [B,3,T,H,W] (half the host->device and read traffic).  The patch
 * embedding rounds its operand to fp16 in either case (swin_backbone.py:707-731 runs it in fp32; the stated 1e-3 score
 * tolerance covers the rounding), so a round-to-nearest conversion on the host gives bit-identical results. */
int kvq_swin3d_forward_x16(const KvqSwinConfig* cfg, const void* const* weights, int num_weights, const void* x_f16,
                           int B, int T, int H, int W, float* feat_out, float* score_out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- weight packing (once per load_state_dict) ---- */
int kvq_cast_f16(const float* in, void* out_f16, size_t n, void* stream);
/* fp32 [rows, K] -> fp16 [rows, 2*ceil64(K)] = [hi | 0 | lo | 0] with hi = fp16(w), lo = fp16(w - hi) */
int kvq_pack_split_f16(const float* in, void* out_f16, int rows, int K, void* stream);
/* entries per head of a packed bias table for base window (wd,wh,ww) */
int kvq_attn_table_len(int wd, int wh, int ww);
/* relative_position_bias_table / fragment_position_bias_table [L, heads] (frag may be NULL) -> packed table
 * (swin_backbone.py:202-243, :291-309) */
int kvq_pack_bias_table(const float* rel, const float* frag, float* out, int wd, int wh, int ww, int heads,
                        void* stream);

/* ---- single operators (unit-testable pieces of the path) ---- */
/* out = A[M,K] * W[N,K]^T + bias, optional exact-erf GELU; fp16 in/out, fp32 accumulate (nn.Linear, Mlp :64-89) */
int kvq_linear_f16(const void* a_f16, const void* w_f16, const float* bias, void* out_f16, int M, int N, int K,
                   int gelu, void* stream);
/* out_f32[M,N] = resid + A*W^T + bias (resid / bias may be NULL; resid may alias out) */
int kvq_linear_resid_f32(const void* a_f16, const void* w_f16, const float* bias, const float* resid, float* out,
                         int M, int N, int K, void* stream);
/* x[M,C] += fc2(gelu(fc1(a) + b1)) + b2 with the [M,4C] hidden kept on chip (Mlp.forward + residual,
 * swin_backbone.py:64-89, :490-491, :509).  a f16 [M,C]; w1 f16 [4C,C]; w2 f16 [C,4C]; built for C = 96, 192 */
int kvq_mlp_fused(const void* a_f16, const void* w1_f16, const float* b1, const void* w2_f16, const float* b2, float* x,
                  int M, int C, void* stream);
/* norm1 + cyclic shift + window_partition (swin_backbone.py:416-449): x f32 [B,D,H,W,C] -> f16 [B*nW*N, C] */
int kvq_ln_window(const float* x, void* out_f16, const float* gamma, const float* beta, float eps, int B, int D,
                  int H, int W, int C, const int32_t window[3], const int32_t shift[3], void* stream);
/* rows of the window-ordered matrix for a geometry: B*nW*N */
int64_t kvq_window_rows(int B, int D, int H, int W, const int32_t window[3], const int32_t shift[3]);
/* Host-only: the row maps the kernels use for one clip (roll by -shift + window_partition, swin_backbone.py:92-117,
 * :430-435), computed by the same inline functions the device code calls (float-reciprocal divisions included).
 * row_to_src [nW*N]: window-order row -> flat token index (d*H + h)*W + w, or -1 for a padded slot;
 * src_to_row [D*H*W] (may be NULL; filled only when the grid needs no padding): the inverse.
 * d_fastest = 0: rows of a window in (d,h,w) order (window_partition); 1: (h,w,d), the order of the fused forward.
 * Returns nW*N, or a negative error code. */
int64_t kvq_window_row_map(int D, int H, int W, const int32_t window[3], const int32_t shift[3], int d_fastest,
                           int32_t* row_to_src, int32_t* src_to_row);
/* scratch bytes for kvq_window_attention */
size_t kvq_window_attention_workspace_bytes(int B, int D, int H, int W, int C, const int32_t window[3],
                                            const int32_t shift[3]);
/* WindowAttention3D.forward up to (not including) proj (swin_backbone.py:245-322) on window-ordered rows:
 * xw f16 [B*nW*N, C] -> out f16 [B*nW*N, C] */
int kvq_window_attention(const void* xw_f16, const void* qkv_w_f16, const float* qkv_b, const float* packed_table,
                         void* out_f16, int B, int D, int H, int W, int C, int heads, const int32_t window[3],
                         const int32_t shift[3], void* workspace, size_t workspace_bytes, int debug_variant,
                         void* stream);
/* VQAHead.forward (models/head.py:60-68) on a channels-first feature map: feat f32 [B,C,tokens] -> score f32 [B]
 * (1x1x1 conv C->hidden, GELU(erf), hidden->1, mean over tokens).  w1 f16 [hidden,C]; b1,w2 f32 [hidden]; b2 f32 [1] */
size_t kvq_vqa_head_workspace_bytes(int B, int C, int tokens);
int kvq_vqa_head(const float* feat, const void* w1_f16, const float* b1, const float* w2, const float* b2,
                 float* score_out, int B, int C, int tokens, int hidden, void* workspace, size_t workspace_bytes,
                 void* stream);
/* Grid mini-patch sampling (datasets/fusion_datasets.py:22-121) fused with (v - mean)/std (:1017-1020):
 * frames u8 [B,T,3,Hs,Ws] -> out f32 [B,3,T,fh*fs,fw*fs]; offsets i32 [B,2,fh,fw,T/aligned] (h then w).
 * A source smaller than the fh*fs x fw*fs canvas takes the reference's "upsample" fallback (:43-50): bilinear
 * enlargement by 1 / min(Hs/(fh*fs), Ws/(fw*fs)) evaluated on the fly, cell grid and offsets on the ORIGINAL size */
int kvq_fragment_gather_u8(const uint8_t* frames, const int32_t* offsets, float* out, int B, int T, int Hs, int Ws,
                           int fragments_h, int fragments_w, int fsize, int aligned, const float mean[3],
                           const float std[3], void* stream);

/* First piece of the literal KSVQE key: quality-aware region selection, RegionNet_CLIP.forward eval
 * branch (models/backbones/patchnet.py:461-550) + obtain_keyframes grouping (KSVQE_model.py:1352-1376), without the
 * reference's per-frame .item() host syncs (graph-capturable).
 *   fragment f32 [B,3,T,H,W] (H = W = g*anchor), cls_attn f32 [B*n_key, L] (n_key = 4 key frames, L = l*l CLIP patch
 *   tokens) -> region_out i32 [B, n_key] (arg-max of the mean score over every region_patches^2 window of the
 *   nearest-resized g x g score map) and x_sel_out f32 [B,3,T,anchor*region_patches,anchor*region_patches] */
int kvq_qrs_select_gather(const float* fragment, const float* cls_attn, float* x_sel_out, int32_t* region_out, int B,
                          int T, int H, int W, int n_key, int L, int anchor, int region_patches, void* stream);

/* Second piece of the literal KSVQE key: the CONTRIQUE distortion encoder (CONTRIQUE_model.forward,
 * models/backbones/KSVQE_model.py:1622-1665; called on x_sel_ori[:, :, ::2], :1425).
 *   x f32 [B,3,T,H,W] -> every frame_step-th frame is cut into anchor x anchor patches (order b, t, gy, gx) -> torchvision
 *   ResNet-50 trunk (children()[:-2]) -> [N,2048] -> F.normalize -> Linear(2048,2048)+BN1d+ReLU -> Linear(2048,128)+BN1d
 *   z_out f32 [B, T/frame_step, (H/anchor)*(W/anchor), 128]
 * Weight table (BatchNorm folded as in kvq_simplevqa_forward): conv1 in the stem layout; per layer / block conv1,
 * conv2, conv3 (+ downsample for block 0); projector.0+projector.1 (fp16 [2048,2048], f32 [2048]);
 * projector.3+projector.4 (fp16 [128,2048], f32 [128]). */
int kvq_contrique_num_weights(void);
size_t kvq_contrique_workspace_bytes(int B, int T, int H, int W, int anchor, int frame_step);
int kvq_contrique_forward(const void* const* weights, int num_weights, const float* x, int B, int T, int H, int W,
                          int anchor, int frame_step, float* z_out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* Third piece of the literal KSVQE key: the CLIP ViT-B/16 visual tower with CLS adapters that KSVQE runs on its four
 * key frames (CLIP_extractor_addadapter_cls.forward, models/backbones/CLIP_backbone.py:156-202). */
typedef struct KvqClipConfig {
  int32_t width;         /* 768 */
  int32_t heads;         /* 12 (head_dim 64) */
  int32_t layers;        /* 12 */
  int32_t patch;         /* 16 */
  int32_t adapter_from;  /* CLIP_location (8): layers >= this blend the CLS token with its adapter output */
} KvqClipConfig;
/*
 * Weight table: conv1 (f16 [768, 768], K index (ky, kx, c)); class_embedding f32 [768]; positional_embedding f32
 * [1 + g*g, 768] ALREADY resized to the g x g patch grid (resize_pos_embed2d, bicubic, at load time); ln_pre weight,
 * bias; then per layer ln_1 weight, bias; attn.in_proj_weight (f16 [2304,768]), in_proj_bias (f32); attn.out_proj
 * weight (f16), bias; ln_2 weight, bias; mlp.c_fc weight (f16 [3072,768]), bias; mlp.c_proj weight (f16 [768,3072]),
 * bias; and for layers >= adapter_from the adapter's Linear(768,192) / Linear(192,768) weight, bias in f32.
 *   images f32 [n_img,3,H,W] (KSVQE: the 4 key frames of each clip, 112x112)
 *   cls_attn_out f32 [n_img, g*g] = cos(CLS, patch token);  tokens_out f32 [n_img, 1 + g*g, 768] (cls_token = row 0,
 *   pat_token = rows 1..)
 */
int kvq_clip_num_weights(const KvqClipConfig* cfg);
size_t kvq_clip_workspace_bytes(const KvqClipConfig* cfg, int n_img, int H, int W);
int kvq_clip_visual_forward(const KvqClipConfig* cfg, const void* const* weights, int num_weights, const float* images,
                            int n_img, int H, int W, float* cls_attn_out, float* tokens_out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Building blocks of KSVQE's cross-gating modulation (CDM, models/backbones/KSVQE_model.py:1436-1482); the Linears on
 * token rows are kvq_linear_f16 / kvq_conv_gemm_f16. */
/* softmax(q k^T * scale) v per (sequence, head), head_dim 64, at most 64 tokens (crossattention1 :1553-1587 with
 * scale = dim^-0.5; Attention :1508-1551 with scale = 64^-0.5).  Token t of sequence (o, i), o < n_outer, i < n_inner,
 * is row o*outer_stride + i*inner_stride + t*t_stride of q / k / v / out (f16, row strides ld* halfs). */
int kvq_mha_f16(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                int n_outer, int n_inner, long long outer_stride, long long inner_stride, long long t_stride, int Lq,
                int Lkv, int heads, float scale, void* stream);
/* x f32 [rows, C] in place: (a1 * (sigmoid(gd_pre[b]) * x + bd[b]) + a2 * (gs * x + bs)) / 2 with the per-token
 * gs = sigmoid(<es, wg> + bg), bs = <es, wb> + bb (Semantic_Transformation2 :817-835), the per-clip, per-channel
 * gd_pre / bd f32 [B, C] (Dist_Transformation3 :934-960 before its sigmoid) and the mix of :1482 */
int kvq_cdm_mix(float* x, const void* es_f16, const float* wg, const float* bg, const float* wb, const float* bb,
                const float* gd_pre, const float* bd, const float* a1, const float* a2, int rows, int C,
                int rows_per_clip, void* stream);
/* out f32 [M, N] = x f32 [M, K] * w f32 [N, K]^T + b (tiny M: per-clip statistics) */
int kvq_small_linear_f32(const float* x, const float* w, const float* b, float* out, int M, int N, int K, void* stream);
/* wa * a (f16) + wb * z (f32) -> f16 and / or f32 (dist_token = 0.2 * dist_adapter(z) + 0.8 * z, :1426) */
int kvq_blend_f16_f32(const void* a_f16, const float* z, float wa, float wb, void* out_f16, float* out_f32, size_t n,
                      void* stream);

/* ---- SimpleVQA spatial branch (config/kwai_simpleVQA_test.yml): ResNet-50 per frame + mean/std pools + head ---- */
typedef struct KvqResNetConfig {
  int32_t layers[4];   /* 3,4,6,3 Bottleneck blocks (models/backbones/simpleVQA_model.py:276 resnet50) */
  int32_t feat3d_dim;  /* 2304: width of batch['feat'], the pre-extracted SlowFast features (:226); 0 = none */
  int32_t head;        /* 1: simpleVQAHead (models/head.py:10-31) follows, score_out is written */
} KvqResNetConfig;

/*
 * Weight pointer table for kvq_simplevqa_forward.  Every convolution is passed with its BatchNorm folded in
 * (eval mode): w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps); weights are fp16
 * [Cout, Kp] with the K index tap-major / channel-minor ((dh*kw + dw)*Cin + c, zero padded to Kp % 64 == 0 so that
 * every 64-wide TMA box of the contraction is full), biases fp32 [Cout]:
 *   [0],[1]  conv1+bn1 (7x7/2) in the stem layout of kvq_stem_conv_f16
 *   then per layer, per block: conv1+bn1, conv2+bn2, conv3+bn3 and, for the first block of a layer,
 *   downsample.0+downsample.1  (w, b each)
 *   then (head) w_eff f32 [feature_dim] = quality.1.weight @ quality.0.weight,
 *               b_eff f32 [1] = quality.1.weight @ quality.0.bias + quality.1.bias
 */
int kvq_resnet_num_weights(const KvqResNetConfig* cfg);
/* 2*(512+1024+2048) + feat3d_dim = 9472 */
int kvq_resnet_feature_dim(const KvqResNetConfig* cfg);
size_t kvq_simplevqa_workspace_bytes(const KvqResNetConfig* cfg, int B, int T, int H, int W);
/*
 * Replaces VQA_Network.forward for the 'simpleVQA' key (models/model.py:93-121): ResNet.forward
 * (simpleVQA_model.py:220-264) + simpleVQAHead.forward (head.py:28-31).
 *   x         f32 [B,3,T,H,W]   (batch['simpleVQA'])
 *   feat3d    f32 [B,T,feat3d_dim] (batch['feat'])
 *   feat_out  f32 [B,T,feature_dim]  (the backbone's return value)
 *   score_out f32 [B]                (head output: mean over frames)
 */
int kvq_simplevqa_forward(const KvqResNetConfig* cfg, const void* const* weights, int num_weights, const float* x,
                          const float* feat3d, int B, int T, int H, int W, float* feat_out, float* score_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- SlowFast-R50 motion features (SlowFast_features.py; trunk = pytorchvideo.models.hub.slowfast_r50 blocks 0..4,
 * absent offline -> restated in oracle/slowfast.py, PARITY UNPINNED) ---- */
typedef struct KvqSlowFastConfig {
  int32_t depths[4];    /* 3,4,6,3 bottleneck blocks per stage, both pathways */
  int32_t alpha;        /* 4: fast frames per slow frame = temporal stride of conv_fast_to_slow (7,1,1) */
  int32_t slow_pool[3]; /* 8,7,7   blocks[5].pool[0] = AvgPool3d kernel, stride 1 (SlowFast_features.py:148) */
  int32_t fast_pool[3]; /* 32,7,7  blocks[5].pool[1] (:149) */
} KvqSlowFastConfig;

/*
 * Weight pointer table for kvq_slowfast_forward: every convolution with its eval-mode BatchNorm folded in, as an
 * (fp16 [round64(Cout), round64(K)] tap-major / channel-minor weight, fp32 [round64(Cout)] shift) pair, in this order:
 *   slow stem, fast stem (both in the stem layout of kvq_stem_conv_f16), block-0 fusion (conv_fast_to_slow + norm);
 *   then per stage s = 0..3: every slow res block (branch1 first for block 0; conv_a, conv_b, conv_c), every fast
 *   res block (same), then for s < 3 the stage's fusion.
 * Twins: every fast-pathway convolution that READS 8, 16 or 32 channels is followed in the table by a second pair:
 *   - 1x1x1, stride 1 (conv_c of stages 0..2, branch1 of stage 0): the row-folded pair (fp16 [g*Cout, 64],
 *     fp32 [g*Cout]), g = 64 / C: block-diagonal W'[j*Cout + n, j*C + k] = w'[n, k], shift repeated g times.  The
 *     library contracts g consecutive activation rows as one 64-wide row (same output bytes, full TMA boxes, g times
 *     fewer tiles) whenever the row count is a multiple of g;
 *   - anything else (conv_a (3,1,1), conv_b (1,3,3), the strided branch1 of stage 1, the fusions after the stem and
 *     after stage 0): (the smem image of kvq_conv_narrow_f16, fp32 [64] shift).
 */
int kvq_slowfast_num_weights(const KvqSlowFastConfig* cfg);
size_t kvq_slowfast_workspace_bytes(const KvqSlowFastConfig* cfg, int B, int Ts, int Tf, int H, int W);
/*
 * Replaces slowfast.forward (SlowFast_features.py:155-165):
 *   slow f32 [B,3,Ts,H,W], fast f32 [B,3,Tf,H,W]  (the two pathways pack_pathway_output returns)
 *   slow_out f32 [B,2048], fast_out f32 [B,256]   (= slow_feature / fast_feature [B,C,1,1,1])
 */
int kvq_slowfast_forward(const KvqSlowFastConfig* cfg, const void* const* weights, int num_weights, const float* slow,
                         const float* fast, int B, int Ts, int Tf, int H, int W, float* slow_out, float* fast_out,
                         void* workspace, size_t workspace_bytes, void* stream);
/* torch.linspace(0, T-1, T // alpha).long() (SlowFast_features.py:127-131), fp32 arithmetic as torch evaluates it;
 * HOST function: writes the indices to out[0..cap) and returns their number */
int kvq_slow_frame_indices(int T, int alpha, int32_t* out, int cap);
/* device side of pack_pathway_output (:112-135): slow_out[B,3,T//alpha,H,W] = index_select(frames[B,3,T,H,W], 2, idx) */
int kvq_pack_pathway_slow_f32(const float* frames, float* slow_out, int B, int T, int H, int W, int alpha,
                              void* stream);

/* ---- convolution building blocks (channels-last fp16 activations [B,T,H,W,C]) ---- */
/* out[M,ldo] = act(A[M,K] * W[N,K]^T + bias + resid): conv-as-GEMM with folded BN (Bottleneck.forward :106-126).
 * N (rows of W) is a multiple of 64; columns >= nvalid (0 = N) are padding and never stored */
int kvq_conv_gemm_f16(const void* a_f16, int lda, const void* w_f16, const float* bias, const void* resid_f16, int ldr,
                      void* out_f16, int ldo, int M, int N, int K, int nvalid, int relu, void* stream);
/* Implicit-GEMM convolution (nn.Conv2d / nn.Conv3d + folded BN + residual + ReLU; Bottleneck.forward :106-126, the
 * SlowFast res blocks) on a channels-last fp16 activation in [B,T,H,W,C], C % 64 == 0: no patch matrix -- the GEMM's
 * TMA producer walks (tap, 64-channel block) over a 5-D tensor map, stride = TMA traversal stride, padding = TMA
 * out-of-bounds zero fill.  w f16 [N, kt*kh*kw*C] tap-major / channel-minor, N % 64 == 0;
 * out[m, 0:nvalid] (row stride ldo), m = ((b*To + t)*Ho + h)*Wo + w */
int kvq_conv_implicit_f16(const void* in_f16, const void* w_f16, const float* bias, const void* resid_f16, int ldr,
                          void* out_f16, int ldo, int B, int T, int H, int W, int C, const int32_t kernel[3],
                          const int32_t stride[3], const int32_t pad[3], int N, int nvalid, int relu, void* stream);
/* Narrow-channel implicit GEMM (C = 8, 16 or 32 input channels, at most 64 output channels): a K block of the
 * contraction is 64 / C TAPS, each fetched by TMA as a [128 pixels x C] sub-tile (no swizzle / 32 B / 64 B swizzle).
 * w_image: kvq_conv_image_kblocks(C, taps) images of 8192 bytes, image kb = the [64 x 64] fp16 weight block of taps
 * [kb*64/C, (kb+1)*64/C) laid out exactly as the kernel's shared-memory B tile:
 *   byte(t, n, k) = t*128*C + n*2*C + ((k/8 ^ s(n)) * 16) + (k%8)*2,  s(n) = 0 (C=8), (n>>2)&1 (C=16), (n>>1)&3 (C=32)
 * with zero rows for n >= Cout and zero taps past the last one; bias fp32 [64] */
int kvq_conv_image_kblocks(int C, int taps);
int kvq_conv_narrow_f16(const void* in_f16, const void* w_image, const float* bias, const void* resid_f16, int ldr,
                        void* out_f16, int ldo, int B, int T, int H, int W, int C, const int32_t kernel[3],
                        const int32_t stride[3], const int32_t pad[3], int nvalid, int relu, void* stream);
/* gather [B,T,H,W,C] -> [B*To*Ho*Wo, Kp] patches, K index ((dt*kh + dh)*kw + dw)*C + c, zero padding (nn.Conv3d) */
int kvq_im2col_cl_f16(const void* in_f16, void* out_f16, int B, int T, int H, int W, int C, const int32_t kernel[3],
                      const int32_t stride[3], const int32_t pad[3], int Kp, void* stream);
/* same for the 3-channel fp32 NCDHW network input */
int kvq_im2col_stem_f32(const float* in, void* out_f16, int N, int T, int H, int W, const int32_t kernel[3],
                        const int32_t stride[3], const int32_t pad[3], int Kp, void* stream);
/* Implicit-GEMM stem: Conv3d(3 -> cout, (kt,7,7), stride (1,2,2), padding (kt/2,3,3)) + folded BN shift + ReLU
 * (simpleVQA_model.py:235-237 with kt = 1, cout = 64; SlowFast stems kt = 1 / cout = 64 and kt = 5 / cout = 8).
 *   x f32 [N,3,T,H,W]; out f16 [N*T*Hs*Ws, cout] channels-last
 *   w_packed f16 [kt*3][rows][64], rows = kvq_stem_weight_rows(cout): block (dt*3 + c), row n, column dy*8 + dx holds
 *   w[n][c][dt][dy][dx] * bn_scale[n] (zero for dy = 7, dx = 7, n >= cout); shift f32 [rows] */
int kvq_stem_weight_rows(int cout);
int kvq_stem_conv_f16(const float* x, const void* w_packed_f16, const float* shift, void* out_f16, int N, int T, int H,
                      int W, int kt, int cout, void* stream);
/* nn.MaxPool2d(3, 2, 1) (:153) on [N,H,W,C] */
int kvq_maxpool_hw_f16(const void* in_f16, void* out_f16, int N, int H, int W, int C, void* stream);
/* per (n, c) weighted mean over L (weights NULL = 1/L) and, if out_std != NULL, the unbiased std
 * (global_std_pool2d :18-20, nn.AdaptiveAvgPool2d :165); in f16 [N,L,C]; outputs f32 with row stride ldo */
int kvq_pool_stats_f16(const void* in_f16, const float* weights, float* out_mean, float* out_std, int N, int L, int C,
                       int ldo, void* stream);
/* score[g] = mean over the `group` rows of g of dot(x[row,:K], w) + *b */
int kvq_rowdot_mean_f32(const float* x, const float* w, const float* b, float* score, int rows, int K, int group,
                        void* stream);

/* ---- view pipeline in front of the models ----
 * torchvision.transforms.Resize((out_h, out_w)) on uint8 frames as the reference's get_resized_video
 * (datasets/fusion_datasets.py:229-252) and get_resizecrop_video (:299-316, test phase: centre crop) apply it, fused with
 * the datasets' normalisation lines (:1017-1027: (v - mean) / std for fragments, (v / 255 - clip_mean) / clip_std for the
 * CLIP view; :902-905 SimpleVQA).  Resize on uint8 = float32 anti-aliased bilinear interpolate (W axis, then H axis) ->
 * round half to even -> uint8; the kernels keep ATen's order of roundings, so out_u8 equals the reference byte for byte
 * and out_f32 float for float.
 *   frames u8: layout 0 = [B,T,3,Hs,Ws] (decoder order, as kvq_fragment_gather_u8), 1 = [B,3,T,Hs,Ws] (the argument
 *   order of the reference functions); outputs [B,3,T,crop_h,crop_w]: out_u8 (resized bytes) and / or out_f32 =
 *   ((v / divisor) - mean[c]) / std[c] (divisor 1 skips the first division), either may be NULL.
 *   crop window = rows [crop_y, crop_y+crop_h) x columns [crop_x, crop_x+crop_w) of the resized frame; crop_h = crop_w
 *   = 0 keeps the whole frame.  Only the source rows / output columns the window needs are computed.
 * kvq_resize_aa_taps / kvq_resize_aa_weights: the per-axis tap windows and weights (host, no GPU needed):
 *   xmin, xsize i32 [out_size], weights f32 [out_size, taps] */
int kvq_resize_aa_taps(int in_size, int out_size);
int kvq_resize_aa_weights(int in_size, int out_size, int32_t* xmin, int32_t* xsize, float* weights);
size_t kvq_resize_view_workspace_bytes(int B, int T, int Hs, int Ws, int out_h, int out_w, int crop_y, int crop_x,
                                       int crop_h, int crop_w);
int kvq_resize_view_u8(const uint8_t* frames, int layout, int B, int T, int Hs, int Ws, int out_h, int out_w, int crop_y,
                       int crop_x, int crop_h, int crop_w, float divisor, const float mean[3], const float std[3],
                       uint8_t* out_u8, float* out_f32, void* workspace, size_t workspace_bytes, void* stream);
/* The same view with the NON-antialiased bilinear filter: what torchvision.transforms.Resize computes on tensors before
 * 0.17 (antialias=None means off) -- the torch ~= 1.10 environment the reference's requirements.txt pins -- and with
 * antialias=False today.  Bit-exact to F.interpolate(bilinear, align_corners=False, antialias=False) + torch.round.
 * Two taps per axis, one fused kernel, no workspace. */
int kvq_resize_view_bilinear_u8(const uint8_t* frames, int layout, int B, int T, int Hs, int Ws, int out_h, int out_w,
                                int crop_y, int crop_x, int crop_h, int crop_w, float divisor, const float mean[3],
                                const float std[3], uint8_t* out_u8, float* out_f32, void* stream);

/* ---- measurement hooks (bench.py): kernels launched so far by this process, and optional CUDA-event timing of
 * every kernel of kvq_swin3d_forward grouped by (kind, stage).  Timing is OFF unless enabled. ---- */
long long kvq_launch_count(void);
void kvq_profile_enable(int on);
int kvq_profile_num_categories(void);
const char* kvq_profile_category_name(int category);
/* synchronises on the last recorded event; returns the number of timed launches (or a negative error) */
int kvq_profile_collect(float* ms_per_category, int* launches_per_category, int num_categories);

/* debug builds (-DKVQ_TIMING) only: per-phase cycle counters of the fast attention kernel; returns 0 otherwise */
int kvq_debug_attn_timers(unsigned long long* out16, int reset);

const char* kvq_last_error_string(void);
/* library / build identification, e.g. "kvq_b200 sm_100a" */
const char* kvq_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* KVQ_B200_H_ */
