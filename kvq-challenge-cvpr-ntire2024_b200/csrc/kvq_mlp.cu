// Fused Swin MLP for the HBM-bound stages (C = 96 / 192):   x += fc2(gelu(fc1(a)))   with a = LN2(x) in fp16
// (SwinTransformerBlock3D.forward_part2 + residual, swin_backbone.py:490-491, :509; Mlp :64-89).
//
// The [tokens, 4C] hidden activation never leaves the SM: per 128-token tile the hidden dimension is walked in
// chunks of 64 channels
//      H_j = A W1_j^T                 tcgen05.mma M128 x N64 x K=C     -> TMEM (double-buffered)
//      G_j = gelu(H_j + b1_j)         8 epilogue warps, packed fp32x2 erf GELU -> fp16, written into shared memory in
//                                     the 128B-swizzled K-major layout the next MMA reads (double-buffered)
//      O  += G_j W2[:, j]^T           tcgen05.mma M128 x N=C x K64     -> TMEM accumulator (double-buffered per tile)
// and the tile ends with  x = x + O + b2  through the same smem-transposed, residual-prefetching epilogue as the
// stand-alone GEMM.  Compared with fc1 / fc2 as two GEMMs this removes the hidden write + read
// (2 x tokens x 4C x 2 B: 616 MB per stage-0 block at batch 8).
//
//   warp 0 : TMA producer (A tile once per tile; W1_j / W2_j chunk tiles through 2-deep rings)
//   warp 1 : MMA issuer; GEMM1(j+1) is issued before GEMM2(j) so the tensor pipe runs under the GELU of chunk j
//   warps 2..9 : GELU epilogue per chunk, residual epilogue per tile
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int MLP_THREADS = 320;
constexpr int HC = 64;                 // hidden channels per chunk
constexpr int TILE_M = 128;

template <int C>
struct MlpCfg {
  static constexpr int KB_A = (C + 63) / 64;                 // 64-wide K blocks of the A tile / W1 chunk
  static constexpr int A_BYTES = KB_A * TILE_M * 128;        // 128 rows x 128 B per K block
  static constexpr int W1_BYTES = KB_A * HC * 128;           // [64 rows x C] as KB_A blocks of 64 rows x 128 B
  static constexpr int W2_BYTES = C * 128;                   // [C rows x 64 k]
  static constexpr int G_BYTES = TILE_M * 128;               // [128 rows x 64 k]
  static constexpr int NCHUNK = 4 * C / HC;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W1 = OFF_A + A_BYTES;
  static constexpr int OFF_W2 = OFF_W1 + 2 * W1_BYTES;
  static constexpr int OFF_G = OFF_W2 + 2 * W2_BYTES;
  static constexpr int OFF_STAGE = OFF_G + 2 * G_BYTES;      // per-warp fp32 transpose tiles (residual epilogue)
  static constexpr int STAGE_BYTES = 8 * (32 * 36 * 4);
  static constexpr int OFF_BIAS = OFF_STAGE + STAGE_BYTES;   // b1 [4C] + b2 [C] fp32
  static constexpr int OFF_BAR = OFF_BIAS + 5 * C * 4;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_H = 0;                           // 2 x 64 columns
  static constexpr int TMEM_O = 128;                         // 2 x C columns
  static constexpr int TMEM_COLS = (128 + 2 * C <= 256) ? 256 : 512;
  static_assert(C % 32 == 0 && 128 + 2 * C <= 512 && SMEM <= 227 * 1024, "unsupported channel count");
};

__device__ __forceinline__ float2 gelu2(float2 x) {
  const float2 z = fmul2(x, splat2(0.70710678118654752f));
  const float2 az = make_float2(fabsf(z.x), fabsf(z.y));
  const float2 den = ffma2(splat2(0.3275911f), az, splat2(1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  float2 p = ffma2(splat2(1.061405429f), t, splat2(-1.453152027f));
  p = ffma2(p, t, splat2(1.421413741f));
  p = ffma2(p, t, splat2(-0.284496736f));
  p = ffma2(p, t, splat2(0.254829592f));
  p = fmul2(p, t);
  const float2 q = fmul2(fmul2(az, az), splat2(-1.4426950408889634f));
  const float2 e = make_float2(fast_exp2(q.x), fast_exp2(q.y));
  const float2 er = ffma2(make_float2(-p.x, -p.y), e, splat2(1.0f));
  const float2 hx = fmul2(x, splat2(0.5f));
  return ffma2(hx, make_float2(copysignf(er.x, x.x), copysignf(er.y, x.y)), hx);
}

template <int C>
__global__ void __launch_bounds__(MLP_THREADS, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const float* __restrict__ b1, const float* __restrict__ b2,
                 float* __restrict__ x, int M) {
  using Cfg = MlpCfg<C>;
  constexpr int KB_A = Cfg::KB_A;
  constexpr int NCHUNK = Cfg::NCHUNK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* w1_full = bars + 2;    // [2]
  uint64_t* w1_empty = bars + 4;   // [2]
  uint64_t* w2_full = bars + 6;    // [2]
  uint64_t* w2_empty = bars + 8;   // [2]
  uint64_t* h_full = bars + 10;    // [2]  GEMM1(j) landed in TMEM
  uint64_t* h_empty = bars + 12;   // [2]  epilogue has read H
  uint64_t* g_full = bars + 14;    // [2]  G written to smem
  uint64_t* g_empty = bars + 16;   // [2]  GEMM2 has read G
  uint64_t* o_full = bars + 18;    // [2]
  uint64_t* o_empty = bars + 20;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (M + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w1_full[i], 1);
      mbar_init(&w1_empty[i], 1);
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 1);
      mbar_init(&h_full[i], 1);
      mbar_init(&h_empty[i], 8);
      mbar_init(&g_full[i], 8);
      mbar_init(&g_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 8);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 5 * C; i += MLP_THREADS) sbias[i] = i < 4 * C ? __ldg(b1 + i) : __ldg(b2 + i - 4 * C);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t n_a = 0, n_w = 0;   // completed uses of the A slot / of chunk slots (chunk counter across tiles)
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(a_empty, (n_a & 1) ^ 1);
        mbar_expect_tx(a_full, Cfg::A_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB_A; ++kb) tma_load_2d(smem + Cfg::OFF_A + kb * TILE_M * 128, &tmA, a_full, kb * 64, tile * TILE_M);
        ++n_a;
        for (int j = 0; j < NCHUNK; ++j, ++n_w) {
          const int s = n_w & 1;
          const uint32_t ph = ((n_w >> 1) & 1) ^ 1;
          mbar_wait(&w1_empty[s], ph);
          mbar_expect_tx(&w1_full[s], Cfg::W1_BYTES);
#pragma unroll
          for (int kb = 0; kb < KB_A; ++kb)
            tma_load_2d(smem + Cfg::OFF_W1 + s * Cfg::W1_BYTES + kb * HC * 128, &tmW1, &w1_full[s], kb * 64, j * HC);
          mbar_wait(&w2_empty[s], ph);
          mbar_expect_tx(&w2_full[s], Cfg::W2_BYTES);
          tma_load_2d(smem + Cfg::OFF_W2 + s * Cfg::W2_BYTES, &tmW2, &w2_full[s], j * HC, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_f16(TILE_M, HC, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_f16(TILE_M, C, 0, 0);
      const uint32_t sA = smem_u32(smem + Cfg::OFF_A);
      uint32_t n_a = 0, n_c = 0, n_t = 0;   // tiles (A uses), global chunk counter, tiles (O buffer uses)
      auto gemm1 = [&](uint32_t cidx) {     // H[cidx & 1] = A * W1_chunk^T
        const int s = cidx & 1;
        const uint32_t ph = (cidx >> 1) & 1;
        mbar_wait(&w1_full[s], ph);
        mbar_wait(&h_empty[s], ph ^ 1);
        tc_fence_after();
        const uint32_t sW = smem_u32(smem + Cfg::OFF_W1 + s * Cfg::W1_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB_A; ++kb) {
          const uint64_t da = umma_smem_desc(sA + kb * TILE_M * 128, 16, 1024, UMMA_SW_128);
          const uint64_t db = umma_smem_desc(sW + kb * HC * 128, 16, 1024, UMMA_SW_128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + Cfg::TMEM_H + s * HC, da + 2 * k, db + 2 * k, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&w1_empty[s]);
        umma_commit(&h_full[s]);
      };
      auto gemm2 = [&](uint32_t cidx, int j, int ob) {   // O[ob] (+)= G[cidx & 1] * W2_chunk^T
        const int s = cidx & 1;
        const uint32_t ph = (cidx >> 1) & 1;
        mbar_wait(&w2_full[s], ph);
        mbar_wait(&g_full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_smem_desc(smem_u32(smem + Cfg::OFF_G + s * Cfg::G_BYTES), 16, 1024, UMMA_SW_128);
        const uint64_t db = umma_smem_desc(smem_u32(smem + Cfg::OFF_W2 + s * Cfg::W2_BYTES), 16, 1024, UMMA_SW_128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_base + Cfg::TMEM_O + ob * C, da + 2 * k, db + 2 * k, idesc2, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&w2_empty[s]);
        umma_commit(&g_empty[s]);
      };
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
        const int ob = n_t & 1;
        mbar_wait(a_full, n_a & 1);
        mbar_wait(&o_empty[ob], ((n_t >> 1) & 1) ^ 1);
        tc_fence_after();
        gemm1(n_c);
        for (int j = 0; j < NCHUNK; ++j) {
          if (j + 1 < NCHUNK) gemm1(n_c + j + 1);        // keep the tensor pipe busy under the GELU of chunk j
          else umma_commit(a_empty);                      // all GEMM1s of this tile issued: A may be refilled
          gemm2(n_c + j, j, ob);
        }
        umma_commit(&o_full[ob]);
        n_c += NCHUNK;
        ++n_a;
      }
    }
  } else {
    // ============================ epilogue warps ============================
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* stile = reinterpret_cast<float*>(smem + Cfg::OFF_STAGE + ew * (32 * 36 * 4));
    const int r_loc = q * 32 + lane;                       // row inside the tile
    uint32_t n_c = 0, n_t = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
      // ---- per chunk: H -> gelu -> G (swizzled smem) ----
      for (int j = 0; j < NCHUNK; ++j, ++n_c) {
        const int s = n_c & 1;
        const uint32_t ph = (n_c >> 1) & 1;
        mbar_wait(&h_full[s], ph);
        __syncwarp();
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_x32(lane_addr + Cfg::TMEM_H + s * HC + half * 32, r);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_empty[s]);           // H buffer may be overwritten by GEMM1(j+2)
        const float* bj = sbias + j * HC + half * 32;
        uint32_t h[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 bb = *reinterpret_cast<const float4*>(bj + 4 * k);
          const float2 g01 = gelu2(fadd2(make_float2(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1])), make_float2(bb.x, bb.y)));
          const float2 g23 = gelu2(fadd2(make_float2(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3])), make_float2(bb.z, bb.w)));
          h[2 * k] = pack_half2(g01.x, g01.y);
          h[2 * k + 1] = pack_half2(g23.x, g23.y);
        }
        mbar_wait(&g_empty[s], ph ^ 1);                    // GEMM2(j-2) has consumed this G buffer
        // G tile = one 128B-swizzled K block: row r_loc, 16-byte chunk c lives at chunk (c ^ (r & 7))
        uint8_t* grow = smem + Cfg::OFF_G + s * Cfg::G_BYTES + r_loc * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(grow + (((half * 4 + k) ^ (r_loc & 7)) << 4)) = make_uint4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_full[s]);
      }
      // ---- per tile: x = x + O + b2 (transposed through smem, residual prefetched before the accumulator wait) ----
      const int ob = n_t & 1;
      constexpr int NCH = C / 32;                          // 32-column chunks of the output row
      constexpr int MYCH = (NCH + 1) / 2;
      const int sub = lane & 7, rsel = lane >> 3;          // 8 lanes per row on the global side, 4 rows per instruction
      const int row0 = tile * TILE_M + q * 32;
      float4 pre[MYCH][8];
#pragma unroll
      for (int ci = 0; ci < MYCH; ++ci) {
        const int c0 = (half + 2 * ci) * 32;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow_i = row0 + it * 4 + rsel;
          pre[ci][it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c0 < C && grow_i < M) pre[ci][it] = *reinterpret_cast<const float4*>(x + static_cast<size_t>(grow_i) * C + c0 + 4 * sub);
        }
      }
      mbar_wait(&o_full[ob], (n_t >> 1) & 1);
      __syncwarp();
      tc_fence_after();
#pragma unroll
      for (int ci = 0; ci < MYCH; ++ci) {
        const int c0 = (half + 2 * ci) * 32;
        if (c0 >= C) break;
        uint32_t r[32];
        tmem_ld_x32(lane_addr + Cfg::TMEM_O + ob * C + c0, r);
        const float4 bb = *reinterpret_cast<const float4*>(sbias + 4 * C + c0 + 4 * sub);
        tmem_wait_ld();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<uint4*>(stile + lane * 36 + 4 * k) = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rsel;
          const int grow_i = row0 + rr;
          if (grow_i < M) {
            float4 v = *reinterpret_cast<const float4*>(stile + rr * 36 + 4 * sub);
            v.x += bb.x + pre[ci][it].x; v.y += bb.y + pre[ci][it].y; v.z += bb.z + pre[ci][it].z; v.w += bb.w + pre[ci][it].w;
            *reinterpret_cast<float4*>(x + static_cast<size_t>(grow_i) * C + c0 + 4 * sub) = v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ob]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int C>
int launch_mlp_impl(const __half* a, const __half* w1, const float* b1, const __half* w2, const float* b2, float* x,
                    int M, cudaStream_t stream) {
  using Cfg = MlpCfg<C>;
  static bool attr = false;
  if (!attr) {
    KVQ_CUDA(cudaFuncSetAttribute(fused_mlp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  CUtensorMap tmA, tmW1, tmW2;
  int rc = make_tmap_2d(&tmA, a, M, C, static_cast<uint64_t>(C) * 2, TILE_M, 64, 2, 128);
  if (rc != 0) return rc;
  rc = make_tmap_2d(&tmW1, w1, 4 * C, C, static_cast<uint64_t>(C) * 2, HC, 64, 2, 128);
  if (rc != 0) return rc;
  rc = make_tmap_2d(&tmW2, w2, C, 4 * C, static_cast<uint64_t>(4 * C) * 2, C, 64, 2, 128);
  if (rc != 0) return rc;
  const int tiles = (M + TILE_M - 1) / TILE_M;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  count_launch();
  return launch_pdl(fused_mlp_kernel<C>, dim3(grid), dim3(MLP_THREADS), Cfg::SMEM, stream, tmA, tmW1, tmW2, b1, b2, x, M);
}

}  // namespace

bool fused_mlp_supported(int C) { return C == 96 || C == 192; }

int launch_fused_mlp(const __half* a, const __half* w1, const float* b1, const __half* w2, const float* b2, float* x,
                     int M, int C, cudaStream_t stream) {
  KVQ_REQUIRE(M > 0 && a && w1 && b1 && w2 && b2 && x, KVQ_ERR_BAD_SHAPE, "fused_mlp: bad arguments");
  if (C == 96) return launch_mlp_impl<96>(a, w1, b1, w2, b2, x, M, stream);
  if (C == 192) return launch_mlp_impl<192>(a, w1, b1, w2, b2, x, M, stream);
  set_error("fused_mlp: C=%d is not built (96, 192)", C);
  return KVQ_ERR_BAD_SHAPE;
}

}  // namespace kvq
