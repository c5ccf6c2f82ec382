// Fused Swin MLP for the HBM-bound stages (C = 96 / 192):   x += fc2(gelu(fc1(a)))   with a = LN2(x) in fp16
// (SwinTransformerBlock3D.forward_part2 + residual, swin_backbone.py:490-491, :509; Mlp :64-89).
//
// The [tokens, 4C] hidden activation never leaves the SM: per 128-token tile the hidden dimension is walked in
// chunks of 64 channels
//      H_j = A W1_j^T                 tcgen05.mma M128 x N64 x K=C     -> TMEM (double-buffered)
//      G_j = gelu(H_j + b1_j)         8 GELU warps, erf GELU (gelu_erf_fast) -> fp16.  C = 96: packed pairs written back
//                                     into TMEM (double-buffered), never into shared memory.  C = 192 (no TMEM columns
//                                     left next to a double-buffered O): shared memory, 128B-swizzled K-major.
//      O  += G_j W2[:, j]^T           tcgen05.mma, A = G_j from TMEM (C = 96) or smem, M128 x N=C x K64 -> TMEM accumulator
//                                     (double-buffered per tile)
// and the tile ends with  x = x + O + b2  through a smem-transposed, residual-prefetching epilogue.  Compared with
// fc1 / fc2 as two GEMMs this removes the hidden write + read (2 x tokens x 4C x 2 B: 616 MB per stage-0 block at
// batch 8).
//
//   warp 0       : TMA producer.  A tile per token tile (2 slots at C = 96); W1_j / W2_j slices through 3-4 / 2-3 deep
//                  rings, loaded in the order the issuers consume them (W1 runs W1_ST - 1 chunks ahead of W2).
//   warp 1       : GEMM1 issuer (lane 0).  H(c) is issued as soon as its W1 slice has landed and the GELU warps have
//                  READ H(c - 2) out of TMEM, so it lands while GELU(c - 1) runs.
//   warp 2       : GEMM2 issuer (lane 0).  O += G(c) W2_c^T as soon as G(c) is written.  Two issuing threads because
//                  each barrier wait costs 90-200 cycles even on a completed phase (tools/ubench/mbar_lat.cu) and one
//                  thread interleaving both streams was the serial resource of the kernel (r02 timing build).
//   warp 3       : idle (fills the warpgroup)
//   warps 4..11  : GELU per chunk (H -> G), nothing else; lane 0 polls the barriers for its warp
//   warps 12..15 : residual epilogue per tile (x += O + b2), concurrent with the GELU of the next tile
// r02 (batch 8, 32x224x224): 196 -> 132 us per stage-0 launch, 138 -> 101 us per stage-1 launch.
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int MLP_THREADS = 512;          // 16 warps: 0 TMA, 1-2 MMA, 3 idle, 4-11 GELU, 12-15 residual
constexpr int HC = 64;                 // hidden channels per chunk
constexpr int TILE_M = 128;

#ifdef MLP_TIMING   // debug build: cycle accounting of the issuer and one GELU warp, printed by one CTA
#define MLP_T_DECL long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long t_begin = clock64(); long long t_mark = t_begin
#define MLP_LAP(k) do { const long long t_now = clock64(); tacc[k] += t_now - t_mark; t_mark = t_now; } while (0)
#else
#define MLP_T_DECL
#define MLP_LAP(k)
#endif

template <int C>
struct MlpCfg {
  static constexpr int KB_A = (C + 63) / 64;                 // 64-wide K blocks of the A tile / W1 chunk
  static constexpr int A_BYTES = KB_A * TILE_M * 128;        // 128 rows x 128 B per K block
  static constexpr int W1_BYTES = KB_A * HC * 128;           // [64 rows x C] as KB_A blocks of 64 rows x 128 B
  static constexpr int W2_BYTES = C * 128;                   // [C rows x 64 k]
  static constexpr int NCHUNK = 4 * C / HC;
  static constexpr int KSTEPS = C / 16;                      // K = 16 steps of GEMM1 (the last K block may be partial)
  // ring depths.  GEMM1 runs two chunks ahead of GEMM2 (see the issuer), so W1 needs > 2 slots to hide the TMA latency
  // (r02 timing build: with 2 + 2 slots the issuer spent 19 % of its time on w1_full and 11 % on a_full).
  // G (the GELU output, A operand of GEMM2) lives in TMEM when the column budget allows a double-buffered O next to
  // it (C = 96); otherwise (C = 192) it goes through shared memory in the 128B-swizzled K-major layout.
  static constexpr bool G_TMEM = 192 + 2 * C <= 512;
  static constexpr int G_BYTES = G_TMEM ? 0 : TILE_M * 128; // [128 rows x 64 k]
  static constexpr int A_ST = C <= 96 ? 2 : 1;
  static constexpr int W1_ST = C <= 96 ? 4 : 3;
  static constexpr int W2_ST = C <= 96 ? 3 : 2;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W1 = OFF_A + A_ST * A_BYTES;
  static constexpr int OFF_W2 = OFF_W1 + W1_ST * W1_BYTES;
  static constexpr int OFF_G = OFF_W2 + W2_ST * W2_BYTES;
  static constexpr int OFF_STAGE = OFF_G + 2 * G_BYTES;      // per-warp fp32 transpose tiles (residual epilogue)
  static constexpr int STAGE_BYTES = 4 * (32 * 36 * 4);
  static constexpr int OFF_BIAS = OFF_STAGE + STAGE_BYTES;   // b1 [4C] + b2 [C] fp32
  static constexpr int OFF_BAR = OFF_BIAS + 5 * C * 4;
  static constexpr int SMEM = OFF_BAR + 512 + 1024;
  static constexpr int TMEM_H = 0;                           // 2 x 64 fp32 columns
  static constexpr int TMEM_G = 128;                         // G_TMEM: 2 x 32 columns of packed fp16 pairs
  static constexpr int TMEM_O = G_TMEM ? 192 : 128;          // 2 x C fp32 columns
  static constexpr int O_ST = 2;
  static constexpr int TMEM_COLS = (TMEM_O + 2 * C <= 256) ? 256 : 512;
  static_assert(C % 32 == 0 && TMEM_O + O_ST * C <= 512 && SMEM <= 227 * 1024, "unsupported channel count");
  static_assert(W1_ST >= 3 && W2_ST >= 2, "GEMM1 runs two chunks ahead");
  static_assert(!G_TMEM || TMEM_G + 64 <= TMEM_O, "TMEM map");
};

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
  uint64_t* a_full = bars + 0;     // [2]
  uint64_t* a_empty = bars + 2;    // [2]
  uint64_t* w1_full = bars + 4;    // [4]
  uint64_t* w1_empty = bars + 8;   // [4]
  uint64_t* w2_full = bars + 12;   // [3]
  uint64_t* w2_empty = bars + 15;  // [3]
  uint64_t* h_full = bars + 18;    // [2]  GEMM1(c) landed in TMEM
  uint64_t* h_empty = bars + 20;   // [2]  GELU warps have read H
  uint64_t* g_full = bars + 22;    // [2]  G written to smem
  uint64_t* g_empty = bars + 24;   // [2]  GEMM2 has read G
  uint64_t* o_full = bars + 26;    // [2]
  uint64_t* o_empty = bars + 28;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (M + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&h_full[i], 1);
      mbar_init(&h_empty[i], 8);
      mbar_init(&g_full[i], 8);
      mbar_init(&g_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&w1_full[i], 1);
      mbar_init(&w1_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 1);
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
  if (tmem_base != 0) {   // one CTA per SM (launch bounds + shared memory): the allocation starts at column 0, which the
    if (threadIdx.x == 0) printf("kvq fused_mlp: unexpected TMEM base %u\n", tmem_base);   // MMA issuer relies on
    __trap();
  }
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ============================ TMA producer ============================
    // Loads follow the order the issuer consumes them in: W1 (and the A tile in front of a tile's first chunk) runs
    // W1_ST - 1 chunks ahead of W2.
    if (lane == 0) {
      const uint32_t my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const uint32_t total = my_tiles * NCHUNK;
      uint32_t p1 = 0;                                   // next chunk whose W1 slice is to be loaded
      MLP_T_DECL;
      auto next_w1 = [&]() {
        if (p1 >= total) return;
        const uint32_t t1 = p1 / NCHUNK, j1 = p1 - t1 * NCHUNK;
        if (j1 == 0) {
          const uint32_t sa = t1 % Cfg::A_ST;
          MLP_LAP(7);
          mbar_wait(&a_empty[sa], ((t1 / Cfg::A_ST) & 1) ^ 1);
          MLP_LAP(0);
          mbar_expect_tx(&a_full[sa], Cfg::A_BYTES);
          const int tile = blockIdx.x + t1 * gridDim.x;
#pragma unroll
          for (int kb = 0; kb < KB_A; ++kb)
            tma_load_2d(smem + Cfg::OFF_A + sa * Cfg::A_BYTES + kb * TILE_M * 128, &tmA, &a_full[sa], kb * 64, tile * TILE_M);
        }
        const uint32_t s = p1 % Cfg::W1_ST;
        MLP_LAP(7);
        mbar_wait(&w1_empty[s], ((p1 / Cfg::W1_ST) & 1) ^ 1);
        MLP_LAP(1);
        mbar_expect_tx(&w1_full[s], Cfg::W1_BYTES);
#pragma unroll
        for (int kb = 0; kb < KB_A; ++kb)
          tma_load_2d(smem + Cfg::OFF_W1 + s * Cfg::W1_BYTES + kb * HC * 128, &tmW1, &w1_full[s], kb * 64, j1 * HC);
        ++p1;
      };
#pragma unroll
      for (int i = 0; i < Cfg::W1_ST - 1; ++i) next_w1();
      for (uint32_t c = 0; c < total; ++c) {
        next_w1();
        const uint32_t s = c % Cfg::W2_ST;
        MLP_LAP(7);
        mbar_wait(&w2_empty[s], ((c / Cfg::W2_ST) & 1) ^ 1);
        MLP_LAP(2);
        mbar_expect_tx(&w2_full[s], Cfg::W2_BYTES);
        tma_load_2d(smem + Cfg::OFF_W2 + s * Cfg::W2_BYTES, &tmW2, &w2_full[s], (c % NCHUNK) * HC, 0);
      }
#ifdef MLP_TIMING
      if (blockIdx.x == 5)
        printf("MLPT producer: total %lld | a_empty wait %lld w1_empty wait %lld w2_empty wait %lld | issue+other %lld\n",
               clock64() - t_begin, tacc[0], tacc[1], tacc[2], tacc[7]);
#endif
    }
  } else if (warp == 1 || warp == 2) {
    // ============================ MMA issuers (warp 1: GEMM1 stream, warp 2: GEMM2 stream) ============================
    // Lane 0 of each warp runs its stream alone.  (A warp-uniform loop with elect.sync around the MMAs and a
    // __syncwarp per step measured 20-30 % slower here, r02.)  Descriptors are a base built once plus an offset; the
    // TMEM base is 0 (one CTA per SM).
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_f16(TILE_M, HC, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_f16(TILE_M, C, 0, 0);
      const uint32_t sbase = smem_u32(smem);
      const uint32_t my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      MLP_T_DECL;
      if (warp == 1) {
        // GEMM1 stream: H(c) as soon as its W1 slice has landed and the GELU warps have read H(c - 2)
        const uint64_t dA0 = umma_smem_desc(sbase + Cfg::OFF_A, 16, 1024, UMMA_SW_128);
        const uint64_t dW10 = umma_smem_desc(sbase + Cfg::OFF_W1, 16, 1024, UMMA_SW_128);
        uint32_t c = 0;
        for (uint32_t t = 0; t < my_tiles; ++t) {
          const uint32_t sa = t % Cfg::A_ST;
          MLP_LAP(7);
          mbar_wait(&a_full[sa], (t / Cfg::A_ST) & 1);
          MLP_LAP(0);
          const uint64_t da = dA0 + static_cast<uint64_t>(sa * (Cfg::A_BYTES >> 4));
          for (uint32_t j = 0; j < NCHUNK; ++j, ++c) {
            const uint32_t s = c & 1, sw = c % Cfg::W1_ST;
            MLP_LAP(7);
            mbar_wait(&w1_full[sw], (c / Cfg::W1_ST) & 1);
            MLP_LAP(1);
            mbar_wait(&h_empty[s], ((c >> 1) & 1) ^ 1);
            MLP_LAP(2);
            tc_fence_after();
            const uint64_t db = dW10 + static_cast<uint64_t>(sw * (Cfg::W1_BYTES >> 4));
            const uint32_t tH = Cfg::TMEM_H + s * HC;
#pragma unroll
            for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
              const int kb = ks >> 2, k = ks & 3;
              umma_f16_ss(tH, da + static_cast<uint64_t>(((kb * TILE_M * 128) >> 4) + 2 * k),
                          db + static_cast<uint64_t>(((kb * HC * 128) >> 4) + 2 * k), idesc1, ks > 0 ? 1u : 0u);
            }
            umma_commit(&w1_empty[sw]);
            umma_commit(&h_full[s]);
            if (j == NCHUNK - 1) umma_commit(&a_empty[sa]);   // all GEMM1s of the tile issued: its A slot may be refilled
            MLP_LAP(3);
          }
        }
#ifdef MLP_TIMING
        if (blockIdx.x == 5)
          printf("MLPT GEMM1 issuer: total %lld | a_full wait %lld w1_full wait %lld h_empty wait %lld issue %lld | other %lld (chunks %u)\n",
                 clock64() - t_begin, tacc[0], tacc[1], tacc[2], tacc[3], tacc[7], c);
#endif
      } else {
        // GEMM2 stream: O += G(c) W2_c^T once the GELU warps have written G(c)
        const uint64_t dW20 = umma_smem_desc(sbase + Cfg::OFF_W2, 16, 1024, UMMA_SW_128);
        [[maybe_unused]] const uint64_t dG0 = umma_smem_desc(sbase + Cfg::OFF_G, 16, 1024, UMMA_SW_128);
        uint32_t c = 0;
        for (uint32_t t = 0; t < my_tiles; ++t) {
          const uint32_t ob = t & 1;
          MLP_LAP(7);
          mbar_wait(&o_empty[ob], ((t >> 1) & 1) ^ 1);
          MLP_LAP(0);
          const uint32_t tOut = Cfg::TMEM_O + ob * C;
          for (uint32_t j = 0; j < NCHUNK; ++j, ++c) {
            const uint32_t s = c & 1, sw = c % Cfg::W2_ST;
            MLP_LAP(7);
            mbar_wait(&w2_full[sw], (c / Cfg::W2_ST) & 1);
            MLP_LAP(1);
            mbar_wait(&g_full[s], (c >> 1) & 1);
            MLP_LAP(2);
            tc_fence_after();
            const uint64_t db = dW20 + static_cast<uint64_t>(sw * (Cfg::W2_BYTES >> 4));
            if constexpr (Cfg::G_TMEM) {
              const uint32_t tG = Cfg::TMEM_G + s * 32;      // A operand straight from TMEM: 8 columns per K = 16 step
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(tOut, tG + 8 * k, db + static_cast<uint64_t>(2 * k), idesc2, (j > 0 || k > 0) ? 1u : 0u);
            } else {
              const uint64_t da = dG0 + static_cast<uint64_t>(s * (Cfg::G_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ss(tOut, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc2, (j > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&w2_empty[sw]);
            umma_commit(&g_empty[s]);
            if (j == NCHUNK - 1) umma_commit(&o_full[ob]);
            MLP_LAP(3);
          }
        }
#ifdef MLP_TIMING
        if (blockIdx.x == 5)
          printf("MLPT GEMM2 issuer: total %lld | o_empty wait %lld w2_full wait %lld g_full wait %lld issue %lld | other %lld (chunks %u)\n",
                 clock64() - t_begin, tacc[0], tacc[1], tacc[2], tacc[3], tacc[7], c);
#endif
      }
    }
  } else if (warp < 4) {
    // warp 3: idle (fills the aux warpgroup)
  } else if (warp < 12) {
    // ============================ GELU warps ============================
    const int ew = warp - 4, q = warp & 3, half = ew >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t n_c = 0;
    MLP_T_DECL;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      // ---- per chunk: H -> gelu -> G (swizzled smem) ----
      for (int j = 0; j < NCHUNK; ++j, ++n_c) {
        const int s = n_c & 1;
        const uint32_t ph = (n_c >> 1) & 1;
        MLP_LAP(7);
        if (lane == 0) mbar_wait(&h_full[s], ph);          // one lane polls: 256 lanes on try_wait slow every barrier operation
        MLP_LAP(0);
        __syncwarp();
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_x32(lane_addr + Cfg::TMEM_H + s * HC + half * 32, r);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_empty[s]);           // H buffer may be overwritten by GEMM1(j+2)
        MLP_LAP(1);
        const float* bj = sbias + j * HC + half * 32;
        uint32_t h[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 bb = *reinterpret_cast<const float4*>(bj + 4 * k);
          const float2 g01 = make_float2(gelu_erf_fast(__uint_as_float(r[4 * k]) + bb.x), gelu_erf_fast(__uint_as_float(r[4 * k + 1]) + bb.y));
          const float2 g23 = make_float2(gelu_erf_fast(__uint_as_float(r[4 * k + 2]) + bb.z), gelu_erf_fast(__uint_as_float(r[4 * k + 3]) + bb.w));
          h[2 * k] = pack_half2(g01.x, g01.y);
          h[2 * k + 1] = pack_half2(g23.x, g23.y);
        }
        MLP_LAP(2);
        if (lane == 0) mbar_wait(&g_empty[s], ph ^ 1);     // GEMM2(j-2) has consumed this G buffer
        __syncwarp();
        MLP_LAP(3);
        // G_TMEM: G never touches shared memory: packed fp16 pairs go back into TMEM (row = lane, 16 columns per warp)
        // where GEMM2 reads them as its A operand (tcgen05.mma with A in TMEM), like P in the attention kernel
        if constexpr (Cfg::G_TMEM) {
          tmem_st_x16(lane_addr + Cfg::TMEM_G + s * 32 + half * 16, h);
          tmem_wait_st();
          tc_fence_before();
        } else {
          // G tile = one 128B-swizzled K block: row r_loc, 16-byte chunk c lives at chunk (c ^ (r & 7))
          const int r_loc = q * 32 + lane;
          uint8_t* grow = smem + Cfg::OFF_G + s * Cfg::G_BYTES + r_loc * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(grow + (((half * 4 + k) ^ (r_loc & 7)) << 4)) = make_uint4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_full[s]);
        MLP_LAP(4);
      }
    }
#ifdef MLP_TIMING
    if (blockIdx.x == 5 && lane == 0)
      printf("MLPT gelu warp %d: total %lld | h_full wait %lld ldtm %lld gelu %lld g_empty wait %lld store+fence %lld | other %lld\n",
             warp, clock64() - t_begin, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[7]);
#endif
  } else {
    // ============================ residual warps ============================
    // per tile: x = x + O + b2.  Warp q owns TMEM lanes / tile rows [32q, 32q+32) and walks the row in 32-column
    // chunks: the accumulator chunk is transposed through the warp's smem tile so that the global side moves whole
    // rows (8 lanes x 16 B per row, 4 rows per instruction); the residual of up to GRP chunks is prefetched before the
    // accumulator wait / the previous group's stores.
    const int q = warp & 3;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* stile = reinterpret_cast<float*>(smem + Cfg::OFF_STAGE + q * (32 * 36 * 4));
    constexpr int NCH = C / 32;                            // 32-column chunks of the output row
    constexpr int GRP = 2;                                 // chunks per prefetch group (16 float4 registers)
    const int sub = lane & 7, rsel = lane >> 3;
    uint32_t n_t = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
      const int ob = n_t % Cfg::O_ST;
      const int row0 = tile * TILE_M + q * 32;
#pragma unroll
      for (int g0 = 0; g0 < NCH; g0 += GRP) {
        float4 pre[GRP][8];
#pragma unroll
        for (int ci = 0; ci < GRP; ++ci) {
          const int c0 = (g0 + ci) * 32;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int grow_i = row0 + it * 4 + rsel;
            pre[ci][it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g0 + ci < NCH && grow_i < M) pre[ci][it] = *reinterpret_cast<const float4*>(x + static_cast<size_t>(grow_i) * C + c0 + 4 * sub);
          }
        }
        if (g0 == 0) {
          if (lane == 0) mbar_wait(&o_full[ob], (n_t / Cfg::O_ST) & 1);
          __syncwarp();
          tc_fence_after();
        }
#pragma unroll
        for (int ci = 0; ci < GRP; ++ci) {
          if (g0 + ci >= NCH) break;
          const int c0 = (g0 + ci) * 32;
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
