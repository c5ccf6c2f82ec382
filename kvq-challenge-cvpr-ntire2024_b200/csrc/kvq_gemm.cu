// Persistent warp-specialised tcgen05 GEMM for every Linear / 1x1x1-conv / patch-embed contraction on the
// Swin3D path:   D[M,N] = A[M,K] * B[N,K]^T,  fp16 operands (TMA, 128 B swizzle), fp32 accumulators in TMEM,
// double-buffered so the epilogue of tile t overlaps the MMAs of tile t+1.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d -> 4-stage smem ring, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BLOCK_N, K=16 per instruction)
//   warps 2..5  : epilogue: tcgen05.ld 32x32b (thread = output row) -> fused bias / GELU / residual /
//                 LayerNorm / attention-image scatter / head reduction -> vectorised global stores
//
// The epilogues replace, per reference call site (swin_backbone.py unless noted):
//   EPI_QKV_IMG   WindowAttention3D.forward :253-261  (qkv Linear + bias, q*scale, head split, window layout)
//   EPI_RESID_F32 :322-324 + :472-488 + :509 (proj + window_reverse + un-roll + crop + shortcut),
//                 Mlp.fc2 + residual :490-491, PatchMerging.reduction :554
//   EPI_GELU_F16  Mlp.fc1 + GELU(erf) :64-89
//   EPI_LN_F32    PatchEmbed3D proj bias + LayerNorm :715-733
//   EPI_HEAD      VQAHead fc_hid + GELU + fc_last (models/head.py:60-68)
#include <cstdlib>

#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 halfs = one 128 B swizzle row
constexpr int GEMM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int EPI_WARPS = 8;
constexpr int A_BYTES = BM * BK * 2;
constexpr int WARP_STAGING = 5120;   // per warp: 32x36 fp32 tile (or two 2 KB TMA-store boxes) + 32 row offsets; 512 B aligned
constexpr int STAGING_BYTES = EPI_WARPS * WARP_STAGING;
constexpr int MAX_BIAS_N = 4096;
constexpr int BIAS_BYTES = MAX_BIAS_N * 4;   // the whole bias vector, staged once per CTA (a per-tile __ldg refresh
                                             // of the n-block's slice cost ~13 % of the epilogue's issue samples)

// One CTA per SM: a deep TMA ring (the large-K GEMMs are latency-bound with fewer than 4 stages in flight) and
// double-buffered accumulators so 8 epilogue warps drain tile t while the tensor pipe works on tile t+1.
// MT = 128-row sub-tiles per CTA tile: MT = 2 (256 x BLOCK_N tile, two MMAs per K step sharing the B tile) cuts
// the operand bytes pulled from L2 per FLOP by 30 % -- the stage-2/3 GEMMs (K >= 384) are L2-bandwidth bound
// with 128 x 192 tiles -- at the price of single-buffered accumulators.
// BSTAT ("B stationary", K <= 384): the CTA keeps its [BN x K] weight tile resident in shared memory for the whole kernel
// and walks M tiles, so only the A operand streams through the ring.  The stage-2 GEMMs are bound by the L2 -> SM feed
// (r02 ncu: 40 KB of operands per 560 cycles of MMAs at 128 x 192 tiles = 71 B/clk/SM against ~42 available when every SM
// pulls; tensor pipe 44-46 % active, epilogue warps waiting for accumulators): this form moves 16 KB per K block.
// It did not pay (see use_bstat) and is opt-in.
constexpr int BSTAT_MAX_KB = 6;
template <int BN, int MT = 1, bool BSTAT = false>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = MT * A_BYTES + (BSTAT ? 0 : B_BYTES);
  static constexpr int STAGES = BSTAT ? 4 : MT == 2 ? 3 : (BN >= 256 ? 3 : BN >= 192 ? 4 : BN >= 128 ? 5 : 6);
  static constexpr int BRES_BYTES = BSTAT ? BSTAT_MAX_KB * B_BYTES : 0;
  static constexpr int ACC_STAGES = MT == 2 ? 1 : 2;
  static constexpr int CW = (BN % 64 == 0) ? 32 : 16;       // TMEM columns per epilogue chunk
  static constexpr int NCHUNK = BN / CW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + BRES_BYTES + STAGING_BYTES + 1024 + 256 + BIAS_BYTES;
  static constexpr int TMEM_NEED = ACC_STAGES * MT * BN;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static_assert(TMEM_NEED <= 512 && SMEM <= 227 * 1024, "tile configuration exceeds the SM");
};

__device__ __forceinline__ float gelu_erf(float x) { return gelu_erf_fast(x); }

// QuickGELU of OpenAI CLIP (models/backbones/clip/model.py: x * sigmoid(1.702 * x))
__device__ __forceinline__ float2 quick_gelu2(float2 x) {
  const float2 e = make_float2(fast_exp2(-1.702f * 1.4426950408889634f * x.x), fast_exp2(-1.702f * 1.4426950408889634f * x.y));
  float2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(1.0f + e.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(1.0f + e.y));
  return fmul2(x, r);
}

__device__ __forceinline__ float2 gelu_erf2(float2 x) { return make_float2(gelu_erf_fast(x.x), gelu_erf_fast(x.y)); }

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int CW>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t* r) {
  if constexpr (CW == 32) tmem_ld_x32(taddr, r);
  else tmem_ld_x16(taddr, r);
}

template <int BN, int EPI, int MT, bool BSTAT = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int STAGES = Cfg<BN, MT, BSTAT>::STAGES;
  constexpr int ACC = Cfg<BN, MT, BSTAT>::ACC_STAGES;
  constexpr int CW = Cfg<BN, MT, BSTAT>::CW;
  constexpr int NCHUNK = Cfg<BN, MT, BSTAT>::NCHUNK;
  uint8_t* bres = smem + STAGES * Cfg<BN, MT, BSTAT>::STAGE_BYTES;     // BSTAT: the resident weight tile (K blocks of BN rows)
  uint8_t* staging = bres + Cfg<BN, MT, BSTAT>::BRES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* bfull = bars + 2 * STAGES + 5;    // BSTAT: the weight tile has landed
  float* sbias = reinterpret_cast<float*>(bars + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TM = BM * MT;                 // rows per CTA tile
  const bool conv_mode = (EPI == EPI_CONV_F16) && p.conv.on != 0;
  const int num_m = conv_mode ? p.M / BM : (p.M + TM - 1) / TM;     // conv: p.M = tiles * 128 (padded tile grid)
  const int num_n = p.N / BN;
  const int tiles = num_m * num_n;
  const bool narrow = conv_mode && p.conv.csub != 0;
  const int nkb = (p.K + BK - 1) / BK;   // narrow mode: p.K = 64 * (number of K blocks of 64 / csub taps)
  // split-weight mode: B holds [W_hi | W_lo] along K (each padded to 64); the K loop runs twice over A
  const int nkb_tot = p.split_b ? 2 * nkb : nkb;
  // tile walk.  Default: tile = blockIdx.x + i * gridDim.x over (m, n) with n fastest.  BSTAT: the CTA owns ONE n block and
  // walks every cpn-th m block of it (cpn = CTAs per n block), expressed as (first, stride) over the same tile index.
  int tile_first = blockIdx.x, tile_stride = gridDim.x;
  if constexpr (BSTAT) {
    const int cpn = static_cast<int>(gridDim.x) / num_n;               // the launcher guarantees >= 1
    const int nb = static_cast<int>(blockIdx.x) / cpn, rank = static_cast<int>(blockIdx.x) - nb * cpn;
    tile_first = nb < num_n ? rank * num_n + nb : tiles;               // left-over CTAs idle
    tile_stride = cpn * num_n;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], (EPI == EPI_CONV_F16 && MT == 1) ? EPI_WARPS / 2 : EPI_WARPS);
    }
    mbar_init(bfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg<BN, MT, BSTAT>::TMEM_COLS);
    tmem_relinquish();
  }
  if constexpr (EPI == EPI_CONV_F16) {
    if (narrow) {
      // a K block whose tap count is not a multiple of 64 / csub leaves sub-tiles unloaded; their weights are zero, so
      // the smem behind them only has to be finite: clear the ring once (0 x NaN would poison the accumulator)
      for (int i = threadIdx.x; i < STAGES * Cfg<BN, MT, BSTAT>::STAGE_BYTES / 16; i += GEMM_THREADS)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
    }
  }
  if constexpr (EPI != EPI_LN_F32 && EPI != EPI_HEAD) {
    // bias is a weight (never produced by the previous kernel): stage it before the dependency wait
    const int nb = p.N < MAX_BIAS_N ? p.N : MAX_BIAS_N;
    const int nval = (EPI == EPI_CONV_F16 && p.nvalid > 0) ? p.nvalid : p.N;
    for (int i = threadIdx.x; i < nb; i += GEMM_THREADS)
      sbias[i] = (p.bias != nullptr && i < nval) ? __ldg(p.bias + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      if constexpr (BSTAT) {
        if (tile_first < tiles) {
          const int n_blk = tile_first % num_n;
          mbar_expect_tx(bfull, static_cast<uint32_t>(nkb * Cfg<BN, MT, BSTAT>::B_BYTES));
          for (int kb = 0; kb < nkb; ++kb) tma_load_2d(bres + kb * Cfg<BN, MT, BSTAT>::B_BYTES, &tmB, bfull, kb * BK, n_blk * BN);
        }
      }
      for (int tile = tile_first; tile < tiles; tile += tile_stride) {
        const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
        int cw0 = 0, ch0 = 0, ct0 = 0, cb = 0, cblocks = 1;
        if constexpr (EPI == EPI_CONV_F16) {
          if (conv_mode) {   // tile -> (clip, t/h/w block); coordinates of its first input pixel for tap (0,0,0)
            int r = m_blk;
            const int tw = r % p.conv.nw; r /= p.conv.nw;
            const int th = r % p.conv.nh; r /= p.conv.nh;
            const int tt = r % p.conv.nt;
            cb = r / p.conv.nt;
            cw0 = tw * p.conv.bw * p.conv.sw - p.conv.pw;
            ch0 = th * p.conv.bh * p.conv.sh - p.conv.ph;
            ct0 = tt * p.conv.bt * p.conv.st - p.conv.pt;
            cblocks = p.conv.C / BK;
          }
        }
        int tap_w = 0, tap_h = 0, tap_t = 0, cblk = 0;
        for (int kb = 0; kb < nkb_tot; ++kb) {
          mbar_wait(&empty[stage], ph ^ 1);
          uint8_t* sA = smem + stage * Cfg<BN, MT, BSTAT>::STAGE_BYTES;
          uint8_t* sB = sA + MT * A_BYTES;
          if (EPI == EPI_CONV_F16 && narrow) {
            // 64 / csub taps per K block, one [128 pixels x csub] sub-tile each; the weight image is one bulk copy
            const int tpb = BK / p.conv.csub, sub_bytes = BM * p.conv.csub * 2;
            const int first = kb * tpb;
            const int ntap = p.conv.taps - first < tpb ? p.conv.taps - first : tpb;
            mbar_expect_tx(&full[stage], static_cast<uint32_t>(ntap * sub_bytes + BN * BK * 2));
            for (int t = 0; t < ntap; ++t) {
              tma_load_5d(sA + t * sub_bytes, &tmA, &full[stage], 0, cw0 + tap_w, ch0 + tap_h, ct0 + tap_t, cb);
              if (++tap_w == p.conv.kw) { tap_w = 0; if (++tap_h == p.conv.kh) { tap_h = 0; ++tap_t; } }
            }
            bulk_load_1d(sB, static_cast<const uint8_t*>(p.conv.wimg) + static_cast<size_t>(kb) * (BN * BK * 2),
                         BN * BK * 2, &full[stage]);
            if (++stage == STAGES) { stage = 0; ph ^= 1; }
            continue;
          }
          mbar_expect_tx(&full[stage], Cfg<BN, MT, BSTAT>::STAGE_BYTES);
          if (EPI == EPI_CONV_F16 && conv_mode) {
            tma_load_5d(sA, &tmA, &full[stage], cblk * BK, cw0 + tap_w, ch0 + tap_h, ct0 + tap_t, cb);
            if (++cblk == cblocks) {
              cblk = 0;
              if (++tap_w == p.conv.kw) { tap_w = 0; if (++tap_h == p.conv.kh) { tap_h = 0; ++tap_t; } }
            }
          } else {
            tma_load_2d(sA, &tmA, &full[stage], (kb >= nkb ? kb - nkb : kb) * BK, m_blk * TM);
          }
          if constexpr (!BSTAT) tma_load_2d(sB, &tmB, &full[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN, 0, 0);
      int stage = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      if constexpr (BSTAT) {
        if (tile_first < tiles) mbar_wait(bfull, 0);
      }
      for (int tile = tile_first; tile < tiles; tile += tile_stride) {
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * MT * BN);
        for (int kb = 0; kb < nkb_tot; ++kb) {
          mbar_wait(&full[stage], ph);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * Cfg<BN, MT, BSTAT>::STAGE_BYTES);
          if (EPI == EPI_CONV_F16 && narrow) {
            // four K = 16 steps per block; where they live depends on the sub-tile width (see GemmParams::conv)
            const uint32_t sBa = sA + A_BYTES;
            const int cs = p.conv.csub;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              uint64_t da, db2;
              if (cs == 8) {          // two taps per step: core matrices 8 rows x 16 B, K-adjacent one sub-tile apart
                da = umma_smem_desc(sA + 2 * m * (BM * 16), BM * 16, 128, UMMA_SW_NONE);
                db2 = umma_smem_desc(sBa + 2 * m * (BN * 16), BN * 16, 128, UMMA_SW_NONE);
              } else if (cs == 16) {  // one tap per step, 32 B swizzle
                da = umma_smem_desc(sA + m * (BM * 32), 16, 256, UMMA_SW_32);
                db2 = umma_smem_desc(sBa + m * (BN * 32), 16, 256, UMMA_SW_32);
              } else {                // half a tap per step, 64 B swizzle
                da = umma_smem_desc(sA + (m >> 1) * (BM * 64) + (m & 1) * 32, 16, 512, UMMA_SW_64);
                db2 = umma_smem_desc(sBa + (m >> 1) * (BN * 64) + (m & 1) * 32, 16, 512, UMMA_SW_64);
              }
              umma_f16_ss(d_tmem, da, db2, idesc, (kb > 0 || m > 0) ? 1u : 0u);
            }
            umma_commit(&empty[stage]);
            if (++stage == STAGES) { stage = 0; ph ^= 1; }
            continue;
          }
          const uint64_t db = BSTAT ? umma_smem_desc(smem_u32(bres) + kb * Cfg<BN, MT, BSTAT>::B_BYTES, 16, 1024, UMMA_SW_128)
                                    : umma_smem_desc(sA + MT * A_BYTES, 16, 1024, UMMA_SW_128);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t da = umma_smem_desc(sA + mt * A_BYTES, 16, 1024, UMMA_SW_128);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advancing 16 halfs (32 B) inside the 128 B swizzle row = +2 in the (addr >> 4) field
              umma_f16_ss(d_tmem + mt * BN, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
        umma_commit(&tfull[as]);
        if (++as == ACC) { as = 0; aph ^= 1; }
      }
    }
  } else {
    const int ew = warp - 2;        // 0..7
    const int q = warp & 3;         // TMEM lane quadrant this warp may touch
    const int half = ew >> 2;       // MT == 1: which interleaved set of column chunks this warp drains;
                                    // MT == 2: which 128-row sub-tile (the warp then drains every chunk of its rows)
    // epilogues that need the whole output row in one thread run on the half-0 warps only
    constexpr bool kWholeRow = (EPI == EPI_HEAD);
    // conv epilogue: the two warp sets take alternate TILES (= alternate accumulator stages) instead of alternate
    // column chunks, so two tiles' load -> TMEM -> store latency chains overlap.  The thin-K convolutions
    // (K = 8..96, one k-block per tile) are bound by exactly that chain.
    constexpr bool kTileSplit = (EPI == EPI_CONV_F16 && MT == 1);
    const int c_begin = (kWholeRow || MT == 2 || kTileSplit) ? 0 : half;
    constexpr int c_step = (kWholeRow || MT == 2 || kTileSplit) ? 1 : 2;
    const bool active = !kWholeRow || half == 0;
    const int msub = MT == 2 ? half : 0;
    float* stile = reinterpret_cast<float*>(staging + ew * WARP_STAGING);
    long long* srow = reinterpret_cast<long long*>(stile + 32 * 36);
    int as = 0;
    uint32_t aph = 0;
    int bias_nblk = -1;
    uint32_t nstore = 0;          // conv epilogue: TMA stores issued by this warp (selects the smem box)
    for (int tile = tile_first; tile < tiles; tile += tile_stride) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      if constexpr (kTileSplit) {
        if (as != half) {                     // the other warp set owns this tile / accumulator stage
          if (++as == ACC) { as = 0; aph ^= 1; }
          continue;
        }
      }
      if constexpr (EPI != EPI_RESID_F32 && EPI != EPI_CONV_F16) {   // (those two prefetch their residual first)
        mbar_wait(&tfull[as], aph);
        __syncwarp();
        tc_fence_after();
      }
      const int row = m_blk * TM + msub * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>((as * MT + msub) * BN);
      const int nbase = n_blk * BN;

      if constexpr (EPI == EPI_GELU_F16 || EPI == EPI_STORE_F16) {
        __half* orow = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(row) * p.ldo + nbase;
        constexpr int MYCH = MT == 2 ? NCHUNK : (NCHUNK + 1) / 2;
        // this warp's bias slice lives in its private smem tile; refreshed only when the CTA moves to another n-block
        if (n_blk != bias_nblk) {
          __syncwarp();
#pragma unroll
          for (int ci = 0; ci < MYCH; ++ci) {
            const int c0 = (c_begin + ci * c_step) * CW;
            if (lane < CW && c0 < BN) stile[ci * CW + lane] = sbias[nbase + c0 + lane];
          }
          bias_nblk = n_blk;
          __syncwarp();
        }
        uint32_t r[2][CW];
        tmem_ld_chunk<CW>(taddr + c_begin * CW, r[0]);
#pragma unroll
        for (int ci = 0; ci < MYCH; ++ci) {
          const int c0 = (c_begin + ci * c_step) * CW;
          if (c0 >= BN) break;
          tmem_wait_ld();
          if (ci + 1 < MYCH && c0 + c_step * CW < BN) tmem_ld_chunk<CW>(taddr + c0 + c_step * CW, r[(ci + 1) & 1]);
          const uint32_t* rc = r[ci & 1];
          uint32_t h[CW / 2];
          // the activation switch sits OUTSIDE the element loop: with the branch inside, every float4 group was its own
          // basic block and the GELU chains of only four elements overlapped (r02 ncu: the fc1 epilogue warps were
          // stalled on dependencies 52 % of the time, the kernel epilogue-bound at 28 % tensor-pipe activity)
          auto act_chunk = [&](auto act) {
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const float4 bb = *reinterpret_cast<const float4*>(stile + ci * CW + 4 * j);
              const float2 v01 = act(fadd2(make_float2(__uint_as_float(rc[4 * j]), __uint_as_float(rc[4 * j + 1])), make_float2(bb.x, bb.y)));
              const float2 v23 = act(fadd2(make_float2(__uint_as_float(rc[4 * j + 2]), __uint_as_float(rc[4 * j + 3])), make_float2(bb.z, bb.w)));
              h[2 * j] = pack_half2(v01.x, v01.y);
              h[2 * j + 1] = pack_half2(v23.x, v23.y);
            }
          };
          if constexpr (EPI == EPI_GELU_F16) {
            if (p.quick_gelu) act_chunk([](float2 v) { return quick_gelu2(v); });
            else act_chunk([](float2 v) { return gelu_erf2(v); });
          } else {
            act_chunk([](float2 v) { return v; });
          }
          // transpose through the warp's smem tile (80 B row pitch: conflict-free both ways) so that every store
          // instruction writes whole 32 B sectors: 32/SEGS rows x (CW*2) contiguous bytes instead of 32 x 16 B
          constexpr int SEGS = CW / 8;                    // 16-byte segments per row chunk
          uint8_t* st16 = reinterpret_cast<uint8_t*>(stile) + 1024;   // (bias slice occupies the first <= 768 B)
          __syncwarp();
#pragma unroll
          for (int j = 0; j < SEGS; ++j)
            *reinterpret_cast<uint4*>(st16 + lane * 80 + j * 16) = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
          __syncwarp();
          constexpr int RPI = 32 / SEGS;                  // rows per store instruction
          const int seg = lane / RPI, rl = lane % RPI;
#pragma unroll
          for (int it = 0; it < SEGS; ++it) {
            const int rr = it * RPI + rl;
            const uint4 v = *reinterpret_cast<const uint4*>(st16 + rr * 80 + seg * 16);
            const int grow = m_blk * TM + msub * BM + q * 32 + rr;
            if (grow < p.M)
              st_global_v4(reinterpret_cast<__half*>(p.out) + static_cast<size_t>(grow) * p.ldo + nbase + c0 + 8 * seg,
                           v.x, v.y, v.z, v.w);
          }
        }
        (void)orow;
        (void)row_ok;
      } else if constexpr (EPI == EPI_CONV_F16) {
        // conv-as-GEMM epilogue: folded-BN shift, optional fp16 residual, optional ReLU, fp16 channels-last output
        constexpr int MYCH = (MT == 2 || kTileSplit) ? NCHUNK : (NCHUNK + 1) / 2;
        const int nvalid = p.nvalid > 0 ? p.nvalid : p.N;
        // output row of this thread's accumulator row: identity, or (implicit conv) the pixel of the tile's
        // bt x bh x bw block; -1 = outside the map
        long long orow = row_ok ? row : -1;
        if (conv_mode && p.resid_h != nullptr) {   // (only the residual load needs the per-thread row)
          int r = m_blk;
          const int tw = r % p.conv.nw; r /= p.conv.nw;
          const int th = r % p.conv.nh; r /= p.conv.nh;
          const int tt = r % p.conv.nt;
          const int b = r / p.conv.nt;
          int l = q * 32 + lane;
          const int lw = l % p.conv.bw; l /= p.conv.bw;
          const int lh = l % p.conv.bh;
          const int lt = l / p.conv.bh;
          const int ow = tw * p.conv.bw + lw, oh = th * p.conv.bh + lh, ot = tt * p.conv.bt + lt;
          orow = (ow < p.conv.Wo && oh < p.conv.Ho && ot < p.conv.To)
                     ? ((static_cast<long long>(b) * p.conv.To + ot) * p.conv.Ho + oh) * p.conv.Wo + ow
                     : -1;
        }
        // this thread's residual row is fetched one chunk ahead, and the first chunk BEFORE the accumulator wait: the
        // thin-K convolutions are bound by this load's latency, not by the tensor pipe
        uint4 rsb[2][CW / 8];
        auto load_resid = [&](int ci, uint4* dst) {
          const int c0 = (c_begin + ci * c_step) * CW;
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) dst[j] = make_uint4(0u, 0u, 0u, 0u);
          if (p.resid_h != nullptr && orow >= 0 && c0 < BN) {
#pragma unroll
            for (int j = 0; j < CW / 8; ++j)
              if (nbase + c0 + 8 * j < nvalid)
                dst[j] = *reinterpret_cast<const uint4*>(p.resid_h + orow * p.ldr + nbase + c0 + 8 * j);
          }
        };
        load_resid(0, rsb[0]);
        mbar_wait(&tfull[as], aph);
        __syncwarp();
        tc_fence_after();
        // Each 32-row x 32-channel chunk goes registers -> 64B-swizzled smem box -> ONE TMA store (2-D rows, or the
        // 5-D pixel block of the implicit convolution; the tensor map clips rows / pixels / channels outside the
        // output).  The transpose + st.global path it replaces kept l1tex at 71-82 % of peak (ncu).
        static_assert(CW == 32, "conv epilogue stores 32-channel boxes");
        uint8_t* sbuf = reinterpret_cast<uint8_t*>(stile);            // two 2 KB boxes
        int oc1 = m_blk * TM + q * 32, oc2 = 0, oc3 = 0, oc4 = 0;      // store coordinates above the channel dim
        if (conv_mode) {
          int rr = m_blk;
          const int tw = rr % p.conv.nw; rr /= p.conv.nw;
          const int th = rr % p.conv.nh; rr /= p.conv.nh;
          const int tt = rr % p.conv.nt;
          oc4 = rr / p.conv.nt;
          int l = q * 32;
          const int lw = l % p.conv.bw; l /= p.conv.bw;
          oc1 = tw * p.conv.bw + lw;
          oc2 = th * p.conv.bh + l % p.conv.bh;
          oc3 = tt * p.conv.bt + l / p.conv.bh;
        }
        uint32_t r[2][CW];
        tmem_ld_chunk<CW>(taddr, r[0]);
#pragma unroll
        for (int ci = 0; ci < MYCH; ++ci) {
          const int c0 = ci * CW;
          if (ci + 1 < MYCH) load_resid(ci + 1, rsb[(ci + 1) & 1]);
          const uint4* rs = rsb[ci & 1];
          tmem_wait_ld();
          if (ci + 1 < MYCH) tmem_ld_chunk<CW>(taddr + c0 + CW, r[(ci + 1) & 1]);
          if (nbase + c0 >= nvalid) continue;                          // padded output channels: nothing to store
          const uint32_t* rc = r[ci & 1];
          const float* bs = sbias + nbase + c0;
          uint32_t h[CW / 2];
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) {
            const uint32_t rw = reinterpret_cast<const uint32_t*>(rs)[j];
            const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rw));
            const float2 bb = *reinterpret_cast<const float2*>(bs + 2 * j);
            float v0 = __uint_as_float(rc[2 * j]) + bb.x + rf.x;
            float v1 = __uint_as_float(rc[2 * j + 1]) + bb.y + rf.y;
            if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            h[j] = pack_half2(v0, v1);
          }
          uint8_t* buf = sbuf + (nstore & 1) * 2048;
          if (lane == 0) tma_store_wait_read<1>();                     // the store that last used this box has read it
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (conv_mode) tma_store_5d(&tmC, buf, nbase + c0, oc1, oc2, oc3, oc4);
            else tma_store_2d(&tmC, buf, nbase + c0, oc1);
            tma_store_commit();
          }
          ++nstore;
        }
      } else if constexpr (EPI == EPI_RESID_F32) {
        // Row-per-thread TMEM reads are transposed through a per-warp smem tile so that global traffic is
        // row-contiguous: CW/4 lanes cover one output row (full 32 B sectors for the fp32 read-modify-write).
        long long orow_idx = row;
        bool ok = row_ok;
        if (p.remap) {
          const int rows_in = p.geom.nW * p.geom.N;
          const int b = fdiv_i(row, rows_in, p.geom.r_rows);
          const int src = win_row_to_src(p.geom, row - b * rows_in);
          ok = ok && src >= 0;
          orow_idx = static_cast<long long>(b) * p.geom.tokens + src;
        }
        srow[lane] = ok ? orow_idx * p.ldo : -1;
        constexpr int LPRW = CW / 4;          // lanes per row on the global side
        constexpr int RPI = 32 / LPRW;        // rows per instruction
        constexpr int MYCH = MT == 2 ? NCHUNK : (NCHUNK + 1) / 2;  // chunks this warp drains
        const int sub = lane % LPRW, rsel = lane / LPRW;
        __syncwarp();
        // prefetch the residual rows while the tensor pipe is still producing this tile: the HBM latency of the
        // fp32 stream is hidden behind the accumulator wait instead of being paid once per chunk
        constexpr int PF = MYCH < 3 ? MYCH : 3;   // chunks whose residual is prefetched (register budget)
        float4 pre[PF][LPRW];
        if (p.resid != nullptr) {
#pragma unroll
          for (int ci = 0; ci < PF; ++ci) {
            const int c0 = (c_begin + ci * c_step) * CW;
#pragma unroll
            for (int it = 0; it < LPRW; ++it) {
              const long long off = srow[it * RPI + rsel];
              pre[ci][it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (off >= 0 && c0 < BN) pre[ci][it] = *reinterpret_cast<const float4*>(p.resid + off + nbase + c0 + 4 * sub);
            }
          }
        }
        mbar_wait(&tfull[as], aph);
        __syncwarp();
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < MYCH; ++ci) {
          const int c0 = (c_begin + ci * c_step) * CW;
          if (c0 >= BN) break;
          uint32_t r[CW];
          tmem_ld_chunk<CW>(taddr + c0, r);
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          bb = *(reinterpret_cast<const float4*>(sbias + nbase + c0) + sub);
          tmem_wait_ld();
          __syncwarp();                       // previous chunk's readers are done with the tile
#pragma unroll
          for (int j = 0; j < CW / 4; ++j)
            *reinterpret_cast<uint4*>(stile + lane * 36 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < LPRW; ++it) {
            const int rr = it * RPI + rsel;
            const long long off = srow[rr];
            if (off >= 0) {
              float4 v = *reinterpret_cast<const float4*>(stile + rr * 36 + 4 * sub);
              v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
              if (p.resid != nullptr) {
                float4 rv;
                if (ci < PF) rv = pre[ci < PF ? ci : 0][it];
                else rv = *reinterpret_cast<const float4*>(p.resid + off + nbase + c0 + 4 * sub);
                v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
              }
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off + nbase + c0 + 4 * sub) = v;
            }
          }
        }
        __syncwarp();                         // srow / tile are rewritten by the next tile
      } else if constexpr (EPI == EPI_QKV_IMG) {
        const int ntok = p.geom.N;
        const int win_g = fdiv_i(row, ntok, p.geom.r_N);   // (float-reciprocal division: kvq_kernels.cuh)
        const int i = row - win_g * ntok;
        const int slab = fdiv_i(i, p.geom.SL, p.geom.r_SL);
        const int pitch = p.att_pitch;
        // pitch 50 / 52: slab-padded key slots d*pitch + h*7 + w; pitch 56 (third generation): h*56 + w*8 + d
        const bool hwd = pitch == ATT3_PITCH;
        const int kv_rows = hwd ? ATT3_KV_ROWS : 8 * pitch;    // 400 or 416 key slots
        const int kv_bytes = kv_rows * ATT_HD * 2;
        const size_t unit_bytes = static_cast<size_t>(ATT_IMG_BYTES) + 2 * kv_bytes;
        const int pos = i - slab * p.geom.SL;
        // third generation: image rows are (h,w,d) d-fastest for Q, K and V alike.  With geom.dfast the GEMM rows already
        // arrive in that order (row i -> image row i: consecutive tokens fill whole core-matrix lines); natural-order
        // rows (the stand-alone op) are scattered to the same places
        const int hwd_row = p.geom.dfast ? i : pos * 8 + slab;
        const int kv_row = hwd ? hwd_row : slab * pitch + pos;
        constexpr int MYCH = MT == 2 ? NCHUNK : (NCHUNK + 1) / 2;
        if (n_blk != bias_nblk) {     // bias slice of this warp's chunks -> private smem, once per n-block
          __syncwarp();
#pragma unroll
          for (int ci = 0; ci < MYCH; ++ci) {
            const int c0 = (c_begin + ci * c_step) * CW;
            if (lane < CW && c0 < BN) stile[ci * CW + lane] = sbias[nbase + c0 + lane];
          }
          bias_nblk = n_blk;
          __syncwarp();
        }
        uint32_t r[2][CW];
        tmem_ld_chunk<CW>(taddr + c_begin * CW, r[0]);
#pragma unroll
        for (int ci = 0; ci < MYCH; ++ci) {
          const int c0 = (c_begin + ci * c_step) * CW;
          if (c0 >= BN) break;
          tmem_wait_ld();
          if (ci + 1 < MYCH && c0 + c_step * CW < BN) tmem_ld_chunk<CW>(taddr + c0 + c_step * CW, r[(ci + 1) & 1]);
          const uint32_t* rc = r[ci & 1];
          const int n0 = nbase + c0;
          const int which = n0 / p.C;
          const int cin = n0 - which * p.C;       // channel inside q / k / v
          const int head = cin >> 5;
          const int kc0 = (cin & 31) >> 3;        // first 16-byte chunk of the head's 32 channels
          const float2 sc2 = splat2(which == 0 ? p.qscale : 1.0f);
          uint32_t h[CW / 2];
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const float4 bb = *reinterpret_cast<const float4*>(stile + ci * CW + 4 * j);
            const float2 v01 = fmul2(fadd2(make_float2(__uint_as_float(rc[4 * j]), __uint_as_float(rc[4 * j + 1])), make_float2(bb.x, bb.y)), sc2);
            const float2 v23 = fmul2(fadd2(make_float2(__uint_as_float(rc[4 * j + 2]), __uint_as_float(rc[4 * j + 3])), make_float2(bb.z, bb.w)), sc2);
            h[2 * j] = pack_half2(v01.x, v01.y);
            h[2 * j + 1] = pack_half2(v23.x, v23.y);
          }
          if (row_ok) {
            const int rimg = which == 0 ? (hwd ? hwd_row : i) : kv_row;
            uint8_t* dst = reinterpret_cast<uint8_t*>(p.img) + (static_cast<size_t>(win_g) * p.heads + head) * unit_bytes +
                           (which == 0 ? 0 : ATT_IMG_BYTES + (which - 1) * kv_bytes);
#pragma unroll
            for (int j = 0; j < CW / 8; ++j)
              st_global_v4(dst + att_img_offset(rimg, kc0 + j), h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
            // the last token of a slab also zeroes the padded K/V key slots behind it (P is 0 there, but 0 * NaN
            // from uninitialised workspace would poison PV); the last token of the window zeroes all the rest
            const bool pad_owner = hwd ? (hwd_row == 391) : (pos == p.geom.SL - 1);
            if (which != 0 && pad_owner) {
              const int end = hwd ? kv_row + 9 : (i == ntok - 1) ? kv_rows : (slab + 1) * pitch;
              for (int rz = kv_row + 1; rz < end; ++rz) {
#pragma unroll
                for (int j = 0; j < CW / 8; ++j) st_global_v4(dst + att_img_offset(rz, kc0 + j), 0u, 0u, 0u, 0u);
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_LN_F32) {
        // PatchEmbed3D bias + LayerNorm(96): the two warps that share a lane quadrant each own half of the row's
        // column chunks; one TMEM pass accumulates sum / sum-of-squares, the halves are combined through shared
        // memory (named barrier over the 8 epilogue warps), a second pass normalises and stores.
        constexpr int MYCH = MT == 2 ? NCHUNK : (NCHUNK + 1) / 2;
        float* sstat = stile;                         // this warp's 32 x {sum, sumsq} partials
        float* pstat = reinterpret_cast<float*>(staging + (ew ^ 4) * WARP_STAGING);   // partner warp's
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int ci = 0; ci < MYCH; ++ci) {
          const int c0 = (c_begin + ci * c_step) * CW;
          if (c0 >= BN) break;
          uint32_t r[CW];
          tmem_ld_chunk<CW>(taddr + c0, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const float v = __uint_as_float(r[j]) + __ldg(p.bias + c0 + j);
            s1 += v;
            s2 = fmaf(v, v, s2);
          }
        }
        sstat[2 * lane] = s1;
        sstat[2 * lane + 1] = s2;
        named_bar_sync(2, EPI_WARPS * 32);
        s1 += pstat[2 * lane];
        s2 += pstat[2 * lane + 1];
        const float mean = s1 * (1.0f / BN);
        const float rstd = rsqrtf(fmaxf(s2 * (1.0f / BN) - mean * mean, 0.f) + p.eps);
        float* orow = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo;
#pragma unroll
        for (int ci = 0; ci < MYCH; ++ci) {
          const int c0 = (c_begin + ci * c_step) * CW;
          if (c0 >= BN) break;
          uint32_t r[CW];
          tmem_ld_chunk<CW>(taddr + c0, r);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < CW; j += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
              const float4 gg = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + j));
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + j));
              float4 v;
              v.x = (__uint_as_float(r[j]) + bb.x - mean) * rstd * gg.x + be.x;
              v.y = (__uint_as_float(r[j + 1]) + bb.y - mean) * rstd * gg.y + be.y;
              v.z = (__uint_as_float(r[j + 2]) + bb.z - mean) * rstd * gg.z + be.z;
              v.w = (__uint_as_float(r[j + 3]) + bb.w - mean) * rstd * gg.w + be.w;
              *reinterpret_cast<float4*>(orow + c0 + j) = v;
            }
          }
        }
        named_bar_sync(2, EPI_WARPS * 32);            // partials are rewritten by the next tile
      } else if constexpr (EPI == EPI_HEAD) {
        if (active) {
          float acc = 0.f;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += CW) {
            uint32_t r[CW];
            tmem_ld_chunk<CW>(taddr + c0, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < CW; ++j)
              acc += gelu_erf(__uint_as_float(r[j]) + __ldg(p.bias + c0 + j)) * __ldg(p.w2 + c0 + j);
          }
          if (row_ok) p.rowscore[row] = acc + __ldg(p.b2ptr);
        }
      }

      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == ACC) { as = 0; aph ^= 1; }
    }
    if constexpr (EPI == EPI_CONV_F16) {
      if (lane == 0) tma_store_wait_all<0>();   // bulk stores complete (and their smem boxes are free) before exit
    }
    (void)nstore;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg<BN, MT, BSTAT>::TMEM_COLS);
  }
}

template <int BN, int EPI, int MT = 1, bool BSTAT = false>
int launch_impl(const __half* A, int lda, const __half* B, int ldb, const GemmParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    KVQ_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI, MT, BSTAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, MT, BSTAT>::SMEM));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, A, p.M, p.K, static_cast<uint64_t>(lda) * 2, BM * MT, BK, 2, 128);
  if (rc != 0) return rc;
  const int kcols_b = p.split_b ? 2 * ((p.K + BK - 1) / BK * BK) : p.K;
  KVQ_REQUIRE(!p.split_b || ldb == kcols_b, KVQ_ERR_BAD_SHAPE, "gemm: split weights need ldb == 2*ceil64(K) (%d vs %d)",
              ldb, kcols_b);
  rc = make_tmap_2d(&tmB, B, p.N, kcols_b, static_cast<uint64_t>(ldb) * 2, BN, BK, 2, 128);
  if (rc != 0) return rc;
  CUtensorMap tmC = tmB;   // only the conv epilogue stores through a tensor map
  if constexpr (EPI == EPI_CONV_F16) {
    const int cols = p.nvalid > 0 ? p.nvalid : p.N;
    rc = make_tmap_2d(&tmC, p.out, p.M, cols, static_cast<uint64_t>(p.ldo) * 2, 32, 32, 2, 64);
    if (rc != 0) return rc;
  }
  const int tiles = ((p.M + BM * MT - 1) / (BM * MT)) * (p.N / BN);
  int grid = tiles < num_sms() ? tiles : num_sms();
  if constexpr (BSTAT) grid = (num_sms() / (p.N / BN)) * (p.N / BN);   // a whole number of CTAs per n block
  count_launch();
  return launch_pdl(gemm_kernel<BN, EPI, MT, BSTAT>, dim3(grid), dim3(GEMM_THREADS), Cfg<BN, MT, BSTAT>::SMEM, stream, tmA, tmB, tmC,
                    p);
}

// 256-row tiles (opt-in, KVQ_GEMM_BIG_TILES=1): fewer L2 bytes per FLOP, but at batch 8 the stage-2/3 GEMMs then have
// only ~1.3 waves of tiles and lose the accumulator double buffering -- measured 1.3-2x slower, so off by default
inline bool use_big_tiles(const GemmParams& p) {
  static const int mode = []() { const char* e = getenv("KVQ_GEMM_BIG_TILES"); return e ? atoi(e) : 0; }();
  return mode != 0 && p.K >= 384 && (p.M + 255) / 256 * (p.N / 192) >= num_sms();
}

// B-stationary form (see Cfg): K <= 384, N a multiple of 128 with at most one n block per SM, enough m tiles that every
// CTA of an n block gets a few.  Opt-in (KVQ_GEMM_BSTAT=1): measured SLOWER than the 128 x 192 ring form on the stage-2
// GEMMs (qkv 0.223 vs 0.187 ms, fc1 0.336 vs 0.306 ms per step) -- the N = 128 MMA costs 107 cycles for 64 of math, and
// the smaller operand stream does not make up for it.
inline bool use_bstat(const GemmParams& p) {
  static const int mode = []() { const char* e = getenv("KVQ_GEMM_BSTAT"); return e ? atoi(e) : 0; }();
  if (mode == 0 || p.split_b || p.K > 64 * BSTAT_MAX_KB || p.K % 64 != 0 || p.N % 128 != 0) return false;
  const int num_n = p.N / 128, num_m = (p.M + BM - 1) / BM;
  if (num_n > num_sms()) return false;
  return num_m >= 4 * (num_sms() / num_n);
}

template <int EPI>
int launch_bn(const __half* A, int lda, const __half* B, int ldb, const GemmParams& p, cudaStream_t stream) {
  // 128 x 256 tiles where N allows and every SM still gets a few tiles (fc1 of stages 2-3: 0.306 -> 0.286 ms at stage 2;
  // KVQ_GEMM_BN256=0 goes back to 128 x 192)
  static const int bn256 = []() { const char* e = getenv("KVQ_GEMM_BN256"); return e ? atoi(e) : 1; }();
  if constexpr (EPI == EPI_GELU_F16) {
    if (bn256 && p.N % 256 == 0 && ((p.M + BM - 1) / BM) * (p.N / 256) >= 2 * num_sms())
      return launch_impl<256, EPI>(A, lda, B, ldb, p, stream);
  }
  if (use_bstat(p)) return launch_impl<128, EPI, 1, true>(A, lda, B, ldb, p, stream);
  if (p.N % 192 == 0 && use_big_tiles(p)) return launch_impl<192, EPI, 2>(A, lda, B, ldb, p, stream);
  if (p.N % 192 == 0) return launch_impl<192, EPI>(A, lda, B, ldb, p, stream);
  if (p.N % 96 == 0) return launch_impl<96, EPI>(A, lda, B, ldb, p, stream);
  if (p.N % 64 == 0) return launch_impl<64, EPI>(A, lda, B, ldb, p, stream);
  set_error("gemm: N=%d is not a multiple of 64/96/192", p.N);
  return KVQ_ERR_BAD_SHAPE;
}

template <int BN>
int launch_conv_impl(const CUtensorMap& tmA, const CUtensorMap& tmC, const __half* Wt, int K, const GemmParams& p,
                     cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    KVQ_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI_CONV_F16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg<BN, 1>::SMEM));
    attr_set = true;
  }
  CUtensorMap tmB;
  int rc = make_tmap_2d(&tmB, Wt, p.N, K, static_cast<uint64_t>(K) * 2, BN, BK, 2, 128);
  if (rc != 0) return rc;
  const int tiles = (p.M / BM) * (p.N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  count_launch();
  return launch_pdl(gemm_kernel<BN, EPI_CONV_F16, 1>, dim3(grid), dim3(GEMM_THREADS), Cfg<BN, 1>::SMEM, stream, tmA,
                    tmB, tmC, p);
}

}  // namespace

bool conv_implicit_supported(int C, int kt, int kh, int kw, int st, int sh, int sw) {
  static const int off = []() { const char* e = getenv("KVQ_CONV_EXPLICIT"); return e ? atoi(e) : 0; }();
  return off == 0 && C % 64 == 0 && kt * kh * kw >= 1 && st >= 1 && sh >= 1 && sw >= 1;
}

int launch_conv_implicit(const __half* in, int B, int T, int H, int W, int C, int kt, int kh, int kw, int st, int sh,
                         int sw, int pt, int ph, int pw, const __half* Wt, GemmParams p, cudaStream_t stream) {
  KVQ_REQUIRE(conv_implicit_supported(C, kt, kh, kw, st, sh, sw), KVQ_ERR_BAD_SHAPE,
              "implicit conv needs C %% 64 == 0 (C=%d)", C);
  KVQ_REQUIRE(p.N % 64 == 0 && p.N <= MAX_BIAS_N && p.ldo % 8 == 0 && (p.resid_h == nullptr || p.ldr % 8 == 0),
              KVQ_ERR_MISALIGNED, "implicit conv: N=%d ldo=%d ldr=%d", p.N, p.ldo, p.ldr);
  KVQ_REQUIRE(p.nvalid == 0 || (p.nvalid % 8 == 0 && p.nvalid <= p.N), KVQ_ERR_BAD_SHAPE, "implicit conv: nvalid=%d",
              p.nvalid);
  const int To = (T + 2 * pt - kt) / st + 1, Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  KVQ_REQUIRE(To > 0 && Ho > 0 && Wo > 0, KVQ_ERR_BAD_SHAPE, "implicit conv: empty output %dx%dx%d", To, Ho, Wo);
  // 128-pixel tile bt x bh x bw (powers of two) with the least padding; ties go to the widest rows
  int best_bw = 0, best_bh = 0, best_bt = 0;
  long long best = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    if (bw * sw > 256) continue;
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      const int bt = 128 / (bw * bh);
      if (bh * sh > 256 || bt * st > 256) continue;
      const long long cover = static_cast<long long>((To + bt - 1) / bt) * bt * ((Ho + bh - 1) / bh) * bh *
                              ((Wo + bw - 1) / bw) * bw;
      if (best < 0 || cover < best) { best = cover; best_bw = bw; best_bh = bh; best_bt = bt; }
    }
  }
  p.conv.on = 1;
  p.conv.C = C; p.conv.kt = kt; p.conv.kh = kh; p.conv.kw = kw;
  p.conv.st = st; p.conv.sh = sh; p.conv.sw = sw; p.conv.pt = pt; p.conv.ph = ph; p.conv.pw = pw;
  p.conv.To = To; p.conv.Ho = Ho; p.conv.Wo = Wo;
  p.conv.bt = best_bt; p.conv.bh = best_bh; p.conv.bw = best_bw;
  p.conv.nt = (To + best_bt - 1) / best_bt; p.conv.nh = (Ho + best_bh - 1) / best_bh;
  p.conv.nw = (Wo + best_bw - 1) / best_bw;
  const long long tiles_m = static_cast<long long>(B) * p.conv.nt * p.conv.nh * p.conv.nw;
  KVQ_REQUIRE(tiles_m * BM < (1LL << 31), KVQ_ERR_BAD_SHAPE, "implicit conv: %lld tiles", tiles_m);
  p.M = static_cast<int>(tiles_m * BM);
  p.K = kt * kh * kw * C;
  CUtensorMap tmA;
  int rc = make_tmap_conv5d(&tmA, in, B, T, H, W, C, best_bt, best_bh, best_bw, st, sh, sw);
  if (rc != 0) return rc;
  // a warp's 32 accumulator rows are an aligned bw_w x bh_w x bt_w sub-block of the tile's pixel block
  const int bw_w = best_bw < 32 ? best_bw : 32;
  const int bh_w = best_bh < 32 / bw_w ? best_bh : 32 / bw_w;
  const int bt_w = 32 / (bw_w * bh_w);
  CUtensorMap tmC;
  rc = make_tmap_out5d(&tmC, p.out, B, To, Ho, Wo, p.nvalid > 0 ? p.nvalid : p.N, p.ldo, bt_w, bh_w, bw_w);
  if (rc != 0) return rc;
  if (p.N % 256 == 0 && tiles_m * (p.N / 256) >= num_sms()) return launch_conv_impl<256>(tmA, tmC, Wt, p.K, p, stream);
  if (p.N % 192 == 0 && tiles_m * (p.N / 192) >= num_sms()) return launch_conv_impl<192>(tmA, tmC, Wt, p.K, p, stream);
  if (p.N % 128 == 0 && tiles_m * (p.N / 128) >= num_sms()) return launch_conv_impl<128>(tmA, tmC, Wt, p.K, p, stream);
  return launch_conv_impl<64>(tmA, tmC, Wt, p.K, p, stream);
}

bool conv_narrow_supported(int C, int cout) {
  static const int off = []() { const char* e = getenv("KVQ_CONV_EXPLICIT"); return e ? atoi(e) : 0; }();
  return off == 0 && (C == 8 || C == 16 || C == 32) && cout >= 8 && cout <= 64 && cout % 8 == 0;
}

int conv_image_kblocks(int C, int taps) { const int tpb = 64 / C; return (taps + tpb - 1) / tpb; }

int launch_conv_narrow(const __half* in, int B, int T, int H, int W, int C, int kt, int kh, int kw, int st, int sh,
                       int sw, int pt, int ph, int pw, const void* Wimg, GemmParams p, cudaStream_t stream) {
  KVQ_REQUIRE(Wimg != nullptr && (C == 8 || C == 16 || C == 32), KVQ_ERR_BAD_SHAPE,
              "narrow conv: C=%d (8, 16 or 32) and a packed weight image are required", C);
  KVQ_REQUIRE(p.N == 64 && p.ldo % 8 == 0 && (p.resid_h == nullptr || p.ldr % 8 == 0), KVQ_ERR_MISALIGNED,
              "narrow conv: N=%d (must be 64) ldo=%d ldr=%d", p.N, p.ldo, p.ldr);
  KVQ_REQUIRE(p.nvalid == 0 || (p.nvalid % 8 == 0 && p.nvalid <= p.N), KVQ_ERR_BAD_SHAPE, "narrow conv: nvalid=%d",
              p.nvalid);
  const int To = (T + 2 * pt - kt) / st + 1, Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  KVQ_REQUIRE(To > 0 && Ho > 0 && Wo > 0, KVQ_ERR_BAD_SHAPE, "narrow conv: empty output %dx%dx%d", To, Ho, Wo);
  int best_bw = 0, best_bh = 0, best_bt = 0;
  long long best = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    if (bw * sw > 256) continue;
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      const int bt = 128 / (bw * bh);
      if (bh * sh > 256 || bt * st > 256) continue;
      const long long cover = static_cast<long long>((To + bt - 1) / bt) * bt * ((Ho + bh - 1) / bh) * bh *
                              ((Wo + bw - 1) / bw) * bw;
      if (best < 0 || cover < best) { best = cover; best_bw = bw; best_bh = bh; best_bt = bt; }
    }
  }
  p.conv.on = 1;
  p.conv.C = C; p.conv.kt = kt; p.conv.kh = kh; p.conv.kw = kw;
  p.conv.st = st; p.conv.sh = sh; p.conv.sw = sw; p.conv.pt = pt; p.conv.ph = ph; p.conv.pw = pw;
  p.conv.To = To; p.conv.Ho = Ho; p.conv.Wo = Wo;
  p.conv.bt = best_bt; p.conv.bh = best_bh; p.conv.bw = best_bw;
  p.conv.nt = (To + best_bt - 1) / best_bt; p.conv.nh = (Ho + best_bh - 1) / best_bh;
  p.conv.nw = (Wo + best_bw - 1) / best_bw;
  p.conv.csub = C; p.conv.taps = kt * kh * kw; p.conv.wimg = Wimg;
  const long long tiles_m = static_cast<long long>(B) * p.conv.nt * p.conv.nh * p.conv.nw;
  KVQ_REQUIRE(tiles_m * BM < (1LL << 31), KVQ_ERR_BAD_SHAPE, "narrow conv: %lld tiles", tiles_m);
  p.M = static_cast<int>(tiles_m * BM);
  p.K = 64 * conv_image_kblocks(C, p.conv.taps);
  CUtensorMap tmA;
  int rc = make_tmap_conv5d(&tmA, in, B, T, H, W, C, best_bt, best_bh, best_bw, st, sh, sw);
  if (rc != 0) return rc;
  const int bw_w = best_bw < 32 ? best_bw : 32;
  const int bh_w = best_bh < 32 / bw_w ? best_bh : 32 / bw_w;
  const int bt_w = 32 / (bw_w * bh_w);
  CUtensorMap tmC;
  rc = make_tmap_out5d(&tmC, p.out, B, To, Ho, Wo, p.nvalid > 0 ? p.nvalid : p.N, p.ldo, bt_w, bh_w, bw_w);
  if (rc != 0) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    KVQ_CUDA(cudaFuncSetAttribute(gemm_kernel<64, EPI_CONV_F16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg<64, 1>::SMEM));
    attr_set = true;
  }
  const int grid = tiles_m < num_sms() ? static_cast<int>(tiles_m) : num_sms();
  count_launch();
  return launch_pdl(gemm_kernel<64, EPI_CONV_F16, 1>, dim3(grid), dim3(GEMM_THREADS), Cfg<64, 1>::SMEM, stream, tmA, tmA,
                    tmC, p);
}

int launch_gemm(int epi, const __half* A, int lda, const __half* B, int ldb, const GemmParams& p,
                cudaStream_t stream) {
  KVQ_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, KVQ_ERR_BAD_SHAPE, "gemm: empty problem %dx%dx%d", p.M, p.N, p.K);
  KVQ_REQUIRE(p.N <= MAX_BIAS_N, KVQ_ERR_BAD_SHAPE, "gemm: N=%d exceeds the %d-entry bias stage", p.N, MAX_BIAS_N);
  KVQ_REQUIRE(p.K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, KVQ_ERR_MISALIGNED,
              "gemm: K=%d lda=%d ldb=%d must be multiples of 8 halfs (16 B TMA rows)", p.K, lda, ldb);
  switch (epi) {
    case EPI_GELU_F16:
      KVQ_REQUIRE(p.ldo % 8 == 0, KVQ_ERR_MISALIGNED, "gemm: fp16 output stride %d not 16 B aligned", p.ldo);
      return launch_bn<EPI_GELU_F16>(A, lda, B, ldb, p, stream);
    case EPI_STORE_F16:
      KVQ_REQUIRE(p.ldo % 8 == 0, KVQ_ERR_MISALIGNED, "gemm: fp16 output stride %d not 16 B aligned", p.ldo);
      return launch_bn<EPI_STORE_F16>(A, lda, B, ldb, p, stream);
    case EPI_CONV_F16:
      KVQ_REQUIRE(p.ldo % 8 == 0 && (p.resid_h == nullptr || p.ldr % 8 == 0), KVQ_ERR_MISALIGNED,
                  "gemm(conv): fp16 output / residual strides %d / %d not 16 B aligned", p.ldo, p.ldr);
      KVQ_REQUIRE(p.nvalid == 0 || (p.nvalid % 8 == 0 && p.nvalid <= p.N), KVQ_ERR_BAD_SHAPE,
                  "gemm(conv): nvalid=%d must be a multiple of 8 and <= N=%d", p.nvalid, p.N);
      {
        // widest tile that still leaves at least one tile per SM: a 128 x 64 tile pulls 3 operand bytes per 128 FLOP
        // out of L2 (L2-bound at ~1/3 of the tensor peak), a 128 x 256 tile 1.5
        const long long mt = (p.M + BM - 1) / BM;
        if (p.N % 256 == 0 && mt * (p.N / 256) >= num_sms()) return launch_impl<256, EPI_CONV_F16>(A, lda, B, ldb, p, stream);
        if (p.N % 192 == 0 && mt * (p.N / 192) >= num_sms()) return launch_impl<192, EPI_CONV_F16>(A, lda, B, ldb, p, stream);
        if (p.N % 128 == 0 && mt * (p.N / 128) >= num_sms()) return launch_impl<128, EPI_CONV_F16>(A, lda, B, ldb, p, stream);
      }
      if (p.N % 64 == 0) return launch_impl<64, EPI_CONV_F16>(A, lda, B, ldb, p, stream);
      set_error("gemm(conv): N=%d (padded output channels) must be a multiple of 64", p.N);
      return KVQ_ERR_BAD_SHAPE;
    case EPI_RESID_F32:
      KVQ_REQUIRE(p.ldo % 4 == 0, KVQ_ERR_MISALIGNED, "gemm: fp32 output stride %d not 16 B aligned", p.ldo);
      return launch_bn<EPI_RESID_F32>(A, lda, B, ldb, p, stream);
    case EPI_QKV_IMG:
      KVQ_REQUIRE(p.N == 3 * p.C && p.C % 32 == 0 && p.bias != nullptr, KVQ_ERR_BAD_SHAPE,
                  "gemm(qkv): N=%d C=%d (need N == 3C, C %% 32 == 0, bias)", p.N, p.C);
      if (use_bstat(p)) return launch_impl<128, EPI_QKV_IMG, 1, true>(A, lda, B, ldb, p, stream);
      if (p.N % 192 == 0 && use_big_tiles(p)) return launch_impl<192, EPI_QKV_IMG, 2>(A, lda, B, ldb, p, stream);
      if (p.N % 192 == 0) return launch_impl<192, EPI_QKV_IMG>(A, lda, B, ldb, p, stream);
      if (p.N % 96 == 0) return launch_impl<96, EPI_QKV_IMG>(A, lda, B, ldb, p, stream);
      set_error("gemm(qkv): N=%d is not a multiple of 96", p.N);
      return KVQ_ERR_BAD_SHAPE;
    case EPI_LN_F32:
      KVQ_REQUIRE(p.N == 96 && p.bias && p.gamma && p.beta, KVQ_ERR_BAD_SHAPE,
                  "gemm(ln): the fused LayerNorm epilogue needs N == 96 (got %d) and bias/gamma/beta", p.N);
      return launch_impl<96, EPI_LN_F32>(A, lda, B, ldb, p, stream);
    case EPI_HEAD:
      KVQ_REQUIRE(p.N == 64 && p.bias && p.w2 && p.rowscore, KVQ_ERR_BAD_SHAPE,
                  "gemm(head): needs N == 64 hidden channels (got %d)", p.N);
      return launch_impl<64, EPI_HEAD>(A, lda, B, ldb, p, stream);
    default:
      set_error("gemm: unknown epilogue %d", epi);
      return KVQ_ERR_BAD_SHAPE;
  }
}

}  // namespace kvq
