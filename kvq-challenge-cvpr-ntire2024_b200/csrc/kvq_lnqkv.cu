// LN1 + roll + window_partition + qkv Linear + bias + q*scale + head split + attention-image layout as ONE kernel
// (SwinTransformerBlock3D.forward_part1 :407-449 + WindowAttention3D.forward :253-261), C = 96 / 192 / 384.
//
// The LayerNorm output is an A tile that exists only in shared memory: eight producer warps read the fp32 residual
// rows of a 128-row tile in window order (8 / 16 / 32 lanes per row, three float4 each), normalise them (fp32 two-pass
// statistics like nn.LayerNorm) and write fp16 into the 128B-swizzled K-major tile the MMA reads.  A tile is 3C/96
// products M128 x N96 x K=C (N block j = 96 consecutive qkv channels = three heads of q, k or v) whose accumulators go
// through a ring of five 96-column TMEM slots, so the MMAs run under the epilogue of earlier blocks.  The qkv weight is
// resident in shared memory at C = 96 (55 KB) and streamed by TMA as (N block, 64-wide K block) slices of 12 KB through
// an 8-deep ring otherwise.  The epilogue is the third-generation image scatter of gemm_kernel<EPI_QKV_IMG> (rows
// d-fastest, key slot = query row).
// Against ln_window_scatter_kernel + gemm_kernel<*, EPI_QKV_IMG> this removes the fp16 LN output (written and read
// back: 154 MB per stage-0 block at batch 8) and one launch per block.
//
//   warps 0..7   : LN producers          warp 16 : lane 0 issues the MMAs        warp 17 : lane 0 streams the weight
//   warps 8..15  : epilogue, two sets of four warps taking alternate accumulator slots; thread = tile row
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int LQ_THREADS = 576;
constexpr int LQ_M = 128;
constexpr int LQ_NBW = 96;                              // N block: 96 qkv channels = 3 heads
constexpr int LQ_SLOTS = 5;                             // accumulator ring: 5 x 96 TMEM columns

template <int C>
struct LqCfg {
  static constexpr int KB = (C + 63) / 64;              // 64-wide K blocks of the A tile
  static constexpr int KSTEPS = C / 16;
  static constexpr int NB = 3 * C / LQ_NBW;             // N blocks per tile
  static constexpr int NBH = C / LQ_NBW;                // N blocks per q / k / v
  static constexpr int A_BYTES = KB * LQ_M * 128;
  static constexpr int A_ST = C <= 192 ? 2 : 1;
  static constexpr bool RESIDENT = C == 96;             // whole weight in shared memory
  static constexpr int W_SLICE = LQ_NBW * 128;          // one (N block, K block) slice: 96 rows x 128 B
  static constexpr int W_ST = 8;
  static constexpr int W_BYTES = RESIDENT ? NB * KB * W_SLICE : W_ST * W_SLICE;
  static constexpr int LPR = C / 12;                    // lanes per row: 8 / 16 / 32 (three float4 per lane)
  static constexpr int RPP = 32 / LPR;                  // rows per pass of a warp: 4 / 2 / 1
  static constexpr int PASSES = 16 / RPP;               // a warp owns 16 rows of the tile
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = OFF_A + A_ST * A_BYTES;
  static constexpr int OFF_PAR = OFF_W + W_BYTES;       // qkv bias [3C] | gamma [C] | beta [C]
  static constexpr int OFF_BAR = OFF_PAR + 5 * C * 4;
  static constexpr int SMEM = OFF_BAR + 512 + 1024;
  static_assert((C == 96 || C == 192 || C == 384) && SMEM <= 227 * 1024, "unsupported channel count");
};

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int LPR>
__device__ __forceinline__ float row_sum(float v) {   // over the LPR lanes that share a row
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int C>
__global__ void __launch_bounds__(LQ_THREADS, 1)
ln_qkv_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, const float* __restrict__ qkv_b, __half* __restrict__ img,
              int rows, int heads, float qscale, WinGeom g) {
  using Cfg = LqCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* spar = reinterpret_cast<float*>(smem + Cfg::OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* a_full = bars + 0;     // [2]
  uint64_t* a_empty = bars + 2;    // [2]
  uint64_t* o_full = bars + 4;     // [5]
  uint64_t* o_empty = bars + 9;    // [5]
  uint64_t* w_full = bars + 14;    // [8] (resident weight: only [0])
  uint64_t* w_empty = bars + 22;   // [8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (rows + LQ_M - 1) / LQ_M;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 8);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < LQ_SLOTS; ++i) {
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    for (int i = 0; i < Cfg::W_ST; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 16) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 5 * C; i += LQ_THREADS)
    spar[i] = i < 3 * C ? __ldg(qkv_b + i) : i < 4 * C ? __ldg(gamma + i - 3 * C) : __ldg(beta + i - 4 * C);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int rows_in = g.nW * g.N;   // window-order rows per clip

  if (warp < 8) {
    // ============================ LN producers ============================
    // A warp owns 16 rows of the tile and normalises RPP rows per pass (LPR lanes x 3 float4 per row).  Software pipeline
    // over batches of two passes: the loads of the next batch are in flight while this one is normalised and stored.
    constexpr int LPR = Cfg::LPR, RPP = Cfg::RPP, HP = 2, NBATCH = Cfg::PASSES / 2;    // passes per batch, batches per tile
    const int sub = lane % LPR, rq = lane / LPR;
    const float* sgam = spar + 3 * C;
    const float* sbet = spar + 4 * C;
    auto load_half = [&](int tile, int hb, float4 (&v)[HP][3], bool (&pad)[HP]) {
#pragma unroll
      for (int ps = 0; ps < HP; ++ps) {
        const int rt = warp * 16 + (hb * HP + ps) * RPP + rq;         // row inside the tile
        int row = tile * LQ_M + rt;
        if (row >= rows) row = rows - 1;                              // (its output is never read)
        const int b = fdiv_i(row, rows_in, g.r_rows);
        const int src = win_row_to_src(g, row - b * rows_in);
        pad[ps] = src < 0;                                            // padded slot: zeros AFTER the norm (:416-424)
        const float* sp = x + (static_cast<size_t>(b) * g.tokens + (src < 0 ? 0 : src)) * C;
#pragma unroll
        for (int k = 0; k < 3; ++k) v[ps][k] = __ldg(reinterpret_cast<const float4*>(sp) + sub + LPR * k);
      }
    };
    auto store_half = [&](int buf, int hb, const float4 (&v)[HP][3], const bool (&pad)[HP]) {
#pragma unroll
      for (int ps = 0; ps < HP; ++ps) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += (v[ps][k].x + v[ps][k].y) + (v[ps][k].z + v[ps][k].w);
        const float mean = row_sum<LPR>(s) * (1.0f / C);
        float qq = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float a = v[ps][k].x - mean, b2 = v[ps][k].y - mean, c = v[ps][k].z - mean, d = v[ps][k].w - mean;
          qq += (a * a + b2 * b2) + (c * c + d * d);
        }
        const float rstd = rsqrtf(row_sum<LPR>(qq) * (1.0f / C) + eps);
        const int rt = warp * 16 + (hb * HP + ps) * RPP + rq;
        uint8_t* arow = smem + Cfg::OFF_A + buf * Cfg::A_BYTES + rt * 128;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int ch = sub + LPR * k;                               // float4 chunk: K index 4 ch
          const float4 gm = *reinterpret_cast<const float4*>(sgam + 4 * ch);
          const float4 bt = *reinterpret_cast<const float4*>(sbet + 4 * ch);
          const float y0 = (v[ps][k].x - mean) * rstd * gm.x + bt.x;
          const float y1 = (v[ps][k].y - mean) * rstd * gm.y + bt.y;
          const float y2 = (v[ps][k].z - mean) * rstd * gm.z + bt.z;
          const float y3 = (v[ps][k].w - mean) * rstd * gm.w + bt.w;
          const uint2 hv = pad[ps] ? make_uint2(0u, 0u) : make_uint2(pack_half2(y0, y1), pack_half2(y2, y3));
          // K block ch / 16, 16-byte chunk (ch % 16) / 2 XOR-swizzled by the row, half ch & 1
          *reinterpret_cast<uint2*>(arow + (ch >> 4) * (LQ_M * 128) + ((((ch & 15) >> 1) ^ (rt & 7)) << 4) + (ch & 1) * 8) = hv;
        }
      }
    };
    uint32_t n_t = 0;
    float4 va[HP][3], vb[HP][3];
    bool pa[HP], pb[HP];
    if (static_cast<int>(blockIdx.x) < tiles) load_half(blockIdx.x, 0, va, pa);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
      const int buf = n_t % Cfg::A_ST;
      const int next = tile + gridDim.x;
#pragma unroll
      for (int bt = 0; bt < NBATCH; bt += 2) {
        load_half(tile, bt + 1, vb, pb);
        if (bt == 0) {
          if (lane == 0) mbar_wait(&a_empty[buf], ((n_t / Cfg::A_ST) & 1) ^ 1);
          __syncwarp();
        }
        store_half(buf, bt, va, pa);
        if (bt + 2 < NBATCH) load_half(tile, bt + 2, va, pa);
        else if (next < tiles) load_half(next, 0, va, pa);
        store_half(buf, bt + 1, vb, pb);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
    }
  } else if (warp == 17) {
    // ============================ weight loader ============================
    if (lane == 0) {
      if constexpr (Cfg::RESIDENT) {
        mbar_expect_tx(&w_full[0], Cfg::W_BYTES);
        for (int nb = 0; nb < Cfg::NB; ++nb)
          for (int kb = 0; kb < Cfg::KB; ++kb)
            tma_load_2d(smem + Cfg::OFF_W + (nb * Cfg::KB + kb) * Cfg::W_SLICE, &tmW, &w_full[0], kb * 64, nb * LQ_NBW);
      } else {
        uint32_t n_w = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
          for (int nb = 0; nb < Cfg::NB; ++nb)
            for (int kb = 0; kb < Cfg::KB; ++kb, ++n_w) {
              const uint32_t st = n_w % Cfg::W_ST;
              mbar_wait(&w_empty[st], ((n_w / Cfg::W_ST) & 1) ^ 1);
              mbar_expect_tx(&w_full[st], Cfg::W_SLICE);
              tma_load_2d(smem + Cfg::OFF_W + st * Cfg::W_SLICE, &tmW, &w_full[st], kb * 64, nb * LQ_NBW);
            }
        }
      }
    }
  } else if (warp == 16) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(LQ_M, LQ_NBW, 0, 0);
      if constexpr (Cfg::RESIDENT) mbar_wait(&w_full[0], 0);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t dA0 = umma_smem_desc(sbase + Cfg::OFF_A, 16, 1024, UMMA_SW_128);
      const uint64_t dW0 = umma_smem_desc(sbase + Cfg::OFF_W, 16, 1024, UMMA_SW_128);
      uint32_t n_t = 0, n_o = 0, n_w = 0;     // tiles, accumulator slots, weight slices used so far
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
        const uint32_t buf = n_t % Cfg::A_ST;
        mbar_wait(&a_full[buf], (n_t / Cfg::A_ST) & 1);
        const uint64_t da = dA0 + static_cast<uint64_t>(buf * (Cfg::A_BYTES >> 4));
        for (int nb = 0; nb < Cfg::NB; ++nb, ++n_o) {
          const uint32_t slot = n_o % LQ_SLOTS;
          mbar_wait(&o_empty[slot], ((n_o / LQ_SLOTS) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tD = tmem_base + slot * LQ_NBW;
#pragma unroll
          for (int kb = 0; kb < Cfg::KB; ++kb) {
            uint64_t dw;
            if constexpr (Cfg::RESIDENT) {
              dw = dW0 + static_cast<uint64_t>(((nb * Cfg::KB + kb) * Cfg::W_SLICE) >> 4);
            } else {
              const uint32_t st = n_w % Cfg::W_ST;
              mbar_wait(&w_full[st], (n_w / Cfg::W_ST) & 1);
              tc_fence_after();
              dw = dW0 + static_cast<uint64_t>((st * Cfg::W_SLICE) >> 4);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kb * 4 + k < Cfg::KSTEPS)
                umma_f16_ss(tD, da + static_cast<uint64_t>(((kb * LQ_M * 128) >> 4) + 2 * k), dw + static_cast<uint64_t>(2 * k),
                            idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            if constexpr (!Cfg::RESIDENT) {
              umma_commit(&w_empty[n_w % Cfg::W_ST]);
              ++n_w;
            }
          }
          if (nb == Cfg::NB - 1) umma_commit(&a_empty[buf]);
          umma_commit(&o_full[slot]);
        }
      }
    }
  } else if (warp < 16) {
    // ============================ epilogue: bias, q*scale, fp16, image scatter ============================
    const int ew = warp - 8, q = warp & 3, set = ew >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const size_t unit_bytes = ATT3_UNIT_BYTES;
    uint32_t n_o = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int row = tile * LQ_M + q * 32 + lane;
      const bool row_ok = row < rows;
      const int win_g = fdiv_i(row, g.N, g.r_N);
      const int i = row - win_g * g.N;           // row inside the window = image row (d-fastest order)
      for (int nb = 0; nb < Cfg::NB; ++nb, ++n_o) {
        if ((n_o & 1) != static_cast<uint32_t>(set)) continue;      // the other warp set owns this slot
        const uint32_t slot = n_o % LQ_SLOTS;
        if (lane == 0) mbar_wait(&o_full[slot], (n_o / LQ_SLOTS) & 1);
        __syncwarp();
        tc_fence_after();
        const int which = nb / Cfg::NBH;                             // 0 q, 1 k, 2 v
        const int head0 = (nb - which * Cfg::NBH) * 3;
        const float sc = which == 0 ? qscale : 1.0f;
        uint8_t* dst0 = reinterpret_cast<uint8_t*>(img) + (static_cast<size_t>(win_g) * heads + head0) * unit_bytes +
                        which * ATT3_KV_BYTES;
#pragma unroll
        for (int hd = 0; hd < 3; ++hd) {         // 32 columns = one head
          uint32_t r[32];
          tmem_ld_x32(lane_addr + slot * LQ_NBW + hd * 32, r);
          tmem_wait_ld();
          uint32_t hh[16];
          const float* bj = spar + nb * LQ_NBW + hd * 32;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 bb = *reinterpret_cast<const float4*>(bj + 4 * k);
            hh[2 * k] = pack_half2((__uint_as_float(r[4 * k]) + bb.x) * sc, (__uint_as_float(r[4 * k + 1]) + bb.y) * sc);
            hh[2 * k + 1] = pack_half2((__uint_as_float(r[4 * k + 2]) + bb.z) * sc, (__uint_as_float(r[4 * k + 3]) + bb.w) * sc);
          }
          if (row_ok) {
            uint8_t* dst = dst0 + static_cast<size_t>(hd) * unit_bytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) st_global_v4(dst + att_img_offset(i, j), hh[4 * j], hh[4 * j + 1], hh[4 * j + 2], hh[4 * j + 3]);
            // the last token of the window also zeroes the padded K / V key slots behind it (P is 0 there, but
            // 0 * NaN from uninitialised workspace would poison PV)
            if (which != 0 && i == g.N - 1) {
              for (int rz = g.N; rz < ATT3_KV_ROWS; ++rz) {
#pragma unroll
                for (int j = 0; j < 4; ++j) st_global_v4(dst + att_img_offset(rz, j), 0u, 0u, 0u, 0u);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[slot]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int C>
int launch_ln_qkv_impl(const float* x, const float* gamma, const float* beta, float eps, const __half* qkv_w,
                       const float* qkv_b, __half* img, long long rows, int heads, float qscale, const WinGeom& g,
                       cudaStream_t stream) {
  using Cfg = LqCfg<C>;
  static bool attr = false;
  if (!attr) {
    KVQ_CUDA(cudaFuncSetAttribute(ln_qkv_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  CUtensorMap tmW;
  const int rc = make_tmap_2d(&tmW, qkv_w, 3 * C, C, static_cast<uint64_t>(C) * 2, LQ_NBW, 64, 2, 128);
  if (rc != 0) return rc;
  const int tiles = static_cast<int>((rows + LQ_M - 1) / LQ_M);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  count_launch();
  return launch_pdl(ln_qkv_kernel<C>, dim3(grid), dim3(LQ_THREADS), Cfg::SMEM, stream, tmW, x, gamma, beta, eps, qkv_b, img,
                    static_cast<int>(rows), heads, qscale, g);
}

}  // namespace

bool ln_qkv_supported(int C) { return C == 96 || C == 192 || C == 384; }

int launch_ln_qkv(const float* x, const float* gamma, const float* beta, float eps, const __half* qkv_w,
                  const float* qkv_b, __half* img, int B, int C, int heads, float qscale, const WinGeom& g,
                  cudaStream_t stream) {
  KVQ_REQUIRE(g.dfast == 1 && g.N == 392 && heads * 32 == C, KVQ_ERR_BAD_SHAPE,
              "ln_qkv: built for the full (8,7,7) window in d-fastest row order, head_dim 32 (N=%d heads=%d C=%d)", g.N,
              heads, C);
  const long long rows = static_cast<long long>(B) * g.nW * g.N;
  KVQ_REQUIRE(rows > 0 && rows < (1ll << 31), KVQ_ERR_BAD_SHAPE, "ln_qkv: %lld rows", rows);
  if (C == 96) return launch_ln_qkv_impl<96>(x, gamma, beta, eps, qkv_w, qkv_b, img, rows, heads, qscale, g, stream);
  if (C == 192) return launch_ln_qkv_impl<192>(x, gamma, beta, eps, qkv_w, qkv_b, img, rows, heads, qscale, g, stream);
  if (C == 384) return launch_ln_qkv_impl<384>(x, gamma, beta, eps, qkv_w, qkv_b, img, rows, heads, qscale, g, stream);
  set_error("ln_qkv: C=%d is not built (96, 192, 384)", C);
  return KVQ_ERR_BAD_SHAPE;
}

}  // namespace kvq
