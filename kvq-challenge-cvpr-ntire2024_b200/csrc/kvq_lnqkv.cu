// LN1 + roll + window_partition + qkv Linear + bias + q*scale + head split + attention-image layout as ONE kernel for
// the C = 96 stage (SwinTransformerBlock3D.forward_part1 :407-449 + WindowAttention3D.forward :253-261).
//
// The LayerNorm output is an A tile that exists only in shared memory: eight producer warps read the fp32 residual
// rows of a 128-row tile in window order (8 lanes per row, three float4 each), normalise them (fp32 two-pass statistics
// like nn.LayerNorm) and write fp16 into the 128B-swizzled K-major tile the MMA reads.  The 288 x 96 qkv weight stays
// resident in shared memory; a tile is three M128 x N96 x K96 products (q, k, v) whose accumulators go through a ring
// of five 96-column TMEM slots, so the MMAs of the next tile run under the epilogue of this one.  The epilogue is the
// third-generation image scatter of gemm_kernel<EPI_QKV_IMG> (rows d-fastest, key slot = query row).
// Against ln_window_scatter_kernel + gemm_kernel<96, EPI_QKV_IMG> this removes the fp16 LN output (77 MB written and
// read back per stage-0 block at batch 8) and one launch.
//
//   warps 0..7   : LN producers          warp 16 : lane 0 issues the MMAs (and loads the weight once)
//   warps 8..15  : epilogue, two sets of four warps taking alternate accumulator slots; thread = tile row
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int LQ_THREADS = 544;
constexpr int LQ_M = 128;
constexpr int LQ_C = 96;
constexpr int LQ_N = 288;
constexpr int LQ_SLOTS = 5;                             // accumulator ring: 5 x 96 TMEM columns
constexpr int LQ_A_BYTES = 2 * LQ_M * 128;              // two 64-wide K blocks (the second one half used)
constexpr int LQ_W_KB = LQ_N * 128;                     // one 64-wide K block of the weight: 288 rows x 128 B
constexpr int LQ_OFF_A = 0;
constexpr int LQ_OFF_W = LQ_OFF_A + 2 * LQ_A_BYTES;
constexpr int LQ_OFF_PAR = LQ_OFF_W + 2 * LQ_W_KB;      // qkv bias [288] | gamma [96] | beta [96]
constexpr int LQ_OFF_BAR = LQ_OFF_PAR + (LQ_N + 2 * LQ_C) * 4;
constexpr int LQ_SMEM = LQ_OFF_BAR + 256 + 1024;

__device__ __forceinline__ void st_global_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ float sum8(float v) {   // over the 8 lanes that share a row
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__global__ void __launch_bounds__(LQ_THREADS, 1)
ln_qkv96_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ x, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, const float* __restrict__ qkv_b, __half* __restrict__ img,
                int rows, int heads, float qscale, WinGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* spar = reinterpret_cast<float*>(smem + LQ_OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LQ_OFF_BAR);
  uint64_t* a_full = bars + 0;     // [2]
  uint64_t* a_empty = bars + 2;    // [2]
  uint64_t* o_full = bars + 4;     // [5]
  uint64_t* o_empty = bars + 9;    // [5]
  uint64_t* w_full = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

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
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (warp == 16) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < LQ_N + 2 * LQ_C; i += LQ_THREADS)
    spar[i] = i < LQ_N ? __ldg(qkv_b + i) : i < LQ_N + LQ_C ? __ldg(gamma + i - LQ_N) : __ldg(beta + i - LQ_N - LQ_C);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int rows_in = g.nW * g.N;   // window-order rows per clip

  if (warp < 8) {
    // ============================ LN producers ============================
    // a warp normalises 4 rows at a time (8 lanes x 3 float4 per row), 4 passes cover its 16 rows of the tile
    const int sub = lane & 7, rq = lane >> 3;
    const float* sgam = spar + LQ_N;
    const float* sbet = spar + LQ_N + LQ_C;
    // Software pipeline over half tiles (2 passes = 8 rows per warp each): the loads of the next half are in flight
    // while this one is normalised and stored, so the SM always has gather reads outstanding
    auto load_half = [&](int tile, int hb, float4 (&v)[2][3], bool (&pad)[2]) {
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const int rt = warp * 16 + (hb * 2 + ps) * 4 + rq;            // row inside the tile
        int row = tile * LQ_M + rt;
        if (row >= rows) row = rows - 1;                              // (its output is never read)
        const int b = fdiv_i(row, rows_in, g.r_rows);
        const int src = win_row_to_src(g, row - b * rows_in);
        pad[ps] = src < 0;                                            // padded slot: zeros AFTER the norm (:416-424)
        const float* sp = x + (static_cast<size_t>(b) * g.tokens + (src < 0 ? 0 : src)) * LQ_C;
#pragma unroll
        for (int k = 0; k < 3; ++k) v[ps][k] = __ldg(reinterpret_cast<const float4*>(sp) + sub + 8 * k);
      }
    };
    auto store_half = [&](int buf, int hb, const float4 (&v)[2][3], const bool (&pad)[2]) {
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += (v[ps][k].x + v[ps][k].y) + (v[ps][k].z + v[ps][k].w);
        const float mean = sum8(s) * (1.0f / LQ_C);
        float qq = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float a = v[ps][k].x - mean, b2 = v[ps][k].y - mean, c = v[ps][k].z - mean, d = v[ps][k].w - mean;
          qq += (a * a + b2 * b2) + (c * c + d * d);
        }
        const float rstd = rsqrtf(sum8(qq) * (1.0f / LQ_C) + eps);
        const int rt = warp * 16 + (hb * 2 + ps) * 4 + rq;
        uint8_t* arow = smem + LQ_OFF_A + buf * LQ_A_BYTES + rt * 128;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int ch = sub + 8 * k;
          const float4 gm = *reinterpret_cast<const float4*>(sgam + 4 * ch);
          const float4 bt = *reinterpret_cast<const float4*>(sbet + 4 * ch);
          const float y0 = (v[ps][k].x - mean) * rstd * gm.x + bt.x;
          const float y1 = (v[ps][k].y - mean) * rstd * gm.y + bt.y;
          const float y2 = (v[ps][k].z - mean) * rstd * gm.z + bt.z;
          const float y3 = (v[ps][k].w - mean) * rstd * gm.w + bt.w;
          const uint2 hv = pad[ps] ? make_uint2(0u, 0u) : make_uint2(pack_half2(y0, y1), pack_half2(y2, y3));
          // K index of float4 chunk ch is 4 ch: K block ch / 16, 16-byte chunk (ch % 16) / 2 (XOR-swizzled by the row)
          *reinterpret_cast<uint2*>(arow + (ch >> 4) * (LQ_M * 128) + ((((ch & 15) >> 1) ^ (rt & 7)) << 4) + (ch & 1) * 8) = hv;
        }
      }
    };
    uint32_t n_t = 0;
    float4 va[2][3], vb[2][3];
    bool pa[2], pb[2];
    if (static_cast<int>(blockIdx.x) < tiles) load_half(blockIdx.x, 0, va, pa);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
      const int buf = n_t & 1;
      load_half(tile, 1, vb, pb);
      if (lane == 0) mbar_wait(&a_empty[buf], ((n_t >> 1) & 1) ^ 1);
      __syncwarp();
      store_half(buf, 0, va, pa);
      const int next = tile + gridDim.x;
      if (next < tiles) load_half(next, 0, va, pa);
      store_half(buf, 1, vb, pb);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
    }
  } else if (warp == 16) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(LQ_M, LQ_C, 0, 0);
      mbar_expect_tx(w_full, 2 * LQ_W_KB);
      for (int kb = 0; kb < 2; ++kb)
        for (int nb = 0; nb < 3; ++nb)
          tma_load_2d(smem + LQ_OFF_W + kb * LQ_W_KB + nb * (LQ_C * 128), &tmW, w_full, kb * 64, nb * LQ_C);
      mbar_wait(w_full, 0);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t dA0 = umma_smem_desc(sbase + LQ_OFF_A, 16, 1024, UMMA_SW_128);
      const uint64_t dW0 = umma_smem_desc(sbase + LQ_OFF_W, 16, 1024, UMMA_SW_128);
      uint32_t n_t = 0, n_o = 0;     // tiles, accumulator slots used so far
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
        const uint32_t buf = n_t & 1;
        mbar_wait(&a_full[buf], (n_t >> 1) & 1);
        const uint64_t da = dA0 + static_cast<uint64_t>(buf * (LQ_A_BYTES >> 4));
        for (int nb = 0; nb < 3; ++nb, ++n_o) {
          const uint32_t slot = n_o % LQ_SLOTS;
          mbar_wait(&o_empty[slot], ((n_o / LQ_SLOTS) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tD = tmem_base + slot * LQ_C;
          const uint64_t dw = dW0 + static_cast<uint64_t>((nb * (LQ_C * 128)) >> 4);
#pragma unroll
          for (int ks = 0; ks < 6; ++ks) {
            const int kb = ks >> 2, k = ks & 3;
            umma_f16_ss(tD, da + static_cast<uint64_t>(((kb * LQ_M * 128) >> 4) + 2 * k),
                        dw + static_cast<uint64_t>(((kb * LQ_W_KB) >> 4) + 2 * k), idesc, ks > 0 ? 1u : 0u);
          }
          if (nb == 2) umma_commit(&a_empty[buf]);
          umma_commit(&o_full[slot]);
        }
      }
    }
  } else {
    // ============================ epilogue: bias, q*scale, fp16, image scatter ============================
    const int ew = warp - 8, q = warp & 3, set = ew >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const size_t unit_bytes = ATT3_UNIT_BYTES;
    uint32_t n_o = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int row = tile * LQ_M + q * 32 + lane;
      const bool row_ok = row < rows;
      const int win_g = row / g.N;
      const int i = row - win_g * g.N;           // row inside the window = image row (d-fastest order)
      for (int nb = 0; nb < 3; ++nb, ++n_o) {
        if ((n_o & 1) != static_cast<uint32_t>(set)) continue;      // the other warp set owns this slot
        const uint32_t slot = n_o % LQ_SLOTS;
        if (lane == 0) mbar_wait(&o_full[slot], (n_o / LQ_SLOTS) & 1);
        __syncwarp();
        tc_fence_after();
        const float sc = nb == 0 ? qscale : 1.0f;
        uint8_t* dst0 = reinterpret_cast<uint8_t*>(img) + static_cast<size_t>(win_g) * heads * unit_bytes + nb * ATT3_KV_BYTES;
#pragma unroll
        for (int hd = 0; hd < 3; ++hd) {         // 32 columns = one head
          uint32_t r[32];
          tmem_ld_x32(lane_addr + slot * LQ_C + hd * 32, r);
          tmem_wait_ld();
          uint32_t hh[16];
          const float* bj = spar + nb * LQ_C + hd * 32;
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
            if (nb != 0 && i == g.N - 1) {
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

}  // namespace

int launch_ln_qkv96(const float* x, const float* gamma, const float* beta, float eps, const __half* qkv_w,
                    const float* qkv_b, __half* img, int B, int heads, float qscale, const WinGeom& g,
                    cudaStream_t stream) {
  KVQ_REQUIRE(g.dfast == 1 && g.N == 392 && heads == 3, KVQ_ERR_BAD_SHAPE,
              "ln_qkv96: built for the full (8,7,7) window in d-fastest row order with 3 heads (N=%d heads=%d)", g.N, heads);
  const long long rows = static_cast<long long>(B) * g.nW * g.N;
  KVQ_REQUIRE(rows > 0 && rows < (1ll << 31), KVQ_ERR_BAD_SHAPE, "ln_qkv96: %lld rows", rows);
  static bool attr = false;
  if (!attr) {
    KVQ_CUDA(cudaFuncSetAttribute(ln_qkv96_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LQ_SMEM));
    attr = true;
  }
  CUtensorMap tmW;
  const int rc = make_tmap_2d(&tmW, qkv_w, LQ_N, LQ_C, LQ_C * 2, LQ_C, 64, 2, 128);
  if (rc != 0) return rc;
  const int tiles = static_cast<int>((rows + LQ_M - 1) / LQ_M);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  count_launch();
  return launch_pdl(ln_qkv96_kernel, dim3(grid), dim3(LQ_THREADS), LQ_SMEM, stream, tmW, x, gamma, beta, eps, qkv_b, img,
                    static_cast<int>(rows), heads, qscale, g);
}

}  // namespace kvq
