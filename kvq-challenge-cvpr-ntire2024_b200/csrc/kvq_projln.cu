// attention proj Linear + bias + window_reverse + un-roll + shortcut  AND  norm2  as ONE kernel, C = 96 / 192
// (SwinTransformerBlock3D.forward_part1 tail :472-488 + :509 residual, then forward_part2's norm2 :490):
//     x[token] += attn_out[row] W^T + b          (row = window-order row of the token, scattered back through win_row_to_src)
//     a2[token] = LayerNorm(x[token])            fp16, the A operand of the fused MLP
// The proj weight (C x C) is resident in shared memory, attention-output tiles arrive by TMA (double-buffered), one tile
// is C/16 tcgen05.mma M128 x N=C x K16 into one of two TMEM accumulators.  The epilogue warps transpose 32 x 32 accumulator
// chunks through shared memory so that the fp32 residual rows are read and written as whole rows (8 lanes x 16 B), keep
// per-row sum / sum-of-squares while they do, and then re-read the rows they just wrote (L1 / L2 hits) to normalise them.
// Against gemm_kernel<C, EPI_RESID_F32> + ln_rows_kernel this removes a second pass over the fp32 residual stream (154 MB
// read per stage-0 block at batch 8) and one launch per block.
//
//   warps 0..7 : epilogue, two sets of four warps taking alternate tiles; warp q of a set owns TMEM lanes 32q..32q+31
//   warp 8     : lane 0 loads (weight once, A tiles) and issues the MMAs
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int PL_THREADS = 288;
constexpr int PL_M = 128;

template <int C>
struct PlCfg {
  static constexpr int KB = (C + 63) / 64;
  static constexpr int KSTEPS = C / 16;
  static constexpr int NCH = C / 32;                    // 32-column chunks of a row
  static constexpr int A_BYTES = KB * PL_M * 128;
  static constexpr int W_KB = C * 128;                  // one 64-wide K block of the weight: C rows x 128 B
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = OFF_A + 2 * A_BYTES;
  static constexpr int OFF_STAGE = OFF_W + KB * W_KB;   // per epilogue warp: 32 x 36 fp32 transpose tile + 32 row offsets
  static constexpr int WARP_STAGE = 32 * 36 * 4 + 32 * 8;
  static constexpr int OFF_PAR = OFF_STAGE + 8 * WARP_STAGE;   // proj bias | gamma | beta
  static constexpr int OFF_BAR = OFF_PAR + 3 * C * 4;
  static constexpr int SMEM = OFF_BAR + 128 + 1024;
  static constexpr int TMEM_COLS = 2 * C <= 256 ? 256 : 512;
  static_assert((C == 96 || C == 192) && SMEM <= 227 * 1024, "unsupported channel count");
};

template <int C>
__global__ void __launch_bounds__(PL_THREADS, 1)
proj_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
               float* x, __half* __restrict__ a2, int rows, WinGeom g) {
  using Cfg = PlCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* spar = reinterpret_cast<float*>(smem + Cfg::OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* a_full = bars + 0;     // [2]
  uint64_t* a_empty = bars + 2;    // [2]
  uint64_t* o_full = bars + 4;     // [2]
  uint64_t* o_empty = bars + 6;    // [2]
  uint64_t* w_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (rows + PL_M - 1) / PL_M;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 3 * C; i += PL_THREADS)
    spar[i] = i < C ? __ldg(bias + i) : i < 2 * C ? __ldg(gamma + i - C) : __ldg(beta + i - 2 * C);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 8) {
    // ============================ loader + MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(PL_M, C, 0, 0);
      mbar_expect_tx(w_full, Cfg::KB * Cfg::W_KB);
      for (int kb = 0; kb < Cfg::KB; ++kb) tma_load_2d(smem + Cfg::OFF_W + kb * Cfg::W_KB, &tmW, w_full, kb * 64, 0);
      auto load_a = [&](int tile, uint32_t n) {
        const uint32_t buf = n & 1;
        mbar_wait(&a_empty[buf], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&a_full[buf], Cfg::A_BYTES);
        for (int kb = 0; kb < Cfg::KB; ++kb)
          tma_load_2d(smem + Cfg::OFF_A + buf * Cfg::A_BYTES + kb * PL_M * 128, &tmA, &a_full[buf], kb * 64, tile * PL_M);
      };
      const uint32_t sbase = smem_u32(smem);
      const uint64_t dA0 = umma_smem_desc(sbase + Cfg::OFF_A, 16, 1024, UMMA_SW_128);
      const uint64_t dW0 = umma_smem_desc(sbase + Cfg::OFF_W, 16, 1024, UMMA_SW_128);
      uint32_t n_t = 0;
      int tile = blockIdx.x;
      if (tile < tiles) load_a(tile, 0);
      mbar_wait(w_full, 0);
      for (; tile < tiles; tile += gridDim.x, ++n_t) {
        const uint32_t buf = n_t & 1, ph = (n_t >> 1) & 1;
        const int next = tile + gridDim.x;
        if (next < tiles) load_a(next, n_t + 1);               // the next tile streams in under this tile's MMAs
        mbar_wait(&o_empty[buf], ph ^ 1);
        mbar_wait(&a_full[buf], ph);
        tc_fence_after();
        const uint64_t da = dA0 + static_cast<uint64_t>(buf * (Cfg::A_BYTES >> 4));
        const uint32_t tD = tmem_base + buf * C;
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          const int kb = ks >> 2, k = ks & 3;
          umma_f16_ss(tD, da + static_cast<uint64_t>(((kb * PL_M * 128) >> 4) + 2 * k),
                      dW0 + static_cast<uint64_t>(((kb * Cfg::W_KB) >> 4) + 2 * k), idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&o_full[buf]);
      }
    }
  } else {
    // ============================ epilogue ============================
    const int set = warp >> 2, q = warp & 3;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* stile = reinterpret_cast<float*>(smem + Cfg::OFF_STAGE + warp * Cfg::WARP_STAGE);
    long long* srow = reinterpret_cast<long long*>(stile + 32 * 36);
    const int sub = lane & 7, rsel = lane >> 3;
    const int rows_in = g.nW * g.N;
    constexpr int NCH = Cfg::NCH;
    constexpr int PF = NCH < 3 ? NCH : 3;                  // residual chunks in flight (the first PF are fetched before the accumulator wait)
    const float* sgam = spar + C;
    const float* sbet = spar + 2 * C;
    uint32_t n_t = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
      if ((n_t & 1) != static_cast<uint32_t>(set)) continue;         // the other warp set owns this tile
      const uint32_t buf = n_t & 1;
      // window-order row of this lane -> element offset of its token row in x / a2 (-1: padded slot or past the end)
      {
        const int row = tile * PL_M + q * 32 + lane;
        long long off = -1;
        if (row < rows) {
          const int b = fdiv_i(row, rows_in, g.r_rows);
          const int src = win_row_to_src(g, row - b * rows_in);
          if (src >= 0) off = (static_cast<long long>(b) * g.tokens + src) * C;
        }
        __syncwarp();                                       // the previous tile's readers are done with srow / stile
        srow[lane] = off;
        __syncwarp();
      }
      float4 pre[PF][8];
#pragma unroll
      for (int ci = 0; ci < PF; ++ci) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const long long off = srow[it * 4 + rsel];
          pre[ci][it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (off >= 0) pre[ci][it] = *reinterpret_cast<const float4*>(x + off + ci * 32 + 4 * sub);
        }
      }
      if (lane == 0) mbar_wait(&o_full[buf], (n_t >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      float s1[8], s2[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) { s1[it] = 0.f; s2[it] = 0.f; }
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        uint32_t r[32];
        tmem_ld_x32(lane_addr + buf * C + ci * 32, r);
        const float4 bb = *reinterpret_cast<const float4*>(spar + ci * 32 + 4 * sub);
        tmem_wait_ld();
        __syncwarp();                                       // the previous chunk's readers are done with the tile
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<uint4*>(stile + lane * 36 + 4 * k) = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rsel;
          const long long off = srow[rr];
          if (off >= 0) {
            float4 v = *reinterpret_cast<const float4*>(stile + rr * 36 + 4 * sub);
            const float4 rv = pre[ci % PF][it];
            v.x += bb.x + rv.x; v.y += bb.y + rv.y; v.z += bb.z + rv.z; v.w += bb.w + rv.w;
            *reinterpret_cast<float4*>(x + off + ci * 32 + 4 * sub) = v;
            s1[it] += (v.x + v.y) + (v.z + v.w);
            s2[it] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            // the slot is free: fetch the residual of chunk ci + PF into it (PF chunks stay in flight); when the whole
            // row fits the PF slots (C = 96) keep the new value there instead, so norm2 below needs no re-read
            if (ci + PF < NCH) pre[ci % PF][it] = *reinterpret_cast<const float4*>(x + off + (ci + PF) * 32 + 4 * sub);
            else if (NCH <= PF) pre[ci][it] = v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[buf]);            // the accumulator is free: the rest works on global memory
      // row statistics (8 lanes share a row), then norm2 over the rows this thread just wrote: from registers when the row
      // fits (C = 96), else re-read through global memory (same-thread read after write: L1 / L2 hits, but one L2 round
      // trip per chunk -- the reason the C = 192 instantiation is slower than proj GEMM + ln_rows)
      float mean[8], rstd[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        float a = s1[it], b = s2[it];
        a += __shfl_xor_sync(0xffffffffu, a, 4); b += __shfl_xor_sync(0xffffffffu, b, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1);
        mean[it] = a * (1.0f / C);
        const float var = fmaxf(b * (1.0f / C) - mean[it] * mean[it], 0.f);
        rstd[it] = rsqrtf(var + eps);
      }
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const float4 gm = *reinterpret_cast<const float4*>(sgam + ci * 32 + 4 * sub);
        const float4 bt = *reinterpret_cast<const float4*>(sbet + ci * 32 + 4 * sub);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const long long off = srow[it * 4 + rsel];
          if (off >= 0) {
            float4 v;
            if constexpr (NCH <= PF) v = pre[ci][it];
            else v = *reinterpret_cast<const float4*>(x + off + ci * 32 + 4 * sub);
            const float m = mean[it], rs = rstd[it];
            uint2 hv;
            hv.x = pack_half2((v.x - m) * rs * gm.x + bt.x, (v.y - m) * rs * gm.y + bt.y);
            hv.y = pack_half2((v.z - m) * rs * gm.z + bt.z, (v.w - m) * rs * gm.w + bt.w);
            *reinterpret_cast<uint2*>(a2 + off + ci * 32 + 4 * sub) = hv;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int C>
int launch_proj_ln_impl(const __half* attn_out, const __half* w, const float* bias, const float* gamma, const float* beta,
                        float eps, float* x, __half* a2, long long rows, const WinGeom& g, cudaStream_t stream) {
  using Cfg = PlCfg<C>;
  static bool attr = false;
  if (!attr) {
    KVQ_CUDA(cudaFuncSetAttribute(proj_ln_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  CUtensorMap tmA, tmW;
  int rc = make_tmap_2d(&tmA, attn_out, static_cast<uint64_t>(rows), C, static_cast<uint64_t>(C) * 2, PL_M, 64, 2, 128);
  if (rc != 0) return rc;
  rc = make_tmap_2d(&tmW, w, C, C, static_cast<uint64_t>(C) * 2, C, 64, 2, 128);
  if (rc != 0) return rc;
  const int tiles = static_cast<int>((rows + PL_M - 1) / PL_M);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  count_launch();
  return launch_pdl(proj_ln_kernel<C>, dim3(grid), dim3(PL_THREADS), Cfg::SMEM, stream, tmA, tmW, bias, gamma, beta, eps, x, a2,
                    static_cast<int>(rows), g);
}

}  // namespace

bool proj_ln_supported(int C) { return C == 96 || C == 192; }

int launch_proj_ln(const __half* attn_out, const __half* w, const float* bias, const float* gamma, const float* beta,
                   float eps, float* x, __half* a2, int B, int C, const WinGeom& g, cudaStream_t stream) {
  const long long rows = static_cast<long long>(B) * g.nW * g.N;
  KVQ_REQUIRE(rows > 0 && rows < (1ll << 31) && attn_out && w && bias && gamma && beta && x && a2, KVQ_ERR_BAD_SHAPE,
              "proj_ln: bad arguments (%lld rows)", rows);
  if (C == 96) return launch_proj_ln_impl<96>(attn_out, w, bias, gamma, beta, eps, x, a2, rows, g, stream);
  if (C == 192) return launch_proj_ln_impl<192>(attn_out, w, bias, gamma, beta, eps, x, a2, rows, g, stream);
  set_error("proj_ln: C=%d is not built (96, 192)", C);
  return KVQ_ERR_BAD_SHAPE;
}

}  // namespace kvq
