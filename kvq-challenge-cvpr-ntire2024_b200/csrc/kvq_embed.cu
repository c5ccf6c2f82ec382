// PatchEmbed3D as ONE kernel: Conv3d k = s = (2,4,4), 3 -> 96 channels, + bias + LayerNorm(96)
// (swin_backbone.py:690-733), reading the clip where it lies -- no patch matrix.
//
// A row of the implicit [tokens, 96] operand is (c, kt, kh, kw) with the 4 kw taps contiguous in the clip, and for a
// fixed (c, kt, kh) consecutive tokens along w read consecutive 16-byte (fp32) / 8-byte (fp16) pieces of one clip row.
// TMA cannot build the operand (a K granule of the MMA is 16 bytes = 8 halfs, a patch row piece is 4), so eight
// producer warps do: coalesced loads, fp32 -> fp16, 8-byte stores into the 128B-swizzled K-major A tile.  The 96 x 96
// weight (or its [W_hi | W_lo] split pair) stays resident in shared memory for the whole kernel.
//
//   warps 0..7   : producers (thread = tile row x half of the 24 (c, kt, kh) groups; 12 loads in flight per thread)
//   warp 8       : lane 0 issues the 6 (12 with split weights) tcgen05.mma M128 x N96 x K16 of a tile
//   warps 9..11  : idle (fill the warpgroup)
//   warps 12..15 : epilogue, thread = token: bias, two-pass LayerNorm statistics over the 96 TMEM columns, normalise,
//                  fp32 rows through a per-warp shared-memory transpose (whole 32 B sectors per store)
// A tiles and accumulators are double-buffered.  Replaces patch_im2col_kernel + gemm_kernel<96, EPI_LN_F32>
// (0.098 + 0.114 ms per batch-8 step: the patch matrix was a 77 MB write and a 77 MB read).
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int PE_THREADS = 512;
constexpr int PE_M = 128;
constexpr int PE_C = 96;
constexpr int PE_A_BYTES = 2 * PE_M * 128;            // two 64-wide K blocks (the second one half used)
constexpr int PE_W_BLOCK = PE_C * 128;                // one 64-wide K block of the weight: 96 rows x 128 B
constexpr int PE_OFF_A = 0;
constexpr int PE_OFF_W = PE_OFF_A + 2 * PE_A_BYTES;   // up to 4 K blocks: hi0 hi1 lo0 lo1
constexpr int PE_OFF_STAGE = PE_OFF_W + 4 * PE_W_BLOCK;
constexpr int PE_STAGE_BYTES = 4 * (32 * 36 * 4);
constexpr int PE_OFF_PAR = PE_OFF_STAGE + PE_STAGE_BYTES;   // bias | gamma | beta, 96 floats each
constexpr int PE_OFF_BAR = PE_OFF_PAR + 3 * PE_C * 4;
constexpr int PE_SMEM = PE_OFF_BAR + 128 + 1024;

template <typename TIn>
__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmW, const TIn* __restrict__ x, const float* __restrict__ bias,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ out,
                   int B, int T, int H, int W, int D, int Hs, int Ws, int M, int split) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* spar = reinterpret_cast<float*>(smem + PE_OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PE_OFF_BAR);
  uint64_t* a_full = bars + 0;    // [2]  producers wrote the A tile
  uint64_t* a_empty = bars + 2;   // [2]  its MMAs have run
  uint64_t* o_full = bars + 4;    // [2]
  uint64_t* o_empty = bars + 6;   // [2]
  uint64_t* w_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (M + PE_M - 1) / PE_M;
  const int nwb = split ? 4 : 2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 8);
      mbar_init(&a_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 3 * PE_C; i += PE_THREADS)
    spar[i] = i < PE_C ? __ldg(bias + i) : i < 2 * PE_C ? __ldg(gamma + i - PE_C) : __ldg(beta + i - 2 * PE_C);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp < 8) {
    // ============================ producers ============================
    const int r = threadIdx.x & 127, gh = threadIdx.x >> 7;    // tile row, which 12 of the 24 (c, kt, kh) groups
    uint32_t n_t = 0;
    const float r_Ws = 1.0f / Ws, r_Hs = 1.0f / Hs, r_D = 1.0f / D;
    const bool fast = (T % 2 == 0) && (H % 4 == 0) && (W % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
      const int buf = n_t & 1;
      const int tok = tile * PE_M + r;                     // < 2^31 (checked by the launcher)
      const bool live = tok < M;
      // float-reciprocal divisions (fdiv_i): a generic integer division is a ~100-cycle dependent chain per thread
      const int bdh = fdiv_i(tok, Ws, r_Ws), ws = tok - bdh * Ws;
      const int bd = fdiv_i(bdh, Hs, r_Hs), hs = bdh - bd * Hs;
      const int b = fdiv_i(bd, D, r_D), d = bd - b * D;
      const int w0 = 4 * ws;
      const bool wfull = w0 + 3 < W;
      uint2 v[12];
      if (fast) {
        // interior geometry (T even, H and W multiples of 4, 16-byte aligned clip): straight-line vector loads, all
        // twelve in flight before the first conversion (the guarded form below serialises them behind its branches)
        const TIn* base = x + ((static_cast<size_t>(b) * 3 * T + 2 * d) * H + 4 * hs) * W + w0;
        if constexpr (sizeof(TIn) == 4) {
          float4 f[12];
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const int g = gh * 12 + i;
            const int c = g >> 3, kt = (g >> 2) & 1, kh = g & 3;
            f[i] = live ? __ldg(reinterpret_cast<const float4*>(base + (static_cast<size_t>(c) * T + kt) * H * W + kh * W))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 12; ++i) v[i] = make_uint2(pack_half2(f[i].x, f[i].y), pack_half2(f[i].z, f[i].w));
        } else {
#pragma unroll
          for (int i = 0; i < 12; ++i) {
            const int g = gh * 12 + i;
            const int c = g >> 3, kt = (g >> 2) & 1, kh = g & 3;
            v[i] = live ? __ldg(reinterpret_cast<const uint2*>(base + (static_cast<size_t>(c) * T + kt) * H * W + kh * W))
                        : make_uint2(0u, 0u);
          }
        }
      } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int g = gh * 12 + i;
        const int c = g >> 3, kt = (g >> 2) & 1, kh = g & 3;
        const int t = 2 * d + kt, h = 4 * hs + kh;
        v[i] = make_uint2(0u, 0u);
        if (live && t < T && h < H) {
          const TIn* src = x + (((static_cast<size_t>(b) * 3 + c) * T + t) * H + h) * W + w0;
          if constexpr (sizeof(TIn) == 4) {
            float f[4] = {0.f, 0.f, 0.f, 0.f};
            if (wfull && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(src));
              f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (w0 + k < W) f[k] = __ldg(src + k);
            }
            v[i].x = pack_half2(f[0], f[1]);
            v[i].y = pack_half2(f[2], f[3]);
          } else {
            if (wfull && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
              v[i] = __ldg(reinterpret_cast<const uint2*>(src));
            } else {
              unsigned short u[4] = {0, 0, 0, 0};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (w0 + k < W) u[k] = __ldg(reinterpret_cast<const unsigned short*>(src) + k);
              v[i].x = u[0] | (static_cast<uint32_t>(u[1]) << 16);
              v[i].y = u[2] | (static_cast<uint32_t>(u[3]) << 16);
            }
          }
        }
      }
      }
      if (lane == 0) mbar_wait(&a_empty[buf], ((n_t >> 1) & 1) ^ 1);
      __syncwarp();
      // K index of group g is 4g: K block g / 16, 16-byte chunk (g % 16) / 2 (XOR-swizzled by the row), half g & 1
      uint8_t* arow = smem + PE_OFF_A + buf * PE_A_BYTES + r * 128;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int g = gh * 12 + i;
        *reinterpret_cast<uint2*>(arow + (g >> 4) * (PE_M * 128) + ((((g & 15) >> 1) ^ (r & 7)) << 4) + (g & 1) * 8) = v[i];
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(PE_M, PE_C, 0, 0);
      mbar_expect_tx(w_full, static_cast<uint32_t>(nwb * PE_W_BLOCK));
      for (int kb = 0; kb < nwb; ++kb) tma_load_2d(smem + PE_OFF_W + kb * PE_W_BLOCK, &tmW, w_full, kb * 64, 0);
      mbar_wait(w_full, 0);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t dA0 = umma_smem_desc(sbase + PE_OFF_A, 16, 1024, UMMA_SW_128);
      const uint64_t dW0 = umma_smem_desc(sbase + PE_OFF_W, 16, 1024, UMMA_SW_128);
      uint32_t n_t = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
        const uint32_t buf = n_t & 1, ph = (n_t >> 1) & 1;
        mbar_wait(&o_empty[buf], ph ^ 1);
        mbar_wait(&a_full[buf], ph);
        tc_fence_after();
        const uint64_t da = dA0 + static_cast<uint64_t>(buf * (PE_A_BYTES >> 4));
        const uint32_t tD = tmem_base + buf * PE_C;
        const int passes = split ? 2 : 1;
        for (int ps = 0; ps < passes; ++ps) {
          const uint64_t dw = dW0 + static_cast<uint64_t>(ps * ((2 * PE_W_BLOCK) >> 4));
#pragma unroll
          for (int ks = 0; ks < 6; ++ks) {
            const int kb = ks >> 2, k = ks & 3;
            umma_f16_ss(tD, da + static_cast<uint64_t>(((kb * PE_M * 128) >> 4) + 2 * k),
                        dw + static_cast<uint64_t>(((kb * PE_W_BLOCK) >> 4) + 2 * k), idesc, (ps > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&o_full[buf]);
      }
    }
  } else if (warp >= 12) {
    // ============================ epilogue: bias + LayerNorm(96) ============================
    const int q = warp & 3;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* stile = reinterpret_cast<float*>(smem + PE_OFF_STAGE + q * (32 * 36 * 4));
    const int sub = lane & 7, rsel = lane >> 3;
    uint32_t n_t = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_t) {
      const uint32_t buf = n_t & 1;
      if (lane == 0) mbar_wait(&o_full[buf], (n_t >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t tD = lane_addr + buf * PE_C;
      // fp32 two-pass statistics like nn.LayerNorm: mean, then the centred sum of squares (TMEM is re-read, not held)
      float s = 0.f;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        uint32_t rr[32];
        tmem_ld_x32(tD + ci * 32, rr);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) s += __uint_as_float(rr[j]) + spar[ci * 32 + j];
      }
      const float mean = s * (1.0f / PE_C);
      float qq = 0.f;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        uint32_t rr[32];
        tmem_ld_x32(tD + ci * 32, rr);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dlt = __uint_as_float(rr[j]) + spar[ci * 32 + j] - mean;
          qq = fmaf(dlt, dlt, qq);
        }
      }
      const float rstd = rsqrtf(qq * (1.0f / PE_C) + eps);
      const int row0 = tile * PE_M + q * 32;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        uint32_t rr[32];
        tmem_ld_x32(tD + ci * 32, rr);
        tmem_wait_ld();
        __syncwarp();                                   // the previous chunk's readers are done with the tile
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 y;
          y.x = (__uint_as_float(rr[4 * k]) + spar[ci * 32 + 4 * k] - mean) * rstd;
          y.y = (__uint_as_float(rr[4 * k + 1]) + spar[ci * 32 + 4 * k + 1] - mean) * rstd;
          y.z = (__uint_as_float(rr[4 * k + 2]) + spar[ci * 32 + 4 * k + 2] - mean) * rstd;
          y.w = (__uint_as_float(rr[4 * k + 3]) + spar[ci * 32 + 4 * k + 3] - mean) * rstd;
          *reinterpret_cast<float4*>(stile + lane * 36 + 4 * k) = y;
        }
        __syncwarp();
        const float4 gm = *reinterpret_cast<const float4*>(spar + PE_C + ci * 32 + 4 * sub);
        const float4 bt = *reinterpret_cast<const float4*>(spar + 2 * PE_C + ci * 32 + 4 * sub);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rloc = it * 4 + rsel;
          const int grow = row0 + rloc;
          if (grow < M) {
            float4 y = *reinterpret_cast<const float4*>(stile + rloc * 36 + 4 * sub);
            y.x = fmaf(y.x, gm.x, bt.x); y.y = fmaf(y.y, gm.y, bt.y); y.z = fmaf(y.z, gm.z, bt.z); y.w = fmaf(y.w, gm.w, bt.w);
            *reinterpret_cast<float4*>(out + static_cast<size_t>(grow) * PE_C + ci * 32 + 4 * sub) = y;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

int launch_patch_embed(const void* x, int x_is_f16, const __half* w, int ldw, int split, const float* bias,
                       const float* gamma, const float* beta, float eps, float* out, int B, int T, int H, int W,
                       cudaStream_t stream) {
  const int D = (T + 1) / 2, Hs = (H + 3) / 4, Ws = (W + 3) / 4;
  const long long M = static_cast<long long>(B) * D * Hs * Ws;
  KVQ_REQUIRE(M > 0 && M < (1ll << 31), KVQ_ERR_BAD_SHAPE, "patch_embed: %lld tokens", M);
  KVQ_REQUIRE(ldw == (split ? 256 : 96), KVQ_ERR_BAD_SHAPE, "patch_embed: weight row stride %d (need %d)", ldw,
              split ? 256 : 96);
  static bool attr = false;
  if (!attr) {
    KVQ_CUDA(cudaFuncSetAttribute(patch_embed_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM));
    KVQ_CUDA(cudaFuncSetAttribute(patch_embed_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM));
    attr = true;
  }
  CUtensorMap tmW;
  const int rc = make_tmap_2d(&tmW, w, PE_C, static_cast<uint64_t>(ldw), static_cast<uint64_t>(ldw) * 2, PE_C, 64, 2, 128);
  if (rc != 0) return rc;
  const int tiles = static_cast<int>((M + PE_M - 1) / PE_M);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  count_launch();
  if (x_is_f16)
    return launch_pdl(patch_embed_kernel<__half>, dim3(grid), dim3(PE_THREADS), PE_SMEM, stream, tmW,
                      static_cast<const __half*>(x), bias, gamma, beta, eps, out, B, T, H, W, D, Hs, Ws,
                      static_cast<int>(M), split);
  return launch_pdl(patch_embed_kernel<float>, dim3(grid), dim3(PE_THREADS), PE_SMEM, stream, tmW,
                    static_cast<const float*>(x), bias, gamma, beta, eps, out, B, T, H, W, D, Hs, Ws,
                    static_cast<int>(M), split);
}

}  // namespace kvq
