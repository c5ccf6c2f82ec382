// Fused 3-D (shifted-)window attention (WindowAttention3D.forward, swin_backbone.py:245-326, with the index tensors
// of compute_mask :560-586 and global_position_index :22-50 evaluated from token coordinates instead of being
// materialised as [nW,392,392(,3)] tensors):
//
//   S = Q K^T                         tcgen05.mma  M=128 x N=400 (208 + 192) x K=32, fp32 accumulators in TMEM
//   S += rel[rpi]*fg + frag[rpi]*(1-fg)  (GRPB gate fg = L1 fragment distance, :291-309)  [+ -100 region mask]
//   P = softmax(S)                    fp32, exact row max, exp2 on log2-domain logits; P written fp16 to smem
//   O = P V                           tcgen05.mma  M=128 x N=32 x K=400, V read MN-major straight from its image
//   out[row, head*32:+32] = O / rowsum
//
// Operands arrive as one contiguous 76 800 B image per (window, head) unit (Q|K|V, UMMA no-swizzle core-matrix
// order, written by the QKV GEMM epilogue) fetched with cp.async.bulk.  The logits never leave the SM.
// Everything is kept in the log2 domain: q is pre-scaled by hd^-0.5 * log2(e), the packed tables by log2(e).
//
// Two kernels:
//   window_attn_fast_kernel   the (8,7,7) window of every shipped config: persistent CTAs (one head per CTA so the
//                             20 KB bias table is staged once), 7x7-unrolled bias loop with compile-time table
//                             offsets, next unit's operands prefetched under the last tile, 8-row tail tile
//                             replicated into all four TMEM lane quadrants so all 8 warps share it
//   window_attn_kernel        any clamped window <= (8,7,7) (small clips / D <= 8), per-key metadata in smem
#include <cstdlib>

#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int ATT_THREADS = 256;
constexpr int P_SBO = (ATT_ROWS / 8) * 128;  // 6400: byte stride between 8-row groups of the P image
constexpr int P_BYTES = 128 * ATT_ROWS * 2;  // 102 400
constexpr int TMEM_O_COL = 400;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASK_L2 = -100.0f * LOG2E;   // compute_mask's -100 (:583) in the log2 domain

struct ColMeta {  // 8 bytes per key slot (generic kernel)
  int base;       // rd*(2bh-1)(2bw-1) + rh*(2bw-1) + rw   (token enumerated in the BASE window, :264)
  uint32_t pk;    // fh | fw << 8 | region << 16 | valid << 24
};

__device__ __forceinline__ void issue_s_mma(uint32_t tmem_base, uint32_t aQ, uint32_t aK) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint64_t dq = umma_smem_desc(aQ + ks * 256, 128, 512, UMMA_SW_NONE);
    const uint64_t dk0 = umma_smem_desc(aK + ks * 256, 128, 512, UMMA_SW_NONE);
    const uint64_t dk1 = umma_smem_desc(aK + 26 * 512 + ks * 256, 128, 512, UMMA_SW_NONE);
    umma_f16_ss(tmem_base, dq, dk0, umma_idesc_f16(128, 208, 0, 0), ks);
    umma_f16_ss(tmem_base + 208, dq, dk1, umma_idesc_f16(128, 192, 0, 0), ks);
  }
}

__device__ __forceinline__ void issue_pv_mma(uint32_t tmem_o, uint32_t aP, uint32_t aV) {
#pragma unroll 1
  for (int ks = 0; ks < ATT_ROWS / 16; ++ks) {
    const uint64_t dp = umma_smem_desc(aP + ks * 256, 128, P_SBO, UMMA_SW_NONE);
    // V image read MN-major: 8-channel chunks 128 B apart (SBO), 8-key groups 512 B apart (LBO)
    const uint64_t dv = umma_smem_desc(aV + ks * 1024, 512, 128, UMMA_SW_NONE);
    umma_f16_ss(tmem_o, dp, dv, umma_idesc_f16(128, 32, 0, 1), ks);
  }
}

__device__ __forceinline__ void store_o_row(__half* dst, const uint32_t* r, float inv) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 o;
    o.x = pack_half2(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv);
    o.y = pack_half2(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
    o.z = pack_half2(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv);
    o.w = pack_half2(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv);
    *reinterpret_cast<uint4*>(dst + j) = o;
  }
}

__device__ __forceinline__ void store_o_half(__half* dst, const uint32_t* r, float inv) {   // 16 channels
#pragma unroll
  for (int j = 0; j < 16; j += 8) {
    uint4 o;
    o.x = pack_half2(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv);
    o.y = pack_half2(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
    o.z = pack_half2(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv);
    o.w = pack_half2(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv);
    *reinterpret_cast<uint4*>(dst + j) = o;
  }
}

// =====================================================================================================
// fast path: window (8,7,7), N = 392
// =====================================================================================================
constexpr int FQ = 0;                           // Q image
constexpr int FK = ATT_IMG_BYTES;               // K image
constexpr int FV = 2 * ATT_IMG_BYTES;           // V image
constexpr int FQT = 3 * ATT_IMG_BYTES;          // 128-row Q tile holding the 8 tail rows in groups 0,4,8,12
constexpr int FP = FQT + 8192;                  // P image
constexpr int FSMALL = FP + P_BYTES;            // small tables / exchange buffers / barriers (7 KB)
constexpr int FTAB = FSMALL + 7168;             // bias table
// Fast-path table layout: idx = (dd+7)*TS_D + (dh+6)*TS_H + (dw+6) with TS_H = 7 (mod 16) and TS_D = 49 (mod 16), so
// the float2 entries read by consecutive query rows (w fastest, then h, then d) fall on consecutive 8-byte bank
// pairs: a half-warp's LDS.64 is conflict-free wherever the rows wrap.
constexpr int TS_H = 23;
constexpr int TS_D = 289;                       // >= 12*TS_H + 13
constexpr int FAST_TAB_LEN = 15 * TS_D + 1;     // 4336 (even)
constexpr int FAST_SMEM = FTAB + FAST_TAB_LEN * 8 + 128;
static_assert(FAST_SMEM <= 227 * 1024, "fast attention kernel exceeds shared memory");

struct FastSmall {
  float fhc[8], fwc[8];         // fragment id of in-window row h / column w (this window)
  float rhm[8], rwm[8], rdm[8]; // region id per in-window index (as float)
  float smax[2][2][128];        // [tile parity][column half][row]
  float ssum[2][2][128];        // [tile parity][column half][row]
  float tmax[8][8];             // tail tile: [warp][row]
  float tsum[8][8];
  uint64_t bar_tab, bar_qk, bar_v, bar_s, bar_o, bar_p;
  uint32_t tmem_slot;
};
static_assert(sizeof(FastSmall) <= 7168, "FastSmall overflows its slot");

#ifdef KVQ_TIMING
// debug build: per-phase cycle counters of the softmax warps (summed over warps' lane 0), read by kvq_debug_timers
__device__ unsigned long long g_attn_timers[16];
#define TMARK(slot)                                               \
  do {                                                            \
    const long long _now = clock64();                             \
    if (lane == 0) tacc[slot] += _now - tprev;                    \
    tprev = _now;                                                 \
  } while (0)
#else
#define TMARK(slot) do {} while (0)
#endif

constexpr int FAST_SOFTMAX_THREADS = 256;
constexpr int FAST_THREADS = FAST_SOFTMAX_THREADS + 32;   // + one control warp: bulk loads and tcgen05.mma issue

__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// pass 1 over one 50-column slab: v = s + t0[idx] + fg * t1[idx] (+ mask); running max; write back (pad = -inf).
// The table gathers are software-pipelined one key row (7 entries) ahead of the arithmetic: with only two warps per
// scheduler an LDS consumed a few instructions after issue stalls the in-order warp for most of its ~30 cycle latency.
template <bool MASK>
__device__ __forceinline__ void pass1_slab(uint32_t taddr, uint32_t trow_addr, const float (&Ah)[7],
                                           const float (&Aw)[7], float m00, float m01, float m10, float m11,
                                           float (&mx)[4]) {
  uint32_t r[50];
  tmem_ld_x32(taddr, r);
  tmem_ld_x16(taddr + 32, r + 32);
  tmem_ld_x2(taddr + 48, r + 48);
  float2 e[2][7];
#pragma unroll
  for (int wj = 0; wj < 7; ++wj) e[0][wj] = lds_f2(trow_addr - 8u * wj);
  tmem_wait_ld();
#pragma unroll
  for (int hj = 0; hj < 7; ++hj) {
    if (hj < 6) {
#pragma unroll
      for (int wj = 0; wj < 7; ++wj) e[(hj + 1) & 1][wj] = lds_f2(trow_addr - 8u * ((hj + 1) * TS_H + wj));
    }
#pragma unroll
    for (int wj = 0; wj < 7; ++wj) {
      const int j = hj * 7 + wj;
      const float2 ee = e[hj & 1][wj];
      const float fg = Ah[hj] + Aw[wj];
      float v = fmaf(fg, ee.y, __uint_as_float(r[j])) + ee.x;
      if (MASK) v += (hj < 4) ? ((wj < 4) ? m00 : m01) : ((wj < 4) ? m10 : m11);
      mx[j & 3] = fmaxf(mx[j & 3], v);
      r[j] = __float_as_uint(v);
    }
  }
  r[49] = __float_as_uint(-INFINITY);
  tmem_st_x32(taddr, r);
  tmem_st_x16(taddr + 32, r + 32);
  tmem_st_x2(taddr + 48, r + 48);
}

// Q|K images (+ the 8 tail query rows replicated into row groups 0,4,8,12 of the tail tile) of one unit
__device__ __forceinline__ void fast_load_qk(uint8_t* smem, const uint8_t* unit_src, uint64_t* bar) {
  mbar_expect_tx(bar, 2 * ATT_IMG_BYTES + 4 * 512);
  bulk_load_1d(smem + FQ, unit_src, 2 * ATT_IMG_BYTES, bar);
#pragma unroll
  for (int k = 0; k < 4; ++k) bulk_load_1d(smem + FQT + k * 4 * 512, unit_src + 48 * 512, 512, bar);
}

__global__ void __launch_bounds__(FAST_THREADS, 1)
window_attn_fast_kernel(const AttnParams p, const float2* __restrict__ tabs, int units) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw_addr & 127u)) & 127u);
  FastSmall& sm = *reinterpret_cast<FastSmall*>(smem + FSMALL);
  float2* stab = reinterpret_cast<float2*>(smem + FTAB);
  uint8_t* sP = smem + FP;

  const WinGeom& g = p.geom;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;                     // multiple of heads: this CTA always sees the same head
  const int head = blockIdx.x % p.heads;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.img);

  if (tid == FAST_SOFTMAX_THREADS) {           // control thread
    mbar_init(&sm.bar_tab, 1);
    mbar_init(&sm.bar_qk, 1);
    mbar_init(&sm.bar_v, 1);
    mbar_init(&sm.bar_s, 1);
    mbar_init(&sm.bar_o, 1);
    mbar_init(&sm.bar_p, 8);                   // one arrive per softmax warp
    mbar_fence_init();
    // the fast-layout tables follow the compact ones in the packed buffer
    mbar_expect_tx(&sm.bar_tab, FAST_TAB_LEN * 8);
    bulk_load_1d(stab, tabs + static_cast<size_t>(p.heads) * 2536 + static_cast<size_t>(head) * FAST_TAB_LEN,
                 FAST_TAB_LEN * 8, &sm.bar_tab);
    if (static_cast<int>(blockIdx.x) < units) {
      const uint8_t* src = img + static_cast<size_t>(blockIdx.x) * ATT_UNIT_BYTES;
      fast_load_qk(smem, src, &sm.bar_qk);
      mbar_expect_tx(&sm.bar_v, ATT_IMG_BYTES);
      bulk_load_1d(smem + FV, src + 2 * ATT_IMG_BYTES, ATT_IMG_BYTES, &sm.bar_v);
    }
  }
  if (warp == 1) {
    tmem_alloc(&sm.tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;

  if (warp == 8) {
    // =============================== control warp ===============================
    if (lane == 0) {
      const uint32_t aQ = smem_u32(smem + FQ), aK = smem_u32(smem + FK), aV = smem_u32(smem + FV);
      const uint32_t aQT = smem_u32(smem + FQT), aP = smem_u32(sP);
      uint32_t n_load = 0, n_s = 0, n_o = 0, n_p = 0;
      for (int unit = blockIdx.x; unit < units; unit += G) {
        const bool has_next = unit + G < units;
        const uint8_t* next_src = img + static_cast<size_t>(unit + G) * ATT_UNIT_BYTES;
        mbar_wait(&sm.bar_qk, n_load & 1);
        tc_fence_after();
        issue_s_mma(tmem_base, aQ, aK);
        umma_commit(&sm.bar_s);
#pragma unroll 1
        for (int t = 0; t < 4; ++t) {
          if (t == 3) {
            // S(3) must have consumed Q and K before the next unit's images land on top of them
            mbar_wait(&sm.bar_s, n_s & 1);
            if (has_next) fast_load_qk(smem, next_src, &sm.bar_qk);
          }
          ++n_s;
          mbar_wait(&sm.bar_p, n_p & 1);       // softmax warps: P(t) written, S(t) fully consumed
          ++n_p;
          tc_fence_after();
          // S(t+1) first: the next tile's softmax starts as soon as its logits land while PV(t) drains behind it
          if (t < 3) {
            issue_s_mma(tmem_base, t == 2 ? aQT : aQ + (t + 1) * 8192u, aK);
            umma_commit(&sm.bar_s);
          }
          if (t == 0) mbar_wait(&sm.bar_v, n_load & 1);
          issue_pv_mma(tmem_base + TMEM_O_COL, aP, aV);
          umma_commit(&sm.bar_o);
          // observe every phase of bar_o in order (a parity wait two phases behind would alias); the next event this
          // thread needs, P(t+1), is thousands of cycles away, so this costs nothing
          mbar_wait(&sm.bar_o, n_o & 1);
          ++n_o;
          if (t == 3 && has_next) {            // last PV done: V is free
            mbar_expect_tx(&sm.bar_v, ATT_IMG_BYTES);
            bulk_load_1d(smem + FV, next_src + 2 * ATT_IMG_BYTES, ATT_IMG_BYTES, &sm.bar_v);
          }
        }
        ++n_load;
      }
    }
  } else {
    // =============================== softmax warps ===============================
    const int q = warp & 3, hs = warp >> 2;
    const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t n_s = 0, n_o = 0;
#ifdef KVQ_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#endif
    mbar_wait(&sm.bar_tab, 0);
    TMARK(0);  // table wait
    int prev_t = -1, prev_win_g = 0;   // tile whose O = P V is still waiting for its epilogue
    auto o_epilogue = [&](int wg, int tp) {
      if (tp < 3) {
        uint32_t r[16];
        tmem_ld_x16(taddr_row + TMEM_O_COL + 16 * hs, r);
        tmem_wait_ld();
        const int row = q * 32 + lane;
        const float inv = 1.0f / (sm.ssum[tp & 1][0][row] + sm.ssum[tp & 1][1][row]);
        __half* dst = p.out + (static_cast<size_t>(wg) * 392 + tp * 128 + row) * p.C + head * ATT_HD + 16 * hs;
        store_o_half(dst, r, inv);
      } else if (q == 0) {
        uint32_t r[16];
        tmem_ld_x16(taddr_row + TMEM_O_COL + 16 * hs, r);
        tmem_wait_ld();
        if (lane < 8) {
          float tot = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < 8; ++w2) tot += sm.tsum[w2][lane];
          __half* dst = p.out + (static_cast<size_t>(wg) * 392 + 384 + lane) * p.C + head * ATT_HD + 16 * hs;
          store_o_half(dst, r, 1.0f / tot);
        }
      }
      tc_fence_before();
    };

    for (int unit = blockIdx.x; unit < units; unit += G) {
      const int win_g = unit / p.heads;
      const int win = win_g % g.nW;
      const int wdi = win / (g.nwh * g.nww), whi = (win / g.nww) % g.nwh, wwi = win % g.nww;
      // ---- per-window coordinate tables (compute_mask / global_position_index from coordinates) ----
      if (tid < 8) {
        const int k = tid;
        const int ph = whi * 7 + k, pw = wwi * 7 + k, pd = wdi * 8 + k;
        int oh = ph + g.sh; if (oh >= g.Hp) oh -= g.Hp;
        int ow = pw + g.sw; if (ow >= g.Wp) ow -= g.Wp;
        int fh = static_cast<int>(floorf(static_cast<float>(oh) * (7.0f / static_cast<float>(g.Hp)))); if (fh > 6) fh = 6;
        int fw = static_cast<int>(floorf(static_cast<float>(ow) * (7.0f / static_cast<float>(g.Wp)))); if (fw > 6) fw = 6;
        sm.fhc[k] = static_cast<float>(fh);
        sm.fwc[k] = static_cast<float>(fw);
        sm.rhm[k] = g.sh == 0 ? 0.f : static_cast<float>((ph >= g.Hp - 7) + (ph >= g.Hp - g.sh));
        sm.rwm[k] = g.sw == 0 ? 0.f : static_cast<float>((pw >= g.Wp - 7) + (pw >= g.Wp - g.sw));
        sm.rdm[k] = g.sd == 0 ? 0.f : static_cast<float>((pd >= g.Dp - 8) + (pd >= g.Dp - g.sd));
      }
      // a window needs the region mask only when it straddles a region boundary in some dim (last window of a
      // shifted dim); CTA-uniform
      const bool masked = (g.sh != 0 && whi == g.nwh - 1) || (g.sw != 0 && wwi == g.nww - 1) ||
                          (g.sd != 0 && wdi == g.nwd - 1);
      named_bar_sync(1, FAST_SOFTMAX_THREADS);

#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        const bool tail = (t == 3);
        const int par = t & 1;
        TMARK(7);  // unit prologue / previous arrive
        mbar_wait(&sm.bar_s, n_s & 1);
        ++n_s;
        __syncwarp();
        tc_fence_after();
        if (t == 0) TMARK(6); else TMARK(1);  // wait for S(0) (includes the unit's operand load) / S(t>0)

        // ---- this thread's query row and its constants ----
        const int ri = tail ? 384 + (lane & 7) : t * 128 + q * 32 + lane;
        const int d_i = ri / 49, hw_i = ri - d_i * 49, h_i = hw_i / 7, w_i = hw_i - h_i * 7;
        float Ah[7], Aw[7];
        {
          const float fh_i = sm.fhc[h_i], fw_i = sm.fwc[w_i];
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            Ah[k] = fabsf(fh_i - sm.fhc[k]);
            Aw[k] = fabsf(fw_i - sm.fwc[k]);
          }
        }
        float hlo = 0.f, hhi = 0.f, wlo = 0.f, whi_m = 0.f, rd_i = 0.f;
        if (masked) {
          // in-window region changes at most once per dim, at index 4 (= window - shift)
          hlo = (sm.rhm[0] != sm.rhm[h_i]) ? MASK_L2 : 0.f;
          hhi = (sm.rhm[6] != sm.rhm[h_i]) ? MASK_L2 : 0.f;
          wlo = (sm.rwm[0] != sm.rwm[w_i]) ? MASK_L2 : 0.f;
          whi_m = (sm.rwm[6] != sm.rwm[w_i]) ? MASK_L2 : 0.f;
          rd_i = sm.rdm[d_i];
        }
        // table row pointer for key slab d:  idx = (d_i+7)*TS_D + (h_i+6)*TS_H + (w_i+6) - d*TS_D - (hj*TS_H + wj)
        const uint32_t trow0 = smem_u32(stab) + 8u * static_cast<uint32_t>((d_i + 7) * TS_D + (h_i + 6) * TS_H + (w_i + 6));

        // ---- pass 1 ----
        const int slab0 = tail ? warp : hs * 4;
        const int nslab = tail ? 1 : 4;
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (masked) {
#pragma unroll 1
          for (int s = 0; s < nslab; ++s) {
            const int d = slab0 + s;
            const float dm = (sm.rdm[d] != rd_i) ? MASK_L2 : 0.f;
            pass1_slab<true>(taddr_row + d * ATT_SLAB, trow0 - 8u * (d * TS_D), Ah, Aw, fminf(fminf(hlo, wlo), dm),
                             fminf(fminf(hlo, whi_m), dm), fminf(fminf(hhi, wlo), dm), fminf(fminf(hhi, whi_m), dm),
                             mx);
          }
        } else {
#pragma unroll 1
          for (int s = 0; s < nslab; ++s) {
            const int d = slab0 + s;
            pass1_slab<false>(taddr_row + d * ATT_SLAB, trow0 - 8u * (d * TS_D), Ah, Aw, 0.f, 0.f, 0.f, 0.f, mx);
          }
        }
        tmem_wait_st();
        TMARK(2);  // pass 1
        float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (tail) {
          if (lane < 8) sm.tmax[warp][lane] = m;
        } else {
          sm.smax[par][hs][q * 32 + lane] = m;
        }
        named_bar_sync(1, FAST_SOFTMAX_THREADS);
        TMARK(3);  // max exchange barrier
        if (tail) {
          m = sm.tmax[0][lane & 7];
#pragma unroll
          for (int w2 = 1; w2 < 8; ++w2) m = fmaxf(m, sm.tmax[w2][lane & 7]);
        } else {
          m = fmaxf(sm.smax[par][0][q * 32 + lane], sm.smax[par][1][q * 32 + lane]);
        }

        // ---- O epilogue of the previous tile -- of this unit or, for t == 0, the tail tile of the previous unit:
        //      its PV finished long ago, and it must precede overwriting P.  The two column halves split O.
        if (prev_t >= 0) {
          mbar_wait(&sm.bar_o, n_o & 1);
          ++n_o;
          __syncwarp();
          tc_fence_after();
          o_epilogue(prev_win_g, prev_t);
        }
        TMARK(4);  // wait PV(t-1) + O epilogue
        // ---- pass 2: P = exp2(v - max) -> fp16 smem image; row sums ----
        float sum = 0.f;
        if (!tail) {
          uint8_t* prow = sP + (q * 4 + (lane >> 3)) * P_SBO + (lane & 7) * 16;
          const int col0 = hs * 200;
          const float2 negm2 = make_float2(-m, -m);
          float2 sum2 = make_float2(0.f, 0.f);
          auto exp_chunk = [&](const uint32_t* r, int c_abs, int n) {
#pragma unroll
            for (int j = 0; j < n; j += 8) {
              uint32_t h[4];
#pragma unroll
              for (int k = 0; k < 8; k += 2) {
                // packed fp32x2: one FADD2 for the two max subtractions, one for the two row-sum updates
                const float2 d = fadd2(make_float2(__uint_as_float(r[j + k]), __uint_as_float(r[j + k + 1])), negm2);
                const float2 e = make_float2(fast_exp2(d.x), fast_exp2(d.y));
                sum2 = fadd2(sum2, e);
                h[k >> 1] = pack_half2(e.x, e.y);
              }
              *reinterpret_cast<uint4*>(prow + ((c_abs + j) >> 3) * 128) = make_uint4(h[0], h[1], h[2], h[3]);
            }
          };
#pragma unroll 1
          for (int c = 0; c < 192; c += 32) {
            uint32_t r[32];
            tmem_ld_x32(taddr_row + col0 + c, r);
            tmem_wait_ld();
            exp_chunk(r, col0 + c, 32);
          }
          {
            uint32_t r[8];
            tmem_ld_x8(taddr_row + col0 + 192, r);
            tmem_wait_ld();
            exp_chunk(r, col0 + 192, 8);
          }
          sum = sum2.x + sum2.y;
          sm.ssum[par][hs][q * 32 + lane] = sum;
        } else {
          // tail: rows 0..7 of the P tile, this warp's slab of 50 columns (4-byte stores: slab starts are even)
          const int c0 = warp * ATT_SLAB;
          uint32_t r[50];
          tmem_ld_x32(taddr_row + c0, r);
          tmem_ld_x16(taddr_row + c0 + 32, r + 32);
          tmem_ld_x2(taddr_row + c0 + 48, r + 48);
          tmem_wait_ld();
          uint8_t* prow = sP + (lane & 7) * 16;
#pragma unroll
          for (int j = 0; j < 50; j += 2) {
            const float e0 = fast_exp2(__uint_as_float(r[j]) - m);
            const float e1 = fast_exp2(__uint_as_float(r[j + 1]) - m);
            sum += e0 + e1;
            if (lane < 8)
              *reinterpret_cast<uint32_t*>(prow + ((c0 + j) >> 3) * 128 + ((c0 + j) & 7) * 2) = pack_half2(e0, e1);
          }
          if (lane < 8) sm.tsum[warp][lane] = sum;
        }
        TMARK(5);  // pass 2
        // publish P(t) / release S(t) to the control warp
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bar_p);
        prev_t = t;
        prev_win_g = win_g;
      }
    }
    if (prev_t >= 0) {   // the very last tile of this CTA
      mbar_wait(&sm.bar_o, n_o & 1);
      ++n_o;
      __syncwarp();
      tc_fence_after();
      o_epilogue(prev_win_g, prev_t);
    }
#ifdef KVQ_TIMING
    if (lane == 0) {
      for (int k = 0; k < 8; ++k) atomicAdd(&g_attn_timers[k], static_cast<unsigned long long>(tacc[k]));
      atomicAdd(&g_attn_timers[8], 1ull);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================
// generic path: any clamped window
// =====================================================================================================
__global__ void __launch_bounds__(ATT_THREADS, 1) window_attn_kernel(const AttnParams p, const float2* __restrict__ tabs,
                                                                      int tab_len, int rpi_offset) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw_addr & 127u)) & 127u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_IMG_BYTES;
  uint8_t* sV = smem + 2 * ATT_IMG_BYTES;
  uint8_t* sP = smem + ATT_UNIT_BYTES;
  ColMeta* cmeta = reinterpret_cast<ColMeta*>(sP + P_BYTES);                 // [400]
  float* smax = reinterpret_cast<float*>(cmeta + ATT_ROWS);                  // [2][128]
  float* ssum = smax + 256;                                                  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ssum + 256);                  // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  float2* stab = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 64);  // [tab_len]

  const WinGeom& g = p.geom;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hs = warp >> 2;
  const int unit = blockIdx.x;
  const int win_g = unit / p.heads;
  const int head = unit - win_g * p.heads;
  const int win = win_g % g.nW;

  uint64_t* bar_load = &bars[0];
  uint64_t* bar_s = &bars[1];
  uint64_t* bar_o = &bars[2];

  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_fence_init();
    mbar_expect_tx(bar_load, ATT_UNIT_BYTES);
    bulk_load_1d(sQ, reinterpret_cast<const uint8_t*>(p.img) + static_cast<size_t>(unit) * ATT_UNIT_BYTES,
                 ATT_UNIT_BYTES, bar_load);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // per-head bias table {t0, t1}:  bias = t0 + fg * t1   (log2 domain)
  {
    const float4* src = reinterpret_cast<const float4*>(tabs + static_cast<size_t>(head) * tab_len);
    float4* dst = reinterpret_cast<float4*>(stab);
    for (int i = tid; i < tab_len / 2; i += ATT_THREADS) dst[i] = __ldg(src + i);
  }
  // per key-slot metadata from coordinates
  {
    const int wdi = win / (g.nwh * g.nww), whi = (win / g.nww) % g.nwh, wwi = win % g.nww;
    const float fscale_h = static_cast<float>(g.wh) / static_cast<float>(g.Hp);
    const float fscale_w = static_cast<float>(g.ww) / static_cast<float>(g.Wp);
    const int bhw = p.base_wh * p.base_ww;
    const int s1 = (2 * p.base_wh - 1) * (2 * p.base_ww - 1), s2 = 2 * p.base_ww - 1;
    for (int c = tid; c < ATT_ROWS; c += ATT_THREADS) {
      const int d = c / ATT_SLAB, hw = c - d * ATT_SLAB;
      ColMeta m;
      m.base = 0;
      m.pk = 0;
      if (d < g.wd && hw < g.SL) {
        const int th = hw / g.ww, tw = hw - th * g.ww;
        const int i = d * g.SL + hw;
        const int pd = wdi * g.wd + d, ph = whi * g.wh + th, pw = wwi * g.ww + tw;
        const int oh = (ph + g.sh) % g.Hp;
        const int ow = (pw + g.sw) % g.Wp;
        // compute_mask (:560-586) writes slice(-w), slice(-w, -s), slice(-s, None) in that order: the last one wins, which
        // matters when s >= w (adaptive windows keep the base shift)
        const int rd = g.sd == 0 ? 0 : (pd >= g.Dp - g.sd ? 2 : (pd >= g.Dp - g.wd ? 1 : 0));
        const int rh = g.sh == 0 ? 0 : (ph >= g.Hp - g.sh ? 2 : (ph >= g.Hp - g.wh ? 1 : 0));
        const int rw = g.sw == 0 ? 0 : (pw >= g.Wp - g.sw ? 2 : (pw >= g.Wp - g.ww ? 1 : 0));
        int fh = static_cast<int>(floorf(static_cast<float>(oh) * fscale_h)); if (fh > g.wh - 1) fh = g.wh - 1;
        int fw = static_cast<int>(floorf(static_cast<float>(ow) * fscale_w)); if (fw > g.ww - 1) fw = g.ww - 1;
        const int brd = p.rpi_geometric ? d : i / bhw, brh = p.rpi_geometric ? th : (i / p.base_ww) % p.base_wh,
                  brw = p.rpi_geometric ? tw : i % p.base_ww;
        m.base = brd * s1 + brh * s2 + brw;
        m.pk = static_cast<uint32_t>(fh) | (static_cast<uint32_t>(fw) << 8) |
               (static_cast<uint32_t>(9 * rd + 3 * rh + rw) << 16) | (1u << 24);
      }
      cmeta[c] = m;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  mbar_wait(bar_load, 0);
  __syncthreads();

  const int ntiles = (g.N + 127) / 128;
  const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const int col0 = hs * 200;
  const int shifted = p.shifted;

  for (int t = 0; t < ntiles; ++t) {
    const uint32_t ph = static_cast<uint32_t>(t & 1);
    if (tid == 0) {
      tc_fence_after();
      issue_s_mma(tmem_base, smem_u32(sQ) + static_cast<uint32_t>(t) * 8192u, smem_u32(sK));
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    __syncwarp();
    tc_fence_after();

    const int ri = t * 128 + q * 32 + lane;
    const bool row_ok = ri < g.N;
    int base_i = 0, fh_i = 0, fw_i = 0, reg_i = 0;
    if (row_ok) {
      const int sl = ri / g.SL;
      const ColMeta m = cmeta[sl * ATT_SLAB + (ri - sl * g.SL)];
      base_i = m.base + rpi_offset;
      fh_i = m.pk & 255;
      fw_i = (m.pk >> 8) & 255;
      reg_i = (m.pk >> 16) & 255;
    }

    float mx = -INFINITY;
    auto bias_chunk = [&](uint32_t* r, int c_abs, int n) {
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const ColMeta m = cmeta[c_abs + j];
        float v = -INFINITY;
        if (m.pk >> 24) {
          int idx = base_i - m.base;
          idx = idx < 0 ? 0 : (idx >= tab_len ? tab_len - 1 : idx);  // only garbage rows can be out of range
          const float2 e = stab[idx];
          const int fg = __sad(fh_i, static_cast<int>(m.pk & 255), 0) + __sad(fw_i, static_cast<int>((m.pk >> 8) & 255), 0);
          v = __uint_as_float(r[j]) + fmaf(static_cast<float>(fg), e.y, e.x);
          if (shifted && reg_i != static_cast<int>((m.pk >> 16) & 255)) v += MASK_L2;
        }
        r[j] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
    };
#pragma unroll 1
    for (int c = 0; c < 192; c += 32) {
      uint32_t r[32];
      tmem_ld_x32(taddr_row + col0 + c, r);
      tmem_wait_ld();
      bias_chunk(r, col0 + c, 32);
      tmem_st_x32(taddr_row + col0 + c, r);
    }
    {
      uint32_t r[8];
      tmem_ld_x8(taddr_row + col0 + 192, r);
      tmem_wait_ld();
      bias_chunk(r, col0 + 192, 8);
      tmem_st_x8(taddr_row + col0 + 192, r);
    }
    tmem_wait_st();
    smax[hs * 128 + q * 32 + lane] = mx;
    __syncthreads();
    mx = fmaxf(smax[q * 32 + lane], smax[128 + q * 32 + lane]);
    if (!(mx > -INFINITY)) mx = 0.f;  // garbage rows only

    float sum = 0.f;
    uint8_t* prow = sP + (q * 4 + (lane >> 3)) * P_SBO + (lane & 7) * 16;
    auto exp_chunk = [&](const uint32_t* r, int c_abs, int n) {
#pragma unroll
      for (int j = 0; j < n; j += 8) {
        uint32_t h[4];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          const float e0 = fast_exp2(__uint_as_float(r[j + k]) - mx);
          const float e1 = fast_exp2(__uint_as_float(r[j + k + 1]) - mx);
          sum += e0 + e1;
          h[k >> 1] = pack_half2(e0, e1);
        }
        *reinterpret_cast<uint4*>(prow + ((c_abs + j) >> 3) * 128) = make_uint4(h[0], h[1], h[2], h[3]);
      }
    };
#pragma unroll 1
    for (int c = 0; c < 192; c += 32) {
      uint32_t r[32];
      tmem_ld_x32(taddr_row + col0 + c, r);
      tmem_wait_ld();
      exp_chunk(r, col0 + c, 32);
    }
    {
      uint32_t r[8];
      tmem_ld_x8(taddr_row + col0 + 192, r);
      tmem_wait_ld();
      exp_chunk(r, col0 + 192, 8);
    }
    ssum[hs * 128 + q * 32 + lane] = sum;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      issue_pv_mma(tmem_base + TMEM_O_COL, smem_u32(sP), smem_u32(sV));
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    __syncwarp();
    tc_fence_after();
    if (hs == 0) {
      uint32_t r[32];
      tmem_ld_x32(taddr_row + TMEM_O_COL, r);
      tmem_wait_ld();
      if (row_ok) {
        const float inv = 1.0f / (ssum[q * 32 + lane] + ssum[128 + q * 32 + lane]);
        store_o_row(p.out + (static_cast<size_t>(win_g) * g.N + ri) * p.C + head * ATT_HD, r, inv);
      }
    }
    tc_fence_before();
    __syncthreads();
  }

  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// [L, heads] fp32 tables -> [heads][Lp] float2 {t0, t1} (log2 domain) with bias = t0 + fg*t1; for the (8,7,7)
// base window a second copy in the bank-conflict-free fast layout follows at out + heads*Lp
__global__ void pack_bias_kernel(const float* __restrict__ rel, const float* __restrict__ frag, float2* __restrict__ out,
                                 int L, int Lp, int heads, int fast) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_compact = heads * Lp;
  if (i >= n_compact + (fast ? heads * FAST_TAB_LEN : 0)) return;
  int h, e;
  if (i < n_compact) {
    h = i / Lp;
    e = i - h * Lp;
  } else {
    const int k = i - n_compact;
    h = k / FAST_TAB_LEN;
    const int f = k - h * FAST_TAB_LEN;
    const int dd = f / TS_D, rem = f - dd * TS_D, dh = rem / TS_H, dw = rem - dh * TS_H;
    e = (dd < 15 && dh < 13 && dw < 13) ? dd * 169 + dh * 13 + dw : L;
  }
  float2 v = make_float2(0.f, 0.f);
  if (e < L) {
    const float r = rel[e * heads + h];
    if (frag != nullptr) {
      const float f = frag[e * heads + h];
      v = make_float2(f * LOG2E, (r - f) * LOG2E);   // rel*fg + frag*(1-fg) = frag + fg*(rel-frag)
    } else {
      v = make_float2(r * LOG2E, 0.f);
    }
  }
  out[i] = v;
}

}  // namespace

static int compact_len(int bd, int bh, int bw) {
  const int L = (2 * bd - 1) * (2 * bh - 1) * (2 * bw - 1);
  return (L + 1) & ~1;
}

static bool is_fast_window(int bd, int bh, int bw) { return bd == 8 && bh == 7 && bw == 7; }

// float2 entries per head in the packed buffer (compact layout, plus the fast layout for the (8,7,7) window)
int debug_attn_timers(unsigned long long* out16, int reset) {
#ifdef KVQ_TIMING
  if (cudaMemcpyFromSymbol(out16, g_attn_timers, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_attn_timers, z, sizeof(z));
  }
  return 1;
#else
  (void)out16; (void)reset;
  return 0;
#endif
}

int attn_table_len(int bd, int bh, int bw) {
  // float2 entries per head: compact | conflict-free fast layout | paired float4 layout (two float2 per entry)
  return compact_len(bd, bh, bw) + (is_fast_window(bd, bh, bw) ? FAST_TAB_LEN + 2 * ATT3_PAIR_LEN : 0);
}

float attn_qscale() { return 0.17677669529663687f * LOG2E; }  // head_dim^-0.5 (:191), log2 domain

static size_t generic_smem_bytes(int tab_len) {
  return 128 + ATT_UNIT_BYTES + P_BYTES + ATT_ROWS * 8 + 2048 + 64 + static_cast<size_t>(tab_len) * 8;
}

int launch_pack_bias(const float* rel, const float* frag, float* out, int bd, int bh, int bw, int heads,
                     cudaStream_t stream) {
  const int L = (2 * bd - 1) * (2 * bh - 1) * (2 * bw - 1);
  const int Lp = compact_len(bd, bh, bw);
  const int fast = is_fast_window(bd, bh, bw) ? 1 : 0;
  const int n = heads * (Lp + (fast ? FAST_TAB_LEN : 0));
  pack_bias_kernel<<<(n + 255) / 256, 256, 0, stream>>>(rel, frag, reinterpret_cast<float2*>(out), L, Lp, heads, fast);
  count_launch();
  KVQ_CUDA(cudaGetLastError());
  if (fast) return launch_pack_bias_pair(rel, frag, out + static_cast<size_t>(n) * 2, heads, stream);
  return KVQ_OK;
}

static int resolve_variant(int variant) {
  // 6 (default) = third generation (kvq_attn3.cu: one-pass softmax, three tile slots per SM) for full (8,7,7)
  // windows; 5 = two-CTA flash-style kernel (kvq_attn2.cu); tuning knob
  // KVQ_ATTN_VARIANT: 1 = first-generation persistent kernel, 2 = generic kernel everywhere
  if (variant != 0) return variant;
  static int env_variant = -1;
  if (env_variant < 0) {
    const char* e = getenv("KVQ_ATTN_VARIANT");
    env_variant = e ? atoi(e) : 6;
    if (env_variant == 0) env_variant = 6;
  }
  return env_variant;
}

static bool full_window(const WinGeom& g, const int base_win[3]) {
  return g.wd == 8 && g.wh == 7 && g.ww == 7 && base_win[0] == 8 && base_win[1] == 7 && base_win[2] == 7 &&
         (g.sd == 0 || g.sd == 4) && (g.sh == 0 || g.sh == 3) && (g.sw == 0 || g.sw == 3);
}

int window_attn_pitch(const WinGeom& g, const int base_win[3], int variant) {
  const int v = resolve_variant(variant);
  if (!full_window(g, base_win)) return ATT_SLAB;
  return v == 6 ? ATT3_PITCH : v == 5 ? ATT2_PITCH : ATT_SLAB;
}

int launch_window_attn(const AttnParams& p_in, cudaStream_t stream) {
  AttnParams p = p_in;
  p.variant = resolve_variant(p.variant);
  const WinGeom& g = p.geom;
  {
    const int bw[3] = {p.base_wd, p.base_wh, p.base_ww};
    if (p.variant == 6 && full_window(g, bw) && p.heads <= num_sms()) return launch_window_attn3(p, stream);
    if (p.variant == 5 && full_window(g, bw) && 2 * p.heads <= 2 * num_sms()) return launch_window_attn2(p, stream);
  }
  KVQ_REQUIRE(p.C == p.heads * ATT_HD, KVQ_ERR_BAD_SHAPE, "attn: C=%d must be heads(%d) x 32", p.C, p.heads);
  KVQ_REQUIRE(g.wd * ATT_SLAB <= ATT_ROWS && g.SL <= ATT_SLAB - 1 && g.wd <= p.base_wd && g.wh <= p.base_wh &&
                  g.ww <= p.base_ww,
              KVQ_ERR_BAD_SHAPE, "attn: window (%d,%d,%d) exceeds the 8 x 49 key layout", g.wd, g.wh, g.ww);
  KVQ_REQUIRE(g.Hp < 256 && g.Wp < 256, KVQ_ERR_BAD_SHAPE, "attn: geometry out of range");
  const long long units = static_cast<long long>(p.B) * g.nW * p.heads;
  KVQ_REQUIRE(units > 0 && units < (1ll << 31), KVQ_ERR_BAD_SHAPE, "attn: %lld units", units);
  const int tab_len = compact_len(p.base_wd, p.base_wh, p.base_ww);

  const bool fast = p.variant != 2 && g.wd == 8 && g.wh == 7 && g.ww == 7 && p.base_wd == 8 && p.base_wh == 7 &&
                    p.base_ww == 7 && (g.sd == 0 || g.sd == 4) && (g.sh == 0 || g.sh == 3) &&
                    (g.sw == 0 || g.sw == 3) && p.heads <= num_sms();
  if (fast) {
    static bool attr = false;
    if (!attr) {
      KVQ_CUDA(cudaFuncSetAttribute(window_attn_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_SMEM));
      attr = true;
    }
    int grid = num_sms() / p.heads * p.heads;   // multiple of heads so a CTA keeps one head's table
    if (grid > units) grid = static_cast<int>(units);   // units is a multiple of heads too
    window_attn_fast_kernel<<<grid, FAST_THREADS, FAST_SMEM, stream>>>(
        p, reinterpret_cast<const float2*>(p.packed_tab), static_cast<int>(units));
    count_launch();
    return check_cuda(cudaGetLastError(), "window_attn_fast_kernel launch");
  }

  const size_t smem = generic_smem_bytes(tab_len);
  KVQ_REQUIRE(smem <= 227 * 1024, KVQ_ERR_BAD_SHAPE, "attn: bias table of %d entries does not fit shared memory",
              tab_len);
  static size_t attr_smem = 0;
  if (attr_smem < smem) {
    KVQ_CUDA(cudaFuncSetAttribute(window_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    attr_smem = smem;
  }
  const int s1 = (2 * p.base_wh - 1) * (2 * p.base_ww - 1), s2 = 2 * p.base_ww - 1;
  const int rpi_offset = (p.base_wd - 1) * s1 + (p.base_wh - 1) * s2 + (p.base_ww - 1);
  window_attn_kernel<<<static_cast<unsigned>(units), ATT_THREADS, smem, stream>>>(
      p, reinterpret_cast<const float2*>(p.packed_tab), tab_len, rpi_offset);
  count_launch();
  return check_cuda(cudaGetLastError(), "window_attn_kernel launch");
}

}  // namespace kvq
