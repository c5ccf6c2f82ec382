// Fused 3-D window attention, second generation for the (8,7,7) window: TWO CTAs per SM, flash-style.
//
// Same mathematics as kvq_attn.cu (WindowAttention3D.forward, swin_backbone.py:245-326) but organised so that the
// table-gather/FMA phase of one CTA overlaps the exp2 (MUFU) phase and the tensor-pipe waits of the other:
//   * a CTA = 4 softmax warps (one thread per query row of a 128-row tile, no cross-warp exchange) + 1 control warp
//   * keys are walked in two blocks of 208 slots (4 temporal slabs x 52-slot pitch) with an online softmax:
//         S = Q K_blk^T  (tcgen05.mma M128 x N208 x K32, fp32 in TMEM)
//         v = S + GRPB bias (+ -100 region mask); running row max; P = exp2(v - max) written fp16 *into TMEM over S*
//         O (+)= P V_blk  (tcgen05.mma with the A operand read from TMEM); second block rescales O by exp2(m_A - m)
//   * TMEM: 208 (S/P) + 32 (O) = 240 of the CTA's 256 columns, so two CTAs share the SM's 512
//   * shared memory: K | V images (2 x 26 624 B), double-buffered 8 KB Q tiles, 34.7 KB conflict-free bias table
//     = 106.5 KB per CTA
// Operand images use the 52-slot key pitch (416 rows): unit = Q 25 600 B | K 26 624 B | V 26 624 B.
#include <cstdlib>

#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int A2_THREADS = 160;
constexpr int BLK = 4 * ATT2_PITCH;              // 208 key slots per block
constexpr int KSTEPS = BLK / 16;                 // 13
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASK_L2 = -100.0f * LOG2E;
constexpr int TS_H = 23, TS_D = 289;             // conflict-free table strides (see kvq_attn.cu)
constexpr int TAB_LEN = 15 * TS_D + 1;           // 4336 float2 entries per head
constexpr int COMPACT_LEN = 2536;

constexpr int S2_K = 0;
constexpr int S2_V = ATT2_KV_BYTES;
constexpr int S2_Q = 2 * ATT2_KV_BYTES;          // 2 x 8192
constexpr int S2_SMALL = S2_Q + 2 * 8192;        // 6144 B
constexpr int S2_TAB = S2_SMALL + 6144;
constexpr int S2_SMEM = S2_TAB + TAB_LEN * 8 + 128;
constexpr int T_S = 0;                           // TMEM columns
constexpr int T_O = BLK;

struct Small2 {
  float fhc[8], fwc[8], rhm[8], rwm[8], rdm[8];
  uint64_t bar_tab, bar_kv, bar_q[2], bar_qe[2], bar_s, bar_p, bar_c, bar_o;
  uint32_t tmem_slot;
  float tmax[2][4][8];          // tail tile: per-block local maxima [block][warp][row]
  float tl[4][8];               // tail tile: per-warp partial row sums
  alignas(16) float to[4][8][32];   // tail tile: per-warp partial outputs
};
static_assert(sizeof(Small2) <= 6144, "Small2 overflows its slot");

__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// pass 1 over one 52-column slab (49 keys + 3 pad slots): v = s + t0[idx] + fg*t1[idx] (+ mask); running max;
// write back.  Table gathers run one key row ahead of the arithmetic.
template <bool MASK>
__device__ __forceinline__ void pass1_slab52(uint32_t taddr, uint32_t trow_addr, const float (&Ah)[7],
                                             const float (&Aw)[7], float m00, float m01, float m10, float m11,
                                             float (&mx)[4]) {
  uint32_t r[52];
  tmem_ld_x32(taddr, r);
  tmem_ld_x16(taddr + 32, r + 32);
  tmem_ld_x4(taddr + 48, r + 48);
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
  r[49] = r[50] = r[51] = __float_as_uint(-INFINITY);
  tmem_st_x32(taddr, r);
  tmem_st_x16(taddr + 32, r + 32);
  tmem_st_x4(taddr + 48, r + 48);
}

__global__ void __launch_bounds__(A2_THREADS, 2)
window_attn2_kernel(const AttnParams p, const float2* __restrict__ tabs, int units) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw_addr & 127u)) & 127u);
  Small2& sm = *reinterpret_cast<Small2*>(smem + S2_SMALL);
  float2* stab = reinterpret_cast<float2*>(smem + S2_TAB);

  const WinGeom& g = p.geom;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;                     // multiple of heads
  const int head = blockIdx.x % p.heads;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.img);

  auto load_kv = [&](int unit) {
    const uint8_t* src = img + static_cast<size_t>(unit) * ATT2_UNIT_BYTES + ATT2_Q_BYTES;
    mbar_expect_tx(&sm.bar_kv, 2 * ATT2_KV_BYTES);
    bulk_load_1d(smem + S2_K, src, 2 * ATT2_KV_BYTES, &sm.bar_kv);
  };
  auto load_q = [&](int unit, int t, int b) {
    const uint8_t* src = img + static_cast<size_t>(unit) * ATT2_UNIT_BYTES;
    if (t < 3) {
      mbar_expect_tx(&sm.bar_q[b], 8192);
      bulk_load_1d(smem + S2_Q + b * 8192, src + t * 8192, 8192, &sm.bar_q[b]);
    } else {   // the 8 tail rows: one 512 B row group; the MMA descriptor replicates it with SBO = 0
      mbar_expect_tx(&sm.bar_q[b], 512);
      bulk_load_1d(smem + S2_Q + b * 8192, src + 48 * 512, 512, &sm.bar_q[b]);
    }
  };

  if (tid == 128) {
    mbar_init(&sm.bar_tab, 1);
    mbar_init(&sm.bar_kv, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.bar_q[i], 1);
      mbar_init(&sm.bar_qe[i], 1);
    }
    mbar_init(&sm.bar_s, 1);
    mbar_init(&sm.bar_p, 4);
    mbar_init(&sm.bar_c, 1);
    mbar_init(&sm.bar_o, 1);
    mbar_fence_init();
    // the bias table is a weight: fetch it while the previous kernel (the QKV GEMM) is still draining ...
    mbar_expect_tx(&sm.bar_tab, TAB_LEN * 8);
    bulk_load_1d(stab, tabs + static_cast<size_t>(p.heads) * COMPACT_LEN + static_cast<size_t>(head) * TAB_LEN,
                 TAB_LEN * 8, &sm.bar_tab);
    pdl_launch_dependents();
    pdl_wait();   // ... the operand images are its output
    if (static_cast<int>(blockIdx.x) < units) {
      load_kv(blockIdx.x);
      load_q(blockIdx.x, 0, 0);
    }
  }
  if (warp == 1) {
    tmem_alloc(&sm.tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;
  pdl_wait();   // every thread: global stores of this kernel must not pass the previous grid's completion

  if (warp == 4) {
    // =============================== control warp ===============================
    if (lane == 0) {
      const uint32_t aK = smem_u32(smem + S2_K), aV = smem_u32(smem + S2_V), aQ = smem_u32(smem + S2_Q);
      constexpr uint32_t idesc_s = umma_idesc_f16(128, BLK, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 32, 0, 1);
      uint32_t gq = 0, n_kv = 0, n_p = 0, n_c = 0, n_o = 0;
      auto issue_s = [&](int b, bool tail, int blk) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t dq = umma_smem_desc(aQ + b * 8192 + ks * 256, 128, tail ? 0 : 512, UMMA_SW_NONE);
          const uint64_t dk = umma_smem_desc(aK + blk * (BLK / 8) * 512 + ks * 256, 128, 512, UMMA_SW_NONE);
          umma_f16_ss(tmem_base + T_S, dq, dk, idesc_s, ks);
        }
      };
      auto issue_pv = [&](int blk) {
#pragma unroll 1
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint64_t dv = umma_smem_desc(aV + (blk * KSTEPS + ks) * 1024, 512, 128, UMMA_SW_NONE);
          umma_f16_ts(tmem_base + T_O, tmem_base + T_S + 8 * ks, dv, idesc_pv, (blk > 0 || ks > 0) ? 1u : 0u);
        }
      };
      for (int unit = blockIdx.x; unit < units; unit += G) {
        const bool has_next = unit + G < units;
        mbar_wait(&sm.bar_kv, n_kv & 1);
        for (int t = 0; t < 4; ++t, ++gq) {
          const int b = gq & 1;
          mbar_wait(&sm.bar_q[b], (gq >> 1) & 1);
          tc_fence_after();
          issue_s(b, t == 3, 0);
          umma_commit(&sm.bar_s);
          // prefetch the next Q tile (next tile of this unit, or tile 0 of this CTA's next unit)
          if (t < 3 || has_next) {
            const int nb = (gq + 1) & 1;
            if (gq >= 1) mbar_wait(&sm.bar_qe[nb], (((gq + 1) >> 1) - 1) & 1);   // S_B of its previous user is done
            load_q(t < 3 ? unit : unit + G, t < 3 ? t + 1 : 0, nb);
          }
          mbar_wait(&sm.bar_p, n_p & 1); ++n_p;        // block A: P written over S
          tc_fence_after();
          issue_pv(0);
          umma_commit(&sm.bar_c);
          mbar_wait(&sm.bar_c, n_c & 1); ++n_c;        // PV_A has consumed P before S_B overwrites it
          tc_fence_after();
          issue_s(b, t == 3, 1);
          umma_commit(&sm.bar_s);
          umma_commit(&sm.bar_qe[b]);                  // Q tile buffer free once S_B is done
          mbar_wait(&sm.bar_p, n_p & 1); ++n_p;        // block B: P written, O rescaled
          tc_fence_after();
          issue_pv(1);
          umma_commit(&sm.bar_o);
          mbar_wait(&sm.bar_o, n_o & 1); ++n_o;        // PV_B done before the next tile's S_A reuses the columns
        }
        ++n_kv;
        if (has_next) load_kv(unit + G);               // all MMAs of this unit are complete: K | V are free
      }
    }
  } else {
    // =============================== softmax warps: one thread per query row ===============================
    const int q = warp;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t n_s = 0, n_o = 0;
    mbar_wait(&sm.bar_tab, 0);
    for (int unit = blockIdx.x; unit < units; unit += G) {
      const int win_g = unit / p.heads;
      const int win = win_g % g.nW;
      const int wdi = win / (g.nwh * g.nww), whi = (win / g.nww) % g.nwh, wwi = win % g.nww;
      named_bar_sync(1, 128);      // every warp is done with the previous unit's coordinate tables
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
      const bool masked = (g.sh != 0 && whi == g.nwh - 1) || (g.sw != 0 && wwi == g.nww - 1) ||
                          (g.sd != 0 && wdi == g.nwd - 1);
      named_bar_sync(1, 128);

#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        const bool tail = (t == 3);
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
          hlo = (sm.rhm[0] != sm.rhm[h_i]) ? MASK_L2 : 0.f;
          hhi = (sm.rhm[6] != sm.rhm[h_i]) ? MASK_L2 : 0.f;
          wlo = (sm.rwm[0] != sm.rwm[w_i]) ? MASK_L2 : 0.f;
          whi_m = (sm.rwm[6] != sm.rwm[w_i]) ? MASK_L2 : 0.f;
          rd_i = sm.rdm[d_i];
        }
        const uint32_t trow0 = smem_u32(stab) + 8u * static_cast<uint32_t>((d_i + 7) * TS_D + (h_i + 6) * TS_H + (w_i + 6));

        float m_run = -INFINITY, l_run = 0.f;
        if (!tail) {
#pragma unroll 1
          for (int blk = 0; blk < 2; ++blk) {
            mbar_wait(&sm.bar_s, n_s & 1);
            ++n_s;
            __syncwarp();
            tc_fence_after();
            // ---- pass 1: bias (+ mask), block max, write back ----
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
            for (int s = 0; s < 4; ++s) {
              const int d = blk * 4 + s;
              const uint32_t ta = trow + T_S + s * ATT2_PITCH;
              if (masked) {
                const float dm = (sm.rdm[d] != rd_i) ? MASK_L2 : 0.f;
                pass1_slab52<true>(ta, trow0 - 8u * (d * TS_D), Ah, Aw, fminf(fminf(hlo, wlo), dm),
                                   fminf(fminf(hlo, whi_m), dm), fminf(fminf(hhi, wlo), dm),
                                   fminf(fminf(hhi, whi_m), dm), mx);
              } else {
                pass1_slab52<false>(ta, trow0 - 8u * (d * TS_D), Ah, Aw, 0.f, 0.f, 0.f, 0.f, mx);
              }
            }
            tmem_wait_st();
            const float m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            const float m_new = fmaxf(m_run, m_blk);
            const float alpha = fast_exp2(m_run - m_new);          // 0 for the first block (m_run = -inf)
            // ---- pass 2: P = exp2(v - m) packed fp16, written into TMEM over the columns already consumed ----
            const float2 negm2 = make_float2(-m_new, -m_new);
            float2 sum2 = make_float2(0.f, 0.f);
            // TMEM loads run one chunk ahead of the exp2 / pack / store of the previous chunk
            auto exp_pack = [&](const uint32_t* r, uint32_t* h, int n) {
#pragma unroll
              for (int k = 0; k < n; k += 2) {
                const float2 dd = fadd2(make_float2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), negm2);
                const float2 ee = make_float2(fast_exp2(dd.x), fast_exp2(dd.y));
                sum2 = fadd2(sum2, ee);
                h[k >> 1] = pack_half2(ee.x, ee.y);
              }
            };
            {
              uint32_t ra[32], rb[32], h[16];
              tmem_ld_x32(trow + T_S, ra);
#pragma unroll 1
              for (int c = 0; c < 192; c += 64) {
                tmem_wait_ld();
                tmem_ld_x32(trow + T_S + c + 32, rb);
                exp_pack(ra, h, 32);
                tmem_st_x16(trow + T_S + (c >> 1), h);
                tmem_wait_ld();
                if (c + 64 < 192) tmem_ld_x32(trow + T_S + c + 64, ra);
                else tmem_ld_x16(trow + T_S + 192, ra);
                exp_pack(rb, h, 32);
                tmem_st_x16(trow + T_S + ((c + 32) >> 1), h);
              }
              tmem_wait_ld();
              exp_pack(ra, h, 16);
              tmem_st_x8(trow + T_S + 96, h);
            }
            l_run = fmaf(l_run, alpha, sum2.x + sum2.y);
            if (blk == 1) {
              // O currently holds P_A V_A computed against m_A: bring it to the new max before PV_B accumulates
              uint32_t o[32];
              tmem_ld_x32(trow + T_O, o);
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
              tmem_st_x32(trow + T_O, o);
            }
            m_run = m_new;
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.bar_p);
          }
          // ---- epilogue: O / l -> global ----
          mbar_wait(&sm.bar_o, n_o & 1);
          ++n_o;
          __syncwarp();
          tc_fence_after();
          uint32_t o[32];
          tmem_ld_x32(trow + T_O, o);
          tmem_wait_ld();
          const float inv = 1.0f / l_run;
          __half* dst = p.out + (static_cast<size_t>(win_g) * 392 + ri) * p.C + head * ATT_HD;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 v;
            v.x = pack_half2(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
            v.y = pack_half2(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
            v.z = pack_half2(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
            v.w = pack_half2(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + j) = v;
          }
          tc_fence_before();
        } else {
          // ================= tail tile: the 8 rows 384..391 are replicated in every lane group (SBO = 0), so the
          // four warps split the KEYS instead: warp q owns slab q of each block, writes zeros into the P columns of
          // the other slabs, and the four partial outputs / row sums are added through shared memory ==============
#pragma unroll 1
          for (int blk = 0; blk < 2; ++blk) {
            mbar_wait(&sm.bar_s, n_s & 1);
            ++n_s;
            __syncwarp();
            tc_fence_after();
            const int d = blk * 4 + q;
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            const uint32_t ta = trow + T_S + q * ATT2_PITCH;
            if (masked) {
              const float dm = (sm.rdm[d] != rd_i) ? MASK_L2 : 0.f;
              pass1_slab52<true>(ta, trow0 - 8u * (d * TS_D), Ah, Aw, fminf(fminf(hlo, wlo), dm),
                                 fminf(fminf(hlo, whi_m), dm), fminf(fminf(hhi, wlo), dm), fminf(fminf(hhi, whi_m), dm),
                                 mx);
            } else {
              pass1_slab52<false>(ta, trow0 - 8u * (d * TS_D), Ah, Aw, 0.f, 0.f, 0.f, 0.f, mx);
            }
            tmem_wait_st();
            if (lane < 8) sm.tmax[blk][q][lane] = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            named_bar_sync(1, 128);
            const int rl = lane & 7;
            const float m_blk = fmaxf(fmaxf(sm.tmax[blk][0][rl], sm.tmax[blk][1][rl]),
                                      fmaxf(sm.tmax[blk][2][rl], sm.tmax[blk][3][rl]));
            const float m_new = fmaxf(m_run, m_blk);
            const float alpha = fast_exp2(m_run - m_new);
            // own slab -> registers -> probabilities
            uint32_t r[52], h[26];
            tmem_ld_x32(ta, r);
            tmem_ld_x16(ta + 32, r + 32);
            tmem_ld_x4(ta + 48, r + 48);
            tmem_wait_ld();
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 52; k += 2) {
              const float e0 = fast_exp2(__uint_as_float(r[k]) - m_new), e1 = fast_exp2(__uint_as_float(r[k + 1]) - m_new);
              sum += e0 + e1;
              h[k >> 1] = pack_half2(e0, e1);
            }
            // zero the whole P block of this lane group, then drop the own slab's 26 packed columns in
            {
              uint32_t z[32];
#pragma unroll
              for (int k = 0; k < 32; ++k) z[k] = 0u;
              tmem_st_x32(trow + T_S, z);
              tmem_st_x32(trow + T_S + 32, z);
              tmem_st_x32(trow + T_S + 64, z);
              tmem_st_x8(trow + T_S + 96, z);
              tmem_wait_st();
              tmem_st_x16(trow + T_S + 26 * q, h);
              tmem_st_x8(trow + T_S + 26 * q + 16, h + 16);
              tmem_st_x2(trow + T_S + 26 * q + 24, h + 24);
            }
            l_run = fmaf(l_run, alpha, sum);
            if (blk == 1) {
              uint32_t o[32];
              tmem_ld_x32(trow + T_O, o);
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
              tmem_st_x32(trow + T_O, o);
            }
            m_run = m_new;
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.bar_p);
          }
          mbar_wait(&sm.bar_o, n_o & 1);
          ++n_o;
          __syncwarp();
          tc_fence_after();
          {
            uint32_t o[32];
            tmem_ld_x32(trow + T_O, o);
            tmem_wait_ld();
            if (lane < 8) {
              sm.tl[q][lane] = l_run;
#pragma unroll
              for (int k = 0; k < 32; k += 4)
                *reinterpret_cast<uint4*>(&sm.to[q][lane][k]) = make_uint4(o[k], o[k + 1], o[k + 2], o[k + 3]);
            }
          }
          tc_fence_before();
          named_bar_sync(1, 128);
          if (q == 0 && lane < 8) {
            const float inv = 1.0f / (sm.tl[0][lane] + sm.tl[1][lane] + sm.tl[2][lane] + sm.tl[3][lane]);
            __half* dst = p.out + (static_cast<size_t>(win_g) * 392 + 384 + lane) * p.C + head * ATT_HD;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float v[8];
#pragma unroll
              for (int k = 0; k < 8; ++k)
                v[k] = (sm.to[0][lane][j + k] + sm.to[1][lane][j + k] + sm.to[2][lane][j + k] + sm.to[3][lane][j + k]) * inv;
              uint4 pk;
              pk.x = pack_half2(v[0], v[1]);
              pk.y = pack_half2(v[2], v[3]);
              pk.z = pack_half2(v[4], v[5]);
              pk.w = pack_half2(v[6], v[7]);
              *reinterpret_cast<uint4*>(dst + j) = pk;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

int launch_window_attn2(const AttnParams& p, cudaStream_t stream) {
  const WinGeom& g = p.geom;
  KVQ_REQUIRE(g.wd == 8 && g.wh == 7 && g.ww == 7 && p.base_wd == 8 && p.base_wh == 7 && p.base_ww == 7,
              KVQ_ERR_BAD_SHAPE, "attn2: built for the (8,7,7) window");
  KVQ_REQUIRE(p.C == p.heads * ATT_HD, KVQ_ERR_BAD_SHAPE, "attn2: C=%d must be heads(%d) x 32", p.C, p.heads);
  const long long units = static_cast<long long>(p.B) * g.nW * p.heads;
  KVQ_REQUIRE(units > 0 && units < (1ll << 31), KVQ_ERR_BAD_SHAPE, "attn2: %lld units", units);
  static bool attr = false;
  if (!attr) {
    KVQ_CUDA(cudaFuncSetAttribute(window_attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
    attr = true;
  }
  int grid = 2 * num_sms() / p.heads * p.heads;   // two CTAs per SM; multiple of heads so a CTA keeps its table
  if (grid > units) grid = static_cast<int>(units);
  count_launch();
  return launch_pdl(window_attn2_kernel, dim3(grid), dim3(A2_THREADS), S2_SMEM, stream, p,
                    reinterpret_cast<const float2*>(p.packed_tab), static_cast<int>(units));
}

}  // namespace kvq
