// Fused 3-D window attention, third generation for the (8,7,7) window (WindowAttention3D.forward,
// swin_backbone.py:245-326; compute_mask :560-586; global_position_index :22-50).
//
// The second generation (kvq_attn2.cu) was latency-bound in its scalar softmax path: two TMEM round trips per logit,
// MMA and softmax serialised inside a CTA, eight softmax warps per SM.  This kernel is organised around the per-logit
// instruction count and around keeping the LDS, MUFU and tensor pipes busy at the same time:
//   * ONE persistent CTA per SM, 16 warps: warp 0 loads (cp.async.bulk), warps 1..3 each issue the tcgen05.mma of one
//     tile slot, warps 4..15 are three softmax warpgroups (one per slot; thread = query row = TMEM lane).
//     The three slots work on the three 128-row tiles of the same (window, head) unit, sharing its K | V image.
//   * keys are ordered (h, w, d) with d fastest: slot = h*56 + w*8 + d, one 56-key chunk per window row h.  A chunk is
//     S = Q K_c^T (M128 x N64 x K32: the 8 columns past the chunk are the next chunk's keys and are never read),
//     double-buffered in TMEM, read ONCE into registers; bias, exp2 and the fp16 pack happen there in a single fused
//     pass, P_c is written in place over S_c and O += P_c V_c is a TS-form MMA (K = 64: the 8 pad columns of P are 0).
//   * the two temporal neighbours (d = 2k, 2k+1) of a key position sit in adjacent TMEM columns and share the GRPB
//     gate fg = |dfh| + |dfw| (it does not depend on d), so the bias is ONE LDS.128 + FFMA2 + FADD2 per two logits:
//     table entry {t0[e], t0[e-1], t1[e], t1[e-1]} with e = d_i - 2k.  Query rows use the same (h,w,d) order, so a
//     quarter-warp holds the 8 frames of one cell and its LDS.128 is conflict-free for any odd e-stride (169).
//   * lazy row max (as in FlashAttention-4): P = exp2(v - m_ref) with m_ref the exact max of the row's FIRST chunk;
//     later chunks run the fused pass against that m_ref and only when a logit exceeds it by more than 15 (P would
//     leave fp16 range) is the chunk redone in two passes, m_ref raised and O, l rescaled.  The result is exact up to
//     rounding because the row sum uses the same m_ref.
//   * the SW-MSA region mask is folded into QK^T: a third K step contracts [a_d 1 a_h 1 a_w 1] (query region bits,
//     zeroed for dims this window does not mask) with (-100 log2e / 2) * [(1-2b_d) b_d (1-2b_h) b_h (1-2b_w) b_w]
//     (key region bits, static); the second 8-half K chunk aliases the first (LBO = 0), hence the factor 1/2.
//     -100 per differing dim instead of -100 once: exp(-100) and exp(-300) are both 0 next to an unmasked logit.
//   * the 8 tail rows (384..391) are replicated into all lane groups (SBO = 0, as in the second generation); warp q
//     takes the temporal pair d = 2q, 2q+1 of every key position and writes zero P elsewhere, so all four warps work
//     on every tail chunk; the four partial (m, l, O) are merged in smem.
//   * the MMA issuers run warp-uniform code with uniform operands (TMEM base 0, descriptors from uniform offsets) and
//     one elected lane: a per-lane `if (lane == 0)` issue path makes ptxas emit an R2UR waterfall of ~130 cycles per
//     MMA (tools/ubench/mma_lat.cu: 27 cycles for a TS N=32 MMA, 75 for an SS N=64 one when issued uniformly).
// TMEM per slot: S/P buffer 0 (64) | S/P buffer 1 (64) | O (32) = 160 columns, 480 of 512.
#include <cstdlib>

#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int A3_THREADS = 512;
constexpr int NSLOT = 3;
constexpr int NCHUNK3 = 7;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float TAU = 15.0f;                      // lazy-max slack (log2 units): P <= 2^15 stays inside fp16

// shared memory map (bytes)
constexpr int S3_TAB = 0;
constexpr int S3_TAB_BYTES = (ATT3_PAIR_LEN * 16 + 255) / 256 * 256;        // 45 056
constexpr int S3_KV = S3_TAB + S3_TAB_BYTES;                               // 2 x (K | V)
constexpr int S3_Q = S3_KV + 2 * 2 * ATT3_KV_BYTES;                        // 3 slots x 2 x 8192
constexpr int S3_KAUG = S3_Q + NSLOT * 2 * 8192;                           // 400 x 16
constexpr int S3_QAUG = S3_KAUG + ATT3_KV_ROWS * 16;                       // 3 slots x 128 x 16
constexpr int S3_SCR = S3_QAUG + NSLOT * 2048;                             // 3 slots x tail merge scratch
constexpr int SCR_BYTES = 3328;                                            // 3 warps x 8 rows x 34 floats = 3264
constexpr int S3_BARS = S3_SCR + NSLOT * SCR_BYTES;                        // 512
constexpr int S3_SMEM = S3_BARS + 512 + 128;
static_assert(S3_SMEM <= 227 * 1024, "attn3 shared memory exceeds the SM");

// TMEM columns of a slot
constexpr int T3_SLOT = 160, T3_O = 128;

struct Bars3 {
  uint64_t tab, kv[2], kvfree[2];
  uint64_t q[NSLOT][2], qfree[NSLOT][2];
  uint64_t s[NSLOT][2], p[NSLOT][2], pv[NSLOT][2], of[NSLOT], qa[NSLOT];
  uint32_t tmem_slot;
};
static_assert(sizeof(Bars3) <= 512, "Bars3 overflows its slot");

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
template <int REGS>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

// fragment coordinate of window-local row/col k (global_position_index :22-50: nearest-resized fragment index of the
// rolled position); inv = 7 / size
__device__ __forceinline__ float frag_coord(int base, int k, int shift, int size, float inv) {
  int o = base + k + shift;
  if (o >= size) o -= size;
  int f = static_cast<int>(floorf(static_cast<float>(o) * inv));
  return static_cast<float>(f > 6 ? 6 : f);
}

struct UnitInfo {
  int win_g, whi, wwi;
  bool md, mh, mw;     // dims whose SW-MSA region mask applies in this window
};
__device__ __forceinline__ UnitInfo unit_info(const AttnParams& p, int unit) {
  const WinGeom& g = p.geom;
  UnitInfo u;
  u.win_g = unit / p.heads;
  const int win = u.win_g % g.nW;
  const int wdi = win / (g.nwh * g.nww);
  u.whi = (win / g.nww) % g.nwh;
  u.wwi = win % g.nww;
  u.md = g.sd != 0 && wdi == g.nwd - 1;
  u.mh = g.sh != 0 && u.whi == g.nwh - 1;
  u.mw = g.sw != 0 && u.wwi == g.nww - 1;
  return u;
}

// walks the (unit, item, chunk) sequence of one slot: slot s takes tile s of every unit and the tail of units n = s (mod 3)
struct Cursor {
  int n, it, c;
  uint32_t k;          // item ordinal of this slot
  __device__ __forceinline__ void advance(int s) {
    if (++c == NCHUNK3) {
      c = 0;
      ++k;
      if (++it == ((n % NSLOT == s) ? 2 : 1)) {
        it = 0;
        ++n;
      }
    }
  }
};

// bias for one key position (8 logits = 4 temporal pairs): r <- s + t0 + fg * t1; running max
__device__ __forceinline__ void bias_pos(uint32_t (&r)[56], const int wj, const float4 (&e)[4], const float2 fg2,
                                         float& gmax) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const float4 t = e[kk];
    const int j = wj * 8 + 2 * kk;
    const float2 v = fadd2(ffma2(fg2, make_float2(t.z, t.w), make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]))),
                           make_float2(t.x, t.y));
    gmax = fmax3(gmax, v.x, v.y);
    r[j] = __float_as_uint(v.x);
    r[j + 1] = __float_as_uint(v.y);
  }
}

__global__ void __launch_bounds__(A3_THREADS, 1)
window_attn3_kernel(const AttnParams p, const float4* __restrict__ tabs, int units) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw_addr & 127u)) & 127u);
  Bars3& bars = *reinterpret_cast<Bars3*>(smem + S3_BARS);

  const WinGeom& g = p.geom;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;                     // multiple of heads: a CTA keeps one head's table
  const int head = blockIdx.x % p.heads;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.img);

  if (tid == 0) {
    mbar_init(&bars.tab, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.kv[i], 1);
      mbar_init(&bars.kvfree[i], NSLOT);
    }
    for (int s = 0; s < NSLOT; ++s) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars.q[s][i], 1);
        mbar_init(&bars.qfree[s][i], 1);
        mbar_init(&bars.s[s][i], 1);
        mbar_init(&bars.p[s][i], 4);
        mbar_init(&bars.pv[s][i], 1);
      }
      mbar_init(&bars.of[s], 4);
      mbar_init(&bars.qa[s], 4);
    }
    mbar_fence_init();
    // the bias table is a weight: fetch it while the previous kernel (the QKV GEMM) may still be draining
    mbar_expect_tx(&bars.tab, ATT3_PAIR_LEN * 16);
    bulk_load_1d(smem + S3_TAB, tabs + static_cast<size_t>(head) * ATT3_PAIR_LEN, ATT3_PAIR_LEN * 16, &bars.tab);
  }
  // static operand: key side of the folded region mask
  for (int r = tid; r < ATT3_KV_ROWS; r += A3_THREADS) {
    const int hj = r / 56, wj = (r % 56) >> 3, dj = r & 7;
    const float mh = -50.0f * LOG2E;           // half of -100 log2(e): the aliased second K chunk doubles it
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (r < 392) {
      const float bd = dj >= 4 ? 1.f : 0.f, bh = hj >= 4 ? 1.f : 0.f, bw = wj >= 4 ? 1.f : 0.f;
      w.x = pack_half2(mh * (1.f - 2.f * bd), mh * bd);
      w.y = pack_half2(mh * (1.f - 2.f * bh), mh * bh);
      w.z = pack_half2(mh * (1.f - 2.f * bw), mh * bw);
    }
    *reinterpret_cast<uint4*>(smem + S3_KAUG + r * 16) = w;
  }
  fence_proxy_async_smem();
  if (warp == 1) {
    tmem_alloc(&bars.tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_slot;
  if (tmem_base != 0) {   // one CTA per SM owns all 512 columns: the MMA issuers rely on base 0 (uniform operands)
    if (tid == 0) printf("kvq attn3: unexpected TMEM base %u\n", tmem_base);
    __trap();
  }
  pdl_wait();   // the operand images are the previous kernel's output; nothing is written before this point

  const int n_units = (units - static_cast<int>(blockIdx.x) + G - 1) / G;   // units of this CTA: blockIdx.x + n*G

  if (warp < 4) {
    reg_dec<64>();
    if (warp == 0) {
      // =============================== loader ===============================
      if (lane == 0) {
        uint32_t k_item[NSLOT] = {0, 0, 0};
        for (int n = 0; n < n_units; ++n) {
          const int unit = blockIdx.x + n * G;
          const uint8_t* src = img + static_cast<size_t>(unit) * ATT3_UNIT_BYTES;
          const int b = n & 1;
          if (n >= 2) mbar_wait(&bars.kvfree[b], ((n >> 1) - 1) & 1);
          mbar_expect_tx(&bars.kv[b], 2 * ATT3_KV_BYTES);
          bulk_load_1d(smem + S3_KV + b * 2 * ATT3_KV_BYTES, src + ATT_IMG_BYTES, 2 * ATT3_KV_BYTES, &bars.kv[b]);
#pragma unroll
          for (int s = 0; s < NSLOT; ++s) {
            const int items = (n % NSLOT == s) ? 2 : 1;
            for (int it = 0; it < items; ++it) {
              const uint32_t k = k_item[s]++;
              const int qb = k & 1;
              if (k >= 2) mbar_wait(&bars.qfree[s][qb], ((k >> 1) - 1) & 1);
              uint8_t* dst = smem + S3_Q + (s * 2 + qb) * 8192;
              if (it == 0) {
                mbar_expect_tx(&bars.q[s][qb], 8192);
                bulk_load_1d(dst, src + s * 8192, 8192, &bars.q[s][qb]);
              } else {   // the 8 tail rows: one 512 B row group; the MMA descriptor replicates it with SBO = 0
                mbar_expect_tx(&bars.q[s][qb], 512);
                bulk_load_1d(dst, src + 48 * 512, 512, &bars.q[s][qb]);
              }
            }
          }
        }
      }
    } else {
      // =============================== MMA issuer of slot `warp - 1` ===============================
      // Warp-uniform control flow and operands; one elected lane issues (see the header comment).  The per-chunk
      // path is kept to a barrier wait plus the MMAs: everything item-dependent (unit decode, descriptor bases) is
      // computed once per item, the chunk loop is unrolled so descriptor offsets are immediates.
      const int s = warp - 1;
      const uint32_t tB = s * T3_SLOT, tO = s * T3_SLOT + T3_O;
      const uint32_t sbase = smem_u32(smem);
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 64, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, 32, 0, 1);
      auto wait_all = [&](uint64_t* bar, uint32_t parity) {   // one lane polls, the warp follows
        if (lane == 0) mbar_wait(bar, parity);
        __syncwarp();
      };
      // descriptors of one item: advancing the start-address field by (bytes >> 4) walks chunks / K steps
      struct ItemDesc {
        uint64_t dq, dk, dqa, dka, dv;
        bool masked;
        int qb, nb;
      };
      uint32_t n_qa = 0;
      // waits for the operands of the item at cursor x and builds its descriptors
      auto open_item = [&](const Cursor& x) {
        const UnitInfo u = unit_info(p, blockIdx.x + x.n * G);
        ItemDesc d;
        d.masked = u.md || u.mh || u.mw;
        const bool tail = x.it == 1;
        d.qb = x.k & 1;
        d.nb = x.n & 1;
        if (x.it == 0) wait_all(&bars.kv[d.nb], (x.n >> 1) & 1);
        wait_all(&bars.q[s][d.qb], (x.k >> 1) & 1);
        if (d.masked) {
          wait_all(&bars.qa[s], n_qa & 1);
          ++n_qa;
        }
        tc_fence_after();
        const uint32_t aK = sbase + S3_KV + d.nb * 2 * ATT3_KV_BYTES;
        d.dq = umma_smem_desc(sbase + S3_Q + (s * 2 + d.qb) * 8192, 128, tail ? 0u : 512u, UMMA_SW_NONE);
        d.dk = umma_smem_desc(aK, 128, 512, UMMA_SW_NONE);
        d.dqa = umma_smem_desc(sbase + S3_QAUG + s * 2048, 0, tail ? 0u : 128u, UMMA_SW_NONE);
        d.dka = umma_smem_desc(sbase + S3_KAUG, 0, 128, UMMA_SW_NONE);
        d.dv = umma_smem_desc(aK + ATT3_KV_BYTES, 512, 128, UMMA_SW_NONE);
        return d;
      };
      // S_c of item d into S/P buffer `buf` (called by the elected lane only)
      auto issue_s = [&](const ItemDesc& d, int c, uint32_t buf) {
        const uint32_t tS = tB + buf * 64;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          umma_f16_ss(tS, d.dq + ((ks * 256) >> 4), d.dk + ((c * (56 * 64) + ks * 256) >> 4), idesc_s, ks);
        if (d.masked) umma_f16_ss(tS, d.dqa, d.dka + ((c * (56 * 16)) >> 4), idesc_s, 1u);
        umma_commit(&bars.s[s][buf]);
        if (c == NCHUNK3 - 1) umma_commit(&bars.qfree[s][d.qb]);   // every S MMA of this item has read the Q tile
      };
      Cursor cur = {0, 0, 0, 0};
      uint32_t j = 0;
      ItemDesc d;
      if (cur.n < n_units) {
        d = open_item(cur);
        if (elect_one()) {
          issue_s(d, 0, 0);
          issue_s(d, 1, 1);
        }
        __syncwarp();
      }
#pragma unroll 1
      while (cur.n < n_units) {
        const bool last_item = cur.it == ((cur.n % NSLOT == s) ? 1 : 0);
#pragma unroll
        for (int c = 0; c < NCHUNK3; ++c, ++j) {
          const uint32_t buf = j & 1;
          // P_c written over S_c.  One barrier per buffer: a warp can run at most two chunks ahead of its slowest
          // sibling (S_{j+2} is only issued once phase j completed), so its arrivals for chunks j and j+1 never mix
          wait_all(&bars.p[s][buf], (j >> 1) & 1);
          if (c == 0 && cur.k > 0) wait_all(&bars.of[s], (cur.k - 1) & 1);   // previous item's O was read
          tc_fence_after();
          if (elect_one()) {
            const uint32_t tP = tB + buf * 64;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_ts(tO, tP + 8 * ks, d.dv + ((c * (56 * 64) + ks * 1024) >> 4), idesc_pv, (c > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&bars.pv[s][buf]);
            if (c == NCHUNK3 - 1 && last_item) umma_commit(&bars.kvfree[d.nb]);   // this slot is done with the unit
            // the buffer P_c lives in is free once PV_c has run: the tensor pipe executes in issue order
            if (c + 2 < NCHUNK3) issue_s(d, c + 2, buf);
          }
          __syncwarp();
        }
        for (int c = 0; c < NCHUNK3; ++c) cur.advance(s);
        if (cur.n < n_units) {
          // the next item's first two chunks (its Q tile / K | V image / Qaug rows were produced long ago; S_0 runs
          // under the softmax warps' epilogue of the item that just ended)
          d = open_item(cur);
          if (elect_one()) {
            issue_s(d, 0, j & 1);
            issue_s(d, 1, (j + 1) & 1);
          }
          __syncwarp();
        }
      }
    }
  } else {
    reg_inc<144>();
    // =============================== softmax warps: one thread per query row ===============================
    const int s = (warp - 4) >> 2, q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tB = lane_off + s * T3_SLOT, tO = lane_off + s * T3_SLOT + T3_O;
    float* scr = reinterpret_cast<float*>(smem + S3_SCR + s * SCR_BYTES);
    uint8_t* qaug = smem + S3_QAUG + s * 2048;
    const uint32_t stab = smem_u32(smem + S3_TAB);
    const float invH = 7.0f / static_cast<float>(g.Hp), invW = 7.0f / static_cast<float>(g.Wp);

    // per-item thread state: which query row this thread owns and everything derived from the window position
    struct Item {
      int win_g, ri;
      bool tail, md, mh, mw;
      float fval;          // this LANE's fragment coordinate: lanes 0..6 rows h = lane, lanes 8..14 cols w = lane - 8
      float fh_i, fw_i;    // the thread's own row / column fragment coordinates
      float Aw[7];         // |fw_i - fw(w)|: the w part of the GRPB gate
      uint32_t trow0;      // table address of (d_i, h_i, w_i) against key (0, 0, 0)
    };
    const int win_g0 = static_cast<int>(blockIdx.x) / p.heads, win_step = G / p.heads;   // G is a multiple of heads
    const float rcp_nW = 1.0f / static_cast<float>(g.nW), rcp_hw = 1.0f / static_cast<float>(g.nwh * g.nww),
                rcp_w = 1.0f / static_cast<float>(g.nww);
    auto fdiv = [](int x, float rcp) { return __float2int_rz((static_cast<float>(x) + 0.5f) * rcp); };   // exact for x < 2^20
    auto open_item = [&](const Cursor& x) {
      Item it;
      it.win_g = win_g0 + x.n * win_step;
      const int win = it.win_g - fdiv(it.win_g, rcp_nW) * g.nW;
      const int wdi = fdiv(win, rcp_hw), rem = win - wdi * (g.nwh * g.nww);
      const int whi = fdiv(rem, rcp_w), wwi = rem - whi * g.nww;
      it.md = g.sd != 0 && wdi == g.nwd - 1;
      it.mh = g.sh != 0 && whi == g.nwh - 1;
      it.mw = g.sw != 0 && wwi == g.nww - 1;
      it.tail = x.it == 1;
      it.ri = it.tail ? 384 + (lane & 7) : s * 128 + q * 32 + lane;
      // query rows are (h,w,d) with d fastest, like the key slots: the tail tile is the 8 frames of the last cell
      const int d_i = it.ri & 7, hw_i = it.ri >> 3, h_i = hw_i / 7, w_i = hw_i - h_i * 7;
      it.fval = (lane & 8) ? frag_coord(wwi * 7, lane & 7, g.sw, g.Wp, invW) : frag_coord(whi * 7, lane & 7, g.sh, g.Hp, invH);
      it.fh_i = __shfl_sync(0xffffffffu, it.fval, h_i);
      it.fw_i = __shfl_sync(0xffffffffu, it.fval, 8 + w_i);
#pragma unroll
      for (int i = 0; i < 7; ++i) it.Aw[i] = fabsf(it.fw_i - __shfl_sync(0xffffffffu, it.fval, 8 + i));
      it.trow0 = stab + 16u * static_cast<uint32_t>((d_i + 6) * ATT3_SD + (h_i + 6) * ATT3_SH + (w_i + 6));
      // query side of the folded mask: [a_d 1 a_h 1 a_w 1 0 0] per row, a pair zeroed when its dim is not masked in
      // this window.  Written while no S MMA of the previous item can still read the rows (see the call sites).
      if (it.md || it.mh || it.mw) {
        uint4 w;
        w.x = it.md ? pack_half2(d_i >= 4 ? 1.f : 0.f, 1.f) : 0u;
        w.y = it.mh ? pack_half2(h_i >= 4 ? 1.f : 0.f, 1.f) : 0u;
        w.z = it.mw ? pack_half2(w_i >= 4 ? 1.f : 0.f, 1.f) : 0u;
        w.w = 0u;
        const int row = q * 32 + lane;
        if (!it.tail || row < 8) *reinterpret_cast<uint4*>(qaug + row * 16) = w;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.qa[s]);
      }
      return it;
    };

    Cursor cur = {0, 0, 0, 0};
    uint32_t j = 0;                            // chunk ordinal of this slot
#ifdef A3_TIMING   // debug build: per-warp cycle accounting, printed by one CTA (tools/attn_profile.py)
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
    long long t_mark = t_begin;
#define A3_LAP(k) do { const long long t_now = clock64(); tacc[k] += t_now - t_mark; t_mark = t_now; } while (0)
#else
#define A3_LAP(k)
#endif
    mbar_wait(&bars.tab, 0);
    Item item;
    if (cur.n < n_units) item = open_item(cur);
#pragma unroll 1
    while (cur.n < n_units) {
      // ---------------- one item: a 128-row tile or the replicated tail rows ----------------
      const bool tail = item.tail;
      Cursor next = cur;
      next.c = NCHUNK3 - 1;
      next.advance(s);                         // first chunk of the next item of this slot
      Item nitem;

      float m_ref = -INFINITY, l_run = 0.f;
      bool fresh = true;                       // no chunk of this row processed yet (warp-uniform)
#pragma unroll 1
      for (int c = 0; c < NCHUNK3; ++c, ++j) {
        const uint32_t tS = tB + (j & 1) * 64;
        A3_LAP(7);                             // loop top
        mbar_wait(&bars.s[s][j & 1], (j >> 1) & 1);
        A3_LAP(tail ? 1 : 0);                  // wait for S
        __syncwarp();
        tc_fence_after();
        if (!tail) {
          uint32_t r[56], h[32];
          const uint32_t tb = item.trow0 - 16u * static_cast<uint32_t>(c * ATT3_SH);
          const float Ah = fabsf(item.fh_i - __shfl_sync(0xffffffffu, item.fval, c));
          float4 e[2][4];
          // ---- fused pass against the standing m_ref: bias, exp2, pack, row sum, chunk max ----
          tmem_ld_x32(tS, r);
          tmem_ld_x16(tS + 32, r + 32);
          tmem_ld_x8(tS + 48, r + 48);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) e[0][kk] = lds_f4(tb - 16u * (2 * kk * ATT3_SD));
          tmem_wait_ld();
          if (fresh) {
            // the row's first chunk: take the max over its first key position (8 logits) as m_ref.  Any later logit
            // more than 15 above it triggers the two-pass redo below, so the estimate only has to be in the right range
            const float2 fg2 = splat2(Ah + item.Aw[0]);
            float mx = -INFINITY;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const float4 t = e[0][kk];
              const float2 v = fadd2(ffma2(fg2, make_float2(t.z, t.w),
                                           make_float2(__uint_as_float(r[2 * kk]), __uint_as_float(r[2 * kk + 1]))),
                                     make_float2(t.x, t.y));
              mx = fmax3(mx, v.x, v.y);
            }
            m_ref = mx;
          }
          float gmax = -INFINITY;
          {
            const float2 negm2 = splat2(-m_ref);
            float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int wj = 0; wj < 7; ++wj) {
              if (wj < 6) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) e[(wj + 1) & 1][kk] = lds_f4(tb - 16u * (2 * kk * ATT3_SD + wj + 1));
              }
              bias_pos(r, wj, e[wj & 1], splat2(Ah + item.Aw[wj]), gmax);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const int i = wj * 8 + 2 * kk;
                const float2 x = fadd2(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), negm2);
                const float2 pe = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                sum2 = fadd2(sum2, pe);
                h[i >> 1] = pack_half2(pe.x, pe.y);
              }
            }
            const bool redo = __any_sync(0xffffffffu, gmax > m_ref + TAU);
            if (!redo) {
              l_run += sum2.x + sum2.y;
            } else {
              // ---- rare: a logit left the fp16-safe range of m_ref.  Two passes: bias + exact chunk max, raise m_ref
              // (rescaling O and l), then exp2 ----
              tmem_ld_x32(tS, r);
              tmem_ld_x16(tS + 32, r + 32);
              tmem_ld_x8(tS + 48, r + 48);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) e[0][kk] = lds_f4(tb - 16u * (2 * kk * ATT3_SD));
              tmem_wait_ld();
              gmax = -INFINITY;
#pragma unroll
              for (int wj = 0; wj < 7; ++wj) {
                if (wj < 6) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk) e[(wj + 1) & 1][kk] = lds_f4(tb - 16u * (2 * kk * ATT3_SD + wj + 1));
                }
                bias_pos(r, wj, e[wj & 1], splat2(Ah + item.Aw[wj]), gmax);
              }
              const float m_new = fmaxf(m_ref, gmax);
              const float alpha = fast_exp2(m_ref - m_new);
              if (!fresh) {
                // O holds this item's partial sums against the old m_ref; PV of the previous chunk must have landed
                // (one barrier per buffer parity: a single one could still be a phase behind and alias the parity test)
                mbar_wait(&bars.pv[s][(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                uint32_t o[32];
                tmem_ld_x32(tO, o);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_x32(tO, o);
                l_run *= alpha;
              }
              m_ref = m_new;
              const float2 negm2b = splat2(-m_ref);
              sum2 = make_float2(0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 28; ++i) {
                const float2 x = fadd2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), negm2b);
                const float2 pe = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                sum2 = fadd2(sum2, pe);
                h[i] = pack_half2(pe.x, pe.y);
              }
              l_run += sum2.x + sum2.y;
            }
          }
          fresh = false;
          h[28] = h[29] = h[30] = h[31] = 0u;  // 8 zero pad slots complete the fourth K step
          tmem_st_x32(tS, h);                  // P_c in place over S_c
        } else {
          // ---- tail tile: the 8 rows are replicated in every lane group; warp q takes the temporal pair d = 2q, 2q+1 of
          // every key position (14 logits per chunk, all four warps busy on every chunk) and writes zero P elsewhere.
          // Few enough values to keep in registers: always the exact two-pass form. ----
          const uint32_t tb = item.trow0 - 16u * static_cast<uint32_t>(c * ATT3_SH + 2 * q * ATT3_SD);
          const float Ah = fabsf(item.fh_i - __shfl_sync(0xffffffffu, item.fval, c));
          uint32_t r2[7][2];
          float4 e[7];
#pragma unroll
          for (int wj = 0; wj < 7; ++wj) {
            tmem_ld_x2(tS + 8 * wj + 2 * q, r2[wj]);
            e[wj] = lds_f4(tb - 16u * wj);
          }
          tmem_wait_ld();
          float2 v[7];
          float gmax = -INFINITY;
#pragma unroll
          for (int wj = 0; wj < 7; ++wj) {
            v[wj] = fadd2(ffma2(splat2(Ah + item.Aw[wj]), make_float2(e[wj].z, e[wj].w),
                                make_float2(__uint_as_float(r2[wj][0]), __uint_as_float(r2[wj][1]))),
                          make_float2(e[wj].x, e[wj].y));
            gmax = fmax3(gmax, v[wj].x, v[wj].y);
          }
          if (__any_sync(0xffffffffu, fresh || gmax > m_ref + TAU)) {
            const float m_new = fmaxf(m_ref, gmax);
            if (!fresh) {
              const float alpha = fast_exp2(m_ref - m_new);
              mbar_wait(&bars.pv[s][(j - 1) & 1], ((j - 1) >> 1) & 1);
              tc_fence_after();
              uint32_t o[32];
              tmem_ld_x32(tO, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x32(tO, o);
              l_run *= alpha;
            }
            m_ref = m_new;
          }
          fresh = false;
          const float2 negm2 = splat2(-m_ref);
          float2 sum2 = make_float2(0.f, 0.f);
          uint32_t h[32];
#pragma unroll
          for (int wj = 0; wj < 7; ++wj) {
            const float2 x = fadd2(v[wj], negm2);
            const float2 pe = make_float2(fast_exp2(x.x), fast_exp2(x.y));
            sum2 = fadd2(sum2, pe);
            const uint32_t ph = pack_half2(pe.x, pe.y);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) h[4 * wj + kk] = (kk == q) ? ph : 0u;
          }
          l_run += sum2.x + sum2.y;
          h[28] = h[29] = h[30] = h[31] = 0u;
          tmem_st_x32(tS, h);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p[s][j & 1]);
        A3_LAP(tail ? 4 : 3);                  // chunk work
      }
      // every S MMA of this item has completed (its last chunk was visible): the next item's Qaug rows may go in, and
      // the rest of its set-up runs while the MMA issuer finishes PV of the last chunk
      if (next.n < n_units) nitem = open_item(next);
      A3_LAP(5);                               // next item's set-up

      // ---- epilogue: O / l -> global ----
      mbar_wait(&bars.pv[s][(j - 1) & 1], ((j - 1) >> 1) & 1);
      A3_LAP(2);                               // wait for the last PV
      __syncwarp();
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_x32(tO, o);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.of[s]);
      if (!tail) {
        const float inv = 1.0f / l_run;
        const int orow = p.rows_dfast ? item.ri : (item.ri & 7) * 49 + (item.ri >> 3);   // natural rows for the stand-alone op
        __half* dst = p.out + (static_cast<size_t>(item.win_g) * 392 + orow) * p.C + head * ATT_HD;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_half2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          v.y = pack_half2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          v.z = pack_half2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          v.w = pack_half2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + i) = v;
        }
      } else {
        // four partial softmaxes per row (one per warp), each against its own m_ref: merge like split-K
        if (q > 0 && lane < 8) {
          float* d = scr + ((q - 1) * 8 + lane) * 34;
#pragma unroll
          for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(o[i]);
          d[32] = l_run;
          d[33] = m_ref;
        }
        named_bar_sync(1 + s, 128);
        if (q == 0 && lane < 8) {
          float mq[4], wq[4];
          mq[0] = m_ref;
#pragma unroll
          for (int i = 1; i < 4; ++i) mq[i] = scr[((i - 1) * 8 + lane) * 34 + 33];
          const float M = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
#pragma unroll
          for (int i = 0; i < 4; ++i) wq[i] = fast_exp2(mq[i] - M);
          float L = wq[0] * l_run;
#pragma unroll
          for (int i = 1; i < 4; ++i) L = fmaf(wq[i], scr[((i - 1) * 8 + lane) * 34 + 32], L);
          const float inv = 1.0f / L;
          const int orow = p.rows_dfast ? 384 + lane : lane * 49 + 48;
          __half* dst = p.out + (static_cast<size_t>(item.win_g) * 392 + orow) * p.C + head * ATT_HD;
#pragma unroll
          for (int i0 = 0; i0 < 32; i0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float a = wq[0] * __uint_as_float(o[i0 + i]);
#pragma unroll
              for (int w = 1; w < 4; ++w) a = fmaf(wq[w], scr[((w - 1) * 8 + lane) * 34 + i0 + i], a);
              v[i] = a * inv;
            }
            uint4 pk;
            pk.x = pack_half2(v[0], v[1]);
            pk.y = pack_half2(v[2], v[3]);
            pk.z = pack_half2(v[4], v[5]);
            pk.w = pack_half2(v[6], v[7]);
            *reinterpret_cast<uint4*>(dst + i0) = pk;
          }
        }
        named_bar_sync(1 + s, 128);   // the scratch is rewritten by this slot's next tail item
      }
      cur = next;
      item = nitem;
      A3_LAP(6);                               // epilogue
    }
#ifdef A3_TIMING
    if (blockIdx.x == 5 && lane == 0)
      printf("A3T warp %2d slot %d q %d: total %lld | S wait %lld tail %lld | pv wait %lld | chunk full %lld tail %lld | "
             "set-up %lld | epilogue %lld | other %lld\n", warp, s, q, clock64() - t_begin, tacc[0], tacc[1], tacc[2], tacc[3],
             tacc[4], tacc[5], tacc[6], tacc[7]);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// [L, heads] fp32 tables -> paired layout [heads][ATT3_PAIR_LEN] float4 {t0[e], t0[e-1], t1[e], t1[e-1]} (log2 domain),
// index (e + 6) * SD + (dh + 6) * SH + (dw + 6), e in [-6, 7], bias = t0 + fg * t1
__global__ void pack_bias_pair_kernel(const float* __restrict__ rel, const float* __restrict__ frag,
                                      float4* __restrict__ out, int heads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= heads * ATT3_PAIR_LEN) return;
  const int h = i / ATT3_PAIR_LEN, f = i - h * ATT3_PAIR_LEN;
  const int ee = f / ATT3_SD, rem = f - ee * ATT3_SD, dh = rem / ATT3_SH, dw = rem - dh * ATT3_SH;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ee < 14 && dh < 13 && dw < 13) {
    // relative_position_index (:214-231): (dd + 7) * 169 + (dh + 6) * 13 + (dw + 6) with d = query - key
    const int ia = (ee + 1) * 169 + dh * 13 + dw;    // dd = e      = ee - 6
    const int ib = ee * 169 + dh * 13 + dw;          // dd = e - 1
    const float ra = rel[ia * heads + h], rb = rel[ib * heads + h];
    if (frag != nullptr) {
      const float fa = frag[ia * heads + h], fb = frag[ib * heads + h];
      v = make_float4(fa * LOG2E, fb * LOG2E, (ra - fa) * LOG2E, (rb - fb) * LOG2E);
    } else {
      v = make_float4(ra * LOG2E, rb * LOG2E, 0.f, 0.f);
    }
  }
  out[i] = v;
}

}  // namespace

int launch_pack_bias_pair(const float* rel, const float* frag, float* out, int heads, cudaStream_t stream) {
  const int n = heads * ATT3_PAIR_LEN;
  pack_bias_pair_kernel<<<(n + 255) / 256, 256, 0, stream>>>(rel, frag, reinterpret_cast<float4*>(out), heads);
  count_launch();
  return check_cuda(cudaGetLastError(), "pack_bias_pair_kernel launch");
}

int launch_window_attn3(const AttnParams& p, cudaStream_t stream) {
  const WinGeom& g = p.geom;
  KVQ_REQUIRE(g.wd == 8 && g.wh == 7 && g.ww == 7 && p.base_wd == 8 && p.base_wh == 7 && p.base_ww == 7,
              KVQ_ERR_BAD_SHAPE, "attn3: built for the (8,7,7) window");
  KVQ_REQUIRE((g.sd == 0 || g.sd == 4) && (g.sh == 0 || g.sh == 3) && (g.sw == 0 || g.sw == 3), KVQ_ERR_BAD_SHAPE,
              "attn3: shift must be (4,3,3) or clamped to 0 per dim");
  KVQ_REQUIRE(p.C == p.heads * ATT_HD, KVQ_ERR_BAD_SHAPE, "attn3: C=%d must be heads(%d) x 32", p.C, p.heads);
  const long long units = static_cast<long long>(p.B) * g.nW * p.heads;
  KVQ_REQUIRE(units > 0 && units < (1ll << 31), KVQ_ERR_BAD_SHAPE, "attn3: %lld units", units);
  KVQ_REQUIRE(p.heads <= num_sms(), KVQ_ERR_BAD_SHAPE, "attn3: %d heads exceed the SM count", p.heads);
  int dev = 0;
  KVQ_CUDA(cudaGetDevice(&dev));
  static bool attr[64] = {};
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    KVQ_CUDA(cudaFuncSetAttribute(window_attn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S3_SMEM));
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  int grid = num_sms() / p.heads * p.heads;   // multiple of heads so a CTA keeps one head's table
  if (grid > units) grid = static_cast<int>(units);
  count_launch();
  const float* tab3 = p.packed_tab + static_cast<size_t>(p.heads) * (2536 + 4336) * 2;   // after compact + fast sections
  return launch_pdl(window_attn3_kernel, dim3(grid), dim3(A3_THREADS), S3_SMEM, stream, p,
                    reinterpret_cast<const float4*>(tab3), static_cast<int>(units));
}

}  // namespace kvq
