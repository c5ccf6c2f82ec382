// Fused 3-D (shifted-)window attention for one (window, head) unit per CTA  (WindowAttention3D.forward,
// swin_backbone.py:245-326, with the index tensors of compute_mask :560-586 and global_position_index :22-50
// evaluated from token coordinates instead of being materialised):
//
//   S = Q K^T                         tcgen05.mma  M=128 x N=400 (208 + 192) x K=32, accumulators in TMEM
//   S += rel[rpi]*fg + frag[rpi]*(1-fg)  (GRPB gate fg = L1 fragment distance, :291-309)  [+ -100 region mask]
//   P = softmax(S)                    fp32, exact row max, exp2 on pre-scaled logits; P written fp16 to smem
//   O = P V                           tcgen05.mma  M=128 x N=32 x K=400, V read MN-major straight from its image
//   out[row, head*32:+32] = O / rowsum
//
// Operands arrive as one contiguous 76 800 B image (Q|K|V, UMMA no-swizzle core-matrix order, written by the
// QKV GEMM epilogue) fetched with a single cp.async.bulk.  The logits tile never leaves the SM.
// Thread mapping: 8 warps; warp w owns TMEM lanes 32*(w&3).. (its query rows) and key columns
// [200*(w>>2), +200) -- four 50-column temporal slabs.
#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int ATT_THREADS = 256;
constexpr int P_SBO = (ATT_ROWS / 8) * 128;  // 6400: byte stride between 8-row groups of the P image
constexpr int P_BYTES = 128 * ATT_ROWS * 2;  // 102 400
constexpr int TMEM_O_COL = 400;
constexpr float LOG2E = 1.4426950408889634f;

struct ColMeta {  // 8 bytes per key slot
  int base;       // rd*(2bh-1)(2bw-1) + rh*(2bw-1) + rw   (token enumerated in the BASE window, :264)
  uint32_t pk;    // fh | fw << 8 | region << 16 | valid << 24
};

__global__ void __launch_bounds__(ATT_THREADS, 1) window_attn_kernel(const AttnParams p, const float2* __restrict__ tabs,
                                                                      int tab_len, int rpi_offset, int variant) {
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
  // per-head bias table {t0, t1}:  bias = t0 + fg * t1
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
        int oh = ph + g.sh; if (oh >= g.Hp) oh -= g.Hp;
        int ow = pw + g.sw; if (ow >= g.Wp) ow -= g.Wp;
        const int rd = g.sd == 0 ? 0 : (pd >= g.Dp - g.wd) + (pd >= g.Dp - g.sd);
        const int rh = g.sh == 0 ? 0 : (ph >= g.Hp - g.wh) + (ph >= g.Hp - g.sh);
        const int rw = g.sw == 0 ? 0 : (pw >= g.Wp - g.ww) + (pw >= g.Wp - g.sw);
        int fh = static_cast<int>(floorf(static_cast<float>(oh) * fscale_h)); if (fh > g.wh - 1) fh = g.wh - 1;
        int fw = static_cast<int>(floorf(static_cast<float>(ow) * fscale_w)); if (fw > g.ww - 1) fw = g.ww - 1;
        const int brd = i / bhw, brh = (i / p.base_ww) % p.base_wh, brw = i % p.base_ww;
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
  // padded key slots: V rows must be finite (P is exactly 0 there); K rows are masked by select below
  for (int c = tid; c < ATT_ROWS; c += ATT_THREADS) {
    if ((cmeta[c].pk >> 24) == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(sV + att_img_offset(c, k)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();

  const int ntiles = (g.N + 127) / 128;
  const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const int col0 = hs * 200;
  const int shifted = p.shifted;

  for (int t = 0; t < ntiles; ++t) {
    const uint32_t ph = static_cast<uint32_t>(t & 1);
    // ---------------- S = Q_t K^T ----------------
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aQ = smem_u32(sQ) + static_cast<uint32_t>(t) * 8192u;
      const uint32_t aK = smem_u32(sK);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t dq = umma_smem_desc(aQ + ks * 256, 128, 512, UMMA_SW_NONE);
        const uint64_t dk0 = umma_smem_desc(aK + ks * 256, 128, 512, UMMA_SW_NONE);
        const uint64_t dk1 = umma_smem_desc(aK + 26 * 512 + ks * 256, 128, 512, UMMA_SW_NONE);
        umma_f16_ss(tmem_base, dq, dk0, umma_idesc_f16(128, 208, 0, 0), ks);
        umma_f16_ss(tmem_base + 208, dq, dk1, umma_idesc_f16(128, 192, 0, 0), ks);
      }
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    __syncwarp();
    tc_fence_after();

    // this thread's query row
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

    // ---------------- pass 1: logits += bias (+ mask); row max; write back ----------------
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
          if (shifted && reg_i != static_cast<int>((m.pk >> 16) & 255)) v += -100.0f;
          v *= LOG2E;
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

    // ---------------- pass 2: P = exp2(t - max) -> fp16 smem image; row sum ----------------
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

    // ---------------- O = P V ----------------
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
      const uint32_t v_lbo = variant == 1 ? 128u : 512u, v_sbo = variant == 1 ? 512u : 128u;
#pragma unroll 1
      for (int ks = 0; ks < ATT_ROWS / 16; ++ks) {
        const uint64_t dp = umma_smem_desc(aP + ks * 256, 128, P_SBO, UMMA_SW_NONE);
        const uint64_t dv = umma_smem_desc(aV + ks * 1024, v_lbo, v_sbo, UMMA_SW_NONE);
        umma_f16_ss(tmem_base + TMEM_O_COL, dp, dv, umma_idesc_f16(128, 32, 0, 1), ks);
      }
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
        __half* dst = p.out + (static_cast<size_t>(win_g) * g.N + ri) * p.C + head * ATT_HD;
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
    }
    tc_fence_before();
    __syncthreads();
  }

  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// [L, heads] fp32 tables -> [heads][Lp] float2 {t0, t1} with bias = t0 + fg*t1
__global__ void pack_bias_kernel(const float* __restrict__ rel, const float* __restrict__ frag, float2* __restrict__ out,
                                 int L, int Lp, int heads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= heads * Lp) return;
  const int h = i / Lp, e = i - h * Lp;
  float2 v = make_float2(0.f, 0.f);
  if (e < L) {
    const float r = rel[e * heads + h];
    if (frag != nullptr) {
      const float f = frag[e * heads + h];
      v = make_float2(f, r - f);   // rel*fg + frag*(1-fg) = frag + fg*(rel-frag)
    } else {
      v = make_float2(r, 0.f);
    }
  }
  out[i] = v;
}

}  // namespace

int attn_table_len(int bd, int bh, int bw) {
  const int L = (2 * bd - 1) * (2 * bh - 1) * (2 * bw - 1);
  return (L + 1) & ~1;
}

size_t attn_smem_bytes(int tab_len) {
  return 128 + ATT_UNIT_BYTES + P_BYTES + ATT_ROWS * 8 + 2048 + 64 + static_cast<size_t>(tab_len) * 8;
}

int launch_pack_bias(const float* rel, const float* frag, float* out, int bd, int bh, int bw, int heads,
                     cudaStream_t stream) {
  const int L = (2 * bd - 1) * (2 * bh - 1) * (2 * bw - 1);
  const int Lp = attn_table_len(bd, bh, bw);
  const int n = heads * Lp;
  pack_bias_kernel<<<(n + 255) / 256, 256, 0, stream>>>(rel, frag, reinterpret_cast<float2*>(out), L, Lp, heads);
  count_launch();
  return check_cuda(cudaGetLastError(), "pack_bias_kernel launch");
}

int launch_window_attn(const AttnParams& p, cudaStream_t stream) {
  const WinGeom& g = p.geom;
  KVQ_REQUIRE(p.C == p.heads * ATT_HD, KVQ_ERR_BAD_SHAPE, "attn: C=%d must be heads(%d) x 32", p.C, p.heads);
  KVQ_REQUIRE(g.wd * ATT_SLAB <= ATT_ROWS && g.SL <= ATT_SLAB - 1 && g.wd <= p.base_wd && g.wh <= p.base_wh &&
                  g.ww <= p.base_ww,
              KVQ_ERR_BAD_SHAPE, "attn: window (%d,%d,%d) exceeds the 8 x 49 key layout", g.wd, g.wh, g.ww);
  KVQ_REQUIRE(g.Hp < 256 * 1 && g.wh < 256, KVQ_ERR_BAD_SHAPE, "attn: geometry out of range");
  const int tab_len = attn_table_len(p.base_wd, p.base_wh, p.base_ww);
  const size_t smem = attn_smem_bytes(tab_len);
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
  const long long units = static_cast<long long>(p.B) * g.nW * p.heads;
  KVQ_REQUIRE(units > 0 && units < (1ll << 31), KVQ_ERR_BAD_SHAPE, "attn: %lld units", units);
  window_attn_kernel<<<static_cast<unsigned>(units), ATT_THREADS, smem, stream>>>(
      p, reinterpret_cast<const float2*>(p.packed_tab), tab_len, rpi_offset, p.variant);
  count_launch();
  return check_cuda(cudaGetLastError(), "window_attn_kernel launch");
}

}  // namespace kvq
