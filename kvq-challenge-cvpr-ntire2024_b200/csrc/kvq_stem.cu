// Implicit-GEMM stem convolution for 3-channel fp32 network inputs: Conv3d(3 -> cout, (KT,7,7), stride (1,2,2),
// padding (KT/2,3,3)) + folded BatchNorm + ReLU, fp16 channels-last output.  Serves
//   SlowFast slow stem  (KT = 1, cout 64), fast stem (KT = 5, cout 8)      -- oracle/slowfast.py:_stem
//   SimpleVQA ResNet-50 conv1 + bn1 + relu (KT = 1, cout 64; frames = the T axis) -- simpleVQA_model.py:235-237
//
// No patch matrix ever reaches HBM (the explicit im2col of the fast stem is 772 MB per 32x256x256 clip).  A CTA owns an
// 8 x 16 tile of output pixels and walks the output frames of its work item:
//   * the input halo (22 x 38 pixels x 3 channels) of each needed frame is converted to fp16 into a KT-slot ring in
//     shared memory (one new frame per output frame),
//   * per (frame tap dt, channel c) the 128 builder threads assemble the [128 pixels x 64] K-block (k = dy*8 + dx; the
//     dy = 7 / dx = 7 slots carry zero weights) in registers -- 4 aligned 32-bit shared loads per kernel row -- and
//     write it with one tcgen05.st into a 3-stage ring of A tiles IN TENSOR MEMORY (thread = pixel = TMEM lane): the
//     A operand never goes back through shared memory (the shared-memory store path was the measured limiter),
//   * one thread issues 4 tcgen05.mma (A from TMEM, B from smem; M128 x N{16,64} x K16) per K-block into one of two
//     TMEM accumulators,
//   * epilogue: tcgen05.ld (thread = pixel) -> + shift, ReLU -> fp16 -> 16-byte stores.
// Two CTAs per SM overlap each other's load / build / epilogue phases.
#include <algorithm>

#include "kvq_common.cuh"
#include "kvq_kernels.cuh"

namespace kvq {

namespace {

constexpr int ST_BUILDERS = 128;              // 4 builder / epilogue warps: thread = output pixel of the tile
constexpr int ST_THREADS = 160;               // + warp 4: single-thread tcgen05.mma issuer
constexpr int ST_STAGES = 3;                  // A-tile ring
constexpr int ST_TY = 8, ST_TX = 16;          // output pixels per tile
constexpr int ST_PR = 22, ST_PC = 48;         // halo rows / row pitch (halfs); pitch % 64 == 48 keeps the loads conflict-free
constexpr int ST_SLOT = 3 * ST_PR * ST_PC;    // halfs per frame slot
constexpr int ST_A_COLS = 32;                 // one K-block of A in TMEM: 64 halfs per row = 32 columns

struct StemParams {
  const float* x;        // [N,3,T,H,W]
  const __half* w;       // [KT*3][NP][64]
  const float* shift;    // [NP]
  __half* out;           // [N*T*Hs*Ws, COUT]
  int N, T, H, W, Hs, Ws;
  int tseg, nseg, ty, tx;
  int items;
};

template <int KT, int NP>
struct StemCfg {
  static constexpr int KB = KT * 3;
  static constexpr int B_BYTES = KB * NP * 128;
  static constexpr int PATCH_BYTES = KT * ST_SLOT * 2;
  static constexpr int SMEM = B_BYTES + PATCH_BYTES + 128 + 1024;
  static constexpr int T_A = 2 * NP;                                               // A ring after the two accumulators
  static constexpr int TMEM_COLS = T_A + ST_STAGES * ST_A_COLS <= 128 ? 128 : 256;
  static constexpr int CTAS = TMEM_COLS <= 128 ? 3 : 2;
};

// Roles: the 128 builder threads own the halo ring and the A-tile ring and run ahead of the tensor pipe (they block
// only on the `empty` barrier of a stage used three K-blocks earlier); one thread of warp 4 issues the MMAs and commits
// them to `empty` / `accfull`.  The epilogue of frame-tile n runs on the builder warps after they have built the K-blocks
// of frame-tile n+1 (two TMEM accumulators), so nobody waits for the tensor pipe in steady state.
template <int KT, int NP, int COUT>
__global__ void __launch_bounds__(ST_THREADS, (StemCfg<KT, NP>::CTAS))
stem_conv_kernel(const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  using C = StemCfg<KT, NP>;
  uint8_t* sB = smem;
  __half* patch = reinterpret_cast<__half*>(sB + C::B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(patch) + C::PATCH_BYTES);
  uint64_t* full = bars;                       // [ST_STAGES]  builders -> MMA   (one arrive per builder warp)
  uint64_t* empty = bars + ST_STAGES;          // [ST_STAGES]  MMA commit -> builders
  uint64_t* accfull = bars + 2 * ST_STAGES;    // [2]          MMA commit -> epilogue
  uint64_t* accempty = accfull + 2;            // [2]          epilogue -> MMA     (one arrive per builder warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < ST_STAGES; ++i) { mbar_init(&full[i], 4); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accempty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();
  // weights -> 128B-swizzled K-major B tiles, one [NP x 64] tile per K-block
  for (int i = tid; i < C::KB * NP * 8; i += ST_THREADS) {
    const int chunk = i & 7, n = (i >> 3) % NP, kb = i / (8 * NP);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w) + i);
    *reinterpret_cast<uint4*>(sB + kb * (NP * 128) + (n >> 3) * 1024 + (n & 7) * 128 + ((chunk ^ (n & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 4) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, NP, 0, 0);
      uint32_t it = 0, n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int seg = (item / (p.tx * p.ty)) % p.nseg;
        const int t_begin = seg * p.tseg, t_end = min(t_begin + p.tseg, p.T);
        for (int t = t_begin; t < t_end; ++t, ++n) {
          const uint32_t acc = n & 1;
          if (n >= 2) mbar_wait(&accempty[acc], ((n >> 1) - 1) & 1);
          tc_fence_after();
          for (int kb = 0; kb < C::KB; ++kb, ++it) {
            const uint32_t s = it % ST_STAGES;
            mbar_wait(&full[s], (it / ST_STAGES) & 1);
            tc_fence_after();
            const uint32_t ta = tmem_acc + C::T_A + s * ST_A_COLS;
            const uint64_t db = umma_smem_desc(smem_u32(sB + kb * (NP * 128)), 16, 1024, UMMA_SW_128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(tmem_acc + acc * NP, ta + 8 * k, db + static_cast<uint64_t>(2 * k), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          umma_commit(&accfull[acc]);
        }
      }
    }
  } else {
    // ---------------- builders / epilogue ----------------
    const int py = tid >> 4, px = tid & 15;
    const long long plane = static_cast<long long>(p.H) * p.W;
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t it = 0, n = 0;
    long long prev_row = -1;
    bool have_prev = false;

    auto epilogue = [&](uint32_t m, long long row) {
      const uint32_t acc = m & 1;
      mbar_wait(&accfull[acc], (m >> 1) & 1);
      tc_fence_after();
      if constexpr (COUT == 8) {
        uint32_t a[8];
        tmem_ld_x8(taddr + acc * NP, a);
        tmem_wait_ld();
        uint32_t h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          h[j] = pack_half2(fmaxf(__uint_as_float(a[2 * j]) + __ldg(p.shift + 2 * j), 0.f),
                            fmaxf(__uint_as_float(a[2 * j + 1]) + __ldg(p.shift + 2 * j + 1), 0.f));
        if (row >= 0) *reinterpret_cast<uint4*>(p.out + row * COUT) = make_uint4(h[0], h[1], h[2], h[3]);
      } else {
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 32) {
          uint32_t a[32];
          tmem_ld_x32(taddr + acc * NP + c0, a);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              h[e] = pack_half2(fmaxf(__uint_as_float(a[8 * j + 2 * e]) + __ldg(p.shift + c0 + 8 * j + 2 * e), 0.f),
                                fmaxf(__uint_as_float(a[8 * j + 2 * e + 1]) + __ldg(p.shift + c0 + 8 * j + 2 * e + 1), 0.f));
            if (row >= 0) *reinterpret_cast<uint4*>(p.out + row * COUT + c0 + 8 * j) = make_uint4(h[0], h[1], h[2], h[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accempty[acc]);
    };

    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int r = item;
      const int xt = r % p.tx; r /= p.tx;
      const int yt = r % p.ty; r /= p.ty;
      const int seg = r % p.nseg;
      const int nclip = r / p.nseg;
      const int y0 = yt * ST_TY, x0 = xt * ST_TX;
      const int t_begin = seg * p.tseg, t_end = min(t_begin + p.tseg, p.T);
      const float* xn = p.x + static_cast<long long>(nclip) * 3 * p.T * plane;

      // halo of input frame f -> ring slot (f mod KT); zeros outside the clip / frame.  A thread owns one column pair
      // (19 pairs x 6 row groups = 114 active threads) and 3 x 4 halo rows.  The global loads of frame t+1's halo are
      // issued BEFORE the K-blocks of frame t are built and land in registers meanwhile; they are converted and
      // stored after the build (the ring slot they replace is still being read until then).
      const int lf_pair = tid % 19, lf_grp = tid / 19;
      const int lf_ix = 2 * x0 - 3 + 2 * lf_pair;
      const bool lf_ok0 = lf_grp < 6 && lf_ix >= 0 && lf_ix < p.W, lf_ok1 = lf_grp < 6 && lf_ix + 1 >= 0 && lf_ix + 1 < p.W;
      constexpr int LF_ROWS = (ST_PR + 5) / 6;     // 4
      float hv[3][LF_ROWS][2];
      auto fetch_frame = [&](int f) {
        const bool f_ok = f >= 0 && f < p.T;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* src = xn + (static_cast<long long>(c) * p.T + f) * plane + lf_ix;
#pragma unroll
          for (int k = 0; k < LF_ROWS; ++k) {
            const int iy = 2 * y0 - 3 + lf_grp + 6 * k;
            const bool ok = f_ok && iy >= 0 && iy < p.H && lf_grp + 6 * k < ST_PR;
            const float* row = src + static_cast<long long>(iy) * p.W;
            hv[c][k][0] = (ok && lf_ok0) ? __ldg(row) : 0.f;
            hv[c][k][1] = (ok && lf_ok1) ? __ldg(row + 1) : 0.f;
          }
        }
      };
      auto store_frame = [&](int f) {
        const int slot = ((f % KT) + KT) % KT;
        uint32_t* dst = reinterpret_cast<uint32_t*>(patch + slot * ST_SLOT) + lf_pair;
        if (lf_grp < 6) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < LF_ROWS; ++k)
              if (lf_grp + 6 * k < ST_PR) dst[(c * ST_PR + lf_grp + 6 * k) * (ST_PC / 2)] = pack_half2(hv[c][k][0], hv[c][k][1]);
        }
      };
      named_bar_sync(1, ST_BUILDERS);   // every builder is done reading the previous item's ring
      for (int f = t_begin - KT / 2; f < t_begin + KT / 2; ++f) { fetch_frame(f); store_frame(f); }
      fetch_frame(t_begin + KT / 2);

      for (int t = t_begin; t < t_end; ++t, ++n) {
        if (t > t_begin) named_bar_sync(1, ST_BUILDERS);   // the slot about to be overwritten was read by frame t-1
        store_frame(t + KT / 2);
        named_bar_sync(1, ST_BUILDERS);
        if (t + 1 < t_end) fetch_frame(t + 1 + KT / 2);    // in flight while this frame's K-blocks are built
        for (int kb = 0; kb < C::KB; ++kb, ++it) {
          const int dt = kb / 3, c = kb - dt * 3;
          const uint32_t s = it % ST_STAGES;
          if (it >= ST_STAGES) mbar_wait(&empty[s], ((it / ST_STAGES) - 1) & 1);   // MMAs that read this stage are done
          const int f = t + dt - KT / 2;
          const int slot = ((f % KT) + KT) % KT;
          const __half* src = patch + slot * ST_SLOT + c * (ST_PR * ST_PC) + (2 * py) * ST_PC + 2 * px;
          uint32_t a[32];                       // this pixel's 64-half K-block row: column j = halfs (2j, 2j+1)
#pragma unroll
          for (int dy = 0; dy < 8; ++dy) {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + dy * ST_PC);
#pragma unroll
            for (int j = 0; j < 4; ++j) a[4 * dy + j] = s32[j];
          }
          tmem_st_x32(taddr + C::T_A + s * ST_A_COLS, a);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[s]);
        }
        if (have_prev) epilogue(n - 1, prev_row);
        const int y = y0 + py, x = x0 + px;
        prev_row = (y < p.Hs && x < p.Ws) ? ((static_cast<long long>(nclip) * p.T + t) * p.Hs + y) * p.Ws + x : -1;
        have_prev = true;
      }
    }
    if (have_prev) epilogue(n - 1, prev_row);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, C::TMEM_COLS);
  }
}

template <int KT, int NP, int COUT>
int launch_stem_t(const StemParams& p, cudaStream_t stream) {
  using C = StemCfg<KT, NP>;
  static bool configured = false;
  if (!configured) {
    KVQ_CUDA(cudaFuncSetAttribute(stem_conv_kernel<KT, NP, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  const int grid = std::min(p.items, C::CTAS * num_sms());
  count_launch();
  return launch_pdl(stem_conv_kernel<KT, NP, COUT>, dim3(grid), dim3(ST_THREADS), C::SMEM, stream, p);
}

}  // namespace

int stem_weight_rows(int cout) { return cout <= 16 ? 16 : 64; }

int launch_stem_conv(const float* x, const __half* w_packed, const float* shift, __half* out, int N, int T, int H, int W,
                     int kt, int cout, cudaStream_t stream) {
  KVQ_REQUIRE(x && w_packed && shift && out, KVQ_ERR_BAD_SHAPE, "stem_conv: NULL argument");
  KVQ_REQUIRE(N > 0 && T > 0 && H > 0 && W > 0, KVQ_ERR_BAD_SHAPE, "stem_conv: empty input %dx3x%dx%dx%d", N, T, H, W);
  KVQ_REQUIRE((kt == 1 && cout == 64) || (kt == 5 && cout == 8), KVQ_ERR_BAD_SHAPE,
              "stem_conv: built for (kt=1, cout=64) and (kt=5, cout=8); got kt=%d cout=%d", kt, cout);
  StemParams p{};
  p.x = x; p.w = w_packed; p.shift = shift; p.out = out;
  p.N = N; p.T = T; p.H = H; p.W = W;
  p.Hs = (H + 6 - 7) / 2 + 1; p.Ws = (W + 6 - 7) / 2 + 1;
  p.ty = (p.Hs + ST_TY - 1) / ST_TY; p.tx = (p.Ws + ST_TX - 1) / ST_TX;
  // frames of a work item share the (kt-1)-frame halo ring; split T only as far as needed to fill the machine
  int tseg = kt == 1 ? 1 : T;
  const long long per = static_cast<long long>(N) * p.ty * p.tx;
  while (kt > 1 && tseg > 4 && per * ((T + tseg - 1) / tseg) < 4LL * num_sms()) tseg = (tseg + 1) / 2;
  p.tseg = tseg;
  p.nseg = (T + tseg - 1) / tseg;
  const long long items = per * p.nseg;
  KVQ_REQUIRE(items < (1LL << 31), KVQ_ERR_BAD_SHAPE, "stem_conv: %lld work items", items);
  p.items = static_cast<int>(items);
  if (kt == 1) return launch_stem_t<1, 64, 64>(p, stream);
  return launch_stem_t<5, 16, 8>(p, stream);
}

}  // namespace kvq
