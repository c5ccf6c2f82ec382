// Host-side plumbing shared by every entry point: last-error string, CUDA error capture, and the
// driver-API tensor-map encoder (resolved through the runtime so the library does not link libcuda).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "kvq_common.cuh"

namespace kvq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return KVQ_OK;
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return KVQ_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_rows, uint32_t box_cols, int elem_bytes, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  KVQ_REQUIRE(enc != nullptr, KVQ_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  KVQ_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, KVQ_ERR_MISALIGNED, "TMA base %p not 16 B aligned", base);
  KVQ_REQUIRE((row_stride_bytes & 15) == 0, KVQ_ERR_MISALIGNED, "TMA row stride %llu not a multiple of 16 B",
              static_cast<unsigned long long>(row_stride_bytes));
  KVQ_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_cols >= 1 && box_cols <= 256, KVQ_ERR_BAD_SHAPE,
              "TMA box %ux%u out of range", box_rows, box_cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KVQ_REQUIRE(r == CUDA_SUCCESS, KVQ_ERR_DRIVER,
              "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu stride=%llu box=%ux%u", static_cast<int>(r),
              static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols),
              static_cast<unsigned long long>(row_stride_bytes), box_rows, box_cols);
  return KVQ_OK;
}

// channels-last activation [B,T,H,W,C] fp16 as a 5-D tensor (C innermost); box = 64 channels x (bw x bh x bt) pixels of
// one clip, traversed with the convolution stride (TMA loads ceil(box/stride) elements per dim), 128 B swizzle.
// Out-of-range coordinates (negative or past the extent) are zero-filled = the convolution's zero padding.
int make_tmap_conv5d(CUtensorMap* out, const void* base, int B, int T, int H, int W, int C, int bt, int bh, int bw,
                     int st, int sh, int sw) {
  EncodeTiledFn enc = get_encode();
  KVQ_REQUIRE(enc != nullptr, KVQ_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  KVQ_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (C % 64 == 0 || C == 8 || C == 16 || C == 32),
              KVQ_ERR_MISALIGNED, "conv TMA: base %p must be 16 B aligned and C=%d a multiple of 64 (or 8/16/32)", base, C);
  KVQ_REQUIRE(bt * st <= 256 && bh * sh <= 256 && bw * sw <= 256, KVQ_ERR_BAD_SHAPE, "conv TMA: box too large");
  cuuint64_t gdim[5] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[4] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                        static_cast<cuuint64_t>(H) * W * C * 2, static_cast<cuuint64_t>(T) * H * W * C * 2};
  // narrow activations: the box is the whole channel extent, swizzle span = row bytes (16 B rows: no swizzle)
  const int cbox = C < 64 ? C : 64;
  const CUtensorMapSwizzle swz = cbox == 64   ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : cbox == 16 ? CU_TENSOR_MAP_SWIZZLE_32B
                                              : CU_TENSOR_MAP_SWIZZLE_NONE;
  cuuint32_t box[5] = {static_cast<cuuint32_t>(cbox), static_cast<cuuint32_t>(bw * sw), static_cast<cuuint32_t>(bh * sh),
                       static_cast<cuuint32_t>(bt * st), 1};
  cuuint32_t estr[5] = {1, static_cast<cuuint32_t>(sw), static_cast<cuuint32_t>(sh), static_cast<cuuint32_t>(st), 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KVQ_REQUIRE(r == CUDA_SUCCESS, KVQ_ERR_DRIVER,
              "cuTensorMapEncodeTiled(conv) failed (%d) dims=%dx%dx%dx%dx%d box=%dx%dx%d stride=%dx%dx%d",
              static_cast<int>(r), B, T, H, W, C, bt, bh, bw, st, sh, sw);
  return KVQ_OK;
}

int make_tmap_out5d(CUtensorMap* out, const void* base, int B, int To, int Ho, int Wo, int cols, int ldo, int bt, int bh,
                    int bw) {
  EncodeTiledFn enc = get_encode();
  KVQ_REQUIRE(enc != nullptr, KVQ_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  KVQ_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ldo % 8 == 0 && cols >= 1, KVQ_ERR_MISALIGNED,
              "out TMA: base %p / row stride %d halfs must be 16 B aligned", base, ldo);
  cuuint64_t gdim[5] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Ho),
                        static_cast<cuuint64_t>(To), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[4] = {static_cast<cuuint64_t>(ldo) * 2, static_cast<cuuint64_t>(Wo) * ldo * 2,
                        static_cast<cuuint64_t>(Ho) * Wo * ldo * 2, static_cast<cuuint64_t>(To) * Ho * Wo * ldo * 2};
  cuuint32_t box[5] = {32, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), static_cast<cuuint32_t>(bt), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KVQ_REQUIRE(r == CUDA_SUCCESS, KVQ_ERR_DRIVER,
              "cuTensorMapEncodeTiled(out) failed (%d) dims=%dx%dx%dx%dx%d ldo=%d box=%dx%dx%d", static_cast<int>(r), B,
              To, Ho, Wo, cols, ldo, bt, bh, bw);
  return KVQ_OK;
}

// ---- optional per-kernel-category device timing (bench.py's roofline leg) and a launch counter ----
namespace {
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  std::vector<int> cat;  // category of pair i (events 2i, 2i+1)
  size_t used = 0;
};
ProfState g_prof;
long long g_launches = 0;
}  // namespace

void count_launch() { ++g_launches; }
long long launch_count() { return g_launches; }

bool prof_enabled() { return g_prof.on; }

void prof_set(bool on) {
  g_prof.on = on;
  g_prof.used = 0;
  g_prof.cat.clear();
}

void prof_mark(int cat, bool begin, cudaStream_t st) {
  if (!g_prof.on) return;
  if (begin) g_prof.cat.push_back(cat);
  if (g_prof.used >= g_prof.pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { g_prof.on = false; return; }
    g_prof.pool.push_back(e);
  }
  cudaEventRecord(g_prof.pool[g_prof.used++], st);
}

int prof_collect(float* ms, int* launches, int ncat) {
  for (int i = 0; i < ncat; ++i) { ms[i] = 0.f; launches[i] = 0; }
  const size_t pairs = g_prof.used / 2;
  if (pairs == 0) return 0;
  if (cudaEventSynchronize(g_prof.pool[g_prof.used - 1]) != cudaSuccess) return KVQ_ERR_CUDA;
  for (size_t i = 0; i < pairs; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]) != cudaSuccess) return KVQ_ERR_CUDA;
    const int c = g_prof.cat[i];
    if (c >= 0 && c < ncat) { ms[c] += t; launches[c] += 1; }
  }
  g_prof.used = 0;
  g_prof.cat.clear();
  return static_cast<int>(pairs);
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("KVQ_PDL");
    on = (e != nullptr && atoi(e) != 0) ? 1 : 0;   // measured on B200: -3.6 % with PDL on, so it is opt-in
  }
  return on == 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace kvq
