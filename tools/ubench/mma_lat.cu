// Micro-benchmark: cost of small tcgen05.mma instructions (M128, K16, fp16) issued back to back by one thread.
// Reports cycles per MMA for SS (A from smem) and TS (A from TMEM) forms at several N, accumulating into one D or
// rotating over independent D tiles, from 1 or 3 issuing threads (warps) at once.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../kvq-challenge-cvpr-ntire2024_b200/csrc -o mma_lat mma_lat.cu
#include <cstdio>
#include "kvq_common.cuh"
using namespace kvq;
namespace kvq { void set_error(const char*, ...) {} int check_cuda(cudaError_t e, const char*) { return e != cudaSuccess; } bool pdl_enabled() { return false; } }

template <int MODE>
__global__ void __launch_bounds__(128, 1) mma_bench(long long* out, int N, int ts, int nmma, int rotate, int issuers, int sw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tslot;
  // A-in-TMEM operand: zeros are fine
  { uint32_t z[32]; for (int i = 0; i < 32; ++i) z[i] = 0x3C003C00u; tmem_st_x32(tb + (static_cast<uint32_t>(warp * 32) << 16) + 448, z); tmem_wait_st(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (MODE == 1) {
    // warp-uniform control flow: every lane of the issuing warp runs the loop, one elected lane issues
    if (warp < issuers) {
      const uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
      const uint64_t da = umma_smem_desc(smem_u32(smem), 128, 256, UMMA_SW_NONE);
      const uint64_t db = umma_smem_desc(smem_u32(smem) + 16384, 128, 256, UMMA_SW_NONE);
      const uint32_t dbase = tb + warp * 128;
      if (elect_one()) { umma_f16_ss(dbase, da, db, idesc, 0); umma_commit(&bar[warp]); }
      mbar_wait(&bar[warp], 0);
      tc_fence_after();
      long long t0 = clock64();
      for (int i = 0; i < nmma; ++i) {
        const uint32_t d = dbase + (rotate ? (i % rotate) * ((N + 7) / 8 * 8) : 0);
        if (elect_one()) {
          if (ts) umma_f16_ts(d, tb + 448 + 8 * (i & 3), db, idesc, 1);
          else umma_f16_ss(d, da, db, idesc, 1);
        }
      }
      long long t1 = clock64();
      if (elect_one()) umma_commit(&bar[warp]);
      mbar_wait(&bar[warp], 1);
      long long t2 = clock64();
      if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
    }
  } else if (MODE == 2) {
    // everything warp-uniform and compile-time where possible: TMEM base assumed 0 (sole CTA, 512 columns), one
    // elect around a fully unrolled batch of 16 MMAs
    if (warp < issuers) {
      const uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t da = sw ? umma_smem_desc(sbase, 16, 1024, UMMA_SW_128) : umma_smem_desc(sbase, 128, 256, UMMA_SW_NONE);
      const uint64_t db = sw ? umma_smem_desc(sbase + 16384, 16, 1024, UMMA_SW_128) : umma_smem_desc(sbase + 16384, 128, 256, UMMA_SW_NONE);
      const uint32_t dbase = warp * 128;
      if (elect_one()) { umma_f16_ss(dbase, da, db, idesc, 0); umma_commit(&bar[warp]); }
      mbar_wait(&bar[warp], 0);
      tc_fence_after();
      long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < nmma; i += 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (ts) umma_f16_ts(dbase, 448 + 8 * (j & 3), db + (sw ? 2 * (j & 3) : 0), idesc, 1);
            else umma_f16_ss(dbase, da + (sw ? 2 * (j & 3) : 0), db + (sw ? 2 * (j & 3) : 0), idesc, 1);
          }
        }
      }
      __syncwarp();
      long long t1 = clock64();
      if (elect_one()) umma_commit(&bar[warp]);
      mbar_wait(&bar[warp], 1);
      long long t2 = clock64();
      if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
    }
  } else if (warp < issuers && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N, 0, 0);
    const uint64_t da = umma_smem_desc(smem_u32(smem), 128, 256, UMMA_SW_NONE);
    const uint64_t db = umma_smem_desc(smem_u32(smem) + 16384, 128, 256, UMMA_SW_NONE);
    const uint32_t dbase = tb + warp * 128;   // each issuer its own D region (<= 128 cols)
    // warm-up
    umma_f16_ss(dbase, da, db, idesc, 0);
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    tc_fence_after();
    long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const uint32_t d = dbase + (rotate ? (i % rotate) * ((N + 7) / 8 * 8) : 0);
      if (ts) umma_f16_ts(d, tb + 448 + 8 * (i & 3), db, idesc, 1);
      else umma_f16_ss(d, da, db, idesc, 1);
    }
    long long t1 = clock64();
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 1);
    long long t2 = clock64();
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(mma_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(mma_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(mma_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int nmma = 256;
  for (int sw : {0, 1})
  for (int issuers : {1, 2})
    for (int ts : {0, 1})
      for (int N : {64, 96, 128, 256}) {
        if (issuers * N > 512 - 64) continue;
        cudaMemset(out, 0, 64);
        mma_bench<2><<<1, 128, 100 * 1024>>>(out, N, ts, nmma, 0, issuers, sw);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
        printf("%s issuers %d %s N=%3d : issue %.1f clk/MMA, complete %.1f clk/MMA (math %.1f) [%s]\n", sw ? "SW128 K-major" : "no-swizzle   ", issuers, ts ? "TS" : "SS", N,
               double(h[0]) / nmma, double(h[1]) / nmma, 128.0 * N / 256, cudaGetErrorString(e));
      }
  return 0;
}
