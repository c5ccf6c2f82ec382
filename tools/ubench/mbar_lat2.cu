// Does a wait on a COMPLETED mbarrier phase slow down when other warps of the CTA are parked on incomplete barriers?
// Warp 0 lane 0 times 256 waits on a completed phase; warps 1..W wait on another barrier that completes only at the end.
// Parked flavours: 0 = try_wait with 20 us suspend hint, 1 = try_wait without hint, 2 = test_wait spin,
//                  3 = test_wait + __nanosleep(64) back-off
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mbar_lat2 mbar_lat2.cu && ./mbar_lat2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool try_hint(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool try_nohint(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void k(long long* out, int parked_warps, int parked_all_lanes, int flavour, int timer_flavour) {
  __shared__ uint64_t bars[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
  }
  __syncthreads();
  const uint32_t done = smem_u32(&bars[0]), pending = smem_u32(&bars[1]);
  if (warp == 0) {
    if (lane == 0) {
      for (volatile int spin = 0; spin < 2000; ++spin) {}   // let the others park first
      long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < 256; ++i) {
        if (timer_flavour == 0) while (!try_hint(done, 0)) {}
        else while (!test(done, 0)) {}
      }
      long long t1 = clock64();
      out[0] = (t1 - t0) / 256;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pending) : "memory");   // release the parked warps
    }
  } else if (warp <= parked_warps) {
    if (lane == 0 || parked_all_lanes) {
      if (flavour == 0) while (!try_hint(pending, 0)) {}
      else if (flavour == 1) while (!try_nohint(pending, 0)) {}
      else if (flavour == 2) while (!test(pending, 0)) {}
      else while (!test(pending, 0)) __nanosleep(64);
    }
  }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  const char* names[4] = {"try_wait + 20us hint", "try_wait, no hint", "test_wait spin", "test_wait + nanosleep(64)"};
  for (int tf = 0; tf < 2; ++tf)
    for (int parked : {0, 1, 4, 12})
      for (int all : {0, 1})
        for (int f = 0; f < 4; ++f) {
          if (parked == 0 && (all || f)) continue;
          k<<<1, 512>>>(out, parked, all, f, tf);
          cudaError_t e = cudaDeviceSynchronize();
          printf("timer=%s  parked warps %2d (%s) flavour %-26s : %5lld cycles per completed-phase wait%s\n", tf ? "test_wait" : "try_wait ", parked,
                 all ? "32 lanes" : "lane 0  ", names[f], out[0], e == cudaSuccess ? "" : "  [CUDA ERROR]");
        }
  return 0;
}
