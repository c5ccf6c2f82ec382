// Latency of an mbarrier wait on an ALREADY COMPLETED phase, per flavour, with and without FMA-heavy neighbours.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mbar_lat mbar_lat.cu && ./mbar_lat
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__device__ __forceinline__ bool wait_once(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  if constexpr (MODE == 0)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  else if constexpr (MODE == 1)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

template <int MODE>
__global__ void k(long long* out, int busy_warps, int all_lanes, float* sink) {
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0 || all_lanes) {
      long long t0 = clock64();
      int fails = 0;
#pragma unroll 1
      for (int i = 0; i < 256; ++i) {
        while (!wait_once<MODE>(smem_u32(&bar), 0)) ++fails;
        if (all_lanes) __syncwarp();
      }
      long long t1 = clock64();
      if (lane == 0) { out[0] = (t1 - t0) / 256; out[1] = fails; }
    }
  } else if (warp <= busy_warps) {
    float a = threadIdx.x * 0.001f, b = 1.0001f, c0 = 0.f, c1 = 1.f, c2 = 2.f, c3 = 3.f;
#pragma unroll 1
    for (int i = 0; i < 40000; ++i) {
      c0 = fmaf(c0, b, a); c1 = fmaf(c1, b, a); c2 = fmaf(c2, b, a); c3 = fmaf(c3, b, a);
      c0 = fmaf(c0, b, a); c1 = fmaf(c1, b, a); c2 = fmaf(c2, b, a); c3 = fmaf(c3, b, a);
    }
    sink[threadIdx.x] = c0 + c1 + c2 + c3;
  }
}

int main() {
  long long* out; float* sink;
  cudaMallocManaged(&out, 16); cudaMalloc(&sink, 4096 * 4);
  const char* names[3] = {"try_wait hint 20000ns", "try_wait (no hint)", "test_wait"};
  for (int busy : {0, 4, 12}) {
    for (int all : {0, 1}) {
      for (int m = 0; m < 3; ++m) {
        if (m == 0) k<0><<<1, 512>>>(out, busy, all, sink);
        if (m == 1) k<1><<<1, 512>>>(out, busy, all, sink);
        if (m == 2) k<2><<<1, 512>>>(out, busy, all, sink);
        cudaDeviceSynchronize();
        printf("busy FMA warps %2d  %s  %-22s : %lld cycles per wait on a completed phase (%lld failed polls)\n", busy, all ? "all lanes + syncwarp" : "lane 0 only        ", names[m], out[0], out[1]);
      }
    }
  }
  return 0;
}
