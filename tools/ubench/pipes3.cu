// Micro-benchmark 3: scalar vs packed fp32 on the FMA pipe without register-bank conflicts, and the softmax instruction
// mixes with either.  Prints warp-instructions (and fp32 elements) per clock per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
#define U 8
template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cyc, float seed) {
  float a[U], b[U];
  unsigned long long p[U], q[U];
  uint32_t h[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    a[i] = seed + i + threadIdx.x * 1e-6f; b[i] = seed * 0.5f + i;
    p[i] = (static_cast<unsigned long long>(__float_as_uint(seed + i)) << 32) | __float_as_uint(seed * 0.5f + i);
    q[i] = (static_cast<unsigned long long>(__float_as_uint(0.5f + i)) << 32) | __float_as_uint(0.25f * i);
    h[i] = i;
  }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (MODE == 0) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i])); }
      else if (MODE == 1) { asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(b[i])); }
      else if (MODE == 2) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b[i])); }
      else if (MODE == 3) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i])); }
      else if (MODE == 4) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(q[i])); }
      else if (MODE == 5) {   // per 2 logits, packed: FFMA2 + 3 FADD2 + 2 MUFU + F2FP + FMNMX3
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(q[i]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[(i + 1) % U]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[(i + 2) % U]) : "l"(p[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[(i + 3) % U]) : "f"(b[(i + 1) % U]), "f"(b[(i + 2) % U]));
      } else if (MODE == 6) {   // same work, scalar: 2 FFMA + 6 FADD + 2 MUFU + F2FP + FMNMX3
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[(i + 4) % U]) : "f"(b[(i + 4) % U]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[(i + 5) % U]) : "f"(b[(i + 5) % U]));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[(i + 4) % U]) : "f"(b[(i + 6) % U]));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[(i + 5) % U]) : "f"(b[(i + 7) % U]));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[(i + 4) % U]) : "f"(b[(i + 1) % U]));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[(i + 5) % U]) : "f"(b[(i + 2) % U]));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b[(i + 6) % U]) : "f"(a[(i + 4) % U]));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b[(i + 7) % U]) : "f"(a[(i + 5) % U]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[(i + 3) % U]) : "f"(b[(i + 1) % U]), "f"(b[(i + 2) % U]));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < U; ++i) s += a[i] + b[i] + __uint_as_float(static_cast<uint32_t>(p[i])) + __uint_as_float(static_cast<uint32_t>(q[i] >> 32)) + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, double ipu, float* out, long long* cyc) {
  for (int warps : {12, 16, 32}) {
    bench<MODE><<<148, warps * 32>>>(out, cyc, 0.25f);
    cudaDeviceSynchronize();
    bench<MODE><<<148, warps * 32>>>(out, cyc, 0.25f);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[148];
    cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += c[i];
    avg /= 148;
    printf("%-44s warps/SM %2d  warp-instr/clk/SMSP %.3f  clk per unit per SMSP %.2f (%s)\n", name, warps,
           double(ITER) * U * ipu * warps / avg / 4.0, avg * 4.0 / (double(ITER) * U * warps), cudaGetErrorString(e));
  }
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("FADD a+=b", 1, out, cyc);
  run<1>("FFMA a=a*b+a", 1, out, cyc);
  run<2>("FFMA a=a*b+b", 1, out, cyc);
  run<3>("FADD2", 1, out, cyc);
  run<4>("FFMA2", 1, out, cyc);
  run<5>("softmax pair, packed (8 instr / 2 logits)", 8, out, cyc);
  run<6>("softmax pair, scalar (12 instr / 2 logits)", 12, out, cyc);
  return 0;
}
