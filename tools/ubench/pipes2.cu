// Micro-benchmark, second batch: packed fp32x2 ops with operands held in 64-bit registers (no pack/unpack moves),
// MUFU.EX2.F16, LEA-style shift-add, and MUFU + LDS.128 + FFMA2 mixes.  warp-instructions per clock per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
#define U 8
template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cyc, float seed) {
  __shared__ float4 sh[2048];
  sh[threadIdx.x] = make_float4(seed, seed + 1, seed + 2, seed + 3);
  sh[threadIdx.x + 1024] = sh[threadIdx.x];
  __syncthreads();
  unsigned long long a[U], b[U];
  float f[U];
  uint32_t h[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    a[i] = (static_cast<unsigned long long>(__float_as_uint(seed + i)) << 32) | __float_as_uint(seed * 0.5f + i + threadIdx.x * 1e-6f);
    b[i] = (static_cast<unsigned long long>(__float_as_uint(0.5f + i)) << 32) | __float_as_uint(0.25f * i);
    f[i] = seed + i;
    h[i] = 0x3C003C00u + i;
  }
  uint32_t saddr = static_cast<uint32_t>(__cvta_generic_to_shared(sh)) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (MODE == 0) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[i]) : "l"(b[i]));
      } else if (MODE == 1) {
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b[i]));
      } else if (MODE == 2) {
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b[i]));
      } else if (MODE == 3) {   // MUFU.EX2.F16 (one half)
        asm volatile("{.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %0;\n\tex2.approx.f16 lo, lo;\n\tmov.b32 %0, {lo, hi};}" : "+r"(h[i]));
      } else if (MODE == 4) {   // shift-add (LEA / IMAD.SHL)
        asm volatile("{.reg .b32 t;\n\tshl.b32 t, %1, 23;\n\tadd.u32 %0, %0, t;}" : "+r"(h[i]) : "r"(h[(i + 1) % U]));
      } else if (MODE == 5) {   // MUFU + LDS.128 1 : 0.5
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        if (i & 1) {
          float4 v;
          asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr ^ ((it & 1) << 14)));
          a[i] = (static_cast<unsigned long long>(__float_as_uint(v.x)) << 32) | __float_as_uint(v.y);
          b[i] = (static_cast<unsigned long long>(__float_as_uint(v.z)) << 32) | __float_as_uint(v.w);
        }
      } else if (MODE == 6) {   // softmax-like mix per 2 logits: LDS.128, FFMA2, FADD2, FMNMX3, FADD2, 2 MUFU, F2FP, FADD2
        float4 v;
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr ^ ((it & 1) << 14)));
        unsigned long long t0_ = (static_cast<unsigned long long>(__float_as_uint(v.x)) << 32) | __float_as_uint(v.y);
        unsigned long long t1_ = (static_cast<unsigned long long>(__float_as_uint(v.z)) << 32) | __float_as_uint(v.w);
        unsigned long long x = a[i];
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x) : "l"(b[i]), "l"(t1_));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(t0_));
        float x0 = __uint_as_float(static_cast<uint32_t>(x)), x1 = __uint_as_float(static_cast<uint32_t>(x >> 32));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[0]) : "f"(x0), "f"(x1));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(b[(i + 1) % U]));
        x0 = __uint_as_float(static_cast<uint32_t>(x)); x1 = __uint_as_float(static_cast<uint32_t>(x >> 32));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x1), "f"(x0));
        unsigned long long pe = (static_cast<unsigned long long>(__float_as_uint(x1)) << 32) | __float_as_uint(x0);
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[(i + 3) % U]) : "l"(pe));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < U; ++i) s += __uint_as_float(static_cast<uint32_t>(a[i])) + __uint_as_float(static_cast<uint32_t>(b[i] >> 32)) + f[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, double ipu, float* out, long long* cyc) {
  for (int warps : {4, 12, 16, 32}) {
    bench<MODE><<<148, warps * 32>>>(out, cyc, 0.25f);
    cudaDeviceSynchronize();
    bench<MODE><<<148, warps * 32>>>(out, cyc, 0.25f);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[148];
    cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += c[i];
    avg /= 148;
    double winstr = double(ITER) * U * ipu * warps;
    printf("%-34s warps/SM %2d  cycles %9.0f  warp-instr/clk/SMSP %.3f  units/clk/SMSP %.4f (%s)\n", name, warps, avg, winstr / avg / 4.0,
           double(ITER) * U * warps / avg / 4.0, cudaGetErrorString(e));
  }
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("FFMA2 (64-bit regs)", 1, out, cyc);
  run<1>("FADD2 (64-bit regs)", 1, out, cyc);
  run<2>("FMUL2 (64-bit regs)", 1, out, cyc);
  run<3>("MUFU.EX2.F16", 1, out, cyc);
  run<4>("SHL+IADD", 2, out, cyc);
  run<5>("MUFU + 0.5 LDS.128", 1.5, out, cyc);
  run<6>("softmax mix (10 instr / 2 logits)", 10, out, cyc);
  return 0;
}
