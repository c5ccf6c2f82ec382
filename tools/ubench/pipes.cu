// Micro-benchmark of the instruction pipes the attention softmax path leans on (B200, sm_100a).
// Prints warp-instructions per clock per SM sub-partition for each instruction class / mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define U 8

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cyc, float seed) {
  __shared__ float4 sh[2048];
  sh[threadIdx.x] = make_float4(seed, seed + 1, seed + 2, seed + 3); sh[threadIdx.x + 1024] = sh[threadIdx.x];
  __syncthreads();
  float a[U], b[U], c[U];
  uint32_t h[U];
#pragma unroll
  for (int i = 0; i < U; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; b[i] = seed * 0.5f + i; h[i] = i; c[i] = seed - i; }
  uint32_t saddr = static_cast<uint32_t>(__cvta_generic_to_shared(sh)) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (MODE == 0) {  // MUFU.EX2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {  // F2FP pack
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
        a[i] = __uint_as_float(h[i]);
      } else if (MODE == 2) {  // FMNMX3
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) % U]));
      } else if (MODE == 3) {  // FMNMX
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      } else if (MODE == 4) {  // FFMA
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[(i + 1) % U]));
      } else if (MODE == 5) {  // FFMA2 (two floats per instr)
        asm volatile("{.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tfma.rn.f32x2 x, x, y, y;\n\tmov.b64 {%0, %1}, x;}"
                     : "+f"(a[i]), "+f"(b[i]) : "f"(a[(i + 1) % U]), "f"(b[(i + 3) % U]));
      } else if (MODE == 6) {  // FADD2
        asm volatile("{.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;}"
                     : "+f"(a[i]), "+f"(b[i]) : "f"(a[(i + 1) % U]), "f"(b[(i + 3) % U]));
      } else if (MODE == 7) {  // IMAD
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(h[(i + 1) % U] | 3u), "r"(h[(i + 2) % U]));
      } else if (MODE == 8) {  // LDS.128 conflict-free
        float4 v;
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr ^ ((it & 1) << 14)));
        a[i] = v.x; b[i] = v.y; c[i] = v.z; h[i] = __float_as_uint(v.w);
      } else if (MODE == 9) {  // LDS.64
        float2 v;
        asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((saddr - (threadIdx.x & 31) * 8) ^ ((it & 1) << 14)));
        a[i] = v.x; b[i] = v.y;
      } else if (MODE == 10) {  // MUFU + F2FP mix (1 : 0.5)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i & 1) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[i - 1])); b[i] = __uint_as_float(h[i]); }
      } else if (MODE == 11) {  // MUFU + FFMA mix 1:1
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(b[i]) : "f"(b[(i + 1) % U]));
      } else if (MODE == 12) {  // F2FP + FFMA mix 1:1
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(b[i]) : "f"(b[(i + 1) % U]));
      } else if (MODE == 13) {  // F2FP + FMNMX mix 1:1
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(b[i]));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(b[(i + 1) % U]));
      } else if (MODE == 14) {  // FMNMX + FFMA mix 1:1
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % U]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(b[i]) : "f"(b[(i + 1) % U]));
      } else if (MODE == 15) {  // FFMA2 + FMNMX3 1:1
        asm volatile("{.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tfma.rn.f32x2 x, x, y, y;\n\tmov.b64 {%0, %1}, x;}"
                     : "+f"(a[i]), "+f"(b[i]) : "f"(a[(i + 1) % U]), "f"(b[(i + 3) % U]));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(c[i]) : "f"(c[(i + 2) % U]), "f"(c[(i + 1) % U]));
      } else if (MODE == 16) {  // shl+add integer (ALU)
        asm volatile("shl.b32 %0, %0, 3;" : "+r"(h[i]));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i + 1) % U]));
      } else if (MODE == 17) {  // HFMA2
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(h[(i + 1) % U]), "r"(h[(i + 2) % U]));
      } else if (MODE == 18) {  // MUFU + F2FP + 3 FFMA2 (softmax-ish mix)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i & 1) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[i - 1])); }
        asm volatile("{.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tfma.rn.f32x2 x, x, y, y;\n\tmov.b64 {%0, %1}, x;}"
                     : "+f"(b[i]), "+f"(b[(i + 1) % U]) : "f"(b[(i + 2) % U]), "f"(b[(i + 3) % U]));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < U; ++i) s += a[i] + b[i] + c[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double instr_per_iter_per_u, float* out, long long* cyc) {
  for (int warps : {4, 8, 16, 32}) {
    bench<MODE><<<148, warps * 32>>>(out, cyc, 0.25f);
    cudaDeviceSynchronize();
    bench<MODE><<<148, warps * 32>>>(out, cyc, 0.25f);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[148];
    cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += c[i];
    avg /= 148;
    double winstr = double(ITER) * U * instr_per_iter_per_u * warps;  // warp-instructions per SM
    printf("%-28s warps/SM %2d  cycles %9.0f  warp-instr/clk/SMSP %.3f  (%s)\n", name, warps, avg, winstr / avg / 4.0,
           cudaGetErrorString(e));
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("MUFU.EX2", 1, out, cyc);
  run<1>("F2FP.PACK", 1, out, cyc);
  run<2>("FMNMX3", 1, out, cyc);
  run<3>("FMNMX", 1, out, cyc);
  run<4>("FFMA", 1, out, cyc);
  run<5>("FFMA2", 1, out, cyc);
  run<6>("FADD2", 1, out, cyc);
  run<7>("IMAD", 1, out, cyc);
  run<8>("LDS.128", 1, out, cyc);
  run<9>("LDS.64", 1, out, cyc);
  run<10>("MUFU+0.5 F2FP", 1.5, out, cyc);
  run<11>("MUFU+FFMA", 2, out, cyc);
  run<12>("F2FP+FFMA", 2, out, cyc);
  run<13>("F2FP+FMNMX", 2, out, cyc);
  run<14>("FMNMX+FFMA", 2, out, cyc);
  run<15>("FFMA2+FMNMX3", 2, out, cyc);
  run<16>("SHL+IADD", 2, out, cyc);
  run<17>("HFMA2", 1, out, cyc);
  run<18>("MUFU+.5F2FP+FFMA2", 2.5, out, cyc);
  return 0;
}
