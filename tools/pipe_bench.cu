// Development microbenchmark: issue rates of the epilogue instruction kinds (cycles per warp instruction per SMSP)
// for 1, 2 and 4 warps per scheduler.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench.bin tools/pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 512, CH = 8;   // CH independent chains per thread

template <int KIND>
__global__ void k(float* out, unsigned long long* cyc, float seed) {
  float a[CH], b[CH];
  uint64_t p[CH];
  uint32_t u[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    a[i] = seed + threadIdx.x * 1e-3f + i;
    b[i] = seed * 0.5f + i;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
    u[i] = threadIdx.x + i;
  }
  uint64_t c2, c3;
  asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(0.999f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c3) : "f"(0.001f));
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (KIND == 0) a[i] = fmaf(a[i], 0.999f, 0.001f);                                   // FFMA (imm)
      if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(c3));   // FFMA2
      if (KIND == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                       // MUFU
      if (KIND == 3) u[i] = u[i] * 0x800000u + u[(i + 1) % CH];                             // IMAD
      if (KIND == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c3));               // FADD2
      if (KIND == 5) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(c3));
                       asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }                   // FFMA2 + MUFU
      if (KIND == 6) a[i] = fmaf(a[i], b[i], b[(i + 1) % CH]);                              // FFMA (3 regs)
      if (KIND == 7) { a[i] = fmaf(a[i], b[i], b[(i + 1) % CH]); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(c3)); }  // FFMA + FFMA2
      if (KIND == 8) asm volatile("shf.l.wrap.b32 %0, %0, %0, 23;" : "+r"(u[i]));           // SHF (ALU)
      if (KIND == 10) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(c3));
                        asm volatile("shf.l.wrap.b32 %0, %0, %0, 23;" : "+r"(u[i])); }         // FFMA2 + SHF (FMA + ALU pipes)
      if (KIND == 11) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(c3));
                        asm volatile("shf.l.wrap.b32 %0, %0, %0, 23;" : "+r"(u[i]));
                        if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }   // + MUFU every 4th
      if (KIND == 12) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c3));
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]), "r"(u[(i + 2) % CH])); }  // FADD2 + LOP3
      if (KIND == 9) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(c3)); u[i] = u[i] * 0x800000u + u[(i + 1) % CH]; }  // FFMA2 + IMAD
    }
  }
  const unsigned long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
    s += a[i] + lo + hi + __uint_as_float(u[i]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int per_iter) {
  float* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  printf("%-14s", name);
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = warps_per_smsp * 4 * 32;
    k<KIND><<<148, threads>>>(out, cyc, 1.0f);
    k<KIND><<<148, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (auto x : h) s += x;
    s /= 148;
    // cycles per warp instruction per SMSP
    printf("  %dw/smsp: %6.2f cyc/inst", warps_per_smsp, s / (double(ITERS) * CH * per_iter * warps_per_smsp));
  }
  printf("\n");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA imm", 1);
  run<6>("FFMA 3reg", 1);
  run<1>("FFMA2", 1);
  run<4>("FADD2", 1);
  run<2>("MUFU.EX2", 1);
  run<3>("IMAD", 1);
  run<8>("SHF", 1);
  run<5>("FFMA2+MUFU", 2);
  run<7>("FFMA+FFMA2", 2);
  run<9>("FFMA2+IMAD", 2);
  run<10>("FFMA2+SHF", 2);
  run<11>("FFMA2+SHF+MUFU/4", 2);
  run<12>("FADD2+LOP3", 2);
  return 0;
}
