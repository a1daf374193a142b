// micro-benchmark: MUFU.TANH throughput, f32 vs f16 / f16x2 / bf16x2 (B200)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, int iters) {
  unsigned a = threadIdx.x * 3 + 1, b = a + 7, c = a + 11, d = a + 13;
  float fa = __uint_as_float(0x3f000000u | (a & 0xffff)), fb = fa * 0.5f, fc = fa * 0.25f, fd = fa * 0.125f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(fa)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(fb));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(fc)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(fd));
      } else if (MODE == 1) {
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(b));
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(c)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(d));
      } else if (MODE == 2) {
        asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(a)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(b));
        asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(c)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(d));
      } else if (MODE == 3) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fa)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fb));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fc)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fd));
      } else if (MODE == 4) {
        unsigned short ha = a, hb = b, hc = c, hd = d;
        asm volatile("tanh.approx.f16 %0, %0;" : "+h"(ha)); asm volatile("tanh.approx.f16 %0, %0;" : "+h"(hb));
        asm volatile("tanh.approx.f16 %0, %0;" : "+h"(hc)); asm volatile("tanh.approx.f16 %0, %0;" : "+h"(hd));
        a = ha; b = hb; c = hc; d = hd;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __float_as_uint(fa + fb + fc + fd);
}
template <int MODE> void run(const char* name, int values_per_op) {
  unsigned* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096, blocks = 148 * 2, threads = 1024;
  k<MODE><<<blocks, threads>>>(out, 16);
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * 32.0;
  printf("%-22s %.3f ms  %.1f G instr-lanes/s  = %.2f lanes/clk/SM @1.9GHz, %.1f G values/s\n", name, ms, ops / ms / 1e6,
         ops / ms / 1e6 / 148 / 1.9, ops * values_per_op / ms / 1e6);
  cudaFree(out);
}
int main() {
  run<0>("tanh.approx.f32", 1); run<1>("tanh.approx.f16x2", 2); run<2>("tanh.approx.bf16x2", 2);
  run<3>("ex2.approx.ftz.f32", 1); run<4>("tanh.approx.f16", 1);
  return 0;
}
