// micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM on B200, as a function of the
// number of reading warps and of the loads in flight per warp.  The BiLSTM epilogue reads
// 128 lanes x 80 columns x 4 B = 40 KB per N-chunk and SM; this tells what that costs at best.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bench ldtm_bench.cu && ./ldtm_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// DEPTH loads of 16 columns in flight per warp before each wait
template <int DEPTH>
__global__ void __launch_bounds__(1024, 1) k(unsigned* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[DEPTH][16];
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) ld16(base + (((warp >> 2) * 16 * DEPTH + d * 16 + i * 16) & 0x1F0), v[d]);
    wait_ld();
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int e = 0; e < 16; ++e) acc ^= v[d][e];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
  }
}

template <int DEPTH> void run(int warps) {
  unsigned* out; long long* cyc;
  const int blocks = 148, iters = 20000;
  cudaMalloc(&out, blocks * 1024 * 4); cudaMalloc(&cyc, blocks * 8);
  k<DEPTH><<<blocks, warps * 32>>>(out, cyc, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<DEPTH><<<blocks, warps * 32>>>(out, cyc, iters); cudaEventRecord(e1);
  cudaError_t err = cudaEventSynchronize(e1);
  if (err != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(err)); return; }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)warps * 32 * 16 * 4 * DEPTH * iters;       // per SM
  printf("warps %2d depth %d: %.3f ms, %lld clk on SM 0 -> %.1f B/clk/SM (%.1f per sub-partition), %.2f TB/s chip\n", warps, DEPTH, ms,
         h[0], bytes / (double)h[0], bytes / (double)h[0] / 4.0, bytes * blocks / ms / 1e9);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 12, 16, 20, 32}) { run<1>(w); run<2>(w); }
  return 0;
}
