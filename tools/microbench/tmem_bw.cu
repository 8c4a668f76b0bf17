// Microbenchmark: TMEM -> register read bandwidth of tcgen05.ld (sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

template <int INFLIGHT>
__global__ void tmem_read_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t a[INFLIGHT][32];
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < INFLIGHT; ++u) ld32(base + ((i * INFLIGHT + u) & 7) * 32, a[u]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int u = 0; u < INFLIGHT; ++u) acc ^= a[u][0] ^ a[u][31];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

int main() {
  long long* d_cycles; uint32_t* d_sink;
  cudaMalloc(&d_cycles, 148 * 8); cudaMalloc(&d_sink, 4);
  const int iters = 2000;
  for (int warps : {4, 8}) {
    for (int inflight : {1, 2, 4}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (inflight == 1) tmem_read_kernel<1><<<148, warps * 32>>>(iters, d_cycles, d_sink);
        if (inflight == 2) tmem_read_kernel<2><<<148, warps * 32>>>(iters, d_cycles, d_sink);
        if (inflight == 4) tmem_read_kernel<4><<<148, warps * 32>>>(iters, d_cycles, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148]; cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = (double)iters * inflight * warps * 32 * 32 * 4;
      printf("warps=%d inflight=%d: %lld cycles, %.1f B/clk/SM (%.0f cycles per 128x128 fp32 accumulator)\n", warps, inflight, h[0], bytes / h[0],
             65536.0 / (bytes / h[0]));
    }
  }
  return 0;
}
