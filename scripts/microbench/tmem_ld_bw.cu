// Microbenchmark: tcgen05.ld (TMEM -> registers) bandwidth per SM on B200, for 1..4 reading warps and 1..2 CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(128) k(int iters, int nwarps, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      uint32_t r[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(base + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 128; i += 16) acc ^= r[i];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(256) : "memory");
}

int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 8 * 1024); cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int nw = 1; nw <= 4; ++nw) {
      k<<<148 * ctas, 128>>>(10, nw, d, s);  // warm
      k<<<148 * ctas, 128>>>(iters, nw, d, s);
      long long h[8];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      cudaError_t e = cudaDeviceSynchronize();
      const double bytes = (double)iters * nw * 32 * 128 * 4;  // per CTA
      printf("ctas/SM=%d warps=%d: %lld cycles/CTA, %.1f B/clk per CTA, %.1f B/clk per SM (%s)\n", ctas, nw, h[0], bytes / h[0],
             ctas * bytes / h[0], cudaGetErrorString(e));
    }
  return 0;
}
