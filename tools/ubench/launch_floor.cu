// Microbenchmark (GPU box): fixed cost of launching a persistent conv_umma-like kernel: CTA launch with a large dynamic
// shared-memory carve-out, TMEM alloc / dealloc, mbarrier init -- with and without a small-smem kernel in between
// (carve-out reconfiguration).  20 launches back to back, CUDA events, microseconds per launch.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(384) big(int do_alloc, int ncols, float* out) {
  extern __shared__ uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bars[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 24; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (do_alloc && threadIdx.x >= 64 && threadIdx.x < 96) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0 && out) out[blockIdx.x] = (float)smem[blockIdx.x];
  __syncthreads();
  if (do_alloc && threadIdx.x >= 64 && threadIdx.x < 96) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(ncols) : "memory");
}
__global__ void small(float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = out[i] * 1.0001f + 1.f;
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24); cudaMemset(d, 0, 1 << 24);
  cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { const char* name; int grid, threads, smem, alloc, ncols, interleave; };
  Cfg cfgs[] = {{"148 CTAs x 384 thr,  16 KB smem, no TMEM        ", 148, 384, 16 << 10, 0, 0, 0},
                {"148 CTAs x 384 thr, 197 KB smem, no TMEM        ", 148, 384, 197 << 10, 0, 0, 0},
                {"148 CTAs x 384 thr, 197 KB smem, TMEM 512       ", 148, 384, 197 << 10, 1, 512, 0},
                {"296 CTAs x 256 thr, 108 KB smem, TMEM 256       ", 296, 256, 108 << 10, 1, 256, 0},
                {"148 x 384, 197 KB, TMEM 512, small kernel between", 148, 384, 197 << 10, 1, 512, 1},
                {"296 x 256, 108 KB, TMEM 256, small kernel between", 296, 256, 108 << 10, 1, 256, 1},
                {"small kernel alone (4096 blocks x 256)           ", 0, 0, 0, 0, 0, 2}};
  for (auto& c : cfgs) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      for (int i = 0; i < 20; i++) {
        if (c.interleave != 2) big<<<c.grid, c.threads, c.smem>>>(c.alloc, c.ncols, d);
        if (c.interleave) small<<<4096, 256>>>(d + 4096, 1 << 20);
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("%s : %6.2f us per iteration  (%s)\n", c.name, ms * 1e3 / 20, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
