// Microbenchmark (GPU box): cost of the producer <-> issuer ring handshake used by conv_umma.cu, without any data or MMA.
//   mode 0: consumer releases a stage with mbarrier.arrive;  mode 1: with tcgen05.commit (as the kernel does)
//   pollers: extra warps spinning on a barrier that never completes until the end (like the epilogue warps during the main loop)
//   lanes: 32 = every lane of the waiting warp executes try_wait; 1 = one elected lane waits, the rest sit at __syncwarp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(384) k(int iters, int stages, int mode, int pollers, int lanes, int fence, int alloc, long long* out) {
  __shared__ uint64_t full[8], empty[8], never;
  __shared__ uint32_t slot;
  if (alloc && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < 8; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } mbar_init(&never, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < iters; i++) {
      if (lanes == 32 || lane == 0) mbar_wait(&empty[st], ph ^ 1u);
      __syncwarp();
      if (elect_one()) mbar_arrive(&full[st]);
      __syncwarp();
      if (++st == (uint32_t)stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < iters; i++) {
      if (lanes == 32 || lane == 0) mbar_wait(&full[st], ph);
      if (fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __syncwarp();
      if (elect_one()) { if (mode == 1) umma_commit(&empty[st]); else mbar_arrive(&empty[st]); }
      __syncwarp();
      if (++st == (uint32_t)stages) { st = 0; ph ^= 1u; }
    }
    if (elect_one()) mbar_arrive(&never);
    __syncwarp();
    if (lane == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
  } else if (warp >= 4 && warp < 4 + pollers) {
    if (lanes == 32 || lane == 0) mbar_wait(&never, 0);
    __syncwarp();
  }
  __syncthreads();
  if (alloc && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 4096;
  for (int mode = 0; mode < 2; mode++)
    for (int stages : {3})
      for (int pollers : {8})
        for (int fence : {0, 1})
          for (int alloc : {0, 1}) {
            const int lanes = 32;
            for (int rep = 0; rep < 2; rep++) k<<<148, 384>>>(iters, stages, mode, pollers, lanes, fence, alloc, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("mode=%s stages=%d pollers=%d fence=%d tmem_alloc=%d : %7.1f cycles/iteration  (%s)\n", mode ? "commit" : "arrive", stages, pollers, fence, alloc, (double)c / iters, cudaGetErrorString(e));
          }
  return 0;
}
