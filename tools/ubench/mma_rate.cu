// Microbenchmark (GPU box): tcgen05.mma kind::f16 issue/execute rate with the UN-SWIZZLED K-major operand layouts conv_umma.cu
// uses.  One thread per CTA issues `n` MMAs (M=128, N=BN, K=16) back to back on operands resident in shared memory, then commits
// and waits; cycles/MMA is reported for several (BN, A SBO/LBO, B LBO, CTAs per SM) combinations.  Ideal: 64 cycles at N=128,
// 128 at N=256.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return ((uint64_t)((sbo >> 4) | (1u << 14)) << 32) | (((lbo >> 4) << 16) + ((addr & 0x3FFFFu) >> 4));
}
__global__ void __launch_bounds__(384) k(int n, int BN, int a_sbo, int a_lbo, int b_lbo, int ncols, int nacc, int a_span, int b_span, int mode, int pollers, int sleep_ns, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  __shared__ uint64_t bar, never;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (a_span + b_span) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc(128, BN);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + a_span;
    // mode 0: descriptors rebuilt per MMA with the kernel's address arithmetic; mode 1: four precomputed descriptor pairs,
    // 8 MMAs per loop iteration, nothing but tcgen05.mma in the loop
    uint64_t da[4], db[4];
    for (int j = 0; j < 4; j++) { da[j] = desc(a0 + (j & 1) * 2 * a_lbo, a_lbo, a_sbo); db[j] = desc(b0 + (j & 1) * 2 * b_lbo, b_lbo, 128); }
    long long t0 = clock64();
    if (mode == 0) {
      uint32_t ao = 0, bo = 0;
      for (int i = 0; i < n; i++) {
        umma_bf16(tm + (uint32_t)((i % nacc) * BN), desc(a0 + ao, a_lbo, a_sbo), desc(b0 + bo, b_lbo, 128), idesc, i >= nacc);
        ao += 2 * a_lbo; if (ao + 2 * a_lbo > (uint32_t)a_span) ao = 0;
        bo += 2 * b_lbo; if (bo + 2 * b_lbo > (uint32_t)b_span) bo = 0;
      }
    } else {
      for (int j = 0; j < nacc; j++) umma_bf16(tm + (uint32_t)(j * BN), da[0], db[0], idesc, 0);
      const uint32_t t1 = tm + (uint32_t)((nacc > 1 ? 1 : 0) * BN);
      for (int i = 0; i < n; i += 8) {
        umma_bf16(tm, da[0], db[0], idesc, 1); umma_bf16(t1, da[1], db[1], idesc, 1);
        umma_bf16(tm, da[2], db[2], idesc, 1); umma_bf16(t1, da[3], db[3], idesc, 1);
        umma_bf16(tm, da[0], db[1], idesc, 1); umma_bf16(t1, da[1], db[0], idesc, 1);
        umma_bf16(tm, da[2], db[3], idesc, 1); umma_bf16(t1, da[3], db[2], idesc, 1);
      }
    }
    umma_commit(&bar);
    while (!mbar_try_wait(&bar, 0)) {}
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&never)) : "memory");
  } else if (warp >= 2 && warp < 2 + pollers) {
    // like the epilogue warps of conv_umma.cu during the main loop: all 32 lanes poll a barrier in shared memory
    while (!mbar_try_wait(&never, 0)) { if (sleep_ns) __nanosleep(sleep_ns); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(ncols) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n = 4096;
  struct Cfg { const char* name; int BN, a_sbo, a_lbo, b_lbo, ncols, nacc, a_span, b_span, ctas; };
  Cfg cfgs[] = {
      {"1x1 flat MT1 BN128 (A lbo 2048)        ", 128, 128, 2048, 2048, 128, 1, 8192, 8192, 148},
      {"1x1 flat MT1 BN128, 2 CTAs/SM          ", 128, 128, 2048, 2048, 128, 1, 8192, 8192, 296},
      {"1x1 flat MT1 BN128, 2 accumulators     ", 128, 128, 2048, 2048, 256, 2, 8192, 8192, 148},
      {"halo 3x3 MT1 BN128 (A sbo 160 lbo 2880)", 128, 160, 2880, 2048, 128, 1, 2 * 2880, 8192, 148},
      {"halo 3x3 MT1 BN128, 2 CTAs/SM          ", 128, 160, 2880, 2048, 128, 1, 2 * 2880, 8192, 296},
      {"1x1 flat MT1 BN256 (B lbo 4096)        ", 256, 128, 2048, 4096, 256, 1, 8192, 16384, 148},
      {"1x1 flat MT2 BN256 (A lbo 4096)        ", 256, 128, 4096, 4096, 512, 2, 16384, 16384, 148},
      {"A lbo 128 sbo 256 (K-adjacent) BN128   ", 128, 256, 128, 2048, 128, 1, 8192, 8192, 148},
      // narrow N (round 2: is a block-diagonal "depthwise as MMA" affordable? cycles per 128 x N x 16 MMA at N = 16 / 32 / 64)
      {"halo A, BN16 (B lbo 256)               ", 16, 160, 2880, 256, 32, 1, 4 * 2880, 1024, 148},
      {"halo A, BN32 (B lbo 512)               ", 32, 160, 2880, 512, 32, 1, 4 * 2880, 2048, 148},
      {"halo A, BN64 (B lbo 1024)              ", 64, 160, 2880, 1024, 64, 1, 4 * 2880, 4096, 148},
      {"halo A, BN16, 2 accumulators           ", 16, 160, 2880, 256, 32, 2, 4 * 2880, 1024, 148},
  };
  for (auto& c : cfgs) {
    // B descriptor SBO is fixed at 128 in the kernel; for the K-adjacent B variant use sbo 256 via a_sbo trick is not possible -> note
    for (int pollers : {0, 4, 8})
      for (int sleep_ns : {0, 200}) {
        if (pollers == 0 && sleep_ns) continue;
        const int mode = 1;
        for (int rep = 0; rep < 2; rep++) k<<<c.ctas, 384, c.a_span + c.b_span + 256>>>(n, c.BN, c.a_sbo, c.a_lbo, c.b_lbo, c.ncols, c.nacc, c.a_span, c.b_span, mode, pollers, sleep_ns, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        printf("%s pollers=%d sleep=%3dns : %6.1f cycles/MMA (ideal %d)  %s\n", c.name, pollers, sleep_ns, (double)cyc / n, c.BN / 2, cudaGetErrorString(e));
      }
  }
  return 0;
}
