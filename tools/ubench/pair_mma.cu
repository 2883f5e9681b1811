// Microbenchmark + layout check for the NEXT step of conv_umma: tcgen05.mma.cta_group::2 on the UN-SWIZZLED K-major
// operand layouts conv_umma.cu uses.  NOT RUN YET (written at the end of round 1 without GPU minutes left; compiles for sm_100a).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pair_mma tools/ubench/pair_mma.cu && tools/ubench/pair_mma
//
// A CTA pair (cluster of 2) computes D[256 x 256] = A[256 x K] * B[256 x K]^T with K = 16 * ksteps:
//   * each CTA holds ITS 128 rows of A and ITS 128 of the 256 rows of B (= output columns) in its own shared memory, both as
//     [k chunk of 8 elements][row][8 bf16]: core matrix = 8 rows x 16 B, SBO = 128 B between 8-row groups, LBO = rows * 16 B
//     between the two K halves of one MMA -- the layout of conv_umma's A boxes / packed weights.  This is the point of the pair:
//     an SM ingests 128 x K of weights instead of 256 x K (41.7 instead of 62.5 B/clk of L2 -> shared-memory fill for a 1x1 layer);
//   * the leader CTA (rank 0) issues tcgen05.mma.cta_group::2 with M = 256, N = 256; each CTA's TMEM receives its 128 rows x 256
//     columns; tcgen05.commit ... multicast::cluster arrives on the barrier of both CTAs;
//   * both CTAs read their accumulator back and compare it with the integer dot products (exact in bf16 / fp32).
// Prints mismatches per CTA and the cycles per MMA of a back-to-back issue loop (ideal: 128 cycles per 256 x 256 x 16).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();   // a protocol bug must become a launch error, never a hung GPU
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// both CTAs' barriers (same shared-memory offset) get one arrival when the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_pair_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, K-major, N >> 3 at bits 17-22, M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// un-swizzled K-major shared-memory descriptor (version 1 at bits 46-47)
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return ((uint64_t)((sbo >> 4) | (1u << 14)) << 32) | (((lbo >> 4) << 16) + ((addr & 0x3FFFFu) >> 4));
}

constexpr int ROWS = 128;           // rows of A and of B held by one CTA
constexpr int KSTEPS = 4;           // K = 64
constexpr int LBO = ROWS * 16;      // bytes between the two 8-element K halves
constexpr int OP_BYTES = KSTEPS * 2 * LBO;

__device__ __forceinline__ int a_val(int m, int k) { return ((m * 3 + k * 5) % 7) - 3; }   // small integers: exact arithmetic
__device__ __forceinline__ int b_val(int n, int k) { return ((n * 2 + k * 3) % 5) - 2; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(int timing_mmas, long long* out_cycles, int* out_bad) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem);              // [2 * KSTEPS chunks][ROWS][8]
  __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(smem + OP_BYTES);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's rows: A rows [rank*128, +128) of the 256 x K matrix, B rows (= output columns) [rank*128, +128)
  for (int i = threadIdx.x; i < 2 * KSTEPS * ROWS * 8; i += blockDim.x) {
    const int e = i & 7, row = (i >> 3) % ROWS, chunk = i / (8 * ROWS), k = chunk * 8 + e;
    As[i] = __float2bfloat16((float)a_val((int)rank * ROWS + row, k));
    Bs[i] = __float2bfloat16((float)b_val((int)rank * ROWS + row, k));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {   // both CTAs of the pair take part in the paired allocation (512 columns would also hold a second buffer)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy operand writes -> async proxy (MMA) reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                                                  // the peer's operands and barrier are ready too
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  const uint32_t idesc = make_idesc(256, 256);
  long long cycles = 0;
  if (rank == 0 && threadIdx.x == 32) {   // the leader CTA issues for the pair
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    for (int ks = 0; ks < KSTEPS; ks++)
      umma_pair_bf16(tm, desc(a0 + ks * 2 * LBO, LBO, 128), desc(b0 + ks * 2 * LBO, LBO, 128), idesc, ks > 0);
    umma_commit_pair(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- check: thread = accumulator row m of this CTA (global row rank*128 + m), all 256 columns ----
  int bad = 0;
  {
    const int m = warp * 32 + lane, gm = (int)rank * ROWS + m;
    for (int c0 = 0; c0 < 256; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      for (int j = 0; j < 16; j++) {
        int want = 0;
        for (int k = 0; k < 16 * KSTEPS; k++) want += a_val(gm, k) * b_val(c0 + j, k);
        if (__uint_as_float(v[j]) != (float)want) bad++;
      }
    }
  }
  atomicAdd(out_bad + rank, bad);
  // ---- timing: back-to-back pair MMAs on the same operands ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (rank == 0 && threadIdx.x == 32) {
    const uint64_t da = desc(smem_u32(As), LBO, 128), db = desc(smem_u32(Bs), LBO, 128);
    const long long t0 = clock64();
    for (int i = 0; i < timing_mmas; i++) umma_pair_bf16(tm, da, db, idesc, 1);
    umma_commit_pair(&bar);
    mbar_wait(&bar, 1);
    cycles = clock64() - t0;
    *out_cycles = cycles;
  } else {
    mbar_wait(&bar, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();   // nobody frees TMEM / exits while the peer may still be using the pair
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

int main() {
  long long* d_cycles;
  int* d_bad;
  cudaMalloc(&d_cycles, sizeof(long long));
  cudaMalloc(&d_bad, 2 * sizeof(int));
  cudaMemset(d_cycles, 0, sizeof(long long));
  cudaMemset(d_bad, 0, 2 * sizeof(int));
  const int smem = 2 * OP_BYTES + 256, n = 4096;
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pair_kernel<<<2, 128, smem>>>(n, d_cycles, d_bad);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("pair_mma: %s\n", cudaGetErrorString(e)); return 1; }
  long long cycles;
  int bad[2];
  cudaMemcpy(&cycles, d_cycles, sizeof(cycles), cudaMemcpyDeviceToHost);
  cudaMemcpy(bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
  printf("pair_mma: cta_group::2 M=256 N=256 K=%d un-swizzled K-major: mismatches CTA0 %d CTA1 %d of %d each; %.1f cycles per MMA (%d back to back)\n",
         16 * KSTEPS, bad[0], bad[1], 128 * 256, (double)cycles / n, n);
  return (bad[0] || bad[1]) ? 2 : 0;
}
