// Cost-volume correlation, 81 displacements, NCHW fp32 (the C-ABI drop-in for corr_cuda_forward in PWC-Net's configuration:
// pad 4, kernel 1, max displacement 4, strides 1; correlation_package/src/corr_cuda.c:7-82, corr_cuda_kernel.cu:59-127) with
// TMA-staged tiles.
//
// The NCHW layout already is what the inner loop wants -- x contiguous -- so the operands need no transposition at all: per 8
// channels one TMA box {40 x, 16 y, 8 c} of the second feature map (the 8 x 32 pixel tile + its 4-pixel halo; rows / columns
// outside the image arrive as zeros = the reference's zero padding) and one box {32, 8, 8} of the first land in shared memory
// as [c][y][x], through a 3-stage mbarrier ring.  A thread owns 4 consecutive pixels x 3 vertical displacements x all 9
// horizontal ones = 108 accumulators: per channel it reads its 4 f1 values (one LDS.128) and, per displacement row, the 12 f2
// values its pixels can reach (three LDS.128, conflict-free: a quarter warp reads 128 contiguous bytes) and issues 108 FMAs --
// 10.8 FMAs per shared-memory load where the per-pixel kernel (corr81_kernel) has 0.96.  No tensor cores: the cost volume is 81
// short dot products per pixel over a sliding window, not a dense contraction.  Results are divided by C (corr_cuda_kernel.cu:
// 119-121) and stored as float4 (4 consecutive x of one displacement plane).
#include <cuda.h>

#include "common.cuh"

namespace premvos {

namespace {

constexpr int TH = 8, TW = 32, MD = 4;
constexpr int HH = TH + 2 * MD, HW = TW + 2 * MD;   // 16 x 40
constexpr int CK = 8, STAGES = 3;
constexpr int S2_FLOATS = CK * HH * HW, S1_FLOATS = CK * TH * TW;
constexpr int STAGE_BYTES = (S2_FLOATS + S1_FLOATS) * 4;   // 28 672
constexpr int THREADS = 192;                               // 64 groups of 4 pixels x 3 groups of 3 vertical displacements
constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();   // a protocol bug must become a launch error, never a hung GPU
  }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct CorrTmaArgs {
  float* out;   // [B][81][H][W]
  int B, C, H, W, tiles_x, tiles_y;
};

__global__ void __launch_bounds__(THREADS, 2) corr81_tma_kernel(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
                                                                const CorrTmaArgs a) {
  extern __shared__ __align__(128) float corr_sm[];   // [STAGES][s2 [CK][HH][HW] | s1 [CK][TH][TW]], then the barriers
  uint64_t* full = reinterpret_cast<uint64_t*>(corr_sm + (size_t)STAGES * (STAGE_BYTES / 4));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int t = blockIdx.x;
  const int tx = t % a.tiles_x; t /= a.tiles_x;
  const int ty = t % a.tiles_y;
  const int n = t / a.tiles_y;
  const int x0 = tx * TW, y0 = ty * TH;
  const int nchunks = (a.C + CK - 1) / CK;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int chunk, int s) {   // one thread; channels beyond C and pixels outside the image arrive as zeros
    float* dst = corr_sm + (size_t)s * (STAGE_BYTES / 4);
    mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
    tma_load_4d(&tm2, &full[s], dst, x0 - MD, y0 - MD, chunk * CK, n);
    tma_load_4d(&tm1, &full[s], dst + S2_FLOATS, x0, y0, chunk * CK, n);
  };
  if (tid == 0)
    for (int s = 0; s < STAGES && s < nchunks; s++) issue(s, s);

  const int grp = warp >> 1;                               // vertical displacements 3*grp .. 3*grp + 2
  const int py = (warp & 1) * 4 + (lane >> 3), gx = lane & 7;   // pixel row of the tile, group of 4 pixels
  float acc[3][9][4];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int d = 0; d < 9; d++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[r][d][j] = 0.f;

  for (int chunk = 0; chunk < nchunks; chunk++) {
    const int s = chunk % STAGES;
    mbar_wait(&full[s], (uint32_t)(chunk / STAGES) & 1u);
    const float* s2 = corr_sm + (size_t)s * (STAGE_BYTES / 4) + ((py + 3 * grp) * HW + gx * 4);
    const float* s1 = corr_sm + (size_t)s * (STAGE_BYTES / 4) + S2_FLOATS + (py * TW + gx * 4);
#pragma unroll 2
    for (int c = 0; c < CK; c++) {
      const float4 f1 = *reinterpret_cast<const float4*>(s1 + c * TH * TW);
      const float f1v[4] = {f1.x, f1.y, f1.z, f1.w};
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const float4* row = reinterpret_cast<const float4*>(s2 + c * HH * HW + r * HW);
        const float4 q0 = row[0], q1 = row[1], q2 = row[2];
        const float f2v[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
        for (int d = 0; d < 9; d++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[r][d][j] = fmaf(f1v[j], f2v[j + d], acc[r][d][j]);
      }
    }
    __syncthreads();   // every thread is done with stage s
    if (tid == 0 && chunk + STAGES < nchunks) issue(chunk + STAGES, s);
  }

  const int gy = y0 + py, gxx = x0 + gx * 4;
  if (gy < a.H && gxx < a.W) {   // W % 4 == 0: a group of 4 pixels is inside or outside as a whole
    const float cdiv = (float)a.C;
    const long plane = (long)a.H * a.W;
    float* o = a.out + ((long)n * 81 + grp * 27) * plane + (long)gy * a.W + gxx;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int d = 0; d < 9; d++) {
        *reinterpret_cast<float4*>(o + (long)(r * 9 + d) * plane) =
            make_float4(acc[r][d][0] / cdiv, acc[r][d][1] / cdiv, acc[r][d][2] / cdiv, acc[r][d][3] / cdiv);
      }
  }
}

}  // namespace

// -> 0 launched, 1 not applicable (caller falls back to corr81_kernel), < 0 / > 1 error
int corr81_nchw_tma(const float* f1, const float* f2, float* out, int B, int C, int H, int W, cudaStream_t st) {
  if ((W & 3) != 0 || (((uintptr_t)f1 | (uintptr_t)f2 | (uintptr_t)out) & 15) != 0) return 1;   // TMA: 16-byte rows
  alignas(64) unsigned char m1[128], m2[128];
  const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)W * H * 4, (uint64_t)W * H * C * 4};
  const uint32_t box1[4] = {TW, TH, CK, 1}, box2[4] = {HW, HH, CK, 1};
  PV_TRY(encode_tensor_map_f32(m1, f1, 4, dims, strides, box1));
  PV_TRY(encode_tensor_map_f32(m2, f2, 4, dims, strides, box2));
  static bool attr_set = false;
  if (!attr_set) {
    PV_CUDA(cudaFuncSetAttribute(corr81_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_set = true;
  }
  CorrTmaArgs a{out, B, C, H, W, (W + TW - 1) / TW, (H + TH - 1) / TH};
  const long ctas = (long)a.tiles_x * a.tiles_y * B;
  PV_CHECK(ctas < (1L << 31), PREMVOS_ERR_INVALID_ARG, "premvos_corr_forward: too many tiles");
  const double px = (double)B * H * W;
  prof_before(st);
  corr81_tma_kernel<<<(unsigned)ctas, THREADS, SMEM, st>>>(*reinterpret_cast<const CUtensorMap*>(m1), *reinterpret_cast<const CUtensorMap*>(m2), a);
  return after_launch("corr81_tma_kernel", st, 2.0 * 81 * C * px, 4.0 * (2.0 * C + 81) * px);
}

}  // namespace premvos
