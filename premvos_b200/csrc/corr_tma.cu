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

#include <stdlib.h>

#include "cp8.cuh"

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

// ---- the same on CP8 features (inside PWC-Net) -------------------------------------------------------------------------------
// Operands are split-bf16 chunk planes [N][C/8][H][W][8]: per 8-channel chunk the hi and lo planes of the f2 halo tile and of the
// f1 tile arrive by TMA (5-D boxes {8, 40, 16} / {8, 32, 8}; zero fill = padding) through a 2-stage ring; all threads then turn
// the chunk into fp32 channel planes [c][y][x] in a third buffer (hi + lo, 8 conflict-free STS per pixel: 5 % of the chunk's
// instructions) and the inner loop is the one of corr81_tma_kernel.  The 81 results per pixel go through shared memory once more
// to leave as 11 chunk planes of the decoder slab (LeakyReLU fused, split to hi / lo); f1's planes are copied next to them.
constexpr int CP_STAGES = 2;
constexpr int CP_F2_PLANE = HH * HW * 8 * 2, CP_F1_PLANE = TH * TW * 8 * 2;          // bytes of one bf16 plane tile
constexpr int CP_STAGE_BYTES = 2 * CP_F2_PLANE + 2 * CP_F1_PLANE;                      // 28 672
constexpr int CP_CONV_BYTES = (8 * HH * HW + 8 * TH * TW) * 4;                         // 28 672
constexpr int OUT_PITCH = 84;                                                          // floats per pixel in the output staging
constexpr size_t CP_SMEM = (size_t)CP_STAGES * CP_STAGE_BYTES + CP_CONV_BYTES + 64;
static_assert((size_t)TH * TW * OUT_PITCH * 4 <= (size_t)CP_STAGES * CP_STAGE_BYTES + CP_CONV_BYTES, "output staging must fit the operand buffers");

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

struct CorrCp8TmaArgs {
  cp8::CV f1, f2, out, c1;
  int has_c1, tiles_x, tiles_y;
  float slope;
};

__global__ void __launch_bounds__(THREADS, 2)
corr81_cp8_tma_kernel(const __grid_constant__ CUtensorMap tm1h, const __grid_constant__ CUtensorMap tm1l, const __grid_constant__ CUtensorMap tm2h,
                      const __grid_constant__ CUtensorMap tm2l, const CorrCp8TmaArgs a) {
  using namespace cp8;
  extern __shared__ __align__(128) unsigned char cp_sm[];
  unsigned char* conv_b = cp_sm + (size_t)CP_STAGES * CP_STAGE_BYTES;
  float* c2 = reinterpret_cast<float*>(conv_b);                      // [8][HH][HW]
  float* c1 = c2 + 8 * HH * HW;                                      // [8][TH][TW]
  uint64_t* full = reinterpret_cast<uint64_t*>(conv_b + CP_CONV_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int t = blockIdx.x;
  const int tx = t % a.tiles_x; t /= a.tiles_x;
  const int ty = t % a.tiles_y;
  const int n = t / a.tiles_y;
  const int x0 = tx * TW, y0 = ty * TH;
  const int H = a.f1.H, W = a.f1.W, C = a.f1.C;
  const int nchunks = (C + 7) / 8;
  if (tid == 0) {
    for (int s = 0; s < CP_STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int chunk, int s) {   // one thread
    unsigned char* dst = cp_sm + (size_t)s * CP_STAGE_BYTES;
    mbar_arrive_expect_tx(&full[s], CP_STAGE_BYTES);
    tma_load_5d(&tm2h, &full[s], dst, 0, x0 - MD, y0 - MD, a.f2.c0 + chunk, n);
    tma_load_5d(&tm2l, &full[s], dst + CP_F2_PLANE, 0, x0 - MD, y0 - MD, a.f2.c0 + chunk, n);
    tma_load_5d(&tm1h, &full[s], dst + 2 * CP_F2_PLANE, 0, x0, y0, a.f1.c0 + chunk, n);
    tma_load_5d(&tm1l, &full[s], dst + 2 * CP_F2_PLANE + CP_F1_PLANE, 0, x0, y0, a.f1.c0 + chunk, n);
  };
  if (tid == 0)
    for (int s = 0; s < CP_STAGES && s < nchunks; s++) issue(s, s);

  const int grp = warp >> 1;
  const int py = (warp & 1) * 4 + (lane >> 3), gx = lane & 7;
  float acc[3][9][4];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int d = 0; d < 9; d++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[r][d][j] = 0.f;

  for (int chunk = 0; chunk < nchunks; chunk++) {
    const int s = chunk % CP_STAGES;
    mbar_wait(&full[s], (uint32_t)(chunk / CP_STAGES) & 1u);
    const unsigned char* st = cp_sm + (size_t)s * CP_STAGE_BYTES;
    // ---- hi + lo -> fp32 channel planes ----
    for (int u = tid; u < HH * HW + TH * TW; u += THREADS) {
      const bool is2 = u < HH * HW;
      const int p = is2 ? u : u - HH * HW;
      const uint4 h = *reinterpret_cast<const uint4*>(st + (is2 ? 0 : 2 * CP_F2_PLANE) + (size_t)p * 16);
      const uint4 l = *reinterpret_cast<const uint4*>(st + (is2 ? CP_F2_PLANE : 2 * CP_F2_PLANE + CP_F1_PLANE) + (size_t)p * 16);
      const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
      float* dst = is2 ? c2 + p : c1 + p;
      const int plane = is2 ? HH * HW : TH * TW;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        dst[(2 * j) * plane] = __uint_as_float(hh[j] << 16) + __uint_as_float(ll[j] << 16);
        dst[(2 * j + 1) * plane] = __uint_as_float(hh[j] & 0xffff0000u) + __uint_as_float(ll[j] & 0xffff0000u);
      }
      if (!is2 && a.has_c1) {   // fused copy of f1 (raw hi / lo bits) into the decoder slab
        const int gy = y0 + p / TW, gxx = x0 + p % TW;
        if (gy < H && gxx < W) {
          const long o = cv_elem(a.c1, n, chunk, gy, gxx);
          *reinterpret_cast<uint4*>(a.c1.hi + o) = h;
          *reinterpret_cast<uint4*>(a.c1.lo + o) = l;
        }
      }
    }
    __syncthreads();   // planes complete; stage s may be refilled
    if (tid == 0 && chunk + CP_STAGES < nchunks) issue(chunk + CP_STAGES, s);
    const float* s2 = c2 + ((py + 3 * grp) * HW + gx * 4);
    const float* s1 = c1 + (py * TW + gx * 4);
#pragma unroll 2
    for (int c = 0; c < 8; c++) {
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
    __syncthreads();   // the planes may be overwritten by the next chunk
  }

  // ---- 81 results per pixel -> shared memory [pixel][84] -> 11 chunk planes ----
  float* so = reinterpret_cast<float*>(cp_sm);
  const float cdiv = (float)C;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float* o = so + (py * TW + gx * 4 + j) * OUT_PITCH + grp * 27;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int d = 0; d < 9; d++) {
        float v = acc[r][d][j] / cdiv;   // corr_cuda_kernel.cu:119-121
        o[r * 9 + d] = v > 0.f ? v : v * a.slope;
      }
    if (grp == 2) { o[27] = 0.f; o[28] = 0.f; o[29] = 0.f; }   // channels 81 .. 83 of the last chunk
  }
  __syncthreads();
  for (int u = tid; u < 11 * TH * TW; u += THREADS) {
    const int q = u / (TH * TW), tp = u - q * (TH * TW);
    const int yy = y0 + tp / TW, xx = x0 + tp % TW;
    if (yy >= H || xx >= W) continue;
    const float4 v0 = *reinterpret_cast<const float4*>(so + tp * OUT_PITCH + q * 8);
    const float4 v1 = q < 10 ? *reinterpret_cast<const float4*>(so + tp * OUT_PITCH + q * 8 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    F8 f;
    f.v[0] = v0.x; f.v[1] = v0.y; f.v[2] = v0.z; f.v[3] = v0.w; f.v[4] = v1.x; f.v[5] = v1.y; f.v[6] = v1.z; f.v[7] = v1.w;
    st_chunk(a.out.hi, a.out.lo, cv_elem(a.out, n, q, yy, xx), f);
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

// CP8 features inside PWC-Net; -> 0 launched, 1 not selected (the caller runs corr81_cp8_kernel), else error
int corr81_cp8_tma(const CView& f1, const CView& f2, const CView& out, const CView& c1_copy, float slope, cudaStream_t st) {
  // Opt-in (PREMVOS_CORR_CP8_TMA=1).  Measured inside the batch-4 forward (tools/pwc_time.py): 0.254 ms per forward with this kernel
  // on level 2 and 0.311 ms on all five levels against 0.242 ms for the plain-load kernel -- in CP8 the cost volume is bound by
  // its OUTPUT (324 B per pixel leave against 256 B that arrive at C = 32) and by the hi / lo conversions around it, not by the
  // operand staging, and the small levels are a handful of tiles with up to 25 serial chunks.  Kept as a tested alternative.
  const char* env = getenv("PREMVOS_CORR_CP8_TMA");
  if (!(env && atoi(env) != 0)) return 1;
  alignas(64) unsigned char maps[4][128];
  const CView* views[2] = {&f1, &f2};
  for (int k = 0; k < 2; k++) {
    const CView& v = *views[k];
    const uint64_t dims[5] = {8, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.chunks, (uint64_t)v.N};
    const uint64_t strides[4] = {16, (uint64_t)v.W * 16, (uint64_t)v.W * v.H * 16, (uint64_t)v.W * v.H * 16 * v.chunks};
    const uint32_t box[5] = {8, (uint32_t)(k == 0 ? TW : HW), (uint32_t)(k == 0 ? TH : HH), 1, 1};
    PV_TRY(encode_tensor_map_bf16(maps[2 * k], v.hi, 5, dims, strides, box));
    PV_TRY(encode_tensor_map_bf16(maps[2 * k + 1], v.lo, 5, dims, strides, box));
  }
  static bool attr_set = false;
  if (!attr_set) {
    PV_CUDA(cudaFuncSetAttribute(corr81_cp8_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CP_SMEM));
    attr_set = true;
  }
  CorrCp8TmaArgs a{cp8::dev(f1), cp8::dev(f2), cp8::dev(out), cp8::dev(c1_copy), c1_copy.null() ? 0 : 1, (f1.W + TW - 1) / TW, (f1.H + TH - 1) / TH, slope};
  const long ctas = (long)a.tiles_x * a.tiles_y * f1.N;
  const double px = (double)f1.pixels();
  prof_before(st);
  corr81_cp8_tma_kernel<<<(unsigned)ctas, THREADS, CP_SMEM, st>>>(*reinterpret_cast<const CUtensorMap*>(maps[0]), *reinterpret_cast<const CUtensorMap*>(maps[1]),
                                                               *reinterpret_cast<const CUtensorMap*>(maps[2]), *reinterpret_cast<const CUtensorMap*>(maps[3]), a);
  return after_launch("corr81_cp8_kernel", st, 2.0 * 81 * f1.C * px, 4.0 * (2.0 * f1.C + 81) * px);
}

}  // namespace premvos
