// Tensor-core convolution for sm_100a: implicit GEMM on tcgen05.mma with TMA-fed shared memory and
// the accumulator in tensor memory (TMEM).
//
// GEMM view of a stride-1 RxS convolution on channels-last activations:
//     D[m, n] = sum_{tap=(r,s)} sum_{c} A_tap[m, c] * W[tap][n][c]
//   m = output pixel inside a TH x TW tile (TH*TW = 128 = UMMA M), n = output channel, c = input channel.
// im2col never materialises: the A tile of tap (r,s) is ONE 4-D TMA box {64 ch, TW, TH, 1} of the
// input tensor at pixel offset ((s-S/2)*dil, (r-R/2)*dil); out-of-image pixels and channels beyond
// the view are zero-filled by the TMA unit, which is exactly the convolution's zero padding (and the
// K tail).  The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, i.e. the
// canonical K-major UMMA operand layout.
//
// Numerics: activations and weights are stored as SPLIT bf16 -- x = hi + lo with hi = bf16(x),
// lo = bf16(x - hi) (16 mantissa bits in total).  Each K step issues three kind::f16 MMAs into the
// same fp32 TMEM accumulator:  hi*hi + hi*lo + lo*hi  (the dropped lo*lo term is ~2^-16 relative).
// This keeps the 1e-3 parity contract with a large margin at 1.5x the cost of a single TF32 pass
// (bf16 runs at twice the TF32 rate) instead of TF32's 2^-11 operand rounding.
//
// Warp roles (256 threads, 1 CTA/SM): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer,
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM lane quarter = warp_id % 4):
// tcgen05.ld -> +bias -> LeakyReLU/ReLU -> split to hi/lo -> 16-byte stores into the output view
// (which may be a channel range of a DenseNet slab).
#include <cuda.h>
#include <cuda_bf16.h>

#include <vector>

#include "common.cuh"

namespace premvos {

namespace {

constexpr int TILE_M = 128;
constexpr int KBLK = 64;           // bf16 channels per pipeline stage = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = TILE_M * KBLK * 2;  // 16 KB
constexpr int UMMA_THREADS = 256;
constexpr int SMEM_LIMIT = 227 * 1024;

// ---- PTX wrappers -----------------------------------------------------------------------------
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
// Bounded spin: a protocol bug must become a trap (launch error), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle: rows of 128 B, 8-row atoms (1024 B) stacked along M/N.
// bits [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64)
// | [46,48) version=1 (sm_100) | [61,64) layout=2 (SWIZZLE_128B)      (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32 (bits 4-5 =1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct UmmaConvArgs {
  int tiles_x, tiles_y;      // tiles per image
  int TW, TH;                // tile shape, TW*TH == 128
  int H, W;                  // output (== input) spatial size
  int in_coff;               // first input channel inside the pixel
  int kblocks;               // ceil(Cin / 64)
  int R, S, dil;
  int BN;                    // N tile (multiple of 32, <= 256)
  int Cout;
  int stages;
  const float* bias;         // [CoutP]
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; float* out_f32;  // split and/or fp32 output
  int out_cs, out_coff;
  float slope;
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 const UmmaConvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle (TMA write and UMMA read agree on address bits)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int w_tile_bytes = a.BN * KBLK * 2;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * w_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + a.stages;
  uint64_t* tmem_full_bar = empty_bar + a.stages;
  uint32_t* tmem_addr_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int n_img = tile / (a.tiles_x * a.tiles_y);
  const int trem = tile - n_img * a.tiles_x * a.tiles_y;
  const int ty0 = (trem / a.tiles_x) * a.TH, tx0 = (trem % a.tiles_x) * a.TW;
  const int n0 = blockIdx.y * a.BN;
  const int kiters = a.R * a.S * a.kblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
    for (int s = 0; s < a.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_addr_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_addr_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    const uint32_t tx_bytes = (uint32_t)stage_bytes;
    int it = 0;
    for (int tap = 0; tap < a.R * a.S; tap++) {
      const int r = tap / a.S, s = tap - r * a.S;
      const int ix = tx0 + (s - a.S / 2) * a.dil, iy = ty0 + (r - a.R / 2) * a.dil;
      for (int kb = 0; kb < a.kblocks; kb++, it++) {
        const int st = it % a.stages;
        const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
        mbar_wait(&empty_bar[st], ph ^ 1u);
        uint8_t* sp = smem + (size_t)st * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[st], tx_bytes);
        tma_load_4d(&tmA_hi, &full_bar[st], sp, a.in_coff + kb * KBLK, ix, iy, n_img);
        tma_load_4d(&tmA_lo, &full_bar[st], sp + A_TILE_BYTES, a.in_coff + kb * KBLK, ix, iy, n_img);
        tma_load_3d(&tmW_hi, &full_bar[st], sp + 2 * A_TILE_BYTES, kb * KBLK, n0, tap);
        tma_load_3d(&tmW_lo, &full_bar[st], sp + 2 * A_TILE_BYTES + w_tile_bytes, kb * KBLK, n0, tap);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_bf16(TILE_M, a.BN);
    for (int it = 0; it < kiters; it++) {
      const int st = it % a.stages;
      const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
      const uint64_t dA_hi = make_sw128_desc(sa), dA_lo = make_sw128_desc(sa + A_TILE_BYTES);
      const uint64_t dW_hi = make_sw128_desc(sa + 2 * A_TILE_BYTES), dW_lo = make_sw128_desc(sa + 2 * A_TILE_BYTES + w_tile_bytes);
#pragma unroll
      for (int k = 0; k < KBLK / UMMA_K; k++) {
        const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);  // advance inside the 128-byte swizzle row
        umma_bf16(tmem_base, dA_lo + koff, dW_hi + koff, idesc, (it | k) != 0);
        umma_bf16(tmem_base, dA_hi + koff, dW_lo + koff, idesc, 1u);
        umma_bf16(tmem_base, dA_hi + koff, dW_hi + koff, idesc, 1u);
      }
      umma_commit(&empty_bar[st]);  // frees the smem slot once these MMAs have read it
    }
    umma_commit(tmem_full_bar);     // accumulator complete
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;         // accumulator row == pixel inside the tile
    const int oy = ty0 + m / a.TW, ox = tx0 + m % a.TW;
    const bool in_img = oy < a.H && ox < a.W;
    const size_t pix = ((size_t)n_img * a.H + oy) * a.W + ox;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < a.BN; c0 += 32) {
      uint32_t v[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      const int co0 = n0 + c0;
      if (in_img && co0 < a.Cout) {
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; j++) {
        float t = __uint_as_float(v[j]) + __ldg(a.bias + co0 + j);
        f[j] = t > 0.f ? t : t * a.slope;
      }
      const bool full = co0 + 32 <= a.Cout;
      if (a.out_hi) {
        __nv_bfloat16* ph = a.out_hi + pix * a.out_cs + a.out_coff + co0;
        __nv_bfloat16* pl = a.out_lo + pix * a.out_cs + a.out_coff + co0;
        if (full) {
          uint32_t hw[16], lw[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
            __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * j] - __bfloat162float(h0));
            __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * j + 1] - __bfloat162float(h1));
            hw[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lw[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            reinterpret_cast<uint4*>(ph)[j] = make_uint4(hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
            reinterpret_cast<uint4*>(pl)[j] = make_uint4(lw[4 * j], lw[4 * j + 1], lw[4 * j + 2], lw[4 * j + 3]);
          }
        } else {
          for (int j = 0; j < 32 && co0 + j < a.Cout; j++) {
            __nv_bfloat16 h = __float2bfloat16_rn(f[j]);
            ph[j] = h;
            pl[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h));
          }
        }
      }
      if (a.out_f32) {
        float* pf = a.out_f32 + pix * a.out_cs + a.out_coff + co0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 8; j++) reinterpret_cast<float4*>(pf)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
          for (int j = 0; j < 32 && co0 + j < a.Cout; j++) pf[j] = f[j];
        }
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    PV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    PV_CHECK(p && q == cudaDriverEntryPointSuccess, PREMVOS_ERR_NO_DEVICE, "cuTensorMapEncodeTiled is not available in this driver");
    fn = (EncodeTiledFn)p;
  }
  *out = fn;
  return 0;
}

int encode_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn;
  PV_TRY(get_encode_fn(&fn));
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PV_CHECK(r == CUDA_SUCCESS, PREMVOS_ERR_INVALID_ARG, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return 0;
}

inline void split_bf16(float x, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  *hi = __float2bfloat16_rn(x);
  *lo = __float2bfloat16_rn(x - __bfloat162float(*hi));
}

}  // namespace

int pack_conv_weights_umma(ConvWeightsUmma* out, const float* host_w, const float* host_b, int Cout, int Cin, int R, int S) {
  out->R = R; out->S = S; out->Cin = Cin; out->Cout = Cout;
  out->KP = round_up(Cin, KBLK);
  int bn = round_up(Cout, 32);
  if (bn > 128) bn = 128;
  out->BN = bn;
  out->CoutP = round_up(Cout, bn);
  size_t n = (size_t)R * S * out->CoutP * out->KP;
  std::vector<__nv_bfloat16> hi(n, __float2bfloat16_rn(0.f)), lo(n, __float2bfloat16_rn(0.f));
  std::vector<float> b(out->CoutP, 0.f);
  for (int co = 0; co < Cout; co++) {
    b[co] = host_b ? host_b[co] : 0.f;
    for (int ci = 0; ci < Cin; ci++)
      for (int t = 0; t < R * S; t++) {
        size_t idx = ((size_t)t * out->CoutP + co) * out->KP + ci;
        split_bf16(host_w[((size_t)co * Cin + ci) * R * S + t], &hi[idx], &lo[idx]);
      }
  }
  PV_CUDA(cudaMalloc((void**)&out->w_hi, n * 2));
  PV_CUDA(cudaMalloc((void**)&out->w_lo, n * 2));
  PV_CUDA(cudaMalloc((void**)&out->bias, out->CoutP * sizeof(float)));
  PV_CUDA(cudaMemcpy(out->w_hi, hi.data(), n * 2, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(out->w_lo, lo.data(), n * 2, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(out->bias, b.data(), out->CoutP * sizeof(float), cudaMemcpyHostToDevice));
  cuuint64_t dims[3] = {(cuuint64_t)out->KP, (cuuint64_t)out->CoutP, (cuuint64_t)(R * S)};
  cuuint64_t strides[2] = {(cuuint64_t)out->KP * 2, (cuuint64_t)out->KP * out->CoutP * 2};
  cuuint32_t box[3] = {(cuuint32_t)KBLK, (cuuint32_t)out->BN, 1};
  PV_TRY(encode_map((CUtensorMap*)out->map_hi, out->w_hi, 3, dims, strides, box));
  PV_TRY(encode_map((CUtensorMap*)out->map_lo, out->w_lo, 3, dims, strides, box));
  return 0;
}

void free_conv_weights_umma(ConvWeightsUmma* w) {
  cudaFree(w->w_hi); cudaFree(w->w_lo); cudaFree(w->bias);
  w->w_hi = w->w_lo = nullptr; w->bias = nullptr;
}

bool conv_umma_supported(int Cin, int Cout, int R, int S, int stride) {
  return stride == 1 && R == S && (R == 1 || R == 3) && Cin >= 16 && Cout >= 16;
}

// Plans one convolution launch: tensor maps of the input view are encoded here (host only, no device work).
int plan_conv_umma(ConvPlanUmma* plan, const TView& in, const TView& out, const ConvWeightsUmma& w, int dil, float slope) {
  PV_CHECK(in.split() && (out.split() || out.p), PREMVOS_ERR_INVALID_ARG, "conv_umma: input must be split-bf16");
  PV_CHECK(in.C == w.Cin && out.C == w.Cout && in.N == out.N && in.H == out.H && in.W == out.W, PREMVOS_ERR_INVALID_ARG,
           "conv_umma: shape mismatch");
  PV_CHECK((in.cs % 8) == 0 && (out.cs % 8) == 0 && ((out.coff) % 8) == 0, PREMVOS_ERR_INVALID_ARG,
           "conv_umma: views must be 16-byte addressable (cs=%d/%d coff=%d)", in.cs, out.cs, out.coff);
  int TW = 1;
  while (TW < in.W && TW < 32) TW <<= 1;   // 8..32 wide tiles; narrow images get taller tiles
  if (TW < 8) TW = 8;
  int TH = TILE_M / TW;
  UmmaConvArgs& a = *reinterpret_cast<UmmaConvArgs*>(plan->args);
  static_assert(sizeof(UmmaConvArgs) <= sizeof(plan->args), "ConvPlanUmma::args too small");
  a.TW = TW; a.TH = TH;
  a.tiles_x = (in.W + TW - 1) / TW; a.tiles_y = (in.H + TH - 1) / TH;
  a.H = in.H; a.W = in.W; a.in_coff = in.coff; a.kblocks = (in.C + KBLK - 1) / KBLK;
  a.R = w.R; a.S = w.S; a.dil = dil; a.BN = w.BN; a.Cout = w.Cout;
  a.bias = w.bias;
  a.out_hi = out.hi; a.out_lo = out.lo; a.out_f32 = out.split() ? nullptr : out.p;
  a.out_cs = out.cs; a.out_coff = out.coff; a.slope = slope;
  uint32_t cols = 32;
  while ((int)cols < w.BN) cols <<= 1;
  a.tmem_cols = cols;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * w.BN * KBLK * 2;
  int stages = (SMEM_LIMIT - 2048) / stage_bytes;
  if (stages > 6) stages = 6;
  a.stages = stages;
  plan->smem_bytes = stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
  plan->grid_x = a.tiles_x * a.tiles_y * in.N;
  plan->grid_y = w.CoutP / w.BN;
  // input tensor maps: dims {channels addressable = coff + C, W, H, N}; box {64, TW, TH, 1}
  cuuint64_t dims[4] = {(cuuint64_t)(in.coff + in.C), (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.N};
  cuuint64_t strides[3] = {(cuuint64_t)in.cs * 2, (cuuint64_t)in.W * in.cs * 2, (cuuint64_t)in.H * in.W * in.cs * 2};
  cuuint32_t box[4] = {(cuuint32_t)KBLK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  PV_TRY(encode_map((CUtensorMap*)plan->map_a_hi, in.hi, 4, dims, strides, box));
  PV_TRY(encode_map((CUtensorMap*)plan->map_a_lo, in.lo, 4, dims, strides, box));
  memcpy(plan->map_w_hi, w.map_hi, 128);
  memcpy(plan->map_w_lo, w.map_lo, 128);
  plan->flops = 2.0 * in.N * in.H * in.W * (double)w.Cout * w.Cin * w.R * w.S;
  plan->bytes = 4.0 * ((double)in.pixels() * in.C + (double)out.pixels() * w.Cout + (double)w.R * w.S * w.Cin * w.Cout);
  return 0;
}

int launch_conv_umma(const ConvPlanUmma& plan, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    PV_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set = true;
  }
  const UmmaConvArgs& a = *reinterpret_cast<const UmmaConvArgs*>(plan.args);
  prof_before(st);
  conv_umma_kernel<<<dim3(plan.grid_x, plan.grid_y), UMMA_THREADS, plan.smem_bytes, st>>>(
      *reinterpret_cast<const CUtensorMap*>(plan.map_a_hi), *reinterpret_cast<const CUtensorMap*>(plan.map_a_lo),
      *reinterpret_cast<const CUtensorMap*>(plan.map_w_hi), *reinterpret_cast<const CUtensorMap*>(plan.map_w_lo), a);
  return after_launch("conv_umma_kernel", st, plan.flops, plan.bytes);
}

}  // namespace premvos
