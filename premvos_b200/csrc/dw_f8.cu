// Depthwise 3x3 convolution of the refinement network on F8 inputs (fp32 chunk-planar activations, FView in common.cuh):
// persistent CTAs, input tiles (with halo) staged by TMA through a 3-deep ring, zero padding = the TMA unit's out-of-bounds fill.
//
// Why a second format: a pointwise output that only feeds the next depthwise convolution (two of the three separable convolutions
// of every Xception unit, refinement_net/network/deeplab/core/xception.py:251-275) or the next unit's depthwise + residual add
// (middle flow) never is a tensor-core operand, so splitting it into bf16 hi/lo planes in the GEMM epilogue and re-assembling it
// in the depthwise kernel is pure instruction overhead (the CP8 depthwise kernel spends ~2/3 of its instructions on that and on
// staging loads; it reaches 3 TB/s of algorithmic bytes, instruction-bound).  Same bytes per element (4), no conversion: the TMA
// box [rows][cols][8 floats] lands in shared memory exactly as the compute phase reads it (one LDS.128 per 4 channels of a tap
// column, conflict-free), the unit's leading ReLU is applied by the producer where it is the only consumer.
// Output is CP8 (split bf16): it is the A operand of the pointwise GEMM.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "cp8.cuh"

namespace premvos {

using namespace cp8;

namespace {

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

constexpr int DW_STAGES_MAX = 4;

struct DwF8Args {
  const float* w;      // [9][cpad], BatchNorm scale folded in
  const float* bias;   // [cpad]
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int out_chunks, out_c0, Ho, Wo;
  int in_c0, nch, cpad;
  int stride, rate, pad;        // input coordinate = o * stride + tap * rate - pad
  int pre_relu, post_relu;
  int TH, TW, ntx, nty, RH, RW; // output tile, tiles per image, input region of a tile
  int stages, stage_bytes;
  long items;                   // n_active * nch * nty * ntx
  int n_active;
};

__device__ __forceinline__ void split2(float x0, float x1, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __uint_as_float(hw << 16), x1 - __uint_as_float(hw & 0xffff0000u));
  *hi = hw; *lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void fma4(float4& acc, const float4& v, const float4& w) {
  acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y); acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
}
__device__ __forceinline__ float4 relu4(const float4& v) { return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)); }

// FAST: stride 1, rate 1 -- a thread owns a vertical strip of RS outputs of one (column, 4-channel half) and slides a 3-row window
// down it (3 LDS.128 per input row).  Otherwise one output per thread-iteration, 9 LDS.128.  Tap order per output is r-major,
// s-minor in both paths: the same fp32 bits as the CP8 kernel and a plain loop.
template <int RS, bool FAST, bool PRE_RELU>
__global__ void __launch_bounds__(256, FAST ? 2 : 3) depthwise3x3_f8_kernel(const __grid_constant__ CUtensorMap tm_in, const DwF8Args a) {
  extern __shared__ __align__(128) float4 dw_sm[];   // [stages][stage_bytes / 16], then the barriers
  uint64_t* full = reinterpret_cast<uint64_t*>(dw_sm + (size_t)a.stages * (a.stage_bytes >> 4));
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // Items are ordered (chunk, image, tile) and every CTA takes a contiguous range: consecutive items of a CTA share their
  // chunk, the 9 tap weights + bias of a thread are re-read only when the chunk changes, and the (chunk, image, tile) digits
  // of the item are carried instead of divided out (the divisions cost more instructions than the 45 FMAs of an output).
  const int tiles = a.ntx * a.nty;
  const int items = (int)a.items;
  const int per_cta = (items + (int)gridDim.x - 1) / (int)gridDim.x;
  const int first = (int)blockIdx.x * per_cta, last = min(items, first + per_cta);
  struct Digits { int ch, n, ty, tx; };
  auto digits_of = [&](int item) {
    Digits d;
    const int per_chunk = a.n_active * tiles;
    d.ch = item / per_chunk;
    const int r = item - d.ch * per_chunk;
    d.n = r / tiles;
    const int tile = r - d.n * tiles;
    d.ty = tile / a.ntx; d.tx = tile - d.ty * a.ntx;
    return d;
  };
  auto advance = [&](Digits& d) {
    if (++d.tx == a.ntx) { d.tx = 0; if (++d.ty == a.nty) { d.ty = 0; if (++d.n == a.n_active) { d.n = 0; d.ch++; } } }
  };
  auto issue = [&](const Digits& d, int slot) {   // one thread
    mbar_arrive_expect_tx(&full[slot], (uint32_t)(a.RH * a.RW * 32));
    tma_load_4d(&tm_in, &full[slot], dw_sm + (size_t)slot * (a.stage_bytes >> 4), (d.tx * a.TW * a.stride - a.pad) * 8,
                d.ty * a.TH * a.stride - a.pad, a.in_c0 + d.ch, d.n);
  };
  Digits cur = digits_of(min(first, max(items - 1, 0))), ahead = cur;   // `ahead` = the item the next TMA is issued for (thread 0)
  if (threadIdx.x == 0)
    for (int s = 0; s + 1 < a.stages; s++)
      if (first + s < last) { issue(ahead, s); advance(ahead); }
  const int units = FAST ? (a.TH / RS) * a.TW * 2 : a.TH * a.TW * 2;
  const unsigned tw_magic = (1u << 20) / (unsigned)a.TW + 1u;   // p / TW for p < 2^20 / TW (host-checked)
  int slot = 0, cur_ch = -1;
  uint32_t phase = 0;
  float4 wv[9], bv;
  const int rw2 = a.RW * 2;
  for (int item = first; item < last; item++) {
    if (threadIdx.x == 0 && item + (a.stages - 1) < last) {
      issue(ahead, slot == 0 ? a.stages - 1 : slot - 1);   // the slot consumed in the previous iteration
      advance(ahead);
    }
    const int ch = cur.ch, n = cur.n;
    const int oy0 = cur.ty * a.TH, ox0 = cur.tx * a.TW;
    const float4* sm = dw_sm + (size_t)slot * (a.stage_bytes >> 4);   // [RH][RW][2 halves]
    const long out_plane = (((long)n * a.out_chunks + a.out_c0 + ch) * a.Ho) * a.Wo * 8;
    bool waited = false;
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
      const int half = u & 1, p = u >> 1, ty = (int)(((unsigned)p * tw_magic) >> 20), tx = p - ty * a.TW;
      const int ox = ox0 + tx;
      if (ch != cur_ch) {   // blockDim is even: a thread's `half` never changes
#pragma unroll
        for (int t = 0; t < 9; t++) wv[t] = __ldg(reinterpret_cast<const float4*>(a.w + t * a.cpad + ch * 8 + half * 4));
        bv = __ldg(reinterpret_cast<const float4*>(a.bias + ch * 8 + half * 4));
        cur_ch = ch;
      }
      if (!waited) { mbar_wait(&full[slot], phase); waited = true; }
      if (ox >= a.Wo) continue;
      if (FAST) {
        float4 acc[RS];
#pragma unroll
        for (int k = 0; k < RS; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* row = sm + ((ty * RS) * a.RW + tx) * 2 + half;
#pragma unroll
        for (int j = 0; j < RS + 2; j++, row += rw2) {
          float4 v0 = row[0], v1 = row[2], v2 = row[4];
          if (PRE_RELU) { v0 = relu4(v0); v1 = relu4(v1); v2 = relu4(v2); }
#pragma unroll
          for (int rr = 0; rr < 3; rr++) {
            const int k = j - rr;
            if (k < 0 || k >= RS) continue;
            fma4(acc[k], v0, wv[rr * 3]); fma4(acc[k], v1, wv[rr * 3 + 1]); fma4(acc[k], v2, wv[rr * 3 + 2]);
          }
        }
        const int oyb = oy0 + ty * RS;
        __nv_bfloat16* ph = a.out_hi + out_plane + ((long)oyb * a.Wo + ox) * 8 + half * 4;
        __nv_bfloat16* pl = a.out_lo + out_plane + ((long)oyb * a.Wo + ox) * 8 + half * 4;
#pragma unroll
        for (int k = 0; k < RS; k++, ph += a.Wo * 8, pl += a.Wo * 8) {
          if (oyb + k >= a.Ho) break;
          float4 t = make_float4(acc[k].x + bv.x, acc[k].y + bv.y, acc[k].z + bv.z, acc[k].w + bv.w);
          if (a.post_relu) t = relu4(t);
          uint32_t h0, l0, h1, l1;
          split2(t.x, t.y, &h0, &l0); split2(t.z, t.w, &h1, &l1);
          *reinterpret_cast<uint2*>(ph) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(pl) = make_uint2(l0, l1);
        }
      } else {
        const int oy = oy0 + ty;
        if (oy >= a.Ho) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
          const int sy = ty * a.stride + rr * a.rate;
#pragma unroll
          for (int s = 0; s < 3; s++) {
            const int sx = tx * a.stride + s * a.rate;
            float4 v = sm[(sy * a.RW + sx) * 2 + half];
            if (PRE_RELU) v = relu4(v);
            fma4(acc, v, wv[rr * 3 + s]);
          }
        }
        float4 t = make_float4(acc.x + bv.x, acc.y + bv.y, acc.z + bv.z, acc.w + bv.w);
        if (a.post_relu) t = relu4(t);
        uint32_t h0, l0, h1, l1;
        split2(t.x, t.y, &h0, &l0); split2(t.z, t.w, &h1, &l1);
        const long e = out_plane + ((long)oy * a.Wo + ox) * 8 + half * 4;
        *reinterpret_cast<uint2*>(a.out_hi + e) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(a.out_lo + e) = make_uint2(l0, l1);
      }
    }
    if (!waited) mbar_wait(&full[slot], phase);   // threads without a unit still observe the phase (keeps parity bookkeeping uniform)
    __syncthreads();                              // everybody is done with this slot before it is refilled
    if (++slot == a.stages) { slot = 0; phase ^= 1u; }
    advance(cur);
  }
}

}  // namespace

// Plans the launch (tensor map over the input view) once; launch with an active batch.
int plan_depthwise3x3_f8(DwF8Plan* plan, const FView& in, const CView& out, const float* w, const float* bias, int stride, int rate, int pad,
                         bool pre_relu, bool post_relu) {
  PV_CHECK(in.p && out.hi && in.C == out.C && in.N == out.N, PREMVOS_ERR_INVALID_ARG, "depthwise3x3_f8: shape mismatch");
  static_assert(sizeof(DwF8Args) <= sizeof(plan->args), "DwF8Plan::args too small");
  DwF8Args& a = *reinterpret_cast<DwF8Args*>(plan->args);
  memset(&a, 0, sizeof(a));
  const bool fast = stride == 1 && rate == 1;
  const int RS = 5;
  a.w = w; a.bias = bias;
  a.out_hi = out.hi; a.out_lo = out.lo; a.out_chunks = out.chunks; a.out_c0 = out.c0; a.Ho = out.H; a.Wo = out.W;
  a.in_c0 = in.c0; a.nch = (in.C + 7) / 8; a.cpad = a.nch * 8;
  a.stride = stride; a.rate = rate; a.pad = pad; a.pre_relu = pre_relu ? 1 : 0; a.post_relu = post_relu ? 1 : 0;
  if (fast) {   // 5 strips of 5 rows x <= 25 columns x 2 halves = <= 250 thread units
    a.ntx = (out.W + 24) / 25; a.TW = (out.W + a.ntx - 1) / a.ntx;
    const int strips = (out.H + RS - 1) / RS, sp = std::min(5, strips);
    a.TH = sp * RS; a.nty = (strips + sp - 1) / sp;
  } else {   // the widest tile whose input region still is one TMA box row (<= 32 pixels = 256 floats)
    const int tw_max = std::max(1, (32 - 2 * rate - 1) / stride + 1);
    a.ntx = (out.W + tw_max - 1) / tw_max; a.TW = (out.W + a.ntx - 1) / a.ntx;
    const int th_max = std::min(tw_max, 25);
    a.nty = (out.H + th_max - 1) / th_max; a.TH = (out.H + a.nty - 1) / a.nty;
  }
  a.RH = (a.TH - 1) * stride + 2 * rate + 1; a.RW = (a.TW - 1) * stride + 2 * rate + 1;
  PV_CHECK(a.RH <= 256 && a.RW <= 32 && (long)a.TH * a.TW * a.TW < (1 << 20), PREMVOS_ERR_UNSUPPORTED,
           "depthwise3x3_f8: tile too large (stride %d rate %d)", stride, rate);
  a.stage_bytes = round_up(a.RH * a.RW * 32, 128);
  a.stages = std::max(2, std::min(DW_STAGES_MAX, (72 * 1024) / a.stage_bytes));
  if (getenv("PREMVOS_DWF8_STAGES")) a.stages = std::max(2, std::min(DW_STAGES_MAX, atoi(getenv("PREMVOS_DWF8_STAGES"))));
  plan->smem_bytes = a.stages * a.stage_bytes + 64;
  PV_CHECK(plan->smem_bytes <= 200 * 1024, PREMVOS_ERR_UNSUPPORTED, "depthwise3x3_f8: tile does not fit shared memory");
  plan->fast = fast ? 1 : 0; plan->N = in.N; plan->tiles = a.ntx * a.nty; plan->nch = a.nch;
  plan->bytes_per_image = 4.0 * ((double)in.H * in.W + (double)out.H * out.W) * in.C;
  plan->flops_per_image = 18.0 * out.H * out.W * a.nch * 8;
  // rows of W * 8 floats (pixels of a row are contiguous): one TMA request per box row of RW * 32 bytes
  const uint64_t dims[4] = {(uint64_t)in.W * 8, (uint64_t)in.H, (uint64_t)in.chunks, (uint64_t)in.N};
  const uint64_t strides[3] = {(uint64_t)in.W * 32, (uint64_t)in.H * in.W * 32, (uint64_t)in.chunks * in.H * in.W * 32};
  const uint32_t box[4] = {(uint32_t)a.RW * 8, (uint32_t)a.RH, 1, 1};
  return encode_tensor_map_f32(plan->map_in, in.p, 4, dims, strides, box);
}

int launch_depthwise3x3_f8(const DwF8Plan& plan, int n_active, cudaStream_t st) {
  DwF8Args a = *reinterpret_cast<const DwF8Args*>(plan.args);
  if (n_active < 0 || n_active > plan.N) n_active = plan.N;
  a.items = (long)n_active * plan.nch * plan.tiles;
  a.n_active = n_active;
  if (a.items == 0) return 0;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PV_CUDA(cudaGetDevice(&dev));
    PV_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  PV_CHECK(a.items < (1L << 30), PREMVOS_ERR_UNSUPPORTED, "depthwise3x3_f8: too many tiles");
  typedef void (*Kern)(const CUtensorMap, const DwF8Args);
  static const Kern kerns[4] = {depthwise3x3_f8_kernel<5, false, false>, depthwise3x3_f8_kernel<5, false, true>,
                                depthwise3x3_f8_kernel<5, true, false>, depthwise3x3_f8_kernel<5, true, true>};
  static bool attr_set = false;
  if (!attr_set) {
    for (Kern k : kerns) PV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  Kern kern = kerns[(plan.fast ? 2 : 0) + (a.pre_relu ? 1 : 0)];
  // resident CTAs per SM: shared memory and the register budget of the launch bounds (256 threads: 2 for the strip kernels, 3 for
  // the generic ones); one CTA per resident slot -- more would run as a second, half-empty wave
  static const int threads = getenv("PREMVOS_DWF8_THREADS") ? atoi(getenv("PREMVOS_DWF8_THREADS")) : 256;
  int per_sm = std::max(1, std::min((plan.fast ? 2 : 3) * (256 / threads), (220 * 1024) / plan.smem_bytes));
  if (getenv("PREMVOS_DWF8_PER_SM")) per_sm = atoi(getenv("PREMVOS_DWF8_PER_SM"));
  const int grid = (int)std::min<long>(a.items, (long)num_sms * per_sm);
  prof_before(st);
  kern<<<grid, threads, plan.smem_bytes, st>>>(*reinterpret_cast<const CUtensorMap*>(plan.map_in), a);
  const char* label = "depthwise3x3_kernel";
  static const int per_layer = getenv("PREMVOS_PROFILE_LAYERS") ? atoi(getenv("PREMVOS_PROFILE_LAYERS")) : 0;
  if (per_layer && profiling_enabled()) {
    char buf[160];
    snprintf(buf, sizeof(buf), "dw_f8[n%d_%dx%d_c%d_s%d_r%d|tile%dx%d_st%d_%s]", n_active, a.Ho, a.Wo, a.nch * 8, a.stride, a.rate, a.TH, a.TW, a.stages,
             plan.fast ? "strip" : "generic");
    label = prof_intern(buf);
  }
  return after_launch(label, st, plan.flops_per_image * n_active, plan.bytes_per_image * n_active);
}

}  // namespace premvos
