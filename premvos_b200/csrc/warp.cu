// Warping layer and input packing.
//
// warp_kernel restates PWCDCNet.warp (models/PWCNet.py:140-176) as ONE kernel: the reference builds
// the sampling grid on the CPU, uploads it, uploads a ones tensor of the full feature size, and
// runs grid_sample twice plus four elementwise ops.  Here every output pixel computes its own
// sampling position (same normalise -> un-normalise arithmetic as grid_sample with the torch-0.2
// convention align_corners=True), gathers the four taps (zeros padding) and derives the validity
// mask analytically: mask = 1 iff the sum of the in-bounds bilinear weights >= 0.9999 (PWCNet.py:173).
// One warp per pixel, float4 per lane over channels -> coalesced 128..512-byte rows; HBM-bound:
// algorithmic bytes = 4*(2*C + 2)*H*W.
#include "common.cuh"

namespace premvos {

struct WarpArgs {
  const float* x; int x_cs, x_coff;
  const float* flow; const __nv_bfloat16 *flow_hi, *flow_lo; int f_cs, f_coff;
  float* out; int o_cs, o_coff;
  int N, H, W, C;
  float scale;
};

template <bool FLOW_SPLIT>
__global__ void __launch_bounds__(256) warp_kernel(WarpArgs a) {
  const int lane = threadIdx.x & 31;
  const long pix = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long P = (long)a.N * a.H * a.W;
  if (pix >= P) return;
  const int n = (int)(pix / ((long)a.H * a.W));
  const int rem = (int)(pix - (long)n * a.H * a.W);
  const int y = rem / a.W, x = rem - y * a.W;
  const long fi = pix * a.f_cs + a.f_coff;
  const float u = __fmul_rn(FLOW_SPLIT ? ld_split(a.flow_hi, a.flow_lo, fi) : a.flow[fi], a.scale);
  const float v = __fmul_rn(FLOW_SPLIT ? ld_split(a.flow_hi, a.flow_lo, fi + 1) : a.flow[fi + 1], a.scale);
  // PWCNet.py:157-162 then grid_sample's un-normalisation ((g+1)/2)*(size-1)
  const float wm1 = (float)max(a.W - 1, 1), hm1 = (float)max(a.H - 1, 1);
  float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)x, u)), wm1), 1.0f);
  float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)y, v)), hm1), 1.0f);
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(a.W - 1));
  float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(a.H - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  // same weight expressions as torch's grid_sampler: (x_se - ix), (ix - x_nw)
  float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.0f) - ix, wy0 = (fy0 + 1.0f) - iy;
  // guard the int conversion against huge flows
  fx0 = fminf(fmaxf(fx0, -2.0f), (float)a.W + 1.0f);
  fy0 = fminf(fmaxf(fy0, -2.0f), (float)a.H + 1.0f);
  int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  bool vx0 = x0 >= 0 && x0 < a.W, vx1 = x1 >= 0 && x1 < a.W;
  bool vy0 = y0 >= 0 && y0 < a.H, vy1 = y1 >= 0 && y1 < a.H;
  float w00 = (vy0 && vx0) ? wy0 * wx0 : 0.f;
  float w01 = (vy0 && vx1) ? wy0 * wx1 : 0.f;
  float w10 = (vy1 && vx0) ? wy1 * wx0 : 0.f;
  float w11 = (vy1 && vx1) ? wy1 * wx1 : 0.f;
  float msum = ((w00 + w01) + w10) + w11;
  const bool valid = msum >= 0.9999f;
  const long img = (long)n * a.H * a.W;
  const float* p00 = a.x + (img + (long)(vy0 ? y0 : 0) * a.W + (vx0 ? x0 : 0)) * a.x_cs + a.x_coff;
  const float* p01 = a.x + (img + (long)(vy0 ? y0 : 0) * a.W + (vx1 ? x1 : 0)) * a.x_cs + a.x_coff;
  const float* p10 = a.x + (img + (long)(vy1 ? y1 : 0) * a.W + (vx0 ? x0 : 0)) * a.x_cs + a.x_coff;
  const float* p11 = a.x + (img + (long)(vy1 ? y1 : 0) * a.W + (vx1 ? x1 : 0)) * a.x_cs + a.x_coff;
  float* po = a.out + pix * a.o_cs + a.o_coff;
  for (int c = lane * 4; c < a.C; c += 128) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      float4 t00 = *reinterpret_cast<const float4*>(p00 + c);
      float4 t01 = *reinterpret_cast<const float4*>(p01 + c);
      float4 t10 = *reinterpret_cast<const float4*>(p10 + c);
      float4 t11 = *reinterpret_cast<const float4*>(p11 + c);
      r.x = ((t00.x * w00 + t01.x * w01) + t10.x * w10) + t11.x * w11;
      r.y = ((t00.y * w00 + t01.y * w01) + t10.y * w10) + t11.y * w11;
      r.z = ((t00.z * w00 + t01.z * w01) + t10.z * w10) + t11.z * w11;
      r.w = ((t00.w * w00 + t01.w * w01) + t10.w * w10) + t11.w * w11;
    }
    *reinterpret_cast<float4*>(po + c) = r;
  }
}

int warp_nhwc(const TView& x2, const TView& flow, float flow_scale, const TView& out, cudaStream_t st) {
  PV_CHECK(x2.N == out.N && x2.H == out.H && x2.W == out.W && x2.C == out.C && flow.H == x2.H && flow.W == x2.W &&
               flow.N == x2.N && flow.C == 2, PREMVOS_ERR_INVALID_ARG, "warp_nhwc: shape mismatch");
  PV_CHECK((x2.C % 4) == 0 && (x2.cs % 4) == 0 && (x2.coff % 4) == 0 && (out.cs % 4) == 0 && (out.coff % 4) == 0,
           PREMVOS_ERR_INVALID_ARG, "warp_nhwc: views must be float4-addressable");
  WarpArgs a;
  a.x = x2.p; a.x_cs = x2.cs; a.x_coff = x2.coff;
  PV_CHECK(!x2.split() && !out.split(), PREMVOS_ERR_INVALID_ARG, "warp_nhwc: features must be fp32 views");
  a.flow = flow.p; a.flow_hi = flow.hi; a.flow_lo = flow.lo; a.f_cs = flow.cs; a.f_coff = flow.coff;
  a.out = out.p; a.o_cs = out.cs; a.o_coff = out.coff;
  a.N = x2.N; a.H = x2.H; a.W = x2.W; a.C = x2.C; a.scale = flow_scale;
  long P = (long)x2.N * x2.H * x2.W;
  prof_before(st);
  if (flow.split()) warp_kernel<true><<<(unsigned)((P + 7) / 8), 256, 0, st>>>(a);
  else warp_kernel<false><<<(unsigned)((P + 7) / 8), 256, 0, st>>>(a);
  return after_launch("warp_kernel", st, 8.0 * P * x2.C, 4.0 * P * (2.0 * x2.C + 2));
}

// x NCHW [B,6,H,W] -> img channels-last [2B,H,W,4]; images 0..B-1 = frame 1 of each pair, B..2B-1 = frame 2
__global__ void __launch_bounds__(256) pack_pair_kernel(const float* __restrict__ x, float* __restrict__ img, int B, int H, int W) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long HW = (long)H * W;
  long total = 2L * B * HW;
  if (idx >= total) return;
  int n2 = (int)(idx / HW);
  long r = idx - (long)n2 * HW;
  int b = n2 % B, which = n2 / B;
  const float* src = x + ((long)b * 6 + which * 3) * HW + r;
  float4 v = make_float4(src[0], src[HW], src[2 * HW], 0.f);
  *reinterpret_cast<float4*>(img + idx * 4) = v;
}

int pack_pair_input(const float* x_nchw, int B, int H, int W, const TView& img, cudaStream_t st) {
  PV_CHECK(img.N == 2 * B && img.H == H && img.W == W && img.cs == 4 && img.coff == 0, PREMVOS_ERR_INVALID_ARG,
           "pack_pair_input: bad image view");
  long total = 2L * B * H * W;
  prof_before(st);
  pack_pair_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x_nchw, img.p, B, H, W);
  return after_launch("pack_pair_kernel", st, 0.0, 4.0 * total * 7);
}

}  // namespace premvos
