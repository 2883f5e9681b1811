// Single-layer C-ABI entry point of the tensor-core convolution (bring-up / parity hook): NCHW fp32 in,
// NCHW fp32 out, everything in between exactly as the fused networks run it (CP8 split-bf16 planes,
// TMA + tcgen05 implicit GEMM).  Allocates its scratch per call -- not a hot-path function.
#include <vector>

#include <stdlib.h>

#include "common.cuh"

using namespace premvos;

namespace {
struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
  int cview(CView* v, int N, int C, int H, int W) {
    v->N = N; v->H = H; v->W = W; v->chunks = (C + 7) / 8; v->c0 = 0; v->C = C;
    const size_t bytes = (size_t)N * v->chunks * H * W * 8 * sizeof(__nv_bfloat16) + 128;  // slack for flattened 1x1 layers
    for (int k = 0; k < 2; k++) {
      void* p = nullptr;
      PV_CUDA(cudaMalloc(&p, bytes));
      ptrs.push_back(p);
      (k == 0 ? v->hi : v->lo) = (__nv_bfloat16*)p;
    }
    return 0;
  }
};
}  // namespace

extern "C" int premvos_conv2d_forward(const float* x_dev, const float* w_host, const float* bias_host, const float* residual_dev,
                                      float* out_dev, int batch, int cin, int height, int width, int cout, int kh, int kw,
                                      int stride, int dilation, int pad_top, int pad_left, int pad_bottom, int pad_right,
                                      float slope, void* stream) {
  PV_CHECK(x_dev && w_host && out_dev, PREMVOS_ERR_INVALID_ARG, "premvos_conv2d_forward: null argument");
  PV_CHECK(batch > 0 && cin > 0 && cout > 0 && height > 0 && width > 0 && kh > 0 && kw > 0 && dilation > 0, PREMVOS_ERR_INVALID_ARG,
           "premvos_conv2d_forward: non-positive size");
  PV_CHECK(stride == 1 || stride == 2, PREMVOS_ERR_UNSUPPORTED, "premvos_conv2d_forward: stride must be 1 or 2 (got %d)", stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = (height + pad_top + pad_bottom - dilation * (kh - 1) - 1) / stride + 1;
  const int Wo = (width + pad_left + pad_right - dilation * (kw - 1) - 1) / stride + 1;
  PV_CHECK(Ho > 0 && Wo > 0, PREMVOS_ERR_INVALID_ARG, "premvos_conv2d_forward: empty output");
  Scratch sc;
  CView in, out, res;
  PV_TRY(sc.cview(&in, batch, cin, height, width));
  PV_TRY(sc.cview(&out, batch, cout, Ho, Wo));
  PV_TRY(nchw_to_cp8(x_dev, in, st));
  ConvOut o;
  o.cp = out;
  if (residual_dev) {
    PV_TRY(sc.cview(&res, batch, cout, Ho, Wo));
    PV_TRY(nchw_to_cp8(residual_dev, res, st));
    o.res = res;
  }
  ConvWeightsUmma w;
  const bool flat = kh == 1 && kw == 1 && stride == 1 && pad_top == 0 && pad_left == 0 && pad_bottom == 0 && pad_right == 0;
  PV_TRY(pack_conv_weights_umma(&w, w_host, bias_host, cout, cin, kh, kw, nullptr, 0, 0, (long)batch * Ho * Wo, flat));
  ConvGeom g;
  g.stride = stride; g.dil = dilation; g.pad_t = pad_top; g.pad_l = pad_left; g.pad_b = pad_bottom; g.pad_r = pad_right;
  g.slope = slope;
  ConvPlanUmma plan;
  int r = plan_conv_umma(&plan, in, o, w, g);
  if (r == 0) r = launch_conv_umma(plan, st);
  // tuning aid: PREMVOS_CONV_REPEAT=n launches the same convolution n more times back to back (idempotent), so that a
  // profile shows the kernel without a different kernel before it
  if (const char* rep = getenv("PREMVOS_CONV_REPEAT"))
    for (int i = 0; i < atoi(rep) && r == 0; i++) r = launch_conv_umma(plan, st);
  if (r == 0) r = cp8_to_nchw(out, 0, out_dev, st);
  cudaError_t e = cudaStreamSynchronize(st);
  free_conv_weights_umma(&w);
  free_conv_plan_umma(&plan);
  if (r == 0 && e != cudaSuccess) r = fail((int)e, "premvos_conv2d_forward: %s", cudaGetErrorString(e));
  return r;
}

namespace {
// NCHW fp32 -> F8 ([N][C/8][H][W][8] fp32, zero in the padding channels)
__global__ void __launch_bounds__(256) nchw_to_f8_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int C, int H, int W) {
  const int chunks = (C + 7) / 8;
  const long hw = (long)H * W, total = (long)N * chunks * hw;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long p = idx % hw;
  const int ch = (int)((idx / hw) % chunks), n = (int)(idx / (hw * chunks));
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = ch * 8 + j < C ? src[((long)n * C + ch * 8 + j) * hw + p] : 0.f;
  float4* d = reinterpret_cast<float4*>(dst + idx * 8);
  d[0] = make_float4(v[0], v[1], v[2], v[3]);
  d[1] = make_float4(v[4], v[5], v[6], v[7]);
}
}  // namespace

// Separable convolution as the refinement network's entry flow runs it (bring-up / parity hook of the fused kernel):
// out = act(pointwise(relu_out?(depthwise3x3(relu_in?(x)) + dw_bias)) + pw_bias), depthwise SAME padding, stride 1.
extern "C" int premvos_sepconv2d_forward(const float* x_dev, const float* dw_w_host /*[C][3][3]*/, const float* dw_bias_host /*[C] or NULL*/,
                                         const float* pw_w_host /*[Cout][C]*/, const float* pw_bias_host /*[Cout] or NULL*/, float* out_dev,
                                         int batch, int channels, int height, int width, int cout, int relu_in, int relu_mid, float slope,
                                         void* stream) {
  PV_CHECK(x_dev && dw_w_host && pw_w_host && out_dev, PREMVOS_ERR_INVALID_ARG, "premvos_sepconv2d_forward: null argument");
  PV_CHECK(batch > 0 && channels > 0 && cout > 0 && height > 0 && width > 0, PREMVOS_ERR_INVALID_ARG, "premvos_sepconv2d_forward: non-positive size");
  PV_CHECK(cout <= 128, PREMVOS_ERR_UNSUPPORTED, "premvos_sepconv2d_forward: one output-channel tile (cout <= 128), got %d", cout);
  cudaStream_t st = (cudaStream_t)stream;
  Scratch sc;
  const int chunks = (channels + 7) / 8, cpad = round_up(chunks, 4) * 8;
  FView in;
  in.N = batch; in.H = height; in.W = width; in.chunks = chunks; in.c0 = 0; in.C = channels;
  PV_CUDA(cudaMalloc((void**)&in.p, (size_t)batch * chunks * height * width * 8 * sizeof(float)));
  sc.ptrs.push_back(in.p);
  CView out;
  PV_TRY(sc.cview(&out, batch, cout, height, width));
  std::vector<float> w((size_t)9 * cpad, 0.f), b(cpad, 0.f);
  for (int c = 0; c < channels; c++) {
    for (int t = 0; t < 9; t++) w[(size_t)t * cpad + c] = dw_w_host[(size_t)c * 9 + t];
    if (dw_bias_host) b[c] = dw_bias_host[c];
  }
  float *dw = nullptr, *db = nullptr;
  PV_CUDA(cudaMalloc((void**)&dw, w.size() * 4)); sc.ptrs.push_back(dw);
  PV_CUDA(cudaMalloc((void**)&db, b.size() * 4)); sc.ptrs.push_back(db);
  PV_CUDA(cudaMemcpyAsync(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice, st));
  PV_CUDA(cudaMemcpyAsync(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice, st));
  const long total = (long)batch * chunks * height * width;
  nchw_to_f8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x_dev, in.p, batch, channels, height, width);
  PV_TRY(after_launch("nchw_to_f8_kernel", st));
  ConvWeightsUmma pw;
  PV_TRY(pack_conv_weights_umma(&pw, pw_w_host, pw_bias_host, cout, channels, 1, 1, nullptr, 0, 4, 0, false));
  ConvOut o;
  o.cp = out;
  SepConvPlan plan;
  int r = plan_sepconv_fused(&plan, in, dw, db, cpad, relu_in != 0, relu_mid != 0, pw, o, slope);
  if (r == 0) r = launch_sepconv_fused(plan, batch, st);
  if (const char* rep = getenv("PREMVOS_CONV_REPEAT"))
    for (int i = 0; i < atoi(rep) && r == 0; i++) r = launch_sepconv_fused(plan, batch, st);
  if (r == 0) r = cp8_to_nchw(out, 0, out_dev, st);
  cudaError_t e = cudaStreamSynchronize(st);
  free_conv_weights_umma(&pw);
  if (r == 0 && e != cudaSuccess) r = fail((int)e, "premvos_sepconv2d_forward: %s", cudaGetErrorString(e));
  return r;
}
