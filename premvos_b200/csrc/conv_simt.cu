// fp32 SIMT convolutions on channels-last activations.
//
//  * conv_simt_kernel     : implicit GEMM (M = output pixels, N = Cout, K = taps x Cin), 4x4 register
//                           tile per thread, register-staged double buffering through shared memory.
//                           Used for the PWC feature pyramid (Cin 3..196, HBM/L2 bound at levels 1-2)
//                           and as the full-fp32 mode of every other 3x3 convolution.
//  * conv3x3_small_kernel : Cout <= 4 (flow heads), one warp per output pixel, float4 channel loads,
//                           warp-shuffle reduction.
//  * deconv4x4_kernel     : ConvTranspose2d(k=4,s=2,p=1) with 2 output channels, gather form.
//
// Reference semantics: torch Conv2d / ConvTranspose2d as used in models/PWCNet.py:24-34.
#include <vector>

#include "common.cuh"

namespace premvos {

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
int pack_conv_weights_simt(ConvWeightsSimt* out, const float* host_w, const float* host_b, int Cout,
                           int Cin, int R, int S) {
  out->R = R; out->S = S; out->Cin = Cin; out->Cout = Cout;
  out->CinP = round_up(Cin, 16);
  out->CoutP = round_up(Cout, 64);
  std::vector<float> w((size_t)R * S * out->CinP * out->CoutP, 0.f), b(out->CoutP, 0.f);
  for (int co = 0; co < Cout; co++) {
    b[co] = host_b ? host_b[co] : 0.f;
    for (int ci = 0; ci < Cin; ci++)
      for (int r = 0; r < R; r++)
        for (int s = 0; s < S; s++)
          w[((size_t)(r * S + s) * out->CinP + ci) * out->CoutP + co] =
              host_w[(((size_t)co * Cin + ci) * R + r) * S + s];
  }
  PV_CUDA(cudaMalloc(&out->w, w.size() * sizeof(float)));
  PV_CUDA(cudaMalloc(&out->bias, b.size() * sizeof(float)));
  PV_CUDA(cudaMemcpy(out->w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(out->bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}
void free_conv_weights_simt(ConvWeightsSimt* w) {
  cudaFree(w->w); cudaFree(w->bias); w->w = w->bias = nullptr;
}

int pack_small_conv_weights(SmallConvWeights* out, const float* host_w, const float* host_b, int Cout, int Cin) {
  out->Cin = Cin; out->Cout = Cout; out->CinP = round_up(Cin, 4);
  std::vector<float> w((size_t)9 * Cout * out->CinP, 0.f);
  for (int co = 0; co < Cout; co++)
    for (int ci = 0; ci < Cin; ci++)
      for (int t = 0; t < 9; t++)
        w[((size_t)t * Cout + co) * out->CinP + ci] = host_w[((size_t)co * Cin + ci) * 9 + t];
  PV_CUDA(cudaMalloc(&out->w, w.size() * sizeof(float)));
  PV_CUDA(cudaMalloc(&out->bias, Cout * sizeof(float)));
  PV_CUDA(cudaMemcpy(out->w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(out->bias, host_b, Cout * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}
void free_small_conv_weights(SmallConvWeights* w) {
  cudaFree(w->w); cudaFree(w->bias); w->w = w->bias = nullptr;
}

int pack_deconv_weights(DeconvWeights* out, const float* host_w, const float* host_b, int Cin) {
  out->Cin = Cin; out->CinP = round_up(Cin, 4);
  std::vector<float> w((size_t)16 * 2 * out->CinP, 0.f);
  for (int ci = 0; ci < Cin; ci++)
    for (int co = 0; co < 2; co++)
      for (int t = 0; t < 16; t++)
        w[((size_t)t * 2 + co) * out->CinP + ci] = host_w[((size_t)ci * 2 + co) * 16 + t];
  PV_CUDA(cudaMalloc(&out->w, w.size() * sizeof(float)));
  PV_CUDA(cudaMalloc(&out->bias, 2 * sizeof(float)));
  PV_CUDA(cudaMemcpy(out->w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(out->bias, host_b, 2 * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}
void free_deconv_weights(DeconvWeights* w) {
  cudaFree(w->w); cudaFree(w->bias); w->w = w->bias = nullptr;
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM convolution
// ------------------------------------------------------------------------------------------------
struct ConvArgs {
  const float* in; int in_cs, in_coff, Cin;
  int N, H, W;
  const float* w; const float* bias;
  float* out; int out_cs, out_coff, Cout;
  int Ho, Wo, stride, pad, dil, R, S, CinP, CoutP;
  float slope;
};

constexpr int BK = 16;

template <int BM, int BN>
__global__ void __launch_bounds__((BM / 4) * (BN / 4)) conv_simt_kernel(ConvArgs a) {
  constexpr int T = (BM / 4) * (BN / 4);
  constexpr int A_LD = (BM * 4) / T;      // float4 loads of the A tile per thread (>=1)
  constexpr int B_F4 = 4 * BN;            // float4 count of the B tile
  static_assert(A_LD >= 1, "tile config");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int M = a.N * a.Ho * a.Wo;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // per-thread A-load rows
  int a_iy0[A_LD], a_ix0[A_LD];
  long a_base[A_LD];
  bool a_ok[A_LD];
#pragma unroll
  for (int j = 0; j < A_LD; j++) {
    int idx = tid + j * T;
    int m = m0 + idx / 4;
    a_ok[j] = m < M;
    int mm = a_ok[j] ? m : 0;
    int n = mm / (a.Ho * a.Wo);
    int r = mm - n * a.Ho * a.Wo;
    int oy = r / a.Wo, ox = r - oy * a.Wo;
    a_iy0[j] = oy * a.stride - a.pad;
    a_ix0[j] = ox * a.stride - a.pad;
    a_base[j] = (long)n * a.H * a.W;
  }
  const int kq = tid % 4;  // which float4 of the 16-channel chunk this thread loads

  const int cchunks = a.CinP / BK;
  const int KT = a.R * a.S * cchunks;

  float4 ra[A_LD];
  float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);

  auto load_tiles = [&](int kt) {
    int tap = kt / cchunks;
    int c0 = (kt - tap * cchunks) * BK;
    int r = tap / a.S, s = tap - r * a.S;
    int c = c0 + kq * 4;
#pragma unroll
    for (int j = 0; j < A_LD; j++) {
      int iy = a_iy0[j] + r * a.dil, ix = a_ix0[j] + s * a.dil;
      bool ok = a_ok[j] && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W && c < a.Cin;
      if (ok) {
        const float* p = a.in + ((a_base[j] + (long)iy * a.W + ix) * a.in_cs + a.in_coff + c);
        ra[j] = *reinterpret_cast<const float4*>(p);
      } else {
        ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (tid < B_F4) {
      int kk = tid / (BN / 4), nq = tid % (BN / 4);
      const float* p = a.w + ((size_t)(tap * a.CinP + c0 + kk) * a.CoutP + n0 + nq * 4);
      rb = *reinterpret_cast<const float4*>(p);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_LD; j++) {
      int idx = tid + j * T;
      int m = idx / 4;
      As[buf][kq * 4 + 0][m] = ra[j].x;
      As[buf][kq * 4 + 1][m] = ra[j].y;
      As[buf][kq * 4 + 2][m] = ra[j].z;
      As[buf][kq * 4 + 3][m] = ra[j].w;
    }
    if (tid < B_F4) {
      int kk = tid / (BN / 4), nq = tid % (BN / 4);
      *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb;
    }
  };

  const int tx = tid % (BN / 4), ty = tid / (BN / 4);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < KT; kt++) {
    int buf = kt & 1;
    if (kt + 1 < KT) load_tiles(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float aa[4] = {av.x, av.y, av.z, av.w};
      float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // epilogue: bias + LeakyReLU, channels-last store
  const int co0 = n0 + tx * 4;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; j++) bv[j] = a.bias[co0 + j];  // bias is padded to CoutP
  const bool vec_ok = (co0 + 3 < a.Cout) && ((a.out_cs & 3) == 0) && (((a.out_coff + co0) & 3) == 0);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float t = acc[i][j] + bv[j];
      v[j] = t > 0.f ? t : t * a.slope;
    }
    float* p = a.out + (size_t)m * a.out_cs + a.out_coff + co0;
    if (vec_ok) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (co0 + j < a.Cout) p[j] = v[j];
    }
  }
}

int conv2d_simt(const TView& in, const TView& out, const ConvWeightsSimt& w, int stride, int dil,
                float slope, cudaStream_t st) {
  PV_CHECK(!in.split() && !out.split(), PREMVOS_ERR_INVALID_ARG, "conv2d_simt: fp32 views only");
  PV_CHECK(in.C == w.Cin && out.C == w.Cout, PREMVOS_ERR_INVALID_ARG, "conv2d_simt: channel mismatch (%d,%d) vs (%d,%d)",
           in.C, out.C, w.Cin, w.Cout);
  PV_CHECK((in.cs % 4) == 0 && (in.coff % 4) == 0 && in.coff + round_up(in.C, 4) <= in.cs, PREMVOS_ERR_INVALID_ARG,
           "conv2d_simt: input view not float4-addressable (cs=%d coff=%d C=%d)", in.cs, in.coff, in.C);
  int pad = dil * (w.R / 2);
  int Ho = (in.H + 2 * pad - dil * (w.R - 1) - 1) / stride + 1;
  int Wo = (in.W + 2 * pad - dil * (w.S - 1) - 1) / stride + 1;
  PV_CHECK(Ho == out.H && Wo == out.W && in.N == out.N, PREMVOS_ERR_INVALID_ARG, "conv2d_simt: output is %dx%d, expected %dx%d",
           out.H, out.W, Ho, Wo);
  ConvArgs a;
  a.in = in.p; a.in_cs = in.cs; a.in_coff = in.coff; a.Cin = in.C;
  a.N = in.N; a.H = in.H; a.W = in.W;
  a.w = w.w; a.bias = w.bias;
  a.out = out.p; a.out_cs = out.cs; a.out_coff = out.coff; a.Cout = out.C;
  a.Ho = Ho; a.Wo = Wo; a.stride = stride; a.pad = pad; a.dil = dil; a.R = w.R; a.S = w.S;
  a.CinP = w.CinP; a.CoutP = w.CoutP; a.slope = slope;
  long M = (long)in.N * Ho * Wo;
  const double flops = 2.0 * M * w.Cout * w.Cin * w.R * w.S;
  const double bytes = 4.0 * ((double)in.pixels() * in.C + (double)M * w.Cout + (double)w.R * w.S * w.Cin * w.Cout);
  prof_before(st);
  if (w.Cout <= 16) {
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)((w.Cout + 15) / 16));
    conv_simt_kernel<256, 16><<<grid, 256, 0, st>>>(a);
  } else if (w.Cout <= 32) {
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)((w.Cout + 31) / 32));
    conv_simt_kernel<128, 32><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((w.Cout + 63) / 64));
    conv_simt_kernel<64, 64><<<grid, 256, 0, st>>>(a);
  }
  return after_launch("conv_simt_kernel", st, flops, bytes);
}

// ------------------------------------------------------------------------------------------------
// Cout <= 4, 3x3, pad 1: one warp per output pixel
// ------------------------------------------------------------------------------------------------
struct SmallConvArgs {
  const float* in; const __nv_bfloat16 *in_hi, *in_lo; int in_cs, in_coff, Cin, CinP;
  int N, H, W;
  const float* w; const float* bias; int Cout;
  float* out; int out_cs, out_coff;
  const float* add; int add_cs, add_coff;
  float* nchw;
};

template <int COUT, bool IN_SPLIT>
__global__ void __launch_bounds__(256) conv3x3_small_kernel(SmallConvArgs a) {
  const int lane = threadIdx.x & 31;
  const long pix = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long P = (long)a.N * a.H * a.W;
  if (pix >= P) return;
  const int n = (int)(pix / ((long)a.H * a.W));
  const int rem = (int)(pix - (long)n * a.H * a.W);
  const int y = rem / a.W, x = rem - y * a.W;
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; c++) acc[c] = 0.f;
  for (int t = 0; t < 9; t++) {
    int iy = y + t / 3 - 1, ix = x + t % 3 - 1;
    if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W) continue;  // warp-uniform
    const long ibase = (((long)n * a.H + iy) * a.W + ix) * a.in_cs + a.in_coff;
    const float* wp = a.w + (size_t)t * COUT * a.CinP;
    for (int c = lane * 4; c < a.CinP; c += 128) {
      float4 v = IN_SPLIT ? ld4_split(a.in_hi, a.in_lo, ibase + c) : *reinterpret_cast<const float4*>(a.in + ibase + c);
#pragma unroll
      for (int co = 0; co < COUT; co++) {
        float4 wv = *reinterpret_cast<const float4*>(wp + (size_t)co * a.CinP + c);
        acc[co] = fmaf(v.x, wv.x, acc[co]);
        acc[co] = fmaf(v.y, wv.y, acc[co]);
        acc[co] = fmaf(v.z, wv.z, acc[co]);
        acc[co] = fmaf(v.w, wv.w, acc[co]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < COUT; co++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
  }
  if (lane < COUT) {
    float v = 0.f;
#pragma unroll
    for (int co = 0; co < COUT; co++)
      if (lane == co) v = acc[co];
    v += a.bias[lane];
    if (a.add) v += a.add[pix * a.add_cs + a.add_coff + lane];
    if (a.out) a.out[pix * a.out_cs + a.out_coff + lane] = v;
    if (a.nchw) a.nchw[(((long)n * COUT + lane) * a.H + y) * a.W + x] = v;
  }
}

int conv3x3_small_cout(const TView& in, const TView& out, const SmallConvWeights& w, const TView* addend,
                       float* nchw_out, cudaStream_t st) {
  PV_CHECK(in.C == w.Cin && w.Cout == 2, PREMVOS_ERR_INVALID_ARG, "conv3x3_small_cout: bad channels");
  PV_CHECK((in.cs % 4) == 0 && (in.coff % 4) == 0 && in.coff + w.CinP <= in.cs, PREMVOS_ERR_INVALID_ARG,
           "conv3x3_small_cout: input view not float4-addressable");
  SmallConvArgs a;
  a.in = in.p; a.in_hi = in.hi; a.in_lo = in.lo; a.in_cs = in.cs; a.in_coff = in.coff; a.Cin = in.C; a.CinP = w.CinP;
  a.N = in.N; a.H = in.H; a.W = in.W;
  a.w = w.w; a.bias = w.bias; a.Cout = w.Cout;
  PV_CHECK(!out.split() && (!addend || !addend->split()), PREMVOS_ERR_INVALID_ARG, "conv3x3_small_cout: fp32 outputs only");
  a.out = out.p; a.out_cs = out.cs; a.out_coff = out.coff;
  a.add = addend ? addend->p : nullptr;
  a.add_cs = addend ? addend->cs : 0;
  a.add_coff = addend ? addend->coff : 0;
  a.nchw = nchw_out;
  long P = (long)in.N * in.H * in.W;
  prof_before(st);
  if (in.split()) conv3x3_small_kernel<2, true><<<(unsigned)((P + 7) / 8), 256, 0, st>>>(a);
  else conv3x3_small_kernel<2, false><<<(unsigned)((P + 7) / 8), 256, 0, st>>>(a);
  return after_launch("conv3x3_small_kernel", st, 2.0 * P * 9 * w.Cin * 2, 4.0 * P * (w.Cin + 2));
}

// ------------------------------------------------------------------------------------------------
// ConvTranspose2d 4x4 / stride 2 / pad 1 / Cout 2: out[oy,ox] gathers in[(oy+1-ky)/2,(ox+1-kx)/2]
// ------------------------------------------------------------------------------------------------
struct DeconvArgs {
  const float* in; const __nv_bfloat16 *in_hi, *in_lo; int in_cs, in_coff, CinP;
  int N, H, W;  // input size; output is 2H x 2W
  const float* w; const float* bias;
  float* out; __nv_bfloat16 *out_hi, *out_lo; int out_cs, out_coff;
};

template <bool IN_SPLIT, bool OUT_SPLIT>
__global__ void __launch_bounds__(256) deconv4x4_kernel(DeconvArgs a) {
  const int lane = threadIdx.x & 31;
  const int Ho = a.H * 2, Wo = a.W * 2;
  const long pix = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long P = (long)a.N * Ho * Wo;
  if (pix >= P) return;
  const int n = (int)(pix / ((long)Ho * Wo));
  const int rem = (int)(pix - (long)n * Ho * Wo);
  const int oy = rem / Wo, ox = rem - oy * Wo;
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int a_ = 0; a_ < 2; a_++) {
    int ky = ((oy + 1) & 1) + 2 * a_;
    int iy = (oy + 1 - ky) / 2;
    if (oy + 1 - ky < 0 || iy >= a.H) continue;
#pragma unroll
    for (int b_ = 0; b_ < 2; b_++) {
      int kx = ((ox + 1) & 1) + 2 * b_;
      int ix = (ox + 1 - kx) / 2;
      if (ox + 1 - kx < 0 || ix >= a.W) continue;
      const long ibase = (((long)n * a.H + iy) * a.W + ix) * a.in_cs + a.in_coff;
      const float* wp = a.w + (size_t)(ky * 4 + kx) * 2 * a.CinP;
      for (int c = lane * 4; c < a.CinP; c += 128) {
        float4 v = IN_SPLIT ? ld4_split(a.in_hi, a.in_lo, ibase + c) : *reinterpret_cast<const float4*>(a.in + ibase + c);
        float4 w0 = *reinterpret_cast<const float4*>(wp + c);
        float4 w1 = *reinterpret_cast<const float4*>(wp + a.CinP + c);
        acc0 = fmaf(v.x, w0.x, acc0); acc0 = fmaf(v.y, w0.y, acc0);
        acc0 = fmaf(v.z, w0.z, acc0); acc0 = fmaf(v.w, w0.w, acc0);
        acc1 = fmaf(v.x, w1.x, acc1); acc1 = fmaf(v.y, w1.y, acc1);
        acc1 = fmaf(v.z, w1.z, acc1); acc1 = fmaf(v.w, w1.w, acc1);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
    acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
  }
  if (lane < 2) {
    float v = (lane == 0 ? acc0 : acc1) + a.bias[lane];
    if (OUT_SPLIT) st_split(a.out_hi, a.out_lo, pix * a.out_cs + a.out_coff + lane, v);
    else a.out[pix * a.out_cs + a.out_coff + lane] = v;
  }
}

int deconv4x4s2_cout2(const TView& in, const TView& out, const DeconvWeights& w, cudaStream_t st) {
  PV_CHECK(in.C == w.Cin && out.C == 2 && out.H == 2 * in.H && out.W == 2 * in.W && out.N == in.N,
           PREMVOS_ERR_INVALID_ARG, "deconv4x4s2_cout2: shape mismatch");
  PV_CHECK((in.coff % 4) == 0 && in.coff + w.CinP <= in.cs && (in.cs % 4) == 0, PREMVOS_ERR_INVALID_ARG,
           "deconv4x4s2_cout2: input view not addressable");
  DeconvArgs a;
  a.in = in.p; a.in_hi = in.hi; a.in_lo = in.lo; a.in_cs = in.cs; a.in_coff = in.coff; a.CinP = w.CinP;
  a.N = in.N; a.H = in.H; a.W = in.W;
  a.w = w.w; a.bias = w.bias;
  a.out = out.p; a.out_hi = out.hi; a.out_lo = out.lo; a.out_cs = out.cs; a.out_coff = out.coff;
  long P = (long)out.N * out.H * out.W;
  prof_before(st);
  const unsigned nb = (unsigned)((P + 7) / 8);
  if (in.split() && out.split()) deconv4x4_kernel<true, true><<<nb, 256, 0, st>>>(a);
  else if (!in.split() && out.split()) deconv4x4_kernel<false, true><<<nb, 256, 0, st>>>(a);
  else if (in.split() && !out.split()) deconv4x4_kernel<true, false><<<nb, 256, 0, st>>>(a);
  else deconv4x4_kernel<false, false><<<nb, 256, 0, st>>>(a);
  return after_launch("deconv4x4_kernel", st, 2.0 * P * 4 * w.Cin * 2, 4.0 * ((double)in.pixels() * w.Cin + 2.0 * P));
}

}  // namespace premvos
