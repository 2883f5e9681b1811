// Cost-volume correlation (FlowNet-C / PWC-Net), forward, "multiply" variant.
//
// Reference: correlation_package/src/corr_cuda.c:7-82 + corr_cuda_kernel.cu:18-127 -- three memsets,
// two NCHW->padded-NHWC transposes with 16-thread blocks, then one 32-thread block per output pixel
// that re-reads a 9x9xC neighbourhood from global memory and reduces through shared memory with a
// serial 32-add.  Here: no scratch tensors, no transposes.  One CTA owns a 4x32 pixel tile; per
// 32-channel chunk the f1 tile and the f2 tile + 4-pixel halo are staged in shared memory
// (zero padding applied while staging), pitch 33 floats so that 32 lanes = 32 consecutive pixels hit
// 32 different banks.  384 threads = 128 pixels x 3 groups of 27 displacements held in registers;
// f2 is read from HBM exactly once per tile (+halo), f1 once.  HBM-bound by design:
// algorithmic bytes = 4*(2*C + 81)*H*W per image pair.
//
//   corr81_kernel<NHWC=true>  : channels-last in / out (the PWC pipeline; output goes straight
//                               into the decoder slab with LeakyReLU fused, optional c1 copy)
//   corr81_kernel<NHWC=false> : NCHW in / out (the C ABI drop-in for corr_cuda_forward)
//   corr_generic_kernel       : any pad/kernel_size/stride1/stride2/max_displacement (NCHW), slow path
#include <stdlib.h>

#include "common.cuh"

namespace premvos {

constexpr int CT_H = 4, CT_W = 32, MDISP = 4, DW = 9;
constexpr int HALO_H = CT_H + 2 * MDISP, HALO_W = CT_W + 2 * MDISP;  // 12 x 40
constexpr int CK = 32, PITCH = CK + 1;
constexpr int CORR_THREADS = 384;
constexpr size_t CORR_SMEM = (size_t)(HALO_H * HALO_W + CT_H * CT_W) * PITCH * sizeof(float);

struct CorrArgs {
  const float* f1; const float* f2; float* out; float* c1;
  __nv_bfloat16 *out_hi, *out_lo, *c1_hi, *c1_lo;   // split-bf16 destinations (decoder slab in tensor-core mode)
  // channels-last: element (n,y,x,c) at ((n*H+y)*W+x)*cs + coff + c ; NCHW: ((n*C+c)*H+y)*W+x
  int f1_cs, f1_coff, f2_cs, f2_coff, out_cs, out_coff, c1_cs, c1_coff;
  int B, C, H, W;
  float slope;
};

template <bool NHWC, bool OUT_SPLIT>
__global__ void __launch_bounds__(CORR_THREADS) corr81_kernel(CorrArgs a) {
  extern __shared__ float smem[];
  float* s2 = smem;                              // [HALO_H*HALO_W][PITCH]
  float* s1 = smem + HALO_H * HALO_W * PITCH;    // [CT_H*CT_W][PITCH]
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * CT_H, x0 = blockIdx.x * CT_W;
  const int p = tid & 127;         // pixel inside the tile
  const int grp = tid >> 7;        // displacement rows 3*grp .. 3*grp+2
  const int py = p >> 5, px = p & 31;

  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; i++) acc[i] = 0.f;

  for (int c0 = 0; c0 < a.C; c0 += CK) {
    // ---- stage f2 halo tile and f1 tile for channels [c0, c0+32) ----
    if (NHWC) {
      // one warp-iteration = 4 pixels x 8 float4
      for (int i = tid; i < HALO_H * HALO_W * (CK / 4); i += CORR_THREADS) {
        int hp = i >> 3, q = i & 7;
        int hy = hp / HALO_W, hx = hp - hy * HALO_W;
        int gy = y0 + hy - MDISP, gx = x0 + hx - MDISP;
        int c = c0 + q * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && c < a.C) {
          const float* src = a.f2 + (((long)n * a.H + gy) * a.W + gx) * a.f2_cs + a.f2_coff + c;
          if (c + 3 < a.C && ((a.f2_cs | a.f2_coff) & 3) == 0) {
            float4 t = *reinterpret_cast<const float4*>(src);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
            for (int k = 0; k < 4; k++) if (c + k < a.C) v[k] = src[k];
          }
        }
        float* d = s2 + hp * PITCH + q * 4;
        d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
      }
      for (int i = tid; i < CT_H * CT_W * (CK / 4); i += CORR_THREADS) {
        int tp = i >> 3, q = i & 7;
        int gy = y0 + (tp >> 5), gx = x0 + (tp & 31);
        int c = c0 + q * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        bool inb = gy < a.H && gx < a.W && c < a.C;
        if (inb) {
          const float* src = a.f1 + (((long)n * a.H + gy) * a.W + gx) * a.f1_cs + a.f1_coff + c;
          if (c + 3 < a.C && ((a.f1_cs | a.f1_coff) & 3) == 0) {
            float4 t = *reinterpret_cast<const float4*>(src);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
            for (int k = 0; k < 4; k++) if (c + k < a.C) v[k] = src[k];
          }
          if (OUT_SPLIT ? (a.c1_hi != nullptr) : (a.c1 != nullptr)) {  // fused copy of f1 into the decoder slab
            long di = (((long)n * a.H + gy) * a.W + gx) * a.c1_cs + a.c1_coff + c;
            for (int k = 0; k < 4; k++)
              if (c + k < a.C) {
                if (OUT_SPLIT) st_split(a.c1_hi, a.c1_lo, di + k, v[k]);
                else a.c1[di + k] = v[k];
              }
          }
        }
        float* d = s1 + tp * PITCH + q * 4;
        d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
      }
    } else {
      // NCHW: consecutive threads read consecutive x of one channel row
      for (int i = tid; i < CK * HALO_H * HALO_W; i += CORR_THREADS) {
        int cc = i / (HALO_H * HALO_W);
        int hp = i - cc * (HALO_H * HALO_W);
        int hy = hp / HALO_W, hx = hp - hy * HALO_W;
        int gy = y0 + hy - MDISP, gx = x0 + hx - MDISP;
        int c = c0 + cc;
        float v = 0.f;
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && c < a.C)
          v = a.f2[(((long)n * a.C + c) * a.H + gy) * a.W + gx];
        s2[hp * PITCH + cc] = v;
      }
      for (int i = tid; i < CK * CT_H * CT_W; i += CORR_THREADS) {
        int cc = i >> 7, tp = i & 127;
        int gy = y0 + (tp >> 5), gx = x0 + (tp & 31);
        int c = c0 + cc;
        float v = 0.f;
        if (gy < a.H && gx < a.W && c < a.C) v = a.f1[(((long)n * a.C + c) * a.H + gy) * a.W + gx];
        s1[tp * PITCH + cc] = v;
      }
    }
    __syncthreads();
    // ---- 27 displacements x 32 channels per thread ----
    const float* q1 = s1 + p * PITCH;
    const float* q2 = s2 + ((py + 3 * grp) * HALO_W + px) * PITCH;
#pragma unroll 4
    for (int c = 0; c < CK; c++) {
      float f = q1[c];
#pragma unroll
      for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < DW; dx++)
          acc[dy * DW + dx] = fmaf(f, q2[(dy * HALO_W + dx) * PITCH + c], acc[dy * DW + dx]);
    }
    __syncthreads();
  }

  const int gy = y0 + py, gx = x0 + px;
  if (NHWC) {
    // stage the 128 x 81 results in shared memory, then write 81 contiguous floats per pixel
    float* so = smem;  // [128][81]
#pragma unroll
    for (int i = 0; i < 27; i++) {
      float v = acc[i] / (float)a.C;  // corr_cuda_kernel.cu:119-121
      v = v > 0.f ? v : v * a.slope;
      so[p * 81 + grp * 27 + i] = v;
    }
    __syncthreads();
    for (int i = tid; i < CT_H * CT_W * 81; i += CORR_THREADS) {
      int tp = i / 81, k = i - tp * 81;
      int yy = y0 + (tp >> 5), xx = x0 + (tp & 31);
      if (yy < a.H && xx < a.W) {
        long oi = (((long)n * a.H + yy) * a.W + xx) * a.out_cs + a.out_coff + k;
        if (OUT_SPLIT) st_split(a.out_hi, a.out_lo, oi, so[i]);
        else a.out[oi] = so[i];
      }
    }
  } else {
    if (gy < a.H && gx < a.W) {
#pragma unroll
      for (int i = 0; i < 27; i++) {
        float v = acc[i] / (float)a.C;
        v = v > 0.f ? v : v * a.slope;
        a.out[(((long)n * 81 + grp * 27 + i) * a.H + gy) * a.W + gx] = v;
      }
    }
  }
}

int corr81_nhwc(const TView& f1, const TView& f2, const TView& out, const TView& c1_copy, float slope,
                cudaStream_t st) {
  PV_CHECK(f1.C == f2.C && f1.H == f2.H && f1.W == f2.W && f1.N == f2.N && out.C == 81 && out.H == f1.H &&
               out.W == f1.W && out.N == f1.N, PREMVOS_ERR_INVALID_ARG, "corr81_nhwc: shape mismatch");
  PV_CHECK(!f1.split() && !f2.split() && (c1_copy.null() || c1_copy.split() == out.split()), PREMVOS_ERR_INVALID_ARG,
           "corr81_nhwc: features must be fp32 views; c1 copy must use the output's storage format");
  static bool attr_set = false;
  if (!attr_set) {
    PV_CUDA(cudaFuncSetAttribute(corr81_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CORR_SMEM));
    PV_CUDA(cudaFuncSetAttribute(corr81_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CORR_SMEM));
    attr_set = true;
  }
  CorrArgs a;
  a.f1 = f1.p; a.f2 = f2.p; a.out = out.p; a.c1 = c1_copy.p;
  a.out_hi = out.hi; a.out_lo = out.lo; a.c1_hi = c1_copy.hi; a.c1_lo = c1_copy.lo;
  a.f1_cs = f1.cs; a.f1_coff = f1.coff; a.f2_cs = f2.cs; a.f2_coff = f2.coff;
  a.out_cs = out.cs; a.out_coff = out.coff; a.c1_cs = c1_copy.cs; a.c1_coff = c1_copy.coff;
  a.B = f1.N; a.C = f1.C; a.H = f1.H; a.W = f1.W; a.slope = slope;
  dim3 grid((f1.W + CT_W - 1) / CT_W, (f1.H + CT_H - 1) / CT_H, f1.N);
  const double px = (double)f1.pixels();
  prof_before(st);
  if (out.split()) corr81_kernel<true, true><<<grid, CORR_THREADS, CORR_SMEM, st>>>(a);
  else corr81_kernel<true, false><<<grid, CORR_THREADS, CORR_SMEM, st>>>(a);
  return after_launch("corr81_kernel<nhwc>", st, 2.0 * 81 * f1.C * px, 4.0 * (2.0 * f1.C + 81) * px);
}

// ------------------------------------------------------------------------------------------------
// generic NCHW path: one thread per output element (any configuration of the reference op)
// ------------------------------------------------------------------------------------------------
struct CorrGenericArgs {
  const float* in1; const float* in2; float* out;
  int B, C, H, W, pad, ksize, md, s1, s2, OC, OH, OW, gr, gw;
};

__global__ void corr_generic_kernel(CorrGenericArgs a) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)a.B * a.OC * a.OH * a.OW;
  if (idx >= total) return;
  int ox = (int)(idx % a.OW);
  int oy = (int)((idx / a.OW) % a.OH);
  int tc = (int)((idx / ((long)a.OW * a.OH)) % a.OC);
  int n = (int)(idx / ((long)a.OW * a.OH * a.OC));
  int s2o = (tc % a.gw - a.gr) * a.s2;
  int s2p = (tc / a.gw - a.gr) * a.s2;
  // padded coordinates -> image coordinates: subtract pad
  int x1 = ox * a.s1 + a.md - a.pad, y1 = oy * a.s1 + a.md - a.pad;
  float sum = 0.f;
  for (int j = 0; j < a.ksize; j++)
    for (int i = 0; i < a.ksize; i++) {
      int ya = y1 + j, xa = x1 + i, yb = y1 + s2p + j, xb = x1 + s2o + i;
      if (ya < 0 || ya >= a.H || xa < 0 || xa >= a.W || yb < 0 || yb >= a.H || xb < 0 || xb >= a.W) continue;
      const float* pa = a.in1 + ((long)n * a.C * a.H + ya) * a.W + xa;
      const float* pb = a.in2 + ((long)n * a.C * a.H + yb) * a.W + xb;
      long cst = (long)a.H * a.W;
      for (int c = 0; c < a.C; c++) sum = fmaf(pa[c * cst], pb[c * cst], sum);
    }
  a.out[idx] = sum / (float)(a.ksize * a.ksize * a.C);
}

}  // namespace premvos

using namespace premvos;

extern "C" int premvos_corr_output_shape(int height, int width, int pad_size, int kernel_size, int max_displacement,
                                         int stride1, int stride2, int* out_channels, int* out_height, int* out_width) {
  PV_CHECK(height > 0 && width > 0 && pad_size >= 0 && kernel_size >= 1 && (kernel_size & 1) && max_displacement >= 0 &&
               stride1 >= 1 && stride2 >= 1, PREMVOS_ERR_INVALID_ARG, "premvos_corr_output_shape: invalid argument");
  int kr = (kernel_size - 1) / 2;
  int border = max_displacement + kr;
  int ph = height + 2 * pad_size, pw = width + 2 * pad_size;
  int ow = (int)ceilf((float)(pw - border * 2) / (float)stride1);
  int oh = (int)ceilf((float)(ph - border * 2) / (float)stride1);
  int gr = max_displacement / stride2, gw = 2 * gr + 1;
  PV_CHECK(ow > 0 && oh > 0, PREMVOS_ERR_INVALID_ARG, "premvos_corr_output_shape: empty output (%dx%d)", oh, ow);
  if (out_channels) *out_channels = gw * gw;
  if (out_height) *out_height = oh;
  if (out_width) *out_width = ow;
  return 0;
}

extern "C" int premvos_corr_forward(const float* input1, const float* input2, float* output, int batch, int channels,
                                    int height, int width, int pad_size, int kernel_size, int max_displacement,
                                    int stride1, int stride2, int corr_type_multiply, void* stream) {
  PV_CHECK(input1 && input2 && output, PREMVOS_ERR_INVALID_ARG, "premvos_corr_forward: null pointer");
  PV_CHECK(batch > 0 && channels > 0, PREMVOS_ERR_INVALID_ARG, "premvos_corr_forward: batch/channels must be positive");
  PV_CHECK(corr_type_multiply == 1, PREMVOS_ERR_UNSUPPORTED,
           "premvos_corr_forward: only corr_type_multiply=1 is on the hot path (PWCNet.py:69)");
  int oc, oh, ow;
  PV_TRY(premvos_corr_output_shape(height, width, pad_size, kernel_size, max_displacement, stride1, stride2, &oc, &oh, &ow));
  cudaStream_t st = (cudaStream_t)stream;
  if (pad_size == MDISP && max_displacement == MDISP && kernel_size == 1 && stride1 == 1 && stride2 == 1) {
    // PWC-Net's configuration: TMA-staged kernel (corr_tma.cu) when the rows are 16-byte multiples, else the per-pixel kernel
    static const bool use_tma = !(getenv("PREMVOS_CORR_TMA") && atoi(getenv("PREMVOS_CORR_TMA")) == 0);
    if (use_tma) {
      const int r = corr81_nchw_tma(input1, input2, output, batch, channels, height, width, st);
      if (r != 1) return r;
    }
    static bool attr_set = false;
    if (!attr_set) {
      PV_CUDA(cudaFuncSetAttribute(corr81_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CORR_SMEM));
      attr_set = true;
    }
    CorrArgs a = {};
    a.f1 = input1; a.f2 = input2; a.out = output; a.c1 = nullptr;
    a.B = batch; a.C = channels; a.H = height; a.W = width; a.slope = 1.0f;
    dim3 grid((width + CT_W - 1) / CT_W, (height + CT_H - 1) / CT_H, batch);
    const double px = (double)batch * height * width;
    prof_before(st);
    corr81_kernel<false, false><<<grid, CORR_THREADS, CORR_SMEM, st>>>(a);
    return after_launch("corr81_kernel<nchw>", st, 2.0 * 81 * channels * px, 4.0 * (2.0 * channels + 81) * px);
  }
  CorrGenericArgs g;
  g.in1 = input1; g.in2 = input2; g.out = output;
  g.B = batch; g.C = channels; g.H = height; g.W = width; g.pad = pad_size; g.ksize = kernel_size;
  g.md = max_displacement; g.s1 = stride1; g.s2 = stride2; g.OC = oc; g.OH = oh; g.OW = ow;
  g.gr = max_displacement / stride2; g.gw = 2 * g.gr + 1;
  long total = (long)batch * oc * oh * ow;
  prof_before(st);
  corr_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g);
  return after_launch("corr_generic_kernel", st, 2.0 * total * channels * kernel_size * kernel_size,
                      4.0 * (2.0 * batch * channels * height * width + total));
}
