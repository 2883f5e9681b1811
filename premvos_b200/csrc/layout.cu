// Layout / storage-format conversion between the reference's NCHW fp32 tensors and the library's
// channels-last views (fp32 or split bf16).  Used at the C-ABI boundary of the single-op entry points.
#include "common.cuh"

namespace premvos {

struct LayoutArgs {
  float* nchw; float* p; __nv_bfloat16* hi; __nv_bfloat16* lo;
  int N, C, H, W, cs, coff;
};

__global__ void __launch_bounds__(256) nchw_to_view_kernel(LayoutArgs a) {
  // one thread per (pixel, channel); threads of a warp walk x for coalesced NCHW reads
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)a.N * a.C * a.H * a.W;
  if (idx >= total) return;
  int x = (int)(idx % a.W);
  int y = (int)((idx / a.W) % a.H);
  int c = (int)((idx / ((long)a.W * a.H)) % a.C);
  int n = (int)(idx / ((long)a.W * a.H * a.C));
  float v = a.nchw[idx];
  long o = (((long)n * a.H + y) * a.W + x) * a.cs + a.coff + c;
  if (a.hi) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    a.hi[o] = h;
    a.lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
  } else {
    a.p[o] = v;
  }
}

__global__ void __launch_bounds__(256) view_to_nchw_kernel(LayoutArgs a) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long total = (long)a.N * a.C * a.H * a.W;
  if (idx >= total) return;
  int x = (int)(idx % a.W);
  int y = (int)((idx / a.W) % a.H);
  int c = (int)((idx / ((long)a.W * a.H)) % a.C);
  int n = (int)(idx / ((long)a.W * a.H * a.C));
  long o = (((long)n * a.H + y) * a.W + x) * a.cs + a.coff + c;
  a.nchw[idx] = a.hi ? (__bfloat162float(a.hi[o]) + __bfloat162float(a.lo[o])) : a.p[o];
}

static LayoutArgs make_args(const float* nchw, const TView& v) {
  LayoutArgs a;
  a.nchw = const_cast<float*>(nchw); a.p = v.p; a.hi = v.hi; a.lo = v.lo;
  a.N = v.N; a.C = v.C; a.H = v.H; a.W = v.W; a.cs = v.cs; a.coff = v.coff;
  return a;
}

int nchw_to_view(const float* src, const TView& dst, cudaStream_t st) {
  long total = (long)dst.N * dst.C * dst.H * dst.W;
  prof_before(st);
  nchw_to_view_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(make_args(src, dst));
  return after_launch("nchw_to_view_kernel", st, 0.0, 8.0 * total);
}

int view_to_nchw(const TView& src, float* dst, cudaStream_t st) {
  long total = (long)src.N * src.C * src.H * src.W;
  prof_before(st);
  view_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(make_args(dst, src));
  return after_launch("view_to_nchw_kernel", st, 0.0, 8.0 * total);
}

}  // namespace premvos
