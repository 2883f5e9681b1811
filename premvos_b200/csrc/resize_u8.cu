// cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR) for uint8 images on the device, bit-exact with OpenCV.
//
// Replaces the host-side resizes of the reference's stage drivers: script_pwc_multi.py:38-45 (both frames of a pair to
// multiples of 64) and proposal_net eval.py:75-78 / common.py:35-62 (CustomResize -> tensorpack ResizeTransform -> cv2.resize).
// OpenCV's 8-bit linear resize is integer arithmetic on 11-bit fixed-point coefficients (modules/imgproc/src/resize.cpp,
// resizeGeneric_ + HResizeLinear / VResizeLinear): the coefficient tables are built on the host with the same float
// operations OpenCV uses (cached per geometry on the device), the kernel does the two integer passes per output pixel:
//     r(y) = S[y][sx]*a0 + S[y][sx+1]*a1
//     out  = (((b0 * (r(sy0) >> 4)) >> 16) + ((b1 * (r(sy1) >> 4)) >> 16) + 2) >> 2
// HBM-bound byte work: one thread per output pixel (all channels), reads 4 source pixels, coalesced row-major stores.
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace premvos {
namespace {

struct ResizeTables {   // device: per destination column {sx0, sx1, a0, a1}, per destination row {sy0, sy1, b0, b1}
  int4* cols = nullptr;
  int4* rows = nullptr;
};

// OpenCV: fx = (float)((dx + 0.5) * scale - 0.5); sx = cvFloor(fx); fx -= sx; (x only:) clamp to the image with fx = 0;
// coefficient = saturate_cast<short>(c * INTER_RESIZE_COEF_SCALE) = cvRound of the float product (round half to even)
void build_axis(int dst_n, int src_n, bool clamp_fraction, std::vector<int4>* out) {
  const double scale = 1.0 / ((double)dst_n / (double)src_n);
  out->resize(dst_n);
  for (int d = 0; d < dst_n; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_fraction) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
    }
    const int c0 = (int)lrintf((1.f - f) * 2048.f), c1 = (int)lrintf(f * 2048.f);
    int s0 = s, s1 = s + 1;
    if (clamp_fraction) {
      s1 = s1 < src_n - 1 ? s1 : src_n - 1;
    } else {   // rows: indices clipped, fraction kept
      s0 = s0 < 0 ? 0 : (s0 > src_n - 1 ? src_n - 1 : s0);
      s1 = s1 < 0 ? 0 : (s1 > src_n - 1 ? src_n - 1 : s1);
    }
    (*out)[d] = make_int4(s0, s1, c0, c1);
  }
}

std::mutex g_mu;
std::map<std::tuple<int, int, int, int, int>, ResizeTables> g_tables;   // (device, sh, sw, dh, dw)

int get_tables(int sh, int sw, int dh, int dw, ResizeTables* out) {
  int dev = 0;
  PV_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_mu);
  auto key = std::make_tuple(dev, sh, sw, dh, dw);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) { *out = it->second; return 0; }
  std::vector<int4> cols, rows;
  build_axis(dw, sw, true, &cols);
  build_axis(dh, sh, false, &rows);
  ResizeTables t;
  PV_CUDA(cudaMalloc((void**)&t.cols, cols.size() * sizeof(int4)));
  PV_CUDA(cudaMalloc((void**)&t.rows, rows.size() * sizeof(int4)));
  PV_CUDA(cudaMemcpy(t.cols, cols.data(), cols.size() * sizeof(int4), cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(t.rows, rows.data(), rows.size() * sizeof(int4), cudaMemcpyHostToDevice));
  g_tables[key] = t;
  *out = t;
  return 0;
}

template <int C>
__global__ void __launch_bounds__(256) resize_linear_u8_kernel(const unsigned char* __restrict__ src, int sw, long src_img, unsigned char* __restrict__ dst,
                                                              int dh, int dw, long dst_img, const int4* __restrict__ cols,
                                                              const int4* __restrict__ rows, int reverse) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (x >= dw) return;
  const int4 cx = __ldg(cols + x), cy = __ldg(rows + y);
  const unsigned char* s0 = src + n * src_img + (long)cy.x * sw * C;
  const unsigned char* s1 = src + n * src_img + (long)cy.y * sw * C;
  unsigned char* o = dst + n * dst_img + ((long)y * dw + x) * C;
#pragma unroll
  for (int c = 0; c < C; c++) {
    const int r0 = (int)s0[cx.x * C + c] * cx.z + (int)s0[cx.y * C + c] * cx.w;
    const int r1 = (int)s1[cx.x * C + c] * cx.z + (int)s1[cx.y * C + c] * cx.w;
    const int v = (((cy.z * (r0 >> 4)) >> 16) + ((cy.w * (r1 >> 4)) >> 16) + 2) >> 2;
    o[reverse ? C - 1 - c : c] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

}  // namespace
}  // namespace premvos

using namespace premvos;

extern "C" int premvos_resize_linear_u8(const unsigned char* src_dev, int batch, int src_h, int src_w, unsigned char* dst_dev, int dst_h,
                                        int dst_w, int channels, int reverse_channels, void* stream) {
  PV_CHECK(src_dev && dst_dev, PREMVOS_ERR_INVALID_ARG, "premvos_resize_linear_u8: null argument");
  PV_CHECK(batch > 0 && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0 && dst_h <= 65535 && batch <= 65535, PREMVOS_ERR_INVALID_ARG,
           "premvos_resize_linear_u8: bad sizes");
  PV_CHECK(channels == 1 || channels == 3, PREMVOS_ERR_UNSUPPORTED, "premvos_resize_linear_u8: 1 or 3 channels (got %d)", channels);
  // OpenCV silently switches INTER_LINEAR to INTER_AREA for an exact 2x down-scale in both directions
  PV_CHECK(!(src_w == 2 * dst_w && src_h == 2 * dst_h), PREMVOS_ERR_UNSUPPORTED,
           "premvos_resize_linear_u8: exact 2x down-scaling is INTER_AREA in OpenCV, not implemented");
  ResizeTables t;
  PV_TRY(get_tables(src_h, src_w, dst_h, dst_w, &t));
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((dst_w + 255) / 256, dst_h, batch);
  const long src_img = (long)src_h * src_w * channels, dst_img = (long)dst_h * dst_w * channels;
  prof_before(st);
  if (channels == 3)
    resize_linear_u8_kernel<3><<<grid, 256, 0, st>>>(src_dev, src_w, src_img, dst_dev, dst_h, dst_w, dst_img, t.cols, t.rows, reverse_channels);
  else
    resize_linear_u8_kernel<1><<<grid, 256, 0, st>>>(src_dev, src_w, src_img, dst_dev, dst_h, dst_w, dst_img, t.cols, t.rows, 0);
  return after_launch("resize_linear_u8_kernel", st, 0.0, (double)batch * (dst_img + 4.0 * dst_img));
}

// Builds (and uploads, synchronously) the fixed-point coefficient tables of one geometry on the current device.  The enqueue-only
// entry point builds them on first use too, but that first call allocates and copies synchronously and therefore must not happen
// inside a stream capture: call this once per geometry beforehand.
extern "C" int premvos_resize_linear_u8_prepare(int src_h, int src_w, int dst_h, int dst_w) {
  PV_CHECK(src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, PREMVOS_ERR_INVALID_ARG, "premvos_resize_linear_u8_prepare: bad sizes");
  premvos::ResizeTables t;
  return premvos::get_tables(src_h, src_w, dst_h, dst_w, &t);
}

// Frees the coefficient tables of premvos_resize_linear_u8 (all devices, all geometries); they are rebuilt on demand.
extern "C" void premvos_resize_linear_u8_release(void) {
  std::lock_guard<std::mutex> lock(premvos::g_mu);
  for (auto& kv : premvos::g_tables) { cudaFree(kv.second.cols); cudaFree(kv.second.rows); }
  premvos::g_tables.clear();
}
