// MergeTrack's live mask propagation on the device and the flow post-processing that feeds it.
//
//   premvos_warp_masks_u8     MergeTrack/merge_functions.py:209-217 (warp_flow: cv2.remap INTER_LINEAR on map = grid - flow,
//                             then `== 1`) + :231 (toBbox of the warped mask), for all masks of a frame in one launch.
//                             BIT-EXACT with OpenCV: 8-bit remap is integer arithmetic -- the map is quantised to 1/32 pixel
//                             (cvRound(coord * 32)), the four taps are weighted with the 15-bit bilinear table, taps outside
//                             the image read 0, out = (sum + 16384) >> 15 (modules/imgproc/src/imgwarp.cpp).
//   premvos_flow_postprocess  optical_flow_net-PWC-Net/script_pwc_multi.py:59-68 (flow2 * 20 -> HWC -> cv2.resize of u and v
//                             to the frame size -> u *= W/W_, v *= H/H_): float32 cv2.resize INTER_LINEAR, same coordinate
//                             arithmetic as OpenCV (float coefficient tables built on the host).
//
// HBM-bound byte work: the map is computed once per pixel and reused for every mask; each thread owns 4 consecutive pixels
// (two 16-byte flow loads, one 4-byte store per mask), the per-mask bounding boxes are reduced with redux.sync per warp,
// shared-memory atomics per CTA and one set of global atomics per CTA and mask it touches.  Measured (tools/bench_aux.py, 40
// masks of 1080p, smooth flow): the kernel is ISSUE-bound, not latency-bound (batching the tap loads of 4 masks did not help) --
// hence weights and validity are computed once per pixel, and threads whose taps are all inside the image take a branch-free
// path (4 byte loads, 4 IMADs per pixel and mask).
#include <limits.h>
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace premvos {
namespace {

struct Tap {   // quantised source position and fixed-point weights of one destination pixel, the same for every mask of the frame
  int off;                  // iy * W + ix: element offset of the top-left tap inside a mask (0 when no tap is inside)
  int valid;                // bit 0..3: tap (iy,ix), (iy,ix+1), (iy+1,ix), (iy+1,ix+1) lies inside the image
  int w00, w01, w10, w11;   // 15-bit bilinear weights
};

__device__ __forceinline__ int cv_round_x32(float coord) {
  const float p = __fmul_rn(coord, 32.f);
  // x86 cvtss2si: NaN and out-of-range give INT_MIN (CUDA's cvt would saturate / give 0)
  if (!(p >= -2147483648.f && p < 2147483648.f)) return INT_MIN;
  return __float2int_rn(p);
}

__device__ __forceinline__ Tap make_tap(float fx, float fy, int x, int y, int H, int W) {
  const int sx = cv_round_x32(__fadd_rn(-fx, (float)x)), sy = cv_round_x32(__fadd_rn(-fy, (float)y));
  const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
  const int ax = sx & 31, ay = sy & 31;
  Tap t;
  t.w00 = min((32 - ay) * (32 - ax) * 32, 32767);   // saturate_cast<short>(1.0 * 32768)
  t.w01 = (32 - ay) * ax * 32;
  t.w10 = ay * (32 - ax) * 32;
  t.w11 = ay * ax * 32;
  const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
  const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
  t.valid = (y0 && x0 ? 1 : 0) | (y0 && x1 ? 2 : 0) | (y1 && x0 ? 4 : 0) | (y1 && x1 ? 8 : 0);
  t.off = t.valid ? iy * W + ix : 0;   // a valid tap means -1 <= iy < H, -1 <= ix < W: fits an int for H, W <= 32767
  return t;
}

// one pixel of one mask.  INTERIOR: all four taps are inside the image (the common case, decided once per thread for all masks)
template <bool INTERIOR>
__device__ __forceinline__ int remap_pixel(const unsigned char* __restrict__ m, int W, const Tap& t) {
  const unsigned char* r0 = m + t.off;
  const unsigned char* r1 = r0 + W;
  int acc = 16384;
  if (INTERIOR) {
    acc += t.w00 * (int)__ldg(r0) + t.w01 * (int)__ldg(r0 + 1) + t.w10 * (int)__ldg(r1) + t.w11 * (int)__ldg(r1 + 1);
  } else {   // taps outside read the constant border 0
    if (t.valid & 1) acc += t.w00 * (int)__ldg(r0);
    if (t.valid & 2) acc += t.w01 * (int)__ldg(r0 + 1);
    if (t.valid & 4) acc += t.w10 * (int)__ldg(r1);
    if (t.valid & 8) acc += t.w11 * (int)__ldg(r1 + 1);
  }
  return min(255, acc >> 15);
}

__global__ void bbox_init_kernel(int* __restrict__ bbox, int n, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<int4*>(bbox)[i] = make_int4(W, H, -1, -1);
}

// [xmin, ymin, xmax, ymax] (int) -> [x, y, w, h] (float), zeros for an empty mask (pycocotools rleToBbox)
__global__ void bbox_finish_kernel(int* __restrict__ bbox, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 b = reinterpret_cast<int4*>(bbox)[i];
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b.z >= 0) o = make_float4((float)b.x, (float)b.y, (float)(b.z - b.x + 1), (float)(b.w - b.y + 1));
  reinterpret_cast<float4*>(bbox)[i] = o;
}

constexpr int PX = 4;   // pixels per thread

template <bool VEC>
__global__ void __launch_bounds__(256) warp_masks_kernel(const unsigned char* __restrict__ masks, int n, int H, int W,
                                                        const float* __restrict__ flow, unsigned char* __restrict__ out,
                                                        int* __restrict__ bbox, int binarize) {
  extern __shared__ int cta_box[];   // [n][xmin, ymin, xmax, ymax] of this CTA's pixels (only with bbox)
  if (bbox) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<int4*>(cta_box)[i] = make_int4(INT_MAX, INT_MAX, -1, -1);
    __syncthreads();
  }
  const long HW = (long)H * W;
  const long p0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * PX;
  Tap tap[PX];
  int px[PX], py[PX];
  const bool active = p0 < HW;
  if (active) {
    float f[2 * PX];
    if (VEC) {   // HW % 4 == 0: the thread's 4 pixels are inside the image and 32-byte aligned
      const float4 a = __ldg(reinterpret_cast<const float4*>(flow + 2 * p0)), b = __ldg(reinterpret_cast<const float4*>(flow + 2 * p0) + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < PX; k++) {
        const long p = min(p0 + k, HW - 1);
        f[2 * k] = __ldg(flow + 2 * p);
        f[2 * k + 1] = __ldg(flow + 2 * p + 1);
      }
    }
    py[0] = (int)(p0 / W);
    px[0] = (int)(p0 - (long)py[0] * W);
#pragma unroll
    for (int k = 1; k < PX; k++) {
      px[k] = px[k - 1] + 1;
      py[k] = py[k - 1];
      if (px[k] == W) { px[k] = 0; py[k]++; }
    }
#pragma unroll
    for (int k = 0; k < PX; k++) tap[k] = make_tap(f[2 * k], f[2 * k + 1], px[k], py[k], H, W);
  }
  // interior threads (every tap of every pixel inside the image, all pixels real) take the branch-free path for all masks
  bool interior = active && (VEC || p0 + PX <= HW);
#pragma unroll
  for (int k = 0; k < PX; k++) interior = interior && tap[k].valid == 15;
  const bool one_row = active && py[0] == py[PX - 1];
  for (int i = 0; i < n; i++) {
    int xmin = INT_MAX, ymin = INT_MAX, xmax = -1, ymax = -1;
    if (active) {
      const unsigned char* m = masks + i * HW;
      unsigned char v[PX];
      unsigned set = 0;
      if (interior) {
#pragma unroll
        for (int k = 0; k < PX; k++) {
          const int r = remap_pixel<true>(m, W, tap[k]);
          v[k] = (unsigned char)(binarize ? (r == 1) : r);
          set |= (v[k] ? 1u : 0u) << k;
        }
      } else {
#pragma unroll
        for (int k = 0; k < PX; k++) {
          const int r = remap_pixel<false>(m, W, tap[k]);
          v[k] = (unsigned char)(binarize ? (r == 1) : r);
          set |= ((v[k] && (VEC || p0 + k < HW)) ? 1u : 0u) << k;
        }
      }
      if (set) {
        if (one_row) {   // consecutive pixels of one row: the first / last set pixel bound x
          xmin = px[0] + __ffs(set) - 1; xmax = px[0] + 31 - __clz(set);
          ymin = ymax = py[0];
        } else {
#pragma unroll
          for (int k = 0; k < PX; k++)
            if (set >> k & 1u) {
              xmin = min(xmin, px[k]); xmax = max(xmax, px[k]);
              ymin = min(ymin, py[k]); ymax = max(ymax, py[k]);
            }
        }
      }
      if (VEC) {
        *reinterpret_cast<uchar4*>(out + i * HW + p0) = make_uchar4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int k = 0; k < PX; k++)
          if (p0 + k < HW) out[i * HW + p0 + k] = v[k];
      }
    }
    if (bbox) {
      const int wxmax = __reduce_max_sync(0xffffffffu, xmax);
      if (wxmax >= 0) {   // the warp holds a set pixel of mask i: fold its extent into the CTA's box (shared-memory atomics)
        const int wxmin = __reduce_min_sync(0xffffffffu, xmin), wymin = __reduce_min_sync(0xffffffffu, ymin);
        const int wymax = __reduce_max_sync(0xffffffffu, ymax);
        if ((threadIdx.x & 31) == 0) {
          atomicMin(&cta_box[4 * i + 0], wxmin);
          atomicMin(&cta_box[4 * i + 1], wymin);
          atomicMax(&cta_box[4 * i + 2], wxmax);
          atomicMax(&cta_box[4 * i + 3], wymax);
        }
      }
    }
  }
  if (bbox) {   // one set of global atomics per CTA and mask it touches (dense masks otherwise serialise on 4 addresses per mask)
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (cta_box[4 * i + 2] >= 0) {
        atomicMin(bbox + 4 * i + 0, cta_box[4 * i + 0]);
        atomicMin(bbox + 4 * i + 1, cta_box[4 * i + 1]);
        atomicMax(bbox + 4 * i + 2, cta_box[4 * i + 2]);
        atomicMax(bbox + 4 * i + 3, cta_box[4 * i + 3]);
      }
    }
  }
}

// ---- flow post-processing ------------------------------------------------------------------------------------------------
struct AxisF {   // per destination index: two source indices and two float coefficients
  int s0, s1;
  float c0, c1;
};

// OpenCV resize(), INTER_LINEAR, float: fx = (float)((d + 0.5) * scale - 0.5); s = cvFloor(fx); fx -= s; the x axis clamps
// (s, fx) to the image with fx = 0, the y axis keeps fx and clips the two row indices
void build_axis_f(int dst_n, int src_n, bool clamp_fraction, std::vector<AxisF>* out) {
  const double scale = 1.0 / ((double)dst_n / (double)src_n);
  out->resize(dst_n);
  for (int d = 0; d < dst_n; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    AxisF a;
    if (clamp_fraction) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
      a.s0 = s;
      a.s1 = s + 1 < src_n - 1 ? s + 1 : src_n - 1;
    } else {
      a.s0 = s < 0 ? 0 : (s > src_n - 1 ? src_n - 1 : s);
      a.s1 = s + 1 < 0 ? 0 : (s + 1 > src_n - 1 ? src_n - 1 : s + 1);
    }
    a.c0 = 1.f - f;
    a.c1 = f;
    (*out)[d] = a;
  }
}

struct FlowTables {
  AxisF* cols = nullptr;
  AxisF* rows = nullptr;
};
std::mutex g_mu;
std::map<std::tuple<int, int, int, int, int>, FlowTables> g_tables;   // (device, sh, sw, dh, dw)

int get_tables(int sh, int sw, int dh, int dw, FlowTables* out) {
  int dev = 0;
  PV_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_mu);
  auto key = std::make_tuple(dev, sh, sw, dh, dw);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) { *out = it->second; return 0; }
  std::vector<AxisF> cols, rows;
  build_axis_f(dw, sw, true, &cols);
  build_axis_f(dh, sh, false, &rows);
  FlowTables t;
  PV_CUDA(cudaMalloc((void**)&t.cols, cols.size() * sizeof(AxisF)));
  PV_CUDA(cudaMalloc((void**)&t.rows, rows.size() * sizeof(AxisF)));
  PV_CUDA(cudaMemcpy(t.cols, cols.data(), cols.size() * sizeof(AxisF), cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(t.rows, rows.data(), rows.size() * sizeof(AxisF), cudaMemcpyHostToDevice));
  g_tables[key] = t;
  *out = t;
  return 0;
}

// flow2 [B,2,h,w] (network output) -> out [B,H,W,2]: ((S*20) resized) * (W/W_, H/H_), every product rounded to float32 like
// the reference's numpy / OpenCV steps (no FMA contraction across them)
__global__ void __launch_bounds__(256) flow_postprocess_kernel(const float* __restrict__ flow2, int h, int w, float* __restrict__ out,
                                                              int H, int W, const AxisF* __restrict__ cols,
                                                              const AxisF* __restrict__ rows, float su, float sv) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const int4 cxi = __ldg(reinterpret_cast<const int4*>(cols) + x), cyi = __ldg(reinterpret_cast<const int4*>(rows) + y);
  const float a0 = __int_as_float(cxi.z), a1 = __int_as_float(cxi.w), b0 = __int_as_float(cyi.z), b1 = __int_as_float(cyi.w);
  float r[2];
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const float* S = flow2 + ((long)b * 2 + c) * h * w;
    const float* s0 = S + (long)cyi.x * w;
    const float* s1 = S + (long)cyi.y * w;
    const float p00 = __fmul_rn(__ldg(s0 + cxi.x), 20.f), p01 = __fmul_rn(__ldg(s0 + cxi.y), 20.f);
    const float p10 = __fmul_rn(__ldg(s1 + cxi.x), 20.f), p11 = __fmul_rn(__ldg(s1 + cxi.y), 20.f);
    const float h0 = __fadd_rn(__fmul_rn(p00, a0), __fmul_rn(p01, a1)), h1 = __fadd_rn(__fmul_rn(p10, a0), __fmul_rn(p11, a1));
    r[c] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
  }
  reinterpret_cast<float2*>(out)[((long)b * H + y) * W + x] = make_float2(__fmul_rn(r[0], su), __fmul_rn(r[1], sv));
}

}  // namespace
}  // namespace premvos

using namespace premvos;

extern "C" int premvos_warp_masks_u8(const unsigned char* masks_dev, int n, int height, int width, const float* flow_dev,
                                     unsigned char* out_dev, float* bbox_dev, int binarize, void* stream) {
  PV_CHECK(n >= 0 && height > 0 && width > 0, PREMVOS_ERR_INVALID_ARG, "premvos_warp_masks_u8: bad sizes");
  if (n == 0) return 0;
  PV_CHECK(masks_dev && flow_dev && out_dev, PREMVOS_ERR_INVALID_ARG, "premvos_warp_masks_u8: null argument");
  PV_CHECK(masks_dev != out_dev, PREMVOS_ERR_INVALID_ARG, "premvos_warp_masks_u8: cannot warp in place");
  PV_CHECK(height <= 32767 && width <= 32767, PREMVOS_ERR_UNSUPPORTED, "premvos_warp_masks_u8: OpenCV's remap is limited to 32767 x 32767");
  cudaStream_t st = (cudaStream_t)stream;
  const long HW = (long)height * width;
  int* bb = reinterpret_cast<int*>(bbox_dev);
  if (bb) {
    prof_before(st);
    bbox_init_kernel<<<(n + 127) / 128, 128, 0, st>>>(bb, n, height, width);
    PV_TRY(after_launch("bbox_init_kernel", st));
  }
  const unsigned blocks = (unsigned)((HW + 256 * PX - 1) / (256 * PX));
  const bool vec = HW % PX == 0 && (reinterpret_cast<uintptr_t>(out_dev) & 3) == 0 && (reinterpret_cast<uintptr_t>(flow_dev) & 15) == 0;
  prof_before(st);
  const size_t smem = bb ? (size_t)n * 16 : 0;
  PV_CHECK(smem <= 48 * 1024, PREMVOS_ERR_UNSUPPORTED, "premvos_warp_masks_u8: at most 3072 masks per call with boxes");
  if (vec)
    warp_masks_kernel<true><<<blocks, 256, smem, st>>>(masks_dev, n, height, width, flow_dev, out_dev, bb, binarize);
  else
    warp_masks_kernel<false><<<blocks, 256, smem, st>>>(masks_dev, n, height, width, flow_dev, out_dev, bb, binarize);
  // algorithmic bytes: the flow once, every mask read once and written once
  PV_TRY(after_launch("warp_masks_kernel", st, 0.0, (double)HW * (8.0 + 2.0 * n)));
  if (bb) {
    prof_before(st);
    bbox_finish_kernel<<<(n + 127) / 128, 128, 0, st>>>(bb, n);
    PV_TRY(after_launch("bbox_finish_kernel", st));
  }
  return 0;
}

extern "C" int premvos_flow_postprocess(const float* flow2_dev, int batch, int net_h, int net_w, float* out_dev, int height,
                                        int width, void* stream) {
  PV_CHECK(flow2_dev && out_dev, PREMVOS_ERR_INVALID_ARG, "premvos_flow_postprocess: null argument");
  PV_CHECK(batch > 0 && net_h > 0 && net_w > 0 && height > 0 && width > 0 && height <= 65535 && batch <= 65535 && net_h % 64 == 0 &&
               net_w % 64 == 0,
           PREMVOS_ERR_INVALID_ARG, "premvos_flow_postprocess: bad sizes (the network input is a multiple of 64)");
  const int h = net_h / 4, w = net_w / 4;
  // OpenCV turns INTER_LINEAR into INTER_AREA for an exact 2x down-scale; the frame is never smaller than net / 4 here
  PV_CHECK(!(w == 2 * width && h == 2 * height), PREMVOS_ERR_UNSUPPORTED, "premvos_flow_postprocess: 2x down-scaling not implemented");
  FlowTables t;
  PV_TRY(get_tables(h, w, height, width, &t));
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((width + 255) / 256, height, batch);
  // the reference multiplies float32 arrays with the Python doubles W / float(W_), H / float(H_): numpy casts the scalar to float32
  const float su = (float)((double)width / (double)net_w), sv = (float)((double)height / (double)net_h);
  prof_before(st);
  flow_postprocess_kernel<<<grid, 256, 0, st>>>(flow2_dev, h, w, out_dev, height, width, t.cols, t.rows, su, sv);
  return after_launch("flow_postprocess_kernel", st, 0.0, (double)batch * (8.0 * h * w + 8.0 * height * width));
}

// ---- bit-packed masks ---------------------------------------------------------------------------------------------------
// The refinement output is one byte per pixel and proposal (0 / 1): 71 MB per 4-pair step, 98 % of what leaves the device and of
// what a rank sends to rank 0.  Packed 8 pixels per byte (pixel i of a mask -> bit i % 8 of byte i / 8,
// numpy.unpackbits(..., bitorder="little")) before it travels.  One thread per 64 pixels: four 16-byte loads, one 8-byte store.
namespace premvos {
namespace pack_bits {
__global__ void __launch_bounds__(256) pack_mask_bits_kernel(const unsigned char* __restrict__ masks, long n_masks, long hw, long words_per_mask,
                                                             unsigned long long* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_masks * words_per_mask) return;
  const long m = idx / words_per_mask, w = idx - m * words_per_mask;
  const unsigned char* src = masks + m * hw + w * 64;
  const long left = hw - w * 64;   // pixels of this mask from here on
  unsigned long long bits = 0;
  if (left >= 64 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const uint4* p = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint4 v = __ldg(p + q);
      const unsigned long long lo = ((unsigned long long)v.y << 32) | v.x, hi = ((unsigned long long)v.w << 32) | v.z;
      // bytes != 0 -> 0x01 each, then the 8 bytes of a word -> 8 bits (byte i -> bit i)
      auto squeeze = [](unsigned long long x) {
        x = (((x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | x) & 0x8080808080808080ull;
        return (unsigned long long)(((x >> 7) * 0x0102040810204080ull) >> 56);
      };
      bits |= (squeeze(lo) | (squeeze(hi) << 8)) << (16 * q);
    }
  } else {
    for (int i = 0; i < 64 && i < left; i++) bits |= (unsigned long long)(src[i] != 0) << i;
  }
  out[idx] = bits;
}
}  // namespace pack_bits
}  // namespace premvos

// Builds (and uploads, synchronously) the coefficient tables premvos_flow_postprocess needs for this geometry on the current
// device.  The enqueue-only entry point builds them on first use too, but that first call allocates and copies synchronously and
// therefore must not happen inside a stream capture: call this once per geometry beforehand.
extern "C" int premvos_flow_postprocess_prepare(int net_h, int net_w, int height, int width) {
  PV_CHECK(net_h > 0 && net_w > 0 && height > 0 && width > 0 && net_h % 4 == 0 && net_w % 4 == 0, PREMVOS_ERR_INVALID_ARG,
           "premvos_flow_postprocess_prepare: bad sizes");
  premvos::FlowTables t;
  return premvos::get_tables(net_h / 4, net_w / 4, height, width, &t);
}

// Frees the coefficient tables of premvos_flow_postprocess (all devices, all geometries); they are rebuilt on demand.
extern "C" void premvos_flow_postprocess_release(void) {
  std::lock_guard<std::mutex> lock(premvos::g_mu);
  for (auto& kv : premvos::g_tables) { cudaFree(kv.second.cols); cudaFree(kv.second.rows); }
  premvos::g_tables.clear();
}

extern "C" int premvos_pack_mask_bits(const unsigned char* masks_dev, long long n_masks, long long hw, unsigned char* out_dev, void* stream) {
  PV_CHECK(masks_dev && out_dev, PREMVOS_ERR_INVALID_ARG, "premvos_pack_mask_bits: null argument");
  PV_CHECK(n_masks >= 0 && hw > 0, PREMVOS_ERR_INVALID_ARG, "premvos_pack_mask_bits: bad sizes");
  PV_CHECK((reinterpret_cast<uintptr_t>(out_dev) & 7) == 0, PREMVOS_ERR_INVALID_ARG, "premvos_pack_mask_bits: out must be 8-byte aligned");
  if (n_masks == 0) return 0;
  const long words = (hw + 63) / 64;
  const long total = n_masks * words;
  cudaStream_t st = (cudaStream_t)stream;
  prof_before(st);
  premvos::pack_bits::pack_mask_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(masks_dev, n_masks, hw, words, reinterpret_cast<unsigned long long*>(out_dev));
  return after_launch("pack_mask_bits_kernel", st, 0.0, (double)n_masks * hw * 1.125);
}
