// Bandwidth-bound kernels of the refinement network (DeepLabv3+ / Xception-65 on box crops), CP8 activations:
//   * crop + resize + guidance + normalisation chain that builds the 4-channel network input for every proposal of
//     a frame in one launch (the reference re-feeds the whole frame over PCIe per proposal and runs the guidance
//     mask through a CPU py_func, datasets/Dataset.py:141-186, Resize.py:150-193);
//   * depthwise 3x3 convolution (stride, atrous rate, fixed_padding) with the unit's leading ReLU, the folded
//     BatchNorm and the optional trailing ReLU fused (deeplab/core/xception.py:92-190, 251-275);
//   * bilinear resize with align_corners=True (deeplab/model.py:399-400, 570-571), plane broadcast (image pooling);
//   * the output layer: legacy-TF1 bilinear resize of the logits to the input size, softmax, argmax, nearest /
//     bilinear resize to the crop size, zero padding back to the frame (network/SegmentationOutputLayers.py:35-61,
//     106-135) and the conf_score reduction of MergeTrack/refinement_net_functions.py:58-62 -- one kernel, the
//     only thing that leaves the device is a uint8 mask per proposal and one float.
#include <stdlib.h>

#include <algorithm>
#include "cp8.cuh"

namespace premvos {

using namespace cp8;

namespace {

// legacy TF1 resize coordinate (align_corners=False, no half-pixel centres): in = out * (in_size / out_size)
__device__ __forceinline__ void legacy_axis(int o, float scale, int in_size, int* lo, int* hi, float* lerp) {
  const float src = __fmul_rn((float)o, scale);
  const int l = (int)floorf(src);
  *lo = l;
  *hi = min(l + 1, in_size - 1);
  *lerp = __fsub_rn(src, (float)l);
}

// ---- network input ----------------------------------------------------------------------------------------------
struct PreArgs {
  const unsigned char* frame; int H, W;   // uint8 RGB frame
  const float* boxes;                      // [N][4] x, y, w, h (proposal 'bbox')
  int N, S;                                // S = network input size (385)
  CV out;                                  // [N][1 chunk][S][S]: ch 0..2 RGB in [-1,1], ch 3 guidance -1/+1
  int rows3;                               // 1: `out` is the ROW IM2COL of the 3x3 / stride-2 root convolution instead:
                                           // [N][2 chunks][S][(S+1)/2], channel s*4 + c of (y, xo) = input(y, 2*xo + s - 1, c)
  int* crops;                              // [N][4] y0 x0 y1 x1 (CROP_BOXES_y0x0y1x1)
};

__device__ __forceinline__ void crop_of(const float* b, int H, int W, int* cy0, int* cx0, int* cy1, int* cx1, int* gy0, int* gx0,
                                        int* gy1, int* gx1) {
  // FewShotFeedSegmentationDataset.py:40-43 (fp32 placeholder), tf.round / np.round = half to even
  const float y0 = b[1], x0 = b[0], y1 = __fadd_rn(b[3], b[1]), x1 = __fadd_rn(b[2], b[0]);
  *gy0 = (int)rintf(y0); *gx0 = (int)rintf(x0); *gy1 = (int)rintf(y1); *gx1 = (int)rintf(x1);
  *cy0 = max(*gy0 - 50, 0); *cx0 = max(*gx0 - 50, 0); *cy1 = min(*gy1 + 50, H); *cx1 = min(*gx1 + 50, W);   // Resize.py:151-164
}

__global__ void __launch_bounds__(256) refine_input_kernel(PreArgs a) {
  const long total = (long)a.N * a.S * a.S;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % a.S), y = (int)((idx / a.S) % a.S), n = (int)(idx / ((long)a.S * a.S));
  int cy0, cx0, cy1, cx1, gy0, gx0, gy1, gx1;
  crop_of(a.boxes + n * 4, a.H, a.W, &cy0, &cx0, &cy1, &cx1, &gy0, &gx0, &gy1, &gx1);
  if (x == 0 && y == 0) { a.crops[n * 4] = cy0; a.crops[n * 4 + 1] = cx0; a.crops[n * 4 + 2] = cy1; a.crops[n * 4 + 3] = cx1; }
  const int ch = cy1 - cy0, cw = cx1 - cx0;
  F8 f = zero8();
  if (ch > 0 && cw > 0) {
    // image: tf.image.resize_images bilinear (Util.py:25), then (v - mean) / std, then DeepLab's (n*std + mean)*255 and 2/255*u - 1
    int ylo, yhi, xlo, xhi; float yl, xl;
    legacy_axis(y, __fdiv_rn((float)ch, (float)a.S), ch, &ylo, &yhi, &yl);
    legacy_axis(x, __fdiv_rn((float)cw, (float)a.S), cw, &xlo, &xhi, &xl);
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; c++) {
      auto px = [&](int yy, int xx) { return (float)((double)a.frame[((long)(cy0 + yy) * a.W + (cx0 + xx)) * 3 + c] / 255.0); };
      const float tl = px(ylo, xlo), tr = px(ylo, xhi), bl = px(yhi, xlo), br = px(yhi, xhi);
      const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl)), bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
      const float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
      const float nrm = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
      const float u = __fmul_rn(__fadd_rn(__fmul_rn(nrm, stdv[c]), mean[c]), 255.f);
      f.v[c] = __fsub_rn(__fmul_rn(2.0f / 255.0f, u), 1.0f);
    }
    // guidance: nearest-neighbour resize of the box mask (Util.py:27-28, BoundingBox.py:15-19)
    const int sy = min((int)floorf(__fmul_rn((float)y, __fdiv_rn((float)ch, (float)a.S))), ch - 1) + cy0;
    const int sx = min((int)floorf(__fmul_rn((float)x, __fdiv_rn((float)cw, (float)a.S))), cw - 1) + cx0;
    // encoded[y0:y1, x0:x1] = 1 with numpy's slice rules: a negative bound counts from the far edge (so a box that starts left of /
    // above the frame usually selects nothing), bounds are clamped to the array, start >= stop is empty
    auto sl = [](int v, int n) { return v < 0 ? max(v + n, 0) : min(v, n); };
    const int ys = sl(gy0, a.H), ye = sl(gy1, a.H), xs = sl(gx0, a.W), xe = sl(gx1, a.W);
    const float g = (sy >= ys && sy < ye && sx >= xs && sx < xe) ? 1.f : 0.f;
    f.v[3] = __fsub_rn(__fmul_rn(2.0f / 255.0f, __fmul_rn(g, 255.f)), 1.0f);
  }
  if (!a.rows3) {
    st_chunk(a.out.hi, a.out.lo, cv_elem(a.out, n, 0, y, x), f);
    return;
  }
  // row im2col: pixel x is tap 1 of output column x/2 (x even), or tap 0 of column (x+1)/2 and tap 2 of column (x-1)/2 (x odd);
  // the slots no pixel maps to (tap 0 of column 0, taps beyond the right border) keep the zeros the buffer was allocated with
  uint32_t h0, l0, h1, l1;
  {
    const __nv_bfloat16 a0 = __float2bfloat16_rn(f.v[0]), a1 = __float2bfloat16_rn(f.v[1]), a2 = __float2bfloat16_rn(f.v[2]), a3 = __float2bfloat16_rn(f.v[3]);
    const __nv_bfloat16 b0 = __float2bfloat16_rn(f.v[0] - __bfloat162float(a0)), b1 = __float2bfloat16_rn(f.v[1] - __bfloat162float(a1));
    const __nv_bfloat16 b2 = __float2bfloat16_rn(f.v[2] - __bfloat162float(a2)), b3 = __float2bfloat16_rn(f.v[3] - __bfloat162float(a3));
    h0 = (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(a1) << 16);
    h1 = (uint32_t)__bfloat16_as_ushort(a2) | ((uint32_t)__bfloat16_as_ushort(a3) << 16);
    l0 = (uint32_t)__bfloat16_as_ushort(b0) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
    l1 = (uint32_t)__bfloat16_as_ushort(b2) | ((uint32_t)__bfloat16_as_ushort(b3) << 16);
  }
  auto put = [&](int chunk, int xo, int half) {
    if (xo < 0 || xo >= a.out.W) return;
    const long e = cv_elem(a.out, n, chunk, y, xo) + half * 4;
    *reinterpret_cast<uint2*>(a.out.hi + e) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(a.out.lo + e) = make_uint2(l0, l1);
  };
  if ((x & 1) == 0) put(0, x >> 1, 1);
  else { put(0, (x + 1) >> 1, 0); put(1, (x - 1) >> 1, 0); }
}

// ---- depthwise 3x3 -----------------------------------------------------------------------------------------------
// One CTA = one output tile of one 8-channel chunk plane of one crop.  The input region of the tile (with its halo) is
// read once with coalesced 8-byte loads, hi + lo are summed to fp32 (leading ReLU applied) and parked in shared memory
// as [row][col][8 floats]; after that every thread works on 4 channels ("half" of a chunk) so that neighbouring threads
// read neighbouring 16-byte words of shared memory (conflict-free) and write neighbouring 8-byte words of both planes.
// Stride-1 / rate-1 layers (62 of the 68 depthwise layers of Xception-65 + decoder) take the FAST path: a thread owns a
// vertical strip of RS outputs and slides a 3-row window down it, 3 LDS.128 per input row.  Tap order per output is
// r-major, s-minor in both paths (same bits as a plain loop).
struct DwArgs {
  CV in, out;
  const float* w;     // [9][Cpad] (BatchNorm scale folded in), Cpad = 8 * chunks
  const float* bias;  // [Cpad]
  int stride, rate, pad;   // input coordinate = o*stride + t*rate - pad
  int pre_relu, post_relu;
  int TH, TW, ntx;         // output tile, tiles per row of tiles
  int RH, RW;              // unclipped input region of a tile
  int clip;                // 1: keep only the part of the region inside the image, taps test the bounds
  int w_c0, cpad;          // first 8-channel chunk of this launch inside w / bias, their padded channel count
  // fused align-corners upsampling: the depthwise input is `src` bilinearly resized to in.H x in.W (in.hi / in.lo unused); the
  // staging loop samples it instead of loading -- the resized tensor never exists in HBM (decoder: 25x25 -> 97x97, 256 channels)
  CV src;
  int upsample;
};

__device__ __forceinline__ void split2(float x0, float x1, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);   // one F2FP for the pair
  const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __uint_as_float(hw << 16), x1 - __uint_as_float(hw & 0xffff0000u));
  *hi = hw; *lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void st_half(const CV& o, long elem, int half, const float4& f) {
  uint32_t h0, l0, h1, l1;
  split2(f.x, f.y, &h0, &l0); split2(f.z, f.w, &h1, &l1);
  *reinterpret_cast<uint2*>(o.hi + elem + half * 4) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(o.lo + elem + half * 4) = make_uint2(l0, l1);
}
__device__ __forceinline__ void fma4(float4& acc, const float4& v, const float4& w) {
  acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y); acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
}

template <int RS, bool FAST>
__global__ void __launch_bounds__(256) depthwise3x3_tile_kernel(DwArgs a) {
  extern __shared__ float4 dw_sm[];   // [RH][RW][2 halves]
  const int tile = blockIdx.x, ch = blockIdx.y, n = blockIdx.z;
  const int oy0 = (tile / a.ntx) * a.TH, ox0 = (tile % a.ntx) * a.TW;
  int ry0 = oy0 * a.stride - a.pad, rx0 = ox0 * a.stride - a.pad, RH = a.RH, RW = a.RW;
  if (a.clip) {
    const int y1 = min(ry0 + RH, a.in.H), x1 = min(rx0 + RW, a.in.W);
    ry0 = max(ry0, 0); rx0 = max(rx0, 0);
    RH = max(y1 - ry0, 0); RW = max(x1 - rx0, 0);
  }
  // ---- stage the input region: 8 independent 8-byte loads per plane in flight per thread ----
  const long plane = cv_elem(a.in, n, ch, 0, 0);
  const int total = RH * RW * 2;
  const unsigned magic = (1u << 20) / (unsigned)max(RW, 1) + 1u;   // e / RW for e < 2^20 / RW (host-checked)
  if (a.upsample) {
    // resize_bilinear(align_corners=True) of `src`, the operations of resize_ac_kernel in the same order, kept in fp32
    const float sy = a.in.H > 1 ? __fdiv_rn((float)(a.src.H - 1), (float)(a.in.H - 1)) : __fdiv_rn((float)a.src.H, (float)a.in.H);
    const float sx = a.in.W > 1 ? __fdiv_rn((float)(a.src.W - 1), (float)(a.in.W - 1)) : __fdiv_rn((float)a.src.W, (float)a.in.W);
    const long splane = cv_elem(a.src, n, ch, 0, 0);
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int half = idx & 1, e = idx >> 1, ry = (int)(((unsigned)e * magic) >> 20), rx = e - ry * RW;
      const int gy = ry0 + ry, gx = rx0 + rx;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < a.in.H && gx >= 0 && gx < a.in.W) {
        int ylo, yhi, xlo, xhi; float yl, xl;
        legacy_axis(gy, sy, a.src.H, &ylo, &yhi, &yl);
        legacy_axis(gx, sx, a.src.W, &xlo, &xhi, &xl);
        auto ld = [&](int yy, int xx) {
          const long el = splane + (long)(yy * a.src.W + xx) * 8 + half * 4;
          const uint2 h = *reinterpret_cast<const uint2*>(a.src.hi + el), l = *reinterpret_cast<const uint2*>(a.src.lo + el);
          return make_float4(__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16), __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u),
                             __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16), __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u));
        };
        const float4 tl = ld(ylo, xlo), tr = ld(ylo, xhi), bl = ld(yhi, xlo), br = ld(yhi, xhi);
        auto mix = [&](float p, float q, float r, float t) {
          const float top = __fadd_rn(p, __fmul_rn(__fsub_rn(q, p), xl)), bot = __fadd_rn(r, __fmul_rn(__fsub_rn(t, r), xl));
          return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
        };
        f = make_float4(mix(tl.x, tr.x, bl.x, br.x), mix(tl.y, tr.y, bl.y, br.y), mix(tl.z, tr.z, bl.z, br.z), mix(tl.w, tr.w, bl.w, br.w));
        if (a.pre_relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
      }
      dw_sm[idx] = f;
    }
  } else
  for (int base = threadIdx.x; base < total; base += 8 * blockDim.x) {
    uint2 h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int idx = base + q * blockDim.x;
      const int half = idx & 1, e = idx >> 1, ry = (int)(((unsigned)e * magic) >> 20), rx = e - ry * RW;
      const int gy = ry0 + ry, gx = rx0 + rx;
      h[q] = make_uint2(0u, 0u); l[q] = make_uint2(0u, 0u);
      if (idx < total && gy >= 0 && gy < a.in.H && gx >= 0 && gx < a.in.W) {
        const long el = plane + (long)(gy * a.in.W + gx) * 8 + half * 4;
        h[q] = *reinterpret_cast<const uint2*>(a.in.hi + el);
        l[q] = *reinterpret_cast<const uint2*>(a.in.lo + el);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int idx = base + q * blockDim.x;
      if (idx >= total) break;
      float4 f;
      f.x = __uint_as_float(h[q].x << 16) + __uint_as_float(l[q].x << 16);
      f.y = __uint_as_float(h[q].x & 0xffff0000u) + __uint_as_float(l[q].x & 0xffff0000u);
      f.z = __uint_as_float(h[q].y << 16) + __uint_as_float(l[q].y << 16);
      f.w = __uint_as_float(h[q].y & 0xffff0000u) + __uint_as_float(l[q].y & 0xffff0000u);
      if (a.pre_relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
      dw_sm[idx] = f;
    }
  }
  __syncthreads();
  // ---- compute ----
  const int cpad = a.cpad, wch = a.w_c0 + ch;
  const int units = FAST ? (a.TH / RS) * a.TW * 2 : a.TH * a.TW * 2;
  const unsigned tw_magic = (1u << 20) / (unsigned)a.TW + 1u;
  for (int u = threadIdx.x; u < units; u += blockDim.x) {
    const int half = u & 1, p = u >> 1, ty = (int)(((unsigned)p * tw_magic) >> 20), tx = p - ty * a.TW;
    const int ox = ox0 + tx;
    if (ox >= a.out.W) continue;
    float4 wv[9];
#pragma unroll
    for (int t = 0; t < 9; t++) wv[t] = *reinterpret_cast<const float4*>(a.w + t * cpad + wch * 8 + half * 4);
    const float4 bv = *reinterpret_cast<const float4*>(a.bias + wch * 8 + half * 4);
    if (FAST) {
      float4 acc[RS];
#pragma unroll
      for (int k = 0; k < RS; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* base = dw_sm + ((long)(ty * RS) * RW + tx) * 2 + half;
#pragma unroll
      for (int j = 0; j < RS + 2; j++) {
        const float4 v0 = base[(j * RW) * 2], v1 = base[(j * RW + 1) * 2], v2 = base[(j * RW + 2) * 2];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const int k = j - r;
          if (k < 0 || k >= RS) continue;
          fma4(acc[k], v0, wv[r * 3]); fma4(acc[k], v1, wv[r * 3 + 1]); fma4(acc[k], v2, wv[r * 3 + 2]);
        }
      }
#pragma unroll
      for (int k = 0; k < RS; k++) {
        const int oy = oy0 + ty * RS + k;
        if (oy >= a.out.H) break;
        float4 t = make_float4(acc[k].x + bv.x, acc[k].y + bv.y, acc[k].z + bv.z, acc[k].w + bv.w);
        if (a.post_relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
        st_half(a.out, cv_elem(a.out, n, ch, oy, ox), half, t);
      }
    } else {
      const int oy = oy0 + ty;
      if (oy >= a.out.H) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int sy = oy * a.stride + r * a.rate - a.pad - ry0;
        if (sy < 0 || sy >= RH) continue;
#pragma unroll
        for (int s = 0; s < 3; s++) {
          const int sx = ox * a.stride + s * a.rate - a.pad - rx0;
          if (sx < 0 || sx >= RW) continue;
          fma4(acc, dw_sm[(sy * RW + sx) * 2 + half], wv[r * 3 + s]);
        }
      }
      float4 t = make_float4(acc.x + bv.x, acc.y + bv.y, acc.z + bv.z, acc.w + bv.w);
      if (a.post_relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
      st_half(a.out, cv_elem(a.out, n, ch, oy, ox), half, t);
    }
  }
}

// ---- bilinear resize, align_corners=True ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_ac_kernel(CV in, CV out, int n_active) {
  const int nch = (in.C + 7) / 8;
  const long total = (long)n_active * nch * out.H * out.W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % out.W), oy = (int)((idx / out.W) % out.H);
  const int ch = (int)((idx / ((long)out.W * out.H)) % nch), n = (int)(idx / ((long)out.W * out.H * nch));
  const float sy = out.H > 1 ? __fdiv_rn((float)(in.H - 1), (float)(out.H - 1)) : __fdiv_rn((float)in.H, (float)out.H);
  const float sx = out.W > 1 ? __fdiv_rn((float)(in.W - 1), (float)(out.W - 1)) : __fdiv_rn((float)in.W, (float)out.W);
  int ylo, yhi, xlo, xhi; float yl, xl;
  legacy_axis(oy, sy, in.H, &ylo, &yhi, &yl);
  legacy_axis(ox, sx, in.W, &xlo, &xhi, &xl);
  const F8 tl = ld_chunk(in.hi, in.lo, cv_elem(in, n, ch, ylo, xlo)), tr = ld_chunk(in.hi, in.lo, cv_elem(in, n, ch, ylo, xhi));
  const F8 bl = ld_chunk(in.hi, in.lo, cv_elem(in, n, ch, yhi, xlo)), br = ld_chunk(in.hi, in.lo, cv_elem(in, n, ch, yhi, xhi));
  F8 r;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float top = __fadd_rn(tl.v[j], __fmul_rn(__fsub_rn(tr.v[j], tl.v[j]), xl));
    const float bot = __fadd_rn(bl.v[j], __fmul_rn(__fsub_rn(br.v[j], bl.v[j]), xl));
    r.v[j] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
  }
  st_chunk(out.hi, out.lo, cv_elem(out, n, ch, oy, ox), r);
}

// vec [N][C] (fp32, optional ReLU) -> every pixel of out (image-level feature, model.py:397-400)
__global__ void __launch_bounds__(256) broadcast_kernel(const float* __restrict__ vec, int C, int relu, CV out, int n_active) {
  const int nch = (C + 7) / 8;
  const long total = (long)n_active * nch * out.H * out.W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % out.W), oy = (int)((idx / out.W) % out.H);
  const int ch = (int)((idx / ((long)out.W * out.H)) % nch), n = (int)(idx / ((long)out.W * out.H * nch));
  F8 f = zero8();
#pragma unroll
  for (int j = 0; j < 8; j++)
    if (ch * 8 + j < C) {
      const float v = vec[(long)n * C + ch * 8 + j];
      f.v[j] = relu ? fmaxf(v, 0.f) : v;
    }
  st_chunk(out.hi, out.lo, cv_elem(out, n, ch, oy, ox), f);
}

// ---- output layer -----------------------------------------------------------------------------------------------------
struct OutArgs {
  const float* logits; int lh, lw, lcs;   // fp32 channels-last [N][lh][lw][lcs], classes 0/1
  const int* crops;                        // [N][4]
  int N, S, H, W;
  unsigned char* mask;                     // [N][H][W] 0/1
  float* posterior;                        // [N][H][W] or null
  double* conf_sum;                        // [N] sum over the frame of 2*p' - 1
};

// logits at input resolution (SegmentationOutputLayers.py:36): legacy bilinear of the [lh,lw] logits
__device__ __forceinline__ void logits_at(const OutArgs& a, int n, int y, int x, float* l0, float* l1) {
  int ylo, yhi, xlo, xhi; float yl, xl;
  legacy_axis(y, __fdiv_rn((float)a.lh, (float)a.S), a.lh, &ylo, &yhi, &yl);
  legacy_axis(x, __fdiv_rn((float)a.lw, (float)a.S), a.lw, &xlo, &xhi, &xl);
  const float* base = a.logits + (long)n * a.lh * a.lw * a.lcs;
  const float* ptl = base + ((long)ylo * a.lw + xlo) * a.lcs; const float* ptr = base + ((long)ylo * a.lw + xhi) * a.lcs;
  const float* pbl = base + ((long)yhi * a.lw + xlo) * a.lcs; const float* pbr = base + ((long)yhi * a.lw + xhi) * a.lcs;
  float o[2];
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const float top = __fadd_rn(ptl[c], __fmul_rn(__fsub_rn(ptr[c], ptl[c]), xl));
    const float bot = __fadd_rn(pbl[c], __fmul_rn(__fsub_rn(pbr[c], pbl[c]), xl));
    o[c] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
  }
  *l0 = o[0]; *l1 = o[1];
}
__device__ __forceinline__ float fg_prob(float l0, float l1) {  // softmax(...)[1], max-subtracted like tf.nn.softmax
  const float m = fmaxf(l0, l1);
  const float e0 = expf(l0 - m), e1 = expf(l1 - m);
  return e1 / (e0 + e1);
}

__global__ void conf_finish_kernel(const double* __restrict__ sum, int N, double hw, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = (float)(sum[i] / hw);
}

__global__ void __launch_bounds__(256) refine_output_kernel(OutArgs a) {
  __shared__ double red[8];
  const int n = blockIdx.y;
  const long hw = (long)a.H * a.W;
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double contrib = 0.0;
  if (p < hw) {
    const int y = (int)(p / a.W), x = (int)(p % a.W);
    const int cy0 = a.crops[n * 4], cx0 = a.crops[n * 4 + 1], cy1 = a.crops[n * 4 + 2], cx1 = a.crops[n * 4 + 3];
    unsigned char m = 0;
    float post = 0.f;
    if (y >= cy0 && y < cy1 && x >= cx0 && x < cx1) {
      const int ch = cy1 - cy0, cw = cx1 - cx0, yy = y - cy0, xx = x - cx0;
      // class map: nearest-neighbour resize S x S -> crop size (:120-125), argmax of the logits (first max wins)
      const int sy = min((int)floorf(__fmul_rn((float)yy, __fdiv_rn((float)a.S, (float)ch))), a.S - 1);
      const int sx = min((int)floorf(__fmul_rn((float)xx, __fdiv_rn((float)a.S, (float)cw))), a.S - 1);
      float l0, l1;
      logits_at(a, n, sy, sx, &l0, &l1);
      m = l1 > l0 ? 1 : 0;
      // posterior: bilinear (legacy) resize of softmax(logits)[..., 1] from S x S to the crop size
      int ylo, yhi, xlo, xhi; float yl, xl;
      legacy_axis(yy, __fdiv_rn((float)a.S, (float)ch), a.S, &ylo, &yhi, &yl);
      legacy_axis(xx, __fdiv_rn((float)a.S, (float)cw), a.S, &xlo, &xhi, &xl);
      float q[4];
      const int ys[2] = {ylo, yhi}, xs[2] = {xlo, xhi};
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          logits_at(a, n, ys[i], xs[j], &l0, &l1);
          q[i * 2 + j] = fg_prob(l0, l1);
        }
      const float top = __fadd_rn(q[0], __fmul_rn(__fsub_rn(q[1], q[0]), xl)), bot = __fadd_rn(q[2], __fmul_rn(__fsub_rn(q[3], q[2]), xl));
      post = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
    }
    a.mask[(long)n * hw + p] = m;
    if (a.posterior) a.posterior[(long)n * hw + p] = post;
    const float c = m ? post : __fsub_rn(1.0f, post);          // refinement_net_functions.py:58-61
    contrib = (double)__fsub_rn(__fmul_rn(2.0f, c), 1.0f);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; i++) s += red[i];
    atomicAdd(a.conf_sum + n, s);
  }
}

// global average pool: one warp per (image, chunk plane), lanes stride over the pixels
__global__ void __launch_bounds__(256) gap_kernel(CV feat, float* __restrict__ pooled /*[N][C]*/, int n_active) {
  const int nch = (feat.C + 7) / 8, hw = feat.H * feat.W;
  const int gw = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (gw >= n_active * nch) return;
  const int n = gw / nch, ch = gw - n * nch;
  F8 s = zero8();
  for (int p = lane; p < hw; p += 32) {
    const F8 v = ld_chunk(feat.hi, feat.lo, (((long)n * feat.chunks + feat.c0 + ch) * hw + p) * 8);
#pragma unroll
    for (int j = 0; j < 8; j++) s.v[j] += v.v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s.v[j] += __shfl_xor_sync(0xffffffffu, s.v[j], off);
  }
  if (lane < 8 && ch * 8 + lane < feat.C) {
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) if (j == lane) m = s.v[j];
    pooled[(long)n * feat.C + ch * 8 + lane] = m / (float)hw;
  }
}

// out[n][o] = act(pooled[n] . W[:, o] + bias[o]); one warp per output
__global__ void __launch_bounds__(256) fc_kernel(const float* __restrict__ pooled, int C, const float* __restrict__ Wt,
                                                 const float* __restrict__ bias, int nout, int relu, float* __restrict__ out, int n_active) {
  const int gw = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (gw >= n_active * nout) return;
  const int n = gw / nout, o = gw - n * nout;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc = fmaf(pooled[(long)n * C + c], Wt[(long)c * nout + o], acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    const float v = acc + bias[o];
    out[(long)n * nout + o] = relu ? fmaxf(v, 0.f) : v;
  }
}

}  // namespace

int refine_make_input(const unsigned char* frame_rgb, int H, int W, const float* boxes_xywh, int N, int S, const CView& out, int* crops,
                      cudaStream_t st) {
  const int rows3 = (out.C == 16 && out.W == (S - 1) / 2 + 1) ? 1 : 0;
  PV_CHECK(out.H == S && (rows3 || out.W == S) && N <= out.N, PREMVOS_ERR_INVALID_ARG, "refine_make_input: shape mismatch");
  PreArgs a{frame_rgb, H, W, boxes_xywh, N, S, dev(out), rows3, crops};
  const long total = (long)N * S * S;
  if (total == 0) return 0;
  prof_before(st);
  refine_input_kernel<<<blocks_for(total), 256, 0, st>>>(a);
  return after_launch("refine_input_kernel", st, 60.0 * total, (double)total * (12.0 + 32.0));
}

int depthwise3x3_cp8(const CView& in, const CView& out, const float* w, const float* bias, int stride, int rate, int pad, bool pre_relu,
                     bool post_relu, int n_active, cudaStream_t st) {
  return depthwise3x3_cp8_ex(in, nullptr, out, w, bias, 0, round_up(in.C, 8), stride, rate, pad, pre_relu, post_relu, n_active, st);
}

// General form: weights / bias of this launch start at chunk w_c0 of arrays padded to cpad channels (one depthwise layer run as
// several launches over channel ranges); up_src != nullptr: the input is *up_src resized (align_corners=True) to in.H x in.W,
// sampled while staging (in's planes are not read).
int depthwise3x3_cp8_ex(const CView& in, const CView* up_src, const CView& out, const float* w, const float* bias, int w_c0, int cpad, int stride,
                        int rate, int pad, bool pre_relu, bool post_relu, int n_active, cudaStream_t st) {
  PV_CHECK(in.C == out.C && in.N == out.N, PREMVOS_ERR_INVALID_ARG, "depthwise3x3_cp8: shape mismatch");
  PV_CHECK(!up_src || (up_src->C == in.C && up_src->N == in.N && up_src->hi), PREMVOS_ERR_INVALID_ARG, "depthwise3x3_cp8: upsampling source mismatch");
  const long total = (long)n_active * in.vchunks() * out.H * out.W;
  if (total == 0) return 0;
  const bool fast = stride == 1 && rate == 1;
  const int RS = (fast && out.H % 5 == 0) ? 5 : 4;
  const int target = stride == 1 ? 32 : 16;
  const int ntx = (out.W + target - 1) / target, TW = (out.W + ntx - 1) / ntx;
  int TH, nty;
  if (fast) {   // strips of RS rows; 5..8 strips per tile, least padding wins
    const int strips = (out.H + RS - 1) / RS;
    int best = 8, waste = 1 << 30;
    for (int sp = 8; sp >= 5; sp--) {
      const int wst = (strips + sp - 1) / sp * sp - strips;
      if (wst < waste) { waste = wst; best = sp; }
    }
    if (strips < 5) best = strips;
    TH = best * RS; nty = (strips + best - 1) / best;
  } else {
    nty = (out.H + target - 1) / target; TH = (out.H + nty - 1) / nty;
  }
  const int clip = rate > 2 ? 1 : 0;
  const int RH = (TH - 1) * stride + 2 * rate + 1, RW = (TW - 1) * stride + 2 * rate + 1;
  const size_t smem = (size_t)(clip ? std::min(RH, in.H) : RH) * (clip ? std::min(RW, in.W) : RW) * 32;
  PV_CHECK(smem <= 200 * 1024 && (long)RH * RW * RW < (1 << 20) && (long)TH * TW * TW < (1 << 20), PREMVOS_ERR_UNSUPPORTED,
           "depthwise3x3_cp8: tile does not fit shared memory");
  const int units = fast ? (TH / RS) * TW * 2 : TH * TW * 2;
  const int iters = (units + 255) / 256;
  const int threads = std::min(256, ((units + iters - 1) / iters + 31) / 32 * 32);
  DwArgs a{dev(in), dev(out), w, bias, stride, rate, pad, pre_relu ? 1 : 0, post_relu ? 1 : 0, TH, TW, ntx, RH, RW, clip, w_c0, cpad,
           up_src ? dev(*up_src) : dev(in), up_src ? 1 : 0};
  auto kern = fast ? (RS == 5 ? depthwise3x3_tile_kernel<5, true> : depthwise3x3_tile_kernel<4, true>) : depthwise3x3_tile_kernel<4, false>;
  if (smem > 48 * 1024) PV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_before(st);
  kern<<<dim3(ntx * nty, in.vchunks(), n_active), threads, smem, st>>>(a);
  const double frac = (double)n_active / in.N;
  const char* label = "depthwise3x3_kernel";
  static const int per_layer = getenv("PREMVOS_PROFILE_LAYERS") ? atoi(getenv("PREMVOS_PROFILE_LAYERS")) : 0;
  if (per_layer && profiling_enabled()) {
    char buf[160];
    snprintf(buf, sizeof(buf), "dw_cp8[n%d_%dx%d_c%d_s%d_r%d%s]", n_active, out.H, out.W, in.C, stride, rate, up_src ? "_upsampled" : "");
    label = prof_intern(buf);
  }
  return after_launch(label, st, 18.0 * total * 8, 4.0 * frac * ((double)(up_src ? up_src->pixels() : in.pixels()) + (double)out.pixels()) * in.C);
}

int resize_bilinear_ac_cp8(const CView& in, const CView& out, int n_active, cudaStream_t st) {
  PV_CHECK(in.C == out.C && in.N == out.N, PREMVOS_ERR_INVALID_ARG, "resize_bilinear_ac_cp8: shape mismatch");
  const long total = (long)n_active * in.vchunks() * out.H * out.W;
  if (total == 0) return 0;
  prof_before(st);
  resize_ac_kernel<<<blocks_for(total), 256, 0, st>>>(dev(in), dev(out), n_active);
  return after_launch("resize_ac_kernel", st, 6.0 * total * 8, 4.0 * (double)n_active * in.C * ((double)in.H * in.W + (double)out.H * out.W));
}

int broadcast_vec_cp8(const float* vec, int C, bool relu, const CView& out, int n_active, cudaStream_t st) {
  const long total = (long)n_active * ((C + 7) / 8) * out.H * out.W;
  if (total == 0) return 0;
  prof_before(st);
  broadcast_kernel<<<blocks_for(total), 256, 0, st>>>(vec, C, relu ? 1 : 0, dev(out), n_active);
  return after_launch("broadcast_kernel", st, 0.0, 32.0 * total);
}

int gap_fc_relu(const CView& feat, const float* Wt, const float* bias, int nout, bool relu, float* pooled_scratch, float* out,
                int n_active, cudaStream_t st) {
  if (n_active == 0) return 0;
  const long warps = (long)n_active * feat.vchunks();
  prof_before(st);
  gap_kernel<<<blocks_for(warps * 32), 256, 0, st>>>(dev(feat), pooled_scratch, n_active);
  PV_TRY(after_launch("gap_kernel", st, 0.0, 4.0 * (double)n_active * feat.H * feat.W * feat.C));
  prof_before(st);
  fc_kernel<<<blocks_for((long)n_active * nout * 32), 256, 0, st>>>(pooled_scratch, feat.C, Wt, bias, nout, relu ? 1 : 0, out, n_active);
  return after_launch("fc_kernel", st, 2.0 * n_active * feat.C * nout, 4.0 * (double)feat.C * nout);
}

int refine_output(const TView& logits, const int* crops, int N, int S, int H, int W, unsigned char* mask, float* posterior,
                  double* conf_sum, cudaStream_t st) {
  PV_CHECK(logits.p && logits.coff == 0 && N <= logits.N, PREMVOS_ERR_INVALID_ARG, "refine_output: bad logits view");
  if (N == 0) return 0;
  OutArgs a{logits.p, logits.H, logits.W, logits.cs, crops, N, S, H, W, mask, posterior, conf_sum};
  PV_CUDA(cudaMemsetAsync(conf_sum, 0, (size_t)N * sizeof(double), st));
  dim3 grid(blocks_for((long)H * W), N);
  prof_before(st);
  refine_output_kernel<<<grid, 256, 0, st>>>(a);
  return after_launch("refine_output_kernel", st, 100.0 * N * H * W, (double)N * H * W * (posterior ? 5.0 : 1.0));
}

int refine_conf_finish(const double* conf_sum, int N, long hw, float* conf_out, cudaStream_t st) {
  if (N == 0) return 0;
  prof_before(st);
  conf_finish_kernel<<<blocks_for(N), 256, 0, st>>>(conf_sum, N, (double)hw, conf_out);
  return after_launch("conf_finish_kernel", st, 0.0, 12.0 * N);
}

}  // namespace premvos
