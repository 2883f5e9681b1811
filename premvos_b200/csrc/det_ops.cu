// Detection-side kernels of the proposal network (everything that is not a convolution):
// image normalisation, 3x3/s2 max pooling, RPN box decoding, exact top-k (radix select + bitonic sort),
// greedy NMS (IoU bitmask + one-warp scan, TensorFlow's IoU arithmetic so that the kept INDICES are
// bit-exact with the oracle on identical inputs), RoIAlign as the reference builds it
// (tf.image.crop_and_resize at 2x resolution + 2x2 average pooling, model.py:301-374),
// global-average-pool + fully connected heads, and the Fast R-CNN inference tail (train.py:275-295).
// All HBM-bound or tiny; activations are CP8 split-bf16 planes (common.cuh).
#include "common.cuh"

namespace premvos {

namespace {

struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld_chunk(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long elem) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi + elem);
  const uint4 l = *reinterpret_cast<const uint4*>(lo + elem);
  const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
  F8 r;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    r.v[2 * j] = __uint_as_float(hh[j] << 16) + __uint_as_float(ll[j] << 16);
    r.v[2 * j + 1] = __uint_as_float(hh[j] & 0xffff0000u) + __uint_as_float(ll[j] & 0xffff0000u);
  }
  return r;
}
__device__ __forceinline__ void st_chunk(__nv_bfloat16* hi, __nv_bfloat16* lo, long elem, const F8& f) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float x0 = f.v[2 * j], x1 = f.v[2 * j + 1];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
    hw[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lw[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(hi + elem) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *reinterpret_cast<uint4*>(lo + elem) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
struct CV { __nv_bfloat16* hi; __nv_bfloat16* lo; int N, H, W, chunks, c0, C; };
CV dev(const CView& v) { return CV{v.hi, v.lo, v.N, v.H, v.W, v.chunks, v.c0, v.C}; }
__device__ __forceinline__ long cv_elem(const CV& v, int n, int chunk, int y, int x) {
  return ((((long)n * v.chunks + v.c0 + chunk) * v.H + y) * v.W + x) * 8;
}
inline unsigned blocks_for(long total, int bs = 256) { return (unsigned)((total + bs - 1) / bs); }

// ---- basemodel.py:12-26: (x/255 - mean_bgr) / std_bgr, HWC -> CP8 (3 channels + 5 zeros) -------------
__global__ void __launch_bounds__(256) preprocess_kernel(const float* __restrict__ img, CV out) {
  const long hw = (long)out.H * out.W;
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hw) return;
  const float mean[3] = {0.406f, 0.456f, 0.485f}, stdv[3] = {0.225f, 0.224f, 0.229f};
  F8 f;
#pragma unroll
  for (int j = 0; j < 8; j++) f.v[j] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; c++) f.v[c] = __fdiv_rn(__fsub_rn(__fmul_rn(img[p * 3 + c], 1.0f / 255), mean[c]), stdv[c]);
  st_chunk(out.hi, out.lo, ((long)out.c0 * hw + p) * 8, f);
}

// Row im2col of the stem (see common.cuh): one thread per (pixel of the half-width grid, chunk of 8 of the 21 + 3 channels)
__global__ void __launch_bounds__(256) preprocess_rows7_kernel(const float* __restrict__ img, int W, int pad_l, CV out) {
  const long hw = (long)out.H * out.W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= hw * 3) return;
  const int ch = (int)(idx / hw);
  const long p = idx - (long)ch * hw;
  const int y = (int)(p / out.W), xo = (int)(p - (long)y * out.W);
  const float mean[3] = {0.406f, 0.456f, 0.485f}, stdv[3] = {0.225f, 0.224f, 0.229f};
  F8 f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = ch * 8 + j, s = k / 3, c = k - s * 3, x = 2 * xo + s - pad_l;
    f.v[j] = 0.f;
    if (k < 21 && x >= 0 && x < W) f.v[j] = __fdiv_rn(__fsub_rn(__fmul_rn(img[((long)y * W + x) * 3 + c], 1.0f / 255), mean[c]), stdv[c]);
  }
  st_chunk(out.hi, out.lo, (((long)(out.c0 + ch)) * hw + p) * 8, f);
}

// uint8 frame -> the fp32 image tensor the graph is fed with (eval.py:75-78 hands pred_func the resized uint8 image)
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, long n,
                                                        int aligned) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (aligned && i + 4 <= n) {
    const uchar4 u = *reinterpret_cast<const uchar4*>(src + i);
    *reinterpret_cast<float4*>(dst + i) = make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
  } else {
    for (long j = i; j < n && j < i + 4; j++) dst[j] = (float)src[j];
  }
}

// ---- basemodel.py:81-82: pad (0,1) zeros, 3x3 stride-2 VALID max pooling -------------------------------
__global__ void __launch_bounds__(256) maxpool_kernel(CV in, CV out) {
  const int nch = (in.C + 7) / 8;
  const long total = (long)out.N * nch * out.H * out.W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % out.W), y = (int)((idx / out.W) % out.H);
  const int ch = (int)((idx / ((long)out.W * out.H)) % nch), n = (int)(idx / ((long)out.W * out.H * nch));
  F8 m;
#pragma unroll
  for (int j = 0; j < 8; j++) m.v[j] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; dy++)
#pragma unroll
    for (int dx = 0; dx < 3; dx++) {
      const int iy = 2 * y + dy, ix = 2 * x + dx;
      F8 v;
      if (iy < in.H && ix < in.W) v = ld_chunk(in.hi, in.lo, cv_elem(in, n, ch, iy, ix));
      else {
#pragma unroll
        for (int j = 0; j < 8; j++) v.v[j] = 0.f;  // the explicit zero padding takes part in the max
      }
#pragma unroll
      for (int j = 0; j < 8; j++) m.v[j] = fmaxf(m.v[j], v.v[j]);
    }
  st_chunk(out.hi, out.lo, cv_elem(out, n, ch, y, x), m);
}

// ---- RPN decode (model.py:114-139) ---------------------------------------------------------------------
// rpn: fp32 channels-last [fh*fw][cs]: channels [0,NA) logits, [NA + a*4 + k] box deltas.
__global__ void __launch_bounds__(256) rpn_decode_kernel(const float* __restrict__ rpn, int cs, int fh, int fw, int na,
                                                         const float* __restrict__ cell_anchors /*[na][4]*/, float stride,
                                                         float clip, float* __restrict__ scores, float* __restrict__ boxes) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)fh * fw * na;
  if (idx >= total) return;
  const int a = (int)(idx % na);
  const long pix = idx / na;
  const int x = (int)(pix % fw), y = (int)(pix / fw);
  const float* p = rpn + pix * cs;
  scores[idx] = p[a];
  const float tx = p[na + a * 4 + 0], ty = p[na + a * 4 + 1], tw = p[na + a * 4 + 2], th = p[na + a * 4 + 3];
  // data.py:35-74: integer-valued anchors shifted by the stride, x2/y2 + 1
  const float ax1 = cell_anchors[a * 4 + 0] + x * stride, ay1 = cell_anchors[a * 4 + 1] + y * stride;
  const float ax2 = cell_anchors[a * 4 + 2] + x * stride + 1.f, ay2 = cell_anchors[a * 4 + 3] + y * stride + 1.f;
  const float wa = __fsub_rn(ax2, ax1), ha = __fsub_rn(ay2, ay1);
  const float xa = __fmul_rn(__fadd_rn(ax2, ax1), 0.5f), ya = __fmul_rn(__fadd_rn(ay2, ay1), 0.5f);
  const float wb = __fmul_rn(expf(fminf(tw, clip)), wa), hb = __fmul_rn(expf(fminf(th, clip)), ha);
  const float xb = __fadd_rn(__fmul_rn(tx, wa), xa), yb = __fadd_rn(__fmul_rn(ty, ha), ya);
  float4 o;
  o.x = __fsub_rn(xb, __fmul_rn(wb, 0.5f));
  o.y = __fsub_rn(yb, __fmul_rn(hb, 0.5f));
  o.z = __fadd_rn(xb, __fmul_rn(wb, 0.5f));
  o.w = __fadd_rn(yb, __fmul_rn(hb, 0.5f));
  reinterpret_cast<float4*>(boxes)[idx] = o;
}

// ---- exact top-k: radix select on order-preserving keys, ordered compaction, bitonic sort ---------------
__device__ __forceinline__ uint32_t float_key(float f) {  // larger float -> larger key
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One block of 1024 threads.  Output: the k best (score desc, index asc) as sorted (index, score) pairs.
__global__ void __launch_bounds__(1024) topk_kernel(const float* __restrict__ scores, int n, int k, int* __restrict__ out_idx,
                                                    float* __restrict__ out_score, int* __restrict__ out_count) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_remaining;
  __shared__ unsigned long long keys[1024];
  __shared__ int s_pos;
  __shared__ int warp_sums[32];
  const int tid = threadIdx.x;
  k = min(k, n);
  k = min(k, 1024);
  // radix select: find the key T of the k-th largest element
  uint32_t prefix = 0, mask = 0;
  uint32_t remaining = (uint32_t)k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
      const uint32_t key = float_key(scores[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t rem = remaining;
      int d = 255;
      for (; d > 0; d--) {
        if (hist[d] >= rem) break;
        rem -= hist[d];
      }
      s_prefix = prefix | ((uint32_t)d << shift);
      s_remaining = rem;
    }
    __syncthreads();
    prefix = s_prefix;
    remaining = s_remaining;
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t T = prefix;  // exactly `remaining` elements with key == T belong to the top k (lowest indices first)
  // ordered compaction: keys > T all, keys == T the first `remaining` in index order
  if (tid == 0) s_pos = 0;
  keys[tid] = 0ull;
  __syncthreads();
  int eq_taken = 0;  // uniform across the block (recomputed from shared state each chunk)
  __shared__ int s_eq_taken;
  if (tid == 0) s_eq_taken = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    uint32_t key = 0;
    int gt = 0, eq = 0;
    if (i < n) {
      key = float_key(scores[i]);
      gt = key > T;
      eq = key == T;
    }
    // block-wide exclusive scan of eq (to honour "lowest index first") and of the selected flags
    eq_taken = s_eq_taken;
    const unsigned lane = tid & 31, wid = tid >> 5;
    const unsigned eq_ballot = __ballot_sync(0xffffffffu, eq);
    const int eq_before_warp = __popc(eq_ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(eq_ballot);
    __syncthreads();
    int eq_before = eq_before_warp;
    for (unsigned w2 = 0; w2 < wid; w2++) eq_before += warp_sums[w2];
    int eq_total = 0;
    for (unsigned w2 = 0; w2 < 32; w2++) eq_total += warp_sums[w2];
    __syncthreads();
    const int sel = gt || (eq && (eq_taken + eq_before) < (int)remaining);
    const unsigned sel_ballot = __ballot_sync(0xffffffffu, sel);
    const int sel_before_warp = __popc(sel_ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(sel_ballot);
    __syncthreads();
    int sel_before = sel_before_warp;
    for (unsigned w2 = 0; w2 < wid; w2++) sel_before += warp_sums[w2];
    int sel_total = 0;
    for (unsigned w2 = 0; w2 < 32; w2++) sel_total += warp_sums[w2];
    const int pos0 = s_pos;
    if (sel) keys[pos0 + sel_before] = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    __syncthreads();
    if (tid == 0) {
      s_pos = pos0 + sel_total;
      s_eq_taken = eq_taken + min(eq_total, (int)remaining - eq_taken);
    }
    __syncthreads();
  }
  // bitonic sort, descending, of 1024 composite keys (unused slots are 0 = smallest)
  for (int size = 2; size <= 1024; size <<= 1) {
    for (int strd = size >> 1; strd > 0; strd >>= 1) {
      const int j = tid ^ strd;
      if (j > tid) {
        const bool desc = (tid & size) == 0;
        const unsigned long long a = keys[tid], b = keys[j];
        if (desc ? (a < b) : (a > b)) { keys[tid] = b; keys[j] = a; }
      }
      __syncthreads();
    }
  }
  if (tid < k) {
    const int idx = (int)(0xffffffffu - (uint32_t)(keys[tid] & 0xffffffffull));
    out_idx[tid] = idx;
    out_score[tid] = scores[idx];
  }
  if (tid == 0) *out_count = k;
}

// gather + clip (model.py:18-27) + drop empty boxes (model.py:193-200), order preserved.  One block of 1024.
__global__ void __launch_bounds__(1024) gather_clip_valid_kernel(const float* __restrict__ boxes, const int* __restrict__ idx,
                                                                 const float* __restrict__ score, const int* __restrict__ count,
                                                                 float img_h, float img_w, float min_size, float* __restrict__ out_boxes,
                                                                 float* __restrict__ out_scores, int* __restrict__ out_src,
                                                                 int* __restrict__ out_count) {
  __shared__ int warp_sums[32];
  const int tid = threadIdx.x, n = *count;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  int valid = 0;
  if (tid < n) {
    b = reinterpret_cast<const float4*>(boxes)[idx[tid]];
    if (!isinf(img_w)) {  // clip_boxes (model.py:18-27); an infinite window means "no clipping" (generic NMS hook)
      b.x = fminf(fmaxf(b.x, 0.f), img_w); b.y = fminf(fmaxf(b.y, 0.f), img_h);
      b.z = fminf(fmaxf(b.z, 0.f), img_w); b.w = fminf(fmaxf(b.w, 0.f), img_h);
    }
    valid = (__fsub_rn(b.z, b.x) > min_size) && (__fsub_rn(b.w, b.y) > min_size);
  }
  const unsigned lane = tid & 31, wid = tid >> 5;
  const unsigned ballot = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) warp_sums[wid] = __popc(ballot);
  __syncthreads();
  int before = __popc(ballot & ((1u << lane) - 1u)), total = 0;
  for (unsigned w2 = 0; w2 < 32; w2++) {
    if (w2 < wid) before += warp_sums[w2];
    total += warp_sums[w2];
  }
  if (valid) {
    reinterpret_cast<float4*>(out_boxes)[before] = b;
    out_scores[before] = score[tid];
    out_src[before] = idx[tid];
  }
  if (tid == 0) *out_count = total;
}

// ---- NMS ---------------------------------------------------------------------------------------------------
// IoU exactly as tensorflow/core/kernels/non_max_suppression_op.cc computes it (fp32, no contraction).
__device__ __forceinline__ float tf_iou(const float4 a, const float4 b) {
  const float ymin_i = fminf(a.x, a.z), xmin_i = fminf(a.y, a.w), ymax_i = fmaxf(a.x, a.z), xmax_i = fmaxf(a.y, a.w);
  const float ymin_j = fminf(b.x, b.z), xmin_j = fminf(b.y, b.w), ymax_j = fmaxf(b.x, b.z), xmax_j = fmaxf(b.y, b.w);
  const float area_i = __fmul_rn(__fsub_rn(ymax_i, ymin_i), __fsub_rn(xmax_i, xmin_i));
  const float area_j = __fmul_rn(__fsub_rn(ymax_j, ymin_j), __fsub_rn(xmax_j, xmin_j));
  if (area_i <= 0.f || area_j <= 0.f) return 0.f;
  const float iy0 = fmaxf(ymin_i, ymin_j), ix0 = fmaxf(xmin_i, xmin_j), iy1 = fminf(ymax_i, ymax_j), ix1 = fminf(xmax_i, xmax_j);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(iy1, iy0), 0.f), fmaxf(__fsub_rn(ix1, ix0), 0.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_j), inter));
}

// boxes sorted by descending score (any consistent corner layout).  mask[i][w] bit b: box j = 32w+b (j > i) has IoU > thr with i.
__global__ void __launch_bounds__(256) nms_mask_kernel(const float* __restrict__ boxes, const int* __restrict__ count, float thr,
                                                       uint32_t* __restrict__ mask /*[1024][32]*/) {
  const int n = *count;
  const int i = blockIdx.x;
  if (i >= n) return;
  const float4 bi = reinterpret_cast<const float4*>(boxes)[i];
  for (int w = threadIdx.x >> 5; w < 32; w += 8) {
    const int j = w * 32 + (threadIdx.x & 31);
    int sup = 0;
    if (j < n && j > i) sup = tf_iou(bi, reinterpret_cast<const float4*>(boxes)[j]) > thr;
    const unsigned bal = __ballot_sync(0xffffffffu, sup);
    if ((threadIdx.x & 31) == 0) mask[i * 32 + w] = bal;
  }
}

// one warp: greedy scan over the bitmask; lane l owns the "removed" word l.
__global__ void __launch_bounds__(32) nms_scan_kernel(const uint32_t* __restrict__ mask, const int* __restrict__ count, int max_out,
                                                      int* __restrict__ keep, int* __restrict__ keep_count) {
  const int n = *count, lane = threadIdx.x;
  uint32_t removed = 0;
  int kept = 0;
  for (int i = 0; i < n && kept < max_out; i++) {
    const uint32_t word = __shfl_sync(0xffffffffu, removed, i >> 5);
    if (!((word >> (i & 31)) & 1u)) {
      if (lane == 0) keep[kept] = i;
      kept++;
      removed |= mask[i * 32 + lane];
    }
  }
  if (lane == 0) *keep_count = kept;
}

// proposals = gather(valid boxes, keep) (model.py:211-216); rows beyond the count are zero-filled
__global__ void __launch_bounds__(128) gather_proposals_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                               const int* __restrict__ keep, const int* __restrict__ keep_count,
                                                               int max_out, float* __restrict__ out_boxes, float* __restrict__ out_scores) {
  const int t = threadIdx.x;
  if (t >= max_out) return;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  float s = 0.f;
  if (t < *keep_count) {
    b = reinterpret_cast<const float4*>(boxes)[keep[t]];
    s = scores[keep[t]];
  }
  reinterpret_cast<float4*>(out_boxes)[t] = b;
  out_scores[t] = s;
}

// ---- Mask R-CNN mask head tail (model.py:495-509, train.py:297-309) ------------------------------------------------------
// d = ReLU(Deconv2D 2x2 stride 2) kept in its GEMM form: CP8 [N, 4*256, 7, 7] with channel (dy*2+dx)*256 + co (a 2x2 stride-2
// transposed convolution has no overlap: out[2y+dy, 2x+dx, co] = in[y, x, :] . W[dy, dx, co, :]).  The 1x1 convolution to the
// single class-agnostic category acts per output pixel, so it is applied per phase here, followed by the sigmoid:
// masks[n, 2y+dy, 2x+dx] = sigmoid(b + sum_co d[n, (dy*2+dx)*256 + co, y, x] * w[co]).  One warp per output pixel.
__global__ void __launch_bounds__(256) mask_head_kernel(CV d, const float* __restrict__ w /*[256]*/, const float* __restrict__ b,
                                                        float* __restrict__ masks /*[N][2H][2W]*/) {
  const int lane = threadIdx.x & 31;
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int OH = 2 * d.H, OW = 2 * d.W;
  const long total = (long)d.N * OH * OW;
  if (wid >= total) return;
  const int ox = (int)(wid % OW), oy = (int)((wid / OW) % OH), n = (int)(wid / ((long)OW * OH));
  const int phase = (oy & 1) * 2 + (ox & 1);
  const F8 v = ld_chunk(d.hi, d.lo, cv_elem(d, n, phase * 32 + lane, oy >> 1, ox >> 1));   // 32 chunks = 256 channels per phase
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) acc = fmaf(v.v[j], __ldg(w + lane * 8 + j), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) masks[wid] = 1.f / (1.f + expf(-(acc + __ldg(b))));
}

// ---- fill_full_mask (eval.py:35-58) for all boxes of an image ---------------------------------------------------------------
// canvas[n, y, x] = (cv2.resize(mask[n] (MxM float32), (w, h)) > 0.5) inside the box's integer rectangle, 0 outside.
// OpenCV's float32 INTER_LINEAR arithmetic (coordinates (float)((d + 0.5) * scale - 0.5) with a double scale, x clamps the
// fraction, y clips the row indices; products and sums rounded one by one), INTER_AREA's 2x2 mean for an exact 2x down-scale,
// a plain copy for w = h = M.  One thread per canvas pixel: warp-level bilinear taps of a 784-byte mask that stays in L1.
__device__ __forceinline__ void cv_axis(int d, int dst_n, int src_n, bool clamp_fraction, int* s0, int* s1, float* f) {
  const double scale = 1.0 / ((double)dst_n / (double)src_n);
  float fx = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);   // no FMA contraction: the host rounds the product first
  int s = (int)floorf(fx);
  fx -= (float)s;
  if (clamp_fraction) {
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= src_n - 1) { s = src_n - 1; fx = 0.f; }
    *s0 = s; *s1 = min(s + 1, src_n - 1);
  } else {
    *s0 = max(0, min(src_n - 1, s)); *s1 = max(0, min(src_n - 1, s + 1));
  }
  *f = fx;
}

__global__ void __launch_bounds__(256) fill_full_mask_kernel(const float* __restrict__ masks, const float* __restrict__ boxes, int M, int H, int W,
                                                             unsigned char* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (x >= W) return;
  const float4 b = reinterpret_cast<const float4*>(boxes)[n];
  const int x0 = (int)__fadd_rn(b.x, 0.5f), y0 = (int)__fadd_rn(b.y, 0.5f);      // int() truncates
  const int x1 = max(x0, (int)__fsub_rn(b.z, 0.5f)), y1 = max(y0, (int)__fsub_rn(b.w, 0.5f));
  unsigned char v = 0;
  if (x >= x0 && x <= x1 && y >= y0 && y <= y1) {
    const int w = x1 + 1 - x0, h = y1 + 1 - y0, dx = x - x0, dy = y - y0;
    const float* m = masks + (long)n * M * M;
    float r;
    if (w == M && h == M) {
      r = __ldg(m + dy * M + dx);
    } else if (M == 2 * w && M == 2 * h) {
      const float* p = m + (2 * dy) * M + 2 * dx;
      r = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(__ldg(p), __ldg(p + 1)), __ldg(p + M)), __ldg(p + M + 1)), 0.25f);
    } else {
      int sx0, sx1, sy0, sy1; float fx, fy;
      cv_axis(dx, w, M, true, &sx0, &sx1, &fx);
      cv_axis(dy, h, M, false, &sy0, &sy1, &fy);
      const float a0 = __fsub_rn(1.f, fx), b0 = __fsub_rn(1.f, fy);
      const float r0 = __fadd_rn(__fmul_rn(__ldg(m + sy0 * M + sx0), a0), __fmul_rn(__ldg(m + sy0 * M + sx1), fx));
      const float r1 = __fadd_rn(__fmul_rn(__ldg(m + sy1 * M + sx0), a0), __fmul_rn(__ldg(m + sy1 * M + sx1), fx));
      r = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, fy));
    }
    v = r > 0.5f ? 1 : 0;
  }
  out[((long)n * H + y) * W + x] = v;
}

// ---- RoIAlign (model.py:301-374) ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) roi_align_kernel(CV fm, const float* __restrict__ rois /*[N][4] x1y1x2y2 image coords*/,
                                                        float spatial_scale, int out_size, CV out) {
  const int nch = (fm.C + 7) / 8;
  const long total = (long)out.N * nch * out_size * out_size;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % out_size), oy = (int)((idx / out_size) % out_size);
  const int ch = (int)((idx / ((long)out_size * out_size)) % nch), n = (int)(idx / ((long)out_size * out_size * nch));
  const float4 r = reinterpret_cast<const float4*>(rois)[n];
  const float x0 = __fmul_rn(r.x, spatial_scale), y0 = __fmul_rn(r.y, spatial_scale);
  const float x1 = __fmul_rn(r.z, spatial_scale), y1 = __fmul_rn(r.w, spatial_scale);
  const int crop = out_size * 2;
  const float H1 = (float)(fm.H - 1), W1 = (float)(fm.W - 1);
  // transform_fpcoor_for_tf (model.py:316-344), then crop_and_resize_op.cc's coordinate arithmetic
  const float sw = __fdiv_rn(__fsub_rn(x1, x0), (float)crop), sh = __fdiv_rn(__fsub_rn(y1, y0), (float)crop);
  const float nx0 = __fdiv_rn(__fsub_rn(__fadd_rn(x0, __fdiv_rn(sw, 2.f)), 0.5f), W1);
  const float ny0 = __fdiv_rn(__fsub_rn(__fadd_rn(y0, __fdiv_rn(sh, 2.f)), 0.5f), H1);
  const float nw = __fdiv_rn(__fmul_rn(sw, (float)(crop - 1)), W1), nh = __fdiv_rn(__fmul_rn(sh, (float)(crop - 1)), H1);
  const float bx1 = nx0, by1 = ny0, bx2 = __fadd_rn(nx0, nw), by2 = __fadd_rn(ny0, nh);
  const float hs = __fdiv_rn(__fmul_rn(__fsub_rn(by2, by1), H1), (float)(crop - 1));
  const float ws = __fdiv_rn(__fmul_rn(__fsub_rn(bx2, bx1), W1), (float)(crop - 1));
  F8 acc;
#pragma unroll
  for (int j = 0; j < 8; j++) acc.v[j] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; dy++) {
    const float in_y = __fadd_rn(__fmul_rn(by1, H1), __fmul_rn((float)(2 * oy + dy), hs));
    const bool yv = in_y >= 0.f && in_y <= H1;
    const int ty = yv ? (int)floorf(in_y) : 0, byy = yv ? (int)ceilf(in_y) : 0;
    const float yl = __fsub_rn(in_y, (float)ty);
#pragma unroll
    for (int dx = 0; dx < 2; dx++) {
      const float in_x = __fadd_rn(__fmul_rn(bx1, W1), __fmul_rn((float)(2 * ox + dx), ws));
      const bool xv = in_x >= 0.f && in_x <= W1;
      if (!(yv && xv)) continue;  // extrapolation value 0
      const int lx = (int)floorf(in_x), rx = (int)ceilf(in_x);
      const float xl = __fsub_rn(in_x, (float)lx);
      const F8 tl = ld_chunk(fm.hi, fm.lo, cv_elem(fm, 0, ch, ty, lx)), tr = ld_chunk(fm.hi, fm.lo, cv_elem(fm, 0, ch, ty, rx));
      const F8 bl = ld_chunk(fm.hi, fm.lo, cv_elem(fm, 0, ch, byy, lx)), br = ld_chunk(fm.hi, fm.lo, cv_elem(fm, 0, ch, byy, rx));
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float top = __fadd_rn(tl.v[j], __fmul_rn(__fsub_rn(tr.v[j], tl.v[j]), xl));
        const float bot = __fadd_rn(bl.v[j], __fmul_rn(__fsub_rn(br.v[j], bl.v[j]), xl));
        acc.v[j] = __fadd_rn(acc.v[j], __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl)));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) acc.v[j] = __fmul_rn(acc.v[j], 0.25f);
  st_chunk(out.hi, out.lo, cv_elem(out, n, ch, oy, ox), acc);
}

// ---- global average pool + fully connected (model.py:378-395, 552-565) ----------------------------------------
// one block per RoI; feat CP8 [N][C/8][h][w]; W [C][nout] (tensorpack FullyConnected layout), out [N][nout]
__global__ void __launch_bounds__(256) gap_fc_kernel(CV feat, const float* __restrict__ Wt, const float* __restrict__ bias, int nout,
                                                     float* __restrict__ pooled_out /*[N][C] or null*/, float* __restrict__ out) {
  extern __shared__ float pooled[];  // [C]
  const int n = blockIdx.x, C = feat.C, nch = (C + 7) / 8, hw = feat.H * feat.W;
  for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
    F8 s;
#pragma unroll
    for (int j = 0; j < 8; j++) s.v[j] = 0.f;
    for (int p = 0; p < hw; p++) {
      const F8 v = ld_chunk(feat.hi, feat.lo, (((long)n * feat.chunks + feat.c0 + ch) * hw + p) * 8);
#pragma unroll
      for (int j = 0; j < 8; j++) s.v[j] += v.v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (ch * 8 + j < C) {
        const float m = s.v[j] / (float)hw;
        pooled[ch * 8 + j] = m;
        if (pooled_out) pooled_out[(long)n * C + ch * 8 + j] = m;
      }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = wid; o < nout; o += nw) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(pooled[c], Wt[(long)c * nout + o], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[(long)n * nout + o] = acc + bias[o];
  }
}

// ---- Fast R-CNN inference tail, class-agnostic (train.py:275-295, model.py:439-491) ----------------------------
// logits [N][nfc]: [0,2) class logits, [2,6) box deltas, [6, 6+nsecond) second-head logits.  One block of 128 threads.
struct TailArgs {
  const float* logits; int nfc, nsecond;
  const float* prop_boxes; const int* prop_count;
  float img_h, img_w, clip, score_thresh, nms_thresh;
  int max_rois, results_per_im;
  float* all_probs;   // [max_rois][2]
  float* all_boxes;   // [max_rois][4]
  float* second_probs;  // [max_rois][nsecond]
  // outputs
  int* n_out; float* final_boxes; float* final_probs; long long* final_labels; float* final_posterior;
  long long* second_final_labels; float* second_final_posterior; int* final_box_index;
};

__global__ void __launch_bounds__(128) frcnn_tail_kernel(TailArgs a) {
  __shared__ float s_prob[128];
  __shared__ float4 s_box[128];
  __shared__ int s_order[128];
  __shared__ int s_ncand;
  __shared__ unsigned char s_keep[128];
  const int t = threadIdx.x, n = min(*a.prop_count, a.max_rois);
  float prob = 0.f;
  float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < n) {
    const float* l = a.logits + (long)t * a.nfc;
    // softmax over the two class logits (max-subtracted, like tf.nn.softmax)
    const float m = fmaxf(l[0], l[1]);
    const float e0 = expf(l[0] - m), e1 = expf(l[1] - m), s = e0 + e1;
    const float p0 = e0 / s, p1 = e1 / s;
    a.all_probs[t * 2] = p0; a.all_probs[t * 2 + 1] = p1;
    prob = p1;
    // decode_bbox_target(box_logits / [10,10,5,5], proposal box), clip
    const float4 pb = reinterpret_cast<const float4*>(a.prop_boxes)[t];
    const float tx = __fdiv_rn(l[2], 10.f), ty = __fdiv_rn(l[3], 10.f), tw = __fdiv_rn(l[4], 5.f), th = __fdiv_rn(l[5], 5.f);
    const float wa = __fsub_rn(pb.z, pb.x), ha = __fsub_rn(pb.w, pb.y);
    const float xa = __fmul_rn(__fadd_rn(pb.z, pb.x), 0.5f), ya = __fmul_rn(__fadd_rn(pb.w, pb.y), 0.5f);
    const float wb = __fmul_rn(expf(fminf(tw, a.clip)), wa), hb = __fmul_rn(expf(fminf(th, a.clip)), ha);
    const float xb = __fadd_rn(__fmul_rn(tx, wa), xa), yb = __fadd_rn(__fmul_rn(ty, ha), ya);
    box.x = fminf(fmaxf(__fsub_rn(xb, __fmul_rn(wb, 0.5f)), 0.f), a.img_w);
    box.y = fminf(fmaxf(__fsub_rn(yb, __fmul_rn(hb, 0.5f)), 0.f), a.img_h);
    box.z = fminf(fmaxf(__fadd_rn(xb, __fmul_rn(wb, 0.5f)), 0.f), a.img_w);
    box.w = fminf(fmaxf(__fadd_rn(yb, __fmul_rn(hb, 0.5f)), 0.f), a.img_h);
    reinterpret_cast<float4*>(a.all_boxes)[t] = box;
    // second head softmax
    if (a.nsecond > 0) {
      const float* s2 = l + 6;
      float mx = -INFINITY;
      for (int j = 0; j < a.nsecond; j++) mx = fmaxf(mx, s2[j]);
      float sum = 0.f;
      for (int j = 0; j < a.nsecond; j++) sum += expf(s2[j] - mx);
      for (int j = 0; j < a.nsecond; j++) a.second_probs[(long)t * a.nsecond + j] = expf(s2[j] - mx) / sum;
    }
  }
  s_prob[t] = prob;
  s_box[t] = box;
  s_keep[t] = 0;
  if (t == 0) s_ncand = 0;
  __syncthreads();
  // candidates: prob > thresh, ranked by (prob desc, index asc)
  const bool cand = t < n && prob > a.score_thresh;
  if (cand) {
    int rank = 0;
    for (int j = 0; j < n; j++) {
      const float pj = s_prob[j];
      if (pj > a.score_thresh && (pj > prob || (pj == prob && j < t))) rank++;
    }
    s_order[rank] = t;
    atomicAdd(&s_ncand, 1);
  }
  __syncthreads();
  if (t == 0) {  // greedy NMS over <= 100 candidates
    int kept = 0;
    int kept_idx[32];
    for (int c = 0; c < s_ncand && kept < a.results_per_im && kept < 32; c++) {
      const int i = s_order[c];
      bool ok = true;
      for (int k2 = kept - 1; k2 >= 0; k2--)
        if (tf_iou(s_box[i], s_box[kept_idx[k2]]) > a.nms_thresh) { ok = false; break; }
      if (ok) { kept_idx[kept++] = i; s_keep[i] = 1; }
    }
  }
  __syncthreads();
  // results in ascending proposal index (the reference's top_k(sorted=False) order is unspecified)
  if (t < n && s_keep[t]) {
    int pos = 0;
    for (int j = 0; j < t; j++) pos += s_keep[j];
    reinterpret_cast<float4*>(a.final_boxes)[pos] = box;
    a.final_probs[pos] = prob;
    a.final_labels[pos] = 1;
    a.final_box_index[pos] = t;
    // train.py:287-288: label_probs gathered with the CATEGORY index (0 when class-agnostic) -- kept as is
    const float q0 = a.all_probs[0], q1 = a.all_probs[1];
    a.final_posterior[pos * 2] = q0; a.final_posterior[pos * 2 + 1] = q1;
    a.second_final_labels[pos] = (q1 > q0) ? 2 : 1;  // argmax(final_posterior) + 1 (first max wins)
    for (int j = 0; j < a.nsecond; j++) a.second_final_posterior[(long)pos * a.nsecond + j] = a.second_probs[j];
  }
  if (t == 0) {
    int tot = 0;
    for (int j = 0; j < n; j++) tot += s_keep[j];
    *a.n_out = tot;
  }
}

}  // namespace

// ---- launchers ------------------------------------------------------------------------------------------------
int det_preprocess(const float* img_hwc, const CView& out, cudaStream_t st) {
  const long hw = (long)out.H * out.W;
  prof_before(st);
  preprocess_kernel<<<blocks_for(hw), 256, 0, st>>>(img_hwc, dev(out));
  return after_launch("preprocess_kernel", st, 0.0, (double)hw * (12.0 + 32.0));
}

int det_preprocess_rows7(const float* img_hwc, int W, int pad_l, const CView& out, cudaStream_t st) {
  PV_CHECK(out.C == 24 && out.N == 1, PREMVOS_ERR_INVALID_ARG, "det_preprocess_rows7: the output view is one image of 3 chunks");
  const long total = (long)out.H * out.W * 3;
  prof_before(st);
  preprocess_rows7_kernel<<<blocks_for(total), 256, 0, st>>>(img_hwc, W, pad_l, dev(out));
  return after_launch("preprocess_kernel", st, 0.0, (double)out.H * W * 12.0 + (double)total * 32.0);
}

int det_u8_to_f32(const unsigned char* src, float* dst, long n, cudaStream_t st) {
  const int aligned = ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 15) == 0;
  prof_before(st);
  u8_to_f32_kernel<<<blocks_for((n + 3) / 4), 256, 0, st>>>(src, dst, n, aligned);
  return after_launch("u8_to_f32_kernel", st, 0.0, 5.0 * n);
}

int det_maxpool3x3s2(const CView& in, const CView& out, cudaStream_t st) {
  PV_CHECK(out.H == (in.H + 1 - 3) / 2 + 1 && out.W == (in.W + 1 - 3) / 2 + 1 && out.C == in.C && out.N == in.N,
           PREMVOS_ERR_INVALID_ARG, "det_maxpool3x3s2: shape mismatch");
  const long total = (long)out.N * out.vchunks() * out.H * out.W;
  prof_before(st);
  maxpool_kernel<<<blocks_for(total), 256, 0, st>>>(dev(in), dev(out));
  return after_launch("maxpool_kernel", st, 0.0, 4.0 * ((double)in.pixels() + (double)out.pixels()) * in.C);
}

int det_rpn_decode(const float* rpn, int cs, int fh, int fw, int na, const float* cell_anchors, float stride, float clip,
                   float* scores, float* boxes, cudaStream_t st) {
  const long total = (long)fh * fw * na;
  prof_before(st);
  rpn_decode_kernel<<<blocks_for(total), 256, 0, st>>>(rpn, cs, fh, fw, na, cell_anchors, stride, clip, scores, boxes);
  return after_launch("rpn_decode_kernel", st, 20.0 * total, 40.0 * total);
}

int det_topk(const float* scores, int n, int k, int* out_idx, float* out_score, int* out_count, cudaStream_t st) {
  PV_CHECK(k >= 1 && k <= 1024 && n >= 1, PREMVOS_ERR_INVALID_ARG, "det_topk: k=%d (max 1024), n=%d", k, n);
  prof_before(st);
  topk_kernel<<<1, 1024, 0, st>>>(scores, n, k, out_idx, out_score, out_count);
  return after_launch("topk_kernel", st, 0.0, 4.0 * 6 * n);
}

int det_gather_clip_valid(const float* boxes, const int* idx, const float* score, const int* count, float img_h, float img_w,
                          float min_size, float* out_boxes, float* out_scores, int* out_src, int* out_count, cudaStream_t st) {
  prof_before(st);
  gather_clip_valid_kernel<<<1, 1024, 0, st>>>(boxes, idx, score, count, img_h, img_w, min_size, out_boxes, out_scores, out_src, out_count);
  return after_launch("gather_clip_valid_kernel", st, 0.0, 1024.0 * 40);
}

int det_nms(const float* boxes_sorted, const int* count, float thr, int max_out, uint32_t* mask_scratch, int* keep, int* keep_count,
            cudaStream_t st) {
  prof_before(st);
  nms_mask_kernel<<<1024, 256, 0, st>>>(boxes_sorted, count, thr, mask_scratch);
  PV_TRY(after_launch("nms_mask_kernel", st, 1024.0 * 1024 * 20, 1024.0 * 144));
  prof_before(st);
  nms_scan_kernel<<<1, 32, 0, st>>>(mask_scratch, count, max_out, keep, keep_count);
  return after_launch("nms_scan_kernel", st, 0.0, 1024.0 * 128);
}

int det_gather_proposals(const float* boxes, const float* scores, const int* keep, const int* keep_count, int max_out,
                         float* out_boxes, float* out_scores, cudaStream_t st) {
  PV_CHECK(max_out <= 128, PREMVOS_ERR_INVALID_ARG, "det_gather_proposals: max_out %d > 128", max_out);
  prof_before(st);
  gather_proposals_kernel<<<1, 128, 0, st>>>(boxes, scores, keep, keep_count, max_out, out_boxes, out_scores);
  return after_launch("gather_proposals_kernel", st, 0.0, 128.0 * 40);
}

int det_roi_align(const CView& fm, const float* rois, float spatial_scale, int out_size, const CView& out, cudaStream_t st) {
  PV_CHECK(fm.N == 1 && out.C == fm.C && out.H == out_size && out.W == out_size, PREMVOS_ERR_INVALID_ARG, "det_roi_align: shape mismatch");
  const long total = (long)out.N * fm.vchunks() * out_size * out_size;
  prof_before(st);
  roi_align_kernel<<<blocks_for(total), 256, 0, st>>>(dev(fm), rois, spatial_scale, out_size, dev(out));
  return after_launch("roi_align_kernel", st, 30.0 * total * 8, 4.0 * ((double)fm.pixels() * fm.C + (double)out.pixels() * out.C));
}

int det_gap_fc(const CView& feat, const float* Wt, const float* bias, int nout, float* pooled_out, float* out, cudaStream_t st) {
  prof_before(st);
  gap_fc_kernel<<<feat.N, 256, round_up(feat.C, 8) * sizeof(float), st>>>(dev(feat), Wt, bias, nout, pooled_out, out);
  return after_launch("gap_fc_kernel", st, 2.0 * feat.N * feat.C * nout, 4.0 * ((double)feat.pixels() * feat.C + (double)feat.C * nout));
}

int det_mask_head(const CView& d, const float* w, const float* b, float* masks, cudaStream_t st) {
  PV_CHECK(d.C == 1024, PREMVOS_ERR_INVALID_ARG, "det_mask_head: expected the 4 x 256 phase channels of the 2x2 deconvolution");
  const long total = (long)d.N * 4 * d.H * d.W;
  prof_before(st);
  mask_head_kernel<<<blocks_for(total * 32), 256, 0, st>>>(dev(d), w, b, masks);
  return after_launch("mask_head_kernel", st, 2.0 * total * 256, 4.0 * ((double)d.pixels() * d.C + total));
}

int det_fill_full_masks(const float* masks, const float* boxes, int n, int M, int H, int W, unsigned char* out, cudaStream_t st) {
  if (n == 0) return 0;
  prof_before(st);
  fill_full_mask_kernel<<<dim3((W + 255) / 256, H, n), 256, 0, st>>>(masks, boxes, M, H, W, out);
  return after_launch("fill_full_mask_kernel", st, 0.0, (double)n * ((double)H * W + 4.0 * M * M));
}

int det_frcnn_tail(const DetTailArgs& h, cudaStream_t st) {
  PV_CHECK(h.max_rois <= 128 && h.results_per_im <= 32, PREMVOS_ERR_INVALID_ARG, "det_frcnn_tail: limits are 128 RoIs / 32 results");
  TailArgs a;
  a.logits = h.logits; a.nfc = h.nfc; a.nsecond = h.nsecond; a.prop_boxes = h.prop_boxes; a.prop_count = h.prop_count;
  a.img_h = h.img_h; a.img_w = h.img_w; a.clip = h.clip; a.score_thresh = h.score_thresh; a.nms_thresh = h.nms_thresh;
  a.max_rois = h.max_rois; a.results_per_im = h.results_per_im;
  a.all_probs = h.all_probs; a.all_boxes = h.all_boxes; a.second_probs = h.second_probs;
  a.n_out = h.n_out; a.final_boxes = h.final_boxes; a.final_probs = h.final_probs; a.final_labels = (long long*)h.final_labels;
  a.final_posterior = h.final_posterior; a.second_final_labels = (long long*)h.second_final_labels;
  a.second_final_posterior = h.second_final_posterior; a.final_box_index = h.final_box_index;
  prof_before(st);
  frcnn_tail_kernel<<<1, 128, 0, st>>>(a);
  return after_launch("frcnn_tail_kernel", st, 0.0, 128.0 * 400);
}

}  // namespace premvos
