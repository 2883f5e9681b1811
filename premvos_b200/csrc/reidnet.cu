// ReID network forward: the 128-d appearance embedding MergeTrack attaches to every proposal
// (MergeTrack/ReID_net_functions.py:26-45 `add_ReID`), a wide pre-activation ResNet on 128 x 128 box crops
// (ReID_net/configs/run:33-65: conv0, res0..res16, conv1 + max pool, fc1, fc2, outputTriplet).
//
// Restated as a fixed list of launches over pre-allocated buffers:
//   * crops: every box of a frame is cut out of ONE uploaded uint8 frame on the device (context region x1.2, tf.round, the
//     reference's clipping incl. its `maximum(excess, 1)`, TF1 legacy bilinear resize, ImageNet normalisation:
//     ReID_net/datasets/Similarity/DAVIS_Forward_Feed.py:34-120) -- the reference feeds the float frame once and crops inside
//     the graph with tf.map_fn;
//   * a ResidualUnit2 (ReID_net/network/NetworkLayers.py:157-210) is BN0 + ReLU (one elementwise kernel, F8 -> CP8), an optional
//     1x1 projection W0 of the activated input, then its convolutions on tcgen05 (conv_umma.cu): the BatchNorm + ReLU in front
//     of convolution i+1 is folded into convolution i's weights / bias / epilogue, the last convolution adds the residual in
//     its epilogue and writes the raw sum in F8 (it only ever feeds the next unit's BN0 and skip);
//   * conv1's max pool and the three BatchNorm + fully-connected layers (2000 -> 500 -> 500 -> 128) are one kernel per crop.
#include <math.h>
#include <stdlib.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "cp8.cuh"

using namespace premvos;
using namespace premvos::cp8;

namespace {

const float REID_BN_EPS = 1e-5f;
const int REID_S = 128;

struct UnitSpec { const char* name; int n; int feats[3]; int ks[3]; int strides[3]; };
const UnitSpec REID_UNITS[] = {
    {"res0", 2, {128, 128, 0}, {3, 3, 0}, {2, 1, 1}},   {"res1", 2, {128, 128, 0}, {3, 3, 0}, {1, 1, 1}},
    {"res2", 2, {128, 128, 0}, {3, 3, 0}, {1, 1, 1}},   {"res3", 2, {256, 256, 0}, {3, 3, 0}, {2, 1, 1}},
    {"res4", 2, {256, 256, 0}, {3, 3, 0}, {1, 1, 1}},   {"res5", 2, {256, 256, 0}, {3, 3, 0}, {1, 1, 1}},
    {"res6", 2, {512, 512, 0}, {3, 3, 0}, {2, 1, 1}},   {"res7", 2, {512, 512, 0}, {3, 3, 0}, {1, 1, 1}},
    {"res8", 2, {512, 512, 0}, {3, 3, 0}, {1, 1, 1}},   {"res9", 2, {512, 512, 0}, {3, 3, 0}, {1, 1, 1}},
    {"res10", 2, {512, 512, 0}, {3, 3, 0}, {1, 1, 1}},  {"res11", 2, {512, 512, 0}, {3, 3, 0}, {1, 1, 1}},
    {"res12", 2, {512, 1024, 0}, {3, 3, 0}, {1, 2, 1}}, {"res13", 2, {512, 1024, 0}, {3, 3, 0}, {1, 1, 1}},
    {"res14", 2, {512, 1024, 0}, {3, 3, 0}, {1, 1, 1}}, {"res15", 3, {512, 1024, 2048}, {1, 3, 1}, {1, 2, 1}},
    {"res16", 3, {1024, 2048, 4096}, {1, 3, 1}, {1, 1, 1}}};

typedef std::function<int(cudaStream_t, int)> Step;

// ---- kernels ----------------------------------------------------------------------------------------------------------------
struct CropArgs {
  const unsigned char* frame; int H, W;   // uint8 RGB frame
  const float* boxes;                      // [N][4] x, y, w, h (proposal 'bbox')
  int N;
  CV out;                                  // [N][1 chunk][128][128]: ch 0..2 normalised RGB
  int* crops;                              // [N][4] x y w h after context region + clipping (test hook)
};

__device__ __forceinline__ void legacy_axis(int o, float scale, int in_size, int* lo, int* hi, float* lerp) {
  const float src = __fmul_rn((float)o, scale);
  const int l = (int)floorf(src);
  *lo = l;
  *hi = min(l + 1, in_size - 1);
  *lerp = __fsub_rn(src, (float)l);
}

// DAVIS_Forward_Feed.py:36-58 in float32 / int32 as TensorFlow evaluates it
__device__ __forceinline__ void reid_crop_box(const float* b, int H, int W, int* x, int* y, int* w, int* h) {
  const float f = 1.2f, fm1 = (float)(1.2 - 1.0);   // python evaluates `factor - 1.0` in double before it meets the float32 tensor
  float xs = __fsub_rn(b[0], __fmul_rn(__fmul_rn(0.5f, b[2]), fm1));
  float ys = __fsub_rn(b[1], __fmul_rn(__fmul_rn(0.5f, b[3]), fm1));
  float ws = __fmul_rn(b[2], f), hs = __fmul_rn(b[3], f);
  int xi = (int)rintf(xs), yi = (int)rintf(ys), wi = (int)rintf(ws), hi = (int)rintf(hs);   // tf.round: half to even
  xi = max(xi, 0); yi = max(yi, 0);
  wi -= max(xi + wi - W, 1);   // sic
  hi -= max(yi + hi - H, 1);
  *x = xi; *y = yi; *w = wi; *h = hi;
}

__global__ void __launch_bounds__(256) reid_input_kernel(CropArgs a) {
  const long total = (long)a.N * REID_S * REID_S;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % REID_S), y = (int)((idx / REID_S) % REID_S), n = (int)(idx / ((long)REID_S * REID_S));
  int cx, cy, cw, ch;
  reid_crop_box(a.boxes + n * 4, a.H, a.W, &cx, &cy, &cw, &ch);
  if (x == 0 && y == 0) { a.crops[n * 4] = cx; a.crops[n * 4 + 1] = cy; a.crops[n * 4 + 2] = cw; a.crops[n * 4 + 3] = ch; }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  F8 f = zero8();
  if (min(cw, ch) > 10) {
    int ylo, yhi, xlo, xhi; float yl, xl;
    legacy_axis(y, __fdiv_rn((float)ch, (float)REID_S), ch, &ylo, &yhi, &yl);
    legacy_axis(x, __fdiv_rn((float)cw, (float)REID_S), cw, &xlo, &xhi, &xl);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      auto px = [&](int yy, int xx) { return __fdiv_rn((float)a.frame[((long)(cy + yy) * a.W + (cx + xx)) * 3 + c], 255.0f); };
      const float tl = px(ylo, xlo), tr = px(ylo, xhi), bl = px(yhi, xlo), br = px(yhi, xhi);
      const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl)), bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
      const float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
      f.v[c] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
    }
  } else {   // tf.cond(min_dim > 10, resize, zeros) -> normalize(zeros)
#pragma unroll
    for (int c = 0; c < 3; c++) f.v[c] = __fdiv_rn(__fsub_rn(0.f, mean[c]), stdv[c]);
  }
  st_chunk(a.out.hi, a.out.lo, cv_elem(a.out, n, 0, y, x), f);
}

// out (CP8) = ReLU(scale * in + shift) per channel, in = F8: the BN0 + ReLU at the head of a residual unit / of conv1
__global__ void __launch_bounds__(256) bn_relu_f8_kernel(const float* __restrict__ in, int in_chunks, int in_c0, CV out, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, long hw, int n_active) {
  const int nch = (out.C + 7) / 8;
  const long total = (long)n_active * nch * hw;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long p = idx % hw;
  const int ch = (int)((idx / hw) % nch), n = (int)(idx / (hw * nch));
  const float4* src = reinterpret_cast<const float4*>(in + (((long)n * in_chunks + in_c0 + ch) * hw + p) * 8);
  const float4 a0 = src[0], a1 = src[1];
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + ch * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + ch * 8) + 1);
  const float4 t0 = __ldg(reinterpret_cast<const float4*>(shift + ch * 8)), t1 = __ldg(reinterpret_cast<const float4*>(shift + ch * 8) + 1);
  F8 f;
  // (x - mean) * (gamma * rsqrt(var + eps)) + beta, rounded like tf.nn.batch_normalization: x * scale + (beta - mean * scale)
  f.v[0] = fmaxf(__fadd_rn(__fmul_rn(a0.x, s0.x), t0.x), 0.f); f.v[1] = fmaxf(__fadd_rn(__fmul_rn(a0.y, s0.y), t0.y), 0.f);
  f.v[2] = fmaxf(__fadd_rn(__fmul_rn(a0.z, s0.z), t0.z), 0.f); f.v[3] = fmaxf(__fadd_rn(__fmul_rn(a0.w, s0.w), t0.w), 0.f);
  f.v[4] = fmaxf(__fadd_rn(__fmul_rn(a1.x, s1.x), t1.x), 0.f); f.v[5] = fmaxf(__fadd_rn(__fmul_rn(a1.y, s1.y), t1.y), 0.f);
  f.v[6] = fmaxf(__fadd_rn(__fmul_rn(a1.z, s1.z), t1.z), 0.f); f.v[7] = fmaxf(__fadd_rn(__fmul_rn(a1.w, s1.w), t1.w), 0.f);
  st_chunk(out.hi, out.lo, (((long)n * out.chunks + out.c0 + ch) * hw + p) * 8, f);
}

// conv1's max pool (3x3 / 3, SAME on 4x4 -> 2x2: every window covers a 2x2 block) + flatten (h, w, c) + the three folded
// BatchNorm + fully-connected layers.  One CTA per crop; weights [in][out] so that the threads of a warp read consecutive floats.
struct HeadArgs {
  const float* conv1;   // fp32 channels-last [N][4][4][cs]
  int cs;
  const float *w1, *b1, *w2, *b2, *w3, *b3;   // BN folded: [2000][500], [500], [500][500], [500], [500][128], [128]
  float* out;           // [N][128]
};
__global__ void __launch_bounds__(256) reid_head_kernel(HeadArgs a) {
  __shared__ float x0[2000], x1[500], x2[500];
  const int n = blockIdx.x;
  const float* src = a.conv1 + (long)n * 16 * a.cs;
  for (int i = threadIdx.x; i < 2000; i += blockDim.x) {
    const int c = i % 500, q = i / 500, qy = q >> 1, qx = q & 1;
    float m = -INFINITY;
    for (int dy = 0; dy < 2; dy++)
      for (int dx = 0; dx < 2; dx++) m = fmaxf(m, src[((2 * qy + dy) * 4 + 2 * qx + dx) * a.cs + c]);
    x0[i] = m;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 500; o += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < 2000; i++) acc = fmaf(x0[i], __ldg(a.w1 + (long)i * 500 + o), acc);
    x1[o] = fmaxf(acc + a.b1[o], 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 500; o += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < 500; i++) acc = fmaf(x1[i], __ldg(a.w2 + (long)i * 500 + o), acc);
    x2[o] = fmaxf(acc + a.b2[o], 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 128; o += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < 500; i++) acc = fmaf(x2[i], __ldg(a.w3 + (long)i * 128 + o), acc);
    a.out[(long)n * 128 + o] = acc + a.b3[o];
  }
}

}  // namespace

struct premvos_reidnet {
  int NB = 0;
  bool finalized = false;
  std::map<std::string, std::vector<float>> params;
  std::map<std::string, std::vector<int64_t>> shapes;
  std::vector<void*> allocs;
  std::vector<std::unique_ptr<ConvWeightsUmma>> conv_weights;
  std::vector<std::unique_ptr<ConvPlanUmma>> conv_plans;
  ConvWorkspace conv_ws;
  std::vector<Step> steps;
  std::map<std::string, FView> named;   // test hook: raw unit outputs
  cudaStream_t stream = nullptr;
  CView input;
  TView conv1_out;
  float* emb_dev = nullptr;
  float* boxes_dev = nullptr; int* crops = nullptr;
  unsigned char* frame_dev = nullptr; size_t frame_cap = 0;
  int launches_per_forward = 0;
};

namespace {

void build_shape_table(premvos_reidnet* n) {
  auto bn = [&](const std::string& s, int c) {
    for (const char* v : {"beta", "gamma", "mean_ema", "var_ema"}) n->shapes[s + "/" + v] = {c};
  };
  n->shapes["conv0/W"] = {3, 3, 3, 64};
  int cin = 64;
  for (const UnitSpec& u : REID_UNITS) {
    const std::string s = u.name;
    bn(s + "/bn0", cin);
    int sres = 1;
    for (int i = 0; i < u.n; i++) sres *= u.strides[i];
    if (u.feats[u.n - 1] != cin || sres != 1) n->shapes[s + "/W0"] = {1, 1, cin, u.feats[u.n - 1]};
    int c = cin;
    for (int i = 0; i < u.n; i++) {
      if (i > 0) bn(s + "/bn" + std::to_string(i + 1), c);
      n->shapes[s + "/W" + std::to_string(i + 1)] = {u.ks[i], u.ks[i], c, u.feats[i]};
      c = u.feats[i];
    }
    cin = u.feats[u.n - 1];
  }
  bn("conv1/bn", cin);
  n->shapes["conv1/W"] = {3, 3, cin, 500};
  const int fc[3][2] = {{2000, 500}, {500, 500}, {500, 128}};
  const char* names[3] = {"fc1", "fc2", "outputTriplet"};
  for (int i = 0; i < 3; i++) {
    bn(std::string(names[i]) + "/bn", fc[i][0]);
    n->shapes[std::string(names[i]) + "/W"] = {fc[i][0], fc[i][1]};
    n->shapes[std::string(names[i]) + "/b"] = {fc[i][1]};
  }
}

int64_t numel_of(const std::vector<int64_t>& s) {
  int64_t k = 1;
  for (auto d : s) k *= d;
  return k;
}

template <typename T>
int dev_alloc(premvos_reidnet* n, T** p, size_t count) {
  PV_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  PV_CUDA(cudaMemset(*p, 0, count * sizeof(T)));
  n->allocs.push_back(*p);
  return 0;
}
int alloc_cview(premvos_reidnet* n, CView* v, int C, int H, int W) {
  v->N = n->NB; v->H = H; v->W = W; v->chunks = (C + 7) / 8; v->c0 = 0; v->C = C;
  const size_t elems = (size_t)v->N * v->chunks * H * W * 8 + 64;
  PV_TRY(dev_alloc(n, &v->hi, elems));
  PV_TRY(dev_alloc(n, &v->lo, elems));
  return 0;
}
int alloc_fview(premvos_reidnet* n, FView* v, int C, int H, int W) {
  v->N = n->NB; v->H = H; v->W = W; v->chunks = (C + 7) / 8; v->c0 = 0; v->C = C;
  PV_TRY(dev_alloc(n, &v->p, (size_t)v->N * v->chunks * H * W * 8 + 64));
  return 0;
}

void bn_fold(premvos_reidnet* n, const std::string& scope, std::vector<float>* scale, std::vector<float>* shift) {
  const std::vector<float>&g = n->params[scope + "/gamma"], &b = n->params[scope + "/beta"];
  const std::vector<float>&m = n->params[scope + "/mean_ema"], &v = n->params[scope + "/var_ema"];
  scale->resize(g.size()); shift->resize(g.size());
  for (size_t i = 0; i < g.size(); i++) {
    (*scale)[i] = g[i] / sqrtf(v[i] + REID_BN_EPS);
    (*shift)[i] = b[i] - m[i] * (*scale)[i];
  }
}

int upload(premvos_reidnet* n, const std::vector<float>& h, float** d, size_t padded = 0) {
  PV_TRY(dev_alloc(n, d, std::max(h.size(), padded)));
  PV_CUDA(cudaMemcpy(*d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

// BN + ReLU of an F8 tensor into a CP8 tensor (the activated input of a unit's convolutions)
int add_bn_relu(premvos_reidnet* n, const std::string& bn_scope, const FView& in, const CView& out) {
  std::vector<float> scale, shift;
  bn_fold(n, bn_scope, &scale, &shift);
  float *ds = nullptr, *dt = nullptr;
  const size_t padded = (size_t)round_up(in.C, 8);
  PV_TRY(upload(n, scale, &ds, padded)); PV_TRY(upload(n, shift, &dt, padded));
  const FView i2 = in; const CView o2 = out;
  n->steps.push_back([=](cudaStream_t st, int na) {
    const long hw = (long)i2.H * i2.W, total = (long)na * ((o2.C + 7) / 8) * hw;
    if (total == 0) return 0;
    prof_before(st);
    bn_relu_f8_kernel<<<blocks_for(total), 256, 0, st>>>(i2.p, i2.chunks, i2.c0, dev(o2), ds, dt, hw, na);
    return after_launch("bn_relu_f8_kernel", st, 3.0 * total * 8, 64.0 * total);
  });
  return 0;
}

// TensorFlow SAME padding of one axis: total = max((ceil(n / s) - 1) * s + k - n, 0), the odd pixel goes to the end
void same_pad(int nin, int k, int s, int* before, int* after) {
  const int out = (nin + s - 1) / s, total = std::max((out - 1) * s + k - nin, 0);
  *before = total / 2; *after = total - total / 2;
}

// convolution W (HWIO, no bias) with an optional BatchNorm + ReLU folded in BEHIND it (scale into the kernel's output channels,
// shift as bias, ReLU in the epilogue)
int add_conv(premvos_reidnet* n, const std::string& wname, const std::string& bn_after, const CView& in, const ConvOut& out, int stride,
             const int* cin_map = nullptr, int cin_phys = 0) {
  const std::vector<int64_t>& ws = n->shapes[wname];
  const int kh = (int)ws[0], kw = (int)ws[1], cin = (int)ws[2], cout = (int)ws[3];
  const std::vector<float>& W = n->params[wname];
  std::vector<float> w((size_t)cout * cin * kh * kw), scale(cout, 1.f), shift(cout, 0.f);
  if (!bn_after.empty()) bn_fold(n, bn_after, &scale, &shift);
  for (int y = 0; y < kh; y++)
    for (int x = 0; x < kw; x++)
      for (int i = 0; i < cin; i++)
        for (int o = 0; o < cout; o++)
          w[(((size_t)o * cin + i) * kh + y) * kw + x] = W[(((size_t)y * kw + x) * cin + i) * cout + o] * scale[o];
  ConvGeom g;
  g.stride = stride;
  same_pad(in.H, kh, stride, &g.pad_t, &g.pad_b);
  same_pad(in.W, kw, stride, &g.pad_l, &g.pad_r);
  g.slope = bn_after.empty() ? 1.f : 0.f;
  n->conv_weights.emplace_back(new ConvWeightsUmma());
  n->conv_plans.emplace_back(new ConvPlanUmma());
  ConvWeightsUmma* cw = n->conv_weights.back().get();
  ConvPlanUmma* pl = n->conv_plans.back().get();
  const int Ho = (in.H + stride - 1) / stride, Wo = (in.W + stride - 1) / stride;
  const bool flat = kh == 1 && kw == 1 && stride == 1 && !out.f32.p && Ho * Wo >= 256;   // CTA-pair tiles are 256 pixels of ONE image
  PV_TRY(pack_conv_weights_umma(cw, w.data(), shift.data(), cout, cin, kh, kw, cin_map, cin_phys, 0, (long)in.N * Ho * Wo, flat));
  PV_TRY(plan_conv_umma(pl, in, out, *cw, g, &n->conv_ws));
  n->steps.push_back([pl](cudaStream_t st, int na) { return launch_conv_umma(*pl, st, na); });
  return 0;
}

int build_network(premvos_reidnet* n) {
  PV_TRY(alloc_cview(n, &n->input, 8, REID_S, REID_S));
  const int map3[3] = {0, 1, 2};
  FView x;   // raw output of the previous layer (F8)
  PV_TRY(alloc_fview(n, &x, 64, REID_S, REID_S));
  {
    ConvOut o; o.f8 = x;
    PV_TRY(add_conv(n, "conv0/W", "", n->input, o, 1, map3, 8));
  }
  n->named["conv0"] = x;
  for (const UnitSpec& u : REID_UNITS) {
    const std::string s = u.name;
    const int cin = x.C, cl = u.feats[u.n - 1];
    int sres = 1;
    for (int i = 0; i < u.n; i++) sres *= u.strides[i];
    CView t;
    PV_TRY(alloc_cview(n, &t, cin, x.H, x.W));
    PV_TRY(add_bn_relu(n, s + "/bn0", x, t));
    const int Ho = (x.H + sres - 1) / sres, Wo = (x.W + sres - 1) / sres;
    FView res = x;
    if (cl != cin || sres != 1) {
      PV_TRY(alloc_fview(n, &res, cl, Ho, Wo));
      ConvOut o; o.f8 = res;
      PV_TRY(add_conv(n, s + "/W0", "", t, o, sres));
    }
    CView cur = t;
    FView y;
    for (int i = 0; i < u.n; i++) {
      const int st = u.strides[i];
      const int h2 = (cur.H + st - 1) / st, w2 = (cur.W + st - 1) / st;
      const bool last = i == u.n - 1;
      ConvOut o;
      CView nxt;
      if (last) {
        PV_TRY(alloc_fview(n, &y, u.feats[i], h2, w2));
        o.f8 = y; o.res_f8 = res;
      } else {
        PV_TRY(alloc_cview(n, &nxt, u.feats[i], h2, w2));
        o.cp = nxt;
      }
      PV_TRY(add_conv(n, s + "/W" + std::to_string(i + 1), last ? "" : s + "/bn" + std::to_string(i + 2), cur, o, st));
      cur = nxt;
    }
    x = y;
    n->named[s] = x;
  }
  // conv1: BN -> ReLU -> 3x3 conv (500 features, raw) -> max pool; then the fully-connected head
  CView t;
  PV_TRY(alloc_cview(n, &t, x.C, x.H, x.W));
  PV_TRY(add_bn_relu(n, "conv1/bn", x, t));
  PV_CHECK(x.H == 4 && x.W == 4, PREMVOS_ERR_UNSUPPORTED, "reidnet: the head expects a 4 x 4 map (128 x 128 crops)");
  n->conv1_out.N = n->NB; n->conv1_out.H = 4; n->conv1_out.W = 4; n->conv1_out.cs = 512; n->conv1_out.coff = 0; n->conv1_out.C = 500;
  PV_TRY(dev_alloc(n, &n->conv1_out.p, (size_t)n->NB * 16 * 512));
  {
    ConvOut o; o.f32 = n->conv1_out;
    PV_TRY(add_conv(n, "conv1/W", "", t, o, 1));
  }
  // fully-connected layers with their leading BatchNorm folded in: z = W^T (s * x + t) + b = (s (.) W)^T x + (W^T t + b)
  HeadArgs ha;
  ha.conv1 = n->conv1_out.p; ha.cs = 512;
  const char* names[3] = {"fc1", "fc2", "outputTriplet"};
  const float** wp[3] = {&ha.w1, &ha.w2, &ha.w3};
  const float** bp[3] = {&ha.b1, &ha.b2, &ha.b3};
  for (int l = 0; l < 3; l++) {
    const std::string s = names[l];
    const std::vector<int64_t>& ws = n->shapes[s + "/W"];
    const int fin = (int)ws[0], fout = (int)ws[1];
    std::vector<float> scale, shift;
    bn_fold(n, s + "/bn", &scale, &shift);
    const std::vector<float>&W = n->params[s + "/W"], &b = n->params[s + "/b"];
    std::vector<float> w2((size_t)fin * fout);
    std::vector<double> b2(b.begin(), b.end());
    for (int i = 0; i < fin; i++)
      for (int o = 0; o < fout; o++) {
        w2[(size_t)i * fout + o] = W[(size_t)i * fout + o] * scale[i];
        b2[o] += (double)W[(size_t)i * fout + o] * shift[i];
      }
    std::vector<float> b2f(b2.begin(), b2.end());
    float *dw = nullptr, *db = nullptr;
    PV_TRY(upload(n, w2, &dw)); PV_TRY(upload(n, b2f, &db));
    *wp[l] = dw; *bp[l] = db;
  }
  PV_TRY(dev_alloc(n, &n->emb_dev, (size_t)n->NB * 128));
  ha.out = n->emb_dev;
  n->steps.push_back([ha](cudaStream_t st, int na) {
    if (na == 0) return 0;
    prof_before(st);
    reid_head_kernel<<<na, 256, 0, st>>>(ha);
    return after_launch("reid_head_kernel", st, 2.0 * na * (2000.0 * 500 + 500.0 * 500 + 500.0 * 128), 4.0 * (2000.0 * 500 + 500.0 * 500 + 500.0 * 128));
  });
  PV_TRY(dev_alloc(n, &n->boxes_dev, (size_t)n->NB * 4));
  PV_TRY(dev_alloc(n, &n->crops, (size_t)n->NB * 4));
  return 0;
}

int run_batch(premvos_reidnet* n, int na, cudaStream_t st) {
  for (auto& s : n->steps) PV_TRY(s(st, na));
  return 0;
}

}  // namespace

extern "C" int premvos_reidnet_create(premvos_reidnet_t** out, int max_batch) {
  PV_CHECK(out, PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_create: out is null");
  *out = nullptr;
  PV_CHECK(max_batch >= 1 && max_batch <= 256, PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_create: max_batch in [1,256] (configs/run: batch_size 256)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PREMVOS_ERR_NO_DEVICE, "premvos_reidnet_create: no CUDA device visible");
  premvos_reidnet* n = new premvos_reidnet();
  n->NB = max_batch;
  build_shape_table(n);
  *out = n;
  return 0;
}

extern "C" int premvos_reidnet_set_param(premvos_reidnet_t* n, const char* name, const float* host_data, int64_t numel) {
  PV_CHECK(n && name && host_data, PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_set_param: null argument");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_reidnet_set_param: network already finalized");
  auto it = n->shapes.find(name);
  if (it == n->shapes.end()) return fail(PREMVOS_ERR_UNKNOWN_PARAM, "premvos_reidnet_set_param: unexpected variable '%s'", name);
  const int64_t want = numel_of(it->second);
  if (numel != want)
    return fail(PREMVOS_ERR_BAD_SHAPE, "premvos_reidnet_set_param: '%s' has %lld elements, expected %lld", name, (long long)numel, (long long)want);
  n->params[name].assign(host_data, host_data + numel);
  return 0;
}

extern "C" int premvos_reidnet_finalize(premvos_reidnet_t* n) {
  PV_CHECK(n, PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_finalize: null handle");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_reidnet_finalize: already finalized");
  for (auto& kv : n->shapes)
    if (!n->params.count(kv.first)) return fail(PREMVOS_ERR_NOT_READY, "premvos_reidnet_finalize: missing variable '%s'", kv.first.c_str());
  PV_CUDA(cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking));
  PV_TRY(build_network(n));
  n->params.clear();
  const int64_t before = g_launch_count.load();
  PV_TRY(run_batch(n, n->NB, n->stream));   // warm-up on the zero input: validates every launch configuration
  PV_CUDA(cudaStreamSynchronize(n->stream));
  n->launches_per_forward = (int)(g_launch_count.load() - before) + 1;
  n->finalized = true;
  return 0;
}

// One frame, n boxes, everything on the device and on `st`; no synchronisation.
static int enqueue_reid(premvos_reidnet* n, const unsigned char* frame_dev, int height, int width, const float* boxes_dev, int num_boxes,
                        float* emb_dev, cudaStream_t st) {
  for (int b0 = 0; b0 < num_boxes; b0 += n->NB) {
    const int na = std::min(n->NB, num_boxes - b0);
    CropArgs ca{frame_dev, height, width, boxes_dev + (size_t)b0 * 4, na, dev(n->input), n->crops};
    const long total = (long)na * REID_S * REID_S;
    prof_before(st);
    reid_input_kernel<<<blocks_for(total), 256, 0, st>>>(ca);
    PV_TRY(after_launch("reid_input_kernel", st, 40.0 * total, (double)total * 44.0));
    PV_TRY(run_batch(n, na, st));
    PV_CUDA(cudaMemcpyAsync(emb_dev + (size_t)b0 * 128, n->emb_dev, (size_t)na * 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

extern "C" int premvos_reidnet_forward(premvos_reidnet_t* n, const unsigned char* frame_rgb_dev, int height, int width,
                                       const float* boxes_xywh_dev, int num_boxes, float* embeddings_dev, void* stream) {
  PV_CHECK(n && frame_rgb_dev && (num_boxes == 0 || (boxes_xywh_dev && embeddings_dev)), PREMVOS_ERR_INVALID_ARG,
           "premvos_reidnet_forward: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_reidnet_forward: call premvos_reidnet_finalize first");
  PV_CHECK(height > 0 && width > 0 && num_boxes >= 0, PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_forward: bad sizes");
  return enqueue_reid(n, frame_rgb_dev, height, width, boxes_xywh_dev, num_boxes, embeddings_dev, (cudaStream_t)stream);
}

extern "C" int premvos_reidnet_forward_host(premvos_reidnet_t* n, const unsigned char* frame_rgb, int height, int width, const float* boxes_xywh,
                                            int num_boxes, float* embeddings_out) {
  PV_CHECK(n && frame_rgb && (num_boxes == 0 || (boxes_xywh && embeddings_out)), PREMVOS_ERR_INVALID_ARG,
           "premvos_reidnet_forward_host: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_reidnet_forward_host: call premvos_reidnet_finalize first");
  PV_CHECK(height > 0 && width > 0 && num_boxes >= 0, PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_forward_host: bad sizes");
  if (num_boxes == 0) return 0;
  cudaStream_t st = n->stream;
  const size_t hw = (size_t)height * width;
  if (hw * 3 > n->frame_cap) {
    if (n->frame_dev) cudaFree(n->frame_dev);
    n->frame_dev = nullptr; n->frame_cap = 0;
    PV_CUDA(cudaMalloc((void**)&n->frame_dev, hw * 3));
    n->frame_cap = hw * 3;
  }
  float *bdev = nullptr, *edev = nullptr;
  PV_CUDA(cudaMalloc((void**)&bdev, (size_t)num_boxes * 4 * sizeof(float)));
  if (cudaMalloc((void**)&edev, (size_t)num_boxes * 128 * sizeof(float)) != cudaSuccess) {
    cudaFree(bdev);
    return fail(PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_forward_host: out of device memory for %d embeddings", num_boxes);
  }
  int r = 0;
  cudaError_t e = cudaMemcpyAsync(n->frame_dev, frame_rgb, hw * 3, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(bdev, boxes_xywh, (size_t)num_boxes * 4 * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) r = enqueue_reid(n, n->frame_dev, height, width, bdev, num_boxes, edev, st);
  if (e == cudaSuccess && r == 0) e = cudaMemcpyAsync(embeddings_out, edev, (size_t)num_boxes * 128 * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && r == 0) e = cudaStreamSynchronize(st);
  cudaFree(bdev); cudaFree(edev);
  if (r != 0) return r;
  if (e != cudaSuccess) return fail((int)e, "premvos_reidnet_forward_host: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int premvos_reidnet_launches_per_forward(const premvos_reidnet_t* n) { return n ? n->launches_per_forward : 0; }

// Test hook (state of the LAST batch): "net_input" [NB,8,128,128] (CP8 -> NCHW), "conv0", "res0" .. "res16" (raw unit outputs, F8 ->
// NCHW [NB,C,H,W]), "conv1" (fp32 [NB,4,4,512] channels-last, before the max pool), "crops" ([NB,4] x y w h as fp32).
extern "C" int premvos_reidnet_get_tensor(premvos_reidnet_t* n, const char* name, float* host_out, int64_t* numel) {
  PV_CHECK(n && name && numel, PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_get_tensor: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_reidnet_get_tensor: network not finalized");
  PV_CUDA(cudaDeviceSynchronize());
  const std::string k(name);
  if (k == "conv1") {
    *numel = (int64_t)n->NB * 16 * 512;
    if (host_out) PV_CUDA(cudaMemcpy(host_out, n->conv1_out.p, (size_t)(*numel) * 4, cudaMemcpyDeviceToHost));
    return 0;
  }
  if (k == "crops") {
    *numel = (int64_t)n->NB * 4;
    if (host_out) {
      std::vector<int> t((size_t)n->NB * 4);
      PV_CUDA(cudaMemcpy(t.data(), n->crops, t.size() * 4, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < t.size(); i++) host_out[i] = (float)t[i];
    }
    return 0;
  }
  if (k == "net_input") {
    const CView& cv = n->input;
    *numel = (int64_t)cv.N * cv.C * cv.H * cv.W;
    if (!host_out) return 0;
    float* dtmp = nullptr;
    PV_CUDA(cudaMalloc((void**)&dtmp, (size_t)(*numel) * sizeof(float)));
    int r = cp8_to_nchw(cv, 0, dtmp, nullptr);
    if (r == 0 && cudaMemcpy(host_out, dtmp, (size_t)(*numel) * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
      r = fail(PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_get_tensor: copy failed");
    cudaFree(dtmp);
    return r;
  }
  auto it = n->named.find(k);
  if (it == n->named.end()) return fail(PREMVOS_ERR_INVALID_ARG, "premvos_reidnet_get_tensor: unknown tensor '%s'", name);
  const FView& f = it->second;
  *numel = (int64_t)f.N * f.C * f.H * f.W;
  if (!host_out) return 0;
  // F8 [N][chunks][H][W][8] -> NCHW on the host
  std::vector<float> raw((size_t)f.N * f.chunks * f.H * f.W * 8);
  PV_CUDA(cudaMemcpy(raw.data(), f.p, raw.size() * 4, cudaMemcpyDeviceToHost));
  const long hw = (long)f.H * f.W;
  for (int b = 0; b < f.N; b++)
    for (int c = 0; c < f.C; c++)
      for (long p = 0; p < hw; p++)
        host_out[((long)b * f.C + c) * hw + p] = raw[((((long)b * f.chunks + c / 8) * hw) + p) * 8 + (c & 7)];
  return 0;
}

extern "C" void premvos_reidnet_destroy(premvos_reidnet_t* n) {
  if (!n) return;
  cudaDeviceSynchronize();
  for (void* p : n->allocs) cudaFree(p);
  for (auto& w : n->conv_weights) free_conv_weights_umma(w.get());
  for (auto& pl : n->conv_plans) free_conv_plan_umma(pl.get());
  conv_workspace_free(&n->conv_ws);
  if (n->frame_dev) cudaFree(n->frame_dev);
  if (n->stream) cudaStreamDestroy(n->stream);
  delete n;
}
