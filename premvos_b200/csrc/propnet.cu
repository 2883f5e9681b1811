// Proposal network forward (ResNet-101 C4 Faster R-CNN, class-agnostic + second classification head):
// host-side orchestration and C ABI.
//
// Restates the inference branch of proposal_net/train.py:Model._build_graph (:107-189, :274-295) as a fixed
// sequence of kernels over pre-allocated CP8 buffers:
//   * frozen BatchNorm (basemodel.py:78, use_local_stat=False) is folded into the convolution weights/bias;
//   * every convolution runs on tcgen05 (conv_umma.cu); ReLU and the residual add of a bottleneck are fused
//     into the epilogue of its conv3 (basemodel.py:51-60, 63-71);
//   * the reference's explicit asymmetric pads + VALID (basemodel.py:54-56, 79-82) and the cropped stride-2
//     shortcut (basemodel.py:40-43) are expressed as per-side paddings of the TMA box origin;
//   * rpn/class and rpn/box share one 1x1 convolution (75 outputs) written as fp32 channels-last;
//   * top-k, NMS, RoIAlign, heads and the final selection never leave the device (the reference's
//     tf.image.non_max_suppression is a CPU kernel in TF 1.8 -> host round trip mid-graph);
//   * the number of RoIs is fixed at TEST_POST_NMS_TOPK = 100 with a device-side count, so shapes are static.
#include <math.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

using namespace premvos;

namespace {

const int NUM_ANCHOR = 15, ANCHOR_STRIDE = 16, MAX_SIZE = 1333;
const int PRE_NMS_TOPK = 1000, POST_NMS_TOPK = 100, RESULTS_PER_IM = 20;
// per-image strides of the detection buffers
const int S_TOPK = 1024, S_KEEP = 128, S_MASK = 1024 * 32;
const float RPN_NMS_THRESH = 0.7f, FRCNN_NMS_THRESH = 0.5f, SCORE_THRESH = 0.5f, BN_EPS = 1e-5f;

struct ConvLayer {
  ConvWeightsUmma w;
  ConvPlanUmma plan;
  bool used = false;
};

struct Bottleneck {
  ConvLayer c1, c2, c3, sc;
  bool has_sc = false;
  int stride = 1;
  CView t1, t2, scb, out;
};

}  // namespace

struct premvos_propnet {
  int H = 0, W = 0, num_class = 2, second_num_class = 81;
  int blocks[4] = {3, 4, 23, 3};
  int mode_mask = 0;         // option "mode_mask": also run the Mask R-CNN mask head on the final boxes (config.MODE_MASK)
  int batch = 1;             // frames per forward (option "batch"): one launch group for all of them
  int stem_rows = 1;         // conv0 as a 7x1 convolution over the row im2col of the input (see build_network)
  bool finalized = false;
  std::map<std::string, std::vector<float>> params;
  std::map<std::string, std::vector<int64_t>> shapes;
  std::vector<void*> allocs;
  cudaStream_t stream = nullptr;
  // the per-image detection tails (top-k, NMS, RoIAlign, final selection: short latency-bound kernels) of a batched forward run as
  // parallel branches: image b > 0 on side[b - 1], forked from / joined to the launching stream by events (also under capture)
  std::vector<cudaStream_t> side;
  std::vector<cudaEvent_t> ev_join;
  cudaEvent_t ev_fork = nullptr;

  float* img_dev = nullptr;  // [batch][H][W][3] fp32 BGR 0..255
  // every per-image buffer below holds `batch` consecutive copies (strides: the S_* constants)
  CView img, c0, pool;
  ConvLayer conv0;
  std::vector<std::unique_ptr<Bottleneck>> backbone, head;
  CView featuremap;
  int fh = 0, fw = 0, n_anchor_total = 0;
  ConvLayer rpn0, rpn_heads;
  CView rpn_hidden;
  TView rpn_out;  // fp32 [1,fh,fw,80]: 15 logits + 60 deltas
  float* cell_anchors = nullptr;
  float *d_scores = nullptr, *d_boxes = nullptr, *topk_score = nullptr, *valid_boxes = nullptr, *valid_scores = nullptr;
  int *topk_idx = nullptr, *topk_count = nullptr, *valid_src = nullptr, *valid_count = nullptr, *keep = nullptr, *keep_count = nullptr;
  uint32_t* nms_mask = nullptr;
  float *prop_boxes = nullptr, *prop_scores = nullptr;
  CView roi;
  int nfc = 0;
  float *fc_w = nullptr, *fc_b = nullptr, *fc_out = nullptr, *pooled = nullptr;
  float *all_probs = nullptr, *all_boxes = nullptr, *second_probs = nullptr;
  // results (device) and their pinned host mirror
  int* n_out = nullptr;
  float *final_boxes = nullptr, *final_probs = nullptr, *final_posterior = nullptr, *second_final_posterior = nullptr;
  int64_t *final_labels = nullptr, *second_final_labels = nullptr;
  int* final_box_index = nullptr;
  // mask head (mode_mask): RoIAlign on the final boxes -> conv5 again -> deconv (as a 1x1 GEMM to 4 x 256 phases) -> 1x1 + sigmoid
  CView roi_m, deconv_out;
  std::vector<std::unique_ptr<Bottleneck>> head_m;
  ConvLayer deconv;
  float *mask_w = nullptr, *mask_b = nullptr, *final_masks = nullptr;   // final_masks [batch][RESULTS_PER_IM][14][14]
  int launches_per_forward = 0;
  ConvWorkspace conv_ws;             // stream-K scratch of the CTA-pair convolution launches
  cudaGraph_t graph = nullptr;       // the whole forward (fixed shapes per handle, device-side counts)
  cudaGraphExec_t exec = nullptr;
  int graph_nodes = 0, opt_cuda_graph = 1;
};

namespace {

int64_t numel_of(const std::vector<int64_t>& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}

void build_shape_table(premvos_propnet* n) {
  auto conv_bn = [&](const std::string& s, int k, int cin, int cout) {
    n->shapes[s + "/W"] = {k, k, cin, cout};
    for (const char* v : {"gamma", "beta", "mean/EMA", "variance/EMA"}) n->shapes[s + "/bn/" + v] = {cout};
  };
  conv_bn("conv0", 7, 3, 64);
  int cin = 64;
  const int chs[4] = {64, 128, 256, 512};
  for (int g = 0; g < 4; g++)
    for (int b = 0; b < n->blocks[g]; b++) {
      std::string s = "group" + std::to_string(g) + "/block" + std::to_string(b);
      conv_bn(s + "/conv1", 1, cin, chs[g]);
      conv_bn(s + "/conv2", 3, chs[g], chs[g]);
      conv_bn(s + "/conv3", 1, chs[g], chs[g] * 4);
      if (cin != chs[g] * 4) conv_bn(s + "/convshortcut", 1, cin, chs[g] * 4);
      cin = chs[g] * 4;
    }
  n->shapes["rpn/conv0/W"] = {3, 3, 1024, 1024};
  n->shapes["rpn/conv0/b"] = {1024};
  n->shapes["rpn/class/W"] = {1, 1, 1024, NUM_ANCHOR};
  n->shapes["rpn/class/b"] = {NUM_ANCHOR};
  n->shapes["rpn/box/W"] = {1, 1, 1024, 4 * NUM_ANCHOR};
  n->shapes["rpn/box/b"] = {4 * NUM_ANCHOR};
  n->shapes["fastrcnn/class/W"] = {2048, n->num_class};
  n->shapes["fastrcnn/class/b"] = {n->num_class};
  n->shapes["fastrcnn/box/W"] = {2048, (n->num_class - 1) * 4};
  n->shapes["fastrcnn/box/b"] = {(n->num_class - 1) * 4};
  if (n->second_num_class > 0) {
    n->shapes["secondclassification/class/W"] = {2048, n->second_num_class};
    n->shapes["secondclassification/class/b"] = {n->second_num_class};
  }
  if (n->mode_mask) {   // model.py:495-509: tensorpack Deconv2D W [kh, kw, out, in], Conv2D W HWIO
    n->shapes["maskrcnn/deconv/W"] = {2, 2, 256, 2048};
    n->shapes["maskrcnn/deconv/b"] = {256};
    n->shapes["maskrcnn/conv/W"] = {1, 1, 256, n->num_class - 1};
    n->shapes["maskrcnn/conv/b"] = {n->num_class - 1};
  }
}

template <typename T>
int dev_alloc(premvos_propnet* n, T** p, size_t count) {
  PV_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  PV_CUDA(cudaMemset(*p, 0, count * sizeof(T)));
  n->allocs.push_back(*p);
  return 0;
}

int alloc_cview(premvos_propnet* n, CView* v, int N, int C, int H, int W) {
  v->N = N; v->H = H; v->W = W; v->chunks = (C + 7) / 8; v->c0 = 0; v->C = C;
  const size_t elems = (size_t)N * v->chunks * H * W * 8 + 64;  // + 128 B slack for flattened 1x1 layers (conv_umma.cu)
  PV_TRY(dev_alloc(n, &v->hi, elems));
  PV_TRY(dev_alloc(n, &v->lo, elems));
  return 0;
}

// Conv (+ folded BatchNorm or bias) from tensorpack variables: W is HWIO.
int make_conv(premvos_propnet* n, ConvLayer* L, const std::string& scope, bool bn, const CView& in, const ConvOut& out,
              const ConvGeom& g, const int* cin_map = nullptr, int cin_phys = 0) {
  const std::vector<int64_t>& ws = n->shapes[scope + "/W"];
  const int kh = (int)ws[0], kw = (int)ws[1], cin = (int)ws[2], cout = (int)ws[3];
  const std::vector<float>& W = n->params[scope + "/W"];
  std::vector<float> w((size_t)cout * cin * kh * kw), b(cout, 0.f);
  std::vector<float> scale(cout, 1.f);
  if (bn) {
    const std::vector<float>&gm = n->params[scope + "/bn/gamma"], &bt = n->params[scope + "/bn/beta"];
    const std::vector<float>&mu = n->params[scope + "/bn/mean/EMA"], &var = n->params[scope + "/bn/variance/EMA"];
    for (int o = 0; o < cout; o++) {
      scale[o] = gm[o] / sqrtf(var[o] + BN_EPS);
      b[o] = bt[o] - mu[o] * scale[o];
    }
  } else if (n->params.count(scope + "/b")) {
    b = n->params[scope + "/b"];
  }
  for (int y = 0; y < kh; y++)
    for (int x = 0; x < kw; x++)
      for (int i = 0; i < cin; i++)
        for (int o = 0; o < cout; o++)
          w[(((size_t)o * cin + i) * kh + y) * kw + x] = W[(((size_t)y * kw + x) * cin + i) * cout + o] * scale[o];
  const long m_out = out.cp.hi ? (long)out.cp.N * out.cp.H * out.cp.W : (long)out.f32.N * out.f32.H * out.f32.W;
  // The CTA-pair kernel (conv_umma.cu) is NOT used here (PREMVOS_PROPNET_PAIR=1 switches it on for experiments): measured inside
  // the batch-4 forward, the bottleneck 1x1 layers (K = 256 .. 1024, residual add in the epilogue) are bound by their epilogue and
  // memory traffic, not by the operand fill the pair halves (256 -> 1024: 62 us single-CTA, 71 us pair; profiles/r02_*propnet*), and
  // its stream-K split makes the fp32 summation order of an image depend on its position in the batch.
  static const bool pair_ok = getenv("PREMVOS_PROPNET_PAIR") && atoi(getenv("PREMVOS_PROPNET_PAIR")) != 0;
  const bool flat = pair_ok && kh == 1 && kw == 1 && g.stride == 1 && g.pad_t == 0 && g.pad_l == 0 && g.pad_b == 0 && g.pad_r == 0 &&
                    out.cp.hi && !out.f32.p;
  PV_TRY(pack_conv_weights_umma(&L->w, w.data(), b.data(), cout, cin, kh, kw, cin_map, cin_phys, 0, m_out, flat));
  PV_TRY(plan_conv_umma(&L->plan, in, out, L->w, g, &n->conv_ws));
  L->used = true;
  return 0;
}

// basemodel.py:51-60 + 36-48 + 63-71 (ReLU after the residual add)
int make_bottleneck(premvos_propnet* n, Bottleneck* B, const std::string& scope, const CView& in, int ch, int stride) {
  B->stride = stride;
  B->has_sc = in.C != ch * 4;
  const int Ho = stride == 2 ? (in.H + 1 - 3) / 2 + 1 : in.H, Wo = stride == 2 ? (in.W + 1 - 3) / 2 + 1 : in.W;
  PV_TRY(alloc_cview(n, &B->t1, in.N, ch, in.H, in.W));
  PV_TRY(alloc_cview(n, &B->t2, in.N, ch, Ho, Wo));
  PV_TRY(alloc_cview(n, &B->out, in.N, ch * 4, Ho, Wo));
  ConvGeom g1;  // 1x1, BN + ReLU
  g1.slope = 0.f;
  ConvOut o1; o1.cp = B->t1;
  PV_TRY(make_conv(n, &B->c1, scope + "/conv1", true, in, o1, g1));
  ConvGeom g2;
  g2.slope = 0.f;
  if (stride == 2) { g2.stride = 2; g2.pad_b = g2.pad_r = 1; }   // tf.pad [0,1] + VALID
  else { g2.pad_t = g2.pad_l = g2.pad_b = g2.pad_r = 1; }         // SAME
  ConvOut o2; o2.cp = B->t2;
  PV_TRY(make_conv(n, &B->c2, scope + "/conv2", true, B->t1, o2, g2));
  CView res = in;
  if (B->has_sc) {
    PV_TRY(alloc_cview(n, &B->scb, in.N, ch * 4, Ho, Wo));
    ConvGeom gs;  // 1x1 stride s on l[:, :, :-1, :-1] (stride 2), BN, no activation
    gs.stride = stride;
    if (stride == 2) gs.pad_b = gs.pad_r = -1;
    ConvOut os; os.cp = B->scb;
    PV_TRY(make_conv(n, &B->sc, scope + "/convshortcut", true, in, os, gs));
    res = B->scb;
  }
  ConvGeom g3;  // 1x1, BN, + shortcut, ReLU
  g3.slope = 0.f;
  ConvOut o3; o3.cp = B->out; o3.res = res;
  PV_TRY(make_conv(n, &B->c3, scope + "/conv3", true, B->t2, o3, g3));
  return 0;
}

int run_bottleneck(Bottleneck* B, cudaStream_t st) {
  PV_TRY(launch_conv_umma(B->c1.plan, st));
  PV_TRY(launch_conv_umma(B->c2.plan, st));
  if (B->has_sc) PV_TRY(launch_conv_umma(B->sc.plan, st));
  PV_TRY(launch_conv_umma(B->c3.plan, st));
  return 0;
}

// utils/generate_anchors.py:40-99 with base 16, ratios (.5,1,2), scales sizes/16 (data.py:49-52)
void make_cell_anchors(float* out /*[15][4]*/) {
  const double ratios[3] = {0.5, 1.0, 2.0}, scales[5] = {2, 4, 8, 16, 32};
  const double w0 = 16, h0 = 16, xc0 = 7.5, yc0 = 7.5;
  int k = 0;
  for (int r = 0; r < 3; r++) {
    const double size = w0 * h0, ws = nearbyint(sqrt(size / ratios[r])), hs = nearbyint(ws * ratios[r]);
    // ratio anchor = centred (ws, hs) box; its (w, h, ctr) as _whctrs recomputes them
    const double x1 = xc0 - 0.5 * (ws - 1), y1 = yc0 - 0.5 * (hs - 1), x2 = xc0 + 0.5 * (ws - 1), y2 = yc0 + 0.5 * (hs - 1);
    const double w = x2 - x1 + 1, h = y2 - y1 + 1, xc = x1 + 0.5 * (w - 1), yc = y1 + 0.5 * (h - 1);
    for (int s = 0; s < 5; s++, k++) {
      const double wss = w * scales[s], hss = h * scales[s];
      out[k * 4 + 0] = (float)(xc - 0.5 * (wss - 1));
      out[k * 4 + 1] = (float)(yc - 0.5 * (hss - 1));
      out[k * 4 + 2] = (float)(xc + 0.5 * (wss - 1));
      out[k * 4 + 3] = (float)(yc + 0.5 * (hss - 1));
    }
  }
}

int build_network(premvos_propnet* n) {
  const int H = n->H, W = n->W, NB = n->batch;
  PV_TRY(dev_alloc(n, &n->img_dev, (size_t)NB * H * W * 3));
  // conv0: pad (2,3) + 7x7 stride 2 VALID + BN + ReLU (basemodel.py:79-80).  With 3 input channels a per-tap K step would be
  // 13/16 zeros and the layer 49 operand boxes per tile; the preprocessing kernel therefore writes the ROW IM2COL of the
  // normalised image -- [H][W0] pixels of 7 horizontal taps x 3 channels = 21 (+3 zero) channels, the horizontal stride and
  // padding already applied -- and conv0 runs as a 7x1 convolution over it (vertical stride 2, horizontal stride 1, K = 32 per
  // tap: 14 boxes per tile).  Same products, same zero padding; the fp32 sum of a pixel is taken in a different order.
  const int H0 = (H + 5 - 7) / 2 + 1, W0 = (W + 5 - 7) / 2 + 1;
  static const bool rows7 = !(getenv("PREMVOS_STEM_ROWS") && atoi(getenv("PREMVOS_STEM_ROWS")) == 0);
  n->stem_rows = rows7 ? 1 : 0;
  if (rows7) PV_TRY(alloc_cview(n, &n->img, NB, 24, H, W0));
  else { PV_TRY(alloc_cview(n, &n->img, NB, 8, H, W)); n->img.C = 8; }
  PV_TRY(alloc_cview(n, &n->c0, NB, 64, H0, W0));
  if (rows7) {
    const std::vector<float>& W7 = n->params["conv0/W"];   // HWIO [7][7][3][64]
    const std::vector<float>&gm = n->params["conv0/bn/gamma"], &bt = n->params["conv0/bn/beta"];
    const std::vector<float>&mu = n->params["conv0/bn/mean/EMA"], &var = n->params["conv0/bn/variance/EMA"];
    std::vector<float> w((size_t)64 * 21 * 7), b(64);
    for (int o = 0; o < 64; o++) {
      const float scale = gm[o] / sqrtf(var[o] + BN_EPS);
      b[o] = bt[o] - mu[o] * scale;
      for (int r = 0; r < 7; r++)
        for (int s = 0; s < 7; s++)
          for (int c = 0; c < 3; c++) w[((size_t)o * 21 + s * 3 + c) * 7 + r] = W7[(((size_t)r * 7 + s) * 3 + c) * 64 + o] * scale;   // [Cout][21][7][1]
    }
    ConvGeom g; g.stride = 2; g.stride_x = 1; g.pad_t = 2; g.pad_b = 3; g.slope = 0.f;
    ConvOut o; o.cp = n->c0;
    PV_TRY(pack_conv_weights_umma(&n->conv0.w, w.data(), b.data(), 64, 21, 7, 1, nullptr, 0, 0, (long)NB * H0 * W0, false));
    PV_TRY(plan_conv_umma(&n->conv0.plan, n->img, o, n->conv0.w, g, &n->conv_ws));
    n->conv0.used = true;
  } else {
    ConvGeom g; g.stride = 2; g.pad_t = g.pad_l = 2; g.pad_b = g.pad_r = 3; g.slope = 0.f;
    ConvOut o; o.cp = n->c0;
    const int map3[3] = {0, 1, 2};
    PV_TRY(make_conv(n, &n->conv0, "conv0", true, n->img, o, g, map3, 8));
  }
  const int H1 = (H0 + 1 - 3) / 2 + 1, W1 = (W0 + 1 - 3) / 2 + 1;
  PV_TRY(alloc_cview(n, &n->pool, NB, 64, H1, W1));
  CView cur = n->pool;
  const int chs[4] = {64, 128, 256, 512};
  for (int g = 0; g < 3; g++)
    for (int b = 0; b < n->blocks[g]; b++) {
      n->backbone.emplace_back(new Bottleneck());
      PV_TRY(make_bottleneck(n, n->backbone.back().get(), "group" + std::to_string(g) + "/block" + std::to_string(b), cur, chs[g],
                             (b == 0 && g > 0) ? 2 : 1));
      cur = n->backbone.back()->out;
    }
  n->featuremap = cur;
  n->fh = cur.H; n->fw = cur.W;
  PV_CHECK(n->fh == H / ANCHOR_STRIDE && n->fw == W / ANCHOR_STRIDE, PREMVOS_ERR_INVALID_ARG,
           "featuremap %dx%d != image//16 %dx%d (train.py:100-104 slices the anchor field by floor division)", n->fh, n->fw,
           H / ANCHOR_STRIDE, W / ANCHOR_STRIDE);
  n->n_anchor_total = n->fh * n->fw * NUM_ANCHOR;
  // RPN head (model.py:31-51)
  PV_TRY(alloc_cview(n, &n->rpn_hidden, NB, 1024, n->fh, n->fw));
  {
    ConvGeom g = ConvGeom::same3x3(1, 0.f);
    ConvOut o; o.cp = n->rpn_hidden;
    PV_TRY(make_conv(n, &n->rpn0, "rpn/conv0", false, cur, o, g));
  }
  {
    // class (15) and box (60) 1x1 convolutions fused into one 75-output layer
    std::vector<float>&cw = n->params["rpn/class/W"], &bw = n->params["rpn/box/W"];
    std::vector<float> w((size_t)75 * 1024), b(75);
    for (int i = 0; i < 1024; i++) {
      for (int o = 0; o < 15; o++) w[(size_t)o * 1024 + i] = cw[(size_t)i * 15 + o];
      for (int o = 0; o < 60; o++) w[(size_t)(15 + o) * 1024 + i] = bw[(size_t)i * 60 + o];
    }
    for (int o = 0; o < 15; o++) b[o] = n->params["rpn/class/b"][o];
    for (int o = 0; o < 60; o++) b[15 + o] = n->params["rpn/box/b"][o];
    n->rpn_out.N = NB; n->rpn_out.H = n->fh; n->rpn_out.W = n->fw; n->rpn_out.cs = 80; n->rpn_out.coff = 0; n->rpn_out.C = 75;
    PV_TRY(dev_alloc(n, &n->rpn_out.p, (size_t)NB * n->fh * n->fw * 80));
    PV_TRY(pack_conv_weights_umma(&n->rpn_heads.w, w.data(), b.data(), 75, 1024, 1, 1));
    ConvOut o; o.f32 = n->rpn_out;
    ConvGeom g;
    PV_TRY(plan_conv_umma(&n->rpn_heads.plan, n->rpn_hidden, o, n->rpn_heads.w, g));
    n->rpn_heads.used = true;
  }
  float ca[60];
  make_cell_anchors(ca);
  PV_TRY(dev_alloc(n, &n->cell_anchors, 60));
  PV_CUDA(cudaMemcpy(n->cell_anchors, ca, sizeof(ca), cudaMemcpyHostToDevice));
  const size_t nb = (size_t)NB;
  PV_TRY(dev_alloc(n, &n->d_scores, nb * n->n_anchor_total));
  PV_TRY(dev_alloc(n, &n->d_boxes, nb * n->n_anchor_total * 4));
  PV_TRY(dev_alloc(n, &n->topk_idx, nb * S_TOPK)); PV_TRY(dev_alloc(n, &n->topk_score, nb * S_TOPK)); PV_TRY(dev_alloc(n, &n->topk_count, nb));
  PV_TRY(dev_alloc(n, &n->valid_boxes, nb * S_TOPK * 4)); PV_TRY(dev_alloc(n, &n->valid_scores, nb * S_TOPK));
  PV_TRY(dev_alloc(n, &n->valid_src, nb * S_TOPK)); PV_TRY(dev_alloc(n, &n->valid_count, nb));
  PV_TRY(dev_alloc(n, &n->nms_mask, nb * S_MASK)); PV_TRY(dev_alloc(n, &n->keep, nb * S_KEEP)); PV_TRY(dev_alloc(n, &n->keep_count, nb));
  PV_TRY(dev_alloc(n, &n->prop_boxes, nb * S_KEEP * 4)); PV_TRY(dev_alloc(n, &n->prop_scores, nb * S_KEEP));
  // RoIAlign + conv5 head on a fixed batch of POST_NMS_TOPK RoIs per image
  PV_TRY(alloc_cview(n, &n->roi, NB * POST_NMS_TOPK, 1024, 14, 14));
  cur = n->roi;
  for (int b = 0; b < n->blocks[3]; b++) {
    n->head.emplace_back(new Bottleneck());
    PV_TRY(make_bottleneck(n, n->head.back().get(), "group3/block" + std::to_string(b), cur, 512, b == 0 ? 2 : 1));
    cur = n->head.back()->out;
  }
  // fully connected heads: [class 2 | box 4 | second 81]
  const int ncls = n->num_class, nbox = (n->num_class - 1) * 4, nsec = n->second_num_class;
  n->nfc = ncls + nbox + nsec;
  std::vector<float> fw((size_t)2048 * n->nfc), fb(n->nfc);
  for (int c = 0; c < 2048; c++) {
    for (int o = 0; o < ncls; o++) fw[(size_t)c * n->nfc + o] = n->params["fastrcnn/class/W"][(size_t)c * ncls + o];
    for (int o = 0; o < nbox; o++) fw[(size_t)c * n->nfc + ncls + o] = n->params["fastrcnn/box/W"][(size_t)c * nbox + o];
    for (int o = 0; o < nsec; o++) fw[(size_t)c * n->nfc + ncls + nbox + o] = n->params["secondclassification/class/W"][(size_t)c * nsec + o];
  }
  for (int o = 0; o < ncls; o++) fb[o] = n->params["fastrcnn/class/b"][o];
  for (int o = 0; o < nbox; o++) fb[ncls + o] = n->params["fastrcnn/box/b"][o];
  for (int o = 0; o < nsec; o++) fb[ncls + nbox + o] = n->params["secondclassification/class/b"][o];
  PV_TRY(dev_alloc(n, &n->fc_w, fw.size())); PV_TRY(dev_alloc(n, &n->fc_b, fb.size()));
  PV_CUDA(cudaMemcpy(n->fc_w, fw.data(), fw.size() * 4, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(n->fc_b, fb.data(), fb.size() * 4, cudaMemcpyHostToDevice));
  PV_TRY(dev_alloc(n, &n->fc_out, nb * POST_NMS_TOPK * n->nfc));
  PV_TRY(dev_alloc(n, &n->pooled, nb * POST_NMS_TOPK * 2048));
  PV_TRY(dev_alloc(n, &n->all_probs, nb * S_KEEP * 2)); PV_TRY(dev_alloc(n, &n->all_boxes, nb * S_KEEP * 4));
  PV_TRY(dev_alloc(n, &n->second_probs, nb * S_KEEP * (nsec > 0 ? nsec : 1)));
  PV_TRY(dev_alloc(n, &n->n_out, nb));
  PV_TRY(dev_alloc(n, &n->final_boxes, nb * RESULTS_PER_IM * 4)); PV_TRY(dev_alloc(n, &n->final_probs, nb * RESULTS_PER_IM));
  PV_TRY(dev_alloc(n, &n->final_labels, nb * RESULTS_PER_IM)); PV_TRY(dev_alloc(n, &n->final_posterior, nb * RESULTS_PER_IM * 2));
  PV_TRY(dev_alloc(n, &n->second_final_labels, nb * RESULTS_PER_IM));
  PV_TRY(dev_alloc(n, &n->second_final_posterior, nb * RESULTS_PER_IM * (nsec > 0 ? nsec : 1)));
  PV_TRY(dev_alloc(n, &n->final_box_index, nb * RESULTS_PER_IM));
  if (n->mode_mask) {
    // train.py:297-309: a second RoIAlign on the (<= 20) final boxes of every image and the conv5 group again (same weights)
    PV_TRY(alloc_cview(n, &n->roi_m, NB * RESULTS_PER_IM, 1024, 14, 14));
    cur = n->roi_m;
    for (int b = 0; b < n->blocks[3]; b++) {
      n->head_m.emplace_back(new Bottleneck());
      PV_TRY(make_bottleneck(n, n->head_m.back().get(), "group3/block" + std::to_string(b), cur, 512, b == 0 ? 2 : 1));
      cur = n->head_m.back()->out;
    }
    // Deconv2D(256, 2, stride 2) + ReLU as ONE 1x1 GEMM with 4 x 256 outputs: channel (dy*2+dx)*256 + co
    const std::vector<float>&dw = n->params["maskrcnn/deconv/W"], &db = n->params["maskrcnn/deconv/b"];
    std::vector<float> w((size_t)1024 * 2048), b(1024);
    for (int ph = 0; ph < 4; ph++)
      for (int co = 0; co < 256; co++) {
        b[ph * 256 + co] = db[co];
        for (int ci = 0; ci < 2048; ci++) w[(size_t)(ph * 256 + co) * 2048 + ci] = dw[((size_t)ph * 256 + co) * 2048 + ci];
      }
    PV_TRY(alloc_cview(n, &n->deconv_out, NB * RESULTS_PER_IM, 1024, cur.H, cur.W));
    PV_TRY(pack_conv_weights_umma(&n->deconv.w, w.data(), b.data(), 1024, 2048, 1, 1, nullptr, 0, 0,
                                  (long)n->deconv_out.N * cur.H * cur.W));
    ConvGeom g; g.slope = 0.f;
    ConvOut o; o.cp = n->deconv_out;
    PV_TRY(plan_conv_umma(&n->deconv.plan, cur, o, n->deconv.w, g));
    n->deconv.used = true;
    PV_CHECK(n->num_class == 2, PREMVOS_ERR_UNSUPPORTED, "mask head: class-agnostic only");
    PV_TRY(dev_alloc(n, &n->mask_w, 256)); PV_TRY(dev_alloc(n, &n->mask_b, 1));
    PV_CUDA(cudaMemcpy(n->mask_w, n->params["maskrcnn/conv/W"].data(), 256 * 4, cudaMemcpyHostToDevice));
    PV_CUDA(cudaMemcpy(n->mask_b, n->params["maskrcnn/conv/b"].data(), 4, cudaMemcpyHostToDevice));
    PV_TRY(dev_alloc(n, &n->final_masks, nb * RESULTS_PER_IM * 4 * cur.H * cur.W));
  }
  return 0;
}

// branch of image b: the launching stream for image 0, a side stream that waits for everything enqueued so far otherwise
static int fork_branch(premvos_propnet* n, cudaStream_t st, int b, cudaStream_t* out) {
  *out = st;
  if (n->side.empty() || profiling_enabled()) return 0;   // the per-launch profile is serial by definition: one stream, clean brackets
  if (b == 0) { PV_CUDA(cudaEventRecord(n->ev_fork, st)); return 0; }   // the fork point: before image 0's own kernels
  PV_CUDA(cudaStreamWaitEvent(n->side[b - 1], n->ev_fork, 0));
  *out = n->side[b - 1];
  return 0;
}
static int join_branches(premvos_propnet* n, cudaStream_t st) {
  if (n->side.empty() || profiling_enabled()) return 0;
  for (int b = 1; b < n->batch; b++) {
    PV_CUDA(cudaEventRecord(n->ev_join[b - 1], n->side[b - 1]));
    PV_CUDA(cudaStreamWaitEvent(st, n->ev_join[b - 1], 0));
  }
  return 0;
}

int run_network(premvos_propnet* n, cudaStream_t st0) {
  const int NB = n->batch, nsec = n->second_num_class > 0 ? n->second_num_class : 1;
  cudaStream_t st = st0;
  for (int b = 0; b < NB; b++) {
    if (n->stem_rows) PV_TRY(det_preprocess_rows7(n->img_dev + (size_t)b * n->H * n->W * 3, n->W, 2, n->img.batch_range(b, 1), st));
    else PV_TRY(det_preprocess(n->img_dev + (size_t)b * n->H * n->W * 3, n->img.batch_range(b, 1), st));
  }
  // backbone + RPN head: every frame of the batch in the same launches
  PV_TRY(launch_conv_umma(n->conv0.plan, st));
  PV_TRY(det_maxpool3x3s2(n->c0, n->pool, st));
  for (auto& b : n->backbone) PV_TRY(run_bottleneck(b.get(), st));
  PV_TRY(launch_conv_umma(n->rpn0.plan, st));
  PV_TRY(launch_conv_umma(n->rpn_heads.plan, st));
  for (int b = 0; b < NB; b++) {
    PV_TRY(fork_branch(n, st0, b, &st));
    const size_t na = (size_t)n->n_anchor_total;
    float *scores = n->d_scores + b * na, *boxes = n->d_boxes + b * na * 4;
    PV_TRY(det_rpn_decode(n->rpn_out.p + (size_t)b * n->fh * n->fw * n->rpn_out.cs, n->rpn_out.cs, n->fh, n->fw, NUM_ANCHOR, n->cell_anchors,
                          (float)ANCHOR_STRIDE, logf((float)MAX_SIZE / 16.0f), scores, boxes, st));
    // generate_rpn_proposals (model.py:170-217)
    PV_TRY(det_topk(scores, n->n_anchor_total, PRE_NMS_TOPK, n->topk_idx + b * S_TOPK, n->topk_score + b * S_TOPK, n->topk_count + b, st));
    PV_TRY(det_gather_clip_valid(boxes, n->topk_idx + b * S_TOPK, n->topk_score + b * S_TOPK, n->topk_count + b, (float)n->H, (float)n->W, 0.f,
                                 n->valid_boxes + b * S_TOPK * 4, n->valid_scores + b * S_TOPK, n->valid_src + b * S_TOPK, n->valid_count + b, st));
    PV_TRY(det_nms(n->valid_boxes + b * S_TOPK * 4, n->valid_count + b, RPN_NMS_THRESH, POST_NMS_TOPK, n->nms_mask + (size_t)b * S_MASK,
                   n->keep + b * S_KEEP, n->keep_count + b, st));
    PV_TRY(det_gather_proposals(n->valid_boxes + b * S_TOPK * 4, n->valid_scores + b * S_TOPK, n->keep + b * S_KEEP, n->keep_count + b,
                                POST_NMS_TOPK, n->prop_boxes + b * S_KEEP * 4, n->prop_scores + b * S_KEEP, st));
    // RoIAlign on the /16 grid (train.py:159)
    PV_TRY(det_roi_align(n->featuremap.batch_range(b, 1), n->prop_boxes + b * S_KEEP * 4, 1.0f / ANCHOR_STRIDE, 14,
                         n->roi.batch_range(b * POST_NMS_TOPK, POST_NMS_TOPK), st));
  }
  st = st0;
  PV_TRY(join_branches(n, st));
  // conv5 head on the RoIs of all frames, pooled features -> heads
  for (auto& b : n->head) PV_TRY(run_bottleneck(b.get(), st));
  PV_TRY(det_gap_fc(n->head.back()->out, n->fc_w, n->fc_b, n->nfc, n->pooled, n->fc_out, st));
  for (int b = 0; b < NB; b++) {
    PV_TRY(fork_branch(n, st0, b, &st));
    DetTailArgs t;
    t.logits = n->fc_out + (size_t)b * POST_NMS_TOPK * n->nfc; t.nfc = n->nfc; t.nsecond = n->second_num_class;
    t.prop_boxes = n->prop_boxes + b * S_KEEP * 4; t.prop_count = n->keep_count + b;
    t.img_h = (float)n->H; t.img_w = (float)n->W; t.clip = logf((float)MAX_SIZE / 16.0f); t.score_thresh = SCORE_THRESH;
    t.nms_thresh = FRCNN_NMS_THRESH; t.max_rois = POST_NMS_TOPK; t.results_per_im = RESULTS_PER_IM;
    t.all_probs = n->all_probs + b * S_KEEP * 2; t.all_boxes = n->all_boxes + b * S_KEEP * 4; t.second_probs = n->second_probs + (size_t)b * S_KEEP * nsec;
    t.n_out = n->n_out + b; t.final_boxes = n->final_boxes + b * RESULTS_PER_IM * 4; t.final_probs = n->final_probs + b * RESULTS_PER_IM;
    t.final_labels = n->final_labels + b * RESULTS_PER_IM; t.final_posterior = n->final_posterior + b * RESULTS_PER_IM * 2;
    t.second_final_labels = n->second_final_labels + b * RESULTS_PER_IM;
    t.second_final_posterior = n->second_final_posterior + (size_t)b * RESULTS_PER_IM * nsec; t.final_box_index = n->final_box_index + b * RESULTS_PER_IM;
    PV_TRY(det_frcnn_tail(t, st));
  }
  st = st0;
  PV_TRY(join_branches(n, st));
  if (n->mode_mask) {
    // rows >= n_out of final_boxes hold older / zero boxes: their masks are computed and never read
    for (int b = 0; b < NB; b++)
      PV_TRY(det_roi_align(n->featuremap.batch_range(b, 1), n->final_boxes + b * RESULTS_PER_IM * 4, 1.0f / ANCHOR_STRIDE, 14,
                           n->roi_m.batch_range(b * RESULTS_PER_IM, RESULTS_PER_IM), st));
    for (auto& b : n->head_m) PV_TRY(run_bottleneck(b.get(), st));
    PV_TRY(launch_conv_umma(n->deconv.plan, st));
    PV_TRY(det_mask_head(n->deconv_out, n->mask_w, n->mask_b, n->final_masks, st));
  }
  return 0;
}

// img_dev already holds the frame
int enqueue_network(premvos_propnet* n, cudaStream_t st) {
  if (n->exec && !profiling_enabled()) {
    PV_CUDA(cudaGraphLaunch(n->exec, st));
    count_launch(n->graph_nodes);
    return 0;
  }
  return run_network(n, st);
}

void free_layer(ConvLayer* L) {
  if (!L->used) return;
  free_conv_weights_umma(&L->w);
  free_conv_plan_umma(&L->plan);
}

}  // namespace

extern "C" int premvos_propnet_create(premvos_propnet_t** out, int height, int width, int num_class, int second_num_class) {
  PV_CHECK(out, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_create: out is null");
  *out = nullptr;
  PV_CHECK(height >= 64 && width >= 64 && height <= MAX_SIZE && width <= MAX_SIZE, PREMVOS_ERR_INVALID_ARG,
           "premvos_propnet_create: image %dx%d outside [64, %d]", height, width, MAX_SIZE);
  PV_CHECK(num_class == 2, PREMVOS_ERR_UNSUPPORTED,
           "premvos_propnet_create: only the class-agnostic head (NUM_CLASS=2, simple_run.sh --agnostic) is implemented");
  PV_CHECK(second_num_class >= 0 && second_num_class <= 128, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_create: second_num_class");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PREMVOS_ERR_NO_DEVICE, "premvos_propnet_create: no CUDA device visible");
  premvos_propnet* n = new premvos_propnet();
  n->H = height; n->W = width; n->num_class = num_class; n->second_num_class = second_num_class;
  build_shape_table(n);
  *out = n;
  return 0;
}

extern "C" int premvos_propnet_set_option(premvos_propnet_t* n, const char* key, int value) {
  PV_CHECK(n && key, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_set_option: null argument");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_propnet_set_option: options must be set before the first set_param");
  std::string k(key);
  if (k.rfind("num_blocks", 0) == 0 && k.size() == 11 && k[10] >= '0' && k[10] <= '3' && value >= 1 && value <= 64) {
    PV_CHECK(n->params.empty(), PREMVOS_ERR_NOT_READY, "premvos_propnet_set_option: set num_blocks before loading parameters");
    n->blocks[k[10] - '0'] = value;
    n->shapes.clear();
    build_shape_table(n);
    return 0;
  }
  if (k == "cuda_graph") { n->opt_cuda_graph = value; return 0; }
  if (k == "mode_mask") {
    PV_CHECK(n->params.empty(), PREMVOS_ERR_NOT_READY, "premvos_propnet_set_option: set mode_mask before loading parameters");
    n->mode_mask = value ? 1 : 0;
    n->shapes.clear();
    build_shape_table(n);
    return 0;
  }
  if (k == "batch") {
    PV_CHECK(value >= 1 && value <= 16, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_set_option: batch in [1,16]");
    n->batch = value;
    return 0;
  }
  return fail(PREMVOS_ERR_INVALID_ARG, "premvos_propnet_set_option: unknown option '%s'", key);
}

extern "C" int premvos_propnet_set_param(premvos_propnet_t* n, const char* name, const float* host_data, int64_t numel) {
  PV_CHECK(n && name && host_data, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_set_param: null argument");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_propnet_set_param: network already finalized");
  auto it = n->shapes.find(name);
  if (it == n->shapes.end()) return fail(PREMVOS_ERR_UNKNOWN_PARAM, "premvos_propnet_set_param: unexpected variable '%s'", name);
  const int64_t want = numel_of(it->second);
  if (numel != want)
    return fail(PREMVOS_ERR_BAD_SHAPE, "premvos_propnet_set_param: '%s' has %lld elements, expected %lld", name, (long long)numel, (long long)want);
  n->params[name].assign(host_data, host_data + numel);
  return 0;
}

extern "C" int premvos_propnet_finalize(premvos_propnet_t* n) {
  PV_CHECK(n, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_finalize: null handle");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_propnet_finalize: already finalized");
  for (auto& kv : n->shapes)
    if (!n->params.count(kv.first)) return fail(PREMVOS_ERR_NOT_READY, "premvos_propnet_finalize: missing variable '%s'", kv.first.c_str());
  PV_CUDA(cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking));
  const char* fork_env = getenv("PREMVOS_PROPNET_FORK");
  if (n->batch > 1 && (fork_env == nullptr || atoi(fork_env) != 0)) {
    PV_CUDA(cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming));
    n->side.resize(n->batch - 1); n->ev_join.resize(n->batch - 1);
    for (int b = 0; b + 1 < n->batch; b++) {
      PV_CUDA(cudaStreamCreateWithFlags(&n->side[b], cudaStreamNonBlocking));
      PV_CUDA(cudaEventCreateWithFlags(&n->ev_join[b], cudaEventDisableTiming));
    }
  }
  PV_TRY(build_network(n));
  n->params.clear();
  const int64_t before = g_launch_count.load();
  PV_TRY(run_network(n, n->stream));  // warm-up: validates every launch configuration
  PV_CUDA(cudaStreamSynchronize(n->stream));
  n->launches_per_forward = (int)(g_launch_count.load() - before);
  if (n->opt_cuda_graph) {
    const int64_t b2 = g_launch_count.load();
    PV_CUDA(cudaStreamBeginCapture(n->stream, cudaStreamCaptureModeThreadLocal));
    int r = run_network(n, n->stream);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(n->stream, &g);
    if (r != 0) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) return fail((int)e, "premvos_propnet_finalize: graph capture failed: %s", cudaGetErrorString(e));
    n->graph = g;
    n->graph_nodes = (int)(g_launch_count.load() - b2);
    g_launch_count.fetch_sub(n->graph_nodes);  // captured, not executed
    PV_CUDA(cudaGraphInstantiate(&n->exec, n->graph, 0));
  }
  n->finalized = true;
  return 0;
}

extern "C" int premvos_propnet_forward(premvos_propnet_t* n, const float* img_dev, void* stream) {
  PV_CHECK(n && img_dev, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_forward: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_propnet_forward: call premvos_propnet_finalize first");
  cudaStream_t st = (cudaStream_t)stream;
  PV_CUDA(cudaMemcpyAsync(n->img_dev, img_dev, (size_t)n->batch * n->H * n->W * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return enqueue_network(n, st);
}

extern "C" int premvos_propnet_forward_u8(premvos_propnet_t* n, const unsigned char* img_bgr_dev, void* stream) {
  PV_CHECK(n && img_bgr_dev, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_forward_u8: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_propnet_forward_u8: call premvos_propnet_finalize first");
  cudaStream_t st = (cudaStream_t)stream;
  PV_TRY(det_u8_to_f32(img_bgr_dev, n->img_dev, (long)n->batch * n->H * n->W * 3, st));
  return enqueue_network(n, st);
}

extern "C" int premvos_propnet_read_results_image(premvos_propnet_t* n, void* stream, int image, int* n_out, float* final_boxes,
                                                  float* final_probs, int64_t* final_labels, float* final_posterior,
                                                  int64_t* second_final_labels, float* second_final_posterior) {
  PV_CHECK(n && n_out, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_read_results: null argument");
  PV_CHECK(image >= 0 && image < n->batch, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_read_results: image %d outside the batch of %d", image,
           n->batch);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t o = (size_t)image * RESULTS_PER_IM;
  PV_CUDA(cudaMemcpyAsync(n_out, n->n_out + image, sizeof(int), cudaMemcpyDeviceToHost, st));
  PV_CUDA(cudaStreamSynchronize(st));
  const int m = *n_out;
  PV_CHECK(m >= 0 && m <= RESULTS_PER_IM, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_read_results: corrupt result count %d", m);
  if (m == 0) return 0;
  if (final_boxes) PV_CUDA(cudaMemcpyAsync(final_boxes, n->final_boxes + o * 4, (size_t)m * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (final_probs) PV_CUDA(cudaMemcpyAsync(final_probs, n->final_probs + o, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (final_labels) PV_CUDA(cudaMemcpyAsync(final_labels, n->final_labels + o, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (final_posterior)
    PV_CUDA(cudaMemcpyAsync(final_posterior, n->final_posterior + o * 2, (size_t)m * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (second_final_labels)
    PV_CUDA(cudaMemcpyAsync(second_final_labels, n->second_final_labels + o, (size_t)m * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (second_final_posterior && n->second_num_class > 0)
    PV_CUDA(cudaMemcpyAsync(second_final_posterior, n->second_final_posterior + o * n->second_num_class,
                            (size_t)m * n->second_num_class * sizeof(float), cudaMemcpyDeviceToHost, st));
  PV_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int premvos_propnet_read_results(premvos_propnet_t* n, void* stream, int* n_out, float* final_boxes, float* final_probs,
                                            int64_t* final_labels, float* final_posterior, int64_t* second_final_labels,
                                            float* second_final_posterior) {
  return premvos_propnet_read_results_image(n, stream, 0, n_out, final_boxes, final_probs, final_labels, final_posterior,
                                            second_final_labels, second_final_posterior);
}

// Device-to-device hand-over of the last forward's results (fixed RESULTS_PER_IM rows + a device count): lets a resident
// pipeline run the next frame on this handle, or feed the boxes to the refinement network, without a host round trip.
extern "C" int premvos_propnet_copy_results(premvos_propnet_t* n, void* stream, int* n_out_dev, float* final_boxes_dev,
                                            float* final_probs_dev, float* second_final_posterior_dev) {
  PV_CHECK(n && n_out_dev, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_copy_results: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_propnet_copy_results: network not finalized");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nb = (size_t)n->batch;   // `batch` images: [batch] counts, [batch][RESULTS_PER_IM] rows
  PV_CUDA(cudaMemcpyAsync(n_out_dev, n->n_out, nb * sizeof(int), cudaMemcpyDeviceToDevice, st));
  if (final_boxes_dev)
    PV_CUDA(cudaMemcpyAsync(final_boxes_dev, n->final_boxes, nb * RESULTS_PER_IM * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (final_probs_dev)
    PV_CUDA(cudaMemcpyAsync(final_probs_dev, n->final_probs, nb * RESULTS_PER_IM * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (second_final_posterior_dev && n->second_num_class > 0)
    PV_CUDA(cudaMemcpyAsync(second_final_posterior_dev, n->second_final_posterior,
                            nb * RESULTS_PER_IM * n->second_num_class * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int premvos_propnet_forward_host(premvos_propnet_t* n, const float* img_host, int* n_out, float* final_boxes, float* final_probs,
                                            int64_t* final_labels, float* final_posterior, int64_t* second_final_labels,
                                            float* second_final_posterior) {
  PV_CHECK(n && img_host && n_out, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_forward_host: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_propnet_forward_host: call premvos_propnet_finalize first");
  PV_CHECK(n->batch == 1, PREMVOS_ERR_UNSUPPORTED,
           "premvos_propnet_forward_host: one image per call; a batched handle is driven with premvos_propnet_forward[_u8] + read_results_image");
  PV_CUDA(cudaMemcpyAsync(n->img_dev, img_host, (size_t)n->batch * n->H * n->W * 3 * sizeof(float), cudaMemcpyHostToDevice, n->stream));
  PV_TRY(enqueue_network(n, n->stream));
  return premvos_propnet_read_results(n, n->stream, n_out, final_boxes, final_probs, final_labels, final_posterior, second_final_labels,
                                      second_final_posterior);
}

extern "C" int premvos_propnet_read_masks(premvos_propnet_t* n, void* stream, int image, float* final_masks, int max_rows) {
  PV_CHECK(n && final_masks, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_read_masks: null argument");
  PV_CHECK(n->finalized && n->mode_mask, PREMVOS_ERR_NOT_READY, "premvos_propnet_read_masks: the handle was not built with option mode_mask");
  PV_CHECK(image >= 0 && image < n->batch && max_rows >= 0, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_read_masks: image %d / rows %d", image, max_rows);
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = max_rows < RESULTS_PER_IM ? max_rows : RESULTS_PER_IM;
  if (rows)
    PV_CUDA(cudaMemcpyAsync(final_masks, n->final_masks + (size_t)image * RESULTS_PER_IM * 196, (size_t)rows * 196 * sizeof(float),
                            cudaMemcpyDeviceToHost, st));
  PV_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// eval.py:35-58 for the n boxes of an image (single-op entry point; host pointers, synchronous)
extern "C" int premvos_fill_full_masks_host(const float* masks_host, const float* boxes_host, int n, int mask_size, int height, int width,
                                            unsigned char* out_host) {
  PV_CHECK(n >= 0 && mask_size >= 2 && mask_size % 2 == 0 && height > 0 && width > 0 && height <= 65535 && n <= 65535, PREMVOS_ERR_INVALID_ARG,
           "premvos_fill_full_masks_host: bad sizes");
  if (n == 0) return 0;
  PV_CHECK(masks_host && boxes_host && out_host, PREMVOS_ERR_INVALID_ARG, "premvos_fill_full_masks_host: null argument");
  float *d_m = nullptr, *d_b = nullptr; unsigned char* d_o = nullptr;
  const size_t mb = (size_t)n * mask_size * mask_size * 4, ob = (size_t)n * height * width;
  PV_CUDA(cudaMalloc((void**)&d_m, mb)); PV_CUDA(cudaMalloc((void**)&d_b, (size_t)n * 16)); PV_CUDA(cudaMalloc((void**)&d_o, ob));
  PV_CUDA(cudaMemcpy(d_m, masks_host, mb, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(d_b, boxes_host, (size_t)n * 16, cudaMemcpyHostToDevice));
  int r = det_fill_full_masks(d_m, d_b, n, mask_size, height, width, d_o, nullptr);
  if (r == 0) {
    cudaError_t e = cudaMemcpy(out_host, d_o, ob, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) r = fail((int)e, "premvos_fill_full_masks_host: %s", cudaGetErrorString(e));
  }
  cudaFree(d_m); cudaFree(d_b); cudaFree(d_o);
  return r;
}

extern "C" int premvos_propnet_launches_per_forward(const premvos_propnet_t* n) { return n ? n->launches_per_forward : 0; }

// Test hook: intermediates of the LAST forward as fp32 (integer tensors converted), see include/premvos_b200.h.
extern "C" int premvos_propnet_get_tensor(premvos_propnet_t* n, const char* name, float* host_out, int64_t* numel) {
  PV_CHECK(n && name && numel, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_get_tensor: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_propnet_get_tensor: network not finalized");
  PV_CUDA(cudaDeviceSynchronize());
  std::string k(name);
  int img = 0;   // "<tensor>@<b>": the per-image tensor of image b of the batch (default image 0); activations hold all images
  const size_t at = k.find('@');
  if (at != std::string::npos) { img = atoi(k.c_str() + at + 1); k = k.substr(0, at); }
  PV_CHECK(img >= 0 && img < n->batch, PREMVOS_ERR_INVALID_ARG, "premvos_propnet_get_tensor: image %d outside the batch of %d", img, n->batch);
  CView cv; bool is_cv = false;
  const float* fptr = nullptr; const int* iptr = nullptr; int64_t cnt = 0;
  int h_topk = 0, h_valid = 0, h_keep = 0;
  PV_CUDA(cudaMemcpy(&h_topk, n->topk_count + img, 4, cudaMemcpyDeviceToHost));
  PV_CUDA(cudaMemcpy(&h_valid, n->valid_count + img, 4, cudaMemcpyDeviceToHost));
  PV_CUDA(cudaMemcpy(&h_keep, n->keep_count + img, 4, cudaMemcpyDeviceToHost));
  if (k == "featuremap") { cv = n->featuremap; is_cv = true; }
  else if (k == "conv0") { cv = n->c0; is_cv = true; }
  else if (k == "pool0") { cv = n->pool; is_cv = true; }
  else if (k.rfind("block", 0) == 0) {   // "block<i>": output of the i-th backbone bottleneck
    const int i = atoi(k.c_str() + 5);
    PV_CHECK(i >= 0 && i < (int)n->backbone.size(), PREMVOS_ERR_INVALID_ARG, "premvos_propnet_get_tensor: no such block");
    cv = n->backbone[i]->out; is_cv = true;
  }
  else if (k == "rpn_hidden") { cv = n->rpn_hidden; is_cv = true; }
  else if (k == "roi_resized") { cv = n->roi; is_cv = true; }
  else if (k == "feature_fastrcnn") { cv = n->head.back()->out; is_cv = true; }
  else if (k == "cell_anchors") { fptr = n->cell_anchors; cnt = 60; }
  else if (k == "final_masks" && n->mode_mask) {
    int m = 0; PV_CUDA(cudaMemcpy(&m, n->n_out + img, 4, cudaMemcpyDeviceToHost));
    fptr = n->final_masks + (size_t)img * RESULTS_PER_IM * 196; cnt = (int64_t)m * 196;
  }
  else if (k == "rpn_out") { fptr = n->rpn_out.p + (size_t)img * n->fh * n->fw * 80; cnt = (int64_t)n->fh * n->fw * 80; }
  else if (k == "rpn_scores") { fptr = n->d_scores + (size_t)img * n->n_anchor_total; cnt = n->n_anchor_total; }
  else if (k == "rpn_decoded_boxes") { fptr = n->d_boxes + (size_t)img * n->n_anchor_total * 4; cnt = (int64_t)n->n_anchor_total * 4; }
  else if (k == "topk_indices") { iptr = n->valid_src + img * S_TOPK; cnt = h_valid; }
  else if (k == "nms_keep") { iptr = n->keep + img * S_KEEP; cnt = h_keep; }
  else if (k == "proposal_boxes") { fptr = n->prop_boxes + img * S_KEEP * 4; cnt = (int64_t)h_keep * 4; }
  else if (k == "proposal_scores") { fptr = n->prop_scores + img * S_KEEP; cnt = h_keep; }
  else if (k == "head_logits") { fptr = n->fc_out + (size_t)img * POST_NMS_TOPK * n->nfc; cnt = (int64_t)POST_NMS_TOPK * n->nfc; }
  else if (k == "pooled") { fptr = n->pooled + (size_t)img * POST_NMS_TOPK * 2048; cnt = (int64_t)POST_NMS_TOPK * 2048; }
  else if (k == "fastrcnn_all_probs") { fptr = n->all_probs + img * S_KEEP * 2; cnt = (int64_t)h_keep * 2; }
  else if (k == "fastrcnn_all_boxes") { fptr = n->all_boxes + img * S_KEEP * 4; cnt = (int64_t)h_keep * 4; }
  else if (k == "final_box_index") { int m = 0; PV_CUDA(cudaMemcpy(&m, n->n_out + img, 4, cudaMemcpyDeviceToHost)); iptr = n->final_box_index + img * RESULTS_PER_IM; cnt = m; }
  else return fail(PREMVOS_ERR_INVALID_ARG, "premvos_propnet_get_tensor: unknown tensor '%s'", name);
  (void)h_topk;
  if (is_cv) {
    *numel = (int64_t)cv.N * cv.C * cv.H * cv.W;
    if (!host_out) return 0;
    float* dtmp = nullptr;
    PV_CUDA(cudaMalloc((void**)&dtmp, (size_t)(*numel) * sizeof(float)));
    int r = cp8_to_nchw(cv, 0, dtmp, nullptr);
    if (r == 0) {
      cudaError_t e = cudaMemcpy(host_out, dtmp, (size_t)(*numel) * sizeof(float), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) r = fail((int)e, "premvos_propnet_get_tensor: %s", cudaGetErrorString(e));
    }
    cudaFree(dtmp);
    return r;
  }
  *numel = cnt;
  if (!host_out || cnt == 0) return 0;
  if (fptr) {
    PV_CUDA(cudaMemcpy(host_out, fptr, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost));
  } else {
    std::vector<int> tmp((size_t)cnt);
    PV_CUDA(cudaMemcpy(tmp.data(), iptr, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < cnt; i++) host_out[i] = (float)tmp[(size_t)i];
  }
  return 0;
}

extern "C" void premvos_propnet_destroy(premvos_propnet_t* n) {
  if (!n) return;
  cudaDeviceSynchronize();
  for (void* p : n->allocs) cudaFree(p);
  free_layer(&n->conv0); free_layer(&n->rpn0); free_layer(&n->rpn_heads); free_layer(&n->deconv);
  for (auto* vec : {&n->backbone, &n->head, &n->head_m})
    for (auto& b : *vec) { free_layer(&b->c1); free_layer(&b->c2); free_layer(&b->c3); free_layer(&b->sc); }
  conv_workspace_free(&n->conv_ws);
  if (n->exec) cudaGraphExecDestroy(n->exec);
  if (n->graph) cudaGraphDestroy(n->graph);
  if (n->stream) cudaStreamDestroy(n->stream);
  for (cudaStream_t s : n->side) cudaStreamDestroy(s);
  for (cudaEvent_t e : n->ev_join) cudaEventDestroy(e);
  if (n->ev_fork) cudaEventDestroy(n->ev_fork);
  delete n;
}

// ---- single-op entry points (parity hooks for the index-exact pieces) ---------------------------------------
// tf.nn.top_k(scores, k) (model.py:189-190): indices of the k largest scores, ordered by (score desc, index asc).
extern "C" int premvos_topk_host(const float* scores_host, int n, int k, int* indices_out, int* count_out) {
  PV_CHECK(scores_host && indices_out && count_out && n >= 1 && k >= 1 && k <= 1024, PREMVOS_ERR_INVALID_ARG,
           "premvos_topk_host: need n >= 1 and 1 <= k <= 1024");
  float *d_s = nullptr, *d_o = nullptr; int *d_i = nullptr, *d_c = nullptr;
  PV_CUDA(cudaMalloc((void**)&d_s, (size_t)n * 4)); PV_CUDA(cudaMalloc((void**)&d_o, 1024 * 4));
  PV_CUDA(cudaMalloc((void**)&d_i, 1024 * 4)); PV_CUDA(cudaMalloc((void**)&d_c, 4));
  PV_CUDA(cudaMemcpy(d_s, scores_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  int r = det_topk(d_s, n, k, d_i, d_o, d_c, nullptr);
  if (r == 0) {
    cudaError_t e = cudaMemcpy(count_out, d_c, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(indices_out, d_i, (size_t)(*count_out) * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) r = fail((int)e, "premvos_topk_host: %s", cudaGetErrorString(e));
  }
  cudaFree(d_s); cudaFree(d_o); cudaFree(d_i); cudaFree(d_c);
  return r;
}

// tf.image.non_max_suppression(boxes, scores, max_output_size, iou_threshold) (model.py:205-209, 466-467) for
// n <= 1024 boxes: greedy by descending score (ties: lower index first), suppress when IoU > iou_threshold.
extern "C" int premvos_nms_host(const float* boxes_host, const float* scores_host, int n, float iou_threshold, int max_output_size,
                                int* selected_out, int* count_out) {
  PV_CHECK(boxes_host && scores_host && selected_out && count_out && n >= 0 && n <= 1024 && max_output_size >= 0,
           PREMVOS_ERR_INVALID_ARG, "premvos_nms_host: need 0 <= n <= 1024");
  *count_out = 0;
  if (n == 0 || max_output_size == 0) return 0;
  float *d_b = nullptr, *d_s = nullptr, *d_so = nullptr, *d_sb = nullptr, *d_ss = nullptr;
  int *d_i = nullptr, *d_c = nullptr, *d_src = nullptr, *d_vc = nullptr, *d_keep = nullptr, *d_kc = nullptr; uint32_t* d_m = nullptr;
  PV_CUDA(cudaMalloc((void**)&d_b, (size_t)n * 16)); PV_CUDA(cudaMalloc((void**)&d_s, (size_t)n * 4));
  PV_CUDA(cudaMalloc((void**)&d_so, 4096)); PV_CUDA(cudaMalloc((void**)&d_sb, 16384)); PV_CUDA(cudaMalloc((void**)&d_ss, 4096));
  PV_CUDA(cudaMalloc((void**)&d_i, 4096)); PV_CUDA(cudaMalloc((void**)&d_c, 4)); PV_CUDA(cudaMalloc((void**)&d_src, 4096));
  PV_CUDA(cudaMalloc((void**)&d_vc, 4)); PV_CUDA(cudaMalloc((void**)&d_keep, 4096)); PV_CUDA(cudaMalloc((void**)&d_kc, 4));
  PV_CUDA(cudaMalloc((void**)&d_m, 1024 * 32 * 4));
  PV_CUDA(cudaMemcpy(d_b, boxes_host, (size_t)n * 16, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(d_s, scores_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  int r = det_topk(d_s, n, n, d_i, d_so, d_c, nullptr);
  // no clipping / filtering here: an "image" large enough and min_size = -inf keep every box, in sorted order
  if (r == 0) r = det_gather_clip_valid(d_b, d_i, d_so, d_c, INFINITY, INFINITY, -INFINITY, d_sb, d_ss, d_src, d_vc, nullptr);
  const int mo = max_output_size < 1024 ? max_output_size : 1024;
  if (r == 0) r = det_nms(d_sb, d_vc, iou_threshold, mo, d_m, d_keep, d_kc, nullptr);
  if (r == 0) {
    std::vector<int> keep(1024), src(1024);
    cudaError_t e = cudaMemcpy(count_out, d_kc, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(keep.data(), d_keep, 4096, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(src.data(), d_src, 4096, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) r = fail((int)e, "premvos_nms_host: %s", cudaGetErrorString(e));
    else for (int i = 0; i < *count_out; i++) selected_out[i] = src[keep[i]];
  }
  for (void* p : {(void*)d_b, (void*)d_s, (void*)d_so, (void*)d_sb, (void*)d_ss, (void*)d_i, (void*)d_c, (void*)d_src, (void*)d_vc,
                  (void*)d_keep, (void*)d_kc, (void*)d_m}) cudaFree(p);
  return r;
}
