// Refinement network forward (DeepLabv3+ with an Xception-65 backbone on 385x385 box crops, 4 input channels =
// RGB + box guidance): host-side orchestration and C ABI.
//
// Restates refinement_net's inference path (network/deeplab/DeepLabV3Plus.py:13-39, deeplab/model.py:200-707,
// deeplab/core/xception.py:70-560, network/SegmentationOutputLayers.py:35-61,106-135, datasets/Dataset.py:141-186,
// datasets/Resize.py:150-193) as a fixed list of launches over pre-allocated CP8 buffers:
//   * all proposals of a frame are cropped/resized on the device from ONE uploaded uint8 frame and run as a batch
//     (the reference runs batch 1 and re-feeds the whole float frame per proposal, configs/run:16-17);
//   * every BatchNorm (frozen: "freeze_batchnorm": true, eps 1e-3 in the backbone, 1e-5 in ASPP/decoder) is folded
//     into the preceding depthwise / pointwise / regular convolution;
//   * separable convolution = depthwise kernel (unit ReLU in, BN, optional ReLU out fused) + pointwise 1x1 GEMM on
//     tcgen05 (conv_umma.cu) whose epilogue adds the unit's shortcut;
//   * ASPP branches and the decoder concat write straight into chunk ranges of one buffer (no tf.concat copies);
//   * the output layer, the paste-back into the frame and the conf_score reduction are one kernel.
#include <math.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

using namespace premvos;

namespace {

struct BlockSpec { const char* scope; int depth[3]; int skip; /*0 none 1 conv 2 sum*/ bool act_in_sep; int units; int stride; };
const float XC_EPS = 1e-3f, ASPP_EPS = 1e-5f;

typedef std::function<int(cudaStream_t, int)> Step;

}  // namespace

struct premvos_refnet {
  int NB = 0, S = 0, middle_units = 16, n_classes = 2;
  int stem_rows = 1;   // conv1_1 as a 3x1 convolution over the row im2col of the network input (see build_network)
  std::map<std::string, std::function<int(cudaStream_t, int)>> lazy;   // test-hook tensors that only exist on request
  std::vector<std::unique_ptr<SepConvPlan>> sep_plans;
  bool finalized = false;
  std::map<std::string, std::vector<float>> params;
  std::map<std::string, std::vector<int64_t>> shapes;
  std::vector<void*> allocs;
  std::vector<std::unique_ptr<ConvWeightsUmma>> conv_weights;
  std::vector<std::unique_ptr<ConvPlanUmma>> conv_plans;
  std::vector<std::unique_ptr<DwF8Plan>> dw_plans;
  ConvWorkspace conv_ws;   // stream-K scratch shared by the pointwise layers (they run one after the other on one stream)
  std::vector<Step> steps;
  std::map<std::string, CView> named;   // test hook
  cudaStream_t stream = nullptr;
  CView input;
  TView logits;
  float* boxes_dev = nullptr; int* crops = nullptr; double* conf_sum = nullptr;
  unsigned char* frame_dev = nullptr; size_t frame_cap = 0;
  unsigned char* mask_dev = nullptr; size_t mask_cap = 0;
  float* post_dev = nullptr; size_t post_cap = 0;
  void* hboxes_dev = nullptr; size_t hboxes_cap = 0;   // forward_host staging: boxes / conf_scores of one frame
  void* hconf_dev = nullptr; size_t hconf_cap = 0;
  cudaGraph_t graph = nullptr;                          // the network body for a full launch group (na == NB)
  cudaGraphExec_t exec = nullptr;
  int graph_nodes = 0, opt_cuda_graph = 1;
  int launches_per_forward = 0;
};

namespace {

std::vector<BlockSpec> block_specs(int middle_units) {
  return {{"entry_flow/block1", {128, 128, 128}, 1, false, 1, 2},   {"entry_flow/block2", {256, 256, 256}, 1, false, 1, 2},
          {"entry_flow/block3", {728, 728, 728}, 1, false, 1, 2},   {"middle_flow/block1", {728, 728, 728}, 2, false, middle_units, 1},
          {"exit_flow/block1", {728, 1024, 1024}, 1, false, 1, 2},  {"exit_flow/block2", {1536, 1536, 2048}, 0, true, 1, 1}};
}

int64_t numel_of(const std::vector<int64_t>& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}

void build_shape_table(premvos_refnet* n) {
  auto bn = [&](const std::string& s, int c) {
    for (const char* v : {"gamma", "beta", "moving_mean", "moving_variance"}) n->shapes[s + "/BatchNorm/" + v] = {c};
  };
  auto conv = [&](const std::string& s, int k, int cin, int cout) { n->shapes[s + "/weights"] = {k, k, cin, cout}; bn(s, cout); };
  auto sep = [&](const std::string& s, int cin, int cout) {
    n->shapes[s + "_depthwise/depthwise_weights"] = {3, 3, cin, 1};
    bn(s + "_depthwise", cin);
    conv(s + "_pointwise", 1, cin, cout);
  };
  const std::string x = "xception_65/";
  conv(x + "entry_flow/conv1_1", 3, 4, 32);
  conv(x + "entry_flow/conv1_2", 3, 32, 64);
  int cin = 64;
  for (const BlockSpec& b : block_specs(n->middle_units))
    for (int u = 0; u < b.units; u++) {
      const std::string s = x + b.scope + "/unit_" + std::to_string(u + 1) + "/xception_module";
      int c = cin;
      for (int i = 0; i < 3; i++) { sep(s + "/separable_conv" + std::to_string(i + 1), c, b.depth[i]); c = b.depth[i]; }
      if (b.skip == 1) conv(s + "/shortcut", 1, cin, b.depth[2]);
      cin = b.depth[2];
    }
  conv("image_pooling", 1, cin, 256);
  conv("aspp0", 1, cin, 256);
  for (int i = 1; i <= 3; i++) sep("aspp" + std::to_string(i), cin, 256);
  conv("concat_projection", 1, 1280, 256);
  conv("decoder/feature_projection0", 1, 256, 48);
  sep("decoder/decoder_conv0", 304, 256);
  sep("decoder/decoder_conv1", 256, 256);
  n->shapes["logits/features/weights"] = {1, 1, 256, n->n_classes};
  n->shapes["logits/features/biases"] = {n->n_classes};
}

template <typename T>
int dev_alloc(premvos_refnet* n, T** p, size_t count) {
  PV_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  PV_CUDA(cudaMemset(*p, 0, count * sizeof(T)));
  n->allocs.push_back(*p);
  return 0;
}

int alloc_cview(premvos_refnet* n, CView* v, int C, int H, int W) {
  v->N = n->NB; v->H = H; v->W = W; v->chunks = (C + 7) / 8; v->c0 = 0; v->C = C;
  const size_t elems = (size_t)v->N * v->chunks * H * W * 8 + 64;  // + 128 B slack for flattened 1x1 layers (conv_umma.cu)
  PV_TRY(dev_alloc(n, &v->hi, elems));
  PV_TRY(dev_alloc(n, &v->lo, elems));
  return 0;
}

int alloc_fview(premvos_refnet* n, FView* v, int C, int H, int W) {
  v->N = n->NB; v->H = H; v->W = W; v->chunks = (C + 7) / 8; v->c0 = 0; v->C = C;
  PV_TRY(dev_alloc(n, &v->p, (size_t)v->N * v->chunks * H * W * 8 + 64));
  return 0;
}

// BatchNorm inference folded to (scale, shift): slim.batch_norm, (x - mean) * rsqrt(var + eps) * gamma + beta
void bn_fold(premvos_refnet* n, const std::string& scope, float eps, std::vector<float>* scale, std::vector<float>* shift) {
  const std::vector<float>&g = n->params[scope + "/BatchNorm/gamma"], &b = n->params[scope + "/BatchNorm/beta"];
  const std::vector<float>&m = n->params[scope + "/BatchNorm/moving_mean"], &v = n->params[scope + "/BatchNorm/moving_variance"];
  scale->resize(g.size()); shift->resize(g.size());
  for (size_t i = 0; i < g.size(); i++) {
    (*scale)[i] = g[i] / sqrtf(v[i] + eps);
    (*shift)[i] = b[i] - m[i] * (*scale)[i];
  }
}

// regular / pointwise convolution (+ folded BN or bias) on tcgen05; appends the launch to the step list
int add_conv(premvos_refnet* n, const std::string& scope, bool bn, float eps, const CView& in, const ConvOut& out, ConvGeom g,
             const int* cin_map = nullptr, int cin_phys = 0) {
  const std::vector<int64_t>& ws = n->shapes[scope + "/weights"];
  const int kh = (int)ws[0], kw = (int)ws[1], cin = (int)ws[2], cout = (int)ws[3];
  const std::vector<float>& W = n->params[scope + "/weights"];
  std::vector<float> w((size_t)cout * cin * kh * kw), scale(cout, 1.f), shift(cout, 0.f);
  if (bn) bn_fold(n, scope, eps, &scale, &shift);
  else if (n->params.count(scope + "/biases")) shift = n->params[scope + "/biases"];
  for (int y = 0; y < kh; y++)
    for (int x = 0; x < kw; x++)
      for (int i = 0; i < cin; i++)
        for (int o = 0; o < cout; o++)
          w[(((size_t)o * cin + i) * kh + y) * kw + x] = W[(((size_t)y * kw + x) * cin + i) * cout + o] * scale[o];
  n->conv_weights.emplace_back(new ConvWeightsUmma());
  n->conv_plans.emplace_back(new ConvPlanUmma());
  ConvWeightsUmma* cw = n->conv_weights.back().get();
  ConvPlanUmma* pl = n->conv_plans.back().get();
  const long m_out = out.cp.hi ? (long)out.cp.N * out.cp.H * out.cp.W
                               : (out.f8.p ? (long)out.f8.N * out.f8.H * out.f8.W : (long)out.f32.N * out.f32.H * out.f32.W);
  const bool flat = kh == 1 && kw == 1 && g.stride == 1 && g.pad_t == 0 && g.pad_l == 0 && g.pad_b == 0 && g.pad_r == 0;
  PV_TRY(pack_conv_weights_umma(cw, w.data(), shift.data(), cout, cin, kh, kw, cin_map, cin_phys, 0, m_out, flat));
  PV_TRY(plan_conv_umma(pl, in, out, *cw, g, &n->conv_ws));
  n->steps.push_back([pl](cudaStream_t st, int na) { return launch_conv_umma(*pl, st, na); });
  return 0;
}

// depthwise 3x3 + folded BN (+ ReLU in / out)
int add_depthwise(premvos_refnet* n, const std::string& scope, float eps, const CView& in, const CView& out, int stride, int rate,
                  bool pre_relu, bool post_relu) {
  const int C = in.C, cpad = round_up(C, 8);
  const std::vector<float>& W = n->params[scope + "/depthwise_weights"];  // [3][3][C][1]
  std::vector<float> scale, shift, w((size_t)9 * cpad, 0.f), b(cpad, 0.f);
  bn_fold(n, scope, eps, &scale, &shift);
  for (int t = 0; t < 9; t++)
    for (int c = 0; c < C; c++) w[(size_t)t * cpad + c] = W[(size_t)t * C + c] * scale[c];
  for (int c = 0; c < C; c++) b[c] = shift[c];
  float *dw = nullptr, *db = nullptr;
  PV_TRY(dev_alloc(n, &dw, w.size())); PV_TRY(dev_alloc(n, &db, b.size()));
  PV_CUDA(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  CView i2 = in, o2 = out;
  n->steps.push_back([=](cudaStream_t st, int na) { return depthwise3x3_cp8(i2, o2, dw, db, stride, rate, rate, pre_relu, post_relu, na, st); });
  return 0;
}

// the same on an F8 input (a pointwise output that is not a tensor-core operand): TMA-pipelined kernel, dw_f8.cu
int add_depthwise_f8(premvos_refnet* n, const std::string& scope, float eps, const FView& in, const CView& out, int stride, int rate,
                     bool pre_relu, bool post_relu) {
  const int C = in.C, cpad = round_up(C, 8);
  const std::vector<float>& W = n->params[scope + "/depthwise_weights"];  // [3][3][C][1]
  std::vector<float> scale, shift, w((size_t)9 * cpad, 0.f), b(cpad, 0.f);
  bn_fold(n, scope, eps, &scale, &shift);
  for (int t = 0; t < 9; t++)
    for (int c = 0; c < C; c++) w[(size_t)t * cpad + c] = W[(size_t)t * C + c] * scale[c];
  for (int c = 0; c < C; c++) b[c] = shift[c];
  float *dw = nullptr, *db = nullptr;
  PV_TRY(dev_alloc(n, &dw, w.size())); PV_TRY(dev_alloc(n, &db, b.size()));
  PV_CUDA(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  n->dw_plans.emplace_back(new DwF8Plan());
  DwF8Plan* pl = n->dw_plans.back().get();
  PV_TRY(plan_depthwise3x3_f8(pl, in, out, dw, db, stride, rate, rate, pre_relu, post_relu));
  n->steps.push_back([pl](cudaStream_t st, int na) { return launch_depthwise3x3_f8(*pl, na, st); });
  return 0;
}

// depthwise 3x3 + BN fused into the pointwise 1x1 + BN GEMM (conv_umma.cu: sepconv_fused_kernel): F8 in, no depthwise output in HBM
int add_sepconv_fused(premvos_refnet* n, const std::string& dw_scope, const std::string& pw_scope, float eps, const FView& in, const ConvOut& out,
                      bool pre_relu, bool post_relu, float slope) {
  const int C = in.C, cpad = round_up((C + 7) / 8, 4) * 8;
  const std::vector<float>& Wd = n->params[dw_scope + "/depthwise_weights"];  // [3][3][C][1]
  std::vector<float> scale, shift, w((size_t)9 * cpad, 0.f), b(cpad, 0.f);
  bn_fold(n, dw_scope, eps, &scale, &shift);
  for (int t = 0; t < 9; t++)
    for (int c = 0; c < C; c++) w[(size_t)t * cpad + c] = Wd[(size_t)t * C + c] * scale[c];
  for (int c = 0; c < C; c++) b[c] = shift[c];
  float *dw = nullptr, *db = nullptr;
  PV_TRY(dev_alloc(n, &dw, w.size())); PV_TRY(dev_alloc(n, &db, b.size()));
  PV_CUDA(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  const std::vector<int64_t>& ws = n->shapes[pw_scope + "/weights"];   // [1][1][C][Cout]
  const int cout = (int)ws[3];
  const std::vector<float>& Wp = n->params[pw_scope + "/weights"];
  std::vector<float> pscale, pshift, wp((size_t)cout * C);
  bn_fold(n, pw_scope, eps, &pscale, &pshift);
  for (int i = 0; i < C; i++)
    for (int o = 0; o < cout; o++) wp[(size_t)o * C + i] = Wp[(size_t)i * cout + o] * pscale[o];
  n->conv_weights.emplace_back(new ConvWeightsUmma());
  ConvWeightsUmma* cw = n->conv_weights.back().get();
  PV_TRY(pack_conv_weights_umma(cw, wp.data(), pshift.data(), cout, C, 1, 1, nullptr, 0, 4, 0, false));
  n->sep_plans.emplace_back(new SepConvPlan());
  SepConvPlan* pl = n->sep_plans.back().get();
  PV_TRY(plan_sepconv_fused(pl, in, dw, db, cpad, pre_relu, post_relu, *cw, out, slope));
  n->steps.push_back([pl](cudaStream_t st, int na) { return launch_sepconv_fused(*pl, na, st); });
  return 0;
}

int build_network(premvos_refnet* n) {
  const int S = n->S;
  const std::string x = "xception_65/";
  const int map4[4] = {0, 1, 2, 3};
  // root: conv2d_same 3x3 s2 (explicit pad 1,1 + VALID) and 3x3 s1 SAME (xception.py:430-433)
  const int S1 = (S + 2 - 3) / 2 + 1;
  // With 4 input channels a per-tap K step would be 3/4 zeros; the input kernel therefore writes the ROW IM2COL of the network
  // input ([S][S1] pixels of 3 horizontal taps x 4 channels, horizontal stride and padding applied) and conv1_1 runs as a 3x1
  // convolution over it (vertical stride 2, horizontal stride 1): 3 operand boxes per tile instead of 9, K = 16 with 12 used.
  n->stem_rows = !(getenv("PREMVOS_STEM_ROWS") && atoi(getenv("PREMVOS_STEM_ROWS")) == 0);
  if (n->stem_rows) PV_TRY(alloc_cview(n, &n->input, 16, S, S1));
  else PV_TRY(alloc_cview(n, &n->input, 8, S, S));
  const bool use_f8 = getenv("PREMVOS_REFNET_F8") ? atoi(getenv("PREMVOS_REFNET_F8")) != 0 : true;
  const bool fuse_sep = use_f8 && (getenv("PREMVOS_REFNET_FUSE_SEP") ? atoi(getenv("PREMVOS_REFNET_FUSE_SEP")) != 0 : true);
  CView c11, c12;
  FView c12f;   // conv1_2's output a second time in F8: the input of the fused first separable convolution (the CP8 copy feeds the shortcut)
  PV_TRY(alloc_cview(n, &c11, 32, S1, S1));
  PV_TRY(alloc_cview(n, &c12, 64, S1, S1));
  {
    ConvGeom g; g.stride = 2; g.pad_t = g.pad_l = g.pad_b = g.pad_r = 1; g.slope = 0.f;
    ConvOut o; o.cp = c11;
    if (n->stem_rows) {
      const std::string scope = x + "entry_flow/conv1_1";
      const std::vector<float>& W = n->params[scope + "/weights"];   // HWIO [3][3][4][32]
      std::vector<float> w((size_t)32 * 12 * 3), scale(32, 1.f), shift(32, 0.f);
      bn_fold(n, scope, XC_EPS, &scale, &shift);
      for (int oc = 0; oc < 32; oc++)
        for (int r = 0; r < 3; r++)
          for (int sx = 0; sx < 3; sx++)
            for (int c = 0; c < 4; c++) w[((size_t)oc * 12 + sx * 4 + c) * 3 + r] = W[(((size_t)r * 3 + sx) * 4 + c) * 32 + oc] * scale[oc];   // [Cout][12][3][1]
      g.stride_x = 1; g.pad_l = g.pad_r = 0;
      n->conv_weights.emplace_back(new ConvWeightsUmma());
      n->conv_plans.emplace_back(new ConvPlanUmma());
      ConvWeightsUmma* cw = n->conv_weights.back().get();
      ConvPlanUmma* pl = n->conv_plans.back().get();
      PV_TRY(pack_conv_weights_umma(cw, w.data(), shift.data(), 32, 12, 3, 1, nullptr, 0, 0, (long)n->NB * S1 * S1, false));
      PV_TRY(plan_conv_umma(pl, n->input, o, *cw, g, &n->conv_ws));
      n->steps.push_back([pl](cudaStream_t st, int na) { return launch_conv_umma(*pl, st, na); });
    } else {
      PV_TRY(add_conv(n, x + "entry_flow/conv1_1", true, XC_EPS, n->input, o, g, map4, 8));
    }
    ConvOut o2; o2.cp = c12;
    if (fuse_sep && (long)S1 * S1 >= 4096) { PV_TRY(alloc_fview(n, &c12f, 64, S1, S1)); o2.f8 = c12f; }
    PV_TRY(add_conv(n, x + "entry_flow/conv1_2", true, XC_EPS, c11, o2, ConvGeom::same3x3(1, 0.f)));
  }
  // Formats (see dw_f8.cu): a tensor that is a tensor-core operand (input of a shortcut / ASPP / decoder convolution) is CP8; a
  // pointwise output consumed only by the next depthwise convolution is F8 with the consumer's leading ReLU already applied; a
  // unit output consumed by the next unit's first depthwise + sum skip (middle flow, exit block 2) is F8, raw.
  CView cur = c12, low_level;
  FView cur_f = c12f;   // the unit input when it is F8 (cur is null then; the root's output exists in both formats when it is fused)
  const int target = 16 / 2;  // output_stride 16, halved by the stride-2 root conv (xception.py:424-429)
  int current_stride = 1, rate = 1;
  const std::vector<BlockSpec> specs = block_specs(n->middle_units);
  for (size_t bi = 0; bi < specs.size(); bi++) {
    const BlockSpec& b = specs[bi];
    for (int u = 0; u < b.units; u++) {
      const std::string s = x + b.scope + "/unit_" + std::to_string(u + 1) + "/xception_module";
      int stride = b.stride, unit_rate = 1;
      if (current_stride == target) { stride = 1; unit_rate = rate; rate *= b.stride; }   // xception.py:341-355
      else current_stride *= b.stride;
      const int inH = cur_f.p ? cur_f.H : cur.H, inW = cur_f.p ? cur_f.W : cur.W, inC = cur_f.p ? cur_f.C : cur.C;
      const int Ho = stride == 2 ? (inH - 1) / 2 + 1 : inH, Wo = stride == 2 ? (inW - 1) / 2 + 1 : inW;
      // what consumes this unit's output: the next unit (same block or first unit of the next block) or the ASPP
      const BlockSpec* next = (u + 1 < b.units) ? &b : nullptr;
      for (size_t bj = bi + 1; !next && bj < specs.size(); bj++)
        if (specs[bj].units > 0) next = &specs[bj];
      const bool out_f8 = use_f8 && next && next->skip != 1;   // no shortcut convolution reads it
      CView sc;
      if (b.skip == 1) {  // 1x1 stride-s shortcut + BN, no activation
        PV_CHECK(!cur.null(), PREMVOS_ERR_INVALID_ARG, "refnet: shortcut convolution needs a CP8 input");
        PV_TRY(alloc_cview(n, &sc, b.depth[2], Ho, Wo));
        ConvGeom g; g.stride = stride;
        ConvOut o; o.cp = sc;
        PV_TRY(add_conv(n, s + "/shortcut", true, XC_EPS, cur, o, g));
      }
      CView t = cur;    // input of the next separable convolution: CP8 ...
      FView tf = cur_f;  // ... or F8
      bool t_relu_applied = false;   // the F8 producer already applied the consumer's leading ReLU
      int tC = inC, tH = inH, tW = inW;
      for (int i = 0; i < 3; i++) {
        const int st_i = i == 2 ? stride : 1;
        const int h_i = st_i == 2 ? Ho : tH, w_i = st_i == 2 ? Wo : tW;
        const std::string ss = s + "/separable_conv" + std::to_string(i + 1);
        // HBM-bound separable convolutions with ONE output-channel tile: the depthwise tile is computed inside the pointwise GEMM
        const bool has_res = i == 2 && b.skip != 0;
        const bool fuse = fuse_sep && tf.p && st_i == 1 && unit_rate == 1 && b.depth[i] <= 128 && b.depth[i] % 32 == 0 && !has_res &&
                          (long)h_i * w_i >= 4096;
        const bool dw_pre_relu = !b.act_in_sep && !(tf.p && t_relu_applied);
        CView d;
        if (!fuse) {
          PV_TRY(alloc_cview(n, &d, tC, h_i, w_i));
          if (tf.p)
            PV_TRY(add_depthwise_f8(n, ss + "_depthwise", XC_EPS, tf, d, st_i, unit_rate, dw_pre_relu, b.act_in_sep));
          else
            PV_TRY(add_depthwise(n, ss + "_depthwise", XC_EPS, t, d, st_i, unit_rate, !b.act_in_sep, b.act_in_sep));
        }
        const bool is_low_level = std::string(b.scope) == "entry_flow/block2" && i == 1;   // feature_extractor.py:89-94
        ConvGeom g; g.slope = b.act_in_sep ? 0.f : 1.f;
        ConvOut o;
        CView p; FView pf;
        if (i < 2 && use_f8 && !is_low_level) {   // only the next depthwise reads it: F8, ReLU of the next separable conv applied here
          PV_TRY(alloc_fview(n, &pf, b.depth[i], h_i, w_i));
          o.f8 = pf; g.slope = 0.f; t_relu_applied = true;
        } else if (i == 2 && out_f8) {
          PV_TRY(alloc_fview(n, &pf, b.depth[i], h_i, w_i));
          o.f8 = pf; t_relu_applied = false;
        } else {
          PV_TRY(alloc_cview(n, &p, b.depth[i], h_i, w_i));
          o.cp = p; t_relu_applied = false;
        }
        if (i == 2 && b.skip == 1) o.res = sc;
        if (i == 2 && b.skip == 2) { if (cur_f.p) o.res_f8 = cur_f; else o.res = cur; }
        if (fuse) PV_TRY(add_sepconv_fused(n, ss + "_depthwise", ss + "_pointwise", XC_EPS, tf, o, dw_pre_relu, b.act_in_sep, g.slope));
        else PV_TRY(add_conv(n, ss + "_pointwise", true, XC_EPS, d, o, g));
        if (is_low_level) low_level = p;
        t = p; tf = pf; tC = b.depth[i]; tH = h_i; tW = w_i;
      }
      cur = t; cur_f = tf;
    }
  }
  PV_CHECK(!cur.null(), PREMVOS_ERR_INVALID_ARG, "refnet: the backbone output must be CP8");
  n->named["xception_out"] = cur;
  n->named["low_level"] = low_level;
  const int fh = cur.H, fw = cur.W;
  // ---- ASPP (model.py:361-435) ----
  CView cat, aspp_out;
  PV_TRY(alloc_cview(n, &cat, 1280, fh, fw));
  PV_TRY(alloc_cview(n, &aspp_out, 256, fh, fw));
  {
    // image pooling: global mean -> 1x1 conv + BN + ReLU -> broadcast
    std::vector<float> scale, shift;
    bn_fold(n, "image_pooling", ASPP_EPS, &scale, &shift);
    const std::vector<float>& W = n->params["image_pooling/weights"];  // [1][1][2048][256] == [in][out]
    std::vector<float> w(W.size());
    const int cin = cur.C;
    for (int i = 0; i < cin; i++)
      for (int o = 0; o < 256; o++) w[(size_t)i * 256 + o] = W[(size_t)i * 256 + o] * scale[o];
    float *dw = nullptr, *db = nullptr, *vec = nullptr, *pooled = nullptr;
    PV_TRY(dev_alloc(n, &dw, w.size())); PV_TRY(dev_alloc(n, &db, (size_t)256)); PV_TRY(dev_alloc(n, &vec, (size_t)n->NB * 256));
    PV_TRY(dev_alloc(n, &pooled, (size_t)n->NB * cur.C));
    PV_CUDA(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    PV_CUDA(cudaMemcpy(db, shift.data(), 256 * 4, cudaMemcpyHostToDevice));
    CView feat = cur, dst = cat.slice(0, 256);
    n->steps.push_back([=](cudaStream_t st, int na) {
      PV_TRY(gap_fc_relu(feat, dw, db, 256, true, pooled, vec, na, st));
      return broadcast_vec_cp8(vec, 256, false, dst, na, st);
    });
  }
  {
    ConvGeom g; g.slope = 0.f;
    ConvOut o; o.cp = cat.slice(32, 256);
    PV_TRY(add_conv(n, "aspp0", true, ASPP_EPS, cur, o, g));
    const int rates[3] = {6, 12, 18};
    for (int i = 1; i <= 3; i++) {
      CView d;
      PV_TRY(alloc_cview(n, &d, cur.C, fh, fw));
      PV_TRY(add_depthwise(n, "aspp" + std::to_string(i) + "_depthwise", ASPP_EPS, cur, d, 1, rates[i - 1], false, true));
      ConvOut o2; o2.cp = cat.slice(32 * (i + 1), 256);
      PV_TRY(add_conv(n, "aspp" + std::to_string(i) + "_pointwise", true, ASPP_EPS, d, o2, g));
    }
    ConvOut o3; o3.cp = aspp_out;
    PV_TRY(add_conv(n, "concat_projection", true, ASPP_EPS, cat, o3, g));
  }
  n->named["aspp_concat"] = cat;
  n->named["aspp_out"] = aspp_out;
  // ---- decoder (model.py:503-598) ----
  const int dh = (int)(((float)S - 1.0f) * 0.25f + 1.0f);
  CView low48, dec_in, d0, p0, d1, p1;
  PV_TRY(alloc_cview(n, &low48, 48, low_level.H, low_level.W));
  PV_TRY(alloc_cview(n, &dec_in, 304, dh, dh));
  PV_TRY(alloc_cview(n, &d0, 304, dh, dh)); PV_TRY(alloc_cview(n, &p0, 256, dh, dh));
  PV_TRY(alloc_cview(n, &d1, 256, dh, dh)); PV_TRY(alloc_cview(n, &p1, 256, dh, dh));
  {
    ConvGeom g; g.slope = 0.f;
    ConvOut o; o.cp = low48;
    PV_TRY(add_conv(n, "decoder/feature_projection0", true, ASPP_EPS, low_level, o, g));
    CView a_src = aspp_out, a_dst = dec_in.slice(0, 256), l_src = low48, l_dst = dec_in.slice(32, 48);
    auto make_concat = [=](cudaStream_t st, int na) {   // model.py:566-577: resize both to the decoder size, concat (views)
      PV_TRY(resize_bilinear_ac_cp8(a_src, a_dst, na, st));
      return resize_bilinear_ac_cp8(l_src, l_dst, na, st);
    };
    const bool fuse_up = low48.H == dh && low48.W == dh && !(getenv("PREMVOS_REFNET_FUSE_UP") && atoi(getenv("PREMVOS_REFNET_FUSE_UP")) == 0);
    if (fuse_up) {
      // The concat is consumed by decoder_conv0's depthwise convolution only: that layer runs as two launches over its channel
      // ranges -- 256 channels sampled straight from the 25 x 25 ASPP output while staging (the 97 x 97 x 256 resized tensor, 385 MB
      // per 40 crops written and read back, never exists), 48 channels from the projected low-level features, which already
      // have the decoder size.  The concat buffer is only filled when the test hook asks for it.
      n->lazy["decoder_in"] = make_concat;
      const std::string scope = "decoder/decoder_conv0_depthwise";
      const int C = 304, cpad = 304;
      const std::vector<float>& W = n->params[scope + "/depthwise_weights"];  // [3][3][C][1]
      std::vector<float> scale, shift, w((size_t)9 * cpad, 0.f), b(cpad, 0.f);
      bn_fold(n, scope, ASPP_EPS, &scale, &shift);
      for (int t = 0; t < 9; t++)
        for (int c = 0; c < C; c++) w[(size_t)t * cpad + c] = W[(size_t)t * C + c] * scale[c];
      for (int c = 0; c < C; c++) b[c] = shift[c];
      float *dw = nullptr, *db = nullptr;
      PV_TRY(dev_alloc(n, &dw, w.size())); PV_TRY(dev_alloc(n, &db, b.size()));
      PV_CUDA(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
      PV_CUDA(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
      CView up_geom = a_dst;            // geometry of the resized tensor (its planes are not touched)
      CView d0a = d0.slice(0, 256), d0l = d0.slice(32, 48);
      n->steps.push_back([=](cudaStream_t st, int na) {
        PV_TRY(depthwise3x3_cp8_ex(up_geom, &a_src, d0a, dw, db, 0, cpad, 1, 1, 1, false, true, na, st));
        return depthwise3x3_cp8_ex(l_src, nullptr, d0l, dw, db, 32, cpad, 1, 1, 1, false, true, na, st);
      });
    } else {
      n->steps.push_back(make_concat);
      PV_TRY(add_depthwise(n, "decoder/decoder_conv0_depthwise", ASPP_EPS, dec_in, d0, 1, 1, false, true));
    }
    ConvOut o0;
    FView p0f;
    if (use_f8) { PV_TRY(alloc_fview(n, &p0f, 256, dh, dh)); o0.f8 = p0f; } else o0.cp = p0;
    PV_TRY(add_conv(n, "decoder/decoder_conv0_pointwise", true, ASPP_EPS, d0, o0, g));
    if (use_f8) PV_TRY(add_depthwise_f8(n, "decoder/decoder_conv1_depthwise", ASPP_EPS, p0f, d1, 1, 1, false, true));
    else PV_TRY(add_depthwise(n, "decoder/decoder_conv1_depthwise", ASPP_EPS, p0, d1, 1, 1, false, true));
    ConvOut o1; o1.cp = p1;
    PV_TRY(add_conv(n, "decoder/decoder_conv1_pointwise", true, ASPP_EPS, d1, o1, g));
  }
  n->named["decoder_in"] = dec_in;
  n->named["decoder_out"] = p1;
  // logits: 1x1 conv with bias, no BN (model.py:639-659); the resize to [dh, dw] (:295-297) is the identity
  n->logits.N = n->NB; n->logits.H = dh; n->logits.W = dh; n->logits.cs = 16; n->logits.coff = 0; n->logits.C = n->n_classes;
  PV_TRY(dev_alloc(n, &n->logits.p, (size_t)n->NB * dh * dh * 16));
  {
    ConvGeom g;
    ConvOut o; o.f32 = n->logits;
    PV_TRY(add_conv(n, "logits/features", false, 0.f, p1, o, g));
  }
  PV_TRY(dev_alloc(n, &n->boxes_dev, (size_t)n->NB * 4));
  PV_TRY(dev_alloc(n, &n->crops, (size_t)n->NB * 4));
  PV_TRY(dev_alloc(n, &n->conf_sum, (size_t)n->NB));
  return 0;
}

int ensure(void** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  PV_CUDA(cudaMalloc(p, need));
  *cap = need;
  return 0;
}

int run_batch(premvos_refnet* n, int na, cudaStream_t st) {
  for (auto& s : n->steps) PV_TRY(s(st, na));
  return 0;
}

}  // namespace

extern "C" int premvos_refnet_create(premvos_refnet_t** out, int max_batch, int input_size, int middle_units) {
  PV_CHECK(out, PREMVOS_ERR_INVALID_ARG, "premvos_refnet_create: out is null");
  *out = nullptr;
  PV_CHECK(max_batch >= 1 && max_batch <= 256 && input_size >= 33 && input_size <= 1025 && middle_units >= 0 && middle_units <= 16,
           PREMVOS_ERR_INVALID_ARG, "premvos_refnet_create: max_batch in [1,256], input_size in [33,1025], middle_units in [0,16]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(PREMVOS_ERR_NO_DEVICE, "premvos_refnet_create: no CUDA device visible");
  premvos_refnet* n = new premvos_refnet();
  n->NB = max_batch; n->S = input_size; n->middle_units = middle_units;
  build_shape_table(n);
  *out = n;
  return 0;
}

extern "C" int premvos_refnet_set_param(premvos_refnet_t* n, const char* name, const float* host_data, int64_t numel) {
  PV_CHECK(n && name && host_data, PREMVOS_ERR_INVALID_ARG, "premvos_refnet_set_param: null argument");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_refnet_set_param: network already finalized");
  auto it = n->shapes.find(name);
  if (it == n->shapes.end()) return fail(PREMVOS_ERR_UNKNOWN_PARAM, "premvos_refnet_set_param: unexpected variable '%s'", name);
  const int64_t want = numel_of(it->second);
  if (numel != want)
    return fail(PREMVOS_ERR_BAD_SHAPE, "premvos_refnet_set_param: '%s' has %lld elements, expected %lld", name, (long long)numel, (long long)want);
  n->params[name].assign(host_data, host_data + numel);
  return 0;
}

extern "C" int premvos_refnet_finalize(premvos_refnet_t* n) {
  PV_CHECK(n, PREMVOS_ERR_INVALID_ARG, "premvos_refnet_finalize: null handle");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_refnet_finalize: already finalized");
  for (auto& kv : n->shapes)
    if (!n->params.count(kv.first)) return fail(PREMVOS_ERR_NOT_READY, "premvos_refnet_finalize: missing variable '%s'", kv.first.c_str());
  PV_CUDA(cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking));
  PV_TRY(build_network(n));
  n->params.clear();
  const int64_t before = g_launch_count.load();
  PV_TRY(run_batch(n, n->NB, n->stream));  // warm-up on the zero-initialised input: validates every launch configuration
  PV_CUDA(cudaStreamSynchronize(n->stream));
  n->launches_per_forward = (int)(g_launch_count.load() - before) + 3;
  if (n->opt_cuda_graph) {
    const int64_t b2 = g_launch_count.load();
    PV_CUDA(cudaStreamBeginCapture(n->stream, cudaStreamCaptureModeThreadLocal));
    int r = run_batch(n, n->NB, n->stream);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(n->stream, &g);
    if (r != 0) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) return fail((int)e, "premvos_refnet_finalize: graph capture failed: %s", cudaGetErrorString(e));
    n->graph = g;
    n->graph_nodes = (int)(g_launch_count.load() - b2);
    g_launch_count.fetch_sub(n->graph_nodes);  // captured, not executed
    PV_CUDA(cudaGraphInstantiate(&n->exec, n->graph, 0));
  }
  n->finalized = true;
  return 0;
}

// One frame, n proposal boxes, everything on the device and on `st`; no synchronisation.  The batched equivalent of
// MergeTrack/refinement_net_functions.py:do_refinement's inner loop.
static int enqueue_refine(premvos_refnet* n, const unsigned char* frame_dev, int height, int width, const float* boxes_dev, int num_boxes,
                          unsigned char* masks_dev, float* conf_dev, float* post_dev, cudaStream_t st) {
  const size_t hw = (size_t)height * width;
  for (int b0 = 0; b0 < num_boxes; b0 += n->NB) {
    const int na = std::min(n->NB, num_boxes - b0);
    PV_TRY(refine_make_input(frame_dev, height, width, boxes_dev + (size_t)b0 * 4, na, n->S, n->input, n->crops, st));
    if (na == n->NB && n->exec && !profiling_enabled()) {
      PV_CUDA(cudaGraphLaunch(n->exec, st));
      count_launch(n->graph_nodes);
    } else {
      PV_TRY(run_batch(n, na, st));
    }
    PV_TRY(refine_output(n->logits, n->crops, na, n->S, height, width, masks_dev + (size_t)b0 * hw,
                         post_dev ? post_dev + (size_t)b0 * hw : nullptr, n->conf_sum, st));
    PV_TRY(refine_conf_finish(n->conf_sum, na, (long)hw, conf_dev + b0, st));
  }
  return 0;
}

extern "C" int premvos_refnet_forward(premvos_refnet_t* n, const unsigned char* frame_rgb_dev, int height, int width,
                                      const float* boxes_xywh_dev, int num_boxes, unsigned char* masks_dev, float* conf_scores_dev,
                                      float* posteriors_dev, void* stream) {
  PV_CHECK(n && frame_rgb_dev && (num_boxes == 0 || (boxes_xywh_dev && masks_dev && conf_scores_dev)), PREMVOS_ERR_INVALID_ARG,
           "premvos_refnet_forward: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_refnet_forward: call premvos_refnet_finalize first");
  PV_CHECK(height > 0 && width > 0 && num_boxes >= 0, PREMVOS_ERR_INVALID_ARG, "premvos_refnet_forward: bad sizes");
  return enqueue_refine(n, frame_rgb_dev, height, width, boxes_xywh_dev, num_boxes, masks_dev, conf_scores_dev, posteriors_dev,
                        (cudaStream_t)stream);
}

extern "C" int premvos_refnet_forward_host(premvos_refnet_t* n, const unsigned char* frame_rgb, int height, int width, const float* boxes_xywh,
                                           int num_boxes, unsigned char* masks_out, float* conf_scores_out, float* posteriors_out) {
  PV_CHECK(n && frame_rgb && (num_boxes == 0 || (boxes_xywh && masks_out && conf_scores_out)), PREMVOS_ERR_INVALID_ARG,
           "premvos_refnet_forward_host: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_refnet_forward_host: call premvos_refnet_finalize first");
  PV_CHECK(height > 0 && width > 0 && num_boxes >= 0, PREMVOS_ERR_INVALID_ARG, "premvos_refnet_forward_host: bad sizes");
  if (num_boxes == 0) return 0;
  cudaStream_t st = n->stream;
  const size_t hw = (size_t)height * width;
  // posteriors (4 B/pixel/proposal, a test / debugging output) are staged one launch group at a time; masks and
  // conf_scores of all proposals stay on the device until the end: one synchronisation per frame
  const int group = posteriors_out ? n->NB : num_boxes;
  PV_TRY(ensure((void**)&n->frame_dev, &n->frame_cap, hw * 3));
  PV_TRY(ensure((void**)&n->mask_dev, &n->mask_cap, hw * group));
  PV_TRY(ensure((void**)&n->hboxes_dev, &n->hboxes_cap, (size_t)num_boxes * 4 * sizeof(float)));
  PV_TRY(ensure((void**)&n->hconf_dev, &n->hconf_cap, (size_t)num_boxes * sizeof(float)));
  if (posteriors_out) PV_TRY(ensure((void**)&n->post_dev, &n->post_cap, hw * group * sizeof(float)));
  PV_CUDA(cudaMemcpyAsync(n->frame_dev, frame_rgb, hw * 3, cudaMemcpyHostToDevice, st));
  PV_CUDA(cudaMemcpyAsync(n->hboxes_dev, boxes_xywh, (size_t)num_boxes * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
  for (int b0 = 0; b0 < num_boxes; b0 += group) {
    const int na = std::min(group, num_boxes - b0);
    PV_TRY(enqueue_refine(n, n->frame_dev, height, width, (const float*)n->hboxes_dev + (size_t)b0 * 4, na, n->mask_dev,
                          (float*)n->hconf_dev + b0, posteriors_out ? n->post_dev : nullptr, st));
    PV_CUDA(cudaMemcpyAsync(masks_out + (size_t)b0 * hw, n->mask_dev, hw * na, cudaMemcpyDeviceToHost, st));
    if (posteriors_out) {
      PV_CUDA(cudaMemcpyAsync(posteriors_out + (size_t)b0 * hw, n->post_dev, hw * na * sizeof(float), cudaMemcpyDeviceToHost, st));
      PV_CUDA(cudaStreamSynchronize(st));
    }
  }
  PV_CUDA(cudaMemcpyAsync(conf_scores_out, n->hconf_dev, (size_t)num_boxes * sizeof(float), cudaMemcpyDeviceToHost, st));
  PV_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int premvos_refnet_launches_per_forward(const premvos_refnet_t* n) { return n ? n->launches_per_forward : 0; }

// Test hook (state of the LAST batch): "net_input" [NB,8,S,S], "xception_out", "low_level", "aspp_concat", "aspp_out",
// "decoder_in", "decoder_out" (CP8 -> NCHW), "logits" (fp32 [NB,h,w,16] channels-last), "crops" ([NB,4] as fp32).
extern "C" int premvos_refnet_get_tensor(premvos_refnet_t* n, const char* name, float* host_out, int64_t* numel) {
  PV_CHECK(n && name && numel, PREMVOS_ERR_INVALID_ARG, "premvos_refnet_get_tensor: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_refnet_get_tensor: network not finalized");
  PV_CUDA(cudaDeviceSynchronize());
  std::string k(name);
  if (k == "logits") {
    *numel = (int64_t)n->NB * n->logits.H * n->logits.W * 16;
    if (host_out) PV_CUDA(cudaMemcpy(host_out, n->logits.p, (size_t)(*numel) * 4, cudaMemcpyDeviceToHost));
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
  if (n->lazy.count(k)) {   // a tensor the production path never materialises: produce it now from the last batch's buffers
    PV_TRY(n->lazy[k](n->stream, n->NB));
    PV_CUDA(cudaStreamSynchronize(n->stream));
  }
  CView cv;
  if (k == "net_input" && n->stem_rows) {   // undo the row im2col: even x = tap 1 of column x/2, odd x = tap 0 of column (x+1)/2
    const CView& im = n->input;
    const int S = n->S, Wo = im.W;
    *numel = (int64_t)n->NB * 8 * S * S;
    if (!host_out) return 0;
    const size_t cols = (size_t)im.N * 16 * S * Wo;
    float* dtmp = nullptr;
    PV_CUDA(cudaMalloc((void**)&dtmp, cols * sizeof(float)));
    std::vector<float> t(cols);
    int r = cp8_to_nchw(im, 0, dtmp, nullptr);
    if (r == 0 && cudaMemcpy(t.data(), dtmp, cols * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
      r = fail(PREMVOS_ERR_INVALID_ARG, "premvos_refnet_get_tensor: copy failed");
    cudaFree(dtmp);
    if (r != 0) return r;
    for (int b = 0; b < n->NB; b++)
      for (int c = 0; c < 8; c++)
        for (int y = 0; y < S; y++)
          for (int xx = 0; xx < S; xx++) {
            float v = 0.f;
            if (c < 4) {
              if (!(xx & 1)) v = t[(((size_t)b * 16 + 4 + c) * S + y) * Wo + xx / 2];
              else if ((xx + 1) / 2 < Wo) v = t[(((size_t)b * 16 + c) * S + y) * Wo + (xx + 1) / 2];
              else v = t[(((size_t)b * 16 + 8 + c) * S + y) * Wo + (xx - 1) / 2];   // last column of an even-sized input: tap 2
            }
            host_out[(((size_t)b * 8 + c) * S + y) * S + xx] = v;
          }
    return 0;
  }
  if (k == "net_input") cv = n->input;
  else if (n->named.count(k)) cv = n->named[k];
  else return fail(PREMVOS_ERR_INVALID_ARG, "premvos_refnet_get_tensor: unknown tensor '%s'", name);
  *numel = (int64_t)cv.N * cv.C * cv.H * cv.W;
  if (!host_out) return 0;
  float* dtmp = nullptr;
  PV_CUDA(cudaMalloc((void**)&dtmp, (size_t)(*numel) * sizeof(float)));
  int r = cp8_to_nchw(cv, 0, dtmp, nullptr);
  if (r == 0) {
    cudaError_t e = cudaMemcpy(host_out, dtmp, (size_t)(*numel) * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) r = fail((int)e, "premvos_refnet_get_tensor: %s", cudaGetErrorString(e));
  }
  cudaFree(dtmp);
  return r;
}

extern "C" void premvos_refnet_destroy(premvos_refnet_t* n) {
  if (!n) return;
  cudaDeviceSynchronize();
  for (void* p : n->allocs) cudaFree(p);
  for (auto& w : n->conv_weights) free_conv_weights_umma(w.get());
  for (auto& pl : n->conv_plans) free_conv_plan_umma(pl.get());
  conv_workspace_free(&n->conv_ws);
  if (n->frame_dev) cudaFree(n->frame_dev);
  if (n->mask_dev) cudaFree(n->mask_dev);
  if (n->post_dev) cudaFree(n->post_dev);
  if (n->hboxes_dev) cudaFree(n->hboxes_dev);
  if (n->hconf_dev) cudaFree(n->hconf_dev);
  if (n->exec) cudaGraphExecDestroy(n->exec);
  if (n->graph) cudaGraphDestroy(n->graph);
  if (n->stream) cudaStreamDestroy(n->stream);
  delete n;
}
