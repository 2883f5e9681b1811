// PWC-DC-Net forward: host-side orchestration and C ABI.
//
// Restates PWCDCNet.forward (models/PWCNet.py:179-272) as a fixed sequence of kernels over
// pre-allocated channels-last buffers:
//   * both frames of every pair go through the siamese pyramid as one batch of 2B images;
//   * each decoder level owns one "slab" [B,H,W,Ctot] whose channel order is the reference's concat
//     order (new features first): [conv_4 32 | conv_3 64 | conv_2 96 | conv_1 128 | conv_0 128 |
//     corr 81 | c1 C_L | up_flow 2 | up_feat 2].  Convolutions read a channel suffix and write the
//     range in front of it, so torch.cat never materialises (the reference re-copies the growing
//     tensor five times per level);
//   * correlation writes (LeakyReLU fused) into the slab and copies c1 next to it; the deconvs of
//     level L+1 write up_flow / up_feat straight into the slab of level L; the warp reads up_flow
//     from there;
//   * the whole middle of the network is captured once into a CUDA graph (fixed shapes per handle).
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

using namespace premvos;

namespace {

const int LEVEL_CH[7] = {3, 16, 32, 64, 96, 128, 196};
const int DEC_OUT[5] = {128, 128, 96, 64, 32};
const int DEC_OUT_OFF[5] = {320, 192, 96, 32, 0};
const int DEC_IN_OFF[5] = {448, 320, 192, 96, 32};
const int BASE_OFF = 448;
const float WARP_SCALE[7] = {0, 0, 5.0f, 2.5f, 1.25f, 0.625f, 0};
const char* PYR_NAMES[7][3] = {{0, 0, 0},
                               {"conv1a", "conv1aa", "conv1b"},
                               {"conv2a", "conv2aa", "conv2b"},
                               {"conv3a", "conv3aa", "conv3b"},
                               {"conv4a", "conv4aa", "conv4b"},
                               {"conv5a", "conv5aa", "conv5b"},
                               {"conv6aa", "conv6a", "conv6b"}};
const int DC_DIL[6] = {1, 2, 4, 8, 16, 1};
const int DC_OUT[6] = {128, 128, 128, 96, 64, 32};

int level_od(int L) { return L == 6 ? 81 : 81 + LEVEL_CH[L] + 4; }

}  // namespace

struct premvos_pwc {
  int B = 0, H = 0, W = 0;
  bool finalized = false;
  int opt_tensor_cores = 1;
  int opt_cuda_graph = 1;
  std::map<std::string, std::vector<float>> params;
  std::map<std::string, std::vector<int64_t>> shapes;  // expected shapes

  std::vector<void*> allocs;
  TView img;
  TView pyr[7][3];
  TView slab[7], warpbuf[7], flow[7];
  TView ctxA, ctxB;

  ConvWeightsSimt w_pyr[7][3];
  ConvWeightsSimt w_dec[7][5];
  SmallConvWeights w_pf[7];
  DeconvWeights w_deconv[7], w_upfeat[7];
  ConvWeightsSimt w_dc[6];
  SmallConvWeights w_dc7;
  // tensor-core mode: decoder + context convolutions on tcgen05 (split-bf16 slabs)
  ConvWeightsUmma wt_dec[7][5], wt_dc[6];
  ConvPlanUmma plan_dec[7][5], plan_dc[6];

  cudaStream_t stream = nullptr;  // internal stream (forward_host, capture)
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int graph_nodes = 0;
  float* x_in = nullptr;      // device staging for forward_host
  float* flow_out = nullptr;  // device staging for forward_host
  int launches_per_forward = 0;
  int tensor_core_layers = 0;
};

namespace {

void build_shape_table(premvos_pwc* n) {
  auto conv = [&](const std::string& name, int cin, int cout, bool seq) {
    std::string k = name + (seq ? ".0" : "");
    n->shapes[k + ".weight"] = {cout, cin, 3, 3};
    n->shapes[k + ".bias"] = {cout};
  };
  auto deconv = [&](const std::string& name, int cin) {
    n->shapes[name + ".weight"] = {cin, 2, 4, 4};
    n->shapes[name + ".bias"] = {2};
  };
  for (int L = 1; L <= 6; L++) {
    conv(PYR_NAMES[L][0], LEVEL_CH[L - 1], LEVEL_CH[L], true);
    conv(PYR_NAMES[L][1], LEVEL_CH[L], LEVEL_CH[L], true);
    conv(PYR_NAMES[L][2], LEVEL_CH[L], LEVEL_CH[L], true);
  }
  for (int L = 6; L >= 2; L--) {
    int od = level_od(L);
    int cin = od;
    for (int i = 0; i < 5; i++) {
      conv("conv" + std::to_string(L) + "_" + std::to_string(i), cin, DEC_OUT[i], true);
      cin += DEC_OUT[i];
    }
    conv("predict_flow" + std::to_string(L), cin, 2, false);
    deconv("deconv" + std::to_string(L), 2);
    if (L != 2) deconv("upfeat" + std::to_string(L), cin);
  }
  int cin = level_od(2) + 448;
  for (int i = 0; i < 6; i++) {
    conv("dc_conv" + std::to_string(i + 1), cin, DC_OUT[i], true);
    cin = DC_OUT[i];
  }
  conv("dc_conv7", 32, 2, false);
}

int64_t numel_of(const std::vector<int64_t>& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}

int alloc_view(premvos_pwc* n, TView* v, int N, int H, int W, int C, int cs, bool split = false) {
  v->N = N; v->H = H; v->W = W; v->C = C; v->cs = cs; v->coff = 0;
  size_t elems = (size_t)N * H * W * cs;
  for (int k = 0; k < (split ? 2 : 1); k++) {
    size_t bytes = elems * (split ? sizeof(__nv_bfloat16) : sizeof(float));
    void* p = nullptr;
    PV_CUDA(cudaMalloc(&p, bytes));
    PV_CUDA(cudaMemset(p, 0, bytes));
    n->allocs.push_back(p);
    if (!split) v->p = (float*)p;
    else if (k == 0) v->hi = (__nv_bfloat16*)p;
    else v->lo = (__nv_bfloat16*)p;
  }
  return 0;
}

const float* P(premvos_pwc* n, const std::string& k) { return n->params[k].data(); }

int pack_all_weights(premvos_pwc* n) {
  for (int L = 1; L <= 6; L++)
    for (int j = 0; j < 3; j++) {
      std::string k = std::string(PYR_NAMES[L][j]) + ".0";
      int cin = (j == 0) ? LEVEL_CH[L - 1] : LEVEL_CH[L];
      PV_TRY(pack_conv_weights_simt(&n->w_pyr[L][j], P(n, k + ".weight"), P(n, k + ".bias"), LEVEL_CH[L], cin, 3, 3));
    }
  for (int L = 6; L >= 2; L--) {
    int cin = level_od(L);
    for (int i = 0; i < 5; i++) {
      std::string k = "conv" + std::to_string(L) + "_" + std::to_string(i) + ".0";
      if (n->opt_tensor_cores)
        PV_TRY(pack_conv_weights_umma(&n->wt_dec[L][i], P(n, k + ".weight"), P(n, k + ".bias"), DEC_OUT[i], cin, 3, 3));
      else
        PV_TRY(pack_conv_weights_simt(&n->w_dec[L][i], P(n, k + ".weight"), P(n, k + ".bias"), DEC_OUT[i], cin, 3, 3));
      cin += DEC_OUT[i];
    }
    std::string pf = "predict_flow" + std::to_string(L);
    PV_TRY(pack_small_conv_weights(&n->w_pf[L], P(n, pf + ".weight"), P(n, pf + ".bias"), 2, cin));
    std::string dk = "deconv" + std::to_string(L);
    PV_TRY(pack_deconv_weights(&n->w_deconv[L], P(n, dk + ".weight"), P(n, dk + ".bias"), 2));
    if (L != 2) {
      std::string uk = "upfeat" + std::to_string(L);
      PV_TRY(pack_deconv_weights(&n->w_upfeat[L], P(n, uk + ".weight"), P(n, uk + ".bias"), cin));
    }
  }
  int cin = level_od(2) + 448;
  for (int i = 0; i < 6; i++) {
    std::string k = "dc_conv" + std::to_string(i + 1) + ".0";
    if (n->opt_tensor_cores)
      PV_TRY(pack_conv_weights_umma(&n->wt_dc[i], P(n, k + ".weight"), P(n, k + ".bias"), DC_OUT[i], cin, 3, 3));
    else
      PV_TRY(pack_conv_weights_simt(&n->w_dc[i], P(n, k + ".weight"), P(n, k + ".bias"), DC_OUT[i], cin, 3, 3));
    cin = DC_OUT[i];
  }
  PV_TRY(pack_small_conv_weights(&n->w_dc7, P(n, "dc_conv7.weight"), P(n, "dc_conv7.bias"), 2, 32));
  return 0;
}

int alloc_activations(premvos_pwc* n) {
  const int B = n->B;
  PV_TRY(alloc_view(n, &n->img, 2 * B, n->H, n->W, 3, 4));
  for (int L = 1; L <= 6; L++) {
    int h = n->H >> L, w = n->W >> L, c = LEVEL_CH[L], cs = round_up(c, 4);
    for (int j = 0; j < 3; j++) PV_TRY(alloc_view(n, &n->pyr[L][j], 2 * B, h, w, c, cs));
  }
  for (int L = 6; L >= 2; L--) {
    int h = n->H >> L, w = n->W >> L;
    int ctot = level_od(L) + 448;
    PV_TRY(alloc_view(n, &n->slab[L], B, h, w, ctot, round_up(ctot, 8), n->opt_tensor_cores != 0));
    PV_TRY(alloc_view(n, &n->flow[L], B, h, w, 2, 4));
    if (L != 6) PV_TRY(alloc_view(n, &n->warpbuf[L], B, h, w, LEVEL_CH[L], round_up(LEVEL_CH[L], 4)));
  }
  PV_TRY(alloc_view(n, &n->ctxA, B, n->H >> 2, n->W >> 2, 128, 128, n->opt_tensor_cores != 0));
  PV_TRY(alloc_view(n, &n->ctxB, B, n->H >> 2, n->W >> 2, 128, 128, n->opt_tensor_cores != 0));
  if (n->opt_tensor_cores) {  // one launch plan (tensor maps + arguments) per tensor-core layer
    for (int L = 6; L >= 2; L--) {
      const int ctot = level_od(L) + 448;
      for (int i = 0; i < 5; i++) {
        TView in = n->slab[L].slice(DEC_IN_OFF[i], ctot - DEC_IN_OFF[i]);
        TView out = n->slab[L].slice(DEC_OUT_OFF[i], DEC_OUT[i]);
        PV_TRY(plan_conv_umma(&n->plan_dec[L][i], in, out, n->wt_dec[L][i], 1, 0.1f));
        n->tensor_core_layers++;
      }
    }
    TView in = n->slab[2].slice(0, level_od(2) + 448);
    TView bufs[2] = {n->ctxA, n->ctxB};
    for (int i = 0; i < 6; i++) {
      TView out = bufs[i & 1].slice(0, DC_OUT[i]);
      PV_TRY(plan_conv_umma(&n->plan_dc[i], in, out, n->wt_dc[i], DC_DIL[i], 0.1f));
      n->tensor_core_layers++;
      in = out;
    }
  }
  size_t xin = (size_t)B * 6 * n->H * n->W * sizeof(float);
  size_t fout = (size_t)B * 2 * (n->H / 4) * (n->W / 4) * sizeof(float);
  PV_CUDA(cudaMalloc((void**)&n->x_in, xin));
  PV_CUDA(cudaMalloc((void**)&n->flow_out, fout));
  n->allocs.push_back(n->x_in);
  n->allocs.push_back(n->flow_out);
  return 0;
}

// ---- the network ---------------------------------------------------------------------------------
int run_front(premvos_pwc* n, const float* x_dev, cudaStream_t st) {
  return pack_pair_input(x_dev, n->B, n->H, n->W, n->img, st);
}

int run_middle(premvos_pwc* n, cudaStream_t st) {
  const int B = n->B;
  // feature pyramid, both frames at once (PWCNet.py:183-194)
  TView cur = n->img;
  for (int L = 1; L <= 6; L++) {
    PV_TRY(conv2d_simt(cur, n->pyr[L][0], n->w_pyr[L][0], 2, 1, 0.1f, st));
    PV_TRY(conv2d_simt(n->pyr[L][0], n->pyr[L][1], n->w_pyr[L][1], 1, 1, 0.1f, st));
    PV_TRY(conv2d_simt(n->pyr[L][1], n->pyr[L][2], n->w_pyr[L][2], 1, 1, 0.1f, st));
    cur = n->pyr[L][2];
  }
  for (int L = 6; L >= 2; L--) {
    const int CL = LEVEL_CH[L];
    const int od = level_od(L);
    const int ctot = od + 448;
    TView c1 = n->pyr[L][2].batch_range(0, B);
    TView c2 = n->pyr[L][2].batch_range(B, B);
    TView& slab = n->slab[L];
    TView f2 = c2;
    TView c1_slot;  // stays null at level 6 (x = corr6 only, PWCNet.py:201)
    if (L != 6) {
      TView upflow = slab.slice(BASE_OFF + 81 + CL, 2);
      PV_TRY(warp_nhwc(c2, upflow, WARP_SCALE[L], n->warpbuf[L], st));  // :211,225,239,255
      f2 = n->warpbuf[L];
      c1_slot = slab.slice(BASE_OFF + 81, CL);
    }
    PV_TRY(corr81_nhwc(c1, f2, slab.slice(BASE_OFF, 81), c1_slot, 0.1f, st));  // corr + leakyRELU
    for (int i = 0; i < 5; i++) {
      TView in = slab.slice(DEC_IN_OFF[i], ctot - DEC_IN_OFF[i]);
      TView out = slab.slice(DEC_OUT_OFF[i], DEC_OUT[i]);
      if (n->opt_tensor_cores) PV_TRY(launch_conv_umma(n->plan_dec[L][i], st));
      else PV_TRY(conv2d_simt(in, out, n->w_dec[L][i], 1, 1, 0.1f, st));
    }
    TView all = slab.slice(0, ctot);
    TView flow = n->flow[L].slice(0, 2);
    PV_TRY(conv3x3_small_cout(all, flow, n->w_pf[L], nullptr, nullptr, st));
    if (L != 2) {
      TView& nslab = n->slab[L - 1];
      const int CN = LEVEL_CH[L - 1];
      PV_TRY(deconv4x4s2_cout2(flow, nslab.slice(BASE_OFF + 81 + CN, 2), n->w_deconv[L], st));
      PV_TRY(deconv4x4s2_cout2(all, nslab.slice(BASE_OFF + 81 + CN + 2, 2), n->w_upfeat[L], st));
    }
  }
  // context network (PWCNet.py:266-267)
  TView in = n->slab[2].slice(0, level_od(2) + 448);
  TView bufs[2] = {n->ctxA, n->ctxB};
  for (int i = 0; i < 6; i++) {
    TView out = bufs[i & 1].slice(0, DC_OUT[i]);
    if (n->opt_tensor_cores) PV_TRY(launch_conv_umma(n->plan_dc[i], st));
    else PV_TRY(conv2d_simt(in, out, n->w_dc[i], 1, DC_DIL[i], 0.1f, st));
    in = out;
  }
  return 0;
}

int run_back(premvos_pwc* n, float* flow_dev, cudaStream_t st) {
  TView dc6 = n->ctxB.slice(0, 32);
  TView flow2 = n->flow[2].slice(0, 2);
  TView none;
  return conv3x3_small_cout(dc6, none, n->w_dc7, &flow2, flow_dev, st);  // flow2 += dc_conv7(...), NCHW out
}

int enqueue_forward(premvos_pwc* n, const float* x_dev, float* flow_dev, cudaStream_t st) {
  PV_TRY(run_front(n, x_dev, st));
  if (n->exec && !profiling_enabled()) {
    PV_CUDA(cudaGraphLaunch(n->exec, st));
    count_launch(n->graph_nodes);
  } else {
    PV_TRY(run_middle(n, st));
  }
  PV_TRY(run_back(n, flow_dev, st));
  return 0;
}

}  // namespace

extern "C" int premvos_pwc_create(premvos_pwc_t** out, int batch, int height, int width) {
  PV_CHECK(out, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_create: out is null");
  *out = nullptr;
  PV_CHECK(batch >= 1 && height >= 64 && width >= 64 && (height % 64) == 0 && (width % 64) == 0, PREMVOS_ERR_INVALID_ARG,
           "premvos_pwc_create: batch >= 1 and height/width positive multiples of 64 required (got %d,%d,%d)", batch,
           height, width);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PREMVOS_ERR_NO_DEVICE, "premvos_pwc_create: no CUDA device visible");
  premvos_pwc* n = new premvos_pwc();
  n->B = batch; n->H = height; n->W = width;
  build_shape_table(n);
  *out = n;
  return 0;
}

extern "C" int premvos_pwc_set_option(premvos_pwc_t* n, const char* key, int value) {
  PV_CHECK(n && key, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_set_option: null argument");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_set_option: options must be set before finalize");
  std::string k(key);
  if (k == "tensor_cores") n->opt_tensor_cores = value;
  else if (k == "cuda_graph") n->opt_cuda_graph = value;
  else return fail(PREMVOS_ERR_INVALID_ARG, "premvos_pwc_set_option: unknown option '%s'", key);
  return 0;
}

extern "C" int premvos_pwc_set_param(premvos_pwc_t* n, const char* name, const float* host_data, int64_t numel) {
  PV_CHECK(n && name && host_data, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_set_param: null argument");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_set_param: network already finalized");
  auto it = n->shapes.find(name);
  if (it == n->shapes.end()) return fail(PREMVOS_ERR_UNKNOWN_PARAM, "premvos_pwc_set_param: unexpected key '%s'", name);
  int64_t want = numel_of(it->second);
  if (numel != want)
    return fail(PREMVOS_ERR_BAD_SHAPE, "premvos_pwc_set_param: '%s' has %lld elements, expected %lld", name,
                (long long)numel, (long long)want);
  n->params[name].assign(host_data, host_data + numel);
  return 0;
}

extern "C" int premvos_pwc_finalize(premvos_pwc_t* n) {
  PV_CHECK(n, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_finalize: null handle");
  PV_CHECK(!n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_finalize: already finalized");
  for (auto& kv : n->shapes)
    if (!n->params.count(kv.first))
      return fail(PREMVOS_ERR_NOT_READY, "premvos_pwc_finalize: missing key '%s' in state_dict", kv.first.c_str());
  PV_CUDA(cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking));
  PV_TRY(pack_all_weights(n));
  PV_TRY(alloc_activations(n));
  n->params.clear();
  // warm-up run outside capture: sets function attributes and validates every launch configuration
  PV_CUDA(cudaMemsetAsync(n->x_in, 0, (size_t)n->B * 6 * n->H * n->W * sizeof(float), n->stream));
  int64_t before = g_launch_count.load();
  PV_TRY(run_front(n, n->x_in, n->stream));
  PV_TRY(run_middle(n, n->stream));
  PV_TRY(run_back(n, n->flow_out, n->stream));
  PV_CUDA(cudaStreamSynchronize(n->stream));
  n->launches_per_forward = (int)(g_launch_count.load() - before);
  if (n->opt_cuda_graph) {
    int64_t b2 = g_launch_count.load();
    PV_CUDA(cudaStreamBeginCapture(n->stream, cudaStreamCaptureModeThreadLocal));
    int r = run_middle(n, n->stream);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(n->stream, &g);
    if (r != 0) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) return fail((int)e, "premvos_pwc_finalize: graph capture failed: %s", cudaGetErrorString(e));
    n->graph = g;
    n->graph_nodes = (int)(g_launch_count.load() - b2);
    g_launch_count.fetch_sub(n->graph_nodes);  // captured, not executed
    PV_CUDA(cudaGraphInstantiate(&n->exec, n->graph, 0));
  }
  n->finalized = true;
  return 0;
}

extern "C" int premvos_pwc_forward(premvos_pwc_t* n, const float* x_dev, float* flow_dev, void* stream) {
  PV_CHECK(n && x_dev && flow_dev, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_forward: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_forward: call premvos_pwc_finalize first");
  return enqueue_forward(n, x_dev, flow_dev, (cudaStream_t)stream);
}

extern "C" int premvos_pwc_forward_host(premvos_pwc_t* n, const float* x_host, float* flow_host) {
  PV_CHECK(n && x_host && flow_host, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_forward_host: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_forward_host: call premvos_pwc_finalize first");
  size_t xin = (size_t)n->B * 6 * n->H * n->W * sizeof(float);
  size_t fout = (size_t)n->B * 2 * (n->H / 4) * (n->W / 4) * sizeof(float);
  PV_CUDA(cudaMemcpyAsync(n->x_in, x_host, xin, cudaMemcpyHostToDevice, n->stream));
  PV_TRY(enqueue_forward(n, n->x_in, n->flow_out, n->stream));
  PV_CUDA(cudaMemcpyAsync(flow_host, n->flow_out, fout, cudaMemcpyDeviceToHost, n->stream));
  PV_CUDA(cudaStreamSynchronize(n->stream));
  return 0;
}

extern "C" int premvos_pwc_launches_per_forward(const premvos_pwc_t* n) { return n ? n->launches_per_forward : 0; }

extern "C" int premvos_pwc_tensor_core_layers(const premvos_pwc_t* n) { return n ? n->tensor_core_layers : 0; }

extern "C" int premvos_pwc_get_tensor(premvos_pwc_t* n, const char* name, float* host_out, int64_t* numel) {
  PV_CHECK(n && name && numel, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_get_tensor: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_get_tensor: network not finalized");
  std::string k(name);
  TView v;
  bool found = false;
  auto lvl = [&](size_t pos) { return (k.size() > pos && k[pos] >= '1' && k[pos] <= '6') ? k[pos] - '0' : -1; };
  if (k.size() == 3 && k[0] == 'c' && (k[1] == '1' || k[1] == '2') && lvl(2) > 0) {
    int L = lvl(2);
    v = n->pyr[L][2].batch_range(k[1] == '1' ? 0 : n->B, n->B);
    found = true;
  } else if (k.rfind("corr", 0) == 0 && lvl(4) >= 2) {
    v = n->slab[lvl(4)].slice(BASE_OFF, 81); found = true;
  } else if (k.rfind("warp", 0) == 0 && lvl(4) >= 2 && lvl(4) <= 5) {
    v = n->warpbuf[lvl(4)]; found = true;
  } else if (k.rfind("flow", 0) == 0 && lvl(4) >= 2) {
    v = n->flow[lvl(4)].slice(0, 2); found = true;
  } else if (k.rfind("up_flow", 0) == 0 && lvl(7) >= 3) {
    int L = lvl(7);
    v = n->slab[L - 1].slice(BASE_OFF + 81 + LEVEL_CH[L - 1], 2); found = true;
  } else if (k.rfind("up_feat", 0) == 0 && lvl(7) >= 3) {
    int L = lvl(7);
    v = n->slab[L - 1].slice(BASE_OFF + 81 + LEVEL_CH[L - 1] + 2, 2); found = true;
  } else if (k.rfind("slab", 0) == 0 && lvl(4) >= 2) {
    int L = lvl(4);
    v = n->slab[L].slice(0, level_od(L) + 448); found = true;
  } else if (k == "dc6") {
    v = n->ctxB.slice(0, 32); found = true;
  }
  if (!found) return fail(PREMVOS_ERR_INVALID_ARG, "premvos_pwc_get_tensor: unknown tensor '%s'", name);
  *numel = (int64_t)v.N * v.C * v.H * v.W;
  if (!host_out) return 0;
  PV_CUDA(cudaDeviceSynchronize());
  float* dtmp = nullptr;
  PV_CUDA(cudaMalloc((void**)&dtmp, (size_t)(*numel) * sizeof(float)));
  int r = view_to_nchw(v, dtmp, nullptr);
  if (r == 0) {
    cudaError_t e = cudaMemcpy(host_out, dtmp, (size_t)(*numel) * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) r = fail((int)e, "premvos_pwc_get_tensor: %s", cudaGetErrorString(e));
  }
  cudaFree(dtmp);
  return r;
}

extern "C" void premvos_pwc_destroy(premvos_pwc_t* n) {
  if (!n) return;
  cudaDeviceSynchronize();
  if (n->exec) cudaGraphExecDestroy(n->exec);
  if (n->graph) cudaGraphDestroy(n->graph);
  for (void* p : n->allocs) cudaFree(p);
  for (int L = 1; L <= 6; L++)
    for (int j = 0; j < 3; j++) free_conv_weights_simt(&n->w_pyr[L][j]);
  for (int L = 2; L <= 6; L++) {
    for (int i = 0; i < 5; i++) free_conv_weights_simt(&n->w_dec[L][i]);
    free_small_conv_weights(&n->w_pf[L]);
    free_deconv_weights(&n->w_deconv[L]);
    free_deconv_weights(&n->w_upfeat[L]);
  }
  for (int i = 0; i < 6; i++) free_conv_weights_simt(&n->w_dc[i]);
  for (int L = 2; L <= 6; L++)
    for (int i = 0; i < 5; i++) free_conv_weights_umma(&n->wt_dec[L][i]);
  for (int i = 0; i < 6; i++) free_conv_weights_umma(&n->wt_dc[i]);
  free_small_conv_weights(&n->w_dc7);
  if (n->stream) cudaStreamDestroy(n->stream);
  delete n;
}
