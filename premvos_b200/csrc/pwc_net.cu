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
  // tensor-core mode: every convolution on tcgen05, activations in CP8 split-bf16 planes
  CView c_img, c_pyr[7][3], c_slab[7], c_warp[7], c_ctxA, c_ctxB;
  TView head[7], dc7out;   // fp32 channels-last [B,h,w,16]: {flow 2, upfeat phases 8} / {dc_conv7 2}
  ConvWeightsUmma wt_pyr[7][3], wt_dec[7][5], wt_head[7], wt_dc[6], wt_dc7;
  ConvPlanUmma pl_pyr[7][3], pl_dec[7][5], pl_head[7], pl_dc[6], pl_dc7;
  float* d_deconv_w[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // deconvL weights [2][2][4][4] (64 floats) + bias [2]

  cudaStream_t stream = nullptr;  // internal stream (forward_host, capture)
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int graph_nodes = 0;
  float* x_in = nullptr;      // device staging for forward_host
  float* flow_out = nullptr;  // device staging for forward_host
  unsigned char* frames_u8 = nullptr;  // device staging for forward_host_u8
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

int pack_weights_tc(premvos_pwc* n);
int alloc_activations_tc(premvos_pwc* n);

int pack_all_weights(premvos_pwc* n) {
  if (n->opt_tensor_cores) return pack_weights_tc(n);
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
    PV_TRY(pack_conv_weights_simt(&n->w_dc[i], P(n, k + ".weight"), P(n, k + ".bias"), DC_OUT[i], cin, 3, 3));
    cin = DC_OUT[i];
  }
  PV_TRY(pack_small_conv_weights(&n->w_dc7, P(n, "dc_conv7.weight"), P(n, "dc_conv7.bias"), 2, 32));
  return 0;
}

int alloc_io(premvos_pwc* n);

int alloc_activations(premvos_pwc* n) {
  if (n->opt_tensor_cores) return alloc_activations_tc(n);
  const int B = n->B;
  PV_TRY(alloc_view(n, &n->img, 2 * B, n->H, n->W, 3, 4));
  for (int L = 1; L <= 6; L++) {
    int h = n->H >> L, w = n->W >> L, c = LEVEL_CH[L], cs = round_up(c, 4);
    for (int j = 0; j < 3; j++) PV_TRY(alloc_view(n, &n->pyr[L][j], 2 * B, h, w, c, cs));
  }
  for (int L = 6; L >= 2; L--) {
    int h = n->H >> L, w = n->W >> L;
    int ctot = level_od(L) + 448;
    PV_TRY(alloc_view(n, &n->slab[L], B, h, w, ctot, round_up(ctot, 8)));
    PV_TRY(alloc_view(n, &n->flow[L], B, h, w, 2, 4));
    if (L != 6) PV_TRY(alloc_view(n, &n->warpbuf[L], B, h, w, LEVEL_CH[L], round_up(LEVEL_CH[L], 4)));
  }
  PV_TRY(alloc_view(n, &n->ctxA, B, n->H >> 2, n->W >> 2, 128, 128));
  PV_TRY(alloc_view(n, &n->ctxB, B, n->H >> 2, n->W >> 2, 128, 128));
  return alloc_io(n);
}

// ---- tensor-core mode: slab layout (chunk planes) ---------------------------------------------------
//   [conv_4 4 | conv_3 8 | conv_2 12 | conv_1 16 | conv_0 16 | corr 11 (81 ch + 7 zeros) | c1 C_L/8 |
//    up 1 (up_flow 2, up_feat 2, 4 zeros)]        -- every segment starts on a chunk boundary; the weight
// packer maps the reference's input-channel order (PWCNet.py:201-205) onto these physical channels.
const int CORR_CHUNK = 56, C1_CHUNK = 67;
int up_chunk(int L) { return C1_CHUNK + LEVEL_CH[L] / 8; }
int slab_chunks(int L) { return L == 6 ? C1_CHUNK : up_chunk(L) + 1; }

// physical channel (relative to a view that starts at dense channel `in_off`) of every reference input channel
std::vector<int> slab_cin_map(int L, int in_off) {
  std::vector<int> m;
  for (int r = in_off; r < 448; r++) m.push_back(r - in_off);
  const int base = 448 - in_off;
  for (int r = 0; r < 81; r++) m.push_back(base + r);
  if (L != 6) {
    for (int r = 0; r < LEVEL_CH[L]; r++) m.push_back(base + 88 + r);
    for (int r = 0; r < 4; r++) m.push_back(base + 88 + LEVEL_CH[L] + r);
  }
  return m;
}

// true when a decoder layer of level L launches fewer 128-pixel tiles than the GPU has SMs
bool small_grid(const premvos_pwc* n, int L) {
  const long tiles = (long)n->B * (((n->H >> L) + 15) / 16) * (((n->W >> L) + 7) / 8);
  return tiles < 148;
}

int pack_weights_tc(premvos_pwc* n) {
  for (int L = 1; L <= 6; L++)
    for (int j = 0; j < 3; j++) {
      std::string k = std::string(PYR_NAMES[L][j]) + ".0";
      int cin = (j == 0) ? LEVEL_CH[L - 1] : LEVEL_CH[L];
      if (L == 1 && j == 0) {
        // conv1a runs as a 1x1 GEMM over the im2col image written by pack_pair_im2col_cp8: K index (r*3+s)*3 + c
        const float* w3 = P(n, k + ".weight");  // [16][3][3][3]
        std::vector<float> w1((size_t)LEVEL_CH[1] * 27);
        for (int co = 0; co < LEVEL_CH[1]; co++)
          for (int c = 0; c < 3; c++)
            for (int t = 0; t < 9; t++) w1[(size_t)co * 27 + t * 3 + c] = w3[((size_t)co * 3 + c) * 9 + t];
        PV_TRY(pack_conv_weights_umma(&n->wt_pyr[L][j], w1.data(), P(n, k + ".bias"), LEVEL_CH[L], 27, 1, 1));
      } else {
        PV_TRY(pack_conv_weights_umma(&n->wt_pyr[L][j], P(n, k + ".weight"), P(n, k + ".bias"), LEVEL_CH[L], cin, 3, 3));
      }
    }
  for (int L = 6; L >= 2; L--) {
    const int phys_total = slab_chunks(L) * 8;
    int cin = level_od(L);
    const int cin_all = cin + 448;               // input channels of predict_flowL / upfeatL (the whole slab)
    // head = predict_flowL (2 ch) fused with upfeatL: ConvTranspose2d(cin, 2, 4, 2, 1) == 3x3 convolution with
    // 8 outputs ((py*2+px)*2+co) at the input resolution followed by a pixel shuffle:
    //   out[2y+py, 2x+px, co] = sum_{r,s} in[y+r-1, x+s-1] * Wd[ci][co][3-2r+py][3-2s+px]   (taps outside 0..3 vanish)
    const int hc = (L != 2) ? 10 : 2;
    std::vector<float> hw((size_t)hc * cin_all * 9, 0.f), hb(hc, 0.f);
    {
      std::string pf = "predict_flow" + std::to_string(L);
      const float* pw = P(n, pf + ".weight");
      std::copy(pw, pw + (size_t)2 * cin_all * 9, hw.begin());
      hb[0] = P(n, pf + ".bias")[0]; hb[1] = P(n, pf + ".bias")[1];
    }
    if (L != 2) {
      std::string uk = "upfeat" + std::to_string(L);
      const float* uw = P(n, uk + ".weight");  // [cin_all][2][4][4]
      const float* ub = P(n, uk + ".bias");
      for (int py = 0; py < 2; py++)
        for (int px = 0; px < 2; px++)
          for (int co = 0; co < 2; co++) {
            const int oc = 2 + (py * 2 + px) * 2 + co;
            hb[oc] = ub[co];
            for (int ci = 0; ci < cin_all; ci++)
              for (int r = 0; r < 3; r++)
                for (int s2 = 0; s2 < 3; s2++) {
                  const int ky = 3 - 2 * r + py, kx = 3 - 2 * s2 + px;
                  if (ky < 0 || ky > 3 || kx < 0 || kx > 3) continue;
                  hw[((size_t)oc * cin_all + ci) * 9 + r * 3 + s2] = uw[(((size_t)ci * 2 + co) * 4 + ky) * 4 + kx];
                }
          }
      std::string dk = "deconv" + std::to_string(L);
      std::vector<float> dw(66);  // [ci 2][co 2][4][4] + bias [2]
      std::copy(P(n, dk + ".weight"), P(n, dk + ".weight") + 64, dw.begin());
      dw[64] = P(n, dk + ".bias")[0]; dw[65] = P(n, dk + ".bias")[1];
      PV_CUDA(cudaMalloc((void**)&n->d_deconv_w[L], 66 * sizeof(float)));
      PV_CUDA(cudaMemcpy(n->d_deconv_w[L], dw.data(), 66 * sizeof(float), cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < 5; i++) {
      std::string k = "conv" + std::to_string(L) + "_" + std::to_string(i) + ".0";
      std::vector<int> map = slab_cin_map(L, DEC_IN_OFF[i]);
      PV_CHECK((int)map.size() == cin, PREMVOS_ERR_INVALID_ARG, "internal: slab map size %d != %d", (int)map.size(), cin);
      // levels 6..4 launch fewer CTAs than there are SMs: split-K with 32-channel k-blocks
      const int kc = small_grid(n, L) ? 4 : 0;
      if (i < 4) {
        PV_TRY(pack_conv_weights_umma(&n->wt_dec[L][i], P(n, k + ".weight"), P(n, k + ".bias"), DEC_OUT[i], cin, 3, 3, map.data(),
                                      phys_total - DEC_IN_OFF[i], kc));
      } else {
        // convL_4 reads everything the flow head reads except its own 32 outputs: the head's partial sums over those
        // `cin` channels ride along as hc extra (linear) output channels; the head proper only adds the 32-channel rest
        const int co = DEC_OUT[4] + hc;
        std::vector<float> w4((size_t)co * cin * 9), b4(co);
        std::copy(P(n, k + ".weight"), P(n, k + ".weight") + (size_t)DEC_OUT[4] * cin * 9, w4.begin());
        std::copy(P(n, k + ".bias"), P(n, k + ".bias") + DEC_OUT[4], b4.begin());
        for (int o = 0; o < hc; o++) {
          b4[DEC_OUT[4] + o] = hb[o];
          for (int ci = 0; ci < cin; ci++)
            for (int t = 0; t < 9; t++)
              w4[((size_t)(DEC_OUT[4] + o) * cin + ci) * 9 + t] = hw[((size_t)o * cin_all + DEC_OUT[4] + ci) * 9 + t];
        }
        PV_TRY(pack_conv_weights_umma(&n->wt_dec[L][i], w4.data(), b4.data(), co, cin, 3, 3, map.data(), phys_total - DEC_IN_OFF[i], kc));
        std::vector<float> wr((size_t)hc * DEC_OUT[4] * 9);
        for (int o = 0; o < hc; o++)
          for (int ci = 0; ci < DEC_OUT[4]; ci++)
            for (int t = 0; t < 9; t++) wr[((size_t)o * DEC_OUT[4] + ci) * 9 + t] = hw[((size_t)o * cin_all + ci) * 9 + t];
        PV_TRY(pack_conv_weights_umma(&n->wt_head[L], wr.data(), nullptr, hc, DEC_OUT[4], 3, 3));
      }
      cin += DEC_OUT[i];
    }
  }
  int cin = level_od(2) + 448;
  for (int i = 0; i < 6; i++) {
    std::string k = "dc_conv" + std::to_string(i + 1) + ".0";
    if (i == 0) {
      std::vector<int> map = slab_cin_map(2, 0);
      PV_TRY(pack_conv_weights_umma(&n->wt_dc[i], P(n, k + ".weight"), P(n, k + ".bias"), DC_OUT[i], cin, 3, 3, map.data(),
                                    slab_chunks(2) * 8));
    } else {
      PV_TRY(pack_conv_weights_umma(&n->wt_dc[i], P(n, k + ".weight"), P(n, k + ".bias"), DC_OUT[i], cin, 3, 3));
    }
    cin = DC_OUT[i];
  }
  PV_TRY(pack_conv_weights_umma(&n->wt_dc7, P(n, "dc_conv7.weight"), P(n, "dc_conv7.bias"), 2, 32, 3, 3));
  return 0;
}

int alloc_cview(premvos_pwc* n, CView* v, int N, int H, int W, int chunks) {
  v->N = N; v->H = H; v->W = W; v->chunks = chunks; v->c0 = 0; v->C = chunks * 8;
  // + 128 B: a flattened 1x1 layer may read the last 8-pixel row of the last plane past its end (conv_umma.cu)
  const size_t bytes = (size_t)N * chunks * H * W * 8 * sizeof(__nv_bfloat16) + 128;
  for (int k = 0; k < 2; k++) {
    void* p = nullptr;
    PV_CUDA(cudaMalloc(&p, bytes));
    PV_CUDA(cudaMemset(p, 0, bytes));
    n->allocs.push_back(p);
    (k == 0 ? v->hi : v->lo) = (__nv_bfloat16*)p;
  }
  return 0;
}

int alloc_io(premvos_pwc* n) {
  size_t xin = (size_t)n->B * 6 * n->H * n->W * sizeof(float);
  size_t fout = (size_t)n->B * 2 * (n->H / 4) * (n->W / 4) * sizeof(float);
  PV_CUDA(cudaMalloc((void**)&n->x_in, xin));
  PV_CUDA(cudaMalloc((void**)&n->flow_out, fout));
  n->allocs.push_back(n->x_in);
  PV_CUDA(cudaMalloc((void**)&n->frames_u8, (size_t)n->B * 2 * n->H * n->W * 3));
  n->allocs.push_back(n->frames_u8);
  n->allocs.push_back(n->flow_out);
  return 0;
}

int alloc_activations_tc(premvos_pwc* n) {
  const int B = n->B;
  PV_TRY(alloc_cview(n, &n->c_img, 2 * B, n->H / 2, n->W / 2, 4));
  n->c_img.C = 32;
  for (int L = 1; L <= 6; L++)
    for (int j = 0; j < 3; j++) {
      PV_TRY(alloc_cview(n, &n->c_pyr[L][j], 2 * B, n->H >> L, n->W >> L, (LEVEL_CH[L] + 7) / 8));
      n->c_pyr[L][j].C = LEVEL_CH[L];
    }
  for (int L = 6; L >= 2; L--) {
    const int h = n->H >> L, w = n->W >> L;
    PV_TRY(alloc_cview(n, &n->c_slab[L], B, h, w, slab_chunks(L)));
    PV_TRY(alloc_view(n, &n->head[L], B, h, w, 16, 16));
    if (L != 6) {
      PV_TRY(alloc_cview(n, &n->c_warp[L], B, h, w, LEVEL_CH[L] / 8));
    }
  }
  PV_TRY(alloc_cview(n, &n->c_ctxA, B, n->H >> 2, n->W >> 2, 16));
  PV_TRY(alloc_cview(n, &n->c_ctxB, B, n->H >> 2, n->W >> 2, 16));
  PV_TRY(alloc_view(n, &n->dc7out, B, n->H >> 2, n->W >> 2, 16, 16));
  // one launch plan (tensor maps + arguments) per layer
  auto plan = [&](ConvPlanUmma* pl, const CView& in, const CView* out_cp, const TView* out_f32, const ConvWeightsUmma& w,
                  const ConvGeom& g) {
    ConvOut o;
    if (out_cp) o.cp = *out_cp;
    if (out_f32) o.f32 = *out_f32;
    n->tensor_core_layers++;
    return plan_conv_umma(pl, in, o, w, g);
  };
  CView cur = n->c_img.slice(0, 27);
  for (int L = 1; L <= 6; L++) {
    ConvGeom g2 = ConvGeom::same3x3(1, 0.1f);
    g2.stride = 2;
    if (L == 1) {  // 1x1 over the im2col image (the stride and the padding were applied by the packing kernel)
      g2 = ConvGeom();
      g2.slope = 0.1f;
    }
    PV_TRY(plan(&n->pl_pyr[L][0], cur, &n->c_pyr[L][0], nullptr, n->wt_pyr[L][0], g2));
    PV_TRY(plan(&n->pl_pyr[L][1], n->c_pyr[L][0], &n->c_pyr[L][1], nullptr, n->wt_pyr[L][1], ConvGeom::same3x3(1, 0.1f)));
    PV_TRY(plan(&n->pl_pyr[L][2], n->c_pyr[L][1], &n->c_pyr[L][2], nullptr, n->wt_pyr[L][2], ConvGeom::same3x3(1, 0.1f)));
    cur = n->c_pyr[L][2];
  }
  for (int L = 6; L >= 2; L--) {
    const int tot = slab_chunks(L);
    const int hc = n->wt_head[L].Cout;
    TView ho = n->head[L].slice(0, hc);
    for (int i = 0; i < 5; i++) {
      CView in = n->c_slab[L].slice(DEC_IN_OFF[i] / 8, (tot - DEC_IN_OFF[i] / 8) * 8);
      CView out = n->c_slab[L].slice(DEC_OUT_OFF[i] / 8, DEC_OUT[i]);
      if (i < 4) {
        PV_TRY(plan(&n->pl_dec[L][i], in, &out, nullptr, n->wt_dec[L][i], ConvGeom::same3x3(1, 0.1f)));
      } else {  // convL_4 + the flow head's partial sums (linear, fp32) in one launch
        ConvOut o;
        o.cp = out; o.cp_channels = DEC_OUT[4];
        o.f32 = ho; o.f32_first = DEC_OUT[4]; o.f32_linear = true;
        n->tensor_core_layers++;
        PV_TRY(plan_conv_umma(&n->pl_dec[L][i], in, o, n->wt_dec[L][i], ConvGeom::same3x3(1, 0.1f)));
      }
    }
    {  // head remainder: the 32 channels convL_4 just produced, accumulated onto the partial sums
      ConvOut o;
      o.f32 = ho; o.f32_linear = true; o.f32_accumulate = true;
      n->tensor_core_layers++;
      PV_TRY(plan_conv_umma(&n->pl_head[L], n->c_slab[L].slice(0, DEC_OUT[4]), o, n->wt_head[L], ConvGeom::same3x3(1, 1.0f)));
    }
  }
  CView in = n->c_slab[2].slice(0, slab_chunks(2) * 8);
  CView bufs[2] = {n->c_ctxA, n->c_ctxB};
  for (int i = 0; i < 6; i++) {
    CView out = bufs[i & 1].slice(0, DC_OUT[i]);
    PV_TRY(plan(&n->pl_dc[i], in, &out, nullptr, n->wt_dc[i], ConvGeom::same3x3(DC_DIL[i], 0.1f)));
    in = out;
  }
  TView d7 = n->dc7out.slice(0, 2);
  PV_TRY(plan(&n->pl_dc7, in, nullptr, &d7, n->wt_dc7, ConvGeom::same3x3(1, 1.0f)));
  return alloc_io(n);
}

int run_middle_tc(premvos_pwc* n, cudaStream_t st) {
  const int B = n->B;
  for (int L = 1; L <= 6; L++)  // feature pyramid, both frames at once (PWCNet.py:183-194)
    for (int j = 0; j < 3; j++) PV_TRY(launch_conv_umma(n->pl_pyr[L][j], st));
  for (int L = 6; L >= 2; L--) {
    const int CL = LEVEL_CH[L];
    CView c1 = n->c_pyr[L][2].batch_range(0, B);
    CView c2 = n->c_pyr[L][2].batch_range(B, B);
    CView& slab = n->c_slab[L];
    CView f2 = c2, c1_slot;  // c1 slot stays null at level 6 (x = corr6 only, PWCNet.py:201)
    if (L != 6) {
      PV_TRY(warp_cp8(c2, slab.slice(up_chunk(L), 8), 0, WARP_SCALE[L], n->c_warp[L], st));  // :211,225,239,255
      f2 = n->c_warp[L];
      c1_slot = slab.slice(C1_CHUNK, CL);
    }
    PV_TRY(corr81_cp8(c1, f2, slab.slice(CORR_CHUNK, 81), c1_slot, 0.1f, st));  // corr + LeakyReLU (:197-198)
    for (int i = 0; i < 5; i++) PV_TRY(launch_conv_umma(n->pl_dec[L][i], st));
    PV_TRY(launch_conv_umma(n->pl_head[L], st));  // predict_flowL + upfeatL
    if (L != 2)
      PV_TRY(level_up_cp8(n->head[L], n->d_deconv_w[L], n->d_deconv_w[L] + 64, n->c_slab[L - 1].slice(up_chunk(L - 1), 8), st));
  }
  for (int i = 0; i < 6; i++) PV_TRY(launch_conv_umma(n->pl_dc[i], st));  // context network (:266)
  PV_TRY(launch_conv_umma(n->pl_dc7, st));
  return 0;
}

// ---- the network ---------------------------------------------------------------------------------
int run_front(premvos_pwc* n, const float* x_dev, cudaStream_t st) {
  if (n->opt_tensor_cores) return pack_pair_im2col_cp8(x_dev, n->B, n->H, n->W, n->c_img, st);
  return pack_pair_input(x_dev, n->B, n->H, n->W, n->img, st);
}

int run_middle(premvos_pwc* n, cudaStream_t st) {
  if (n->opt_tensor_cores) return run_middle_tc(n, st);
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
      PV_TRY(conv2d_simt(in, out, n->w_dec[L][i], 1, 1, 0.1f, st));
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
    PV_TRY(conv2d_simt(in, out, n->w_dc[i], 1, DC_DIL[i], 0.1f, st));
    in = out;
  }
  return 0;
}

int run_back(premvos_pwc* n, float* flow_dev, cudaStream_t st) {
  if (n->opt_tensor_cores)  // flow2 + dc_conv7(...) (PWCNet.py:267), NCHW out
    return flow_finish(n->head[2].slice(0, 2), n->dc7out.slice(0, 2), flow_dev, st);
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

// Stage-1 unit of work of the reference (calculate_flow, script_pwc_multi.py:33-70) minus file I/O and cv2.resize: takes the
// uint8 RGB frames (already resized to multiples of 64), does BGR / 255 / planar on the device.  4x less H2D traffic.
extern "C" int premvos_pwc_forward_host_u8(premvos_pwc_t* n, const unsigned char* frames_rgb_host, float* flow_host) {
  PV_CHECK(n && frames_rgb_host && flow_host, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_forward_host_u8: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_forward_host_u8: call premvos_pwc_finalize first");
  const size_t fin = (size_t)n->B * 2 * n->H * n->W * 3;
  const size_t fout = (size_t)n->B * 2 * (n->H / 4) * (n->W / 4) * sizeof(float);
  PV_CUDA(cudaMemcpyAsync(n->frames_u8, frames_rgb_host, fin, cudaMemcpyHostToDevice, n->stream));
  PV_TRY(frames_u8_to_x(n->frames_u8, n->x_in, n->B, n->H, n->W, n->stream));
  PV_TRY(enqueue_forward(n, n->x_in, n->flow_out, n->stream));
  PV_CUDA(cudaMemcpyAsync(flow_host, n->flow_out, fout, cudaMemcpyDeviceToHost, n->stream));
  PV_CUDA(cudaStreamSynchronize(n->stream));
  return 0;
}

// Same unit of work with the frames already on the device (a resident pipeline decodes / uploads each frame once).
extern "C" int premvos_pwc_forward_u8(premvos_pwc_t* n, const unsigned char* frames_rgb_dev, float* flow_dev, void* stream) {
  PV_CHECK(n && frames_rgb_dev && flow_dev, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_forward_u8: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_forward_u8: call premvos_pwc_finalize first");
  cudaStream_t st = (cudaStream_t)stream;
  PV_TRY(frames_u8_to_x(frames_rgb_dev, n->x_in, n->B, n->H, n->W, st));
  return enqueue_forward(n, n->x_in, flow_dev, st);
}

extern "C" int premvos_pwc_launches_per_forward(const premvos_pwc_t* n) { return n ? n->launches_per_forward : 0; }

extern "C" int premvos_pwc_tensor_core_layers(const premvos_pwc_t* n) { return n ? n->tensor_core_layers : 0; }

extern "C" int premvos_pwc_get_tensor(premvos_pwc_t* n, const char* name, float* host_out, int64_t* numel) {
  PV_CHECK(n && name && numel, PREMVOS_ERR_INVALID_ARG, "premvos_pwc_get_tensor: null argument");
  PV_CHECK(n->finalized, PREMVOS_ERR_NOT_READY, "premvos_pwc_get_tensor: network not finalized");
  const bool tc = n->opt_tensor_cores != 0;
  std::string k(name);
  TView v;       // fp32 channels-last source ...
  CView cv;      // ... or CP8 source (+ channel offset inside its first chunk plane)
  int ch_off = 0;
  bool found = false, is_cp = false;
  auto lvl = [&](size_t pos) { return (k.size() > pos && k[pos] >= '1' && k[pos] <= '6') ? k[pos] - '0' : -1; };
  if (k.size() == 3 && k[0] == 'c' && (k[1] == '1' || k[1] == '2') && lvl(2) > 0) {
    int L = lvl(2);
    if (tc) { cv = n->c_pyr[L][2].batch_range(k[1] == '1' ? 0 : n->B, n->B); is_cp = true; }
    else v = n->pyr[L][2].batch_range(k[1] == '1' ? 0 : n->B, n->B);
    found = true;
  } else if (k.rfind("corr", 0) == 0 && lvl(4) >= 2) {
    if (tc) { cv = n->c_slab[lvl(4)].slice(CORR_CHUNK, 81); is_cp = true; }
    else v = n->slab[lvl(4)].slice(BASE_OFF, 81);
    found = true;
  } else if (k.rfind("warp", 0) == 0 && lvl(4) >= 2 && lvl(4) <= 5) {
    if (tc) { cv = n->c_warp[lvl(4)]; is_cp = true; }
    else v = n->warpbuf[lvl(4)];
    found = true;
  } else if (k.rfind("flow", 0) == 0 && lvl(4) >= 2) {
    v = tc ? n->head[lvl(4)].slice(0, 2) : n->flow[lvl(4)].slice(0, 2); found = true;
  } else if ((k.rfind("up_flow", 0) == 0 || k.rfind("up_feat", 0) == 0) && lvl(7) >= 3) {
    int L = lvl(7);
    const int off = k[3] == 'f' && k[4] == 'l' ? 0 : 2;
    if (tc) { cv = n->c_slab[L - 1].slice(up_chunk(L - 1), 2); ch_off = off; is_cp = true; }
    else v = n->slab[L - 1].slice(BASE_OFF + 81 + LEVEL_CH[L - 1] + off, 2);
    found = true;
  } else if (k == "dc6") {
    if (tc) { cv = n->c_ctxB.slice(0, 32); is_cp = true; }
    else v = n->ctxB.slice(0, 32);
    found = true;
  }
  if (!found) return fail(PREMVOS_ERR_INVALID_ARG, "premvos_pwc_get_tensor: unknown tensor '%s'", name);
  *numel = is_cp ? (int64_t)cv.N * cv.C * cv.H * cv.W : (int64_t)v.N * v.C * v.H * v.W;
  if (!host_out) return 0;
  PV_CUDA(cudaDeviceSynchronize());
  float* dtmp = nullptr;
  PV_CUDA(cudaMalloc((void**)&dtmp, (size_t)(*numel) * sizeof(float)));
  int r = is_cp ? cp8_to_nchw(cv, ch_off, dtmp, nullptr) : view_to_nchw(v, dtmp, nullptr);
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
  for (int L = 1; L <= 6; L++)
    for (int j = 0; j < 3; j++) free_conv_weights_umma(&n->wt_pyr[L][j]);
  for (int L = 2; L <= 6; L++) {
    for (int i = 0; i < 5; i++) free_conv_weights_umma(&n->wt_dec[L][i]);
    free_conv_weights_umma(&n->wt_head[L]);
    if (n->d_deconv_w[L]) cudaFree(n->d_deconv_w[L]);
  }
  for (int i = 0; i < 6; i++) free_conv_weights_umma(&n->wt_dc[i]);
  free_conv_weights_umma(&n->wt_dc7);
  for (int L = 1; L <= 6; L++)
    for (int j = 0; j < 3; j++) free_conv_plan_umma(&n->pl_pyr[L][j]);
  for (int L = 2; L <= 6; L++) {
    for (int i = 0; i < 5; i++) free_conv_plan_umma(&n->pl_dec[L][i]);
    free_conv_plan_umma(&n->pl_head[L]);
  }
  for (int i = 0; i < 6; i++) free_conv_plan_umma(&n->pl_dc[i]);
  free_conv_plan_umma(&n->pl_dc7);
  free_small_conv_weights(&n->w_dc7);
  if (n->stream) cudaStreamDestroy(n->stream);
  delete n;
}
