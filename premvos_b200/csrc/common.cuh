// Shared declarations for the premvos_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <string.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/premvos_b200.h"

namespace premvos {

// ---- error plumbing --------------------------------------------------------------------------
void set_last_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define PV_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::premvos::fail((int)_e, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                             cudaGetErrorString(_e));                                      \
  } while (0)

#define PV_CHECK(cond, code, ...)                          \
  do {                                                     \
    if (!(cond)) return ::premvos::fail((code), __VA_ARGS__); \
  } while (0)

#define PV_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

extern std::atomic<int64_t> g_launch_count;
// Called once per kernel launch (also while capturing: one graph node == one launch).
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
// Per-launch profiling (bench.py's roofline leg): when enabled every launcher brackets its kernel
// with CUDA events on the launching stream and records the launch's algorithmic FLOPs / bytes.
bool profiling_enabled();
void prof_before(cudaStream_t st);
const char* prof_intern(const std::string& s);
void prof_after(const char* what, cudaStream_t st, double flops, double bytes);
inline int after_launch(const char* what, cudaStream_t st = nullptr, double flops = 0.0, double bytes = 0.0) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "launch of %s failed: %s", what, cudaGetErrorString(e));
  if (profiling_enabled()) prof_after(what, st, flops, bytes);
  return 0;
}

// ---- tensor view: channels-last activations -----------------------------------------------------
// Element (n,y,x,c) lives at p[((n*H + y)*W + x)*cs + coff + c].  A view can therefore name a
// channel range inside a wider "slab" (the DenseNet concat buffers of the PWC decoder).
// Two storage formats share the view: fp32 (p) or SPLIT bf16 (hi, lo with x = hi + lo, the operand
// format of the tensor-core convolutions; same bytes per element as fp32).
struct TView {
  float* p = nullptr;
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int N = 0, H = 0, W = 0;
  int cs = 0;    // pixel stride in floats
  int coff = 0;  // first channel of this view inside the pixel
  int C = 0;     // logical channels of this view
  TView slice(int off, int c) const {
    TView v = *this;
    v.coff = coff + off;
    v.C = c;
    return v;
  }
  TView batch_range(int n0, int n) const {
    TView v = *this;
    size_t off = (size_t)n0 * H * W * cs;
    if (p) v.p = p + off;
    if (hi) { v.hi = hi + off; v.lo = lo + off; }
    v.N = n;
    return v;
  }
  bool split() const { return hi != nullptr; }
  bool null() const { return p == nullptr && hi == nullptr; }
  size_t pixels() const { return (size_t)N * H * W; }
};

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

#ifdef __CUDACC__
// split-bf16 element access: x = hi + lo
__device__ __forceinline__ float ld_split(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long i) {
  return __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}
__device__ __forceinline__ float4 ld4_split(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long i) {  // i % 4 == 0
  const uint2 h = *reinterpret_cast<const uint2*>(hi + i);
  const uint2 l = *reinterpret_cast<const uint2*>(lo + i);
  float4 r;
  r.x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
  r.y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
  r.z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
  r.w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
  return r;
}
__device__ __forceinline__ void st_split(__nv_bfloat16* hi, __nv_bfloat16* lo, long i, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
#endif

// ---- convolution (fp32 SIMT implicit GEMM), conv_simt.cu ---------------------------------------
struct ConvWeightsSimt {
  float* w = nullptr;     // device [R*S][CinP][CoutP], zero padded
  float* bias = nullptr;  // device [CoutP]
  int R = 3, S = 3, Cin = 0, Cout = 0, CinP = 0, CoutP = 0;
};
// host_w: torch Conv2d layout [Cout][Cin][R][S]; host_b: [Cout]
int pack_conv_weights_simt(ConvWeightsSimt* out, const float* host_w, const float* host_b, int Cout,
                           int Cin, int R, int S);
void free_conv_weights_simt(ConvWeightsSimt* w);
// out = act(conv(in) + bias); act = LeakyReLU(slope) (slope 1 = identity, 0 = ReLU)
int conv2d_simt(const TView& in, const TView& out, const ConvWeightsSimt& w, int stride, int dil,
                float slope, cudaStream_t st);

// Small-Cout (<=4) 3x3 convolution, one warp per output pixel (predict_flow / dc_conv7).
struct SmallConvWeights {
  float* w = nullptr;     // device [9][Cout][CinP4]
  float* bias = nullptr;  // device [Cout]
  int Cin = 0, CinP = 0, Cout = 0;
};
int pack_small_conv_weights(SmallConvWeights* out, const float* host_w, const float* host_b, int Cout, int Cin);
void free_small_conv_weights(SmallConvWeights* w);
// out (channels-last view) = conv3x3(in) + bias [+ addend]; if nchw_out != nullptr the result is
// (also) written as contiguous NCHW [N,Cout,H,W] there.
int conv3x3_small_cout(const TView& in, const TView& out, const SmallConvWeights& w, const TView* addend,
                       float* nchw_out, cudaStream_t st);

// ConvTranspose2d 4x4 stride 2 pad 1 with Cout == 2 (deconvL / upfeatL), one warp per output pixel.
struct DeconvWeights {
  float* w = nullptr;     // device [16][2][CinP4]   (tap = ky*4+kx)
  float* bias = nullptr;  // device [2]
  int Cin = 0, CinP = 0;
};
int pack_deconv_weights(DeconvWeights* out, const float* host_w /*[Cin][2][4][4]*/, const float* host_b, int Cin);
void free_deconv_weights(DeconvWeights* w);
int deconv4x4s2_cout2(const TView& in, const TView& out /*[N,2H,2W], C=2*/, const DeconvWeights& w, cudaStream_t st);

// ---- CP8: channel-chunk-planar split-bf16 activations (tensor-core path) ----------------------------
// Element (n, c, y, x) of a view lives at  plane[(((n*chunks + c0 + c/8) * H + y) * W + x) * 8 + c%8]
// in both the hi and the lo plane (x = hi + lo).  A view names a range of 8-channel chunk planes of a
// wider buffer (the DenseNet slabs of the PWC decoder); channels of the last chunk beyond C hold zeros.
struct CView {
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int N = 0, H = 0, W = 0;
  int chunks = 0;  // chunk planes per image in the underlying buffer
  int c0 = 0;      // first chunk plane of this view
  int C = 0;       // logical channels (the view spans (C+7)/8 planes)
  CView slice(int chunk_off, int c) const {
    CView v = *this;
    v.c0 = c0 + chunk_off;
    v.C = c;
    return v;
  }
  CView batch_range(int n0, int n) const {
    CView v = *this;
    size_t off = (size_t)n0 * chunks * H * W * 8;
    v.hi = hi + off; v.lo = lo + off; v.N = n;
    return v;
  }
  bool null() const { return hi == nullptr; }
  int vchunks() const { return (C + 7) / 8; }
  size_t pixels() const { return (size_t)N * H * W; }
};

// ---- F8: channel-chunk-planar fp32 activations ----------------------------------------------------------
// Same geometry as CP8 with one fp32 plane instead of two bf16 planes: element (n, c, y, x) lives at
// p[(((n*chunks + c0 + c/8) * H + y) * W + x) * 8 + c%8].  Used for tensors that never are tensor-core operands (inputs of
// depthwise convolutions, residuals): no hi/lo split on the way out of a GEMM, no re-assembly on the way into the next kernel.
struct FView {
  float* p = nullptr;
  int N = 0, H = 0, W = 0;
  int chunks = 0, c0 = 0, C = 0;
  bool null() const { return p == nullptr; }
  int vchunks() const { return (C + 7) / 8; }
};
// TMA descriptor over fp32 data (rank <= 5, no swizzle, zero fill outside): map128 = 128 bytes, 64-byte aligned
// TMA descriptor over bf16 data (one plane of a CP8 buffer), same conventions as encode_tensor_map_f32
int encode_tensor_map_bf16(void* map128, const __nv_bfloat16* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);
// corr_tma.cu: the same correlation on CP8 features inside PWC-Net (output = 11 chunk planes of the decoder slab, LeakyReLU fused,
// optional copy of f1 next to it), TMA-staged
int corr81_cp8_tma(const CView& f1, const CView& f2, const CView& out, const CView& c1_copy, float slope, cudaStream_t st);
// corr_tma.cu: 81-displacement correlation, NCHW fp32, TMA-staged; returns 1 when not applicable (W % 4 != 0 or unaligned pointers)
int corr81_nchw_tma(const float* f1, const float* f2, float* out, int B, int C, int H, int W, cudaStream_t st);
int encode_tensor_map_f32(void* map128, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

// ---- convolution on tcgen05 tensor cores (split-bf16 x3, fp32 accumulate), conv_umma.cu ----------
struct ConvWeightsUmma {
  __nv_bfloat16* w = nullptr;  // device, packed shared-memory images [ntile][kblock][tap][hi|lo][KC][BN][8]
  float* bias = nullptr;       // device [ntiles*BN]
  int R = 3, S = 3, Cin = 0, CinPhys = 0, Cout = 0, KC = 2, kblocks = 0, BN = 0, ntiles = 0;
  int pair = 0;  // packed for conv_pair_kernel (cta_group::2): [ntile][kblock][CTA rank][hi|lo][KC][128][8]
};
struct ConvGeom {
  int stride = 1, dil = 1;
  int stride_x = 0;   // horizontal stride when it differs from the vertical one (0: same as `stride`); tap mode only
  int pad_t = 0, pad_l = 0, pad_b = 0, pad_r = 0;
  float slope = 1.f;  // LeakyReLU slope applied after bias (+ residual): 1 = identity, 0 = ReLU
  static ConvGeom same3x3(int dil, float slope) { ConvGeom g; g.dil = dil; g.pad_t = g.pad_l = g.pad_b = g.pad_r = dil; g.slope = slope; return g; }
};
struct ConvOut {
  CView cp;    // CP8 split output (optional)
  TView f32;   // fp32 channels-last output (optional; small heads)
  CView res;   // optional residual added before the activation (same shape as the output)
  FView f8;    // F8 output (optional, instead of or next to cp): all output channels, after the activation
  FView res_f8;  // residual in F8 (instead of res)
  // channel routing (fused heads): output channels [0, cp_channels) go to `cp` (cp_channels < 0: all of them);
  // channels [f32_first, Cout) go to `f32` at channel index c - f32_first
  int cp_channels = -1, f32_first = 0;
  bool f32_linear = false;      // fp32 outputs skip the activation
  bool f32_accumulate = false;  // fp32 outputs are added to the buffer's current contents
};
struct ConvPlanUmma {  // everything one launch needs; built once per layer at finalize time
  alignas(64) unsigned char map_a_hi[128];
  alignas(64) unsigned char map_a_lo[128];
  alignas(64) unsigned char map_w[128];   // CTA-pair kernel: packed weights as 128-byte rows
  alignas(16) unsigned char args[384];
  int grid_x = 0, grid_y = 0, grid_z = 1, smem_bytes = 0, halo = 0, MT = 0, N = 0, ctas_per_sm = 1;
  void* scratch = nullptr;   // split-K partial sums (owned by the plan, see free_conv_plan_umma)
  double flops = 0, bytes = 0;
};
// host_w: torch Conv2d layout [Cout][Cin][R][S].  cin_map (optional): physical channel (inside the input
// view) of every reference input channel, cin_phys = physical channel count of that view.
int pack_conv_weights_umma(ConvWeightsUmma* out, const float* host_w, const float* host_b, int Cout, int Cin, int R, int S,
                           const int* cin_map = nullptr, int cin_phys = 0, int kc_hint = 0, long m_hint = 0, bool allow_pair = false);
void free_conv_weights_umma(ConvWeightsUmma* w);
struct ConvPlanUmma;
void free_conv_plan_umma(ConvPlanUmma* plan);
// Scratch of the CTA-pair kernel's stream-K schedule (partial sums + flags).  One per network / stream: the layers of a network
// run one after the other and share it; launches that may run CONCURRENTLY must not (ws == nullptr: the plan owns a private one).
struct ConvWorkspace {
  void* base = nullptr;
  float* partial = nullptr;
  unsigned* flags = nullptr;
};
int conv_workspace_reserve(ConvWorkspace* ws);
void conv_workspace_free(ConvWorkspace* ws);
int plan_conv_umma(ConvPlanUmma* plan, const CView& in, const ConvOut& out, const ConvWeightsUmma& w, const ConvGeom& g,
                   ConvWorkspace* ws = nullptr);
// active_n >= 0: process only the first active_n images of the planned batch
int launch_conv_umma(const ConvPlanUmma& plan, cudaStream_t st, int active_n = -1);

// ---- CP8 companions of the tensor-core path, cp8_ops.cu -----------------------------------------------
// x: NCHW fp32 [B,6,H,W] -> img CP8 [2B, 1 chunk, H, W] (ch 0..2 = BGR of image 1 for n < B, image 2 for n >= B)
int pack_pair_input_cp8(const float* x_nchw, int B, int H, int W, const CView& img, cudaStream_t st);
// uint8 RGB frame pairs [B][2][H][W][3] -> x NCHW fp32 [B][6][H][W] (BGR / 255, script_pwc_multi.py:47-56)
int frames_u8_to_x(const unsigned char* frames_rgb, float* x_nchw, int B, int H, int W, cudaStream_t st);
// same input -> im2col of the first 3x3/stride-2 convolution: img CP8 [2B, 4 chunks, H/2, W/2], channel (r*3+s)*3 + c (27 of 32 used)
int pack_pair_im2col_cp8(const float* x_nchw, int B, int H, int W, const CView& img, cudaStream_t st);
// PWCDCNet.warp on CP8 features; flow = channels [flow_ch, flow_ch+2) of chunk plane flow.c0
int warp_cp8(const CView& x2, const CView& flow, int flow_ch, float flow_scale, const CView& out, cudaStream_t st);
// 9x9 cost volume: out = 11 chunk planes (81 channels + 7 zeros), LeakyReLU fused; c1_copy (optional) receives f1's planes
int corr81_cp8(const CView& f1, const CView& f2, const CView& out, const CView& c1_copy, float slope, cudaStream_t st);
// head: fp32 channels-last [B,H,W,16] = {flow 2, upfeat phases 8 ((py*2+px)*2+co), ...}.  Writes chunk plane `dst`
// of the next level's slab: [deconv(flow) 2, pixel-shuffled up_feat 2, 0 0 0 0] at [B,2H,2W].
int level_up_cp8(const TView& head, const float* deconv_w /*device [2][2][4][4]*/, const float* deconv_b, const CView& dst,
                 cudaStream_t st);
// out NCHW [B,2,H,W] = a[...,0:2] + b[...,0:2]   (flow2 + dc_conv7)
int flow_finish(const TView& a, const TView& b, float* out_nchw, cudaStream_t st);
// NCHW fp32 <-> CP8 (ch_off = first channel inside chunk plane v.c0; writes zero to the padding channels when ch_off == 0)
int nchw_to_cp8(const float* src, const CView& dst, cudaStream_t st);
int cp8_to_nchw(const CView& src, int ch_off, float* dst, cudaStream_t st);

// ---- detection-side kernels of the proposal network, det_ops.cu ---------------------------------------
int det_preprocess(const float* img_hwc_dev, const CView& out, cudaStream_t st);           // basemodel.py:12-26
// the same normalisation written as the ROW IM2COL of the 7x7 / stride-2 stem: out [1][3 chunks][H][Wo], channel s*3 + c of pixel
// (y, xo) = normalised img(y, 2*xo + s - pad_l, c) (0 outside the image: the stem pads the NORMALISED image with zeros)
int det_preprocess_rows7(const float* img_hwc_dev, int W, int pad_l, const CView& out, cudaStream_t st);
int det_maxpool3x3s2(const CView& in, const CView& out, cudaStream_t st);                  // basemodel.py:81-82
int det_rpn_decode(const float* rpn, int cs, int fh, int fw, int na, const float* cell_anchors, float stride, float clip,
                   float* scores, float* boxes, cudaStream_t st);                          // model.py:114-139
// k best of n scores as (index, score) sorted by (score desc, index asc); k <= 1024
int det_topk(const float* scores, int n, int k, int* out_idx, float* out_score, int* out_count, cudaStream_t st);
int det_gather_clip_valid(const float* boxes, const int* idx, const float* score, const int* count, float img_h, float img_w,
                          float min_size, float* out_boxes, float* out_scores, int* out_src, int* out_count, cudaStream_t st);
// greedy NMS over <= 1024 boxes already sorted by descending score; keep[] = positions, TensorFlow IoU arithmetic
int det_nms(const float* boxes_sorted, const int* count, float thr, int max_out, uint32_t* mask_scratch /*[1024*32]*/, int* keep,
            int* keep_count, cudaStream_t st);
int det_gather_proposals(const float* boxes, const float* scores, const int* keep, const int* keep_count, int max_out,
                         float* out_boxes, float* out_scores, cudaStream_t st);
int det_roi_align(const CView& fm, const float* rois, float spatial_scale, int out_size, const CView& out, cudaStream_t st);
int det_gap_fc(const CView& feat, const float* Wt, const float* bias, int nout, float* pooled_out, float* out, cudaStream_t st);
// mask head tail: d = ReLU(deconv 2x2 s2) as CP8 [N, 4*256, 7, 7] (phase-major channels) -> sigmoid(1x1 conv) fp32 [N,14,14]
int det_mask_head(const CView& d, const float* w /*[256]*/, const float* b /*[1]*/, float* masks, cudaStream_t st);   // model.py:495-509
// eval.py:35-58 for n boxes: masks fp32 [n,M,M], boxes fp32 [n,4] x1y1x2y2 (clipped to the image) -> uint8 [n,H,W]
int det_fill_full_masks(const float* masks, const float* boxes, int n, int M, int H, int W, unsigned char* out, cudaStream_t st);
struct DetTailArgs {
  const float* logits; int nfc, nsecond;
  const float* prop_boxes; const int* prop_count;
  float img_h, img_w, clip, score_thresh, nms_thresh;
  int max_rois, results_per_im;
  float* all_probs; float* all_boxes; float* second_probs;
  int* n_out; float* final_boxes; float* final_probs; int64_t* final_labels; float* final_posterior;
  int64_t* second_final_labels; float* second_final_posterior; int* final_box_index;
};
int det_u8_to_f32(const unsigned char* src, float* dst, long n, cudaStream_t st);
int det_frcnn_tail(const DetTailArgs& a, cudaStream_t st);                                  // train.py:275-295

// ---- depthwise 3x3 on F8 inputs (TMA-pipelined), dw_f8.cu ------------------------------------------------------
struct DwF8Plan {
  alignas(64) unsigned char map_in[128];
  alignas(16) unsigned char args[192];
  int smem_bytes = 0, fast = 0, N = 0, tiles = 0, nch = 0;
  double bytes_per_image = 0, flops_per_image = 0;
};
// w: [9][round_up(C,8)] with BN scale folded, bias [round_up(C,8)]; input coordinate = o*stride + tap*rate - pad; out is CP8
int plan_depthwise3x3_f8(DwF8Plan* plan, const FView& in, const CView& out, const float* w, const float* bias, int stride, int rate, int pad,
                         bool pre_relu, bool post_relu);
int launch_depthwise3x3_f8(const DwF8Plan& plan, int n_active, cudaStream_t st);

// ---- fused separable convolution (depthwise 3x3 computed into the pointwise GEMM's A operand), conv_umma.cu -----------------------
// Eligible: F8 input, stride 1, rate 1, SAME padding, a pointwise layer with ONE channel tile (Cout <= 128, packed with kc_hint 4),
// F8 and / or CP8 output, no residual.  dw_w [9][cpad] (BN scale folded), dw_b [cpad]; cpad >= 32 * k-blocks.
struct SepConvPlan {
  alignas(64) unsigned char map_in[128];
  alignas(16) unsigned char args[192];
  int smem_bytes = 0, N = 0;
  double flops_per_image = 0, bytes_per_image = 0;
};
int plan_sepconv_fused(SepConvPlan* plan, const FView& in, const float* dw_w, const float* dw_b, int cpad, bool pre_relu, bool post_relu,
                       const ConvWeightsUmma& pw, const ConvOut& out, float slope);
int launch_sepconv_fused(const SepConvPlan& plan, int n_active, cudaStream_t st);

// ---- refinement-network kernels, refine_ops.cu ------------------------------------------------------------
// frame uint8 RGB [H,W,3] + boxes [N,4] (x,y,w,h) -> network input CP8 [N,1 chunk,S,S] (RGB in [-1,1], guidance -1/+1) + crop boxes
int refine_make_input(const unsigned char* frame_rgb, int H, int W, const float* boxes_xywh, int N, int S, const CView& out, int* crops,
                      cudaStream_t st);
// w: [9][round_up(C,8)] with BN scale folded, bias [round_up(C,8)]; input coordinate = o*stride + tap*rate - pad
int depthwise3x3_cp8(const CView& in, const CView& out, const float* w, const float* bias, int stride, int rate, int pad, bool pre_relu,
                     bool post_relu, int n_active, cudaStream_t st);
int depthwise3x3_cp8_ex(const CView& in, const CView* up_src, const CView& out, const float* w, const float* bias, int w_c0, int cpad, int stride,
                        int rate, int pad, bool pre_relu, bool post_relu, int n_active, cudaStream_t st);
int resize_bilinear_ac_cp8(const CView& in, const CView& out, int n_active, cudaStream_t st);   // align_corners=True
int broadcast_vec_cp8(const float* vec, int C, bool relu, const CView& out, int n_active, cudaStream_t st);
int gap_fc_relu(const CView& feat, const float* Wt /*[C][nout]*/, const float* bias, int nout, bool relu, float* pooled_scratch /*[N][C]*/,
                float* out, int n_active, cudaStream_t st);
// logits fp32 channels-last [N,h,w,cs] -> per-proposal full-frame masks (0/1), optional posteriors, sum of (2p'-1) over the frame
int refine_output(const TView& logits, const int* crops, int N, int S, int H, int W, unsigned char* mask, float* posterior,
                  double* conf_sum, cudaStream_t st);
// conf_score[i] = (float)(conf_sum[i] / hw)  (refinement_net_functions.py:58-62: mean over the whole frame)
int refine_conf_finish(const double* conf_sum, int N, long hw, float* conf_out, cudaStream_t st);

// ---- layout / format conversion, layout.cu ---------------------------------------------------------
// NCHW fp32 [N,C,H,W] -> channels-last view (fp32 or split)
int nchw_to_view(const float* src, const TView& dst, cudaStream_t st);
// channels-last view (fp32 or split) -> NCHW fp32
int view_to_nchw(const TView& src, float* dst, cudaStream_t st);

// ---- correlation / warp / layout, corr.cu warp.cu ------------------------------------------------
// 9x9 (md=4) cost volume on channels-last features; writes 81 channels (LeakyReLU(slope) fused) into
// `out` and, if c1_copy.p != nullptr, copies f1's channels there (the decoder slab's c1 slot).
int corr81_nhwc(const TView& f1, const TView& f2, const TView& out, const TView& c1_copy, float slope,
                cudaStream_t st);
// PWCDCNet.warp: out = bilinear(x2, pos + flow*flow_scale) * validity mask
int warp_nhwc(const TView& x2, const TView& flow, float flow_scale, const TView& out, cudaStream_t st);
// x: NCHW [B,6,H,W] -> img: channels-last [2B,H,W,4] (image 1 of every pair first, then image 2)
int pack_pair_input(const float* x_nchw, int B, int H, int W, const TView& img, cudaStream_t st);

}  // namespace premvos
