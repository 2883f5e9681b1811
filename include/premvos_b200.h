/* premvos_b200 -- C ABI of the B200-native PReMVOS hot path.
 *
 * Plain pointers and sizes only (no torch / THC types).  Every entry point returns 0 on success,
 * a positive cudaError_t value when CUDA failed, or a negative PREMVOS_ERR_* code for argument
 * errors; premvos_last_error() returns a human-readable description for the calling thread.
 * Nothing here allocates behind the caller's back inside a forward call: all device memory of a
 * network handle is allocated by *_create / *_finalize.
 *
 * Reference interfaces replaced (paths relative to the reference repo, code/optical_flow_net-PWC-Net):
 *   premvos_corr_forward   <- int corr_cuda_forward(THCudaTensor* input1, input2, rbot1, rbot2, output,
 *                                 int pad_size, kernel_size, max_displacement, stride1, stride2,
 *                                 corr_type_multiply)
 *                             external_packages/correlation-pytorch-master/correlation-pytorch/
 *                             correlation_package/src/corr_cuda.h:1-11, corr_cuda.c:7-82
 *                             (bound by cffi in correlation_package/_ext/corr/__init__.py:1-12 and
 *                             called from functions/corr.py:24-32)
 *   premvos_pwc_*          <- models.pwc_dc_net(path) / PWCDCNet.forward, models/PWCNet.py:179-272,
 *                             496-505, driven by script_pwc_multi.py:33-70 (one device round trip
 *                             per frame pair)
 *   premvos_propnet_*      <- tensorpack OfflinePredictor `pred_func(img)` over proposal_net/train.py's
 *                             Model (train.py:107-309, 653-657), called by eval.py:61-110 / train.py:508
 *   premvos_refnet_*       <- refinement_net Engine: `engine.trainer.validation_step(feed_dict, extraction_keys)` per
 *                             proposal (core/Trainer.py:128-133), as driven by MergeTrack/refinement_net_functions.py:38-65
 *   premvos_reidnet_*      <- MergeTrack/ReID_net_functions.py:26-45 `add_ReID`: ReID_net.session.run([outputs, crop_list],
 *                             feed_dict={image, boxes}) on the graph of ReID_net/configs/live
 *   premvos_conv2d_forward <- one nn.Conv2d / tensorpack Conv2D layer (bring-up hook)
 */
#ifndef PREMVOS_B200_H
#define PREMVOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PREMVOS_OK 0
#define PREMVOS_ERR_INVALID_ARG (-1)   /* null pointer, non-positive size, size not a multiple of 64 ... */
#define PREMVOS_ERR_UNSUPPORTED (-2)   /* a configuration the hot path never uses (see each function)  */
#define PREMVOS_ERR_UNKNOWN_PARAM (-3) /* premvos_pwc_set_param: name is not a PWC-DC-Net state_dict key */
#define PREMVOS_ERR_BAD_SHAPE (-4)     /* premvos_pwc_set_param: numel does not match the parameter    */
#define PREMVOS_ERR_NOT_READY (-5)     /* forward before finalize, or parameters missing at finalize   */
#define PREMVOS_ERR_NO_DEVICE (-6)     /* no sm_100 device visible                                      */

/* Library / build identification, e.g. "premvos_b200 0.1 sm_100a". */
const char* premvos_version(void);
/* Description of the last error on this thread ("" if none). */
const char* premvos_last_error(void);
/* Total number of CUDA kernels launched (or captured into graphs) by this library so far in this
 * process; bench.py reads it around the timed region to report `gpu_launches`. */
int64_t premvos_kernel_launch_count(void);
/* Per-launch profiling for bench.py's roofline leg.  Between begin and end every kernel launched by
 * this library is bracketed by CUDA events on its launching stream (CUDA graphs are bypassed while
 * profiling).  premvos_profile_end synchronises and writes one text line per kernel name into buf:
 * "name launches total_ms algorithmic_flops algorithmic_bytes\n".  Returns 0; when buflen is too small for the report, nothing is
 * written and the number of bytes needed is returned (> 0): call again with a buffer of that size, the report is kept.  The record
 * list is guarded by a mutex (launchers may run on several host threads). */
int premvos_profile_begin(void);
int premvos_profile_end(char* buf, int buflen);

/* ---------------------------------------------------------------------------------------------
 * Correlation (cost volume), forward only, "multiply" variant.
 *
 * input1, input2 : device, contiguous fp32 NCHW [batch, channels, height, width]
 * output         : device, contiguous fp32 NCHW [batch, (2*(max_displacement/stride2)+1)^2, OH, OW]
 *                  with OH = ceil((height + 2*pad_size - 2*(max_displacement + (kernel_size-1)/2)) / stride1)
 *                  (corr_cuda.c:23-45); caller-owned -- unlike the reference the callee never resizes,
 *                  zero-fills scratch tensors or frees anything (corr_cuda.c:52-60,77-78).
 * out[b, (dy/stride2+r)*(2r+1) + (dx/stride2+r), y, x] =
 *        (1/(k*k*C)) * sum_{j,i<k} sum_c in1pad[b,c,y1+j,x1+i] * in2pad[b,c,y1+dy+j,x1+dx+i]
 * with x1 = x*stride1 + max_displacement (corr_cuda_kernel.cu:59-127), zero outside the images.
 * corr_type_multiply must be 1 (PWCNet.py:69); 0 (the L1 "subtract" variant) returns
 * PREMVOS_ERR_UNSUPPORTED.  stream is a cudaStream_t (NULL = default stream).
 * The reference returns 1 always and exit(-1)s on a launch error; this returns an error code.
 * PWC-Net's configuration (pad 4, kernel 1, displacement 4, strides 1) runs the TMA-staged kernel (csrc/corr_tma.cu) when
 * width % 4 == 0 and the pointers are 16-byte aligned, the per-pixel tile kernel otherwise; any other configuration the generic one.
 * --------------------------------------------------------------------------------------------- */
int premvos_corr_output_shape(int height, int width, int pad_size, int kernel_size,
                              int max_displacement, int stride1, int stride2, int* out_channels,
                              int* out_height, int* out_width);
int premvos_corr_forward(const float* input1, const float* input2, float* output, int batch,
                         int channels, int height, int width, int pad_size, int kernel_size,
                         int max_displacement, int stride1, int stride2, int corr_type_multiply,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * One convolution layer on the tensor-core path (bring-up / parity hook; the networks below fuse their
 * layers and never go through NCHW).  Replaces a torch nn.Conv2d(+LeakyReLU) forward as the reference
 * builds it in models/PWCNet.py:24-28, or a tensorpack Conv2D(+BN folded)+ReLU in proposal_net/basemodel.py:51-60.
 *
 * x        : device fp32 NCHW [batch, cin, height, width]
 * w, bias  : HOST fp32, torch layout [cout, cin, kh, kw] / [cout] (bias may be NULL)
 * residual : device fp32 NCHW [batch, cout, OH, OW] or NULL; added before the activation
 * out      : device fp32 NCHW [batch, cout, OH, OW], OH = (height + pad_top + pad_bottom - dilation*(kh-1) - 1)/stride + 1
 * out = act(conv(x) + bias + residual), act = LeakyReLU(slope) (slope 1 = identity, 0 = ReLU); cross-correlation,
 * explicit zero padding per side (so TF "SAME"/tensorpack asymmetric pads are expressible).  stride 1 or 2.
 * Arithmetic: split-bf16 x3 tcgen05 MMAs, fp32 accumulate (~1e-5 relative).  Synchronises `stream`.
 * --------------------------------------------------------------------------------------------- */
int premvos_conv2d_forward(const float* x, const float* w, const float* bias, const float* residual, float* out,
                           int batch, int cin, int height, int width, int cout, int kh, int kw, int stride,
                           int dilation, int pad_top, int pad_left, int pad_bottom, int pad_right, float slope,
                           void* stream);

/* One separable convolution on the FUSED path (bring-up / parity hook): depthwise 3x3 (SAME, stride 1) computed into the A
 * operand of the pointwise tcgen05 GEMM -- slim `separable_conv2d` + folded BatchNorms as refinement_net/network/deeplab/core/
 * xception.py:152-190 builds it.  x device fp32 NCHW [batch, channels, height, width]; dw_w HOST [channels][3][3], dw_bias HOST
 * [channels] or NULL; pw_w HOST [cout][channels], pw_bias HOST [cout] or NULL; out device fp32 NCHW [batch, cout, height, width];
 * out = act(pw(relu_mid?(dw(relu_in?(x)) + dw_bias)) + pw_bias), act = LeakyReLU(slope).  cout <= 128.  Synchronises `stream`. */
int premvos_sepconv2d_forward(const float* x, const float* dw_w, const float* dw_bias, const float* pw_w, const float* pw_bias,
                              float* out, int batch, int channels, int height, int width, int cout, int relu_in, int relu_mid,
                              float slope, void* stream);

/* ---------------------------------------------------------------------------------------------
 * cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 images, on the device, BIT-EXACT with OpenCV
 * (8-bit linear resize is integer arithmetic on 11-bit fixed-point coefficients, modules/imgproc/src/resize.cpp).
 * Replaces the host-side resizes of the reference's stage drivers: script_pwc_multi.py:38-45 (frames -> multiples of 64) and
 * proposal_net eval.py:75-78 / common.py:35-62 (CustomResize -> cv2.resize).
 * src_dev  uint8 [batch, src_h, src_w, channels] (device), dst_dev uint8 [batch, dst_h, dst_w, channels] (device);
 * channels 1 or 3; reverse_channels != 0 writes the channels in reverse order (RGB frame -> resized BGR image: resize acts per
 * channel).  Enqueues on `stream`, does not synchronise.  The coefficient tables of a geometry are built on the first call with
 * it (one small device allocation, cached for the life of the process).  An exact 2x down-scale in both directions (which
 * OpenCV turns into INTER_AREA) returns PREMVOS_ERR_UNSUPPORTED.
 * --------------------------------------------------------------------------------------------- */
int premvos_resize_linear_u8(const unsigned char* src_dev, int batch, int src_h, int src_w, unsigned char* dst_dev, int dst_h,
                             int dst_w, int channels, int reverse_channels, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MergeTrack's live mask propagation, on the device.
 *
 * premvos_warp_masks_u8 replaces MergeTrack/merge_functions.py:209-217 `warp_flow(img, flow, binarize)` for all masks of a
 * frame at once (the loop of warp_proposals, :226) and the `toBbox` of :231:
 *     map = (x, y) - flow;  res = cv2.remap(mask, map, None, cv2.INTER_LINEAR);  if binarize: res = (res == 1)
 * BIT-EXACT with OpenCV's 8-bit remap (map quantised to 1/32 pixel, 15-bit bilinear table, constant border 0).
 *   masks_dev uint8 [n, height, width] (device), flow_dev float32 [height, width, 2] (device, (u, v) interleaved = the
 *   layout of a .flo file / of premvos_flow_postprocess), out_dev uint8 [n, height, width] (device, != masks_dev),
 *   bbox_dev float32 [n, 4] (device, 16-byte aligned, may be NULL): the tight box [x, y, w, h] of every warped mask's
 *   non-zero pixels, zeros for an empty one (pycocotools rleToBbox) -- the `bbox` do_refinement reads
 *   (MergeTrack/refinement_net_functions.py:44), so premvos_refnet_forward can consume it without a host round trip.
 * Enqueues on `stream` (1 launch, 3 with bbox_dev), never synchronises, allocates nothing.  n == 0 is a no-op; with bbox_dev
 * n <= 3072 (the per-CTA boxes live in shared memory).
 *
 * premvos_flow_postprocess replaces optical_flow_net-PWC-Net/script_pwc_multi.py:59-68 for a batch: flow2 float32
 * [batch, 2, net_h/4, net_w/4] (what premvos_pwc_forward writes) -> x20 -> cv2.resize of u and v to (width, height)
 * (float32 INTER_LINEAR) -> u *= width/net_w, v *= height/net_h -> out float32 [batch, height, width, 2] (the payload of a
 * .flo file, what MergeTrack's get_flow returns).  Same coordinate arithmetic as OpenCV; products and sums are rounded
 * to float32 one by one.  Enqueues on `stream`; the coefficient tables of a geometry are cached on the first call.
 * --------------------------------------------------------------------------------------------- */
int premvos_warp_masks_u8(const unsigned char* masks_dev, int n, int height, int width, const float* flow_dev,
                          unsigned char* out_dev, float* bbox_dev, int binarize, void* stream);
int premvos_flow_postprocess(const float* flow2_dev, int batch, int net_h, int net_w, float* out_dev, int height, int width,
                             void* stream);

/*
 * premvos_pack_mask_bits: the per-proposal full-frame masks the refinement stage returns (refinement_net/forwarding/
 * FewShotSegmentationForwarder.py:139-141 hands `mask` to pycocotools' RLE encoder one proposal at a time) packed 8 pixels per
 * byte on the device before they leave it: masks_dev uint8 [n_masks, hw] (any non-zero = 1) -> out_dev [n_masks, 8 * ceil(hw / 64)]
 * bytes, pixel i of a mask = bit (i % 8) of byte (i / 8) (numpy.unpackbits(..., bitorder="little")); every mask starts on an
 * 8-byte boundary, out_dev must be 8-byte aligned.  Enqueues on `stream`, never synchronises.
 */
int premvos_pack_mask_bits(const unsigned char* masks_dev, long long n_masks, long long hw, unsigned char* out_dev, void* stream);

/*
 * Coefficient tables of premvos_resize_linear_u8 / premvos_flow_postprocess (OpenCV's per-axis source indices and weights for one
 * (source size, destination size) pair, a few KB).  Both enqueue-only entry points build the tables of a new geometry on first use:
 * that FIRST call allocates device memory and copies synchronously, i.e. it is not capturable into a CUDA graph and may block.
 * Call the matching *_prepare once per geometry beforehand (same device) to take that cost out of the stream; *_release frees every
 * table of this process (they are rebuilt on demand).
 */
int premvos_resize_linear_u8_prepare(int src_h, int src_w, int dst_h, int dst_w);
void premvos_resize_linear_u8_release(void);
int premvos_flow_postprocess_prepare(int net_h, int net_w, int height, int width);
void premvos_flow_postprocess_release(void);

/* ---------------------------------------------------------------------------------------------
 * PWC-DC-Net forward (optical flow).
 *
 * Life cycle: create(batch,H,W) -> set_param(name, host fp32 data) for each of the 128 state_dict
 * tensors (names exactly as in the reference checkpoint: "conv1a.0.weight", "predict_flow6.bias",
 * "upfeat5.weight", "dc_conv7.weight" ...; layouts as torch stores them: Conv2d [Cout,Cin,3,3],
 * ConvTranspose2d [Cin,Cout,4,4]) -> finalize() (packs + uploads weights, allocates all activation
 * buffers, captures the CUDA graph) -> forward()* -> destroy().
 *
 * x    : fp32 NCHW [batch, 6, height, width], channels = BGR/255 of frame 1 then frame 2
 *        (script_pwc_multi.py:47-56); height and width must be multiples of 64.
 * flow : fp32 NCHW [batch, 2, height/4, width/4] = the eval-mode output `flow2` (PWCNet.py:269-272),
 *        in units of input pixels / 20.
 * premvos_pwc_forward      : x and flow are DEVICE pointers; enqueues on `stream`, does not sync.
 * premvos_pwc_forward_host : x and flow are HOST pointers (pinned or pageable); copies x to the
 *        device, runs the network, copies flow back and synchronises -- the end-to-end call.
 * --------------------------------------------------------------------------------------------- */
typedef struct premvos_pwc premvos_pwc_t;

int premvos_pwc_create(premvos_pwc_t** out, int batch, int height, int width);
int premvos_pwc_set_param(premvos_pwc_t* net, const char* name, const float* host_data, int64_t numel);
int premvos_pwc_finalize(premvos_pwc_t* net);
int premvos_pwc_forward(premvos_pwc_t* net, const float* x_dev, float* flow_dev, void* stream);
int premvos_pwc_forward_host(premvos_pwc_t* net, const float* x_host, float* flow_host);
/* The stage-1 unit of work (calculate_flow, script_pwc_multi.py:33-70) on decoded frames: frames = HOST uint8 RGB
 * [batch, 2, height, width, 3] (frame 1 then frame 2 of every pair, already cv2.resize'd to multiples of 64, :38-45);
 * the BGR swap, /255 (float32(double(u)/255.0)), planar layout (:47-56) happen on the device.  Same result, bit for bit, as
 * premvos_pwc_forward_host on the float tensor the reference builds; a quarter of its host->device bytes. */
int premvos_pwc_forward_host_u8(premvos_pwc_t* net, const unsigned char* frames_rgb_host, float* flow_host);
/* The same with DEVICE pointers (frames uint8 RGB [batch, 2, height, width, 3], flow fp32 NCHW); enqueues on `stream`,
 * does not synchronise.  One forward in flight per handle. */
int premvos_pwc_forward_u8(premvos_pwc_t* net, const unsigned char* frames_rgb_dev, float* flow_dev, void* stream);
/* Number of kernel launches one forward() performs (nodes of the captured graph). */
int premvos_pwc_launches_per_forward(const premvos_pwc_t* net);
/* Number of convolution layers of this handle that run on the tcgen05 tensor-core path (0 = pure
 * fp32 SIMT); valid after finalize. */
int premvos_pwc_tensor_core_layers(const premvos_pwc_t* net);
/* Set before finalize: 0 = fp32 SIMT convolutions everywhere; 1 (default) = tcgen05 tensor-core
 * implicit-GEMM convolutions (split-bf16 x3, fp32 accumulate) for every convolution of the network. */
int premvos_pwc_set_option(premvos_pwc_t* net, const char* key, int value);
/* Test hook: copy an intermediate of the LAST forward to a host fp32 NCHW buffer.  Names follow the
 * reference's variable names in PWCDCNet.forward: "c11".."c16", "c21".."c26", "corr6".."corr2",
 * "warp5".."warp2", "flow6".."flow2" (flow2 = before the context net), "up_flow6".."up_flow3",
 * "up_feat6".."up_feat3", "dc6".  *numel receives the element count; pass host_out = NULL to query. */
int premvos_pwc_get_tensor(premvos_pwc_t* net, const char* name, float* host_out, int64_t* numel);
void premvos_pwc_destroy(premvos_pwc_t* net);

/* ---------------------------------------------------------------------------------------------
 * Index-exact pieces of the proposal network as single ops (host pointers; parity hooks).
 * premvos_topk_host <- tf.nn.top_k(scores, k, sorted=False) (proposal_net/model.py:189-190): the k (<= 1024)
 *   largest of n scores; returned ordered by (score descending, index ascending).
 * premvos_nms_host  <- tf.image.non_max_suppression(boxes, scores, max_output_size, iou_threshold)
 *   (model.py:205-209, 466-467), n <= 1024 boxes [n,4] (either corner order): greedy by descending score, ties
 *   broken by the lower index, a box is dropped when its IoU with a kept box is > iou_threshold; IoU in fp32 with
 *   TensorFlow's operation order, so the selected INDICES are bit-exact with the oracle on identical inputs.
 * --------------------------------------------------------------------------------------------- */
int premvos_topk_host(const float* scores, int n, int k, int* indices_out, int* count_out);
int premvos_nms_host(const float* boxes, const float* scores, int n, float iou_threshold, int max_output_size,
                     int* selected_out, int* count_out);

/* ---------------------------------------------------------------------------------------------
 * Proposal network forward: ResNet-101 C4 Faster R-CNN, class-agnostic, with the second classification
 * head -- the graph `train.py --forward --agnostic --second_head` builds (proposal_net/train.py:107-189,
 * 274-295; basemodel.py:12-99; model.py:18-51, 114-139, 170-217, 301-395, 439-491, 552-565) and that
 * tensorpack's OfflinePredictor exposes as `pred_func(img)` (train.py:653-657, 508; eval.py:61-110).
 *
 * Life cycle: create(H, W of the ALREADY RESIZED image, num_class = 2, second_num_class = 81)
 *   [-> set_option("num_blocks0".."num_blocks3", count): ResNet depth, default 3,4,23,3 (config.py:61)]
 *   [-> set_option("batch", B), 1..16, default 1: B frames per forward.  The reference runs one frame per session.run
 *       (eval.py:61-110); a resident pipeline hands over several frames at once so that every convolution launch covers all of
 *       them (batch-1 ResNet layers have 23..184 work items for 148 SMs).  img is then [B, H, W, 3], copy_results hands over B
 *       images, read_results_image reads one; per-image results equal the batch-1 results up to fp32 summation order.]
 *   -> set_param(name, host fp32, numel) for every tensorpack variable, names and layouts as in a
 *      tensorpack checkpoint: "conv0/W" (HWIO), "conv0/bn/gamma|beta|mean/EMA|variance/EMA",
 *      "group{0..3}/block{i}/{conv1,conv2,conv3,convshortcut}/...", "rpn/conv0/{W,b}", "rpn/class/{W,b}",
 *      "rpn/box/{W,b}", "fastrcnn/{class,box}/{W,b}" (FullyConnected: [in,out]), "secondclassification/class/{W,b}"
 *   -> finalize() (folds the frozen BatchNorms, packs weights, allocates every buffer, plans every launch)
 *   -> forward*() -> destroy().
 *
 * img : fp32 [H, W, 3], BGR, 0..255, un-normalised (the graph normalises, basemodel.py:12-26).
 * Results = the six tensors of get_model_output_names() (train.py:52-62), at most 20 rows
 * (RESULTS_PER_IM, config.py:123): final_boxes [n,4] x1y1x2y2 in resized-image coordinates, final_probs [n],
 * final_labels [n] (int64, always 1 when class-agnostic), final_posterior [n,2], second_final_labels [n]
 * (int64), second_final_posterior [n,second_num_class].  Row order: ascending proposal index (the reference
 * leaves it unspecified: tf.nn.top_k(sorted=False), model.py:485-488).  The reference's quirk of gathering
 * label_probs with the CATEGORY index (train.py:287-288) is reproduced.
 *
 * premvos_propnet_forward       : img is a DEVICE pointer; enqueues on `stream`, does not synchronise;
 *                                 fetch with premvos_propnet_read_results (synchronises `stream`).
 * premvos_propnet_forward_host  : img and all outputs are HOST pointers; copies, runs, synchronises (batch 1 only).
 * Any output pointer except n_out may be NULL.  The whole forward is one CUDA graph captured at finalize
 * (set_option("cuda_graph", 0) before finalize turns that off); one forward in flight per handle.
 * --------------------------------------------------------------------------------------------- */
typedef struct premvos_propnet premvos_propnet_t;

int premvos_propnet_create(premvos_propnet_t** out, int height, int width, int num_class, int second_num_class);
int premvos_propnet_set_option(premvos_propnet_t* net, const char* key, int value);
int premvos_propnet_set_param(premvos_propnet_t* net, const char* name, const float* host_data, int64_t numel);
int premvos_propnet_finalize(premvos_propnet_t* net);
int premvos_propnet_forward(premvos_propnet_t* net, const float* img_dev, void* stream);
/* Same with the frame as the DEVICE uint8 BGR [H, W, 3] image eval.py:75-78 hands to pred_func (cv2.resize of a uint8
 * frame stays uint8); the uint8 -> fp32 conversion is exact, so results are bit-identical to premvos_propnet_forward. */
int premvos_propnet_forward_u8(premvos_propnet_t* net, const unsigned char* img_bgr_dev, void* stream);
int premvos_propnet_read_results(premvos_propnet_t* net, void* stream, int* n_out, float* final_boxes, float* final_probs,
                                 int64_t* final_labels, float* final_posterior, int64_t* second_final_labels,
                                 float* second_final_posterior);
/* read_results for image `image` (0 .. batch-1) of the last forward of a batched handle; read_results reads image 0. */
int premvos_propnet_read_results_image(premvos_propnet_t* net, void* stream, int image, int* n_out, float* final_boxes,
                                       float* final_probs, int64_t* final_labels, float* final_posterior,
                                       int64_t* second_final_labels, float* second_final_posterior);
/* Device-to-device hand-over of the last forward's results, enqueued on `stream` without synchronising: n_out_dev int[B],
 * final_boxes_dev fp32 [B,20,4], final_probs_dev fp32 [B,20], second_final_posterior_dev fp32 [B,20,second_num_class] with
 * B = the handle's batch (rows >= n_out_dev[b] are unspecified; any pointer except n_out_dev may be NULL). */
int premvos_propnet_copy_results(premvos_propnet_t* net, void* stream, int* n_out_dev, float* final_boxes_dev,
                                 float* final_probs_dev, float* second_final_posterior_dev);
int premvos_propnet_forward_host(premvos_propnet_t* net, const float* img_host, int* n_out, float* final_boxes,
                                 float* final_probs, int64_t* final_labels, float* final_posterior,
                                 int64_t* second_final_labels, float* second_final_posterior);
/* Mask R-CNN mask head (proposal_net/model.py:495-509, train.py:297-309; config.MODE_MASK -- inactive under simple_run.sh's
 * `--forward`).  set_option("mode_mask", 1) before the first set_param adds the variables
 * "maskrcnn/deconv/{W [2,2,256,2048], b [256]}" and "maskrcnn/conv/{W [1,1,256,1], b [1]}"; every forward then also runs
 * RoIAlign on the final boxes -> conv5 group -> Deconv2D 2x2 stride 2 + ReLU (one tcgen05 GEMM to the 4 x 256 phase channels) ->
 * 1x1 conv + sigmoid.  read_masks copies the first min(max_rows, 20) masks of image `image` (the `final_masks` output,
 * fp32 [rows, 14, 14], row i belongs to final_boxes[i]) to the host and synchronises `stream`. */
int premvos_propnet_read_masks(premvos_propnet_t* net, void* stream, int image, float* final_masks, int max_rows);
/* eval.py:35-58 `fill_full_mask(box, mask, shape)` for the n boxes of an image, on the device (host pointers, synchronous):
 * out[i] uint8 [height, width] = (cv2.resize(masks[i] (mask_size x mask_size fp32), (w, h)) > 0.5) pasted at the box's integer
 * rectangle x0 = int(x1 + 0.5) .. int(x2 - 0.5), zero elsewhere; boxes fp32 [n,4] x1y1x2y2 already clipped to the image.
 * OpenCV's float32 INTER_LINEAR arithmetic (INTER_AREA's 2x2 mean for an exact 2x down-scale, as cv2.resize substitutes). */
int premvos_fill_full_masks_host(const float* masks_host, const float* boxes_host, int n, int mask_size, int height, int width,
                                 unsigned char* out_host);
int premvos_propnet_launches_per_forward(const premvos_propnet_t* net);
/* Test hook: intermediates of the LAST forward as fp32 (CP8 activations are returned NCHW; index tensors are
 * converted to fp32).  Names: "conv0", "pool0", "block<i>", "featuremap", "rpn_hidden", "rpn_out" ([fh,fw,80]:
 * 15 logits + 60 deltas), "rpn_scores", "rpn_decoded_boxes", "topk_indices", "nms_keep", "proposal_boxes",
 * "proposal_scores", "roi_resized", "feature_fastrcnn", "pooled", "head_logits", "fastrcnn_all_probs",
 * "fastrcnn_all_boxes", "final_box_index", "cell_anchors".  Pass host_out = NULL to query *numel.  On a batched handle
 * activations hold all images ([B,C,H,W]); "<name>@<b>" selects image b for the per-image detection tensors (default 0). */
int premvos_propnet_get_tensor(premvos_propnet_t* net, const char* name, float* host_out, int64_t* numel);
void premvos_propnet_destroy(premvos_propnet_t* net);

/* ---------------------------------------------------------------------------------------------
 * Refinement network forward: DeepLabv3+ (Xception-65, output stride 16, ASPP 6/12/18, decoder stride 4) on box crops
 * with a 4th guidance channel -- the graph refinement_net/configs/{run,live} build (network/deeplab/DeepLabV3Plus.py:13-39,
 * deeplab/model.py:200-707, deeplab/core/xception.py:70-560, network/SegmentationOutputLayers.py:35-61,106-135) together
 * with its in-graph input pipeline (datasets/Dataset.py:141-186, datasets/Resize.py:150-193) and the per-proposal loop of
 * MergeTrack/refinement_net_functions.py:38-65 / forwarding/FewShotSegmentationForwarder.py:85-155, which calls
 * `engine.trainer.validation_step(feed_dict, extraction_keys)` once per proposal with the whole frame in the feed.
 *
 * Life cycle: create(max_batch proposals per launch group, input_size = 385 ("input_size_train", configs/run:27),
 *   middle_units = 16 (Xception-65; fewer only for tests)) -> set_param(name, host fp32, numel) for every slim variable,
 *   names and layouts as in the TF checkpoint: "xception_65/entry_flow/conv1_1/weights" (HWIO, 4 input channels),
 *   ".../BatchNorm/{gamma,beta,moving_mean,moving_variance}",
 *   "xception_65/<flow>/block<b>/unit_<u>/xception_module/separable_conv<i>_depthwise/depthwise_weights" ([3,3,C,1]),
 *   ".../separable_conv<i>_pointwise/weights", ".../shortcut/weights", "image_pooling/...", "aspp0/...",
 *   "aspp{1,2,3}_{depthwise,pointwise}/...", "concat_projection/...", "decoder/feature_projection0/...",
 *   "decoder/decoder_conv{0,1}_{depthwise,pointwise}/...", "logits/features/{weights,biases}"
 *   -> finalize() -> forward_host()* -> destroy().
 *
 * forward_host: frame = uint8 RGB [height, width, 3] (as scipy imread returns it; the reference divides by 255,
 *   FewShotFeedSegmentationDataset.py:37); boxes = [num_boxes, 4] fp32 x, y, w, h (proposal['bbox']).  Any number of
 *   boxes (processed max_batch at a time).  Outputs, all HOST pointers:
 *     masks        uint8 [num_boxes, height, width], 0/1 = SEGMENTATION_MASK_ORIGINAL_SIZE
 *     conf_scores  fp32  [num_boxes] = mean over the frame of 2*p' - 1, p' = p where mask else 1 - p
 *                  (refinement_net_functions.py:58-62)
 *     posteriors   fp32  [num_boxes, height, width] = SEGMENTATION_POSTERIORS_ORIGINAL_SIZE, or NULL to skip the copy
 * --------------------------------------------------------------------------------------------- */
typedef struct premvos_refnet premvos_refnet_t;

int premvos_refnet_create(premvos_refnet_t** out, int max_batch, int input_size, int middle_units);
int premvos_refnet_set_param(premvos_refnet_t* net, const char* name, const float* host_data, int64_t numel);
int premvos_refnet_finalize(premvos_refnet_t* net);
/* forward: the same with every pointer a DEVICE pointer (frame uint8 RGB, boxes fp32 xywh, masks uint8 [num_boxes, height,
 *   width], conf_scores fp32 [num_boxes], posteriors fp32 or NULL); enqueues on `stream`, never synchronises -- the entry
 *   point of a resident pipeline (frames decoded once, proposals handed over on the device).  One forward in flight per
 *   handle.  A full launch group (max_batch proposals) replays a CUDA graph captured at finalize. */
int premvos_refnet_forward(premvos_refnet_t* net, const unsigned char* frame_rgb_dev, int height, int width,
                           const float* boxes_xywh_dev, int num_boxes, unsigned char* masks_dev, float* conf_scores_dev,
                           float* posteriors_dev, void* stream);
int premvos_refnet_forward_host(premvos_refnet_t* net, const unsigned char* frame_rgb, int height, int width,
                                const float* boxes_xywh, int num_boxes, unsigned char* masks_out, float* conf_scores_out,
                                float* posteriors_out);
int premvos_refnet_launches_per_forward(const premvos_refnet_t* net);
/* Test hook (state of the LAST batch): "net_input", "xception_out", "low_level", "aspp_concat", "aspp_out", "decoder_in",
 * "decoder_out" (NCHW), "logits" ([max_batch,h,w,16] channels-last, classes in channels 0..1), "crops" ([max_batch,4]). */
int premvos_refnet_get_tensor(premvos_refnet_t* net, const char* name, float* host_out, int64_t* numel);
void premvos_refnet_destroy(premvos_refnet_t* net);

/* ---------------------------------------------------------------------------------------------
 * ReID network forward: the 128-d appearance embedding MergeTrack attaches to every proposal
 * (MergeTrack/ReID_net_functions.py:26-45 `add_ReID`: `engine.forward(net, [{"boxes", "image_paths"}])["ys"]`), i.e. the graph
 * ReID_net/configs/run:33-65 builds -- conv0, 17 pre-activation ResidualUnit2 (network/NetworkLayers.py:157-210), conv1 +
 * 3x3/3 max pool, fc1, fc2, outputTriplet (FullyConnected, NetworkLayers.py:536-567) -- together with the in-graph crop
 * pipeline of ReID_net/datasets/Similarity/DAVIS_Forward_Feed.py:34-120 (context region x1.2, tf.round, clipping, 128 x 128
 * legacy bilinear resize, zeros for crops with a side <= 10 px, ImageNet normalisation).
 *
 * Life cycle: create(max_batch crops per launch group, <= 256 = configs/run "batch_size") -> set_param(name, host fp32,
 *   numel) for every variable, names and layouts as in the TF checkpoint: "conv0/W" (HWIO), "res<k>/W0" (1x1 projection,
 *   only where the unit changes shape), "res<k>/W<i>", "res<k>/bn<j>/{beta,gamma,mean_ema,var_ema}" (bn0 in front of the
 *   unit, bn<i> in front of W<i> for i >= 2), "conv1/{W,bn/...}", "{fc1,fc2,outputTriplet}/{W,b,bn/...}"
 *   -> finalize() -> forward_host()* -> destroy().
 *
 * forward_host: frame = uint8 RGB [height, width, 3]; boxes = [num_boxes, 4] fp32 x, y, w, h (proposal['bbox']); any
 *   number of boxes.  Output (HOST): embeddings fp32 [num_boxes, 128] = "ys".
 * forward: the same with DEVICE pointers, enqueued on `stream`, never synchronises.
 * get_tensor (test hook, state of the LAST batch): "net_input" (NCHW [max_batch,8,128,128], channels 0..2 used), "conv0",
 *   "res0".."res16" (NCHW raw unit outputs), "conv1" ([max_batch,4,4,512] channels-last, 500 used, before the max pool),
 *   "crops" ([max_batch,4] x y w h after context region + clipping).
 * --------------------------------------------------------------------------------------------- */
typedef struct premvos_reidnet premvos_reidnet_t;

int premvos_reidnet_create(premvos_reidnet_t** out, int max_batch);
int premvos_reidnet_set_param(premvos_reidnet_t* net, const char* name, const float* host_data, int64_t numel);
int premvos_reidnet_finalize(premvos_reidnet_t* net);
int premvos_reidnet_forward(premvos_reidnet_t* net, const unsigned char* frame_rgb_dev, int height, int width,
                            const float* boxes_xywh_dev, int num_boxes, float* embeddings_dev, void* stream);
int premvos_reidnet_forward_host(premvos_reidnet_t* net, const unsigned char* frame_rgb, int height, int width,
                                 const float* boxes_xywh, int num_boxes, float* embeddings_out);
int premvos_reidnet_launches_per_forward(const premvos_reidnet_t* net);
int premvos_reidnet_get_tensor(premvos_reidnet_t* net, const char* name, float* host_out, int64_t* numel);
void premvos_reidnet_destroy(premvos_reidnet_t* net);

#ifdef __cplusplus
}
#endif
#endif /* PREMVOS_B200_H */
