/* Plain-C CPU restatement of the reference correlation forward -- TEST INFRASTRUCTURE ONLY.
 *
 * Follows (paths relative to /root/reference/code/optical_flow_net-PWC-Net/external_packages/
 * correlation-pytorch-master/correlation-pytorch/correlation_package/src/):
 *   corr_cuda.c:23-45            output/padded shape math
 *   corr_cuda.c:52-60            zero-filled output and zero-padded NHWC scratch ("rbot")
 *   corr_cuda_kernel.cu:18-37    blob_rearrange: NCHW -> padded NHWC
 *   corr_cuda_kernel.cu:59-127   CorrelateData: per output pixel, per displacement channel,
 *                                sum over kernel window and channels, divide by k*k*C
 * The reference's own CPU entry point (corr.c:3-16) is a stub that returns 1 without touching
 * the output, so this file restates the CUDA kernel's arithmetic in scalar C instead.
 * Summation order: the GPU kernel accumulates 32 per-lane partial sums (channel c goes to lane
 * c%32) and then adds the 32 partials serially (corr_cuda_kernel.cu:107-118); this file does the
 * same so that single-threaded results are as close as fp32 allows to the recompiled kernel.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int corr_oracle_output_shape(int H, int W, int pad_size, int kernel_size, int max_displacement,
                             int stride1, int stride2, int *out_c, int *out_h, int *out_w) {
  int kr = (kernel_size - 1) / 2;
  int border = max_displacement + kr;
  int ph = H + 2 * pad_size, pw = W + 2 * pad_size;
  *out_w = (int)ceilf((float)(pw - border * 2) / (float)stride1);
  *out_h = (int)ceilf((float)(ph - border * 2) / (float)stride1);
  int gr = max_displacement / stride2;
  int gw = gr * 2 + 1;
  *out_c = gw * gw;
  return 0;
}

/* in1,in2: [B,C,H,W] fp32; out: [B, gw*gw, OH, OW] fp32.  Returns 0 on success. */
int corr_oracle_forward(const float *in1, const float *in2, float *out, int B, int C, int H, int W,
                        int pad_size, int kernel_size, int max_displacement, int stride1,
                        int stride2) {
  int oc, oh, ow;
  corr_oracle_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2, &oc,
                           &oh, &ow);
  if (oh <= 0 || ow <= 0) return -1;
  const int gr = max_displacement / stride2;
  const int gw = 2 * gr + 1;
  /* guard band so that pad_size < max_displacement reads zeros instead of faulting */
  const int extra = (max_displacement > pad_size ? max_displacement - pad_size : 0) + kernel_size;
  const int ph = H + 2 * pad_size + 2 * extra, pw = W + 2 * pad_size + 2 * extra;
  size_t pbytes = (size_t)B * ph * pw * C * sizeof(float);
  float *r1 = (float *)calloc(1, pbytes), *r2 = (float *)calloc(1, pbytes);
  if (!r1 || !r2) { free(r1); free(r2); return -2; }
  for (int n = 0; n < B; n++)
    for (int c = 0; c < C; c++)
      for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
          size_t src = (((size_t)n * C + c) * H + y) * W + x;
          size_t dst = (((size_t)n * ph + (y + pad_size + extra)) * pw + (x + pad_size + extra)) * C + c;
          r1[dst] = in1[src];
          r2[dst] = in2[src];
        }
  const float sumelems = (float)(kernel_size * kernel_size * C);
  for (int n = 0; n < B; n++)
    for (int by = 0; by < oh; by++)
      for (int bx = 0; bx < ow; bx++) {
        int x1 = bx * stride1 + max_displacement + extra;
        int y1 = by * stride1 + max_displacement + extra;
        for (int tc = 0; tc < oc; tc++) {
          int s2o = (tc % gw - gr) * stride2;
          int s2p = (tc / gw - gr) * stride2;
          float lane[32];
          memset(lane, 0, sizeof(lane));
          for (int j = 0; j < kernel_size; j++)
            for (int i = 0; i < kernel_size; i++) {
              const float *a = r1 + (((size_t)n * ph + (y1 + j)) * pw + (x1 + i)) * C;
              const float *b = r2 + (((size_t)n * ph + (y1 + s2p + j)) * pw + (x1 + s2o + i)) * C;
              for (int c = 0; c < C; c++) lane[c & 31] += a[c] * b[c];
            }
          float total = 0.f;
          for (int l = 0; l < 32; l++) total += lane[l];
          out[(((size_t)n * oc + tc) * oh + by) * ow + bx] = total / sumelems;
        }
      }
  free(r1);
  free(r2);
  return 0;
}
