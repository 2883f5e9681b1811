// Bandwidth-bound companions of the tensor-core convolutions, all on CP8 activations
// ([N][C/8][H][W][8] split-bf16 planes, see CView in common.cuh): input packing, the PWC warping layer,
// the 9x9 cost volume, the flow/feature up-sampling glue and the NCHW <-> CP8 converters of the C ABI.
// One (pixel, chunk) element is 16 bytes per plane, so "one thread per (pixel, chunk), x fastest" gives
// 512-byte coalesced warp transactions everywhere.
#include "common.cuh"

namespace premvos {

namespace {

struct F8 { float v[8]; };

__device__ __forceinline__ F8 ld_chunk(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long elem /* multiple of 8 */) {
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

struct CV {  // device copy of a CView
  __nv_bfloat16* hi; __nv_bfloat16* lo; int N, H, W, chunks, c0, C;
};
CV dev(const CView& v) { return CV{v.hi, v.lo, v.N, v.H, v.W, v.chunks, v.c0, v.C}; }
__device__ __forceinline__ long cv_elem(const CV& v, int n, int chunk, int y, int x) {
  return ((((long)n * v.chunks + v.c0 + chunk) * v.H + y) * v.W + x) * 8;
}

// ---- input packing ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_pair_cp8_kernel(const float* __restrict__ x, CV img, int B) {
  const long hw = (long)img.H * img.W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2L * B * hw) return;
  const int n = (int)(idx / hw);
  const long p = idx - (long)n * hw;
  const int b = n % B, im = n / B;
  const float* src = x + ((long)b * 6 + im * 3) * hw + p;
  F8 f;
#pragma unroll
  for (int j = 0; j < 8; j++) f.v[j] = 0.f;
  f.v[0] = src[0]; f.v[1] = src[hw]; f.v[2] = src[2 * hw];
  st_chunk(img.hi, img.lo, (((long)n * img.chunks + img.c0) * hw + p) * 8, f);
}

// im2col of the FIRST pyramid convolution (3x3, stride 2, pad 1, 3 input channels: PWCNet.py:50): the 27 inputs of every
// output pixel are gathered here so that conv1a becomes a 1x1 GEMM with K = 27 (4 chunk planes, channel (r*3+s)*3 + c).
// Same bytes written as a plain 8-channel full-resolution image, but the convolution reads them once instead of through
// nine strided TMA boxes.
__global__ void __launch_bounds__(256) pack_pair_im2col_kernel(const float* __restrict__ x, CV img, int B, int H, int W) {
  const int Ho = img.H, Wo = img.W;
  const long total = 2L * B * 4 * Ho * Wo;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho);
  const int ch = (int)((idx / ((long)Wo * Ho)) % 4), n = (int)(idx / ((long)Wo * Ho * 4));
  const int b = n % B, im = n / B;
  const float* src = x + ((long)b * 6 + im * 3) * H * W;
  F8 f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = ch * 8 + j;          // (r*3+s)*3 + c
    float v = 0.f;
    if (k < 27) {
      const int c = k % 3, tap = k / 3, r = tap / 3, s = tap % 3;
      const int iy = 2 * oy + r - 1, ix = 2 * ox + s - 1;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = src[(long)c * H * W + (long)iy * W + ix];
    }
    f.v[j] = v;
  }
  st_chunk(img.hi, img.lo, cv_elem(img, n, ch, oy, ox), f);
}

// uint8 RGB frames [B][2][H][W][3] -> the network input the reference builds on the host (script_pwc_multi.py:47-56):
// fp32 NCHW [B][6][H][W], BGR, float32(double(u) / 255.0)
__global__ void __launch_bounds__(256) frames_u8_to_x_kernel(const unsigned char* __restrict__ fr, float* __restrict__ x, int B, int H, int W) {
  const long hw = (long)H * W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * 2 * hw) return;
  const long p = idx % hw;
  const int im = (int)((idx / hw) % 2), b = (int)(idx / (2 * hw));
  const unsigned char* s = fr + (((long)b * 2 + im) * hw + p) * 3;
  float* d = x + ((long)b * 6 + im * 3) * hw + p;
#pragma unroll
  for (int c = 0; c < 3; c++) d[(long)c * hw] = (float)((double)s[2 - c] / 255.0);
}

// ---- PWCDCNet.warp (PWCNet.py:140-176) -------------------------------------------------------------
struct WarpCp8Args { CV x, flow, out; int flow_ch; float scale; };

__global__ void __launch_bounds__(256) warp_cp8_kernel(WarpCp8Args a) {
  const int W = a.x.W, H = a.x.H;
  const int nch = (a.x.C + 7) / 8;
  const long total = (long)a.x.N * nch * H * W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int ch = (int)((idx / ((long)W * H)) % nch);
  const int n = (int)(idx / ((long)W * H * nch));
  const F8 fl = ld_chunk(a.flow.hi, a.flow.lo, cv_elem(a.flow, n, 0, y, x));
  float fu = fl.v[0], fv = fl.v[1];
#pragma unroll
  for (int j = 2; j < 8; j += 2)
    if (a.flow_ch == j) { fu = fl.v[j]; fv = fl.v[j + 1]; }
  const float u = __fmul_rn(fu, a.scale);
  const float v = __fmul_rn(fv, a.scale);
  // PWCNet.py:157-162 then grid_sample's un-normalisation ((g+1)/2)*(size-1)   (align_corners=True semantics)
  const float wm1 = (float)max(W - 1, 1), hm1 = (float)max(H - 1, 1);
  float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)x, u)), wm1), 1.0f);
  float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)y, v)), hm1), 1.0f);
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
  float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.0f) - ix, wy0 = (fy0 + 1.0f) - iy;
  fx0 = fminf(fmaxf(fx0, -2.0f), (float)W + 1.0f);
  fy0 = fminf(fmaxf(fy0, -2.0f), (float)H + 1.0f);
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
  const float w00 = (vy0 && vx0) ? wy0 * wx0 : 0.f;
  const float w01 = (vy0 && vx1) ? wy0 * wx1 : 0.f;
  const float w10 = (vy1 && vx0) ? wy1 * wx0 : 0.f;
  const float w11 = (vy1 && vx1) ? wy1 * wx1 : 0.f;
  const float msum = ((w00 + w01) + w10) + w11;   // bilinear sample of the all-ones image (:166-167)
  const bool valid = msum >= 0.9999f;             // mask[mask<0.9999] = 0 (:173-174)
  F8 r;
#pragma unroll
  for (int j = 0; j < 8; j++) r.v[j] = 0.f;
  if (valid) {
    const F8 t00 = ld_chunk(a.x.hi, a.x.lo, cv_elem(a.x, n, ch, vy0 ? y0 : 0, vx0 ? x0 : 0));
    const F8 t01 = ld_chunk(a.x.hi, a.x.lo, cv_elem(a.x, n, ch, vy0 ? y0 : 0, vx1 ? x1 : 0));
    const F8 t10 = ld_chunk(a.x.hi, a.x.lo, cv_elem(a.x, n, ch, vy1 ? y1 : 0, vx0 ? x0 : 0));
    const F8 t11 = ld_chunk(a.x.hi, a.x.lo, cv_elem(a.x, n, ch, vy1 ? y1 : 0, vx1 ? x1 : 0));
#pragma unroll
    for (int j = 0; j < 8; j++) r.v[j] = ((t00.v[j] * w00 + t01.v[j] * w01) + t10.v[j] * w10) + t11.v[j] * w11;
  }
  st_chunk(a.out.hi, a.out.lo, cv_elem(a.out, n, ch, y, x), r);
}

// ---- 9x9 cost volume (corr_cuda_kernel.cu:59-127 semantics, pad 4, k 1, md 4, strides 1) ---------------
// One CTA owns a 4x32 pixel tile; per 32-channel group the f1 tile and the f2 tile + 4-pixel halo are
// staged in shared memory as fp32 (pitch 33 floats: 32 lanes = 32 consecutive pixels hit 32 banks).
// 288 threads = 32 pixel columns x 9 x-displacements; a thread owns its column's 4 pixels x all 9 y-displacements
// (36 accumulators): per channel 4 f1 + 12 f2 shared-memory reads feed 36 FMAs (register reuse along y), and the
// 32 lanes of a warp read 32 consecutive pixels = 32 different banks.
constexpr int CT_H = 4, CT_W = 32, MD = 4, DW = 9;
constexpr int HALO_H = CT_H + 2 * MD, HALO_W = CT_W + 2 * MD;  // 12 x 40
constexpr int CK = 32, PITCH = CK + 1;
constexpr int CORR_THREADS = 288;
constexpr size_t CORR_SMEM = (size_t)(HALO_H * HALO_W + CT_H * CT_W) * PITCH * sizeof(float);

struct CorrCp8Args { CV f1, f2, out, c1; int has_c1; float slope; };

__global__ void __launch_bounds__(CORR_THREADS) corr81_cp8_kernel(CorrCp8Args a) {
  extern __shared__ float smem[];
  float* s2 = smem;                            // [HALO_H*HALO_W][PITCH]
  float* s1 = smem + HALO_H * HALO_W * PITCH;  // [CT_H*CT_W][PITCH]
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * CT_H, x0 = blockIdx.x * CT_W;
  const int px = tid & 31, dxi = tid >> 5;   // pixel column, x-displacement index (dx = dxi - 4)
  const int H = a.f1.H, W = a.f1.W, C = a.f1.C;
  const int nchunks = (C + 7) / 8;

  float acc[CT_H][DW];   // [pixel row][y-displacement]
#pragma unroll
  for (int i = 0; i < CT_H; i++)
#pragma unroll
    for (int j = 0; j < DW; j++) acc[i][j] = 0.f;

  for (int cg = 0; cg < nchunks; cg += 4) {  // 4 chunks = 32 channels per pass
    for (int i = tid; i < 4 * HALO_H * HALO_W; i += CORR_THREADS) {
      const int q = i / (HALO_H * HALO_W), hp = i - q * (HALO_H * HALO_W);
      const int hy = hp / HALO_W, hx = hp - hy * HALO_W;
      const int gy = y0 + hy - MD, gx = x0 + hx - MD;
      F8 f;
#pragma unroll
      for (int j = 0; j < 8; j++) f.v[j] = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W && cg + q < nchunks) f = ld_chunk(a.f2.hi, a.f2.lo, cv_elem(a.f2, n, cg + q, gy, gx));
      float* d = s2 + hp * PITCH + q * 8;
#pragma unroll
      for (int j = 0; j < 8; j++) d[j] = f.v[j];
    }
    for (int i = tid; i < 4 * CT_H * CT_W; i += CORR_THREADS) {
      const int q = i >> 7, tp = i & 127;
      const int gy = y0 + (tp >> 5), gx = x0 + (tp & 31);
      F8 f;
#pragma unroll
      for (int j = 0; j < 8; j++) f.v[j] = 0.f;
      if (gy < H && gx < W && cg + q < nchunks) {
        const long e = cv_elem(a.f1, n, cg + q, gy, gx);
        f = ld_chunk(a.f1.hi, a.f1.lo, e);
        if (a.has_c1) {  // fused copy of f1 (raw hi/lo bits) into the decoder slab
          const long o = cv_elem(a.c1, n, cg + q, gy, gx);
          *reinterpret_cast<uint4*>(a.c1.hi + o) = *reinterpret_cast<const uint4*>(a.f1.hi + e);
          *reinterpret_cast<uint4*>(a.c1.lo + o) = *reinterpret_cast<const uint4*>(a.f1.lo + e);
        }
      }
      float* d = s1 + tp * PITCH + q * 8;
#pragma unroll
      for (int j = 0; j < 8; j++) d[j] = f.v[j];
    }
    __syncthreads();
    const float* q1 = s1 + px * PITCH;                 // f1 pixel (row r, column px) at q1[r * CT_W * PITCH]
    const float* q2 = s2 + (px + dxi) * PITCH;         // f2 halo pixel (row r, column px + dxi) at q2[r * HALO_W * PITCH]
#pragma unroll 2
    for (int c = 0; c < CK; c++) {
      float f1v[CT_H], f2v[HALO_H];
#pragma unroll
      for (int r = 0; r < CT_H; r++) f1v[r] = q1[r * CT_W * PITCH + c];
#pragma unroll
      for (int r = 0; r < HALO_H; r++) f2v[r] = q2[r * HALO_W * PITCH + c];
#pragma unroll
      for (int r = 0; r < CT_H; r++)
#pragma unroll
        for (int dy = 0; dy < DW; dy++) acc[r][dy] = fmaf(f1v[r], f2v[r + dy], acc[r][dy]);
    }
    __syncthreads();
  }
  // stage the 128 x 81 results (pitch 89 floats: 88 = 11 chunks + 1 to spread banks), then write chunk planes
  constexpr int OP = 89;
  float* so = smem;
#pragma unroll
  for (int r = 0; r < CT_H; r++)
#pragma unroll
    for (int dy = 0; dy < DW; dy++) {
      float v = acc[r][dy] / (float)C;  // corr_cuda_kernel.cu:119-121
      v = v > 0.f ? v : v * a.slope;
      so[(r * CT_W + px) * OP + dy * DW + dxi] = v;   // output channel = (dy+4)*9 + (dx+4), corr_cuda_kernel.cu:94-95
    }
  if (dxi < 7) {
#pragma unroll
    for (int r = 0; r < CT_H; r++) so[(r * CT_W + px) * OP + 81 + dxi] = 0.f;
  }
  __syncthreads();
  for (int i = tid; i < 11 * CT_H * CT_W; i += CORR_THREADS) {
    const int q = i >> 7, tp = i & 127;
    const int yy = y0 + (tp >> 5), xx = x0 + (tp & 31);
    if (yy < H && xx < W) {
      F8 f;
#pragma unroll
      for (int j = 0; j < 8; j++) f.v[j] = so[tp * OP + q * 8 + j];
      st_chunk(a.out.hi, a.out.lo, cv_elem(a.out, n, q, yy, xx), f);
    }
  }
}

// ---- level glue ------------------------------------------------------------------------------------
struct LevelUpArgs {
  const float* head; int cs; int N, H, W;  // head fp32 channels-last [N,H,W,cs]
  const float* dw; const float* db;        // deconvL weights [2][2][4][4] (ci, co, ky, kx), bias [2]
  CV dst;                                  // chunk plane of the next slab at [N,2H,2W]
};

__global__ void __launch_bounds__(256) level_up_kernel(LevelUpArgs a) {
  const int Ho = 2 * a.H, Wo = 2 * a.W;
  const long total = (long)a.N * Ho * Wo;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho), n = (int)(idx / ((long)Wo * Ho));
  const int y = oy >> 1, x = ox >> 1, py = oy & 1, px = ox & 1;
  // ConvTranspose2d(2, 2, 4, stride 2, pad 1) of the flow (PWCNet.py:83): oy = 2*iy - 1 + ky
  float up0 = a.db[0], up1 = a.db[1];
#pragma unroll
  for (int a_ = 0; a_ < 2; a_++) {
    const int ky = ((oy + 1) & 1) + 2 * a_;
    const int iy = (oy + 1 - ky) / 2;
    if (oy + 1 - ky < 0 || iy >= a.H) continue;
#pragma unroll
    for (int b_ = 0; b_ < 2; b_++) {
      const int kx = ((ox + 1) & 1) + 2 * b_;
      const int ix = (ox + 1 - kx) / 2;
      if (ox + 1 - kx < 0 || ix >= a.W) continue;
      const float* fp = a.head + (((long)n * a.H + iy) * a.W + ix) * a.cs;
      const float f0 = fp[0], f1 = fp[1];
      const int t = ky * 4 + kx;
      up0 = fmaf(f0, a.dw[(0 * 2 + 0) * 16 + t], up0);
      up0 = fmaf(f1, a.dw[(1 * 2 + 0) * 16 + t], up0);
      up1 = fmaf(f0, a.dw[(0 * 2 + 1) * 16 + t], up1);
      up1 = fmaf(f1, a.dw[(1 * 2 + 1) * 16 + t], up1);
    }
  }
  const float* hp = a.head + (((long)n * a.H + y) * a.W + x) * a.cs + 2 + (py * 2 + px) * 2;
  F8 f;
  f.v[0] = up0; f.v[1] = up1; f.v[2] = hp[0]; f.v[3] = hp[1];
  f.v[4] = f.v[5] = f.v[6] = f.v[7] = 0.f;
  st_chunk(a.dst.hi, a.dst.lo, cv_elem(a.dst, n, 0, oy, ox), f);
}

__global__ void __launch_bounds__(256) flow_finish_kernel(const float* a, int a_cs, const float* b, int b_cs, float* out, int N, int H, int W) {
  const long hw = (long)H * W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)N * 2 * hw) return;
  const long p = idx % hw;
  const int c = (int)((idx / hw) % 2), n = (int)(idx / (2 * hw));
  out[idx] = a[((long)n * hw + p) * a_cs + c] + b[((long)n * hw + p) * b_cs + c];
}

// ---- NCHW <-> CP8 -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_cp8_kernel(const float* __restrict__ src, CV v) {
  const int nch = (v.C + 7) / 8;
  const long hw = (long)v.H * v.W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)v.N * nch * hw) return;
  const long p = idx % hw;
  const int ch = (int)((idx / hw) % nch), n = (int)(idx / (hw * nch));
  F8 f;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = ch * 8 + j;
    f.v[j] = c < v.C ? src[((long)n * v.C + c) * hw + p] : 0.f;
  }
  st_chunk(v.hi, v.lo, (((long)n * v.chunks + v.c0 + ch) * hw + p) * 8, f);
}

__global__ void __launch_bounds__(256) cp8_to_nchw_kernel(CV v, int ch_off, float* __restrict__ dst) {
  const long hw = (long)v.H * v.W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)v.N * v.C * hw) return;
  const long p = idx % hw;
  const int c = (int)((idx / hw) % v.C), n = (int)(idx / (hw * v.C));
  const int pc = c + ch_off;
  const long e = (((long)n * v.chunks + v.c0 + (pc >> 3)) * hw + p) * 8 + (pc & 7);
  dst[idx] = __bfloat162float(v.hi[e]) + __bfloat162float(v.lo[e]);
}

inline unsigned blocks_for(long total) { return (unsigned)((total + 255) / 256); }

}  // namespace

int pack_pair_input_cp8(const float* x_nchw, int B, int H, int W, const CView& img, cudaStream_t st) {
  PV_CHECK(img.N == 2 * B && img.H == H && img.W == W, PREMVOS_ERR_INVALID_ARG, "pack_pair_input_cp8: shape mismatch");
  const long total = 2L * B * H * W;
  prof_before(st);
  pack_pair_cp8_kernel<<<blocks_for(total), 256, 0, st>>>(x_nchw, dev(img), B);
  return after_launch("pack_pair_cp8_kernel", st, 0.0, (double)total * (12.0 + 32.0));
}

int pack_pair_im2col_cp8(const float* x_nchw, int B, int H, int W, const CView& img, cudaStream_t st) {
  PV_CHECK(img.N == 2 * B && img.H == H / 2 && img.W == W / 2 && img.C == 32, PREMVOS_ERR_INVALID_ARG, "pack_pair_im2col_cp8: shape mismatch");
  const long total = 2L * B * 4 * img.H * img.W;
  prof_before(st);
  pack_pair_im2col_kernel<<<blocks_for(total), 256, 0, st>>>(x_nchw, dev(img), B, H, W);
  return after_launch("pack_pair_im2col_kernel", st, 0.0, (double)B * 6 * H * W * 4 + (double)total * 32.0);
}

int frames_u8_to_x(const unsigned char* frames_rgb, float* x_nchw, int B, int H, int W, cudaStream_t st) {
  const long total = (long)B * 2 * H * W;
  prof_before(st);
  frames_u8_to_x_kernel<<<blocks_for(total), 256, 0, st>>>(frames_rgb, x_nchw, B, H, W);
  return after_launch("frames_u8_to_x_kernel", st, 0.0, (double)total * 15.0);
}

int warp_cp8(const CView& x2, const CView& flow, int flow_ch, float flow_scale, const CView& out, cudaStream_t st) {
  PV_CHECK(x2.N == out.N && x2.H == out.H && x2.W == out.W && x2.C == out.C && flow.H == x2.H && flow.W == x2.W &&
               flow.N == x2.N && flow_ch >= 0 && flow_ch <= 6 && (flow_ch & 1) == 0, PREMVOS_ERR_INVALID_ARG, "warp_cp8: shape mismatch");
  WarpCp8Args a{dev(x2), dev(flow), dev(out), flow_ch, flow_scale};
  const long total = (long)x2.N * x2.vchunks() * x2.H * x2.W;
  prof_before(st);
  warp_cp8_kernel<<<blocks_for(total), 256, 0, st>>>(a);
  const double P = (double)x2.pixels();
  return after_launch("warp_cp8_kernel", st, 8.0 * P * x2.C, 4.0 * P * (2.0 * x2.C + 2));
}

int corr81_cp8(const CView& f1, const CView& f2, const CView& out, const CView& c1_copy, float slope, cudaStream_t st) {
  PV_CHECK(f1.C == f2.C && f1.H == f2.H && f1.W == f2.W && f1.N == f2.N && out.C == 81 && out.H == f1.H && out.W == f1.W &&
               out.N == f1.N, PREMVOS_ERR_INVALID_ARG, "corr81_cp8: shape mismatch");
  PV_CHECK(c1_copy.null() || (c1_copy.C == f1.C && c1_copy.H == f1.H && c1_copy.W == f1.W && c1_copy.N == f1.N),
           PREMVOS_ERR_INVALID_ARG, "corr81_cp8: c1 copy shape mismatch");
  {   // TMA-staged alternative (corr_tma.cu, opt-in: measured slower in CP8, see there)
    const int r = corr81_cp8_tma(f1, f2, out, c1_copy, slope, st);
    if (r != 1) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    PV_CUDA(cudaFuncSetAttribute(corr81_cp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CORR_SMEM));
    attr_set = true;
  }
  CorrCp8Args a{dev(f1), dev(f2), dev(out), dev(c1_copy), c1_copy.null() ? 0 : 1, slope};
  dim3 grid((f1.W + CT_W - 1) / CT_W, (f1.H + CT_H - 1) / CT_H, f1.N);
  const double px = (double)f1.pixels();
  prof_before(st);
  corr81_cp8_kernel<<<grid, CORR_THREADS, CORR_SMEM, st>>>(a);
  return after_launch("corr81_cp8_kernel", st, 2.0 * 81 * f1.C * px, 4.0 * (2.0 * f1.C + 81) * px);
}

int level_up_cp8(const TView& head, const float* deconv_w, const float* deconv_b, const CView& dst, cudaStream_t st) {
  PV_CHECK(head.p && head.cs >= 10 && dst.N == head.N && dst.H == 2 * head.H && dst.W == 2 * head.W, PREMVOS_ERR_INVALID_ARG,
           "level_up_cp8: shape mismatch");
  LevelUpArgs a{head.p, head.cs, head.N, head.H, head.W, deconv_w, deconv_b, dev(dst)};
  const long total = (long)dst.N * dst.H * dst.W;
  prof_before(st);
  level_up_kernel<<<blocks_for(total), 256, 0, st>>>(a);
  return after_launch("level_up_kernel", st, 2.0 * total * 16, (double)total * (10.0 + 32.0));
}

int flow_finish(const TView& a, const TView& b, float* out_nchw, cudaStream_t st) {
  PV_CHECK(a.p && b.p && a.N == b.N && a.H == b.H && a.W == b.W, PREMVOS_ERR_INVALID_ARG, "flow_finish: shape mismatch");
  const long total = (long)a.N * 2 * a.H * a.W;
  prof_before(st);
  flow_finish_kernel<<<blocks_for(total), 256, 0, st>>>(a.p + a.coff, a.cs, b.p + b.coff, b.cs, out_nchw, a.N, a.H, a.W);
  return after_launch("flow_finish_kernel", st, (double)total, 12.0 * total);
}

int nchw_to_cp8(const float* src, const CView& dst, cudaStream_t st) {
  const long total = (long)dst.N * dst.vchunks() * dst.H * dst.W;
  prof_before(st);
  nchw_to_cp8_kernel<<<blocks_for(total), 256, 0, st>>>(src, dev(dst));
  return after_launch("nchw_to_cp8_kernel", st, 0.0, 8.0 * (double)dst.pixels() * dst.C);
}

int cp8_to_nchw(const CView& src, int ch_off, float* dst, cudaStream_t st) {
  const long total = (long)src.N * src.C * src.H * src.W;
  prof_before(st);
  cp8_to_nchw_kernel<<<blocks_for(total), 256, 0, st>>>(dev(src), ch_off, dst);
  return after_launch("cp8_to_nchw_kernel", st, 0.0, 8.0 * (double)total);
}

}  // namespace premvos
