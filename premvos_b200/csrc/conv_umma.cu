// Tensor-core convolution for sm_100a: implicit GEMM on tcgen05.mma, operands staged in shared memory
// by TMA, accumulators in tensor memory (TMEM).
//
// Activation layout ("CP8", CView in common.cuh): split-bf16 planes (x = hi + lo) stored channel-chunk
// planar, [N][C/8][H][W][8].  One 8-channel chunk of one pixel is 16 bytes -- exactly one row of a UMMA
// "core matrix" in the un-swizzled K-major operand layout, so a TMA box {8 ch, BW px, BH px, KC chunks}
// lands in shared memory as [KC][BH][BW][8ch] and IS a valid A operand with
//      SBO (next 8 pixels = next tile row) = BW*16 B,    LBO (next 8 channels) = BH*BW*16 B.
// im2col never materialises, and for stride-1 convolutions not even per tap: the box is the output
// tile (16*MT rows x 8 cols) PLUS its halo, loaded ONCE per 16-channel k-block; the R*S taps are R*S
// descriptors whose start address is shifted by ((r*dil)*BW + s*dil)*16 B.  Shared-memory fill traffic
// for A drops from 9x to ~1.4x of the tile ("halo mode").  Strided / strongly dilated / 1x1
// convolutions use one box per (tap, k-block) ("tap mode"; stride 2 = TMA elementStrides).
// Out-of-image pixels and channels beyond the view are zero-filled by the TMA unit = zero padding.
//
// Weights are pre-packed on the host into the exact shared-memory image of every (n-tile, tap,
// k-block) stage, [KC][BN][8ch] (SBO = 128 B, LBO = BN*16 B), and fetched with 1-D bulk copies.
//
// Numerics: x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits).  Each K step issues
// three kind::f16 MMAs into the same fp32 TMEM accumulator: lo*hi + hi*lo + hi*hi (the dropped lo*lo
// term is ~2^-16 relative).  1e-3 parity with a wide margin; fp32 accumulate.
//
// CTA = 256 threads: warp 0 lane 0 TMA producer, warp 1 lane 0 MMA issuer, warp 2 TMEM allocator,
// warps 4..7 epilogue (TMEM lane quarter = warp % 4).  M = 128*MT pixels per CTA (MT accumulators
// share every weight stage), N = BN <= 128.  ~80 KB shared memory and <= 256 TMEM columns per CTA so
// two CTAs co-reside per SM: one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace premvos {

namespace {

constexpr int UMMA_THREADS = 256;
constexpr int MAX_A_STAGES = 4, MAX_W_STAGES = 8;
constexpr int SMEM_LIMIT = 226 * 1024;   // dynamic part; 1 KB of the 227 KB stays for static shared memory (debug timestamps)

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (the issuer looks at the NEXT stage's barrier before it issues the current stage's MMAs)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must become a trap (launch error), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// one lane of the (converged) warp; the ELECT form lets ptxas keep MMA/TMA operands in uniform registers
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the 64-bit descriptors given as {low word, constant high word}: the issuer advances only the low words
// (start address >> 4 inside bits [0,14)) with 32-bit adds, so that one MMA costs the issuing thread a handful of
// instructions.  (tools/ubench/mma_rate.cu: descriptors rebuilt with shifts/masks per MMA cap one issuing thread at
// ~150 cycles per MMA; precomputed ones reach the tensor pipe's 64 / 128 cycles per 128xNx16 MMA.)
__device__ __forceinline__ void umma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: a cluster of two CTAs on the two SMs of a TPC runs ONE 256 x N MMA per instruction ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// executed by every thread of both CTAs (convergent warps)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Same without the cluster-scope release (which costs a MEMBAR.ALL.GPU per arrival): what is handed over here is tensor memory,
// ordered by tcgen05.fence::before_thread_sync on this side and ::after_thread_sync on the waiting side, not generic memory.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA box into THIS CTA's shared memory, bytes credited to the barrier at `mbar_cluster_addr` (the pair leader's)
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t mbar_cluster_addr, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t mbar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows of each CTA] * B[N/2 rows of each CTA]^T; issued by one thread of the leader CTA
__device__ __forceinline__ void umma_pair_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one arrival on the barrier at this shared-memory offset in BOTH CTAs once the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// Un-swizzled ("interleaved") K-major operand descriptor: core matrix = 8 rows x 16 B, rows 16 B apart.
//   bits [0,14) start>>4 | [16,30) LBO>>4 (byte distance between the two 8-element K halves of one MMA)
//   | [32,46) SBO>>4 (byte distance between 8-row groups along M/N) | [46,48) version = 1 (sm_100)
//   | [61,64) layout = 0 (SWIZZLE_NONE)     (cute/arch/mma_sm100_desc.hpp, mma_traits_sm100.hpp:194)
// The MMA issuer assembles it from a constant high word and (LBO | address >> 4) in the low word.
// kind::f16 instruction descriptor: D=f32 (bits 4-5 =1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// PREMVOS_DBG & 32: per-launch CTA timing statistics (ns, %globaltimer): [0] min start, [1] max end, [2] sum of CTA durations,
// [3] max CTA duration, [4] min CTA duration, [5] CTAs
__device__ unsigned long long g_dbg_cta[8];
__device__ unsigned long long g_dbg_rec[512][6];   // per CTA: smid, start, end, issuer first MMA, issuer last commit, epilogue end
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct UmmaConvArgs {
  int tiles_x, tiles_y, MT;  // tiles per image; MT accumulators (16 rows x 8 cols each) per CTA ...
  int mt_horizontal;         // ... side by side in x (1) or stacked in y (0), whichever pads the image less
  int Ho, Wo;                // output size
  int stride, R, S, dil, pad_t, pad_l;
  int stride_x;              // horizontal stride (== stride unless the layer was planned with ConvGeom::stride_x)
  int halo;                  // 1: one A box per k-block serves all taps; 0: one A box per (tap, k-block)
  int merged_x;              // 1: tensor map dim 0 = W*8 elements (stride-1 layers); 0: dims {8, W, ...}
  int box_w, box_h;          // A box (pixels) as it lies in shared memory
  int KC, kblocks;           // 8-channel chunks per k-block (even); k-blocks
  int BN, Cout;
  int dbg;                   // ablation switches (PREMVOS_DBG): 1 = producer skips the copies, 2 = issuer skips the MMAs,
                             // 4 = epilogue only drains TMEM, 8 = one k-block per work item
  int NACC;                  // accumulator replicas (1..3): product p of {lo*hi, hi*lo, hi*hi} accumulates into replica p % NACC;
                             // independent accumulators let the tensor pipe overlap the otherwise serial MMA chain
  int TPS;                   // taps per weight stage (divides R*S): one bulk copy fetches TPS taps x {hi, lo}
  int a_stages, w_stages;
  int a_plane;               // bytes of one (hi or lo) A plane of a stage, rounded up to 128
  int a_box_bytes;           // bytes one A TMA actually transfers
  int w_plane;               // bytes of one (hi or lo) plane of one tap = KC*BN*16
  int w_stage;               // bytes of a weight stage (TPS * 2 * w_plane), rounded up to 128
  const __nv_bfloat16* w;    // packed [ntile][kblock][tap][plane hi|lo][KC][BN][8]
  const float* bias;         // [ntiles*BN]
  float slope;               // LeakyReLU slope (1 = identity, 0 = ReLU)
  // outputs: CP8 split planes and/or fp32 channels-last
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int out_chunks, out_c0;
  float* out_f32; int out_cs, out_coff;
  // channel routing: [0, cp_cout) -> CP8 planes (activated); [f32_first, Cout) -> fp32 channels-last at channel c - f32_first
  int cp_cout, f32_first, f32_linear /*1: no activation on the fp32 outputs*/, f32_accum /*1: add to what is there*/;
  const __nv_bfloat16* res_hi; const __nv_bfloat16* res_lo; int res_chunks, res_c0;  // optional residual
  float* out_f8; int f8_chunks, f8_c0;                 // optional F8 (fp32 chunk-planar) output, activated like the CP8 one
  const float* res_f8; int resf_chunks, resf_c0;       // optional residual in F8 (instead of res_hi / res_lo)
  uint32_t tmem_cols;
  // split-K (grids smaller than the machine): blockIdx.z owns k-blocks [z*kb_per, (z+1)*kb_per) and stores raw fp32
  // partial sums; conv_finish_kernel adds them in a fixed order (deterministic) and applies the epilogue
  int ksplit, kb_per, cout_pad;
  float* partial;            // [ksplit][N*Ho*Wo][cout_pad]
  long partial_stride;       // elements per split
  // persistent scheduling: work item t = ((z * ntiles + ntile) * n_images + n) * tiles_per_img + tile, CTA b takes b, b+grid, ...
  int ntiles, n_images, total_work;
  int flat_hw;               // > 0: 1x1 stride-1 layer run over the FLATTENED pixel list of each plane (tiles of 128 consecutive
                             // pixels, no 2-D tile padding); value = real H*W, geometry fields describe an [ceil(HW/8)][8] image
  int nbuf;                  // TMEM accumulator buffers (2: the epilogue of item i overlaps the MMAs of item i+1)
  // tail split: when the item count is a little more than a whole number of waves, the last `tail_items` items are split along K
  // into `tail_split` sub-items each (work indices >= main_work), which store raw fp32 partial sums [z][tail item][row][BN] into
  // `partial`; conv_finish_tail_kernel adds them in a fixed order and applies the epilogue.  Without it those few items cost a
  // whole extra wave (e.g. 300 items on 148 CTAs: 3 waves instead of 2.03).
  int tail_items, tail_split, tail_kb_per, main_work;
  int lockstep;              // 1x1 layers: A and weight rings advance together and share one barrier pair per k-block
  int epi_warps;             // 4 or 8 epilogue warps (8: kernel instantiation with 384 threads, one CTA per SM)
  int pair;                  // 1: conv_pair_kernel (cta_group::2, 256 pixels x 256 output channels per CTA pair); see there
  int bias_smem;             // pair kernel: floats of bias staged in shared memory (ntiles * BN, or 0: read from global memory)
  int stream_k_all;          // pair kernel: 1 = no round-robin phase, all items are cut into equal k-block ranges (tuning switch)
  unsigned* sk_flags;        // pair kernel, stream-K: [pairs][2 ranks][8 epilogue warps] flags; `partial` = [pairs][2][128][256] fp32
  // folded small maps (Ho*Wo <= 64): one 128-row tile holds `fold` whole images, row m = g * slot + oy * Wo + ox.  The A box of a
  // (tap, k-block) is {8 channels, Wo, Ho, fold images, KC chunks} of a 5-D map whose 4th dimension is the image, so every image
  // gets its own zero padding from the TMA's out-of-bounds fill; `n_images` counts image GROUPS, `n_real` images.
  int fold, slot, n_real;
  // stacked halo tiles (3x3 stride-1 layers on maps at most 8 wide): the halo box of a k-block is {W*8, hf_pitch rows, fold images,
  // KC chunks}; image g of the group occupies tile rows [g * hf_pitch, g * hf_pitch + Ho) and brings its own zero padding rows
  int hf_pitch;
  // work order: 0 = output-channel tile outermost (weights of a tile stay hot, A streamed once per tile: right while A fits the L2),
  // 1 = output-channel tile innermost (the ntiles CTAs that share an A tile run side by side, A comes from DRAM once: layers whose
  // input is larger than the L2, e.g. the RoI head's entry on 400 x 14 x 14 x 1024)
  int n_fastest;
  int a_lbo;                 // bytes between the two 8-channel chunks of a K step in the A stage (0: box_h * box_w * 16)
};

// EPI_WARPS = 4: 256 threads, up to two CTAs per SM.  EPI_WARPS = 8: 384 threads, one CTA per SM owning the whole TMEM (256-wide N
// tiles); two warps share each TMEM lane quarter and split the accumulator columns.
template <int EPI_WARPS>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, EPI_WARPS == 4 ? 2 : 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo, const UmmaConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* a_smem = smem;
  uint8_t* w_smem = a_smem + (size_t)a.a_stages * 2 * a.a_plane;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + (size_t)a.w_stages * a.w_stage);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + MAX_A_STAGES;
  uint64_t* w_full = bars + 2 * MAX_A_STAGES;
  uint64_t* w_empty = w_full + MAX_W_STAGES;
  uint64_t* tmem_full_bar = w_empty + MAX_W_STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_addr_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ long long dbg_ts[5][24];   // PREMVOS_DBG & 32: per-k-block timestamps of CTA 0 (tuning aid)
  int dbg_i = 0, dbg_j = 0;
  const bool dbg_on = (a.dbg & 32) && blockIdx.x == 0;
  const unsigned long long dbg_t0 = (a.dbg & 32) ? globaltimer_ns() : 0ull;
  __shared__ unsigned long long dbg_ev[4];
  if ((a.dbg & 32) && threadIdx.x < 4) dbg_ev[threadIdx.x] = 0;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int taps = a.R * a.S;
  // decode of a work item (once per item, the only integer divisions in the kernel)
  struct Work { int n_img, ty0, tx0, ntile, z, kb_begin, kb_end, tail_slot; };
  auto decode = [&](int t) {
    Work w;
    w.tail_slot = -1;
    int tz = 0;
    if (a.tail_items > 0 && t >= a.main_work) {   // K-split sub-item of one of the last items
      const int u = t - a.main_work;
      w.tail_slot = u / a.tail_split;
      tz = u - w.tail_slot * a.tail_split;
      t = a.main_work + w.tail_slot;
    }
    int tile, r;
    if (a.n_fastest) {
      w.ntile = t % a.ntiles; r = t / a.ntiles;
      tile = r % tiles_per_img; r /= tiles_per_img;
      w.n_img = (r % a.n_images) * (a.fold > 0 ? a.fold : 1);
      w.z = r / a.n_images;
    } else {
      tile = t % tiles_per_img;
      r = t / tiles_per_img;
      w.n_img = (r % a.n_images) * (a.fold > 0 ? a.fold : 1); r /= a.n_images;
      w.ntile = r % a.ntiles;
      w.z = r / a.ntiles;
    }
    w.ty0 = (tile / a.tiles_x) * (a.mt_horizontal ? 16 : 16 * a.MT);
    w.tx0 = (tile % a.tiles_x) * (a.mt_horizontal ? 8 * a.MT : 8);
    w.kb_begin = a.ksplit > 1 ? w.z * a.kb_per : 0;
    w.kb_end = a.ksplit > 1 ? min(a.kblocks, w.kb_begin + a.kb_per) : a.kblocks;
    if (w.tail_slot >= 0) { w.z = tz; w.kb_begin = tz * a.tail_kb_per; w.kb_end = min(a.kblocks, w.kb_begin + a.tail_kb_per); }
    if (a.dbg & 8) w.kb_end = w.kb_begin + 1;   // ablation: one k-block per item (wrong results)
    return w;
  };

  if (warp == 0 && lane == 0) {  // one-time setup
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo);
    for (int s = 0; s < a.a_stages && s < MAX_A_STAGES; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < a.w_stages; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_addr_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_addr_slot, 0);

  if (warp == 0) {
    // ===== TMA producer =====  the whole warp walks the warp-uniform loop, one elected lane issues the copies
    uint32_t a_st = 0, a_ph = 0, w_st = 0, w_ph = 0;
    const uint32_t w_bytes = (uint32_t)a.TPS * 2u * (uint32_t)a.w_plane;   // one bulk copy
    for (int t = blockIdx.x; t < a.total_work; t += gridDim.x) {
    const Work wk = decode(t);
    const int n_img = wk.n_img;
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w) + ((size_t)wk.ntile * a.kblocks + wk.kb_begin) * taps * 2 * a.w_plane;
    const int bx = wk.tx0 * a.stride_x - a.pad_l, by = wk.ty0 * a.stride - a.pad_t;
    if (a.lockstep) {
      // 1x1 layers: one ring, one barrier pair per k-block -- the A box and the weight block of a k-block share the stage
      // index and the w_full / w_empty barriers (a tcgen05.commit per ring per k-block costs more than the k-block's MMAs)
      for (int kb = wk.kb_begin; kb < wk.kb_end; kb++) {
        mbar_wait(&w_empty[w_st], w_ph ^ 1u);
        if (dbg_on && lane == 0 && dbg_j < 24) dbg_ts[3][dbg_j] = clock64();
        if (a.dbg & 1) {
          if (elect_one()) mbar_arrive(&w_full[w_st]);
        } else if (elect_one()) {
          uint8_t* dst = a_smem + (size_t)w_st * 2 * a.a_plane;
          mbar_arrive_expect_tx(&w_full[w_st], 2u * (uint32_t)a.a_box_bytes + w_bytes);
          if (a.fold && !a.hf_pitch) {
            tma_load_5d(&tmA_hi, &w_full[w_st], dst, 0, bx, by, n_img, kb * a.KC);
            tma_load_5d(&tmA_lo, &w_full[w_st], dst + a.a_plane, 0, bx, by, n_img, kb * a.KC);
          } else if (a.merged_x) {
            tma_load_4d(&tmA_hi, &w_full[w_st], dst, bx * 8, by, kb * a.KC, n_img);
            tma_load_4d(&tmA_lo, &w_full[w_st], dst + a.a_plane, bx * 8, by, kb * a.KC, n_img);
          } else {
            tma_load_5d(&tmA_hi, &w_full[w_st], dst, 0, bx, by, kb * a.KC, n_img);
            tma_load_5d(&tmA_lo, &w_full[w_st], dst + a.a_plane, 0, bx, by, kb * a.KC, n_img);
          }
          bulk_load_1d(w_smem + (size_t)w_st * a.w_stage, wsrc, w_bytes, &w_full[w_st]);
        }
        __syncwarp();
        if (dbg_on && lane == 0 && dbg_j < 24) dbg_ts[4][dbg_j++] = clock64();
        wsrc += w_bytes;
        if (++w_st == (uint32_t)a.w_stages) { w_st = 0; w_ph ^= 1u; }
      }
      continue;
    }
    for (int kb = wk.kb_begin; kb < wk.kb_end; kb++) {
      int tin = 0;  // tap index inside the current weight stage
      for (int r = 0; r < a.R; r++)
        for (int s = 0; s < a.S; s++) {
          if (!a.halo || (r | s) == 0) {
            const int px = a.halo ? bx : bx + s * a.dil, py = a.halo ? by : by + r * a.dil;
            mbar_wait(&a_empty[a_st], a_ph ^ 1u);
            if (a.dbg & 1) {
              if (elect_one()) mbar_arrive(&a_full[a_st]);
            } else if (elect_one()) {
              uint8_t* dst = a_smem + (size_t)a_st * 2 * a.a_plane;
              mbar_arrive_expect_tx(&a_full[a_st], 2u * (uint32_t)a.a_box_bytes);
              if (a.hf_pitch) {
                tma_load_4d(&tmA_hi, &a_full[a_st], dst, px * 8, py, n_img, kb * a.KC);
                tma_load_4d(&tmA_lo, &a_full[a_st], dst + a.a_plane, px * 8, py, n_img, kb * a.KC);
              } else if (a.fold) {
                tma_load_5d(&tmA_hi, &a_full[a_st], dst, 0, px, py, n_img, kb * a.KC);
                tma_load_5d(&tmA_lo, &a_full[a_st], dst + a.a_plane, 0, px, py, n_img, kb * a.KC);
              } else if (a.merged_x) {
                tma_load_4d(&tmA_hi, &a_full[a_st], dst, px * 8, py, kb * a.KC, n_img);
                tma_load_4d(&tmA_lo, &a_full[a_st], dst + a.a_plane, px * 8, py, kb * a.KC, n_img);
              } else {
                tma_load_5d(&tmA_hi, &a_full[a_st], dst, 0, px, py, kb * a.KC, n_img);
                tma_load_5d(&tmA_lo, &a_full[a_st], dst + a.a_plane, 0, px, py, kb * a.KC, n_img);
              }
            }
            __syncwarp();
            if (++a_st == (uint32_t)a.a_stages) { a_st = 0; a_ph ^= 1u; }
          }
          if (tin == 0) {
            mbar_wait(&w_empty[w_st], w_ph ^ 1u);
            if (a.dbg & 1) {
              if (elect_one()) mbar_arrive(&w_full[w_st]);
            } else if (elect_one()) {
              mbar_arrive_expect_tx(&w_full[w_st], w_bytes);
              bulk_load_1d(w_smem + (size_t)w_st * a.w_stage, wsrc, w_bytes, &w_full[w_st]);
            }
            __syncwarp();
            wsrc += w_bytes;
            if (++w_st == (uint32_t)a.w_stages) { w_st = 0; w_ph ^= 1u; }
          }
          if (++tin == a.TPS) tin = 0;
        }
    }
    }  // work loop
  } else if (warp == 1) {
    // ===== MMA issuer =====  the whole warp walks the (warp-uniform) loop, one elected lane issues
    const uint32_t idesc = make_idesc_bf16(128, a.BN);
    const uint32_t a_sbo = (uint32_t)a.box_w * 16u, a_lbo = a.a_lbo ? (uint32_t)a.a_lbo : (uint32_t)a.box_h * a.box_w * 16u;
    const uint32_t w_sbo = 128u, w_lbo = (uint32_t)a.BN * 16u;
    // descriptor = constant high word (SBO, version) + low word (LBO | start address >> 4)
    const uint32_t a_hi32 = (a_sbo >> 4) | (1u << 14), w_hi32 = (w_sbo >> 4) | (1u << 14);
    const uint32_t a_lo32 = (a_lbo >> 4) << 16, w_lo32 = (w_lbo >> 4) << 16;
    const uint32_t a_base = smem_u32(a_smem), w_base = smem_u32(w_smem);
    const uint32_t a_stage_bytes = 2u * (uint32_t)a.a_plane, w_stage_bytes = (uint32_t)a.w_stage;
    const uint32_t a_kstep16 = (2u * a_lbo) >> 4, w_kstep16 = (2u * w_lbo) >> 4, a_mstep16 = (a.mt_horizontal ? 128u : 16u * a_sbo) >> 4;
    const uint32_t a_plane16 = (uint32_t)a.a_plane >> 4, w_plane16 = (uint32_t)a.w_plane >> 4;
    const uint32_t tap_row = a.halo ? (uint32_t)(a.dil * a.box_w) * 16u : 0u, tap_col = a.halo ? (uint32_t)a.dil * 16u : 0u;
    uint32_t a_st = 0, a_ph = 0, w_st = 0, w_ph = 0, cur_a = 0, a_addr_stage = 0;
    uint32_t buf = 0, empty_ph = 0;   // bit b = phase of tmem_empty_bar[b]
    const int ksteps = a.KC / 2;
    const uint32_t rep_cols = (uint32_t)(a.MT * a.BN);
    // One barrier handshake per weight stage (TPS taps): all of its MMAs are issued in one go, one commit
    // releases the stage.  (A per-tap handshake costs ~500 cycles of mbarrier/commit latency -- more than the
    // MMAs of one tap.)  Tap mode additionally waits for / releases one A stage per tap.
    const int wgroups = taps / a.TPS;
    for (int t = blockIdx.x; t < a.total_work; t += gridDim.x) {
    const Work wk = decode(t);
    // the accumulator buffer must have been drained by the epilogue of the item that used it last
    mbar_wait(&tmem_empty_bar[buf], ((empty_ph >> buf) & 1u) ^ 1u);
    empty_ph ^= 1u << buf;
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base + buf * rep_cols * (uint32_t)a.NACC;
    uint32_t accum = 0;
    for (int kb = wk.kb_begin; kb < wk.kb_end; kb++) {
      int r = 0, s = 0;
      uint32_t row_off = 0, tap_off = 0;
      for (int wg = 0; wg < wgroups; wg++) {
        if (a.halo && wg == 0) {
          mbar_wait(&a_full[a_st], a_ph);
          cur_a = a_st;
          a_addr_stage = a_base + a_st * a_stage_bytes;
          if (++a_st == (uint32_t)a.a_stages) { a_st = 0; a_ph ^= 1u; }
        }
        mbar_wait(&w_full[w_st], w_ph);
        tc_fence_after();
        uint32_t w_addr_tap = w_base + w_st * w_stage_bytes;
        if (a.halo || a.lockstep) {
          // One elected region per weight stage: every MMA of its TPS taps and the commits that release the stage.  The
          // tensor pipe queues very few MMAs per CTA (tools/ubench + PREMVOS_DBG ablations: MMA time adds to the issuer's
          // control time instead of overlapping it), so every instruction between two UTCHMMAs of consecutive stages is
          // idle tensor time: no second elect, no __syncwarp, tap bookkeeping only in the elected lane (always the same one).
          // (Tried and dropped: one elected lane running the whole K loop with a non-blocking probe of the next stage's
          // barrier -- 416 instead of 512 control cycles per k-block, but lane-private ring state costs R2UR moves in front of
          // every UTCHMMA: +70 issue cycles, slower for halo layers.  In-kernel timestamps, PREMVOS_DBG=96: a 1x1 k-block at
          // 128x256 takes ~1050 cycles = 46 B/clk of L2 -> shared-memory fill, i.e. these layers sit at the fill cap.)
          if (a.lockstep) { cur_a = w_st; a_addr_stage = a_base + w_st * a_stage_bytes; }
          if (elect_one()) {
            if ((a.dbg & 32) && dbg_ev[0] == 0) dbg_ev[0] = globaltimer_ns();
            if (dbg_on && dbg_i < 24) dbg_ts[0][dbg_i] = clock64();
            for (int t = 0; t < a.TPS; t++) {
              uint32_t aH = a_lo32 + ((a_addr_stage + tap_off) >> 4), wH = w_lo32 + (w_addr_tap >> 4);
              for (int ks = 0; ks < ((a.dbg & 2) ? 0 : ksteps); ks++, aH += a_kstep16, wH += w_kstep16) {
                const uint32_t wL = wH + w_plane16;
                uint32_t ah = aH, d = tmem_acc;
                for (int mt = 0; mt < a.MT; mt++, ah += a_mstep16, d += (uint32_t)a.BN) {
                  const uint32_t al = ah + a_plane16;
                  umma_bf16_lo(d, al, a_hi32, wH, w_hi32, idesc, accum);
                  umma_bf16_lo(d, ah, a_hi32, wL, w_hi32, idesc, 1u);
                  umma_bf16_lo(d, ah, a_hi32, wH, w_hi32, idesc, 1u);
                }
                accum = 1u;
              }
              w_addr_tap += 2u * (uint32_t)a.w_plane;
              if (++s == a.S) { s = 0; r++; row_off += tap_row; tap_off = row_off; } else { tap_off += tap_col; }
            }
            if (dbg_on && dbg_i < 24) dbg_ts[1][dbg_i] = clock64();
            umma_commit(&w_empty[w_st]);                                   // frees the weight slot (and the A box in lockstep mode)
            if (a.halo && wg == wgroups - 1) umma_commit(&a_empty[cur_a]);  // ... and the halo box after the last tap
            if (dbg_on && dbg_i < 24) dbg_ts[2][dbg_i++] = clock64();
          }
          accum = 1u;   // lanes that were not elected: keep the (unused) copy consistent
        } else {
          // tap mode: one A box per tap, waited for by the whole warp
          for (int t = 0; t < a.TPS; t++) {
            mbar_wait(&a_full[a_st], a_ph);
            tc_fence_after();
            cur_a = a_st;
            a_addr_stage = a_base + a_st * a_stage_bytes;
            if (++a_st == (uint32_t)a.a_stages) { a_st = 0; a_ph ^= 1u; }
            if (elect_one()) {
              uint32_t aH = a_lo32 + (a_addr_stage >> 4), wH = w_lo32 + (w_addr_tap >> 4);
              for (int ks = 0; ks < ((a.dbg & 2) ? 0 : ksteps); ks++, aH += a_kstep16, wH += w_kstep16) {
                const uint32_t wL = wH + w_plane16;
                uint32_t ah = aH, d = tmem_acc;
                for (int mt = 0; mt < a.MT; mt++, ah += a_mstep16, d += (uint32_t)a.BN) {
                  const uint32_t al = ah + a_plane16;
                  umma_bf16_lo(d, al, a_hi32, wH, w_hi32, idesc, accum);
                  umma_bf16_lo(d, ah, a_hi32, wL, w_hi32, idesc, 1u);
                  umma_bf16_lo(d, ah, a_hi32, wH, w_hi32, idesc, 1u);
                }
                accum = 1u;
              }
              umma_commit(&a_empty[cur_a]);
              if (t == a.TPS - 1) umma_commit(&w_empty[w_st]);
            }
            accum = 1u;
            w_addr_tap += 2u * (uint32_t)a.w_plane;
          }
        }
        if (++w_st == (uint32_t)a.w_stages) { w_st = 0; w_ph ^= 1u; }
      }
    }
    if ((a.dbg & 32) && lane == 0) dbg_ev[1] = globaltimer_ns();
    if (elect_one()) umma_commit(&tmem_full_bar[buf]);  // accumulators of this item complete
    __syncwarp();
    if (a.nbuf > 1) buf ^= 1u;
    }  // work loop
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;  // accumulator row == pixel inside the 16x8 sub-tile
    const long hw = a.flat_hw > 0 ? (long)a.flat_hw : (long)a.Ho * a.Wo;
    uint32_t buf = 0, full_ph = 0;    // bit b = phase of tmem_full_bar[b]
    const uint32_t buf_cols = (uint32_t)(a.MT * a.BN * a.NACC);
    // accumulator columns of this warp: all of them, or one half when two warps share a lane quarter
    const int col_part = (warp - 4) >> 2, col_span = EPI_WARPS == 8 ? a.BN / 2 : a.BN;
    const int col_begin = col_part * col_span, col_end = col_begin + col_span;
    // folded small maps: this thread's accumulator row is pixel (f_oy, f_ox) of image f_g of the group, for every item
    const bool tfold = a.fold && !a.hf_pitch;
    const int f_g = tfold ? m / a.slot : 0, f_p = tfold ? m - f_g * a.slot : 0, f_oy = tfold ? f_p / a.Wo : 0, f_ox = tfold ? f_p - f_oy * a.Wo : 0;
    for (int t = blockIdx.x; t < a.total_work; t += gridDim.x) {
    const Work wk = decode(t);
    const int n_img0 = wk.n_img + f_g, ty0 = wk.ty0, tx0 = wk.tx0, ntile = wk.ntile;
    mbar_wait(&tmem_full_bar[buf], (full_ph >> buf) & 1u);
    full_ph ^= 1u << buf;
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base + buf * buf_cols;
    for (int mt = 0; mt < a.MT; mt++) {
      int oy = tfold ? f_oy : ty0 + (a.mt_horizontal ? 0 : mt * 16) + (m >> 3), n_img = n_img0;
      const int ox = tfold ? f_ox : tx0 + (a.mt_horizontal ? mt * 8 : 0) + (m & 7);
      bool in_group = true;
      if (a.hf_pitch) {   // stacked halo tile: tile row -> (image of the group, row of that image)
        const int g = oy / a.hf_pitch;
        oy -= g * a.hf_pitch;
        n_img += g;
        in_group = g < a.fold && n_img < a.n_real;
      }
      const long pix = (long)oy * a.Wo + ox;
      const bool in_img = tfold ? (f_g < a.fold && n_img < a.n_real) : a.flat_hw > 0 ? pix < hw : (in_group && oy < a.Ho && ox < a.Wo);
      for (int c0 = col_begin; c0 < col_end; c0 += 16) {
        uint32_t v[16];
        __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * a.BN + c0);
        const bool last_ld = (mt == a.MT - 1) && (c0 + 16 >= col_end);
        tmem_ld16(taddr, v);
        // global loads of this group (bias, residual) are issued while the TMEM load is in flight
        const int co0 = ntile * a.BN + c0;
        const bool live = in_img && co0 < a.Cout && a.ksplit == 1 && wk.tail_slot < 0;
        float4 bv[4];
        uint4 rres[4];
        if (live) {
#pragma unroll
          for (int j = 0; j < 4; j++) bv[j] = __ldg(reinterpret_cast<const float4*>(a.bias + co0) + j);
          if (a.res_hi) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
              if (co0 + 8 * h >= a.Cout) continue;
              const long ri = (((long)n_img * a.res_chunks + a.res_c0 + (co0 >> 3) + h) * hw + pix) * 8;
              rres[2 * h] = *reinterpret_cast<const uint4*>(a.res_hi + ri);
              rres[2 * h + 1] = *reinterpret_cast<const uint4*>(a.res_lo + ri);
            }
          } else if (a.res_f8) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
              if (co0 + 8 * h >= a.Cout) continue;
              const long ri = (((long)n_img * a.resf_chunks + a.resf_c0 + (co0 >> 3) + h) * hw + pix) * 8;
              rres[2 * h] = *reinterpret_cast<const uint4*>(a.res_f8 + ri);
              rres[2 * h + 1] = *reinterpret_cast<const uint4*>(a.res_f8 + ri + 4);
            }
          }
        }
        tmem_ld_wait();
        for (int rep = 1; rep < a.NACC; rep++) {
          uint32_t v2[16];
          tmem_ld16(taddr + (uint32_t)(rep * a.MT * a.BN), v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
        if (last_ld) {  // every accumulator column of this item is in registers: hand the TMEM buffer back to the issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        if (a.dbg & 4) continue;   // ablation: epilogue drains TMEM only
        if (wk.tail_slot >= 0) {  // tail sub-item: raw partial sums of this tile, [z][tail item][MT*128 rows][BN]
          float* pp = a.partial + (((long)(wk.z * a.tail_items + wk.tail_slot) * a.MT + mt) * 128 + m) * a.BN + c0;
#pragma unroll
          for (int j = 0; j < 4; j++)
            reinterpret_cast<float4*>(pp)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          continue;
        }
        if (a.ksplit > 1) {  // raw partial sums; bias / residual / activation happen in conv_finish_kernel
          if (in_img) {
            float* pp = a.partial + (long)wk.z * a.partial_stride + ((long)n_img * hw + pix) * a.cout_pad + co0;
#pragma unroll
            for (int j = 0; j < 4; j++)
              reinterpret_cast<float4*>(pp)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
          continue;
        }
        if (!in_img || co0 >= a.Cout) continue;
        float f[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          f[4 * j] = __uint_as_float(v[4 * j]) + bv[j].x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bv[j].y;
          f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bv[j].z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bv[j].w;
        }
        if (a.res_hi) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (co0 + 8 * h >= a.Cout) continue;
            const uint4 rh = rres[2 * h], rl = rres[2 * h + 1];
            const uint32_t hh[4] = {rh.x, rh.y, rh.z, rh.w}, ll[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              f[8 * h + 2 * j] += __uint_as_float(hh[j] << 16) + __uint_as_float(ll[j] << 16);
              f[8 * h + 2 * j + 1] += __uint_as_float(hh[j] & 0xffff0000u) + __uint_as_float(ll[j] & 0xffff0000u);
            }
          }
        }
        else if (a.res_f8) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (co0 + 8 * h >= a.Cout) continue;
            const uint4 r0 = rres[2 * h], r1 = rres[2 * h + 1];
            f[8 * h] += __uint_as_float(r0.x); f[8 * h + 1] += __uint_as_float(r0.y); f[8 * h + 2] += __uint_as_float(r0.z); f[8 * h + 3] += __uint_as_float(r0.w);
            f[8 * h + 4] += __uint_as_float(r1.x); f[8 * h + 5] += __uint_as_float(r1.y); f[8 * h + 6] += __uint_as_float(r1.z); f[8 * h + 7] += __uint_as_float(r1.w);
          }
        }
        if (a.out_f32 && a.f32_linear) {  // fp32 outputs without activation (fused flow heads)
          float* pf = a.out_f32 + ((long)n_img * hw + pix) * a.out_cs + a.out_coff - a.f32_first + co0;
#pragma unroll
          for (int j = 0; j < 16; j++)
            if (co0 + j >= a.f32_first && co0 + j < a.Cout) pf[j] = a.f32_accum ? pf[j] + f[j] : f[j];
        }
#pragma unroll
        for (int j = 0; j < 16; j++) f[j] = f[j] > 0.f ? f[j] : f[j] * a.slope;
        if (a.out_f8) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (co0 + 8 * h >= a.Cout) continue;
            float* pf8 = a.out_f8 + (((long)n_img * a.f8_chunks + a.f8_c0 + (co0 >> 3) + h) * hw + pix) * 8;
            reinterpret_cast<float4*>(pf8)[0] = make_float4(f[8 * h], f[8 * h + 1], f[8 * h + 2], f[8 * h + 3]);
            reinterpret_cast<float4*>(pf8)[1] = make_float4(f[8 * h + 4], f[8 * h + 5], f[8 * h + 6], f[8 * h + 7]);
          }
        }
        if (a.out_hi) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (co0 + 8 * h >= a.cp_cout) continue;  // chunk beyond the CP8 range (padding channels inside a chunk are exact zeros)
            uint32_t hw4[4], lw4[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float x0 = f[8 * h + 2 * j], x1 = f[8 * h + 2 * j + 1];
              const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
              const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
              const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
              hw4[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              lw4[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            const long oi = (((long)n_img * a.out_chunks + a.out_c0 + (co0 >> 3) + h) * hw + pix) * 8;
            *reinterpret_cast<uint4*>(a.out_hi + oi) = make_uint4(hw4[0], hw4[1], hw4[2], hw4[3]);
            *reinterpret_cast<uint4*>(a.out_lo + oi) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
          }
        }
        if (a.out_f32 && !a.f32_linear) {
          float* pf = a.out_f32 + ((long)n_img * hw + pix) * a.out_cs + a.out_coff - a.f32_first + co0;
          if (a.f32_first == 0 && !a.f32_accum && co0 + 16 <= a.Cout && ((a.out_cs | a.out_coff) & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) reinterpret_cast<float4*>(pf)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
              if (co0 + j >= a.f32_first && co0 + j < a.Cout) pf[j] = a.f32_accum ? pf[j] + f[j] : f[j];
          }
        }
      }
    }
    if (a.nbuf > 1) buf ^= 1u;
    if ((a.dbg & 32) && warp == 4 && lane == 0) dbg_ev[2] = globaltimer_ns();
    }  // work loop
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, a.tmem_cols);
  if ((a.dbg & 32) && threadIdx.x == 0) {
    const unsigned long long t1 = globaltimer_ns(), d = t1 - dbg_t0;
    atomicMin(&g_dbg_cta[0], dbg_t0); atomicMax(&g_dbg_cta[1], t1); atomicAdd(&g_dbg_cta[2], d);
    atomicMax(&g_dbg_cta[3], d); atomicMin(&g_dbg_cta[4], d); atomicAdd(&g_dbg_cta[5], 1ull);
    if (blockIdx.x < 512) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long* rr = g_dbg_rec[blockIdx.x];
      rr[0] = smid; rr[1] = dbg_t0; rr[2] = t1; rr[3] = dbg_ev[0]; rr[4] = dbg_ev[1]; rr[5] = dbg_ev[2];
    }
  }
  if (dbg_on && threadIdx.x == 32 && (a.dbg & 64)) {
    for (int i = 1; i < dbg_i; i++)
      printf("kb %2d: issuer wait->start %6lld  issue %6lld  commit %5lld | producer empty-wait %6lld issue %5lld (start-to-start %6lld)\n", i,
             dbg_ts[0][i] - dbg_ts[2][i - 1], dbg_ts[1][i] - dbg_ts[0][i], dbg_ts[2][i] - dbg_ts[1][i],
             dbg_ts[3][i] - dbg_ts[4][i - 1], dbg_ts[4][i] - dbg_ts[3][i], dbg_ts[0][i] - dbg_ts[0][i - 1]);
  }
}

// Shared tail of the split-K finish kernels: bias, residual, fp32 outputs, activation, CP8 split store for one (pixel, 8-channel
// chunk).  pg = global pixel index (n_img * hw + pix), c8 = 8-channel chunk of the output.
__device__ __forceinline__ void finish_store(const UmmaConvArgs& a, int n_img, long pix, long pg, int c8, float (&f)[8]) {
  const long hw = a.flat_hw > 0 ? (long)a.flat_hw : (long)a.Ho * a.Wo;
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] += __ldg(a.bias + c8 * 8 + j);
  if (a.res_hi) {
    const long ri = (((long)n_img * a.res_chunks + a.res_c0 + c8) * hw + pix) * 8;
    const uint4 rh = *reinterpret_cast<const uint4*>(a.res_hi + ri);
    const uint4 rl = *reinterpret_cast<const uint4*>(a.res_lo + ri);
    const uint32_t hh[4] = {rh.x, rh.y, rh.z, rh.w}, ll[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      f[2 * j] += __uint_as_float(hh[j] << 16) + __uint_as_float(ll[j] << 16);
      f[2 * j + 1] += __uint_as_float(hh[j] & 0xffff0000u) + __uint_as_float(ll[j] & 0xffff0000u);
    }
  } else if (a.res_f8) {
    const float4* rp = reinterpret_cast<const float4*>(a.res_f8 + (((long)n_img * a.resf_chunks + a.resf_c0 + c8) * hw + pix) * 8);
    const float4 r0 = rp[0], r1 = rp[1];
    f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w; f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
  }
  float* pf = a.out_f32 ? a.out_f32 + pg * a.out_cs + a.out_coff - a.f32_first + c8 * 8 : nullptr;
  if (pf && a.f32_linear) {
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (c8 * 8 + j >= a.f32_first && c8 * 8 + j < a.Cout) pf[j] = a.f32_accum ? pf[j] + f[j] : f[j];
  }
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] = f[j] > 0.f ? f[j] : f[j] * a.slope;
  if (a.out_f8) {
    float4* pf8 = reinterpret_cast<float4*>(a.out_f8 + (((long)n_img * a.f8_chunks + a.f8_c0 + c8) * hw + pix) * 8);
    pf8[0] = make_float4(f[0], f[1], f[2], f[3]);
    pf8[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  if (a.out_hi && c8 * 8 < a.cp_cout) {
    uint32_t hw4[4], lw4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float x0 = f[2 * j], x1 = f[2 * j + 1];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
      hw4[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lw4[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const long oi = (((long)n_img * a.out_chunks + a.out_c0 + c8) * hw + pix) * 8;
    *reinterpret_cast<uint4*>(a.out_hi + oi) = make_uint4(hw4[0], hw4[1], hw4[2], hw4[3]);
    *reinterpret_cast<uint4*>(a.out_lo + oi) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
  }
  if (pf && !a.f32_linear) {
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (c8 * 8 + j >= a.f32_first && c8 * 8 + j < a.Cout) pf[j] = a.f32_accum ? pf[j] + f[j] : f[j];
  }
}

// Sums the split-K partials in split order and applies the same epilogue as the fused path (bias, residual, LeakyReLU/ReLU,
// CP8 split store and/or fp32 channels-last store).  One thread per (pixel, 8-channel chunk).
__global__ void __launch_bounds__(256) conv_finish_kernel(const UmmaConvArgs a, int n_active) {
  const long hw = a.flat_hw > 0 ? (long)a.flat_hw : (long)a.Ho * a.Wo;
  const int cchunks = (a.Cout + 7) / 8;
  const long total = (long)n_active * hw * cchunks;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long pg = idx % ((long)n_active * hw);   // pixel fastest -> coalesced CP8 stores
  const int c8 = (int)(idx / ((long)n_active * hw));
  const int n_img = (int)(pg / hw);
  const long pix = pg - (long)n_img * hw;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] = 0.f;
  for (int z = 0; z < a.ksplit; z++) {
    const float* pp = a.partial + (long)z * a.partial_stride + pg * a.cout_pad + c8 * 8;
    const float4 p0 = reinterpret_cast<const float4*>(pp)[0], p1 = reinterpret_cast<const float4*>(pp)[1];
    f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w; f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
  }
  finish_store(a, n_img, pix, pg, c8, f);
}

// Finishes the K-split tail items of a launch (see UmmaConvArgs::tail_items): one thread per (tail item, tile row, 8-channel chunk
// of the item's N tile); partial sums are added in split order (deterministic).
__global__ void __launch_bounds__(256) conv_finish_tail_kernel(const UmmaConvArgs a) {
  const int rows = a.MT * 128, cch = a.BN / 8;
  const long total = (long)a.tail_items * rows * cch;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int row = (int)(idx % rows);            // row fastest -> coalesced CP8 stores
  const int cc = (int)((idx / rows) % cch);
  const int slot = (int)(idx / ((long)rows * cch));
  // decode the item exactly as the convolution kernel does
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int t = a.main_work + slot;
  int tile, r, n_img, ntile;
  if (a.n_fastest) {
    ntile = t % a.ntiles; r = t / a.ntiles;
    tile = r % tiles_per_img; r /= tiles_per_img;
    n_img = r % a.n_images;
  } else {
    tile = t % tiles_per_img;
    r = t / tiles_per_img;
    n_img = r % a.n_images; r /= a.n_images;
    ntile = r % a.ntiles;
  }
  const int ty0 = (tile / a.tiles_x) * (a.mt_horizontal ? 16 : 16 * a.MT), tx0 = (tile % a.tiles_x) * (a.mt_horizontal ? 8 * a.MT : 8);
  const int mt = row >> 7, m = row & 127;
  const int oy = ty0 + (a.mt_horizontal ? 0 : mt * 16) + (m >> 3), ox = tx0 + (a.mt_horizontal ? mt * 8 : 0) + (m & 7);
  const long hw = a.flat_hw > 0 ? (long)a.flat_hw : (long)a.Ho * a.Wo;
  const long pix = (long)oy * a.Wo + ox;
  const bool in_img = a.flat_hw > 0 ? pix < hw : (oy < a.Ho && ox < a.Wo);
  const int c8 = (ntile * a.BN) / 8 + cc;       // global 8-channel chunk
  if (!in_img || c8 * 8 >= a.Cout) return;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; j++) f[j] = 0.f;
  for (int z = 0; z < a.tail_split; z++) {
    const float* pp = a.partial + ((long)(z * a.tail_items + slot) * rows + row) * a.BN + cc * 8;
    const float4 p0 = reinterpret_cast<const float4*>(pp)[0], p1 = reinterpret_cast<const float4*>(pp)[1];
    f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w; f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
  }
  finish_store(a, n_img, pix, (long)n_img * hw + pix, c8, f);
}


// ---- fused separable convolution: depthwise 3x3 (+BN) computed INTO the pointwise GEMM's A operand -------------------------------
// For the HBM-bound separable convolutions of the Xception entry flow (193 x 193 x 128: 763 MB per tensor and 40 crops) the
// depthwise output is the largest tensor of the layer pair and lives only to be read back by the pointwise GEMM.  Here it never
// leaves the SM: per 16 x 8 pixel tile and 32-channel k-block
//   warp 0        TMA: the fp32 (F8) input tile + halo, box {10 px * 8 ch, 18 rows, 4 chunks}, zero fill = SAME padding; the packed
//                 pointwise weights of the k-block (bulk copy);
//   warps 4-11    depthwise: 9 LDS.128 + 36 FMA per 4 channels of a pixel (BatchNorm folded, optional ReLU in / out), split into
//                 bf16 hi / lo and written straight into the A stage in the un-swizzled K-major layout [4 chunks][128 pixels][8]
//                 (generic-proxy writes, made visible to the tensor core with fence.proxy.async);
//   warp 1        three tcgen05.mma per 16 channels into one of two TMEM accumulators;
//   warps 12-19   epilogue of the previous tile (bias, ReLU, F8 and / or CP8 stores) while the next one is produced.
// A layer pair with ONE output-channel tile (Cout <= 128) is eligible: the depthwise tile is computed once.  The 728-channel
// middle flow is not (three channel tiles = three recomputations, and it is tensor-bound, not HBM-bound).
constexpr int SF_THREADS = 640, SF_DW_WARPS = 8, SF_EPI_WARPS = 8;
constexpr int SF_XS = 4, SF_AS = 4, SF_WS = 3;
constexpr int SF_HW = 10, SF_HH = 18;                              // halo tile of a 16 x 8 output tile
constexpr int SF_X_BYTES = 4 * SF_HH * SF_HW * 32;                 // 4 chunks x 18 x 10 pixels x 8 fp32 = 23 040
constexpr int SF_A_PLANE = 4 * 128 * 16;                           // one bf16 plane of an A stage: [4 chunks][128 rows][8] = 8 192
constexpr int SF_A_BYTES = 2 * SF_A_PLANE;

struct SepFusedArgs {
  const float* dw_w;     // [9][cpad], BatchNorm scale folded
  const float* dw_b;     // [cpad]
  const __nv_bfloat16* pw_w;   // packed [kblock][hi|lo][4][BN][8] (pack_conv_weights_umma, 1x1, KC = 4, one channel tile)
  const float* pw_b;     // [BN]
  float* out_f8; int f8_chunks, f8_c0;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int out_chunks, out_c0;
  int H, W, in_c0, chunks, cpad, kblocks, BN, Cout;
  int pre_relu, post_relu;   // around the depthwise convolution
  float slope;               // after the pointwise convolution (1 = identity, 0 = ReLU)
  int tiles_x, tiles_y, n_active, total;
  int w_stage;               // bytes of one pointwise weight stage = 2 * 4 * BN * 16
  int dbg;                   // PREMVOS_DBG ablations: 1 = depthwise producers skip the arithmetic, 2 = epilogue drains TMEM only
};

__global__ void __launch_bounds__(SF_THREADS, 1) sepconv_fused_kernel(const __grid_constant__ CUtensorMap tm_in, const SepFusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* x_smem = smem;                                   // [XS][4][18][10][8] fp32
  uint8_t* a_smem = x_smem + SF_XS * SF_X_BYTES;             // [AS][hi | lo][4][128][8] bf16
  uint8_t* w_smem = a_smem + SF_AS * SF_A_BYTES;             // [WS][hi | lo][4][BN][8] bf16
  float* dww = reinterpret_cast<float*>(w_smem + SF_WS * a.w_stage);   // [9][cpad]
  float* dwb = dww + 9 * a.cpad;                                        // [cpad]
  float* pwb = dwb + a.cpad;                                            // [BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(pwb + a.BN);
  uint64_t *x_full = bars, *x_empty = bars + SF_XS, *a_full = bars + 2 * SF_XS, *a_empty = a_full + SF_AS, *w_full = a_empty + SF_AS,
           *w_empty = w_full + SF_WS, *t_full = w_empty + SF_WS, *t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = a.tiles_x * a.tiles_y;

  for (int i = threadIdx.x; i < 10 * a.cpad; i += SF_THREADS) dww[i] = i < 9 * a.cpad ? a.dw_w[i] : a.dw_b[i - 9 * a.cpad];
  for (int i = threadIdx.x; i < a.BN; i += SF_THREADS) pwb[i] = a.pw_b[i];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_in);
    for (int s = 0; s < SF_XS; s++) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], SF_DW_WARPS); }
    for (int s = 0; s < SF_AS; s++) { mbar_init(&a_full[s], SF_DW_WARPS); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SF_WS; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], SF_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2u * (uint32_t)a.BN < 32u ? 32u : (2u * (uint32_t)a.BN <= 64u ? 64u : (2u * (uint32_t)a.BN <= 128u ? 128u : 256u)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===== TMA producer =====
    uint32_t xs = 0, xph = 0, ws = 0, wph = 0;
    for (int t = blockIdx.x; t < a.total; t += gridDim.x) {
      const int tile = t % tiles_per_img, n = t / tiles_per_img;
      const int ty0 = (tile / a.tiles_x) * 16, tx0 = (tile % a.tiles_x) * 8;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.pw_w);
      for (int kb = 0; kb < a.kblocks; kb++) {
        mbar_wait(&x_empty[xs], xph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&x_full[xs], SF_X_BYTES);
          tma_load_4d(&tm_in, &x_full[xs], x_smem + (size_t)xs * SF_X_BYTES, (tx0 - 1) * 8, ty0 - 1, a.in_c0 + kb * 4, n);
        }
        __syncwarp();
        if (++xs == SF_XS) { xs = 0; xph ^= 1u; }
        mbar_wait(&w_empty[ws], wph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&w_full[ws], (uint32_t)a.w_stage);
          bulk_load_1d(w_smem + (size_t)ws * a.w_stage, wsrc, (uint32_t)a.w_stage, &w_full[ws]);
        }
        __syncwarp();
        wsrc += a.w_stage;
        if (++ws == SF_WS) { ws = 0; wph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_bf16(128, a.BN);
    const uint32_t a_hi32 = (128u >> 4) | (1u << 14), w_hi32 = (128u >> 4) | (1u << 14);           // SBO = 128 B
    const uint32_t a_lo32 = (2048u >> 4) << 16, w_lo32 = (((uint32_t)a.BN * 16u) >> 4) << 16;         // LBO: next 8 channels
    const uint32_t a_base = smem_u32(a_smem), w_base = smem_u32(w_smem);
    const uint32_t w_plane16 = ((uint32_t)a.w_stage / 2u) >> 4, a_plane16 = (uint32_t)SF_A_PLANE >> 4;
    const uint32_t a_k16 = (2u * 2048u) >> 4, w_k16 = (2u * (uint32_t)a.BN * 16u) >> 4;
    uint32_t as = 0, aph = 0, ws = 0, wph = 0, buf = 0, eph = 0;
    for (int t = blockIdx.x; t < a.total; t += gridDim.x) {
      mbar_wait(&t_empty[buf], ((eph >> buf) & 1u) ^ 1u);
      eph ^= 1u << buf;
      tc_fence_after();
      const uint32_t d = tmem_base + buf * (uint32_t)a.BN;
      uint32_t accum = 0;
      for (int kb = 0; kb < a.kblocks; kb++) {
        mbar_wait(&a_full[as], aph);
        mbar_wait(&w_full[ws], wph);
        tc_fence_after();
        if (elect_one()) {
          uint32_t aH = a_lo32 + ((a_base + as * (uint32_t)SF_A_BYTES) >> 4), wH = w_lo32 + ((w_base + ws * (uint32_t)a.w_stage) >> 4);
#pragma unroll
          for (int ks = 0; ks < 2; ks++, aH += a_k16, wH += w_k16) {
            umma_bf16_lo(d, aH + a_plane16, a_hi32, wH, w_hi32, idesc, accum);           // lo * hi
            umma_bf16_lo(d, aH, a_hi32, wH + w_plane16, w_hi32, idesc, 1u);              // hi * lo
            umma_bf16_lo(d, aH, a_hi32, wH, w_hi32, idesc, 1u);                          // hi * hi
            accum = 1u;
          }
          umma_commit(&a_empty[as]);
          umma_commit(&w_empty[ws]);
        }
        accum = 1u;
        __syncwarp();
        if (++as == SF_AS) { as = 0; aph ^= 1u; }
        if (++ws == SF_WS) { ws = 0; wph ^= 1u; }
      }
      if (elect_one()) umma_commit(&t_full[buf]);
      __syncwarp();
      buf ^= 1u;
    }
  } else if (warp >= 4 && warp < 4 + SF_DW_WARPS) {
    // ===== depthwise producers: 256 threads; thread = (chunk q of the k-block, strip of 4 rows, column, half of the chunk): its 9 tap
    // weights stay in registers for the whole k-block, a 6 x 3 window of LDS.128 feeds 4 vertically adjacent outputs, and the 16
    // lanes of a half warp read one 256-byte tile row (conflict-free) =====
    const int dt = (warp - 4) * 32 + lane;
    const int half = dt & 1, tx = (dt >> 1) & 7, sr = (dt >> 4) & 3, q = dt >> 6;
    uint32_t xs = 0, xph = 0, as = 0, aph = 0;
    for (int t = blockIdx.x; t < a.total; t += gridDim.x) {
      for (int kb = 0; kb < a.kblocks; kb++) {
        const int c4 = (kb * 4 + q) * 8 + half * 4;
        float4 wv[9];
#pragma unroll
        for (int k = 0; k < 9; k++) wv[k] = *reinterpret_cast<const float4*>(dww + k * a.cpad + c4);
        const float4 bv = *reinterpret_cast<const float4*>(dwb + c4);
        mbar_wait(&x_full[xs], xph);
        mbar_wait(&a_empty[as], aph ^ 1u);
        const float4* xp = reinterpret_cast<const float4*>(x_smem + (size_t)xs * SF_X_BYTES) + ((q * SF_HH + sr * 4) * SF_HW + tx) * 2 + half;
        uint8_t* dst = a_smem + (size_t)as * SF_A_BYTES + q * 2048 + ((sr * 4) * 8 + tx) * 16 + half * 8;
        if (!(a.dbg & 1)) {
          float4 acc[4] = {bv, bv, bv, bv};
#pragma unroll
          for (int r = 0; r < 6; r++) {
            float4 v[3];
#pragma unroll
            for (int s2 = 0; s2 < 3; s2++) {
              v[s2] = xp[(r * SF_HW + s2) * 2];
              if (a.pre_relu) { v[s2].x = fmaxf(v[s2].x, 0.f); v[s2].y = fmaxf(v[s2].y, 0.f); v[s2].z = fmaxf(v[s2].z, 0.f); v[s2].w = fmaxf(v[s2].w, 0.f); }
            }
#pragma unroll
            for (int o = 0; o < 4; o++) {
              const int tr = r - o;   // tap row of output o fed by window row r
              if (tr < 0 || tr > 2) continue;
#pragma unroll
              for (int s2 = 0; s2 < 3; s2++) {
                const float4 w4 = wv[tr * 3 + s2];
                acc[o].x = fmaf(v[s2].x, w4.x, acc[o].x); acc[o].y = fmaf(v[s2].y, w4.y, acc[o].y);
                acc[o].z = fmaf(v[s2].z, w4.z, acc[o].z); acc[o].w = fmaf(v[s2].w, w4.w, acc[o].w);
              }
            }
          }
#pragma unroll
          for (int o = 0; o < 4; o++) {
            float4 f = acc[o];
            if (a.post_relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(f.x, f.y), h1 = __floats2bfloat162_rn(f.z, f.w);
            const uint32_t hw0 = *reinterpret_cast<const uint32_t*>(&h0), hw1 = *reinterpret_cast<const uint32_t*>(&h1);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(f.x - __uint_as_float(hw0 << 16), f.y - __uint_as_float(hw0 & 0xffff0000u));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(f.z - __uint_as_float(hw1 << 16), f.w - __uint_as_float(hw1 & 0xffff0000u));
            *reinterpret_cast<uint2*>(dst + o * 128) = make_uint2(hw0, hw1);
            *reinterpret_cast<uint2*>(dst + o * 128 + SF_A_PLANE) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) { mbar_arrive(&a_full[as]); mbar_arrive(&x_empty[xs]); }
        if (++xs == SF_XS) { xs = 0; xph ^= 1u; }
        if (++as == SF_AS) { as = 0; aph ^= 1u; }
      }
    }
  } else if (warp >= 4 + SF_DW_WARPS) {
    // ===== epilogue: two warps per TMEM lane quarter, each one half of the accumulator columns =====
    const int ew = warp - (4 + SF_DW_WARPS), q = warp & 3, part = ew >> 2;
    const int m = q * 32 + lane;
    const int col_span = a.BN / 2, col_begin = part * col_span, col_end = col_begin + col_span;
    const long hw = (long)a.H * a.W;
    uint32_t buf = 0, fph = 0;
    for (int t = blockIdx.x; t < a.total; t += gridDim.x) {
      const int tile = t % tiles_per_img, n = t / tiles_per_img;
      const int oy = (tile / a.tiles_x) * 16 + (m >> 3), ox = (tile % a.tiles_x) * 8 + (m & 7);
      const bool in_img = oy < a.H && ox < a.W;
      const long pix = (long)oy * a.W + ox;
      mbar_wait(&t_full[buf], (fph >> buf) & 1u);
      fph ^= 1u << buf;
      tc_fence_after();
      for (int c0 = col_begin; c0 < col_end; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)a.BN + (uint32_t)c0, v);
        tmem_ld_wait();
        if (c0 + 16 >= col_end) {   // all columns of this warp are in registers: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
        if (!in_img || c0 >= a.Cout || (a.dbg & 2)) continue;
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const float x = __uint_as_float(v[j]) + pwb[c0 + j];
          f[j] = x > 0.f ? x : x * a.slope;
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (c0 + 8 * h >= a.Cout) continue;
          if (a.out_f8) {
            float* pf = a.out_f8 + (((long)n * a.f8_chunks + a.f8_c0 + (c0 >> 3) + h) * hw + pix) * 8;
            reinterpret_cast<float4*>(pf)[0] = make_float4(f[8 * h], f[8 * h + 1], f[8 * h + 2], f[8 * h + 3]);
            reinterpret_cast<float4*>(pf)[1] = make_float4(f[8 * h + 4], f[8 * h + 5], f[8 * h + 6], f[8 * h + 7]);
          }
          if (a.out_hi) {
            uint32_t hw4[4], lw4[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float x0 = f[8 * h + 2 * j], x1 = f[8 * h + 2 * j + 1];
              const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
              hw4[j] = *reinterpret_cast<const uint32_t*>(&hh);
              const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __uint_as_float(hw4[j] << 16), x1 - __uint_as_float(hw4[j] & 0xffff0000u));
              lw4[j] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            const long oi = (((long)n * a.out_chunks + a.out_c0 + (c0 >> 3) + h) * hw + pix) * 8;
            *reinterpret_cast<uint4*>(a.out_hi + oi) = make_uint4(hw4[0], hw4[1], hw4[2], hw4[3]);
            *reinterpret_cast<uint4*>(a.out_lo + oi) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
          }
        }
      }
      buf ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2u * (uint32_t)a.BN < 32u ? 32u : (2u * (uint32_t)a.BN <= 64u ? 64u : (2u * (uint32_t)a.BN <= 128u ? 128u : 256u)));
}

// ---- CTA-pair kernel for 1x1 / stride-1 layers -------------------------------------------------------------------------
// A 1x1 layer moves 9x more A bytes per MMA than a 3x3 layer in halo mode; at 128 x 256 tiles one SM has to ingest 62.5 B/clk of
// operands to keep its tensor pipe busy, the L2 delivers ~43-46.  A CTA PAIR (cluster of 2 = the two SMs of a TPC) computes a
// 256 pixel x 256 channel tile with tcgen05.mma.cta_group::2: each CTA loads ITS 128 pixels of A and ITS 128 of the 256 weight
// rows, the MMA reads both shared memories and writes 128 accumulator rows into each CTA's tensor memory -- 41.7 B/clk per SM and
// still 256 TMEM columns per buffer, so the epilogue of item i overlaps the MMAs of item i+1 (tools/ubench/pair_mma.cu: the
// un-swizzled CP8 layouts are valid split operands, 128 cycles per 256 x 256 x 16 MMA = the tensor pipe's rate).
//
// Protocol (the one CUTLASS's 2-SM kernels use): both CTAs run a producer warp; every TMA (.cta_group::2) credits its bytes to the
// LEADER's full barrier, which the leader's producer arms with the bytes of both CTAs; the leader's issuer waits on it, issues the
// MMAs for the pair and commits (multicast) to the empty barrier of BOTH CTAs and, per item, to both tmem_full barriers; the
// epilogue warps of both CTAs arrive on the leader's tmem_empty barrier (the peer's over the cluster).
// Work item = (pixel-tile pair, 256-wide channel tile); CTA r of a pair owns pixel tile 2*mp + r of the list (image, tile), which
// may lie in another image than its peer's or beyond the end (zero rows, nothing stored).  Items are ordered channel-tile-minor:
// pairs running at the same time share their A tiles through the L2.  Tail items are split along K exactly as in conv_umma_kernel.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmW, const UmmaConvArgs a) {
  constexpr int EPI_WARPS = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  uint8_t* a_smem = smem;
  uint8_t* w_smem = a_smem + (size_t)a.w_stages * 2 * a.a_plane;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + (size_t)a.w_stages * a.w_stage);
  uint64_t* full = bars;                      // [MAX_W_STAGES]  (used in the leader CTA)
  uint64_t* empty = bars + MAX_W_STAGES;      // [MAX_W_STAGES]  (each CTA waits on its own copy)
  uint64_t* tmem_full_bar = empty + MAX_W_STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]  (used in the leader CTA, 2 * EPI_WARPS arrivals)
  uint32_t* tmem_addr_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  // The layer's whole bias vector lives in shared memory (<= 8 KB): with 192 KB of operand stages the L1 has no capacity left, so
  // a bias load in the epilogue is an L2 round trip per 16 columns (38 % of the epilogue's stall samples were waiting for it)
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
  for (int i = threadIdx.x; i < a.bias_smem; i += blockDim.x) bias_s[i] = __ldg(a.bias + i);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_per_img = a.tiles_y;
  // Schedule.  Phase 1: all but the last full round of items go round-robin, item t to pair t % P, so that the pairs running at the
  // same time work on neighbouring items = share their A tiles (same pixel tiles) and weight tiles in the L2 (an all-stream-K order
  // measured 10 % slower on the big layers: 74 pairs reading 74 different tiles).  Phase 2 ("stream-K" over the last round plus the
  // remainder, P <= items < 2P, or over everything when there are fewer items than pairs): the (item, k-block) units of those
  // items, item-major, are cut into P equal contiguous ranges, pair p takes range p as fragments (item, [kb_begin, kb_end)); a
  // range is at least half an item long (the host picks P), so an item is finished from at most two partial sums.  A fragment that does not start at k-block 0 is a CONTRIBUTOR: it is the first
  // fragment of its pair's range, stores raw fp32 partial sums into workspace slot `pair` and raises one flag per epilogue warp.
  // The fragment that starts at k-block 0 but ends early is the FINISHER: it is the last fragment of its pair's range, waits for the
  // flags of the following pair(s), adds their partial sums in pair order (deterministic) and runs the epilogue.  A pair runs its
  // contributor before anything it could wait for, and only ever waits for pairs with a higher index.
  const int K = a.kblocks;
  const int rounds = a.total_work / n_pairs;
  const int n_main = a.stream_k_all ? 0 : (a.total_work % n_pairs == 0 ? a.total_work : max(0, rounds - 1) * n_pairs);
  const long U = (long)(a.total_work - n_main) * K;
  auto range_begin = [&](int p) { return U * p / n_pairs; };
  const long u_end = range_begin(pair_id + 1);
  struct Work { int ntile, mp, item, kb_begin, kb_end; };
  struct Cursor { int t; long u; };
  auto next_fragment = [&](Cursor& c, Work& w) {
    if (c.t < n_main) {
      w.item = c.t; w.kb_begin = 0; w.kb_end = K;
      c.t += n_pairs;
    } else if (c.u < u_end) {
      const int ri = (int)(c.u / K);
      w.item = n_main + ri;
      w.kb_begin = (int)(c.u - (long)ri * K);
      w.kb_end = (int)min((long)K, (long)w.kb_begin + (u_end - c.u));
      c.u += w.kb_end - w.kb_begin;
    } else {
      return false;
    }
    w.mp = w.item / a.ntiles;
    w.ntile = w.item - w.mp * a.ntiles;
    return true;
  };
  const Cursor cursor0{pair_id, range_begin(pair_id)};
  const bool dbg_on = (a.dbg & 64) && (blockIdx.x == 0 || blockIdx.x == 41);
  __shared__ long long dbg_frag[16][4];
  int dbg_nf = 0;
  const long long dbg_c0 = clock64();
  const unsigned long long dbg_t0 = (a.dbg & 32) ? globaltimer_ns() : 0ull;
  __shared__ unsigned long long dbg_ev[4];
  if ((a.dbg & 32) && threadIdx.x < 4) dbg_ev[threadIdx.x] = 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmW);
    for (int s = 0; s < a.w_stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_addr_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_addr_slot, 0);

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    uint32_t st = 0, ph = 0;
    const uint32_t half_bytes = 2u * (uint32_t)a.a_box_bytes + 2u * (uint32_t)a.w_plane;   // what ONE CTA loads per k-block
    const int w_rows = (2 * a.w_plane) >> 7;                                             // 128-byte rows of one CTA's weight block
    Cursor cur = cursor0;
    Work wk;
    while (next_fragment(cur, wk)) {
      const int mtile = 2 * wk.mp + (int)rank;
      const int n_img = mtile / tiles_per_img, ty0 = (mtile - n_img * tiles_per_img) * 16;
      int wrow = ((wk.ntile * a.kblocks + wk.kb_begin) * 2 + (int)rank) * w_rows;
      for (int kb = wk.kb_begin; kb < wk.kb_end; kb++, wrow += 2 * w_rows) {
        mbar_wait(&empty[st], ph ^ 1u);
        if (elect_one()) {
          const uint32_t fb = mapa_u32(smem_u32(&full[st]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full[st], 2u * half_bytes);
          uint8_t* dst = a_smem + (size_t)st * 2 * a.a_plane;
          tma_load_4d_pair(&tmA_hi, fb, dst, 0, ty0, kb * a.KC, n_img);
          tma_load_4d_pair(&tmA_lo, fb, dst + a.a_plane, 0, ty0, kb * a.KC, n_img);
          tma_load_2d_pair(&tmW, fb, w_smem + (size_t)st * a.w_stage, 0, wrow);
        }
        __syncwarp();
        if (++st == (uint32_t)a.w_stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader CTA only) =====
    const uint32_t idesc = make_idesc_bf16(256, a.BN);
    const uint32_t lbo = 128u * 16u;   // both operands: [k chunk][128 rows][8 ch], SBO = 128 B
    const uint32_t hi32 = (128u >> 4) | (1u << 14), lo32 = (lbo >> 4) << 16;
    const uint32_t a_base = smem_u32(a_smem) & 0x3FFFFu, w_base = smem_u32(w_smem) & 0x3FFFFu;
    const uint32_t a_stage_bytes = 2u * (uint32_t)a.a_plane, w_stage_bytes = (uint32_t)a.w_stage;
    const uint32_t kstep16 = (2u * lbo) >> 4, a_plane16 = (uint32_t)a.a_plane >> 4, w_plane16 = (uint32_t)a.w_plane >> 4;
    const int ksteps = a.KC / 2;
    uint32_t st = 0, ph = 0, buf = 0, empty_ph = 0;
    Cursor cur = cursor0;
    Work wk;
    while (next_fragment(cur, wk)) {
      mbar_wait(&tmem_empty_bar[buf], ((empty_ph >> buf) & 1u) ^ 1u);
      empty_ph ^= 1u << buf;
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * 256u;
      uint32_t accum = 0;
      for (int kb = wk.kb_begin; kb < wk.kb_end; kb++) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          if ((a.dbg & 32) && dbg_ev[0] == 0) dbg_ev[0] = globaltimer_ns();
          uint32_t aH = lo32 + ((a_base + st * a_stage_bytes) >> 4), wH = lo32 + ((w_base + st * w_stage_bytes) >> 4);
          for (int ks = 0; ks < ((a.dbg & 2) ? 0 : ksteps); ks++, aH += kstep16, wH += kstep16) {
            umma_pair_lo(tmem_acc, aH + a_plane16, hi32, wH, hi32, idesc, accum);
            umma_pair_lo(tmem_acc, aH, hi32, wH + w_plane16, hi32, idesc, 1u);
            umma_pair_lo(tmem_acc, aH, hi32, wH, hi32, idesc, 1u);
            accum = 1u;
          }
          umma_commit_pair(&empty[st]);
        }
        accum = 1u;
        if (++st == (uint32_t)a.w_stages) { st = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit_pair(&tmem_full_bar[buf]);
      __syncwarp();
      if ((a.dbg & 32) && lane == 0) dbg_ev[1] = globaltimer_ns();
      buf ^= 1u;
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): 128 accumulator rows x BN columns of this CTA =====
    const int q = warp & 3, m = q * 32 + lane;
    const long hw = (long)a.flat_hw;
    const uint32_t te_leader = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);   // [buf] at +8 bytes
    uint32_t buf = 0, full_ph = 0;
    const int col_part = (warp - 4) >> 2, col_span = a.BN / 2;
    const int col_begin = col_part * col_span, col_end = col_begin + col_span;
    const int ew = warp - 4;   // epilogue warp 0..7 of this CTA
    Cursor cur = cursor0;
    Work wk;
    while (next_fragment(cur, wk)) {
      const int mtile = 2 * wk.mp + (int)rank;
      const int n_img = mtile / tiles_per_img;
      const long pix = (long)(mtile - n_img * tiles_per_img) * 128 + m;
      const bool in_img = n_img < a.n_images && pix < hw;
      const bool contributor = wk.kb_begin > 0;
      // finisher: the pairs after this one whose range starts inside this item hold the rest of its k-blocks
      int nq = 0;
      if (!contributor && wk.kb_end < K) {
        const long item_end = (long)(wk.item - n_main + 1) * K;
        while (pair_id + 1 + nq < n_pairs && range_begin(pair_id + 1 + nq) < item_end) nq++;
      }
      long long dbg_f0 = 0, dbg_f1 = 0;
      if (dbg_on && warp == 4 && lane == 0) dbg_f0 = clock64();
      mbar_wait(&tmem_full_bar[buf], (full_ph >> buf) & 1u);
      full_ph ^= 1u << buf;
      tc_fence_after();
      if (nq > 0) {   // the partial sums this warp will add were written by the same (rank, warp) of the contributing pairs
        if (lane == 0) {
          for (int j = 0; j < nq; j++) {
            unsigned* flag = a.sk_flags + ((pair_id + 1 + j) * 2 + (int)rank) * 8 + ew;
            uint32_t spins = 0;
            while (ld_acquire_u32(flag) == 0u) {
              if (++spins > (1u << 24)) __trap();
              __nanosleep(64);
            }
            *flag = 0u;   // self-cleaning: the next launch (stream order) finds it lowered
          }
        }
        __syncwarp();
      }
      if (dbg_on && warp == 4 && lane == 0) dbg_f1 = clock64();
      const uint32_t tmem_acc = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
      // Column loop, 16 accumulator columns (= two 8-channel chunks) per step.  The TMEM load of step g+1 is in flight while step g
      // is processed (two register sets), addresses are running pointers (one 64-bit multiply per item, not per store), the hi/lo
      // split converts two values per instruction (F2FP.PACK_AB; the scalar F2F conversions run at a quarter of that rate).
      const long row = ((long)n_img * hw + pix);                          // pixel index inside its plane stack
      const long chunk_stride = hw * 8;                                     // elements between two chunk planes
      const int cg0 = (wk.ntile * a.BN + col_begin) >> 3;                   // first 8-channel chunk of this warp
      const long o_off = (((long)n_img * a.out_chunks + a.out_c0 + cg0) * hw + pix) << 3;
      __nv_bfloat16* ohi = a.out_hi + o_off;
      __nv_bfloat16* olo = a.out_lo + o_off;
      const __nv_bfloat16* rhi = a.res_hi ? a.res_hi + ((((long)n_img * a.res_chunks + a.res_c0 + cg0) * hw + pix) << 3) : nullptr;
      const __nv_bfloat16* rlo = a.res_hi ? a.res_lo + (rhi - a.res_hi) : nullptr;
      float* of8 = a.out_f8 ? a.out_f8 + ((((long)n_img * a.f8_chunks + a.f8_c0 + cg0) * hw + pix) << 3) : nullptr;
      const float* rf8 = a.res_f8 ? a.res_f8 + ((((long)n_img * a.resf_chunks + a.resf_c0 + cg0) * hw + pix) << 3) : nullptr;
      const float* bias = (a.bias_smem ? bias_s : a.bias) + wk.ntile * a.BN + col_begin;
      // partial sums: [pair slot][rank][column / 4 (64)][row (128)] float4 -- a warp's 32 rows of one float4 column are contiguous
      float4* ppart = reinterpret_cast<float4*>(a.partial) + (((long)pair_id * 2 + rank) * 64 + (col_begin >> 2)) * 128 + m;
      (void)row;
      uint32_t va[16], vb[16];
      __syncwarp();
      tmem_ld16(tmem_acc + (uint32_t)col_begin, va);
      auto process = [&](uint32_t (&v)[16], int c0) {
        const int co0 = wk.ntile * a.BN + c0, g = (c0 - col_begin) >> 4;
        if (contributor) {   // raw partial sums into this pair's workspace slot, [pair][rank][128 rows][256]
          float4* pp = ppart + (g << 2) * 128;
#pragma unroll
          for (int j = 0; j < 4; j++)
            __stcg(pp + j * 128, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                       __uint_as_float(v[4 * j + 3])));
          return;
        }
        if (!in_img || co0 >= a.Cout) return;
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; j++) f[j] = __uint_as_float(v[j]);
        for (int jq = 0; jq < nq; jq++) {   // finisher: partial sums of the following pairs, in pair order (deterministic)
          const float4* pp = reinterpret_cast<const float4*>(a.partial) + (((long)(pair_id + 1 + jq) * 2 + rank) * 64 + (c0 >> 2)) * 128 + m;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 t4 = __ldcg(pp + i * 128);
            f[4 * i] += t4.x; f[4 * i + 1] += t4.y; f[4 * i + 2] += t4.z; f[4 * i + 3] += t4.w;
          }
        }
        const float4* bp = reinterpret_cast<const float4*>(bias + (g << 4));
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float4 b4 = bp[j];
          f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
        }
        const bool second = co0 + 8 < a.Cout;   // the second chunk of this step exists
        if (rhi) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (h == 1 && !second) break;
            const uint4 rh = *reinterpret_cast<const uint4*>(rhi + (2 * g + h) * chunk_stride);
            const uint4 rl = *reinterpret_cast<const uint4*>(rlo + (2 * g + h) * chunk_stride);
            const uint32_t hh[4] = {rh.x, rh.y, rh.z, rh.w}, ll[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              f[8 * h + 2 * j] += __uint_as_float(hh[j] << 16) + __uint_as_float(ll[j] << 16);
              f[8 * h + 2 * j + 1] += __uint_as_float(hh[j] & 0xffff0000u) + __uint_as_float(ll[j] & 0xffff0000u);
            }
          }
        }
        if (rf8) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (h == 1 && !second) break;
            const float4* rp = reinterpret_cast<const float4*>(rf8 + (2 * g + h) * chunk_stride);
            const float4 r0 = rp[0], r1 = rp[1];
            f[8 * h] += r0.x; f[8 * h + 1] += r0.y; f[8 * h + 2] += r0.z; f[8 * h + 3] += r0.w;
            f[8 * h + 4] += r1.x; f[8 * h + 5] += r1.y; f[8 * h + 6] += r1.z; f[8 * h + 7] += r1.w;
          }
        }
        if (a.slope != 1.f) {   // LeakyReLU / ReLU: max(f, slope * f) for 0 <= slope < 1
#pragma unroll
          for (int j = 0; j < 16; j++) f[j] = fmaxf(f[j], f[j] * a.slope);
        }
        if (of8) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (h == 1 && !second) break;
            float4* pf8 = reinterpret_cast<float4*>(of8 + (2 * g + h) * chunk_stride);
            pf8[0] = make_float4(f[8 * h], f[8 * h + 1], f[8 * h + 2], f[8 * h + 3]);
            pf8[1] = make_float4(f[8 * h + 4], f[8 * h + 5], f[8 * h + 6], f[8 * h + 7]);
          }
        }
        if (!a.out_hi) return;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (h == 1 && (!second || co0 + 8 >= a.cp_cout)) break;
          uint32_t hw4[4], lw4[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float x0 = f[8 * h + 2 * j], x1 = f[8 * h + 2 * j + 1];
            const __nv_bfloat162 hp = __floats2bfloat162_rn(x0, x1);
            const uint32_t hwd = *reinterpret_cast<const uint32_t*>(&hp);
            const __nv_bfloat162 lp = __floats2bfloat162_rn(x0 - __uint_as_float(hwd << 16), x1 - __uint_as_float(hwd & 0xffff0000u));
            hw4[j] = hwd; lw4[j] = *reinterpret_cast<const uint32_t*>(&lp);
          }
          *reinterpret_cast<uint4*>(ohi + (2 * g + h) * chunk_stride) = make_uint4(hw4[0], hw4[1], hw4[2], hw4[3]);
          *reinterpret_cast<uint4*>(olo + (2 * g + h) * chunk_stride) = make_uint4(lw4[0], lw4[1], lw4[2], lw4[3]);
        }
      };
      for (int c0 = col_begin; c0 < col_end; c0 += 32) {
        tmem_ld_wait();                                                    // va (columns c0 ..) has landed
        __syncwarp();
        if (c0 + 16 < col_end) tmem_ld16(tmem_acc + (uint32_t)(c0 + 16), vb);
        else {   // every column of this warp is in registers: hand the accumulator buffer back to the issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(te_leader + buf * 8u);
        }
        if (!(a.dbg & 4)) process(va, c0);
        if (c0 + 16 >= col_end) break;
        tmem_ld_wait();
        __syncwarp();
        if (c0 + 32 < col_end) tmem_ld16(tmem_acc + (uint32_t)(c0 + 32), va);
        else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(te_leader + buf * 8u);
        }
        if (!(a.dbg & 4)) process(vb, c0 + 16);
      }
      if (contributor) {   // this warp's share of the partial sums is written: publish it
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_u32(a.sk_flags + (pair_id * 2 + (int)rank) * 8 + ew, 1u);
      }
      buf ^= 1u;
      if ((a.dbg & 32) && lane == 0) dbg_ev[2] = max(dbg_ev[2], globaltimer_ns());
      if (dbg_on && warp == 4 && lane == 0 && dbg_nf < 16) {
        dbg_frag[dbg_nf][0] = dbg_f0 - dbg_c0; dbg_frag[dbg_nf][1] = dbg_f1 - dbg_c0; dbg_frag[dbg_nf][2] = clock64() - dbg_c0;
        dbg_frag[dbg_nf][3] = wk.item * 1000 + wk.kb_begin * 10 + (contributor ? 1 : 0) + (nq ? 2 : 0);
        dbg_nf++;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody frees TMEM / exits while the peer's MMAs, TMA credits or barrier arrivals may still be in flight
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
  if ((a.dbg & 32) && threadIdx.x == 0) {
    const unsigned long long t1 = globaltimer_ns(), d = t1 - dbg_t0;
    atomicMin(&g_dbg_cta[0], dbg_t0); atomicMax(&g_dbg_cta[1], t1); atomicAdd(&g_dbg_cta[2], d);
    atomicMax(&g_dbg_cta[3], d); atomicMin(&g_dbg_cta[4], d); atomicAdd(&g_dbg_cta[5], 1ull);
    if (blockIdx.x < 512) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long* rr = g_dbg_rec[blockIdx.x];
      rr[0] = smid; rr[1] = dbg_t0; rr[2] = t1; rr[3] = dbg_ev[0]; rr[4] = dbg_ev[1]; rr[5] = dbg_ev[2];
    }
  }
  if (dbg_on && warp == 4 && lane == 0)
    for (int i = 0; i < dbg_nf; i++)
      printf("cta %d fragment %d (item*1000 + kb_begin*10 + contributor + 2*finisher = %lld): accumulator wait from %lld, flags done %lld, epilogue done %lld cycles\n",
             (int)blockIdx.x, i, dbg_frag[i][3], dbg_frag[i][0], dbg_frag[i][1], dbg_frag[i][2]);
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    PV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    PV_CHECK(p && q == cudaDriverEntryPointSuccess, PREMVOS_ERR_NO_DEVICE, "cuTensorMapEncodeTiled is not available in this driver");
    fn = (EncodeTiledFn)p;
  }
  *out = fn;
  return 0;
}

int encode_map(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* estr) {
  EncodeTiledFn fn = nullptr;
  PV_TRY(get_encode_fn(&fn));
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PV_CHECK(r == CUDA_SUCCESS, PREMVOS_ERR_INVALID_ARG,
           "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu, box %u %u %u %u)", (int)r, rank,
           (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3],
           box[0], box[1], box[2], box[3]);
  return 0;
}

}  // namespace

int encode_tensor_map_bf16(void* map128, const __nv_bfloat16* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  cuuint64_t d[5] = {1, 1, 1, 1, 1}, s[4] = {0, 0, 0, 0};
  cuuint32_t b[5] = {1, 1, 1, 1, 1}, e[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; i++) s[i] = strides_bytes[i];
  return encode_map((CUtensorMap*)map128, (void*)base, rank, d, s, b, e);
}

int encode_tensor_map_f32(void* map128, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = nullptr;
  PV_TRY(get_encode_fn(&fn));
  cuuint64_t d[5] = {1, 1, 1, 1, 1}, s[4] = {0, 0, 0, 0};
  cuuint32_t b[5] = {1, 1, 1, 1, 1}, e[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; i++) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; i++) s[i] = strides_bytes[i];
  CUresult r = fn((CUtensorMap*)map128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PV_CHECK(r == CUDA_SUCCESS, PREMVOS_ERR_INVALID_ARG, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u)",
           (int)r, rank, (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2], b[0], b[1], b[2]);
  return 0;
}

namespace {

inline void split_bf16(float x, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  *hi = __float2bfloat16_rn(x);
  *lo = __float2bfloat16_rn(x - __bfloat162float(*hi));
}

}  // namespace

// Channels per k-block (KC chunks of 8).  Bigger stages amortise the fixed cost of a TMA request; halo-mode
// 3x3 layers keep the A box small instead (its halo is loaded once per k-block and serves all taps).
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

int pack_conv_weights_umma(ConvWeightsUmma* out, const float* host_w, const float* host_b, int Cout, int Cin, int R, int S,
                           const int* cin_map, int cin_phys, int kc_hint, long m_hint, bool allow_pair) {
  out->R = R; out->S = S; out->Cin = Cin; out->Cout = Cout;
  const int phys = cin_map ? cin_phys : Cin;   // physical input channels (after the view's chunk padding)
  out->CinPhys = phys;
  const int chunks = (phys + 7) / 8;
  // measured on B200 (tools/conv_sweep.py): 3x3 with a wide N tile runs best with 16-channel k-blocks and two
  // co-resident CTAs per SM; narrow N tiles and 1x1 layers want longer k-blocks (fewer barrier handshakes)
  int kc = kc_hint > 0 ? kc_hint : (R * S > 1 ? 2 : (phys <= 160 ? 2 : 4));
  kc = env_int("PREMVOS_KC", kc);
  if (kc > round_up(chunks, 2)) kc = round_up(chunks, 2);
  PV_CHECK(kc >= 2 && (kc % 2) == 0 && kc <= 16, PREMVOS_ERR_INVALID_ARG, "pack_conv_weights_umma: KC=%d", kc);
  out->KC = kc;
  out->kblocks = (chunks + kc - 1) / kc;
  // N tile: 128 columns, or 256 for 1x1 layers whose output (m_hint pixels) still fills the machine with 256 x 256 tiles: a
  // 1x1 layer moves 9x more A bytes per MMA than a 3x3 layer in halo mode, so at 128 x 128 it is bound by the L2 -> shared
  // memory fill and by the per-k-block barrier round trip; 256 x 256 halves the bytes and quarters the handshakes per FLOP
  int bn = round_up(Cout, 16);
  int bn_cap = 128;
  if (R * S == 1 && Cout >= 256 && m_hint > 0 && (m_hint / 256) * ((Cout + 255) / 256) >= 120) bn_cap = 256;
  // CTA-pair kernel (conv_pair_kernel, cta_group::2, stream-K): layers the caller declares flat (1x1, stride 1, no padding) with at
  // least 256 output channels and enough (item, k-block) units to give every one of the 74 pairs a few
  bool pair = false;
  if (allow_pair && R * S == 1 && Cout >= 256 && m_hint > 0 && kc_hint == 0 && env_int("PREMVOS_PAIR", 1) != 0 && env_int("PREMVOS_KC", 0) == 0 &&
      env_int("PREMVOS_BN", 0) == 0) {
    const int kc4 = std::min(4, round_up(chunks, 2));
    const long units = ((m_hint + 255) / 256) * ((Cout + 255) / 256) * ((chunks + kc4 - 1) / kc4);
    pair = kc4 == 4 && units >= env_int("PREMVOS_PAIR_MIN_UNITS", 296);
  }
  if (pair) bn_cap = 256;
  bn_cap = env_int("PREMVOS_BN", bn_cap);
  if (bn_cap != 256 || R * S != 1) bn_cap = std::min(bn_cap, 128);
  if (bn > bn_cap) bn = bn_cap;
  if (bn > 128 && kc_hint == 0 && env_int("PREMVOS_KC", 0) == 0) { kc = std::min(4, round_up(chunks, 2)); out->KC = kc; out->kblocks = (chunks + kc - 1) / kc; }
  out->BN = bn;
  out->ntiles = (Cout + bn - 1) / bn;
  out->pair = (pair && bn == 256 && kc == 4) ? 1 : 0;
  const int KP = out->kblocks * kc * 8;
  const int taps = R * S;
  const size_t plane_elems = (size_t)kc * bn * 8;   // one (tap, k-block) operand image of one plane
  const size_t n = (size_t)out->ntiles * out->kblocks * taps * 2 * plane_elems;
  std::vector<__nv_bfloat16> w(n, __float2bfloat16_rn(0.f));
  std::vector<float> b((size_t)out->ntiles * bn, 0.f);
  for (int co = 0; co < Cout; co++) {
    b[co] = host_b ? host_b[co] : 0.f;
    const int nt = co / bn, row = co - nt * bn;
    for (int ci = 0; ci < Cin; ci++) {
      const int pc = cin_map ? cin_map[ci] : ci;
      if (pc < 0 || pc >= KP) return fail(PREMVOS_ERR_INVALID_ARG, "pack_conv_weights_umma: channel map out of range");
      const int kb = pc / (kc * 8), kcc = (pc / 8) % kc, e = pc & 7;
      for (int t = 0; t < taps; t++) {
        // [ntile][kblock][tap][plane][KC][BN][8]
        size_t base = ((((size_t)nt * out->kblocks + kb) * taps + t) * 2) * plane_elems + ((size_t)kcc * bn + row) * 8 + e;
        size_t lo_off = plane_elems;
        if (out->pair) {   // [ntile][kblock][CTA rank][plane][KC][128][8]
          const size_t half = plane_elems / 2;
          base = ((((size_t)nt * out->kblocks + kb) * 2 + row / 128) * 2) * half + ((size_t)kcc * 128 + row % 128) * 8 + e;
          lo_off = half;
        }
        split_bf16(host_w[((size_t)co * Cin + ci) * taps + t], &w[base], &w[base + lo_off]);
      }
    }
  }
  PV_CUDA(cudaMalloc((void**)&out->w, n * 2));
  PV_CUDA(cudaMalloc((void**)&out->bias, b.size() * sizeof(float)));
  PV_CUDA(cudaMemcpy(out->w, w.data(), n * 2, cudaMemcpyHostToDevice));
  PV_CUDA(cudaMemcpy(out->bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

void free_conv_plan_umma(ConvPlanUmma* plan) {
  if (plan->scratch) cudaFree(plan->scratch);
  plan->scratch = nullptr;
}

void free_conv_weights_umma(ConvWeightsUmma* w) {
  cudaFree(w->w); cudaFree(w->bias);
  w->w = nullptr; w->bias = nullptr;
}


constexpr int PAIR_SLOTS = 80;   // >= CTA pairs of a launch (74 on B200)
int conv_workspace_reserve(ConvWorkspace* ws) {
  if (ws->base) return 0;
  const size_t partial_bytes = (size_t)PAIR_SLOTS * 2 * 128 * 256 * sizeof(float), flag_bytes = (size_t)PAIR_SLOTS * 2 * 8 * sizeof(unsigned);
  PV_CUDA(cudaMalloc(&ws->base, partial_bytes + flag_bytes));
  PV_CUDA(cudaMemset((char*)ws->base + partial_bytes, 0, flag_bytes));   // flags are lowered again by the kernel that consumes them
  ws->partial = (float*)ws->base;
  ws->flags = (unsigned*)((char*)ws->base + partial_bytes);
  return 0;
}
void conv_workspace_free(ConvWorkspace* ws) {
  if (ws->base) cudaFree(ws->base);
  ws->base = nullptr; ws->partial = nullptr; ws->flags = nullptr;
}

// Plan of a CTA-pair launch (conv_pair_kernel): flattened 128-pixel tiles per image, pairs of tiles x 256-wide channel tiles.
static int plan_conv_pair(ConvPlanUmma* plan, const CView& in, const ConvOut& out, const ConvWeightsUmma& w, const ConvGeom& g, bool flat,
                          int real_hw, int cp_cout, int f32_first, ConvWorkspace* ws) {
  PV_CHECK(flat && w.BN == 256 && w.KC == 4, PREMVOS_ERR_INVALID_ARG,
           "conv_umma: weights were packed for the CTA-pair kernel, which needs a 1x1 / stride-1 / unpadded layer (BN %d KC %d)", w.BN, w.KC);
  PV_CHECK((out.cp.hi || out.f8.p) && !out.f32.p, PREMVOS_ERR_UNSUPPORTED, "conv_umma: the CTA-pair kernel writes CP8 / F8 outputs only");
  UmmaConvArgs& a = *reinterpret_cast<UmmaConvArgs*>(plan->args);
  const int geoH = (real_hw + 7) / 8;
  a.pair = 1;
  a.MT = 1; a.mt_horizontal = 0; a.halo = 0; a.merged_x = 1; a.lockstep = 1; a.NACC = 1; a.TPS = 1; a.nbuf = 2; a.epi_warps = 8;
  a.tiles_x = 1; a.tiles_y = (real_hw + 127) / 128;
  a.box_w = 8; a.box_h = 16;
  a.a_box_bytes = w.KC * 128 * 16;
  a.a_plane = a.a_box_bytes;
  a.w_plane = w.KC * 128 * 16;       // ONE CTA's half of a (hi or lo) weight plane
  a.w_stage = 2 * a.w_plane;
  const int stage_bytes = 2 * a.a_plane + a.w_stage;
  // short K loops only (epilogue-bound layers: -14 % at K = 256); long ones keep the seventh operand stage instead
  a.bias_smem = (w.ntiles * w.BN * 4 <= 8192 && w.kblocks <= 16 && env_int("PREMVOS_PAIR_BIAS_SMEM", 1)) ? w.ntiles * w.BN : 0;
  int st = (SMEM_LIMIT - 1024 - a.bias_smem * 4) / stage_bytes;
  st = std::max(2, std::min(st, MAX_W_STAGES));
  st = env_int("PREMVOS_PAIR_STAGES", st);
  PV_CHECK(st >= 2 && st <= MAX_W_STAGES, PREMVOS_ERR_INVALID_ARG, "conv_umma: PREMVOS_PAIR_STAGES=%d", st);
  a.a_stages = a.w_stages = st;
  plan->smem_bytes = st * stage_bytes + 128 + 512 + a.bias_smem * 4;
  PV_CHECK(plan->smem_bytes <= SMEM_LIMIT, PREMVOS_ERR_UNSUPPORTED, "conv_umma: %d bytes of shared memory needed", plan->smem_bytes);
  a.w = w.w; a.bias = w.bias; a.slope = g.slope;
  a.out_hi = out.cp.hi; a.out_lo = out.cp.lo; a.out_chunks = out.cp.chunks; a.out_c0 = out.cp.c0;
  a.out_f32 = out.f32.p; a.out_cs = out.f32.cs; a.out_coff = out.f32.coff;
  a.cp_cout = cp_cout; a.f32_first = f32_first; a.f32_linear = out.f32_linear ? 1 : 0; a.f32_accum = out.f32_accumulate ? 1 : 0;
  a.res_hi = out.res.hi; a.res_lo = out.res.lo; a.res_chunks = out.res.chunks; a.res_c0 = out.res.c0;
  a.out_f8 = out.f8.p; a.f8_chunks = out.f8.chunks; a.f8_c0 = out.f8.c0;
  a.res_f8 = out.res_f8.p; a.resf_chunks = out.res_f8.chunks; a.resf_c0 = out.res_f8.c0;
  a.dbg = env_int("PREMVOS_DBG", 0);
  a.tmem_cols = 512;
  a.ksplit = 1; a.kb_per = a.kblocks; a.cout_pad = w.ntiles * w.BN; a.partial = nullptr; a.partial_stride = 0;
  a.ntiles = w.ntiles; a.n_images = in.N;
  const int mtiles = in.N * a.tiles_y;
  a.total_work = ((mtiles + 1) / 2) * w.ntiles;
  plan->grid_x = a.total_work; plan->grid_y = 1; plan->grid_z = 1; plan->ctas_per_sm = 1;
  // stream-K workspace (partial sums + flags of up to PAIR_SLOTS pairs): the caller's (one per network: layers run one after the
  // other on one stream and reuse it, so it stays in the L2) or a private one
  a.tail_items = 0; a.tail_split = 1; a.tail_kb_per = a.kblocks; a.main_work = a.total_work;
  if (ws) {
    PV_TRY(conv_workspace_reserve(ws));
    a.partial = ws->partial; a.sk_flags = ws->flags;
  } else {
    ConvWorkspace own;
    PV_TRY(conv_workspace_reserve(&own));
    a.partial = own.partial; a.sk_flags = own.flags;
    plan->scratch = own.base;
  }
  // tensor maps: activations [N][chunks][ceil(HW/8)][64], weights [rows][64] (128-byte rows of the packed image)
  const int vchunks = (in.C + 7) / 8;
  const size_t plane_bytes = (size_t)in.H * in.W * 16;
  __nv_bfloat16* bases[2] = {in.hi + (size_t)in.c0 * in.H * in.W * 8, in.lo + (size_t)in.c0 * in.H * in.W * 8};
  CUtensorMap* maps[2] = {(CUtensorMap*)plan->map_a_hi, (CUtensorMap*)plan->map_a_lo};
  for (int k = 0; k < 2; k++) {
    cuuint64_t dims[4] = {64, (cuuint64_t)geoH, (cuuint64_t)vchunks, (cuuint64_t)in.N};
    cuuint64_t strides[3] = {128, plane_bytes, plane_bytes * in.chunks};
    cuuint32_t box[4] = {64, 16, (cuuint32_t)w.KC, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    PV_TRY(encode_map(maps[k], bases[k], 4, dims, strides, box, estr));
  }
  {
    const cuuint64_t rows = (cuuint64_t)w.ntiles * w.kblocks * 2 * (2 * a.w_plane / 128);
    cuuint64_t dims[4] = {64, rows, 1, 1};
    cuuint64_t strides[3] = {128, 0, 0};
    cuuint32_t box[4] = {64, (cuuint32_t)(2 * a.w_plane / 128), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    PV_TRY(encode_map((CUtensorMap*)plan->map_w, w.w, 2, dims, strides, box, estr));
  }
  const double opx = (double)in.N * real_hw;
  plan->flops = 2.0 * opx * (double)w.Cout * w.Cin;
  plan->bytes = 4.0 * ((double)in.N * in.H * in.W * w.Cin + opx * w.Cout + (double)w.Cin * w.Cout);
  plan->halo = 0; plan->MT = 1; plan->N = in.N;
  return 0;
}

// Plans one convolution launch: tensor maps of the input view are encoded here (host only, no device work).
int plan_conv_umma(ConvPlanUmma* plan, const CView& in, const ConvOut& out, const ConvWeightsUmma& w, const ConvGeom& g, ConvWorkspace* ws) {
  PV_CHECK(in.hi && in.lo, PREMVOS_ERR_INVALID_ARG, "conv_umma: null input view");
  PV_CHECK(round_up(in.C, 8) == round_up(w.CinPhys, 8), PREMVOS_ERR_INVALID_ARG,
           "conv_umma: input view has %d channels, weights were packed for %d", in.C, w.CinPhys);
  PV_CHECK(g.stride == 1 || g.stride == 2, PREMVOS_ERR_UNSUPPORTED, "conv_umma: stride %d", g.stride);
  const int Ho = (in.H + g.pad_t + g.pad_b - g.dil * (w.R - 1) - 1) / g.stride + 1;
  const int sx = g.stride_x > 0 ? g.stride_x : g.stride;   // horizontal stride; a layer with sx != stride runs in plain tap mode
  PV_CHECK(sx == 1 || sx == 2, PREMVOS_ERR_UNSUPPORTED, "conv_umma: horizontal stride %d", sx);
  const bool mixed = sx != g.stride;
  const int Wo = (in.W + g.pad_l + g.pad_r - g.dil * (w.S - 1) - 1) / sx + 1;
  PV_CHECK(Ho > 0 && Wo > 0, PREMVOS_ERR_INVALID_ARG, "conv_umma: empty output");
  const int cp_cout = out.cp.hi ? (out.cp_channels >= 0 ? out.cp_channels : w.Cout) : 0;
  const int f32_first = out.f32.p ? out.f32_first : 0;
  PV_CHECK(cp_cout <= w.Cout && (cp_cout == w.Cout || cp_cout % 8 == 0) && f32_first >= 0 && f32_first <= w.Cout, PREMVOS_ERR_INVALID_ARG,
           "conv_umma: bad channel routing (cp %d, f32 from %d, Cout %d)", cp_cout, f32_first, w.Cout);
  if (out.cp.hi) {
    PV_CHECK(out.cp.N == in.N && out.cp.H == Ho && out.cp.W == Wo && out.cp.C == cp_cout, PREMVOS_ERR_INVALID_ARG,
             "conv_umma: CP8 output view is [%d,%d,%d,%d], expected [%d,%d,%d,%d]", out.cp.N, out.cp.C, out.cp.H, out.cp.W, in.N,
             w.Cout, Ho, Wo);
  }
  if (out.f32.p) {
    PV_CHECK(out.f32.N == in.N && out.f32.H == Ho && out.f32.W == Wo && out.f32.C == w.Cout - f32_first, PREMVOS_ERR_INVALID_ARG,
             "conv_umma: fp32 output view shape mismatch");
  }
  PV_CHECK(out.cp.hi || out.f32.p || out.f8.p, PREMVOS_ERR_INVALID_ARG, "conv_umma: no output");
  if (out.f8.p)
    PV_CHECK(out.f8.N == in.N && out.f8.H == Ho && out.f8.W == Wo && out.f8.C == w.Cout, PREMVOS_ERR_INVALID_ARG,
             "conv_umma: F8 output view shape mismatch");
  if (out.res_f8.p)
    PV_CHECK(!out.res.hi && out.res_f8.N == in.N && out.res_f8.H == Ho && out.res_f8.W == Wo && out.res_f8.C == w.Cout,
             PREMVOS_ERR_INVALID_ARG, "conv_umma: F8 residual view shape mismatch");
  if (out.res.hi)
    PV_CHECK(out.res.N == in.N && out.res.H == Ho && out.res.W == Wo && out.res.C == w.Cout, PREMVOS_ERR_INVALID_ARG,
             "conv_umma: residual view shape mismatch");

  UmmaConvArgs& a = *reinterpret_cast<UmmaConvArgs*>(plan->args);
  static_assert(sizeof(UmmaConvArgs) <= sizeof(plan->args), "ConvPlanUmma::args too small");
  memset(&a, 0, sizeof(a));
  const int taps = w.R * w.S;
  // 1x1 / stride 1 / no padding: the spatial structure is irrelevant -> tile the flattened pixel list (needs the
  // 128 bytes of slack every activation buffer is allocated with: the last 8-pixel row may straddle the plane end)
  // small output maps (RoI heads at 7x7, the ReID network's 8x8 / 4x4 stages): several whole images share one 128-row tile
  int fold = 0;
  if (!w.pair && !mixed && env_int("PREMVOS_FOLD", 1) != 0 && Ho * Wo <= 64 && in.N >= 2 && Wo * g.stride <= 256 && Ho * g.stride <= 256)
    fold = std::min(128 / (Ho * Wo), in.N);
  if (fold < 2) fold = 0;
  // measured (profiles/r02_fold_layers.txt): 3x3 stride-1 layers keep halo mode (one box per k-block serves all taps) down to two
  // images per tile -- at 7x7 / 8x8 the per-tap boxes of a folded tile cost more than the half-empty halo tile; from 4 images per
  // tile on, and for every layer that runs in tap mode anyway (1x1, strided), folding wins (4x4 maps: 2-4x)
  if (fold && fold < 4 && taps > 1 && g.stride == 1 && env_int("PREMVOS_FOLD", 1) != 2) fold = 0;
  // stacked halo tiles for the 3x3 stride-1 layers that stay in halo mode: images of at most 8 x 16 share a tile at a row pitch of
  // Ho + halo rows (RoI head 7x7: 2 images per 16 x 8 tile, 38 % -> 77 % of the rows; 8x8 maps: 3 per 32 x 8 tile, 50 % -> 75 %)
  int hfold = 0, hf_pitch = 0, hf_mt = 1;
  if (!fold && !w.pair && env_int("PREMVOS_FOLD", 1) != 0 && env_int("PREMVOS_HFOLD", 1) != 0 && taps > 1 && g.stride == 1 && !mixed && g.dil == 1 &&
      Wo <= 8 && in.N >= 2 && Wo == in.W + g.pad_l + g.pad_r - (w.S - 1)) {
    const int pitch = Ho + (w.R - 1) * g.dil;
    double best = 1.3 * Ho * Wo / (128.0 * ((Ho + 15) / 16));   // must beat one image per tile column by a margin
    for (int mt = 1; mt <= 2; mt++) {
      if (Ho > 16 * mt) continue;
      const int G = std::min((16 * mt - Ho) / pitch + 1, in.N);
      const double util = (double)G * Ho * Wo / (128.0 * mt);
      if (G >= 2 && util > best) { best = util; hfold = G; hf_pitch = pitch; hf_mt = mt; }
    }
  }
  const bool flat = !fold && !mixed && taps == 1 && g.stride == 1 && g.pad_t == 0 && g.pad_l == 0 && g.pad_b == 0 && g.pad_r == 0 &&
                    env_int("PREMVOS_FLAT", 1) != 0 && (long)in.H * in.W >= 8;
  const int real_hw = Ho * Wo;
  const int geoH = flat ? (real_hw + 7) / 8 : Ho, geoW = flat ? 8 : Wo;
  a.Ho = geoH; a.Wo = geoW; a.flat_hw = flat ? real_hw : 0;
  a.stride = g.stride; a.stride_x = sx; a.R = w.R; a.S = w.S; a.dil = g.dil; a.pad_t = g.pad_t; a.pad_l = g.pad_l;
  a.KC = w.KC; a.kblocks = w.kblocks; a.BN = w.BN; a.Cout = w.Cout;
  a.merged_x = (g.stride == 1 && !mixed) ? 1 : 0;
  a.w_plane = w.KC * w.BN * 16;
  if (w.pair) return plan_conv_pair(plan, in, out, w, g, flat, real_hw, cp_cout, f32_first, ws);
  // MT = 2 halves the weight traffic per pixel; only worth it when the grid still fills the machine twice over
  // two sub-tiles per CTA either stacked (32 x 8 pixels) or side by side (16 x 16): take the one that pads less
  const long tiles_v = (long)((geoW + 7) / 8) * ((geoH + 31) / 32), tiles_h = (long)((geoW + 15) / 16) * ((geoH + 15) / 16);
  const bool horiz = tiles_h < tiles_v;
  const long ctas_mt2 = std::min(tiles_v, tiles_h) * in.N * w.ntiles;
  int mt_pref = (ctas_mt2 >= 2 * 148 && geoH > 16 && taps > 1) ? 2 : 1;   // 1x1 layers measured best with one sub-tile at BN = 128
  const bool wide = w.BN > 128;                                           // 256-wide N tile: one CTA per SM owns the whole TMEM
  if (hfold) mt_pref = hf_mt;
  if (wide || fold) mt_pref = 1;   // measured (tools/conv_sweep2.py): 256 x 256 single-buffered tiles lose to 128 x 256 double-buffered ones
  mt_pref = env_int("PREMVOS_MT", mt_pref);
  PV_CHECK(mt_pref == 1 || mt_pref == 2, PREMVOS_ERR_INVALID_ARG, "conv_umma: MT=%d", mt_pref);
  const int tps_env = env_int("PREMVOS_TPS", 0);
  PV_CHECK(tps_env == 0 || taps % tps_env == 0, PREMVOS_ERR_INVALID_ARG, "conv_umma: TPS=%d does not divide %d taps", tps_env, taps);
  // candidate configurations in order of preference; the first whose minimal pipeline fits shared memory wins
  bool found = false;
  for (int cand = 0; cand < 4 && !found; cand++) {
    const int mt = (cand & 1) ? 1 : mt_pref;
    const int halo_env = env_int("PREMVOS_HALO", -1);
    const bool want_halo = cand < 2;
    if ((halo_env == 0 && want_halo) || (halo_env == 1 && !want_halo)) continue;   // tuning override
    if ((cand & 1) && mt_pref == 1) continue;
    const bool hz = horiz && mt == 2 && !hfold;
    const int halo_w = (hz ? 16 : 8) + (w.S - 1) * g.dil, halo_h = (hz ? 16 : 16 * mt) + (w.R - 1) * g.dil;
    const long halo_px = (long)halo_w * halo_h, tap_px = (long)taps * 128 * mt;
    // halo mode only when it moves fewer bytes into shared memory than per-tap boxes
    // (measured: from dilation 4 on the halo box is so large that one CTA per SM remains; per-tap boxes win)
    const bool halo_ok = g.stride == 1 && !mixed && taps > 1 && g.dil < 4 && halo_px * 5 < tap_px * 4 && halo_w * 8 <= 256 && halo_h <= 256;
    if (want_halo && (!halo_ok || fold)) continue;
    a.MT = mt;
    a.mt_horizontal = hz ? 1 : 0;
    a.halo = want_halo ? 1 : 0;
    a.box_w = a.halo ? halo_w : (hz ? 16 : 8);
    a.box_h = a.halo ? halo_h : (hz ? 16 : 16 * mt);
    a.a_box_bytes = w.KC * a.box_h * a.box_w * 16;
    a.a_plane = round_up(a.a_box_bytes, 128);
    a.fold = 0; a.hf_pitch = 0;
    if (hfold && a.halo && mt == hf_mt) {   // the box holds hfold images of hf_pitch rows; the MMAs still walk halo_h rows per chunk
      a.fold = hfold; a.hf_pitch = hf_pitch; a.n_real = in.N;
      a.box_h = hfold * hf_pitch;
      a.a_box_bytes = w.KC * a.box_h * a.box_w * 16;
      a.a_plane = round_up(std::max(a.a_box_bytes, ((w.KC - 1) * a.box_h + halo_h) * a.box_w * 16), 128);
    }
    if (fold) {   // rows are consecutive pixels of consecutive images; the MMA always reads 128 rows per chunk
      a.fold = fold; a.slot = Ho * Wo; a.n_real = in.N; a.merged_x = 0;
      a.a_lbo = fold * a.slot * 16;
      a.a_box_bytes = w.KC * a.a_lbo;
      a.a_plane = round_up(std::max(a.a_box_bytes, (w.KC - 1) * a.a_lbo + 2048), 128);
    }
    a.a_stages = a.halo ? 2 : 3;
    // taps per weight stage: fewer, larger bulk copies (a TMA request has a fixed cost) -- the largest group of <= 32 KB
    // that still leaves room for two CTAs per SM; only if nothing fits 110 KB is the whole SM used
    for (int pass = 0; pass < 2 && !found; pass++) {
      const int limit = pass == 0 ? 110 * 1024 : SMEM_LIMIT;
      for (int t = taps; t >= 1 && !found; t--) {
        if (taps % t != 0 || !(t <= w.S || t % w.S == 0)) continue;
        if (tps_env ? (t != tps_env) : (t > 1 && t * 2 * a.w_plane > 32 * 1024)) continue;
        a.TPS = t;
        a.w_stage = round_up(t * 2 * a.w_plane, 128);
        a.w_stages = 2;
        if (a.a_stages * 2 * a.a_plane + a.w_stages * a.w_stage + 1024 <= limit) found = true;
      }
    }
  }
  PV_CHECK(found, PREMVOS_ERR_UNSUPPORTED, "conv_umma: no pipeline configuration fits shared memory (KC=%d BN=%d dil=%d)", w.KC, w.BN, g.dil);
  const int tile_w = a.mt_horizontal ? 16 : 8, tile_h = a.mt_horizontal ? 16 : 16 * a.MT;
  if (!a.hf_pitch) hfold = 0;   // the stacked-halo candidate did not fit: plain tiles
  a.tiles_x = (fold || hfold) ? 1 : (geoW + tile_w - 1) / tile_w;
  a.tiles_y = (fold || hfold) ? 1 : (geoH + tile_h - 1) / tile_h;
  const int n_groups = a.fold ? (in.N + a.fold - 1) / a.fold : in.N;   // images, or groups of `fold` images, per output-channel tile
  // deepen the rings; stay under 110 KB when the minimal pipeline does (two CTAs per SM), else use the whole SM
  int budget = (!wide && a.a_stages * 2 * a.a_plane + 2 * a.w_stage + 1024 <= 110 * 1024) ? 110 * 1024 : SMEM_LIMIT - 1024;
  if (env_int("PREMVOS_BUDGET_KB", 0) > 0) budget = env_int("PREMVOS_BUDGET_KB", 0) * 1024;
  if (a.halo && budget > 120 * 1024) a.a_stages = 3;
  while (a.w_stages < MAX_W_STAGES && a.w_stages * a.TPS < 3 * taps &&
         a.a_stages * 2 * a.a_plane + (a.w_stages + 1) * a.w_stage + 1024 <= budget)
    a.w_stages++;
  while (!a.halo && a.a_stages < MAX_A_STAGES && a.a_stages * 2 * a.a_plane + a.w_stages * a.w_stage + 2 * a.a_plane + 1024 <= budget)
    a.a_stages++;
  // 1x1 layers: one ring for A and weights (see the kernel), as deep as the budget allows
  a.lockstep = (!a.halo && taps == 1 && env_int("PREMVOS_LOCKSTEP", 1) != 0) ? 1 : 0;
  if (a.lockstep) {
    int st = (budget - 1024) / (2 * a.a_plane + a.w_stage);
    st = std::max(2, std::min(st, MAX_W_STAGES));
    st = std::min(st, std::max(2, a.kblocks + 1));
    a.a_stages = a.w_stages = st;
  }
  plan->smem_bytes = a.a_stages * 2 * a.a_plane + a.w_stages * a.w_stage + 128 /*align slack*/ + 512 /*barriers*/;
  PV_CHECK(plan->smem_bytes <= SMEM_LIMIT, PREMVOS_ERR_UNSUPPORTED, "conv_umma: %d bytes of shared memory needed", plan->smem_bytes);
  a.w = w.w; a.bias = w.bias; a.slope = g.slope;
  a.out_hi = out.cp.hi; a.out_lo = out.cp.lo; a.out_chunks = out.cp.chunks; a.out_c0 = out.cp.c0;
  a.out_f32 = out.f32.p; a.out_cs = out.f32.cs; a.out_coff = out.f32.coff;
  a.cp_cout = cp_cout; a.f32_first = f32_first; a.f32_linear = out.f32_linear ? 1 : 0; a.f32_accum = out.f32_accumulate ? 1 : 0;
  a.res_hi = out.res.hi; a.res_lo = out.res.lo; a.res_chunks = out.res.chunks; a.res_c0 = out.res.c0;
  a.out_f8 = out.f8.p; a.f8_chunks = out.f8.chunks; a.f8_c0 = out.f8.c0;
  a.res_f8 = out.res_f8.p; a.resf_chunks = out.res_f8.chunks; a.resf_c0 = out.res_f8.c0;
  a.dbg = env_int("PREMVOS_DBG", 0);
  a.NACC = 1;   // accumulator replicas were an experiment (no gain: the dependent-MMA chain is not the limiter)
  PV_CHECK(a.NACC >= 1 && a.NACC <= 3 && a.NACC * a.MT * w.BN <= 512, PREMVOS_ERR_INVALID_ARG, "conv_umma: NACC=%d does not fit TMEM", a.NACC);
  uint32_t cols = 32;
  while ((int)cols < a.NACC * a.MT * w.BN) cols <<= 1;
  a.tmem_cols = cols;
  plan->grid_x = a.tiles_x * a.tiles_y * n_groups;
  plan->grid_y = w.ntiles;
  // split-K when the grid cannot fill the machine: the K loop of a CTA is a serial chain of barrier handshakes
  a.ksplit = 1; a.kb_per = a.kblocks; a.cout_pad = w.ntiles * w.BN; a.partial = nullptr; a.partial_stride = 0;
  {
    const long ctas = (long)plan->grid_x * plan->grid_y;
    int split = 1;
    if (ctas * 2 <= 148 && a.kblocks >= 4) {
      split = (int)std::min<long>(std::min<long>(a.kblocks / 2, 148 / ctas), 16);
    }
    split = env_int("PREMVOS_KSPLIT", split);
    if (split > a.kblocks) split = a.kblocks;
    if (split > 1) {
      a.kb_per = (a.kblocks + split - 1) / split;
      a.ksplit = (a.kblocks + a.kb_per - 1) / a.kb_per;   // every split owns at least one k-block
      a.partial_stride = (long)in.N * Ho * Wo * a.cout_pad;
      PV_CUDA(cudaMalloc((void**)&a.partial, (size_t)a.ksplit * a.partial_stride * sizeof(float)));
      PV_CUDA(cudaMemset(a.partial, 0, (size_t)a.ksplit * a.partial_stride * sizeof(float)));
      plan->scratch = a.partial;
    }
  }
  plan->grid_z = a.ksplit;
  a.ntiles = w.ntiles; a.n_images = n_groups;
  a.total_work = plan->grid_x * plan->grid_y * a.ksplit;
  {
    const double a_bytes = 4.0 * (double)in.N * in.H * in.W * w.CinPhys;
    const int dflt = (w.ntiles > 1 && a_bytes > 96e6) ? 1 : 0;
    a.n_fastest = env_int("PREMVOS_NFAST", dflt) != 0 ? 1 : 0;
  }
  // two CTAs per SM when shared memory and TMEM allow it; a second accumulator buffer when TMEM allows that too
  const int cps = (!wide && plan->smem_bytes <= 112 * 1024 && a.NACC * a.MT * w.BN <= 256) ? 2 : 1;
  a.nbuf = (2 * a.NACC * a.MT * w.BN <= (cps == 2 ? 256 : 512)) ? 2 : 1;
  a.nbuf = env_int("PREMVOS_NBUF", a.nbuf);
  cols = 32;
  while ((int)cols < a.nbuf * a.NACC * a.MT * w.BN) cols <<= 1;
  a.tmem_cols = cols;
  plan->ctas_per_sm = cps;
  // tail split (see UmmaConvArgs): only for un-split layers with at least one full wave and a small remainder
  a.tail_items = 0; a.tail_split = 1; a.tail_kb_per = a.kblocks; a.main_work = a.total_work;
  if (a.ksplit == 1 && !a.fold && env_int("PREMVOS_TAIL", 1) != 0) {
    int num_sms = 148;
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, 0);
    const int slots = num_sms * cps, rem = a.total_work % slots;
    // ideal tensor-pipe cycles of one item; the extra finish launch (~5 us) only pays off when a wave is long
    const long item_cycles = (long)a.kblocks * taps * (w.KC / 2) * a.MT * 3 * (w.BN / 2);
    if (a.total_work > slots && rem > 0 && rem * 3 <= slots && a.kblocks >= 4 && item_cycles >= env_int("PREMVOS_TAIL_MIN_CYCLES", 12000)) {
      int split = std::min(slots / rem, a.kblocks / 2);
      split = std::min(split, env_int("PREMVOS_TAIL_SPLIT", 16));
      if (split >= 2) {
        a.tail_kb_per = (a.kblocks + split - 1) / split;
        a.tail_split = (a.kblocks + a.tail_kb_per - 1) / a.tail_kb_per;
        a.tail_items = rem;
        a.main_work = a.total_work - rem;
        const size_t elems = (size_t)a.tail_split * rem * a.MT * 128 * w.BN;
        PV_CUDA(cudaMalloc((void**)&a.partial, elems * sizeof(float)));
        PV_CUDA(cudaMemset(a.partial, 0, elems * sizeof(float)));
        plan->scratch = a.partial;
        a.total_work = a.main_work + rem * a.tail_split;
      }
    }
  }
  a.epi_warps = (wide && w.BN % 32 == 0 && env_int("PREMVOS_EPI8", 1) != 0) ? 8 : 4;

  // input tensor maps over the view's chunk planes: [N][chunks][H][W][8]
  const int vchunks = (in.C + 7) / 8;
  const size_t plane_bytes = (size_t)in.H * in.W * 16;
  __nv_bfloat16* bases[2] = {in.hi + (size_t)in.c0 * in.H * in.W * 8, in.lo + (size_t)in.c0 * in.H * in.W * 8};
  CUtensorMap* maps[2] = {(CUtensorMap*)plan->map_a_hi, (CUtensorMap*)plan->map_a_lo};
  for (int k = 0; k < 2; k++) {
    if (hfold) {
      cuuint64_t dims[4] = {(cuuint64_t)in.W * 8, (cuuint64_t)in.H, (cuuint64_t)in.N, (cuuint64_t)vchunks};
      cuuint64_t strides[3] = {(cuuint64_t)in.W * 16, plane_bytes * in.chunks, plane_bytes};
      cuuint32_t box[4] = {(cuuint32_t)a.box_w * 8, (cuuint32_t)hf_pitch, (cuuint32_t)hfold, (cuuint32_t)w.KC};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      PV_TRY(encode_map(maps[k], bases[k], 4, dims, strides, box, estr));
    } else if (fold) {
      cuuint64_t dims[5] = {8, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.N, (cuuint64_t)vchunks};
      cuuint64_t strides[4] = {16, (cuuint64_t)in.W * 16, plane_bytes * in.chunks, plane_bytes};
      cuuint32_t box[5] = {8, (cuuint32_t)(Wo * g.stride), (cuuint32_t)(Ho * g.stride), (cuuint32_t)fold, (cuuint32_t)w.KC};
      cuuint32_t estr[5] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1, 1};
      PV_TRY(encode_map(maps[k], bases[k], 5, dims, strides, box, estr));
    } else if (flat) {
      cuuint64_t dims[4] = {64, (cuuint64_t)geoH, (cuuint64_t)vchunks, (cuuint64_t)in.N};
      cuuint64_t strides[3] = {128, plane_bytes, plane_bytes * in.chunks};
      cuuint32_t box[4] = {(cuuint32_t)a.box_w * 8, (cuuint32_t)a.box_h, (cuuint32_t)w.KC, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      PV_TRY(encode_map(maps[k], bases[k], 4, dims, strides, box, estr));
    } else if (a.merged_x) {
      cuuint64_t dims[4] = {(cuuint64_t)in.W * 8, (cuuint64_t)in.H, (cuuint64_t)vchunks, (cuuint64_t)in.N};
      cuuint64_t strides[3] = {(cuuint64_t)in.W * 16, plane_bytes, plane_bytes * in.chunks};
      cuuint32_t box[4] = {(cuuint32_t)a.box_w * 8, (cuuint32_t)a.box_h, (cuuint32_t)w.KC, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      PV_TRY(encode_map(maps[k], bases[k], 4, dims, strides, box, estr));
    } else {
      cuuint64_t dims[5] = {8, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)vchunks, (cuuint64_t)in.N};
      cuuint64_t strides[4] = {16, (cuuint64_t)in.W * 16, plane_bytes, plane_bytes * in.chunks};
      cuuint32_t box[5] = {8, (cuuint32_t)(a.box_w * sx), (cuuint32_t)(a.box_h * g.stride), (cuuint32_t)w.KC, 1};
      cuuint32_t estr[5] = {1, (cuuint32_t)sx, (cuuint32_t)g.stride, 1, 1};
      PV_TRY(encode_map(maps[k], bases[k], 5, dims, strides, box, estr));
    }
  }
  const double opx = (double)in.N * Ho * Wo;
  plan->flops = 2.0 * opx * (double)w.Cout * w.Cin * taps;
  plan->bytes = 4.0 * ((double)in.N * in.H * in.W * w.Cin + opx * w.Cout + (double)taps * w.Cin * w.Cout);
  plan->halo = a.halo; plan->MT = a.MT; plan->N = in.N;
  return 0;
}



// ---- fused separable convolution: host side ----
int plan_sepconv_fused(SepConvPlan* plan, const FView& in, const float* dw_w, const float* dw_b, int cpad, bool pre_relu, bool post_relu,
                       const ConvWeightsUmma& pw, const ConvOut& out, float slope) {
  PV_CHECK(in.p && pw.R == 1 && pw.S == 1 && pw.ntiles == 1 && pw.KC == 4 && !pw.pair && pw.BN % 32 == 0 && pw.BN <= 128, PREMVOS_ERR_UNSUPPORTED,
           "sepconv_fused: needs a 1x1 pointwise layer with one channel tile of 32..128 (BN %d, KC %d, tiles %d)", pw.BN, pw.KC, pw.ntiles);
  PV_CHECK(round_up(in.C, 8) == round_up(pw.CinPhys, 8) && cpad >= round_up(in.C, 8) && cpad % 8 == 0, PREMVOS_ERR_INVALID_ARG,
           "sepconv_fused: channel mismatch (input %d, pointwise %d, depthwise arrays %d)", in.C, pw.CinPhys, cpad);
  PV_CHECK((out.f8.p || out.cp.hi) && !out.f32.p && !out.res.hi && !out.res_f8.p, PREMVOS_ERR_UNSUPPORTED,
           "sepconv_fused: F8 / CP8 outputs without residual only");
  if (out.f8.p) PV_CHECK(out.f8.N == in.N && out.f8.H == in.H && out.f8.W == in.W && out.f8.C == pw.Cout, PREMVOS_ERR_INVALID_ARG, "sepconv_fused: F8 output shape");
  if (out.cp.hi) PV_CHECK(out.cp.N == in.N && out.cp.H == in.H && out.cp.W == in.W && out.cp.C == pw.Cout, PREMVOS_ERR_INVALID_ARG, "sepconv_fused: CP8 output shape");
  static_assert(sizeof(SepFusedArgs) <= sizeof(plan->args), "SepConvPlan::args too small");
  SepFusedArgs& a = *reinterpret_cast<SepFusedArgs*>(plan->args);
  memset(&a, 0, sizeof(a));
  a.dw_w = dw_w; a.dw_b = dw_b; a.pw_w = pw.w; a.pw_b = pw.bias;
  a.out_f8 = out.f8.p; a.f8_chunks = out.f8.chunks; a.f8_c0 = out.f8.c0;
  a.out_hi = out.cp.hi; a.out_lo = out.cp.lo; a.out_chunks = out.cp.chunks; a.out_c0 = out.cp.c0;
  a.H = in.H; a.W = in.W; a.in_c0 = in.c0; a.chunks = (in.C + 7) / 8; a.cpad = cpad; a.kblocks = pw.kblocks; a.BN = pw.BN; a.Cout = pw.Cout;
  PV_CHECK(a.kblocks * 32 <= cpad, PREMVOS_ERR_INVALID_ARG, "sepconv_fused: depthwise arrays must cover %d k-blocks of 32 channels (cpad %d)", a.kblocks, cpad);
  a.pre_relu = pre_relu ? 1 : 0; a.post_relu = post_relu ? 1 : 0; a.slope = slope;
  a.tiles_x = (in.W + 7) / 8; a.tiles_y = (in.H + 15) / 16;
  a.w_stage = 2 * 4 * pw.BN * 16;
  a.dbg = env_int("PREMVOS_DBG", 0);
  plan->N = in.N;
  plan->smem_bytes = SF_XS * SF_X_BYTES + SF_AS * SF_A_BYTES + SF_WS * a.w_stage + (10 * cpad + pw.BN) * 4 + 256 + 128;
  PV_CHECK(plan->smem_bytes <= SMEM_LIMIT, PREMVOS_ERR_UNSUPPORTED, "sepconv_fused: %d bytes of shared memory needed", plan->smem_bytes);
  plan->flops_per_image = 2.0 * in.H * in.W * ((double)pw.Cout * pw.Cin + 9.0 * in.C);
  plan->bytes_per_image = 4.0 * in.H * in.W * ((double)in.C + pw.Cout);
  const uint64_t dims[4] = {(uint64_t)in.W * 8, (uint64_t)in.H, (uint64_t)in.chunks, (uint64_t)in.N};
  const uint64_t strides[3] = {(uint64_t)in.W * 32, (uint64_t)in.H * in.W * 32, (uint64_t)in.chunks * in.H * in.W * 32};
  const uint32_t box[4] = {SF_HW * 8, SF_HH, 4, 1};
  return encode_tensor_map_f32(plan->map_in, in.p, 4, dims, strides, box);
}

int launch_sepconv_fused(const SepConvPlan& plan, int n_active, cudaStream_t st) {
  SepFusedArgs a = *reinterpret_cast<const SepFusedArgs*>(plan.args);
  if (n_active < 0 || n_active > plan.N) n_active = plan.N;
  a.n_active = n_active;
  a.total = a.tiles_x * a.tiles_y * n_active;
  if (a.total == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    PV_CUDA(cudaFuncSetAttribute(sepconv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PV_CUDA(cudaGetDevice(&dev));
    PV_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = std::min(a.total, num_sms);
  prof_before(st);
  sepconv_fused_kernel<<<grid, SF_THREADS, plan.smem_bytes, st>>>(*reinterpret_cast<const CUtensorMap*>(plan.map_in), a);
  const char* label = "sepconv_fused_kernel";
  static const int per_layer = env_int("PREMVOS_PROFILE_LAYERS", 0);
  if (per_layer && profiling_enabled()) {
    char buf[160];
    snprintf(buf, sizeof(buf), "sepconv_fused[n%d_%dx%d_cin%d_cout%d]", n_active, a.H, a.W, a.chunks * 8, a.Cout);
    label = prof_intern(buf);
  }
  return after_launch(label, st, plan.flops_per_image * n_active, plan.bytes_per_image * n_active);
}

static int dbg_report(int dbg, int grid, cudaStream_t st) {
  unsigned long long r[8];
  PV_CUDA(cudaStreamSynchronize(st));
  PV_CUDA(cudaMemcpyFromSymbol(r, g_dbg_cta, sizeof(r)));
  fprintf(stderr, "conv_umma dbg: %llu CTAs, first start -> last end %.2f us, CTA duration min %.2f avg %.2f max %.2f us\n", r[5],
          (r[1] - r[0]) * 1e-3, r[4] * 1e-3, r[5] ? r[2] * 1e-3 / r[5] : 0.0, r[3] * 1e-3);
  if (dbg & 128) {
    static unsigned long long rec[512][6];
    PV_CUDA(cudaMemcpyFromSymbol(rec, g_dbg_rec, sizeof(rec)));
    std::vector<int> order;
    for (int i = 0; i < grid && i < 512; i++) order.push_back(i);
    std::sort(order.begin(), order.end(), [&](int x, int y) { return rec[x][2] - rec[x][1] > rec[y][2] - rec[y][1]; });
    for (size_t k = 0; k < order.size(); k += (k < 12 ? 1 : 16)) {
      const unsigned long long* q = rec[order[k]];
      fprintf(stderr, "  cta %3d sm %3llu: start +%6.2f  first MMA +%6.2f  last item MMAs issued +%6.2f  epilogue done +%6.2f  end +%6.2f us\n", order[k], q[0],
              (q[1] - r[0]) * 1e-3, (q[3] - q[1]) * 1e-3, (q[4] - q[1]) * 1e-3, (q[5] - q[1]) * 1e-3, (q[2] - q[1]) * 1e-3);
    }
  }
  return 0;
}
static int dbg_reset() {
  const unsigned long long init[8] = {~0ull, 0, 0, 0, ~0ull, 0, 0, 0};
  PV_CUDA(cudaMemcpyToSymbol(g_dbg_cta, init, sizeof(init)));
  return 0;
}

static int launch_conv_pair(const ConvPlanUmma& plan, cudaStream_t st, int active_n) {
  UmmaConvArgs a = *reinterpret_cast<const UmmaConvArgs*>(plan.args);
  if (active_n >= 0 && active_n < plan.N) {   // a smaller active batch: a shorter item list
    a.n_images = active_n;
    a.total_work = ((active_n * a.tiles_y + 1) / 2) * a.ntiles;
  }
  if (a.total_work == 0) return 0;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PV_CUDA(cudaGetDevice(&dev));
    PV_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // at most two pairs per item: a stream-K range is then at least half an item, an item is finished from at most two partial sums
  const int pairs = (int)std::max<long>(1, std::min<long>(std::min(num_sms / 2, PAIR_SLOTS), 2L * a.total_work));
  a.stream_k_all = env_int("PREMVOS_STREAMK_ALL", 0);
  if (a.dbg & 32) PV_TRY(dbg_reset());
  prof_before(st);
  conv_pair_kernel<<<2 * pairs, 384, plan.smem_bytes, st>>>(*reinterpret_cast<const CUtensorMap*>(plan.map_a_hi),
                                                           *reinterpret_cast<const CUtensorMap*>(plan.map_a_lo),
                                                           *reinterpret_cast<const CUtensorMap*>(plan.map_w), a);
  const double frac = (double)a.n_images / plan.N;
  const char* label = "conv_pair_kernel";
  static const int per_layer = env_int("PREMVOS_PROFILE_LAYERS", 0);
  if (per_layer && profiling_enabled()) {
    char buf[256];
    snprintf(buf, sizeof(buf), "conv_pair[n%d_%dx1_cin%d_cout%d|BN%d_KC%d_st%d_items%d_pairs%d]", a.n_images, a.flat_hw, a.kblocks * a.KC * 8,
             a.Cout, a.BN, a.KC, a.w_stages, a.total_work, pairs);
    label = prof_intern(buf);
  }
  PV_TRY(after_launch(label, st, plan.flops * frac, plan.bytes * frac));
  if (a.dbg & 32) PV_TRY(dbg_report(a.dbg, 2 * pairs, st));
  return 0;
}

int launch_conv_umma(const ConvPlanUmma& plan, cudaStream_t st, int active_n) {
  static bool attr_set = false;
  if (!attr_set) {
    // experiment: one shared-memory carve-out for every kernel of the process, so that alternating small kernels and
    // 110-224 KB convolution kernels never reconfigure the SMs
    if (env_int("PREMVOS_PREFER_SHARED", 0)) PV_CUDA(cudaDeviceSetCacheConfig(cudaFuncCachePreferShared));
    PV_CUDA(cudaFuncSetAttribute(conv_umma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PV_CUDA(cudaFuncSetAttribute(conv_umma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    PV_CUDA(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set = true;
  }
  // persistent CTAs: at most ctas_per_sm per SM, each walks the work items b, b + grid, ...  A smaller active batch
  // just shortens the list (the image index is a digit of the work index).
  UmmaConvArgs a = *reinterpret_cast<const UmmaConvArgs*>(plan.args);
  if (a.pair) return launch_conv_pair(plan, st, active_n);
  if (active_n >= 0 && active_n < plan.N) {   // a smaller active batch: plain item list, no tail split
    const int items = a.tail_items > 0 ? a.main_work + a.tail_items : a.total_work;
    const int groups = a.fold ? (active_n + a.fold - 1) / a.fold : active_n;
    a.total_work = items / a.n_images * groups;
    a.n_images = groups;
    a.n_real = active_n;
    a.tail_items = 0; a.main_work = a.total_work;
  }
  const int n_act = (active_n >= 0 && active_n < plan.N) ? active_n : plan.N;
  if (a.total_work == 0) return 0;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    PV_CUDA(cudaGetDevice(&dev));
    PV_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = std::min(a.total_work, num_sms * plan.ctas_per_sm);
  if (a.dbg & 32) {
    const unsigned long long init[8] = {~0ull, 0, 0, 0, ~0ull, 0, 0, 0};
    PV_CUDA(cudaMemcpyToSymbol(g_dbg_cta, init, sizeof(init)));
  }
  prof_before(st);
  if (a.epi_warps == 8)
    conv_umma_kernel<8><<<grid, 128 + 32 * 8, plan.smem_bytes, st>>>(
        *reinterpret_cast<const CUtensorMap*>(plan.map_a_hi), *reinterpret_cast<const CUtensorMap*>(plan.map_a_lo), a);
  else
    conv_umma_kernel<4><<<grid, UMMA_THREADS, plan.smem_bytes, st>>>(
        *reinterpret_cast<const CUtensorMap*>(plan.map_a_hi), *reinterpret_cast<const CUtensorMap*>(plan.map_a_lo), a);
  const double frac = (double)n_act / plan.N;
  const char* label = "conv_umma_kernel";
  static const int per_layer = env_int("PREMVOS_PROFILE_LAYERS", 0);
  if (per_layer && profiling_enabled()) {   // per-layer breakdown for tools/profile_nets.py
    char buf[256];
    snprintf(buf, sizeof(buf), "conv_umma[n%d_%dx%d_cin%d_cout%d_k%d_s%d_d%d|MT%d_BN%d_KC%d_halo%d_ks%d_cps%d_nbuf%d_ws%d_as%d_ls%d_tail%dx%d_fold%d_hp%d]", n_act,
             a.flat_hw > 0 ? a.flat_hw : a.Ho, a.flat_hw > 0 ? 1 : a.Wo, a.kblocks * a.KC * 8, a.Cout, a.R, a.stride, a.dil, a.MT, a.BN,
             a.KC, a.halo, a.ksplit, plan.ctas_per_sm, a.nbuf, a.w_stages, a.a_stages, a.lockstep, a.tail_items, a.tail_split, a.fold, a.hf_pitch);
    label = prof_intern(buf);
  }
  PV_TRY(after_launch(label, st, plan.flops * frac, plan.bytes * frac));
  if (a.dbg & 32) PV_TRY(dbg_report(a.dbg, grid, st));
  if (a.tail_items > 0) {
    const long total = (long)a.tail_items * a.MT * 128 * (a.BN / 8);
    prof_before(st);
    conv_finish_tail_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    PV_TRY(after_launch("conv_finish_kernel", st, 0.0, (double)total * 32.0 * (a.tail_split + 1)));
  }
  if (plan.grid_z > 1) {
    const int na = (active_n >= 0 && active_n < plan.N) ? active_n : plan.N;
    const long total = (long)na * a.Ho * a.Wo * ((a.Cout + 7) / 8);
    prof_before(st);
    conv_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, na);
    PV_TRY(after_launch("conv_finish_kernel", st, 0.0, (double)total * 32.0 * (a.ksplit + 1)));
  }
  return 0;
}

}  // namespace premvos
