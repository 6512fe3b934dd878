// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//     C[M, N] = A[M, K] (bf16, row-major)  x  W[N, K]^T (bf16, row-major = nn.Linear.weight)
// fp32 accumulation in TMEM, fused epilogues.  Both operands are K-major, so the natural
// PyTorch layouts are consumed as they are, with no transposes anywhere.
//
// CTAS = 2 (default): a CTA pair on one TPC computes a 256 x BN tile with
// tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and HALF of the BN rows of
// W, so the pair pulls (256 + BN) x K operand bytes from L2 for 256 x BN x K MACs (1.5x fewer
// bytes per FLOP than two independent 128 x BN tiles; this GEMM is L2->SM-bandwidth-bound
// otherwise) and each SM's shared-memory operand traffic halves.  CTAS = 1 is the single-CTA
// 128 x BN variant of the same code.
//
// CTA = 19 warps, one CTA per SM, static round-robin tile scheduler over clusters:
//   warp 17     TMA producer   (one lane, both CTAs): A/B tiles -> 128B-swizzled smem ring; all
//               completion bytes are credited to the LEADER CTA's "full" barrier
//   warp 18     MMA issuer     (one lane, leader CTA only): tcgen05.mma, accumulators
//               double-buffered in TMEM; tcgen05.commit multicasts "slot free" / "accumulator ready"
//               to both CTAs
//   warp 16     TMEM allocator
//   warps 0-15  epilogue (both CTAs, own 128 rows; warp = TMEM lane quarter x column quarter):
//               tcgen05.ld -> registers -> bias / erf-GELU / layer-scale -> swizzled smem staging
//               -> TMA store (bf16 outputs) or TMA reduce-add into the fp32 residual stream
//
// EPI_CONV_BIAS runs the FPN's 3x3 convolutions (TV:ops/feature_pyramid_network.py:118-124, 193) through the
// same pipeline as an implicit GEMM: the A tile of k-block (tap, channel block) is ONE 4-D TMA box of the
// channels-last input shifted by the tap offset -- out-of-image coordinates are zero-filled by the TMA unit,
// which is the convolution's padding -- so no im2col matrix ever exists in memory.
//
// Replaces the cuBLAS calls behind nn.Linear at HF:324-338 (QKV), HF:383 (+HF:488-492),
// HF:429-430 and HF:442 (+HF:500-504), and the conv at HF:218 (as an im2col GEMM).
#pragma once

#include "ptx.cuh"

namespace ldit {

enum : int { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_SCALE_RESID = 2, EPI_PATCH = 3, EPI_CONV_BIAS = 4, EPI_CONV_BIAS_F32 = 5, EPI_BIAS_SCALE = 6, EPI_PATCH_TMA = 7,
              EPI_WGRAD = 8, EPI_DGRAD = 9 };   // EPI_WGRAD: C f32 [M, N] += A^T B with A [K, M] and B [K, N] row-major (both operands MN-major in smem);
                                 // the epilogue is EPI_SCALE_RESID's fp32 reduce-add without bias / scale / rounding
// EPI_DGRAD: C bf16 [M, N] = A B with A [M, K] K-major as usual and B [K, N] row-major (MN-major in smem): dA = dY W for
// nn.Linear's W [N_out, K_in] as it is stored -- no transposed weight copy; the epilogue is EPI_BIAS's (bias = nullptr)
__host__ __device__ constexpr bool epi_is_resid(int epi) { return epi == EPI_SCALE_RESID || epi == EPI_WGRAD; }
__host__ __device__ constexpr bool epi_b_mn(int epi) { return epi == EPI_WGRAD || epi == EPI_DGRAD; }
// EPI_PATCH_TMA: the patch embedding with its A operand gathered by TMA straight out of the NCHW page batch (16-bit pixels):
// no im2col matrix, CLS rows written by the same kernel
__host__ __device__ constexpr bool epi_is_patch(int epi) { return epi == EPI_PATCH || epi == EPI_PATCH_TMA; }
// EPI_BIAS_SCALE: bf16 out = scale (.) (acc + bias) -- the layer-scaled branch of a residual block, stored for a fused
// residual-add + LayerNorm kernel to pick up (rowwise.cuh) instead of being reduce-added into the fp32 stream here
__host__ __device__ constexpr bool epi_has_scale(int epi) { return epi == EPI_SCALE_RESID || epi == EPI_BIAS_SCALE || epi == EPI_WGRAD; }
// EPI_CONV_BIAS_F32: the same convolution with an fp32 output map (the detection heads behind the FPN hold fp32 weights)
__host__ __device__ constexpr bool epi_is_conv(int epi) { return epi == EPI_CONV_BIAS || epi == EPI_CONV_BIAS_F32; }

struct GemmArgs {
  int M, N, K;
  const float* bias;   // [N] or nullptr
  const float* scale;  // [N] layer-scale (EPI_SCALE_RESID) or nullptr (== 1)
  void* out;           // bf16 [M, ldo] (EPI_BIAS, EPI_BIAS_GELU); fp32 [M, ldo] written (EPI_PATCH) or
                       // accumulated into (EPI_SCALE_RESID: the residual stream itself)
  int ldo;             // output row pitch in elements
  int P;               // EPI_PATCH: patches per image; GEMM row b*P+p -> token row b*(P+1)+1+p
  const float* posb;   // EPI_PATCH: [P, N] fp32 = position rows 1..P + conv bias
  // EPI_PATCH_TMA (HF:176-180, 218 without an im2col pass): pixels [B, 3, H, W] of a 16-bit float type are a 5-D tensor
  // (px 16, patch column pe_gw, py 16, patch row pe_gh, image x channel) -- dimensions in order of increasing stride, which
  // the TMA unit needs.  One UMMA_K step (16 K elements = pixel row py of channel c) of a CTA's 128 patches (cv_tw columns
  // x cv_th rows of the patch grid; the pair's second CTA takes the cv_th rows below) is ONE box [16, cv_tw, 1, cv_th, 1]
  // = 128 rows of 32 B, a K-major operand tile with 32-byte swizzle; a 64-wide k-block is four of them.  Patches outside
  // the grid are zero filled by the TMA unit and skipped by the epilogue.  cv_tx x cv_ty pair tiles per image.
  int pe_gw, pe_gh;
  int a_f16;           // both operands fp16 (pixels as the reference feeds them under autocast + an fp16 copy of the weights)
  const float* cls;    // [N] fp32 = cls_token + position row 0, written to token row 0 of every image
  int num_m_blocks, num_n_blocks;
  int ksplit, kb_per_split;   // EPI_WGRAD: pieces of the K range and 64-wide k-blocks per piece
  int round_bf16;      // EPI_SCALE_RESID: round scale * (acc + bias) to bf16 before the fp32 reduce-add, so that the residual stream
                       // gets bit for bit what the deferred form (EPI_BIAS_SCALE + add_layernorm_kernel) adds -- the forward's
                       // numbers then do not depend on which of the two forms a geometry uses.  0 for the wgrad accumulation.
  int m_reverse;       // 1: row blocks are visited last-to-first (consume a just-written A operand freshest-first, see ldit_api.cu)
  // EPI_CONV_BIAS (3x3 convolution, stride 1, zero padding 1, over a channels-last image [B, H, W, Cin] as an
  // implicit GEMM): a CTA's 128 rows are a cv_th x cv_tw patch of output pixels, a CTA pair covers two patches
  // side by side; an image is cv_ty x cv_tx pair tiles; K = 9 taps x Cin in (ky, kx, cin) order, cv_cblocks = Cin/64
  int cv_tw, cv_th, cv_tx, cv_ty, cv_cblocks;
#ifdef LDIT_DEBUG_HOOKS   // diagnosis builds only (tools/gemm_timeline.py, LDIT_GEMM_DBG): never in the product library
  int dbg;             // bit 0 = epilogue drains TMEM but stores nothing, ...
  long long* tl;       // clock64 timeline [cluster][16 tiles][8] (leader CTA), or nullptr
#endif
};

#ifdef LDIT_DEBUG_HOOKS
#define LDIT_DBG(g, bit) ((g).dbg & (bit))
#define LDIT_TL(g) ((g).tl)
#else
#define LDIT_DBG(g, bit) false
#define LDIT_TL(g) static_cast<long long*>(nullptr)
#endif

#ifndef LDIT_EPI_BUFS
#define LDIT_EPI_BUFS 1        // staging buffers per epilogue warp (1: the smem goes to the operand ring instead)
#endif
#ifndef LDIT_KSTEP
#define LDIT_KSTEP 1           // 2: producer / MMA loops take two ring slots per iteration (measured slower: coarser turnaround)
#endif
#ifndef LDIT_TAIL_RING
#define LDIT_TAIL_RING 0       // 1: the last tile's epilogue stages its chunks in the (then idle) operand ring (measured: no gain, same-box A/B 2.68 vs 2.69 ms/step)
#endif
#ifndef LDIT_EPI_ROLLED
#define LDIT_EPI_ROLLED 0      // 1: chunk loop not unrolled (no cross-chunk overlap inside a warp)
#endif
#ifndef LDIT_EPI_SYNC_PAIRS
#define LDIT_EPI_SYNC_PAIRS 0  // n > 0: __syncwarp() after every n GELU pairs (caps the ILP of the math burst)
#endif
constexpr int kBM = 128;  // rows per CTA
constexpr int kBK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
// Warp roles.  16 epilogue warps = 4 per SM sub-partition (a TMEM lane quarter is only reachable
// from warps with warp_id % 4 == quarter): with fewer, the fixed latencies of the epilogue chain
// (tcgen05.ld -> math -> st.shared -> proxy fence -> TMA store) are exposed and the tile time
// is set by the epilogue instead of the tensor core.  The SM's warp arbiter favours the highest
// warp id on each scheduler, so the two latency-critical single-lane roles (TMA producer, MMA
// issuer) get the highest ids: a busy epilogue must never delay an MMA issue.
constexpr int kGemmEpiWarps = 16;
constexpr int kWarpAlloc = kGemmEpiWarps;         // 16
constexpr int kWarpProducer = kGemmEpiWarps + 1;  // 17
constexpr int kWarpMma = kGemmEpiWarps + 2;       // 18
constexpr int kGemmThreads = (kGemmEpiWarps + 3) * 32;  // 608
constexpr int kEpiCols = 16;     // accumulator columns per epilogue chunk (one tcgen05.ld.x16)
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr int kTmemCols = 512;
constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA

template <int BN, int EPI, int CTAS>
struct GemmCfg {
  static_assert(BN == 128 || BN == 192 || BN == 256, "BN");
  static_assert(CTAS == 1 || CTAS == 2, "CTAS");
  static constexpr bool OUT_F32 = (epi_is_resid(EPI) || epi_is_patch(EPI) || EPI == EPI_CONV_BIAS_F32);
  static constexpr int TILE_M = kBM * CTAS;
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_ROWS = BN / CTAS;            // rows of W staged by each CTA
  static constexpr int B_BYTES = B_ROWS * kBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue: warp (quarter, column group) owns 32 rows x BN/4 columns, in chunks of 16 columns;
  // per warp two staging buffers of 32 rows x 16 columns (rows of 32 B bf16 / 64 B fp32)
  static constexpr int CG_COLS = BN / 4;
  static constexpr int CHUNKS = CG_COLS / kEpiCols;
  // Tried and rejected for the bf16 outputs (same-box A/B, base224 step): one wide box per warp and
  // tile (96-byte rows; stores arrive in bursts, main loop bimodal 4400 / 7500 cycles, +3 %), and
  // 256-bit stores straight from registers (no TMA store at all: the operand loads are then on time,
  // but the LSU traffic delays the MMA warp's own dispatch, +3 %).  L2 prefetch of the A operand
  // ahead of the ring: +8 % (the prefetches compete with the loads in the TMA unit).
  static constexpr int CHUNK_BYTES = 32 * kEpiCols * (OUT_F32 ? 4 : 2);
  static constexpr int STAGING_BYTES = kGemmEpiWarps * LDIT_EPI_BUFS * CHUNK_BYTES;
  // per epilogue warp: bias and layer-scale of the warp's CG_COLS columns, staged once per tile
  static constexpr int COLOP_BYTES = epi_is_patch(EPI) ? 0 : kGemmEpiWarps * 2 * 64 * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAGES_FIT = (kMaxSmem - 1024 - BAR_BYTES - STAGING_BYTES - COLOP_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = (STAGES_FIT > 8 ? 8 : STAGES_FIT) & ~(LDIT_KSTEP - 1);  // even when the loops take slots in pairs
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "swizzle-128B tiles must stay 1 KB aligned");
  static_assert(kGemmEpiWarps * CHUNKS * CHUNK_BYTES <= STAGES * STAGE_BYTES, "tail staging must fit the operand ring");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + COLOP_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

// exact-erf GELU (HF:430, ACT2FN["gelu"]):  gelu(x) = x Phi(x) = max(x, 0) - |x| * 0.5 erfc(|x| / sqrt 2).
// 0.5 erfc(z) is evaluated as exp2(P(z)), P = degree-7 minimax fit of log2(0.5 erfc(z)) on
// z in [0, 5] (z is clamped there: 0.5 erfc(5) = 7.7e-13, i.e. x Phi(x) for x < -7.07 is returned
// as ~-5e-12 instead of something even smaller).  Evaluated in fp32 the result is within 1.5e-5
// relative / 1.5e-6 absolute of the exact function -- under 1 % of a bf16 half-ulp, so the rounded
// bf16 output is the one an erff()-based epilogue produces -- for 9 FMA-pipe ops and ONE MUFU per
// element (erff() costs ~2x that, the rational A&S 7.1.26 form needs two MUFU).
// Coefficients: Chebyshev-node least-squares fit, tools/fit_gelu_poly.py.
// Two elements per instruction: Blackwell's packed fp32 pipe (FFMA2 / FMUL2 / FADD2) does the
// Horner chain for a pair at the issue cost of one -- the epilogue warps share their schedulers
// with the MMA issuer and the TMA producer, and every issue slot they do not take shortens the
// main loop (measured: scalar GELU math stretched the MMA issue span of a tile by 14 %).
// P has no clamp: it decreases monotonically beyond the fitted range (P(z) < -40 for z > 5), so
// 0.5 erfc(z) just keeps underflowing towards 0 as it should.
struct GeluCoef {
  uint64_t nk, c0, c1, c2, c3, c4, c5, c6, c7;
  __device__ __forceinline__ GeluCoef()
      : nk(f2_splat(-0.70710678118654752f)), c0(f2_splat(-1.000004768371582f)), c1(f2_splat(-1.627893328666687f)),
        c2(f2_splat(-0.9177651405334473f)), c3(f2_splat(-0.15145094692707062f)), c4(f2_splat(0.03325735405087471f)),
        c5(f2_splat(-0.005002959165722132f)), c6(f2_splat(0.0004494435270316899f)), c7(f2_splat(-1.794292256818153e-05f)) {}
};
__device__ __forceinline__ void gelu_erf_pair(float& x0, float& x1, const GeluCoef& k) {
  const uint64_t nu = f2_pack(__uint_as_float(__float_as_uint(x0) | 0x80000000u),
                              __uint_as_float(__float_as_uint(x1) | 0x80000000u));  // -|x|
  const uint64_t z = f2_mul(nu, k.nk);                                              // |x| / sqrt 2
  uint64_t p = f2_fma(z, k.c7, k.c6);
#if !defined(LDIT_EXP_NO_HORNER)  // experiment only
  p = f2_fma(p, z, k.c5);
  p = f2_fma(p, z, k.c4);
  p = f2_fma(p, z, k.c3);
  p = f2_fma(p, z, k.c2);
  p = f2_fma(p, z, k.c1);
  p = f2_fma(p, z, k.c0);
#endif
  float p0, p1, h0, h1;
  f2_unpack(p, p0, p1);
#if defined(LDIT_EXP_NO_MUFU)   // experiment only (wrong numbers): same instruction count without the MUFU
  h0 = p0 * 1.0001f; h1 = p1 * 1.0001f;
#else
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(p0));  // 0.5 * erfc(|x| / sqrt 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(p1));
#endif
  const uint64_t y = f2_fma(nu, f2_pack(h0, h1), f2_pack(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
  f2_unpack(y, x0, x1);
}

template <int BN, int EPI, int CTAS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
  using Cfg = GemmCfg<BN, EPI, CTAS>;
  constexpr int S = Cfg::STAGES;
  static_assert(!(epi_is_conv(EPI) || EPI == EPI_PATCH_TMA) || (CTAS == 2 && LDIT_KSTEP == 1), "the convolution mode is written for CTA pairs, one ring slot per step");
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * Cfg::A_BYTES;
  uint8_t* sStage = smem + S * Cfg::STAGE_BYTES;  // 1 KB aligned: every tile size is a multiple of 1 KB
  float* sColOp = reinterpret_cast<float*>(sStage + Cfg::STAGING_BYTES);   // [warp][bias 64 | scale 64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES + Cfg::COLOP_BYTES);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;  // 0 = leader of the pair
  const int cluster_id = blockIdx.x / CTAS;
  const int num_clusters = gridDim.x / CTAS;

  if (warp == kWarpProducer && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (!epi_is_patch(EPI)) tma_prefetch_desc(&tmC);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kGemmEpiWarps * CTAS);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == kWarpAlloc) {
    if constexpr (CTAS == 2) {
      tmem_alloc_cg2(tmem_slot, kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tcgen05_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  pdl_wait();   // the prologue above touched no global memory; everything below may
  const uint32_t tmem_base = *tmem_slot;

  // EPI_WGRAD: the contraction (tokens) is long and the output small, so the K range is cut into g.ksplit pieces that all
  // reduce-add into the same fp32 tile; work item = (k piece, m block, n block)
  const int mn_tiles = g.num_m_blocks * g.num_n_blocks;
  const int num_tiles = (EPI == EPI_WGRAD) ? mn_tiles * g.ksplit : mn_tiles;
  const int nkb = (g.K + kBK - 1) / kBK;

  // Producer and MMA loops are executed by ALL lanes of their warp (warp-uniform control flow,
  // so ptxas keeps stage / phase / descriptors in uniform registers); only the TMA / tcgen05
  // instructions themselves are issued by one elected lane.  A single-lane loop costs ~80
  // dependent SASS instructions per k-block (ELECT + R2UR per operand) and starves the tensor
  // core whenever an epilogue warp shares the scheduler.
  // Experiment knob LDIT_KSTEP=2: both loops take two ring slots per iteration.  Measured slower --
  // the ring is then released and refilled 128 of K at a time and its turnaround latency, which is
  // what limits the main loop under epilogue load, gets longer.
  const int kstep = (LDIT_KSTEP == 2 && !(nkb & 1)) ? 2 : 1;
  if (warp == kWarpProducer) {
    int stage = 0;
    uint32_t phase = 0;
    int pti = 0;
    for (int item = cluster_id; item < num_tiles; item += num_clusters, ++pti) {
      const int tile = (EPI == EPI_WGRAD) ? item % mn_tiles : item;
      const int kb0 = (EPI == EPI_WGRAD) ? (item / mn_tiles) * g.kb_per_split : 0;
      const int kb1 = (EPI == EPI_WGRAD) ? min(nkb, kb0 + g.kb_per_split) : nkb;
      const int mblk = g.m_reverse ? g.num_m_blocks - 1 - tile / g.num_n_blocks : tile / g.num_n_blocks;
      const int m0 = mblk * Cfg::TILE_M + static_cast<int>(rank) * kBM;
      const int n0 = (tile % g.num_n_blocks) * BN + static_cast<int>(rank) * Cfg::B_ROWS;
      int cvx = 0, cvy = 0, cvb = 0, cv_c = 0, cv_kx = 0, cv_ky = 0;   // EPI_CONV_BIAS: patch origin, running (ky, kx, channel block)
      if constexpr (epi_is_conv(EPI)) {
        const int mb = tile / g.num_n_blocks, per_img = g.cv_tx * g.cv_ty;
        cvb = mb / per_img;
        const int r = mb - cvb * per_img, ty = r / g.cv_tx;
        cvx = ((r - ty * g.cv_tx) * 2 + static_cast<int>(rank)) * g.cv_tw;
        cvy = ty * g.cv_th;
      }
      if constexpr (EPI == EPI_PATCH_TMA) {   // patch rectangle of this CTA: columns cvx.., rows cvy.. of image cvb's patch grid
        const int mb = tile / g.num_n_blocks, per_img = g.cv_tx * g.cv_ty;
        cvb = mb / per_img;
        const int r = mb - cvb * per_img, ty = r / g.cv_tx;
        cvx = (r - ty * g.cv_tx) * g.cv_tw;
        cvy = (ty * 2 + static_cast<int>(rank)) * g.cv_th;
      }
      long long* ptl = (LDIT_TL(g) != nullptr && rank == 0 && lane == 0 && pti < 16) ? LDIT_TL(g) + (static_cast<size_t>(cluster_id) * 16 + pti) * 16 : nullptr;
      long long wempty = 0;
      for (int kb = kb0; kb < kb1; kb += kstep) {
        const long long w0 = ptl ? clock64() : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (kstep == 2) mbar_wait(&empty_bar[stage + 1], phase ^ 1);
        if (ptl) { wempty += clock64() - w0; if (kb + kstep >= kb1) ptl[9] = wempty; }
        if (elect_one_sync()) {
          for (int j = 0; j < kstep; ++j) {
            const int st = stage + j;
            if constexpr (CTAS == 2) {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[st], Cfg::STAGE_BYTES * 2);
              const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[st]), 0);
              if constexpr (epi_is_conv(EPI))
                tma_load_4d_cg2(sA + st * Cfg::A_BYTES, &tmA, leader_full, cv_c * kBK, cvx + cv_kx - 1, cvy + cv_ky - 1, cvb);
              else if constexpr (EPI == EPI_PATCH_TMA) {   // k-block kb = (channel kb / 4, pixel rows 4 (kb % 4) ..): one box per pixel row
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  tma_load_5d_cg2(sA + st * Cfg::A_BYTES + q * (Cfg::A_BYTES / 4), &tmA, leader_full, 0, cvx, ((kb + j) & 3) * 4 + q, cvy,
                                  cvb * 3 + ((kb + j) >> 2));
              } else if constexpr (EPI == EPI_WGRAD) {
                // MN-major operands: a stage holds [64 k-rows x 64 columns] atoms of 8 KB, one per 64 columns of the tile
                tma_load_2d_cg2(sA + st * Cfg::A_BYTES, &tmA, leader_full, m0, (kb + j) * kBK);
                tma_load_2d_cg2(sA + st * Cfg::A_BYTES + 8192, &tmA, leader_full, m0 + 64, (kb + j) * kBK);
              } else
                tma_load_2d_cg2(sA + st * Cfg::A_BYTES, &tmA, leader_full, (kb + j) * kBK, m0);
              if constexpr (epi_b_mn(EPI)) {
#pragma unroll
                for (int q = 0; q < Cfg::B_ROWS / 64; ++q)
                  tma_load_2d_cg2(sB + st * Cfg::B_BYTES + q * 8192, &tmB, leader_full, n0 + 64 * q, (kb + j) * kBK);
              } else
                tma_load_2d_cg2(sB + st * Cfg::B_BYTES, &tmB, leader_full, (kb + j) * kBK, n0);
            } else {
              mbar_arrive_expect_tx(&full_bar[st], Cfg::STAGE_BYTES);
              if constexpr (EPI == EPI_WGRAD) {
                tma_load_2d(sA + st * Cfg::A_BYTES, &tmA, &full_bar[st], m0, (kb + j) * kBK);
                tma_load_2d(sA + st * Cfg::A_BYTES + 8192, &tmA, &full_bar[st], m0 + 64, (kb + j) * kBK);
#pragma unroll
                for (int q = 0; q < Cfg::B_ROWS / 64; ++q)
                  tma_load_2d(sB + st * Cfg::B_BYTES + q * 8192, &tmB, &full_bar[st], n0 + 64 * q, (kb + j) * kBK);
              } else {
                tma_load_2d(sA + st * Cfg::A_BYTES, &tmA, &full_bar[st], (kb + j) * kBK, m0);
                if constexpr (epi_b_mn(EPI)) {
#pragma unroll
                  for (int q = 0; q < Cfg::B_ROWS / 64; ++q)
                    tma_load_2d(sB + st * Cfg::B_BYTES + q * 8192, &tmB, &full_bar[st], n0 + 64 * q, (kb + j) * kBK);
                } else
                  tma_load_2d(sB + st * Cfg::B_BYTES, &tmB, &full_bar[st], (kb + j) * kBK, n0);
              }
            }
          }
        }
        __syncwarp();
        if constexpr (epi_is_conv(EPI)) {   // next k-block: channel block fastest, then kx, then ky
          if (++cv_c == g.cv_cblocks) { cv_c = 0; if (++cv_kx == 3) { cv_kx = 0; ++cv_ky; } }
        }
        stage += kstep;
        if (stage >= S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kWarpMma) {
    if (rank == 0) {
      constexpr uint32_t idesc_bf16 = umma_idesc_bf16(Cfg::TILE_M, BN, EPI == EPI_WGRAD ? 1 : 0, epi_b_mn(EPI) ? 1 : 0);
      // EPI_PATCH_TMA with fp16 pixels: both operand format fields ([7,10) A, [10,13) B) = 0 (f16); the weights are then an
      // fp16 copy (mixing an f16 A with a bf16 B traps with "illegal instruction" on sm_100a, measured)
      const uint32_t idesc = (EPI == EPI_PATCH_TMA && g.a_f16) ? (idesc_bf16 & ~((7u << 7) | (7u << 10))) : idesc_bf16;
      const uint64_t adesc0 = (EPI == EPI_PATCH_TMA) ? umma_desc_kmajor_sw32(smem_u32(sA))
                            : (EPI == EPI_WGRAD)   ? umma_desc_mnmajor_sw128_atoms(smem_u32(sA), 8192)
                                                   : umma_desc_kmajor_sw128(smem_u32(sA));
      const uint64_t bdesc0 = epi_b_mn(EPI) ? umma_desc_mnmajor_sw128_atoms(smem_u32(sB), 8192) : umma_desc_kmajor_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int ti = 0;
      for (int item = cluster_id; item < num_tiles; item += num_clusters, ++ti) {
        const int kb0 = (EPI == EPI_WGRAD) ? (item / mn_tiles) * g.kb_per_split : 0;
        const int kb1 = (EPI == EPI_WGRAD) ? min(nkb, kb0 + g.kb_per_split) : nkb;
        long long* tl = (LDIT_TL(g) != nullptr && lane == 0 && ti < 16) ? LDIT_TL(g) + (static_cast<size_t>(cluster_id) * 16 + ti) * 16 : nullptr;
        if (tl) tl[0] = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        if (tl) tl[1] = clock64();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        long long wfull = 0;
        for (int kb = kb0; kb < kb1; kb += kstep) {
          const long long w0 = tl ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (kstep == 2) mbar_wait(&full_bar[stage + 1], phase);
          if (tl) wfull += clock64() - w0;
          tcgen05_fence_after();
          if (elect_one_sync()) {
            for (int j = 0; j < kstep; ++j) {
              const int st = stage + j;
              const uint64_t adesc = adesc0 + static_cast<uint32_t>(st * (Cfg::A_BYTES >> 4));
              const uint64_t bdesc = bdesc0 + static_cast<uint32_t>(st * (Cfg::B_BYTES >> 4));
              // A advances by 32 B inside the 128-byte swizzle atom per UMMA_K step; the TMA-gathered patch operand instead
              // holds one 32-byte-swizzled [128 x 16] sub-tile per step (A_BYTES / 4 apart)
              // MN-major operands (EPI_WGRAD) advance by 16 k-rows of 128 B
              constexpr uint32_t a_step = (EPI == EPI_PATCH_TMA) ? (Cfg::A_BYTES / 4) >> 4 : (EPI == EPI_WGRAD) ? 128 : 2;
              constexpr uint32_t b_step = epi_b_mn(EPI) ? 128 : 2;
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k) {
                if constexpr (CTAS == 2) umma_bf16_ss_cg2(d_tmem, adesc + a_step * k, bdesc + b_step * k, idesc, ((kb - kb0) | j | k) != 0);
                else umma_bf16_ss(d_tmem, adesc + a_step * k, bdesc + b_step * k, idesc, ((kb - kb0) | j | k) != 0);
              }
              // smem slot reusable (in both CTAs) once these MMAs have read it
              if constexpr (CTAS == 2) tcgen05_commit_cg2(&empty_bar[st], 3); else tcgen05_commit(&empty_bar[st]);
            }
            // last k-block: the accumulator is complete (both CTAs' epilogues)
            if (kb + kstep >= kb1) {
              if constexpr (CTAS == 2) tcgen05_commit_cg2(&tfull_bar[acc], 3); else tcgen05_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          stage += kstep;
          if (stage >= S) { stage = 0; phase ^= 1; }
          if (tl && kb == 0) tl[2] = clock64();
        }
        if (tl) { tl[3] = clock64(); tl[7] = wfull; }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < kGemmEpiWarps) {
    // ------------------------------------------------------------------------- epilogue
    // warp -> (TMEM lane quarter, column group); thread <-> accumulator row; per chunk of 16
    // columns: tcgen05.ld (the next chunk's load is already in flight) -> fused math in
    // registers -> swizzled smem staging -> one TMA op per chunk issued by lane 0:
    //   EPI_BIAS / EPI_BIAS_GELU  bf16 tile store
    //   EPI_SCALE_RESID           fp32 tile REDUCE-ADD into the residual stream: x += scale * (acc + bias)
    //                             is performed by the L2 atomic units, the residual is never
    //                             loaded into the SM (every element has exactly one contributor,
    //                             so the result is deterministic)
    //   EPI_PATCH                 no TMA: rows are re-indexed per image, so the staged chunk is read
    //                             back row-contiguously (4 lanes x 16 B per row) and written with the
    //                             position / conv-bias rows added
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int cgrp = warp >> 2;     // which quarter of the BN columns
    constexpr int kChunks = Cfg::CHUNKS;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int row_in_tile = static_cast<int>(rank) * kBM + quarter * 32;
    uint8_t* my_stage = sStage + warp * LDIT_EPI_BUFS * Cfg::CHUNK_BYTES;
    uint32_t gc = 0;  // chunks processed by this warp so far: staging buffer = gc & 1
    const GeluCoef gelu_k;

    auto release_accumulator = [&](int a) {
      // all TMEM reads of this accumulator are done: hand it back to the (leader's) MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[a]), 0));
        else mbar_arrive(&tempty_bar[a]);
      }
    };

    int ti = 0;
    for (int item = cluster_id; item < num_tiles; item += num_clusters, ++ti) {
      const int tile = (EPI == EPI_WGRAD) ? item % mn_tiles : item;
      const int row0 = (g.m_reverse ? g.num_m_blocks - 1 - tile / g.num_n_blocks : tile / g.num_n_blocks) * Cfg::TILE_M + row_in_tile;
      const int col0 = (tile % g.num_n_blocks) * BN + cgrp * Cfg::CG_COLS;
      int cvx = 0, cvy = 0, cvb = 0;   // EPI_CONV_BIAS: first pixel of this warp's 32 rows (32 / cv_tw image rows of cv_tw pixels)
      if constexpr (epi_is_conv(EPI)) {
        const int mb = tile / g.num_n_blocks, per_img = g.cv_tx * g.cv_ty;
        cvb = mb / per_img;
        const int r = mb - cvb * per_img, ty = r / g.cv_tx;
        cvx = ((r - ty * g.cv_tx) * 2 + static_cast<int>(rank)) * g.cv_tw;
        cvy = ty * g.cv_th + quarter * (32 / g.cv_tw);
      }
      long long* tl = (LDIT_TL(g) != nullptr && rank == 0 && warp == 0 && lane == 0 && ti < 16)
                          ? LDIT_TL(g) + (static_cast<size_t>(cluster_id) * 16 + ti) * 16 : nullptr;

      // EPI_PATCH: this lane reads back rows (lane >> 2) + 8 i, 16-byte piece lane & 3
      size_t p_orow[4];
      int p_prow[4];
      if constexpr (EPI == EPI_PATCH) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int gm = row0 + (lane >> 2) + 8 * i;
          const int b = gm / g.P, p = gm - b * g.P;
          p_prow[i] = gm < g.M ? p : -1;
          p_orow[i] = static_cast<size_t>(b) * (g.P + 1) + 1 + p;
        }
      }
      if constexpr (EPI == EPI_PATCH_TMA) {
        const int mb = tile / g.num_n_blocks, per_img = g.cv_tx * g.cv_ty;
        const int b = mb / per_img, r = mb - b * per_img, ty = r / g.cv_tx, tx = r - ty * g.cv_tx;
        const int gx0 = tx * g.cv_tw, gy0 = (ty * 2 + static_cast<int>(rank)) * g.cv_th;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rl = quarter * 32 + (lane >> 2) + 8 * i;        // accumulator row = patch (rl / tw, rl % tw) of the rectangle
          const int gy = gy0 + rl / g.cv_tw, gx = gx0 + rl % g.cv_tw;
          const int p = gy * g.pe_gw + gx;
          p_prow[i] = (gy < g.pe_gh && gx < g.pe_gw) ? p : -1;
          p_orow[i] = static_cast<size_t>(b) * (g.P + 1) + 1 + p;
        }
        // token row 0 of the image (cls_token + position row 0, HF:176-180): written once per image and column block,
        // by the CTA that owns patch (0, 0)
        if (tx == 0 && ty == 0 && rank == 0 && quarter == 0) {
          float* dst = reinterpret_cast<float*>(g.out) + static_cast<size_t>(b) * (g.P + 1) * g.ldo;
          for (int jj = lane; jj < Cfg::CG_COLS; jj += 32)
            if (col0 + jj < g.N) dst[col0 + jj] = __ldg(g.cls + col0 + jj);
        }
      }

      // bias / layer-scale of this warp's columns -> smem, requested before the accumulator wait so the
      // global-load latency never sits in the per-chunk chain (CG_COLS <= 64 = 2 floats per lane)
      float* my_colop = sColOp + warp * 128;
      if constexpr (!epi_is_patch(EPI)) {
        const int cc = col0 + 2 * lane;
        const bool ok = 2 * lane < Cfg::CG_COLS && cc < g.N;
        float2 bb = make_float2(0.f, 0.f), ss = make_float2(1.f, 1.f);
        if (ok && g.bias != nullptr) bb = __ldg(reinterpret_cast<const float2*>(g.bias + cc));
        if constexpr (epi_has_scale(EPI)) {
          if (ok && g.scale != nullptr) ss = __ldg(reinterpret_cast<const float2*>(g.scale + cc));
        }
        *reinterpret_cast<float2*>(my_colop + 2 * lane) = bb;
        if constexpr (epi_has_scale(EPI)) *reinterpret_cast<float2*>(my_colop + 64 + 2 * lane) = ss;
        __syncwarp();
      }
      if (tl) tl[4] = clock64();
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      // Experiment knob LDIT_TAIL_RING: the last tile of a CTA overlaps nothing, so its epilogue is pure kernel
      // tail; its MMAs have completed (tfull) and no further load will be issued, i.e. the operand ring is free,
      // and every chunk can have a staging buffer of its own there (no per-chunk wait for the previous TMA
      // store to have read the single regular buffer).  Measured: no change -- the tail is not bound by that wait.
      const bool last_tile = LDIT_TAIL_RING && (item + num_clusters >= num_tiles);
      uint8_t* tail_stage = smem + static_cast<size_t>(warp) * kChunks * Cfg::CHUNK_BYTES;
      if (tl) tl[5] = clock64();
      const uint32_t taddr = tmem_base + acc * kAccStride + cgrp * Cfg::CG_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
      uint32_t r[2][16];
#if LDIT_EPI_ROLLED
#pragma unroll 1
#else
      tmem_ld_32x32b_x16(taddr, r[0]);
#pragma unroll
#endif
      for (int c = 0; c < kChunks; ++c, ++gc) {
        const int col = col0 + c * kEpiCols;
        const bool col_ok = col < g.N;
#if LDIT_EPI_ROLLED
        uint32_t (&rc)[16] = r[0];
        tmem_ld_32x32b_x16(taddr + c * kEpiCols, rc);
#else
        uint32_t (&rc)[16] = r[c & 1];
#endif
        // per-column operands of this chunk (smem broadcast reads)
        float4 b4[4], s4[4];
        if constexpr (!epi_is_patch(EPI)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            b4[j] = *reinterpret_cast<const float4*>(my_colop + c * kEpiCols + 4 * j);
            if constexpr (epi_has_scale(EPI)) s4[j] = *reinterpret_cast<const float4*>(my_colop + 64 + c * kEpiCols + 4 * j);
          }
        }
        tmem_wait_ld16(rc);
#if LDIT_EPI_ROLLED
        if (c + 1 == kChunks) release_accumulator(acc);
#else
        if (c + 1 < kChunks) tmem_ld_32x32b_x16(taddr + (c + 1) * kEpiCols, r[(c + 1) & 1]);
        else release_accumulator(acc);
#endif
        if (LDIT_DBG(g, 1)) continue;
        uint8_t* buf = last_tile ? tail_stage + c * Cfg::CHUNK_BYTES : my_stage + (gc & (LDIT_EPI_BUFS - 1)) * Cfg::CHUNK_BYTES;

        if constexpr (!Cfg::OUT_F32) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            f2_unpack(f2_add(f2_pack(__uint_as_float(rc[4 * j + 0]), __uint_as_float(rc[4 * j + 1])), f2_pack(b4[j].x, b4[j].y)),
                      v[4 * j + 0], v[4 * j + 1]);
            f2_unpack(f2_add(f2_pack(__uint_as_float(rc[4 * j + 2]), __uint_as_float(rc[4 * j + 3])), f2_pack(b4[j].z, b4[j].w)),
                      v[4 * j + 2], v[4 * j + 3]);
          }
          if constexpr (EPI == EPI_BIAS_SCALE) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[4 * j + 0] *= s4[j].x; v[4 * j + 1] *= s4[j].y; v[4 * j + 2] *= s4[j].z; v[4 * j + 3] *= s4[j].w;
            }
          }
          if constexpr (EPI == EPI_BIAS_GELU) {
            if (!LDIT_DBG(g, 8)) {
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                gelu_erf_pair(v[e], v[e + 1], gelu_k);
#if LDIT_EPI_SYNC_PAIRS > 0
                if (((e / 2 + 1) % LDIT_EPI_SYNC_PAIRS) == 0) __syncwarp();
#endif
              }
            }
          }
          uint4 o[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            o[j].x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
            o[j].y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o[j].z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            o[j].w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
          }
          if (lane == 0 && !last_tile) tma_store_wait_read<LDIT_EPI_BUFS - 1>();  // the last store out of this buffer has finished reading it
          __syncwarp();
          // bf16 rows of 32 B, 32B swizzle: 16-byte piece j of row `lane` sits at j ^ ((lane >> 2) & 1)
          if (!LDIT_DBG(g, 16)) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              *reinterpret_cast<uint4*>(buf + lane * 32 + ((j ^ ((lane >> 2) & 1)) << 4)) = o[j];
          } else if (o[0].x == 0x12345678u && o[1].y == 0x9abcdef0u) {
            *reinterpret_cast<uint4*>(buf) = o[0];
          }
        } else {
          float4 o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[j] = make_float4(__uint_as_float(rc[4 * j + 0]), __uint_as_float(rc[4 * j + 1]),
                               __uint_as_float(rc[4 * j + 2]), __uint_as_float(rc[4 * j + 3]));
            if constexpr (epi_is_resid(EPI)) {
              o[j].x = s4[j].x * (o[j].x + b4[j].x);
              o[j].y = s4[j].y * (o[j].y + b4[j].y);
              o[j].z = s4[j].z * (o[j].z + b4[j].z);
              o[j].w = s4[j].w * (o[j].w + b4[j].w);
              if (g.round_bf16) {
                const uint32_t lo = pack_bf16x2(o[j].x, o[j].y), hi = pack_bf16x2(o[j].z, o[j].w);
                o[j] = make_float4(__uint_as_float(lo << 16), __uint_as_float(lo & 0xffff0000u), __uint_as_float(hi << 16),
                                   __uint_as_float(hi & 0xffff0000u));
              }
            }
            if constexpr (EPI == EPI_CONV_BIAS_F32) {
              o[j].x += b4[j].x; o[j].y += b4[j].y; o[j].z += b4[j].z; o[j].w += b4[j].w;
            }
          }
          if constexpr (epi_is_resid(EPI) || EPI == EPI_CONV_BIAS_F32) {
            if (lane == 0 && !last_tile) tma_store_wait_read<LDIT_EPI_BUFS - 1>();
          }
          __syncwarp();  // EPI_PATCH: every lane has finished reading this buffer (chunk gc-2) long ago; keeps the warp converged
          // fp32 rows of 64 B, 64B swizzle: 16-byte piece j of row `lane` sits at j ^ ((lane >> 1) & 3)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o[j];
        }

        if constexpr (epi_is_patch(EPI)) {
          // GEMM row b*P+p lands on token row b*(P+1)+1+p, plus position/conv-bias row p (HF:176-180)
          __syncwarp();
          if (col_ok) {
            const int piece = lane & 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = (lane >> 2) + 8 * i;
              if (p_prow[i] >= 0) {
                float4 x = *reinterpret_cast<const float4*>(buf + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4));
                const float4 pb = __ldg(reinterpret_cast<const float4*>(g.posb + static_cast<size_t>(p_prow[i]) * g.N + col) + piece);
                x.x += pb.x; x.y += pb.y; x.z += pb.z; x.w += pb.w;
                reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + p_orow[i] * g.ldo + col)[piece] = x;
              }
            }
          }
        } else {
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0 && col_ok && !LDIT_DBG(g, 2)) {
            if constexpr (epi_is_resid(EPI)) { if (LDIT_DBG(g, 32)) tma_store_2d(&tmC, buf, col, row0); else tma_reduce_add_2d(&tmC, buf, col, row0); }
            else if constexpr (epi_is_conv(EPI)) tma_store_4d(&tmC, buf, col, cvx, cvy, cvb);   // pixels past the image edge are clipped
            else tma_store_2d(&tmC, buf, col, row0);
            tma_store_commit();
          }
        }
      }
      if (tl) tl[6] = clock64();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if constexpr (!epi_is_patch(EPI)) {
      if (lane == 0) tma_store_wait<0>();  // smem must stay valid until the last stores have read it
    }
  }

  __syncwarp();  // single-lane branches above: reconverge before the .aligned barriers
  tcgen05_fence_before();
  if constexpr (CTAS == 2) {
    cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading its smem
    if (warp == kWarpAlloc) tmem_dealloc_cg2(tmem_base, kTmemCols);
  } else {
    __syncthreads();
    if (warp == kWarpAlloc) tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace ldit
