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
// CTA = 128 + 256 threads, one CTA per SM, static round-robin tile scheduler over clusters:
//   warp 10   TMA producer   (one lane, both CTAs): A/B tiles -> 128B-swizzled smem ring; all
//             completion bytes are credited to the LEADER CTA's "full" barrier
//   warp 11   MMA issuer     (one lane, leader CTA only): tcgen05.mma, accumulators
//             double-buffered in TMEM; tcgen05.commit multicasts "slot free" / "accumulator ready"
//             to both CTAs
//   warp 9    TMEM allocator
//   warps 0-7 epilogue (both CTAs, own 128 rows): tcgen05.ld -> registers -> bias / erf-GELU /
//             layer-scale + residual -> swizzled smem staging -> TMA store (the residual tile
//             arrives by TMA load into the same staging buffer, prefetched one chunk ahead)
//
// Replaces the cuBLAS calls behind nn.Linear at HF:324-338 (QKV), HF:383 (+HF:488-492),
// HF:429-430 and HF:442 (+HF:500-504), and the conv at HF:218 (as an im2col GEMM).
#pragma once

#include "ptx.cuh"

namespace ldit {

enum : int { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_SCALE_RESID = 2, EPI_PATCH = 3 };

struct GemmArgs {
  int M, N, K;
  const float* bias;   // [N] or nullptr
  const float* scale;  // [N] layer-scale (EPI_SCALE_RESID) or nullptr (== 1)
  const float* resid;  // fp32 [M, ldo] residual stream (EPI_SCALE_RESID); may alias out
  void* out;           // bf16 [M, ldo] (EPI_BIAS, EPI_BIAS_GELU) or fp32 (EPI_SCALE_RESID, EPI_PATCH)
  int ldo;             // output row pitch in elements
  int P;               // EPI_PATCH: patches per image; GEMM row b*P+p -> token row b*(P+1)+1+p
  const float* posb;   // EPI_PATCH: [P, N] fp32 = position rows 1..P + conv bias
  int num_m_blocks, num_n_blocks;
  int dbg;             // experiments only (LDIT_GEMM_DBG): bit 0 = epilogue drains TMEM but stores nothing
  long long* tl;       // experiments only: clock64 timeline [cluster][16 tiles][8] (leader CTA), or nullptr
};

constexpr int kBM = 128;  // rows per CTA
constexpr int kBK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmEpiWarps = 8;
constexpr int kGemmThreads = 128 + kGemmEpiWarps * 32;
// Warp roles.  The SM's warp arbiter favours the highest warp id on each scheduler, so the two
// latency-critical single-lane roles (TMA producer, MMA issuer) get the highest ids and the
// compute-heavy epilogue warps the lowest: a busy epilogue must never delay an MMA issue.
constexpr int kWarpAlloc = kGemmEpiWarps + 1;     // 9
constexpr int kWarpProducer = kGemmEpiWarps + 2;  // 10
constexpr int kWarpMma = kGemmEpiWarps + 3;       // 11
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr int kTmemCols = 512;
constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA

template <int BN, int EPI, int CTAS>
struct GemmCfg {
  static_assert(BN == 128 || BN == 192 || BN == 256, "BN");
  static_assert(CTAS == 1 || CTAS == 2, "CTAS");
  static constexpr bool OUT_F32 = (EPI == EPI_SCALE_RESID || EPI == EPI_PATCH);
  static constexpr bool STAGED = (EPI != EPI_PATCH);  // patch rows are re-indexed per image: direct stores
  static constexpr int TILE_M = kBM * CTAS;
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_ROWS = BN / CTAS;            // rows of W staged by each CTA
  static constexpr int B_BYTES = B_ROWS * kBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue staging: per warp two buffers of 32 rows x 32 columns
  static constexpr int CHUNK_BYTES = 32 * 32 * (OUT_F32 ? 4 : 2);
  static constexpr int STAGING_BYTES = STAGED ? kGemmEpiWarps * 2 * CHUNK_BYTES : 0;
  static constexpr int BAR_BYTES = 512;
  static constexpr int STAGES_FIT = (kMaxSmem - 1024 - BAR_BYTES - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "swizzle-128B tiles must stay 1 KB aligned");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

// exact-erf GELU (HF:430, ACT2FN["gelu"]) evaluated as x * Phi(x) with
// Phi(x) = 0.5 erfc(-x / sqrt 2) from the Abramowitz-Stegun 7.1.26 rational form (|erf error|
// < 1.5e-7, i.e. far below the bf16 rounding of the result): ~13 FMA-pipe ops + 2 MUFU, so
// the epilogue stays under the MMA time of its tile -- erff() costs ~2x that.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float h = p * e;  // 0.5 * erfc(|x| / sqrt 2)
  return x * (x >= 0.f ? 1.0f - h : h);
}

template <int BN, int EPI, int CTAS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
  using Cfg = GemmCfg<BN, EPI, CTAS>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * Cfg::A_BYTES;
  uint8_t* sStage = smem + S * Cfg::STAGE_BYTES;  // 1 KB aligned: every tile size is a multiple of 1 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* resid_bar = tempty_bar + 2;  // [kGemmEpiWarps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resid_bar + 2 * kGemmEpiWarps);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;  // 0 = leader of the pair
  const int cluster_id = blockIdx.x / CTAS;
  const int num_clusters = gridDim.x / CTAS;

  if (warp == kWarpProducer && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (Cfg::STAGED) tma_prefetch_desc(&tmC);
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
    for (int i = 0; i < 2 * kGemmEpiWarps; ++i) mbar_init(&resid_bar[i], 1);
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
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = g.num_m_blocks * g.num_n_blocks;
  const int nkb = (g.K + kBK - 1) / kBK;

  // Producer and MMA loops are executed by ALL lanes of their warp (warp-uniform control flow,
  // so ptxas keeps stage / phase / descriptors in uniform registers); only the TMA / tcgen05
  // instructions themselves are issued by one elected lane.  A single-lane loop costs ~80
  // dependent SASS instructions per k-block (ELECT + R2UR per operand) and starves the tensor
  // core whenever an epilogue warp shares the scheduler.
  if (warp == kWarpProducer) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / g.num_n_blocks) * Cfg::TILE_M + static_cast<int>(rank) * kBM;
      const int n0 = (tile % g.num_n_blocks) * BN + static_cast<int>(rank) * Cfg::B_ROWS;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if constexpr (CTAS == 2) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES * 2);
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_cg2(sA + stage * Cfg::A_BYTES, &tmA, leader_full, kb * kBK, m0);
            tma_load_2d_cg2(sB + stage * Cfg::B_BYTES, &tmB, leader_full, kb * kBK, n0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * kBK, m0);
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * kBK, n0);
          }
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kWarpMma) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(Cfg::TILE_M, BN, 0, 0);
      const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(sA));
      const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int ti = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++ti) {
        long long* tl = (g.tl != nullptr && lane == 0 && ti < 16) ? g.tl + (static_cast<size_t>(cluster_id) * 16 + ti) * 8 : nullptr;
        if (tl) tl[0] = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        if (tl) tl[1] = clock64();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one_sync()) {
            const uint64_t adesc = adesc0 + static_cast<uint32_t>(stage * (Cfg::A_BYTES >> 4));
            const uint64_t bdesc = bdesc0 + static_cast<uint32_t>(stage * (Cfg::B_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              if constexpr (CTAS == 2) umma_bf16_ss_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              else umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            // smem slot reusable (in both CTAs) once these MMAs have read it
            if constexpr (CTAS == 2) tcgen05_commit_cg2(&empty_bar[stage], 3); else tcgen05_commit(&empty_bar[stage]);
            // last k-block: the accumulator is complete (both CTAs' epilogues)
            if (kb == nkb - 1) {
              if constexpr (CTAS == 2) tcgen05_commit_cg2(&tfull_bar[acc], 3); else tcgen05_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
          if (tl && kb == 0) tl[2] = clock64();
        }
        if (tl) tl[3] = clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < kGemmEpiWarps) {
    const int ew = warp;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int half = ew >> 2;       // which half of the BN columns
    constexpr int kChunks = BN / 2 / 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int row_in_tile = static_cast<int>(rank) * kBM + quarter * 32;

    auto release_accumulator = [&](int a) {
      // all TMEM reads of this accumulator are done: hand it back to the (leader's) MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[a]), 0));
        else mbar_arrive(&tempty_bar[a]);
      }
    };

    if constexpr (Cfg::STAGED) {
      uint8_t* my_stage = sStage + ew * 2 * Cfg::CHUNK_BYTES;
      uint64_t* my_bar = resid_bar + ew * 2;
      uint32_t gc = 0;  // chunks processed by this warp so far: buffer = gc & 1, residual phase = (gc >> 1) & 1
      if constexpr (EPI == EPI_SCALE_RESID) {
        if (lane == 0 && cluster_id < num_tiles) {  // residual of the very first chunk
          const int tile = cluster_id;
          mbar_arrive_expect_tx(&my_bar[0], Cfg::CHUNK_BYTES);
          tma_load_2d(my_stage, &tmC, &my_bar[0], (tile % g.num_n_blocks) * BN + half * (BN / 2),
                      (tile / g.num_n_blocks) * Cfg::TILE_M + row_in_tile);
        }
      }
      int ti = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++ti) {
        const int row0 = (tile / g.num_n_blocks) * Cfg::TILE_M + row_in_tile;
        const int col0 = (tile % g.num_n_blocks) * BN + half * (BN / 2);
        long long* tl = (g.tl != nullptr && rank == 0 && ew == 0 && lane == 0 && ti < 16)
                            ? g.tl + (static_cast<size_t>(cluster_id) * 16 + ti) * 8 : nullptr;
        if (tl) tl[4] = clock64();
        mbar_wait(&tfull_bar[acc], acc_phase);
        tcgen05_fence_after();
        if (tl) tl[5] = clock64();
        const uint32_t taddr = tmem_base + acc * kAccStride + half * (BN / 2) + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c, ++gc) {
          uint8_t* buf = my_stage + (gc & 1) * Cfg::CHUNK_BYTES;
          uint8_t* nbuf = my_stage + ((gc + 1) & 1) * Cfg::CHUNK_BYTES;
          if (lane == 0) {
            if constexpr (EPI == EPI_SCALE_RESID) {
              // the other buffer was last read by the store of chunk gc-1: it must have drained
              // before the next residual chunk is prefetched into it
              tma_store_wait_read<0>();
              int ntile = tile, nc = c + 1;
              if (nc == kChunks) { ntile = tile + num_clusters; nc = 0; }
              if (ntile < num_tiles) {
                mbar_arrive_expect_tx(&my_bar[(gc + 1) & 1], Cfg::CHUNK_BYTES);
                tma_load_2d(nbuf, &tmC, &my_bar[(gc + 1) & 1], (ntile % g.num_n_blocks) * BN + half * (BN / 2) + nc * 32,
                            (ntile / g.num_n_blocks) * Cfg::TILE_M + row_in_tile);
              }
            } else {
              tma_store_wait_read<1>();  // only the store of chunk gc-2 (this chunk's buffer) must have drained
            }
          }
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tcgen05_wait_ld();
          if (c == kChunks - 1) release_accumulator(acc);
          if (g.dbg & 1) continue;
          __syncwarp();  // lane 0's wait_group.read above covers the whole warp's writes into buf
          const int col = col0 + c * 32;
          const bool col_ok = col < g.N;
          if constexpr (EPI == EPI_SCALE_RESID) {
            mbar_wait(&my_bar[gc & 1], (gc >> 1) & 1);
            // fp32 rows of 128 B, 128B swizzle: 16-byte chunk j of row `lane` sits at j ^ (lane & 7)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4* p = reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4));
              float4 x = *p;
              float v0 = __uint_as_float(r[4 * j]), v1 = __uint_as_float(r[4 * j + 1]);
              float v2 = __uint_as_float(r[4 * j + 2]), v3 = __uint_as_float(r[4 * j + 3]);
              if (g.bias != nullptr && col_ok) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col) + j);
                v0 += b.x; v1 += b.y; v2 += b.z; v3 += b.w;
              }
              if (g.scale != nullptr && col_ok) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(g.scale + col) + j);
                x.x = fmaf(s.x, v0, x.x); x.y = fmaf(s.y, v1, x.y); x.z = fmaf(s.z, v2, x.z); x.w = fmaf(s.w, v3, x.w);
              } else {
                x.x += v0; x.y += v1; x.z += v2; x.w += v3;
              }
              *p = x;
            }
          } else {
            // bf16 rows of 64 B, 64B swizzle: 16-byte chunk j of row `lane` sits at j ^ ((lane >> 1) & 3)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * j + e]);
              if (g.bias != nullptr && col_ok) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + col) + 2 * j);
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + col) + 2 * j + 1);
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = gelu_erf_fast(v[e]);
              }
              uint4 o;
              o.x = pack_bf16x2(v[0], v[1]);
              o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]);
              o.w = pack_bf16x2(v[6], v[7]);
              *reinterpret_cast<uint4*>(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o;
            }
          }
          if (!(g.dbg & 4)) fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0 && col_ok && !(g.dbg & 2)) {
            tma_store_2d(&tmC, buf, col, row0);
            tma_store_commit();
          }
        }
        if (tl) tl[6] = clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (lane == 0) tma_store_wait<0>();  // smem must stay valid until the last stores have read it
    } else {
      // EPI_PATCH: GEMM row b*P+p lands on token row b*(P+1)+1+p, plus position/conv-bias row p
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int n0 = (tile % g.num_n_blocks) * BN;
        const int row = (tile / g.num_n_blocks) * Cfg::TILE_M + row_in_tile + lane;
        const bool row_ok = row < g.M;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + acc * kAccStride + half * (BN / 2) + (static_cast<uint32_t>(quarter * 32) << 16);
        const int b = row / g.P, p = row - b * g.P;
        const size_t orow = static_cast<size_t>(b) * (g.P + 1) + 1 + p;
        const float* posb_row = g.posb + static_cast<size_t>(p) * g.N;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tcgen05_wait_ld();
          if (c == kChunks - 1) release_accumulator(acc);
          const int col = n0 + half * (BN / 2) + c * 32;
          if (row_ok && col < g.N) {
            float* op = reinterpret_cast<float*>(g.out) + orow * g.ldo + col;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 pb = __ldg(reinterpret_cast<const float4*>(posb_row + col) + j);
              float4 x = make_float4(__uint_as_float(r[4 * j]) + pb.x, __uint_as_float(r[4 * j + 1]) + pb.y,
                                     __uint_as_float(r[4 * j + 2]) + pb.z, __uint_as_float(r[4 * j + 3]) + pb.w);
              reinterpret_cast<float4*>(op)[j] = x;
            }
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  }

  __syncwarp();  // warps 0/1 ran single-lane loops: reconverge before the .aligned barriers
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
