// The MLP of a layer as ONE persistent kernel (experimental, behind ldit_mlp_fused):
//     h = gelu_erf(a W1^T + b1)            BeitIntermediate, HF:428-432   ("fc1", bf16 out)
//     x += lambda2 (.) (h W2^T + b2)       BeitOutput + layer scale + residual, HF:442, 500-504   ("fc2")
// Both GEMMs run through the pipeline of gemm.cuh (CTA pairs, 256 x 192 tiles, TMA ring, TMEM double
// buffering, 16 epilogue warps); what changes is the schedule.  As two launches the pair costs
// ceil(10.8) + 4.6 * ceil(2.7) = 24.8 fc1-tile units per CTA pair at base224 plus a full kernel boundary
// (drain, launch, prologue, first loads ~ 5 us); the work itself is 23.2 units.  Here every CTA pair owns a
// list of tiles -- its fc1 tiles first, then its fc2 tiles -- built on the host so that the lists are equally
// long (pairs that take three fc2 tiles take fewer fc1 tiles), and an fc2 tile only needs the 256 rows of h
// it reads: fc1 tiles publish finished row blocks in global counters, the fc2 producer polls the counter of its
// row block before its first load.  fc1 tiles are dealt in row-block order and every pair finishes its fc1
// list before it touches fc2, so nothing ever waits on a tile that waits on it, and in practice the counters
// are long satisfied when they are read.  The last reader of a row block re-arms its counters.
#pragma once

#include "gemm.cuh"

namespace ldit {

struct MlpArgs {
  int M, D, I;
  const float* b1;      // [I]
  const float* b2;      // [D]
  const float* lam2;    // [D] or nullptr
  int nb1, nb2;         // column tiles of fc1 (I / BN) and fc2 (D / BN)
  int num_m_blocks, tiles1;
  const int* sched;     // [clusters][sched_stride] tile ids (fc1: [0, tiles1), fc2: tiles1 + ...), -1 terminated
  int sched_stride;
  int* ready;           // [num_m_blocks] columns of h completely stored, per 32-row warp slice and CTA  | [num_m_blocks] readers seen
  int ready_target;     // 4 * 2 * I
  int dbg;              // experiments only (LDIT_MLP_DBG): bit 0 = publish after the stores were READ, not completed (racy; timing only)
};

template <int BN>
struct MlpCfg {
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_ROWS = BN / 2;
  static constexpr int B_BYTES = B_ROWS * kBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int CG_COLS = BN / 4;
  static constexpr int CHUNKS = CG_COLS / kEpiCols;
  static constexpr int CHUNK_BYTES = 32 * kEpiCols * 4;   // sized for the fp32 (fc2) chunks; the bf16 ones use half
  static constexpr int STAGING_BYTES = kGemmEpiWarps * CHUNK_BYTES;
  static constexpr int COLOP_BYTES = kGemmEpiWarps * 2 * 64 * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAGES_FIT = (kMaxSmem - 1024 - BAR_BYTES - STAGING_BYTES - COLOP_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + COLOP_BYTES + BAR_BYTES + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
mlp_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ CUtensorMap tmC1, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC2, const MlpArgs g) {
  using Cfg = MlpCfg<BN>;
  constexpr int S = Cfg::STAGES;
  constexpr int TILE_M = 2 * kBM;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * Cfg::A_BYTES;
  uint8_t* sStage = smem + S * Cfg::STAGE_BYTES;
  float* sColOp = reinterpret_cast<float*>(sStage + Cfg::STAGING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGING_BYTES + Cfg::COLOP_BYTES);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / 2;

  if (warp == kWarpProducer && lane == 0) {
    tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmC1);
    tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmC2);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kGemmEpiWarps * 2);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == kWarpAlloc) {
    tmem_alloc_cg2(tmem_slot, kTmemCols);
    tmem_relinquish_cg2();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  const int* my_sched = g.sched + static_cast<size_t>(cluster_id) * g.sched_stride;
  const int nkb1 = (g.D + kBK - 1) / kBK, nkb2 = (g.I + kBK - 1) / kBK;

  if (warp == kWarpProducer) {
    int stage = 0;
    uint32_t phase = 0;
    for (int si = 0;; ++si) {
      const int tile = __ldg(my_sched + si);
      if (tile < 0) break;
      const bool second = tile >= g.tiles1;
      const int t = second ? tile - g.tiles1 : tile;
      const int nb = second ? g.nb2 : g.nb1;
      const int mb = t / nb;
      const int m0 = mb * TILE_M + static_cast<int>(rank) * kBM;
      const int n0 = (t - mb * nb) * BN + static_cast<int>(rank) * Cfg::B_ROWS;
      const int nkb = second ? nkb2 : nkb1;
      const CUtensorMap* mA = second ? &tmA2 : &tmA1;
      const CUtensorMap* mB = second ? &tmB2 : &tmB1;
      if (second) {
        // the 256 rows of h this tile reads are complete once every fc1 tile of the row block has published
        if (lane == 0) {
          const uint64_t t0 = global_timer_ns();
          while (ld_acquire_gpu(g.ready + mb) < g.ready_target) {
            __nanosleep(64);
            if (global_timer_ns() - t0 > 4000000000ull) asm volatile("trap;");   // 4 s: a scheduling bug, never a legitimate wait
          }
          // both CTAs of all nb2 tiles of this row block read the counter; the last one re-arms it
          if (atomicAdd(g.ready + g.num_m_blocks + mb, 1) == 2 * g.nb2 - 1) {
            g.ready[g.num_m_blocks + mb] = 0;
            g.ready[mb] = 0;
          }
          fence_proxy_async_all();   // the TMA loads below (async proxy) are ordered behind the acquire
        }
        __syncwarp();
      }
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES * 2);
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_cg2(sA + stage * Cfg::A_BYTES, mA, leader_full, kb * kBK, m0);
          tma_load_2d_cg2(sB + stage * Cfg::B_BYTES, mB, leader_full, kb * kBK, n0);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kWarpMma) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN, 0, 0);
      const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(sA));
      const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int si = 0;; ++si) {
        const int tile = __ldg(my_sched + si);
        if (tile < 0) break;
        const int nkb = tile >= g.tiles1 ? nkb2 : nkb1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one_sync()) {
            const uint64_t adesc = adesc0 + static_cast<uint32_t>(stage * (Cfg::A_BYTES >> 4));
            const uint64_t bdesc = bdesc0 + static_cast<uint32_t>(stage * (Cfg::B_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) umma_bf16_ss_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            tcgen05_commit_cg2(&empty_bar[stage], 3);
            if (kb + 1 == nkb) tcgen05_commit_cg2(&tfull_bar[acc], 3);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < kGemmEpiWarps) {
    const int quarter = warp & 3;
    const int cgrp = warp >> 2;
    constexpr int kChunks = Cfg::CHUNKS;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int row_in_tile = static_cast<int>(rank) * kBM + quarter * 32;
    uint8_t* buf = sStage + warp * Cfg::CHUNK_BYTES;
    float* my_colop = sColOp + warp * 128;
    const GeluCoef gelu_k;

    // fc1 tiles publish their row block only after the warp has seen its stores COMPLETE; the wait is deferred to
    // the top of the next tile, where the warp would otherwise sit on the accumulator barrier
    int pending_mb = -1, pending_cols = 0;
    auto publish_pending = [&]() {
      if (pending_mb >= 0) {
        if (lane == 0) {
          if (g.dbg & 1) tma_store_wait_read<0>(); else tma_store_wait<0>();
          fence_proxy_async_all();
          __threadfence();
          red_release_gpu_add(g.ready + pending_mb, pending_cols);
        }
        __syncwarp();
        pending_mb = -1;
      }
    };

    for (int si = 0;; ++si) {
      const int tile = __ldg(my_sched + si);
      if (tile < 0) break;
      publish_pending();
      const bool second = tile >= g.tiles1;
      const int t = second ? tile - g.tiles1 : tile;
      const int nb = second ? g.nb2 : g.nb1;
      const int N = second ? g.D : g.I;
      const int mb = t / nb;
      const int row0 = mb * TILE_M + row_in_tile;
      const int col0 = (t - mb * nb) * BN + cgrp * Cfg::CG_COLS;
      {
        const int cc = col0 + 2 * lane;
        const bool ok = 2 * lane < Cfg::CG_COLS && cc < N;
        float2 bb = make_float2(0.f, 0.f), ss = make_float2(1.f, 1.f);
        const float* bias = second ? g.b2 : g.b1;
        if (ok && bias != nullptr) bb = __ldg(reinterpret_cast<const float2*>(bias + cc));
        if (second && ok && g.lam2 != nullptr) ss = __ldg(reinterpret_cast<const float2*>(g.lam2 + cc));
        *reinterpret_cast<float2*>(my_colop + 2 * lane) = bb;
        *reinterpret_cast<float2*>(my_colop + 64 + 2 * lane) = ss;
        __syncwarp();
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * kAccStride + cgrp * Cfg::CG_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
      uint32_t r[2][16];
      tmem_ld_32x32b_x16(taddr, r[0]);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col = col0 + c * kEpiCols;
        const bool col_ok = col < N;
        uint32_t (&rc)[16] = r[c & 1];
        float4 b4[4], s4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          b4[j] = *reinterpret_cast<const float4*>(my_colop + c * kEpiCols + 4 * j);
          s4[j] = *reinterpret_cast<const float4*>(my_colop + 64 + c * kEpiCols + 4 * j);
        }
        tmem_wait_ld16(rc);
        if (c + 1 < kChunks) {
          tmem_ld_32x32b_x16(taddr + (c + 1) * kEpiCols, r[(c + 1) & 1]);
        } else {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
        }
        if (!second) {
          // fc1: + bias, erf-GELU, bf16 tile store
          float v[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            f2_unpack(f2_add(f2_pack(__uint_as_float(rc[4 * j + 0]), __uint_as_float(rc[4 * j + 1])), f2_pack(b4[j].x, b4[j].y)),
                      v[4 * j + 0], v[4 * j + 1]);
            f2_unpack(f2_add(f2_pack(__uint_as_float(rc[4 * j + 2]), __uint_as_float(rc[4 * j + 3])), f2_pack(b4[j].z, b4[j].w)),
                      v[4 * j + 2], v[4 * j + 3]);
          }
#pragma unroll
          for (int e = 0; e < 16; e += 2) gelu_erf_pair(v[e], v[e + 1], gelu_k);
          uint4 o[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            o[j].x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
            o[j].y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o[j].z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            o[j].w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(buf + lane * 32 + ((j ^ ((lane >> 2) & 1)) << 4)) = o[j];
        } else {
          // fc2: lambda2 (.) (acc + bias), fp32 tile reduce-add into the residual stream
          float4 o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[j].x = s4[j].x * (__uint_as_float(rc[4 * j + 0]) + b4[j].x);
            o[j].y = s4[j].y * (__uint_as_float(rc[4 * j + 1]) + b4[j].y);
            o[j].z = s4[j].z * (__uint_as_float(rc[4 * j + 2]) + b4[j].z);
            o[j].w = s4[j].w * (__uint_as_float(rc[4 * j + 3]) + b4[j].w);
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o[j];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && col_ok) {
          if (second) tma_reduce_add_2d(&tmC2, buf, col, row0);
          else tma_store_2d(&tmC1, buf, col, row0);
          tma_store_commit();
        }
      }
      if (!second) {
        pending_mb = mb;
        pending_cols = max(0, min(Cfg::CG_COLS, N - col0));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();
    publish_pending();
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == kWarpAlloc) tmem_dealloc_cg2(tmem_base, kTmemCols);
}

}  // namespace ldit
