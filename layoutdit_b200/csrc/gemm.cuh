// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//     C[M, N] = A[M, K] (bf16, row-major)  x  W[N, K]^T (bf16, row-major = nn.Linear.weight)
// fp32 accumulation in TMEM, fused epilogues.  Both operands are K-major, so the natural
// PyTorch layouts are consumed as they are, with no transposes anywhere.
//
// CTA = 128 + 256 threads, one CTA per SM, static round-robin tile scheduler:
//   warp 0    TMA producer   (one lane): cp.async.bulk.tensor A/B tiles -> 128B-swizzled smem ring
//   warp 1    MMA issuer     (one lane): tcgen05.mma 128 x BN x 16, accumulators double-buffered in TMEM
//   warp 2    TMEM allocator
//   warps 4-11 epilogue: tcgen05.ld -> registers -> bias / erf-GELU / layer-scale + residual -> global
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and TMEM full/empty mbarriers
// (MMA <-> epilogue), so the epilogue of tile i overlaps the main loop of tile i+1.
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
};

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmEpiWarps = 8;
constexpr int kGemmThreads = 128 + kGemmEpiWarps * 32;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr int kTmemCols = 512;

template <int BN>
struct GemmCfg {
  static_assert(BN == 128 || BN == 192 || BN == 256, "BN");
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kGemmEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = g.num_m_blocks * g.num_n_blocks;
  const int nkb = (g.K + kBK - 1) / kBK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / g.num_n_blocks) * kBM;
        const int n0 = (tile % g.num_n_blocks) * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * kBK, m0);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * kBK, n0);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          tcgen05_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tfull_bar[acc]);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;          // which half of the BN columns
    constexpr int kChunks = BN / 2 / 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / g.num_n_blocks) * kBM;
      const int n0 = (tile % g.num_n_blocks) * BN;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < g.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * kAccStride + half * (BN / 2) + (static_cast<uint32_t>(quarter * 32) << 16);

      size_t orow = static_cast<size_t>(row);
      const float* posb_row = nullptr;
      if constexpr (EPI == EPI_PATCH) {
        const int b = row / g.P, p = row - b * g.P;
        orow = static_cast<size_t>(b) * (g.P + 1) + 1 + p;
        posb_row = g.posb + static_cast<size_t>(p) * g.N;
      }
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tcgen05_wait_ld();
        if (c == kChunks - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        const int col = n0 + half * (BN / 2) + c * 32;
        if (row_ok && col < g.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[j + e]);
            if (g.bias != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + col + j));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + col + j + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if constexpr (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
              if constexpr (EPI == EPI_BIAS_GELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = gelu_erf(v[e]);
              }
              uint4 o;
              o.x = pack_bf16x2(v[0], v[1]);
              o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]);
              o.w = pack_bf16x2(v[6], v[7]);
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(g.out) + orow * g.ldo + col + j;
              *reinterpret_cast<uint4*>(op) = o;
            } else if constexpr (EPI == EPI_SCALE_RESID) {
              const float* rp = g.resid + orow * g.ldo + col + j;
              float4 x0 = *reinterpret_cast<const float4*>(rp);
              float4 x1 = *reinterpret_cast<const float4*>(rp + 4);
              if (g.scale != nullptr) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(g.scale + col + j));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(g.scale + col + j + 4));
                x0.x = fmaf(s0.x, v[0], x0.x); x0.y = fmaf(s0.y, v[1], x0.y);
                x0.z = fmaf(s0.z, v[2], x0.z); x0.w = fmaf(s0.w, v[3], x0.w);
                x1.x = fmaf(s1.x, v[4], x1.x); x1.y = fmaf(s1.y, v[5], x1.y);
                x1.z = fmaf(s1.z, v[6], x1.z); x1.w = fmaf(s1.w, v[7], x1.w);
              } else {
                x0.x += v[0]; x0.y += v[1]; x0.z += v[2]; x0.w += v[3];
                x1.x += v[4]; x1.y += v[5]; x1.z += v[6]; x1.w += v[7];
              }
              float* op = reinterpret_cast<float*>(g.out) + orow * g.ldo + col + j;
              *reinterpret_cast<float4*>(op) = x0;
              *reinterpret_cast<float4*>(op + 4) = x1;
            } else {  // EPI_PATCH
              const float4 p0 = __ldg(reinterpret_cast<const float4*>(posb_row + col + j));
              const float4 p1 = __ldg(reinterpret_cast<const float4*>(posb_row + col + j + 4));
              float4 x0 = make_float4(v[0] + p0.x, v[1] + p0.y, v[2] + p0.z, v[3] + p0.w);
              float4 x1 = make_float4(v[4] + p1.x, v[5] + p1.y, v[6] + p1.z, v[7] + p1.w);
              float* op = reinterpret_cast<float*>(g.out) + orow * g.ldo + col + j;
              *reinterpret_cast<float4*>(op) = x0;
              *reinterpret_cast<float4*>(op + 4) = x1;
            }
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace ldit
