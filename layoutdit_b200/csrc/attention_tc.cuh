// tcgen05 fused attention for sm_100a (head_dim 64):  ctx = softmax(q k^T / 8 + bias) v
// per (image, head), HF:249-306 / HF:310-368, reading Q/K/V in place from the fused QKV GEMM
// output [B, N, 3D] and writing the merged-heads context [B, N, D] (HF:365-367).
//
// CTA = one (image, head, 128-query tile); 4 softmax warps (thread <-> query row <-> TMEM lane)
// + 1 warp that issues TMA and tcgen05.mma.  256 TMEM columns per CTA, two CTAs per SM, so one
// CTA's loads / MMAs run under the other CTA's exponentials.
//   S = Q K^T   tcgen05.mma SS: A = Q tile (K-major, TMA 128B swizzle), B = K tile (K-major),
//               D = fp32 [128 x kv_tile] in TMEM columns [0, kv_tile), kv_tile <= 208
//   softmax     registers: pass 1 row max, pass 2 exp2 / row sum; relative-position bias is
//               gathered in-tile from the per-head table in smem (index rule of HF:522-544);
//               P goes back to TMEM as packed bf16 over the S columns it came from
//   O = P V     tcgen05.mma TS: A = P from TMEM, B = V tile exactly as it sits in the QKV buffer
//               (rows = keys: an MN-major operand), D = fp32 [128 x 64] in TMEM columns [128, 192)
//   KV tiles    N <= 208 (224x224: N = 197) is ONE tile: a plain exact softmax, no rescaling.
//               Longer sequences (512x512: N = 1025 -> 5 x 208) use the online-softmax
//               recurrence with the running O kept in registers.
#pragma once

#include "ptx.cuh"

namespace ldit {

struct AttnTcArgs {
  __nv_bfloat16* ctx;       // [B*N, D]
  const float* bias_table;  // [heads, T] fp32 or nullptr
  int B, N, heads, D;
  int Gh, Gw, T;
  int kv_tile, n_kv_tiles;
  float scale_log2e;
};

constexpr int kAtcThreads = 160;
constexpr int kAtcQ = 128;
constexpr int kAtcMaxKv = 208;
constexpr int kAtcTmemCols = 256;
constexpr int kAtcOCol = 128;  // O accumulator columns [128, 192): the dead upper part of S

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(kAtcThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnTcArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // 128 x 128 B
  uint8_t* sK = smem + 16384;                          // kv_tile x 128 B (<= 26 KB)
  uint8_t* sV = sK + kAtcMaxKv * 128;                  // kv_tile x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAtcMaxKv * 128);
  uint64_t* bar_k = bars + 0;   // K tile (and Q the first time) landed
  uint64_t* bar_v = bars + 1;   // V tile landed
  uint64_t* bar_s = bars + 2;   // S = Q K^T complete (also: K buffer reusable)
  uint64_t* bar_p = bars + 3;   // P written to TMEM by all 4 softmax warps
  uint64_t* bar_o = bars + 4;   // O = P V complete (also: V buffer and P columns reusable)
  uint64_t* bar_r = bars + 5;   // O read back by all 4 softmax warps: TMEM columns reusable by the next S
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* sTab = reinterpret_cast<float*>(bars + 8);
  int* sCol = reinterpret_cast<int*>(sTab + (HAS_BIAS ? a.T : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAtcQ, h = blockIdx.y, b = blockIdx.z;
  const int kvt = a.kv_tile;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], (i == 3 || i == 5) ? 4 : 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kAtcTmemCols);
    tmem_relinquish();
  }
  if constexpr (HAS_BIAS) {
    const float* tab = a.bias_table + static_cast<size_t>(h) * a.T;
    for (int i = threadIdx.x; i < a.T; i += kAtcThreads) sTab[i] = tab[i] * 1.4426950408889634f;
    for (int k = threadIdx.x; k < a.N; k += kAtcThreads) {
      const int p = k - 1;
      sCol[k] = (k == 0) ? 0 : (p / a.Gw) * (2 * a.Gw - 1) + (p % a.Gw);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t kv_bytes = static_cast<uint32_t>(kvt) * 128;
      constexpr uint32_t idesc_s_base = umma_idesc_bf16(kAtcQ, 0, 0, 0);       // N filled in below
      const uint32_t idesc_s = idesc_s_base | (static_cast<uint32_t>(kvt >> 3) << 17);
      constexpr uint32_t idesc_o = umma_idesc_bf16(kAtcQ, 64, 0, 1);           // B = V is MN-major
      mbar_arrive_expect_tx(bar_k, 16384 + kv_bytes);
      tma_load_3d(sQ, &tmQ, bar_k, h * 64, q0, b);
      tma_load_3d(sK, &tmKV, bar_k, a.D + h * 64, 0, b);
      mbar_arrive_expect_tx(bar_v, kv_bytes);
      tma_load_3d(sV, &tmKV, bar_v, 2 * a.D + h * 64, 0, b);
      const uint64_t qdesc = umma_desc_kmajor_sw128(smem_u32(sQ));
      const uint64_t kdesc = umma_desc_kmajor_sw128(smem_u32(sK));
      const uint64_t vdesc = umma_desc_mnmajor_sw128(smem_u32(sV));
      for (int t = 0; t < a.n_kv_tiles; ++t) {
        const uint32_t ph = t & 1;
        mbar_wait(bar_k, ph);
        if (t > 0) mbar_wait(bar_r, ph ^ 1);  // O of tile t-1 (columns 128..191, inside the S range) has been read back
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        tcgen05_commit(bar_s);
        if (t + 1 < a.n_kv_tiles) {  // K buffer is free once S is complete
          mbar_wait(bar_s, ph);
          mbar_arrive_expect_tx(bar_k, kv_bytes);
          tma_load_3d(sK, &tmKV, bar_k, a.D + h * 64, (t + 1) * kvt, b);
        }
        mbar_wait(bar_p, ph);
        mbar_wait(bar_v, ph);
        tcgen05_fence_after();
        for (int k = 0; k < kvt / 16; ++k)
          umma_bf16_ts(tmem_base + kAtcOCol, tmem_base + 8 * k, vdesc + 128 * k, idesc_o, k != 0);
        tcgen05_commit(bar_o);
        if (t + 1 < a.n_kv_tiles) {  // V buffer is free once O is complete
          mbar_wait(bar_o, ph);
          mbar_arrive_expect_tx(bar_v, kv_bytes);
          tma_load_3d(sV, &tmKV, bar_v, 2 * a.D + h * 64, (t + 1) * kvt, b);
        }
      }
    }
  } else {
    const int row = warp * 32 + lane;      // row of the q tile == TMEM lane
    const int q = q0 + row;                // token index inside the image
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    int rowterm = 0;
    if constexpr (HAS_BIAS) {
      if (q >= 1) { const int p = q - 1; rowterm = (p / a.Gw + a.Gh - 1) * (2 * a.Gw - 1) + (p % a.Gw) + a.Gw - 1; }
    }
    auto biased = [&](float s, int kc) -> float {   // kc = key index inside the image (kc < N)
      float v = s * a.scale_log2e;
      if constexpr (HAS_BIAS) {
        int idx;
        if (q == 0) idx = (kc == 0) ? a.T - 1 : a.T - 3;
        else if (kc == 0) idx = a.T - 2;
        else idx = rowterm - sCol[kc];
        if (q < a.N) v += sTab[idx];
      }
      return v;
    };
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const int nchunks = (kvt + 31) / 32;   // 32-column chunks; the last may reach past kv_tile (masked)
    // a warp whose 32 query rows all lie past the end of the image only keeps the barriers in step
    const bool warp_active = (q0 + warp * 32) < a.N;
    const float sc = a.scale_log2e;

    for (int t = 0; t < a.n_kv_tiles; ++t) {
      const uint32_t ph = t & 1;
      const int k0 = t * kvt;
      const int valid = min(kvt, a.N - k0);        // keys of this tile that exist
      const int nfull = valid >> 5;                // chunks in which every key exists
      mbar_wait(bar_s, ph);
      tcgen05_fence_after();
      float alpha = 0.f;
      if (warp_active) {
        // ---- pass 1: row max.  Without bias the max is taken on the raw scores (scale > 0).
        float mx = -INFINITY;
        for (int c = 0; c < nchunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + c * 32, r);
          tcgen05_wait_ld();
          if constexpr (!HAS_BIAS) {
            if (c < nfull) {
              float m0 = fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
              float m1 = fmax3(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
              float m2 = fmax3(__uint_as_float(r[6]), __uint_as_float(r[7]), __uint_as_float(r[8]));
              float m3 = fmax3(__uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
#pragma unroll
              for (int i = 12; i < 32; i += 8) {
                m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                if (i + 4 < 32) {
                  m2 = fmax3(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
                  m3 = fmax3(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
                }
              }
              mx = fmax3(mx, fmaxf(m0, m1), fmaxf(m2, m3));
            } else {
              const int rem = valid - c * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < rem) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
          } else {
            const int rem = valid - c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < rem) mx = fmaxf(mx, biased(__uint_as_float(r[i]), k0 + c * 32 + i));
          }
        }
        if constexpr (!HAS_BIAS) mx *= sc;
        const float m_new = fmaxf(m_run, mx);
        alpha = fast_exp2(m_run - m_new);   // first tile: exp2(-inf) = 0
        m_run = m_new;
        // ---- pass 2: p = exp2(s - m), row sum, P -> TMEM as packed bf16 over columns [0, kv_tile/2)
        float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
        const float neg_m = -m_new;
        for (int c = 0; c < nchunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + c * 32, r);
          tcgen05_wait_ld();
          uint32_t pk[16];
          if (!HAS_BIAS && c < nfull) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), sc, neg_m));
              const float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), sc, neg_m));
              const float p2 = fast_exp2(fmaf(__uint_as_float(r[i + 2]), sc, neg_m));
              const float p3 = fast_exp2(fmaf(__uint_as_float(r[i + 3]), sc, neg_m));
              ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
              pk[i >> 1] = pack_bf16x2(p0, p1);
              pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            }
          } else {
            const int rem = valid - c * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = 0.f, p1 = 0.f;
              if (i < rem) p0 = fast_exp2(biased(__uint_as_float(r[i]), k0 + c * 32 + i) + neg_m);
              if (i + 1 < rem) p1 = fast_exp2(biased(__uint_as_float(r[i + 1]), k0 + c * 32 + i + 1) + neg_m);
              ps0 += p0; ps1 += p1;
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st_32x32b_x16(lane_addr + c * 16, pk);
        }
        l_run = l_run * alpha + ((ps0 + ps1) + (ps2 + ps3));
        tcgen05_wait_st();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      // O_tile = P V ; running O (registers) = alpha * O + O_tile
      mbar_wait(bar_o, ph);
      tcgen05_fence_after();
      if (warp_active) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + kAtcOCol + c * 32, r);
          tcgen05_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(r[i]));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_r);
    }

    if (q < a.N) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* dst = a.ctx + (static_cast<size_t>(b) * a.N + q) * a.D + h * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 v;
        v.x = pack_bf16x2(o[8 * j] * inv, o[8 * j + 1] * inv);
        v.y = pack_bf16x2(o[8 * j + 2] * inv, o[8 * j + 3] * inv);
        v.z = pack_bf16x2(o[8 * j + 4] * inv, o[8 * j + 5] * inv);
        v.w = pack_bf16x2(o[8 * j + 6] * inv, o[8 * j + 7] * inv);
        reinterpret_cast<uint4*>(dst)[j] = v;
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kAtcTmemCols);
}

}  // namespace ldit
