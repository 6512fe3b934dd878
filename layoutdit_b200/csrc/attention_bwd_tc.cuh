// Attention backward on the tensor cores (tcgen05 / TMEM), head_dim 64, up to 256 tokens per image:
//     dQ, dK, dV  from  Q, K, V (the forward's fused QKV buffer [B, N, 3D]) and dO = d ctx [B, N, D]
// for ctx = softmax(q k^T / 8) v, HF:249-306.  One CTA per (image, head); P is recomputed, nothing but QKV is
// needed from the forward:
//     S = Q K^T            P = softmax(S / 8)          delta_i = sum_j P_ij dP_ij
//     dP = dO V^T          dS = P (.) (dP - delta) / 8
//     dV = P^T dO          dK = dS^T Q                  dQ = dS K
// All five products are tcgen05.mma with fp32 accumulators in TMEM; the element-wise part runs with thread <-> query
// row, exactly as in the forward kernel.  Work is cut into (query tile t, key half kh) blocks of 128 x 128:
//   TMEM columns   [0,128) S(t,kh)   [128,256) dP(t,kh)   [256,320) dV(kh)   [320,384) dK(kh)   [384,448) dQ(0)   [448,512) dQ(1)
//   shared memory  Q, K, V, dO as two 128 x 64 K-major tiles each (TMA, 128-byte swizzle; the same tiles serve as
//                  MN-major B operands where the contraction runs over their rows), P^T and dS^T as [128 keys x 128
//                  rows] K-major (written transposed by the row threads), dS as [128 rows x 128 keys] K-major
// A statistics pre-pass computes S and dP for the whole row (all 512 columns are free then) and leaves the row max,
// 1 / row sum and delta in registers; the main pass loops kh (outer: dV / dK of one key half stay in TMEM) over t.
// 8 warps: warp w owns TMEM lanes 32 (w & 3) .. +32 and the 16-column chunks c with (c & 1) == (w >> 2).
// Correctness-first scheduling (no overlap of MMA and element-wise phases): this kernel replaces a CUDA-core
// version that took 10.7 ms per layer at B = 64, N = 197; it is not tuned beyond that.
#pragma once

#include "ptx.cuh"

namespace ldit {

struct AttnBwdArgs {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* dqkv;
  int B, N, heads, D;
  float scale_log2e, scale;
};

constexpr int kAbtThreads = 256;
constexpr int kAbtTile = 16384;                         // one 128 x 64 bf16 tile
constexpr int kAbtSmemTiles = 14 * kAbtTile;            // Q, K, V, dO (2 each) + P^T, dS^T, dS (2 atoms each)
constexpr int kAbtTmemCols = 512;

__device__ __forceinline__ float abt_ex2(float x) {   // MUFU.EX2 (2 ulp; exp2(-inf) = +0), as the forward
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// element (row r, column c) of a [128 x 64] bf16 tile with rows of 128 B and the 128-byte swizzle: byte offset
__device__ __forceinline__ uint32_t abt_sw128(int r, int c) {
  return static_cast<uint32_t>(r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1)));
}

__global__ void __launch_bounds__(kAbtThreads, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // [t]
  uint8_t* sK = smem + 2 * kAbtTile;        // [kh]
  uint8_t* sV = smem + 4 * kAbtTile;        // [kh]
  uint8_t* sDO = smem + 6 * kAbtTile;       // [t]
  uint8_t* sPT = smem + 8 * kAbtTile;       // [row atom]: 128 keys x 64 rows
  uint8_t* sDST = smem + 10 * kAbtTile;     // [row atom]
  uint8_t* sDS = smem + 12 * kAbtTile;      // [key atom]: 128 rows x 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAbtSmemTiles);
  uint64_t* bar_load = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane;                   // row of a 128-row tile = TMEM lane
  const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int NT = (a.N + 127) / 128;                      // query tiles = key halves (1 or 2)

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kAbtTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, static_cast<uint32_t>(NT) * 4u * kAbtTile);
    for (int t = 0; t < NT; ++t) {
      tma_load_3d(sQ + t * kAbtTile, &tmQKV, bar_load, h * 64, t * 128, b);
      tma_load_3d(sK + t * kAbtTile, &tmQKV, bar_load, a.D + h * 64, t * 128, b);
      tma_load_3d(sV + t * kAbtTile, &tmQKV, bar_load, 2 * a.D + h * 64, t * 128, b);
      tma_load_3d(sDO + t * kAbtTile, &tmDO, bar_load, h * 64, t * 128, b);
    }
  }
  mbar_wait(bar_load, 0);
  tcgen05_fence_after();

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);    // S, dP: both operands K-major
  constexpr uint32_t idesc_g = umma_idesc_bf16(128, 64, 0, 1);     // dV, dK, dQ: B (dO, Q, K tiles) is MN-major
  uint32_t mma_phase = 0;
  // S(t, kh) -> columns col_s, dP(t, kh) -> columns col_p  (one thread issues; the commit covers everything before it)
  auto issue_s_dp = [&](int t, int kh, uint32_t col_s, uint32_t col_p) {
    const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ + t * kAbtTile));
    const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK + kh * kAbtTile));
    const uint64_t dod = umma_desc_kmajor_sw128(smem_u32(sDO + t * kAbtTile));
    const uint64_t vd = umma_desc_kmajor_sw128(smem_u32(sV + kh * kAbtTile));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + col_s, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + col_p, dod + 2 * k, vd + 2 * k, idesc_s, k != 0);
  };
  auto wait_mma = [&]() {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tcgen05_fence_after();
  };

  // ------------------------------------------------------------------ statistics pre-pass
  float m_row[2], linv_row[2], delta_row[2];
  float* xch = reinterpret_cast<float*>(sPT);            // [2 halves][128 rows][2]: partials of the other warp of a row
  for (int t = 0; t < NT; ++t) {
    if (threadIdx.x == 0) {
      for (int kh = 0; kh < NT; ++kh) issue_s_dp(t, kh, kh * 128, 256 + kh * 128);
      tcgen05_commit(bar_mma);
    }
    wait_mma();
    const int ncols = NT * 128;
    // pass A: row max and sum over this warp's chunks, merged with the partner warp's through shared memory
    float m = -INFINITY, l = 0.f;
    for (int c = half; c < ncols / 16; c += 2) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(lane_addr + c * 16, r);
      tmem_wait_ld16(r);
      float cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float s = (c * 16 + i < a.N) ? __uint_as_float(r[i]) * a.scale_log2e : -INFINITY;
        r[i] = __float_as_uint(s);
        cm = fmaxf(cm, s);
      }
      if (cm != -INFINITY) {                             // a chunk entirely past the last key contributes nothing
        if (cm > m) { l *= abt_ex2(m - cm); m = cm; }    // first chunk: 0 * exp2(-inf) = 0
#pragma unroll
        for (int i = 0; i < 16; ++i) l += abt_ex2(__uint_as_float(r[i]) - m);
      }
    }
    xch[(half * 128 + row) * 2] = m;
    xch[(half * 128 + row) * 2 + 1] = l;
    __syncthreads();
    {
      const float m2 = xch[((half ^ 1) * 128 + row) * 2], l2 = xch[((half ^ 1) * 128 + row) * 2 + 1];
      const float mm = fmaxf(m, m2);                     // key 0 always exists: at least one of the two is finite
      l = ((m == -INFINITY) ? 0.f : l * abt_ex2(m - mm)) + ((m2 == -INFINITY) ? 0.f : l2 * abt_ex2(m2 - mm));
      m = mm;
    }
    __syncthreads();
    const float linv = 1.0f / l;
    // pass B: delta = sum_j P_ij dP_ij
    float dl = 0.f;
    for (int c = half; c < ncols / 16; c += 2) {
      uint32_t r[16], q[16];
      tmem_ld_32x32b_x16(lane_addr + c * 16, r);
      tmem_ld_32x32b_x16(lane_addr + 256 + c * 16, q);
      tmem_wait_ld16(r);
      tmem_wait_ld16(q);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c * 16 + i < a.N) dl = fmaf(abt_ex2(__uint_as_float(r[i]) * a.scale_log2e - m) * linv, __uint_as_float(q[i]), dl);
    }
    xch[half * 128 + row] = dl;
    __syncthreads();
    dl += xch[(half ^ 1) * 128 + row];
    m_row[t] = m; linv_row[t] = linv; delta_row[t] = dl;
    tcgen05_fence_before();
    __syncthreads();                                     // xch and the TMEM columns are rewritten by the next tile
  }

  // ------------------------------------------------------------------ main pass
  for (int kh = 0; kh < NT; ++kh) {
    for (int t = 0; t < NT; ++t) {
      if (threadIdx.x == 0) {
        tcgen05_fence_after();
        issue_s_dp(t, kh, 0, 128);
        tcgen05_commit(bar_mma);
      }
      wait_mma();      // ... and with it every earlier MMA: P^T / dS^T / dS of the previous block are free to rewrite
      const float m = m_row[t], linv = linv_row[t], dl = delta_row[t];
      for (int c = half; c < 8; c += 2) {
        uint32_t r[16], q[16];
        tmem_ld_32x32b_x16(lane_addr + c * 16, r);
        tmem_ld_32x32b_x16(lane_addr + 128 + c * 16, q);
        tmem_wait_ld16(r);
        tmem_wait_ld16(q);
        uint32_t dsp[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float p[2], ds[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int key = kh * 128 + c * 16 + i + e;
            p[e] = (key < a.N) ? abt_ex2(__uint_as_float(r[i + e]) * a.scale_log2e - m) * linv : 0.f;
            ds[e] = p[e] * (__uint_as_float(q[i + e]) - dl) * a.scale;
            const int kl = c * 16 + i + e;                // key inside the half = row of the transposed tiles
            const uint32_t off = static_cast<uint32_t>(row >> 6) * kAbtTile + abt_sw128(kl, row & 63);
            *reinterpret_cast<__nv_bfloat16*>(sPT + off) = __float2bfloat16_rn(p[e]);
            *reinterpret_cast<__nv_bfloat16*>(sDST + off) = __float2bfloat16_rn(ds[e]);
          }
          dsp[i >> 1] = pack_bf16x2(ds[0], ds[1]);
        }
        // dS, K-major: this thread's row, keys c*16 .. +16 = two 16-byte pieces of the row's 128-byte line in key atom c / 4
        {
          uint8_t* line = sDS + (c >> 2) * kAbtTile + row * 128;
          const int piece = (c & 3) * 2;
          *reinterpret_cast<uint4*>(line + (((piece) ^ (row & 7)) << 4)) = make_uint4(dsp[0], dsp[1], dsp[2], dsp[3]);
          *reinterpret_cast<uint4*>(line + (((piece + 1) ^ (row & 7)) << 4)) = make_uint4(dsp[4], dsp[5], dsp[6], dsp[7]);
        }
      }
      fence_proxy_async_smem();     // the generic-proxy writes above are operands of the MMAs below
      tcgen05_fence_before();
      __syncthreads();
      if (threadIdx.x == 0) {
        tcgen05_fence_after();
        const uint64_t ptd = umma_desc_kmajor_sw128(smem_u32(sPT));
        const uint64_t dstd = umma_desc_kmajor_sw128(smem_u32(sDST));
        const uint64_t dsd = umma_desc_kmajor_sw128(smem_u32(sDS));
        const uint64_t dod = umma_desc_mnmajor_sw128(smem_u32(sDO + t * kAbtTile));
        const uint64_t qd = umma_desc_mnmajor_sw128(smem_u32(sQ + t * kAbtTile));
        const uint64_t kd = umma_desc_mnmajor_sw128(smem_u32(sK + kh * kAbtTile));
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // contraction over the 128 rows of the tile, 16 per MMA; atom k / 4, 32 B per step inside it
          const uint32_t aoff = static_cast<uint32_t>((k >> 2) * (kAbtTile >> 4) + (k & 3) * 2);
          umma_bf16_ss(tmem_base + 256, ptd + aoff, dod + 128 * k, idesc_g, (t | k) != 0);    // dV(kh) += P^T dO_t
          umma_bf16_ss(tmem_base + 320, dstd + aoff, qd + 128 * k, idesc_g, (t | k) != 0);    // dK(kh) += dS^T Q_t
          umma_bf16_ss(tmem_base + 384 + 64 * t, dsd + aoff, kd + 128 * k, idesc_g, (kh | k) != 0);   // dQ(t) += dS K_kh
        }
        // no commit here: the next block's commit (or the one below) covers these
      }
    }
    // dV(kh), dK(kh): thread <-> key row
    if (threadIdx.x == 0) tcgen05_commit(bar_mma);
    wait_mma();
    {
      const int key = kh * 128 + row;
      for (int c = half; c < 8; c += 2) {                 // chunks 0-3: dV, 4-7: dK (64 columns each)
        uint32_t r[16];
        tmem_ld_32x32b_x16(lane_addr + 256 + c * 16, r);
        tmem_wait_ld16(r);
        if (key < a.N) {
          const int which = c >> 2;                       // 0 = dV -> third third of the row, 1 = dK -> second third
          __nv_bfloat16* dst = a.dqkv + (static_cast<size_t>(b) * a.N + key) * 3 * a.D + (which ? a.D : 2 * a.D) + h * 64 + (c & 3) * 16;
          uint4 v0, v1;
          v0.x = pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));   v0.y = pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
          v0.z = pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));   v0.w = pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
          v1.x = pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9]));   v1.y = pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11]));
          v1.z = pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13])); v1.w = pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15]));
          reinterpret_cast<uint4*>(dst)[0] = v0;
          reinterpret_cast<uint4*>(dst)[1] = v1;
        }
      }
    }
    tcgen05_fence_before();
    __syncthreads();     // dV / dK columns are read out before the next key half overwrites them
  }
  // dQ(t): thread <-> query row (every MMA has completed: the last commit above covered them)
  for (int t = 0; t < NT; ++t) {
    const int qrow = t * 128 + row;
    for (int c = half; c < 4; c += 2) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(lane_addr + 384 + 64 * t + c * 16, r);
      tmem_wait_ld16(r);
      if (qrow < a.N) {
        __nv_bfloat16* dst = a.dqkv + (static_cast<size_t>(b) * a.N + qrow) * 3 * a.D + h * 64 + c * 16;
        uint4 v0, v1;
        v0.x = pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));   v0.y = pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
        v0.z = pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));   v0.w = pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
        v1.x = pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9]));   v1.y = pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11]));
        v1.z = pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13])); v1.w = pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15]));
        reinterpret_cast<uint4*>(dst)[0] = v0;
        reinterpret_cast<uint4*>(dst)[1] = v1;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kAbtTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------------
// Any sequence length: the flash-style variant.  One CTA per (image, head, 128-key tile) keeps dV / dK of its keys in
// TMEM and walks the query tiles (Q / dO double-buffered through TMA); P is recomputed from the forward's row
// log-sum-exp (`lse`, written by ldit_attention_lse), delta = rowsum(dO (.) O) comes from attention_delta_kernel, and
// the dQ contribution of every (query tile, key tile) block leaves through fp32 vector reductions into `dq_acc`
// (cast to bf16 by dq_cast_kernel afterwards).  Same 128 x 128 block body as attention_bwd_tc_kernel above.
//   TMEM columns  [0,128) S   [128,256) dP   [256,320) dV   [320,384) dK   [384,448) dQ block
struct AttnBwdFlashArgs {
  const float* lse;       // [B, heads, N], log2 units
  const float* delta;     // [B, heads, N]
  float* dq_acc;          // [B*N, D] f32, zeroed by the caller of the kernel
  __nv_bfloat16* dqkv;
  int B, N, heads, D, nqt;
  float scale_log2e, scale;
  // relative-position bias (HAS_BIAS): table [heads, T] in natural units as the forward takes it, its gradient (+=)
  const float* bias_table;
  float* dbias;
  int Gh, Gw, T;
};
constexpr int kAbfSmemTiles = 12 * kAbtTile;   // K, V, Q[2], dO[2], P^T (2 atoms), dS^T (2), dS (2)
#ifndef LDIT_ABF_THREADS
#define LDIT_ABF_THREADS 512
#endif
constexpr int kAbfThreads = LDIT_ABF_THREADS;   // 16 warps: the element-wise phase between the MMAs is what a block costs

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// delta[b, h, q] = sum_d dO[b, q, h, d] * O[b, q, h, d]   (one thread per (b, q, h); rows of 128 B)
__global__ void __launch_bounds__(256)
attention_delta_kernel(const __nv_bfloat16* __restrict__ dctx, const __nv_bfloat16* __restrict__ ctx, float* __restrict__ delta,
                       int B, int N, int heads) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * N * heads) return;
  const int h = static_cast<int>(i % heads);
  const size_t bq = i / heads;
  const int q = static_cast<int>(bq % N), b = static_cast<int>(bq / N);
  const uint4* a = reinterpret_cast<const uint4*>(dctx + i * 64);
  const uint4* o = reinterpret_cast<const uint4*>(ctx + i * 64);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 ua = __ldg(a + c), uo = __ldg(o + c);
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wo[4] = {uo.x, uo.y, uo.z, uo.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc = fmaf(__uint_as_float(wa[e] << 16), __uint_as_float(wo[e] << 16), acc);
      acc = fmaf(__uint_as_float(wa[e] & 0xffff0000u), __uint_as_float(wo[e] & 0xffff0000u), acc);
    }
  }
  delta[(static_cast<size_t>(b) * heads + h) * N + q] = acc;
}

// dqkv[:, 0:D] (row pitch 3D) bf16 = dq_acc f32 [rows, D]
__global__ void __launch_bounds__(256)
dq_cast_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, size_t rows, int D) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // 8 elements each
  const int d8 = D / 8;
  if (i >= rows * d8) return;
  const size_t r = i / d8;
  const int c = static_cast<int>(i % d8) * 8;
  const float4 v0 = __ldg(reinterpret_cast<const float4*>(dq_acc + r * D + c)), v1 = __ldg(reinterpret_cast<const float4*>(dq_acc + r * D + c) + 1);
  uint4 o;
  o.x = pack_bf16x2(v0.x, v0.y); o.y = pack_bf16x2(v0.z, v0.w); o.z = pack_bf16x2(v1.x, v1.y); o.w = pack_bf16x2(v1.z, v1.w);
  *reinterpret_cast<uint4*>(dqkv + r * 3 * D + c) = o;
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(kAbfThreads, 1)
attention_bwd_flash_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttnBwdFlashArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + kAbtTile;
  uint8_t* sQ = smem + 2 * kAbtTile;        // [slot]
  uint8_t* sDO = smem + 4 * kAbtTile;       // [slot]
  uint8_t* sPT = smem + 6 * kAbtTile;
  uint8_t* sDST = smem + 8 * kAbtTile;
  uint8_t* sDS = smem + 10 * kAbtTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAbfSmemTiles);
  uint64_t* bar_kv = bars;
  uint64_t* bar_q = bars + 1;               // [slot]
  uint64_t* bar_mma = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  // HAS_BIAS: the head's table (x log2 e), its gradient accumulator, the column terms of this CTA's 128 keys (HF:522-544)
  float* sTab = reinterpret_cast<float*>(bars + 6);
  float* sGrad = sTab + (HAS_BIAS ? a.T : 0);
  int* sColK = reinterpret_cast<int*>(sGrad + (HAS_BIAS ? a.T : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, half = warp >> 2;     // 16 warps: four per TMEM lane quarter, each takes every fourth 16-column chunk
  const int row = quarter * 32 + lane;
  const int kh = blockIdx.x;
  const int b = blockIdx.y / a.heads, h = blockIdx.y % a.heads;

  if (threadIdx.x == 0) {
    mbar_init(bar_kv, 1);
    mbar_init(&bar_q[0], 1);
    mbar_init(&bar_q[1], 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kAbtTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);

  auto load_q = [&](int t, int slot) {
    mbar_arrive_expect_tx(&bar_q[slot], 2u * kAbtTile);
    tma_load_3d(sQ + slot * kAbtTile, &tmQKV, &bar_q[slot], h * 64, t * 128, b);
    tma_load_3d(sDO + slot * kAbtTile, &tmDO, &bar_q[slot], h * 64, t * 128, b);
  };
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_kv, 2u * kAbtTile);
    tma_load_3d(sK, &tmQKV, bar_kv, a.D + h * 64, kh * 128, b);
    tma_load_3d(sV, &tmQKV, bar_kv, 2 * a.D + h * 64, kh * 128, b);
    load_q(0, 0);
  }
  if constexpr (HAS_BIAS) {
    const float* tab = a.bias_table + static_cast<size_t>(h) * a.T;
    for (int i = threadIdx.x; i < a.T; i += kAbfThreads) { sTab[i] = tab[i] * 1.4426950408889634f; sGrad[i] = 0.f; }
    if (threadIdx.x < 128) {
      const int key = kh * 128 + threadIdx.x, p = key - 1;
      sColK[threadIdx.x] = (key == 0 || key >= a.N) ? 0 : (p / a.Gw) * (2 * a.Gw - 1) + (p % a.Gw);
    }
    __syncthreads();
  }
  mbar_wait(bar_kv, 0);

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_g = umma_idesc_bf16(128, 64, 0, 1);
  uint32_t mma_phase = 0;
  auto wait_mma = [&]() {
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tcgen05_fence_after();
  };
  const float* lse = a.lse + (static_cast<size_t>(b) * a.heads + h) * a.N;
  const float* dlt = a.delta + (static_cast<size_t>(b) * a.heads + h) * a.N;

  for (int t = 0; t < a.nqt; ++t) {
    const int slot = t & 1;
    // every MMA of block t-1 has completed (its dQ was read): the other slot is free for the next query tile
    if (threadIdx.x == 0 && t + 1 < a.nqt) load_q(t + 1, slot ^ 1);
    mbar_wait(&bar_q[slot], (t >> 1) & 1);
    tcgen05_fence_after();
    if (threadIdx.x == 0) {
      const uint64_t qd = umma_desc_kmajor_sw128(smem_u32(sQ + slot * kAbtTile));
      const uint64_t kd = umma_desc_kmajor_sw128(smem_u32(sK));
      const uint64_t dod = umma_desc_kmajor_sw128(smem_u32(sDO + slot * kAbtTile));
      const uint64_t vd = umma_desc_kmajor_sw128(smem_u32(sV));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + 128, dod + 2 * k, vd + 2 * k, idesc_s, k != 0);
      tcgen05_commit(bar_mma);
    }
    const int qrow = t * 128 + row;
    const float ls = qrow < a.N ? __ldg(lse + qrow) : 0.f;     // padding rows: Q = dO = 0, any finite statistic does
    const float dl = qrow < a.N ? __ldg(dlt + qrow) : 0.f;
    int rowterm = 0;
    if constexpr (HAS_BIAS) {
      if (qrow >= 1 && qrow < a.N) { const int pp = qrow - 1; rowterm = (pp / a.Gw + a.Gh - 1) * (2 * a.Gw - 1) + (pp % a.Gw) + a.Gw - 1; }
    }
    wait_mma();
    for (int c = half; c < 8; c += kAbfThreads / 128) {
      uint32_t r[16], q[16];
      tmem_ld_32x32b_x16(lane_addr + c * 16, r);
      tmem_ld_32x32b_x16(lane_addr + 128 + c * 16, q);
      tmem_wait_ld16(r);
      tmem_wait_ld16(q);
      uint32_t dsp[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float p[2], ds[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = kh * 128 + c * 16 + i + e;
          const int kl = c * 16 + i + e;
          float logit = __uint_as_float(r[i + e]) * a.scale_log2e;
          int idx = 0;
          if constexpr (HAS_BIAS) {     // the index rule of HF:522-544: CLS row / column entries T-3, T-2, T-1
            idx = (qrow == 0 || qrow >= a.N) ? (key == 0 ? a.T - 1 : a.T - 3) : (key == 0 ? a.T - 2 : rowterm - sColK[kl]);
            logit += sTab[idx];
          }
          p[e] = (key < a.N) ? abt_ex2(logit - ls) : 0.f;
          const float dsn = p[e] * (__uint_as_float(q[i + e]) - dl);     // d loss / d logit (natural units) = d loss / d bias
          ds[e] = dsn * a.scale;
          if constexpr (HAS_BIAS) {
            if (key < a.N && qrow < a.N) atomicAdd(&sGrad[idx], dsn);
          }
          const uint32_t off = static_cast<uint32_t>(row >> 6) * kAbtTile + abt_sw128(kl, row & 63);
          *reinterpret_cast<__nv_bfloat16*>(sPT + off) = __float2bfloat16_rn(p[e]);
          *reinterpret_cast<__nv_bfloat16*>(sDST + off) = __float2bfloat16_rn(ds[e]);
        }
        dsp[i >> 1] = pack_bf16x2(ds[0], ds[1]);
      }
      uint8_t* line = sDS + (c >> 2) * kAbtTile + row * 128;
      const int piece = (c & 3) * 2;
      *reinterpret_cast<uint4*>(line + (((piece) ^ (row & 7)) << 4)) = make_uint4(dsp[0], dsp[1], dsp[2], dsp[3]);
      *reinterpret_cast<uint4*>(line + (((piece + 1) ^ (row & 7)) << 4)) = make_uint4(dsp[4], dsp[5], dsp[6], dsp[7]);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tcgen05_fence_after();
      const uint64_t ptd = umma_desc_kmajor_sw128(smem_u32(sPT));
      const uint64_t dstd = umma_desc_kmajor_sw128(smem_u32(sDST));
      const uint64_t dsd = umma_desc_kmajor_sw128(smem_u32(sDS));
      const uint64_t dod = umma_desc_mnmajor_sw128(smem_u32(sDO + slot * kAbtTile));
      const uint64_t qd = umma_desc_mnmajor_sw128(smem_u32(sQ + slot * kAbtTile));
      const uint64_t kd = umma_desc_mnmajor_sw128(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t aoff = static_cast<uint32_t>((k >> 2) * (kAbtTile >> 4) + (k & 3) * 2);
        umma_bf16_ss(tmem_base + 256, ptd + aoff, dod + 128 * k, idesc_g, (t | k) != 0);    // dV += P^T dO_t
        umma_bf16_ss(tmem_base + 320, dstd + aoff, qd + 128 * k, idesc_g, (t | k) != 0);    // dK += dS^T Q_t
        umma_bf16_ss(tmem_base + 384, dsd + aoff, kd + 128 * k, idesc_g, k != 0);           // dQ block = dS K
      }
      tcgen05_commit(bar_mma);
    }
    wait_mma();
    // dQ block -> fp32 reductions into dq_acc (this warp's half of the 64 columns)
    for (int c = half; c < 4; c += kAbfThreads / 128) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(lane_addr + 384 + c * 16, r);
      tmem_wait_ld16(r);
      if (qrow < a.N) {
        float* dst = a.dq_acc + (static_cast<size_t>(b) * a.N + qrow) * a.D + h * 64 + c * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          red_add_v4(dst + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
    }
    tcgen05_fence_before();
    __syncthreads();     // S / dP / dQ columns and the operand tiles are free for the next block
  }
  // dV, dK of this key tile: thread <-> key row (every MMA completed: the last wait above)
  {
    const int key = kh * 128 + row;
    for (int c = half; c < 8; c += kAbfThreads / 128) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(lane_addr + 256 + c * 16, r);
      tmem_wait_ld16(r);
      if (key < a.N) {
        const int which = c >> 2;
        __nv_bfloat16* dst = a.dqkv + (static_cast<size_t>(b) * a.N + key) * 3 * a.D + (which ? a.D : 2 * a.D) + h * 64 + (c & 3) * 16;
        uint4 v0, v1;
        v0.x = pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));   v0.y = pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
        v0.z = pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));   v0.w = pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
        v1.x = pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9]));   v1.y = pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11]));
        v1.z = pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13])); v1.w = pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15]));
        reinterpret_cast<uint4*>(dst)[0] = v0;
        reinterpret_cast<uint4*>(dst)[1] = v1;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (HAS_BIAS) {   // this CTA's share of the table gradient (every shared-memory atomic above is behind the barrier)
    float* dst = a.dbias + static_cast<size_t>(h) * a.T;
    for (int i = threadIdx.x; i < a.T; i += kAbfThreads) {
      const float v = sGrad[i];
      if (v != 0.f) atomicAdd(dst + i, v);
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, kAbtTmemCols);
}

}  // namespace ldit
