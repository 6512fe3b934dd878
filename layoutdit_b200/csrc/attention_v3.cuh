// Persistent tcgen05 attention for sm_100a (head_dim 64), two CTAs per SM:
//     ctx = softmax(q k^T / 8 + bias) v      per (image, head)          HF:249-306 / HF:310-368
// reading Q/K/V in place from the fused QKV GEMM output [B, N, 3D] and writing the
// merged-heads context [B, N, D] (HF:365-367).
//
// What bounds this op at DiT's sizes is not the tensor pipe but the softmax: ~6 issue slots and
// 0.6 MUFU operations per (row, key), executed by one thread per query row, with the FMA, ALU and XU
// pipes loaded about equally.  Throughput is therefore set by (a) how many softmax warps an SM
// sub-partition can switch between and (b) whether a warp ever has to wait for the tensor pipe.
//   (a) The kernel is sized to be resident TWICE per SM -- 256 TMEM columns, ~100 KB of shared
//       memory, <= 96 registers per thread -- i.e. 16 softmax warps per SM (4 per scheduler).
//   (b) K/V stream through smem in tiles of 32 keys and S = Q K^T is DOUBLE-BUFFERED inside a
//       warpgroup's 64 score columns: S(t+1) is already complete when the warpgroup finishes the
//       exponentials of S(t), so the MMA round trip is off the warpgroup's serial chain.
//
// One CTA loops over work items (image, head, pair of 128-query tiles):
//   warps 0-3   softmax warpgroup 0  (thread <-> query row <-> TMEM lane)
//   warps 4-7   softmax warpgroup 1
//   warp  8     TMA producer + TMEM allocator
//   warp  9     MMA issuer
// Per warpgroup 128 TMEM columns: S / P buffers [0, 32) and [32, 64), O in [64, 128).  Per step t
//   S(t) = Q K(t)^T    tcgen05.mma SS, fp32 [128 x 32] into buffer t & 1 (issued behind P V (t-2))
//   softmax            pass 1: row max straight out of TMEM; pass 2: exp2, row sum, P packed to bf16
//                      back into TMEM over the first 16 columns of the same buffer
//   O += P(t) V(t)     tcgen05.mma TS (P from TMEM; V MN-major exactly as it sits in the QKV buffer)
// The running max is LAZY: O and the row sum are only rescaled when the max grows by more than
// 2^8 (p then stays <= 256, harmless in bf16 / fp32), so most steps never touch O; when one does it
// first waits for P V (t-1).  6 of every 16 exponentials run on the FMA pipe (cubic in packed fp32),
// the bf16 packing of P on the integer pipe: MUFU.EX2 and F2FP share the quarter-rate XU pipe.
// The finished O tile is staged in the (now dead) Q tile of the item and leaves with one TMA store
// per warp; Q is double-buffered across items.
#pragma once

#include "ptx.cuh"

namespace ldit {

struct AttnV3Args {
  const float* bias_table;  // [heads, T] fp32 or nullptr
  int B, N, heads, D;
  int Gh, Gw, T;
  int n_ktiles, n_qpairs, num_items;
  float scale_log2e;
  float* lse;               // optional [B, heads, N]: log2-sum-exp of every row's scaled logits (what the backward needs to recompute P)
  long long* dbg;           // LDIT_A3_TIMELINE builds only: clock64 stamps [CTA < 8][warp 0..9][512]
};

constexpr int kA3Threads = 320;
constexpr int kA3KT = 32;                        // keys per K/V tile (= per softmax step)
constexpr int kA3Chunks = kA3KT / 16;
constexpr int kA3TileBytes = kA3KT * 128;        // 4 KB
constexpr int kA3KStages = 4, kA3VStages = 4;
constexpr int kA3QBytes = 4 * 16384;             // Q[buffer][warpgroup], 128 rows x 128 B each
constexpr int kA3SmemTiles = kA3QBytes + (kA3KStages + kA3VStages) * kA3TileBytes;
constexpr int kA3WarpProducer = 8, kA3WarpMma = 9;
constexpr int kA3OCol = 64;
constexpr int kA3TmemCols = 256;
constexpr float kA3LazyLog2 = 8.0f;              // rescale only when the row max grows by more than this (log2 units)
// barrier slots
constexpr int kA3BarQFull = 0, kA3BarQEmpty = 2, kA3BarKFull = 4, kA3BarKEmpty = kA3BarKFull + kA3KStages,
              kA3BarVFull = kA3BarKEmpty + kA3KStages, kA3BarVEmpty = kA3BarVFull + kA3VStages,
              kA3BarSFull = kA3BarVEmpty + kA3VStages,   // [warpgroup][S buffer]
              kA3BarPFull = kA3BarSFull + 4,     // [warpgroup][S buffer]
              kA3BarPvDone = kA3BarPFull + 4,    // [warpgroup][S buffer]
              kA3NumBars = kA3BarPvDone + 4;

// length of a warpgroup's bias table in shared memory: T entries + (largest column term + 1) copies of entry T - 3,
// rounded up to 4 floats (the column terms behind it are read as int4)
__host__ __device__ inline int a3_ext_len(int Gh, int Gw, int T) {
  return (T + (Gh - 1) * (2 * Gw - 1) + Gw - 1 + 1 + 3) & ~3;
}
__device__ __forceinline__ float a3_fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// barrier of one softmax warpgroup (128 threads); immediate ids, so ptxas reserves 3 named barriers and not all 16
__device__ __forceinline__ void a3_wg_sync(int g) {
  if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
__device__ __forceinline__ float a3_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bf16x2 packing of P: 0 = integer pipe (2 adds + PRMT, round-half-up), 1 = F2FP (cvt.rn.bf16x2.f32)
#ifndef LDIT_A3_PACK
#define LDIT_A3_PACK 1
#endif
// pairs of every 8 (= 16 keys) whose exponentials run on the FMA pipe instead of MUFU
#ifndef LDIT_A3_POLY_PAIRS
#define LDIT_A3_POLY_PAIRS 3
#endif

#ifdef LDIT_A3_TIMELINE
#define A3_STAMP() do { if (tl && tln < 512) tl[tln++] = clock64(); } while (0)
#else
#define A3_STAMP() do { } while (0)
#endif

template <bool HAS_BIAS>
__global__ void __launch_bounds__(kA3Threads, 2)
attention_v3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmO, const AttnV3Args a) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [qb][g] 16 KB each; doubles as the O staging of the item
  uint8_t* sK = smem + kA3QBytes;                       // [stage] 4 KB
  uint8_t* sV = sK + kA3KStages * kA3TileBytes;         // [stage] 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kA3SmemTiles);
  uint64_t* q_full = bars + kA3BarQFull;
  uint64_t* q_empty = bars + kA3BarQEmpty;    // 8 arrivals: every softmax warp, once its O store has left the buffer
  uint64_t* k_full = bars + kA3BarKFull;
  uint64_t* k_empty = bars + kA3BarKEmpty;
  uint64_t* v_full = bars + kA3BarVFull;
  uint64_t* v_empty = bars + kA3BarVEmpty;
  uint64_t* s_full = bars + kA3BarSFull;      // [warpgroup][S buffer]: scores complete
  uint64_t* p_full = bars + kA3BarPFull;      // [warpgroup][S buffer]: P written, O rescaled (4 warp arrivals).  One barrier per
                                              // buffer: warps of a warpgroup may be a step apart, their arrivals for steps t and t+1 must not mix
  uint64_t* pv_done = bars + kA3BarPvDone;    // [warpgroup][S buffer]: P V of a step on that buffer (and everything issued before it) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kA3NumBars);
  // relative-position bias: per warpgroup the head's table (x log2 e) followed by a flat run of the "CLS row" entry, so
  // that bias(q, k) = sTab[rowterm(q) - sCol[k]] holds for EVERY row without a per-element special case (a3_ext_len)
  float* sTab = reinterpret_cast<float*>(bars + kA3NumBars + 2);  // [2][ext] (one copy per warpgroup)
  const int ext = HAS_BIAS ? a3_ext_len(a.Gh, a.Gw, a.T) : 0;
  int* sCol = reinterpret_cast<int*>(sTab + 2 * ext);             // [n_ktiles * 32]: column term of key k (0 past the last key)

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == kA3WarpMma && lane == 0) {
    for (int i = 0; i < kA3NumBars; ++i) {
      const bool q_e = (i == kA3BarQEmpty || i == kA3BarQEmpty + 1);
      const bool p_f = (i >= kA3BarPFull && i < kA3BarPFull + 4);
      mbar_init(&bars[i], q_e ? 8 : (p_f ? 4 : 1));
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == kA3WarpProducer) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      tma_prefetch_desc(&tmO);
    }
    tmem_alloc(tmem_slot, kA3TmemCols);
    tmem_relinquish();
  }
  if constexpr (HAS_BIAS) {
    for (int k = threadIdx.x; k < a.n_ktiles * kA3KT; k += kA3Threads) {
      const int p = k - 1;
      sCol[k] = (k == 0 || k >= a.N) ? 0 : (p / a.Gw) * (2 * a.Gw - 1) + (p % a.Gw);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_wait();   // the prologue above touched no global memory; everything below may
  const uint32_t tmem_base = *tmem_slot;
  const int T = a.n_ktiles;
#ifdef LDIT_A3_TIMELINE
  long long* tl = (a.dbg != nullptr && blockIdx.x < 8 && lane == 0) ? a.dbg + (static_cast<size_t>(blockIdx.x) * 10 + warp) * 512 : nullptr;
  int tln = 0;
#endif

  if (warp == kA3WarpProducer) {
    // ------------------------------------------------------------------ TMA producer
    // order of use by the issuer: K(0), K(1), then per step t: V(t), K(t+2)
    uint32_t it = 0, kcount = 0, vcount = 0;
    auto load_kv = [&](uint8_t* ring, uint64_t* full, uint64_t* empty, uint32_t cnt, int stages, int col, int t, int b) {
      const uint32_t st = cnt % stages, ph = (cnt / stages) & 1;
      mbar_wait_relaxed(&empty[st], ph ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[st], kA3TileBytes);
        tma_load_3d(ring + st * kA3TileBytes, &tmKV, &full[st], col, t * kA3KT, b);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
      const int p = item % a.n_qpairs, bh = item / a.n_qpairs;
      const int h = bh % a.heads, b = bh / a.heads;
      const int nvalid = (256 * p + 128 < a.N) ? 2 : 1;
      const uint32_t qb = it & 1, qph = (it >> 1) & 1;
      mbar_wait_relaxed(&q_empty[qb], qph ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&q_full[qb], 16384u * nvalid);
        tma_load_3d(sQ + (qb * 2 + 0) * 16384, &tmQ, &q_full[qb], h * 64, 256 * p, b);
        if (nvalid == 2) tma_load_3d(sQ + (qb * 2 + 1) * 16384, &tmQ, &q_full[qb], h * 64, 256 * p + 128, b);
      }
      __syncwarp();
      for (int j = 0; j < T + 2; ++j) {
        if (j < T) { load_kv(sK, k_full, k_empty, kcount, kA3KStages, a.D + h * 64, j, b); ++kcount; }
        if (j >= 2) { load_kv(sV, v_full, v_empty, vcount, kA3VStages, 2 * a.D + h * 64, j - 2, b); ++vcount; }
      }
    }
  } else if (warp == kA3WarpMma) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);  // B = V is MN-major
    const uint64_t qdesc0 = umma_desc_kmajor_sw128(smem_u32(sQ));
    const uint64_t kdesc0 = umma_desc_kmajor_sw128(smem_u32(sK));
    const uint64_t vdesc0 = umma_desc_mnmajor_sw128(smem_u32(sV));
    uint32_t it = 0, kcount = 0, vcount = 0;
    uint32_t pbits = 0;          // parity of the next phase of p_full[g][S buffer], one bit each
    auto issue_s = [&](int g, uint32_t qb, uint32_t ks, int sbuf, int nch) {
      if (elect_one_sync()) {
        const uint32_t idesc_s = umma_idesc_bf16(128, 16, 0, 0) + (static_cast<uint32_t>((nch - 1) * 2) << 17);   // N = 16 nch
        const uint64_t qd = qdesc0 + static_cast<uint32_t>((qb * 2 + g) * (16384 >> 4));
        const uint64_t kd = kdesc0 + static_cast<uint32_t>(ks * (kA3TileBytes >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + g * 128 + sbuf * kA3KT, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        tcgen05_commit(&s_full[g * 2 + sbuf]);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
      const int p = item % a.n_qpairs;
      const int nvalid = (256 * p + 128 < a.N) ? 2 : 1;
      const uint32_t qb = it & 1, qph = (it >> 1) & 1;
      mbar_wait(&q_full[qb], qph);
      // S(0) and S(1): both buffers are free once the previous item's last P V has been issued (the tensor pipe
      // executes in order); the O columns are not touched before P V (0)
      for (int j = 0; j < 2 && j < T; ++j, ++kcount) {
        const uint32_t ks = kcount % kA3KStages, kph = (kcount / kA3KStages) & 1;
        mbar_wait(&k_full[ks], kph);
        tcgen05_fence_after();
        const int nch = (min(kA3KT, a.N - j * kA3KT) + 15) >> 4;
#pragma unroll
        for (int g = 0; g < 2; ++g)
          if (g < nvalid) issue_s(g, qb, ks, j, nch);
        if (elect_one_sync()) tcgen05_commit(&k_empty[ks]);
        __syncwarp();
      }
      for (int j = 0; j < T; ++j, ++vcount) {
        const uint32_t vs = vcount % kA3VStages, vph = (vcount / kA3VStages) & 1;
        const uint32_t ks2 = kcount % kA3KStages, kph2 = (kcount / kA3KStages) & 1;   // K(j+2), if there is one
        const bool more = j + 2 < T;
        A3_STAMP();
        mbar_wait(&v_full[vs], vph);
        if (more) mbar_wait(&k_full[ks2], kph2);
        A3_STAMP();
        const int nchv = (min(kA3KT, a.N - j * kA3KT) + 15) >> 4;
        const int nchn = more ? (min(kA3KT, a.N - (j + 2) * kA3KT) + 15) >> 4 : 0;
        const int sbuf = j & 1;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g < nvalid) {
            mbar_wait(&p_full[g * 2 + sbuf], (pbits >> (g * 2 + sbuf)) & 1);
            pbits ^= 1u << (g * 2 + sbuf);
            A3_STAMP();
            tcgen05_fence_after();
            if (elect_one_sync()) {
              const uint64_t vd = vdesc0 + static_cast<uint32_t>(vs * (kA3TileBytes >> 4));
#pragma unroll
              for (int k = 0; k < kA3Chunks; ++k)
                if (k < nchv)   // chunks past the last key were never written by the softmax warps
                  umma_bf16_ts(tmem_base + g * 128 + kA3OCol, tmem_base + g * 128 + sbuf * kA3KT + 8 * k, vd + 128 * k, idesc_o, (j | k) != 0);
              tcgen05_commit(&pv_done[g * 2 + sbuf]);
            }
            __syncwarp();
            if (more) issue_s(g, qb, ks2, sbuf, nchn);   // in order behind P V (j): S(j+2) may overwrite the P(j) columns
          }
        }
        if (elect_one_sync()) {
          tcgen05_commit(&v_empty[vs]);
          if (more) tcgen05_commit(&k_empty[ks2]);
        }
        __syncwarp();
        if (more) ++kcount;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int g = warp >> 2;
    const int quarter = warp & 3;
    const uint32_t lane_addr = tmem_base + g * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
    float* myTab = sTab + g * ext;
    const int cmax = (a.Gh - 1) * (2 * a.Gw - 1) + a.Gw - 1;   // largest column term
    const float sc = a.scale_log2e;
    uint32_t sbits = 0;          // parity of the next phase of s_full[g][S buffer], one bit each
    // pv_done[g][b] completes once per step on buffer b.  A warp waits on it only when it needs O (a rescale, the
    // epilogue); in between it observes nothing, which is safe: when step j is being processed the barrier of buffer
    // (j-1)&1 has completed either all its phases up to P V (j-1) or all but that one (P V (j-3) is implied by
    // s_full(j-1), P V (j+1) cannot be issued before this warp's arrival for step j+1), so a parity wait cannot alias.
    uint32_t pvuses = 0;         // bit b: parity of the number of P V's issued so far on buffer b (for this warpgroup)
    uint32_t it = 0;
    int cur_h = -1;
    int pending_qb = -1;         // Q buffer whose O store of the previous item may still be reading its staging
    const uint64_t magic2 = f2_splat(12582912.0f), nmagic2 = f2_splat(-12582912.0f), mone2 = f2_splat(-1.0f);
    const uint64_t e3 = f2_splat(0.05508868396282196f), e2 = f2_splat(0.24260404706001282f),
                   e1 = f2_splat(0.6932762265205383f), e0 = f2_splat(0.9999289512634277f);
    for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
      const int p = item % a.n_qpairs, bh = item / a.n_qpairs;
      const int h = bh % a.heads, b = bh / a.heads;
      const int nvalid = (256 * p + 128 < a.N) ? 2 : 1;
      const uint32_t qb = it & 1;
      if (g >= nvalid) {   // nothing of this item belongs to this warpgroup; its buffer slot is free as far as it is concerned
        if (lane == 0) {
          if (pending_qb >= 0) {
            tma_store_wait_read<0>();
            mbar_arrive(&q_empty[pending_qb]);
          }
          mbar_arrive(&q_empty[qb]);
        }
        pending_qb = -1;
        continue;
      }
      const int row0 = 256 * p + 128 * g;
      const int q = row0 + quarter * 32 + lane;          // token index inside the image
      const bool warp_active = (row0 + quarter * 32) < a.N;
      int rowterm = 0;
      if constexpr (HAS_BIAS) {
        if (h != cur_h) {  // (re)load this head's table; 128 threads of the warpgroup
          a3_wg_sync(g);
          const float* tab = a.bias_table + static_cast<size_t>(h) * a.T;
          for (int i = quarter * 32 + lane; i < ext; i += 128) myTab[i] = tab[i < a.T ? i : a.T - 3] * 1.4426950408889634f;
          a3_wg_sync(g);
          cur_h = h;
        }
        // patch rows: (row term) - (column term) is the index rule of HF:522-544; the CLS row (q = 0) and the padding rows
        // past N point at the flat run behind the table, which holds the "CLS to token" entry T - 3 for every column
        rowterm = a.T + cmax;
        if (q >= 1 && q < a.N) { const int pp = q - 1; rowterm = (pp / a.Gw + a.Gh - 1) * (2 * a.Gw - 1) + (pp % a.Gw) + a.Gw - 1; }
      }
      float m_run = -INFINITY, l_run = 0.f;

#pragma unroll 1
      for (int j = 0; j < T; ++j) {
        const int k0 = j * kA3KT;
        const int valid = min(kA3KT, a.N - k0);   // keys of this tile that exist
        const int nch = (valid + 15) >> 4;        // 16-key chunks that hold at least one of them
        const int sbuf = j & 1;
        const uint32_t s_addr = lane_addr + sbuf * kA3KT;
        A3_STAMP();
        mbar_wait(&s_full[g * 2 + sbuf], (sbits >> sbuf) & 1);
        A3_STAMP();
        sbits ^= 1u << sbuf;
        tcgen05_fence_after();
        if (warp_active) {
          // ---- pass 1: row max (bias added and written back first, when there is one)
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
          uint32_t r[2][16];
          tmem_ld_32x32b_x16(s_addr, r[0]);
#pragma unroll
          for (int c = 0; c < kA3Chunks; ++c) {
            if (c < nch) {
              uint32_t (&rc)[16] = r[c & 1];
              tmem_wait_ld16(rc);
              if (c + 1 < nch) tmem_ld_32x32b_x16(s_addr + (c + 1) * 16, r[(c + 1) & 1]);
              const bool partial = (c + 1) * 16 > valid;
              if constexpr (HAS_BIAS) {
                // logits = s * scale * log2 e + bias: one table load and one FMA per element; the column terms of four
                // keys come with one broadcast 16-byte load
                const int4* colv = reinterpret_cast<const int4*>(sCol + k0 + c * 16);
                const float s00 = __uint_as_float(rc[0]);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const int4 cc = colv[i4];
                  rc[4 * i4 + 0] = __float_as_uint(fmaf(__uint_as_float(rc[4 * i4 + 0]), sc, myTab[rowterm - cc.x]));
                  rc[4 * i4 + 1] = __float_as_uint(fmaf(__uint_as_float(rc[4 * i4 + 1]), sc, myTab[rowterm - cc.y]));
                  rc[4 * i4 + 2] = __float_as_uint(fmaf(__uint_as_float(rc[4 * i4 + 2]), sc, myTab[rowterm - cc.z]));
                  rc[4 * i4 + 3] = __float_as_uint(fmaf(__uint_as_float(rc[4 * i4 + 3]), sc, myTab[rowterm - cc.w]));
                }
                if (c == 0 && k0 == 0)   // the CLS column: "token to CLS" T - 2, "CLS to CLS" T - 1
                  rc[0] = __float_as_uint(fmaf(s00, sc, myTab[q == 0 ? a.T - 1 : a.T - 2]));
              }
              if (partial) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (c * 16 + i >= valid) rc[i] = __float_as_uint(-INFINITY);
                if constexpr (!HAS_BIAS) tmem_st_32x32b_x16(s_addr + c * 16, rc);   // pass 2 reads the masked logits
              }
#pragma unroll
              for (int i = 0; i < 16; i += 8) {
                mx0 = a3_fmax3(mx0, __uint_as_float(rc[i]), __uint_as_float(rc[i + 1]));
                mx1 = a3_fmax3(mx1, __uint_as_float(rc[i + 2]), __uint_as_float(rc[i + 3]));
                mx2 = a3_fmax3(mx2, __uint_as_float(rc[i + 4]), __uint_as_float(rc[i + 5]));
                mx3 = a3_fmax3(mx3, __uint_as_float(rc[i + 6]), __uint_as_float(rc[i + 7]));
              }
            }
          }
          // with a bias table the finished logits of both chunks stay in r[0] / r[1] for pass 2 (no write-back, no
          // second read); without one pass 2 reads S again, which keeps 16 registers free for the common case
          if (!HAS_BIAS && (valid & 15)) tcgen05_wait_st();
          A3_STAMP();
          if constexpr (!HAS_BIAS) tmem_ld_32x32b_x16(s_addr, r[0]);   // pass 2, chunk 0
          float mt = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          if constexpr (!HAS_BIAS) mt *= sc;   // scale > 0: max commutes with it
          // ---- lazy running max: move it (and rescale O, l) only when it grows by more than 2^kA3LazyLog2
          const bool need = mt > m_run + kA3LazyLog2;
          const float m_new = need ? mt : m_run;
          const float alpha = need ? a3_exp2(m_run - m_new) : 1.0f;   // first tile: exp2(-inf) = 0
          if (j > 0 && __any_sync(0xffffffffu, need)) {
            // O must hold every P V up to tile j-1 (tile j-2 is implied by s_full(j), tile j-1 may still be in flight)
            mbar_wait(&pv_done[g * 2 + (sbuf ^ 1)], ((pvuses >> (sbuf ^ 1)) & 1) ^ 1);   // P V (j-1): the last use of the other buffer
            tcgen05_fence_after();
            uint32_t ob[16];
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              tmem_ld_32x32b_x16(lane_addr + kA3OCol + c * 16, ob);
              tmem_wait_ld16(ob);
#pragma unroll
              for (int i = 0; i < 16; ++i) ob[i] = __float_as_uint(__uint_as_float(ob[i]) * alpha);
              tmem_st_32x32b_x16(lane_addr + kA3OCol + c * 16, ob);
            }
          }
          l_run *= alpha;
          m_run = m_new;
          // ---- pass 2: p = exp2(s - m), row sum, P -> TMEM as packed bf16 over the first 16 columns of the buffer
          const uint64_t sc2 = f2_splat(HAS_BIAS ? 1.0f : sc), negm2 = f2_splat(-m_new);
          uint64_t ps_a = f2_splat(0.f), ps_b = f2_splat(0.f);
#pragma unroll
          for (int c = 0; c < kA3Chunks; ++c) {
            if (c < nch) {
              uint32_t (&rc)[16] = r[c & 1];
              if constexpr (!HAS_BIAS) {
                tmem_wait_ld16(rc);
                if (c + 1 < nch) tmem_ld_32x32b_x16(s_addr + (c + 1) * 16, r[(c + 1) & 1]);
              }
              if (!HAS_BIAS && (c + 1) * 16 > valid) {   // partial chunk: keys past the end of the image
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (c * 16 + i >= valid) rc[i] = __float_as_uint(-INFINITY);
              }
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                uint64_t x = f2_fma(f2_pack(__uint_as_float(rc[i]), __uint_as_float(rc[i + 1])), sc2, negm2);
                float x0, x1, p0, p1;
                f2_unpack(x, x0, x1);
                if (i < 16 - 2 * LDIT_A3_POLY_PAIRS) {
                  p0 = a3_exp2(x0);
                  p1 = a3_exp2(x1);
                } else {
                  // 2^x on the FMA pipe: x = n + f (round to nearest via the 1.5*2^23 trick), 2^f by a cubic
                  // (7.7e-5 relative error, far below the bf16 rounding of P), 2^n by an integer add into the exponent
                  x = f2_pack(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));   // masked keys are -inf
                  const uint64_t t = f2_add(x, magic2);                  // low mantissa bits = n = round(x)
                  const uint64_t f = f2_fma(f2_add(t, nmagic2), mone2, x);   // x - n in [-0.5, 0.5]
                  uint64_t qq = f2_fma(f, e3, e2);
                  qq = f2_fma(qq, f, e1);
                  qq = f2_fma(qq, f, e0);
                  float q0, q1, t0, t1;
                  f2_unpack(qq, q0, q1);
                  f2_unpack(t, t0, t1);
                  p0 = __int_as_float(__float_as_int(t0) * 0x800000 + __float_as_int(q0));        // 2^f * 2^n
                  p1 = __int_as_float(__float_as_int(t1) * 0x800000 + __float_as_int(q1));
                }
#if LDIT_A3_PACK == 1
                pk[i >> 1] = pack_bf16x2(p0, p1);
#else
                pk[i >> 1] = pack_bf16x2_alu(p0, p1);
#endif
                if ((i >> 1) & 1) ps_b = f2_add(ps_b, f2_pack(p0, p1)); else ps_a = f2_add(ps_a, f2_pack(p0, p1));
              }
              tmem_st_32x32b_x8(s_addr + c * 8, pk);
            }
          }
          float s0, s1;
          f2_unpack(f2_add(ps_a, ps_b), s0, s1);
          l_run += s0 + s1;
          tcgen05_wait_st();
        }
        A3_STAMP();
        A3_STAMP();
        pvuses ^= 1u << sbuf;    // P V (j) will be issued on this buffer
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_full[g * 2 + sbuf]);
          if (j == 0 && pending_qb >= 0) {
            // the previous item's O store has long finished reading its staging (= that item's Q tile): hand the
            // buffer back to the producer now, a whole item ahead of when it is needed
            tma_store_wait_read<0>();
            mbar_arrive(&q_empty[pending_qb]);
          }
        }
        if (j == 0) pending_qb = -1;
      }

      // ---- O of the item: normalise and store (merged heads)
      {   // the last P V of the item (and with it every earlier one)
        const int lb = (T - 1) & 1;
        mbar_wait(&pv_done[g * 2 + lb], ((pvuses >> lb) & 1) ^ 1);
      }
      tcgen05_fence_after();
      if (warp_active) {
        // 32 rows x 64 bf16 of this warp -> swizzled smem (this warp's quarter of the item's Q tile: every MMA that
        // read it has completed) -> one TMA store (rows past N are clipped by the 3-D tensor map)
        const float inv = 1.0f / l_run;
        if (a.lse != nullptr && q < a.N)   // exact whatever the lazy running max was: m + log2(sum exp2(s - m))
          a.lse[(static_cast<size_t>(b) * a.heads + h) * a.N + q] = m_run + log2f(l_run);
        uint8_t* stage = sQ + (qb * 2 + g) * 16384 + quarter * 4096;
        uint32_t oa[16], ob[16];
        tmem_ld_32x32b_x16(lane_addr + kA3OCol, oa);
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          tmem_wait_ld16(oa);
          tmem_ld_32x32b_x16(lane_addr + kA3OCol + (c + 1) * 16, ob);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(oa[8 * jj]) * inv, __uint_as_float(oa[8 * jj + 1]) * inv);
            v.y = pack_bf16x2(__uint_as_float(oa[8 * jj + 2]) * inv, __uint_as_float(oa[8 * jj + 3]) * inv);
            v.z = pack_bf16x2(__uint_as_float(oa[8 * jj + 4]) * inv, __uint_as_float(oa[8 * jj + 5]) * inv);
            v.w = pack_bf16x2(__uint_as_float(oa[8 * jj + 6]) * inv, __uint_as_float(oa[8 * jj + 7]) * inv);
            *reinterpret_cast<uint4*>(stage + lane * 128 + (((2 * c + jj) ^ (lane & 7)) << 4)) = v;
          }
          tmem_wait_ld16(ob);
          if (c + 2 < 4) tmem_ld_32x32b_x16(lane_addr + kA3OCol + (c + 2) * 16, oa);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(ob[8 * jj]) * inv, __uint_as_float(ob[8 * jj + 1]) * inv);
            v.y = pack_bf16x2(__uint_as_float(ob[8 * jj + 2]) * inv, __uint_as_float(ob[8 * jj + 3]) * inv);
            v.z = pack_bf16x2(__uint_as_float(ob[8 * jj + 4]) * inv, __uint_as_float(ob[8 * jj + 5]) * inv);
            v.w = pack_bf16x2(__uint_as_float(ob[8 * jj + 6]) * inv, __uint_as_float(ob[8 * jj + 7]) * inv);
            *reinterpret_cast<uint4*>(stage + lane * 128 + (((2 * (c + 1) + jj) ^ (lane & 7)) << 4)) = v;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, stage, h * 64, row0 + quarter * 32, b);
          tma_store_commit();
        }
      }
      // the O columns are read out (tcgen05.wait::ld above): P V (0) of the next item may overwrite them -- it is only
      // issued after this warpgroup's next p_full arrival, which is ordered behind this point
      tcgen05_fence_before();
      __syncwarp();
      pending_qb = static_cast<int>(qb);
    }
    if (lane == 0 && pending_qb >= 0) {
      tma_store_wait_read<0>();
      mbar_arrive(&q_empty[pending_qb]);   // nobody waits for it any more; keeps the arrival count per buffer use uniform
    }
  }

  if (warp < 8 && lane == 0) tma_store_wait<0>();   // global writes of the last ctx stores complete before exit
  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kA3WarpProducer) tmem_dealloc(tmem_base, kA3TmemCols);
}

}  // namespace ldit
