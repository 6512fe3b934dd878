// Flash-style fused attention, register-resident online softmax, optional in-tile
// relative-position bias.  softmax(q k^T / sqrt(64) + bias) v per (image, head)
// (HF:249-306 / HF:310-368), reading Q/K/V straight out of the fused QKV GEMM output
// [B*N, 3D] and writing the merged-heads context [B*N, D] (HF:365-367) -- no head
// transposes are ever materialised.
//
// This is the warp-level (mma.sync m16n8k16) variant: CTA = 4 warps x 16 query rows, 64-key
// K/V tiles double-buffered with cp.async into XOR-swizzled smem.
#pragma once

#include "ptx.cuh"

namespace ldit {

struct AttnArgs {
  const __nv_bfloat16* qkv;  // [B*N, 3*D]
  __nv_bfloat16* ctx;        // [B*N, D]
  const float* bias_table;   // [heads, T] fp32 (already resized to this window) or nullptr
  int B, N, heads, D;
  int Gh, Gw, T;             // T = (2Gh-1)(2Gw-1)+3
  float scale_log2e;         // (1/sqrt(head_dim)) * log2(e)
};

constexpr int kAttBQ = 64, kAttBK = 64, kAttDh = 64;

__device__ __forceinline__ uint32_t swz_off(int row, int chunk) {  // byte offset in a [64][64] bf16 tile
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void att_load_tile(uint8_t* dst, const __nv_bfloat16* src_base, int ld, int row0, int nrows_valid) {
  // 64 rows x 8 chunks of 16 B; 128 threads x 4
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = threadIdx.x + 128 * i;
    const int r = id >> 3, c = id & 7;
    const bool ok = (row0 + r) < nrows_valid;
    const __nv_bfloat16* src = src_base + static_cast<size_t>(ok ? row0 + r : 0) * ld + c * 8;
    cp_async_16(dst + swz_off(r, c), src, ok);
  }
}

template <bool HAS_BIAS>
__global__ void __launch_bounds__(128)
attention_mma_kernel(const AttnArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 8192;
  uint8_t* sV = smem + 8192 + 16384;
  float* sTab = reinterpret_cast<float*>(smem + 8192 + 32768);
  int* sCol = reinterpret_cast<int*>(sTab + (HAS_BIAS ? a.T : 0));

  const int q0 = blockIdx.x * kAttBQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ld = 3 * a.D;
  const __nv_bfloat16* qbase = a.qkv + static_cast<size_t>(b) * a.N * ld + h * kAttDh;
  const __nv_bfloat16* kbase = qbase + a.D;
  const __nv_bfloat16* vbase = qbase + 2 * a.D;

  att_load_tile(sQ, qbase, ld, q0, a.N);
  att_load_tile(sK, kbase, ld, 0, a.N);
  att_load_tile(sV, vbase, ld, 0, a.N);
  cp_async_commit();

  if constexpr (HAS_BIAS) {
    const float* tab = a.bias_table + static_cast<size_t>(h) * a.T;
    for (int i = threadIdx.x; i < a.T; i += 128) sTab[i] = tab[i];
    for (int k = threadIdx.x; k < a.N; k += 128) {
      const int p = k - 1;
      sCol[k] = (k == 0) ? 0 : (p / a.Gw) * (2 * a.Gw - 1) + (p % a.Gw);
    }
  }
  // per-thread row terms of the relative-position index (rows g and g+8 of this warp's slab)
  int rowterm[2] = {0, 0};
  int qrow[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    qrow[r] = q0 + warp * 16 + g + 8 * r;
    if constexpr (HAS_BIAS) {
      const int p = qrow[r] - 1;
      if (qrow[r] >= 1) rowterm[r] = (p / a.Gw + a.Gh - 1) * (2 * a.Gw - 1) + (p % a.Gw) + a.Gw - 1;
    }
  }

  const int ntiles = (a.N + kAttBK - 1) / kAttBK;
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[4][4];

  for (int kt = 0; kt < ntiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ntiles) {
      att_load_tile(sK + (buf ^ 1) * 8192, kbase, ld, (kt + 1) * kAttBK, a.N);
      att_load_tile(sV + (buf ^ 1) * 8192, vbase, ld, (kt + 1) * kAttBK, a.N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int mi = lane >> 3, r = lane & 7;
        ldmatrix_x4(qf[ks], sQ + swz_off(warp * 16 + (mi & 1) * 8 + r, ks * 2 + (mi >> 1)));
      }
    }
    const uint8_t* cK = sK + buf * 8192;
    const uint8_t* cV = sV + buf * 8192;

    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t kf[4];
        const int mi = lane >> 3, r = lane & 7;
        ldmatrix_x4(kf, cK + swz_off((2 * jj + (mi >> 1)) * 8 + r, ks * 2 + (mi & 1)));
        mma_bf16_16816(s[2 * jj], qf[ks], kf[0], kf[1]);
        mma_bf16_16816(s[2 * jj + 1], qf[ks], kf[2], kf[3]);
      }
    }

    // scale, bias, mask; online softmax in registers
    const int kcol0 = kt * kAttBK;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int kc = kcol0 + j * 8 + 2 * t + (e & 1);
        float v = s[j][e] * a.scale_log2e;
        if (kc < a.N) {
          if constexpr (HAS_BIAS) {
            int idx;
            if (qrow[r] == 0) idx = (kc == 0) ? a.T - 1 : a.T - 3;
            else if (kc == 0) idx = a.T - 2;
            else idx = rowterm[r] - sCol[kc];
            if (qrow[r] < a.N) v = fmaf(sTab[idx], 1.4426950408889634f, v);
          }
        } else {
          v = -INFINITY;
        }
        s[j][e] = v;
        mx[r] = fmaxf(mx[r], v);
      }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const float p = exp2f(s[j][e] - m_run[r]);
        s[j][e] = p;
        psum[r] += p;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + psum[r];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
      o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
    }

    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint32_t vf[4];
        const int mi = lane >> 3, r = lane & 7;
        ldmatrix_x4_trans(vf, cV + swz_off(kk * 16 + (mi & 1) * 8 + r, 2 * jj + (mi >> 1)));
        mma_bf16_16816(o[2 * jj], pf, vf[0], vf[1]);
        mma_bf16_16816(o[2 * jj + 1], pf, vf[2], vf[3]);
      }
    }
    __syncthreads();  // everyone is done with buf before it is refilled two tiles later
  }

  // finalise: O / l, merged-heads store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = 1.0f / l;
    if (qrow[r] < a.N) {
      __nv_bfloat16* dst = a.ctx + (static_cast<size_t>(b) * a.N + qrow[r]) * a.D + h * kAttDh + 2 * t;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint32_t*>(dst + j * 8) = pack_bf16x2(o[j][2 * r] * inv, o[j][2 * r + 1] * inv);
      }
    }
  }
}

}  // namespace ldit
