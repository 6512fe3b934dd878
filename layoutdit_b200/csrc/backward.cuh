// Backward of one BeitLayer (HF:469-508) -- first vertical slice of SURVEY.md section 8 row f2.
//
// The eight GEMMs of the backward (four dgrad, four wgrad) run on the forward's tcgen05 kernel (gemm.cuh):
//   dgrad  dA[M, K] = dY[M, N] . W[N, K]        = ldit_gemm_bias(dY, W^T)             (W^T packed once per step, K-major)
//   wgrad  dW[N, K] += dY^T[N, M] . A[M, K]     = ldit_gemm_bias_scale_residual(dY^T, A^T) accumulating into fp32 dW
// with the transposed activations produced by transpose_bf16_kernel (a follow-up replaces those copies by MN-major
// shared-memory descriptors, which tcgen05 takes natively -- the attention kernel already feeds V that way).
// Everything else the backward needs is here: LayerNorm backward, erf-GELU forward / backward as element-wise
// kernels (the training forward keeps the pre-activation), layer-scale + residual forward / backward, column
// sums for the bias gradients, and a correctness-first attention backward (one block per (image, head), thread <-> key,
// softmax statistics recomputed) that stands in until the tcgen05 flash backward exists.  fp32 accumulation throughout.
#pragma once

#include "ptx.cuh"

namespace ldit {

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  return o;
}

// ---------------------------------------------------------------------------------- transpose
// in bf16 [R, C] -> out bf16 [C, ldo] (ldo >= R: the row pitch of a TMA operand must be a multiple of 16 bytes; the
// caller zero-fills the padding once).  32 x 32 tiles through padded shared memory; block (32, 8).
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C, int ldo) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[static_cast<size_t>(r) * C + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) out[static_cast<size_t>(c) * ldo + r] = tile[threadIdx.x][i];
  }
}

// --------------------------------------------------------------------------------- column sums
// out f32 [C] += sum over rows of in bf16 [R, ld] columns [0, C)  (bias gradients).  Block = 32 column-pairs x 8 row
// lanes; a block covers 64 columns and a slice of the rows, one atomicAdd per column and block.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int R, int C, int ld, int rows_per_block) {
  __shared__ float red[8][64];
  const int c = blockIdx.x * 64 + threadIdx.x * 2;
  const int r_lo = blockIdx.y * rows_per_block, r_hi = min(R, r_lo + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    for (int r = r_lo + threadIdx.y; r < r_hi; r += 8) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(in + static_cast<size_t>(r) * ld + c));
      s0 += f.x; s1 += f.y;
    }
  }
  red[threadIdx.y][threadIdx.x * 2] = s0;
  red[threadIdx.y][threadIdx.x * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += red[i][threadIdx.x * 2]; b += red[i][threadIdx.x * 2 + 1]; }
    atomicAdd(out + c, a);
    if (c + 1 < C) atomicAdd(out + c + 1, b);
  }
}

// -------------------------------------------------------------------------------------- GELU
// exact erf GELU (HF:430): forward h = x Phi(x) on a saved pre-activation, backward dx = dh (Phi(x) + x phi(x))
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ h, size_t n8) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float v[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(pre) + i), v);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.5f * v[e] * (1.0f + erff(v[e] * 0.70710678118654752f));
  reinterpret_cast<uint4*>(h)[i] = pack8(v);
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ dpre, size_t n8) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float g[8], x[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(dh) + i), g);
  unpack8(__ldg(reinterpret_cast<const uint4*>(pre) + i), x);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float cdf = 0.5f * (1.0f + erff(x[e] * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x[e] * x[e]);
    g[e] *= cdf + x[e] * pdf;
  }
  reinterpret_cast<uint4*>(dpre)[i] = pack8(g);
}

// ----------------------------------------------------------------- layer scale + residual (HF:488-492, 500-504)
// forward: y f32 [R, D] = x + lam (.) branch        (lam may be nullptr = 1)
__global__ void __launch_bounds__(256)
scale_residual_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ branch, const float* __restrict__ lam,
                          float* __restrict__ y, int R, int D) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // 8 elements each
  const int d8 = D / 8;
  if (i >= static_cast<size_t>(R) * d8) return;
  const int c = static_cast<int>(i % d8) * 8;
  float b[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(branch) + i), b);
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), x1 = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
  float l[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
  if (lam != nullptr) {
    const float4 l0 = __ldg(reinterpret_cast<const float4*>(lam + c)), l1 = __ldg(reinterpret_cast<const float4*>(lam + c) + 1);
    l[0] = l0.x; l[1] = l0.y; l[2] = l0.z; l[3] = l0.w; l[4] = l1.x; l[5] = l1.y; l[6] = l1.z; l[7] = l1.w;
  }
  reinterpret_cast<float4*>(y)[2 * i] = make_float4(x0.x + l[0] * b[0], x0.y + l[1] * b[1], x0.z + l[2] * b[2], x0.w + l[3] * b[3]);
  reinterpret_cast<float4*>(y)[2 * i + 1] = make_float4(x1.x + l[4] * b[4], x1.y + l[5] * b[5], x1.z + l[6] * b[6], x1.w + l[7] * b[7]);
}
// backward: dbranch bf16 = lam (.) dy;  dlam f32 [D] += column sums of dy (.) branch (skipped when dlam == nullptr).
// Block (D/8 column groups up to 128, rows slice): thread owns 8 columns, walks its slice of the rows.
__global__ void __launch_bounds__(256)
scale_residual_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ branch, const float* __restrict__ lam,
                          __nv_bfloat16* __restrict__ dbranch, float* __restrict__ dlam, int R, int D, int rows_per_block) {
  const int d8 = D / 8;
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;      // column group
  if (cg >= d8) return;
  const int c = cg * 8;
  float l[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (lam != nullptr) {
    const float4 l0 = __ldg(reinterpret_cast<const float4*>(lam + c)), l1 = __ldg(reinterpret_cast<const float4*>(lam + c) + 1);
    l[0] = l0.x; l[1] = l0.y; l[2] = l0.z; l[3] = l0.w; l[4] = l1.x; l[5] = l1.y; l[6] = l1.z; l[7] = l1.w;
  }
  const int r_lo = blockIdx.y * rows_per_block, r_hi = min(R, r_lo + rows_per_block);
  for (int r = r_lo; r < r_hi; ++r) {
    const size_t i = static_cast<size_t>(r) * d8 + cg;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(dy) + 2 * i), g1 = __ldg(reinterpret_cast<const float4*>(dy) + 2 * i + 1);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float b[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(branch) + i), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) { o[e] = l[e] * g[e]; acc[e] += g[e] * b[e]; }
    reinterpret_cast<uint4*>(dbranch)[i] = pack8(o);
  }
  if (dlam != nullptr) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dlam + c + e, acc[e]);
  }
}

// --------------------------------------------------------------------------- LayerNorm backward
// y = LN(x) gamma + beta (HF:458,460).  One warp per row (row in registers), persistent over rows:
//   dx_out[r] = dx_in[r] + rstd (g - mean(g) - xhat mean(g xhat)),   g = dy gamma,  xhat = (x - mean) rstd
//   dgamma += sum_r dy xhat,  dbeta += sum_r dy        (per-lane partials, one atomicAdd per column and warp at the end)
// dx_in may be nullptr (= 0) and may alias dx_out.  Statistics are recomputed from x (two-pass, as the forward).
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dy,
                     const float* dx_in, float* dx_out, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, float eps) {
  constexpr int D = 128 * VPL;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  float4 pg[VPL], pb[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) { pg[i] = make_float4(0.f, 0.f, 0.f, 0.f); pb[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
  for (int row = warp; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    const uint2* gr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(row) * D);
    float4 v[VPL], g[VPL];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i] = xr[lane + 32 * i];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / D) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const uint2 u = gr[lane + 32 * i];
      const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      const float4 w = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;           // xhat
      pb[i].x += lo.x; pb[i].y += lo.y; pb[i].z += hi.x; pb[i].w += hi.y;
      pg[i].x += lo.x * v[i].x; pg[i].y += lo.y * v[i].y; pg[i].z += hi.x * v[i].z; pg[i].w += hi.y * v[i].w;
      g[i] = make_float4(lo.x * w.x, lo.y * w.y, hi.x * w.z, hi.y * w.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const float m1 = s1 * (1.0f / D), m2 = s2 * (1.0f / D);
    float4* outr = reinterpret_cast<float4*>(dx_out + static_cast<size_t>(row) * D);
    const float4* inr = dx_in ? reinterpret_cast<const float4*>(dx_in + static_cast<size_t>(row) * D) : nullptr;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 o = make_float4(rstd * (g[i].x - m1 - v[i].x * m2), rstd * (g[i].y - m1 - v[i].y * m2),
                             rstd * (g[i].z - m1 - v[i].z * m2), rstd * (g[i].w - m1 - v[i].w * m2));
      if (inr) { const float4 a = inr[lane + 32 * i]; o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
      outr[lane + 32 * i] = o;
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (lane + 32 * i) * 4;
    atomicAdd(dgamma + c, pg[i].x); atomicAdd(dgamma + c + 1, pg[i].y); atomicAdd(dgamma + c + 2, pg[i].z); atomicAdd(dgamma + c + 3, pg[i].w);
    atomicAdd(dbeta + c, pb[i].x); atomicAdd(dbeta + c + 1, pb[i].y); atomicAdd(dbeta + c + 2, pb[i].z); atomicAdd(dbeta + c + 3, pb[i].w);
  }
}

// The attention backward lives in attention_bwd_tc.cuh (tcgen05).

}  // namespace ldit
