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
// caller zero-fills the padding once).  64 x 64 tiles through padded shared memory, block (32, 8); both global sides
// move 4 bytes per thread (128-byte warp rows) when C, R and ldo are even, element by element otherwise.
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C, int ldo) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool in_pairs = (C & 1) == 0, out_pairs = (ldo & 1) == 0;
  for (int i = ty; i < 64; i += 8) {
    const int r = r0 + i, c = c0 + 2 * tx;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (r < R) {
      if (in_pairs && c + 1 < C) v = *reinterpret_cast<const __nv_bfloat162*>(in + static_cast<size_t>(r) * C + c);
      else {
        if (c < C) v.x = in[static_cast<size_t>(r) * C + c];
        if (c + 1 < C) v.y = in[static_cast<size_t>(r) * C + c + 1];
      }
    }
    *reinterpret_cast<__nv_bfloat162*>(&tile[i][2 * tx]) = v;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i, r = r0 + 2 * tx;
    if (c >= C) continue;
    __nv_bfloat162 v;
    v.x = tile[2 * tx][i];
    v.y = tile[2 * tx + 1][i];
    __nv_bfloat16* dst = out + static_cast<size_t>(c) * ldo + r;
    if (out_pairs && r + 1 < R) *reinterpret_cast<__nv_bfloat162*>(dst) = v;
    else {
      if (r < R) dst[0] = v.x;
      if (r + 1 < R) dst[1] = v.y;
    }
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
                          const float* __restrict__ row_scale, int rows_per_image, float* __restrict__ y, int R, int D) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // 8 elements each
  const int d8 = D / 8;
  if (i >= static_cast<size_t>(R) * d8) return;
  const int c = static_cast<int>(i % d8) * 8;
  // drop-path (HF:61-73): the whole branch of image b is scaled by row_scale[b] (0 or 1 / keep_prob)
  const float rs = row_scale ? __ldg(row_scale + static_cast<int>(i / d8) / rows_per_image) : 1.f;
  float b[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(branch) + i), b);
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), x1 = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
  float l[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
  if (lam != nullptr) {
    const float4 l0 = __ldg(reinterpret_cast<const float4*>(lam + c)), l1 = __ldg(reinterpret_cast<const float4*>(lam + c) + 1);
    l[0] = l0.x; l[1] = l0.y; l[2] = l0.z; l[3] = l0.w; l[4] = l1.x; l[5] = l1.y; l[6] = l1.z; l[7] = l1.w;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) l[e] *= rs;
  reinterpret_cast<float4*>(y)[2 * i] = make_float4(x0.x + l[0] * b[0], x0.y + l[1] * b[1], x0.z + l[2] * b[2], x0.w + l[3] * b[3]);
  reinterpret_cast<float4*>(y)[2 * i + 1] = make_float4(x1.x + l[4] * b[4], x1.y + l[5] * b[5], x1.z + l[6] * b[6], x1.w + l[7] * b[7]);
}
// backward: dbranch bf16 = lam (.) dy;  dlam f32 [D] += column sums of dy (.) branch (skipped when dlam == nullptr).
// Block (min(D/8, 128) column groups, 8 row lanes) owns rows_per_block rows: thread (cg, ry) walks rows ry, ry + 8, ..
// of the slice, the 8 row lanes are summed through shared memory, one atomicAdd per column and block.
__global__ void __launch_bounds__(1024)
scale_residual_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ branch, const float* __restrict__ lam,
                          const float* __restrict__ row_scale, int rows_per_image, __nv_bfloat16* __restrict__ dbranch,
                          float* __restrict__ dlam, int R, int D, int rows_per_block) {
  __shared__ float red[8][128 * 8 + 8];
  const int d8 = D / 8;
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;      // column group
  const bool live = cg < d8;
  const int c = cg * 8;
  float l[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (live && lam != nullptr) {
    const float4 l0 = __ldg(reinterpret_cast<const float4*>(lam + c)), l1 = __ldg(reinterpret_cast<const float4*>(lam + c) + 1);
    l[0] = l0.x; l[1] = l0.y; l[2] = l0.z; l[3] = l0.w; l[4] = l1.x; l[5] = l1.y; l[6] = l1.z; l[7] = l1.w;
  }
  const int r_lo = blockIdx.y * rows_per_block, r_hi = min(R, r_lo + rows_per_block);
  if (live) {
#pragma unroll 2
    for (int r = r_lo + threadIdx.y; r < r_hi; r += 8) {
      const size_t i = static_cast<size_t>(r) * d8 + cg;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(dy) + 2 * i), g1 = __ldg(reinterpret_cast<const float4*>(dy) + 2 * i + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float b[8], o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(branch) + i), b);
      const float rs = row_scale ? __ldg(row_scale + r / rows_per_image) : 1.f;   // drop-path scale of this row's image
#pragma unroll
      for (int e = 0; e < 8; ++e) { o[e] = rs * l[e] * g[e]; acc[e] += rs * g[e] * b[e]; }
      reinterpret_cast<uint4*>(dbranch)[i] = pack8(o);
    }
  }
  if (dlam == nullptr) return;
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.y][threadIdx.x * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.y == 0 && live) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x * 8 + e];
      atomicAdd(dlam + c + e, t);
    }
  }
}

// --------------------------------------------------------------------------- LayerNorm backward
// y = LN(x) gamma + beta (HF:458,460).  One warp per row (row in registers), persistent over rows:
//   dx_out[r] = dx_in[r] + rstd (g - mean(g) - xhat mean(g xhat)),   g = dy gamma,  xhat = (x - mean) rstd
//   dgamma += sum_r dy xhat,  dbeta += sum_r dy        (per-lane partials, one atomicAdd per column and warp at the end)
// dx_in may be nullptr (= 0) and may alias dx_out.  Statistics are recomputed from x (two-pass, as the forward).
template <int VPL>
__global__ void __launch_bounds__(256, (VPL <= 6) ? 2 : 1)   // 152 registers unconstrained at VPL = 6: one block per SM
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
    if (row + nwarps < rows) {
      // the three streams of this warp's NEXT row into L2 while this row's dependent load / reduce phases run
      // (one 128-byte line per lane and pass; no registers are held)
      const size_t nxt = static_cast<size_t>(row + nwarps) * D;
      for (int l = lane * 32; l < D; l += 32 * 32) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(x + nxt + l));
        if (dx_in) asm volatile("prefetch.global.L2 [%0];" ::"l"(dx_in + nxt + l));
      }
      for (int l = lane * 64; l < D; l += 32 * 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(dy + nxt + l));
    }
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
  // per-lane partials -> one sum per block in shared memory (warps take turns: no shared atomics) -> one atomicAdd per
  // column and block.  (One atomicAdd per column and WARP was 7 M atomics on 1536 addresses: 220 us at 12608 x 768.)
  __shared__ float red[2][D];
  const int wib = threadIdx.x >> 5;
  for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) {
    if (wib == w) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float4* rg = reinterpret_cast<float4*>(&red[0][0]) + lane + 32 * i;
        float4* rb = reinterpret_cast<float4*>(&red[1][0]) + lane + 32 * i;
        if (w == 0) { *rg = pg[i]; *rb = pb[i]; }
        else {
          float4 a = *rg, b = *rb;
          a.x += pg[i].x; a.y += pg[i].y; a.z += pg[i].z; a.w += pg[i].w;
          b.x += pb[i].x; b.y += pb[i].y; b.z += pb[i].z; b.w += pb[i].w;
          *rg = a; *rb = b;
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(dgamma + c, red[0][c]);
    atomicAdd(dbeta + c, red[1][c]);
  }
}

// The attention backward lives in attention_bwd_tc.cuh (tcgen05).

// ------------------------------------------------------------------------------- taps backward
// Adjoint of ldit_resample_taps (R:dit_backbone.py:50-61): dout bf16 [B, oh, ow, D] (channels-last, the forward's
// output layout) -> dx f32 [B, N, D] rows 1..P (the CLS row is not read by a tap: the caller zeroes it).  Gather form,
// no atomics: the thread of token cell (gy, gx) walks the output pixels whose bilinear footprint (ATen rule,
// align_corners = False, exactly as resample_taps_kernel) contains the cell and sums weight x gradient.
// blockDim = (D/8, kTapPix), gridDim = (ceil(Gh*Gw / kTapPix), B).
__device__ __forceinline__ float tap_axis_weight(int o, int g, int G, float inv_scale) {
  const float s = fmaxf((o + 0.5f) * inv_scale - 0.5f, 0.f);
  const int i0 = min(static_cast<int>(s), G - 1), i1 = min(i0 + 1, G - 1);
  const float l = s - i0;
  return (i0 == g ? 1.f - l : 0.f) + (i1 == g ? l : 0.f);
}
__global__ void __launch_bounds__(1024)
resample_taps_bwd_kernel(const __nv_bfloat16* __restrict__ dout, float* __restrict__ dx, int N, int D, int Gh, int Gw, int oh, int ow,
                         float inv_scale, float scale) {
  const int cell = blockIdx.x * blockDim.y + threadIdx.y;
  if (cell >= Gh * Gw) return;
  const int b = blockIdx.y;
  const int gy = cell / Gw, gx = cell - gy * Gw;
  // candidate outputs: source coordinate within one cell of (gy, gx); everything towards the border at the border cells
  const int oy_lo = gy == 0 ? 0 : max(0, static_cast<int>(floorf((gy - 0.5f) * scale - 0.5f)) - 1);
  const int oy_hi = gy == Gh - 1 ? oh - 1 : min(oh - 1, static_cast<int>(ceilf((gy + 1.5f) * scale - 0.5f)) + 1);
  const int ox_lo = gx == 0 ? 0 : max(0, static_cast<int>(floorf((gx - 0.5f) * scale - 0.5f)) - 1);
  const int ox_hi = gx == Gw - 1 ? ow - 1 : min(ow - 1, static_cast<int>(ceilf((gx + 1.5f) * scale - 0.5f)) + 1);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const __nv_bfloat16* base = dout + static_cast<size_t>(b) * oh * ow * D + threadIdx.x * 8;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    const float wy = tap_axis_weight(oy, gy, Gh, inv_scale);
    if (wy == 0.f) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const float w = wy * tap_axis_weight(ox, gx, Gw, inv_scale);
      if (w == 0.f) continue;
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(oy * ow + ox) * D)), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, v[i], acc[i]);
    }
  }
  float4* dst = reinterpret_cast<float4*>(dx + (static_cast<size_t>(b) * N + 1 + cell) * D + threadIdx.x * 8);
  dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// out f32 [R] += sum over b of x f32 [B, R]   (position / CLS embedding gradients: every image adds the same rows; R % 4 == 0)
__global__ void __launch_bounds__(256)
batch_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int R) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= R) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + static_cast<size_t>(b) * R + i));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float4* o = reinterpret_cast<float4*>(out + i);
  float4 c = *o;
  c.x += acc.x; c.y += acc.y; c.z += acc.z; c.w += acc.w;
  *o = c;
}

}  // namespace ldit
