// HBM-bound row-wise kernels: LayerNorm, im2col (+cast), CLS rows, tap resampling.
// None of these has data reuse worth staging in shared memory; they are written for
// coalesced 128-bit accesses, one pass over their bytes, and enough bytes in flight.
#pragma once

#include "ptx.cuh"

namespace ldit {

// ------------------------------------------------------------------------------ LayerNorm
// nn.LayerNorm(D, eps=1e-12) at HF:458,460 (called HF:478,495).  One warp per row, the row
// lives in registers: mean, then the centred second moment (eps is ~0, so the one-pass
// E[x^2]-E[x]^2 form is not safe), all in fp32.  fp32 residual stream in, bf16 out
// (the next kernel is a bf16 GEMM).
// Persistent: the 85 rows an SM owns at base224 do not fit its register file at once (85 x 3 KB), so a
// one-row-per-warp grid runs in 1.3 waves of fully exposed load latency.  Here two blocks of 8-16 warps per SM
// each walk their rows with the NEXT row's loads already in flight while the current one is reduced,
// normalised and stored; rows are dealt so that every block gets the same number (+-1).
// Warps per block by row width, so that two rows of registers per lane fit without spills at two blocks per SM.
__host__ __device__ constexpr int ln_warps(int vpl) { return vpl <= 4 ? 16 : (vpl <= 6 ? 10 : (vpl <= 8 ? 8 : 4)); }
template <int VPL>  // float4 per lane: D = 128 * VPL
__global__ void __launch_bounds__(ln_warps(VPL) * 32, 2)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y, int rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = 128 * VPL;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kLnWarps = ln_warps(VPL);
  const int stride = gridDim.x * kLnWarps;
  int row = blockIdx.x + gridDim.x * warp;
  if (row >= rows) return;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4 v[VPL], nx[VPL];
  {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = xr[lane + 32 * i];
  }
  for (; row < rows; row += stride) {
    const int nrow = row + stride;
    if (nrow < rows) {
      const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(nrow) * D);
#pragma unroll
      for (int i = 0; i < VPL; ++i) nx[i] = xr[lane + 32 * i];
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
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
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 g = __ldg(g4 + lane + 32 * i);
      const float4 b = __ldg(b4 + lane + 32 * i);
      uint2 o;
      o.x = pack_bf16x2(fmaf(v[i].x * rstd, g.x, b.x), fmaf(v[i].y * rstd, g.y, b.y));
      o.y = pack_bf16x2(fmaf(v[i].z * rstd, g.z, b.z), fmaf(v[i].w * rstd, g.w, b.w));
      yr[lane + 32 * i] = o;
    }
    if (nrow < rows) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) v[i] = nx[i];
    }
  }
}

// Residual add fused into the LayerNorm that follows it (HF:488-495, 500-504 then the next layer's HF:478):
//     x <- x + branch;   y = LayerNorm(x)
// `branch` is the bf16, already layer-scaled output of the out-projection / fc2 GEMM (EPI_BIAS_SCALE).  The GEMM then
// ends in a plain bf16 tile store instead of an fp32 reduce-add through the L2 atomic units, and the residual stream
// is updated by the kernel that has to stream it anyway.  y may alias branch (same thread reads then writes the same
// elements).  Same persistent structure as layernorm_kernel.
__host__ __device__ constexpr int aln_warps(int vpl) { return vpl <= 4 ? 12 : (vpl <= 6 ? 8 : (vpl <= 8 ? 6 : 3)); }
template <int VPL>
__global__ void __launch_bounds__(aln_warps(VPL) * 32, 2)
add_layernorm_kernel(float* __restrict__ x, const __nv_bfloat16* branch, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __nv_bfloat16* y, int rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = 128 * VPL;
  constexpr int kLnWarps = aln_warps(VPL);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * kLnWarps;
  int row = blockIdx.x + gridDim.x * warp;
  if (row >= rows) return;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  float4 v[VPL], nx[VPL];
  uint2 br[VPL], nb[VPL];
  {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    const uint2* rr = reinterpret_cast<const uint2*>(branch + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) { v[i] = xr[lane + 32 * i]; br[i] = rr[lane + 32 * i]; }
  }
  for (; row < rows; row += stride) {
    const int nrow = row + stride;
    if (nrow < rows) {
      const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(nrow) * D);
      const uint2* rr = reinterpret_cast<const uint2*>(branch + static_cast<size_t>(nrow) * D);
#pragma unroll
      for (int i = 0; i < VPL; ++i) { nx[i] = xr[lane + 32 * i]; nb[i] = rr[lane + 32 * i]; }
    }
    float4* xw = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * D);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&br[i].x));
      const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&br[i].y));
      v[i].x += lo.x; v[i].y += lo.y; v[i].z += hi.x; v[i].w += hi.y;
      xw[lane + 32 * i] = v[i];
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
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 g = __ldg(g4 + lane + 32 * i);
      const float4 b = __ldg(b4 + lane + 32 * i);
      uint2 o;
      o.x = pack_bf16x2(fmaf(v[i].x * rstd, g.x, b.x), fmaf(v[i].y * rstd, g.y, b.y));
      o.y = pack_bf16x2(fmaf(v[i].z * rstd, g.z, b.z), fmaf(v[i].w * rstd, g.w, b.w));
      yr[lane + 32 * i] = o;
    }
    if (nrow < rows) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) { v[i] = nx[i]; br[i] = nb[i]; }
    }
  }
}

// ------------------------------------------------------------------------- im2col (+cast)
// Conv2d(3, D, k=16, s=16) at HF:209,218 is a GEMM over 16x16 patches.  This pass rewrites
// the NCHW page batch as the GEMM's A operand [B*P, 768] bf16 with column order (c, py, px)
// -- the order of the flattened conv weight -- casting from the caller's dtype on the way
// (the reference feeds fp32 on CPU and fp16 under autocast, R:trainer.py:155,168).
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

template <typename T>
__global__ void __launch_bounds__(256)
im2col_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ a, int B, int H, int W, int Gh, int Gw) {
  pdl_launch_dependents();
  pdl_wait();
  // one thread = 8 consecutive pixels of one image row (half a patch row)
  const int wchunks = (Gw * 16) / 8;
  const size_t total = static_cast<size_t>(B) * 3 * (Gh * 16) * wchunks;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xc = static_cast<int>(idx % wchunks);
  size_t t = idx / wchunks;
  const int yy = static_cast<int>(t % (Gh * 16)); t /= (Gh * 16);
  const int c = static_cast<int>(t % 3);
  const int b = static_cast<int>(t / 3);
  float v[8];
  load8<T>(x + ((static_cast<size_t>(b) * 3 + c) * H + yy) * W + xc * 8, v);
  const int gy = yy >> 4, py = yy & 15, gx = xc >> 1, px = (xc & 1) * 8;
  const size_t row = (static_cast<size_t>(b) * Gh + gy) * Gw + gx;
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(a + row * 768 + c * 256 + py * 16 + px) = o;
}

// Input transform fused into the patch gather.  The detector hands the backbone
// GeneralizedRCNNTransform(pages) (R:src/layoutdit/modeling/model.py:44-56: fixed_size (224, 224), mean/std 0.5):
// per page normalize (TV:models/detection/transform.py normalize(): (image - mean) / std), then
// F.interpolate(size=fixed_size, mode="bilinear", align_corners=False) (_resize_image_and_masks), then batching.
// This kernel reads the raw pages -- a device array of B pointers to [3, Hs_b, Ws_b] images of any size --
// and writes the normalized, resized pixels straight into the patch-embed GEMM's A operand ([B*P, 768] bf16,
// (c, py, px) columns): the resized fp32 batch is never materialised.  ATen's rule: ratio = in / out,
// src = max(ratio * (dst + .5) - .5, 0), i0 = min(floor(src), in - 1), i1 = i0 + (i0 < in - 1), l = src - i0;
// value = l0y (l0x p00 + l1x p01) + l1y (l0x p10 + l1x p11) on normalized pixels.
// One thread = 8 consecutive output pixels of one output row and channel (one 16-byte store).
template <typename T> __device__ __forceinline__ float load1(const T* p);
template <> __device__ __forceinline__ float load1<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load1<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }

template <typename T>
__global__ void __launch_bounds__(256)
pages_im2col_kernel(const T* const* __restrict__ pages, const int* __restrict__ page_hw, __nv_bfloat16* __restrict__ a, int B, int H,
                    int W, int Gh, int Gw, float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_launch_dependents();
  pdl_wait();
  const int wchunks = (Gw * 16) / 8;
  const size_t total = static_cast<size_t>(B) * 3 * (Gh * 16) * wchunks;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xc = static_cast<int>(idx % wchunks);
  size_t t = idx / wchunks;
  const int yy = static_cast<int>(t % (Gh * 16)); t /= (Gh * 16);
  const int c = static_cast<int>(t % 3);
  const int b = static_cast<int>(t / 3);
  const int Hs = page_hw[2 * b], Ws = page_hw[2 * b + 1];
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
  const T* src = pages[b] + static_cast<size_t>(c) * Hs * Ws;
  const float ry = static_cast<float>(Hs) / H, rx = static_cast<float>(Ws) / W;
  const float sy = fmaxf(ry * (yy + 0.5f) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), Hs - 1);
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0);
  const float l1y = sy - y0, l0y = 1.f - l1y;
  const T* r0 = src + static_cast<size_t>(y0) * Ws;
  const T* r1 = src + static_cast<size_t>(y1) * Ws;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float sx = fmaxf(rx * (xc * 8 + i + 0.5f) - 0.5f, 0.f);
    const int x0 = min(static_cast<int>(sx), Ws - 1);
    const int x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
    const float l1x = sx - x0, l0x = 1.f - l1x;
    const float p00 = (load1<T>(r0 + x0) - mean) / sd, p01 = (load1<T>(r0 + x1) - mean) / sd;
    const float p10 = (load1<T>(r1 + x0) - mean) / sd, p11 = (load1<T>(r1 + x1) - mean) / sd;
    v[i] = l0y * (l0x * p00 + l1x * p01) + l1y * (l0x * p10 + l1x * p11);
  }
  const int gy = yy >> 4, py = yy & 15, gx = xc >> 1, px = (xc & 1) * 8;
  const size_t row = (static_cast<size_t>(b) * Gh + gy) * Gw + gx;
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(a + row * 768 + c * 256 + py * 16 + px) = o;
}

// Same transform with the source rows staged in shared memory: a block owns one output row of one page (all
// three channels); the two source rows each channel blends are copied in with coalesced (16-byte when the page
// allows it) loads and the strided bilinear gather then runs out of shared memory.  The direct kernel above
// issues 32 scattered 4-byte loads per thread (one sector each at a 1024 -> 224 stride); this one reads each
// needed source row once, contiguously.  Arithmetic identical to the direct kernel.  max_w = row pitch of the
// staging buffer (>= every page's width); dynamic smem = 6 * max_w * sizeof(T).
template <typename T>
__global__ void __launch_bounds__(128)
pages_rows_im2col_kernel(const T* const* __restrict__ pages, const int* __restrict__ page_hw, __nv_bfloat16* __restrict__ a, int H,
                         int W, int Gh, int Gw, float m0, float m1, float m2, float s0, float s1, float s2, int max_w) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t rows_raw[];
  T* rows = reinterpret_cast<T*>(rows_raw);   // [c][r][max_w]
  const int b = blockIdx.y, yy = blockIdx.x;
  const int Hs = page_hw[2 * b], Ws = page_hw[2 * b + 1];
  const float ry = static_cast<float>(Hs) / H, rx = static_cast<float>(Ws) / W;
  const float sy = fmaxf(ry * (yy + 0.5f) - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), Hs - 1);
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0);
  const float l1y = sy - y0, l0y = 1.f - l1y;
  const T* page = pages[b];
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (reinterpret_cast<uintptr_t>(page) % 16 == 0) && (Ws % VEC == 0);
#pragma unroll
  for (int cr = 0; cr < 6; ++cr) {
    const T* src = page + (static_cast<size_t>(cr >> 1) * Hs + ((cr & 1) ? y1 : y0)) * Ws;
    T* dst = rows + cr * max_w;
    if (vec_ok) {
      for (int i = threadIdx.x; i < Ws / VEC; i += blockDim.x)
        reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    } else {
      for (int i = threadIdx.x; i < Ws; i += blockDim.x) dst[i] = src[i];
    }
  }
  __syncthreads();
  const int wchunks = (Gw * 16) / 8;
  for (int t = threadIdx.x; t < 3 * wchunks; t += blockDim.x) {
    const int c = t / wchunks, xc = t - c * wchunks;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const T* r0 = rows + (2 * c) * max_w;
    const T* r1 = r0 + max_w;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float sx = fmaxf(rx * (xc * 8 + i + 0.5f) - 0.5f, 0.f);
      const int x0 = min(static_cast<int>(sx), Ws - 1);
      const int x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
      const float l1x = sx - x0, l0x = 1.f - l1x;
      const float p00 = (static_cast<float>(r0[x0]) - mean) / sd, p01 = (static_cast<float>(r0[x1]) - mean) / sd;
      const float p10 = (static_cast<float>(r1[x0]) - mean) / sd, p11 = (static_cast<float>(r1[x1]) - mean) / sd;
      v[i] = l0y * (l0x * p00 + l1x * p01) + l1y * (l0x * p10 + l1x * p11);
    }
    const int gy = yy >> 4, py = yy & 15, gx = xc >> 1, px = (xc & 1) * 8;
    const size_t row = (static_cast<size_t>(b) * Gh + gy) * Gw + gx;
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(a + row * 768 + c * 256 + py * 16 + px) = o;
  }
}

// --------------------------------------------------------------- weight-preparation resizes
// Once per (weights, H, W), never per forward: the position table at a non-native patch grid (HF
// interpolate_pos_encoding, HF:138-159: F.interpolate(mode="bicubic", align_corners=False, size=...)) and the
// relative-position bias table at a non-native window (HF:556-571: F.interpolate(mode="bilinear", size=...)).
// src f32 [h*w, C] (C contiguous) -> dst f32 [oh*ow, C] (+ add[C] if given: the conv bias that rides on the
// position rows).  ATen's rules for `size=`: ratio = in/out; bilinear: src = max(ratio (dst+.5) - .5, 0), two taps;
// bicubic: src = ratio (dst+.5) - .5 (not clamped), four taps at floor(src)-1..+2 with clamped indices and the
// cubic-convolution coefficients for A = -0.75.  Same size in and out is an exact copy (+ add).
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  constexpr float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

template <bool CUBIC>
__global__ void __launch_bounds__(256)
resize_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ add, int h, int w, int oh,
                   int ow, int C) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(oh) * ow * C) return;
  const int c = static_cast<int>(idx % C);
  const int pix = static_cast<int>(idx / C);
  const int oy = pix / ow, ox = pix - oy * ow;
  const float ry = static_cast<float>(h) / oh, rx = static_cast<float>(w) / ow;
  float acc = 0.f;
  if (h == oh && w == ow) {
    acc = src[idx];
  } else if constexpr (CUBIC) {
    const float sy = ry * (oy + 0.5f) - 0.5f, sx = rx * (ox + 0.5f) - 0.5f;
    const int iy = static_cast<int>(floorf(sy)), ix = static_cast<int>(floorf(sx));
    float wy[4], wx[4];
    cubic_coeffs(sy - iy, wy);
    cubic_coeffs(sx - ix, wx);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), h - 1);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), w - 1);
        row += wx[b] * src[(static_cast<size_t>(yy) * w + xx) * C + c];
      }
      acc += wy[a] * row;
    }
  } else {
    const float sy = fmaxf(ry * (oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(rx * (ox + 0.5f) - 0.5f, 0.f);
    const int y0 = min(static_cast<int>(sy), h - 1), x0 = min(static_cast<int>(sx), w - 1);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - y0, lx = sx - x0;
    const float* p = src + c;
    acc = (1.f - ly) * ((1.f - lx) * p[(static_cast<size_t>(y0) * w + x0) * C] + lx * p[(static_cast<size_t>(y0) * w + x1) * C]) +
          ly * ((1.f - lx) * p[(static_cast<size_t>(y1) * w + x0) * C] + lx * p[(static_cast<size_t>(y1) * w + x1) * C]);
  }
  dst[idx] = acc + (add != nullptr ? add[c] : 0.f);
}

// Token row 0 of every image: cls_token + position row 0 (HF:176-180); cls_pos = their sum.
__global__ void __launch_bounds__(256)
cls_rows_kernel(const float* __restrict__ cls_pos, float* __restrict__ xres, int B, int N, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  const int d4 = D / 4;
  if (idx >= B * d4) return;
  const int b = idx / d4, j = idx - b * d4;
  reinterpret_cast<float4*>(xres + static_cast<size_t>(b) * N * D)[j] = __ldg(reinterpret_cast<const float4*>(cls_pos) + j);
}

// ------------------------------------------------------------------------ tap resampling
// R:dit_backbone.py:50-61: drop CLS, view tokens as a Gh x Gw grid and resample with
// F.interpolate(scale_factor=s, mode="bilinear", align_corners=False), s in {4, 2, 1, .5}.
// Output is written channels-last ([B, oh, ow, D] in memory, handed to the caller as a
// [B, D, oh, ow] view) -- the same strides the reference's interpolate outputs have on the
// NHWC-strided view it is given, and the layout in which both the token reads and the
// (dominant) tap writes are fully coalesced.  ATen's source-index rule for a given
// scale_factor: src = max((dst + .5) / s - .5, 0); i0 = min(floor(src), in-1); i1 = min(i0+1, in-1).
// blockDim = (D/8, kTapPix): threadIdx.x = 8-channel chunk, threadIdx.y = output pixel of the block;
// gridDim = (ceil(oh*ow / kTapPix), B).  One 32-bit division per thread; every global access is 16 B
// per lane with consecutive lanes on consecutive addresses (a pixel's D channels are contiguous in
// both the token rows and the channels-last output).
constexpr int kTapPix = 4;

__global__ void __launch_bounds__(1024)
resample_taps_kernel(const float* __restrict__ xres, __nv_bfloat16* __restrict__ out, int N, int D, int Gh, int Gw,
                     int oh, int ow, float inv_scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int pix = blockIdx.x * kTapPix + threadIdx.y;
  if (pix >= oh * ow) return;
  const int b = blockIdx.y;
  const int oy = pix / ow, ox = pix - oy * ow;
  const float sy = fmaxf((oy + 0.5f) * inv_scale - 0.5f, 0.f);
  const float sx = fmaxf((ox + 0.5f) * inv_scale - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), Gh - 1), x0 = min(static_cast<int>(sx), Gw - 1);
  const int y1 = min(y0 + 1, Gh - 1), x1 = min(x0 + 1, Gw - 1);
  const float ly = sy - y0, lx = sx - x0;
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const float* base = xres + (static_cast<size_t>(b) * N + 1) * D + threadIdx.x * 8;
  const float4* p00 = reinterpret_cast<const float4*>(base + static_cast<size_t>(y0 * Gw + x0) * D);
  const float4* p01 = reinterpret_cast<const float4*>(base + static_cast<size_t>(y0 * Gw + x1) * D);
  const float4* p10 = reinterpret_cast<const float4*>(base + static_cast<size_t>(y1 * Gw + x0) * D);
  const float4* p11 = reinterpret_cast<const float4*>(base + static_cast<size_t>(y1 * Gw + x1) * D);
  float acc[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 a = __ldg(p00 + h), bq = __ldg(p01 + h), c = __ldg(p10 + h), d = __ldg(p11 + h);
    // same association order as ATen: w00*a + w01*b + w10*c + w11*d
    acc[4 * h + 0] = w00 * a.x + w01 * bq.x + w10 * c.x + w11 * d.x;
    acc[4 * h + 1] = w00 * a.y + w01 * bq.y + w10 * c.y + w11 * d.y;
    acc[4 * h + 2] = w00 * a.z + w01 * bq.z + w10 * c.z + w11 * d.z;
    acc[4 * h + 3] = w00 * a.w + w01 * bq.w + w10 * c.w + w11 * d.w;
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  __nv_bfloat16* dst = out + ((static_cast<size_t>(b) * oh * ow + pix) * D) + threadIdx.x * 8;
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
}

// Integer up-sampling (S = 2 or 4).  Every output pixel of source cell (y, x) only needs the 3x3
// token neighbourhood of that cell, so one thread loads it once for its 8 channels (from L1/L2,
// each token row is touched ~9x instead of 4*S*S times) and emits all S*S outputs of the cell.
// Weights follow ATen exactly: dst row S*y+a has src = y + fa, fa = (a + .5)/S - .5; fa < 0 blends
// rows (y-1, y) with l = 1 + fa, fa > 0 blends (y, y+1) with l = fa; at the top/left border ATen
// clamps src to 0 (weight 1 on row 0), at the bottom/right it blends row Gh-1 with itself.
// blockDim = (D/8, kUpCells), gridDim = (ceil(Gh*Gw / kUpCells), B).
constexpr int kUpCells = 2;

template <int S>
__global__ void __launch_bounds__(512)
upsample_taps_kernel(const float* __restrict__ xres, __nv_bfloat16* __restrict__ out, int N, int D, int Gh, int Gw) {
  pdl_launch_dependents();
  pdl_wait();
  const int cell = blockIdx.x * kUpCells + threadIdx.y;
  if (cell >= Gh * Gw) return;
  const int b = blockIdx.y;
  const int y = cell / Gw, x = cell - y * Gw;
  const int yy[3] = {max(y - 1, 0), y, min(y + 1, Gh - 1)};
  const int xx[3] = {max(x - 1, 0), x, min(x + 1, Gw - 1)};
  const float* base = xres + (static_cast<size_t>(b) * N + 1) * D + threadIdx.x * 8;
  float w[3][3][8];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4* p = reinterpret_cast<const float4*>(base + static_cast<size_t>(yy[r] * Gw + xx[c]) * D);
      const float4 lo = __ldg(p), hi = __ldg(p + 1);
      w[r][c][0] = lo.x; w[r][c][1] = lo.y; w[r][c][2] = lo.z; w[r][c][3] = lo.w;
      w[r][c][4] = hi.x; w[r][c][5] = hi.y; w[r][c][6] = hi.z; w[r][c][7] = hi.w;
    }
  const int ow = Gw * S;
  __nv_bfloat16* obase = out + ((static_cast<size_t>(b) * Gh * S + static_cast<size_t>(y) * S) * ow + static_cast<size_t>(x) * S) * D +
                         threadIdx.x * 8;
#pragma unroll
  for (int a = 0; a < S; ++a) {
    constexpr float inv = 1.0f / S;
    const float fa = (a + 0.5f) * inv - 0.5f;
    const int r0 = fa < 0.f ? 0 : 1;
    const float ly = fa < 0.f ? (y == 0 ? 1.0f : 1.0f + fa) : fa;
#pragma unroll
    for (int bb = 0; bb < S; ++bb) {
      const float fb = (bb + 0.5f) * inv - 0.5f;
      const int c0 = fb < 0.f ? 0 : 1;
      const float lx = fb < 0.f ? (x == 0 ? 1.0f : 1.0f + fb) : fb;
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        acc[e] = w00 * w[r0][c0][e] + w01 * w[r0][c0 + 1][e] + w10 * w[r0 + 1][c0][e] + w11 * w[r0 + 1][c0 + 1][e];
      uint4 o;
      o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
      __nv_bfloat16* dst = obase + (static_cast<size_t>(a) * ow + bb) * D;
      asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
  }
}

// ------------------------------------------------------------------ FPN top-down pathway
// torchvision FeaturePyramidNetwork.forward (TV:ops/feature_pyramid_network.py:172-196) on a DiT pyramid:
//     inner_i = inner_blocks[i](resample_s(h_i)) + F.interpolate(inner_{i+1}, size=..., mode="nearest")
// The 1x1 lateral convolution and the bilinear resampling of R:dit_backbone.py:57-59 are both linear and the
// bilinear weights sum to one, so they commute: the lateral runs FIRST, as a GEMM on the Gh x Gw token grid
// (256 instead of D channels, 1/16 .. 4x fewer pixels), and this kernel resamples its output `lat`
// [B, Gh, Gw, C] bf16 to [B, oh, ow, C] (ATen's bilinear rule, align_corners=False, as in
// resample_taps_kernel) and adds the nearest-neighbour up-sampled coarser level `top` [B, th, tw, C]
// (ATen nearest: src = min(floor(dst * in / out), in - 1)), fp32 math, bf16 out.  The D-channel taps of the
// reference are never written.  blockDim = (C/8, kTapPix), gridDim = (ceil(oh*ow / kTapPix), B).
__global__ void __launch_bounds__(1024)
fpn_merge_kernel(const __nv_bfloat16* __restrict__ lat, const __nv_bfloat16* __restrict__ top, __nv_bfloat16* __restrict__ out,
                 int C, int Gh, int Gw, int oh, int ow, float inv_scale, int th, int tw) {
  pdl_launch_dependents();
  pdl_wait();
  const int pix = blockIdx.x * blockDim.y + threadIdx.y;
  if (pix >= oh * ow) return;
  const int b = blockIdx.y;
  const int oy = pix / ow, ox = pix - oy * ow;
  const float sy = fmaxf((oy + 0.5f) * inv_scale - 0.5f, 0.f);
  const float sx = fmaxf((ox + 0.5f) * inv_scale - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(sy), Gh - 1), x0 = min(static_cast<int>(sx), Gw - 1);
  const int y1 = min(y0 + 1, Gh - 1), x1 = min(x0 + 1, Gw - 1);
  const float ly = sy - y0, lx = sx - x0;
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const __nv_bfloat16* base = lat + static_cast<size_t>(b) * Gh * Gw * C + threadIdx.x * 8;
  float v00[8], v01[8], v10[8], v11[8], acc[8];
  load8<__nv_bfloat16>(base + static_cast<size_t>(y0 * Gw + x0) * C, v00);
  load8<__nv_bfloat16>(base + static_cast<size_t>(y0 * Gw + x1) * C, v01);
  load8<__nv_bfloat16>(base + static_cast<size_t>(y1 * Gw + x0) * C, v10);
  load8<__nv_bfloat16>(base + static_cast<size_t>(y1 * Gw + x1) * C, v11);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = w00 * v00[e] + w01 * v01[e] + w10 * v10[e] + w11 * v11[e];
  if (top != nullptr) {
    const int ty = min(static_cast<int>(floorf(oy * (static_cast<float>(th) / oh))), th - 1);
    const int tx = min(static_cast<int>(floorf(ox * (static_cast<float>(tw) / ow))), tw - 1);
    float t[8];
    load8<__nv_bfloat16>(top + ((static_cast<size_t>(b) * th + ty) * tw + tx) * C + threadIdx.x * 8, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += t[e];
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * oh * ow + pix) * C + threadIdx.x * 8) = o;
}

// Integer up-sampling merge (S = 2 or 4) when the coarser level is exactly half the output size (always the
// case on a DiT pyramid: levels are 4Gh, 2Gh, Gh): the cell scheme of upsample_taps_kernel -- one thread loads
// the 3x3 lateral neighbourhood of source cell (y, x) and the (S/2)^2 coarser-level pixels above it once and
// emits the S*S outputs of the cell, 13 (S=4) or 10 (S=2) 16-byte loads per S*S stores instead of 5 per store.
// blockDim = (C/8, cells), gridDim = (ceil(Gh*Gw / cells), B).
template <int S>
__global__ void __launch_bounds__(256)
fpn_merge_up_kernel(const __nv_bfloat16* __restrict__ lat, const __nv_bfloat16* __restrict__ top, __nv_bfloat16* __restrict__ out,
                    int C, int Gh, int Gw) {
  pdl_launch_dependents();
  pdl_wait();
  const int cell = blockIdx.x * blockDim.y + threadIdx.y;
  if (cell >= Gh * Gw) return;
  const int b = blockIdx.y;
  const int y = cell / Gw, x = cell - y * Gw;
  const int yy[3] = {max(y - 1, 0), y, min(y + 1, Gh - 1)};
  const int xx[3] = {max(x - 1, 0), x, min(x + 1, Gw - 1)};
  const __nv_bfloat16* base = lat + static_cast<size_t>(b) * Gh * Gw * C + threadIdx.x * 8;
  float w[3][3][8];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) load8<__nv_bfloat16>(base + static_cast<size_t>(yy[r] * Gw + xx[c]) * C, w[r][c]);
  constexpr int T = S / 2;                       // coarser-level pixels per cell and axis
  const int th = Gh * T, tw = Gw * T;
  float t[T][T][8];
#pragma unroll
  for (int r = 0; r < T; ++r)
#pragma unroll
    for (int c = 0; c < T; ++c)
      load8<__nv_bfloat16>(top + ((static_cast<size_t>(b) * th + y * T + r) * tw + x * T + c) * C + threadIdx.x * 8, t[r][c]);
  const int ow = Gw * S;
  __nv_bfloat16* obase = out + ((static_cast<size_t>(b) * Gh * S + static_cast<size_t>(y) * S) * ow + static_cast<size_t>(x) * S) * C +
                         threadIdx.x * 8;
#pragma unroll
  for (int a = 0; a < S; ++a) {
    constexpr float inv = 1.0f / S;
    const float fa = (a + 0.5f) * inv - 0.5f;
    const int r0 = fa < 0.f ? 0 : 1;
    const float ly = fa < 0.f ? (y == 0 ? 1.0f : 1.0f + fa) : fa;
#pragma unroll
    for (int bb = 0; bb < S; ++bb) {
      const float fb = (bb + 0.5f) * inv - 0.5f;
      const int c0 = fb < 0.f ? 0 : 1;
      const float lx = fb < 0.f ? (x == 0 ? 1.0f : 1.0f + fb) : fb;
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        acc[e] = (w00 * w[r0][c0][e] + w01 * w[r0][c0 + 1][e] + w10 * w[r0 + 1][c0][e] + w11 * w[r0 + 1][c0 + 1][e]) + t[a / 2][bb / 2][e];
      uint4 o;
      o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(obase + (static_cast<size_t>(a) * ow + bb) * C) = o;
    }
  }
}

// Tokens without CLS, fp32 -> bf16: the scale-1 tap (R:dit_backbone.py:50-56, no interpolate at :57) and the
// A operand of the FPN lateral GEMM.  One 16-byte store per thread, rows of D contiguous on both sides.
__global__ void __launch_bounds__(256)
cast_tokens_kernel(const float* __restrict__ xres, __nv_bfloat16* __restrict__ out, int B, int P, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int d8 = D / 8;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(B) * P * d8) return;
  const int c = static_cast<int>(idx % d8);
  const size_t row = idx / d8;            // b * P + p
  const size_t b = row / P;
  const float4* src = reinterpret_cast<const float4*>(xres + (row + b + 1) * D) + 2 * c;   // token row b*(P+1) + 1 + p
  const float4 lo = __ldg(src), hi = __ldg(src + 1);
  uint4 o;
  o.x = pack_bf16x2(lo.x, lo.y); o.y = pack_bf16x2(lo.z, lo.w);
  o.z = pack_bf16x2(hi.x, hi.y); o.w = pack_bf16x2(hi.z, hi.w);
  reinterpret_cast<uint4*>(out)[idx] = o;
}

// LastLevelMaxPool (TV:ops/feature_pyramid_network.py:231-249): max_pool2d(kernel 1, stride 2) == every second
// pixel of every second row.  in [B, H, W, C] -> out [B, ceil(H/2), ceil(W/2), C], any element type: a pixel is
// c16 16-byte pieces (C * sizeof(element) / 16), one piece per thread.
__global__ void __launch_bounds__(256)
subsample2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int c16) {
  pdl_launch_dependents();
  pdl_wait();
  const int oh = (H + 1) / 2, ow = (W + 1) / 2;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(B) * oh * ow * c16) return;
  const int c = static_cast<int>(idx % c16);
  size_t t = idx / c16;
  const int x = static_cast<int>(t % ow); t /= ow;
  const int y = static_cast<int>(t % oh);
  const int b = static_cast<int>(t / oh);
  out[idx] = __ldg(in + ((static_cast<size_t>(b) * H + 2 * y) * W + 2 * x) * c16 + c);
}

}  // namespace ldit
