// Throughput of a few instruction classes on one SM sub-partition (warp instructions per clock), to decide
// which pipe the softmax packs P on: F2FP (cvt.rn.bf16x2.f32), MUFU.EX2, PRMT + 2 IADD, FFMA2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  float a[8];
  uint32_t r[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; r[i] = threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(a[i]), "f"(__uint_as_float(r[i])));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i]));
      if (OP == 2) { uint32_t x = __float_as_uint(a[i]) + 0x8000u, y = r[i] + 0x8000u; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r[i]) : "r"(x), "r"(y)); }
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(a[i]), "f"(__uint_as_float(r[i])));
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) acc += r[i] + __float_as_uint(a[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const char* names[5] = {"F2FP.BF16 pack", "MUFU.EX2", "2xIADD+PRMT", "FFMA", "F2FP.F16 pack"};
  for (int warps : {4, 8, 16}) {
    for (int op = 0; op < 5; ++op) {
      const int iters = 2000;
      void (*fn)(uint32_t*, long long*, int) = op == 0 ? k<0> : op == 1 ? k<1> : op == 2 ? k<2> : op == 3 ? k<3> : k<4>;
      fn<<<148, warps * 32>>>(out, cyc, iters); fn<<<148, warps * 32>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      double ops = double(iters) * 8 * warps;   // warp-instructions of the class per SM
      printf("%2d warps/SM  %-16s %8.3f warp-instr/clk/SM  (%.1f lanes/clk/SM)\n", warps, names[op], ops / c, 32 * ops / c);
    }
  }
  return 0;
}
