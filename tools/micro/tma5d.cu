// Where does a 5-D TMA box with a 32-byte inner dimension and SWIZZLE_128B land in shared memory?
// (layout probe for the TMA-fed patch embedding, EPI_PATCH_TMA).  nvcc -gencode arch=compute_100a,code=sm_100a -o tma5d tma5d.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
__global__ void k(const __grid_constant__ CUtensorMap tm, uint16_t* out, int c1, int c2, int c3, int c4) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(&bar)), s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4096));
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(s), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");   // (px, gx, py, gy, bc)
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(b));
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
int main() {
  const int B = 1, H = 64, W = 96, Gh = 4, Gw = 6;
  std::vector<uint16_t> px(B * 3 * H * W);
  // value encodes (c, y, x): c*4096*... keep it simple: 16-bit code = (y << 7) | x  for channel 0; others offset
  for (int c = 0; c < 3; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) px[(c * H + y) * W + x] = (c << 14) | (y << 7) | x;
  uint16_t *d, *o; cudaMalloc(&d, px.size() * 2); cudaMalloc(&o, 16384);
  cudaMemcpy(d, px.data(), px.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[5] = {16, (cuuint64_t)Gw, 16, (cuuint64_t)Gh, (cuuint64_t)B * 3};
  cuuint64_t strides[4] = {32, (cuuint64_t)W * 2, (cuuint64_t)W * 32, (cuuint64_t)H * W * 2};
  cuuint32_t box[5] = {16, 16, 1, 8, 1}, estr[5] = {1, 1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
  k<<<1, 128, 20480>>>(tm, o, 0, 5, 0, 1);   // gx0 = 0, py = 5, gy0 = 0, channel 1
  printf("launch: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  std::vector<uint16_t> h(8192); cudaMemcpy(h.data(), o, 16384, cudaMemcpyDeviceToHost);
  // print, for the first 24 16-byte chunks and some later rows, what (y, x) the chunk starts with
  // rows of 32 B (2 chunks of 16 B); 32-byte swizzle: chunk bit ^= address bit 7, i.e. (row >> 2) & 1
  for (int row : {0, 1, 2, 3, 4, 5, 6, 7, 8, 15, 16, 17, 20, 21, 22, 32, 37, 48, 53, 64}) {
    for (int piece = 0; piece < 2; ++piece) {
      uint16_t v = h[row * 16 + piece * 8];
      printf("row %3d piece %d: c=%d y=%2d x=%2d | patch (gy %d, gx %d) expects y=%d, x=%d for logical piece %d\n", row, piece, v >> 14, (v >> 7) & 127,
             v & 127, row / 16, row % 16, (row / 16) * 16 + 5, (row % 16) * 16 + 8 * (piece ^ ((row >> 2) & 1)), piece ^ ((row >> 2) & 1));
    }
  }
  return 0;
}
