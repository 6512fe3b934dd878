// Does programmatic dependent launch survive stream capture into a CUDA graph on this driver?
// Kernel A spins ~30 us after calling launch_dependents; kernel B stamps the global timer before and
// after griddepcontrol.wait.  If B's first stamp precedes A's end, B was resident while A ran.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void A(unsigned long long* out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned long long t0 = gt();
  while (gt() - t0 < 30000) {}
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t0; out[1] = gt(); }
}
__global__ void B(unsigned long long* out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  unsigned long long t0 = gt();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned long long t1 = gt();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[2] = t0; out[3] = t1; }
}
static void launch(void (*k)(unsigned long long*), unsigned long long* d, cudaStream_t st, bool pdl) {
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(74); cfg.blockDim = dim3(128); cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, d);
}
int main() {
  unsigned long long *d, h[4];
  cudaMalloc(&d, 32); cudaStream_t st; cudaStreamCreate(&st);
  for (int pdl = 0; pdl < 2; ++pdl) {
    launch(A, d, st, pdl); launch(B, d, st, pdl); cudaStreamSynchronize(st);
    launch(A, d, st, pdl); launch(B, d, st, pdl); cudaStreamSynchronize(st);
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("stream pdl=%d: A ran %lld ns; B resident %lld ns before A's end; B released %lld ns after A's end\n", pdl, (long long)(h[1]-h[0]), (long long)(h[1]-h[2]), (long long)(h[3]-h[1]));
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    launch(A, d, st, pdl); launch(B, d, st, pdl);
    cudaStreamEndCapture(st, &g);
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("graph  pdl=%d (%s): A ran %lld ns; B resident %lld ns before A's end; B released %lld ns after A's end\n", pdl, cudaGetErrorString(e), (long long)(h[1]-h[0]), (long long)(h[1]-h[2]), (long long)(h[3]-h[1]));
  }
  return 0;
}
