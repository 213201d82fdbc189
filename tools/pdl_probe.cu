// Probe: cost of a kernel->kernel dependency inside a CUDA graph, with and without programmatic dependent launch.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pdl_probe tools/pdl_probe.cu && ./pdl_probe
// A chain of N dependent kernels (each CTA reads what the previous kernel wrote and writes its own slot) is captured into
// a graph; "work" spins each CTA for a given number of clock cycles to emulate a short kernel body.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <bool PDL>
__global__ void link_kernel(const float* in, float* out, int spin) {
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ float s[256];
  s[threadIdx.x] = (float)threadIdx.x;             // "prologue": work that does not depend on the predecessor
  __syncthreads();
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = in[i] + s[(threadIdx.x + 1) & 255];
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  out[i] = v;
}

static float run(bool pdl, int n, int ctas, int spin, float* a, float* b, cudaStream_t st) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    const float* in = (i & 1) ? b : a;
    float* out = (i & 1) ? a : b;
    if (pdl) CK(cudaLaunchKernelEx(&cfg, link_kernel<true>, in, out, spin));
    else     CK(cudaLaunchKernelEx(&cfg, link_kernel<false>, in, out, spin));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  const int reps = 10;
  for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
  return ms * 1e3f / (reps * n);
}

int main() {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  const int n = 200;
  float *a, *b;
  CK(cudaMalloc(&a, 148 * 8 * 256 * 4)); CK(cudaMalloc(&b, 148 * 8 * 256 * 4));
  CK(cudaMemset(a, 0, 148 * 8 * 256 * 4)); CK(cudaMemset(b, 0, 148 * 8 * 256 * 4));
  for (int ctas : {148, 148 * 4}) {
    for (int spin : {0, 2000, 10000}) {
      const float plain = run(false, n, ctas, spin, a, b, st);
      const float pdl = run(true, n, ctas, spin, a, b, st);
      printf("{\"ctas\": %d, \"spin_cycles\": %d, \"us_per_kernel_plain\": %.3f, \"us_per_kernel_pdl\": %.3f}\n", ctas, spin, plain, pdl);
    }
  }
  return 0;
}
