// Micro-probe: how fast can ONE thread per CTA pull a [80 x 2048] bf16 operand (327 KB) from L2 into shared memory
// through an 8-slot ring, when C CTAs read the SAME data at the same time (the access pattern of the persistent
// LSTM kernels)?  Variants: 2-D TMA boxes over strided rows (what lstm_seq.cu does), 2-D TMA boxes over a
// k-block-major (contiguous 10 KB) layout, 1-D bulk copies of the contiguous layout, and plain LDG by 512 threads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tools/tma_probe.cu -lcuda && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(ph) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)), "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

constexpr int ROWS = 80, KB = 32, SLOT = 16384, S = 8, BOX = ROWS * 128;

// mode 0: 2-D TMA strided rows; 1: 2-D TMA over the contiguous layout; 2: 1-D bulk; 3: LDG by all threads
__global__ void __launch_bounds__(512) probe(const __grid_constant__ CUtensorMap mS, const __grid_constant__ CUtensorMap mC, const uint8_t* contig,
                                             int mode, int steps, int distinct, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[S];
  if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int cta_off = distinct ? blockIdx.x : 0;
  long long t0 = clock64();
  if (mode < 3) {
    if (threadIdx.x == 0) {
      int it = 0;
      for (int st = 0; st < steps; ++st)
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          if (it >= S) mbar_wait(&full[s], ((it / S) - 1) & 1);      // previous occupant landed ("consumed" instantly)
          mbar_expect(&full[s], BOX);
          if (mode == 0) tma2d(sm + s * SLOT, &mS, kb * 64, cta_off * ROWS, &full[s]);
          else if (mode == 1) tma2d(sm + s * SLOT, &mC, 0, (cta_off * KB + kb) * ROWS, &full[s]);
          else bulk1d(sm + s * SLOT, contig + ((size_t)cta_off * KB + kb) * BOX, BOX, &full[s]);
        }
      for (int s = 0; s < S; ++s) { const int last = (it - 1 - s) / S; if (it - 1 - s >= 0) mbar_wait(&full[(it - 1 - s) % S], last & 1); }
    }
  } else {
    float4 acc = make_float4(0, 0, 0, 0);
    const float4* src = reinterpret_cast<const float4*>(contig + (size_t)cta_off * KB * BOX);
    for (int st = 0; st < steps; ++st)
      for (int i = threadIdx.x; i < KB * BOX / 16; i += 512 * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = i + u * 512 < KB * BOX / 16 ? __ldcg(src + i + u * 512) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
    if (acc.x == 1234.5f) out[1000] = 1;
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int maxC = 148, T = 18;
  // strided layout: [rows = maxC*80][cols = T*2048] bf16, we read the first 2048 columns (row stride 73728 B)
  size_t strided_bytes = (size_t)maxC * ROWS * T * 2048 * 2, contig_bytes = (size_t)maxC * KB * BOX;
  uint8_t *dS, *dC; long long* dOut;
  CK(cudaMalloc(&dS, strided_bytes)); CK(cudaMalloc(&dC, contig_bytes)); CK(cudaMalloc(&dOut, 2048 * 8));
  CK(cudaMemset(dS, 1, strided_bytes)); CK(cudaMemset(dC, 1, contig_bytes));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncFn enc = (EncFn)fp;
  CUtensorMap mS, mC;
  { cuuint64_t gd[2] = {(cuuint64_t)T * 2048, (cuuint64_t)maxC * ROWS}; cuuint64_t gs[1] = {(cuuint64_t)T * 2048 * 2}; cuuint32_t bx[2] = {64, ROWS}, es[2] = {1, 1};
    if (enc(&mS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dS, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc S failed\n"); return 1; } }
  { cuuint64_t gd[2] = {64, (cuuint64_t)maxC * KB * ROWS}; cuuint64_t gs[1] = {128}; cuuint32_t bx[2] = {64, ROWS}, es[2] = {1, 1};
    if (enc(&mC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dC, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc C failed\n"); return 1; } }
  const size_t smem = S * SLOT + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const char* names[4] = {"tma2d_strided", "tma2d_contig", "bulk1d_contig", "ldg_512thr"};
  const int steps = 8;
  for (int distinct = 0; distinct < 2; ++distinct)
    for (int C : {1, 32, 128})
      for (int mode = 0; mode < 4; ++mode) {
        long long best = 1LL << 60;
        for (int rep = 0; rep < 5; ++rep) {
          probe<<<C, 512, smem>>>(mS, mC, dC, mode, steps, distinct, dOut);
          CK(cudaDeviceSynchronize());
          long long h[148]; CK(cudaMemcpy(h, dOut, C * 8, cudaMemcpyDeviceToHost));
          long long mx = 0; for (int i = 0; i < C; ++i) mx = h[i] > mx ? h[i] : mx;
          if (mx < best) best = mx;
        }
        const double cyc_per_step = (double)best / steps;
        printf("{\"variant\": \"%s\", \"ctas\": %d, \"distinct_data\": %d, \"cycles_per_327KB_step\": %.0f, \"us_at_1.9GHz\": %.2f, \"bytes_per_clk_per_sm\": %.1f}\n",
               names[mode], C, distinct, cyc_per_step, cyc_per_step / 1900.0, (double)KB * BOX / cyc_per_step);
      }
  return 0;
}
