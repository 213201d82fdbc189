// Micro-probe: cost of back-to-back tcgen05.mma (kind::f16, bf16 operands from shared memory, SS mode) issued by one
// thread, as a function of the tile N, the UMMA M and the number of independent TMEM accumulators the k-steps rotate
// over.  Answers: what does one small-N MMA cost inside the persistent LSTM kernels (N = 16), and does spreading
// dependent accumulations over several accumulators help?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/mma_probe.cu && ./mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(s32(b)), "r"(ph) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // SW128, K-major, SBO 1024
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((16 >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// TF32: kind::tf32 (K = 8 per instruction, same 32-byte k-step in shared memory)
template <bool TF32>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (TF32)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred;
}

// warp_wide = 0: `if (threadIdx.x == 0)` issues (what lstm_seq.cu / gemm_tc.cu do); 1: the whole warp runs the loop
// converged and the MMA is issued under elect.sync (operands are warp-uniform)
template <bool TF32>
__global__ void __launch_bounds__(128) probe(int N, int M, int nacc, int iters, int commit_every, int warp_wide, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  uint8_t* sA = sm;             // 4 x 16 KB
  uint8_t* sB = sm + 65536;     // 32 KB (N <= 256)
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // bf16 ~0.0078
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp_wide ? threadIdx.x < 32 : threadIdx.x == 0) {
    const uint32_t fmt = TF32 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t db0 = make_desc(s32(sB));
    long long t0 = clock64();
    int j = 0;
    for (int it = 0; it < iters; ++it) {
      const uint64_t da0 = make_desc(s32(sA + (it & 3) * 16384));
      if (warp_wide) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) mma<TF32>(tmem + (uint32_t)(((j + k) % nacc) * N), da0 + 2 * k, db0 + 2 * k, idesc, (j + k) >= nacc ? 1u : 0u);
          if (commit_every) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar2)) : "memory");
        }
        j += 4;
        __syncwarp();
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k, ++j) mma<TF32>(tmem + (uint32_t)((j % nacc) * N), da0 + 2 * k, db0 + 2 * k, idesc, j >= nacc ? 1u : 0u);
        if (commit_every) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar2)) : "memory");
      }
    }
    long long t1 = clock64();
    if (!warp_wide || elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 16));
  const size_t smem = 65536 + 32768 + 1024;
  CK(cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 256;   // x4 MMAs
  for (int M : {128})
    for (int N : {16, 32, 64, 128, 256})
      for (int nacc : {1, 4})
        for (int ce : {0, 1})
         for (int ww : {0, 1}) {
          if (N * nacc > 512) continue;
          long long best[2] = {1LL << 60, 1LL << 60};
          for (int rep = 0; rep < 3; ++rep) {
            probe<false><<<1, 128, smem>>>(N, M, nacc, iters, ce, ww, d);
            CK(cudaDeviceSynchronize());
            long long h[2];
            CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
            if (h[1] < best[1]) { best[0] = h[0]; best[1] = h[1]; }
          }
          printf("{\"M\": %d, \"N\": %d, \"accumulators\": %d, \"commit_per_kblock\": %d, \"warp_wide_elect\": %d, \"issue_cycles_per_mma\": %.1f, \"total_cycles_per_mma\": %.1f}\n", M, N, nacc, ce, ww,
                 (double)best[0] / (iters * 4), (double)best[1] / (iters * 4));
        }
  // kind::tf32 against kind::f16 at the tile widths the contractions use (converged warp, one accumulator)
  for (int tf : {0, 1})
    for (int N : {64, 128, 256}) {
      long long best[2] = {1LL << 60, 1LL << 60};
      for (int rep = 0; rep < 3; ++rep) {
        if (tf) probe<true><<<1, 128, smem>>>(N, 128, 1, iters, 1, 1, d);
        else probe<false><<<1, 128, smem>>>(N, 128, 1, iters, 1, 1, d);
        CK(cudaDeviceSynchronize());
        long long h[2];
        CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        if (h[1] < best[1]) { best[0] = h[0]; best[1] = h[1]; }
      }
      const double cyc = (double)best[1] / (iters * 4);
      printf("{\"kind\": \"%s\", \"M\": 128, \"N\": %d, \"K_per_mma\": %d, \"total_cycles_per_mma\": %.1f, \"flop_per_clk\": %.0f}\n", tf ? "tf32" : "f16(bf16)", N,
             tf ? 8 : 16, cyc, 2.0 * 128 * N * (tf ? 8 : 16) / cyc);
    }
  return 0;
}
