// Shared device/host helpers for libadaptive_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/adaptive_b200.h"

namespace aa {

// ---- error plumbing (thread-local message, C-ABI returns int codes) ------------------
// (codes AA_OK / AA_ERR_* come from the public header)
void set_error(const char* fmt, ...);
const char* get_error();

#define AA_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      aa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AA_ERR_CUDA;                                                                \
    }                                                                                        \
  } while (0)

#define AA_CHECK_LAUNCH(name)                                                                \
  do {                                                                                       \
    aa::count_launch();                                                                      \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      aa::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AA_ERR_CUDA;                                                                \
    }                                                                                        \
  } while (0)

#define AA_REQUIRE(cond, ...)                                                                \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      aa::set_error(__VA_ARGS__);                                                            \
      return AA_ERR_INVALID;                                                             \
    }                                                                                        \
  } while (0)

#define AA_TRY(expr)                                                                         \
  do {                                                                                       \
    int _rc = (expr);                                                                        \
    if (_rc != 0) return _rc;                                                                \
  } while (0)

// run `call` (an int-returning launcher) inside a profiling scope named `tag`
#define AA_PROF(tag, stream, call)          \
  do {                                      \
    aa::ProfScope _ps(tag, stream);         \
    int _rc = (call);                       \
    if (_rc != 0) return _rc;               \
  } while (0)

int num_sms();

// ---- instrumentation: launch counter + optional per-kernel CUDA-event timing ---------------
void count_launch();
// Records a cudaEvent pair around the enclosed launches on `stream` when profiling is enabled
// (aa_profile_enable); totals are resolved lazily by aa_profile_get.  No-op otherwise.
struct ProfScope {
  int slot;
  cudaStream_t stream;
  ProfScope(const char* tag, cudaStream_t s);
  ~ProfScope();
};

// ---- device math -----------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// fast variants (ex2.approx / rcp.approx based, abs error ~1e-7): bf16 mixed-precision path and attention scores
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_fast(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// fp32 -> (hi, lo) split for the fp32-accurate "3xTF32" contractions: hi = round-to-nearest tf32 (low 13 mantissa
// bits zero), lo = tf32(x - hi) (x - hi is exact in fp32).  hi + lo carries 21+ mantissa bits of x.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = tf32_rn(x);
  lo = tf32_rn(x - hi);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- generic fp32 GEMM (gemm_simt.cu) --------------------------------------------------
// D[m,n] = alpha * sum_k A(m,k) * B(k,n) + beta * Cin[m,n] + bias1[n] + bias2[n]
//   a_kcontig: A(m,k) = A[m*lda + k]   else A(m,k) = A[k*lda + m]
//   b_kcontig: B(k,n) = B[n*ldb + k]   else B(k,n) = B[k*ldb + n]
// splitk > 1: D must be pre-zeroed, Cin/bias must be null; partial sums are atomically added.
struct GemmArgs {
  int M, N, K;
  const float* A; long long lda; int a_kcontig;
  const float* B; long long ldb; int b_kcontig;
  const float* Cin; long long ldcin;
  float* D; long long ldd;
  const float* bias1; const float* bias2;
  float alpha, beta;
  int splitk;
};
int launch_sgemm(const GemmArgs& g, cudaStream_t stream);

// convenience wrappers
// Y[M,N] = X[M,K] * W[N,K]^T (+Cin) (+bias)    -- forward linear
int gemm_nt(int M, int N, int K, const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy,
            const float* Cin, long long ldcin, const float* b1, const float* b2, cudaStream_t s);
// dX[M,K] = dY[M,N] * W[N,K] (+Cin)            -- input gradient
int gemm_nn(int M, int K, int N, const float* dY, long long lddy, const float* W, long long ldw, float* dX, long long lddx,
            const float* Cin, long long ldcin, cudaStream_t s);
// dW[N,K] (+)= dY[M,N]^T * X[M,K]              -- weight gradient (reduction over rows)
int gemm_tn(int N, int K, int M, const float* dY, long long lddy, const float* X, long long ldx, float* dW, long long lddw,
            bool accumulate, cudaStream_t s);

}  // namespace aa
