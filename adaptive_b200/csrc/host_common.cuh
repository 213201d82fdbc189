// Host-side helpers shared by the orchestration files (capi.cu, encoder.cu): blob carving and the
// precision-dispatching contraction wrappers.
#pragma once
#include "kernels.cuh"

namespace aa {
namespace {

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  float* take(size_t nfloats) {
    float* p = reinterpret_cast<float*>(base + off);
    off += align_up(nfloats * sizeof(float), 256);
    return p;
  }
};

typedef __nv_bfloat16 bf16;

struct Carver16 {   // carve helper for bf16 arrays sharing the same blob
  Carver& c;
  bf16* take(size_t n) { return reinterpret_cast<bf16*>(c.take((n + 1) / 2)); }
};

// ---- precision-dispatching contractions --------------------------------------------------
// A matrix that exists as fp32 and (in bf16 mode) as a bf16 mirror, each with its own row stride.
struct Mat {
  const float* f; long long ldf;
  const bf16* h; long long ldh;
};
inline Mat M32(const float* f, long long ld) { return Mat{f, ld, nullptr, 0}; }
inline Mat M2(const float* f, long long ldf, const bf16* h, long long ldh) { return Mat{f, ldf, h, ldh}; }

struct Ctx {
  int prec;
  cudaStream_t st;
  bool tc() const { return prec == AA_PREC_BF16; }
};

// Y[M,N] = X[M,K] W[N,K]^T (+Cin) (+b1+b2)
int mm_nt(const Ctx& c, const char* tag, int M, int N, int K, Mat X, Mat W, float* Y, long long ldy, const float* Cin,
          long long ldcin, const float* b1, const float* b2) {
  ProfScope ps(tag, c.st);
  if (!c.tc()) return gemm_nt(M, N, K, X.f, X.ldf, W.f, W.ldf, Y, ldy, Cin, ldcin, b1, b2, c.st);
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.elem_size = 2;
  g.A = X.h; g.lda = X.ldh; g.a_mn = 0;
  g.B = W.h; g.ldb = W.ldh; g.b_mn = 0;
  g.D32 = Y; g.ldd32 = ldy; g.Cin = Cin; g.ldcin = ldcin; g.beta = 1.f; g.bias1 = b1; g.bias2 = b2;
  return launch_gemm_tc(g, c.st);
}
// dX[M,K] = dY[M,N] W[N,K] (+Cin)
int mm_nn(const Ctx& c, const char* tag, int M, int K, int N, Mat dY, Mat W, float* dX, long long lddx, const float* Cin,
          long long ldcin) {
  ProfScope ps(tag, c.st);
  if (!c.tc()) return gemm_nn(M, K, N, dY.f, dY.ldf, W.f, W.ldf, dX, lddx, Cin, ldcin, c.st);
  TcGemmArgs g{};
  g.M = M; g.N = K; g.K = N; g.elem_size = 2;
  g.A = dY.h; g.lda = dY.ldh; g.a_mn = 0;
  g.B = W.h; g.ldb = W.ldh; g.b_mn = 1;          // W stored [N(reduction), K(out)]: MN-major B
  g.D32 = dX; g.ldd32 = lddx; g.Cin = Cin; g.ldcin = ldcin; g.beta = 1.f;
  return launch_gemm_tc(g, c.st);
}
// dW[N,K] (+)= dY[M,N]^T X[M,K]
int mm_tn(const Ctx& c, const char* tag, int N, int K, int M, Mat dY, Mat X, float* dW, long long lddw, bool accumulate) {
  ProfScope ps(tag, c.st);
  if (!c.tc()) return gemm_tn(N, K, M, dY.f, dY.ldf, X.f, X.ldf, dW, lddw, accumulate, c.st);
  TcGemmArgs g{};
  g.M = N; g.N = K; g.K = M; g.elem_size = 2;
  g.A = dY.h; g.lda = dY.ldh; g.a_mn = 1;        // both operands are stored reduction-major
  g.B = X.h; g.ldb = X.ldh; g.b_mn = 1;
  g.D32 = dW; g.ldd32 = lddw; g.Cin = accumulate ? dW : nullptr; g.ldcin = lddw; g.beta = 1.f;
  return launch_gemm_tc(g, c.st);
}

}  // namespace
}  // namespace aa
