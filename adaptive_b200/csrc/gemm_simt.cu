// Generic fp32 SIMT GEMM for sm_100a: register-blocked outer product, double-buffered
// shared-memory tiles, 128-bit global loads.  It is the exact-fp32 contraction engine of the
// library (decode path, weight gradients in fp32 mode) and the on-device reference the
// tcgen05 GEMMs are validated against.
#include <stdlib.h>

#include "common.cuh"

namespace aa {

namespace {

constexpr int PAD = 4;

// Load one [R x BK] operand tile into registers.  kcontig: elem(r,k) = p[r*ld + k];
// otherwise elem(r,k) = p[k*ld + r].
template <int R, int BK, int NT, bool KCONTIG>
struct TileLoader {
  static constexpr int NV = R * BK / 4 / NT;
  static_assert(NV >= 1 && (R * BK / 4) % NT == 0, "tile/thread mismatch");
  float4 v[NV];

  __device__ __forceinline__ void load(const float* __restrict__ p, long long ld, int r0, int rmax, int k0, int kmax,
                                       bool vec_ok, int tid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = tid + i * NT;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (KCONTIG) {
        const int r = r0 + f / (BK / 4);
        const int k = k0 + (f % (BK / 4)) * 4;
        if (r < rmax && k < kmax) {
          const float* q = p + (long long)r * ld + k;
          if (vec_ok && k + 3 < kmax) {
            x = ldg4(q);
          } else {
            x.x = __ldg(q);
            if (k + 1 < kmax) x.y = __ldg(q + 1);
            if (k + 2 < kmax) x.z = __ldg(q + 2);
            if (k + 3 < kmax) x.w = __ldg(q + 3);
          }
        }
      } else {
        const int k = k0 + f / (R / 4);
        const int r = r0 + (f % (R / 4)) * 4;
        if (k < kmax && r < rmax) {
          const float* q = p + (long long)k * ld + r;
          if (vec_ok && r + 3 < rmax) {
            x = ldg4(q);
          } else {
            x.x = __ldg(q);
            if (r + 1 < rmax) x.y = __ldg(q + 1);
            if (r + 2 < rmax) x.z = __ldg(q + 2);
            if (r + 3 < rmax) x.w = __ldg(q + 3);
          }
        }
      }
      v[i] = x;
    }
  }

  __device__ __forceinline__ void store(float (*S)[R + PAD], int tid) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = tid + i * NT;
      if (KCONTIG) {
        const int r = f / (BK / 4);
        const int k = (f % (BK / 4)) * 4;
        S[k + 0][r] = v[i].x;
        S[k + 1][r] = v[i].y;
        S[k + 2][r] = v[i].z;
        S[k + 3][r] = v[i].w;
      } else {
        const int k = f / (R / 4);
        const int r = (f % (R / 4)) * 4;
        *reinterpret_cast<float4*>(&S[k][r]) = v[i];
      }
    }
  }
};

template <int BM, int BN, int BK, int TM, int TN, bool AK, bool BKC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) sgemm_kernel(const GemmArgs g) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int CM = TM / 4, CN = TN / 4;
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile must be a multiple of 4");
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  int kbeg = 0, kend = g.K;
  if (g.splitk > 1) {
    const int per = ((g.K + BK - 1) / BK + g.splitk - 1) / g.splitk * BK;
    kbeg = blockIdx.z * per;
    kend = min(g.K, kbeg + per);
    if (kbeg >= kend) return;
  }
  const int nk = (kend - kbeg + BK - 1) / BK;

  const bool a_vec = ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && (g.lda % 4 == 0);
  const bool b_vec = ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && (g.ldb % 4 == 0);

  TileLoader<BM, BK, NT, AK> la;
  TileLoader<BN, BK, NT, BKC> lb;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  la.load(g.A, g.lda, m0, g.M, kbeg, kend, a_vec, tid);
  lb.load(g.B, g.ldb, n0, g.N, kbeg, kend, b_vec, tid);
  la.store(As[0], tid);
  lb.store(Bs[0], tid);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      la.load(g.A, g.lda, m0, g.M, kbeg + (kt + 1) * BK, kend, a_vec, tid);
      lb.load(g.B, g.ldb, n0, g.N, kbeg + (kt + 1) * BK, kend, b_vec, tid);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int c = 0; c < CM; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][c * (BM / CM) + ty * 4]);
        a[c * 4 + 0] = t.x; a[c * 4 + 1] = t.y; a[c * 4 + 2] = t.z; a[c * 4 + 3] = t.w;
      }
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][c * (BN / CN) + tx * 4]);
        b[c * 4 + 0] = t.x; b[c * 4 + 1] = t.y; b[c * 4 + 2] = t.z; b[c * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      la.store(As[buf ^ 1], tid);
      lb.store(Bs[buf ^ 1], tid);
    }
    __syncthreads();
  }

  // epilogue
  const bool d_vec = ((reinterpret_cast<uintptr_t>(g.D) & 15) == 0) && (g.ldd % 4 == 0) &&
                     (g.Cin == nullptr || (((reinterpret_cast<uintptr_t>(g.Cin) & 15) == 0) && (g.ldcin % 4 == 0)));
#pragma unroll
  for (int cm = 0; cm < CM; ++cm)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + cm * (BM / CM) + ty * 4 + i;
      if (m >= g.M) continue;
#pragma unroll
      for (int cn = 0; cn < CN; ++cn) {
        const int n = n0 + cn * (BN / CN) + tx * 4;
        if (n >= g.N) continue;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = g.alpha * acc[cm * 4 + i][cn * 4 + j];
        float* dp = g.D + (long long)m * g.ldd + n;
        if (g.splitk > 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < g.N) atomicAdd(dp + j, o[j]);
          continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j < g.N) {
            if (g.bias1) o[j] += __ldg(g.bias1 + n + j);
            if (g.bias2) o[j] += __ldg(g.bias2 + n + j);
          }
        }
        if (g.Cin) {
          const float* cp = g.Cin + (long long)m * g.ldcin + n;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < g.N) o[j] += g.beta * cp[j];
        }
        if (d_vec && n + 3 < g.N) {
          *reinterpret_cast<float4*>(dp) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < g.N) dp[j] = o[j];
        }
      }
    }
}

template <int BM, int BN, int BK, int TM, int TN>
int dispatch_layout(const GemmArgs& g, cudaStream_t stream) {
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), g.splitk > 1 ? g.splitk : 1);
  dim3 block((BM / TM) * (BN / TN));
  if (g.a_kcontig && g.b_kcontig)
    sgemm_kernel<BM, BN, BK, TM, TN, true, true><<<grid, block, 0, stream>>>(g);
  else if (g.a_kcontig && !g.b_kcontig)
    sgemm_kernel<BM, BN, BK, TM, TN, true, false><<<grid, block, 0, stream>>>(g);
  else if (!g.a_kcontig && g.b_kcontig)
    sgemm_kernel<BM, BN, BK, TM, TN, false, true><<<grid, block, 0, stream>>>(g);
  else
    sgemm_kernel<BM, BN, BK, TM, TN, false, false><<<grid, block, 0, stream>>>(g);
  AA_CHECK_LAUNCH("sgemm_kernel");
  return AA_OK;
}

}  // namespace

int launch_sgemm(const GemmArgs& gin, cudaStream_t stream) {
  GemmArgs g = gin;
  if (g.M <= 0 || g.N <= 0) return AA_OK;
  AA_REQUIRE(g.K > 0, "sgemm: K must be positive (got %d)", g.K);
  AA_REQUIRE(g.A && g.B && g.D, "sgemm: null operand");
  if (g.splitk > 1) AA_REQUIRE(!g.Cin && !g.bias1 && !g.bias2, "sgemm: split-K needs a pre-zeroed D and no Cin/bias");
  const long long sms = num_sms();
  auto tiles = [&](int bm, int bn) { return (long long)ceil_div(g.M, bm) * ceil_div(g.N, bn); };
  // pick the largest tile that still fills the chip; fall back to split-K for long, thin reductions
  // (a 128-column tile wastes most of its work on N <= 64 outputs -- P = V W_v^T has 49 -- when the 64-column tiles fill the chip too)
  if (g.N <= 64 && tiles(128, 64) >= sms) {   // tall and narrow (P): 128 x 64 tiles, 8 x 4 outputs per thread
    static const bool tall = [] { const char* e = getenv("AA_SGEMM_128X64"); return !e || e[0] != '0'; }();
    if (tall) return dispatch_layout<128, 64, 16, 8, 4>(g, stream);
  }
  if (tiles(128, 128) >= sms && !(g.N <= 64 && tiles(64, 64) >= sms)) return dispatch_layout<128, 128, 16, 8, 8>(g, stream);
  if (tiles(64, 64) >= sms || g.M > 32) {
    if (g.splitk == 0) {  // auto split-K
      const long long t = tiles(64, 64);
      int want = (int)((sms + t - 1) / t);
      int maxs = g.K / 256;
      g.splitk = (want > 1 && maxs > 1 && !g.Cin && !g.bias1 && !g.bias2 && g.beta == 0.f) ? (want < maxs ? want : maxs) : 1;
      if (g.splitk > 1) AA_CHECK_CUDA(cudaMemset2DAsync(g.D, g.ldd * sizeof(float), 0, (size_t)g.N * sizeof(float), g.M, stream));
    }
    return dispatch_layout<64, 64, 16, 4, 4>(g, stream);
  }
  if (g.splitk == 0) g.splitk = 1;
  return dispatch_layout<32, 64, 16, 4, 4>(g, stream);
}

int gemm_nt(int M, int N, int K, const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy,
            const float* Cin, long long ldcin, const float* b1, const float* b2, cudaStream_t s) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = X; g.lda = ldx; g.a_kcontig = 1;
  g.B = W; g.ldb = ldw; g.b_kcontig = 1;
  g.Cin = Cin; g.ldcin = ldcin; g.D = Y; g.ldd = ldy;
  g.bias1 = b1; g.bias2 = b2; g.alpha = 1.f; g.beta = Cin ? 1.f : 0.f; g.splitk = 1;
  return launch_sgemm(g, s);
}

int gemm_nn(int M, int K, int N, const float* dY, long long lddy, const float* W, long long ldw, float* dX, long long lddx,
            const float* Cin, long long ldcin, cudaStream_t s) {
  // dX[m,k] = sum_n dY[m,n] W[n,k] : reduction index n; B(n,k) = W[n*ldw + k] -> not k-contig
  GemmArgs g{};
  g.M = M; g.N = K; g.K = N;
  g.A = dY; g.lda = lddy; g.a_kcontig = 1;
  g.B = W; g.ldb = ldw; g.b_kcontig = 0;
  g.Cin = Cin; g.ldcin = ldcin; g.D = dX; g.ldd = lddx;
  g.alpha = 1.f; g.beta = Cin ? 1.f : 0.f; g.splitk = 1;
  return launch_sgemm(g, s);
}

int gemm_tn(int N, int K, int M, const float* dY, long long lddy, const float* X, long long ldx, float* dW, long long lddw,
            bool accumulate, cudaStream_t s) {
  // dW[n,k] = sum_m dY[m,n] X[m,k] : A(n,m) = dY[m*lddy + n] (not k-contig), B(m,k) = X[m*ldx + k] (not k-contig)
  GemmArgs g{};
  g.M = N; g.N = K; g.K = M;
  g.A = dY; g.lda = lddy; g.a_kcontig = 0;
  g.B = X; g.ldb = ldx; g.b_kcontig = 0;
  g.Cin = accumulate ? dW : nullptr; g.ldcin = lddw; g.D = dW; g.ldd = lddw;
  g.alpha = 1.f; g.beta = accumulate ? 1.f : 0.f;
  g.splitk = accumulate ? 1 : 0;  // 0 = let the launcher split the reduction when the grid is small
  return launch_sgemm(g, s);
}

}  // namespace aa
