// Optimizer step next to the hot path: what the reference does after loss.backward() --
//   torch.nn.utils.clip_grad_norm_(model.decoder.LSTM.parameters(), 5)        train.py:213-214
//   torch.optim.Adam(params, lr=1e-3, betas=(0.8, 0.999), weight_decay=0)     model_factory.py:69-77, cfg_wzn.py:47-51
// -- as two launches over ALL parameter tensors: a sum-of-squares reduction over the clipped group, then one
// multi-tensor Adam update that applies the clip coefficient on the fly (the gradients are not rewritten unless asked).
// HBM-bound: reads p, g, m, v and writes p, m, v once (28 bytes per parameter).
#include "../../include/adaptive_b200.h"
#include "kernels.cuh"

namespace aa {

namespace {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) sumsq_kernel(const aa_opt_tensors t, float* __restrict__ out) {
  // grid.y = tensor; only tensors of the clipped group contribute
  const int ti = blockIdx.y;
  if (!t.clip[ti]) return;
  const float* g = t.grad[ti];
  const long long n = t.n[ti];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * OPT_THREADS) {
    const float x = g[i];
    acc = fmaf(x, x, acc);
  }
  acc = warp_sum(acc);
  __shared__ float red[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < OPT_THREADS / 32; ++i) s += red[i];
    atomicAdd(out, s);
  }
}

// torch.optim.Adam (no amsgrad, maximize = False), torch >= 1.x formulation:
//   g += wd * p ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(OPT_THREADS) adam_kernel(const aa_opt_tensors t, const float* __restrict__ sumsq, float max_norm, float lr,
                                                           float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                           int write_clipped_grads, float* __restrict__ norm_out) {
  const int ti = blockIdx.y;
  float* p = t.param[ti];
  float* g = t.grad[ti];
  float* m = t.m[ti];
  float* v = t.v[ti];
  const long long n = t.n[ti];
  float coef = 1.f;
  if (t.clip[ti] && max_norm > 0.f) {      // clip_grad_norm_: coef = min(1, max_norm / (||g|| + 1e-6))
    const float norm = sqrtf(*sumsq);
    coef = fminf(1.f, max_norm / (norm + 1e-6f));
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  }
  const float step = lr / bc1;
  for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * OPT_THREADS) {
    float gi = g[i] * coef;
    if (write_clipped_grads && coef != 1.f) g[i] = gi;
    const float pi = p[i];
    gi = fmaf(wd, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

}  // namespace
}  // namespace aa

using namespace aa;

extern "C" int aa_clip_adam_step(const aa_opt_tensors* t, int n_tensors, int step, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, float clip_max_norm, int write_clipped_grads, float* scratch, float* norm_out,
                                 void* stream) {
  AA_REQUIRE(t && n_tensors >= 1 && n_tensors <= AA_OPT_MAX_TENSORS, "aa_clip_adam_step: need 1..%d tensors", AA_OPT_MAX_TENSORS);
  AA_REQUIRE(step >= 1 && scratch, "aa_clip_adam_step: step counts from 1; scratch (one float) is required");
  long long nmax = 0;
  bool any_clip = false;
  for (int i = 0; i < n_tensors; ++i) {
    AA_REQUIRE(t->param[i] && t->grad[i] && t->m[i] && t->v[i] && t->n[i] >= 0, "aa_clip_adam_step: tensor %d has a null pointer", i);
    nmax = t->n[i] > nmax ? t->n[i] : nmax;
    any_clip |= t->clip[i] != 0;
  }
  if (nmax == 0) return AA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  long long bx = (nmax + OPT_THREADS * 4 - 1) / (OPT_THREADS * 4);
  const long long cap = (long long)num_sms() * 8 / n_tensors + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  const dim3 grid((unsigned)bx, (unsigned)n_tensors);
  AA_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float), st));
  if (any_clip && clip_max_norm > 0.f) {
    sumsq_kernel<<<grid, OPT_THREADS, 0, st>>>(*t, scratch);
    AA_CHECK_LAUNCH("sumsq");
  }
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<grid, OPT_THREADS, 0, st>>>(*t, scratch, clip_max_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, write_clipped_grads,
                                            norm_out);
  AA_CHECK_LAUNCH("adam");
  return AA_OK;
}
