// Fused adaptive-attention kernels for the teacher-forced (training) path, fp32.
//
// Forward  (Atten.forward, adaptive_attention.py:26-58 + the `c_hat + h` of :132), per image b and step t:
//   z_i = w_h . tanh(P_i + q_t)  (i<k),  z_s = w_h . tanh(r_t)          r_t = s_t W_s^T + q_t
//   alpha = softmax_k(z) ; beta = softmax_{k+1}([z; z_s])[k]
//   ctx = sum_i alpha_i V_i ; c_hat = beta s + (1-beta) ctx ; u = c_hat + h
// Nothing of size [B,T,k,a] is ever materialised (the reference writes it twice); the backward
// recomputes the tanh terms from P and q.
#include "kernels.cuh"

namespace aa {

int g_atten_sequential = 0;   // diagnostics (aa_debug_set_atten_sequential): force the step-by-step kernels

namespace {

constexpr int AT_THREADS = 256;
constexpr int AT_WARPS = AT_THREADS / 32;
constexpr int MAXJ = 4;  // a <= 32*MAXJ

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS) atten_fwd_kernel(const AttenFwdArgs p, int t_per_cta) {
  extern __shared__ __align__(16) float sm[];
  const int k = p.k, a = p.a, H = p.H, T = p.T;
  float* Ps = sm;              // [k*a]
  float* whs = Ps + k * a;     // [a]
  float* qs = whs + a;         // [a]
  float* rs = qs + a;          // [a]
  float* zs = rs + a;          // [k+1]
  float* als = zs + (k + 1);   // [k]
  float* misc = als + k;       // [4]

  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t_begin = blockIdx.y * t_per_cta;
  const int t_end = min(T, t_begin + t_per_cta);

  const float* Pb = p.P + (long long)b * k * a;
  for (int i = tid; i < k * a; i += AT_THREADS) Ps[i] = Pb[i];
  for (int j = tid; j < a; j += AT_THREADS) whs[j] = p.wh[j];
  const float* Vb = p.V + (long long)b * k * H;

  for (int t = t_begin; t < t_end; ++t) {
    const long long row = (long long)b * T + t;
    for (int j = tid; j < a; j += AT_THREADS) {
      qs[j] = p.q[row * a + j];
      rs[j] = p.r[row * a + j];
    }
    __syncthreads();
    // scores: one warp per region row (row k = sentinel)
    for (int i = warp; i <= k; i += AT_WARPS) {
      float acc = 0.f;
      if (i < k) {
        for (int j = lane; j < a; j += 32) acc = fmaf(whs[j], tanhf(Ps[i * a + j] + qs[j]), acc);
      } else {
        for (int j = lane; j < a; j += 32) acc = fmaf(whs[j], tanhf(rs[j]), acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) zs[i] = acc;
    }
    __syncthreads();
    if (warp == 0) {
      float m = -INFINITY;
      for (int i = lane; i < k; i += 32) m = fmaxf(m, zs[i]);
      m = warp_max(m);
      float sum = 0.f;
      for (int i = lane; i < k; i += 32) sum += expf(zs[i] - m);
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      for (int i = lane; i < k; i += 32) {
        const float al = expf(zs[i] - m) * inv;
        als[i] = al;
        p.alpha[row * k + i] = al;
      }
      // (k+1)-way softmax, last entry
      const float zsent = zs[k];
      const float m1 = fmaxf(m, zsent);
      float sum1 = 0.f;
      for (int i = lane; i < k; i += 32) sum1 += expf(zs[i] - m1);
      sum1 = warp_sum(sum1);
      const float es = expf(zsent - m1);
      const float beta = p.no_sentinel ? 0.f : es / (sum1 + es);   // (sentinel-less baseline block: c_hat = ctx)
      if (lane == 0) {
        misc[0] = beta;
        p.beta[row] = beta;
      }
    }
    __syncthreads();
    const float beta = misc[0];
    for (int c = tid * 4; c < H; c += AT_THREADS * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 7
      for (int i = 0; i < k; ++i) {
        const float4 v = ldg4(Vb + (long long)i * H + c);
        const float al = als[i];
        acc.x = fmaf(al, v.x, acc.x); acc.y = fmaf(al, v.y, acc.y);
        acc.z = fmaf(al, v.z, acc.z); acc.w = fmaf(al, v.w, acc.w);
      }
      const float4 sv = *reinterpret_cast<const float4*>(p.s + row * H + c);
      const float4 hv = *reinterpret_cast<const float4*>(p.h + row * H + c);
      float4 ch;
      ch.x = beta * sv.x + (1.f - beta) * acc.x;
      ch.y = beta * sv.y + (1.f - beta) * acc.y;
      ch.z = beta * sv.z + (1.f - beta) * acc.z;
      ch.w = beta * sv.w + (1.f - beta) * acc.w;
      if (p.ctx) *reinterpret_cast<float4*>(p.ctx + row * H + c) = acc;
      if (p.c_hat) *reinterpret_cast<float4*>(p.c_hat + row * H + c) = ch;
      if (p.u) *reinterpret_cast<float4*>(p.u + row * H + c) = make_float4(ch.x + hv.x, ch.y + hv.y, ch.z + hv.z, ch.w + hv.w);
      if (p.u16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(ch.x + hv.x, ch.y + hv.y), hi = __floats2bfloat162_rn(ch.z + hv.z, ch.w + hv.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p.u16 + row * H + c) = pk;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS) atten_bwd_kernel(const AttenBwdArgs p, int TC, int t_per_cta, int atomic_out) {
  extern __shared__ __align__(16) float sm[];
  const int k = p.k, a = p.a, H = p.H, T = p.T;
  float* Ps = sm;                       // [k*a]
  float* dPs = Ps + k * a;              // [k*a]
  float* whs = dPs + k * a;             // [a]
  float* qs = whs + a;                  // [a]
  float* rs = qs + a;                   // [a]
  float* das = rs + a;                  // [k]   d_alpha -> dz
  float* colred = das + k;              // [AT_WARPS * a]
  float* misc = colred + AT_WARPS * a;  // [8 + AT_WARPS]
  float* als = misc + 8 + AT_WARPS;     // [TC*k]
  // dctx rows are accessed as float4: round their offset up to 4 floats (sm itself is 16B aligned)
  const size_t dcs_off = ((size_t)(als - sm) + (size_t)TC * k + 3) & ~(size_t)3;
  float* dcs = sm + dcs_off;            // [TC*H]

  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t_begin = blockIdx.y * t_per_cta;
  const int t_end = min(T, t_begin + t_per_cta);

  const float* Pb = p.P + (long long)b * k * a;
  for (int i = tid; i < k * a; i += AT_THREADS) { Ps[i] = Pb[i]; dPs[i] = 0.f; }
  for (int j = tid; j < a; j += AT_THREADS) whs[j] = p.wh[j];
  const float* Vb = p.V + (long long)b * k * H;
  float* dVb = p.dV + (long long)b * k * H;

  float dwh_part[MAXJ];   // per (warp,lane) partial of dw_h over this CTA's rows and steps
#pragma unroll
  for (int jj = 0; jj < MAXJ; ++jj) dwh_part[jj] = 0.f;
  float dwh_sent = 0.f;   // thread j < a: sentinel-row contribution
  __syncthreads();

  bool first_chunk = true;
  for (int t0 = t_begin; t0 < t_end; t0 += TC) {
    const int t1 = min(t_end, t0 + TC);
    for (int t = t0; t < t1; ++t) {
      const int tl = t - t0;
      const long long row = (long long)b * T + t;
      const float beta = p.beta[row];
      // (a) beta gate: dbeta, ds = beta*dchat, dctx = (1-beta)*dchat
      float part = 0.f;
      for (int c = tid * 4; c < H; c += AT_THREADS * 4) {
        const float4 d = *reinterpret_cast<const float4*>(p.dchat + row * H + c);
        const float4 sv = *reinterpret_cast<const float4*>(p.s + row * H + c);
        const float4 cx = *reinterpret_cast<const float4*>(p.ctx + row * H + c);
        part += d.x * (sv.x - cx.x) + d.y * (sv.y - cx.y) + d.z * (sv.z - cx.z) + d.w * (sv.w - cx.w);
        *reinterpret_cast<float4*>(p.ds + row * H + c) = make_float4(beta * d.x, beta * d.y, beta * d.z, beta * d.w);
        const float ob = 1.f - beta;
        *reinterpret_cast<float4*>(dcs + (size_t)tl * H + c) = make_float4(ob * d.x, ob * d.y, ob * d.z, ob * d.w);
      }
      part = warp_sum(part);
      if (lane == 0) misc[8 + warp] = part;
      for (int i = tid; i < k; i += AT_THREADS) als[(size_t)tl * k + i] = p.alpha[row * k + i];
      for (int j = tid; j < a; j += AT_THREADS) { qs[j] = p.q[row * a + j]; rs[j] = p.r[row * a + j]; }
      __syncthreads();
      // (b) d_alpha_i = dctx . V_i
      for (int i = warp; i < k; i += AT_WARPS) {
        float acc = 0.f;
        for (int c = lane * 4; c < H; c += 128) {
          const float4 v = ldg4(Vb + (long long)i * H + c);
          const float4 d = *reinterpret_cast<const float4*>(dcs + (size_t)tl * H + c);
          acc += v.x * d.x + v.y * d.y + v.z * d.z + v.w * d.w;
        }
        acc = warp_sum(acc);
        if (lane == 0) das[i] = acc + (p.d_alpha ? p.d_alpha[row * k + i] : 0.f);
      }
      __syncthreads();
      // (c) the two softmaxes
      if (warp == 0) {
        float dbeta = 0.f;
#pragma unroll
        for (int w = 0; w < AT_WARPS; ++w) dbeta += misc[8 + w];
        if (p.d_beta) dbeta += p.d_beta[row];
        float S = 0.f;
        for (int i = lane; i < k; i += 32) S += als[(size_t)tl * k + i] * das[i];
        S = warp_sum(S);
        const float b1 = beta * (1.f - beta) * dbeta;
        for (int i = lane; i < k; i += 32) {
          const float al = als[(size_t)tl * k + i];
          das[i] = al * (das[i] - S) - b1 * al;
        }
        if (lane == 0) misc[0] = b1;  // dz_s
      }
      __syncthreads();
      // (d) score backward with tanh recompute
      float dq_part[MAXJ];
#pragma unroll
      for (int jj = 0; jj < MAXJ; ++jj) dq_part[jj] = 0.f;
      for (int i = warp; i < k; i += AT_WARPS) {
        const float dzv = das[i];
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
          const int j = lane + jj * 32;
          if (j < a) {
            const float tp = tanhf(Ps[i * a + j] + qs[j]);
            const float dp = dzv * whs[j] * (1.f - tp * tp);
            dPs[i * a + j] += dp;
            dq_part[jj] += dp;
            dwh_part[jj] = fmaf(dzv, tp, dwh_part[jj]);
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < MAXJ; ++jj) {
        const int j = lane + jj * 32;
        if (j < a) colred[warp * a + j] = dq_part[jj];
      }
      __syncthreads();
      if (tid < a) {
        const int j = tid;
        float dq = 0.f;
#pragma unroll
        for (int w = 0; w < AT_WARPS; ++w) dq += colred[w * a + j];
        const float dzs = misc[0];
        const float tr = tanhf(rs[j]);
        const float drj = dzs * whs[j] * (1.f - tr * tr);
        dwh_sent = fmaf(dzs, tr, dwh_sent);
        p.dr[row * a + j] = drj;
        p.dq[row * a + j] = dq + drj;
        if (p.dq16) {
          p.dr16[row * p.a_pad + j] = __float2bfloat16(drj);
          p.dq16[row * p.a_pad + j] = __float2bfloat16(dq + drj);
        }
      }
      __syncthreads();
    }
    // (e) dV_i (+)= sum_t alpha_{t,i} dctx_t   for this chunk of steps
    const int nt = t1 - t0;
    for (int c = tid * 4; c < H; c += AT_THREADS * 4) {
      for (int i = 0; i < k; ++i) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int tl = 0; tl < nt; ++tl) {
          const float al = als[(size_t)tl * k + i];
          const float4 d = *reinterpret_cast<const float4*>(dcs + (size_t)tl * H + c);
          acc.x = fmaf(al, d.x, acc.x); acc.y = fmaf(al, d.y, acc.y);
          acc.z = fmaf(al, d.z, acc.z); acc.w = fmaf(al, d.w, acc.w);
        }
        float* dst = dVb + (long long)i * H + c;
        if (atomic_out) {
          atomicAdd(dst + 0, acc.x); atomicAdd(dst + 1, acc.y); atomicAdd(dst + 2, acc.z); atomicAdd(dst + 3, acc.w);
        } else if (first_chunk) {
          *reinterpret_cast<float4*>(dst) = acc;
        } else {
          float4 o = *reinterpret_cast<float4*>(dst);
          o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
          *reinterpret_cast<float4*>(dst) = o;
        }
      }
    }
    first_chunk = false;
    __syncthreads();
  }

  // dP for this image
  float* dPb = p.dP + (long long)b * k * a;
  for (int i = tid; i < k * a; i += AT_THREADS) {
    if (atomic_out) atomicAdd(dPb + i, dPs[i]);
    else {
      dPb[i] = dPs[i];
      if (p.dP16) p.dP16[((long long)b * k + i / a) * p.a_pad + (i % a)] = __float2bfloat16(dPs[i]);
    }
  }
  // dw_h: reduce the per-warp partials, add the sentinel rows, one atomic per column
#pragma unroll
  for (int jj = 0; jj < MAXJ; ++jj) {
    const int j = lane + jj * 32;
    if (j < a) colred[warp * a + j] = dwh_part[jj];
  }
  __syncthreads();
  if (tid < a) {
    float t = dwh_sent;
#pragma unroll
    for (int w = 0; w < AT_WARPS; ++w) t += colred[w * a + tid];
    atomicAdd(p.dwh + tid, t);
  }
}


// =======================================================================================
// Step-parallel variants (the default when an image's operands fit shared memory).
//
// The kernels above walk the steps of an image one after the other, five block-wide barriers per step, half of the
// threads idle in the context phase and an accurate tanhf per score term: ncu shows them latency-bound (5.7 % and
// 3.8 % of the HBM roofline at B=80, T=18).  Teacher forcing makes every step's inputs (q_t, r_t, s_t, h_t and, in the
// backward, dc_hat_t) available up front, so one 512-thread CTA per image can run each PHASE over all of its steps at
// once: scores for T*(k+1) units with 4 threads per unit, one warp per step for the softmaxes, and the V-sized
// contractions with V_i (resp. dctx_t) loaded once and reused across the steps held in registers.
// =======================================================================================
constexpr int TP_THREADS = 512;
constexpr int TP_WARPS = TP_THREADS / 32;

// tanh from two MUFU ops: 1 - 2 / (1 + 2^(2x log2 e)); abs error ~1e-7, saturates correctly at +-inf
__device__ __forceinline__ float tanh_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

// TT = steps accumulated per thread in the context phase (steps of the CTA = TC <= ngrp * TT)
template <int TT>
__global__ void __launch_bounds__(TP_THREADS) atten_fwd_tpar_kernel(const AttenFwdArgs p, int t_per_cta) {
  extern __shared__ __align__(16) float sm[];
  const int k = p.k, a = p.a, H = p.H, T = p.T;
  const int t_begin = blockIdx.y * t_per_cta;
  const int TC = min(T, t_begin + t_per_cta) - t_begin;
  float* Ps = sm;                          // [k*a]
  float* whs = Ps + k * a;                 // [a]
  float* qs = whs + a;                     // [t_per*a]
  float* rs = qs + t_per_cta * a;          // [t_per*a]
  float* zs = rs + t_per_cta * a;          // [t_per*(k+1)]
  float* als = zs + t_per_cta * (k + 1);   // [t_per*k]
  float* bts = als + t_per_cta * k;        // [t_per]

  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)b * T + t_begin;
  const float* Pb = p.P + (long long)b * k * a;
  for (int i = tid; i < k * a; i += TP_THREADS) Ps[i] = Pb[i];
  for (int j = tid; j < a; j += TP_THREADS) whs[j] = p.wh[j];
  for (int i = tid; i < TC * a; i += TP_THREADS) {       // rows t_begin.. of q and r are contiguous
    qs[i] = p.q[row0 * a + i];
    rs[i] = p.r[row0 * a + i];
  }
  __syncthreads();
  // ---- scores: unit = (step, region) with region k = sentinel; 4 threads per unit over the attention dim ----
  const int units = TC * (k + 1);
  const int sub = tid & 3;
  for (int ub = 0; ub < units; ub += TP_THREADS / 4) {
    const int unit = ub + (tid >> 2);
    float acc = 0.f;
    if (unit < units) {
      const int tl = unit / (k + 1), i = unit - tl * (k + 1);
      if (i < k) {          // z_i = w_h . tanh(P_i + q_t)                              adaptive_attention.py:37-38
        const float* prow = Ps + i * a;
        const float* qrow = qs + tl * a;
#pragma unroll 4
        for (int j = sub; j < a; j += 4) acc = fmaf(whs[j], tanh_mufu(prow[j] + qrow[j]), acc);
      } else {              // z_s = w_h . tanh(r_t)                                    adaptive_attention.py:46-47
        const float* rrow = rs + tl * a;
#pragma unroll 4
        for (int j = sub; j < a; j += 4) acc = fmaf(whs[j], tanh_mufu(rrow[j]), acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0 && unit < units) zs[unit] = acc;
  }
  __syncthreads();
  // ---- softmaxes (adaptive_attention.py:39,51): one warp per step ----
  for (int tl = warp; tl < TC; tl += TP_WARPS) {
    const float* z = zs + tl * (k + 1);
    const long long row = row0 + tl;
    float m = -INFINITY;
    for (int i = lane; i < k; i += 32) m = fmaxf(m, z[i]);
    m = warp_max(m);
    float sum = 0.f;
    for (int i = lane; i < k; i += 32) sum += expf(z[i] - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    const float zsent = z[k];
    const float m1 = fmaxf(m, zsent);
    float sum1 = 0.f;
    for (int i = lane; i < k; i += 32) sum1 += expf(z[i] - m1);
    sum1 = warp_sum(sum1);
    const float es = expf(zsent - m1);
    const float beta = p.no_sentinel ? 0.f : es / (sum1 + es);   // (sentinel-less baseline block: c_hat = ctx)
    for (int i = lane; i < k; i += 32) {
      const float al = expf(z[i] - m) * inv;
      als[tl * k + i] = al;
      p.alpha[row * k + i] = al;
    }
    if (lane == 0) {
      bts[tl] = beta;
      p.beta[row] = beta;
    }
  }
  __syncthreads();
  // ---- context: thread = (float4 column, step group); V_i is loaded once per thread and reused for its TT steps ----
  const int nvec = H / 4;
  const int ngrp = TP_THREADS / nvec;
  const int vec = tid % nvec, tg = tid / nvec;
  if (tg < ngrp) {
    const float* Vb = p.V + (long long)b * k * H + vec * 4;
    float4 acc[TT];
#pragma unroll
    for (int m = 0; m < TT; ++m) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int VB = 7;    // V rows in flight per thread (global-load latency, not bandwidth, bounds this phase)
    for (int i0 = 0; i0 < k; i0 += VB) {
      float4 v[VB];
#pragma unroll
      for (int u = 0; u < VB; ++u) v[u] = i0 + u < k ? ldg4(Vb + (long long)(i0 + u) * H) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < VB; ++u) {
        if (i0 + u < k) {
#pragma unroll
          for (int m = 0; m < TT; ++m) {
            const int tl = tg + ngrp * m;
            if (tl < TC) {
              const float al = als[tl * k + i0 + u];
              acc[m].x = fmaf(al, v[u].x, acc[m].x); acc[m].y = fmaf(al, v[u].y, acc[m].y);
              acc[m].z = fmaf(al, v[u].z, acc[m].z); acc[m].w = fmaf(al, v[u].w, acc[m].w);
            }
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < TT; ++m) {
      const int tl = tg + ngrp * m;
      if (tl < TC) {
        const long long o = (row0 + tl) * H + vec * 4;
        const float beta = bts[tl];
        const float4 sv = *reinterpret_cast<const float4*>(p.s + o);
        const float4 hv = *reinterpret_cast<const float4*>(p.h + o);
        float4 ch;
        ch.x = beta * sv.x + (1.f - beta) * acc[m].x;
        ch.y = beta * sv.y + (1.f - beta) * acc[m].y;
        ch.z = beta * sv.z + (1.f - beta) * acc[m].z;
        ch.w = beta * sv.w + (1.f - beta) * acc[m].w;
        if (p.ctx) *reinterpret_cast<float4*>(p.ctx + o) = acc[m];
        if (p.c_hat) *reinterpret_cast<float4*>(p.c_hat + o) = ch;
        if (p.u) *reinterpret_cast<float4*>(p.u + o) = make_float4(ch.x + hv.x, ch.y + hv.y, ch.z + hv.z, ch.w + hv.w);
        if (p.u16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(ch.x + hv.x, ch.y + hv.y), hi = __floats2bfloat162_rn(ch.z + hv.z, ch.w + hv.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(p.u16 + o) = pk;
        }
      }
    }
  }
}

// Backward, one CTA per image, all T steps of the image inside (no atomics on dV / dP).
//   NV = float4 chunks of 128 columns a lane holds of one V row (H <= 128 * NV), RI = regions per thread in the score
//   backward (k <= (TP_THREADS / a) * RI)
template <int NV, int RI>
__global__ void __launch_bounds__(TP_THREADS) atten_bwd_tpar_kernel(const AttenBwdArgs p, int t_per_cta) {
  extern __shared__ __align__(16) float sm[];
  const int k = p.k, a = p.a, H = p.H;
  // this CTA's steps [t_begin, t_begin + T); with more than one CTA per image dV, dP are accumulated with atomics
  const int t_begin = blockIdx.y * t_per_cta;
  const int T = min(p.T, t_begin + t_per_cta) - t_begin;
  const bool atomic_out = gridDim.y > 1;
  const int TS = t_per_cta;              // shared-memory strides are sized for the largest chunk
  float* dcs = sm;                       // [T*H]   dctx_t = (1-beta_t) dc_hat_t          (16-byte aligned rows)
  float* Ps = dcs + (size_t)TS * H;      // [k*a]
  float* whs = Ps + k * a;               // [a]
  float* qs = whs + a;                   // [T*a]
  float* rs = qs + TS * a;               // [T*a]
  float* als = rs + TS * a;              // [T*k]
  float* dzs = als + TS * k;             // [T*k]   d_alpha, then dz
  float* dqs = dzs + TS * k;             // [T*a]   sum_i dp_i (shared-memory atomics)
  float* drs = dqs + TS * a;             // [T*a]
  float* dwhs = drs + TS * a;            // [a]
  float* bts = dwhs + a;                 // [T]
  float* dbs = bts + TS;                 // [T]     dbeta
  float* dzsent = dbs + TS;              // [T]

  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)b * p.T + t_begin;
  const float* Pb = p.P + (long long)b * k * a;
  for (int i = tid; i < k * a; i += TP_THREADS) Ps[i] = Pb[i];
  for (int j = tid; j < a; j += TP_THREADS) { whs[j] = p.wh[j]; dwhs[j] = 0.f; }
  for (int i = tid; i < T * a; i += TP_THREADS) {
    qs[i] = p.q[row0 * a + i];
    rs[i] = p.r[row0 * a + i];
    dqs[i] = 0.f;
  }
  for (int i = tid; i < T * k; i += TP_THREADS) als[i] = p.alpha[row0 * k + i];
  // ---- (a) beta gate: warp per step.  dbeta = dc_hat . (s - ctx), ds = beta dc_hat, dctx = (1-beta) dc_hat ----
  for (int t = warp; t < T; t += TP_WARPS) {
    const long long row = row0 + t;
    const float beta = p.beta[row];
    float part = 0.f;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 d = *reinterpret_cast<const float4*>(p.dchat + row * H + c);
      const float4 sv = *reinterpret_cast<const float4*>(p.s + row * H + c);
      const float4 cx = *reinterpret_cast<const float4*>(p.ctx + row * H + c);
      part += d.x * (sv.x - cx.x) + d.y * (sv.y - cx.y) + d.z * (sv.z - cx.z) + d.w * (sv.w - cx.w);
      *reinterpret_cast<float4*>(p.ds + row * H + c) = make_float4(beta * d.x, beta * d.y, beta * d.z, beta * d.w);
      const float ob = 1.f - beta;
      *reinterpret_cast<float4*>(dcs + (size_t)t * H + c) = make_float4(ob * d.x, ob * d.y, ob * d.z, ob * d.w);
    }
    part = warp_sum(part);
    if (lane == 0) {
      bts[t] = beta;
      dbs[t] = part + (p.d_beta ? p.d_beta[row] : 0.f);
    }
  }
  __syncthreads();
  // ---- (b) d_alpha[t][i] = dctx_t . V_i: warp per region, V_i held in registers across the steps ----
  const float* Vb = p.V + (long long)b * k * H;
  for (int i = warp; i < k; i += TP_WARPS) {
    float4 v[NV];
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      const int c = lane * 4 + n * 128;
      v[n] = c < H ? ldg4(Vb + (long long)i * H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int t = 0; t < T; ++t) {
      float acc = 0.f;
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const int c = lane * 4 + n * 128;
        if (c < H) {
          const float4 d = *reinterpret_cast<const float4*>(dcs + (size_t)t * H + c);
          acc += v[n].x * d.x + v[n].y * d.y + v[n].z * d.z + v[n].w * d.w;
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) dzs[t * k + i] = acc + (p.d_alpha ? p.d_alpha[(row0 + t) * k + i] : 0.f);
    }
  }
  __syncthreads();
  // ---- (c) the two softmaxes, warp per step: dz_i = alpha_i (dalpha_i - S) - b1 alpha_i, dz_s = b1 ----
  for (int t = warp; t < T; t += TP_WARPS) {
    float S = 0.f;
    for (int i = lane; i < k; i += 32) S += als[t * k + i] * dzs[t * k + i];
    S = warp_sum(S);
    const float beta = bts[t];
    const float b1 = beta * (1.f - beta) * dbs[t];
    for (int i = lane; i < k; i += 32) {
      const float al = als[t * k + i];
      dzs[t * k + i] = al * (dzs[t * k + i] - S) - b1 * al;
    }
    if (lane == 0) dzsent[t] = b1;
  }
  __syncthreads();
  // ---- (d) score backward with tanh recompute: thread = (attention column j, region residue g); dP in registers ----
  {
    const int NG = TP_THREADS / a;
    const int j = tid % a, g = tid / a;
    if (g < NG) {
      const float wj = whs[j];
      float pv[RI], dP[RI];
#pragma unroll
      for (int ii = 0; ii < RI; ++ii) {
        const int i = g + ii * NG;
        pv[ii] = i < k ? Ps[i * a + j] : 0.f;
        dP[ii] = 0.f;
      }
      float dwh = 0.f;
      for (int t = 0; t < T; ++t) {
        const float qv = qs[t * a + j];
        float dq = 0.f;
#pragma unroll
        for (int ii = 0; ii < RI; ++ii) {
          const int i = g + ii * NG;
          if (i < k) {
            const float dz = dzs[t * k + i];
            const float tp = tanh_mufu(pv[ii] + qv);
            const float dp = dz * wj * (1.f - tp * tp);
            dP[ii] += dp;
            dq += dp;
            dwh = fmaf(dz, tp, dwh);
          }
        }
        atomicAdd(&dqs[t * a + j], dq);
        if (g == 0) {       // sentinel row of this step                                adaptive_attention.py:46-47
          const float dzs_t = dzsent[t];
          const float tr = tanh_mufu(rs[t * a + j]);
          drs[t * a + j] = dzs_t * wj * (1.f - tr * tr);
          dwh = fmaf(dzs_t, tr, dwh);
        }
      }
      atomicAdd(&dwhs[j], dwh);
      float* dPb = p.dP + (long long)b * k * a;
#pragma unroll
      for (int ii = 0; ii < RI; ++ii) {
        const int i = g + ii * NG;
        if (i < k) {
          if (atomic_out) atomicAdd(dPb + i * a + j, dP[ii]);
          else {
            dPb[i * a + j] = dP[ii];
            if (p.dP16) p.dP16[((long long)b * k + i) * p.a_pad + j] = __float2bfloat16(dP[ii]);
          }
        }
      }
    }
  }
  // ---- (e) dV_i = sum_t alpha_{t,i} dctx_t: thread = (float4 column, region residue) (needs only als / dcs) ----
  {
    const int nvec = H / 4;
    const int ngrp = TP_THREADS / nvec;
    const int vec = tid % nvec, rg = tid / nvec;
    if (rg < ngrp) {
      float* dVb = p.dV + (long long)b * k * H + vec * 4;
      const float* dc = dcs + vec * 4;
      for (int i = rg; i < k; i += ngrp) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 6
        for (int t = 0; t < T; ++t) {
          const float al = als[t * k + i];
          const float4 d = *reinterpret_cast<const float4*>(dc + (size_t)t * H);
          acc.x = fmaf(al, d.x, acc.x); acc.y = fmaf(al, d.y, acc.y);
          acc.z = fmaf(al, d.z, acc.z); acc.w = fmaf(al, d.w, acc.w);
        }
        if (atomic_out) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dVb + (long long)i * H), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
        else *reinterpret_cast<float4*>(dVb + (long long)i * H) = acc;
      }
    }
  }
  __syncthreads();
  // ---- dq = sum_i dp_i + dr, dr; dw_h ----
  for (int i = tid; i < T * a; i += TP_THREADS) {
    const float drj = drs[i], dqj = dqs[i] + drj;
    const int t = i / a, j = i - t * a;
    p.dr[row0 * a + i] = drj;
    p.dq[row0 * a + i] = dqj;
    if (p.dq16) {
      p.dr16[(row0 + t) * p.a_pad + j] = __float2bfloat16(drj);
      p.dq16[(row0 + t) * p.a_pad + j] = __float2bfloat16(dqj);
    }
  }
  if (tid < a) atomicAdd(p.dwh + tid, dwhs[tid]);
}

size_t fwd_tpar_smem(const AttenFwdArgs& p, int t_per) {
  return sizeof(float) * ((size_t)p.k * p.a + p.a + (size_t)t_per * (2 * p.a + (p.k + 1) + p.k + 1) + 4);
}
size_t bwd_tpar_smem(const AttenBwdArgs& p, int t_per) {
  return sizeof(float) * ((size_t)t_per * p.H + (size_t)p.k * p.a + 2 * p.a + (size_t)t_per * (4 * p.a + 2 * p.k + 3) + 4);
}

template <int TT>
int launch_fwd_tpar(const AttenFwdArgs& p, int t_per, size_t smem, cudaStream_t s) {
  auto kern = atten_fwd_tpar_kernel<TT>;
  static size_t attr = 0;
  if (smem > attr) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  kern<<<dim3(p.B, ceil_div(p.T, t_per)), TP_THREADS, smem, s>>>(p, t_per);
  AA_CHECK_LAUNCH("atten_fwd_tpar");
  return AA_OK;
}

template <int NV, int RI>
int launch_bwd_tpar(const AttenBwdArgs& p, int t_per, size_t smem, cudaStream_t s) {
  auto kern = atten_bwd_tpar_kernel<NV, RI>;
  static size_t attr = 0;
  if (smem > attr) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int ny = ceil_div(p.T, t_per);
  if (ny > 1 && !p.prezeroed) {   // several CTAs per image: dV / dP are accumulated
    AA_CHECK_CUDA(cudaMemsetAsync(p.dV, 0, sizeof(float) * (size_t)p.B * p.k * p.H, s));
    AA_CHECK_CUDA(cudaMemsetAsync(p.dP, 0, sizeof(float) * (size_t)p.B * p.k * p.a, s));
  }
  kern<<<dim3(p.B, ny), TP_THREADS, smem, s>>>(p, t_per);
  AA_CHECK_LAUNCH("atten_bwd_tpar");
  return AA_OK;
}

}  // namespace

int launch_atten_fwd(const AttenFwdArgs& p, cudaStream_t s) {
  AA_REQUIRE(p.H % 4 == 0, "atten_fwd: H must be a multiple of 4 (got %d)", p.H);
  AA_REQUIRE(p.k >= 1 && p.a >= 1, "atten_fwd: bad k/a");
  if (p.B == 0 || p.T == 0) return AA_OK;
  // step-parallel kernel: steps per CTA bounded by the context accumulators (ngrp * 10) and by shared memory
  if (!g_atten_sequential && p.H <= 2048 && TP_THREADS / (p.H / 4) >= 1) {
    const int ngrp = TP_THREADS / (p.H / 4);
    int t_per = p.T < ngrp * 10 ? p.T : ngrp * 10;
    // few images: split the steps over more CTAs, aiming at one full wave of two co-resident CTAs per SM
    {
      const long long slots = 2LL * num_sms();
      int best = t_per;
      double best_cost = 1e30;
      for (int y = 1; y <= p.T; ++y) {
        const int tp = ceil_div(p.T, y);
        if (tp > ngrp * 10) continue;
        const long long ctas = (long long)p.B * ceil_div(p.T, tp);
        const double cost = (double)((ctas + slots - 1) / slots) * (1.5 + tp) * (ctas > num_sms() ? 1.25 : 1.0);
        if (cost < best_cost) { best_cost = cost; best = tp; }
      }
      t_per = best;
    }
    while (t_per > 1 && fwd_tpar_smem(p, t_per) > 160 * 1024) --t_per;
    const size_t smem_tp = fwd_tpar_smem(p, t_per);
    if (smem_tp <= 160 * 1024) {
      const int tt = ceil_div(t_per, ngrp);
      if (tt <= 2) return launch_fwd_tpar<2>(p, t_per, smem_tp, s);
      if (tt <= 5) return launch_fwd_tpar<5>(p, t_per, smem_tp, s);
      return launch_fwd_tpar<10>(p, t_per, smem_tp, s);
    }
  }
  const size_t smem = sizeof(float) * ((size_t)p.k * p.a + 3 * p.a + (p.k + 1) + p.k + 4);
  AA_REQUIRE(smem <= 200 * 1024, "atten_fwd: k*a too large for shared memory (%zu B)", smem);
  static bool attr_done = false;
  if (!attr_done) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(atten_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  // split T so that the grid covers the chip at least ~2x
  const int sms = num_sms();
  int splits = 1;
  while (p.B * splits < 2 * sms && splits < p.T) ++splits;
  const int t_per = ceil_div(p.T, splits);
  dim3 grid(p.B, ceil_div(p.T, t_per));
  atten_fwd_kernel<<<grid, AT_THREADS, smem, s>>>(p, t_per);
  AA_CHECK_LAUNCH("atten_fwd");
  return AA_OK;
}

int launch_atten_bwd(const AttenBwdArgs& p, cudaStream_t s) {
  AA_REQUIRE(p.H % 4 == 0, "atten_bwd: H must be a multiple of 4 (got %d)", p.H);
  AA_REQUIRE(p.a <= 32 * MAXJ, "atten_bwd: attention dim a=%d exceeds %d", p.a, 32 * MAXJ);
  if (p.B == 0 || p.T == 0) return AA_OK;
  // step-parallel kernel: one CTA per image with every step inside (dV / dP written once, no atomics)
  if (!g_atten_sequential && p.H <= 1024 && p.a <= TP_THREADS) {
    // steps per CTA: all of them when the images alone cover the chip, else split (dV / dP then go through atomics);
    // shrink further until the chunk fits shared memory
    int t_per = p.T;
    if ((long long)p.B * 4 < 3LL * num_sms() && p.T > 1) t_per = ceil_div(p.T, (int)((2 * num_sms()) / p.B));   // ~two CTAs per SM
    while (t_per > 1 && bwd_tpar_smem(p, t_per) > 100 * 1024) --t_per;     // (<= 100 KB: two CTAs per SM)
    const int ri = ceil_div(p.k, TP_THREADS / p.a);
    const size_t smem_tp = bwd_tpar_smem(p, t_per);
    if (ri <= 24 && smem_tp <= 200 * 1024) {
      const bool wide = p.H > 512;
      if (ri <= 6) return wide ? launch_bwd_tpar<8, 6>(p, t_per, smem_tp, s) : launch_bwd_tpar<4, 6>(p, t_per, smem_tp, s);
      return wide ? launch_bwd_tpar<8, 24>(p, t_per, smem_tp, s) : launch_bwd_tpar<4, 24>(p, t_per, smem_tp, s);
    }
  }
  static bool attr_done = false;
  const size_t budget = 200 * 1024;
  if (!attr_done) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(atten_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    attr_done = true;
  }
  size_t fixed = (size_t)2 * p.k * p.a + 3 * p.a + p.k + (size_t)AT_WARPS * p.a + 8 + AT_WARPS;
  const int sms = num_sms();
  int splits = 1;
  while (p.B * splits < sms && splits < p.T && splits < 4) ++splits;
  const int t_per = ceil_div(p.T, splits);
  const int ny = ceil_div(p.T, t_per);
  // choose TC = number of steps whose alpha / dctx rows are staged in shared memory at once
  int TC = t_per;
  auto bytes = [&](int tc) { return sizeof(float) * (align_up(fixed + (size_t)tc * p.k, 4) + (size_t)tc * p.H); };
  while (TC > 1 && bytes(TC) > budget) --TC;
  AA_REQUIRE(bytes(TC) <= budget, "atten_bwd: k=%d a=%d H=%d does not fit shared memory", p.k, p.a, p.H);
  const int atomic_out = ny > 1 ? 1 : 0;
  if (atomic_out && !p.prezeroed) {
    AA_CHECK_CUDA(cudaMemsetAsync(p.dV, 0, sizeof(float) * (size_t)p.B * p.k * p.H, s));
    AA_CHECK_CUDA(cudaMemsetAsync(p.dP, 0, sizeof(float) * (size_t)p.B * p.k * p.a, s));
  }
  dim3 grid(p.B, ny);
  atten_bwd_kernel<<<grid, AT_THREADS, bytes(TC), s>>>(p, TC, t_per, atomic_out);
  AA_CHECK_LAUNCH("atten_bwd");
  return AA_OK;
}

}  // namespace aa
