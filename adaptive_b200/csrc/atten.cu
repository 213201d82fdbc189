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
      const float beta = es / (sum1 + es);
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

}  // namespace

int launch_atten_fwd(const AttenFwdArgs& p, cudaStream_t s) {
  AA_REQUIRE(p.H % 4 == 0, "atten_fwd: H must be a multiple of 4 (got %d)", p.H);
  AA_REQUIRE(p.k >= 1 && p.a >= 1, "atten_fwd: bad k/a");
  if (p.B == 0 || p.T == 0) return AA_OK;
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
  if (atomic_out) {
    AA_CHECK_CUDA(cudaMemsetAsync(p.dV, 0, sizeof(float) * (size_t)p.B * p.k * p.H, s));
    AA_CHECK_CUDA(cudaMemsetAsync(p.dP, 0, sizeof(float) * (size_t)p.B * p.k * p.a, s));
  }
  dim3 grid(p.B, ny);
  atten_bwd_kernel<<<grid, AT_THREADS, bytes(TC), s>>>(p, TC, t_per, atomic_out);
  AA_CHECK_LAUNCH("atten_bwd");
  return AA_OK;
}

}  // namespace aa
