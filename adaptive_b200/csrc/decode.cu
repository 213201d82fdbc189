// Fused per-step decode kernel (fp32): LSTM pointwise + sentinel gate + q/r projections +
// scores + tanh + k-way and (k+1)-way softmax + beta-gated context + (c_hat + h).
// One launch per decode step covers what the reference issues as ~40 device kernels
// (Decoder.forward with seq-len 1: baseline_attention.py:167-178, adaptive_attention.py:75-85,
// 26-58, 132).  HBM-bound: per image and step it streams V (k*H*4 B) once plus ~26 KB of
// state; see DESIGN.md for the byte accounting.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace aa {

namespace {

// bf16 mirror of four consecutive values (operand of the single-pass arg-max contraction, vocab_refine.cu)
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* dst, const float4& v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a);
  pk.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(dst) = pk;
}

constexpr int DS_THREADS = 256;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int G = 4;  // images per group: W_g/W_s rows are read once per group

__global__ void __launch_bounds__(DS_THREADS) decode_step_kernel(const DecodeStepArgs p, int ngroups) {
  extern __shared__ __align__(16) float sm[];
  const int k = p.k, a = p.a, H = p.H;
  float* hs = sm;                  // [G*H]  h_t
  float* ss = hs + G * H;          // [G*H]  s_t
  float* qs = ss + G * H;          // [G*a]
  float* rs = qs + G * a;          // [G*a]  r = s W_s^T + q
  float* zs = rs + G * a;          // [G*(k+1)]
  float* als = zs + G * (k + 1);   // [G*k]
  float* whs = als + G * k;        // [a]
  float* bts = whs + a;            // [G]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int j = tid; j < a; j += DS_THREADS) whs[j] = p.wh[j];

  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int b0 = grp * G;
    const int ng = min(G, p.B - b0);
    __syncthreads();  // previous group's smem fully consumed
    // ---- phase 1: LSTM cell + sentinel gate (adaptive_attention.py:79-83 with h~ = 0, Q3) ----
    for (int item = tid; item < ng * (H / 4); item += DS_THREADS) {
      const int g = item / (H / 4), c = (item % (H / 4)) * 4;
      const long long b = b0 + g;
      const float* pre = p.gates + b * 5 * H;
      const float4 pi = *reinterpret_cast<const float4*>(pre + c);
      const float4 pf = *reinterpret_cast<const float4*>(pre + H + c);
      const float4 pg = *reinterpret_cast<const float4*>(pre + 2 * H + c);
      const float4 po = *reinterpret_cast<const float4*>(pre + 3 * H + c);
      const float4 ps = *reinterpret_cast<const float4*>(pre + 4 * H + c);
      const float4 cp = *reinterpret_cast<const float4*>(p.c + b * H + c);
      float4 cn, hn, sn;
#define AA_CELL(X)                                                               \
  {                                                                              \
    const float cc = sigmoidf_acc(pf.X) * cp.X + sigmoidf_acc(pi.X) * tanhf(pg.X); \
    const float tc = tanhf(cc);                                                  \
    cn.X = cc;                                                                   \
    hn.X = sigmoidf_acc(po.X) * tc;                                              \
    sn.X = sigmoidf_acc(ps.X) * tc;                                              \
  }
      AA_CELL(x) AA_CELL(y) AA_CELL(z) AA_CELL(w)
#undef AA_CELL
      *reinterpret_cast<float4*>(p.c + b * H + c) = cn;
      if (p.split) {
        float4 hi, lo;
        split_tf32(hn.x, hi.x, lo.x); split_tf32(hn.y, hi.y, lo.y); split_tf32(hn.z, hi.z, lo.z); split_tf32(hn.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(p.h_out + b * p.ld_h + c) = hi;
        *reinterpret_cast<float4*>(p.h_out + b * p.ld_h + p.h_lo_off + c) = lo;
      } else {
        *reinterpret_cast<float4*>(p.h_out + b * p.ld_h + c) = hn;
      }
      *reinterpret_cast<float4*>(hs + g * H + c) = hn;
      *reinterpret_cast<float4*>(ss + g * H + c) = sn;
    }
    __syncthreads();
    // ---- phase 2: q = h W_g^T, r = s W_s^T (+ q)  -- each weight row read once per group ----
    for (int rrow = warp; rrow < 2 * a; rrow += DS_WARPS) {
      const bool is_s = rrow >= a;
      const int j = is_s ? rrow - a : rrow;
      if (is_s && !p.Ws) {   // baseline model: no sentinel branch (r stays q, beta is forced to 0 below)
        if (lane == 0)
          for (int g = 0; g < G; ++g) rs[g * a + j] = 0.f;
        continue;
      }
      const float* wrow = (is_s ? p.Ws : p.Wg) + (long long)j * H;
      const float* act = is_s ? ss : hs;
      float acc[G];
#pragma unroll
      for (int g = 0; g < G; ++g) acc[g] = 0.f;
      for (int c = lane * 4; c < H; c += 128) {
        const float4 w4 = ldg4(wrow + c);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float4 x4 = *reinterpret_cast<const float4*>(act + g * H + c);
          acc[g] += w4.x * x4.x + w4.y * x4.y + w4.z * x4.z + w4.w * x4.w;
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float v = warp_sum(acc[g]);
        if (lane == 0) (is_s ? rs : qs)[g * a + j] = v;
      }
    }
    __syncthreads();
    for (int i = tid; i < ng * a; i += DS_THREADS) rs[i] += qs[i];
    __syncthreads();
    // ---- phase 3: scores z_i = w_h . tanh(P_i + q), z_s = w_h . tanh(r) ----
    for (int item = warp; item < ng * (k + 1); item += DS_WARPS) {
      const int g = item / (k + 1), i = item % (k + 1);
      float acc = 0.f;
      if (i < k) {
        const float* prow = p.P + (((long long)(b0 + g) / p.beam) * k + i) * a;
        for (int j = lane; j < a; j += 32) acc = fmaf(whs[j], tanhf(__ldg(prow + j) + qs[g * a + j]), acc);
      } else {
        for (int j = lane; j < a; j += 32) acc = fmaf(whs[j], tanhf(rs[g * a + j]), acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) zs[g * (k + 1) + i] = acc;
    }
    __syncthreads();
    // ---- phase 4: softmaxes (adaptive_attention.py:39,51) ----
    if (warp < ng) {
      const int g = warp;
      const long long b = b0 + g;
      const float* z = zs + g * (k + 1);
      float m = -INFINITY;
      for (int i = lane; i < k; i += 32) m = fmaxf(m, z[i]);
      m = warp_max(m);
      float sum = 0.f;
      for (int i = lane; i < k; i += 32) sum += expf(z[i] - m);
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      for (int i = lane; i < k; i += 32) {
        const float al = expf(z[i] - m) * inv;
        als[g * k + i] = al;
        p.alpha[b * p.ld_alpha + i] = al;
      }
      const float zsent = z[k];
      const float m1 = fmaxf(m, zsent);
      float sum1 = 0.f;
      for (int i = lane; i < k; i += 32) sum1 += expf(z[i] - m1);
      sum1 = warp_sum(sum1);
      const float es = expf(zsent - m1);
      const float beta = p.Ws ? es / (sum1 + es) : 0.f;
      if (lane == 0) {
        bts[g] = beta;
        p.beta[b * p.ld_beta] = beta;
      }
    }
    __syncthreads();
    // ---- phase 5: context over V (streamed once), c_hat, u = c_hat + h ----
    for (int item = tid; item < ng * (H / 4); item += DS_THREADS) {
      const int g = item / (H / 4), c = (item % (H / 4)) * 4;
      const long long b = b0 + g;
      const float* vb = p.V + ((b / p.beam) * k) * H + c;
      const float* al = als + g * k;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int i = 0;
      for (; i + 7 <= k; i += 7) {
        float4 v[7];
#pragma unroll
        for (int u = 0; u < 7; ++u) v[u] = ldg4_stream(vb + (long long)(i + u) * H);
#pragma unroll
        for (int u = 0; u < 7; ++u) {
          const float w = al[i + u];
          acc.x = fmaf(w, v[u].x, acc.x); acc.y = fmaf(w, v[u].y, acc.y);
          acc.z = fmaf(w, v[u].z, acc.z); acc.w = fmaf(w, v[u].w, acc.w);
        }
      }
      for (; i < k; ++i) {
        const float4 v = ldg4_stream(vb + (long long)i * H);
        const float w = al[i];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      const float beta = bts[g];
      const float4 sv = *reinterpret_cast<const float4*>(ss + g * H + c);
      const float4 hv = *reinterpret_cast<const float4*>(hs + g * H + c);
      float4 o;
      o.x = beta * sv.x + (1.f - beta) * acc.x + hv.x;
      o.y = beta * sv.y + (1.f - beta) * acc.y + hv.y;
      o.z = beta * sv.z + (1.f - beta) * acc.z + hv.z;
      o.w = beta * sv.w + (1.f - beta) * acc.w + hv.w;
      float* ur = p.u + b * (p.ld_u ? p.ld_u : (long long)H) + c;
      if (p.split) {
        float4 hi, lo;
        split_tf32(o.x, hi.x, lo.x); split_tf32(o.y, hi.y, lo.y); split_tf32(o.z, hi.z, lo.z); split_tf32(o.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(ur) = hi;
        *reinterpret_cast<float4*>(ur + p.u_lo_off) = lo;
      } else {
        *reinterpret_cast<float4*>(ur) = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core decode pipeline (AA_PREC_TF32X3): the per-step work is cut where the data dependencies are
//   gate GEMM -> dec_cell_kernel -> q/r GEMM -> dec_atten_kernel -> vocabulary GEMM (+ arg-max)
// so that W_g / W_s run on tcgen05 like the gate and vocabulary contractions, and the attention kernel is a pure
// stream over V with no block-level synchronisation.
// ---------------------------------------------------------------------------------------------

// LSTM pointwise + sentinel gate (baseline_attention.py:172, adaptive_attention.py:79-83 with h~ = 0, Q3).
// Writes c in place, h and s as tf32 (hi, lo) pairs into the A-operand rows [emb | h | s] and as plain fp32 [h | s].
__global__ void __launch_bounds__(256) dec_cell_kernel(const DecodeCellArgs p) {
  const int H = p.H, H4 = H / 4;
  const long long total = (long long)p.R * H4;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < total; item += (long long)gridDim.x * blockDim.x) {
    const long long r = item / H4;
    const int c = (int)(item % H4) * 4;
    const float* pre = p.gates + r * 5 * H;
    float4 pi = ldg4_stream(pre + c);
    float4 pf = ldg4_stream(pre + H + c);
    float4 pg = ldg4_stream(pre + 2 * H + c);
    float4 po = ldg4_stream(pre + 3 * H + c);
    float4 ps;
    if (p.EG) {
      const long long word = p.prev_ids ? p.prev_ids[r * p.ld_ids] : (long long)p.start_id;
      const float* eg = p.EG + word * 5 * H;
      const float4 ei = ldg4(eg + c), ef = ldg4(eg + H + c), eg4 = ldg4(eg + 2 * H + c), eo = ldg4(eg + 3 * H + c), es = ldg4(eg + 4 * H + c);
      pi.x += ei.x; pi.y += ei.y; pi.z += ei.z; pi.w += ei.w;
      pf.x += ef.x; pf.y += ef.y; pf.z += ef.z; pf.w += ef.w;
      pg.x += eg4.x; pg.y += eg4.y; pg.z += eg4.z; pg.w += eg4.w;
      po.x += eo.x; po.y += eo.y; po.z += eo.z; po.w += eo.w;
      const float4 ss = ldg4_stream(p.stat + r * 5 * H + 4 * H + c);
      ps = make_float4(ss.x + es.x, ss.y + es.y, ss.z + es.z, ss.w + es.w);
    } else {
      ps = ldg4_stream(pre + 4 * H + c);
    }
    const float4 cp = *reinterpret_cast<const float4*>(p.c + r * H + c);
    float4 cn, hn, sn;
#define AA_CELL(X)                                                               \
  {                                                                              \
    const float cc = sigmoidf_acc(pf.X) * cp.X + sigmoidf_acc(pi.X) * tanhf(pg.X); \
    const float tc = tanhf(cc);                                                  \
    cn.X = cc;                                                                   \
    hn.X = sigmoidf_acc(po.X) * tc;                                              \
    sn.X = sigmoidf_acc(ps.X) * tc;                                              \
  }
    AA_CELL(x) AA_CELL(y) AA_CELL(z) AA_CELL(w)
#undef AA_CELL
    *reinterpret_cast<float4*>(p.c + r * H + c) = cn;
    *reinterpret_cast<float4*>(p.hs + r * 2 * H + c) = hn;
    *reinterpret_cast<float4*>(p.hs + r * 2 * H + H + c) = sn;
    float4 hi, lo;
    float* arow = p.A + r * p.ldA;
    split_tf32(hn.x, hi.x, lo.x); split_tf32(hn.y, hi.y, lo.y); split_tf32(hn.z, hi.z, lo.z); split_tf32(hn.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(arow + p.h_off + c) = hi;
    *reinterpret_cast<float4*>(arow + p.lo_off + p.h_off + c) = lo;
    split_tf32(sn.x, hi.x, lo.x); split_tf32(sn.y, hi.y, lo.y); split_tf32(sn.z, hi.z, lo.z); split_tf32(sn.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(arow + p.h_off + H + c) = hi;
    *reinterpret_cast<float4*>(arow + p.lo_off + p.h_off + H + c) = lo;
  }
}

// Scores + tanh + k-way / (k+1)-way softmax + beta-gated context + (c_hat + h): ONE WARP PER ROW, no __syncthreads.
//   z_i = w_h . tanh(P_i + q)  (lanes over the attention dim, coalesced P rows, warp-shuffle sums)
//   alpha, beta: warp-shuffle max / sum
//   ctx = sum_i alpha_i V_i: each lane owns 4*NCH columns, V rows streamed with 128-bit coalesced loads
//   (R rows in flight per lane), read exactly once (L1::no_allocate)
// HBM-bound on V (k*H*4 bytes per image and step).
__device__ __forceinline__ float tanh_mufu(float x);
constexpr int DA_WARPS = 8;
template <int NCH, int RU>
__global__ void __launch_bounds__(DA_WARPS * 32) dec_atten_kernel(const DecodeAttenArgs p) {
  extern __shared__ float da_sm[];
  const int k = p.k, a = p.a, H = p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* zs = da_sm + (size_t)warp * (k + 1);    // this warp's scores, then alphas
  const int j0 = lane, j1 = lane + 32, j2 = lane + 64, j3 = lane + 96;       // a <= 128
  const float w0 = j0 < a ? p.wh[j0] : 0.f, w1 = j1 < a ? p.wh[j1] : 0.f, w2 = j2 < a ? p.wh[j2] : 0.f, w3 = j3 < a ? p.wh[j3] : 0.f;
  for (long long r = (long long)blockIdx.x * DA_WARPS + warp; r < p.R; r += (long long)gridDim.x * DA_WARPS) {
    const long long b = r / p.beam;
    const float* qr = p.qr + r * p.ld_qr;
    const float q0 = j0 < a ? qr[j0] : 0.f, q1 = j1 < a ? qr[j1] : 0.f, q2 = j2 < a ? qr[j2] : 0.f, q3 = j3 < a ? qr[j3] : 0.f;
    // ---- scores ----
    const float* Pb = p.P + b * k * p.ldP;
    for (int i0 = 0; i0 < k; i0 += 4) {
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u] = 0.f;
        if (i0 + u < k) {
          const float* prow = Pb + (long long)(i0 + u) * p.ldP;
          if (j0 < a) acc[u] = w0 * tanh_mufu(__ldg(prow + j0) + q0);
          if (j1 < a) acc[u] = fmaf(w1, tanh_mufu(__ldg(prow + j1) + q1), acc[u]);
          if (j2 < a) acc[u] = fmaf(w2, tanh_mufu(__ldg(prow + j2) + q2), acc[u]);
          if (j3 < a) acc[u] = fmaf(w3, tanh_mufu(__ldg(prow + j3) + q3), acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float z = warp_sum(acc[u]);
        if (lane == u && i0 + u < k) zs[i0 + u] = z;
      }
    }
    {   // sentinel score z_s = w_h . tanh(r)
      float acc = 0.f;
      const float* rr = qr + (p.r_off ? p.r_off : a);
      const float pq = p.r_partial ? 1.f : 0.f;       // r = r' + q when the row holds the sentinel's half alone
      if (j0 < a) acc = w0 * tanh_mufu(rr[j0] + pq * q0);
      if (j1 < a) acc = fmaf(w1, tanh_mufu(rr[j1] + pq * q1), acc);
      if (j2 < a) acc = fmaf(w2, tanh_mufu(rr[j2] + pq * q2), acc);
      if (j3 < a) acc = fmaf(w3, tanh_mufu(rr[j3] + pq * q3), acc);
      acc = warp_sum(acc);
      if (lane == 0) zs[k] = acc;
    }
    __syncwarp();
    // ---- softmaxes (adaptive_attention.py:39,51) ----
    float m = -INFINITY;
    for (int i = lane; i < k; i += 32) m = fmaxf(m, zs[i]);
    m = warp_max(m);
    float sum = 0.f;
    for (int i = lane; i < k; i += 32) sum += expf(zs[i] - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    const float zsent = zs[k];
    const float m1 = fmaxf(m, zsent);
    float sum1 = 0.f;
    for (int i = lane; i < k; i += 32) sum1 += expf(zs[i] - m1);
    sum1 = warp_sum(sum1);
    const float es = expf(zsent - m1);
    const float beta = p.no_sentinel ? 0.f : es / (sum1 + es);
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
      const float al = expf(zs[i] - m) * inv;
      zs[i] = al;
      p.alpha[r * p.ld_alpha + i] = al;
    }
    if (lane == 0) p.beta[r * p.ld_beta] = beta;
    __syncwarp();
    // ---- context over V, c_hat, u = c_hat + h ----
    const float* Vb = p.V + b * k * H;
    float4 acc[NCH];
#pragma unroll
    for (int n = 0; n < NCH; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = 0;
    for (; i + RU <= k; i += RU) {
      float4 v[RU][NCH];
#pragma unroll
      for (int u = 0; u < RU; ++u)
#pragma unroll
        for (int n = 0; n < NCH; ++n) {
          const int c = lane * 4 + n * 128;
          v[u][n] = c < H ? ldg4_stream(Vb + (long long)(i + u) * H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const float al = zs[i + u];
#pragma unroll
        for (int n = 0; n < NCH; ++n) {
          acc[n].x = fmaf(al, v[u][n].x, acc[n].x); acc[n].y = fmaf(al, v[u][n].y, acc[n].y);
          acc[n].z = fmaf(al, v[u][n].z, acc[n].z); acc[n].w = fmaf(al, v[u][n].w, acc[n].w);
        }
      }
    }
    for (; i < k; ++i) {
      const float al = zs[i];
#pragma unroll
      for (int n = 0; n < NCH; ++n) {
        const int c = lane * 4 + n * 128;
        if (c < H) {
          const float4 v = ldg4_stream(Vb + (long long)i * H + c);
          acc[n].x = fmaf(al, v.x, acc[n].x); acc[n].y = fmaf(al, v.y, acc[n].y);
          acc[n].z = fmaf(al, v.z, acc[n].z); acc[n].w = fmaf(al, v.w, acc[n].w);
        }
      }
    }
    const float* hs = p.hs + r * 2 * H;
    float* urow = p.u + r * p.ld_u;
#pragma unroll
    for (int n = 0; n < NCH; ++n) {
      const int c = lane * 4 + n * 128;
      if (c < H) {
        const float4 hv = *reinterpret_cast<const float4*>(hs + c);
        const float4 sv = *reinterpret_cast<const float4*>(hs + H + c);
        float4 o;
        o.x = beta * sv.x + (1.f - beta) * acc[n].x + hv.x;
        o.y = beta * sv.y + (1.f - beta) * acc[n].y + hv.y;
        o.z = beta * sv.z + (1.f - beta) * acc[n].z + hv.z;
        o.w = beta * sv.w + (1.f - beta) * acc[n].w + hv.w;
        float4 hi, lo;
        split_tf32(o.x, hi.x, lo.x); split_tf32(o.y, hi.y, lo.y); split_tf32(o.z, hi.z, lo.z); split_tf32(o.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(urow + c) = hi;
        *reinterpret_cast<float4*>(urow + p.u_lo_off + c) = lo;
        if (p.u16) store_bf16x4(p.u16 + r * p.ld_u16 + c, o);
      }
    }
    __syncwarp();   // zs is reused by this warp's next row
  }
}


// ---------------------------------------------------------------------------------------------
// The same attention step as dec_atten_kernel, restructured as a bulk-copy pipeline (the default whenever the operand
// layout allows 16-byte bulk copies):
//   * one persistent CTA per SM; work item = one image (all of its beam rows, NB at a time, share one pass over V);
//   * a producer warp streams V[b] (k*H*4 contiguous bytes) through an S-stage shared-memory ring with
//     cp.async.bulk + mbarrier complete_tx, L2 evict-first (V is read once per step and is far larger than L2), and
//     the small per-item operands (P[b], the rows' [q | r] and [h | s]) through a 3-slot side buffer.  It runs up to a
//     full ring ahead of the consumers, across item boundaries, so HBM never waits for the score / softmax phases --
//     this is what the register-staged kernel above cannot do (its loads of V only start once alpha is known, and ncu
//     shows it issue-latency bound at ~1 IPC per SM: r01_v7 capture);
//   * two consumer groups of 8 warps take alternate items, so one group's score phase (MUFU) overlaps the other's
//     context phase (LDS + FMA).  Per item: scores (4 threads per (row, region), interleaved over the attention dim,
//     two shuffles), softmaxes (warp per row), context sum_i alpha_i V_i straight out of the ring (thread = one float4
//     column x one region residue class), beta gate and the tf32 (hi, lo) split of u = c_hat + h.
// HBM-bound on V; the V bytes in flight per SM are the ring size (>= 100 KB), not registers x occupancy.
constexpr int DT_CG = 2;                  // consumer groups
constexpr int DT_GT = 256;                // threads per consumer group
constexpr int DT_GW = DT_GT / 32;
constexpr int DT_THREADS = DT_CG * DT_GT + 32;   // + producer warp
constexpr int DT_MAXS = 16;
constexpr int DT_AUX = 3;                 // side-buffer slots: items n, n+1 in work, n+2 loading
constexpr size_t DT_SMEM_BUDGET = 220 * 1024;
constexpr uint32_t DT_STAGE_TARGET = 16 * 1024;

struct DaTmaCfg {
  int S, rps, nchunks, G, nvec;
  uint32_t stage_bytes, aux_stride, p_bytes, qr_off, hs_off;
  uint32_t off_aux, off_wh, off_grp, grp_stride, off_als, off_bts, off_red, off_bar;   // byte offsets (128-byte aligned base)
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)), "l"(policy)
               : "memory");
}
// tanh from two MUFU ops, 5 instructions: 1 - 2 / (1 + 2^(2x log2 e)); abs error ~1e-7, saturates correctly at +-inf
__device__ __forceinline__ float tanh_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
// mbarrier wait without the clock64 guard of tc::mbar_wait (the pipeline's consumers poll often): traps after ~2^28 polls
__device__ __forceinline__ void mbar_wait_lite(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(bar);
  uint32_t done, polls = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++polls > (1u << 28)) __trap();
  } while (!done);
}
__device__ __forceinline__ void group_bar(int cg) { asm volatile("bar.sync %0, %1;" ::"r"(cg + 1), "n"(DT_GT) : "memory"); }

template <int NB>
__global__ void __launch_bounds__(DT_THREADS, 1) dec_atten_tma_kernel(const DecodeAttenArgs p, const DaTmaCfg cf) {
  using namespace tc;
  extern __shared__ uint8_t dt_raw[];
  // (offset arithmetic on the __shared__ array keeps the address space: LDS/STS instead of generic LD/ST)
  uint8_t* sm = dt_raw + ((128u - (smem_u32(dt_raw) & 127u)) & 127u);
  uint8_t* ring = sm;
  uint8_t* auxb = sm + cf.off_aux;
  float* whs = reinterpret_cast<float*>(sm + cf.off_wh);
  uint64_t* fullV = reinterpret_cast<uint64_t*>(sm + cf.off_bar);
  uint64_t* emptyV = fullV + DT_MAXS;
  uint64_t* fullA = emptyV + DT_MAXS;
  uint64_t* emptyA = fullA + DT_AUX;

  const int k = p.k, a = p.a, H = p.H, S = cf.S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int groups = (p.beam + NB - 1) / NB;
  const long long nitems = (long long)(p.R / p.beam) * groups;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&fullV[s], 1); mbar_init(&emptyV[s], DT_GW); }
    for (int s = 0; s < DT_AUX; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], DT_GW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = tid; j < a; j += DT_THREADS) whs[j] = p.wh[j];
  __syncthreads();

  if (warp == DT_CG * DT_GW) {
    // ===== producer =====
    if (lane == 0) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      int s = 0;
      uint32_t ph = 0;
      int n = 0;
      for (long long item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
        const long long b = item / groups;
        const int g = (int)(item % groups);
        const int nrows = min(NB, p.beam - g * NB);
        const long long r0 = b * p.beam + (long long)g * NB;
        const int slot = n % DT_AUX;
        mbar_wait(&emptyA[slot], ((n / DT_AUX) & 1) ^ 1);
        uint8_t* aux = auxb + (size_t)slot * cf.aux_stride;
        const uint32_t qb = (uint32_t)nrows * (uint32_t)p.ld_qr * 4u, hb = (uint32_t)nrows * 2u * (uint32_t)H * 4u;
        mbar_expect_tx(&fullA[slot], cf.p_bytes + qb + hb);
        bulk_g2s(aux, p.P + b * k * p.ldP, cf.p_bytes, &fullA[slot]);
        bulk_g2s(aux + cf.qr_off, p.qr + r0 * p.ld_qr, qb, &fullA[slot]);
        bulk_g2s(aux + cf.hs_off, p.hs + r0 * 2 * H, hb, &fullA[slot]);
        const float* Vb = p.V + b * k * H;
        for (int c = 0; c < cf.nchunks; ++c) {
          const int nreg = min(cf.rps, k - c * cf.rps);
          const uint32_t bytes = (uint32_t)nreg * (uint32_t)H * 4u;
          mbar_wait(&emptyV[s], ph ^ 1);
          mbar_expect_tx(&fullV[s], bytes);
          bulk_g2s_hint(ring + (size_t)s * cf.stage_bytes, Vb + (long long)c * cf.rps * H, bytes, &fullV[s], pol);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
    return;
  }

  // ===== consumers: group cg takes the CTA's items n = cg, cg + DT_CG, ... =====
  const int cg = warp / DT_GW;
  const int gt = tid - cg * DT_GT;            // thread within the group
  const int gw = gt >> 5;                     // warp within the group
  float* zs = reinterpret_cast<float*>(sm + cf.off_grp + (size_t)cg * cf.grp_stride);   // [NB][k+1]
  float* als = reinterpret_cast<float*>(sm + cf.off_grp + (size_t)cg * cf.grp_stride + cf.off_als);   // [NB][k]
  float* bts = reinterpret_cast<float*>(sm + cf.off_grp + (size_t)cg * cf.grp_stride + cf.off_bts);   // [NB]
  float* red = reinterpret_cast<float*>(sm + cf.off_grp + (size_t)cg * cf.grp_stride + cf.off_red);   // [G-1][NB][H]
  const int nvec = cf.nvec, G = cf.G;
  const bool act = gt < G * nvec;
  const int vec = gt % nvec, grp = gt / nvec;
  const int sub = gt & 3;
  const int vcol = vec * 4, GH = G * H, rps_mod = cf.rps % G;
  for (int n = cg;; n += DT_CG) {
    const long long item = blockIdx.x + (long long)n * gridDim.x;
    if (item >= nitems) break;
    const long long b = item / groups;
    const int g = (int)(item % groups);
    const int nrows = min(NB, p.beam - g * NB);
    const long long r0 = b * p.beam + (long long)g * NB;
    const int slot = n % DT_AUX;
    const uint8_t* aux = auxb + (size_t)slot * cf.aux_stride;
    const float* Ps = reinterpret_cast<const float*>(aux);
    const float* qs = reinterpret_cast<const float*>(aux + cf.qr_off);
    const float* hss = reinterpret_cast<const float*>(aux + cf.hs_off);
    mbar_wait_lite(&fullA[slot], (n / DT_AUX) & 1);
    // ---- scores: unit = (row j, region i), i == k the sentinel; 4 threads per unit over the attention dim ----
    const int units = nrows * (k + 1);
    for (int ub = 0; ub < units; ub += DT_GT / 4) {
      const int unit = ub + (gt >> 2);
      float acc = 0.f;
      if (unit < units) {
        const int j = unit / (k + 1), i = unit - j * (k + 1);
        const float* qrow = qs + j * p.ld_qr;
        if (i < k) {          // z_i = w_h . tanh(P_i + q)                                adaptive_attention.py:37-38
          const float* prow = Ps + (long long)i * p.ldP;
#pragma unroll 4
          for (int jj = sub; jj < a; jj += 4) acc = fmaf(whs[jj], tanh_mufu(prow[jj] + qrow[jj]), acc);
        } else {              // z_s = w_h . tanh(r)                                      adaptive_attention.py:46-47
          const float* rrow = qrow + (p.r_off ? p.r_off : a);
          if (p.r_partial) {
#pragma unroll 4
            for (int jj = sub; jj < a; jj += 4) acc = fmaf(whs[jj], tanh_mufu(rrow[jj] + qrow[jj]), acc);
          } else {
#pragma unroll 4
            for (int jj = sub; jj < a; jj += 4) acc = fmaf(whs[jj], tanh_mufu(rrow[jj]), acc);
          }
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (sub == 0 && unit < units) zs[unit] = acc;
    }
    group_bar(cg);
    // ---- softmaxes (adaptive_attention.py:39,51): warp j owns row j ----
    for (int j = gw; j < nrows; j += DT_GW) {
      const float* z = zs + j * (k + 1);
      const long long r = r0 + j;
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
      const float beta = p.no_sentinel ? 0.f : es / (sum1 + es);
      for (int i = lane; i < k; i += 32) {
        const float al = expf(z[i] - m) * inv;
        als[j * k + i] = al;
        p.alpha[r * p.ld_alpha + i] = al;
      }
      if (lane == 0) {
        bts[j] = beta;
        p.beta[r * p.ld_beta] = beta;
      }
    }
    group_bar(cg);
    // ---- context out of the ring ----
    float4 acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int it0 = n * cf.nchunks;
    int s = it0 % S;
    uint32_t ph = (uint32_t)(it0 / S) & 1u;
    const uint8_t* stage = ring + (size_t)s * cf.stage_bytes;
    int resid = 0;                              // (c * rps) % G: this thread takes the regions with global index == grp (mod G)
    for (int c = 0; c < cf.nchunks; ++c) {
      const int i0 = c * cf.rps;
      const int nreg = min(cf.rps, k - i0);
      mbar_wait_lite(&fullV[s], ph);
      if (act) {
        int first = grp - resid;
        if (first < 0) first += G;
        const float* vp = reinterpret_cast<const float*>(stage) + vcol + first * H;
        const float* ap = als + i0 + first;
        for (int i = first; i < nreg; i += G, vp += GH, ap += G) {
          const float4 v = *reinterpret_cast<const float4*>(vp);
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            if (NB == 1 || j < nrows) {
              const float al = ap[j * k];
              acc[j].x = fmaf(al, v.x, acc[j].x); acc[j].y = fmaf(al, v.y, acc[j].y);
              acc[j].z = fmaf(al, v.z, acc[j].z); acc[j].w = fmaf(al, v.w, acc[j].w);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&emptyV[s]);
      resid += rps_mod;
      if (resid >= G) resid -= G;
      stage += cf.stage_bytes;
      if (++s == S) { s = 0; ph ^= 1; stage = ring; }
    }
    if (act && grp > 0) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (j < nrows) *reinterpret_cast<float4*>(red + ((size_t)(grp - 1) * NB + j) * H + vec * 4) = acc[j];
    }
    group_bar(cg);   // partial sums visible; also: every read of zs / als of this item is done
    if (act && grp == 0) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < nrows) {
          float4 t = acc[j];
          for (int gg = 1; gg < G; ++gg) {
            const float4 o = *reinterpret_cast<const float4*>(red + ((size_t)(gg - 1) * NB + j) * H + vec * 4);
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
          }
          const float beta = bts[j];
          const float4 hv = *reinterpret_cast<const float4*>(hss + (size_t)j * 2 * H + vec * 4);
          const float4 sv = *reinterpret_cast<const float4*>(hss + (size_t)j * 2 * H + H + vec * 4);
          float4 o;
          o.x = beta * sv.x + (1.f - beta) * t.x + hv.x;
          o.y = beta * sv.y + (1.f - beta) * t.y + hv.y;
          o.z = beta * sv.z + (1.f - beta) * t.z + hv.z;
          o.w = beta * sv.w + (1.f - beta) * t.w + hv.w;
          float4 hi, lo;
          split_tf32(o.x, hi.x, lo.x); split_tf32(o.y, hi.y, lo.y); split_tf32(o.z, hi.z, lo.z); split_tf32(o.w, hi.w, lo.w);
          float* urow = p.u + (r0 + j) * p.ld_u + vec * 4;
          *reinterpret_cast<float4*>(urow) = hi;
          *reinterpret_cast<float4*>(urow + p.u_lo_off) = lo;
          if (p.u16) store_bf16x4(p.u16 + (r0 + j) * p.ld_u16 + vec * 4, o);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&emptyA[slot]);   // this warp is done with the item's P / qr / hs
  }
}

// shared-memory plan of the pipeline for one shape; false when the shape does not fit (the caller falls back)
bool plan_da_tma(const DecodeAttenArgs& p, int NB, DaTmaCfg& cf, size_t& smem) {
  const int k = p.k, H = p.H;
  const uint32_t row_bytes = (uint32_t)H * 4u;
  int max_rps = (int)(DT_STAGE_TARGET / row_bytes);
  if (max_rps < 1) max_rps = 1;
  if (max_rps > k) max_rps = k;
  cf.nchunks = ceil_div(k, max_rps);
  cf.rps = ceil_div(k, cf.nchunks);
  cf.nchunks = ceil_div(k, cf.rps);
  cf.stage_bytes = (uint32_t)align_up((size_t)cf.rps * row_bytes, 128);
  cf.nvec = H / 4;
  cf.G = DT_GT / cf.nvec;
  if (cf.G < 1) return false;                     // H > 1024
  if (cf.G > k) cf.G = k;
  cf.p_bytes = (uint32_t)k * (uint32_t)p.ldP * 4u;
  cf.qr_off = (uint32_t)align_up(cf.p_bytes, 128);
  cf.hs_off = cf.qr_off + (uint32_t)align_up((size_t)NB * p.ld_qr * 4, 128);
  cf.aux_stride = cf.hs_off + (uint32_t)align_up((size_t)NB * 2 * H * 4, 128);
  const size_t wh_bytes = align_up(sizeof(float) * (size_t)p.a, 16);
  const size_t g_zs = align_up(sizeof(float) * NB * (size_t)(k + 1), 16);
  const size_t g_als = align_up(sizeof(float) * NB * (size_t)k, 16);
  const size_t g_bts = 16 * ((NB + 3) / 4);
  const size_t g_red = sizeof(float) * (size_t)(cf.G - 1) * NB * H;
  const size_t g_all = align_up(g_zs + g_als + g_bts + g_red, 16);
  const size_t bars = sizeof(uint64_t) * (2 * DT_MAXS + 2 * DT_AUX);
  const size_t fixed = DT_AUX * (size_t)cf.aux_stride + wh_bytes + DT_CG * g_all + bars + 128 /* base alignment */;
  if (fixed + 3 * (size_t)cf.stage_bytes > DT_SMEM_BUDGET) return false;
  size_t S = (DT_SMEM_BUDGET - fixed) / cf.stage_bytes;
  if (S > (size_t)DT_MAXS) S = DT_MAXS;
  cf.S = (int)S;
  cf.off_aux = (uint32_t)(S * cf.stage_bytes);
  cf.off_wh = cf.off_aux + DT_AUX * cf.aux_stride;
  cf.off_grp = cf.off_wh + (uint32_t)wh_bytes;
  cf.grp_stride = (uint32_t)g_all;
  cf.off_als = (uint32_t)g_zs;
  cf.off_bts = cf.off_als + (uint32_t)g_als;
  cf.off_red = cf.off_bts + (uint32_t)g_bts;
  cf.off_bar = (uint32_t)align_up(cf.off_grp + DT_CG * g_all, 8);
  smem = cf.off_bar + bars + 128;
  return true;
}

template <int NB>
int launch_da_tma(const DecodeAttenArgs& p, const DaTmaCfg& cf, size_t smem, cudaStream_t s) {
  auto kern = dec_atten_tma_kernel<NB>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int groups = (p.beam + NB - 1) / NB;
  const long long items = (long long)(p.R / p.beam) * groups;
  const int sms = num_sms();
  kern<<<(unsigned)(items < sms ? items : sms), DT_THREADS, smem, s>>>(p, cf);
  AA_CHECK_LAUNCH("dec_atten_tma");
  return AA_OK;
}

template <int NCH, int RU>
int launch_da(const DecodeAttenArgs& p, cudaStream_t s) {
  const size_t smem = sizeof(float) * DA_WARPS * (size_t)(p.k + 1);
  auto kern = dec_atten_kernel<NCH, RU>;
  int occ = 1;
  AA_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DA_WARPS * 32, smem));
  if (occ < 1) occ = 1;
  const long long blocks = (p.R + DA_WARPS - 1) / DA_WARPS;
  const long long cap = (long long)num_sms() * occ;
  kern<<<(unsigned)(blocks < cap ? blocks : cap), DA_WARPS * 32, smem, s>>>(p);
  AA_CHECK_LAUNCH("dec_atten");
  return AA_OK;
}

}  // namespace

int launch_decode_cell(const DecodeCellArgs& p, cudaStream_t s) {
  AA_REQUIRE(p.H % 4 == 0 && p.ldA % 4 == 0 && p.h_off % 4 == 0 && p.lo_off % 4 == 0, "decode_cell: H, ldA and the offsets must be multiples of 4");
  if (p.R == 0) return AA_OK;
  const long long total = (long long)p.R * (p.H / 4);
  const long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  dec_cell_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(p);
  AA_CHECK_LAUNCH("dec_cell");
  return AA_OK;
}

int launch_decode_atten(const DecodeAttenArgs& p, cudaStream_t s) {
  AA_REQUIRE(p.H % 4 == 0 && p.H <= 1024, "decode_atten: H must be a multiple of 4 and <= 1024 (got %d)", p.H);
  AA_REQUIRE(p.a <= 128 && p.k >= 1 && p.k <= 4096, "decode_atten: need a <= 128 and 1 <= k <= 4096");
  AA_REQUIRE(p.beam >= 1 && p.ld_u % 4 == 0 && p.u_lo_off % 4 == 0, "decode_atten: bad beam / u layout");
  AA_REQUIRE(p.ldP >= p.a && p.ld_qr >= (p.r_off ? p.r_off : p.a) + p.a && p.R % p.beam == 0, "decode_atten: bad P / qr strides or R not a multiple of beam");
  if (p.R == 0) return AA_OK;
  // bulk-copy pipeline whenever every streamed operand is 16-byte addressable and the shape fits shared memory
  const bool aligned = p.ldP % 4 == 0 && p.ld_qr % 4 == 0 && (reinterpret_cast<uintptr_t>(p.V) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(p.P) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.qr) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(p.hs) & 15) == 0 && !p.force_simple;
  if (aligned) {
    const int NB = p.beam >= 4 ? 4 : p.beam;
    DaTmaCfg cf{};
    size_t smem = 0;
    if (plan_da_tma(p, NB, cf, smem)) {
      switch (NB) {
        case 1: return launch_da_tma<1>(p, cf, smem, s);
        case 2: return launch_da_tma<2>(p, cf, smem, s);
        case 3: return launch_da_tma<3>(p, cf, smem, s);
        default: return launch_da_tma<4>(p, cf, smem, s);
      }
    }
  }
  if (p.H <= 128) return launch_da<1, 8>(p, s);
  if (p.H <= 256) return launch_da<2, 8>(p, s);
  if (p.H <= 512) return launch_da<4, 4>(p, s);
  return launch_da<8, 2>(p, s);
}

int launch_decode_step(const DecodeStepArgs& p, cudaStream_t s) {
  AA_REQUIRE(p.H % 4 == 0, "decode_step: H must be a multiple of 4 (got %d)", p.H);
  AA_REQUIRE(p.beam >= 1, "decode_step: beam must be >= 1");
  AA_REQUIRE(p.ld_h % 4 == 0 && p.ld_u % 4 == 0 && p.h_lo_off % 4 == 0 && p.u_lo_off % 4 == 0,
             "decode_step: row strides and lo offsets must be multiples of 4");
  if (p.B == 0) return AA_OK;
  const size_t smem = sizeof(float) * ((size_t)2 * G * p.H + 2 * G * p.a + G * (p.k + 1) + G * p.k + p.a + G);
  AA_REQUIRE(smem <= 200 * 1024, "decode_step: H=%d k=%d too large for shared memory", p.H, p.k);
  static bool attr_done = false;
  if (!attr_done) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  const int ngroups = ceil_div(p.B, G);
  int occ = 1;
  AA_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_step_kernel, DS_THREADS, smem));
  if (occ < 1) occ = 1;
  const int grid = ngroups < num_sms() * occ ? ngroups : num_sms() * occ;
  decode_step_kernel<<<grid, DS_THREADS, smem, s>>>(p, ngroups);
  AA_CHECK_LAUNCH("decode_step");
  return AA_OK;
}

}  // namespace aa
