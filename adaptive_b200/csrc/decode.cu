// Fused per-step decode kernel (fp32): LSTM pointwise + sentinel gate + q/r projections +
// scores + tanh + k-way and (k+1)-way softmax + beta-gated context + (c_hat + h).
// One launch per decode step covers what the reference issues as ~40 device kernels
// (Decoder.forward with seq-len 1: baseline_attention.py:167-178, adaptive_attention.py:75-85,
// 26-58, 132).  HBM-bound: per image and step it streams V (k*H*4 B) once plus ~26 KB of
// state; see DESIGN.md for the byte accounting.
#include "kernels.cuh"

namespace aa {

namespace {

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
      *reinterpret_cast<float4*>(p.h_out + b * p.ld_h + c) = hn;
      *reinterpret_cast<float4*>(hs + g * H + c) = hn;
      *reinterpret_cast<float4*>(ss + g * H + c) = sn;
    }
    __syncthreads();
    // ---- phase 2: q = h W_g^T, r = s W_s^T (+ q)  -- each weight row read once per group ----
    for (int rrow = warp; rrow < 2 * a; rrow += DS_WARPS) {
      const bool is_s = rrow >= a;
      const int j = is_s ? rrow - a : rrow;
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
      const float beta = es / (sum1 + es);
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
      *reinterpret_cast<float4*>(p.u + b * H + c) = o;
    }
  }
}

}  // namespace

int launch_decode_step(const DecodeStepArgs& p, cudaStream_t s) {
  AA_REQUIRE(p.H % 4 == 0, "decode_step: H must be a multiple of 4 (got %d)", p.H);
  AA_REQUIRE(p.beam >= 1, "decode_step: beam must be >= 1");
  AA_REQUIRE(p.ld_h % 4 == 0, "decode_step: ld_h must be a multiple of 4");
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
