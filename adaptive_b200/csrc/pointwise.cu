// Element-wise / gather / reduction kernels of the decoder hot path (fp32).
#include "kernels.cuh"

namespace aa {

namespace {

constexpr int PW_THREADS = 256;

typedef __nv_bfloat16 bf16;

__global__ void build_x_kernel(const long long* __restrict__ cap, const float* __restrict__ embed,
                               const float* __restrict__ v_g, float* __restrict__ x, bf16* __restrict__ x16, int B, int T, int E,
                               int Vc, float* __restrict__ zrow32, bf16* __restrict__ zrow16, int H) {
  const int row = blockIdx.x;  // b*T + t
  const int b = row / T;
  if (row == b * T) {   // optional side job of the t = 0 blocks: the h~_0 = 0 rows of the shifted hidden-state arrays [B,T,H]
    if (zrow32) for (int j = threadIdx.x; j < H; j += blockDim.x) zrow32[(long long)row * H + j] = 0.f;
    if (zrow16) for (int j = threadIdx.x; j < H; j += blockDim.x) zrow16[(long long)row * H + j] = __float2bfloat16(0.f);
  }
  long long id = cap[row];
  id = id < 0 ? 0 : (id >= Vc ? Vc - 1 : id);
  const float* src = embed + id * (long long)E;
  const float* vg = v_g + (long long)b * E;
  float* dst = x + (long long)row * 2 * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float a = __ldg(src + e), b2 = __ldg(vg + e);
    dst[e] = a;
    dst[E + e] = b2;
    if (x16) {
      x16[(long long)row * 2 * E + e] = __float2bfloat16(a);
      x16[(long long)row * 2 * E + E + e] = __float2bfloat16(b2);
    }
  }
}

__global__ void lstm_cell_fwd_kernel(const float* __restrict__ pre, long long ld_pre, const float* __restrict__ c_prev,
                                     long long ld_cprev, float* __restrict__ acts, long long ld_acts,
                                     float* __restrict__ c_out, long long ld_c, float* __restrict__ h_out, long long ld_h,
                                     float* __restrict__ hs_next, long long ld_hs, bf16* __restrict__ h16,
                                     bf16* __restrict__ hs_next16, int B, int H) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  const float* p = pre + b * ld_pre;
  const float ig = sigmoidf_acc(p[j]);
  const float fg = sigmoidf_acc(p[H + j]);
  const float gg = tanhf(p[2 * H + j]);
  const float og = sigmoidf_acc(p[3 * H + j]);
  const float c = fg * c_prev[b * ld_cprev + j] + ig * gg;
  const float h = og * tanhf(c);
  float* a = acts + b * ld_acts;
  a[j] = ig; a[H + j] = fg; a[2 * H + j] = gg; a[3 * H + j] = og;
  c_out[b * ld_c + j] = c;
  h_out[b * ld_h + j] = h;
  if (hs_next) hs_next[b * ld_hs + j] = h;
  if (h16) h16[b * ld_h + j] = __float2bfloat16(h);               // bf16 mirrors share the fp32 strides
  if (hs_next16) hs_next16[b * ld_hs + j] = __float2bfloat16(h);
}

// pre and g may alias (in-place gate)
__global__ void sentinel_fwd_kernel(const float* pre, const float* __restrict__ cells, float* g, float* __restrict__ s_out,
                                    bf16* __restrict__ s16, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gv = sigmoidf_acc(pre[i]);
  g[i] = gv;
  const float sv = gv * tanhf(cells[i]);
  s_out[i] = sv;
  if (s16) s16[i] = __float2bfloat16(sv);
}

__global__ void sentinel_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ g, const float* __restrict__ cells,
                                    float* __restrict__ da, float* __restrict__ dcell, bf16* __restrict__ da16, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = ds[i], gv = g[i], tc = tanhf(cells[i]);
  const float dav = d * tc * gv * (1.f - gv);
  da[i] = dav;
  if (da16) da16[i] = __float2bfloat16(dav);
  dcell[i] = d * gv * (1.f - tc * tc);
}

__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dh_attn, long long ld_dh, const float* __restrict__ dhs_next,
                                     long long ld_dhs, const float* __restrict__ dh_rec, const float* __restrict__ dcell,
                                     long long ld_dcell, const float* dc_rec, const float* __restrict__ acts,
                                     long long ld_acts, const float* __restrict__ cells, long long ld_c,
                                     const float* __restrict__ c_prev, long long ld_cprev, float* __restrict__ dgates,
                                     long long ld_dg, bf16* __restrict__ dgates16, float* dc_out, int B, int H) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  float dh = dh_attn[b * ld_dh + j];
  if (dhs_next) dh += dhs_next[b * ld_dhs + j];
  if (dh_rec) dh += dh_rec[(long long)b * H + j];
  const float* a = acts + b * ld_acts;
  const float ig = a[j], fg = a[H + j], gg = a[2 * H + j], og = a[3 * H + j];
  const float tc = tanhf(cells[b * ld_c + j]);
  float dc = dcell[b * ld_dcell + j] + dh * og * (1.f - tc * tc);
  if (dc_rec) dc += dc_rec[(long long)b * H + j];
  float* dg = dgates + b * ld_dg;
  const float d0 = dc * gg * ig * (1.f - ig), d1 = dc * c_prev[b * ld_cprev + j] * fg * (1.f - fg);
  const float d2 = dc * ig * (1.f - gg * gg), d3 = dh * tc * og * (1.f - og);
  dg[j] = d0; dg[H + j] = d1; dg[2 * H + j] = d2; dg[3 * H + j] = d3;
  if (dgates16) {
    bf16* dg16 = dgates16 + b * ld_dg;
    dg16[j] = __float2bfloat16(d0); dg16[H + j] = __float2bfloat16(d1);
    dg16[2 * H + j] = __float2bfloat16(d2); dg16[3 * H + j] = __float2bfloat16(d3);
  }
  dc_out[(long long)b * H + j] = dc * fg;
}

__global__ void embed_scatter_kernel(const long long* __restrict__ cap, const float* __restrict__ dx, float* __restrict__ dE,
                                     int E, int Vc) {
  const int row = blockIdx.x;
  long long id = cap[row];
  id = id < 0 ? 0 : (id >= Vc ? Vc - 1 : id);
  const float* src = dx + (long long)row * 2 * E;
  float* dst = dE + id * (long long)E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(dst + e, src[e]);
}

__global__ void dvg_kernel(const float* __restrict__ dx, float* __restrict__ dvg, int B, int T, int E) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * E) return;
  const int b = (int)(idx / E), e = (int)(idx % E);
  const float* p = dx + ((long long)b * T) * 2 * E + E + e;
  float acc = 0.f;
  for (int t = 0; t < T; ++t) acc += p[(long long)t * 2 * E];
  dvg[idx] = acc;
}

__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

__global__ void copy2d_kernel(float* __restrict__ dst, long long ld_dst, const float* __restrict__ src, long long ld_src,
                              int rows, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int r = (int)(idx / cols), c = (int)(idx % cols);
  dst[r * ld_dst + c] = src[r * ld_src + c];
}

__global__ void gather_rows_kernel(const long long* __restrict__ ids, long long ld_ids, const float* __restrict__ table, int E,
                                   int Vc, float* __restrict__ dst, long long ld_dst) {
  const int b = blockIdx.x;
  long long id = ids[b * ld_ids];
  id = id < 0 ? 0 : (id >= Vc ? Vc - 1 : id);
  for (int e = threadIdx.x; e < E; e += blockDim.x) dst[b * ld_dst + e] = __ldg(table + id * E + e);
}

// One CTA per row.  Lowest index wins ties (torch.max on CPU).
__global__ void __launch_bounds__(PW_THREADS) argmax_gather_kernel(const float* __restrict__ logits, long long ld_logits,
                                                                   int Vc, long long* __restrict__ ids_out, long long ld_ids,
                                                                   const float* __restrict__ embed, int E,
                                                                   float* __restrict__ emb_dst, long long ld_emb) {
  __shared__ float sv[PW_THREADS / 32];
  __shared__ int si[PW_THREADS / 32];
  __shared__ int winner;
  const int b = blockIdx.x;
  const float* row = logits + b * ld_logits;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int n = threadIdx.x; n < Vc; n += blockDim.x) {
    const float v = row[n];
    if (v > best || (v == best && n < bi)) { best = v; bi = n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sv[w] = best; si[w] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < PW_THREADS / 32; ++i)
      if (sv[i] > best || (sv[i] == best && si[i] < bi)) { best = sv[i]; bi = si[i]; }
    if (bi == 0x7fffffff) bi = 0;  // all-NaN row
    winner = bi;
    ids_out[b * ld_ids] = bi;
  }
  __syncthreads();
  if (emb_dst) {
    const float* src = embed + (long long)winner * E;
    for (int e = threadIdx.x; e < E; e += blockDim.x) emb_dst[b * ld_emb + e] = __ldg(src + e);
  }
}

// Final reduction of the vocabulary GEMM's per-(row, n-tile) arg-max partials (gemm_tc.cu epilogue) + gather of the
// winner's embedding into the next step's A operand.  One warp per row; partial j covers columns [j*tile_n, ...), so
// comparing (value, index) with "greater, or equal and lower index" keeps torch.max's lowest-index tie-break (Q12).
// emb_dst row layout: split != 0 -> hi at [0,E), lo at [lo_off, lo_off+E); else plain fp32 at [0,E).
__global__ void __launch_bounds__(PW_THREADS) argmax_finalize_kernel(const float* __restrict__ pmax, const int* __restrict__ pidx,
                                                                     int tiles_n, int B, long long* __restrict__ ids_out,
                                                                     long long ld_ids, const float* __restrict__ embed, int E,
                                                                     float* __restrict__ emb_dst, long long ld_emb, int split,
                                                                     long long lo_off, const int* __restrict__ ncand) {
  const int b = blockIdx.x * (PW_THREADS / 32) + (threadIdx.x >> 5);
  const int l = threadIdx.x & 31;
  if (b >= B) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const int nv = ncand ? ncand[b] : tiles_n;
  for (int j = l; j < nv; j += 32) {
    const float v = pmax[(long long)b * tiles_n + j];
    const int i = pidx[(long long)b * tiles_n + j];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;  // all-NaN row
  if (l == 0) ids_out[b * ld_ids] = bi;
  if (emb_dst) {
    const float* src = embed + (long long)bi * E;
    float* dst = emb_dst + b * ld_emb;
    for (int e = l; e < E; e += 32) {
      const float x = __ldg(src + e);
      if (split) {
        float hi, lo;
        split_tf32(x, hi, lo);
        dst[e] = hi;
        dst[lo_off + e] = lo;
      } else {
        dst[e] = x;
      }
    }
  }
}

// dst[r, 0:Kp] = tf32 hi of src[r, 0:cols] (zero padded), dst[r, Kp:2Kp] = lo
__global__ void __launch_bounds__(PW_THREADS) split_tf32_kernel(const float* __restrict__ src, long long ld_src, long long rows, int cols,
                                                                float* __restrict__ dst, int Kp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Kp) return;
  const long long r = idx / Kp;
  const int c = (int)(idx % Kp);
  float hi = 0.f, lo = 0.f;
  if (c < cols) split_tf32(src[r * ld_src + c], hi, lo);
  dst[r * 2 * Kp + c] = hi;
  dst[r * 2 * Kp + Kp + c] = lo;
}

__global__ void __launch_bounds__(PW_THREADS) ce_fwd_bwd_kernel(const float* __restrict__ logits, long long ld,
                                                                const long long* __restrict__ tgt, int n, int Vc,
                                                                float* __restrict__ loss, float* __restrict__ dlogits,
                                                                long long ldd, float inv_n) {
  __shared__ float red[PW_THREADS / 32];
  __shared__ float bcast;
  const int r = blockIdx.x;
  const float* row = logits + r * ld;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < Vc; i += blockDim.x) m = fmaxf(m, row[i]);
  m = warp_max(m);
  if (l == 0) red[w] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < PW_THREADS / 32; ++i) t = fmaxf(t, red[i]);
    bcast = t;
  }
  __syncthreads();
  m = bcast;
  float sum = 0.f;
  for (int i = threadIdx.x; i < Vc; i += blockDim.x) sum += expf(row[i] - m);
  sum = warp_sum(sum);
  __syncthreads();
  if (l == 0) red[w] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < PW_THREADS / 32; ++i) t += red[i];
    bcast = t;
  }
  __syncthreads();
  sum = bcast;
  const long long t = tgt[r];
  if (threadIdx.x == 0) atomicAdd(loss, (logf(sum) + m - row[t]) * inv_n);
  if (dlogits) {
    const float inv = 1.f / sum;
    float* d = dlogits + r * ldd;
    for (int i = threadIdx.x; i < Vc; i += blockDim.x) {
      float p = expf(row[i] - m) * inv;
      if (i == t) p -= 1.f;
      d[i] = p * inv_n;
    }
  }
}

// Same result with the row held in registers: one read of the logits, one exp per element, one write of the gradient
// (the kernel above reads every row three times and evaluates exp twice per element: 36.7 us for 840 x 10000 inside the
// training step, profiles/r01_v42_timeline.txt; 67 MB of traffic is ~11 us of HBM time).  Needs Vc % 4 == 0, 16-byte
// aligned rows and Vc <= 4 * NT * Q4.
template <int NT, int Q4>
__global__ void __launch_bounds__(NT) ce_fwd_bwd_reg_kernel(const float* __restrict__ logits, long long ld,
                                                            const long long* __restrict__ tgt, int Vc, float* __restrict__ loss,
                                                            float* __restrict__ dlogits, long long ldd, float inv_n,
                                                            bf16* __restrict__ dlogits16) {
  __shared__ float red[NT / 32];
  __shared__ float bcast[2];
  const int r = blockIdx.x;
  const float4* row = reinterpret_cast<const float4*>(logits + r * ld);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int n4 = Vc >> 2;
  float4 v[Q4];
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < Q4; ++q) {
    const int i = q * NT + threadIdx.x;
    v[q] = i < n4 ? ldg4_stream(reinterpret_cast<const float*>(row + i)) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    m = fmaxf(fmaxf(m, fmaxf(v[q].x, v[q].y)), fmaxf(v[q].z, v[q].w));
  }
  m = warp_max(m);
  if (l == 0) red[w] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < NT / 32; ++i) t = fmaxf(t, red[i]);
    bcast[0] = t;
  }
  __syncthreads();
  m = bcast[0];
  const long long t = tgt[r];
  float sum = 0.f, xt = 0.f;
#pragma unroll
  for (int q = 0; q < Q4; ++q) {
    const int i = q * NT + threadIdx.x;
    if (i == (int)(t >> 2)) xt = (t & 3) == 0 ? v[q].x : (t & 3) == 1 ? v[q].y : (t & 3) == 2 ? v[q].z : v[q].w;
    v[q].x = expf(v[q].x - m); v[q].y = expf(v[q].y - m); v[q].z = expf(v[q].z - m); v[q].w = expf(v[q].w - m);   // (exp(-inf) = 0 past the row)
    sum += (v[q].x + v[q].y) + (v[q].z + v[q].w);
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (l == 0) red[w] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s2 = 0.f;
    for (int i = 0; i < NT / 32; ++i) s2 += red[i];
    bcast[1] = s2;
  }
  __syncthreads();
  sum = bcast[1];
  if ((int)(t >> 2) % NT == (int)threadIdx.x && (t >> 2) < n4) atomicAdd(loss, (logf(sum) + m - xt) * inv_n);   // the thread that held x_t
  if (dlogits) {
    const float sc = inv_n / sum;
    float4* d = reinterpret_cast<float4*>(dlogits + r * ldd);
#pragma unroll
    for (int q = 0; q < Q4; ++q) {
      const int i = q * NT + threadIdx.x;
      if (i < n4) {
        float4 p = make_float4(v[q].x * sc, v[q].y * sc, v[q].z * sc, v[q].w * sc);
        if (i == (int)(t >> 2)) {
          if ((t & 3) == 0) p.x -= inv_n; else if ((t & 3) == 1) p.y -= inv_n; else if ((t & 3) == 2) p.z -= inv_n; else p.w -= inv_n;
        }
        d[i] = p;
        if (dlogits16) {   // bf16 mirror of the gradient (operand of the vocabulary projection's backward), same row stride
          __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          reinterpret_cast<uint2*>(dlogits16 + r * ldd)[i] = pk;
        }
      }
    }
  }
}

// x *= *g unless *g == 1 (the upstream gradient of a loss that is the root of the backward pass): every CTA reads the device
// scalar and leaves at once in the common case instead of a full read-modify-write pass over x
__global__ void scale_unless_one_kernel(float* __restrict__ x, const float* __restrict__ g, long long n, bf16* __restrict__ x16) {
  const float s = *g;
  if (s == 1.0f) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i] * s;
    x[i] = v;
    if (x16) x16[i] = __float2bfloat16(v);     // (keeps an optional bf16 mirror consistent)
  }
}

// up to 8 device-to-device copies in one launch (the step's input tensors into a CUDA graph's static buffers)
struct CopySegs {
  const void* src[8];
  void* dst[8];
  long long bytes[8];
};
__global__ void copy_multi_kernel(const CopySegs segs) {
  const int sidx = blockIdx.y;
  const char* src = static_cast<const char*>(segs.src[sidx]);
  char* dst = static_cast<char*>(segs.dst[sidx]);
  const long long n = segs.bytes[sidx];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const long long n16 = n >> 4;
    for (long long i = tid; i < n16; i += nth) reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    for (long long i = (n16 << 4) + tid; i < n; i += nth) dst[i] = src[i];
  } else {
    for (long long i = tid; i < n; i += nth) dst[i] = src[i];
  }
}

// fp32 -> bf16, 2-D with independent row strides (also used for the zero-padded a -> a_pad copies)
__global__ void cast2d_kernel(const float* __restrict__ src, long long ld_src, bf16* __restrict__ dst, long long ld_dst, long long rows,
                              int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols;
  const int c = (int)(idx % cols);
  dst[r * ld_dst + c] = __float2bfloat16(src[r * ld_src + c]);
}

// out[n] += sum_m X[m,n] over this block's row chunk (out pre-zeroed); optionally writes the bf16 mirror of X.
// block (32,8): 128 columns x RB rows per block, float4 per thread.
constexpr int CS_RB = 64;
__global__ void colsum_cast_kernel(const float* __restrict__ X, long long ldx, int M, int N, float* __restrict__ out,
                                   float* __restrict__ out2, bf16* __restrict__ X16, long long ld16) {
  __shared__ float4 red[8][32];
  const int c = blockIdx.x * 128 + threadIdx.x * 4;
  const int r_begin = blockIdx.y * CS_RB, r_end = min(M, r_begin + CS_RB);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < N) {
    const bool vec = (c + 3 < N) && ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    for (int m = r_begin + threadIdx.y; m < r_end; m += 8) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* p = X + (long long)m * ldx + c;
      if (vec) v = *reinterpret_cast<const float4*>(p);
      else { v.x = p[0]; if (c + 1 < N) v.y = p[1]; if (c + 2 < N) v.z = p[2]; if (c + 3 < N) v.w = p[3]; }
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      if (X16) {
        bf16* q = X16 + (long long)m * ld16 + c;
        if (vec && ((ld16 & 3) == 0) && ((reinterpret_cast<uintptr_t>(X16) & 7) == 0)) {
          __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&a);
          pk.y = *reinterpret_cast<uint32_t*>(&b);
          *reinterpret_cast<uint2*>(q) = pk;
        } else {
          q[0] = __float2bfloat16(v.x);
          if (c + 1 < N) q[1] = __float2bfloat16(v.y);
          if (c + 2 < N) q[2] = __float2bfloat16(v.z);
          if (c + 3 < N) q[3] = __float2bfloat16(v.w);
        }
      }
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float4 t = red[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < 8; ++y) {
      const float4 u = red[y][threadIdx.x];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    const float tv[4] = {t.x, t.y, t.z, t.w};
    for (int j = 0; j < 4 && c + j < N; ++j) {
      atomicAdd(out + c + j, tv[j]);
      if (out2) atomicAdd(out2 + c + j, tv[j]);
    }
  }
}

__global__ void cast_multi_kernel(const CastSegs segs) {
  const int sidx = blockIdx.y;
  const float* __restrict__ src = segs.src[sidx];
  bf16* __restrict__ dst = segs.dst[sidx];
  const long long n = segs.n[sidx];
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(dst)[i] = pk;
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16(src[i]);
}

inline unsigned blocks_for(long long n) { return (unsigned)((n + PW_THREADS - 1) / PW_THREADS); }

}  // namespace

int launch_build_x(const long long* cap, const float* embed, const float* v_g, float* x, __nv_bfloat16* x16, int B, int T, int E,
                   int Vc, cudaStream_t s, float* zrow32, __nv_bfloat16* zrow16, int H) {
  if (B * T == 0) return AA_OK;
  build_x_kernel<<<B * T, 128, 0, s>>>(cap, embed, v_g, x, x16, B, T, E, Vc, zrow32, zrow16, H);
  AA_CHECK_LAUNCH("build_x");
  return AA_OK;
}

int launch_lstm_cell_fwd(const float* pre, long long ld_pre, const float* c_prev, long long ld_cprev, float* acts,
                         long long ld_acts, float* c_out, long long ld_c, float* h_out, long long ld_h, float* hs_next,
                         long long ld_hs, __nv_bfloat16* h16, __nv_bfloat16* hs_next16, int B, int H, cudaStream_t s) {
  lstm_cell_fwd_kernel<<<blocks_for((long long)B * H), PW_THREADS, 0, s>>>(pre, ld_pre, c_prev, ld_cprev, acts, ld_acts, c_out,
                                                                           ld_c, h_out, ld_h, hs_next, ld_hs, h16, hs_next16, B, H);
  AA_CHECK_LAUNCH("lstm_cell_fwd");
  return AA_OK;
}

int launch_sentinel_fwd(const float* pre, const float* cells, float* g, float* s_out, __nv_bfloat16* s16, long long n,
                        cudaStream_t s) {
  sentinel_fwd_kernel<<<blocks_for(n), PW_THREADS, 0, s>>>(pre, cells, g, s_out, s16, n);
  AA_CHECK_LAUNCH("sentinel_fwd");
  return AA_OK;
}

int launch_sentinel_bwd(const float* ds, const float* g, const float* cells, float* da, float* dcell, __nv_bfloat16* da16,
                        long long n, cudaStream_t s) {
  sentinel_bwd_kernel<<<blocks_for(n), PW_THREADS, 0, s>>>(ds, g, cells, da, dcell, da16, n);
  AA_CHECK_LAUNCH("sentinel_bwd");
  return AA_OK;
}

int launch_lstm_cell_bwd(const float* dh_attn, long long ld_dh, const float* dhs_next, long long ld_dhs,
                         const float* dh_rec, const float* dcell, long long ld_dcell, const float* dc_rec,
                         const float* acts, long long ld_acts, const float* cells, long long ld_c, const float* c_prev,
                         long long ld_cprev, float* dgates, long long ld_dg, __nv_bfloat16* dgates16, float* dc_out, int B,
                         int H, cudaStream_t s) {
  lstm_cell_bwd_kernel<<<blocks_for((long long)B * H), PW_THREADS, 0, s>>>(dh_attn, ld_dh, dhs_next, ld_dhs, dh_rec, dcell,
                                                                           ld_dcell, dc_rec, acts, ld_acts, cells, ld_c,
                                                                           c_prev, ld_cprev, dgates, ld_dg, dgates16, dc_out, B, H);
  AA_CHECK_LAUNCH("lstm_cell_bwd");
  return AA_OK;
}

int launch_colsum(const float* X, long long ldx, int M, int N, float* out, float* out2, cudaStream_t s) {
  return launch_colsum_cast(X, ldx, M, N, out, out2, nullptr, 0, s);
}

int launch_colsum_cast(const float* X, long long ldx, int M, int N, float* out, float* out2, __nv_bfloat16* X16, long long ld16,
                       cudaStream_t s) {
  if (N == 0) return AA_OK;
  AA_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, s));
  if (out2) AA_CHECK_CUDA(cudaMemsetAsync(out2, 0, sizeof(float) * N, s));
  if (M == 0) return AA_OK;
  colsum_cast_kernel<<<dim3(ceil_div(N, 128), ceil_div(M, CS_RB)), dim3(32, 8), 0, s>>>(X, ldx, M, N, out, out2, X16, ld16);
  AA_CHECK_LAUNCH("colsum_cast");
  return AA_OK;
}

int launch_embed_bwd(const long long* cap, const float* dx, float* dE, float* dvg, int B, int T, int E, int Vc,
                     cudaStream_t s) {
  if (dE) {
    embed_scatter_kernel<<<B * T, 128, 0, s>>>(cap, dx, dE, E, Vc);
    AA_CHECK_LAUNCH("embed_scatter");
  }
  if (dvg) {
    dvg_kernel<<<blocks_for((long long)B * E), PW_THREADS, 0, s>>>(dx, dvg, B, T, E);
    AA_CHECK_LAUNCH("dvg");
  }
  return AA_OK;
}

int launch_add_inplace(float* y, const float* x, long long n, cudaStream_t s) {
  add_inplace_kernel<<<blocks_for(n), PW_THREADS, 0, s>>>(y, x, n);
  AA_CHECK_LAUNCH("add_inplace");
  return AA_OK;
}

int launch_argmax_gather(const float* logits, long long ld_logits, int B, int Vc, long long* ids_out, long long ld_ids,
                         const float* embed, int E, float* emb_dst, long long ld_emb, cudaStream_t s) {
  argmax_gather_kernel<<<B, PW_THREADS, 0, s>>>(logits, ld_logits, Vc, ids_out, ld_ids, embed, E, emb_dst, ld_emb);
  AA_CHECK_LAUNCH("argmax_gather");
  return AA_OK;
}

int launch_argmax_finalize(const float* pmax, const int* pidx, int tiles_n, int B, long long* ids_out, long long ld_ids,
                           const float* embed, int E, float* emb_dst, long long ld_emb, int split, long long lo_off, cudaStream_t s,
                           const int* ncand) {
  if (B == 0) return AA_OK;
  argmax_finalize_kernel<<<ceil_div(B, PW_THREADS / 32), PW_THREADS, 0, s>>>(pmax, pidx, tiles_n, B, ids_out, ld_ids, embed, E, emb_dst,
                                                                             ld_emb, split, lo_off, ncand);
  AA_CHECK_LAUNCH("argmax_finalize");
  return AA_OK;
}

int launch_split_tf32(const float* src, long long ld_src, long long rows, int cols, float* dst, int Kp, cudaStream_t s) {
  if (rows == 0) return AA_OK;
  AA_REQUIRE(Kp >= cols && Kp % 32 == 0, "split_tf32: Kp=%d must be a multiple of 32 and >= cols=%d", Kp, cols);
  split_tf32_kernel<<<blocks_for(rows * Kp), PW_THREADS, 0, s>>>(src, ld_src, rows, cols, dst, Kp);
  AA_CHECK_LAUNCH("split_tf32");
  return AA_OK;
}

int launch_gather_rows(const long long* ids, long long ld_ids, const float* table, int E, int Vc, float* dst, long long ld_dst,
                       int B, cudaStream_t s) {
  gather_rows_kernel<<<B, 128, 0, s>>>(ids, ld_ids, table, E, Vc, dst, ld_dst);
  AA_CHECK_LAUNCH("gather_rows");
  return AA_OK;
}

int launch_copy2d(float* dst, long long ld_dst, const float* src, long long ld_src, int rows, int cols, cudaStream_t s) {
  copy2d_kernel<<<blocks_for((long long)rows * cols), PW_THREADS, 0, s>>>(dst, ld_dst, src, ld_src, rows, cols);
  AA_CHECK_LAUNCH("copy2d");
  return AA_OK;
}

int launch_cast2d(const float* src, long long ld_src, __nv_bfloat16* dst, long long ld_dst, long long rows, int cols,
                  cudaStream_t s) {
  if (rows * cols == 0) return AA_OK;
  if (ld_src == cols && ld_dst == cols) {   // contiguous: vectorised 1-D path
    CastSegs cs{};
    cs.src[0] = src; cs.dst[0] = dst; cs.n[0] = rows * cols;
    return launch_cast_multi(cs, 1, s);
  }
  cast2d_kernel<<<blocks_for(rows * cols), PW_THREADS, 0, s>>>(src, ld_src, dst, ld_dst, rows, cols);
  AA_CHECK_LAUNCH("cast2d");
  return AA_OK;
}

int launch_cast_multi(const CastSegs& segs, int nsegs, cudaStream_t s, int max_blocks_per_seg) {
  if (nsegs == 0) return AA_OK;
  long long nmax = 0;
  for (int i = 0; i < nsegs; ++i) nmax = segs.n[i] > nmax ? segs.n[i] : nmax;
  long long nb = (nmax / 4 + PW_THREADS * 4 - 1) / (PW_THREADS * 4);   // ~4 float4 per thread
  const long long cap = max_blocks_per_seg > 0 ? max_blocks_per_seg : 1184;
  nb = nb < 1 ? 1 : (nb > cap ? cap : nb);
  cast_multi_kernel<<<dim3((unsigned)nb, nsegs), PW_THREADS, 0, s>>>(segs);
  AA_CHECK_LAUNCH("cast_multi");
  return AA_OK;
}

int launch_ce_fwd_bwd(const float* logits, long long ld, const long long* tgt, int n, int Vc, float* loss, float* dlogits,
                      long long ldd, long long denom, cudaStream_t s, __nv_bfloat16* dlogits16, int* mirror_written) {
  if (mirror_written) *mirror_written = 0;
  if (n == 0) return AA_OK;
  const float inv_n = 1.f / (float)(denom > 0 ? denom : n);
  const bool vec = Vc % 4 == 0 && ld % 4 == 0 && ldd % 4 == 0 && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0;
  const int n4 = Vc / 4;
  bf16* m16 = (vec && dlogits && (reinterpret_cast<uintptr_t>(dlogits16) & 7) == 0) ? dlogits16 : nullptr;
  if (vec && n4 <= 256 * 4) ce_fwd_bwd_reg_kernel<256, 4><<<n, 256, 0, s>>>(logits, ld, tgt, Vc, loss, dlogits, ldd, inv_n, m16);
  else if (vec && n4 <= 256 * 10) ce_fwd_bwd_reg_kernel<256, 10><<<n, 256, 0, s>>>(logits, ld, tgt, Vc, loss, dlogits, ldd, inv_n, m16);
  else if (vec && n4 <= 512 * 10) ce_fwd_bwd_reg_kernel<512, 10><<<n, 512, 0, s>>>(logits, ld, tgt, Vc, loss, dlogits, ldd, inv_n, m16);
  else if (vec && n4 <= 1024 * 12) ce_fwd_bwd_reg_kernel<1024, 12><<<n, 1024, 0, s>>>(logits, ld, tgt, Vc, loss, dlogits, ldd, inv_n, m16);
  else {
    m16 = nullptr;
    ce_fwd_bwd_kernel<<<n, PW_THREADS, 0, s>>>(logits, ld, tgt, n, Vc, loss, dlogits, ldd, inv_n);
  }
  AA_CHECK_LAUNCH("ce_fwd_bwd");
  if (mirror_written) *mirror_written = m16 ? 1 : 0;
  return AA_OK;
}

// ---- cross-entropy fused into the vocabulary projection (gemm_tc.cu, PM == 3): the two small passes after the contraction ----
namespace {
// one warp per row: lse = M + log(sum_c s_c exp(m_c - M)); loss += (lse - x_t) / denom; scale[c] = exp(m_c - lse) / denom
__global__ void __launch_bounds__(PW_THREADS) ce_merge_kernel(const float* __restrict__ part, int chunks, int n_rows, const float* __restrict__ xt,
                                                              float inv_n, float* __restrict__ loss, float* __restrict__ scale) {
  const int r = blockIdx.x * (PW_THREADS / 32) + (threadIdx.x >> 5);
  const int l = threadIdx.x & 31;
  if (r >= n_rows) return;
  const float2* pr = reinterpret_cast<const float2*>(part) + (long long)r * chunks;
  float M = -INFINITY;
  for (int c = l; c < chunks; c += 32) M = fmaxf(M, pr[c].x);
  M = warp_max(M);
  float S = 0.f;
  for (int c = l; c < chunks; c += 32) {
    const float2 v = pr[c];
    S += v.y * expf(v.x - M);
  }
  S = warp_sum(S);
  const float lse = M + logf(S);
  for (int c = l; c < chunks; c += 32) scale[(long long)r * chunks + c] = expf(pr[c].x - lse) * inv_n;
  if (l == 0) atomicAdd(loss, (lse - xt[r]) * inv_n);
}

// dlogits = e * scale - onehot / denom (bf16, in place), dbp[j] += column sums.  CTA = (strip of 2 * PW_THREADS columns, block of
// CE_RB rows); a thread owns two adjacent columns.
constexpr int CE_RB = 32;
__global__ void __launch_bounds__(PW_THREADS) ce_fixup_kernel(bf16* __restrict__ e16, long long ld, int n_rows, int Vc, const float* __restrict__ scale,
                                                              int chunks, const long long* __restrict__ tgt, float inv_n, float* __restrict__ dbp,
                                                              const float* __restrict__ loss_acc, float* __restrict__ loss_out) {
  if (loss_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *loss_out = *loss_acc;      // (ce_merge finished before this launch)
  const int j = (blockIdx.x * PW_THREADS + threadIdx.x) * 2;
  if (j >= Vc) return;
  const int r0 = blockIdx.y * CE_RB, r1 = min(n_rows, r0 + CE_RB);
  const int c = j >> 5;
  float s0 = 0.f, s1 = 0.f;
  constexpr int U = 8;       // rows in flight per thread
  for (int rb = r0; rb < r1; rb += U) {
    __nv_bfloat162 v[U];
    float f[U];
    long long t[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = min(rb + u, r1 - 1);
      v[u] = *reinterpret_cast<const __nv_bfloat162*>(e16 + (long long)r * ld + j);
      f[u] = __ldg(scale + (long long)r * chunks + c);
      t[u] = __ldg(tgt + r);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (rb + u < r1) {
        const float2 x = __bfloat1622float2(v[u]);
        float d0 = x.x * f[u], d1 = x.y * f[u];
        if (t[u] == j) d0 -= inv_n;
        if (t[u] == j + 1) d1 -= inv_n;
        if (j + 1 >= Vc) d1 = 0.f;
        *reinterpret_cast<__nv_bfloat162*>(e16 + (long long)(rb + u) * ld + j) = __floats2bfloat162_rn(d0, d1);
        s0 += d0;
        s1 += d1;
      }
    }
  }
  atomicAdd(dbp + j, s0);
  if (j + 1 < Vc) atomicAdd(dbp + j + 1, s1);
}
}  // namespace

namespace {
__global__ void scale_bf16_unless_one_kernel(bf16* __restrict__ x, const float* __restrict__ g, long long n2, float* __restrict__ y, long long ny) {
  const float f = *g;
  if (f == 1.f) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(x) + i;
    const float2 v = __bfloat1622float2(*p);
    *p = __floats2bfloat162_rn(v.x * f, v.y * f);
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ny; i += (long long)gridDim.x * blockDim.x) y[i] *= f;
}
}  // namespace

int launch_scale_bf16_unless_one(__nv_bfloat16* x, const float* g, long long n, float* y, long long ny, cudaStream_t s) {
  if (n == 0 && ny == 0) return AA_OK;
  AA_REQUIRE(n % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0, "scale_bf16: even length and 4-byte alignment needed");
  long long nb = (n / 2 + PW_THREADS * 4 - 1) / (PW_THREADS * 4);
  nb = nb > 1184 ? 1184 : nb;
  if (nb < 1) nb = 1;
  scale_bf16_unless_one_kernel<<<(unsigned)nb, PW_THREADS, 0, s>>>(x, g, n / 2, y, ny);
  AA_CHECK_LAUNCH("scale_bf16_unless_one");
  return AA_OK;
}

int launch_ce_merge(const float* part, int chunks, int n_rows, const float* xt, long long denom, float* loss, float* scale, cudaStream_t s) {
  if (n_rows == 0) return AA_OK;
  const float inv_n = 1.f / (float)(denom > 0 ? denom : n_rows);
  ce_merge_kernel<<<ceil_div(n_rows, PW_THREADS / 32), PW_THREADS, 0, s>>>(part, chunks, n_rows, xt, inv_n, loss, scale);
  AA_CHECK_LAUNCH("ce_merge");
  return AA_OK;
}

int launch_ce_fixup(__nv_bfloat16* e16, long long ld, int n_rows, int Vc, const float* scale, int chunks, const long long* tgt, long long denom,
                    float* dbp, cudaStream_t s, const float* loss_acc, float* loss_out) {
  if (n_rows == 0) return AA_OK;
  AA_REQUIRE(ld % 2 == 0 && (reinterpret_cast<uintptr_t>(e16) & 3) == 0, "ce_fixup: bf16 rows must be 4-byte aligned");
  const float inv_n = 1.f / (float)(denom > 0 ? denom : n_rows);
  ce_fixup_kernel<<<dim3(ceil_div(Vc, 2 * PW_THREADS), ceil_div(n_rows, CE_RB)), PW_THREADS, 0, s>>>(e16, ld, n_rows, Vc, scale, chunks, tgt, inv_n, dbp, loss_acc, loss_out);
  AA_CHECK_LAUNCH("ce_fixup");
  return AA_OK;
}

// zero-fill of up to 8 buffers in one launch (16-byte vector stores where the buffer allows)
namespace {
__global__ void zero_multi_kernel(const CopySegs segs) {
  const int sidx = blockIdx.y;
  char* dst = static_cast<char*>(segs.dst[sidx]);
  const long long n = segs.bytes[sidx];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const long long n16 = n >> 4;
    for (long long i = tid; i < n16; i += nth) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0, 0, 0, 0);
    for (long long i = (n16 << 4) + tid; i < n; i += nth) dst[i] = 0;
  } else {
    for (long long i = tid; i < n; i += nth) dst[i] = 0;
  }
}
}  // namespace

int launch_zero_multi(int nsegs, void* const* dst, const long long* bytes, cudaStream_t s) {
  AA_REQUIRE(nsegs >= 0 && nsegs <= 8, "zero_multi: at most 8 segments (got %d)", nsegs);
  CopySegs segs{};
  long long nmax = 0;
  int n = 0;
  for (int i = 0; i < nsegs; ++i) {
    if (!dst[i] || bytes[i] <= 0) continue;
    segs.dst[n] = dst[i]; segs.bytes[n] = bytes[i];
    nmax = bytes[i] > nmax ? bytes[i] : nmax;
    ++n;
  }
  if (n == 0) return AA_OK;
  long long nb = (nmax / 16 + PW_THREADS * 4 - 1) / (PW_THREADS * 4);
  nb = nb < 1 ? 1 : (nb > 296 ? 296 : nb);
  zero_multi_kernel<<<dim3((unsigned)nb, n), PW_THREADS, 0, s>>>(segs);
  AA_CHECK_LAUNCH("zero_multi");
  return AA_OK;
}

int launch_scale_unless_one(float* x, const float* g, long long n, cudaStream_t s, __nv_bfloat16* x16) {
  if (n == 0) return AA_OK;
  long long nb = (n + PW_THREADS * 8 - 1) / (PW_THREADS * 8);
  nb = nb > 1184 ? 1184 : nb;
  scale_unless_one_kernel<<<(unsigned)nb, PW_THREADS, 0, s>>>(x, g, n, x16);
  AA_CHECK_LAUNCH("scale_unless_one");
  return AA_OK;
}

int launch_copy_multi(int nsegs, const void* const* src, void* const* dst, const long long* bytes, cudaStream_t s) {
  AA_REQUIRE(nsegs >= 0 && nsegs <= 8, "copy_multi: at most 8 segments (got %d)", nsegs);
  if (nsegs == 0) return AA_OK;
  CopySegs segs{};
  long long nmax = 0;
  for (int i = 0; i < nsegs; ++i) {
    segs.src[i] = src[i]; segs.dst[i] = dst[i]; segs.bytes[i] = bytes[i];
    nmax = bytes[i] > nmax ? bytes[i] : nmax;
  }
  long long nb = (nmax / 16 + PW_THREADS * 4 - 1) / (PW_THREADS * 4);   // ~4 uint4 per thread of the largest segment
  nb = nb < 1 ? 1 : (nb > 592 ? 592 : nb);
  copy_multi_kernel<<<dim3((unsigned)nb, nsegs), PW_THREADS, 0, s>>>(segs);
  AA_CHECK_LAUNCH("copy_multi");
  return AA_OK;
}

}  // namespace aa
