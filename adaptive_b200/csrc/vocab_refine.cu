// Filter-and-refine arg-max of the vocabulary projection for greedy decoding (adaptive_attention.py:132, 201:
// predicted = (mlp(c_hat + h)).max(2)[1]).
//
// The sampler only needs arg-max_j (u . W_j + b_j), exact in fp32.  A fp32-accurate (3xTF32) contraction spends three
// tensor-core products and twice the operand traffic on every one of the Vc logits although all but a handful are
// far below the row's maximum.  Instead:
//   1. ONE single-pass tensor-core contraction over the tf32 "hi" halves of the operands, whose epilogue keeps only
//      the per-(row, 64-column tile) maximum (gemm_tc.cu, arg-max partials);
//   2. argmax_filter: with the rigorous bound |approx_j - exact_j| <= c ||u||_2 ||W_j||_2 (Cauchy-Schwarz over the
//      per-product rounding errors of the two tf32 roundings, plus the fp32 accumulation), a tile can hold the exact
//      arg-max only if its approximate maximum + bound reaches the best (approximate maximum - bound) of the row;
//      those (row, tile) pairs -- ~1.2 per row for tf32 -- are appended to the tile's row list, every other partial
//      is set to -inf;
//   3. argmax_refine: one CTA per tile with a non-empty list keeps the tile's 64 fp32 weight rows in shared memory and
//      recomputes the listed rows' 64 logits in plain fp32 FMA arithmetic, writing the tile's exact (max, index) back;
//   4. the existing argmax_finalize reduces the partials (lowest index wins ties) and gathers the next embedding.
// Every column that could be the exact arg-max is recomputed exactly, so the ids equal those of an exact fp32 projection
// (up to fp32 summation order, like any fp32 implementation).
#include "kernels.cuh"

namespace aa {
namespace {

constexpr int RF_THREADS = 256;
constexpr int RF_WARPS = RF_THREADS / 32;
constexpr int RF_ROWS = 4;       // rows a warp carries through one pass over the tile's weights

__device__ unsigned long long d_refine_pairs = 0;    // diagnostics: (row, tile) pairs handed to the refinement so far

// wnorm[t] = max_{j in tile t} ||W[j, :]||_2, rounded up   (once per decode call)
__global__ void tile_wnorm_kernel(const float* __restrict__ W, int Vc, int H, int tile_n, float* __restrict__ wnorm) {
  __shared__ float red[RF_WARPS];
  const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float best = 0.f;
  for (int j = t * tile_n + warp; j < min(Vc, (t + 1) * tile_n); j += RF_WARPS) {
    float ss = 0.f;
    for (int k = lane; k < H; k += 32) {
      const float x = __ldg(W + (long long)j * H + k);
      ss = fmaf(x, x, ss);
    }
    ss = warp_sum(ss);
    best = fmaxf(best, ss);
  }
  if (lane == 0) red[warp] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < RF_WARPS; ++i) best = fmaxf(best, red[i]);
    wnorm[t] = sqrtf(best) * (1.f + 1e-6f);
  }
}

// one warp per row: row norm, row threshold, candidate lists
__global__ void __launch_bounds__(RF_THREADS) argmax_filter_kernel(float* __restrict__ pmax, int tiles, int R, const float* __restrict__ u,
                                                                   long long ldu, long long lo_off, int H, const float* __restrict__ wnorm,
                                                                   float c, int* __restrict__ counts, int* __restrict__ list) {
  const int r = blockIdx.x * RF_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* ur = u + (long long)r * ldu;
  float ss = 0.f;
  for (int k = lane * 4; k < H; k += 128) {
    const float4 hi = *reinterpret_cast<const float4*>(ur + k), lo = *reinterpret_cast<const float4*>(ur + lo_off + k);
    const float x0 = hi.x + lo.x, x1 = hi.y + lo.y, x2 = hi.z + lo.z, x3 = hi.w + lo.w;
    ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
  }
  const float cn = c * sqrtf(warp_sum(ss)) * (1.f + 1e-6f);
  float* pr = pmax + (long long)r * tiles;
  float L = -INFINITY;      // best lower bound of the row's exact maximum
  for (int t = lane; t < tiles; t += 32) L = fmaxf(L, pr[t] - cn * wnorm[t]);
  L = warp_max(L);
  int mine = 0;
  for (int t = lane; t < tiles; t += 32) {
    const float p = pr[t];
    if (p + cn * wnorm[t] >= L) {           // the tile's exact maximum may reach the row's: refine it
      const int pos = atomicAdd(&counts[t], 1);
      list[(long long)t * R + pos] = r;
      ++mine;
    } else {
      pr[t] = -INFINITY;                    // cannot hold the arg-max
    }
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) atomicAdd(&d_refine_pairs, (unsigned long long)mine);
}

// Exact fp32 logits of the listed rows.  Work unit = (64-column tile, chunk of up to 32 listed rows); the units are dealt round
// robin to a persistent grid of one CTA per SM.  (One CTA per tile is not enough: at a given step most rows of a batch favour the
// same few words -- with the synthetic weights nearly all 4096 rows list the same tile -- and that CTA would do all the work.)
// A CTA keeps the unit's 64 weight rows in shared memory; each warp carries RF_ROWS rows through one pass over them.
__global__ void __launch_bounds__(RF_THREADS, 1) argmax_refine_kernel(const float* __restrict__ W, const float* __restrict__ bias, int Vc, int H,
                                                                      const float* __restrict__ u, long long ldu, long long lo_off, int R,
                                                                      int* __restrict__ counts, const int* __restrict__ list,
                                                                      float* __restrict__ pmax, int* __restrict__ pidx, int tiles) {
  extern __shared__ __align__(16) float sm[];
  constexpr int CH = RF_WARPS * RF_ROWS;               // listed rows per work unit
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ldw = H + 4;                               // (H % 32 == 0: 8 lanes x 16 B of one LDS.128 phase hit 32 distinct banks)
  float* Ws = sm;                                      // [64][H + 4]
  float* us = sm + 64 * ldw + warp * RF_ROWS * H;      // [RF_ROWS][H] per warp
  int* pre = reinterpret_cast<int*>(sm + 64 * ldw + RF_WARPS * RF_ROWS * H);   // [tiles + 1] exclusive prefix of the units per tile
  if (warp == 0) {     // blocked exclusive scan of the units per tile: lane l owns tiles [l * per, (l + 1) * per)
    const int per = (tiles + 31) / 32;
    int loc = 0;
    for (int i = 0; i < per; ++i) {
      const int t = lane * per + i;
      const int c = t < tiles ? (__ldcg(counts + t) + CH - 1) / CH : 0;
      if (t < tiles) pre[t] = c;
      loc += c;
    }
    int inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    int acc = inc - loc;
    for (int i = 0; i < per; ++i) {
      const int t = lane * per + i;
      if (t < tiles) {
        const int c = pre[t];
        pre[t] = acc;
        acc += c;
      }
    }
    if (lane == 31) pre[tiles] = inc;
  }
  __syncthreads();
  const int total = pre[tiles];
  const int h4 = H / 4;
  int cur_tile = -1;
  for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
    int lo_t = 0, hi_t = tiles - 1;                    // largest t with pre[t] <= unit (tiles without units share the next one's prefix)
    while (lo_t < hi_t) {
      const int mid = (lo_t + hi_t + 1) >> 1;
      if (pre[mid] <= unit) lo_t = mid; else hi_t = mid - 1;
    }
    const int t = lo_t, chunk = unit - pre[t];
    const int n = counts[t];
    const int j0 = t * 64, nc = min(64, Vc - j0);
    if (t != cur_tile) {
      __syncthreads();                                 // every warp is done with the previous tile's weights
      for (int i = tid; i < 64 * h4; i += RF_THREADS) {     // asynchronous copies: all of a thread's 16-byte pieces are in flight at once
        const int j = i / h4, k4 = i - j * h4;
        float* dst = Ws + j * ldw + k4 * 4;
        if (j < nc) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(W + (long long)(j0 + j) * H + k4 * 4) : "memory");
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      cur_tile = t;
    }
    const int i0 = chunk * CH + warp * RF_ROWS;
    if (i0 >= n) continue;                             // (warp-uniform; the barriers above are reached by every warp of the CTA)
    const float b0 = lane < nc ? (bias ? __ldg(bias + j0 + lane) : 0.f) : -INFINITY;
    const float b1 = lane + 32 < nc ? (bias ? __ldg(bias + j0 + lane + 32) : 0.f) : -INFINITY;
    const float4* w0 = reinterpret_cast<const float4*>(Ws + lane * ldw);
    const float4* w1 = reinterpret_cast<const float4*>(Ws + (lane + 32) * ldw);
    int rows[RF_ROWS];
#pragma unroll
    for (int m = 0; m < RF_ROWS; ++m) rows[m] = list[(long long)t * R + min(i0 + m, n - 1)];
    for (int k = lane * 4; k < H; k += 128) {            // (the RF_ROWS rows' loads of one k are issued together)
      float4 hi[RF_ROWS], lo[RF_ROWS];
#pragma unroll
      for (int m = 0; m < RF_ROWS; ++m) {
        const float* ur = u + (long long)rows[m] * ldu;
        hi[m] = __ldcg(reinterpret_cast<const float4*>(ur + k));
        lo[m] = __ldcg(reinterpret_cast<const float4*>(ur + lo_off + k));
      }
#pragma unroll
      for (int m = 0; m < RF_ROWS; ++m)
        *reinterpret_cast<float4*>(us + m * H + k) = make_float4(hi[m].x + lo[m].x, hi[m].y + lo[m].y, hi[m].z + lo[m].z, hi[m].w + lo[m].w);
    }
    __syncwarp();
    float4 a0[RF_ROWS], a1[RF_ROWS];                    // four partial sums (k mod 4) per (row, column)
#pragma unroll
    for (int m = 0; m < RF_ROWS; ++m) a0[m] = a1[m] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int k4 = 0; k4 < h4; ++k4) {
      const float4 x0 = w0[k4], x1 = w1[k4];
#pragma unroll
      for (int m = 0; m < RF_ROWS; ++m) {
        const float4 uv = *reinterpret_cast<const float4*>(us + m * H + k4 * 4);
        a0[m].x = fmaf(uv.x, x0.x, a0[m].x); a0[m].y = fmaf(uv.y, x0.y, a0[m].y);
        a0[m].z = fmaf(uv.z, x0.z, a0[m].z); a0[m].w = fmaf(uv.w, x0.w, a0[m].w);
        a1[m].x = fmaf(uv.x, x1.x, a1[m].x); a1[m].y = fmaf(uv.y, x1.y, a1[m].y);
        a1[m].z = fmaf(uv.z, x1.z, a1[m].z); a1[m].w = fmaf(uv.w, x1.w, a1[m].w);
      }
    }
#pragma unroll
    for (int m = 0; m < RF_ROWS; ++m) {
      const float v0 = ((a0[m].x + a0[m].y) + (a0[m].z + a0[m].w)) + b0;
      const float v1 = ((a1[m].x + a1[m].y) + (a1[m].z + a1[m].w)) + b1;
      float best = v0;
      int bi = j0 + lane;
      if (v1 > best) { best = v1; bi = j0 + lane + 32; }      // (strict: the lower column wins a tie)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0 && i0 + m < n) {
        pmax[(long long)rows[m] * tiles + t] = best;
        pidx[(long long)rows[m] * tiles + t] = bi;
      }
    }
    __syncwarp();
  }
  // the CTA that finishes last clears the lists for the next step's filter (counts[tiles] = number of finished CTAs)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&counts[tiles], 1) == (int)gridDim.x - 1) {
      for (int t = 0; t <= tiles; ++t) counts[t] = 0;
      __threadfence();
    }
  }
}

size_t refine_smem(int H, int tiles) { return sizeof(float) * ((size_t)64 * (H + 4) + (size_t)RF_WARPS * RF_ROWS * H + tiles + 1); }

}  // namespace

long long refine_pairs(int reset) {
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, d_refine_pairs, sizeof(v)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned long long z = 0;
    cudaMemcpyToSymbol(d_refine_pairs, &z, sizeof(z));
  }
  return (long long)v;
}

bool argmax_refine_supported(int Vc, int H) { return Vc > 64 && H % 4 == 0 && refine_smem(H, ceil_div(Vc, 64)) <= 220 * 1024; }

int launch_tile_wnorm(const float* W, int Vc, int H, float* wnorm, cudaStream_t s) {
  tile_wnorm_kernel<<<ceil_div(Vc, 64), RF_THREADS, 0, s>>>(W, Vc, H, 64, wnorm);
  AA_CHECK_LAUNCH("tile_wnorm");
  return AA_OK;
}

int launch_argmax_filter(float* pmax, int tiles, int R, const float* u, long long ldu, long long lo_off, int H, const float* wnorm, float c,
                         int* counts, int* list, cudaStream_t s) {
  argmax_filter_kernel<<<ceil_div(R, RF_WARPS), RF_THREADS, 0, s>>>(pmax, tiles, R, u, ldu, lo_off, H, wnorm, c, counts, list);
  AA_CHECK_LAUNCH("argmax_filter");
  return AA_OK;
}

int launch_argmax_refine(const float* W, const float* bias, int Vc, int H, const float* u, long long ldu, long long lo_off, int R, int* counts,
                         const int* list, float* pmax, int* pidx, int tiles, cudaStream_t s) {
  static size_t granted = 0;
  const size_t smem = refine_smem(H, tiles);
  if (smem > granted) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(argmax_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    granted = smem;
  }
  argmax_refine_kernel<<<num_sms(), RF_THREADS, smem, s>>>(W, bias, Vc, H, u, ldu, lo_off, R, counts, list, pmax, pidx, tiles);
  AA_CHECK_LAUNCH("argmax_refine");
  return AA_OK;
}

}  // namespace aa
