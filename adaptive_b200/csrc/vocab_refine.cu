// Filter-and-refine arg-max of the vocabulary projection for greedy decoding (adaptive_attention.py:132, 201:
// predicted = (mlp(c_hat + h)).max(2)[1]).
//
// The sampler only needs arg-max_j (u . W_j + b_j), exact in fp32.  A fp32-accurate (3xTF32) contraction spends three
// tensor-core products and twice the operand traffic on every one of the Vc logits although all but a handful are
// far below the row's maximum.  Instead:
//   1. ONE single-pass tensor-core contraction over bf16 mirrors (or the tf32 "hi" halves) of the operands, whose epilogue
//      keeps only the per-(row, 16-column tile) maximum (gemm_tc.cu, arg-max partials);
//   2. argmax_filter: with the rigorous bound |approx_j - exact_j| <= c ||u||_2 ||W_j||_2 (Cauchy-Schwarz over the
//      per-product rounding errors of the two operand roundings -- unit roundoff 2^-8 each for bf16 (8 significant bits,
//      round to nearest), 2^-11 for tf32 (cvt.rna) -- plus an allowance for the fp32 accumulation: c = 2.1 * 2^-8 /
//      1.1 * 2^-10), a tile can hold the exact arg-max only if its approximate
//      maximum + bound reaches the best (approximate maximum - bound) of the row; those (row, tile) pairs -- ~1.3 per row for
//      tf32, ~2.2 for bf16 at config 3 (625 tiles) -- are appended to the tile's row list together with their rank ("slot")
//      among the row's candidates (tests/test_host_cpu.py::test_argmax_filter_bound_never_drops_the_exact_argmax checks the bound in numpy);
//   3. argmax_refine: work units (tile, chunk of listed rows) dealt to a persistent grid; a unit keeps the tile's 16 fp32 weight
//      rows in shared memory and recomputes the listed rows' 16 logits in plain fp32 FMA arithmetic, writing the tile's exact
//      (max, index) to the row's slot;
//   4. argmax_finalize reduces the row's ncand[row] refined candidates (lowest index wins ties) and gathers the next embedding.
// Every column that could be the exact arg-max is recomputed exactly, so the ids equal those of an exact fp32 projection
// (up to fp32 summation order, like any fp32 implementation).
#include "kernels.cuh"

namespace aa {
namespace {

constexpr int RF_THREADS = 256;
constexpr int RF_WARPS = RF_THREADS / 32;
constexpr int RF_ROWS = 4;       // rows a warp carries through one pass over the tile's weights
constexpr int RF_TN = 16;        // columns per tile = granularity of the first pass's maxima (gemm_tc_argmax_tile_n_plain)

__device__ unsigned long long d_refine_pairs = 0;    // diagnostics: (row, tile) pairs handed to the refinement so far

// wnorm[t] = max_{j in tile t} ||W[j, :]||_2 and (optional) dwnorm[t] = max_j ||W[j, :] - bf16(W[j, :])||_2, both rounded up
// (once per decode call).  bf16() is the round-to-nearest-even conversion the bf16 mirror of W is made with (launch_cast2d).
__global__ void tile_wnorm_kernel(const float* __restrict__ W, int Vc, int H, int tile_n, float* __restrict__ wnorm, float* __restrict__ dwnorm) {
  __shared__ float red[RF_WARPS], redd[RF_WARPS];
  const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float best = 0.f, bestd = 0.f;
  for (int j = t * tile_n + warp; j < min(Vc, (t + 1) * tile_n); j += RF_WARPS) {
    float ss = 0.f, sd = 0.f;
    for (int k = lane; k < H; k += 32) {
      const float x = __ldg(W + (long long)j * H + k);
      const float dx = x - __bfloat162float(__float2bfloat16(x));
      ss = fmaf(x, x, ss);
      sd = fmaf(dx, dx, sd);
    }
    ss = warp_sum(ss);
    sd = warp_sum(sd);
    best = fmaxf(best, ss);
    bestd = fmaxf(bestd, sd);
  }
  if (lane == 0) { red[warp] = best; redd[warp] = bestd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < RF_WARPS; ++i) { best = fmaxf(best, red[i]); bestd = fmaxf(bestd, redd[i]); }
    wnorm[t] = sqrtf(best) * (1.f + 1e-5f);
    if (dwnorm) dwnorm[t] = sqrtf(bestd) * (1.f + 1e-5f);
  }
}

// Row norms, row thresholds, candidate lists: one warp per row, the row's maxima staged in shared memory.  A listed
// (row, tile) pair gets a SLOT = its rank among the row's candidates: the refinement writes the tile's exact (max, index) to
// pmax / pidx[row * tiles + slot], so that the final reduction reads ncand[row] entries instead of every tile of the row.
// The positions in a tile's list come from ONE global atomic per (CTA, tile): most rows of a batch list the same tile (the word
// most of them favour), and 4096 same-address atomics cost ~10 us.
constexpr int FL_THREADS = 256, FL_ROWS = FL_THREADS / 32;      // one warp per row
// Error bound of the first pass for row r and tile t:  cw[r] * wnorm[t] + cd[r] * dwnorm[t].
//   tf32 first pass (u16 == null): cw = c ||u||, cd = 0 -- the relative bound of two tf32 roundings per product.
//   bf16 first pass (u16, dwnorm given): the EXACT decomposition  u.W_j - u^.W^_j = u^.(W_j - W^_j) + (u - u^).W_j  with u^ = the bf16
//   mirror the pass actually read and W^ = the bf16 mirror of the weights, bounded term by term with Cauchy-Schwarz:
//   cd = ||u^||, cw = ||u - u^|| (+ 2^-13 ||u|| for the fp32 accumulation of the tensor pipe and the bias add).  Rigorous like the
//   relative bound 2.1 * 2^-8 ||u|| ||W_j|| it replaces, but ~2.3x tighter on real data (rounding errors average to ulp / sqrt(12),
//   they do not all sit at the half-ulp worst case).
__global__ void __launch_bounds__(FL_THREADS) argmax_filter_kernel(const float* __restrict__ pmax, int tiles, int R, const float* __restrict__ u,
                                                                   long long ldu, long long lo_off, int H, const float* __restrict__ wnorm,
                                                                   float c, int* __restrict__ counts, unsigned* __restrict__ list,
                                                                   int* __restrict__ ncand, int stage, const __nv_bfloat16* __restrict__ u16,
                                                                   long long ld16, const float* __restrict__ dwnorm) {
  extern __shared__ int fsm[];
  int* scnt = fsm;                                              // [tiles] pairs this CTA lists per tile, then the running position inside the CTA's range
  float* wn = reinterpret_cast<float*>(fsm + tiles);            // [tiles] weight norms
  float* dwn = wn + tiles;                                      // [tiles] norms of the weights' bf16 residuals (zero without dwnorm)
  __shared__ int spairs;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = blockIdx.x * FL_ROWS + warp;
  const bool live = r < R;
  // this warp's row of maxima is read three times: staged in shared memory when it fits (stage != 0), else re-read from L2
  float* srow = dwn + tiles + warp * tiles;
  const float* prow = stage ? srow : pmax + (long long)(live ? r : 0) * tiles;
  for (int t = tid; t < tiles; t += FL_THREADS) {
    scnt[t] = 0;
    wn[t] = __ldg(wnorm + t);
    dwn[t] = dwnorm ? __ldg(dwnorm + t) : 0.f;
  }
  if (tid == 0) spairs = 0;
  float cn = 0.f, cd = 0.f, L = INFINITY;
  if (live) {
    if (stage) {
      const float* pr = pmax + (long long)r * tiles;
#pragma unroll 4
      for (int t = lane; t < tiles; t += 32) srow[t] = __ldcg(pr + t);
    }
    const float* ur = u + (long long)r * ldu;
    float ss = 0.f, sh = 0.f, sd = 0.f;
    for (int k = lane * 4; k < H; k += 128) {
      const float4 hi = __ldcg(reinterpret_cast<const float4*>(ur + k)), lo = __ldcg(reinterpret_cast<const float4*>(ur + lo_off + k));
      const float x0 = hi.x + lo.x, x1 = hi.y + lo.y, x2 = hi.z + lo.z, x3 = hi.w + lo.w;
      ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
      if (u16) {      // the mirror the first pass actually contracted with (never re-derived: a second rounding could differ in the last place)
        const uint2 pk = __ldcg(reinterpret_cast<const uint2*>(u16 + (long long)r * ld16 + k));
        const float h0 = __uint_as_float(pk.x << 16), h1 = __uint_as_float(pk.x & 0xffff0000u);
        const float h2 = __uint_as_float(pk.y << 16), h3 = __uint_as_float(pk.y & 0xffff0000u);
        sh = fmaf(h0, h0, sh); sh = fmaf(h1, h1, sh); sh = fmaf(h2, h2, sh); sh = fmaf(h3, h3, sh);
        const float d0 = x0 - h0, d1 = x1 - h1, d2 = x2 - h2, d3 = x3 - h3;
        sd = fmaf(d0, d0, sd); sd = fmaf(d1, d1, sd); sd = fmaf(d2, d2, sd); sd = fmaf(d3, d3, sd);
      }
    }
    const float nu = sqrtf(warp_sum(ss));
    if (u16) {
      cn = sqrtf(warp_sum(sd)) * (1.f + 1e-5f) + 1.220703125e-4f * nu;      // ||u - u^|| + 2^-13 ||u||
      cd = sqrtf(warp_sum(sh)) * (1.f + 1e-5f);
    } else {
      cn = c * nu * (1.f + 1e-6f);
    }
  }
  __syncthreads();
  if (live) {
    float lo_best = -INFINITY;      // best lower bound of the row's exact maximum
    for (int t = lane; t < tiles; t += 32) lo_best = fmaxf(lo_best, prow[t] - (cn * wn[t] + cd * dwn[t]));
    L = warp_max(lo_best);
    int mine = 0;
    for (int t = lane; t < tiles; t += 32)
      if (prow[t] + (cn * wn[t] + cd * dwn[t]) >= L) {     // the tile's exact maximum may reach the row's: refine it
        atomicAdd(&scnt[t], 1);
        ++mine;
      }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) {
      ncand[r] = mine;
      atomicAdd(&spairs, mine);
    }
  }
  __syncthreads();
  for (int t = tid; t < tiles; t += FL_THREADS) {
    const int n = scnt[t];
    scnt[t] = n > 0 ? atomicAdd(&counts[t], n) : 0;      // this CTA's range in the tile's list
  }
  if (tid == 0) atomicAdd(&d_refine_pairs, (unsigned long long)spairs);
  __syncthreads();
  if (live) {
    int nrow = 0;
    for (int t0 = 0; t0 < tiles; t0 += 32) {
      const int t = t0 + lane;
      const bool flag = t < tiles && prow[t] + (cn * wn[t] + cd * dwn[t]) >= L;
      const unsigned m = __ballot_sync(0xffffffffu, flag);
      if (flag) {
        const int slot = nrow + __popc(m & ((1u << lane) - 1u));
        const int pos = atomicAdd(&scnt[t], 1);
        list[(long long)t * R + pos] = (unsigned)r | ((unsigned)slot << 20);
      }
      nrow += __popc(m);
    }
  }
}

// Exact fp32 logits of the listed rows.  Work unit = (16-column tile, chunk of up to 32 listed rows); the units are dealt round
// robin to a persistent grid.  (One CTA per tile is not enough: at a given step most rows of a batch favour the same few words --
// with the synthetic weights nearly all 4096 rows list the same tile -- and that CTA would do all the work.)
// A CTA keeps the unit's 16 weight rows in shared memory; each warp carries RF_ROWS rows through one pass over them, lane =
// (column, half of the reduction range).
__global__ void __launch_bounds__(RF_THREADS, 2) argmax_refine_kernel(const float* __restrict__ W, const float* __restrict__ bias, int Vc, int H,
                                                                      const float* __restrict__ u, long long ldu, long long lo_off, int R,
                                                                      int* __restrict__ counts, const unsigned* __restrict__ list,
                                                                      float* __restrict__ pmax, int* __restrict__ pidx, int tiles) {
  extern __shared__ __align__(16) float sm[];
  constexpr int CH = RF_WARPS * RF_ROWS;               // listed rows per work unit
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ldw = H + 4;                               // (H % 32 == 0: 8 lanes x 16 B of one LDS.128 phase hit 32 distinct banks)
  float* Ws = sm;                                      // [RF_TN][H + 4]
  float* us = sm + RF_TN * ldw + warp * RF_ROWS * H;   // [RF_ROWS][H] per warp
  int* pre = reinterpret_cast<int*>(sm + RF_TN * ldw + RF_WARPS * RF_ROWS * H);   // [tiles + 1] exclusive prefix of the units per tile
  {   // blocked exclusive scan of the units per tile: thread i owns tiles [i * per, (i + 1) * per)
    __shared__ int wsum[RF_WARPS];
    const int per = (tiles + RF_THREADS - 1) / RF_THREADS;
    int loc = 0;
    for (int i = 0; i < per; ++i) {
      const int t = tid * per + i;
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
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int base = 0;
    for (int w2 = 0; w2 < warp; ++w2) base += wsum[w2];
    int acc = base + inc - loc;
    for (int i = 0; i < per; ++i) {
      const int t = tid * per + i;
      if (t < tiles) {
        const int c = pre[t];
        pre[t] = acc;
        acc += c;
      }
    }
    if (tid == RF_THREADS - 1) pre[tiles] = base + inc;
  }
  __syncthreads();
  const int total = pre[tiles];
  const int h4 = H / 4, hh4 = H / 8;                   // float4s per row / per half row
  const int col = lane & (RF_TN - 1), kh = lane >> 4;  // this lane's column of the tile and half of the reduction range
  int cur_tile = -1;
  for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
    int lo_t = 0, hi_t = tiles - 1;                    // largest t with pre[t] <= unit (tiles without units share the next one's prefix)
    while (lo_t < hi_t) {
      const int mid = (lo_t + hi_t + 1) >> 1;
      if (pre[mid] <= unit) lo_t = mid; else hi_t = mid - 1;
    }
    const int t = lo_t, chunk = unit - pre[t];
    const int n = __ldcg(counts + t);
    const int j0 = t * RF_TN, nc = min(RF_TN, Vc - j0);
    if (t != cur_tile) {
      __syncthreads();                                 // every warp is done with the previous tile's weights
      for (int i = tid; i < RF_TN * h4; i += RF_THREADS) {     // asynchronous copies: all of a thread's 16-byte pieces are in flight at once
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
    const float bj = col < nc ? (bias ? __ldg(bias + j0 + col) : 0.f) : -INFINITY;
    const float4* wp = reinterpret_cast<const float4*>(Ws + col * ldw) + kh * hh4;
    int rows[RF_ROWS], slots[RF_ROWS];
#pragma unroll
    for (int m = 0; m < RF_ROWS; ++m) {
      const unsigned e = list[(long long)t * R + min(i0 + m, n - 1)];
      rows[m] = (int)(e & 0xFFFFFu);
      slots[m] = (int)(e >> 20);
    }
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
    float4 a[RF_ROWS];                                   // four partial sums (k mod 4) per row, this lane's column and half range
#pragma unroll
    for (int m = 0; m < RF_ROWS; ++m) a[m] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* up = reinterpret_cast<const float4*>(us) + kh * hh4;
#pragma unroll 4
    for (int k4 = 0; k4 < hh4; ++k4) {
      const float4 x = wp[k4];
#pragma unroll
      for (int m = 0; m < RF_ROWS; ++m) {
        const float4 uv = up[m * h4 + k4];
        a[m].x = fmaf(uv.x, x.x, a[m].x); a[m].y = fmaf(uv.y, x.y, a[m].y);
        a[m].z = fmaf(uv.z, x.z, a[m].z); a[m].w = fmaf(uv.w, x.w, a[m].w);
      }
    }
#pragma unroll
    for (int m = 0; m < RF_ROWS; ++m) {
      float v = (a[m].x + a[m].y) + (a[m].z + a[m].w);
      v += __shfl_xor_sync(0xffffffffu, v, 16);           // the two halves of the reduction range (same sum in both lanes)
      float best = v + bj;
      int bi = j0 + col;
#pragma unroll
      for (int o = RF_TN / 2; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }   // (the lower column wins a tie)
      }
      if (lane == 0 && i0 + m < n) {          // (compacted: the row's slot-th candidate)
        pmax[(long long)rows[m] * tiles + slots[m]] = best;
        pidx[(long long)rows[m] * tiles + slots[m]] = bi;
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

size_t refine_smem(int H, int tiles) { return sizeof(float) * ((size_t)RF_TN * (H + 4) + (size_t)RF_WARPS * RF_ROWS * H + tiles + 1); }

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

bool argmax_refine_supported(int Vc, int H) {
  // (12-bit slots in the list entries: at most 4096 tiles; the refinement's shared memory -- 16 weight rows + 32 rows of u -- must fit one SM)
  return Vc > 64 && Vc <= 4096 * RF_TN && H % 8 == 0 && refine_smem(H, ceil_div(Vc, RF_TN)) <= 220 * 1024;
}

int launch_tile_wnorm(const float* W, int Vc, int H, float* wnorm, cudaStream_t s, float* dwnorm) {
  tile_wnorm_kernel<<<ceil_div(Vc, RF_TN), RF_THREADS, 0, s>>>(W, Vc, H, RF_TN, wnorm, dwnorm);
  AA_CHECK_LAUNCH("tile_wnorm");
  return AA_OK;
}

int launch_argmax_filter(const float* pmax, int tiles, int R, const float* u, long long ldu, long long lo_off, int H, const float* wnorm, float c,
                         int* counts, unsigned* list, int* ncand, cudaStream_t s, const __nv_bfloat16* u16, long long ld16, const float* dwnorm) {
  AA_REQUIRE(R <= (1 << 20) && tiles <= 3900, "argmax_filter: at most 2^20 rows and 3900 tiles (got %d, %d)", R, tiles);
  AA_REQUIRE((u16 == nullptr) == (dwnorm == nullptr) && (!u16 || (ld16 % 4 == 0 && H % 4 == 0)), "argmax_filter: the bf16 mirror and the residual norms go together");
  const int stage = tiles <= 1100 ? 1 : 0;      // (3 + 8 arrays of one entry per tile within the default 48 KB of shared memory)
  argmax_filter_kernel<<<ceil_div(R, FL_ROWS), FL_THREADS, sizeof(int) * tiles * (3 + (stage ? FL_ROWS : 0)), s>>>(pmax, tiles, R, u, ldu, lo_off, H, wnorm,
                                                                                                            c, counts, list, ncand, stage, u16, ld16, dwnorm);
  AA_CHECK_LAUNCH("argmax_filter");
  return AA_OK;
}

int launch_argmax_refine(const float* W, const float* bias, int Vc, int H, const float* u, long long ldu, long long lo_off, int R, int* counts,
                         const unsigned* list, float* pmax, int* pidx, int tiles, cudaStream_t s) {
  static size_t granted = 0;
  const size_t smem = refine_smem(H, tiles);
  if (smem > granted) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(argmax_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    granted = smem;
  }
  argmax_refine_kernel<<<2 * num_sms(), RF_THREADS, smem, s>>>(W, bias, Vc, H, u, ldu, lo_off, R, counts, list, pmax, pidx, tiles);
  AA_CHECK_LAUNCH("argmax_refine");
  return AA_OK;
}

}  // namespace aa
