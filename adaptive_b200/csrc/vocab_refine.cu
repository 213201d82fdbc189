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
//   3. argmax_refine: work units (tile, chunk of listed rows) dealt to a persistent grid -- per CTA with the tile's 16 fp32 weight
//      rows in shared memory for the tiles many rows list, per warp straight from L2 for the others -- recompute the listed rows'
//      16 logits in plain fp32 FMA arithmetic and write the tile's exact (max, index) to the row's slot;
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
__device__ unsigned long long d_refine_units[4] = {0, 0, 0, 0};   // diagnostics: CTA units, warp units, tiles refined by CTAs, by warps

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

// The same filter with the row's maxima and bounds held in REGISTERS (NT values per lane, tiles <= 32 NT): one read of the maxima,
// one evaluation of the bounds, the candidates remembered as a bit mask per lane -- the staged kernel above reads its shared-memory copy
// three times and re-evaluates every bound twice (1745 instructions per row at config 3, latency-bound at 19 us per step:
// profiles/r02_decode_ncu_details.txt).  A candidate's slot is its rank in (lane, tile) order: any numbering 0 .. ncand - 1 serves, the
// final reduction is order-independent (lowest index wins ties).
template <int NT>
__global__ void __launch_bounds__(FL_THREADS) argmax_filter_reg_kernel(const float* __restrict__ pmax, int tiles, int R, const float* __restrict__ u,
                                                                       long long ldu, long long lo_off, int H, const float* __restrict__ wnorm,
                                                                       float c, int* __restrict__ counts, unsigned* __restrict__ list,
                                                                       int* __restrict__ ncand, const __nv_bfloat16* __restrict__ u16,
                                                                       long long ld16, const float* __restrict__ dwnorm) {
  extern __shared__ int fsm[];
  int* scnt = fsm;                                              // [tiles]
  float* wn = reinterpret_cast<float*>(fsm + tiles);            // [tiles]
  float* dwn = wn + tiles;                                      // [tiles]
  __shared__ int spairs;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = blockIdx.x * FL_ROWS + warp;
  const bool live = r < R;
  for (int t = tid; t < tiles; t += FL_THREADS) {
    scnt[t] = 0;
    wn[t] = __ldg(wnorm + t);
    dwn[t] = dwnorm ? __ldg(dwnorm + t) : 0.f;
  }
  if (tid == 0) spairs = 0;
  float p[NT];
  float cn = 0.f, cd = 0.f;
  if (live) {
    const float* pr = pmax + (long long)r * tiles;
#pragma unroll
    for (int i = 0; i < NT; ++i) p[i] = (i * 32 + lane < tiles) ? __ldcg(pr + i * 32 + lane) : -INFINITY;
    const float* ur = u + (long long)r * ldu;
    float ss = 0.f, sh = 0.f, sd = 0.f;
    for (int k = lane * 4; k < H; k += 128) {
      const float4 hi = __ldcg(reinterpret_cast<const float4*>(ur + k)), lo = __ldcg(reinterpret_cast<const float4*>(ur + lo_off + k));
      const float x0 = hi.x + lo.x, x1 = hi.y + lo.y, x2 = hi.z + lo.z, x3 = hi.w + lo.w;
      ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
      if (u16) {      // the mirror the first pass actually contracted with
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
  unsigned fm = 0;                  // bit i: tile i * 32 + lane is a candidate of this row
  int slot0 = 0;
  if (live) {
    float b[NT];
    float lo_best = -INFINITY;      // best lower bound of the row's exact maximum
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int t = min(i * 32 + lane, tiles - 1);
      b[i] = cn * wn[t] + cd * dwn[t];
      lo_best = fmaxf(lo_best, p[i] - b[i]);        // (p = -inf past the last tile)
    }
    const float L = warp_max(lo_best);
#pragma unroll
    for (int i = 0; i < NT; ++i)
      if (i * 32 + lane < tiles && p[i] + b[i] >= L) {       // the tile's exact maximum may reach the row's: refine it
        fm |= 1u << i;
        atomicAdd(&scnt[i * 32 + lane], 1);
      }
    const int mine = __popc(fm);
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    slot0 = inc - mine;
    if (lane == 31) {
      ncand[r] = inc;
      atomicAdd(&spairs, inc);
    }
  }
  __syncthreads();
  for (int t = tid; t < tiles; t += FL_THREADS) {
    const int n = scnt[t];
    scnt[t] = n > 0 ? atomicAdd(&counts[t], n) : 0;      // this CTA's range in the tile's list
  }
  if (tid == 0) atomicAdd(&d_refine_pairs, (unsigned long long)spairs);
  __syncthreads();
  while (fm) {
    const int i = __ffs(fm) - 1;
    fm &= fm - 1;
    const int t = i * 32 + lane;
    const int pos = atomicAdd(&scnt[t], 1);
    list[(long long)t * R + pos] = (unsigned)r | ((unsigned)slot0++ << 20);
  }
}

// Exact fp32 logits of the listed rows, two kinds of work units dealt to a persistent grid:
//   HOT  tiles (listed by more than RF_COLD_MAX rows -- at a given step most rows of a batch favour the same few words; with the
//        synthetic weights nearly all 4096 rows list the same tile): (tile, chunk of up to 32 listed rows) per CTA, the tile's 16 fp32
//        weight rows staged in shared memory once per CTA and tile, RF_ROWS rows per warp;
//   COLD tiles (a few rows each, several hundred of them per step): (tile, up to RF_COLD_ROWS rows) per WARP, the weights read straight
//        from L2.  As CTA units each of them cost a whole CTA the fixed latency of a 32 KB tile load for one busy warp, and the
//        kernel ran 2-3 such rounds deep (29 us at config 3, 60 % of the SMs idle at the end: profiles/r02_decode_ncu_details.txt).
// Both kinds go through refine_rows(): lanes along the reduction dimension (coalesced 512-byte reads of u and W), rows x 16 accumulators
// per lane, ONE fixed butterfly to reduce them -- the same summation order whichever kind of unit a (row, tile) pair lands in, so an
// image's ids do not depend on how many batch mates list the same tile (the shard == slice property of the decode tests).
constexpr int RF_COLD_MAX = 8;
constexpr int RF_COLD_ROWS = 2;     // rows per cold (warp) unit: leaves the registers for 16 weight reads in flight

template <bool W_SMEM, int NR>
__device__ __forceinline__ void refine_rows(const float* __restrict__ Wt, long long ldw, int nc, const float* __restrict__ bias, int j0,
                                            const float* __restrict__ u, long long ldu, long long lo_off, int H,
                                            const unsigned* __restrict__ lst, int i0, int n, float* __restrict__ pmax, int* __restrict__ pidx,
                                            int tiles, int lane) {
  static_assert(NR == 2 || NR == 4, "2 or 4 rows per warp");
  int rows[NR];
#pragma unroll
  for (int m = 0; m < NR; ++m) rows[m] = (int)(lst[min(i0 + m, n - 1)] & 0xFFFFFu);
  float acc[NR * RF_TN];
#pragma unroll
  for (int i = 0; i < NR * RF_TN; ++i) acc[i] = 0.f;
  for (int k = lane * 4; k < H; k += 128) {
    float4 uv[NR];
#pragma unroll
    for (int m = 0; m < NR; ++m) {
      const float* ur = u + (long long)rows[m] * ldu + k;
      const float4 hi = __ldcg(reinterpret_cast<const float4*>(ur)), lo = __ldcg(reinterpret_cast<const float4*>(ur + lo_off));
      uv[m] = make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
    }
    if constexpr (W_SMEM) {
#pragma unroll
      for (int j = 0; j < RF_TN; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(Wt + j * ldw + k);       // (rows past nc are zero-filled)
#pragma unroll
        for (int m = 0; m < NR; ++m) {
          float a = acc[m * RF_TN + j];
          a = fmaf(uv[m].x, x.x, a); a = fmaf(uv[m].y, x.y, a); a = fmaf(uv[m].z, x.z, a); a = fmaf(uv[m].w, x.w, a);
          acc[m * RF_TN + j] = a;
        }
      }
    } else {
      float4 x[RF_TN];            // all 16 reads from L2 in flight before the first FMA (the point of carrying only NR = 2 rows here)
#pragma unroll
      for (int j = 0; j < RF_TN; ++j)
        x[j] = j < nc ? __ldg(reinterpret_cast<const float4*>(Wt + j * ldw + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < RF_TN; ++j) {
#pragma unroll
        for (int m = 0; m < NR; ++m) {
          float a = acc[m * RF_TN + j];
          a = fmaf(uv[m].x, x[j].x, a); a = fmaf(uv[m].y, x[j].y, a); a = fmaf(uv[m].z, x[j].z, a); a = fmaf(uv[m].w, x[j].w, a);
          acc[m * RF_TN + j] = a;
        }
      }
    }
  }
  // NR * 16 sums x 32 lanes -> NR / 2 per lane: at every stage (lane bits 4, 3, 2, 1, 0 in this order -- the same reduction tree over
  // the lanes' partial sums for every value and for either NR) a lane keeps the half of its values whose index bit matches its lane bit
  // and hands the other half to its partner.
#pragma unroll
  for (int half = NR * RF_TN / 2, bit = 16; bit >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? acc[i] : acc[i + half];
      const float keep = up ? acc[i + half] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  // NR == 4: lane l holds i = 2 l, 2 l + 1 (i = row * 16 + column): row l / 8, columns 2 (l % 8), + 1.   NR == 2: lane l holds i = l.
  constexpr int LPR = 32 / NR;                     // lanes per row
  constexpr int CPL = RF_TN / LPR;                 // columns per lane (2 or 1)
  const int m = lane / LPR, c0 = (lane % LPR) * CPL;
  float best = c0 < nc ? acc[0] + (bias ? __ldg(bias + j0 + c0) : 0.f) : -INFINITY;
  int bi = j0 + c0;
  if constexpr (CPL == 2) {
    const float b1 = c0 + 1 < nc ? acc[1] + (bias ? __ldg(bias + j0 + c0 + 1) : 0.f) : -INFINITY;
    if (b1 > best) { best = b1; bi = j0 + c0 + 1; }            // (the lower column wins a tie)
  }
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane % LPR == 0 && i0 + m < n) {          // (compacted: the row's slot-th candidate)
    const unsigned e = lst[i0 + m];
    const long long o = (long long)(e & 0xFFFFFu) * tiles + (e >> 20);
    pmax[o] = best;
    pidx[o] = bi;
  }
}

__global__ void __launch_bounds__(RF_THREADS, 2) argmax_refine_kernel(const float* __restrict__ W, const float* __restrict__ bias, int Vc, int H,
                                                                      const float* __restrict__ u, long long ldu, long long lo_off, int R,
                                                                      int* __restrict__ counts, const unsigned* __restrict__ list,
                                                                      float* __restrict__ pmax, int* __restrict__ pidx, int tiles) {
  extern __shared__ __align__(16) float sm[];
  constexpr int CH = RF_WARPS * RF_ROWS;               // listed rows per CTA unit
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ldw = H + 4;
  float* Ws = sm;                                      // [RF_TN][H + 4]
  int* preH = reinterpret_cast<int*>(sm + RF_TN * ldw);   // [tiles + 1] exclusive prefix of the CTA units per tile
  int* preC = preH + tiles + 1;                           // [tiles + 1] exclusive prefix of the warp units per tile
  {   // blocked exclusive scans: thread i owns tiles [i * per, (i + 1) * per)
    __shared__ int wsumH[RF_WARPS], wsumC[RF_WARPS];
    const int per = (tiles + RF_THREADS - 1) / RF_THREADS;
    int locH = 0, locC = 0;
    for (int i = 0; i < per; ++i) {
      const int t = tid * per + i;
      const int n = t < tiles ? __ldcg(counts + t) : 0;
      const int cH = n > RF_COLD_MAX ? (n + CH - 1) / CH : 0, cC = n > RF_COLD_MAX ? 0 : (n + RF_COLD_ROWS - 1) / RF_COLD_ROWS;
      if (t < tiles) { preH[t] = cH; preC[t] = cC; }
      locH += cH;
      locC += cC;
      if (blockIdx.x == 0 && n > 0) atomicAdd(&d_refine_units[n > RF_COLD_MAX ? 2 : 3], 1ull);
    }
    int incH = locH, incC = locC;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int vH = __shfl_up_sync(0xffffffffu, incH, o), vC = __shfl_up_sync(0xffffffffu, incC, o);
      if (lane >= o) { incH += vH; incC += vC; }
    }
    if (lane == 31) { wsumH[warp] = incH; wsumC[warp] = incC; }
    __syncthreads();
    int baseH = 0, baseC = 0;
    for (int w2 = 0; w2 < warp; ++w2) { baseH += wsumH[w2]; baseC += wsumC[w2]; }
    int accH = baseH + incH - locH, accC = baseC + incC - locC;
    for (int i = 0; i < per; ++i) {
      const int t = tid * per + i;
      if (t < tiles) {
        const int cH = preH[t], cC = preC[t];
        preH[t] = accH; preC[t] = accC;
        accH += cH; accC += cC;
      }
    }
    if (tid == RF_THREADS - 1) {
      preH[tiles] = baseH + incH; preC[tiles] = baseC + incC;
      if (blockIdx.x == 0) {
        atomicAdd(&d_refine_units[0], (unsigned long long)(baseH + incH));
        atomicAdd(&d_refine_units[1], (unsigned long long)(baseC + incC));
      }
    }
  }
  __syncthreads();
  auto find = [&](const int* pre, int unit) {          // largest t with pre[t] <= unit (tiles without units share the next one's prefix)
    int lo_t = 0, hi_t = tiles - 1;
    while (lo_t < hi_t) {
      const int mid = (lo_t + hi_t + 1) >> 1;
      if (pre[mid] <= unit) lo_t = mid; else hi_t = mid - 1;
    }
    return lo_t;
  };
  const int totalH = preH[tiles], totalC = preC[tiles];
  const int h4 = H / 4;
  int cur_tile = -1;
  for (int unit = blockIdx.x; unit < totalH; unit += gridDim.x) {
    const int t = find(preH, unit), chunk = unit - preH[t];
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
    refine_rows<true, RF_ROWS>(Ws, ldw, nc, bias, j0, u, ldu, lo_off, H, list + (long long)t * R, i0, n, pmax, pidx, tiles, lane);
  }
  // cold units, one per warp; dealt from the far end of the grid so that the CTAs without a hot unit start on them at once
  for (int unit = ((int)gridDim.x - 1 - (int)blockIdx.x) * RF_WARPS + warp; unit < totalC; unit += gridDim.x * RF_WARPS) {
    const int t = find(preC, unit), chunk = unit - preC[t];
    const int n = __ldcg(counts + t);
    const int j0 = t * RF_TN, nc = min(RF_TN, Vc - j0);
    refine_rows<false, RF_COLD_ROWS>(W + (long long)j0 * H, H, nc, bias, j0, u, ldu, lo_off, H, list + (long long)t * R, chunk * RF_COLD_ROWS, n, pmax, pidx, tiles, lane);
  }
  // the CTA that finishes last clears the lists for the next step's filter (counts[tiles] = number of finished CTAs)
  __syncthreads();
  if (tid == 0) __threadfence();
  int last = 0;
  if (tid == 0) last = atomicAdd(&counts[tiles], 1) == (int)gridDim.x - 1;
  if (__syncthreads_or(last)) {
    for (int t = tid; t <= tiles; t += RF_THREADS) counts[t] = 0;
  }
}

size_t refine_smem(int H, int tiles) { return sizeof(float) * ((size_t)RF_TN * (H + 4) + 2 * ((size_t)tiles + 1)); }

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

int refine_units(long long* out4, int reset) {
  unsigned long long v[4] = {0, 0, 0, 0};
  AA_CHECK_CUDA(cudaMemcpyFromSymbol(v, d_refine_units, sizeof(v)));
  for (int i = 0; i < 4; ++i) out4[i] = (long long)v[i];
  if (reset) {
    const unsigned long long z[4] = {0, 0, 0, 0};
    AA_CHECK_CUDA(cudaMemcpyToSymbol(d_refine_units, z, sizeof(z)));
  }
  return AA_OK;
}

bool argmax_refine_supported(int Vc, int H) {
  // (12-bit slots in the list entries: at most 4096 tiles; the refinement's shared memory -- 16 weight rows + two prefix arrays -- must fit one SM)
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
  if (tiles <= 1024) {
    const unsigned grid = ceil_div(R, FL_ROWS);
    const size_t sm = sizeof(int) * tiles * 3;
    if (tiles <= 256) argmax_filter_reg_kernel<8><<<grid, FL_THREADS, sm, s>>>(pmax, tiles, R, u, ldu, lo_off, H, wnorm, c, counts, list, ncand, u16, ld16, dwnorm);
    else if (tiles <= 640) argmax_filter_reg_kernel<20><<<grid, FL_THREADS, sm, s>>>(pmax, tiles, R, u, ldu, lo_off, H, wnorm, c, counts, list, ncand, u16, ld16, dwnorm);
    else argmax_filter_reg_kernel<32><<<grid, FL_THREADS, sm, s>>>(pmax, tiles, R, u, ldu, lo_off, H, wnorm, c, counts, list, ncand, u16, ld16, dwnorm);
    AA_CHECK_LAUNCH("argmax_filter");
    return AA_OK;
  }
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
