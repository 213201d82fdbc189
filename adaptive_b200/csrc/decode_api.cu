// Greedy and beam decoding drivers: the whole sampler loop of
// Encoder2Decoder.sampler (adaptive_attention.py:186-216) runs on the device, four launches
// per step (gate GEMM, fused step kernel, vocabulary GEMM, arg-max + next-token gather) and
// no host round trip — the arg-max feeds the next step's A operand directly.
#include <stdlib.h>
#include "../../include/adaptive_b200.h"
#include "kernels.cuh"

namespace aa {

namespace {

constexpr int START_ID = 1;  // adaptive_attention.py:188
constexpr int END_ID = 2;    // build_vocab.py:48-51
constexpr int MAX_BEAM = 8;
int g_decode_table = [] { const char* e = getenv("AA_DECODE_TABLE"); return (e && e[0] == '0') ? 0 : 1; }();
int g_decode_table_min_rows = [] { const char* e = getenv("AA_DECODE_TABLE_MIN_ROWS"); return e ? atoi(e) : 0; }();
int g_decode_p_tc = [] { const char* e = getenv("AA_DECODE_P_TC"); return (e && e[0] == '0') ? 0 : 1; }();
int g_decode_qr_split = [] { const char* e = getenv("AA_DECODE_QR_SPLIT"); return (e && e[0] == '0') ? 0 : 1; }();
int g_force_simple_atten = 0;   // diagnostics (aa_debug_set_decode_atten_simple): register-staged attention kernel
// filter-and-refine arg-max of the greedy vocabulary projection (vocab_refine.cu); AA_DECODE_REFINE=0 / aa_debug_set_decode_argmax_refine(0)
// keep the fp32-accurate 3xTF32 contraction over the whole vocabulary
// (1 = tf32 first pass, 2 = bf16 first pass: half the operand bytes, ~8x the error bound -> a few candidate tiles per row instead of ~1.3)
int g_argmax_refine = [] { const char* e = getenv("AA_DECODE_REFINE"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2; }();

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T), 256);
    return p;
  }
};

struct DecodeWs {
  float *Wcat, *P, *stat, *Acat, *Acat2, *gates, *c, *c2, *u, *logits;
  // AA_PREC_TF32X3: the A/B operands of the two per-step contractions are stored as tf32 (hi | lo) row pairs
  // rows of the A operand: [emb (E) | h (H) | s (H)] hi half, padded to `lo` columns, then the lo half; the gate GEMM
  // reads the [emb | h] window, the q/r GEMM the [h | s] window.  W2 = [[W_g, 0], [W_g, W_s]] gives [q | r] in one GEMM.
  float *Wp_s, *pmax, *W2, *hs, *qr; int* pidx;
  int split, bm, K, Kp, lo, K2p, Hp, ldA, ldU, tiles_n;
  int qr_split;     // [q | r'] from two K = H tiles of one launch (needs a <= 64, H % 32 == 0)
  // filter-and-refine arg-max (vocab_refine.cu): 16-column partial tiles (`tiles16` of them), per-tile weight norms and candidate row lists
  int refine, tiles16; float *wnorm, *dwnorm; int *counts, *ncand; unsigned* list;   // refine: 0 = off, 1 = tf32 first pass, 2 = bf16 first pass
  __nv_bfloat16 *u16, *Wp16;
  // table mode (greedy, tensor-core pipeline, large batches): EG [Vc,5H] = embed . [W_ih[:, :E]; W_x[:, :E]]^T takes the word's half of
  // the gate contraction out of the loop (K = E+H -> H, N = 5H -> 4H: the sentinel block has no recurrent half in decode mode, Q3)
  int table; float *EG, *Whh_s, *emb_s, *wxe_s;
  int p_tc; float* wv_s;     // P = V W_v^T on tcgen05 with V split into tf32 (hi, lo) on the fly (TcGemmArgs::a_raw); wv_s = W_v split
  float *vg_s, *wx_s; int Ep;     // split pipeline: (hi | lo) copies of v_g [B, 2*Ep] and of the v_g columns of [W_ih; W_x] [5H, 2*Ep]
  int ldP, ld_qr;   // row strides of P and [q | r]: padded to 4 floats in the split pipeline (16-byte bulk copies)
  // beam only
  float *cum, *row_max, *row_lsum, *rec_alpha, *rec_beta;
  int *rec_word, *rec_src, *rec_wasdone, *done;
  size_t bytes;
};

// beam == 0: greedy layout (one row per image, no bookkeeping); beam >= 1: beam layout
DecodeWs carve_decode(const aa_dims& d, int beam, void* base) {
  const bool bm = beam >= 1;
  const size_t B = d.B, R = B * (bm ? beam : 1), H = d.H, E = d.E, K = E + H;
  const size_t L = d.T;  // max_len travels in d.T for sizing
  Carver c(base);
  DecodeWs w{};
  w.split = d.precision == AA_PREC_TF32X3;
  w.bm = bm ? 1 : 0;
  w.K = (int)K;
  w.Kp = (int)((K + 31) / 32 * 32);
  w.Hp = (int)((H + 31) / 32 * 32);
  w.K2p = (int)((2 * H + 31) / 32 * 32);
  w.lo = (int)((E + 2 * H + 31) / 32 * 32);
  w.ldA = w.split ? 2 * w.lo : (int)K;
  w.ldU = w.split ? 2 * w.Hp : (int)H;
  w.tiles_n = ceil_div(d.Vc, gemm_tc_argmax_tile_n(d.Vc));
  // (list entries pack the row into 20 bits: larger batches take the 3xTF32 projection of every logit)
  w.refine = (w.split && !bm && g_argmax_refine && argmax_refine_supported(d.Vc, d.H) && R <= ((size_t)1 << 20))
                 ? ((g_argmax_refine >= 2 && d.H % 8 == 0) ? 2 : 1) : 0;
  w.tiles16 = ceil_div(d.Vc, gemm_tc_argmax_tile_n_plain(d.Vc));
  const size_t ptiles = w.refine ? (size_t)w.tiles16 : (size_t)w.tiles_n;
  w.ldP = w.split ? (d.a + 3) / 4 * 4 : d.a;
  w.qr_split = (w.split && d.a <= 64 && H % 32 == 0 && g_decode_qr_split) ? 1 : 0;
  w.ld_qr = w.qr_split ? 128 : (2 * d.a + 3) / 4 * 4;
  // (building the table costs one [Vc x 5H x E] contraction per call, ~0.1 ms at cfgA: repaid from ~1k rows on; taken for every
  //  batch size all the same, so that a shard of a batch decodes bit-identically to the whole -- AA_DECODE_TABLE_MIN_ROWS overrides)
  w.table = (w.split && !bm && g_decode_table && R >= (size_t)g_decode_table_min_rows) ? 1 : 0;
  w.Ep = (int)((E + 31) / 32 * 32);
  w.Wcat = c.take<float>(w.table ? 0 : (size_t)5 * H * (w.split ? 2 * w.Kp : (int)K));
  w.EG = c.take<float>(w.table ? (size_t)d.Vc * 5 * H : 0);
  w.Whh_s = c.take<float>(w.table ? (size_t)4 * H * 2 * w.Hp : 0);
  w.emb_s = c.take<float>(w.table ? (size_t)d.Vc * 2 * w.Ep : 0);
  w.wxe_s = c.take<float>(w.table ? (size_t)5 * H * 2 * w.Ep : 0);
  w.p_tc = (w.split && d.a <= 64 && H % 32 == 0 && g_decode_p_tc) ? 1 : 0;
  w.wv_s = c.take<float>(w.p_tc ? (size_t)d.a * 2 * w.Hp : 0);
  w.W2 = c.take<float>(w.split ? (w.qr_split ? (size_t)(64 + d.a) * 2 * w.Hp : (size_t)2 * d.a * 2 * w.K2p) : 0);
  w.hs = c.take<float>(w.split ? R * 2 * H : 0);
  w.qr = c.take<float>(w.split ? R * w.ld_qr : 0);
  w.P = c.take<float>(B * d.k * w.ldP);
  w.stat = c.take<float>(R * 5 * H);
  w.Acat = c.take<float>(R * w.ldA);
  w.gates = c.take<float>(R * 5 * H);
  w.c = c.take<float>(R * H);
  w.u = c.take<float>(R * w.ldU);
  w.logits = c.take<float>((w.split && !bm) ? 0 : R * d.Vc);     // split greedy never materialises the logits
  w.Wp_s = c.take<float>(w.split ? (size_t)d.Vc * 2 * w.Hp : 0);
  w.pmax = c.take<float>((w.split && !bm) ? R * ptiles : 0);
  w.pidx = c.take<int>((w.split && !bm) ? R * ptiles : 0);
  w.wnorm = c.take<float>(w.refine ? w.tiles16 : 0);
  w.dwnorm = c.take<float>(w.refine == 2 ? w.tiles16 : 0);
  w.counts = c.take<int>(w.refine ? w.tiles16 + 1 : 0);      // (+ 1: finished-CTA ticket of the refinement kernel)
  w.list = c.take<unsigned>(w.refine ? (size_t)w.tiles16 * R : 0);
  w.ncand = c.take<int>(w.refine ? R : 0);
  w.u16 = reinterpret_cast<__nv_bfloat16*>(c.take<unsigned short>(w.refine == 2 ? R * H : 0));
  w.Wp16 = reinterpret_cast<__nv_bfloat16*>(c.take<unsigned short>(w.refine == 2 ? (size_t)d.Vc * H : 0));
  w.vg_s = c.take<float>(w.split ? B * 2 * w.Ep : 0);
  w.wx_s = c.take<float>(w.split ? (size_t)5 * H * 2 * w.Ep : 0);
  w.Acat2 = c.take<float>(bm ? R * w.ldA : 0);
  w.c2 = c.take<float>(bm ? R * H : 0);
  w.cum = c.take<float>(bm ? R : 0);
  w.row_max = c.take<float>(bm ? R : 0);
  w.row_lsum = c.take<float>(bm ? R : 0);
  w.rec_alpha = c.take<float>(bm ? L * R * d.k : 0);
  w.rec_beta = c.take<float>(bm ? L * R : 0);
  w.rec_word = c.take<int>(bm ? L * R : 0);
  w.rec_src = c.take<int>(bm ? L * R : 0);
  w.rec_wasdone = c.take<int>(bm ? L * R : 0);
  w.done = c.take<int>(bm ? R : 0);
  w.bytes = c.off;
  return w;
}

// Wcat [5H, E+H]: rows 0..4H = [W_ih[:, :E] | W_hh], rows 4H..5H = [W_x[:, :E] | 0]
// (decode-mode sentinel gate has no recurrent term: h~ = 0)
// split != 0: row = [tf32 hi (Kp, zero padded) | lo (Kp)], row stride ld = 2*Kp; else plain fp32, ld = K
__global__ void pack_wcat_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ sen_wx,
                                 float* __restrict__ Wcat, int H, int E, int split, int Kp, int ld) {
  const int n = blockIdx.x;  // output row
  const int K = E + H;
  float* dst = Wcat + (long long)n * ld;
  for (int c = threadIdx.x; c < (split ? Kp : K); c += blockDim.x) {
    float v = 0.f;
    if (c < K) {
      if (n < 4 * H) v = c < E ? w_ih[(long long)n * 2 * E + c] : w_hh[(long long)n * H + (c - E)];
      else v = (c < E && sen_wx) ? sen_wx[(long long)(n - 4 * H) * 2 * E + c] : 0.f;   // (sen_wx == null: baseline model, zero rows)
    }
    if (split) {
      float hi, lo;
      split_tf32(v, hi, lo);
      dst[c] = hi;
      dst[Kp + c] = lo;
    } else {
      dst[c] = v;
    }
  }
}

// W2s [64 + a, 2*Hp] (tf32 hi | lo): row j < a = W_g[j], rows a..63 zero, row 64 + j = W_s[j]: with the A window of the second
// 64-column tile shifted from h to s (TcGemmArgs::ashift_*) one launch gives [q | r'] = [h W_g^T | s W_s^T] with K = H per tile
// instead of K = 2H against a zero block; the attention kernel adds q to r'
__global__ void pack_wqr_split_kernel(const float* __restrict__ Wg, const float* __restrict__ Ws, float* __restrict__ W2, int a, int H, int Hp) {
  const int n = blockIdx.x;                      // 0 .. 64 + a
  const float* src = n < a ? Wg + (long long)n * H : (n >= 64 && Ws) ? Ws + (long long)(n - 64) * H : nullptr;
  float* dst = W2 + (long long)n * 2 * Hp;
  for (int c = threadIdx.x; c < Hp; c += blockDim.x) {
    float hi = 0.f, lo = 0.f;
    if (src && c < H) split_tf32(src[c], hi, lo);
    dst[c] = hi;
    dst[Hp + c] = lo;
  }
}

// Acat[r, :E] = embed[<start>], Acat[r, E:] = h0[r / beam], c[r] = c0[r / beam]
// W2 [2a, 2*K2p] (tf32 hi | lo): row j < a = [W_g[j] | 0], row a + j = [W_g[j] | W_s[j]]  ->  [h | s] W2^T = [q | s W_s^T + q]
__global__ void pack_wqr_kernel(const float* __restrict__ Wg, const float* __restrict__ Ws, float* __restrict__ W2, int a, int H, int K2p) {
  const int n = blockIdx.x;
  const int j = n < a ? n : n - a;
  float* dst = W2 + (long long)n * 2 * K2p;
  for (int c = threadIdx.x; c < K2p; c += blockDim.x) {
    float v = 0.f;
    if (c < H) v = Wg[(long long)j * H + c];
    else if (c < 2 * H && n >= a && Ws) v = Ws[(long long)j * H + (c - H)];   // (Ws == null: baseline model)
    float hi, lo;
    split_tf32(v, hi, lo);
    dst[c] = hi;
    dst[K2p + c] = lo;
  }
}

// (split layout: hi at column j, lo at column Kp + j of the row -- Kp here is the row's lo offset; pads are pre-zeroed)
__global__ void init_state_kernel(const float* __restrict__ embed, const float* __restrict__ h0, const float* __restrict__ c0,
                                  float* __restrict__ Acat, float* __restrict__ c, int H, int E, int beam, int split, int Kp, int ld) {
  const int r = blockIdx.x;
  const int b = r / beam;
  float* row = Acat + (long long)r * ld;
  for (int i = threadIdx.x; i < E + H; i += blockDim.x) {
    const float v = i < E ? embed[(long long)START_ID * E + i] : (h0 ? h0[(long long)b * H + (i - E)] : 0.f);
    if (split) {
      float hi, lo;
      split_tf32(v, hi, lo);
      row[i] = hi;
      row[Kp + i] = lo;
    } else {
      row[i] = v;
    }
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) c[(long long)r * H + i] = c0 ? c0[(long long)b * H + i] : 0.f;
}

__global__ void expand_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int cols, int beam) {
  const int r = blockIdx.x;
  const int b = r / beam;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) dst[(long long)r * cols + i] = src[(long long)b * cols + i];
}

__global__ void beam_init_kernel(float* __restrict__ cum, int* __restrict__ done, int R, int beam) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  cum[r] = (r % beam == 0) ? 0.f : -INFINITY;   // only hypothesis 0 is live at step 0
  done[r] = 0;
}

// per row: max and log-sum-exp denominator of the logits (log_softmax pieces)
__global__ void __launch_bounds__(256) row_lse_kernel(const float* __restrict__ logits, int Vc, float* __restrict__ row_max,
                                                      float* __restrict__ row_lsum) {
  __shared__ float red[8];
  __shared__ float bc;
  const int r = blockIdx.x;
  const float* row = logits + (long long)r * Vc;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < Vc; i += 256) m = fmaxf(m, row[i]);
  m = warp_max(m);
  if (l == 0) red[w] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
    bc = t;
  }
  __syncthreads();
  m = bc;
  float s = 0.f;
  for (int i = threadIdx.x; i < Vc; i += 256) s += expf(row[i] - m);
  s = warp_sum(s);
  __syncthreads();
  if (l == 0) red[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    row_max[r] = m;
    row_lsum[r] = logf(t);
  }
}

// One CTA per image: pick the `beam` best candidates among beam*Vc (ties -> lowest flat index),
// then move the winners' states into the next-step buffers and record the step.
__global__ void __launch_bounds__(256) beam_select_kernel(const float* __restrict__ logits, const float* __restrict__ row_max,
                                                          const float* __restrict__ row_lsum, float* __restrict__ cum,
                                                          int* __restrict__ done, int Vc, int beam, int H, int E,
                                                          const float* __restrict__ embed, const float* __restrict__ Acat,
                                                          const float* __restrict__ c, float* __restrict__ Acat_next,
                                                          float* __restrict__ c_next, int* __restrict__ rec_word,
                                                          int* __restrict__ rec_src, int* __restrict__ rec_wasdone, int split,
                                                          int Kp, int ldA) {
  __shared__ float s_cum[MAX_BEAM], s_max[MAX_BEAM], s_lsum[MAX_BEAM];
  __shared__ int s_done[MAX_BEAM];
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ int sel_idx[MAX_BEAM];
  __shared__ float sel_val[MAX_BEAM];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int r0 = b * beam;
  if (tid < beam) {
    s_cum[tid] = cum[r0 + tid];
    s_max[tid] = row_max[r0 + tid];
    s_lsum[tid] = row_lsum[r0 + tid];
    s_done[tid] = done[r0 + tid];
  }
  __syncthreads();
  const int total = beam * Vc;
  for (int round = 0; round < beam; ++round) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int f = tid; f < total; f += 256) {
      const int j = f / Vc, v = f - j * Vc;
      float val;
      if (s_done[j]) {
        val = (v == END_ID) ? s_cum[j] : -INFINITY;   // frozen hypothesis: one candidate
      } else {
        val = s_cum[j] + ((logits[(long long)(r0 + j) * Vc + v] - s_max[j]) - s_lsum[j]);
      }
      bool taken = false;
      for (int q = 0; q < round; ++q) taken |= (sel_idx[q] == f);
      if (taken) continue;
      if (val > best || (val == best && f < bi)) { best = val; bi = f; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (l == 0) { red_v[w] = best; red_i[w] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < 8; ++i)
        if (red_v[i] > best || (red_v[i] == best && red_i[i] < bi)) { best = red_v[i]; bi = red_i[i]; }
      if (bi == 0x7fffffff) {  // fewer than `beam` finite candidates: lowest untaken flat index
        bi = 0;
        bool again = true;
        while (again) {
          again = false;
          for (int q = 0; q < round; ++q)
            if (sel_idx[q] == bi) { ++bi; again = true; }
        }
        best = -INFINITY;
      }
      sel_idx[round] = bi;
      sel_val[round] = best;
    }
    __syncthreads();
  }
  // commit: scores, done flags, records
  if (tid < beam) {
    const int f = sel_idx[tid];
    const int src = f / Vc, word = f - src * Vc;
    const int t_was = s_done[src];
    cum[r0 + tid] = sel_val[tid];
    done[r0 + tid] = t_was | (word == END_ID);
    rec_word[r0 + tid] = word;
    rec_src[r0 + tid] = src;
    rec_wasdone[r0 + tid] = t_was;
  }
  // move state: Acat_next[slot] = [embed[word] | h[src]], c_next[slot] = c[src]
  for (int slot = 0; slot < beam; ++slot) {
    const int f = sel_idx[slot];
    const int src = f / Vc, word = f - src * Vc;
    const float* a_src = Acat + (long long)(r0 + src) * ldA;
    float* a_dst = Acat_next + (long long)(r0 + slot) * ldA;
    for (int i = tid; i < E; i += 256) {
      const float x = embed[(long long)word * E + i];
      if (split) {
        float hi, lo;
        split_tf32(x, hi, lo);
        a_dst[i] = hi;
        a_dst[Kp + i] = lo;
      } else {
        a_dst[i] = x;
      }
    }
    for (int i = tid; i < H; i += 256) {
      a_dst[E + i] = a_src[E + i];
      if (split) a_dst[Kp + E + i] = a_src[Kp + E + i];
      c_next[(long long)(r0 + slot) * H + i] = c[(long long)(r0 + src) * H + i];
    }
  }
}

// follow the back-pointers of the best final hypothesis (slot 0) of each image
__global__ void beam_backtrack_kernel(const int* __restrict__ rec_word, const int* __restrict__ rec_src,
                                      const int* __restrict__ rec_wasdone, const float* __restrict__ rec_alpha,
                                      const float* __restrict__ rec_beta, const float* __restrict__ cum, int L, int R, int beam,
                                      int k, long long* __restrict__ ids, float* __restrict__ attention,
                                      float* __restrict__ Beta, float* __restrict__ score) {
  const int b = blockIdx.x;
  const int r0 = b * beam;
  int slot = 0;
  if (threadIdx.x == 0 && score) score[b] = cum[r0];
  for (int t = L - 1; t >= 0; --t) {
    const int src = rec_src[(long long)t * R + r0 + slot];
    const int was = rec_wasdone[(long long)t * R + r0 + slot];
    if (threadIdx.x == 0) {
      ids[(long long)b * L + t] = rec_word[(long long)t * R + r0 + slot];
      Beta[(long long)b * L + t] = was ? 0.f : rec_beta[(long long)t * R + r0 + src];
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x)
      attention[((long long)b * L + t) * k + i] = was ? 0.f : rec_alpha[((long long)t * R + r0 + src) * k + i];
    slot = src;
  }
}

int check_decode(const aa_dims* d, const aa_weights* w, int max_len, int beam) {
  AA_REQUIRE(d && w, "decode: null dims/weights");
  {   // the baseline model (no sentinel) passes sen_wx = sen_wh = att_ws = NULL, all three together
    const int ns = (w->sen_wx != nullptr) + (w->sen_wh != nullptr) + (w->att_ws != nullptr);
    AA_REQUIRE(ns == 0 || ns == 3, "decode: sen_wx, sen_wh and att_ws must be given together (adaptive) or all be NULL (baseline)");
  }
  AA_REQUIRE(d->B >= 0 && d->k >= 1 && d->a >= 1 && d->a <= 128 && d->Vc >= 3, "decode: bad dims");
  AA_REQUIRE(d->H % 4 == 0 && d->E % 4 == 0 && d->H >= 4 && d->E >= 4, "decode: H and E must be multiples of 4");
  AA_REQUIRE(max_len >= 1, "decode: max_len must be >= 1");
  AA_REQUIRE(beam >= 1 && beam <= MAX_BEAM, "decode: beam must be in [1,%d]", MAX_BEAM);
  AA_REQUIRE(beam <= d->Vc, "decode: beam larger than the vocabulary");
  AA_REQUIRE(d->precision == AA_PREC_FP32 || d->precision == AA_PREC_TF32X3,
             "decode: precision must be AA_PREC_FP32 (SIMT) or AA_PREC_TF32X3 (tensor cores), got %d", d->precision);
  return AA_OK;
}

// One per-step contraction D[M,N] = A W^T (+Cin) (+bias): exact-fp32 SIMT, or 3xTF32 on tcgen05 over pre-split operands
// (then optionally with the fused arg-max partials instead of / next to D).
// (split: A is a column window starting at A of rows of lda floats whose lo half sits lo_a columns further; W rows are
// [hi (Kp) | lo (Kp)])
int dec_gemm(const DecodeWs& ws, int M, int N, int K, int Kp, const float* A, long long lda, int lo_a, long long a_cols, const float* W,
             long long ldw, float* D, long long ldd, const float* Cin, long long ldcin, const float* bias, float* pmax, int* pidx,
             cudaStream_t st, int kcut_n0 = 0, int kcut_cols = 0) {
  if (!ws.split) return gemm_nt(M, N, K, A, lda, W, ldw, D, ldd, Cin, ldcin, bias, nullptr, st);
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = Kp; g.elem_size = 4; g.split3 = 1;
  g.A = A; g.lda = lda; g.lo_a = lo_a; g.a_cols = a_cols; g.B = W; g.ldb = ldw;
  g.D32 = D; g.ldd32 = ldd; g.Cin = Cin; g.ldcin = ldcin; g.beta = 1.f; g.bias1 = bias;
  g.pmax = pmax; g.pidx = pidx;
  g.kcut_n0 = kcut_n0; g.kcut_cols = kcut_cols;
  return launch_gemm_tc(g, st);
}

// shared prologue: weight packing, P, static gate terms, initial state
int decode_prologue(const aa_dims& d, const aa_weights& w, const float* V, const float* v_g, const float* h0, const float* c0,
                    int beam, DecodeWs& ws, cudaStream_t st) {
  const int B = d.B, H = d.H, E = d.E, R = B * beam, K = E + H;
  if (ws.table) {
    // recurrent weights alone, and the per-word table of the input half (fp32-accurate 3xTF32 like the per-step contractions)
    AA_TRY(launch_split_tf32(w.w_hh, H, 4 * H, H, ws.Whh_s, ws.Hp, st));
    AA_TRY(launch_split_tf32(w.embed, E, d.Vc, E, ws.emb_s, ws.Ep, st));
    AA_TRY(launch_split_tf32(w.w_ih, 2 * E, 4 * H, E, ws.wxe_s, ws.Ep, st));
    if (w.sen_wx) AA_TRY(launch_split_tf32(w.sen_wx, 2 * E, H, E, ws.wxe_s + (size_t)4 * H * 2 * ws.Ep, ws.Ep, st));
    else AA_CHECK_CUDA(cudaMemsetAsync(ws.wxe_s + (size_t)4 * H * 2 * ws.Ep, 0, sizeof(float) * (size_t)H * 2 * ws.Ep, st));
    TcGemmArgs g{};
    g.M = d.Vc; g.N = 5 * H; g.K = ws.Ep; g.elem_size = 4; g.split3 = 1;
    g.A = ws.emb_s; g.lda = 2 * ws.Ep; g.B = ws.wxe_s; g.ldb = 2 * ws.Ep;
    g.D32 = ws.EG; g.ldd32 = 5 * H;
    AA_PROF("dec_gate_table", st, launch_gemm_tc(g, st));
  } else {
    pack_wcat_kernel<<<5 * H, 256, 0, st>>>(w.w_ih, w.w_hh, w.sen_wx, ws.Wcat, H, E, ws.split, ws.Kp, ws.split ? 2 * ws.Kp : K);
    AA_CHECK_LAUNCH("pack_wcat");
  }
  if (ws.split) {
    AA_TRY(launch_split_tf32(w.mlp_w, H, d.Vc, H, ws.Wp_s, ws.Hp, st));
    if (ws.refine) {
      AA_TRY(launch_tile_wnorm(w.mlp_w, d.Vc, H, ws.wnorm, st, ws.refine == 2 ? ws.dwnorm : nullptr));
      if (ws.refine == 2) AA_TRY(launch_cast2d(w.mlp_w, H, ws.Wp16, H, d.Vc, H, st));
      AA_CHECK_CUDA(cudaMemsetAsync(ws.counts, 0, sizeof(int) * (size_t)(ws.tiles16 + 1), st));
    }
    if (ws.qr_split) pack_wqr_split_kernel<<<64 + d.a, 256, 0, st>>>(w.att_wg, w.att_ws, ws.W2, d.a, H, ws.Hp);
    else pack_wqr_kernel<<<2 * d.a, 256, 0, st>>>(w.att_wg, w.att_ws, ws.W2, d.a, H, ws.K2p);
    AA_CHECK_LAUNCH("pack_wqr");
    // the operand windows read past the columns they need (up to a multiple of 32, against zero weight columns):
    // everything they can touch must be finite from the start, and the pad columns stay zero for the whole decode
    AA_CHECK_CUDA(cudaMemsetAsync(ws.Acat, 0, sizeof(float) * (size_t)R * ws.ldA, st));
    if (ws.bm) AA_CHECK_CUDA(cudaMemsetAsync(ws.Acat2, 0, sizeof(float) * (size_t)R * ws.ldA, st));
    if (ws.Hp != H) AA_CHECK_CUDA(cudaMemsetAsync(ws.u, 0, sizeof(float) * (size_t)R * ws.ldU, st));
  }
  if (ws.split && (ws.ldP != d.a || ws.ld_qr != 2 * d.a || ws.qr_split)) {   // pad columns are streamed into shared memory with their rows (never read): keep them finite
    AA_CHECK_CUDA(cudaMemsetAsync(ws.P, 0, sizeof(float) * (size_t)B * d.k * ws.ldP, st));
    AA_CHECK_CUDA(cudaMemsetAsync(ws.qr, 0, sizeof(float) * (size_t)R * ws.ld_qr, st));
  }
  if (ws.p_tc) {      // fp32-accurate 3xTF32 on tcgen05; V is read once and split in shared memory by the kernel's converter warps
    AA_TRY(launch_split_tf32(w.att_wv, H, d.a, H, ws.wv_s, ws.Hp, st));
    aa::ProfScope ps("dec_prologue_P", st);
    TcGemmArgs g{};
    g.M = B * d.k; g.N = d.a; g.K = H; g.elem_size = 4; g.split3 = 1; g.a_raw = 1;
    g.A = V; g.lda = H; g.B = ws.wv_s; g.ldb = 2 * ws.Hp;
    g.D32 = ws.P; g.ldd32 = ws.ldP;
    AA_TRY(launch_gemm_tc(g, st));
  } else {
    AA_PROF("dec_prologue_P", st, gemm_nt(B * d.k, d.a, H, V, H, w.att_wv, H, ws.P, ws.ldP, nullptr, 0, nullptr, nullptr, st));
  }
  // static (per image) gate terms: v_g half of x and the biases
  float* stat_img = beam > 1 ? ws.gates : ws.stat;   // [B,5H]; `gates` is free before the first step
  if (ws.split) {   // on tensor cores like the per-step contractions (fp32-accurate 3xTF32): 4096 x 2560 x 256 took 165 us on the SIMT path
    AA_TRY(launch_split_tf32(v_g, E, B, E, ws.vg_s, ws.Ep, st));
    AA_TRY(launch_split_tf32(w.w_ih + E, 2 * E, 4 * H, E, ws.wx_s, ws.Ep, st));
    if (w.sen_wx) AA_TRY(launch_split_tf32(w.sen_wx + E, 2 * E, H, E, ws.wx_s + (size_t)4 * H * 2 * ws.Ep, ws.Ep, st));
    TcGemmArgs g{};
    g.M = B; g.N = 4 * H; g.K = ws.Ep; g.elem_size = 4; g.split3 = 1;
    g.A = ws.vg_s; g.lda = 2 * ws.Ep; g.B = ws.wx_s; g.ldb = 2 * ws.Ep;
    g.D32 = stat_img; g.ldd32 = 5 * H; g.bias1 = w.b_ih; g.bias2 = w.b_hh;
    AA_TRY(launch_gemm_tc(g, st));
    if (w.sen_wx) {
      g.N = H; g.B = ws.wx_s + (size_t)4 * H * 2 * ws.Ep; g.D32 = stat_img + 4 * H; g.bias1 = nullptr; g.bias2 = nullptr;
      AA_TRY(launch_gemm_tc(g, st));
    }
  } else {
    AA_TRY(gemm_nt(B, 4 * H, E, v_g, E, w.w_ih + E, 2 * E, stat_img, 5 * H, nullptr, 0, w.b_ih, w.b_hh, st));
    if (w.sen_wx) AA_TRY(gemm_nt(B, H, E, v_g, E, w.sen_wx + E, 2 * E, stat_img + 4 * H, 5 * H, nullptr, 0, nullptr, nullptr, st));
  }
  if (!w.sen_wx) AA_CHECK_CUDA(cudaMemset2DAsync(stat_img + 4 * H, sizeof(float) * 5 * H, 0, sizeof(float) * H, (size_t)B, st));   // baseline model
  if (beam > 1) {
    expand_rows_kernel<<<R, 256, 0, st>>>(stat_img, ws.stat, 5 * H, beam);
    AA_CHECK_LAUNCH("expand_rows");
  }
  init_state_kernel<<<R, 256, 0, st>>>(w.embed, h0, c0, ws.Acat, ws.c, H, E, beam, ws.split, ws.lo, ws.ldA);
  AA_CHECK_LAUNCH("init_state");
  (void)K;
  return AA_OK;
}

// One decode step up to u = c_hat + h for R rows (R = B * beam) whose A operand is `Acur` and cell state `ccur`.
int decode_step_body(const aa_dims& d, const aa_weights& w, const DecodeWs& ws, const float* V, float* Acur, float* ccur, int R, int beam,
                     float* alpha, long long ld_alpha, float* beta, long long ld_beta, cudaStream_t st,
                     const long long* prev_ids = nullptr, long long ld_ids = 0) {
  const int H = d.H, E = d.E, K = E + H;
  if (ws.table) {
    // gates[:, :4H] = h_{t-1} W_hh^T + static; the word's half comes from the table inside dec_cell
    AA_PROF("dec_gate_gemm", st, dec_gemm(ws, R, 4 * H, H, ws.Hp, Acur + E, ws.ldA, ws.lo, ws.ldA - E, ws.Whh_s, 2 * ws.Hp, ws.gates, 5 * H,
                                          ws.stat, 5 * H, nullptr, nullptr, nullptr, st));
  } else
  // gates = [emb(w_t) | h_{t-1}] Wcat^T + static              (LSTM + sentinel-x pre-activations)
  // (the sentinel rows 4H..5H of Wcat are zero beyond the embedding columns -- decode-mode h~ = 0, Q3: their tiles stop the K loop at E)
  {
    AA_PROF("dec_gate_gemm", st, dec_gemm(ws, R, 5 * H, K, ws.Kp, Acur, ws.ldA, ws.lo, ws.ldA, ws.Wcat, ws.split ? 2 * ws.Kp : K, ws.gates,
                                          5 * H, ws.stat, 5 * H, nullptr, nullptr, nullptr, st, 4 * H, E));
  }
  if (!ws.split) {   // exact-fp32 path: one fused kernel incl. the q/r mat-vecs
    DecodeStepArgs p{};
    p.B = R; p.k = d.k; p.a = d.a; p.H = H; p.beam = beam;
    p.gates = ws.gates; p.c = ccur; p.h_out = Acur + E; p.ld_h = ws.ldA;
    p.P = ws.P; p.V = V; p.Wg = w.att_wg; p.Ws = w.att_ws; p.wh = w.att_wh;
    p.alpha = alpha; p.ld_alpha = ld_alpha; p.beta = beta; p.ld_beta = ld_beta;
    p.u = ws.u; p.ld_u = ws.ldU;
    AA_PROF("dec_step_fused", st, launch_decode_step(p, st));
    return AA_OK;
  }
  DecodeCellArgs cp{};
  cp.R = R; cp.H = H; cp.gates = ws.gates; cp.c = ccur; cp.hs = ws.hs; cp.A = Acur; cp.ldA = ws.ldA; cp.h_off = E; cp.lo_off = ws.lo;
  if (ws.table) { cp.EG = ws.EG; cp.stat = ws.stat; cp.prev_ids = prev_ids; cp.ld_ids = ld_ids; cp.start_id = START_ID; }
  AA_PROF("dec_cell", st, launch_decode_cell(cp, st));
  // [q | r] = [h | s] W2^T                                                              adaptive_attention.py:35,45
  if (ws.qr_split) {      // [q | r'] = [h W_g^T | s W_s^T]: two 64-column tiles, K = H each, the second one reading the s window of the rows
    aa::ProfScope ps("dec_qr_gemm", st);
    TcGemmArgs g{};
    g.M = R; g.N = 64 + d.a; g.K = ws.Hp; g.elem_size = 4; g.split3 = 1;
    g.A = Acur + E; g.lda = ws.ldA; g.lo_a = ws.lo; g.a_cols = ws.ldA - E; g.B = ws.W2; g.ldb = 2 * ws.Hp;
    g.D32 = ws.qr; g.ldd32 = ws.ld_qr; g.beta = 1.f;
    g.ashift_n0 = 64; g.ashift_cols = H;
    AA_TRY(launch_gemm_tc(g, st));
  } else {
    AA_PROF("dec_qr_gemm", st, dec_gemm(ws, R, 2 * d.a, 2 * H, ws.K2p, Acur + E, ws.ldA, ws.lo, ws.ldA - E, ws.W2, 2 * ws.K2p, ws.qr,
                                        ws.ld_qr, nullptr, 0, nullptr, nullptr, nullptr, st));
  }
  DecodeAttenArgs ap{};
  ap.R = R; ap.k = d.k; ap.a = d.a; ap.H = H; ap.beam = beam;
  ap.qr = ws.qr; ap.ld_qr = ws.ld_qr; ap.r_off = ws.qr_split ? 64 : 0; ap.r_partial = ws.qr_split; ap.hs = ws.hs; ap.P = ws.P; ap.ldP = ws.ldP; ap.V = V; ap.wh = w.att_wh;
  ap.force_simple = g_force_simple_atten;
  ap.no_sentinel = w.att_ws == nullptr;     // baseline model (baseline_attention.py:79-100): beta = 0
  ap.alpha = alpha; ap.ld_alpha = ld_alpha; ap.beta = beta; ap.ld_beta = ld_beta;
  ap.u = ws.u; ap.ld_u = ws.ldU; ap.u_lo_off = ws.Hp;
  ap.u16 = ws.refine == 2 ? ws.u16 : nullptr; ap.ld_u16 = H;
  AA_PROF("dec_step_fused", st, launch_decode_atten(ap, st));
  return AA_OK;
}

// ---- persistent variant (decode_persist.cu): workspace = [weight-derived region | per-call region] ----
struct PersistWs {
  // weight-derived (reusable across calls while the weights do not change)
  float* Whh; __nv_bfloat16* Wp16; float *wn, *dwn; float* EG;
  // per call
  float *P, *stat, *hA, *c, *part1, *approx; __nv_bfloat16* u16; unsigned* bar;
  int Kp, ldA, ldP, ldv, ks_max;
  size_t bytes;
};

PersistWs carve_persist(const aa_dims& d, void* base) {
  const size_t B = d.B, H = d.H;
  Carver c(base);
  PersistWs w{};
  w.Kp = (int)((H + 31) / 32 * 32);
  w.ldA = 2 * w.Kp;
  w.ldP = (d.a + 3) / 4 * 4;
  w.ldv = (d.Vc + 3) / 4 * 4;
  w.ks_max = 8;
  w.Whh = c.take<float>((size_t)4 * H * 2 * w.Kp);
  w.Wp16 = reinterpret_cast<__nv_bfloat16*>(c.take<unsigned short>((size_t)d.Vc * H));
  w.wn = c.take<float>((size_t)w.ldv);
  w.dwn = c.take<float>((size_t)w.ldv);
  w.EG = c.take<float>((size_t)d.Vc * 5 * H);
  w.P = c.take<float>(B * d.k * w.ldP);
  w.stat = c.take<float>(B * 5 * H);
  w.hA = c.take<float>(B * w.ldA);
  w.c = c.take<float>(B * H);
  w.part1 = c.take<float>((size_t)w.ks_max * B * 4 * H);
  w.approx = c.take<float>(B * w.ldv);
  w.u16 = reinterpret_cast<__nv_bfloat16*>(c.take<unsigned short>(B * H));
  w.bar = c.take<unsigned>(64);
  w.bytes = c.off;
  return w;
}

// hA[r] = tf32 (hi | lo) of h0[r] (zeros without an initial state; pad columns zero), c[r] = c0[r]
__global__ void persist_init_kernel(const float* __restrict__ h0, const float* __restrict__ c0, float* __restrict__ hA, float* __restrict__ c,
                                    int H, int Kp) {
  const int r = blockIdx.x;
  float* row = hA + (long long)r * 2 * Kp;
  for (int i = threadIdx.x; i < Kp; i += blockDim.x) {
    float hi = 0.f, lo = 0.f;
    if (i < H && h0) split_tf32(h0[(long long)r * H + i], hi, lo);
    row[i] = hi;
    row[Kp + i] = lo;
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) c[(long long)r * H + i] = c0 ? c0[(long long)r * H + i] : 0.f;
}

}  // namespace
}  // namespace aa

using namespace aa;

extern "C" {

long long aa_debug_refine_pairs(int reset) { return aa::refine_pairs(reset); }
int aa_debug_refine_units(long long* out4, int reset) { return out4 ? aa::refine_units(out4, reset) : AA_ERR_INVALID; }

int aa_debug_set_decode_argmax_refine(int on) {
  g_argmax_refine = on < 0 ? 0 : (on > 2 ? 2 : on);
  return AA_OK;
}

int aa_debug_set_decode_atten_simple(int on) {
  g_force_simple_atten = on ? 1 : 0;
  return AA_OK;
}

size_t aa_decode_workspace_bytes(const aa_dims* d, int beam) {
  if (!d || beam < 0) return 0;
  return carve_decode(*d, beam, nullptr).bytes;
}

int aa_greedy_decode(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const float* h0,
                     const float* c0, int max_len, int64_t* ids, float* attention, float* Beta, float* logits_out,
                     void* workspace, size_t workspace_bytes, void* stream) {
  AA_TRY(check_decode(d, w, max_len, 1));
  AA_REQUIRE(V && v_g && ids && attention && Beta, "aa_greedy_decode: null pointer");
  if (d->B == 0) return AA_OK;
  aa_dims dd = *d;
  dd.T = max_len;
  if (!workspace || workspace_bytes < aa_decode_workspace_bytes(&dd, 0)) {
    set_error("aa_greedy_decode: workspace too small (%zu < %zu)", workspace_bytes, aa_decode_workspace_bytes(&dd, 0));
    return AA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DecodeWs ws = carve_decode(dd, 0, workspace);
  const int B = d->B, H = d->H, E = d->E, K = E + H, Vc = d->Vc, k = d->k, L = max_len;
  AA_TRY(decode_prologue(dd, *w, V, v_g, h0, c0, 1, ws, st));
  const float* Wp = ws.split ? ws.Wp_s : w->mlp_w;
  for (int t = 0; t < L; ++t) {
    long long* ids_t = reinterpret_cast<long long*>(ids) + t;
    AA_TRY(decode_step_body(dd, *w, ws, V, ws.Acat, ws.c, B, 1, attention + (size_t)t * k, (long long)L * k, Beta + t, L, st,
                            t > 0 ? ids_t - 1 : nullptr, L));
    float* emb_dst = ws.table ? nullptr : ws.Acat;      // table mode: the next step reads the word's gate terms from EG, not its embedding
    if (ws.split) {
      // logits = u W_p^T + b_p on tensor cores; the epilogue keeps a per-(row, column tile) arg-max, so the [B,Vc]
      // logits never touch HBM unless the caller asked for them                                          :132, :201
      float* lg = logits_out ? logits_out + (size_t)t * B * Vc : nullptr;
      if (ws.refine && !lg) {
        // arg-max only: ONE tf32 pass over the hi halves with per-64-column maxima, then exact fp32 logits for the few (row, tile)
        // pairs that can hold the maximum (vocab_refine.cu).  c = 1.1 * 2^-10: two tf32 roundings per product (2^-11 each) and
        // an allowance of 2^-13.3 for the fp32 accumulation of the tensor pipe, relative to ||u|| ||W_j||.
        // (bf16 first pass, operands = the bf16 mirrors of u and W_p: the worst-case relative bound would be 2.1 * 2^-8 -- 8 significant
        // bits, unit roundoff 2^-8 per operand -- ; the filter uses the exact decomposition u.W - u^.W^ = u^.(W - W^) + (u - u^).W with
        // the norms of the ACTUAL residuals instead, ~2.3x tighter on real data and just as rigorous: vocab_refine.cu)
        TcGemmArgs g{};
        g.M = B; g.N = Vc;
        if (ws.refine == 2) {
          g.K = H; g.elem_size = 2; g.A = ws.u16; g.lda = H; g.B = ws.Wp16; g.ldb = H;
        } else {
          g.K = ws.Hp; g.elem_size = 4; g.A = ws.u; g.lda = ws.ldU; g.B = Wp; g.ldb = 2 * ws.Hp;
        }
        g.bias1 = w->mlp_b; g.pmax = ws.pmax; g.pidx = nullptr;     // (maxima only: the refinement writes the indices of the tiles that matter)
        AA_PROF("dec_vocab_gemm1", st, launch_gemm_tc(g, st));
        AA_PROF("dec_argmax_filter", st, launch_argmax_filter(ws.pmax, ws.tiles16, B, ws.u, ws.ldU, ws.Hp, H, ws.wnorm,
                                                              1.1f / 1024.f, ws.counts, ws.list, ws.ncand, st,
                                                              ws.refine == 2 ? ws.u16 : nullptr, H, ws.refine == 2 ? ws.dwnorm : nullptr));
        AA_PROF("dec_argmax_refine", st, launch_argmax_refine(w->mlp_w, w->mlp_b, Vc, H, ws.u, ws.ldU, ws.Hp, B, ws.counts, ws.list,
                                                              ws.pmax, ws.pidx, ws.tiles16, st));
        AA_PROF("dec_argmax", st, launch_argmax_finalize(ws.pmax, ws.pidx, ws.tiles16, B, ids_t, L, w->embed, E, emb_dst, ws.ldA, 1,
                                                         ws.lo, st, ws.ncand));
        continue;
      }
      AA_PROF("dec_vocab_gemm", st, dec_gemm(ws, B, Vc, H, ws.Hp, ws.u, ws.ldU, ws.Hp, ws.ldU, Wp, 2 * ws.Hp, lg, Vc, nullptr, 0, w->mlp_b,
                                             ws.pmax, ws.pidx, st));
      AA_PROF("dec_argmax", st, launch_argmax_finalize(ws.pmax, ws.pidx, ws.tiles_n, B, ids_t, L, w->embed, E, emb_dst, ws.ldA, 1,
                                                       ws.lo, st));
    } else {
      float* lg = logits_out ? logits_out + (size_t)t * B * Vc : ws.logits;
      AA_PROF("dec_vocab_gemm", st, gemm_nt(B, Vc, H, ws.u, H, w->mlp_w, H, lg, Vc, nullptr, 0, w->mlp_b, nullptr, st));   // :132
      AA_PROF("dec_argmax", st, launch_argmax_gather(lg, Vc, B, Vc, ids_t, L, w->embed, E, ws.Acat, K, st));                // :201
    }
  }
  return AA_OK;
}

int aa_debug_set_persist_trace(void* dev_ptr) { return set_persist_trace_buffer(dev_ptr); }

size_t aa_decode_persistent_workspace_bytes(const aa_dims* d) {
  if (!d) return 0;
  return carve_persist(*d, nullptr).bytes;
}

int aa_decode_persistent_supported(const aa_dims* d) {
  if (!d) return 0;
  return decode_persist_supported(d->B, d->k, d->a, d->H, d->E, d->Vc) ? 1 : 0;
}

int aa_decode_persistent(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const float* h0, const float* c0,
                         int max_len, int64_t* ids, float* attention, float* Beta, int flags, int* candidates_out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  AA_TRY(check_decode(d, w, max_len, 1));
  AA_REQUIRE(V && v_g && ids && attention && Beta, "aa_decode_persistent: null pointer");
  if (d->B == 0) return AA_OK;
  if (!decode_persist_supported(d->B, d->k, d->a, d->H, d->E, d->Vc)) {
    set_error("aa_decode_persistent: B=%d k=%d H=%d does not fit one image per SM with V resident in shared memory; use aa_greedy_decode",
              d->B, d->k, d->H);
    return AA_ERR_UNSUPPORTED;
  }
  aa_dims dd = *d;
  dd.T = max_len;
  const size_t need = carve_persist(dd, nullptr).bytes;
  if (!workspace || workspace_bytes < need) {
    set_error("aa_decode_persistent: workspace too small (%zu < %zu)", workspace_bytes, need);
    return AA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PersistWs ws = carve_persist(dd, workspace);
  const int B = d->B, H = d->H, E = d->E, L = max_len;
  if (!(flags & AA_DECODE_REUSE_PACKED_WEIGHTS)) {     // weight-derived operands: once per set of weights
    AA_TRY(launch_split_tf32(w->w_hh, H, 4 * H, H, ws.Whh, ws.Kp, st));
    AA_TRY(launch_cast2d(w->mlp_w, H, ws.Wp16, H, d->Vc, H, st));
    AA_TRY(launch_row_norm(w->mlp_w, d->Vc, H, ws.wn, ws.dwn, st));
    // EG[v] = [W_ih[:, :E]; W_x[:, :E]] emb(v): the input half of the five gate blocks for every word, exact fp32
    AA_TRY(gemm_nt(d->Vc, 4 * H, E, w->embed, E, w->w_ih, 2 * E, ws.EG, 5 * H, nullptr, 0, nullptr, nullptr, st));
    if (w->sen_wx) AA_TRY(gemm_nt(d->Vc, H, E, w->embed, E, w->sen_wx, 2 * E, ws.EG + 4 * H, 5 * H, nullptr, 0, nullptr, nullptr, st));
    else AA_CHECK_CUDA(cudaMemset2DAsync(ws.EG + 4 * H, sizeof(float) * 5 * H, 0, sizeof(float) * H, (size_t)d->Vc, st));
  }
  // per call: P = V W_v^T, the static (v_g, bias) gate terms, the initial operand rows and cell state -- exact fp32
  if (ws.ldP != d->a) AA_CHECK_CUDA(cudaMemsetAsync(ws.P, 0, sizeof(float) * (size_t)B * d->k * ws.ldP, st));
  AA_TRY(gemm_nt(B * d->k, d->a, H, V, H, w->att_wv, H, ws.P, ws.ldP, nullptr, 0, nullptr, nullptr, st));
  AA_TRY(gemm_nt(B, 4 * H, E, v_g, E, w->w_ih + E, 2 * E, ws.stat, 5 * H, nullptr, 0, w->b_ih, w->b_hh, st));
  if (w->sen_wx) AA_TRY(gemm_nt(B, H, E, v_g, E, w->sen_wx + E, 2 * E, ws.stat + 4 * H, 5 * H, nullptr, 0, nullptr, nullptr, st));
  else AA_CHECK_CUDA(cudaMemset2DAsync(ws.stat + 4 * H, sizeof(float) * 5 * H, 0, sizeof(float) * H, (size_t)B, st));
  persist_init_kernel<<<B, 256, 0, st>>>(h0, c0, ws.hA, ws.c, H, ws.Kp);
  AA_CHECK_LAUNCH("persist_init");
  DecodePersistArgs p{};
  p.B = B; p.k = d->k; p.a = d->a; p.H = H; p.E = E; p.Vc = d->Vc; p.L = L;
  p.K1p = ws.Kp; p.lo1 = ws.Kp; p.ldA = ws.ldA; p.ldP = ws.ldP; p.ldv = ws.ldv; p.ks1_max = ws.ks_max; p.start_id = START_ID;
  p.V = V; p.P = ws.P; p.stat = ws.stat; p.c0 = ws.c; p.EG = ws.EG; p.hA = ws.hA; p.part1 = ws.part1; p.u16 = ws.u16; p.approx = ws.approx;
  p.Wg = w->att_wg; p.Ws = w->att_ws; p.wh = w->att_wh; p.Wp = w->mlp_w; p.bp = w->mlp_b; p.wn = ws.wn; p.dwn = ws.dwn;
  p.ids = reinterpret_cast<long long*>(ids); p.alpha = attention; p.beta = Beta; p.ncand_out = candidates_out; p.bar = ws.bar;
  AA_PROF("dec_persistent", st, launch_decode_persist(p, ws.Whh, ws.Wp16, st));
  return AA_OK;
}

int aa_beam_decode(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const float* h0, const float* c0,
                   int beam, int max_len, int64_t* ids, float* attention, float* Beta, float* score, void* workspace,
                   size_t workspace_bytes, void* stream) {
  AA_TRY(check_decode(d, w, max_len, beam));
  AA_REQUIRE(V && v_g && ids && attention && Beta, "aa_beam_decode: null pointer");
  if (d->B == 0) return AA_OK;
  aa_dims dd = *d;
  dd.T = max_len;
  if (!workspace || workspace_bytes < aa_decode_workspace_bytes(&dd, beam)) {
    set_error("aa_beam_decode: workspace too small (%zu < %zu)", workspace_bytes, aa_decode_workspace_bytes(&dd, beam));
    return AA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DecodeWs ws = carve_decode(dd, beam, workspace);
  const int B = d->B, H = d->H, E = d->E, Vc = d->Vc, k = d->k, L = max_len, R = B * beam;
  AA_TRY(decode_prologue(dd, *w, V, v_g, h0, c0, beam, ws, st));
  beam_init_kernel<<<ceil_div(R, 256), 256, 0, st>>>(ws.cum, ws.done, R, beam);
  AA_CHECK_LAUNCH("beam_init");
  float* Acur = ws.Acat;  float* Anext = ws.Acat2;
  float* ccur = ws.c;     float* cnext = ws.c2;
  for (int t = 0; t < L; ++t) {
    AA_TRY(decode_step_body(dd, *w, ws, V, Acur, ccur, R, beam, ws.rec_alpha + (size_t)t * R * k, k, ws.rec_beta + (size_t)t * R, 1, st));
    AA_PROF("dec_vocab_gemm", st, dec_gemm(ws, R, Vc, H, ws.Hp, ws.u, ws.ldU, ws.Hp, ws.ldU, ws.split ? ws.Wp_s : w->mlp_w,
                                           ws.split ? 2 * ws.Hp : H, ws.logits, Vc, nullptr, 0, w->mlp_b, nullptr, nullptr, st));
    row_lse_kernel<<<R, 256, 0, st>>>(ws.logits, Vc, ws.row_max, ws.row_lsum);
    AA_CHECK_LAUNCH("row_lse");
    beam_select_kernel<<<B, 256, 0, st>>>(ws.logits, ws.row_max, ws.row_lsum, ws.cum, ws.done, Vc, beam, H, E, w->embed, Acur,
                                          ccur, Anext, cnext, ws.rec_word + (size_t)t * R, ws.rec_src + (size_t)t * R,
                                          ws.rec_wasdone + (size_t)t * R, ws.split, ws.lo, ws.ldA);
    AA_CHECK_LAUNCH("beam_select");
    float* tmp = Acur; Acur = Anext; Anext = tmp;
    tmp = ccur; ccur = cnext; cnext = tmp;
  }
  beam_backtrack_kernel<<<B, 64, 0, st>>>(ws.rec_word, ws.rec_src, ws.rec_wasdone, ws.rec_alpha, ws.rec_beta, ws.cum, L, R, beam,
                                          k, reinterpret_cast<long long*>(ids), attention, Beta, score);
  AA_CHECK_LAUNCH("beam_backtrack");
  return AA_OK;
}

}  // extern "C"
