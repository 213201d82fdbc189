// Internal launcher declarations (one per kernel family).  All launchers are asynchronous on
// `stream`, return AA_OK / error code and never allocate.
#pragma once
#include "common.cuh"

namespace aa {

// ---- pointwise.cu ----------------------------------------------------------------------
// x[b,t,:] = [embed[cap[b,t]] ; v_g[b]]                      (baseline_attention.py:151-154)
// (every *16 argument is an optional bf16 mirror of the fp32 output, same strides; null in fp32 mode)
int launch_build_x(const long long* cap, const float* embed, const float* v_g, float* x, __nv_bfloat16* x16, int B, int T, int E,
                   int Vc, cudaStream_t s, float* zrow32 = nullptr, __nv_bfloat16* zrow16 = nullptr, int H = 0);   // zrow*: [B,T,H] arrays whose t = 0 rows are zero-filled on the way
// one LSTM cell update from pre-activations (gate order i,f,g,o)   (baseline_attention.py:172)
//   pre [B,4H] row stride ld_pre; c_prev [B,H] stride ld_cprev;
//   writes acts[b, t, 4, H] (post-activation), cells[b,t,:], hiddens[b,t,:], hs_next (= h, may be null)
int launch_lstm_cell_fwd(const float* pre, long long ld_pre, const float* c_prev, long long ld_cprev, float* acts,
                         long long ld_acts, float* c_out, long long ld_c, float* h_out, long long ld_h, float* hs_next,
                         long long ld_hs, __nv_bfloat16* h16, __nv_bfloat16* hs_next16, int B, int H, cudaStream_t s);
// sentinel: g = sigmoid(pre) ; s = g * tanh(cell)                   (adaptive_attention.py:79-83)
int launch_sentinel_fwd(const float* pre, const float* cells, float* g, float* s_out, __nv_bfloat16* s16, long long n,
                        cudaStream_t s);
// sentinel backward: da = ds*tanh(c)*g*(1-g) ; dcell = ds*g*(1-tanh(c)^2)
int launch_sentinel_bwd(const float* ds, const float* g, const float* cells, float* da, float* dcell, __nv_bfloat16* da16,
                        long long n, cudaStream_t s);
// one BPTT step of the LSTM cell.  dh_attn/dhs_next/dcell are the batched (non-recurrent) contributions.
int launch_lstm_cell_bwd(const float* dh_attn, long long ld_dh, const float* dhs_next, long long ld_dhs,
                         const float* dh_rec, const float* dcell, long long ld_dcell, const float* dc_rec,
                         const float* acts, long long ld_acts, const float* cells, long long ld_c, const float* c_prev,
                         long long ld_cprev, float* dgates, long long ld_dg, __nv_bfloat16* dgates16, float* dc_out, int B,
                         int H, cudaStream_t s);
// out[n] = sum_m X[m, n]   (column sums; bias gradients).  out2 (optional) receives a copy.
int launch_colsum(const float* X, long long ldx, int M, int N, float* out, float* out2, cudaStream_t s);
// same, and writes the bf16 mirror X16 (row stride ld16) of X in the same pass when X16 != null
int launch_colsum_cast(const float* X, long long ldx, int M, int N, float* out, float* out2, __nv_bfloat16* X16, long long ld16,
                       cudaStream_t s);
// dE[cap[row], :] += dx[row, 0:E]  (dE pre-zeroed) ; dvg[b, :] = sum_t dx[b,t,E:2E]
int launch_embed_bwd(const long long* cap, const float* dx, float* dE, float* dvg, int B, int T, int E, int Vc,
                     cudaStream_t s);
int launch_add_inplace(float* y, const float* x, long long n, cudaStream_t s);
// row-wise arg-max of logits [B,Vc] (lowest index wins ties); writes ids_out[b*ld_ids] (int64) and, if
// emb_dst != null, gathers embed[id] into emb_dst[b*ld_emb + 0:E] for the next decode step
int launch_argmax_gather(const float* logits, long long ld_logits, int B, int Vc, long long* ids_out, long long ld_ids,
                         const float* embed, int E, float* emb_dst, long long ld_emb, cudaStream_t s);
// reduce the vocabulary GEMM's arg-max partials [B, tiles_n] to ids and gather embed[id] into the next A operand
// (split != 0: as tf32 hi at [0,E) and lo at [lo_off, lo_off+E) of the destination row)
// (ncand != null: row b holds ncand[b] valid partials at the front of its tiles_n entries)
int launch_argmax_finalize(const float* pmax, const int* pidx, int tiles_n, int B, long long* ids_out, long long ld_ids,
                           const float* embed, int E, float* emb_dst, long long ld_emb, int split, long long lo_off, cudaStream_t s,
                           const int* ncand = nullptr);
// dst [rows, 2*Kp] = [tf32 hi | lo] of src [rows, cols] (zero padded to Kp, a multiple of 32)
int launch_split_tf32(const float* src, long long ld_src, long long rows, int cols, float* dst, int Kp, cudaStream_t s);
int launch_gather_rows(const long long* ids, long long ld_ids, const float* table, int E, int Vc, float* dst, long long ld_dst,
                       int B, cudaStream_t s);
int launch_copy2d(float* dst, long long ld_dst, const float* src, long long ld_src, int rows, int cols, cudaStream_t s);
// fp32 -> bf16 casts: 2-D with row strides, and a one-launch multi-segment cast (weights)
int launch_cast2d(const float* src, long long ld_src, __nv_bfloat16* dst, long long ld_dst, long long rows, int cols,
                  cudaStream_t s);
struct CastSegs {
  const float* src[8];
  __nv_bfloat16* dst[8];
  long long n[8];
};
// (max_blocks_per_seg > 0 caps the grid: background casts that must not crowd the SMs a critical kernel is about to need)
int launch_cast_multi(const CastSegs& segs, int nsegs, cudaStream_t s, int max_blocks_per_seg = 0);
// mean cross-entropy over rows + gradient: loss += -log_softmax(logits[r])[tgt[r]] / denom ; dlogits = (softmax - onehot)/denom
// (denom <= 0: the row count n)
int launch_ce_fwd_bwd(const float* logits, long long ld, const long long* tgt, int n, int Vc, float* loss, float* dlogits,
                      long long ldd, long long denom, cudaStream_t s, __nv_bfloat16* dlogits16 = nullptr, int* mirror_written = nullptr);

// x[0:n] *= *g unless *g == 1 (device scalar)
int launch_scale_unless_one(float* x, const float* g, long long n, cudaStream_t s, __nv_bfloat16* x16 = nullptr);
// fused cross-entropy (see TcGemmArgs::ce_*): merge the chunk partials of every row into its log-sum-exp, add the row's loss term
// to *loss, and leave scale[row * chunks + chunk] = exp(m_chunk - lse) / denom ...
// x[0:n] (bf16) *= *g and y[0:ny] (fp32) *= *g unless *g == 1 (decided on the device)
int launch_scale_bf16_unless_one(__nv_bfloat16* x, const float* g, long long n, float* y, long long ny, cudaStream_t s);
int launch_ce_merge(const float* part, int chunks, int n_rows, const float* xt, long long denom, float* loss, float* scale, cudaStream_t s);
// ... then dlogits (bf16, in place over e) = e * scale - onehot / denom, and the bias gradient dbp[j] = sum over rows (dbp pre-zeroed)
int launch_ce_fixup(__nv_bfloat16* e16, long long ld, int n_rows, int Vc, const float* scale, int chunks, const long long* tgt, long long denom,
                    float* dbp, cudaStream_t s, const float* loss_acc = nullptr, float* loss_out = nullptr);   // (*loss_out = *loss_acc on the way)
// zero-fill of up to 8 buffers in one launch (null / empty entries are skipped)
int launch_zero_multi(int nsegs, void* const* dst, const long long* bytes, cudaStream_t s);
// up to 8 device-to-device copies in one launch
int launch_copy_multi(int nsegs, const void* const* src, void* const* dst, const long long* bytes, cudaStream_t s);

// ---- gemm_tc.cu (tcgen05 + TMA) ------------------------------------------------------------
// D[m,n] = sum_k A(m,k) B(n,k) (+ beta*Cin + bias1[n] + bias2[n]); bf16 (elem_size 2) or tf32 (4) inputs,
// fp32 accumulation.  a_mn/b_mn = 0: operand stored [rows, K] with K contiguous (K-major);
// = 1: stored [K, rows] with rows contiguous (MN-major).  Output fp32 (D32) and/or bf16 (D16).
struct TcGemmArgs {
  int M, N, K;
  const void* A; long long lda; int a_mn;
  const void* B; long long ldb; int b_mn;
  int elem_size;
  float* D32; long long ldd32;
  __nv_bfloat16* D16; long long ldd16;
  const float* Cin; long long ldcin; float beta;
  const float* bias1; const float* bias2;
  // split3 != 0: fp32-accurate "3xTF32" mode.  A is [M, 2K], B is [N, 2K], each row = [hi | lo] with
  // hi = tf32(x), lo = tf32(x - hi) (see launch_split_tf32), K = padded length of one half (multiple of 32).
  int split3;
  // split3 only, all optional (0 = default): column offset of the lo half inside a row (default K) and number of
  // valid columns of a row counted from the operand's base pointer (default lo + K) -- lets an operand be a column
  // window of a wider split row, e.g. [emb | h | s] rows serving both the gate and the q/r contraction
  int lo_a, lo_b; long long a_cols, b_cols;
  // optional (split3 only): per-(row, n-tile) arg-max partials of D incl. bias, layout [M, ceil(N / tile_n)] with
  // tile_n = gemm_tc_argmax_tile_n(N); D32/D16 may then both be null (the logits are never written).
  // pidx == null: maxima only (cheaper epilogue; bias2 must be null)
  float* pmax; int* pidx;
  // optional: the rows n >= kcut_n0 of B are zero beyond their first kcut_cols reduction columns (per half in split mode): output
  // tiles starting at or after kcut_n0 stop their K loop there.  Ignored unless kcut_n0 is a multiple of the tile width; no split-K.
  int kcut_n0, kcut_cols;
  // optional (split3 only): output tiles starting at or after column ashift_n0 read their A operand ashift_cols columns further into
  // the row (both halves) -- two contractions over different column windows of the same operand rows in one launch, e.g.
  // [q | r'] = [h W_g^T | s W_s^T] from rows [h | s].  Needs ashift_n0 to be a multiple of the tile width (forces 64-column tiles).
  int ashift_n0, ashift_cols;
  // optional (split3 only): A is a PLAIN fp32 array [M, K] (row stride lda) that the kernel splits into tf32 (hi, lo) on the fly in
  // shared memory; B stays pre-split.  N <= 64.
  int a_raw;
  // optional (plain bf16, K-major): cross-entropy pieces instead of the logits (train.py:63,208 fused into the vocabulary projection's
  // epilogue; D32 / D16 null).  Per row and 32-column chunk of D (+ bias1): the chunk maximum m and s = sum exp(x - m) go to
  // ce_part[(row * ce_chunks + chunk) * 2 + {0, 1}], e = exp(x - m) as bf16 to ce_e16[row * ld_ce + col], and the target column's
  // logit x[ce_tgt[row]] to ce_xt[row].  ce_chunks = ceil(N / 32).
  __nv_bfloat16* ce_e16; long long ld_ce; float* ce_part; int ce_chunks; const long long* ce_tgt; float* ce_xt;
};
int launch_gemm_tc(const TcGemmArgs& g, cudaStream_t s);
int gemm_tc_argmax_tile_n(int N);
int gemm_tc_argmax_tile_n_plain(int N);   // same for a plain (non-split) contraction with arg-max partials
int set_gemm_pair(int on);     // diagnostics: 0 = single-CTA kernels only, 1 = CTA-pair kernels where they apply, < 0 = the environment's choice (AA_GEMM_PAIR, default on)
int set_gemm_splitk(int on);   // diagnostics: 0 = never split K (deterministic summation order)

// ---- vocab_refine.cu (filter-and-refine arg-max of the vocabulary projection, greedy decoding) ----
bool argmax_refine_supported(int Vc, int H);
int refine_units(long long* out4, int reset);   // diagnostics: CTA units, warp units, tiles refined by CTAs, tiles refined by warps (accumulated)
long long refine_pairs(int reset);   // diagnostics: (row, tile) pairs refined so far on the current device (synchronises)
// wnorm[t] = max_j ||W[j,:]||_2 over the 16-column tile t of W [Vc,H] (tile width = gemm_tc_argmax_tile_n_plain)
int launch_tile_wnorm(const float* W, int Vc, int H, float* wnorm, cudaStream_t s, float* dwnorm = nullptr);   // dwnorm: norms of W - bf16(W) per tile
// pmax [R, tiles]: approximate per-tile maxima (single-pass tensor-core contraction, |error_j| <= c ||u|| ||W_j||).  Every (row, tile)
// pair whose tile can hold the row's exact arg-max is appended to list[t*R + counts[t]++] as row | slot << 20, slot = rank among
// the row's candidates; ncand[row] = number of candidates  (u rows: hi at [0,H), lo at +lo_off)
int launch_argmax_filter(const float* pmax, int tiles, int R, const float* u, long long ldu, long long lo_off, int H, const float* wnorm, float c,
                         int* counts, unsigned* list, int* ncand, cudaStream_t s, const __nv_bfloat16* u16 = nullptr, long long ld16 = 0,
                         const float* dwnorm = nullptr);   // (u16, dwnorm: bf16 first pass -- exact-decomposition bound, see vocab_refine.cu)
// exact fp32 (max, index) of every listed pair written to pmax / pidx[row * tiles + slot]; resets counts
int launch_argmax_refine(const float* W, const float* bias, int Vc, int H, const float* u, long long ldu, long long lo_off, int R, int* counts,
                         const unsigned* list, float* pmax, int* pidx, int tiles, cudaStream_t s);

// ---- lstm_seq.cu (persistent recurrence, bf16 tensor-core mode) -----------------------------
// true when the one-launch recurrence kernels can run this shape (H % 64 == 0 and the CTAs fit the chip)
bool lstm_seq_supported(int B, int H, int* units_fwd);
struct LstmSeqFwd {
  int B, T, H;
  const float* w_hh;                  // [4H,H] fp32 master weights (re-packed to bf16 inside)
  const float* xg;                    // [B,T,4H] input-half gate pre-activations incl. biases
  const float* c0;                    // [B,H] or null
  const __nv_bfloat16* h016;          // [B,H] bf16 initial hidden state
  float *hiddens, *cells, *acts, *hs_prev;
  __nv_bfloat16 *hid16, *hsprev16;
  __nv_bfloat16* whh_packed16;        // scratch [4H*H]
  unsigned* counters;                 // scratch [ceil(B/128)]
  const __nv_bfloat16* whh16;         // optional: plain bf16 copy of w_hh [4H,H] (operand of the cluster kernels)
};
int launch_lstm_seq_fwd(const LstmSeqFwd& p, cudaStream_t s);
struct LstmSeqBwd {
  int B, T, H;
  const float* w_hh;
  const float *dh_attn, *dhs, *dcell, *d_hT, *d_cT, *acts, *cells, *c0;
  float* dgates; __nv_bfloat16* dgates16;
  float *dh0, *dc0;
  __nv_bfloat16* whhT16;              // scratch [H*4H]
  unsigned* counters;
  const __nv_bfloat16* whh16;         // optional: plain bf16 copy of w_hh [4H,H] (operand of the cluster kernels)
  __nv_bfloat16* dgates16_t0;         // optional [B,4H]: second copy of the step-0 rows of dgates16 (cluster kernel only)
};
int launch_lstm_seq_bwd(const LstmSeqBwd& p, cudaStream_t s);
int set_seq_trace_buffer(void* dev_ptr);
int set_bptt_ksplit_max(int ks);   // diagnostics: cap on the K-split (cluster size) of the BPTT kernel; 1 = no cluster   // diagnostics: [steps][8] uint64 globaltimer stamps of CTA (0,0); null = off

// ---- lstm_cluster.cu (same recurrences inside thread-block clusters, weights in tensor memory; H in {128,256,512}) ----
// Both return AA_ERR_UNSUPPORTED when the shape / device cannot run them: the caller then takes the lstm_seq.cu kernels.
bool lstm_cluster_supported(int B, int H);
int launch_lstm_cluster_fwd(const LstmSeqFwd& p, cudaStream_t s);
int launch_lstm_cluster_bwd(const LstmSeqBwd& p, cudaStream_t s);
int set_clk_trace_buffer(void* dev_ptr);
int set_lstm_cluster(int on, int nacc);   // diagnostics: on = 0 never takes the cluster kernels; nacc = accumulators of the forward chain

// ---- atten.cu --------------------------------------------------------------------------
struct AttenFwdArgs {
  int B, T, k, a, H;
  const float *P, *q, *r, *s, *h, *V, *wh;   // P[B,k,a] q,r[B,T,a] s,h[B,T,H] V[B,k,H] wh[a]
  float *alpha, *beta, *ctx, *u;            // alpha[B,T,k] beta[B,T] ctx[B,T,H] (may be null) u[B,T,H] = c_hat + h
  float* c_hat;                             // optional [B,T,H]
  __nv_bfloat16* u16;                       // optional bf16 mirror of u
  int no_sentinel;                          // != 0: the baseline model's block (baseline_attention.py:79-100): beta = 0, c_hat = ctx;
                                            // r and s must still point at finite (zero-filled) rows
};
int launch_atten_fwd(const AttenFwdArgs& p, cudaStream_t s);

struct AttenBwdArgs {
  int B, T, k, a, H;
  const float *P, *q, *r, *s, *V, *wh, *alpha, *beta, *ctx;
  const float* dchat;     // [B,T,H]
  const float* d_alpha;   // optional [B,T,k]
  const float* d_beta;    // optional [B,T]
  float *ds;              // [B,T,H]  = beta * dchat         (sentinel part added later by GEMM)
  float *dq, *dr;         // [B,T,a]  dq = sum_i dp_i + dr
  float *dP;              // [B,k,a]  summed over t
  float *dV;              // [B,k,H]  sum_t alpha_i dctx     (dP W_v added later by GEMM)
  float *dwh;             // [a]      pre-zeroed, atomically accumulated
  __nv_bfloat16 *dq16, *dr16, *dP16;   // optional bf16 mirrors with row stride a_pad (dP16 only when not split over T)
  int a_pad;
  int prezeroed;          // != 0: the caller has already zero-filled dV and dP (off the critical path)
};
int launch_atten_bwd(const AttenBwdArgs& p, cudaStream_t s);

// ---- decode.cu -------------------------------------------------------------------------
struct DecodeStepArgs {
  int B, k, a, H;
  int beam;                   // rows per image (1 = greedy); V/P row = b / beam
  const float* gates;         // [B,5H] pre-activations: i,f,g,o,sentinel
  float* c;                   // [B,H]  in: c_{t-1}  out: c_t
  float* h_out; long long ld_h;   // h_t destination (A-operand of the next gate GEMM)
  const float *P, *V;         // [B/beam,k,a], [B/beam,k,H]
  const float *Wg, *Ws, *wh;  // [a,H],[a,H],[a]   (Ws == null: baseline model without the sentinel, beta = 0)
  float* alpha; long long ld_alpha;   // alpha[b*ld_alpha + i]
  float* beta; long long ld_beta;
  float* u; long long ld_u;   // u = c_hat + h, row stride ld_u (0 = H)
  // split != 0: h and u are written as tf32 (hi, lo) pairs for the 3xTF32 tensor-core contractions: hi at the
  // plain position, lo at +h_lo_off / +u_lo_off in the same row
  int split; long long h_lo_off, u_lo_off;
};
int launch_decode_step(const DecodeStepArgs& p, cudaStream_t s);

// tensor-core decode pipeline (AA_PREC_TF32X3): LSTM pointwise + sentinel gate ...
struct DecodeCellArgs {
  int R, H;
  const float* gates;         // [R,5H] pre-activations: i,f,g,o,sentinel
  float* c;                   // [R,H] in/out
  float* hs;                  // [R,2H] fp32 [h_t | s_t]
  float* A; long long ldA;    // A-operand rows: h (hi) at [h_off, h_off+H), s (hi) at [h_off+H, h_off+2H), lo halves at +lo_off
  long long h_off, lo_off;
  // table mode (EG != null): the input half of the gates is not in `gates` but a row of the per-word table
  // EG [Vc,5H] = embed . [W_ih[:, :E]; W_x[:, :E]]^T; `gates` then holds the recurrent half + static term of the four LSTM blocks
  // (columns [0,4H) of rows of 5H) and the sentinel block's pre-activation is stat[r, 4H:] + EG[word, 4H:]
  const float* EG; const float* stat;
  const long long* prev_ids; long long ld_ids;   // word fed at this step: prev_ids[r * ld_ids] (null: start_id for every row)
  int start_id;
};
int launch_decode_cell(const DecodeCellArgs& p, cudaStream_t s);
// ... and the attention stream: scores + both softmaxes + beta-gated context + (c_hat + h), one warp per row
struct DecodeAttenArgs {
  int R, k, a, H, beam;       // R rows; V/P row = r / beam
  const float* qr; long long ld_qr;   // [R, ld_qr >= 2a] = [q | r]   (q = h W_g^T, r = s W_s^T + q)
  int r_off, r_partial;               // r starts r_off columns into a row (0 = a); r_partial != 0: the row holds r' = s W_s^T, r = r' + q
  const float* hs;            // [R,2H] = [h | s]
  const float *P, *V, *wh;    // [R/beam,k,ldP >= a], [R/beam,k,H], [a]
  long long ldP;              // row stride of P (a multiple of 4 enables the bulk-copy pipeline)
  int force_simple;           // != 0: always take the register-staged kernel (tests)
  int no_sentinel;            // != 0: baseline model (no sentinel): beta = 0, c_hat = ctx
  float* alpha; long long ld_alpha;
  float* beta; long long ld_beta;
  float* u; long long ld_u, u_lo_off;    // u = c_hat + h as tf32 (hi, lo): hi at [0,H), lo at [u_lo_off, u_lo_off+H)
  __nv_bfloat16* u16; long long ld_u16;  // optional bf16 mirror of u (ld_u16 % 4 == 0)
};
int launch_decode_atten(const DecodeAttenArgs& p, cudaStream_t s);

// ---- decode_persist.cu: the whole greedy loop in one cooperative launch, one image per SM, V / P / c resident on chip ----
struct DecodePersistArgs {
  int B, NB, k, a, H, E, Vc, L;     // NB: batch padded to the MMA's N (filled by the launcher)
  int K1p, lo1, ldA;                // G1: padded K of W_hh's halves (= column of hA's lo half), hA row stride (floats)
  int ldP, ldv;                     // row strides of P and of the approximate logits
  int nkb1, ks1, kbper1, ks1_max, nkb2;   // tiling (filled by the launcher; ks1_max = K ranges the partial buffer has room for)
  int start_id;                     // <start> token (adaptive_attention.py:188)
  const float *V, *P, *stat, *c0;   // [B,k,H], [B,k,ldP], [B,5H] static gate terms (v_g half + biases), [B,H] (null = zeros)
  const float* EG;                  // [Vc,5H] input-half gate table: embed . [W_ih[:, :E]; W_x[:, :E]]^T
  float* hA;                        // [B, ldA] G1's operand rows: h as tf32 (hi | lo); initialised with h0
  float* part1;                     // [ks1_max, B, 4H] K-split partials of G1
  __nv_bfloat16* u16;               // [B, H] bf16 mirror of u (G2's operand)
  float* approx;                    // [B, ldv] approximate logits (bias included)
  const float *Wg, *Ws, *wh;        // attention weights (Ws == null: baseline model, beta = 0)
  const float *Wp, *bp, *wn, *dwn;  // vocabulary projection (fp32), its bias, the norms of its rows and of their bf16 residuals
  long long* ids; float* alpha; float* beta;   // [B,L], [B,L,k], [B,L]
  int* ncand_out;                   // optional [B,L]: columns recomputed exactly (diagnostics)
  unsigned* bar;                    // grid barrier counter
};
bool decode_persist_supported(int B, int k, int a, int H, int E, int Vc);
int decode_persist_nb(int B);
int launch_row_norm(const float* W, int rows, int cols, float* wn, float* dwn, cudaStream_t s);   // row norms of W and of W - bf16(W)
int set_persist_trace_buffer(void* dev_ptr);   // diagnostics: [steps][8] uint64 globaltimer stamps of CTA 0; null = off
int launch_decode_persist(const DecodePersistArgs& p, const float* Whh_split, const __nv_bfloat16* Wp16, cudaStream_t s);

}  // namespace aa
