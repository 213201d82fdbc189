/* adaptive_b200.h — C ABI of libadaptive_sm100.so
 *
 * B200 (sm_100a) implementation of the caption-decoder hot path of wzn0828/Adaptive
 * ("Knowing When to Look": LSTM step + visual sentinel + adaptive attention + vocabulary
 * projection) for teacher-forced training (forward/backward) and greedy / beam decoding.
 *
 * The reference has no FFI of its own: its boundary for this path is the PyTorch nn.Module
 * surface.  Each entry point below names the reference code it
 * replaces (paths relative to the reference checkout, file:line).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all memory,
 *     including workspaces (sizes from the aa_*_bytes queries); nothing is allocated, freed
 *     or retained by the library;
 *   - all tensors are dense, row-major, contiguous, fp32 ("float") unless the name says bf16;
 *     token ids are int64 ("long long");
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never
 *     synchronises the device; it is safe to capture the calls into a CUDA graph;
 *   - return value: 0 = ok, non-zero = error (AA_ERR_*); aa_last_error() returns a
 *     thread-local message for the last failing call on the calling thread;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef ADAPTIVE_B200_H_
#define ADAPTIVE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AA_OK 0
#define AA_ERR_INVALID 1      /* bad argument (shape, null pointer, alignment) */
#define AA_ERR_CUDA 2         /* a CUDA runtime call or kernel launch failed */
#define AA_ERR_UNSUPPORTED 3  /* valid request the library has no kernel for */
#define AA_ERR_WORKSPACE 4    /* workspace too small */

/* Problem shape.  B images (rows), T teacher-forced steps, k regions per image, a = inner
 * attention dim (49 in the reference regardless of k: adaptive_attention.py:16-19), H LSTM
 * hidden size, E word-embedding size (LSTM input is 2E), Vc vocabulary size. */
#define AA_PREC_FP32 0   /* exact fp32 everywhere (SIMT contractions): the parity path */
#define AA_PREC_BF16 1   /* mixed precision: bf16 operands on tcgen05 tensor cores, fp32 accumulation,
                            fp32 master weights / state / softmax / gradients (teacher-forced path only) */
#define AA_PREC_TF32X3 2 /* decoding: the per-step contractions run on tcgen05 as 3xTF32 over (hi, lo)-split fp32 operands
                            (fp32-accurate to ~1e-6 relative; fp32 accumulate); everything else exact fp32 */
typedef struct aa_dims {
  int32_t B, T, k, a, H, E, Vc;
  int32_t precision;   /* AA_PREC_* */
} aa_dims;

/* Decoder parameters, named after the reference state_dict keys:
 *   decoder.embed.weight [Vc,E]                              baseline_attention.py:137
 *   decoder.LSTM.{weight_ih_l0 [4H,2E], weight_hh_l0 [4H,H], bias_ih_l0, bias_hh_l0 [4H]}  :140
 *   decoder.adaptive.sentinel.{affine_x [H,2E], affine_h [H,H]}.weight   adaptive_attention.py:66-67
 *   decoder.adaptive.atten.{affine_v, affine_g, affine_s [a,H], affine_h [1,a]}.weight      :16-19
 *   decoder.adaptive.mlp.{weight [Vc,H], bias [Vc]}                                         :100   */
typedef struct aa_weights {
  const float* embed;
  const float* w_ih;
  const float* w_hh;
  const float* b_ih;
  const float* b_hh;
  const float* sen_wx;
  const float* sen_wh;
  const float* att_wv;
  const float* att_wg;
  const float* att_ws;
  const float* att_wh;
  const float* mlp_w;
  const float* mlp_b;
} aa_weights;

/* Gradients w.r.t. aa_weights, same shapes.  Every buffer is OVERWRITTEN (not accumulated). */
typedef struct aa_weight_grads {
  float* embed;
  float* w_ih;
  float* w_hh;
  float* b_ih;
  float* b_hh;
  float* sen_wx;
  float* sen_wh;
  float* att_wv;
  float* att_wg;
  float* att_ws;
  float* att_wh;
  float* mlp_w;
  float* mlp_b;
} aa_weight_grads;

/* ---- library ------------------------------------------------------------------------ */
int aa_version(void);                 /* major*10000 + minor*100 + patch */
const char* aa_last_error(void);      /* thread-local, never NULL */
int aa_device_info(int* sm_count, int* cc_major, int* cc_minor);   /* host out-params */

/* Instrumentation (host-side; no reference counterpart).  aa_launch_count: kernels launched by
 * this library in this process so far.  aa_profile_*: when enabled, the major kernels are
 * bracketed by cudaEvent pairs on their own stream; aa_profile_count() waits for the recorded
 * events and returns the number of distinct kernel tags, aa_profile_get(i, ...) their totals. */
long long aa_launch_count(void);
int aa_profile_enable(int on);
int aa_profile_reset(void);
int aa_profile_count(void);
int aa_profile_get(int i, char* name, int name_len, double* total_ms, int* launches);
/* Diagnostics: device buffer [steps][8] of uint64 %globaltimer stamps written by CTA (0,0) of the persistent
 * LSTM kernels (see lstm_seq.cu); NULL switches tracing off. */
int aa_debug_set_trace_buffer(void* dev_ptr);
/* Diagnostics: device buffer [max_len][8] of uint64 %globaltimer stamps written by CTA 0 of the persistent decoder
 * (decode_persist.cu lists the events); NULL switches tracing off. */
int aa_debug_set_persist_trace(void* dev_ptr);
/* Diagnostics: on != 0 makes the tensor-core decode pipeline use the register-staged attention kernel instead of the
 * bulk-copy (cp.async.bulk + mbarrier ring) one; both compute the same step (tests compare them). */
int aa_debug_set_decode_atten_simple(int on);
/* Diagnostics: on == 0 makes the greedy sampler compute every logit with the fp32-accurate 3xTF32 contraction instead of the
 * filter-and-refine arg-max (one low-precision tensor-core pass + exact fp32 logits of the columns that can hold the maximum);
 * 1 = tf32 first pass, 2 = bf16 first pass (default). */
int aa_debug_set_decode_argmax_refine(int on);
/* Diagnostics: number of (row, 16-column tile) pairs the filter has handed to the exact refinement on the current device since
 * the last reset (synchronises the device); -1 on error. */
long long aa_debug_refine_pairs(int reset);
/* Diagnostics: how the refinement dealt its work since the last reset -- out4 = {CTA units (tile, <= 32 rows), warp units
 * (tile, <= 2 rows), tiles refined by CTA units, tiles refined by warp units}, summed over its launches (synchronises). */
int aa_debug_refine_units(long long* out4, int reset);
/* Diagnostics: on != 0 makes the training attention use the step-by-step kernels instead of the step-parallel ones. */
int aa_debug_set_atten_sequential(int on);
/* Diagnostics: on == 0 stops the tcgen05 GEMM from splitting K (partial tiles summed with red.global.add, i.e. a
 * run-to-run varying fp32 summation order in the affected gradients); default on. */
int aa_debug_set_gemm_splitk(int on);
/* Diagnostics: the CTA-pair contraction kernels (tcgen05.mma.cta_group::2, 256 x 256 tiles over the two SMs of a TPC; they serve the
 * decode step's gate contraction and the maxima pass of its vocabulary arg-max).  0 = single-CTA kernels only, 1 = pairs where they
 * apply, < 0 = the environment's choice (AA_GEMM_PAIR, default on). */
int aa_debug_set_gemm_pair(int on);
/* Diagnostics: cap on the K-split (thread-block cluster size 1, 2 or 4) of the persistent BPTT kernel. */
int aa_debug_set_bptt_ksplit(int ks);
/* Diagnostics: on == 0 makes the bf16 recurrences always take the grid-barrier kernels (lstm_seq.cu) instead of the
 * cluster kernels with the weights in tensor memory (lstm_cluster.cu); nacc in {1,2,4} = partial accumulators the
 * forward cluster kernel spreads its MMA chain over (anything else: the default, 4).  Default: on. */
int aa_debug_set_lstm_cluster(int on, int nacc);

/* ---- stage operators (the nn.Module sub-blocks) ------------------------------------- */

/* Generic fp32 contraction Y[M,N] = X[M,K] * W[N,K]^T (+ bias[N]) — nn.Linear as used at
 * adaptive_attention.py:16-19,66-67,100.  ld* are row strides in elements. */
int aa_linear_forward(int M, int N, int K, const float* X, int64_t ldx, const float* W, int64_t ldw,
                      const float* bias, float* Y, int64_t ldy, void* stream);

/* The contraction engines themselves (every nn.Linear / bmm of the path is one of these):
 *   D[M,N] = sum_k A(m,k) B(n,k) + beta*C[M,N] + bias[N]
 * engine 0: exact fp32 SIMT (A, B float);  1: tcgen05 kind::f16 (A, B bf16, fp32 accumulate);
 * 2: tcgen05 kind::tf32 (A, B float read as tf32, fp32 accumulate).
 * a_kmajor != 0: A stored [M,K] (K contiguous, row stride lda); == 0: stored [K,M] (M contiguous).
 * b_kmajor != 0: B stored [N,K] (the nn.Linear weight layout); == 0: stored [K,N].
 * C, bias may be NULL.  tcgen05 engines need 16-byte aligned bases and row strides. */
int aa_gemm(int engine, int M, int N, int K, const void* A, int64_t lda, int a_kmajor, const void* B, int64_t ldb,
            int b_kmajor, const float* C, int64_t ldc, float beta, const float* bias, float* D, int64_t ldd,
            void* stream);

/* fp32-accurate tensor-core contraction ("3xTF32"): operands pre-split by aa_split_tf32 into rows of 2*Kp floats
 * [tf32 hi (cols zero-padded to Kp) | lo], Kp a multiple of 32.  D[M,N] = A B^T (+ bias[N]); lo*lo terms are dropped
 * (relative error ~1e-6, fp32 accumulation).  This is the engine of the two per-step contractions of decoding. */
int aa_split_tf32(const float* src, int64_t ld_src, int64_t rows, int cols, float* dst, int Kp, void* stream);
int aa_gemm_split3(int M, int N, int Kp, const float* A_split, const float* B_split, const float* bias, float* D,
                   int64_t ldd, void* stream);

/* P = V * W_v^T, once per image (adaptive_attention.py:34, `affine_v(V)`).  V [B,k,H] -> P [B,k,a]. */
int aa_precompute_P(const aa_dims* d, const float* V, const float* att_wv, float* P, void* stream);

/* Sentinel.forward (adaptive_attention.py:75-85): s = sigmoid(x W_x^T + h_prev W_h^T) * tanh(cell).
 * x [n,2E]; h_prev, cell [n,H] (h_prev may be NULL = zeros, the sampler case);
 * gate_out [n,H] (optional, the sigmoid) ; s_out [n,H].  n = B*T rows. */
int aa_sentinel_forward(const aa_dims* d, const float* sen_wx, const float* sen_wh, const float* x,
                        const float* h_prev, const float* cell, float* gate_out, float* s_out, void* stream);

/* Atten.forward (adaptive_attention.py:26-58): V [B,k,H], h_t, s_t [B,T,H] ->
 * c_hat [B,T,H], alpha [B,T,k], beta [B,T].  workspace: aa_atten_workspace_bytes(d). */
size_t aa_atten_workspace_bytes(const aa_dims* d);
int aa_atten_forward(const aa_dims* d, const float* att_wv, const float* att_wg, const float* att_ws,
                     const float* att_wh, const float* V, const float* h_t, const float* s_t, float* c_hat,
                     float* alpha, float* beta, void* workspace, size_t workspace_bytes, void* stream);

/* AdaptiveBlock.forward (adaptive_attention.py:110-134): x [B,T,2E], hiddens, cells [B,T,H], V [B,k,H] ->
 * scores [B,T,Vc], alpha [B,T,k], beta [B,T].  Applies the zero-h0 shift of :116-122 itself.
 * workspace: aa_adaptive_workspace_bytes(d). */
size_t aa_adaptive_workspace_bytes(const aa_dims* d);
int aa_adaptive_forward(const aa_dims* d, const aa_weights* w, const float* x, const float* hiddens,
                        const float* cells, const float* V, float* scores, float* alpha, float* beta,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- teacher-forced decoder: Decoder.forward / autograd backward -------------------- */

/* Bytes of the `saved` blob written by aa_decoder_forward and read by aa_decoder_backward, and
 * of the scratch blob aa_decoder_backward needs. */
size_t aa_decoder_saved_bytes(const aa_dims* d);
size_t aa_decoder_bwd_scratch_bytes(const aa_dims* d);

/* Decoder.forward (baseline_attention.py:148-194 with the adaptive block of adaptive_attention.py:155).
 *   V [B,k,H], v_g [B,E], captions [B,T] int64, h0/c0 [B,H] (NULL = zeros)
 *   -> scores [B,T,Vc], alpha [B,T,k], beta [B,T], hT/cT [B,H].
 * T == 1 is the sampler step (sentinel sees h~ = 0).  The `saved` blob
 * (aa_decoder_saved_bytes) is always required: it doubles as the forward workspace. */
int aa_decoder_forward(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g,
                       const int64_t* captions, const float* h0, const float* c0, float* scores, float* alpha,
                       float* beta, float* hT, float* cT, void* saved, size_t saved_bytes, void* stream);

/* Gradient of  <d_scores,scores> + <d_alpha,alpha> + <d_beta,beta> + <d_hT,hT> + <d_cT,cT>
 * (what autograd computes for the reference; formulas in SURVEY.md section 7).  d_alpha, d_beta,
 * d_hT, d_cT may be NULL (= zero).  Outputs: all 13 parameter gradients (overwritten) and
 * dV [B,k,H], dv_g [B,E], dh0, dc0 [B,H] (each may be NULL to skip).  d_scores is read-only. */
int aa_decoder_backward(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g,
                        const int64_t* captions, const float* h0, const float* c0, const float* alpha,
                        const float* beta, const void* saved, size_t saved_bytes, const float* d_scores,
                        const float* d_alpha, const float* d_beta, const float* d_hT, const float* d_cT,
                        const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0, float* dc0,
                        void* scratch, size_t scratch_bytes, void* stream);

/* Data-parallel hook (the reference's multi-GPU scheme is nn.DataParallel's reduce_add_coalesced after
 * backward, baseline_attention.py:184-187; here: one process per GPU, all-reduce overlapped with the
 * rest of the backward).  The parameter gradients become final in four buckets, in this order:
 *   AA_BUCKET_MLP      mlp_w, mlp_b                                      (first: half of all gradient bytes)
 *   AA_BUCKET_ATTEN    att_wv, att_wg, att_ws, att_wh, sen_wx, sen_wh
 *   AA_BUCKET_LSTM     w_ih, w_hh, b_ih, b_hh
 *   AA_BUCKET_EMBED    embed                                             (last)
 * aa_decoder_backward_hooked is aa_decoder_backward plus: as soon as the last kernel writing a bucket has been
 * enqueued, it records ready_events[bucket] (cudaEvent_t passed as void*, may be NULL to skip recording) on the
 * stream that kernel ran on and calls on_ready(bucket, user) on the calling host thread, so that the caller
 * can make a communication stream wait on the event and launch the bucket's all-reduce while the remaining
 * backward kernels are still running.  on_ready may be NULL.
 * ready_events has AA_NUM_READY_EVENTS entries: the four buckets above and, at index AA_EVENT_BPTT_DONE, an event recorded (and
 * reported through on_ready with that index) when the recurrence's backward -- the latency-critical kernel of the step, between
 * the ATTEN and LSTM buckets -- has been enqueued: measured on 8 GPUs, an exchange running next to that kernel doubles its time
 * (profiles/r02_timeline_n8_v2.txt), so the data-parallel driver starts the first exchange on this event. */
#define AA_BUCKET_MLP 0
#define AA_BUCKET_ATTEN 1
#define AA_BUCKET_LSTM 2
#define AA_BUCKET_EMBED 3
#define AA_NUM_BUCKETS 4
#define AA_EVENT_BPTT_DONE 4
#define AA_NUM_READY_EVENTS 5
typedef void (*aa_grad_ready_fn)(int bucket, void* user);
int aa_decoder_backward_hooked(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g,
                               const int64_t* captions, const float* h0, const float* c0, const float* alpha,
                               const float* beta, const void* saved, size_t saved_bytes, const float* d_scores,
                               const float* d_alpha, const float* d_beta, const float* d_hT, const float* d_cT,
                               const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0, float* dc0,
                               void* scratch, size_t scratch_bytes, void* stream, void* const* ready_events,
                               aa_grad_ready_fn on_ready, void* user);

/* Sum all-reduce of one gradient bucket over the GPUs of one NVSwitch box, written against peer-mapped ("symmetric") memory
 * instead of NCCL (the exchange step of data-parallel training).  Every rank holds one allocation of identical
 * size, mapped into every peer: peer_bufs[p] (HOST array of `world` device pointers) is rank p's allocation as seen from this
 * device, peer_bufs[rank] the local one; multicast_buf is the NVLS multicast mapping of the same allocation or NULL.  The bucket
 * is elements [offset_elems, offset_elems + n_elems) of the allocation viewed as float (both multiples of 4), reduced IN PLACE on
 * every rank.  flag_offset_bytes: start of aa_allreduce_flag_bytes() bytes inside the allocation, zero-filled once at allocation,
 * that the kernels use for their cross-GPU flags.  channel in [0, AA_AR_CHANNELS): collectives that may be in flight at the
 * same time must use different channels.  Every rank must make the same sequence of calls; the call is asynchronous on
 * `stream`, capturable into a CUDA graph, and uses at most max_blocks (<= AA_AR_MAX_BLOCKS; 0 = default) CTAs.
 * With multicast_buf: one multimem.ld_reduce + one multimem.st per 16 bytes of this rank's 1/world slice (the switch adds);
 * without: two-shot over peer loads, summed in rank order (deterministic). */
#define AA_AR_MAX_WORLD 8
#define AA_AR_MAX_BLOCKS 64
#define AA_AR_CHANNELS 4
size_t aa_allreduce_flag_bytes(void);
int aa_allreduce_sum_f32(void* const* peer_bufs, void* multicast_buf, long long flag_offset_bytes, int rank, int world,
                         long long offset_elems, long long n_elems, int channel, int max_blocks, void* stream);
/* Same reduction with the bucket travelling as bf16 (half the NVLink bytes; the mixed-precision training path): three launches
 * on `stream` -- the fp32 bucket is rounded into a bf16 staging region of the same allocation (stage_offset_bytes from its
 * start, indexed like the fp32 payload), summed over the ranks with fp32 accumulation (multimem.ld_reduce ... acc::f32 in the
 * switch, or registers without multicast), and the bf16 sum is widened back into the fp32 bucket.  Every rank ends with the
 * same values; they differ from the fp32 exchange by two bf16 roundings (<= 2^-8 relative each).  Multiples of 8 elements. */
int aa_allreduce_sum_bf16(void* const* peer_bufs, void* multicast_buf, long long flag_offset_bytes, long long stage_offset_bytes,
                          int rank, int world, long long offset_elems, long long n_elems, int channel, int max_blocks,
                          void* stream);

/* ---- optimizer step next to the path --------------------------------------------------------------
 * clip_grad_norm_(model.decoder.LSTM.parameters(), clip_max_norm) (train.py:213-214) followed by torch.optim.Adam's update
 * (model_factory.py:69-77; defaults of the reference: lr 1e-3, betas 0.8 / 0.999, eps 1e-8, weight_decay 0) over up to
 * AA_OPT_MAX_TENSORS parameter tensors in two launches.  clip[i] != 0 marks the tensors of the clipped group (the LSTM's
 * four); their gradients enter Adam scaled by min(1, max_norm / (norm + 1e-6)) and are rewritten in place only when
 * write_clipped_grads != 0.  step counts from 1.  scratch: one device float; norm_out (optional): the group's norm. */
#define AA_OPT_MAX_TENSORS 24
typedef struct aa_opt_tensors {
  float* param[AA_OPT_MAX_TENSORS];
  float* grad[AA_OPT_MAX_TENSORS];
  float* m[AA_OPT_MAX_TENSORS];
  float* v[AA_OPT_MAX_TENSORS];
  long long n[AA_OPT_MAX_TENSORS];
  int clip[AA_OPT_MAX_TENSORS];
} aa_opt_tensors;
int aa_clip_adam_step(const aa_opt_tensors* t, int n_tensors, int step, float lr, float beta1, float beta2, float eps,
                      float weight_decay, float clip_max_norm, int write_clipped_grads, float* scratch, float* norm_out,
                      void* stream);

/* Packed variants of aa_decoder_forward / aa_decoder_backward for Encoder2Decoder.forward (baseline_attention.py:206-230):
 * the reference projects all B*T positions onto the vocabulary and then keeps, through pack_padded_sequence (:228), only the
 * n_rows positions inside each caption's length.  Here the projection (and, in the backward, its three contractions, the bias
 * gradient and the bf16 cast) runs over those rows only, directly in packed time-major order.  row_index [n_rows] (device,
 * int64) = b*T+t of every packed row (aa host logic: functional.packed_row_index); scores_packed / d_scores_packed are
 * [n_rows, Vc].  Everything else as in the unpacked entry points; ready_events / on_ready / user as in
 * aa_decoder_backward_hooked (all may be NULL). */
int aa_decoder_forward_packed(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                              const float* h0, const float* c0, const int64_t* row_index, int64_t n_rows, float* scores_packed,
                              float* alpha, float* beta, float* hT, float* cT, void* saved, size_t saved_bytes, void* stream);
int aa_decoder_backward_packed(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                               const float* h0, const float* c0, const float* alpha, const float* beta, const void* saved,
                               size_t saved_bytes, const int64_t* row_index, int64_t n_rows, const float* d_scores_packed,
                               const float* d_alpha, const float* d_beta, const float* d_hT, const float* d_cT,
                               const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0, float* dc0, void* scratch,
                               size_t scratch_bytes, void* stream, void* const* ready_events, aa_grad_ready_fn on_ready, void* user,
                               const void* d_scores_packed_bf16 /* optional: bf16 mirror of d_scores_packed, see aa_cross_entropy_mirror */);

/* Forward + loss in one operator: Encoder2Decoder.forward (baseline_attention.py:206-230) followed by the
 * caller's mean cross-entropy over the packed positions (nn.CrossEntropyLoss at train.py:63,208) WITHOUT ever writing the
 * [n_rows, Vc] logits: the vocabulary projection's tcgen05 epilogue keeps, per row and 32-column chunk, the maximum, the sum of
 * exponentials and the exponentials themselves as bf16; two small passes turn them into the loss and into the bf16 gradient of the
 * logits (kept inside `saved`, together with the bias gradient).  AA_PREC_BF16 only (AA_ERR_UNSUPPORTED otherwise: the exact
 * path is aa_decoder_forward_packed + aa_cross_entropy).  targets [n_rows] int64 = the packed next words (train.py:102);
 * denom = denominator of the mean (0 = n_rows; data-parallel training passes the global count); loss: device scalar.
 * The matching backward is aa_decoder_backward_packed with d_scores_packed = NULL and d_scores_packed_bf16 = NULL: it takes the
 * gradient of the logits from `saved`.  If the loss is not the root of the backward pass, aa_decoder_loss_grad_scale multiplies
 * that stored gradient by the upstream scalar *g (device pointer; nothing is done when *g == 1) first. */
int aa_decoder_forward_loss(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                            const float* h0, const float* c0, const int64_t* row_index, int64_t n_rows, const int64_t* targets,
                            int64_t denom, float* loss, float* alpha, float* beta, float* hT, float* cT, void* saved,
                            size_t saved_bytes, void* stream);
int aa_decoder_loss_grad_scale(const aa_dims* d, void* saved, size_t saved_bytes, int64_t n_rows, const float* g, void* stream);

/* pack_padded_sequence(scores, lengths, batch_first=True).data (baseline_attention.py:228):
 * gathers rows (b,t) with t < lengths[b] in time-major order.  row_index [n_rows] int64 holds
 * b*T+t per packed row (host code builds it from `lengths`).  packed [n_rows,Vc]. */
int aa_pack_rows(const float* scores, int64_t n_cols, const int64_t* row_index, int64_t n_rows, float* packed,
                 void* stream);
/* adjoint: d_scores[B*T, Vc] = 0 ; d_scores[row_index[r]] = d_packed[r] */
int aa_unpack_rows(const float* d_packed, int64_t n_cols, const int64_t* row_index, int64_t n_rows, int64_t total_rows,
                   float* d_scores, void* stream);

/* Mean cross-entropy + gradient over rows (nn.CrossEntropyLoss at train.py:63,208; the step
 * after the hot path).  loss: device scalar (overwritten); dlogits may be NULL. */
int aa_cross_entropy(const float* logits, int64_t n_rows, int64_t Vc, const int64_t* targets, float* loss,
                     float* dlogits, void* stream);
/* Same with an explicit denominator: loss = sum_r nll_r / denom, dlogits = (softmax - onehot) / denom.  Data-parallel
 * training passes the GLOBAL packed-row count so that the all-reduced (summed) gradients equal the reference's
 * single-process mean (train.py:63,208). */
int aa_cross_entropy_denom(const float* logits, int64_t n_rows, int64_t Vc, const int64_t* targets, int64_t denom,
                           float* loss, float* dlogits, void* stream);

/* Same, and (optional) the bf16 mirror of dlogits written by the same pass -- the operand the vocabulary projection's backward
 * contracts with; handed to aa_decoder_backward_packed as d_scores_packed_bf16 it takes the fp32 -> bf16 cast and the bias
 * column sums off that call's critical path.  *mirror_written (host int, optional) says whether the mirror was produced (it
 * needs Vc % 4 == 0 and 16-byte aligned rows; otherwise the backward makes its own copy as before). */
int aa_cross_entropy_mirror(const float* logits, int64_t n_rows, int64_t Vc, const int64_t* targets, int64_t denom,
                            float* loss, float* dlogits, void* dlogits_bf16, int* mirror_written, void* stream);
/* x[0:n] *= *g unless *g == 1 (x_bf16: optional bf16 mirror kept consistent) -- the backward of the loss above when it is not the root of the backward pass
 * (torch.autograd hands the upstream gradient as a device scalar; `loss.backward()` at train.py:210 passes 1). */
int aa_scale_unless_one(float* x, const float* g, int64_t n, void* x_bf16, void* stream);
/* Up to 8 device-to-device copies in one launch (host-side plumbing with no reference counterpart: a step's input tensors
 * into the static buffers of a captured CUDA graph; six separate copies cost 27 us of a 500 us step). */
int aa_copy_multi(int n_segments, const void* const* src, void* const* dst, const int64_t* bytes, void* stream);

/* ---- encoder heads: AttentiveCNN.forward after the (out-of-scope) ResNet trunk ------- */

/* Shape of the heads: B images, C feature channels (2048), hw = h*w positions of the last-conv map (49 or 196),
 * H LSTM hidden size, E word-embedding size; precision AA_PREC_FP32 (exact) or AA_PREC_BF16 (tcgen05). */
typedef struct aa_enc_dims {
  int32_t B, C, hw, H, E;
  int32_t precision;
} aa_enc_dims;

/* encoder.{affine_a [H,C], affine_b [E,C], affine_h0 [H,C], affine_c0 [H,C]}.{weight, bias}   baseline_attention.py:21-34 */
typedef struct aa_enc_weights {
  const float *wa, *ba, *wb, *bb, *wh0, *bh0, *wc0, *bc0;
} aa_enc_weights;
typedef struct aa_enc_weight_grads {   /* same shapes; every buffer is OVERWRITTEN */
  float *wa, *ba, *wb, *bb, *wh0, *bh0, *wc0, *bc0;
} aa_enc_weight_grads;

size_t aa_encoder_saved_bytes(const aa_enc_dims* d);
size_t aa_encoder_bwd_scratch_bytes(const aa_enc_dims* d, int want_dA);

/* AttentiveCNN.forward from the feature map on (baseline_attention.py:46-62):
 *   A [B,C,hw] (the NCHW last-conv map) -> V = relu(affine_a(A^T)) [B,hw,H], v_g = relu(affine_b(a_g)) [B,E],
 *   h0 = tanh(affine_h0(a_g)) [B,H], c0 = tanh(affine_c0(a_g)) [B,H], a_g = mean of A over the map (AvgPool2d(7) on 7x7).
 * `saved` (aa_encoder_saved_bytes) keeps the transposed map, the pooled rows and the bf16 weight copies for the backward. */
int aa_encoder_forward(const aa_enc_dims* d, const aa_enc_weights* w, const float* A, float* V, float* v_g, float* h0, float* c0,
                       void* saved, size_t saved_bytes, void* stream);
/* Gradient of the above (autograd in the reference, train.py:210): V, v_g, h0, c0 are the forward's outputs, dV ... dc0 the
 * upstream gradients; writes the 8 parameter gradients and, when dA != NULL, the gradient towards the trunk [B,C,hw]. */
int aa_encoder_backward(const aa_enc_dims* d, const aa_enc_weights* w, const void* saved, size_t saved_bytes, const float* V,
                        const float* v_g, const float* h0, const float* c0, const float* dV, const float* dv_g, const float* dh0,
                        const float* dc0, const aa_enc_weight_grads* gw, float* dA, void* scratch, size_t scratch_bytes, void* stream);

/* ---- decoding: Encoder2Decoder.sampler -------------------------------------------- */

/* Workspace for decoding d->B images for d->T (= max_len) steps: beam = 0 -> aa_greedy_decode,
 * beam >= 1 -> aa_beam_decode with that width. */
size_t aa_decode_workspace_bytes(const aa_dims* d, int beam);

/* Greedy sampler loop (adaptive_attention.py:186-216) for B images, max_len steps, fully on the
 * device: start id 1, arg-max of raw logits (lowest index wins ties), never stops early.
 *   V [B,k,H], v_g [B,E], h0/c0 [B,H] -> ids [B,max_len] int64, attention [B,max_len,k],
 *   Beta [B,max_len].  logits_out (optional, [max_len,B,Vc]) receives every step's scores. */
int aa_greedy_decode(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const float* h0,
                     const float* c0, int max_len, int64_t* ids, float* attention, float* Beta, float* logits_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Persistent-kernel variant of the greedy sampler (same loop, adaptive_attention.py:186-216; same outputs as aa_greedy_decode)
 * for batches of at most ONE IMAGE PER SM: the whole max_len-step loop is one cooperative launch; each image's V, P = V W_v^T
 * and LSTM cell state stay in the shared memory of the SM that owns it for all steps (24 130 instead of
 * 128 588 algorithmic bytes per image and step), the gate and vocabulary contractions run on tcgen05 inside the kernel
 * (fp32-accurate 3xTF32 gates; bf16 first pass + exact fp32 recomputation of every column that can hold the row maximum), four
 * grid barriers per step replace the per-step launches.  aa_decode_persistent_supported: 1 if d->B (<= SM count) and the shape
 * fit the residency limits on the current device, else 0 -- callers fall back to aa_greedy_decode (aa_decode_persistent itself
 * returns AA_ERR_UNSUPPORTED then).  flags: AA_DECODE_REUSE_PACKED_WEIGHTS = the weight-derived operands at the head of
 * `workspace` (packed gate weights, bf16 projection, row norms) were written by an earlier call with the same weights, dims and
 * workspace and are not rebuilt.  candidates_out (optional, int32 [B,max_len]): columns recomputed exactly per row and step. */
#define AA_DECODE_REUSE_PACKED_WEIGHTS 1
size_t aa_decode_persistent_workspace_bytes(const aa_dims* d);   /* d->T = max_len */
int aa_decode_persistent_supported(const aa_dims* d);
int aa_decode_persistent(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const float* h0,
                         const float* c0, int max_len, int64_t* ids, float* attention, float* Beta, int flags,
                         int* candidates_out, void* workspace, size_t workspace_bytes, void* stream);

/* Beam search (NOT in the reference; definition in SURVEY section 8c / oracle.beam_decode):
 * returns the best hypothesis per image: ids [B,max_len], attention [B,max_len,k], Beta [B,max_len],
 * score [B] (cumulative log-prob). */
int aa_beam_decode(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const float* h0,
                   const float* c0, int beam, int max_len, int64_t* ids, float* attention, float* Beta, float* score,
                   void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADAPTIVE_B200_H_ */
