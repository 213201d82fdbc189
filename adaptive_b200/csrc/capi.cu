// extern "C" boundary of libadaptive_sm100.so (see include/adaptive_b200.h) and the host-side
// orchestration of the teacher-forced forward/backward.  Decoding lives in decode_api.cu.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/adaptive_b200.h"
#include "kernels.cuh"
#include "host_common.cuh"

namespace aa {
extern int g_atten_sequential;

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- side stream (fork/join inside one call) -------------------------------------------------
// The backward has work that is off the critical path (every weight-gradient contraction, dV, the sentinel's
// dx): it runs on a library-owned non-blocking stream, forked from / joined back into the caller's stream with
// events, so the call still looks like ONE stream-ordered operation to the caller (and captures into a CUDA
// graph as a fork/join).  Three side lanes + event pool per (host thread, device), created lazily -- the first
// call on a thread must therefore not be made under stream capture (same rule as the one-time
// cudaFuncSetAttribute calls of the kernels).  AA_NO_SIDE_STREAM=1 serialises everything on the caller's stream.
// The critical path of the training forward / backward itself runs on a library-owned stream of the HIGHEST priority
// (forked from and joined into the caller's stream like the side lanes): when a side-lane contraction and a critical one
// become ready together, the block scheduler hands free SMs to the critical one first (the caller's stream usually has the
// lowest priority, like the side lanes; profiles/r01_v43_timeline.txt shows du = dS W_p waiting 14 us behind dW_p).
namespace {
constexpr int SIDE_EVENTS = 32;
struct SideCtx {
  cudaStream_t side = nullptr, side2 = nullptr, side3 = nullptr;   // lanes A, B, C (lowest priority)
  cudaStream_t crit = nullptr;                                     // the critical path of a call (highest priority)
  cudaStream_t zero = nullptr;                                     // zero-fills the critical path will wait for (highest priority)
  cudaEvent_t ev[SIDE_EVENTS] = {};
  bool ready = false;
};
thread_local SideCtx g_side[16];
int g_side_disabled = -1;
}  // namespace

static int get_side(SideCtx** out) {
  *out = nullptr;
  if (g_side_disabled < 0) {
    const char* e = getenv("AA_NO_SIDE_STREAM");
    g_side_disabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (g_side_disabled) return AA_OK;
  int dev = 0;
  AA_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return AA_OK;
  SideCtx& c = g_side[dev];
  if (!c.ready) {
    int prio_least = 0, prio_greatest = 0;
    AA_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    AA_CHECK_CUDA(cudaStreamCreateWithPriority(&c.side, cudaStreamNonBlocking, prio_least));
    AA_CHECK_CUDA(cudaStreamCreateWithPriority(&c.side2, cudaStreamNonBlocking, prio_least));
    AA_CHECK_CUDA(cudaStreamCreateWithPriority(&c.side3, cudaStreamNonBlocking, prio_least));
    AA_CHECK_CUDA(cudaStreamCreateWithPriority(&c.crit, cudaStreamNonBlocking, prio_greatest));
    AA_CHECK_CUDA(cudaStreamCreateWithPriority(&c.zero, cudaStreamNonBlocking, prio_greatest));
    for (int i = 0; i < SIDE_EVENTS; ++i) AA_CHECK_CUDA(cudaEventCreateWithFlags(&c.ev[i], cudaEventDisableTiming));
    c.ready = true;
  }
  *out = &c;
  return AA_OK;
}

// `to` waits for everything enqueued on `from` so far (event slot i of the pool)
static int stream_dep(SideCtx* sc, int i, cudaStream_t from, cudaStream_t to) {
  if (!sc || from == to) return AA_OK;
  AA_CHECK_CUDA(cudaEventRecord(sc->ev[i], from));
  AA_CHECK_CUDA(cudaStreamWaitEvent(to, sc->ev[i], 0));
  return AA_OK;
}

// ---- instrumentation -------------------------------------------------------------------
namespace {
std::atomic<long long> g_launches{0};
std::atomic<int> g_prof_on{0};
struct ProfRec { const char* tag; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
struct ProfTotal { std::string tag; double ms; int n; };
std::vector<ProfTotal> g_prof_totals;
}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

ProfScope::ProfScope(const char* tag, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{tag, nullptr, nullptr};
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  slot = (int)g_prof_recs.size();
  g_prof_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof_recs[slot].e1, stream);
}

static void prof_resolve() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      bool found = false;
      for (auto& t : g_prof_totals)
        if (t.tag == r.tag) { t.ms += ms; t.n += 1; found = true; break; }
      if (!found) g_prof_totals.push_back({r.tag, ms, 1});
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof_recs.clear();
}

namespace {

inline int a_pad_of(const aa_dims& d) { return (d.a + 7) / 8 * 8; }   // bf16 rows need 16-byte strides for TMA

// bf16 copies of the GEMM weights (mixed-precision mode)
struct W16 {
  bf16 *w_ih, *w_hh, *sen_wx, *sen_wh, *att_wv, *att_wg, *att_ws, *mlp_w;
};

struct Saved {
  float *x, *xg, *acts, *cells, *hiddens, *hs_prev, *g, *s, *P, *q, *r, *ctx, *u, *zeros;
  float* up;      // rows of u in packed order (packed entry points only)
  // bf16 mirrors (only carved when d.precision == AA_PREC_BF16)
  bf16 *x16, *hid16, *hsprev16, *s16, *V16, *u16, *h016, *whh_pack16, *up16;
  unsigned* counters;
  W16 w16;
  // fused cross-entropy (aa_decoder_forward_loss, bf16 mode): exp pieces, then the bf16 gradient of the packed logits, in place; chunk
  // partials (max, sum) and scales; the targets' logits; the bias gradient
  bf16* ce_e16; float *ce_part, *ce_scale, *ce_xt, *ce_dbp; int ce_chunks;
  size_t bytes;
};

Saved carve_saved(const aa_dims& d, void* base) {
  const size_t N = (size_t)d.B * d.T, H = d.H, E = d.E;
  Carver c(base);
  Saved s{};
  s.x = c.take(N * 2 * E);
  s.xg = c.take(N * 4 * H);
  s.acts = c.take(N * 4 * H);
  s.cells = c.take(N * H);
  s.hiddens = c.take(N * H);
  s.hs_prev = c.take(N * H);
  s.g = c.take(N * H);
  s.s = c.take(N * H);
  s.P = c.take((size_t)d.B * d.k * d.a);
  s.q = c.take(N * d.a);
  s.r = c.take(N * d.a);
  s.ctx = c.take(N * H);
  s.u = c.take(N * H);
  s.zeros = c.take((size_t)d.B * H);
  s.up = c.take(N * H);
  if (d.precision == AA_PREC_BF16) {
    Carver16 h{c};
    s.up16 = h.take(N * H);
    s.x16 = h.take(N * 2 * E);
    s.hid16 = h.take(N * H);
    s.hsprev16 = h.take(N * H + (size_t)d.B * H);   // (+ B rows: h0, the K-tail of the dW_hh contraction)
    s.s16 = h.take(N * H);
    s.V16 = h.take((size_t)d.B * d.k * H);
    s.u16 = h.take(N * H);
    s.h016 = h.take((size_t)d.B * H);
    s.w16.w_ih = h.take(4 * H * 2 * E);
    s.w16.w_hh = h.take(4 * H * H);
    s.w16.sen_wx = h.take(H * 2 * E);
    s.w16.sen_wh = h.take(H * H);
    s.w16.att_wv = h.take((size_t)d.a * H);
    s.w16.att_wg = h.take((size_t)d.a * H);
    s.w16.att_ws = h.take((size_t)d.a * H);
    s.w16.mlp_w = h.take((size_t)d.Vc * H);
    s.whh_pack16 = h.take(4 * H * H);
    s.counters = reinterpret_cast<unsigned*>(c.take(64));
    s.ce_chunks = (d.Vc + 31) / 32;
    s.ce_e16 = h.take(N * d.Vc);
    s.ce_part = c.take(N * s.ce_chunks * 2);
    s.ce_scale = c.take(N * s.ce_chunks);
    s.ce_xt = c.take(N);
    s.ce_dbp = c.take((size_t)d.Vc + 1);      // (+ 1: the loss accumulator)
  }
  s.bytes = c.off;
  return s;
}

struct BwdScratch {
  float* dup;     // du rows in packed order (packed entry points only)
  float *du, *ds, *dq, *dr, *dP, *da, *dcell, *dhs, *dgates, *dx, *dh_rec, *dc_rec, *dV;
  bf16 *dS16, *dq16, *dr16, *dP16, *da16, *dgates16, *whhT16;
  unsigned* counters;
  size_t bytes;
};

BwdScratch carve_bwd(const aa_dims& d, void* base) {
  const size_t N = (size_t)d.B * d.T, H = d.H, E = d.E;
  Carver c(base);
  BwdScratch s{};
  s.du = c.take(N * H);
  s.ds = c.take(N * H);
  s.dq = c.take(N * d.a);
  s.dr = c.take(N * d.a);
  s.dP = c.take((size_t)d.B * d.k * d.a);
  s.da = c.take(N * H);
  s.dcell = c.take(N * H);
  s.dhs = c.take(N * H);
  s.dgates = c.take(N * 4 * H);
  s.dx = c.take(N * 2 * E);
  s.dh_rec = c.take((size_t)d.B * H);
  s.dc_rec = c.take((size_t)d.B * H);
  s.dV = c.take((size_t)d.B * d.k * H);
  s.dup = c.take(N * H);
  if (d.precision == AA_PREC_BF16) {
    Carver16 h{c};
    const size_t ap = a_pad_of(d);
    s.dS16 = h.take(N * d.Vc);
    s.dq16 = h.take(N * ap);
    s.dr16 = h.take(N * ap);
    s.dP16 = h.take((size_t)d.B * d.k * ap);
    s.da16 = h.take(N * H);
    s.dgates16 = h.take(N * 4 * H + (size_t)d.B * 4 * H);   // (+ B rows: dgates of step 0, the K-tail of the dW_hh contraction)
    s.whhT16 = h.take(4 * H * H);
    s.counters = reinterpret_cast<unsigned*>(c.take(64));
  }
  s.bytes = c.off;
  return s;
}

int check_dims(const aa_dims* d, bool need_T) {
  AA_REQUIRE(d != nullptr, "dims is NULL");
  AA_REQUIRE(d->B >= 0 && d->k >= 1 && d->a >= 1 && d->H >= 4 && d->E >= 4 && d->Vc >= 1,
             "bad dims B=%d k=%d a=%d H=%d E=%d Vc=%d", d->B, d->k, d->a, d->H, d->E, d->Vc);
  AA_REQUIRE(d->H % 4 == 0 && d->E % 4 == 0, "H and E must be multiples of 4 (H=%d, E=%d)", d->H, d->E);
  AA_REQUIRE(d->a <= 128, "attention dim a=%d exceeds 128", d->a);
  if (need_T) AA_REQUIRE(d->T >= 1, "T must be >= 1 (got %d)", d->T);
  AA_REQUIRE(d->precision == AA_PREC_FP32 || d->precision == AA_PREC_BF16, "unknown precision %d", d->precision);
  if (d->precision == AA_PREC_BF16)
    AA_REQUIRE(d->H % 8 == 0 && d->E % 8 == 0 && d->Vc % 8 == 0,
               "bf16 mode needs H, E and Vc to be multiples of 8 (16-byte bf16 rows for TMA); got H=%d E=%d Vc=%d", d->H, d->E, d->Vc);
  return AA_OK;
}

// The sentinel-less baseline model (baseline_attention.py:66-194) is the same operator with the three sentinel
// weights absent: sen_wx, sen_wh and att_ws are NULL together.  beta is then 0 everywhere, c_hat = ctx, scores = mlp(ctx + h).
int check_sentinel_weights(const aa_weights* w, bool* base) {
  const int n = (w->sen_wx != nullptr) + (w->sen_wh != nullptr) + (w->att_ws != nullptr);
  AA_REQUIRE(n == 0 || n == 3, "sen_wx, sen_wh and att_ws must be given together (adaptive model) or all be NULL (baseline model)");
  *base = n == 0;
  return AA_OK;
}

// dst[r, :] = src[row_index[r], :] (fp32 and, optionally, the bf16 mirror in the same launch); cols % 4 == 0
// (zero / nzero: an fp32 array the same launch zero-fills on the way -- the fused loss's scalar and bias gradient: memset nodes in front
//  of the vocabulary contraction cost ~13 us of dispatch inside the replayed graph, profiles/r02_timeline_n1_fusedce_v1.txt)
__global__ void gather_rows2_kernel(const float* __restrict__ src, const bf16* __restrict__ src16, const long long* __restrict__ row_index,
                                    int cols, float* __restrict__ dst, bf16* __restrict__ dst16, float* __restrict__ zero = nullptr,
                                    long long nzero = 0) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nzero; i += (long long)gridDim.x * blockDim.x) zero[i] = 0.f;
  const long long r = blockIdx.x, sr = row_index[r];
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    *reinterpret_cast<float4*>(dst + r * cols + c) = *reinterpret_cast<const float4*>(src + sr * cols + c);
    if (src16) *reinterpret_cast<uint2*>(dst16 + r * cols + c) = *reinterpret_cast<const uint2*>(src16 + sr * cols + c);
  }
}

__global__ void pack_rows_kernel(const float* __restrict__ src, long long n_cols, const long long* __restrict__ row_index,
                                 float* __restrict__ dst, int gather) {
  const long long r = blockIdx.x;
  const long long sr = row_index[r];
  const float* s = gather ? src + sr * n_cols : src + r * n_cols;
  float* d = gather ? dst + r * n_cols : dst + sr * n_cols;
  if ((n_cols & 3) == 0) {
    for (long long c = threadIdx.x * 4; c < n_cols; c += blockDim.x * 4)
      *reinterpret_cast<float4*>(d + c) = *reinterpret_cast<const float4*>(s + c);
  } else {
    for (long long c = threadIdx.x; c < n_cols; c += blockDim.x) d[c] = s[c];
  }
}

}  // namespace

// attention sub-stage shared by aa_atten_forward / aa_adaptive_forward / aa_decoder_forward
static int atten_stage(const Ctx& c, const aa_dims& d, Mat Wv, Mat Wg, Mat Ws, const float* att_wh, Mat V, Mat h, Mat s, float* P,
                       float* q, float* r, float* c_hat, float* ctx, float* u, bf16* u16, float* alpha, float* beta,
                       bool have_P = false, bool have_q = false, bool base = false) {
  const int N = d.B * d.T;
  if (!have_P) AA_TRY(mm_nt(c, "gemm_P", d.B * d.k, d.a, d.H, V, Wv, P, d.a, nullptr, 0, nullptr, nullptr));   // :34
  if (!have_q) AA_TRY(mm_nt(c, "gemm_qr", N, d.a, d.H, h, Wg, q, d.a, nullptr, 0, nullptr, nullptr));           // :35
  if (!base) AA_TRY(mm_nt(c, "gemm_qr", N, d.a, d.H, s, Ws, r, d.a, q, d.a, nullptr, nullptr));    // :45   (baseline: r, s are zero-filled)
  AttenFwdArgs p{};
  p.B = d.B; p.T = d.T; p.k = d.k; p.a = d.a; p.H = d.H;
  p.P = P; p.q = q; p.r = r; p.s = s.f; p.h = h.f; p.V = V.f; p.wh = att_wh;
  p.alpha = alpha; p.beta = beta; p.ctx = ctx; p.u = u; p.c_hat = c_hat; p.u16 = u16;
  p.no_sentinel = base ? 1 : 0;
  AA_PROF("atten_fwd", c.st, launch_atten_fwd(p, c.st));
  return AA_OK;
}

}  // namespace aa

using namespace aa;

extern "C" {

int aa_debug_set_gemm_splitk(int on) { return aa::set_gemm_splitk(on); }
int aa_debug_set_gemm_pair(int on) { return aa::set_gemm_pair(on); }

int aa_debug_set_bptt_ksplit(int ks) { return aa::set_bptt_ksplit_max(ks); }

int aa_debug_set_atten_sequential(int on) {
  aa::g_atten_sequential = on ? 1 : 0;
  return AA_OK;
}

int aa_version(void) { return 100; /* 0.1.0 */ }

const char* aa_last_error(void) { return get_error(); }

long long aa_launch_count(void) { return g_launches.load(); }

int aa_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return AA_OK;
}

int aa_profile_reset(void) {
  prof_resolve();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_totals.clear();
  return AA_OK;
}

int aa_profile_count(void) {
  prof_resolve();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  return (int)g_prof_totals.size();
}

int aa_profile_get(int i, char* name, int name_len, double* total_ms, int* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  AA_REQUIRE(i >= 0 && i < (int)g_prof_totals.size(), "aa_profile_get: index %d out of range", i);
  if (name && name_len > 0) {
    strncpy(name, g_prof_totals[i].tag.c_str(), name_len - 1);
    name[name_len - 1] = 0;
  }
  if (total_ms) *total_ms = g_prof_totals[i].ms;
  if (launches) *launches = g_prof_totals[i].n;
  return AA_OK;
}

int aa_debug_set_trace_buffer(void* dev_ptr) {
  AA_TRY(set_clk_trace_buffer(dev_ptr));
  return set_seq_trace_buffer(dev_ptr);
}

int aa_debug_set_lstm_cluster(int on, int nacc) { return aa::set_lstm_cluster(on, nacc); }

int aa_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  AA_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  AA_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return AA_OK;
}

int aa_linear_forward(int M, int N, int K, const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                      float* Y, int64_t ldy, void* stream) {
  return gemm_nt(M, N, K, X, ldx, W, ldw, Y, ldy, nullptr, 0, bias, nullptr, (cudaStream_t)stream);
}

int aa_gemm(int engine, int M, int N, int K, const void* A, int64_t lda, int a_kmajor, const void* B, int64_t ldb, int b_kmajor,
            const float* C, int64_t ldc, float beta, const float* bias, float* D, int64_t ldd, void* stream) {
  if (engine == 0) {
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K;
    g.A = static_cast<const float*>(A); g.lda = lda; g.a_kcontig = a_kmajor;
    g.B = static_cast<const float*>(B); g.ldb = ldb; g.b_kcontig = b_kmajor;
    g.Cin = C; g.ldcin = ldc; g.D = D; g.ldd = ldd; g.bias1 = bias; g.alpha = 1.f; g.beta = C ? beta : 0.f; g.splitk = 1;
    return launch_sgemm(g, (cudaStream_t)stream);
  }
  AA_REQUIRE(engine == 1 || engine == 2, "aa_gemm: engine must be 0 (fp32 SIMT), 1 (tcgen05 bf16) or 2 (tcgen05 tf32)");
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_mn = a_kmajor ? 0 : 1;
  g.B = B; g.ldb = ldb; g.b_mn = b_kmajor ? 0 : 1;
  g.elem_size = engine == 1 ? 2 : 4;
  g.D32 = D; g.ldd32 = ldd; g.Cin = C; g.ldcin = ldc; g.beta = beta; g.bias1 = bias;
  return launch_gemm_tc(g, (cudaStream_t)stream);
}

int aa_split_tf32(const float* src, int64_t ld_src, int64_t rows, int cols, float* dst, int Kp, void* stream) {
  AA_REQUIRE(src && dst && rows >= 0 && cols >= 1, "aa_split_tf32: bad argument");
  return launch_split_tf32(src, ld_src, rows, cols, dst, Kp, (cudaStream_t)stream);
}

int aa_gemm_split3(int M, int N, int Kp, const float* A_split, const float* B_split, const float* bias, float* D, int64_t ldd,
                   void* stream) {
  TcGemmArgs g{};
  g.M = M; g.N = N; g.K = Kp; g.elem_size = 4; g.split3 = 1;
  g.A = A_split; g.lda = 2LL * Kp; g.B = B_split; g.ldb = 2LL * Kp;
  g.D32 = D; g.ldd32 = ldd; g.bias1 = bias; g.beta = 0.f;
  return launch_gemm_tc(g, (cudaStream_t)stream);
}

int aa_precompute_P(const aa_dims* d, const float* V, const float* att_wv, float* P, void* stream) {
  AA_TRY(check_dims(d, false));
  AA_REQUIRE(V && att_wv && P, "aa_precompute_P: null pointer");
  return gemm_nt(d->B * d->k, d->a, d->H, V, d->H, att_wv, d->H, P, d->a, nullptr, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int aa_sentinel_forward(const aa_dims* d, const float* sen_wx, const float* sen_wh, const float* x, const float* h_prev,
                        const float* cell, float* gate_out, float* s_out, void* stream) {
  AA_TRY(check_dims(d, true));
  AA_REQUIRE(sen_wx && x && cell && s_out && gate_out, "aa_sentinel_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = d->B * d->T;
  AA_TRY(gemm_nt(N, d->H, 2 * d->E, x, 2 * d->E, sen_wx, 2 * d->E, gate_out, d->H, nullptr, 0, nullptr, nullptr, st));
  if (h_prev) {
    AA_REQUIRE(sen_wh, "aa_sentinel_forward: sen_wh is NULL");
    AA_TRY(gemm_nt(N, d->H, d->H, h_prev, d->H, sen_wh, d->H, gate_out, d->H, gate_out, d->H, nullptr, nullptr, st));
  }
  return launch_sentinel_fwd(gate_out, cell, gate_out, s_out, nullptr, (long long)N * d->H, st);
}

size_t aa_atten_workspace_bytes(const aa_dims* d) {
  if (!d) return 0;
  const size_t N = (size_t)d->B * d->T;
  return align_up((size_t)d->B * d->k * d->a * 4, 256) + 2 * align_up(N * d->a * 4, 256);
}

int aa_atten_forward(const aa_dims* d, const float* att_wv, const float* att_wg, const float* att_ws, const float* att_wh,
                     const float* V, const float* h_t, const float* s_t, float* c_hat, float* alpha, float* beta,
                     void* workspace, size_t workspace_bytes, void* stream) {
  AA_TRY(check_dims(d, true));
  AA_REQUIRE(att_wv && att_wg && att_wh && V && h_t && c_hat && alpha && beta, "aa_atten_forward: null pointer");
  AA_REQUIRE((att_ws != nullptr) == (s_t != nullptr), "aa_atten_forward: att_ws and s_t go together (both NULL: baseline Atten.forward)");
  const bool base = att_ws == nullptr;
  if (workspace_bytes < aa_atten_workspace_bytes(d) || !workspace) {
    set_error("aa_atten_forward: workspace too small (%zu < %zu)", workspace_bytes, aa_atten_workspace_bytes(d));
    return AA_ERR_WORKSPACE;
  }
  Carver c(workspace);
  const size_t N = (size_t)d->B * d->T;
  float* P = c.take((size_t)d->B * d->k * d->a);
  float* q = c.take(N * d->a);
  float* r = c.take(N * d->a);
  const Ctx cx{AA_PREC_FP32, (cudaStream_t)stream};
  if (base) {   // baseline_attention.py:79-100: the kernel still reads (finite) r and s rows -- r is cleared, h_t stands in for s_t
    AA_CHECK_CUDA(cudaMemsetAsync(r, 0, sizeof(float) * N * d->a, cx.st));
    s_t = h_t;
  }
  return atten_stage(cx, *d, M32(att_wv, d->H), M32(att_wg, d->H), M32(att_ws, d->H), att_wh, M32(V, d->H), M32(h_t, d->H),
                     M32(s_t, d->H), P, q, r, c_hat, nullptr, nullptr, nullptr, alpha, beta, false, false, base);
}

size_t aa_adaptive_workspace_bytes(const aa_dims* d) {
  if (!d) return 0;
  const size_t N = (size_t)d->B * d->T;
  return aa_atten_workspace_bytes(d) + 4 * align_up(N * d->H * 4, 256);
}

int aa_adaptive_forward(const aa_dims* d, const aa_weights* w, const float* x, const float* hiddens, const float* cells,
                        const float* V, float* scores, float* alpha, float* beta, void* workspace, size_t workspace_bytes,
                        void* stream) {
  AA_TRY(check_dims(d, true));
  AA_REQUIRE(w && x && hiddens && cells && V && scores && alpha && beta, "aa_adaptive_forward: null pointer");
  bool base = false;
  AA_TRY(check_sentinel_weights(w, &base));
  if (workspace_bytes < aa_adaptive_workspace_bytes(d) || !workspace) {
    set_error("aa_adaptive_forward: workspace too small (%zu < %zu)", workspace_bytes, aa_adaptive_workspace_bytes(d));
    return AA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t N = (size_t)d->B * d->T;
  const int H = d->H, T = d->T;
  Carver c(workspace);
  float* P = c.take((size_t)d->B * d->k * d->a);
  float* q = c.take(N * d->a);
  float* r = c.take(N * d->a);
  float* hs_prev = c.take(N * H);
  float* g = c.take(N * H);
  float* s = c.take(N * H);
  float* u = c.take(N * H);
  if (base) {   // baseline block (baseline_attention.py:121-128): no sentinel; the attention kernel reads zero r and s rows
    AA_CHECK_CUDA(cudaMemsetAsync(s, 0, sizeof(float) * N * H, st));
    AA_CHECK_CUDA(cudaMemsetAsync(r, 0, sizeof(float) * N * d->a, st));
  } else {
    // h~: zero-h0 shift of adaptive_attention.py:116-122
    AA_TRY(gemm_nt((int)N, H, 2 * d->E, x, 2 * d->E, w->sen_wx, 2 * d->E, g, H, nullptr, 0, nullptr, nullptr, st));
    if (T > 1) {
      AA_CHECK_CUDA(cudaMemset2DAsync(hs_prev, (size_t)T * H * 4, 0, (size_t)H * 4, d->B, st));
      AA_TRY(launch_copy2d(hs_prev + H, (long long)T * H, hiddens, (long long)T * H, d->B, (T - 1) * H, st));
      AA_TRY(gemm_nt((int)N, H, H, hs_prev, H, w->sen_wh, H, g, H, g, H, nullptr, nullptr, st));
    }
    AA_TRY(launch_sentinel_fwd(g, cells, g, s, nullptr, (long long)N * H, st));
  }
  const Ctx cx{AA_PREC_FP32, st};
  AA_TRY(atten_stage(cx, *d, M32(w->att_wv, H), M32(w->att_wg, H), M32(w->att_ws, H), w->att_wh, M32(V, H), M32(hiddens, H),
                     M32(s, H), P, q, r, nullptr, nullptr, u, nullptr, alpha, beta, false, false, base));
  return gemm_nt((int)N, d->Vc, H, u, H, w->mlp_w, H, scores, d->Vc, nullptr, 0, w->mlp_b, nullptr, st);   // :132
}

size_t aa_decoder_saved_bytes(const aa_dims* d) { return d ? carve_saved(*d, nullptr).bytes : 0; }
size_t aa_decoder_bwd_scratch_bytes(const aa_dims* d) { return d ? carve_bwd(*d, nullptr).bytes : 0; }

struct CeFuse;
static int decoder_forward_body(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                                const float* h0, const float* c0, float* scores, float* alpha, float* beta, float* hT, float* cT,
                                void* saved, size_t saved_bytes, void* stream, SideCtx* side, const int64_t* row_index, int64_t n_rows,
                                const CeFuse* ce);

// fork the call's critical lane from the caller's stream, run the body on it, join every lane back
// fused loss (aa_decoder_forward_loss): packed targets, denominator of the mean, the loss scalar
struct CeFuse {
  const int64_t* targets; long long denom; float* loss;
};

static int decoder_forward_impl(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                                const float* h0, const float* c0, float* scores, float* alpha, float* beta, float* hT, float* cT,
                                void* saved, size_t saved_bytes, void* stream, const int64_t* row_index, int64_t n_rows,
                                const CeFuse* ce = nullptr) {
  SideCtx* side = nullptr;
  AA_TRY(get_side(&side));
  cudaStream_t caller = (cudaStream_t)stream;
  cudaStream_t st = side ? side->crit : caller;
  AA_TRY(stream_dep(side, SIDE_EVENTS - 4, caller, st));
  AA_TRY(decoder_forward_body(d, w, V, v_g, captions, h0, c0, scores, alpha, beta, hT, cT, saved, saved_bytes, (void*)st, side, row_index, n_rows, ce));
  if (side) AA_TRY(stream_dep(side, SIDE_EVENTS - 5, side->side2, st));   // (hT / cT copies)
  return stream_dep(side, SIDE_EVENTS - 6, st, caller);
}

static int decoder_forward_body(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                                const float* h0, const float* c0, float* scores, float* alpha, float* beta, float* hT, float* cT,
                                void* saved, size_t saved_bytes, void* stream, SideCtx* side, const int64_t* row_index, int64_t n_rows,
                                const CeFuse* ce) {
  AA_TRY(check_dims(d, true));
  AA_REQUIRE(w && V && v_g && captions && (scores || ce) && alpha && beta, "aa_decoder_forward: null pointer");
  bool base = false;
  AA_TRY(check_sentinel_weights(w, &base));
  if (d->B == 0) return AA_OK;
  if (!saved || saved_bytes < aa_decoder_saved_bytes(d)) {
    set_error("aa_decoder_forward: saved blob too small (%zu < %zu)", saved_bytes, aa_decoder_saved_bytes(d));
    return AA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, T = d->T, H = d->H, E = d->E;
  const int N = B * T;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 size mismatch");
  Saved sv = carve_saved(*d, saved);
  const long long* cap = reinterpret_cast<const long long*>(captions);

  const Ctx cx{d->precision, st};
  const bool tc = cx.tc();
  const W16& h = sv.w16;
  // Side lane, forked at once: what the recurrence does not need -- the bf16 copies of every weight but the LSTM's and of V,
  // P = V W_v^T and the sentinel gate's input half -- runs next to the main lane's casts, gate GEMM and recurrence.
  const Ctx cs{d->precision, side ? side->side : st};
  if (!tc) AA_TRY(stream_dep(side, SIDE_EVENTS - 2, st, cs.st));
  if (tc) {   // bf16 copies of the GEMM weights and of V, h0
    const float* srcs[8] = {w->w_ih, w->w_hh, w->sen_wx, w->sen_wh, w->att_wv, w->att_wg, w->att_ws, w->mlp_w};
    bf16* dsts[8] = {h.w_ih, h.w_hh, h.sen_wx, h.sen_wh, h.att_wv, h.att_wg, h.att_ws, h.mlp_w};
    const long long ns[8] = {4LL * H * 2 * E, 4LL * H * H, (long long)H * 2 * E, (long long)H * H, (long long)d->a * H,
                             (long long)d->a * H, (long long)d->a * H, (long long)d->Vc * H};
    CastSegs c_main{}, c_side{};
    for (int i = 0; i < 2; ++i) { c_main.src[i] = srcs[i]; c_main.dst[i] = dsts[i]; c_main.n[i] = ns[i]; }
    for (int i = 2; i < 8; ++i) { c_side.src[i - 2] = srcs[i]; c_side.dst[i - 2] = dsts[i]; c_side.n[i - 2] = srcs[i] ? ns[i] : 0; }   // (baseline: no sentinel weights)
    AA_PROF("cast_weights", st, launch_cast_multi(c_main, 2, st));
    // (the side lane forks behind the main lane's small cast: dispatched first, the 7000-CTA cast of the other weights kept
    //  the main lane's first kernels waiting for ~12 us in some replays, profiles/r01_v54_timeline.txt)
    AA_TRY(stream_dep(side, SIDE_EVENTS - 2, st, cs.st));
    AA_PROF("cast_weights", cs.st, launch_cast_multi(c_side, 6, cs.st));   // (capping its grid at ~2 CTAs per SM made the step slower: 389 us)
    AA_PROF("cast_inputs", cs.st, launch_cast2d(V, H, sv.V16, H, (long long)B * d->k, H, cs.st));
  }
  const Mat Wih = M2(w->w_ih, 2 * E, h.w_ih, 2 * E), Whh = M2(w->w_hh, H, h.w_hh, H);
  const Mat Wx = M2(w->sen_wx, 2 * E, h.sen_wx, 2 * E), Wh = M2(w->sen_wh, H, h.sen_wh, H);
  const Mat Wv = M2(w->att_wv, H, h.att_wv, H), Wg = M2(w->att_wg, H, h.att_wg, H), Ws = M2(w->att_ws, H, h.att_ws, H);
  const Mat Wp = M2(w->mlp_w, H, h.mlp_w, H);
  const Mat X = M2(sv.x, 2 * E, sv.x16, 2 * E);
  // Lane B, forked at once as well: x = [embed(w); v_g] (baseline_attention.py:151-154) next to the main lane's weight casts,
  // then what only the recurrence needs -- the bf16 copy of h0 and the zero-fills (h~_0 = 0 rows, absent initial states)
  const cudaStream_t so = side ? side->side2 : st;
  AA_TRY(stream_dep(side, SIDE_EVENTS - 8, st, so));
  AA_TRY(launch_build_x(cap, w->embed, v_g, sv.x, sv.x16, B, T, E, d->Vc, so, sv.hs_prev, tc ? sv.hsprev16 : nullptr, H));   // (+ h~_0 = 0 rows, Q2)
  AA_TRY(stream_dep(side, SIDE_EVENTS - 3, so, st));          // (x is built: the gate contractions of both lanes wait for it)
  if (side) AA_CHECK_CUDA(cudaStreamWaitEvent(cs.st, side->ev[SIDE_EVENTS - 3], 0));
  if (tc) {
    if (h0) AA_PROF("cast_inputs", so, launch_cast2d(h0, H, sv.h016, H, B, H, so));
    else AA_CHECK_CUDA(cudaMemsetAsync(sv.h016, 0, sizeof(bf16) * (size_t)B * H, so));
  }
  if (!h0 || !c0) AA_CHECK_CUDA(cudaMemsetAsync(sv.zeros, 0, sizeof(float) * (size_t)B * H, so));

  // input halves of the LSTM gates (main lane) and of the sentinel gate (side lane), batched over all T
  AA_TRY(mm_nt(cx, "gemm_gates_in", N, 4 * H, 2 * E, X, Wih, sv.xg, 4 * H, nullptr, 0, w->b_ih, w->b_hh));
  AA_TRY(mm_nt(cs, "gemm_P", B * d->k, d->a, H, M2(V, H, sv.V16, H), Wv, sv.P, d->a, nullptr, 0, nullptr, nullptr));   // :34
  if (!base) AA_TRY(mm_nt(cs, "gemm_gates_in", N, H, 2 * E, X, Wx, sv.g, H, nullptr, 0, nullptr, nullptr));
  if (base) {   // baseline model: the attention kernels (forward and backward) read s and r as zero rows, beta is forced to 0
    void* zp[3] = {(void*)sv.s, (void*)sv.r, tc ? (void*)sv.s16 : nullptr};
    const long long zb[3] = {(long long)sizeof(float) * N * H, (long long)sizeof(float) * N * d->a, (long long)sizeof(bf16) * N * H};
    AA_TRY(launch_zero_multi(3, zp, zb, cs.st));
  }
  AA_TRY(stream_dep(side, SIDE_EVENTS - 9, so, st));          // (h0 cast and zero-fills done)
  // recurrence                                                 baseline_attention.py:167-178
  const bool seq = tc && lstm_seq_supported(B, H, nullptr) && B <= 128 * 64;
  if (seq) {   // all T steps in one cooperative launch (lstm_seq.cu)
    LstmSeqFwd ls{};
    ls.B = B; ls.T = T; ls.H = H; ls.w_hh = w->w_hh; ls.xg = sv.xg; ls.c0 = c0; ls.h016 = sv.h016;
    ls.hiddens = sv.hiddens; ls.cells = sv.cells; ls.acts = sv.acts; ls.hs_prev = sv.hs_prev;
    ls.hid16 = sv.hid16; ls.hsprev16 = sv.hsprev16; ls.whh_packed16 = sv.whh_pack16; ls.counters = sv.counters;
    ls.whh16 = h.w_hh;
    {   // cluster kernels (weights in tensor memory) where the shape allows, else the grid-barrier kernel
      aa::ProfScope ps("lstm_seq_fwd", st);
      int rc = launch_lstm_cluster_fwd(ls, st);
      if (rc == AA_ERR_UNSUPPORTED) rc = launch_lstm_seq_fwd(ls, st);
      if (rc != AA_OK) return rc;
    }
  }
  for (int t = 0; t < T && !seq; ++t) {
    const Mat hp = t == 0 ? M2(h0 ? h0 : sv.zeros, H, sv.h016, H)
                          : M2(sv.hiddens + (size_t)(t - 1) * H, (long long)T * H, tc ? sv.hid16 + (size_t)(t - 1) * H : nullptr,
                               (long long)T * H);
    const float* cp = t == 0 ? (c0 ? c0 : sv.zeros) : sv.cells + (size_t)(t - 1) * H;
    const long long ldcp = t == 0 ? H : (long long)T * H;
    float* pre = sv.xg + (size_t)t * 4 * H;
    AA_TRY(mm_nt(cx, "lstm_rec_gemm", B, 4 * H, H, hp, Whh, pre, (long long)T * 4 * H, pre, (long long)T * 4 * H, nullptr, nullptr));
    AA_PROF("lstm_cell_fwd", st,
            launch_lstm_cell_fwd(pre, (long long)T * 4 * H, cp, ldcp, sv.acts + (size_t)t * 4 * H, (long long)T * 4 * H,
                                 sv.cells + (size_t)t * H, (long long)T * H, sv.hiddens + (size_t)t * H, (long long)T * H,
                                 t + 1 < T ? sv.hs_prev + (size_t)(t + 1) * H : nullptr, (long long)T * H,
                                 tc ? sv.hid16 + (size_t)t * H : nullptr, (tc && t + 1 < T) ? sv.hsprev16 + (size_t)(t + 1) * H : nullptr,
                                 B, H, st));
  }
  {   // final states out: nothing downstream reads them, lane B (joined by the caller of this body)
    AA_TRY(stream_dep(side, SIDE_EVENTS - 7, st, so));
    if (hT) AA_TRY(launch_copy2d(hT, H, sv.hiddens + (size_t)(T - 1) * H, (long long)T * H, B, H, so));
    if (cT) AA_TRY(launch_copy2d(cT, H, sv.cells + (size_t)(T - 1) * H, (long long)T * H, B, H, so));
  }
  AA_TRY(stream_dep(side, SIDE_EVENTS - 1, cs.st, st));       // main lane joins the side lane
  // q = h W_g^T only needs the hidden states: it runs on the side lane next to the sentinel's recurrent half
  AA_TRY(stream_dep(side, SIDE_EVENTS - 2, st, cs.st));
  AA_TRY(mm_nt(cs, "gemm_qr", N, d->a, H, M2(sv.hiddens, H, sv.hid16, H), Wg, sv.q, d->a, nullptr, 0, nullptr, nullptr));   // :35
  // sentinel                                                   adaptive_attention.py:116-125, 75-85
  if (T > 1 && !base) AA_TRY(mm_nt(cx, "gemm_sentinel_h", N, H, H, M2(sv.hs_prev, H, sv.hsprev16, H), Wh, sv.g, H, sv.g, H, nullptr, nullptr));
  if (!base) AA_TRY(launch_sentinel_fwd(sv.g, sv.cells, sv.g, sv.s, sv.s16, (long long)N * H, st));
  AA_TRY(stream_dep(side, SIDE_EVENTS - 1, cs.st, st));       // q is ready
  // attention + vocabulary projection                          adaptive_attention.py:128-132
  AA_TRY(atten_stage(cx, *d, Wv, Wg, Ws, w->att_wh, M2(V, H, sv.V16, H), M2(sv.hiddens, H, sv.hid16, H), M2(sv.s, H, sv.s16, H), sv.P,
                     sv.q, sv.r, nullptr, sv.ctx, sv.u, sv.u16, alpha, beta, /*have_P=*/true, /*have_q=*/true, base));
  if (row_index) {   // only the rows pack_padded_sequence keeps, already in packed order (baseline_attention.py:228, Q13)
    if (n_rows == 0) return AA_OK;
    // (fused loss: ce_dbp [Vc] and the loss accumulator behind it are zero-filled by the same launch)
    gather_rows2_kernel<<<(unsigned)n_rows, 128, 0, st>>>(sv.u, tc ? sv.u16 : nullptr, reinterpret_cast<const long long*>(row_index), H,
                                                          sv.up, tc ? sv.up16 : nullptr, ce ? sv.ce_dbp : nullptr, ce ? (long long)d->Vc + 1 : 0);
    AA_CHECK_LAUNCH("gather_rows2");
    if (ce) {
      // vocabulary projection with the cross-entropy in its epilogue (train.py:63,208): per 32-column chunk the row's maximum, sum of
      // exponentials and the exponentials as bf16; the [n_rows, Vc] logits are never written.  Then two small passes: chunk partials ->
      // log-sum-exp, loss and per-chunk scales; exponentials -> bf16 gradient of the logits (in place) + bias gradient.
      {
        ProfScope ps("gemm_vocab_fwd", st);
        TcGemmArgs g{};
        g.M = (int)n_rows; g.N = d->Vc; g.K = H; g.elem_size = 2;
        g.A = sv.up16; g.lda = H; g.B = Wp.h; g.ldb = Wp.ldh; g.bias1 = w->mlp_b;
        g.ce_e16 = sv.ce_e16; g.ld_ce = d->Vc; g.ce_part = sv.ce_part; g.ce_chunks = sv.ce_chunks;
        g.ce_tgt = reinterpret_cast<const long long*>(ce->targets); g.ce_xt = sv.ce_xt;
        AA_TRY(launch_gemm_tc(g, st));
      }
      AA_PROF("ce_merge", st, launch_ce_merge(sv.ce_part, sv.ce_chunks, (int)n_rows, sv.ce_xt, ce->denom, sv.ce_dbp + d->Vc, sv.ce_scale, st));
      AA_PROF("ce_fixup", st, launch_ce_fixup(sv.ce_e16, d->Vc, (int)n_rows, d->Vc, sv.ce_scale, sv.ce_chunks,
                                              reinterpret_cast<const long long*>(ce->targets), ce->denom, sv.ce_dbp, st, sv.ce_dbp + d->Vc, ce->loss));
      return AA_OK;
    }
    AA_TRY(mm_nt(cx, "gemm_vocab_fwd", (int)n_rows, d->Vc, H, M2(sv.up, H, sv.up16, H), Wp, scores, d->Vc, nullptr, 0, w->mlp_b, nullptr));
    return AA_OK;
  }
  AA_TRY(mm_nt(cx, "gemm_vocab_fwd", N, d->Vc, H, M2(sv.u, H, sv.u16, H), Wp, scores, d->Vc, nullptr, 0, w->mlp_b, nullptr));
  return AA_OK;
}

int aa_decoder_forward(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                       const float* h0, const float* c0, float* scores, float* alpha, float* beta, float* hT, float* cT,
                       void* saved, size_t saved_bytes, void* stream) {
  return decoder_forward_impl(d, w, V, v_g, captions, h0, c0, scores, alpha, beta, hT, cT, saved, saved_bytes, stream, nullptr, 0);
}

int aa_decoder_forward_packed(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                              const float* h0, const float* c0, const int64_t* row_index, int64_t n_rows, float* scores_packed,
                              float* alpha, float* beta, float* hT, float* cT, void* saved, size_t saved_bytes, void* stream) {
  AA_REQUIRE(row_index && d && n_rows >= 0 && n_rows <= (int64_t)d->B * d->T, "aa_decoder_forward_packed: bad row index / count");
  AA_REQUIRE(d->H % 4 == 0, "aa_decoder_forward_packed: H must be a multiple of 4");
  return decoder_forward_impl(d, w, V, v_g, captions, h0, c0, scores_packed, alpha, beta, hT, cT, saved, saved_bytes, stream, row_index,
                              n_rows);
}

int aa_decoder_forward_loss(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                            const float* h0, const float* c0, const int64_t* row_index, int64_t n_rows, const int64_t* targets,
                            int64_t denom, float* loss, float* alpha, float* beta, float* hT, float* cT, void* saved, size_t saved_bytes,
                            void* stream) {
  AA_REQUIRE(row_index && d && n_rows >= 0 && n_rows <= (int64_t)d->B * d->T, "aa_decoder_forward_loss: bad row index / count");
  AA_REQUIRE(targets && loss, "aa_decoder_forward_loss: null targets / loss");
  if (d->precision != AA_PREC_BF16) {
    set_error("aa_decoder_forward_loss: the loss is fused into the tcgen05 vocabulary contraction (AA_PREC_BF16); the exact path uses "
              "aa_decoder_forward_packed + aa_cross_entropy");
    return AA_ERR_UNSUPPORTED;
  }
  const CeFuse ce{targets, (long long)(denom > 0 ? denom : n_rows), loss};
  if (n_rows == 0) AA_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
  return decoder_forward_impl(d, w, V, v_g, captions, h0, c0, nullptr, alpha, beta, hT, cT, saved, saved_bytes, stream, row_index, n_rows, &ce);
}

int aa_decoder_loss_grad_scale(const aa_dims* d, void* saved, size_t saved_bytes, int64_t n_rows, const float* g, void* stream) {
  AA_TRY(check_dims(d, true));
  AA_REQUIRE(saved && g && d->precision == AA_PREC_BF16 && n_rows >= 0 && n_rows <= (int64_t)d->B * d->T, "aa_decoder_loss_grad_scale: bad arguments");
  AA_REQUIRE(saved_bytes >= aa_decoder_saved_bytes(d), "aa_decoder_loss_grad_scale: saved blob too small");
  Saved sv = carve_saved(*d, saved);
  return launch_scale_bf16_unless_one(sv.ce_e16, g, n_rows * (long long)d->Vc, sv.ce_dbp, d->Vc, (cudaStream_t)stream);
}

static int decoder_backward_body(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                                 const float* h0, const float* c0, const float* alpha, const float* beta, const void* saved,
                                 size_t saved_bytes, const float* d_scores, const float* d_alpha, const float* d_beta,
                                 const float* d_hT, const float* d_cT, const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0,
                                 float* dc0, void* scratch, size_t scratch_bytes, void* stream, void* const* ready_events,
                                 aa_grad_ready_fn on_ready, void* user, SideCtx* side, const int64_t* row_index, int64_t n_rows, const bf16* dS16_in);

static int decoder_backward_impl(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                                 const float* h0, const float* c0, const float* alpha, const float* beta, const void* saved,
                                 size_t saved_bytes, const float* d_scores, const float* d_alpha, const float* d_beta,
                                 const float* d_hT, const float* d_cT, const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0,
                                 float* dc0, void* scratch, size_t scratch_bytes, void* stream, void* const* ready_events,
                                 aa_grad_ready_fn on_ready, void* user, const int64_t* row_index = nullptr, int64_t n_rows = 0,
                                 const bf16* dS16_in = nullptr) {
  SideCtx* side = nullptr;
  AA_TRY(get_side(&side));
  cudaStream_t caller = (cudaStream_t)stream;
  cudaStream_t st = side ? side->crit : caller;
  AA_TRY(stream_dep(side, SIDE_EVENTS - 4, caller, st));
  AA_TRY(decoder_backward_body(d, w, V, v_g, captions, h0, c0, alpha, beta, saved, saved_bytes, d_scores, d_alpha, d_beta, d_hT, d_cT, gw, dV,
                               dv_g, dh0, dc0, scratch, scratch_bytes, (void*)st, ready_events, on_ready, user, side, row_index, n_rows, dS16_in));
  return stream_dep(side, SIDE_EVENTS - 6, st, caller);
}

static int decoder_backward_body(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                                 const float* h0, const float* c0, const float* alpha, const float* beta, const void* saved,
                                 size_t saved_bytes, const float* d_scores, const float* d_alpha, const float* d_beta,
                                 const float* d_hT, const float* d_cT, const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0,
                                 float* dc0, void* scratch, size_t scratch_bytes, void* stream, void* const* ready_events,
                                 aa_grad_ready_fn on_ready, void* user, SideCtx* side, const int64_t* row_index, int64_t n_rows, const bf16* dS16_in) {
  AA_TRY(check_dims(d, true));
  AA_REQUIRE(w && V && v_g && captions && alpha && beta && gw, "aa_decoder_backward: null pointer");
  // d_scores == NULL: the gradient of the packed logits is the one aa_decoder_forward_loss left in `saved` (bf16) next to the bias gradient
  const bool fused_ce = d_scores == nullptr;
  AA_REQUIRE(!fused_ce || (row_index && d->precision == AA_PREC_BF16), "aa_decoder_backward: d_scores is NULL (only the packed bf16 path after aa_decoder_forward_loss may omit it)");
  bool base = false;
  AA_TRY(check_sentinel_weights(w, &base));
  AA_REQUIRE(gw->embed && gw->w_ih && gw->w_hh && gw->b_ih && gw->b_hh && gw->att_wv && gw->att_wg && gw->att_wh && gw->mlp_w && gw->mlp_b &&
                 (base || (gw->sen_wx && gw->sen_wh && gw->att_ws)),
             "aa_decoder_backward: every parameter gradient buffer must be provided");
  if (d->B == 0) return AA_OK;
  if (!saved || saved_bytes < aa_decoder_saved_bytes(d)) {
    set_error("aa_decoder_backward: saved blob too small (%zu < %zu)", saved_bytes, aa_decoder_saved_bytes(d));
    return AA_ERR_WORKSPACE;
  }
  if (!scratch || scratch_bytes < aa_decoder_bwd_scratch_bytes(d)) {
    set_error("aa_decoder_backward: scratch too small (%zu < %zu)", scratch_bytes, aa_decoder_bwd_scratch_bytes(d));
    return AA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, T = d->T, H = d->H, E = d->E, a = d->a, k = d->k, Vc = d->Vc;
  const int N = B * T;
  const Saved sv = carve_saved(*d, const_cast<void*>(saved));
  BwdScratch sc = carve_bwd(*d, scratch);
  const long long* cap = reinterpret_cast<const long long*>(captions);
  float* dVb = dV ? dV : sc.dV;

  // Four lanes.  `cx` (the caller's stream) carries the critical path
  //   dS -> du -> attention backward -> ds -> sentinel backward -> dhs -> BPTT -> dx -> embedding scatter;
  // the library's side lanes A, B (and C for the bias sums) carry everything nothing on that path waits for: the 11
  // weight-gradient contractions, dV += dP W_v, dh = du + dq W_g (only the BPTT needs it) and the sentinel's dx.  One
  // side lane serialised those ~14 small contractions and finished ~60 us after the main lane (profiles/r01_v40_bench.json:
  // the side lane's kernels sum to more than the main lane's); small contractions leave most SMs idle, so independent
  // ones now run next to each other.
  const Ctx cx{d->precision, st};
  const cudaStream_t sd = side ? side->side : st, sb = side ? side->side2 : st, sl = side ? side->side3 : st;
  const Ctx cs{d->precision, sd}, cb{d->precision, sb};
  int evi = 0;
  auto dep = [&](cudaStream_t from, cudaStream_t to) { return stream_dep(side, evi++, from, to); };   // `to` waits for `from` so far
  auto to_side = [&]() { return dep(st, sd); };
  auto bucket_ready = [&](int bucket, cudaStream_t on) -> int {
    if (ready_events && ready_events[bucket]) AA_CHECK_CUDA(cudaEventRecord((cudaEvent_t)ready_events[bucket], on));
    if (on_ready) on_ready(bucket, user);
    return AA_OK;
  };
  const bool tc = cx.tc();
  const W16& h = sv.w16;
  const int ap = a_pad_of(*d);
  const Mat Whh = M2(w->w_hh, H, h.w_hh, H), Wih = M2(w->w_ih, 2 * E, h.w_ih, 2 * E);
  const Mat Wx = M2(w->sen_wx, 2 * E, h.sen_wx, 2 * E), Wh = M2(w->sen_wh, H, h.sen_wh, H);
  const Mat Wv = M2(w->att_wv, H, h.att_wv, H), Wg = M2(w->att_wg, H, h.att_wg, H), Ws = M2(w->att_ws, H, h.att_ws, H);
  const Mat Wp = M2(w->mlp_w, H, h.mlp_w, H);
  const Mat X = M2(sv.x, 2 * E, sv.x16, 2 * E);
  // Zero-fills of everything that is accumulated into later run next to the first kernels of the critical lane instead of as
  // memset nodes in front of their consumers.  What the critical lane waits for soon (du, dV, dP, att_wh) goes to a lane of its
  // own priority: on a low-priority lane these small kernels queued behind the pending CTAs of the first weight-gradient
  // contraction and held the attention backward up for 24 us (profiles/r01_v45_timeline.txt).  The rest (embedding and LSTM
  // weight gradients, the h0 rows of the dW_hh operand) is only needed after the BPTT: lane C.
  const cudaStream_t sz = side ? side->zero : st;
  AA_TRY(dep(st, sz));
  const int NR = row_index ? (int)n_rows : N;
  const bool pre_dx = tc && NR > 0;     // du's contraction adds onto a zero-filled output: no memset node in front of it
  {   // ONE kernel per group (memset nodes carry no priority and are dispatched one by one).  The output of du's contraction
      // is needed first and is small: the critical lane clears it itself instead of waiting for the whole group.
    const bool own_dup = pre_dx && row_index;       // (unpacked: du itself is the output and is in the group anyway)
    if (own_dup) {
      void* zp1[1] = {(void*)sc.dup};
      const long long zb1[1] = {(long long)sizeof(float) * NR * H};
      AA_TRY(launch_zero_multi(1, zp1, zb1, st));
    }
    void* zp[5] = {(pre_dx && !own_dup) ? (void*)sc.du : nullptr, row_index ? (void*)sc.du : nullptr, (void*)gw->att_wh, (void*)sc.dP,
                   (void*)dVb};
    const long long zb[5] = {(long long)sizeof(float) * NR * H, (long long)sizeof(float) * N * H, (long long)sizeof(float) * a,
                             (long long)sizeof(float) * B * k * a, (long long)sizeof(float) * B * k * H};
    AA_TRY(launch_zero_multi(5, zp, zb, sz));
  }
  AA_TRY(dep(st, sl));
  {
    void* zp[3] = {(void*)gw->embed, tc ? (void*)gw->w_ih : nullptr, tc ? (void*)gw->w_hh : nullptr};
    const long long zb[3] = {(long long)sizeof(float) * Vc * E, (long long)sizeof(float) * 4 * H * 2 * E, (long long)sizeof(float) * 4 * H * H};
    AA_TRY(launch_zero_multi(3, zp, zb, sl));
    if (tc && h0) AA_CHECK_CUDA(cudaMemcpyAsync(sv.hsprev16 + (size_t)N * H, sv.h016, sizeof(bf16) * (size_t)B * H, cudaMemcpyDeviceToDevice, sl));
  }
  const int ev_late = evi++;
  if (side) AA_CHECK_CUDA(cudaEventRecord(side->ev[ev_late], sl));
  // db_p = column sums of dS, fused with the bf16 cast of dS                     adaptive_attention.py:132
  // (packed entry point: d_scores holds the NR = n_rows packed rows only; the other positions have no gradient)
  // When the caller hands the bf16 mirror of dS along (aa_cross_entropy_mirror writes it in the loss kernel's own pass), nothing
  // of this is on the critical path: the column sums only feed the bias gradient and move to lane A.
  if (fused_ce) dS16_in = sv.ce_e16;
  const bool have16 = tc && dS16_in != nullptr && NR > 0;
  if (have16) {}
  else if (NR > 0) AA_PROF("colsum_cast_dS", st, launch_colsum_cast(d_scores, Vc, NR, Vc, gw->mlp_b, nullptr, tc ? sc.dS16 : nullptr, Vc, st));
  else AA_CHECK_CUDA(cudaMemsetAsync(gw->mlp_b, 0, sizeof(float) * Vc, st));
  const Mat dS = M2(d_scores, Vc, have16 ? dS16_in : sc.dS16, Vc);

  // vocabulary projection: u = c_hat + h                        adaptive_attention.py:132
  // (du is on the critical path: its contraction is enqueued before the weight gradient's so that it gets the SMs first)
  AA_TRY(to_side());
  if (!(pre_dx && row_index)) AA_TRY(dep(sz, st));             // (unpacked: du is zero-filled on the zero lane)
  if (row_index) {   // du rows of the packed positions, scattered back to [B,T,H] (zero elsewhere)
    if (NR > 0) AA_TRY(mm_nn(cx, "gemm_vocab_dx", NR, H, Vc, dS, Wp, sc.dup, H, pre_dx ? sc.dup : nullptr, H));
  } else {
    AA_TRY(mm_nn(cx, "gemm_vocab_dx", N, H, Vc, dS, Wp, sc.du, H, pre_dx ? sc.du : nullptr, H));
  }
  if (fused_ce) AA_CHECK_CUDA(cudaMemcpyAsync(gw->mlp_b, sv.ce_dbp, sizeof(float) * (size_t)Vc, cudaMemcpyDeviceToDevice, sd));
  else if (have16) AA_PROF("colsum_cast_dS", sd, launch_colsum(d_scores, Vc, NR, Vc, gw->mlp_b, nullptr, sd));
  if (NR > 0) AA_TRY(mm_tn(cs, "gemm_vocab_dw", Vc, H, NR, dS, row_index ? M2(sv.up, H, sv.up16, H) : M2(sv.u, H, sv.u16, H), gw->mlp_w, H, false));
  else AA_CHECK_CUDA(cudaMemsetAsync(gw->mlp_w, 0, sizeof(float) * (size_t)Vc * H, sd));
  AA_TRY(bucket_ready(AA_BUCKET_MLP, sd));
  if (pre_dx && row_index) AA_TRY(dep(sz, st));                // the zero-fills of du, dV, dP, att_wh are done
  if (row_index && NR > 0) {
    pack_rows_kernel<<<(unsigned)NR, 128, 0, st>>>(sc.dup, H, reinterpret_cast<const long long*>(row_index), sc.du, 0);
    AA_CHECK_LAUNCH("scatter_rows");
  }
  // attention                                                   adaptive_attention.py:34-56
  AttenBwdArgs ab{};
  ab.B = B; ab.T = T; ab.k = k; ab.a = a; ab.H = H;
  ab.P = sv.P; ab.q = sv.q; ab.r = sv.r; ab.s = sv.s; ab.V = V; ab.wh = w->att_wh;
  ab.alpha = alpha; ab.beta = beta; ab.ctx = sv.ctx; ab.dchat = sc.du; ab.d_alpha = d_alpha; ab.d_beta = d_beta;
  ab.ds = sc.ds; ab.dq = sc.dq; ab.dr = sc.dr; ab.dP = sc.dP; ab.dV = dVb; ab.dwh = gw->att_wh;
  ab.dq16 = tc ? sc.dq16 : nullptr; ab.dr16 = tc ? sc.dr16 : nullptr; ab.dP16 = nullptr; ab.a_pad = ap;
  ab.prezeroed = 1;
  AA_PROF("atten_bwd", st, launch_atten_bwd(ab, st));
  const Mat dR = M2(sc.dr, a, sc.dr16, ap), dQ = M2(sc.dq, a, sc.dq16, ap), dPm = M2(sc.dP, a, sc.dP16, ap);
  AA_TRY(to_side());
  AA_TRY(dep(st, sb));
  AA_TRY(mm_nn(cb, "gemm_att_dx", N, H, a, dQ, Wg, sc.du, H, sc.du, H));                              // dh = du + dq W_g  (lane B)
  const int ev_du = evi++;
  if (side) AA_CHECK_CUDA(cudaEventRecord(side->ev[ev_du], sb));
  if (!base) AA_TRY(mm_tn(cs, "gemm_att_dw", a, H, N, dR, M2(sv.s, H, sv.s16, H), gw->att_ws, H, false));
  AA_TRY(mm_tn(cs, "gemm_att_dw", a, H, N, dQ, M2(sv.hiddens, H, sv.hid16, H), gw->att_wg, H, false));
  if (tc) AA_PROF("cast_inputs", sb, launch_cast2d(sc.dP, a, sc.dP16, ap, (long long)B * k, a, sb));
  AA_TRY(mm_nn(cb, "gemm_att_dx", B * k, H, a, dPm, Wv, dVb, H, dVb, H));                             // dV += dP W_v
  AA_TRY(mm_tn(cb, "gemm_att_dw", a, H, B * k, dPm, M2(V, H, sv.V16, H), gw->att_wv, H, false));
  const int ev_dx = evi++;
  if (base) {   // baseline model: beta = 0 -> ds = 0 and dr = 0; no sentinel, so nothing reaches the cells or x from here
    void* zp[1] = {(void*)sc.dcell};
    const long long zb[1] = {(long long)sizeof(float) * N * H};
    AA_TRY(launch_zero_multi(1, zp, zb, st));
    if (side) AA_CHECK_CUDA(cudaEventRecord(side->ev[ev_dx], sd));
  } else {
  AA_TRY(mm_nn(cx, "gemm_att_dx", N, H, a, dR, Ws, sc.ds, H, sc.ds, H));                              // ds += dr W_s
  // sentinel                                                    adaptive_attention.py:79-83
  AA_TRY(launch_sentinel_bwd(sc.ds, sv.g, sv.cells, sc.da, sc.dcell, tc ? sc.da16 : nullptr, (long long)N * H, st));
  const Mat dA = M2(sc.da, H, sc.da16, H);
  AA_TRY(to_side());
  AA_TRY(dep(st, sb));
  AA_TRY(mm_nn(cs, "gemm_sent_dx", N, 2 * E, H, dA, Wx, sc.dx, 2 * E, nullptr, 0));
  if (side) AA_CHECK_CUDA(cudaEventRecord(side->ev[ev_dx], sd));
  AA_TRY(mm_tn(cs, "gemm_sent_dw", H, 2 * E, N, dA, X, gw->sen_wx, 2 * E, false));
  if (T > 1) {
    AA_TRY(mm_nn(cx, "gemm_sent_dx", N, H, H, dA, Wh, sc.dhs, H, nullptr, 0));
    AA_TRY(mm_tn(cb, "gemm_sent_dw", H, H, N, dA, M2(sv.hs_prev, H, sv.hsprev16, H), gw->sen_wh, H, false));
  } else {
    AA_CHECK_CUDA(cudaMemsetAsync(gw->sen_wh, 0, sizeof(float) * (size_t)H * H, sb));   // h~ = 0: no gradient (Q3)
  }
  }
  AA_TRY(dep(sb, sd));
  AA_TRY(bucket_ready(AA_BUCKET_ATTEN, sd));   // (att_wh was finished by atten_bwd, which the side lanes have waited for)
  if (side) AA_CHECK_CUDA(cudaStreamWaitEvent(st, side->ev[ev_du], 0));   // the BPTT reads dh
  // BPTT                                                        baseline_attention.py:167-178
  const bool seq = tc && lstm_seq_supported(B, H, nullptr) && B <= 128 * 64;
  bool t0_rows_copied = seq;      // the cluster kernel stores the step-0 rows of dgates16 a second time behind the array
  if (seq) {
    LstmSeqBwd ls{};
    ls.B = B; ls.T = T; ls.H = H; ls.w_hh = w->w_hh;
    ls.dh_attn = sc.du; ls.dhs = (T > 1 && !base) ? sc.dhs : nullptr; ls.dcell = sc.dcell; ls.d_hT = d_hT; ls.d_cT = d_cT;
    ls.acts = sv.acts; ls.cells = sv.cells; ls.c0 = c0; ls.dgates = sc.dgates; ls.dgates16 = sc.dgates16;
    ls.dh0 = dh0; ls.dc0 = dc0; ls.whhT16 = sc.whhT16; ls.counters = sc.counters;
    ls.whh16 = sv.w16.w_hh;
    ls.dgates16_t0 = (tc && h0) ? sc.dgates16 + (size_t)N * 4 * H : nullptr;   // (K-tail rows of the dW_hh contraction, see below)
    {
      aa::ProfScope ps("lstm_seq_bwd", st);
      int rc = launch_lstm_cluster_bwd(ls, st);
      if (rc == AA_ERR_UNSUPPORTED) {
        rc = launch_lstm_seq_bwd(ls, st);
        t0_rows_copied = false;
      }
      if (rc != AA_OK) return rc;
    }
  }
  for (int t = T - 1; t >= 0 && !seq; --t) {
    const float* dh_rec_in = t == T - 1 ? d_hT : sc.dh_rec;
    const float* dc_rec_in = t == T - 1 ? d_cT : sc.dc_rec;
    const float* dhs_next = (T > 1 && t + 1 < T && !base) ? sc.dhs + (size_t)(t + 1) * H : nullptr;
    const float* cp = t == 0 ? (c0 ? c0 : sv.zeros) : sv.cells + (size_t)(t - 1) * H;
    const long long ldcp = t == 0 ? H : (long long)T * H;
    AA_PROF("lstm_cell_bwd", st,
            launch_lstm_cell_bwd(sc.du + (size_t)t * H, (long long)T * H, dhs_next, (long long)T * H, dh_rec_in,
                                 sc.dcell + (size_t)t * H, (long long)T * H, dc_rec_in, sv.acts + (size_t)t * 4 * H,
                                 (long long)T * 4 * H, sv.cells + (size_t)t * H, (long long)T * H, cp, ldcp,
                                 sc.dgates + (size_t)t * 4 * H, (long long)T * 4 * H, tc ? sc.dgates16 + (size_t)t * 4 * H : nullptr,
                                 sc.dc_rec, B, H, st));
    AA_TRY(mm_nn(cx, "bptt_rec_gemm", B, H, 4 * H,
                 M2(sc.dgates + (size_t)t * 4 * H, (long long)T * 4 * H, tc ? sc.dgates16 + (size_t)t * 4 * H : nullptr, (long long)T * 4 * H),
                 Whh, sc.dh_rec, H, nullptr, 0));
  }
  if (dh0 && !seq) AA_TRY(launch_copy2d(dh0, H, sc.dh_rec, H, B, H, st));
  if (dc0 && !seq) AA_TRY(launch_copy2d(dc0, H, sc.dc_rec, H, B, H, st));
  // the recurrence (the latency-critical kernel of the backward) is enqueued: a caller that wants its gradient exchange to stay
  // off the SMs / memory system while it runs starts the exchange on this event instead of on the buckets' own
  AA_TRY(bucket_ready(AA_EVENT_BPTT_DONE, st));
  // LSTM parameter gradients, batched over all steps: one lane each
  const Mat dG = M2(sc.dgates, 4 * H, sc.dgates16, 4 * H);
  AA_TRY(to_side());
  AA_TRY(dep(st, sb));
  AA_TRY(dep(st, sl));
  if (side) {   // (the late zero-fills of lane C)
    AA_CHECK_CUDA(cudaStreamWaitEvent(sd, side->ev[ev_late], 0));
    AA_CHECK_CUDA(cudaStreamWaitEvent(sb, side->ev[ev_late], 0));
    AA_CHECK_CUDA(cudaStreamWaitEvent(st, side->ev[ev_late], 0));
  }
  if (tc) {
    // (outputs zero-filled on lane C at the start: accumulate, no memset nodes here.)  The step-0 term dgates_0^T h0 of dW_hh
    // rides along as B extra K rows: h0 behind the rows of h~ and the step-0 rows of dgates behind the rows of dgates
    AA_TRY(mm_tn(cs, "gemm_lstm_dw", 4 * H, 2 * E, N, dG, X, gw->w_ih, 2 * E, true));
    int Kh = N;
    if (h0 && t0_rows_copied) Kh = N + B;
    else if (h0) {
      AA_CHECK_CUDA(cudaMemcpy2DAsync(sc.dgates16 + (size_t)N * 4 * H, sizeof(bf16) * (size_t)4 * H, sc.dgates16, sizeof(bf16) * (size_t)T * 4 * H,
                                      sizeof(bf16) * (size_t)4 * H, (size_t)B, cudaMemcpyDeviceToDevice, sb));
      Kh = N + B;
    }
    AA_TRY(mm_tn(cb, "gemm_lstm_dw", 4 * H, H, Kh, dG, M2(sv.hs_prev, H, sv.hsprev16, H), gw->w_hh, H, true));
  } else {
    AA_TRY(mm_tn(cs, "gemm_lstm_dw", 4 * H, 2 * E, N, dG, X, gw->w_ih, 2 * E, false));
    AA_TRY(mm_tn(cb, "gemm_lstm_dw", 4 * H, H, N, dG, M2(sv.hs_prev, H, sv.hsprev16, H), gw->w_hh, H, false));   // steps t >= 1 (h~_0 rows are 0)
    if (h0)                                                                                                      // step 0
      AA_TRY(mm_tn(cb, "gemm_lstm_dw", 4 * H, H, B, M2(sc.dgates, (long long)T * 4 * H, sc.dgates16, (long long)T * 4 * H),
                   M2(h0, H, sv.h016, H), gw->w_hh, H, true));
  }
  AA_PROF("colsum", sl, launch_colsum(sc.dgates, 4 * H, N, 4 * H, gw->b_ih, gw->b_hh, sl));
  AA_TRY(dep(sb, sd));
  AA_TRY(dep(sl, sd));
  AA_TRY(bucket_ready(AA_BUCKET_LSTM, sd));
  // dx += dgates W_ih needs the sentinel's dx from the side lane
  if (side) AA_CHECK_CUDA(cudaStreamWaitEvent(st, side->ev[ev_dx], 0));
  AA_TRY(mm_nn(cx, "gemm_lstm_dx", N, 2 * E, 4 * H, dG, Wih, sc.dx, 2 * E, base ? nullptr : sc.dx, 2 * E));   // (baseline: no sentinel dx to add onto)
  // x = [embed(w); v_g]                                         baseline_attention.py:151-154   (gw->embed was zero-filled on lane C)
  AA_TRY(launch_embed_bwd(cap, sc.dx, gw->embed, dv_g, B, T, E, Vc, st));
  AA_TRY(bucket_ready(AA_BUCKET_EMBED, st));
  return dep(sd, st);   // join (lane A has joined B and C): the call is complete, in stream order, when `stream` says so
}

int aa_decoder_backward(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                        const float* h0, const float* c0, const float* alpha, const float* beta, const void* saved,
                        size_t saved_bytes, const float* d_scores, const float* d_alpha, const float* d_beta,
                        const float* d_hT, const float* d_cT, const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0,
                        float* dc0, void* scratch, size_t scratch_bytes, void* stream) {
  return decoder_backward_impl(d, w, V, v_g, captions, h0, c0, alpha, beta, saved, saved_bytes, d_scores, d_alpha, d_beta, d_hT,
                               d_cT, gw, dV, dv_g, dh0, dc0, scratch, scratch_bytes, stream, nullptr, nullptr, nullptr);
}

int aa_decoder_backward_hooked(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                               const float* h0, const float* c0, const float* alpha, const float* beta, const void* saved,
                               size_t saved_bytes, const float* d_scores, const float* d_alpha, const float* d_beta,
                               const float* d_hT, const float* d_cT, const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0,
                               float* dc0, void* scratch, size_t scratch_bytes, void* stream, void* const* ready_events,
                               aa_grad_ready_fn on_ready, void* user) {
  return decoder_backward_impl(d, w, V, v_g, captions, h0, c0, alpha, beta, saved, saved_bytes, d_scores, d_alpha, d_beta, d_hT,
                               d_cT, gw, dV, dv_g, dh0, dc0, scratch, scratch_bytes, stream, ready_events, on_ready, user);
}

int aa_decoder_backward_packed(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                               const float* h0, const float* c0, const float* alpha, const float* beta, const void* saved,
                               size_t saved_bytes, const int64_t* row_index, int64_t n_rows, const float* d_scores_packed,
                               const float* d_alpha, const float* d_beta, const float* d_hT, const float* d_cT,
                               const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0, float* dc0, void* scratch,
                               size_t scratch_bytes, void* stream, void* const* ready_events, aa_grad_ready_fn on_ready, void* user,
                               const void* d_scores_packed_bf16) {
  AA_REQUIRE(row_index && d && n_rows >= 0 && n_rows <= (int64_t)d->B * d->T, "aa_decoder_backward_packed: bad row index / count");
  return decoder_backward_impl(d, w, V, v_g, captions, h0, c0, alpha, beta, saved, saved_bytes, d_scores_packed, d_alpha, d_beta, d_hT,
                               d_cT, gw, dV, dv_g, dh0, dc0, scratch, scratch_bytes, stream, ready_events, on_ready, user, row_index,
                               n_rows, static_cast<const bf16*>(d_scores_packed_bf16));
}

int aa_pack_rows(const float* scores, int64_t n_cols, const int64_t* row_index, int64_t n_rows, float* packed, void* stream) {
  if (n_rows == 0) return AA_OK;
  AA_REQUIRE(scores && row_index && packed && n_cols > 0, "aa_pack_rows: bad argument");
  pack_rows_kernel<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(scores, n_cols, reinterpret_cast<const long long*>(row_index),
                                                                       packed, 1);
  AA_CHECK_LAUNCH("pack_rows");
  return AA_OK;
}

int aa_unpack_rows(const float* d_packed, int64_t n_cols, const int64_t* row_index, int64_t n_rows, int64_t total_rows,
                   float* d_scores, void* stream) {
  AA_REQUIRE(d_scores && n_cols > 0 && total_rows >= n_rows, "aa_unpack_rows: bad argument");
  AA_CHECK_CUDA(cudaMemsetAsync(d_scores, 0, sizeof(float) * (size_t)total_rows * n_cols, (cudaStream_t)stream));
  if (n_rows == 0) return AA_OK;
  AA_REQUIRE(d_packed && row_index, "aa_unpack_rows: null pointer");
  pack_rows_kernel<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(d_packed, n_cols, reinterpret_cast<const long long*>(row_index),
                                                                       d_scores, 0);
  AA_CHECK_LAUNCH("unpack_rows");
  return AA_OK;
}

int aa_cross_entropy_mirror(const float* logits, int64_t n_rows, int64_t Vc, const int64_t* targets, int64_t denom, float* loss,
                            float* dlogits, void* dlogits_bf16, int* mirror_written, void* stream) {
  if (mirror_written) *mirror_written = 0;
  AA_REQUIRE(loss, "aa_cross_entropy: loss is NULL");
  {   // (a kernel, not a memset node: inside a captured graph the small memset in front of the loss kernel cost a ~9 us gap)
    void* zp[1] = {loss};
    const long long zb[1] = {(long long)sizeof(float)};
    AA_TRY(launch_zero_multi(1, zp, zb, (cudaStream_t)stream));
  }
  if (n_rows == 0) return AA_OK;
  AA_REQUIRE(logits && targets, "aa_cross_entropy: null pointer");
  return launch_ce_fwd_bwd(logits, Vc, reinterpret_cast<const long long*>(targets), (int)n_rows, (int)Vc, loss, dlogits, Vc, denom,
                           (cudaStream_t)stream, static_cast<bf16*>(dlogits_bf16), mirror_written);
}

int aa_cross_entropy_denom(const float* logits, int64_t n_rows, int64_t Vc, const int64_t* targets, int64_t denom, float* loss,
                           float* dlogits, void* stream) {
  return aa_cross_entropy_mirror(logits, n_rows, Vc, targets, denom, loss, dlogits, nullptr, nullptr, stream);
}

int aa_cross_entropy(const float* logits, int64_t n_rows, int64_t Vc, const int64_t* targets, float* loss, float* dlogits,
                     void* stream) {
  return aa_cross_entropy_denom(logits, n_rows, Vc, targets, n_rows, loss, dlogits, stream);
}

int aa_scale_unless_one(float* x, const float* g, int64_t n, void* x_bf16, void* stream) {
  AA_REQUIRE(n >= 0 && (n == 0 || (x && g)), "aa_scale_unless_one: bad argument");
  return launch_scale_unless_one(x, g, (long long)n, (cudaStream_t)stream, static_cast<bf16*>(x_bf16));
}

int aa_copy_multi(int n_segments, const void* const* src, void* const* dst, const int64_t* bytes, void* stream) {
  AA_REQUIRE(n_segments >= 0 && n_segments <= 8 && (n_segments == 0 || (src && dst && bytes)), "aa_copy_multi: bad argument");
  long long b[8];
  for (int i = 0; i < n_segments; ++i) {
    AA_REQUIRE(bytes[i] >= 0 && (bytes[i] == 0 || (src[i] && dst[i])), "aa_copy_multi: bad segment %d", i);
    b[i] = (long long)bytes[i];
  }
  return launch_copy_multi(n_segments, src, dst, b, (cudaStream_t)stream);
}

}  // extern "C"
