// extern "C" boundary of libadaptive_sm100.so (see include/adaptive_b200.h) and the host-side
// orchestration of the teacher-forced forward/backward.  Decoding lives in decode_api.cu.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/adaptive_b200.h"
#include "kernels.cuh"

namespace aa {

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

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  float* take(size_t nfloats) {
    float* p = reinterpret_cast<float*>(base + off);
    off += align_up(nfloats * sizeof(float), 256);
    return p;
  }
};

struct Saved {
  float *x, *xg, *acts, *cells, *hiddens, *hs_prev, *g, *s, *P, *q, *r, *ctx, *u, *zeros;
  size_t bytes;
};

Saved carve_saved(const aa_dims& d, void* base) {
  const size_t N = (size_t)d.B * d.T, H = d.H, E = d.E;
  Carver c(base);
  Saved s;
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
  s.bytes = c.off;
  return s;
}

struct BwdScratch {
  float *du, *ds, *dq, *dr, *dP, *da, *dcell, *dhs, *dgates, *dx, *dh_rec, *dc_rec, *dV;
  size_t bytes;
};

BwdScratch carve_bwd(const aa_dims& d, void* base) {
  const size_t N = (size_t)d.B * d.T, H = d.H, E = d.E;
  Carver c(base);
  BwdScratch s;
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
  return AA_OK;
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
static int atten_stage(const aa_dims& d, const float* att_wv, const float* att_wg, const float* att_ws, const float* att_wh,
                       const float* V, const float* h, const float* s, float* P, float* q, float* r, float* c_hat,
                       float* ctx, float* u, float* alpha, float* beta, cudaStream_t st) {
  const int N = d.B * d.T;
  AA_PROF("gemm_P", st, gemm_nt(d.B * d.k, d.a, d.H, V, d.H, att_wv, d.H, P, d.a, nullptr, 0, nullptr, nullptr, st));   // :34
  AA_PROF("gemm_qr", st, gemm_nt(N, d.a, d.H, h, d.H, att_wg, d.H, q, d.a, nullptr, 0, nullptr, nullptr, st));           // :35
  AA_PROF("gemm_qr", st, gemm_nt(N, d.a, d.H, s, d.H, att_ws, d.H, r, d.a, q, d.a, nullptr, nullptr, st));               // :45
  AttenFwdArgs p{};
  p.B = d.B; p.T = d.T; p.k = d.k; p.a = d.a; p.H = d.H;
  p.P = P; p.q = q; p.r = r; p.s = s; p.h = h; p.V = V; p.wh = att_wh;
  p.alpha = alpha; p.beta = beta; p.ctx = ctx; p.u = u; p.c_hat = c_hat;
  AA_PROF("atten_fwd", st, launch_atten_fwd(p, st));
  return AA_OK;
}

}  // namespace aa

using namespace aa;

extern "C" {

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
  return launch_sentinel_fwd(gate_out, cell, gate_out, s_out, (long long)N * d->H, st);
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
  AA_REQUIRE(att_wv && att_wg && att_ws && att_wh && V && h_t && s_t && c_hat && alpha && beta, "aa_atten_forward: null pointer");
  if (workspace_bytes < aa_atten_workspace_bytes(d) || !workspace) {
    set_error("aa_atten_forward: workspace too small (%zu < %zu)", workspace_bytes, aa_atten_workspace_bytes(d));
    return AA_ERR_WORKSPACE;
  }
  Carver c(workspace);
  const size_t N = (size_t)d->B * d->T;
  float* P = c.take((size_t)d->B * d->k * d->a);
  float* q = c.take(N * d->a);
  float* r = c.take(N * d->a);
  return atten_stage(*d, att_wv, att_wg, att_ws, att_wh, V, h_t, s_t, P, q, r, c_hat, nullptr, nullptr, alpha, beta,
                     (cudaStream_t)stream);
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
  // h~: zero-h0 shift of adaptive_attention.py:116-122
  AA_TRY(gemm_nt((int)N, H, 2 * d->E, x, 2 * d->E, w->sen_wx, 2 * d->E, g, H, nullptr, 0, nullptr, nullptr, st));
  if (T > 1) {
    AA_CHECK_CUDA(cudaMemset2DAsync(hs_prev, (size_t)T * H * 4, 0, (size_t)H * 4, d->B, st));
    AA_TRY(launch_copy2d(hs_prev + H, (long long)T * H, hiddens, (long long)T * H, d->B, (T - 1) * H, st));
    AA_TRY(gemm_nt((int)N, H, H, hs_prev, H, w->sen_wh, H, g, H, g, H, nullptr, nullptr, st));
  }
  AA_TRY(launch_sentinel_fwd(g, cells, g, s, (long long)N * H, st));
  AA_TRY(atten_stage(*d, w->att_wv, w->att_wg, w->att_ws, w->att_wh, V, hiddens, s, P, q, r, nullptr, nullptr, u, alpha,
                     beta, st));
  return gemm_nt((int)N, d->Vc, H, u, H, w->mlp_w, H, scores, d->Vc, nullptr, 0, w->mlp_b, nullptr, st);   // :132
}

size_t aa_decoder_saved_bytes(const aa_dims* d) { return d ? carve_saved(*d, nullptr).bytes : 0; }
size_t aa_decoder_bwd_scratch_bytes(const aa_dims* d) { return d ? carve_bwd(*d, nullptr).bytes : 0; }

int aa_decoder_forward(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                       const float* h0, const float* c0, float* scores, float* alpha, float* beta, float* hT, float* cT,
                       void* saved, size_t saved_bytes, void* stream) {
  AA_TRY(check_dims(d, true));
  AA_REQUIRE(w && V && v_g && captions && scores && alpha && beta, "aa_decoder_forward: null pointer");
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

  // x = [embed(w); v_g]                                        baseline_attention.py:151-154
  AA_TRY(launch_build_x(cap, w->embed, v_g, sv.x, B, T, E, d->Vc, st));
  // input halves of the LSTM gates and of the sentinel gate, batched over all T
  AA_PROF("gemm_gates_in", st, gemm_nt(N, 4 * H, 2 * E, sv.x, 2 * E, w->w_ih, 2 * E, sv.xg, 4 * H, nullptr, 0, w->b_ih, w->b_hh, st));
  AA_PROF("gemm_gates_in", st, gemm_nt(N, H, 2 * E, sv.x, 2 * E, w->sen_wx, 2 * E, sv.g, H, nullptr, 0, nullptr, nullptr, st));
  if (!h0 || !c0) AA_CHECK_CUDA(cudaMemsetAsync(sv.zeros, 0, sizeof(float) * (size_t)B * H, st));
  AA_CHECK_CUDA(cudaMemset2DAsync(sv.hs_prev, (size_t)T * H * 4, 0, (size_t)H * 4, B, st));   // h~_0 = 0 (Q2)
  // recurrence                                                 baseline_attention.py:167-178
  for (int t = 0; t < T; ++t) {
    const float* hp = t == 0 ? (h0 ? h0 : sv.zeros) : sv.hiddens + (size_t)(t - 1) * H;
    const long long ldhp = t == 0 ? H : (long long)T * H;
    const float* cp = t == 0 ? (c0 ? c0 : sv.zeros) : sv.cells + (size_t)(t - 1) * H;
    const long long ldcp = t == 0 ? H : (long long)T * H;
    float* pre = sv.xg + (size_t)t * 4 * H;
    AA_PROF("lstm_rec_gemm", st, gemm_nt(B, 4 * H, H, hp, ldhp, w->w_hh, H, pre, (long long)T * 4 * H, pre, (long long)T * 4 * H, nullptr, nullptr, st));
    AA_PROF("lstm_cell_fwd", st, launch_lstm_cell_fwd(pre, (long long)T * 4 * H, cp, ldcp, sv.acts + (size_t)t * 4 * H, (long long)T * 4 * H,
                                sv.cells + (size_t)t * H, (long long)T * H, sv.hiddens + (size_t)t * H, (long long)T * H,
                                t + 1 < T ? sv.hs_prev + (size_t)(t + 1) * H : nullptr, (long long)T * H, B, H, st));
  }
  if (hT) AA_TRY(launch_copy2d(hT, H, sv.hiddens + (size_t)(T - 1) * H, (long long)T * H, B, H, st));
  if (cT) AA_TRY(launch_copy2d(cT, H, sv.cells + (size_t)(T - 1) * H, (long long)T * H, B, H, st));
  // sentinel                                                   adaptive_attention.py:116-125, 75-85
  if (T > 1) AA_TRY(gemm_nt(N, H, H, sv.hs_prev, H, w->sen_wh, H, sv.g, H, sv.g, H, nullptr, nullptr, st));
  AA_TRY(launch_sentinel_fwd(sv.g, sv.cells, sv.g, sv.s, (long long)N * H, st));
  // attention + vocabulary projection                          adaptive_attention.py:128-132
  AA_TRY(atten_stage(*d, w->att_wv, w->att_wg, w->att_ws, w->att_wh, V, sv.hiddens, sv.s, sv.P, sv.q, sv.r, nullptr, sv.ctx,
                     sv.u, alpha, beta, st));
  AA_PROF("gemm_vocab_fwd", st, gemm_nt(N, d->Vc, H, sv.u, H, w->mlp_w, H, scores, d->Vc, nullptr, 0, w->mlp_b, nullptr, st));
  return AA_OK;
}

int aa_decoder_backward(const aa_dims* d, const aa_weights* w, const float* V, const float* v_g, const int64_t* captions,
                        const float* h0, const float* c0, const float* alpha, const float* beta, const void* saved,
                        size_t saved_bytes, const float* d_scores, const float* d_alpha, const float* d_beta,
                        const float* d_hT, const float* d_cT, const aa_weight_grads* gw, float* dV, float* dv_g, float* dh0,
                        float* dc0, void* scratch, size_t scratch_bytes, void* stream) {
  AA_TRY(check_dims(d, true));
  AA_REQUIRE(w && V && v_g && captions && alpha && beta && d_scores && gw, "aa_decoder_backward: null pointer");
  AA_REQUIRE(gw->embed && gw->w_ih && gw->w_hh && gw->b_ih && gw->b_hh && gw->sen_wx && gw->sen_wh && gw->att_wv &&
                 gw->att_wg && gw->att_ws && gw->att_wh && gw->mlp_w && gw->mlp_b,
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

  // vocabulary projection: u = c_hat + h                        adaptive_attention.py:132
  AA_PROF("gemm_vocab_dx", st, gemm_nn(N, H, Vc, d_scores, Vc, w->mlp_w, H, sc.du, H, nullptr, 0, st));
  AA_PROF("gemm_vocab_dw", st, gemm_tn(Vc, H, N, d_scores, Vc, sv.u, H, gw->mlp_w, H, false, st));
  AA_PROF("colsum", st, launch_colsum(d_scores, Vc, N, Vc, gw->mlp_b, nullptr, st));
  // attention                                                   adaptive_attention.py:34-56
  AA_CHECK_CUDA(cudaMemsetAsync(gw->att_wh, 0, sizeof(float) * a, st));
  AttenBwdArgs ab{};
  ab.B = B; ab.T = T; ab.k = k; ab.a = a; ab.H = H;
  ab.P = sv.P; ab.q = sv.q; ab.r = sv.r; ab.s = sv.s; ab.V = V; ab.wh = w->att_wh;
  ab.alpha = alpha; ab.beta = beta; ab.ctx = sv.ctx; ab.dchat = sc.du; ab.d_alpha = d_alpha; ab.d_beta = d_beta;
  ab.ds = sc.ds; ab.dq = sc.dq; ab.dr = sc.dr; ab.dP = sc.dP; ab.dV = dVb; ab.dwh = gw->att_wh;
  AA_PROF("atten_bwd", st, launch_atten_bwd(ab, st));
  AA_TRY(gemm_nn(N, H, a, sc.dr, a, w->att_ws, H, sc.ds, H, sc.ds, H, st));            // ds += dr W_s
  AA_TRY(gemm_tn(a, H, N, sc.dr, a, sv.s, H, gw->att_ws, H, false, st));
  AA_TRY(gemm_nn(N, H, a, sc.dq, a, w->att_wg, H, sc.du, H, sc.du, H, st));            // dh = du + dq W_g
  AA_TRY(gemm_tn(a, H, N, sc.dq, a, sv.hiddens, H, gw->att_wg, H, false, st));
  AA_TRY(gemm_nn(B * k, H, a, sc.dP, a, w->att_wv, H, dVb, H, dVb, H, st));            // dV += dP W_v
  AA_TRY(gemm_tn(a, H, B * k, sc.dP, a, V, H, gw->att_wv, H, false, st));
  // sentinel                                                    adaptive_attention.py:79-83
  AA_TRY(launch_sentinel_bwd(sc.ds, sv.g, sv.cells, sc.da, sc.dcell, (long long)N * H, st));
  AA_TRY(gemm_nn(N, 2 * E, H, sc.da, H, w->sen_wx, 2 * E, sc.dx, 2 * E, nullptr, 0, st));
  AA_TRY(gemm_tn(H, 2 * E, N, sc.da, H, sv.x, 2 * E, gw->sen_wx, 2 * E, false, st));
  if (T > 1) {
    AA_TRY(gemm_nn(N, H, H, sc.da, H, w->sen_wh, H, sc.dhs, H, nullptr, 0, st));
    AA_TRY(gemm_tn(H, H, N, sc.da, H, sv.hs_prev, H, gw->sen_wh, H, false, st));
  } else {
    AA_CHECK_CUDA(cudaMemsetAsync(gw->sen_wh, 0, sizeof(float) * (size_t)H * H, st));   // h~ = 0: no gradient (Q3)
  }
  // BPTT                                                        baseline_attention.py:167-178
  for (int t = T - 1; t >= 0; --t) {
    const float* dh_rec_in = t == T - 1 ? d_hT : sc.dh_rec;
    const float* dc_rec_in = t == T - 1 ? d_cT : sc.dc_rec;
    const float* dhs_next = (T > 1 && t + 1 < T) ? sc.dhs + (size_t)(t + 1) * H : nullptr;
    const float* cp = t == 0 ? (c0 ? c0 : sv.zeros) : sv.cells + (size_t)(t - 1) * H;
    const long long ldcp = t == 0 ? H : (long long)T * H;
    AA_PROF("lstm_cell_bwd", st, launch_lstm_cell_bwd(sc.du + (size_t)t * H, (long long)T * H, dhs_next, (long long)T * H, dh_rec_in,
                                sc.dcell + (size_t)t * H, (long long)T * H, dc_rec_in, sv.acts + (size_t)t * 4 * H,
                                (long long)T * 4 * H, sv.cells + (size_t)t * H, (long long)T * H, cp, ldcp,
                                sc.dgates + (size_t)t * 4 * H, (long long)T * 4 * H, sc.dc_rec, B, H, st));
    AA_PROF("bptt_rec_gemm", st, gemm_nn(B, H, 4 * H, sc.dgates + (size_t)t * 4 * H, (long long)T * 4 * H, w->w_hh, H, sc.dh_rec, H, nullptr, 0, st));
  }
  if (dh0) AA_TRY(launch_copy2d(dh0, H, sc.dh_rec, H, B, H, st));
  if (dc0) AA_TRY(launch_copy2d(dc0, H, sc.dc_rec, H, B, H, st));
  // LSTM parameter gradients, batched over all steps
  AA_PROF("gemm_lstm_dw", st, gemm_tn(4 * H, 2 * E, N, sc.dgates, 4 * H, sv.x, 2 * E, gw->w_ih, 2 * E, false, st));
  AA_PROF("gemm_lstm_dw", st, gemm_tn(4 * H, H, N, sc.dgates, 4 * H, sv.hs_prev, H, gw->w_hh, H, false, st));   // steps t >= 1 (h~_0 rows are 0)
  if (h0) AA_TRY(gemm_tn(4 * H, H, B, sc.dgates, (long long)T * 4 * H, h0, H, gw->w_hh, H, true, st));   // step 0
  AA_TRY(launch_colsum(sc.dgates, 4 * H, N, 4 * H, gw->b_ih, gw->b_hh, st));
  AA_PROF("gemm_lstm_dx", st, gemm_nn(N, 2 * E, 4 * H, sc.dgates, 4 * H, w->w_ih, 2 * E, sc.dx, 2 * E, sc.dx, 2 * E, st));   // dx += dgates W_ih
  // x = [embed(w); v_g]                                         baseline_attention.py:151-154
  AA_CHECK_CUDA(cudaMemsetAsync(gw->embed, 0, sizeof(float) * (size_t)Vc * E, st));
  return launch_embed_bwd(cap, sc.dx, gw->embed, dv_g, B, T, E, Vc, st);
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

int aa_cross_entropy(const float* logits, int64_t n_rows, int64_t Vc, const int64_t* targets, float* loss, float* dlogits,
                     void* stream) {
  AA_REQUIRE(loss, "aa_cross_entropy: loss is NULL");
  AA_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
  if (n_rows == 0) return AA_OK;
  AA_REQUIRE(logits && targets, "aa_cross_entropy: null pointer");
  return launch_ce_fwd_bwd(logits, Vc, reinterpret_cast<const long long*>(targets), (int)n_rows, (int)Vc, loss, dlogits, Vc,
                           (cudaStream_t)stream);
}

}  // extern "C"
