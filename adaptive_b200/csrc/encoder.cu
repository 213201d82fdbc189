// Encoder heads of the reference's AttentiveCNN (baseline_attention.py:21-34, 46-62) -- everything after the ResNet trunk,
// which is out of scope: last-conv feature map A [B,C,h*w] (NCHW) -> V = relu(A^T W_a^T + b_a) [B,hw,H],
// v_g = relu(a_g W_b^T + b_b) [B,E], h0 = tanh(a_g W_h0^T + b_h0), c0 = tanh(a_g W_c0^T + b_c0) with a_g = mean over the map.
// SURVEY.md section 8f rank 2.  One pass over A produces its [B,hw,C] transpose (the K-major operand of the W_a contraction,
// bf16 in mixed-precision mode) together with the pooled row; the contractions run on the same engines as the decoder's
// (exact fp32 SIMT, or bf16 on tcgen05), the activations and their adjoints are one multi-segment pointwise launch each.
#include "../../include/adaptive_b200.h"
#include "host_common.cuh"
#include "kernels.cuh"

namespace aa {
namespace {

constexpr int ENC_TC = 64;       // channels per CTA of the transposing kernels
constexpr int ENC_THREADS = 256;

// A [B,C,hw] -> At [B,hw,C] (fp32 and/or bf16) + a_g [B,C] = mean_p A[b,c,p] (fixed summation order).
// One CTA per (64-channel slab, image): the slab is contiguous in A (64*hw floats) and lands as hw segments of 64 channels.
__global__ void __launch_bounds__(ENC_THREADS) enc_transpose_pool_kernel(const float* __restrict__ A, int C, int hw, float* __restrict__ At,
                                                                         bf16* __restrict__ At16, float* __restrict__ ag,
                                                                         bf16* __restrict__ ag16) {
  extern __shared__ float tile[];   // [ENC_TC][hw + 1]
  const int c0 = blockIdx.x * ENC_TC, b = blockIdx.y;
  const int nc = min(ENC_TC, C - c0), ld = hw + 1;
  const float* src = A + ((size_t)b * C + c0) * hw;
  for (int i = threadIdx.x; i < nc * hw; i += ENC_THREADS) tile[(i / hw) * ld + (i % hw)] = __ldg(src + i);
  __syncthreads();
  if (threadIdx.x < nc) {
    float s = 0.f;
    for (int p = 0; p < hw; ++p) s += tile[threadIdx.x * ld + p];
    s /= (float)hw;
    ag[(size_t)b * C + c0 + threadIdx.x] = s;
    if (ag16) ag16[(size_t)b * C + c0 + threadIdx.x] = __float2bfloat16(s);
  }
  for (int i = threadIdx.x; i < hw * ENC_TC; i += ENC_THREADS) {
    const int p = i / ENC_TC, c = i % ENC_TC;
    if (c < nc) {
      const float v = tile[c * ld + p];
      const size_t o = ((size_t)b * hw + p) * C + c0 + c;
      if (At) At[o] = v;
      if (At16) At16[o] = __float2bfloat16(v);
    }
  }
}

// dA [B,C,hw] = dAt [B,hw,C]^T + dag [B,C] / hw      (adjoint of the transpose and of the average pool)
__global__ void __launch_bounds__(ENC_THREADS) enc_untranspose_kernel(const float* __restrict__ dAt, const float* __restrict__ dag, int C,
                                                                      int hw, float* __restrict__ dA) {
  extern __shared__ float tile[];   // [ENC_TC][hw + 1]
  const int c0 = blockIdx.x * ENC_TC, b = blockIdx.y;
  const int nc = min(ENC_TC, C - c0), ld = hw + 1;
  for (int i = threadIdx.x; i < hw * ENC_TC; i += ENC_THREADS) {
    const int p = i / ENC_TC, c = i % ENC_TC;
    if (c < nc) tile[c * ld + p] = __ldg(dAt + ((size_t)b * hw + p) * C + c0 + c);
  }
  __syncthreads();
  const float inv = 1.f / (float)hw;
  float* dst = dA + ((size_t)b * C + c0) * hw;
  for (int i = threadIdx.x; i < nc * hw; i += ENC_THREADS) {
    const int c = i / hw;
    dst[i] = tile[c * ld + (i % hw)] + __ldg(dag + (size_t)b * C + c0 + c) * inv;
  }
}

struct ActSegs {
  float* y[4];            // in place: y = act(y)                                  (forward)
  const float* out[4];    // forward outputs                                      (backward)
  const float* up[4];     // upstream gradients
  float* dpre[4];         // gradient w.r.t. the pre-activation
  bf16* dpre16[4];        // optional bf16 mirror
  long long n[4];
  int tanh_mode[4];       // 0 = relu, 1 = tanh
};

__global__ void enc_act_fwd_kernel(ActSegs a, int nsegs) {
  const int sgm = blockIdx.y;
  if (sgm >= nsegs) return;
  float* y = a.y[sgm];
  const long long n = a.n[sgm];
  const bool th = a.tanh_mode[sgm] != 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = y[i];
    y[i] = th ? tanhf(v) : fmaxf(v, 0.f);
  }
}

// relu: dpre = up * (out > 0);  tanh: dpre = up * (1 - out^2)
__global__ void enc_act_bwd_kernel(ActSegs a, int nsegs) {
  const int sgm = blockIdx.y;
  if (sgm >= nsegs) return;
  const float* out = a.out[sgm];
  const float* up = a.up[sgm];
  float* d = a.dpre[sgm];
  bf16* d16 = a.dpre16[sgm];
  const long long n = a.n[sgm];
  const bool th = a.tanh_mode[sgm] != 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float o = out[i], u = up[i];
    const float g = th ? u * (1.f - o * o) : (o > 0.f ? u : 0.f);
    d[i] = g;
    if (d16) d16[i] = __float2bfloat16(g);
  }
}

size_t enc_tile_bytes(int hw) { return sizeof(float) * (size_t)ENC_TC * (hw + 1); }

int ensure_tile_smem(int hw) {
  static size_t granted_fwd = 48 * 1024, granted_bwd = 48 * 1024;
  const size_t need = enc_tile_bytes(hw);
  AA_REQUIRE(need <= 200 * 1024, "encoder: feature map of %d positions is too large for the transposing kernels", hw);
  if (need > granted_fwd) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(enc_transpose_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    granted_fwd = need;
  }
  if (need > granted_bwd) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(enc_untranspose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    granted_bwd = need;
  }
  return AA_OK;
}

struct EncSaved {
  float *At, *ag;
  bf16 *At16, *ag16, *wa16, *wb16, *wh016, *wc016;
  size_t bytes;
};

EncSaved carve_enc_saved(const aa_enc_dims& d, void* base) {
  const size_t M = (size_t)d.B * d.hw, C = d.C;
  Carver c(base);
  EncSaved s{};
  s.ag = c.take((size_t)d.B * C);
  if (d.precision == AA_PREC_BF16) {
    Carver16 h{c};
    s.At16 = h.take(M * C);
    s.ag16 = h.take((size_t)d.B * C);
    s.wa16 = h.take((size_t)d.H * C);
    s.wb16 = h.take((size_t)d.E * C);
    s.wh016 = h.take((size_t)d.H * C);
    s.wc016 = h.take((size_t)d.H * C);
  } else {
    s.At = c.take(M * C);
  }
  s.bytes = c.off;
  return s;
}

struct EncScratch {
  float *dVp, *dvgp, *dh0p, *dc0p, *dAt, *dag;
  bf16 *dVp16, *dvgp16, *dh0p16, *dc0p16;
  size_t bytes;
};

EncScratch carve_enc_scratch(const aa_enc_dims& d, int want_dA, void* base) {
  const size_t M = (size_t)d.B * d.hw, C = d.C;
  Carver c(base);
  EncScratch s{};
  s.dVp = c.take(M * d.H);
  s.dvgp = c.take((size_t)d.B * d.E);
  s.dh0p = c.take((size_t)d.B * d.H);
  s.dc0p = c.take((size_t)d.B * d.H);
  if (want_dA) {
    s.dAt = c.take(M * C);
    s.dag = c.take((size_t)d.B * C);
  }
  if (d.precision == AA_PREC_BF16) {
    Carver16 h{c};
    s.dVp16 = h.take(M * d.H);
    s.dvgp16 = h.take((size_t)d.B * d.E);
    s.dh0p16 = h.take((size_t)d.B * d.H);
    s.dc0p16 = h.take((size_t)d.B * d.H);
  }
  s.bytes = c.off;
  return s;
}

int check_enc_dims(const aa_enc_dims* d) {
  AA_REQUIRE(d != nullptr, "encoder dims is NULL");
  AA_REQUIRE(d->B >= 0 && d->C >= 4 && d->hw >= 1 && d->H >= 4 && d->E >= 4, "bad encoder dims B=%d C=%d hw=%d H=%d E=%d", d->B, d->C,
             d->hw, d->H, d->E);
  AA_REQUIRE(d->C % 4 == 0 && d->H % 4 == 0 && d->E % 4 == 0, "encoder: C, H and E must be multiples of 4 (C=%d H=%d E=%d)", d->C, d->H,
             d->E);
  AA_REQUIRE(d->precision == AA_PREC_FP32 || d->precision == AA_PREC_BF16, "encoder: unknown precision %d", d->precision);
  if (d->precision == AA_PREC_BF16)
    AA_REQUIRE(d->C % 8 == 0 && d->H % 8 == 0 && d->E % 8 == 0, "encoder: bf16 mode needs C, H and E to be multiples of 8 (C=%d H=%d E=%d)",
               d->C, d->H, d->E);
  return AA_OK;
}

}  // namespace
}  // namespace aa

using namespace aa;

extern "C" {

size_t aa_encoder_saved_bytes(const aa_enc_dims* d) { return d ? carve_enc_saved(*d, nullptr).bytes : 0; }
size_t aa_encoder_bwd_scratch_bytes(const aa_enc_dims* d, int want_dA) { return d ? carve_enc_scratch(*d, want_dA, nullptr).bytes : 0; }

int aa_encoder_forward(const aa_enc_dims* d, const aa_enc_weights* w, const float* A, float* V, float* v_g, float* h0, float* c0,
                       void* saved, size_t saved_bytes, void* stream) {
  AA_TRY(check_enc_dims(d));
  AA_REQUIRE(w && A && V && v_g && h0 && c0, "aa_encoder_forward: null pointer");
  AA_REQUIRE(w->wa && w->ba && w->wb && w->bb && w->wh0 && w->bh0 && w->wc0 && w->bc0, "aa_encoder_forward: null weight pointer");
  if (d->B == 0) return AA_OK;
  if (!saved || saved_bytes < aa_encoder_saved_bytes(d)) {
    set_error("aa_encoder_forward: saved blob too small (%zu < %zu)", saved_bytes, aa_encoder_saved_bytes(d));
    return AA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, C = d->C, hw = d->hw, H = d->H, E = d->E, M = B * hw;
  const EncSaved sv = carve_enc_saved(*d, saved);
  const Ctx cx{d->precision, st};
  const bool tc = cx.tc();
  AA_TRY(ensure_tile_smem(hw));
  {   // a_g (baseline_attention.py:46-47) and the [B,hw,C] view of :50 in one pass over A
    ProfScope ps("enc_transpose_pool", st);
    enc_transpose_pool_kernel<<<dim3(ceil_div(C, ENC_TC), B), ENC_THREADS, enc_tile_bytes(hw), st>>>(A, C, hw, sv.At, sv.At16, sv.ag, sv.ag16);
    AA_CHECK_LAUNCH("enc_transpose_pool");
  }
  if (tc) {
    CastSegs cs{};
    const float* srcs[4] = {w->wa, w->wb, w->wh0, w->wc0};
    bf16* dsts[4] = {sv.wa16, sv.wb16, sv.wh016, sv.wc016};
    const long long ns[4] = {(long long)H * C, (long long)E * C, (long long)H * C, (long long)H * C};
    for (int i = 0; i < 4; ++i) { cs.src[i] = srcs[i]; cs.dst[i] = dsts[i]; cs.n[i] = ns[i]; }
    AA_PROF("enc_cast_weights", st, launch_cast_multi(cs, 4, st));
  }
  const Mat At = M2(sv.At, C, sv.At16, C), Ag = M2(sv.ag, C, sv.ag16, C);
  AA_TRY(mm_nt(cx, "enc_gemm_V", M, H, C, At, M2(w->wa, C, sv.wa16, C), V, H, nullptr, 0, w->ba, nullptr));          // :51
  AA_TRY(mm_nt(cx, "enc_gemm_heads", B, E, C, Ag, M2(w->wb, C, sv.wb16, C), v_g, E, nullptr, 0, w->bb, nullptr));    // :53
  AA_TRY(mm_nt(cx, "enc_gemm_heads", B, H, C, Ag, M2(w->wh0, C, sv.wh016, C), h0, H, nullptr, 0, w->bh0, nullptr));  // :56
  AA_TRY(mm_nt(cx, "enc_gemm_heads", B, H, C, Ag, M2(w->wc0, C, sv.wc016, C), c0, H, nullptr, 0, w->bc0, nullptr));  // :58
  ActSegs a{};
  float* ys[4] = {V, v_g, h0, c0};
  const long long ns[4] = {(long long)M * H, (long long)B * E, (long long)B * H, (long long)B * H};
  for (int i = 0; i < 4; ++i) { a.y[i] = ys[i]; a.n[i] = ns[i]; a.tanh_mode[i] = i >= 2; }
  {
    ProfScope ps("enc_act", st);
    const int gx = (int)std::min<long long>((ns[0] + 255) / 256, 4LL * num_sms());
    enc_act_fwd_kernel<<<dim3(gx, 4), 256, 0, st>>>(a, 4);
    AA_CHECK_LAUNCH("enc_act_fwd");
  }
  return AA_OK;
}

int aa_encoder_backward(const aa_enc_dims* d, const aa_enc_weights* w, const void* saved, size_t saved_bytes, const float* V,
                        const float* v_g, const float* h0, const float* c0, const float* dV, const float* dv_g, const float* dh0,
                        const float* dc0, const aa_enc_weight_grads* gw, float* dA, void* scratch, size_t scratch_bytes, void* stream) {
  AA_TRY(check_enc_dims(d));
  AA_REQUIRE(w && V && v_g && h0 && c0 && dV && dv_g && dh0 && dc0 && gw, "aa_encoder_backward: null pointer");
  AA_REQUIRE(gw->wa && gw->ba && gw->wb && gw->bb && gw->wh0 && gw->bh0 && gw->wc0 && gw->bc0,
             "aa_encoder_backward: every parameter gradient buffer must be provided");
  if (d->B == 0) return AA_OK;
  if (!saved || saved_bytes < aa_encoder_saved_bytes(d)) {
    set_error("aa_encoder_backward: saved blob too small (%zu < %zu)", saved_bytes, aa_encoder_saved_bytes(d));
    return AA_ERR_WORKSPACE;
  }
  const int want_dA = dA != nullptr;
  if (!scratch || scratch_bytes < aa_encoder_bwd_scratch_bytes(d, want_dA)) {
    set_error("aa_encoder_backward: scratch too small (%zu < %zu)", scratch_bytes, aa_encoder_bwd_scratch_bytes(d, want_dA));
    return AA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, C = d->C, hw = d->hw, H = d->H, E = d->E, M = B * hw;
  const EncSaved sv = carve_enc_saved(*d, const_cast<void*>(saved));
  const EncScratch sc = carve_enc_scratch(*d, want_dA, scratch);
  const Ctx cx{d->precision, st};
  const bool tc = cx.tc();
  ActSegs a{};
  const float* outs[4] = {V, v_g, h0, c0};
  const float* ups[4] = {dV, dv_g, dh0, dc0};
  float* dps[4] = {sc.dVp, sc.dvgp, sc.dh0p, sc.dc0p};
  bf16* dps16[4] = {sc.dVp16, sc.dvgp16, sc.dh0p16, sc.dc0p16};
  const long long ns[4] = {(long long)M * H, (long long)B * E, (long long)B * H, (long long)B * H};
  for (int i = 0; i < 4; ++i) {
    a.out[i] = outs[i]; a.up[i] = ups[i]; a.dpre[i] = dps[i]; a.dpre16[i] = tc ? dps16[i] : nullptr; a.n[i] = ns[i]; a.tanh_mode[i] = i >= 2;
  }
  {
    ProfScope ps("enc_act", st);
    const int gx = (int)std::min<long long>((ns[0] + 255) / 256, 4LL * num_sms());
    enc_act_bwd_kernel<<<dim3(gx, 4), 256, 0, st>>>(a, 4);
    AA_CHECK_LAUNCH("enc_act_bwd");
  }
  const Mat At = M2(sv.At, C, sv.At16, C), Ag = M2(sv.ag, C, sv.ag16, C);
  const Mat dVp = M2(sc.dVp, H, sc.dVp16, H), dvgp = M2(sc.dvgp, E, sc.dvgp16, E), dh0p = M2(sc.dh0p, H, sc.dh0p16, H),
            dc0p = M2(sc.dc0p, H, sc.dc0p16, H);
  const Mat Wa = M2(w->wa, C, sv.wa16, C), Wb = M2(w->wb, C, sv.wb16, C), Wh0 = M2(w->wh0, C, sv.wh016, C), Wc0 = M2(w->wc0, C, sv.wc016, C);
  // weight and bias gradients
  AA_TRY(mm_tn(cx, "enc_gemm_dWa", H, C, M, dVp, At, gw->wa, C, false));
  AA_TRY(mm_tn(cx, "enc_gemm_dheads", E, C, B, dvgp, Ag, gw->wb, C, false));
  AA_TRY(mm_tn(cx, "enc_gemm_dheads", H, C, B, dh0p, Ag, gw->wh0, C, false));
  AA_TRY(mm_tn(cx, "enc_gemm_dheads", H, C, B, dc0p, Ag, gw->wc0, C, false));
  AA_PROF("enc_colsum", st, launch_colsum(sc.dVp, H, M, H, gw->ba, nullptr, st));
  AA_TRY(launch_colsum(sc.dvgp, E, B, E, gw->bb, nullptr, st));
  AA_TRY(launch_colsum(sc.dh0p, H, B, H, gw->bh0, nullptr, st));
  AA_TRY(launch_colsum(sc.dc0p, H, B, H, gw->bc0, nullptr, st));
  if (want_dA) {   // gradient towards the trunk (the reference fine-tunes it after cf.fine_tune_start_epoch)
    AA_TRY(mm_nn(cx, "enc_gemm_dA", M, C, H, dVp, Wa, sc.dAt, C, nullptr, 0));
    AA_TRY(mm_nn(cx, "enc_gemm_dheads", B, C, E, dvgp, Wb, sc.dag, C, nullptr, 0));
    AA_TRY(mm_nn(cx, "enc_gemm_dheads", B, C, H, dh0p, Wh0, sc.dag, C, sc.dag, C));
    AA_TRY(mm_nn(cx, "enc_gemm_dheads", B, C, H, dc0p, Wc0, sc.dag, C, sc.dag, C));
    AA_TRY(ensure_tile_smem(hw));
    ProfScope ps("enc_untranspose", st);
    enc_untranspose_kernel<<<dim3(ceil_div(C, ENC_TC), B), ENC_THREADS, enc_tile_bytes(hw), st>>>(sc.dAt, sc.dag, C, hw, dA);
    AA_CHECK_LAUNCH("enc_untranspose");
  }
  return AA_OK;
}

}  // extern "C"
