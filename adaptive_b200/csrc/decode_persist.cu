// Persistent greedy decoder: the WHOLE sampler loop of Encoder2Decoder.sampler (adaptive_attention.py:186-216) in ONE
// cooperative launch, for batches of at most one image per SM.
//
// The per-step pipeline of decode_api.cu re-reads every image's V (k*H*4 = 100 KB at cfgA) from HBM on every step and pays
// ~9 kernel boundaries per step -- invisible at B = 4096, dominant at the reference's evaluation batch sizes.  Here
//   * CTA b owns image b: V_b and P_b = V_b W_v^T are loaded into shared memory ONCE and stay there for all L steps
//     (24 130 instead of 128 588 algorithmic bytes per image and step), the cell state c never leaves the SM;
//   * the two batch-wide contractions of a step run on tcgen05 inside the same kernel, distributed over ALL CTAs with the
//     WEIGHTS as the M = 128 side and the (padded) batch as the N side of the MMA, so that one instruction covers every image:
//       G1  recurrent gate terms  [4H x B] = W_hh (tf32 hi | lo) . h^T     3xTF32 (fp32-accurate), K split over CTAs,
//           64-byte-swizzled TMA tiles (a 128-byte-swizzled stage of four sub-tiles would not fit next to the resident V);
//       G2  approximate logits    [Vc x B] = W_p (bf16) . u^T              one bf16 pass (128-byte swizzle);
//     the INPUT half of the gates never needs a contraction at decode time: x_t = [emb(w_t); v_g], so W_ih[:, :E] emb(w) is a
//     row of the table EG = embed . [W_ih[:, :E]; W_x[:, :E]]^T (built once per set of weights, exact fp32) and the v_g half
//     is the per-image static term.  That removes the dependency of the gate contraction on the word just chosen: G1 of step
//     t+1 only needs h_t and runs in the SAME phase as G2 of step t;
//   * everything between is the owner CTA's business, in exact fp32 (FMA) arithmetic, one uninterrupted phase per step:
//       O2(t)   filter-and-refine arg-max per COLUMN: with |approx_j - exact_j| <= ||u - u^|| ||W_j|| + ||u^|| ||W_j - W^_j|| (the exact
//               decomposition of the two bf16 roundings, vocab_refine.cu)
//               only columns with approx_j + bound_j >= max_j (approx_j - bound_j) can hold the maximum; those few are recomputed
//               exactly from the fp32 u kept in shared memory, lowest index wins ties;
//       O1(t+1) gates = EG[word] + static + K-split partials -> LSTM cell -> sentinel -> q / r mat-vecs -> scores, both softmaxes
//               -> context over the resident V -> u = c_hat + h;
//   * TWO grid barriers per step ([G2(t) | G1(t+1)] -> [O2(t), O1(t+1)]) replace the ~9 launches.
// Nothing of a step but the K-split partials (L2), h and u (operands, L2), the approximate logits (L2) and the outputs
// (ids, alpha, beta) leaves the SM.
#include <stdlib.h>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace aa {
namespace {

using namespace tc;

constexpr int PD_THREADS = 384;          // warp 0: TMA, warp 1: MMA issue, warps 2-5: accumulator drain; all 12 warps in the owner phases
constexpr int PD_WARPS = PD_THREADS / 32;
constexpr int PD_STAGES = 3;
constexpr uint32_t PD_A_BYTES = 128 * 128;   // per stage: G1 = weight rows 128 x (hi 64 B | lo 64 B), G2 = 128 rows x 128 B
constexpr int PD_MAX_CAND = 2048;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) { return make_smem_desc(saddr, 16, 1024); }
// K-major, 64-byte swizzle: rows of 64 B, 8-row groups 512 B apart, layout type 4 (same encoding as lstm_cluster.cu)
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// optional per-step timeline of CTA 0 (aa_debug_set_persist_trace): 8 x uint64 globaltimer stamps per step --
//   0 contraction phase entered   1 this CTA's units done   2 grid barrier passed   3 arg-max done (word known)
//   4 next step's owner work done (u written)   5 second grid barrier passed
__device__ unsigned long long* g_pd_trace = nullptr;
__device__ __forceinline__ void pd_trace(int step, int ev) {
  if (g_pd_trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_pd_trace[step * 8 + ev] = t;
  }
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void pd_grid_wait(const unsigned* counter, unsigned target) {
  long long t0 = 0;
  while (ld_acquire_gpu(counter) < target) {
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > AA_SPIN_LIMIT_CYCLES) __trap();
  }
}

// All CTAs are co-resident (cooperative launch).  Monotonic arrival counter; thread 0 (= the TMA-issuing lane) performs the
// release / acquire, the proxy fences order the generic-proxy stores of this phase against the TMA reads of the next one.
__device__ __forceinline__ void pd_grid_sync(unsigned* bar, unsigned& epoch) {
  fence_proxy_async_global();
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    __threadfence();
    red_release_gpu_add(bar, 1u);
    pd_grid_wait(bar, epoch * gridDim.x);
    fence_proxy_async_global();
  }
  __syncthreads();
}

struct PdSmem {
  uint8_t* ring;        // PD_STAGES stages (GEMM phases) / scratch (owner phases)
  float* V;             // [k, H]   resident
  float* P;             // [k, ldP] resident
  float* c;             // [H]      cell state, resident
  float* u;             // [H]      u = c_hat + h of the current step (O1 -> O2)
  float* red;           // [64] reductions
  uint64_t* full;       // [PD_STAGES]
  uint64_t* empty;      // [PD_STAGES]
  uint64_t* tfull;      // [2]
  uint64_t* tempty;     // [2]
  uint32_t* tmem_slot;
};

__device__ __forceinline__ float block_max(float v, float* red, int tid) {
  v = warp_max(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < PD_WARPS; ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red, int tid) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < PD_WARPS; ++i) r += red[i];      // fixed order: deterministic
  return r;
}

// One distributed contraction phase over a MIXED unit list: units [0, units2) are G2 tiles (bf16, 128-byte swizzle, output
// out2[b * ldv + j] = acc + bias[j]); units [units2, units2 + tiles1 * ks1) are G1 (tile, K range) pairs (tf32 hi/lo sub-tiles,
// 64-byte swizzle, 3 MMAs per k-step, output = raw partial accumulators part1[(ks * B + b) * M1 + j]).  Either list may be empty.
// M = weight rows (tiles of 128), N = NB padded batch columns, k-blocks of 16 floats (G1) / 64 bf16 (G2).
struct PdPhase {
  int units2, tiles1, ks1, kbper1, nkb1, nkb2;
  bool g1, g2;
};

__device__ __forceinline__ void pd_gemm_phase(const PdSmem& sm, const DecodePersistArgs& p, const PdPhase& ph, const CUtensorMap* tmA1,
                                              const CUtensorMap* tmB1, const CUtensorMap* tmA2, const CUtensorMap* tmB2, uint32_t tmem_base,
                                              int& it, int& lt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NB = p.NB, B = p.B, M1 = 4 * p.H;
  const int n2 = ph.g2 ? ph.units2 : 0;
  const int n1 = ph.g1 ? ph.tiles1 * ph.ks1 : 0;
  // unit -> CTA: G2 tiles on CTAs 0, 1, ...; G1 units on the CTAs after them (wrapping), so that the two lists overlap as
  // little as the grid allows
  const int G = gridDim.x, me = blockIdx.x;
  const uint32_t b_half = (uint32_t)NB * 64u;                 // G1: bytes of one B sub-tile (hi or lo)
  const uint32_t stage_bytes = PD_A_BYTES + (uint32_t)NB * 128u;
  // my units, in order: G2 tiles me, me + G, ...; then G1 units u with (n2 + u) % G == me
  const int first1 = ((me - n2) % G + G) % G;
  if (warp == 0) {
    if (lane == 0) {
      int i = it;
      for (int unit = me; unit < n2; unit += G) {
        const int m0 = unit * 128;
        for (int kb = 0; kb < ph.nkb2; ++kb, ++i) {
          const int s = i % PD_STAGES;
          mbar_wait(&sm.empty[s], ((i / PD_STAGES) & 1) ^ 1);
          mbar_expect_tx(&sm.full[s], stage_bytes);
          uint8_t* a_dst = sm.ring + (size_t)s * stage_bytes;
          tma_load_2d(a_dst, tmA2, kb * 64, m0, &sm.full[s]);
          tma_load_2d(a_dst + PD_A_BYTES, tmB2, kb * 64, 0, &sm.full[s]);
        }
      }
      for (int unit = first1; unit < n1; unit += G) {
        const int m0 = (unit % ph.tiles1) * 128, ks = unit / ph.tiles1;
        const int kb0 = ks * ph.kbper1, kb1 = min(ph.nkb1, kb0 + ph.kbper1);
        for (int kb = kb0; kb < kb1; ++kb, ++i) {
          const int s = i % PD_STAGES;
          mbar_wait(&sm.empty[s], ((i / PD_STAGES) & 1) ^ 1);
          mbar_expect_tx(&sm.full[s], stage_bytes);
          uint8_t* a_dst = sm.ring + (size_t)s * stage_bytes;
          uint8_t* b_dst = a_dst + PD_A_BYTES;
          tma_load_2d(a_dst, tmA1, kb * 16, m0, &sm.full[s]);
          tma_load_2d(a_dst + PD_A_BYTES / 2, tmA1, p.K1p + kb * 16, m0, &sm.full[s]);
          tma_load_2d(b_dst, tmB1, kb * 16, 0, &sm.full[s]);
          tma_load_2d(b_dst + b_half, tmB1, p.lo1 + kb * 16, 0, &sm.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D = f32, A/B format (bf16 = 1, tf32 = 2), both K-major, N >> 3 at bits 17-22, M >> 4 at bits 24-28
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int i = it, l = lt;
    for (int unit = me; unit < n2; unit += G, ++l) {
      const int acc = l & 1;
      mbar_wait(&sm.tempty[acc], ((l >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < ph.nkb2; ++kb, ++i) {
        const int s = i % PD_STAGES;
        mbar_wait(&sm.full[s], (i / PD_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sm.ring + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + PD_A_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)        // k-step = 16 bf16 = 32 B inside the 128-byte row
            tc_mma<false>(tmem_d, desc_sw128(a_addr + k * 32), desc_sw128(b_addr + k * 32), idesc2, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&sm.empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&sm.tfull[acc]);
      __syncwarp();
    }
    for (int unit = first1; unit < n1; unit += G, ++l) {
      const int ks = unit / ph.tiles1;
      const int kb0 = ks * ph.kbper1, kb1 = min(ph.nkb1, kb0 + ph.kbper1);
      const int acc = l & 1;
      mbar_wait(&sm.tempty[acc], ((l >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 256);
      for (int kb = kb0; kb < kb1; ++kb, ++i) {
        const int s = i % PD_STAGES;
        mbar_wait(&sm.full[s], (i / PD_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sm.ring + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + PD_A_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {      // k-step = 8 tf32 = 32 B inside the 64-byte row
            const uint64_t ah = desc_sw64(a_addr + k * 32), al = desc_sw64(a_addr + PD_A_BYTES / 2 + k * 32);
            const uint64_t bh = desc_sw64(b_addr + k * 32), bl = desc_sw64(b_addr + b_half + k * 32);
            tc_mma<true>(tmem_d, al, bh, idesc1, ((kb - kb0) | k) != 0 ? 1u : 0u);
            tc_mma<true>(tmem_d, ah, bl, idesc1, 1u);
            tc_mma<true>(tmem_d, ah, bh, idesc1, 1u);
          }
          tc_commit(&sm.empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&sm.tfull[acc]);
      __syncwarp();
    }
  } else if (warp < 6) {
    // accumulator drain: TMEM lane = weight row j (coalesced over the 32 lanes for a fixed image b), columns = images
    const int q = warp & 3;
    int l = lt;
    for (int pass = 0; pass < 2; ++pass) {
      const bool is2 = pass == 0;
      const int nun = is2 ? n2 : n1;
      for (int unit = is2 ? me : first1; unit < nun; unit += G, ++l) {
        const int m0 = (is2 ? unit : unit % ph.tiles1) * 128, ks = is2 ? 0 : unit / ph.tiles1;
        const int M = is2 ? p.Vc : M1;
        const int acc = l & 1;
        mbar_wait(&sm.tfull[acc], (l >> 1) & 1);
        tc_fence_after();
        const int j = m0 + q * 32 + lane;
        const float bj = (is2 && j < M) ? __ldg(p.bp + j) : 0.f;
        float* out = is2 ? p.approx : p.part1 + (size_t)ks * B * M1;
        const long long ldo = is2 ? p.ldv : M1;
        for (int c = 0; c * 32 < NB; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + c * 32), r);
          if (j < M) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int b = c * 32 + e;
              if (b < B) out[(long long)b * ldo + j] = __uint_as_float(r[e]) + bj;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.tempty[acc]);
      }
    }
  }
  // every role advances the shared counters identically
  int n_it = 0, n_lt = 0;
  for (int unit = me; unit < n2; unit += G) { n_it += ph.nkb2; ++n_lt; }
  for (int unit = first1; unit < n1; unit += G) {
    const int ks = unit / ph.tiles1;
    n_it += min(ph.nkb1, ks * ph.kbper1 + ph.kbper1) - ks * ph.kbper1;
    ++n_lt;
  }
  it += n_it;
  lt += n_lt;
}

__global__ void __launch_bounds__(PD_THREADS, 1)
dec_persist_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmB2, const DecodePersistArgs p) {
  extern __shared__ uint8_t pd_raw[];
  uint8_t* base = pd_raw + ((1024u - (smem_u32(pd_raw) & 1023u)) & 1023u);
  const int k = p.k, a = p.a, H = p.H, NB = p.NB, B = p.B;
  const int M1 = 4 * H, G5 = 5 * H;
  const uint32_t stage_bytes = PD_A_BYTES + (uint32_t)NB * 128u;
  PdSmem sm;
  sm.ring = base;
  sm.V = reinterpret_cast<float*>(base + (size_t)PD_STAGES * stage_bytes);
  sm.P = sm.V + (size_t)k * H;
  sm.c = sm.P + (size_t)k * p.ldP;
  sm.u = sm.c + H;
  sm.red = sm.u + H;
  sm.full = reinterpret_cast<uint64_t*>(sm.red + 64);
  sm.empty = sm.full + PD_STAGES;
  sm.tfull = sm.empty + PD_STAGES;
  sm.tempty = sm.tfull + 2;
  sm.tmem_slot = reinterpret_cast<uint32_t*>(sm.tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool owner = (int)blockIdx.x < B;
  const int b = blockIdx.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PD_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.tfull[i], 1);
      mbar_init(&sm.tempty[i], 4);      // one arrive per drain warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {      // two accumulators of up to 256 columns: the whole tensor memory of this SM (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident operands of the owned image: V_b, P_b, c_0
  if (owner) {
    const float* Vb = p.V + (size_t)b * k * H;
    for (int i = tid * 4; i < k * H; i += PD_THREADS * 4) *reinterpret_cast<float4*>(sm.V + i) = ldg4_stream(Vb + i);
    const float* Pb = p.P + (size_t)b * k * p.ldP;
    for (int i = tid * 4; i < k * p.ldP; i += PD_THREADS * 4) *reinterpret_cast<float4*>(sm.P + i) = ldg4(Pb + i);
    for (int i = tid; i < H; i += PD_THREADS) sm.c[i] = p.c0 ? p.c0[(size_t)b * H + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot;

  // owner-phase scratch aliases the stage ring (the phases never overlap: grid barriers in between)
  float* g_s = reinterpret_cast<float*>(sm.ring);          // [5H] gate pre-activations
  float* h_s = g_s + G5;                                   // [H]
  float* s_s = h_s + H;                                    // [H]
  float* q_s = s_s + H;                                    // [a]
  float* r_s = q_s + ((a + 3) & ~3);                       // [a]
  float* z_s = r_s + ((a + 3) & ~3);                       // [k + 1] scores, then alphas
  float* cx_s = z_s + ((k + 4) & ~3);                      // [3][H] partial contexts
  int* cand = reinterpret_cast<int*>(cx_s + 3 * H);        // O2: candidate columns (O2 and O1 share one phase: no aliasing between them)
  int* ncand = cand + PD_MAX_CAND;
  float* wv = reinterpret_cast<float*>(ncand + 4);         // O2: per-warp (value, index) winners
  int* wi = reinterpret_cast<int*>(wv + PD_WARPS);

  unsigned epoch = 0;
  int it = 0, lt = 0;
  const bool sentinel = p.Ws != nullptr;
  PdPhase ph;
  ph.units2 = (p.Vc + 127) / 128; ph.tiles1 = (M1 + 127) / 128; ph.ks1 = p.ks1; ph.kbper1 = p.kbper1; ph.nkb1 = p.nkb1; ph.nkb2 = p.nkb2;

  // ---- O1: gates -> cell -> sentinel -> attention over the resident V -> u.  `word` = the token fed at this step ----
  auto owner_step = [&](int t, int word) {
    // (a) gates = EG[word] (input half: W_ih[:, :E] emb(word), W_x[:, :E] emb(word)) + static (v_g half, biases) + recurrent K-split partials
    const float* eg = p.EG + (size_t)word * G5;
    for (int i = tid * 4; i < G5; i += PD_THREADS * 4) {
      float4 acc = ldg4(eg + i);
      const float4 st4 = ldg4(p.stat + (size_t)b * G5 + i);
      acc.x += st4.x; acc.y += st4.y; acc.z += st4.z; acc.w += st4.w;
      if (i < M1) {
        for (int ks = 0; ks < p.ks1; ++ks) {
          const float4 x = ldcg4(p.part1 + ((size_t)ks * B + b) * M1 + i);
          acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
      }
      *reinterpret_cast<float4*>(g_s + i) = acc;
    }
    __syncthreads();
    // (b) LSTM cell + sentinel gate (baseline_attention.py:172; adaptive_attention.py:79-83 with h~ = 0)
    float* hrow = p.hA + (size_t)b * p.ldA;
    for (int i = tid; i < H; i += PD_THREADS) {
      const float cc = sigmoidf_acc(g_s[H + i]) * sm.c[i] + sigmoidf_acc(g_s[i]) * tanhf(g_s[2 * H + i]);
      const float tcc = tanhf(cc);
      const float hn = sigmoidf_acc(g_s[3 * H + i]) * tcc;
      const float sn = sentinel ? sigmoidf_acc(g_s[4 * H + i]) * tcc : 0.f;
      sm.c[i] = cc;
      h_s[i] = hn;
      s_s[i] = sn;
      float hi, lo;
      split_tf32(hn, hi, lo);
      hrow[i] = hi;                     // operand of the next step's G1
      hrow[p.lo1 + i] = lo;
    }
    __syncthreads();
    // (c) q = W_g h, r = W_s s + q: two weight rows per warp in flight                        adaptive_attention.py:35,45
    for (int row = warp * 2; row < 2 * a; row += PD_WARPS * 2) {
      float acc[2] = {0.f, 0.f};
      float4 w4[2][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r2 = row + rr;
        const bool is_s = r2 >= a;
        const int j = is_s ? r2 - a : r2;
        const bool live = r2 < 2 * a && (!is_s || sentinel);
        const float* wrow = (is_s ? p.Ws : p.Wg) + (size_t)j * H;
#pragma unroll
        for (int c = 0; c < 4; ++c) w4[rr][c] = (live && lane * 4 + c * 128 < H) ? ldg4(wrow + lane * 4 + c * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r2 = row + rr;
        const bool is_s = r2 >= a;
        const int j = is_s ? r2 - a : r2;
        const bool live = r2 < 2 * a && (!is_s || sentinel);
        const float* act = is_s ? s_s : h_s;
        if (live) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int col = lane * 4 + c * 128;
            if (col < H) {
              const float4 x4 = *reinterpret_cast<const float4*>(act + col);
              acc[rr] = fmaf(w4[rr][c].x, x4.x, acc[rr]); acc[rr] = fmaf(w4[rr][c].y, x4.y, acc[rr]);
              acc[rr] = fmaf(w4[rr][c].z, x4.z, acc[rr]); acc[rr] = fmaf(w4[rr][c].w, x4.w, acc[rr]);
            }
          }
          for (int col = lane * 4 + 512; col < H; col += 128) {      // (H > 512: the rest of the row)
            const float4 w = ldg4(((is_s ? p.Ws : p.Wg) + (size_t)j * H) + col);
            const float4 x4 = *reinterpret_cast<const float4*>(act + col);
            acc[rr] = fmaf(w.x, x4.x, acc[rr]); acc[rr] = fmaf(w.y, x4.y, acc[rr]); acc[rr] = fmaf(w.z, x4.z, acc[rr]); acc[rr] = fmaf(w.w, x4.w, acc[rr]);
          }
        }
        const float v = warp_sum(acc[rr]);
        if (lane == 0 && r2 < 2 * a) (is_s ? r_s : q_s)[j] = v;
      }
    }
    __syncthreads();
    for (int j = tid; j < a; j += PD_THREADS) r_s[j] += q_s[j];
    __syncthreads();
    // (d) scores z_i = w_h . tanh(P_i + q), z_s = w_h . tanh(r)                            :36-37, :46-47
    for (int item = warp; item < k + 1; item += PD_WARPS) {
      float acc = 0.f;
      if (item < k) {
        const float* prow = sm.P + (size_t)item * p.ldP;
        for (int j = lane; j < a; j += 32) acc = fmaf(__ldg(p.wh + j), tanhf(prow[j] + q_s[j]), acc);
      } else {
        for (int j = lane; j < a; j += 32) acc = fmaf(__ldg(p.wh + j), tanhf(r_s[j]), acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) z_s[item] = acc;
    }
    __syncthreads();
    // (e) softmax over the k regions, and the sentinel's share of the (k+1)-way softmax                :39, :51
    if (warp == 0) {
      float m = -INFINITY;
      for (int i = lane; i < k; i += 32) m = fmaxf(m, z_s[i]);
      m = warp_max(m);
      float sum = 0.f;
      for (int i = lane; i < k; i += 32) sum += expf(z_s[i] - m);
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      const float zsent = z_s[k];
      const float m1 = fmaxf(m, zsent);
      float sum1 = 0.f;
      for (int i = lane; i < k; i += 32) sum1 += expf(z_s[i] - m1);
      sum1 = warp_sum(sum1);
      const float es = expf(zsent - m1);
      const float beta = sentinel ? es / (sum1 + es) : 0.f;
      __syncwarp();
      float* aout = p.alpha + ((size_t)b * p.L + t) * k;
      for (int i = lane; i < k; i += 32) {
        const float al = expf(z_s[i] - m) * inv;
        z_s[i] = al;
        aout[i] = al;
      }
      if (lane == 0) {
        z_s[k] = beta;
        p.beta[(size_t)b * p.L + t] = beta;
      }
    }
    __syncthreads();
    // (f) context over the resident V: three row groups in parallel, summed in a fixed order
    {
      const int H4 = H / 4;
      for (int item = tid; item < 3 * H4; item += PD_THREADS) {
        const int grp = item / H4, c4 = (item % H4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = grp; i < k; i += 3) {
          const float w = z_s[i];
          const float4 v = *reinterpret_cast<const float4*>(sm.V + (size_t)i * H + c4);
          acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
        }
        *reinterpret_cast<float4*>(cx_s + (size_t)grp * H + c4) = acc;
      }
    }
    __syncthreads();
    // (g) c_hat = beta s + (1 - beta) ctx, u = c_hat + h                                          :54, :132
    {
      const float beta = z_s[k];
      float ss = 0.f, sh = 0.f, sd = 0.f;
      for (int i = tid; i < H; i += PD_THREADS) {
        const float ctx = (cx_s[i] + cx_s[H + i]) + cx_s[2 * H + i];
        const float uu = beta * s_s[i] + (1.f - beta) * ctx + h_s[i];
        const __nv_bfloat16 u16 = __float2bfloat16(uu);
        const float uh = __bfloat162float(u16);
        sm.u[i] = uu;
        p.u16[(size_t)b * H + i] = u16;
        ss = fmaf(uu, uu, ss);
        sh = fmaf(uh, uh, sh);
        sd = fmaf(uu - uh, uu - uh, sd);
      }
      ss = block_sum(ss, sm.red, tid);
      sh = block_sum(sh, sm.red, tid);
      sd = block_sum(sd, sm.red, tid);
      if (tid == 0) {      // coefficients of the first pass's error bound  cw * ||W_j|| + cd * ||W_j - W^_j||  (vocab_refine.cu)
        sm.red[32] = sqrtf(sd) * (1.f + 1e-5f) + 1.220703125e-4f * sqrtf(ss);      // ||u - u^|| + 2^-13 ||u||
        sm.red[33] = sqrtf(sh) * (1.f + 1e-5f);                                      // ||u^||
      }
    }
    __syncthreads();
  };

  // ---- O2: exact arg-max of the row by filter-and-refine; returns the chosen word (uniform over the CTA) ----
  auto owner_argmax = [&](int t) -> int {
    const float cb = sm.red[32], cdb = sm.red[33];
    const float* arow = p.approx + (size_t)b * p.ldv;
    const int Vc = p.Vc;
    constexpr int NC = 8;                         // float4 groups a thread keeps in registers between the two passes
    float4 xv[NC], bv[NC];
    // pass 1: L = max_j (approx_j - bound_j); the first NC groups of every thread stay in registers for pass 2
    float lo = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = (c * PD_THREADS + tid) * 4;
      xv[c] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      bv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j + 3 < Vc) {
        xv[c] = ldcg4(arow + j);
        const float4 w = ldg4(p.wn + j), dw = ldg4(p.dwn + j);
        bv[c] = make_float4(cb * w.x + cdb * dw.x, cb * w.y + cdb * dw.y, cb * w.z + cdb * dw.z, cb * w.w + cdb * dw.w);
      } else if (j < Vc) {
        float xs[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, ws[4] = {0.f, 0.f, 0.f, 0.f};
        for (int e = 0; e < 4 && j + e < Vc; ++e) { xs[e] = __ldcg(arow + j + e); ws[e] = cb * __ldg(p.wn + j + e) + cdb * __ldg(p.dwn + j + e); }
        xv[c] = make_float4(xs[0], xs[1], xs[2], xs[3]);
        bv[c] = make_float4(ws[0], ws[1], ws[2], ws[3]);
      }
      lo = fmaxf(fmaxf(lo, xv[c].x - bv[c].x), fmaxf(xv[c].y - bv[c].y, fmaxf(xv[c].z - bv[c].z, xv[c].w - bv[c].w)));
    }
    for (int j = (NC * PD_THREADS + tid) * 4; j < Vc; j += PD_THREADS * 4)       // (vocabularies beyond NC * 1536 columns)
      for (int e = 0; e < 4 && j + e < Vc; ++e) lo = fmaxf(lo, __ldcg(arow + j + e) - (cb * __ldg(p.wn + j + e) + cdb * __ldg(p.dwn + j + e)));
    if (tid == 0) *ncand = 0;
    const float Lb = block_max(lo, sm.red, tid);
    // pass 2: columns whose upper bound reaches L
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = (c * PD_THREADS + tid) * 4;
      const float xs[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w}, bs[4] = {bv[c].x, bv[c].y, bv[c].z, bv[c].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (xs[e] + bs[e] >= Lb) {                // (-inf for columns beyond Vc: never listed)
          const int slot = atomicAdd(ncand, 1);
          if (slot < PD_MAX_CAND) cand[slot] = j + e;
        }
      }
    }
    for (int j = (NC * PD_THREADS + tid) * 4; j < Vc; j += PD_THREADS * 4)
      for (int e = 0; e < 4 && j + e < Vc; ++e)
        if (__ldcg(arow + j + e) + (cb * __ldg(p.wn + j + e) + cdb * __ldg(p.dwn + j + e)) >= Lb) {
          const int slot = atomicAdd(ncand, 1);
          if (slot < PD_MAX_CAND) cand[slot] = j + e;
        }
    __syncthreads();
    const int nc = *ncand;
    const bool all = nc > PD_MAX_CAND;       // (pathological row: recompute every column exactly)
    const int n = all ? Vc : nc;
    // pass 3: exact fp32 logits of the candidates, lowest index wins ties                         :132, :201
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int ci = warp; ci < n; ci += PD_WARPS) {
      const int j = all ? ci : cand[ci];
      const float* wrow = p.Wp + (size_t)j * H;
      float acc = 0.f;
      for (int c = lane * 4; c < H; c += 128) {
        const float4 w4 = ldg4(wrow + c);
        const float4 x4 = *reinterpret_cast<const float4*>(sm.u + c);
        acc = fmaf(w4.x, x4.x, acc); acc = fmaf(w4.y, x4.y, acc); acc = fmaf(w4.z, x4.z, acc); acc = fmaf(w4.w, x4.w, acc);
      }
      acc = warp_sum(acc) + __ldg(p.bp + j);
      if (acc > best || (acc == best && j < best_i)) { best = acc; best_i = j; }
    }
    if (lane == 0) { wv[warp] = best; wi[warp] = best_i; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < PD_WARPS; ++w)
        if (wv[w] > best || (wv[w] == best && wi[w] < best_i)) { best = wv[w]; best_i = wi[w]; }
      wi[0] = best_i;
      p.ids[(size_t)b * p.L + t] = best_i;                                                   // :201-202
      if (p.ncand_out) p.ncand_out[(size_t)b * p.L + t] = nc;
    }
    __syncthreads();
    const int word = wi[0];
    __syncthreads();
    return word;
  };

  // step 0: recurrent terms of h_0, then the first owner step with <start>                                   :188
  ph.g1 = true; ph.g2 = false;
  pd_gemm_phase(sm, p, ph, &tmA1, &tmB1, &tmA2, &tmB2, tmem_base, it, lt);
  pd_grid_sync(p.bar, epoch);
  if (owner) owner_step(0, p.start_id);
  pd_grid_sync(p.bar, epoch);
  for (int t = 0; t < p.L; ++t) {
    // [ G2(t): approximate logits of u_t | G1(t+1): recurrent gate terms of h_t ]
    ph.g2 = true; ph.g1 = t + 1 < p.L;
    pd_trace(t, 0);
    pd_gemm_phase(sm, p, ph, &tmA1, &tmB1, &tmA2, &tmB2, tmem_base, it, lt);
    __syncthreads();
    pd_trace(t, 1);
    pd_grid_sync(p.bar, epoch);
    pd_trace(t, 2);
    // [ O2(t): the word | O1(t+1): the next step up to u ]
    if (owner) {
      const int word = owner_argmax(t);
      pd_trace(t, 3);
      if (t + 1 < p.L) owner_step(t + 1, word);
      pd_trace(t, 4);
    }
    if (t + 1 < p.L) pd_grid_sync(p.bar, epoch);
    pd_trace(t, 5);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// norms of the projection's rows and of their bf16 rounding residuals, rounded up (once per set of weights):
// wn[j] = ||W_p[j, :]||_2, dwn[j] = ||W_p[j, :] - bf16(W_p[j, :])||_2 (the conversion of launch_cast2d)
__global__ void __launch_bounds__(256) row_norm_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ wn,
                                                       float* __restrict__ dwn) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= rows) return;
  float ss = 0.f, sd = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float x = __ldg(W + (size_t)j * cols + c);
    const float dx = x - __bfloat162float(__float2bfloat16(x));
    ss = fmaf(x, x, ss);
    sd = fmaf(dx, dx, sd);
  }
  ss = warp_sum(ss);
  sd = warp_sum(sd);
  if (lane == 0) {
    wn[j] = sqrtf(ss) * (1.f + 1e-5f);
    dwn[j] = sqrtf(sd) * (1.f + 1e-5f);
  }
}

size_t pd_smem_bytes(int NB, int k, int H, int ldP) {
  return (size_t)PD_STAGES * (PD_A_BYTES + (size_t)NB * 128) + (size_t)k * H * 4 + (size_t)k * ldP * 4 + (size_t)2 * H * 4 + 64 * 4 +
         (2 * PD_STAGES + 4) * 8 + 16 + 1024;
}

}  // namespace

int decode_persist_nb(int B) { return B <= 16 ? 16 : (B + 15) / 16 * 16; }

// The shape fits if the batch has at most one image per SM, the resident operands + the stage ring fit in shared memory and the
// owner-phase scratch fits inside the ring.
bool decode_persist_supported(int B, int k, int a, int H, int E, int Vc) {
  if (B < 1 || B > num_sms() || B > 256) return false;
  if (H % 8 != 0 || E % 4 != 0 || H < 8 || a > 128 || k < 1) return false;
  const int NB = decode_persist_nb(B), ldP = (a + 3) / 4 * 4;
  if (pd_smem_bytes(NB, k, H, ldP) > 227 * 1024) return false;
  const size_t scratch = (size_t)(5 * H + 2 * H + 2 * ((a + 3) & ~3) + ((k + 4) & ~3) + 3 * H + PD_MAX_CAND + 4 + 2 * PD_WARPS) * 4;
  const size_t ring = (size_t)PD_STAGES * (PD_A_BYTES + (size_t)NB * 128);
  (void)Vc;
  return scratch <= ring;
}

int set_persist_trace_buffer(void* dev_ptr) {
  unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
  AA_CHECK_CUDA(cudaMemcpyToSymbol(g_pd_trace, &p, sizeof(p)));
  return AA_OK;
}

int launch_row_norm(const float* W, int rows, int cols, float* wn, float* dwn, cudaStream_t st) {
  row_norm_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(W, rows, cols, wn, dwn);
  AA_CHECK_LAUNCH("row_norm");
  return AA_OK;
}

int launch_decode_persist(const DecodePersistArgs& p0, const float* Whh_split, const __nv_bfloat16* Wp16, cudaStream_t st) {
  DecodePersistArgs p = p0;
  AA_REQUIRE(decode_persist_supported(p.B, p.k, p.a, p.H, p.E, p.Vc), "persistent decode: shape / batch does not fit (B=%d k=%d H=%d)", p.B, p.k, p.H);
  p.NB = decode_persist_nb(p.B);
  const int M1 = 4 * p.H;
  const int sms = num_sms();
  // G1: k-blocks of 16 floats over the padded K (p.K1p, a multiple of 32); K ranges so that its (row tile, range) units fill the
  // CTAs G2's tiles leave free
  p.nkb1 = p.K1p / 16;
  p.nkb2 = ceil_div(p.H, 64);
  const int tiles1 = ceil_div(M1, 128), units2 = ceil_div(p.Vc, 128);
  const int free_ctas = sms - (units2 % sms);
  int ks = (units2 < sms ? free_ctas : sms) / tiles1;
  if (ks < 1) ks = 1;
  if (ks > p.nkb1) ks = p.nkb1;
  if (ks > p.ks1_max) ks = p.ks1_max;
  p.kbper1 = ceil_div(p.nkb1, ks);
  p.ks1 = ceil_div(p.nkb1, p.kbper1);
  CUtensorMap tmA1, tmB1, tmA2, tmB2;
  // W_hh rows [hi (K1p) | lo (K1p)]; h rows [hi ... | lo ...] with the lo half lo1 columns in; rows beyond B / 4H are zero-filled by TMA
  AA_TRY(make_map(&tmA1, Whh_split, 4, M1, 2LL * p.K1p, 2LL * p.K1p, 128, 64));
  AA_TRY(make_map(&tmB1, p.hA, 4, p.B, p.ldA, p.ldA, p.NB, 64));
  AA_TRY(make_map(&tmA2, Wp16, 2, p.Vc, p.H, p.H, 128, 128));
  AA_TRY(make_map(&tmB2, p.u16, 2, p.B, p.H, p.H, p.NB, 128));
  const size_t smem = pd_smem_bytes(p.NB, p.k, p.H, p.ldP);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(dec_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  AA_CHECK_CUDA(cudaMemsetAsync(p.bar, 0, sizeof(unsigned), st));
  void* args[] = {(void*)&tmA1, (void*)&tmB1, (void*)&tmA2, (void*)&tmB2, (void*)&p};
  AA_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)dec_persist_kernel, dim3(sms), dim3(PD_THREADS), args, smem, st));
  count_launch();
  return AA_OK;
}

}  // namespace aa
