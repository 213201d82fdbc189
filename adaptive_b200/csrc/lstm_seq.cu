// Persistent LSTM recurrence kernels (bf16 tensor-core mode): all T steps of the forward
// recurrence (baseline_attention.py:167-178) and of its BPTT run inside ONE cooperative launch
// each, instead of 2 launches per step.
//
//  * every CTA keeps its slice of W_hh resident in shared memory (TMA-loaded once) for all steps;
//  * per step it TMA-streams the previous state (h_{t-1}, resp. dgates_{t+1}) from L2 through an
//    mbarrier ring, runs the slice GEMM on tcgen05 (fp32 accumulator in TMEM), and applies the LSTM
//    cell (resp. its gradient) in the epilogue warps straight out of tensor memory, with the cell
//    state c (resp. dc) living in registers across all steps;
//  * steps are separated by a release/acquire grid barrier per 128-row group (rows are independent
//    sequences), so the launch must be cooperative (all CTAs co-resident).
#include <stdlib.h>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace aa {

namespace {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int SEQ_THREADS = 192;      // forward: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int BWD_THREADS = 576;      // backward: warp 0 TMA, warp 1 MMA, warps 2-17 epilogue (4 units each)
constexpr int MAX_STAGES = 8;
constexpr size_t SEQ_SMEM_BUDGET = 200 * 1024;
constexpr uint32_t A_STAGE_BYTES = 128 * 128;   // 128 rows x 64 bf16

// Optional per-step timeline of CTA (0,0) (diagnostics: aa_debug_set_trace_buffer): 8 x uint64 globaltimer
// stamps per step -- 0 barrier passed, 1 TMA issued, 2 first stage landed, 3 MMA committed, 4 accumulator
// observed, 5 cell math + stores issued, 6 epilogue warps met, 7 arrive published.
__device__ unsigned long long* g_seq_trace = nullptr;
__device__ __forceinline__ void trace(int step, int ev) {
  if (g_seq_trace && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_seq_trace[step * 8 + ev] = t;
  }
}

__device__ __forceinline__ void grid_barrier_wait(const unsigned* counter, unsigned target) {
  long long t0 = 0;
  while (ld_acquire_gpu(counter) < target) {
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > AA_SPIN_LIMIT_CYCLES) __trap();
  }
}

template <int NT>
__device__ __forceinline__ void epilogue_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// packed 4 x bf16 store
__device__ __forceinline__ void st_bf16x4(bf16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

struct SeqFwdArgs {
  int B, T, H;
  const float* xg;   // [B,T,4H] input-half gate pre-activations incl. both biases
  const float* c0;   // [B,H] or null
  float *hiddens, *cells, *acts, *hs_prev;   // [B,T,H] [B,T,H] [B,T,4,H] [B,T,H]
  bf16 *hid16, *hsprev16;                    // bf16 mirrors
  unsigned* counters;                        // [row groups], zeroed by the launcher
  int stages;                                // depth of the h_{t-1} TMA ring (<= MAX_STAGES)
  int box_rows;                              // rows per TMA box (batch rows of a group rounded up to 8)
};

// U hidden units per CTA -> UMMA N = 4U gate columns, packed unit-major: column n = u*4 + g.
template <int U>
__global__ void __launch_bounds__(SEQ_THREADS, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH0,
                    const __grid_constant__ CUtensorMap tmH, const SeqFwdArgs a) {
  constexpr int N = 4 * U;
  // Back-to-back tcgen05.mma into the SAME accumulator serialise on the accumulate dependency (~180 cycles each,
  // measured: 32 dependent N=16 MMAs took 3 us); with the k-steps spread round-robin over NACC accumulators (summed in
  // the epilogue) the pipe is bound by the A-operand read instead.
  constexpr int NACC = N <= 32 ? 4 : 2;
  constexpr int TCOLS = N * NACC < 32 ? 32 : N * NACC;
  constexpr int CH = N < 32 ? N : 32;           // TMEM columns per epilogue chunk
  constexpr int UC = CH / 4;                    // units per chunk
  constexpr uint32_t W_KB_BYTES = N * 128;      // one 64-wide k-block of the weight slice
  const int KB = a.H / 64;
  const int C = gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS)
  const int STAGES = a.stages;
  uint8_t* sA = smem;                                        // [STAGES][16 KB]
  uint8_t* sW = smem + (size_t)STAGES * A_STAGE_BYTES;       // [KB][N*128]  (N*128 is a multiple of 1024 for N >= 8)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)KB * W_KB_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* w_full = bars + 2 * MAX_STAGES;
  uint64_t* tmem_full = w_full + 1;
  uint64_t* tmem_empty = w_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, rg = blockIdx.y;
  const int m0 = rg * 128;
  const unsigned* counter = a.counters + rg;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(w_full, 1);
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // resident weight slice: rows [c*N, (c+1)*N) of the packed W_hh, all k-blocks
      mbar_expect_tx(w_full, (uint32_t)KB * W_KB_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + (size_t)kb * W_KB_BYTES, &tmW, kb * 64, c * N, w_full);
      int it = 0;
      for (int t = 0; t < a.T; ++t) {
        if (t > 0) {
          grid_barrier_wait(counter, (unsigned)t * C);   // every CTA of this row group has published h_{t-1}
          fence_proxy_async_global();
        }
        trace(t, 0);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], (uint32_t)a.box_rows * 128u);
          if (t == 0) tma_load_2d(sA + s * A_STAGE_BYTES, &tmH0, kb * 64, m0, &full_bar[s]);
          else        tma_load_2d(sA + s * A_STAGE_BYTES, &tmH, (t - 1) * a.H + kb * 64, m0, &full_bar[s]);
        }
        trace(t, 1);
      }
    }
  } else if (warp == 1) {
    {   // the MMA warp stays converged; one elected lane issues (see tc::elect_one)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(w_full, 0);
      const uint64_t desc0 = make_smem_desc(0, 16, 1024);
      int it = 0;
      for (int t = 0; t < a.T; ++t) {
        mbar_wait(tmem_empty, (t & 1) ^ 1);      // epilogue has drained the previous step's accumulator
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          if (kb == 0 && lane == 0) trace(t, 2);
          tc_fence_after();
          // descriptors differ only in the start-address field (bits 0-13, 16-byte units): +2 per 32-byte k-step
          const uint64_t da = desc0 + ((smem_u32(sA + s * A_STAGE_BYTES)) >> 4);
          const uint64_t db = desc0 + ((smem_u32(sW + (size_t)kb * W_KB_BYTES)) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)    // k-step j = kb*4 + k -> accumulator j % NACC (NACC divides 4 or is 4)
              tc_mma<false>(tmem_base + (uint32_t)((k % NACC) * N), da + 2 * k, db + 2 * k, idesc, (kb * 4 + k) >= NACC ? 1u : 0u);
            tc_commit(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(tmem_full);
        __syncwarp();
        if (lane == 0) trace(t, 3);
      }
    }
  } else {
    // ===== epilogue warps: LSTM cell out of tensor memory; thread = batch row, cell state in registers =====
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool valid = row < a.B;
    const int H = a.H, T = a.T;
    const int j0 = c * U;                        // first hidden unit of this CTA
    float creg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) creg[u] = (valid && a.c0) ? a.c0[(long long)row * H + j0 + u] : 0.f;
    for (int t = 0; t < T; ++t) {
      const long long bt = (long long)row * T + t;
      float hn[U], ig[U], fg[U], gg[U], og[U];
      // input-half pre-activations of the first chunk: independent of the MMA, fetched while it runs
      float4 x4[4][UC / 4];
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int v = 0; v < UC / 4; ++v)
            x4[g][v] = *reinterpret_cast<const float4*>(a.xg + bt * 4 * H + (long long)g * H + j0 + v * 4);
      }
      mbar_wait(tmem_full, t & 1);
      if (threadIdx.x == 64) trace(t, 4);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < N / CH; ++ch) {
        if (ch > 0 && valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int v = 0; v < UC / 4; ++v)
              x4[g][v] = *reinterpret_cast<const float4*>(a.xg + bt * 4 * H + (long long)g * H + j0 + ch * UC + v * 4);
        }
        uint32_t r[CH];
        tmem_ld<CH>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * CH), r);
#pragma unroll
        for (int ai = 1; ai < NACC; ++ai) {      // sum the partial accumulators
          uint32_t r2[CH];
          tmem_ld<CH>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ai * N + ch * CH), r2);
#pragma unroll
          for (int e = 0; e < CH; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
        }
        if (ch == N / CH - 1) {                  // accumulator fully read: hand TMEM back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);
        }
        if (valid) {
#pragma unroll
          for (int v = 0; v < UC / 4; ++v) {
            const float xi[4] = {x4[0][v].x, x4[0][v].y, x4[0][v].z, x4[0][v].w};
            const float xf[4] = {x4[1][v].x, x4[1][v].y, x4[1][v].z, x4[1][v].w};
            const float xc[4] = {x4[2][v].x, x4[2][v].y, x4[2][v].z, x4[2][v].w};
            const float xo[4] = {x4[3][v].x, x4[3][v].y, x4[3][v].z, x4[3][v].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ul = v * 4 + e;          // unit within chunk; TMEM column = ul*4 + gate
              const int u = ch * UC + ul;
              ig[u] = sigmoidf_fast(__uint_as_float(r[ul * 4 + 0]) + xi[e]);
              fg[u] = sigmoidf_fast(__uint_as_float(r[ul * 4 + 1]) + xf[e]);
              gg[u] = tanhf_fast(__uint_as_float(r[ul * 4 + 2]) + xc[e]);
              og[u] = sigmoidf_fast(__uint_as_float(r[ul * 4 + 3]) + xo[e]);
              creg[u] = fg[u] * creg[u] + ig[u] * gg[u];
              hn[u] = og[u] * tanhf_fast(creg[u]);
            }
            // only the bf16 h_t feeds the next step: store it first, everything else after the arrive
            st_bf16x4(a.hid16 + bt * H + j0 + ch * UC + v * 4, hn[ch * UC + v * 4], hn[ch * UC + v * 4 + 1], hn[ch * UC + v * 4 + 2],
                      hn[ch * UC + v * 4 + 3]);
          }
        }
      }
      // publish h_t (cooperative-groups grid-sync pattern): the epilogue warps meet, one thread fences at gpu
      // scope (cumulative over the stores it observed through the barrier) and release-arrives on the row
      // group's counter; the consumer side acquires and issues fence.proxy.async before its TMA loads
      if (threadIdx.x == 64) trace(t, 5);
      epilogue_bar<128>();
      if (threadIdx.x == 64) trace(t, 6);
      // (red.release.gpu is cumulative over the stores observed through the barrier: no separate __threadfence)
      if (warp == 2 && lane == 0 && t + 1 < T) red_release_gpu_add(a.counters + rg, 1u);
      if (valid) {   // the remaining outputs are only read after the kernel: off the critical path
#pragma unroll
        for (int u = 0; u < U; u += 4) {
          const int j = j0 + u;
          *reinterpret_cast<float4*>(a.hiddens + bt * H + j) = make_float4(hn[u], hn[u + 1], hn[u + 2], hn[u + 3]);
          *reinterpret_cast<float4*>(a.cells + bt * H + j) = make_float4(creg[u], creg[u + 1], creg[u + 2], creg[u + 3]);
          float* ac = a.acts + bt * 4 * H + j;
          *reinterpret_cast<float4*>(ac) = make_float4(ig[u], ig[u + 1], ig[u + 2], ig[u + 3]);
          *reinterpret_cast<float4*>(ac + H) = make_float4(fg[u], fg[u + 1], fg[u + 2], fg[u + 3]);
          *reinterpret_cast<float4*>(ac + 2 * H) = make_float4(gg[u], gg[u + 1], gg[u + 2], gg[u + 3]);
          *reinterpret_cast<float4*>(ac + 3 * H) = make_float4(og[u], og[u + 1], og[u + 2], og[u + 3]);
          if (t + 1 < T) {
            *reinterpret_cast<float4*>(a.hs_prev + (bt + 1) * H + j) = make_float4(hn[u], hn[u + 1], hn[u + 2], hn[u + 3]);
            st_bf16x4(a.hsprev16 + (bt + 1) * H + j, hn[u], hn[u + 1], hn[u + 2], hn[u + 3]);
          }
        }
      }
      if (threadIdx.x == 64) trace(t, 7);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TCOLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// BPTT
// ---------------------------------------------------------------------------------------------
struct SeqBwdArgs {
  int B, T, H;
  const float *dh_attn, *dhs, *dcell;     // [B,T,H]: batched (non-recurrent) gradient contributions; dhs may be null (T == 1)
  const float *d_hT, *d_cT;               // [B,H] or null
  const float *acts, *cells, *c0;         // saved forward state; c0 may be null (zeros)
  float* dgates; bf16* dgates16;          // [B,T,4H]
  float *dh0, *dc0;                       // [B,H] (may be null)
  unsigned* counters;
  int stages;
  int box_rows;
};

// 16 hidden units per CTA: dh_rec[:, j-slice] = dgates_{t+1} [B,4H] * W_hh[:, j-slice]  (K = 4H).
// 16 epilogue warps: warp -> (TMEM lane quarter = warp % 4, column group = (warp-2)/4), i.e. each thread owns
// one batch row and 4 hidden units, so that all of a step's operand loads are in flight at once.
// The same 512 threads also FEED the contraction: per step every CTA needs the whole dgates_t block
// ([rows, 4H] bf16, 320 KB at B=80 / H=512).  A one-thread TMA producer tops out at ~48 B/clk per SM on this
// access pattern (tools/tma_probe.cu: 3.6 us per step, and 12 us inside the kernel), cooperative 16-byte
// cp.async.cg copies into the 128B-swizzled K-major layout run as deep as the ring (slots - 1 super-stages of 2
// k-blocks in flight), completion counted by the slot's mbarrier (cp.async.mbarrier.arrive.noinc).
constexpr int SS_KB = 2;                                   // k-blocks per super-stage
constexpr uint32_t SS_BYTES = SS_KB * A_STAGE_BYTES;       // 32 KB
constexpr int BWD_LOADERS = BWD_THREADS - 64;              // 512
constexpr int BWD_LPT = (128 * 8 * SS_KB) / BWD_LOADERS;   // 16-byte chunks per thread and super-stage at 128 rows (4)

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (does not add to the pending count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- cluster helpers (K-split of the backward contraction) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// release-arrive (cluster scope) on an mbarrier that lives in a peer CTA's shared memory
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the data guarded by the barrier was written by peer CTAs
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > AA_SPIN_LIMIT_CYCLES) __trap();
    }
  } while (!done);
}

// KS = K-split: a cluster of KS CTAs shares one 16-unit output slice, CTA rank kq contracts the K range
// [kq*4H/KS, (kq+1)*4H/KS) (its own quarter of dgates_t and of the W_hh^T slice), the fp32 partials are exchanged through
// distributed shared memory (one 16-byte store per thread to the CTA that owns the thread's 4 units, completion counted
// by an mbarrier in the owner) and rank r finishes the units of column groups cg % KS == r.  A tcgen05.mma with M = 128
// operands from shared memory costs ~103 cycles whatever N is (tools/mma_probe.cu), so the 4H/16 k-steps of this
// contraction bound the step; the split divides them (and the operand bytes each CTA pulls from L2) by KS.
template <int KS>
__global__ void __launch_bounds__(BWD_THREADS, 1)
lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap tmWT, const SeqBwdArgs a) {
  constexpr int U = 16, N = 16, NACC = 8, TCOLS = N * NACC;   // NACC partial accumulators: see lstm_seq_fwd_kernel
  constexpr uint32_t W_KB_BYTES = N * 128;
  const int KB = 4 * a.H / 64 / KS;           // k-blocks of this CTA's K range
  const int NSS = KB / SS_KB;                 // super-stages per step
  const int CT = gridDim.x;                   // CTAs per row group (all of them publish every step)
  const int NSLOT = a.stages;                 // ring slots of SS_BYTES

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS)
  uint8_t* sA = smem;
  uint8_t* sW = smem + (size_t)NSLOT * SS_BYTES;
  float* recv = reinterpret_cast<float*>(sW + (size_t)KB * W_KB_BYTES);          // [KS src][4 cg][128 rows][4] fp32 partials from the peers
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(recv) + (KS > 1 ? (size_t)KS * 4 * 128 * 16 : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* w_full = bars + 2 * MAX_STAGES;
  uint64_t* tmem_full = w_full + 1;
  uint64_t* tmem_empty = w_full + 2;
  uint64_t* recv_bar = w_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kq = KS > 1 ? (int)cluster_ctarank() : 0;
  const int c = blockIdx.x / KS, rg = blockIdx.y;
  const int m0 = rg * 128;
  const int T = a.T, H = a.H;
  const int kcol0 = kq * (4 * H / KS);        // first gate column of this CTA's K range

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&full_bar[s], BWD_LOADERS); mbar_init(&empty_bar[s], 1); }
    mbar_init(w_full, 1);
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 16);
    mbar_init(recv_bar, KS > 1 ? (KS - 1) * (4 / KS) * 4 : 1);   // one arrive per sending warp: (KS-1) peers x owned column groups x 4 lane quarters
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // GEMM i (i = 1..T) consumes dgates of step t = T - i and produces dh_rec for step t - 1 (dh0 when t == 0).
  if (warp == 0) {
    if (lane == 0) {   // resident weight slice: rows [c*U, (c+1)*U) of W_hh^T, this CTA's k-blocks
      mbar_expect_tx(w_full, (uint32_t)KB * W_KB_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + (size_t)kb * W_KB_BYTES, &tmWT, kcol0 + kb * 64, c * U, w_full);
    }
  } else if (warp == 1) {
    {   // the MMA warp stays converged; one elected lane issues (see tc::elect_one)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(w_full, 0);
      const uint64_t desc0 = make_smem_desc(0, 16, 1024);
      int it = 0;
      for (int i = 1; i <= T; ++i) {
        mbar_wait(tmem_empty, ((i - 1) & 1) ^ 1);
        tc_fence_after();
        for (int ss = 0; ss < NSS; ++ss, ++it) {
          const int s = it % NSLOT;
          const uint32_t ph = (it / NSLOT) & 1;
          mbar_wait(&full_bar[s], ph);
          if (ss == 0 && lane == 0) trace(i, 2);
          fence_proxy_async_smem(); // generic-proxy (cp.async) writes -> async-proxy (tcgen05.mma) reads, on the consumer side
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
          for (int kl = 0; kl < SS_KB; ++kl) {
            const int kb = ss * SS_KB + kl;
            // descriptors differ only in the start-address field (bits 0-13, 16-byte units): +2 per 32-byte k-step
            const uint64_t da = desc0 + ((smem_u32(sA + (size_t)s * SS_BYTES + kl * A_STAGE_BYTES)) >> 4);
            const uint64_t db = desc0 + ((smem_u32(sW + (size_t)kb * W_KB_BYTES)) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // k-step j = kb*4 + k -> accumulator j % 8 = kl*4 + k (SS_KB == 2, kb = ss*2 + kl)
              tc_mma<false>(tmem_base + (uint32_t)((kl * 4 + k) * N), da + 2 * k, db + 2 * k, idesc, ss != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(tmem_full);
        __syncwarp();
        if (lane == 0) trace(i, 3);
      }
    }
  } else {
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;            // column group: units cg*4 .. cg*4+3 of this slice's 16
    const bool owner = (cg % KS) == kq;        // this CTA finishes these 4 units (cell gradient, stores)
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const bool valid = row < a.B;
    const int j = c * U + cg * 4;              // first of this thread's 4 hidden units
    // ---- loader role: this thread's 16-byte chunks of a super-stage (fixed for the whole kernel) ----
    const int tl = threadIdx.x - 64;           // 0..511
    const int rows = min(128, a.B - m0);
    const int per_kb = rows * 8;
    uint32_t ld_dst[BWD_LPT];                  // byte offset inside a ring slot (128B swizzle: chunk ^= row & 7)
    long long ld_src[BWD_LPT];                 // element offset from dgates16 + step/super-stage base; < 0: none
#pragma unroll
    for (int m = 0; m < BWD_LPT; ++m) {
      const int qi = tl + m * BWD_LOADERS;
      const int kl = qi / per_kb, rem = qi - kl * per_kb;
      const int r = rem >> 3, ch = rem & 7;
      ld_src[m] = kl < SS_KB ? (long long)(m0 + r) * T * 4 * H + kl * 64 + ch * 8 : -1;
      ld_dst[m] = (uint32_t)(kl * A_STAGE_BYTES + r * 128 + ((ch ^ (r & 7)) << 4));
    }
    const uint32_t sA_u32 = smem_u32(sA);
    // where this thread's partial goes when another CTA owns its units: recv[src = kq][cg][rloc] in CTA (cg % KS)
    uint32_t send_addr = 0, send_bar = 0;
    if (KS > 1 && !owner) {
      send_addr = mapa_u32(smem_u32(recv) + (uint32_t)(((kq * 4 + cg) * 128 + rloc) * 16), (uint32_t)(cg % KS));
      send_bar = mapa_u32(smem_u32(recv_bar), (uint32_t)(cg % KS));
    }
    int it_issue = 0;                          // super-stages issued so far (all steps)
    float dcreg[4], dhrec[4];                  // dc / dh flowing from step t+1 into step t
    {
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 c4 = (owner && valid && a.d_cT) ? *reinterpret_cast<const float4*>(a.d_cT + (long long)row * H + j) : z;
      const float4 h4 = (owner && valid && a.d_hT) ? *reinterpret_cast<const float4*>(a.d_hT + (long long)row * H + j) : z;
      dcreg[0] = c4.x; dcreg[1] = c4.y; dcreg[2] = c4.z; dcreg[3] = c4.w;
      dhrec[0] = h4.x; dhrec[1] = h4.y; dhrec[2] = h4.z; dhrec[3] = h4.w;
    }
    for (int i = 0; i <= T; ++i) {
      const int t = T - 1 - i;
      // operands of step t that do not depend on the recurrence: issue the loads before waiting for the MMA
      float4 dha, dhs4, dcl, ig4, fg4, gg4, og4, ce4, cp4;
      dha = dhs4 = dcl = ig4 = fg4 = gg4 = og4 = ce4 = cp4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const long long bt = (long long)row * T + (t < 0 ? 0 : t);
      if (owner && valid && t >= 0) {
        dha = *reinterpret_cast<const float4*>(a.dh_attn + bt * H + j);
        if (a.dhs && t + 1 < T) dhs4 = *reinterpret_cast<const float4*>(a.dhs + (bt + 1) * H + j);
        dcl = *reinterpret_cast<const float4*>(a.dcell + bt * H + j);
        const float* ac = a.acts + bt * 4 * H + j;
        ig4 = *reinterpret_cast<const float4*>(ac);
        fg4 = *reinterpret_cast<const float4*>(ac + H);
        gg4 = *reinterpret_cast<const float4*>(ac + 2 * H);
        og4 = *reinterpret_cast<const float4*>(ac + 3 * H);
        ce4 = *reinterpret_cast<const float4*>(a.cells + bt * H + j);
        if (t > 0) cp4 = *reinterpret_cast<const float4*>(a.cells + (bt - 1) * H + j);
        else if (a.c0) cp4 = *reinterpret_cast<const float4*>(a.c0 + (long long)row * H + j);
      }
      if (i >= 1) {
        // ---- feed GEMM i: this CTA's K range of dgates of step T - i, complete once the counter says so ----
        if (threadIdx.x == 64) {
          grid_barrier_wait(a.counters + rg, (unsigned)i * CT);
          trace(i, 0);
        }
        epilogue_bar<BWD_LOADERS>();
        const __nv_bfloat16* src0 = a.dgates16 + (long long)(T - i) * 4 * H + kcol0;
        for (int ss = 0; ss < NSS; ++ss, ++it_issue) {
          const int s = it_issue % NSLOT;
          mbar_wait(&empty_bar[s], ((it_issue / NSLOT) & 1) ^ 1);
          const uint32_t dst0 = sA_u32 + (uint32_t)s * SS_BYTES;
          const __nv_bfloat16* src = src0 + ss * (SS_KB * 64);
#pragma unroll
          for (int m = 0; m < BWD_LPT; ++m)
            if (ld_src[m] >= 0) cp_async16(dst0 + ld_dst[m], src + ld_src[m]);
          // completion of this thread's copies is counted by the slot's mbarrier itself (no wait, no fence here: a
          // writer-side fence.proxy.async would drain the younger copies and serialise the pipeline)
          cp_async_arrive_noinc(&full_bar[s]);
        }
        if (threadIdx.x == 64) trace(i, 1);

        mbar_wait(tmem_full, (i - 1) & 1);
        if (threadIdx.x == 64) trace(i, 4);
        tc_fence_after();
        uint32_t r[NACC][4];
#pragma unroll
        for (int ai = 0; ai < NACC; ++ai) tmem_ld<4>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ai * N + cg * 4), r[ai]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty);
        float part[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float acc = __uint_as_float(r[0][u]);
#pragma unroll
          for (int ai = 1; ai < NACC; ++ai) acc += __uint_as_float(r[ai][u]);
          part[u] = acc;
        }
        if (KS > 1) {
          if (!owner) {        // hand the partial to the CTA that owns these units
            st_cluster_f4(send_addr, part[0], part[1], part[2], part[3]);
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(send_bar);
          } else {             // gather the peers' partials of this thread's units
            mbar_wait_cluster(recv_bar, (i - 1) & 1);
#pragma unroll
            for (int src = 0; src < KS; ++src) {
              if (src != kq) {
                const float4 o = *reinterpret_cast<const float4*>(recv + ((src * 4 + cg) * 128 + rloc) * 4);
                part[0] += o.x; part[1] += o.y; part[2] += o.z; part[3] += o.w;
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) dhrec[u] = part[u];
      }
      if (t < 0) break;
      float d0[4], d1[4], d2[4], d3[4];
      if (owner && valid) {
        const float dhav[4] = {dha.x + dhs4.x, dha.y + dhs4.y, dha.z + dhs4.z, dha.w + dhs4.w};
        const float dclv[4] = {dcl.x, dcl.y, dcl.z, dcl.w};
        const float igv[4] = {ig4.x, ig4.y, ig4.z, ig4.w}, fgv[4] = {fg4.x, fg4.y, fg4.z, fg4.w};
        const float ggv[4] = {gg4.x, gg4.y, gg4.z, gg4.w}, ogv[4] = {og4.x, og4.y, og4.z, og4.w};
        const float cev[4] = {ce4.x, ce4.y, ce4.z, ce4.w}, cpv[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dh = dhav[e] + dhrec[e];
          const float tcv = tanhf_fast(cev[e]);
          const float dc = dclv[e] + dcreg[e] + dh * ogv[e] * (1.f - tcv * tcv);
          d0[e] = dc * ggv[e] * igv[e] * (1.f - igv[e]);
          d1[e] = dc * cpv[e] * fgv[e] * (1.f - fgv[e]);
          d2[e] = dc * igv[e] * (1.f - ggv[e] * ggv[e]);
          d3[e] = dh * tcv * ogv[e] * (1.f - ogv[e]);
          dcreg[e] = dc * fgv[e];
        }
        // only the bf16 mirror feeds the next step's contraction: store it first, the fp32 copy after the arrive
        bf16* dg16 = a.dgates16 + bt * 4 * H + j;
        st_bf16x4(dg16, d0[0], d0[1], d0[2], d0[3]);
        st_bf16x4(dg16 + H, d1[0], d1[1], d1[2], d1[3]);
        st_bf16x4(dg16 + 2 * H, d2[0], d2[1], d2[2], d2[3]);
        st_bf16x4(dg16 + 3 * H, d3[0], d3[1], d3[2], d3[3]);
      }
      if (threadIdx.x == 64) trace(i, 5);
      epilogue_bar<BWD_LOADERS>();
      if (threadIdx.x == 64) trace(i, 6);
      // red.release.gpu is cumulative over the stores observed through the barrier: no separate fence
      if (warp == 2 && lane == 0) red_release_gpu_add(a.counters + rg, 1u);
      if (threadIdx.x == 64) trace(i, 7);
      if (owner && valid) {
        float* dg = a.dgates + bt * 4 * H + j;
        *reinterpret_cast<float4*>(dg) = make_float4(d0[0], d0[1], d0[2], d0[3]);
        *reinterpret_cast<float4*>(dg + H) = make_float4(d1[0], d1[1], d1[2], d1[3]);
        *reinterpret_cast<float4*>(dg + 2 * H) = make_float4(d2[0], d2[1], d2[2], d2[3]);
        *reinterpret_cast<float4*>(dg + 3 * H) = make_float4(d3[0], d3[1], d3[2], d3[3]);
      }
    }
    // after the last GEMM: dhrec = dgates_0 W_hh = dh0 ; dcreg = dc0
    if (owner && valid) {
      if (a.dh0) *reinterpret_cast<float4*>(a.dh0 + (long long)row * H + j) = make_float4(dhrec[0], dhrec[1], dhrec[2], dhrec[3]);
      if (a.dc0) *reinterpret_cast<float4*>(a.dc0 + (long long)row * H + j) = make_float4(dcreg[0], dcreg[1], dcreg[2], dcreg[3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TCOLS) : "memory");
  }
  if (KS > 1) {   // no CTA of the cluster may exit while a peer could still address its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

// ---- weight re-layouts (fp32 master -> bf16 operand) -----------------------------------------
// forward: Wp[(c*4U + u*4 + g), k] = W_hh[g*H + c*U + u, k]
__global__ void pack_whh_fwd_kernel(const float* __restrict__ w_hh, bf16* __restrict__ wp, int H, int U) {
  const int n = blockIdx.x;                 // packed row
  const int c = n / (4 * U), rem = n % (4 * U), u = rem / 4, g = rem % 4;
  const float* src = w_hh + ((long long)g * H + c * U + u) * H;
  bf16* dst = wp + (long long)n * H;
  for (int k = threadIdx.x; k < H; k += blockDim.x) dst[k] = __float2bfloat16(src[k]);
}
// backward: WT[j, n] = W_hh[n, j]   ([H, 4H] bf16), 32x32 smem tiles
__global__ void transpose_whh_kernel(const float* __restrict__ w_hh, bf16* __restrict__ wt, int H) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) tile[r][threadIdx.x] = w_hh[(long long)(n0 + r) * H + j0 + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) wt[(long long)(j0 + r) * 4 * H + n0 + threadIdx.x] = __float2bfloat16(tile[threadIdx.x][r]);
}

// depth of the activation ring: as deep as shared memory allows next to the resident weight slice
inline int seq_stages(size_t w_bytes) {
  if (w_bytes + 2 * A_STAGE_BYTES > SEQ_SMEM_BUDGET) return 0;
  size_t n = (SEQ_SMEM_BUDGET - w_bytes) / A_STAGE_BYTES;
  return (int)(n > MAX_STAGES ? MAX_STAGES : n);
}

// backward: ring slots of one super-stage (SS_BYTES) next to the resident W_hh^T slice (works from 2)
inline int bwd_slots(size_t w_bytes) {
  if (w_bytes + 2 * SS_BYTES > SEQ_SMEM_BUDGET) return 0;
  size_t n = (SEQ_SMEM_BUDGET - w_bytes) / SS_BYTES;
  return (int)(n > MAX_STAGES ? MAX_STAGES : n);
}

template <int U>
int launch_fwd_u(const LstmSeqFwd& p, int C, int RG, cudaStream_t st) {
  constexpr int N = 4 * U;
  const int H = p.H, KB = H / 64;
  pack_whh_fwd_kernel<<<4 * H, 128, 0, st>>>(p.w_hh, p.whh_packed16, H, U);
  AA_CHECK_LAUNCH("pack_whh_fwd");
  AA_CHECK_CUDA(cudaMemsetAsync(p.counters, 0, sizeof(unsigned) * RG, st));
  CUtensorMap tmW, tmH0, tmH;
  AA_TRY(make_map(&tmW, p.whh_packed16, 2, 4LL * H, H, H, N));
  const int box_rows = (p.B >= 128 ? 128 : (p.B + 7) / 8 * 8);   // rows beyond the batch are never loaded
  AA_TRY(make_map(&tmH0, p.h016, 2, p.B, H, H, box_rows));
  AA_TRY(make_map(&tmH, p.hid16, 2, p.B, (long long)p.T * H, (long long)p.T * H, box_rows));
  SeqFwdArgs a{};
  a.B = p.B; a.T = p.T; a.H = H; a.xg = p.xg; a.c0 = p.c0;
  a.hiddens = p.hiddens; a.cells = p.cells; a.acts = p.acts; a.hs_prev = p.hs_prev; a.hid16 = p.hid16; a.hsprev16 = p.hsprev16;
  a.counters = p.counters;
  a.stages = seq_stages((size_t)KB * N * 128);
  a.box_rows = box_rows;
  const size_t smem = (size_t)a.stages * A_STAGE_BYTES + (size_t)KB * N * 128 + (2 * MAX_STAGES + 4) * 8 + 16 + 1024;
  auto kern = lstm_seq_fwd_kernel<U>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  void* args[] = {(void*)&tmW, (void*)&tmH0, (void*)&tmH, (void*)&a};
  AA_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(C, RG), dim3(SEQ_THREADS), args, smem, st));
  count_launch();
  return AA_OK;
}

}  // namespace

int set_seq_trace_buffer(void* dev_ptr) {
  unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
  AA_CHECK_CUDA(cudaMemcpyToSymbol(g_seq_trace, &p, sizeof(p)));
  return AA_OK;
}

bool lstm_seq_supported(int B, int H, int* units_fwd) {
  if (H % 64 != 0 || H < 64 || B < 1) return false;
  const int RG = ceil_div(B, 128);
  const int sms = num_sms();
  if ((H / 16) * RG > sms) return false;                         // backward: 16 units per CTA
  if (bwd_slots((size_t)(4 * H / 64) * 16 * 128) < 2) return false;
  for (int U : {4, 8, 16, 32}) {
    if (H % U) continue;
    if ((H / U) * RG <= sms && seq_stages((size_t)(H / 64) * 4 * U * 128) >= 2) {
      if (units_fwd) *units_fwd = U;
      return true;
    }
  }
  return false;
}

int launch_lstm_seq_fwd(const LstmSeqFwd& p, cudaStream_t st) {
  int U = 0;
  AA_REQUIRE(lstm_seq_supported(p.B, p.H, &U), "lstm_seq_fwd: unsupported shape B=%d H=%d", p.B, p.H);
  const int C = p.H / U, RG = ceil_div(p.B, 128);
  switch (U) {
    case 4: return launch_fwd_u<4>(p, C, RG, st);
    case 8: return launch_fwd_u<8>(p, C, RG, st);
    case 16: return launch_fwd_u<16>(p, C, RG, st);
    default: return launch_fwd_u<32>(p, C, RG, st);
  }
}

template <int KS>
int launch_bwd_ks(const LstmSeqBwd& p, cudaStream_t st) {
  const int H = p.H, KB = 4 * H / 64 / KS, C = H / 16, RG = ceil_div(p.B, 128);
  CUtensorMap tmWT;
  AA_TRY(make_map(&tmWT, p.whhT16, 2, H, 4LL * H, 4LL * H, 16));
  SeqBwdArgs a{};
  a.B = p.B; a.T = p.T; a.H = H;
  a.dh_attn = p.dh_attn; a.dhs = p.dhs; a.dcell = p.dcell; a.d_hT = p.d_hT; a.d_cT = p.d_cT;
  a.acts = p.acts; a.cells = p.cells; a.c0 = p.c0; a.dgates = p.dgates; a.dgates16 = p.dgates16; a.dh0 = p.dh0; a.dc0 = p.dc0;
  a.counters = p.counters;
  const size_t fixed = (size_t)KB * 16 * 128 + (KS > 1 ? (size_t)KS * 4 * 128 * 16 : 0);
  a.stages = bwd_slots(fixed);
  if (a.stages > KB / SS_KB * 2) a.stages = KB / SS_KB * 2 > 2 ? KB / SS_KB * 2 : 2;   // no point in more than two steps' worth of slots
  a.box_rows = 0;
  const size_t smem = (size_t)a.stages * SS_BYTES + fixed + (2 * MAX_STAGES + 5) * 8 + 16 + 1024;
  auto kern = lstm_seq_bwd_kernel<KS>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  void* args[] = {(void*)&tmWT, (void*)&a};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C * KS, RG);
  cfg.blockDim = dim3(BWD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  // The grid barrier needs every CTA resident.  Without clusters that is what a cooperative launch checks; with clusters the
  // residency check is cudaOccupancyMaxActiveClusters below (Nsight Compute cannot replay a launch that is both cooperative
  // and clustered: "LaunchFailed"), the launch itself is a plain cluster launch.
  cudaLaunchAttribute at[1];
  if (KS > 1) {
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = KS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
  } else {
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
  }
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (KS > 1) {   // clusters are placed inside one GPC: ask the runtime how many fit at once (varies with floor-sweeping / profilers)
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, (const void*)kern, &cfg) != cudaSuccess || max_clusters < C * RG) {
      cudaGetLastError();
      return AA_ERR_UNSUPPORTED;
    }
  }
  const cudaError_t le = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
  if (le != cudaSuccess) {
    if (KS > 1) {          // a refused cluster launch (e.g. under a profiler that cannot replay cooperative cluster grids) is not
      cudaGetLastError();  // fatal: the caller retries without the K-split
      return AA_ERR_UNSUPPORTED;
    }
    set_error("cudaLaunchKernelExC(lstm_seq_bwd) failed: %s", cudaGetErrorString(le));
    return AA_ERR_CUDA;
  }
  count_launch();
  return AA_OK;
}

// K-split of the backward contraction (cluster size): as wide as the chip and the shape allow
int bwd_ksplit(int B, int H) {
  const int C = H / 16, RG = ceil_div(B, 128), sms = num_sms();
  for (int ks : {4, 2}) {
    if ((4 * H / 64) % (ks * SS_KB) != 0) continue;
    if (C * ks * RG > sms) continue;
    const size_t fixed = (size_t)(4 * H / 64 / ks) * 16 * 128 + (size_t)ks * 4 * 128 * 16;
    if (bwd_slots(fixed) >= 2) return ks;
  }
  return 1;
}

int g_bwd_ksplit_max = 4;   // diagnostics (aa_debug_set_bptt_ksplit): cap on the K-split

int launch_lstm_seq_bwd(const LstmSeqBwd& p, cudaStream_t st) {
  AA_REQUIRE(lstm_seq_supported(p.B, p.H, nullptr), "lstm_seq_bwd: unsupported shape B=%d H=%d", p.B, p.H);
  const int H = p.H, RG = ceil_div(p.B, 128);
  transpose_whh_kernel<<<dim3(4 * H / 32, H / 32), dim3(32, 8), 0, st>>>(p.w_hh, p.whhT16, H);
  AA_CHECK_LAUNCH("transpose_whh");
  AA_CHECK_CUDA(cudaMemsetAsync(p.counters, 0, sizeof(unsigned) * RG, st));
  int ks = bwd_ksplit(p.B, H);
  if (ks > g_bwd_ksplit_max) ks = g_bwd_ksplit_max;
  static const int env_cap = [] {   // AA_BPTT_KSPLIT=1 in the environment: never use the cluster variant
    const char* e = getenv("AA_BPTT_KSPLIT");
    return e ? atoi(e) : 4;
  }();
  if (ks > env_cap) ks = env_cap < 1 ? 1 : env_cap;
  int rc = AA_ERR_UNSUPPORTED;
  if (ks == 4) rc = launch_bwd_ks<4>(p, st);
  else if (ks == 2) rc = launch_bwd_ks<2>(p, st);
  if (rc == AA_ERR_UNSUPPORTED) rc = launch_bwd_ks<1>(p, st);
  return rc;
}

int set_bptt_ksplit_max(int ks) {
  g_bwd_ksplit_max = ks < 1 ? 1 : ks;
  return AA_OK;
}

}  // namespace aa
