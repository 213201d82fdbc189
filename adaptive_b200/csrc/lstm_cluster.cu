// LSTM recurrence and its BPTT inside thread-block clusters, weights resident in TENSOR MEMORY (bf16 mode, H <= 512).
//
// lstm_seq.cu spreads one step over 128 CTAs that meet at a grid barrier in global memory: a step costs two L2 round
// trips (publish h_t, re-read it in every CTA) plus ~32 tcgen05.mma whose A operand (M = 128 batch rows) is re-read from
// shared memory at ~103 cycles each -- 6.7 us (forward) / 8 us (backward) per step at B=80, H=512 however little work
// a step has (baseline_attention.py:167-178 is a chain of T dependent [B,H] x [H,4H] products).
//
// Here nothing of a step leaves the chip:
//  * batch rows are independent sequences, so the batch is cut into groups of <= 16 rows and every group runs on its
//    own cluster of CL = H/32 CTAs (8 clusters x 10 rows at B = 80: 128 SMs busy, no inter-cluster communication);
//  * CTA c of a cluster owns hidden units [32c, 32c+32), i.e. 128 gate columns.  The operands are swapped with
//    respect to lstm_seq.cu: the WEIGHT slice is the M = 128 side of the MMA and lives in tensor memory for the whole
//    kernel (tcgen05.mma with A in TMEM: no per-step operand read from shared memory for it), the batch is the N = 16
//    side, [16 x K] bf16 in shared memory;
//  * forward: gates^T[128, 16] = W_slice[128, H] . h_{t-1}^T; every CTA applies the LSTM cell to its 32 units and
//    writes its bf16 slice of h_t straight into the operand buffer of all CL CTAs (st.shared::cluster), completion
//    counted by an mbarrier in the receiver (remote arrive, release/acquire at cluster scope);
//  * backward: dh_{t-1} = dgates_t W_hh is contracted over the gate columns.  Each CTA contracts over ITS OWN 128 gate
//    columns (the dgates it has just produced: no broadcast) against W_hh[own columns, :]^T held in TMEM as H/128
//    M-blocks, and the fp32 partials [H, 16] are reduce-scattered through distributed shared memory to the CTAs owning
//    the units.
// The weights come from the plain bf16 copy of W_hh (rows g*H + j, gate order i,f,g,o) by TMA; no re-packed copy.
#include <stdlib.h>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace aa {

namespace {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int CLK_THREADS = 160;          // warps 0-3: cell math (TMEM lane quarter = warp), warp 4: TMA + MMA issue
constexpr int NB = 16;                    // UMMA N = batch rows per cluster (padded)
constexpr uint32_t W_TILE = 32 * 128;     // one TMA box of the weight slice: 32 rows x 64 bf16 (128B swizzle)
constexpr uint32_t B_KB = NB * 128;       // one 64-wide k-block of the batch operand: 16 rows x 128 B
constexpr uint32_t D_COL = 256;           // first accumulator column (A occupies columns [0, H/2) <= 256)
constexpr uint32_t TMEM_COLS = 512;

// ---- PTX helpers --------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return (uint32_t)v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint16_t bf16_bits(float x) {
  bf16 h = __float2bfloat16(x);
  return *reinterpret_cast<uint16_t*>(&h);
}
template <int NT>
__device__ __forceinline__ void named_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }

// tcgen05.mma with the A operand in tensor memory (M = 128: lane = row, 32-bit column = two consecutive K elements)
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// bulk copy local shared memory -> a peer CTA's shared memory; completion (bytes) is counted by an mbarrier of the peer
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}
// K-major operand with the 64-byte swizzle: rows of 64 B (32 bf16), 16-byte chunk index ^= (row >> 1) & 3, 8-row groups
// sbo bytes apart (512 when dense); layout type 4
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// optional per-step timeline of CTA 0 of cluster 0 (aa_debug_set_trace_buffer): 8 x uint64 globaltimer stamps per step
//   0 operand complete (MMA warp)   1 MMAs issued   2 accumulator seen   3 activations staged   4 cell done
//   5 slice delivered / partials sent   6 partials received (bwd)   7 outputs stored
__device__ unsigned long long* g_clk_trace = nullptr;
__device__ __forceinline__ void cl_trace(int step, int ev) {
  if (g_clk_trace && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_clk_trace[step * 8 + ev] = t;
  }
}

// Common prologue: barriers, TMEM, the CTA's weight slice (rows g*H + 32c + u, all H columns) by TMA into sW as
// [4 gates][H/64 k-blocks] tiles of 32 rows x 128 B.
struct ClSmem {
  uint8_t* base;       // 1024-aligned
  uint32_t sW;         // shared-space byte addresses
  uint64_t* bars;      // [0] w_full  [1] tmem_full  [2],[3] operand / partials complete (by parity)
  uint32_t* tmem_slot;
};

constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// =============================================================================================================
// forward
// =============================================================================================================
struct ClFwdArgs {
  int B, T, H, rpg;            // rpg = batch rows per cluster (<= 16)
  const float* xg;             // [B,T,4H] input-half gate pre-activations incl. both biases
  const float* c0;             // [B,H] or null
  const bf16* h016;            // [B,H]
  float *hiddens, *cells, *acts, *hs_prev;
  bf16 *hid16, *hsprev16;
};

// Operand layout of h_{t-1} in shared memory: [slice c of 32 units][row (16)][64 B] with the 64-byte swizzle (a row of the
// K-major operand is exactly one CTA's 32 units), so that the slice CTA c produces is one contiguous block in every
// receiver: it is delivered with ONE bulk shared->shared::cluster copy per peer whose bytes the receiver's mbarrier counts --
// no per-thread remote stores, no release fence waiting for their acknowledgements, and the data arrives through the async
// proxy the tensor core reads with.  (A layout without swizzle is contiguous too, but its MMAs ran at 1.8 us per 32 against
// 1.15 us with the 128-byte swizzle: profiles/r01_v36_lstm_trace.txt vs r01_v38_lstm_trace.txt.)
constexpr uint32_t SLICE_BYTES = NB * 64;       // 1 KB
__device__ __forceinline__ uint32_t sw64_off(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

template <int NACC>
__global__ void __launch_bounds__(CLK_THREADS, 1) lstm_clk_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const ClFwdArgs a) {
  const int H = a.H, T = a.T, KB = H / 64, CL = H / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sW = smem_u32(smem);                               // [4][KB] weight tiles (prologue only)
  const uint32_t sB = sW + 4u * KB * W_TILE;                         // [2 parity][CL slices][1 KB]  h_{t-1} operand
  const uint32_t sAct = sB + 2u * CL * SLICE_BYTES;                  // [4 gates][16 rows][32 units] fp32
  const uint32_t sH16 = sAct + 4u * NB * 32 * 4;                     // [2 parity][1 KB]: this CTA's slice of h_t, operand layout
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (sH16 - sW) + 2 * SLICE_BYTES);
  uint64_t* w_full = bars;
  uint64_t* tmem_full = bars + 1;
  uint64_t* full = bars + 2;                                         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = cluster_ctarank();
  const int m0 = blockIdx.y * a.rpg;
  const int rows = min(a.rpg, a.B - m0);
  const uint32_t step_bytes = (uint32_t)(CL * rows * 64);            // what one step delivers into this CTA
  if (threadIdx.x == 0) cl_trace(40, 0);                             // (row 40 of the trace: kernel entry, prologue done, loop done, exit)

  if (warp == 4 && lane == 0) {
    mbar_init(w_full, 1);
    mbar_init(tmem_full, 1);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // operand + staging buffers: zero (rows beyond the group stay zero), then h_0 of this group into parity 0
  for (uint32_t i = threadIdx.x; i < 2u * CL * SLICE_BYTES / 16; i += CLK_THREADS) sts128(sB + i * 16, make_uint4(0, 0, 0, 0));
  for (uint32_t i = threadIdx.x; i < 2u * SLICE_BYTES / 16; i += CLK_THREADS) sts128(sH16 + i * 16, make_uint4(0, 0, 0, 0));
  __syncthreads();
  if (warp == 4 && elect_one()) {
    mbar_expect_tx(w_full, 4u * KB * W_TILE);
    for (int g = 0; g < 4; ++g)
      for (int kb = 0; kb < KB; ++kb)
        tma_load_2d(smem + (size_t)(g * KB + kb) * W_TILE, &tmW, kb * 64, g * H + (int)c * 32, w_full);
  }
  for (int i = threadIdx.x; i < rows * (H / 8); i += CLK_THREADS) {
    const int r = i / (H / 8), cc = i - r * (H / 8);               // cc = 16-byte chunk (8 units) of row r
    const uint4 v = *reinterpret_cast<const uint4*>(a.h016 + (long long)(m0 + r) * H + cc * 8);
    sts128(sB + (uint32_t)(cc >> 2) * SLICE_BYTES + sw64_off(r, cc & 3), v);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();   // every CTA's barriers are initialised and its operand buffers cleared before remote traffic

  if (warp < 4) {       // weight slice -> tensor memory: lane = gate column (gate = warp, unit = lane), columns = K pairs
    mbar_wait(w_full, 0);
    for (int kb = 0; kb < KB; ++kb) {
      const uint32_t row = sW + (uint32_t)(warp * KB + kb) * W_TILE + (uint32_t)lane * 128u;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint4 v = lds128(row + (uint32_t)(((half * 4 + ch) ^ (lane & 7)) << 4));
          r[ch * 4 + 0] = v.x; r[ch * 4 + 1] = v.y; r[ch * 4 + 2] = v.z; r[ch * 4 + 3] = v.w;
        }
        tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kb * 32 + half * 16), r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) cl_trace(40, 1);

  if (warp == 4) {
    // ===== MMA issuer (converged warp, one elected lane) =====
    const uint64_t desc0 = make_smem_desc_sw64(0, 512);
    fence_proxy_async_smem();                    // h_0 was written with st.shared (generic proxy)
    for (int t = 0; t < T; ++t) {
      const uint32_t p = (uint32_t)t & 1u;
      if (t > 0) mbar_wait(&full[p], (uint32_t)((t - 1) >> 1) & 1u);           // all CL slices of h_{t-1} have landed
      if (lane == 0) cl_trace(t, 0);
      tc_fence_after();
      if (elect_one()) {
        if (t + 1 < T) mbar_expect_tx(&full[p ^ 1u], step_bytes);            // arm the buffer h_t will be delivered into
        // k-step ks = 16 units = half a row (32 B) of slice ks/2: the descriptors differ from the parity's base only in the
        // start-address field (16-byte units): + (ks/2) * 64 + (ks%2) * 2 -- one add per MMA in the unrolled loop
        const uint64_t dbase = desc0 + (((sB + p * CL * SLICE_BYTES) & 0x3FFFFu) >> 4);
#pragma unroll 8
        for (int ks = 0; ks < 4 * KB; ++ks)
          tc_mma_ts(tmem_base + D_COL + (uint32_t)((ks % NACC) * NB), tmem_base + (uint32_t)(ks * 8),
                    dbase + (uint64_t)((ks >> 1) * (SLICE_BYTES / 16) + (ks & 1) * 2), IDESC, ks >= NACC ? 1u : 0u);
        tc_commit(tmem_full);
      }
      __syncwarp();
      if (lane == 0) cl_trace(t, 1);
    }
  } else {
    // ===== activations: thread = (gate = warp, unit = lane) over all rows; cell: thread = (rows warp + 4i, unit = lane) =====
    const int g = warp, u = lane;
    const int j = (int)c * 32 + u;
    const long long T4H = (long long)T * 4 * H;
    const float* xg_base = a.xg + (long long)m0 * T4H + (long long)g * H + j;
    // one branch-free form for all four gates: sigmoid(x) = 1/(1+e^-x), tanh(x) = 2 sigmoid(2x) - 1
    const float a_in = g == 2 ? -2.f : -1.f, a_mul = g == 2 ? 2.f : 1.f, a_add = g == 2 ? -1.f : 0.f;
    float xc[NB], creg[4];
#pragma unroll
    for (int b = 0; b < NB; ++b) xc[b] = b < rows ? xg_base[b * T4H] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = warp + 4 * i;
      creg[i] = (b < rows && a.c0) ? a.c0[(long long)(m0 + b) * H + j] : 0.f;
    }
    for (int t = 0; t < T; ++t) {
      float xn[NB];
      if (t + 1 < T) {
#pragma unroll
        for (int b = 0; b < NB; ++b) xn[b] = b < rows ? xg_base[b * T4H + (long long)(t + 1) * 4 * H] : 0.f;
      }
      mbar_wait(tmem_full, (uint32_t)t & 1u);
      if (threadIdx.x == 0) cl_trace(t, 2);
      tc_fence_after();
      uint32_t r[NB];
      tmem_ld<16>(tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL, r);
#pragma unroll
      for (int ai = 1; ai < NACC; ++ai) {
        uint32_t r2[NB];
        tmem_ld<16>(tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL + (uint32_t)(ai * NB), r2);
#pragma unroll
        for (int e = 0; e < NB; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
      }
      tc_fence_before();
      float act[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float e = __expf(a_in * (__uint_as_float(r[b]) + xc[b]));
        act[b] = fmaf(a_mul, __fdividef(1.0f, 1.0f + e), a_add);
      }
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (b < rows) sts_f32(sAct + (uint32_t)(((g * NB + b) * 32 + u) * 4), act[b]);
      named_bar<128>(1);
      if (threadIdx.x == 0) cl_trace(t, 3);
      float hv[4];
      const uint32_t stage = sH16 + ((uint32_t)t & 1u) * SLICE_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = warp + 4 * i;
        hv[i] = 0.f;
        if (b < rows) {
          const float ig = lds_f32(sAct + (uint32_t)(((0 * NB + b) * 32 + u) * 4));
          const float fg = lds_f32(sAct + (uint32_t)(((1 * NB + b) * 32 + u) * 4));
          const float gg = lds_f32(sAct + (uint32_t)(((2 * NB + b) * 32 + u) * 4));
          const float og = lds_f32(sAct + (uint32_t)(((3 * NB + b) * 32 + u) * 4));
          creg[i] = fg * creg[i] + ig * gg;
          hv[i] = og * tanhf_fast(creg[i]);
          sts_u16(stage + sw64_off(b, u >> 3) + (uint32_t)((u & 7) * 2), bf16_bits(hv[i]));
        }
      }
      if (t + 1 < T) {
        fence_proxy_async_smem();                // the staged slice is read by the bulk-copy engine (async proxy)
        named_bar<128>(1);
        if (threadIdx.x == 0) cl_trace(t, 4);
        // this CTA's slice of h_t into every CTA's operand buffer of the other parity: one copy per peer
        if ((int)threadIdx.x < CL) {
          const uint32_t peer = threadIdx.x;
          const uint32_t pn = ((uint32_t)t + 1u) & 1u;
          bulk_copy_to_peer(mapa_u32(sB + (pn * CL + c) * SLICE_BYTES, peer), stage, (uint32_t)rows * 64u, mapa_u32(smem_u32(&full[pn]), peer));
        }
      } else {
        named_bar<128>(1);                       // (sAct is rewritten by the next step: keep the step structure)
      }
      if (threadIdx.x == 0) cl_trace(t, 5);
      // everything below is only read after the kernel
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (b < rows) a.acts[((long long)(m0 + b) * T + t) * 4 * H + (long long)g * H + j] = act[b];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = warp + 4 * i;
        if (b < rows) {
          const long long bt = (long long)(m0 + b) * T + t;
          a.hiddens[bt * H + j] = hv[i];
          a.cells[bt * H + j] = creg[i];
          a.hid16[bt * H + j] = __float2bfloat16(hv[i]);
          if (t + 1 < T) {
            a.hs_prev[(bt + 1) * H + j] = hv[i];
            a.hsprev16[(bt + 1) * H + j] = __float2bfloat16(hv[i]);
          }
        }
      }
      if (threadIdx.x == 0) cl_trace(t, 7);
      if (t + 1 < T) {
#pragma unroll
        for (int b = 0; b < NB; ++b) xc[b] = xn[b];
      }
    }
  }
  if (threadIdx.x == 0) cl_trace(40, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  cluster_sync_all();   // no CTA leaves while a peer could still address its shared memory
  if (threadIdx.x == 0) cl_trace(40, 3);
}

// =============================================================================================================
// backward
// =============================================================================================================
struct ClBwdArgs {
  int B, T, H, rpg, RP;                   // RP = floats per (source, unit) row of the partials buffer (see bwd_rp)
  const float *dh_attn, *dhs, *dcell;     // [B,T,H]: batched (non-recurrent) gradient contributions; dhs may be null
  const float *d_hT, *d_cT;               // [B,H] or null
  const float *acts, *cells, *c0;         // saved forward state; c0 may be null (zeros)
  float* dgates; bf16* dgates16;          // [B,T,4H]
  float *dh0, *dc0;                       // [B,H] (may be null)
  bf16* dg16_t0;                          // optional [B,4H]: second copy of the step-0 rows of dgates16
};

__global__ void __launch_bounds__(CLK_THREADS, 1) lstm_clk_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const ClBwdArgs a) {
  const int H = a.H, T = a.T, KB = H / 64, CL = H / 32, MB = H / 128, RP = a.RP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t blk = 32u * (uint32_t)RP * 4u;                      // partials of 32 units from one source CTA
  const uint32_t region = max(4u * KB * W_TILE, 4u * CL * blk);
  const uint32_t sW = smem_u32(smem);                               // [4][KB] weight tiles -- prologue only; afterwards the same bytes hold
  const uint32_t sRecv = sW;                                         //   [2 parity][CL src][32 units][RP] fp32 partials received, and
  const uint32_t sStage = sW + 2u * CL * blk;                        //   [2 parity][CL dst][32 units][RP] partials staged for the copy engine
  const uint32_t sBd = sW + region;                                  // [2 k-blocks][16 x 128 B]: own dgates_t, K order = gate*32 + unit
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + region + 2u * B_KB);
  uint64_t* w_full = bars;
  uint64_t* tmem_full = bars + 1;
  uint64_t* recv_full = bars + 2;                                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = cluster_ctarank();
  const int m0 = blockIdx.y * a.rpg;
  const int rows = min(a.rpg, a.B - m0);
  const int rows4 = (a.rpg + 3) / 4 * 4;                            // floats staged per (destination, unit)
  if (threadIdx.x == 0) cl_trace(40, 0);

  if (warp == 4 && lane == 0) {
    mbar_init(w_full, 1);
    mbar_init(tmem_full, 1);
    mbar_init(&recv_full[0], 1);
    mbar_init(&recv_full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < 2u * B_KB / 16; i += CLK_THREADS) sts128(sBd + i * 16, make_uint4(0, 0, 0, 0));
  __syncthreads();
  if (warp == 4 && elect_one()) {
    mbar_expect_tx(w_full, 4u * KB * W_TILE);
    for (int g = 0; g < 4; ++g)
      for (int kb = 0; kb < KB; ++kb)
        tma_load_2d(smem + (size_t)(g * KB + kb) * W_TILE, &tmW, kb * 64, g * H + (int)c * 32, w_full);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // W_hh[own gate columns, :]^T -> tensor memory.  M-block m: lane jl = output unit 128m + jl, K index = gate*32 + unit
    // (the order of the dgates operand), two K elements per 32-bit column: column m*64 + gate*16 + unit/2.
    mbar_wait(w_full, 0);
    const int jl = warp * 32 + lane;
    for (int m = 0; m < MB; ++m) {
      const int jj = m * 128 + jl;
      const uint32_t kbj = (uint32_t)jj >> 6, cj = (uint32_t)jj & 63u;
      for (int g = 0; g < 4; ++g) {
        const uint32_t tile = sW + ((uint32_t)g * KB + kbj) * W_TILE + (cj & 7u) * 2u;
        uint32_t r[16];
#pragma unroll
        for (int up = 0; up < 16; ++up) {
          const uint32_t u0 = 2 * up, u1 = 2 * up + 1;
          const uint32_t lo = lds_u16(tile + u0 * 128u + (((cj >> 3) ^ (u0 & 7u)) << 4));
          const uint32_t hi = lds_u16(tile + u1 * 128u + (((cj >> 3) ^ (u1 & 7u)) << 4));
          r[up] = lo | (hi << 16);
        }
        tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(m * 64 + g * 16), r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();   // every CTA is done with its weight tiles: their bytes may now receive partials; barriers are initialised
  if (threadIdx.x == 0) cl_trace(40, 1);

  // GEMM i (i = 1..T) consumes dgates of step T - i and yields dh_rec for step T - i - 1 (dh0 when i == T).
  if (warp == 4) {
    const uint64_t desc0 = make_smem_desc(0, 16, 1024);
    for (int i = 1; i <= T; ++i) {
      named_bar<160>(2);                         // dgates of step T - i are staged (writers fenced for the async proxy)
      tc_fence_after();
      if (lane == 0) cl_trace(i, 0);
      if (elect_one()) {
        mbar_expect_tx(&recv_full[(i - 1) & 1], (uint32_t)CL * blk);          // the partials this GEMM makes the cluster send here
        for (int m = 0; m < MB; ++m) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t db = desc0 + ((sBd + (uint32_t)(ks >> 2) * B_KB) >> 4);
            tc_mma_ts(tmem_base + D_COL + (uint32_t)(m * NB), tmem_base + (uint32_t)(m * 64 + ks * 8), db + 2 * (ks & 3), IDESC, ks > 0 ? 1u : 0u);
          }
        }
        tc_commit(tmem_full);
      }
      __syncwarp();
      if (lane == 0) cl_trace(i, 1);
    }
  } else {
    // cell gradient: thread = (rows 4*warp .. 4*warp+3, unit = lane)
    const int u = lane;
    const int j = (int)c * 32 + u;
    const int b0 = 4 * warp;
    float dcreg[4], dhrec[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = b0 + k;
      dcreg[k] = (b < rows && a.d_cT) ? a.d_cT[(long long)(m0 + b) * H + j] : 0.f;
      dhrec[k] = (b < rows && a.d_hT) ? a.d_hT[(long long)(m0 + b) * H + j] : 0.f;
    }
    for (int i = 0; i <= T; ++i) {
      const int t = T - 1 - i;
      // operands of step t that do not depend on the recurrence: in flight while the contraction and the exchange run
      // (nothing below touches them before the cell gradient)
      float dha[4], dhs[4], dcl[4], ig[4], fg[4], gg[4], og[4], ce[4], cp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = b0 + k;
        dha[k] = dhs[k] = dcl[k] = ig[k] = fg[k] = gg[k] = og[k] = ce[k] = cp[k] = 0.f;
        if (b < rows && t >= 0) {
          const long long bt = (long long)(m0 + b) * T + t;
          dha[k] = a.dh_attn[bt * H + j];
          if (a.dhs && t + 1 < T) dhs[k] = a.dhs[(bt + 1) * H + j];
          dcl[k] = a.dcell[bt * H + j];
          const float* ac = a.acts + bt * 4 * H + j;
          ig[k] = ac[0]; fg[k] = ac[H]; gg[k] = ac[2 * H]; og[k] = ac[3 * H];
          ce[k] = a.cells[bt * H + j];
          if (t > 0) cp[k] = a.cells[(bt - 1) * H + j];
          else if (a.c0) cp[k] = a.c0[(long long)(m0 + b) * H + j];
        }
      }
      if (i >= 1) {
        const uint32_t par = (uint32_t)(i - 1) & 1u;
        mbar_wait(tmem_full, par);
        if (threadIdx.x == 0) cl_trace(i, 2);
        tc_fence_after();
        // partial dh^T[128m + 32*warp + lane, 0..15] of this CTA's K range -> staged as [unit][RP], then one bulk copy per
        // (m, warp) block to the CTA owning those 32 units (rank 4m + warp); the owner's mbarrier counts the bytes
        for (int m = 0; m < MB; ++m) {
          uint32_t v[NB];
          tmem_ld<16>(tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL + (uint32_t)(m * NB), v);
          const uint32_t dst = sStage + (par * CL + (uint32_t)(4 * m + warp)) * blk + (uint32_t)lane * (uint32_t)RP * 4u;
#pragma unroll
          for (int q4 = 0; q4 < NB / 4; ++q4)
            if (q4 * 4 < rows4) sts128(dst + q4 * 16, make_uint4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]));
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane < MB) {
          const uint32_t owner = (uint32_t)(4 * lane + warp);
          bulk_copy_to_peer(mapa_u32(sRecv + (par * CL + c) * blk, owner), sStage + (par * CL + owner) * blk, blk,
                            mapa_u32(smem_u32(&recv_full[par]), owner));
        }
        if (threadIdx.x == 0) cl_trace(i, 5);
        // gather the CL partials of this thread's (rows, unit)
        mbar_wait(&recv_full[par], (uint32_t)((i - 1) >> 1) & 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) dhrec[k] = 0.f;
        if (b0 < rows) {
          for (int src = 0; src < CL; ++src) {
            const float4 o = lds128f(sRecv + (par * CL + (uint32_t)src) * blk + (uint32_t)((u * RP + b0) * 4));
            dhrec[0] += o.x; dhrec[1] += o.y; dhrec[2] += o.z; dhrec[3] += o.w;
          }
        }
        if (threadIdx.x == 0) cl_trace(i, 6);
      }
      if (t < 0) break;
      float d[4][4];   // [row][gate]
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = b0 + k;
        const float dh = dha[k] + dhs[k] + dhrec[k];
        const float tcv = tanhf_fast(ce[k]);
        const float dc = dcl[k] + dcreg[k] + dh * og[k] * (1.f - tcv * tcv);
        d[k][0] = dc * gg[k] * ig[k] * (1.f - ig[k]);
        d[k][1] = dc * cp[k] * fg[k] * (1.f - fg[k]);
        d[k][2] = dc * ig[k] * (1.f - gg[k] * gg[k]);
        d[k][3] = dh * tcv * og[k] * (1.f - og[k]);
        dcreg[k] = dc * fg[k];
        if (b < rows) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t e = (uint32_t)((g & 1) * 32 + u);       // element inside the k-block row
            sts_u16(sBd + (uint32_t)(g >> 1) * B_KB + (uint32_t)b * 128u + (((e >> 3) ^ ((uint32_t)b & 7u)) << 4) + (e & 7u) * 2u,
                    bf16_bits(d[k][g]));
          }
        }
      }
      fence_proxy_async_smem();                  // this thread's operand writes -> visible to the tensor core's reads
      named_bar<160>(2);
      if (threadIdx.x == 0) cl_trace(i, 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = b0 + k;
        if (b < rows) {
          const long long bt = (long long)(m0 + b) * T + t;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            a.dgates[bt * 4 * H + (long long)g * H + j] = d[k][g];
            a.dgates16[bt * 4 * H + (long long)g * H + j] = __float2bfloat16(d[k][g]);
            if (t == 0 && a.dg16_t0) a.dg16_t0[(long long)(m0 + b) * 4 * H + (long long)g * H + j] = __float2bfloat16(d[k][g]);
          }
        }
      }
      if (threadIdx.x == 0) cl_trace(i, 7);
    }
    // after the last GEMM: dhrec = dgates_0 W_hh = dh0 ; dcreg = dc0
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = b0 + k;
      if (b < rows) {
        if (a.dh0) a.dh0[(long long)(m0 + b) * H + j] = dhrec[k];
        if (a.dc0) a.dc0[(long long)(m0 + b) * H + j] = dcreg[k];
      }
    }
  }
  if (threadIdx.x == 0) cl_trace(40, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  cluster_sync_all();
  if (threadIdx.x == 0) cl_trace(40, 3);
}

// floats per (source, unit) row of the partials buffer: >= the rows sent, an odd number of 16-byte quads so that the
// 32 lanes' 16-byte reads (stride RP floats) spread over all banks
inline int bwd_rp(int rpg) {
  int q = (rpg + 3) / 4;
  if (q % 2 == 0) ++q;
  return q * 4;
}

inline size_t fwd_smem(int H) {
  const size_t KB = H / 64;
  return 4 * KB * W_TILE + 2 * (H / 32) * SLICE_BYTES + 4 * NB * 32 * 4 + 2 * SLICE_BYTES + 64 + 1024;
}
inline size_t bwd_smem(int H, int RP) {
  const size_t KB = H / 64, CL = H / 32;
  const size_t tiles = 4 * KB * W_TILE, exch = 4 * CL * 32 * (size_t)RP * 4;   // (the exchange buffers reuse the tiles' bytes)
  return (tiles > exch ? tiles : exch) + 2 * B_KB + 64 + 1024;
}

int g_lstm_cluster = 1;    // diagnostics (aa_debug_set_lstm_cluster): 0 = always take the grid-barrier kernels of lstm_seq.cu
int g_lstm_cluster_nacc = 4;

// Plain cluster launch; co-residency of a cluster is the hardware's business, clusters are independent of each other.
int launch_clk(const void* kern, int CL, int groups, size_t smem, cudaStream_t st, void** args, bool query_only, int* max_clusters) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CL, groups);
  cfg.blockDim = dim3(CLK_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (query_only) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    *max_clusters = n;
    return AA_OK;
  }
  const cudaError_t le = cudaLaunchKernelExC(&cfg, kern, args);
  if (le != cudaSuccess) {
    cudaGetLastError();
    return AA_ERR_UNSUPPORTED;     // the caller falls back to the grid-barrier kernels
  }
  count_launch();
  return AA_OK;
}

// rows per cluster / number of clusters: as many clusters as can be resident at once, at most 16 rows each
int plan_groups(const void* kern, int B, int CL, size_t smem, int* rpg, int* groups) {
  static int cached_n[3] = {-1, -1, -1};     // per cluster size 4 / 8 / 16 (smem differs little; the limit is SMs per GPC)
  const int slot = CL == 4 ? 0 : CL == 8 ? 1 : 2;
  if (cached_n[slot] < 0) {
    int n = 0;
    launch_clk(kern, CL, 1, smem, 0, nullptr, true, &n);
    cached_n[slot] = n;
  }
  const int n = cached_n[slot];
  if (n < 1) return AA_ERR_UNSUPPORTED;
  int r = ceil_div(B, n < B ? n : B);
  if (r > NB) r = NB;
  *rpg = r;
  *groups = ceil_div(B, r);
  return AA_OK;
}

template <typename K>
int prep_kernel(K kern, size_t smem, int CL, size_t* done_smem) {
  if (smem > *done_smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        (CL > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
      cudaGetLastError();
      return AA_ERR_UNSUPPORTED;
    }
    *done_smem = smem;
  }
  return AA_OK;
}

}  // namespace

int set_clk_trace_buffer(void* dev_ptr) {
  unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
  AA_CHECK_CUDA(cudaMemcpyToSymbol(g_clk_trace, &p, sizeof(p)));
  return AA_OK;
}

int set_lstm_cluster(int on, int nacc) {
  g_lstm_cluster = on ? 1 : 0;
  g_lstm_cluster_nacc = nacc == 1 ? 1 : nacc == 2 ? 2 : 4;   // (anything else: the default)
  return AA_OK;
}

bool lstm_cluster_supported(int B, int H) {
  static const int env_on = [] {
    const char* e = getenv("AA_LSTM_CLUSTER");
    return e ? atoi(e) : 1;
  }();
  return g_lstm_cluster && env_on && B >= 1 && (H == 128 || H == 256 || H == 512);
}

int launch_lstm_cluster_fwd(const LstmSeqFwd& p, cudaStream_t st) {
  const int H = p.H, CL = H / 32;
  if (!lstm_cluster_supported(p.B, H) || !p.whh16) return AA_ERR_UNSUPPORTED;
  const size_t smem = fwd_smem(H);
  const void* kern = g_lstm_cluster_nacc == 4 ? (const void*)lstm_clk_fwd_kernel<4>
                   : g_lstm_cluster_nacc == 2 ? (const void*)lstm_clk_fwd_kernel<2> : (const void*)lstm_clk_fwd_kernel<1>;
  static size_t done[3] = {0, 0, 0};
  AA_TRY(prep_kernel(kern, smem, CL, &done[g_lstm_cluster_nacc == 4 ? 2 : g_lstm_cluster_nacc == 2 ? 1 : 0]));
  int rpg = 0, groups = 0;
  AA_TRY(plan_groups(kern, p.B, CL, smem, &rpg, &groups));
  CUtensorMap tmW;
  AA_TRY(make_map(&tmW, p.whh16, 2, 4LL * H, H, H, 32));
  ClFwdArgs a{};
  a.B = p.B; a.T = p.T; a.H = H; a.rpg = rpg;
  a.xg = p.xg; a.c0 = p.c0; a.h016 = p.h016;
  a.hiddens = p.hiddens; a.cells = p.cells; a.acts = p.acts; a.hs_prev = p.hs_prev; a.hid16 = p.hid16; a.hsprev16 = p.hsprev16;
  void* args[] = {(void*)&tmW, (void*)&a};
  return launch_clk(kern, CL, groups, smem, st, args, false, nullptr);
}

int launch_lstm_cluster_bwd(const LstmSeqBwd& p, cudaStream_t st) {
  const int H = p.H, CL = H / 32;
  if (!lstm_cluster_supported(p.B, H) || !p.whh16) return AA_ERR_UNSUPPORTED;
  const void* kern = (const void*)lstm_clk_bwd_kernel;
  static size_t done = 0;
  AA_TRY(prep_kernel(kern, bwd_smem(H, bwd_rp(NB)), CL, &done));      // (attribute for the largest layout)
  int rpg = 0, groups = 0;
  AA_TRY(plan_groups(kern, p.B, CL, bwd_smem(H, bwd_rp(NB)), &rpg, &groups));
  const int RP = bwd_rp(rpg);
  const size_t smem = bwd_smem(H, RP);
  CUtensorMap tmW;
  AA_TRY(make_map(&tmW, p.whh16, 2, 4LL * H, H, H, 32));
  ClBwdArgs a{};
  a.B = p.B; a.T = p.T; a.H = H; a.rpg = rpg; a.RP = RP;
  a.dh_attn = p.dh_attn; a.dhs = p.dhs; a.dcell = p.dcell; a.d_hT = p.d_hT; a.d_cT = p.d_cT;
  a.acts = p.acts; a.cells = p.cells; a.c0 = p.c0; a.dgates = p.dgates; a.dgates16 = p.dgates16; a.dh0 = p.dh0; a.dc0 = p.dc0;
  a.dg16_t0 = p.dgates16_t0;
  void* args[] = {(void*)&tmW, (void*)&a};
  return launch_clk(kern, CL, groups, smem, st, args, false, nullptr);
}

}  // namespace aa
