// EXPERIMENT, not built into the library (kept for the record; see DESIGN.md "What did not work"): the forward recurrence
// inside one 16-CTA cluster with DSMEM exchange was correct but SLOWER than the grid-barrier kernel (8.7 us vs 5.3 us per step,
// profiles/r01_v23_lstm_cluster16_trace.txt): 16-byte st.shared::cluster scatter to 16 peers ran at ~40 cycles per warp store,
// and concentrating the cell math + strided output stores on 32 SMs instead of 128 made the epilogue 4x longer per SM.
// LSTM recurrence inside ONE thread-block cluster per batch group (bf16 tensor-core mode, H = 32 * cluster size <= 512).
//
// lstm_seq.cu synchronises its CTAs through global memory (release/acquire counter + L2 round trip, ~2 us of a 5.3 us
// forward step and ~4 us of a 14 us backward step) and re-reads the whole recurrent operand from L2 in every CTA.
// Here the CL = H/32 CTAs that together own the weight matrix form a cluster: the recurrent operand lives in shared
// memory, every CTA writes its slice of h_t (resp. dgates_t) straight into its peers' operand buffers through
// distributed shared memory, and steps are separated by hardware cluster barriers.  Batch rows are independent sequences,
// so a batch larger than the per-cluster row budget simply runs as several clusters.
//
// Forward (baseline_attention.py:167-178): CTA c owns hidden units [32c, 32c+32) = 128 gate columns (unit-major packed,
// column n = u*4 + gate), keeps that W_hh slice [128 x H] resident (B operand) and h_{t-1} [rows x H] as the A operand;
// gates = h_{t-1} W^T on tcgen05 (M = 128 batch lanes, N = 128), LSTM cell in 16 epilogue warps straight out of TMEM
// (thread = one batch row x 8 units, c_t in registers), h_t as one 16-byte bf16 chunk per thread to every peer.
//
// Backward: dh_{t-1} = dgates_t W_hh has K = 4H.  CTA c = 4*jg + kg multiplies the K-group kg (the gate columns of units
// [128kg, 128kg+128), produced by CTAs 4kg..4kg+3) with the weight block (K-group kg, output units [128jg, 128jg+128)),
// sends each 32-unit block of that partial product to the CTA owning those units (a reduce-scatter among the four CTAs
// of a j-group, fp32), which adds the four partials, applies the cell gradient and broadcasts its bf16 dgates_{t-1} slice
// to the four CTAs that consume it.  Two cluster barriers per step.
#include <cooperative_groups.h>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace aa {

namespace {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int CLT_EPI_WARPS = 16;
constexpr int CLT_THREADS = (CLT_EPI_WARPS + 1) * 32;   // warps 0-15: epilogue (TMEM lane quarter = warp % 4, unit group = warp / 4); warp 16: TMA + MMA
constexpr size_t CLT_SMEM_BUDGET = 222 * 1024;

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
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_bf16x8(bf16* p, const float* v) {
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]); pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = pk;
}

// optional per-step timeline of CTA 0 of cluster 0 (aa_debug_set_trace_buffer): 8 x uint64 globaltimer stamps per step
__device__ unsigned long long* g_cl_trace = nullptr;
__device__ __forceinline__ void cl_trace(int step, int ev) {
  if (g_cl_trace && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_cl_trace[step * 8 + ev] = t;
  }
}

struct ClFwdArgs {
  int B, T, H, rpg, rows_pad;   // rpg = batch rows per cluster, rows_pad = rpg rounded up to 8 (operand k-block stride / 128)
  const float* xg;              // [B,T,4H] input-half gate pre-activations incl. both biases (gate order i,f,g,o)
  const float* c0;              // [B,H] or null
  const bf16* h016;             // [B,H]
  float *hiddens, *cells, *acts, *hs_prev;
  bf16 *hid16, *hsprev16;
};

__global__ void __launch_bounds__(CLT_THREADS, 1) lstm_cl_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const ClFwdArgs a) {
  constexpr int N = 128;
  const int H = a.H, T = a.T, KB = H / 64;
  const uint32_t KBS = (uint32_t)a.rows_pad * 128u;       // bytes per k-block of the h operand
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                     // [KB][128 rows x 128 B]   (resident)
  uint8_t* sA = sW + (size_t)KB * 16384;                  // [KB][rows_pad x 128 B]   h_{t-1}, written by every CTA of the cluster
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)KB * KBS + 16384);   // (+16 KB: the M=128 MMA reads past rows_pad)
  uint64_t* w_full = bars;
  uint64_t* tmem_full = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = cluster_ctarank();
  const int CL = H / 32;
  const int m0 = blockIdx.y * a.rpg;
  const int rows = min(a.rpg, a.B - m0);

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // h_0 -> operand buffer (every CTA loads the whole [rows x H] block itself): 16-byte chunks, 128B swizzle
  {
    const int nchunks = rows * (H / 8);
    const uint32_t sA_u = smem_u32(sA);
    for (int i = threadIdx.x; i < nchunks; i += CLT_THREADS) {
      const int r = i / (H / 8), cc = i - r * (H / 8);
      const int kb = cc >> 3, ch = cc & 7;
      cp_async16(sA_u + (uint32_t)kb * KBS + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4), a.h016 + (long long)(m0 + r) * H + cc * 8);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == CLT_EPI_WARPS) {
    // ===== TMA (once) + MMA issuer: the warp stays converged, one elected lane issues =====
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)KB * 16384u);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + (size_t)kb * 16384, &tmW, kb * 64, (int)c * N, w_full);
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t desc0 = make_smem_desc(0, 16, 1024);
    for (int t = 0; t < T; ++t) {
      if (t > 0) cluster_wait();               // h_{t-1} delivered by every CTA
      if (lane == 0) cl_trace(t, 0);
      fence_proxy_async_smem();                // generic-proxy (cp.async / st.shared::cluster) writes -> tcgen05 reads
      tc_fence_after();
      if (elect_one()) {
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t da = desc0 + ((smem_u32(sA) + (uint32_t)kb * KBS) >> 4);
          const uint64_t db = desc0 + ((smem_u32(sW) + (uint32_t)kb * 16384u) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma<false>(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(tmem_full);
      }
      __syncwarp();
      if (lane == 0) cl_trace(t, 1);
      cluster_arrive();                        // barrier "all MMAs of step t done" (the epilogue threads arrive after seeing tmem_full)
      cluster_wait();
      if (t + 1 < T) cluster_arrive();         // barrier "h_t delivered" (nothing to deliver from this warp)
    }
  } else {
    // ===== epilogue: thread = batch row x 8 hidden units =====
    const int q = warp & 3, e = warp >> 2;
    const int rloc = q * 32 + lane;
    const int row = m0 + rloc;
    const bool valid = rloc < rows;
    const int j = (int)c * 32 + e * 8;         // first of this thread's 8 hidden units
    float creg[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) creg[u] = 0.f;
    if (valid && a.c0) {
      const float4 c0a = *reinterpret_cast<const float4*>(a.c0 + (long long)row * H + j);
      const float4 c0b = *reinterpret_cast<const float4*>(a.c0 + (long long)row * H + j + 4);
      creg[0] = c0a.x; creg[1] = c0a.y; creg[2] = c0a.z; creg[3] = c0a.w;
      creg[4] = c0b.x; creg[5] = c0b.y; creg[6] = c0b.z; creg[7] = c0b.w;
    }
    // this thread's 16-byte slot in every CTA's operand buffer: k-block c/2, chunk (c%2)*4 + e
    const uint32_t slot = smem_u32(sA) + (uint32_t)(c >> 1) * KBS + (uint32_t)rloc * 128u + (uint32_t)(((((int)c & 1) * 4 + e) ^ (rloc & 7)) << 4);
    for (int t = 0; t < T; ++t) {
      const long long bt = (long long)row * T + t;
      float4 x4[4][2];
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int v = 0; v < 2; ++v) x4[g][v] = *reinterpret_cast<const float4*>(a.xg + bt * 4 * H + (long long)g * H + j + v * 4);
      }
      if (t > 0) cluster_wait();               // (pairs with the arrive at the end of the previous step)
      mbar_wait(tmem_full, t & 1);
      if (threadIdx.x == 0) cl_trace(t, 2);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld<32>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(e * 32), r);
      tc_fence_before();
      cluster_arrive();                        // this CTA's MMAs of step t are complete and this thread has read its accumulator slice
      float hn[8], ig[8], fg[8], gg[8], og[8];
      if (valid) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const float xi[4] = {x4[0][v].x, x4[0][v].y, x4[0][v].z, x4[0][v].w};
          const float xf[4] = {x4[1][v].x, x4[1][v].y, x4[1][v].z, x4[1][v].w};
          const float xc[4] = {x4[2][v].x, x4[2][v].y, x4[2][v].z, x4[2][v].w};
          const float xo[4] = {x4[3][v].x, x4[3][v].y, x4[3][v].z, x4[3][v].w};
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const int u = v * 4 + e4;          // TMEM column = u*4 + gate
            ig[u] = sigmoidf_fast(__uint_as_float(r[u * 4 + 0]) + xi[e4]);
            fg[u] = sigmoidf_fast(__uint_as_float(r[u * 4 + 1]) + xf[e4]);
            gg[u] = tanhf_fast(__uint_as_float(r[u * 4 + 2]) + xc[e4]);
            og[u] = sigmoidf_fast(__uint_as_float(r[u * 4 + 3]) + xo[e4]);
            creg[u] = fg[u] * creg[u] + ig[u] * gg[u];
            hn[u] = og[u] * tanhf_fast(creg[u]);
          }
        }
      }
      if (threadIdx.x == 0) cl_trace(t, 3);
      cluster_wait();                          // every CTA's MMAs of step t are done: the operand buffers may be overwritten
      if (threadIdx.x == 0) cl_trace(t, 4);
      if (t + 1 < T) {
        if (valid) {
          uint4 pk;
          pk.x = pack_bf16x2(hn[0], hn[1]); pk.y = pack_bf16x2(hn[2], hn[3]); pk.z = pack_bf16x2(hn[4], hn[5]); pk.w = pack_bf16x2(hn[6], hn[7]);
          for (int d = 0; d < CL; ++d) st_cluster_v4(mapa_u32(slot, (uint32_t)d), pk);
        }
        cluster_arrive();                      // h_t delivered (release)
      }
      if (threadIdx.x == 0) cl_trace(t, 5);
      if (valid) {   // everything else is only read after the kernel
        *reinterpret_cast<float4*>(a.hiddens + bt * H + j) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        *reinterpret_cast<float4*>(a.hiddens + bt * H + j + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
        *reinterpret_cast<float4*>(a.cells + bt * H + j) = make_float4(creg[0], creg[1], creg[2], creg[3]);
        *reinterpret_cast<float4*>(a.cells + bt * H + j + 4) = make_float4(creg[4], creg[5], creg[6], creg[7]);
        float* ac = a.acts + bt * 4 * H + j;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          *reinterpret_cast<float4*>(ac + v * 4) = make_float4(ig[v * 4], ig[v * 4 + 1], ig[v * 4 + 2], ig[v * 4 + 3]);
          *reinterpret_cast<float4*>(ac + H + v * 4) = make_float4(fg[v * 4], fg[v * 4 + 1], fg[v * 4 + 2], fg[v * 4 + 3]);
          *reinterpret_cast<float4*>(ac + 2 * H + v * 4) = make_float4(gg[v * 4], gg[v * 4 + 1], gg[v * 4 + 2], gg[v * 4 + 3]);
          *reinterpret_cast<float4*>(ac + 3 * H + v * 4) = make_float4(og[v * 4], og[v * 4 + 1], og[v * 4 + 2], og[v * 4 + 3]);
        }
        st_bf16x8(a.hid16 + bt * H + j, hn);
        if (t + 1 < T) {
          *reinterpret_cast<float4*>(a.hs_prev + (bt + 1) * H + j) = make_float4(hn[0], hn[1], hn[2], hn[3]);
          *reinterpret_cast<float4*>(a.hs_prev + (bt + 1) * H + j + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
          st_bf16x8(a.hsprev16 + (bt + 1) * H + j, hn);
        }
      }
      if (threadIdx.x == 0) cl_trace(t, 6);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// forward weight packing: Wp[(c*128 + u*4 + g), k] = W_hh[g*H + c*32 + u, k]  (unit-major gate columns per CTA)
__global__ void pack_whh_cl_fwd_kernel(const float* __restrict__ w_hh, bf16* __restrict__ wp, int H) {
  const int n = blockIdx.x;
  const int c = n / 128, rem = n % 128, u = rem / 4, g = rem % 4;
  const float* src = w_hh + ((long long)g * H + c * 32 + u) * H;
  bf16* dst = wp + (long long)n * H;
  for (int k = threadIdx.x; k < H; k += blockDim.x) dst[k] = __float2bfloat16(src[k]);
}

// rows per cluster so that W slice + operand buffer fit shared memory, and the batch splits evenly
int cl_rows_per_group(int B, int H, size_t extra_per_row_bytes, int* groups) {
  const int KB = H / 64;
  const size_t fixed = (size_t)KB * 16384 + 16384 + 64 + 1024;
  const size_t per_row = (size_t)KB * 128 + extra_per_row_bytes;
  int maxr = (int)((CLT_SMEM_BUDGET - fixed) / per_row);
  maxr = maxr / 8 * 8;
  if (maxr > 128) maxr = 128;
  if (maxr < 8) { *groups = 0; return 0; }
  const int g = ceil_div(B, maxr);
  *groups = g;
  return ceil_div(B, g);
}

int launch_cluster(const void* kern, dim3 grid, int cl, size_t smem, cudaStream_t st, void** args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(CLT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  AA_CHECK_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  count_launch();
  return AA_OK;
}

}  // namespace

int set_cl_trace_buffer(void* dev_ptr) {
  unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
  AA_CHECK_CUDA(cudaMemcpyToSymbol(g_cl_trace, &p, sizeof(p)));
  return AA_OK;
}

int g_lstm_cluster = 1;   // diagnostics (aa_debug_set_lstm_cluster): 0 = always take the grid-barrier kernels of lstm_seq.cu

bool lstm_cluster_supported(int B, int H) {
  if (!g_lstm_cluster || B < 1 || H % 64 != 0 || H < 128 || H > 512) return false;
  int groups = 0;
  return cl_rows_per_group(B, H, 0, &groups) > 0 && cl_rows_per_group(B, H, 512, &groups) > 0;
}

int launch_lstm_cluster_fwd(const LstmSeqFwd& p, cudaStream_t st) {
  const int H = p.H, KB = H / 64, CL = H / 32;
  int groups = 0;
  const int rpg = cl_rows_per_group(p.B, H, 0, &groups);
  AA_REQUIRE(rpg > 0, "lstm_cluster_fwd: unsupported shape B=%d H=%d", p.B, H);
  pack_whh_cl_fwd_kernel<<<4 * H, 128, 0, st>>>(p.w_hh, p.whh_packed16, H);
  AA_CHECK_LAUNCH("pack_whh_cl_fwd");
  CUtensorMap tmW;
  AA_TRY(make_map(&tmW, p.whh_packed16, 2, 4LL * H, H, H, 128));
  ClFwdArgs a{};
  a.B = p.B; a.T = p.T; a.H = H; a.rpg = rpg; a.rows_pad = (rpg + 7) / 8 * 8;
  a.xg = p.xg; a.c0 = p.c0; a.h016 = p.h016;
  a.hiddens = p.hiddens; a.cells = p.cells; a.acts = p.acts; a.hs_prev = p.hs_prev; a.hid16 = p.hid16; a.hsprev16 = p.hsprev16;
  const size_t smem = (size_t)KB * 16384 + (size_t)KB * a.rows_pad * 128 + 16384 + 64 + 1024;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(lstm_cl_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AA_CHECK_CUDA(cudaFuncSetAttribute(lstm_cl_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_smem = smem;
  }
  void* args[] = {(void*)&tmW, (void*)&a};
  return launch_cluster((const void*)lstm_cl_fwd_kernel, dim3(CL, groups), CL, smem, st, args);
}

}  // namespace aa
