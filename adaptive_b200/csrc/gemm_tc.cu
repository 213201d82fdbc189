// tcgen05 GEMM for sm_100a: TMA-staged 128B-swizzled tiles in shared memory, one elected
// thread issuing tcgen05.mma (kind::f16 for bf16, kind::tf32 for fp32 inputs), fp32
// accumulators in tensor memory, warp-specialised (TMA producer / MMA issuer / 4 epilogue
// warps), mbarrier pipelines.  D[M,N] = A * B^T-style contraction with either operand
// K-major (reduction index contiguous) or MN-major (output index contiguous), so the forward
// (X W^T), input-gradient (dY W) and weight-gradient (dY^T X) forms all run without a
// transpose pass.
//
// Descriptor encodings follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction
// descriptor" tables (same bit layout as cute::UMMA::SmemDescriptor / InstrDescriptor).
#include <stdlib.h>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace aa {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

// 2-D tensor map over a row-major [rows, cols] array (cols contiguous, row stride ld elements):
// box = {swizzle_bytes (128 or 64) of the contiguous dim, box_rows}, matching swizzle, zero fill out of bounds.
int make_map(CUtensorMap* map, const void* base, int es, long long rows, long long cols, long long ld, int box_rows, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return AA_ERR_CUDA;
  }
  AA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tcgen05 GEMM: operand base must be 16-byte aligned");
  AA_REQUIRE((ld * es) % 16 == 0, "tcgen05 GEMM: operand row stride must be a multiple of 16 bytes (ld=%lld, es=%d)", ld, es);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld * es)};
  AA_REQUIRE(swizzle_bytes == 128 || swizzle_bytes == 64, "tcgen05 GEMM: swizzle span must be 128 or 64 bytes");
  AA_REQUIRE(box_rows >= 1 && box_rows <= 256, "tcgen05 GEMM: TMA box rows must be in [1, 256] (got %d)", box_rows);
  cuuint32_t box[2] = {(cuuint32_t)(swizzle_bytes / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim,
                   gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld es=%d)", (int)r, rows, cols, ld, es);
    return AA_ERR_CUDA;
  }
  return AA_OK;
}

namespace {

using namespace tc;

constexpr int BM = 128;
int g_tc_splitk = 1;   // diagnostics (aa_debug_set_gemm_splitk)
// warp 0: TMA, warp 1: MMA, warps 2..: epilogue (TMEM lane quarter = warp % 4).  EW epilogue warps = EW/4 "parts" per
// lane quarter, each part draining a contiguous range of the tile's 32-column chunks: with 4 warps the TMEM -> smem
// transpose -> global chain of one warp per quarter bounded the K=512 contractions (1.4 TB/s of output at 148 SMs).
constexpr int epi_warps(bool split) { return split ? 4 : 8; }   // (split stages leave no room for 8 transpose tiles)
constexpr int tc_threads(bool split, bool a_raw = false) { return 64 + 32 * epi_warps(split) + (a_raw ? 128 : 0); }   // a_raw: + 4 converter warps

struct TcEpilogue {
  int M, N, K;
  float* D32; long long ldd32;
  __nv_bfloat16* D16; long long ldd16;
  const float* Cin; long long ldcin; float beta;
  const float* bias1; const float* bias2;
  float* pmax; int* pidx; int tiles_n;      // optional per-(row, n-tile) arg-max partials of D (bias included)
  int lo_a, lo_b;                           // split mode: column offset of the lo half inside a row of A / B
  int ksplit, kb_per;                       // split-K: work unit = (tile, K range of kb_per k-blocks); partial tiles are added with red.global.add
  int kcut_n0, kcut_nkb;                    // tiles whose first column is >= kcut_n0 stop after kcut_nkb k-blocks (their B rows are zero beyond)
  int ashift_n0, ashift_cols;               // split mode: tiles whose first column is >= ashift_n0 read A ashift_cols columns further in
  __nv_bfloat16* ce_e16; long long ld_ce; float* ce_part; int ce_chunks; const long long* ce_tgt; float* ce_xt;   // PM == 3 (TcGemmArgs::ce_*)
};

// Tile geometry (bytes): every smem row is 128 B (the swizzle span).
//   K-major operand, R rows (M or N): one TMA box {128B/ES elements of K, R rows}      -> R*128 B
//   MN-major operand, R columns     : R*ES/128 TMA boxes {128B/ES elements of MN, BK rows of K} -> BK*128 B each
//
// SPLIT (fp32-accurate "3xTF32"): the operands arrive pre-split as [rows, 2*K] = [hi | lo] with hi = tf32(x),
// lo = tf32(x - hi); a stage holds the four sub-tiles A_hi, A_lo, B_hi, B_lo and every k-step issues
// lo*hi + hi*lo + hi*hi into the same fp32 accumulator (the lo*lo term, ~2^-22 relative, is dropped).  Loading
// each sub-tile once and using it in two products gives 1.5x the arithmetic intensity against L2 of a plain
// 3K-long tf32 contraction.
// PM: arg-max partials of the epilogue -- 0 = none (every training contraction: the scan code and its registers stay out of those
// instantiations), 1 = per-(row, part) maximum and its column index, 2 = maximum only, 3 = cross-entropy pieces per 32-column
// chunk (maximum, sum of exponentials, exponentials as bf16, the target's logit) instead of the logits
// A_RAW (split mode only): the A operand is a PLAIN fp32 array -- one TMA box per stage lands in the hi slot and four converter warps
// split it there into tf32 (hi, lo) sub-tiles (the split is elementwise, so the swizzled layout stays what the MMA expects).  The
// operand is read from HBM once instead of being split into a 2x larger array first: P = V W_v^T of the decode prologue, whose
// 411 MB of V took 317 us of exact-fp32 SIMT work per call.
template <int BN, int ES, int STAGES, bool A_MN, bool B_MN, bool SPLIT, int PM = 0, bool A_RAW = false>
__global__ void __launch_bounds__(tc_threads(SPLIT, A_RAW), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcEpilogue e, int tiles_m,
               int num_tiles, int num_units) {
  // Persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ... (m fastest, so neighbouring CTAs share the
  // B/weight tile in L2).  Two TMEM accumulators: the epilogue of tile i overlaps the TMA/MMA of tile i+1.
  constexpr bool TF32 = (ES == 4);
  constexpr int EPR = 128 / ES;            // elements per 128-byte smem row
  constexpr int BK = EPR;                  // reduction elements per stage (64 bf16 / 32 tf32)
  constexpr int UMMA_K = 32 / ES;          // 16 bf16 / 8 tf32
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128;   // BM*BK*ES, BN*BK*ES
  constexpr uint32_t A_STAGE = SPLIT ? 2 * A_BYTES : A_BYTES, B_STAGE = SPLIT ? 2 * B_BYTES : B_BYTES;
  static_assert(!SPLIT || (ES == 4 && !A_MN && !B_MN), "split mode: K-major fp32 operands only");
  constexpr uint32_t A_BOX = A_MN ? BK * 128 : A_BYTES;        // bytes per TMA box
  constexpr uint32_t B_BOX = B_MN ? BK * 128 : B_BYTES;
  constexpr uint32_t TCOLS = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int EW = epi_warps(SPLIT);
  constexpr int NCH = BN / 32;                               // 32-column chunks per tile
  constexpr int NPARTS = (EW / 4) < NCH ? (EW / 4) : NCH;    // parts with work (a 32-wide tile has one chunk)
  constexpr int CPP = NCH / NPARTS;                          // chunks per part

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS)
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* conv_bar = tmem_empty + 2;           // [STAGES] (A_RAW: the stage's A tile has been split)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(conv_bar + (A_RAW ? STAGES : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (e.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      if constexpr (A_RAW) mbar_init(&conv_bar[s], 4);   // one arrive per converter warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EW);   // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int tile = unit % num_tiles, sp = unit / num_tiles;
        const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
        const int kb0 = sp * e.kb_per, kb1 = min(n0 >= e.kcut_n0 ? e.kcut_nkb : nkb, kb0 + e.kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if constexpr (!(SPLIT && A_RAW)) mbar_expect_tx(&full_bar[s], A_STAGE + B_STAGE);
          uint8_t* a_dst = sA + s * A_STAGE;
          uint8_t* b_dst = sB + s * B_STAGE;
          if constexpr (SPLIT && A_RAW) {      // plain fp32 A: one box into the hi slot (the converter warps fill the lo slot)
            mbar_expect_tx(&full_bar[s], A_BYTES + B_STAGE);
            tma_load_2d(a_dst, &tmA, kb * BK, m0, &full_bar[s]);
            tma_load_2d(b_dst, &tmB, kb * BK, n0, &full_bar[s]);
            tma_load_2d(b_dst + B_BYTES, &tmB, e.lo_b + kb * BK, n0, &full_bar[s]);
            continue;
          }
          if constexpr (SPLIT) {      // hi halves at column kb*BK, lo halves at column K + kb*BK of the same arrays
            const int ash = n0 >= e.ashift_n0 ? e.ashift_cols : 0;
            tma_load_2d(a_dst, &tmA, kb * BK + ash, m0, &full_bar[s]);
            tma_load_2d(a_dst + A_BYTES, &tmA, e.lo_a + kb * BK + ash, m0, &full_bar[s]);
            tma_load_2d(b_dst, &tmB, kb * BK, n0, &full_bar[s]);
            tma_load_2d(b_dst + B_BYTES, &tmB, e.lo_b + kb * BK, n0, &full_bar[s]);
            continue;
          }
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < (int)(A_BYTES / A_BOX); ++j) tma_load_2d(a_dst + j * A_BOX, &tmA, m0 + j * EPR, kb * BK, &full_bar[s]);
          } else {
            tma_load_2d(a_dst, &tmA, kb * BK, m0, &full_bar[s]);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int j = 0; j < (int)(B_BYTES / B_BOX); ++j) tma_load_2d(b_dst + j * B_BOX, &tmB, n0 + j * EPR, kb * BK, &full_bar[s]);
          } else {
            tma_load_2d(b_dst, &tmB, kb * BK, n0, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp stays converged, one elected lane issues (see tc::elect_one) =====
    {
      // instruction descriptor: D=f32 (bits 4-5 = 1), A/B format (bf16 = 1, tf32 = 2) at bits 7-9 / 10-12,
      // A/B major at bits 15/16 (0 = K, 1 = MN), N>>3 at bits 17-22, M>>4 at bits 24-28
      constexpr uint32_t fmt = TF32 ? 2u : 1u;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0, lt = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++lt) {
        const int n0 = ((unit % num_tiles) / tiles_m) * BN;
        const int kb0 = (unit / num_tiles) * e.kb_per, kb1 = min(n0 >= e.kcut_n0 ? e.kcut_nkb : nkb, kb0 + e.kb_per);
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if constexpr (A_RAW) mbar_wait(&conv_bar[s], ph);      // (implies the stage's TMA loads: the converters waited for them)
          else mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE);
          const uint32_t b_addr = smem_u32(sB + s * B_STAGE);
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if constexpr (SPLIT) {
              const uint64_t ah = make_smem_desc(a_addr + k * 32, 16, 1024), al = make_smem_desc(a_addr + A_BYTES + k * 32, 16, 1024);
              const uint64_t bh = make_smem_desc(b_addr + k * 32, 16, 1024), bl = make_smem_desc(b_addr + B_BYTES + k * 32, 16, 1024);
              tc_mma<true>(tmem_d, al, bh, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
              tc_mma<true>(tmem_d, ah, bl, idesc, 1u);
              tc_mma<true>(tmem_d, ah, bh, idesc, 1u);
              continue;
            }
            // K-major : LBO unused (1), SBO = 8 rows * 128 B; advance 32 B per UMMA_K inside the swizzle atom
            // MN-major: LBO = bytes between 128B-wide MN chunks (one TMA box), SBO = 8 k-rows * 128 B;
            //           advance UMMA_K k-rows * 128 B
            const uint64_t da = A_MN ? make_smem_desc(a_addr + k * UMMA_K * 128, A_BOX, 1024) : make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(b_addr + k * UMMA_K * 128, B_BOX, 1024) : make_smem_desc(b_addr + k * 32, 16, 1024);
            tc_mma<TF32>(tmem_d, da, db, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);   // frees the smem stage when these MMAs have read it
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(&tmem_full[acc]);   // accumulator complete
        __syncwarp();
      }
    }
  } else if (A_RAW && warp >= 2 + EW) {
    // ===== converters (A_RAW): raw fp32 A tile -> tf32 (hi, lo) sub-tiles, in place + the lo slot =====
    if constexpr (A_RAW) {
      const int ct = threadIdx.x - (2 + EW) * 32;      // 0 .. 127
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int n0 = ((unit % num_tiles) / tiles_m) * BN;
        const int kb0 = (unit / num_tiles) * e.kb_per, kb1 = min(n0 >= e.kcut_n0 ? e.kcut_nkb : nkb, kb0 + e.kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          float4* hi4 = reinterpret_cast<float4*>(sA + s * A_STAGE);
          float4* lo4 = reinterpret_cast<float4*>(sA + s * A_STAGE + A_BYTES);
#pragma unroll
          for (int i = 0; i < (int)(A_BYTES / 16 / 128); ++i) {      // consecutive threads, consecutive 16-byte chunks: conflict-free
            const float4 x = hi4[i * 128 + ct];
            float4 h, l;
            split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
            hi4[i * 128 + ct] = h;
            lo4[i * 128 + ct] = l;
          }
          fence_proxy_async_smem();      // generic-proxy writes -> the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv_bar[s]);
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (smem transpose) -> coalesced global stores =====
    // tcgen05.ld hands each thread one accumulator ROW (32 consecutive columns); storing that directly
    // touches 32 different 128-byte lines per instruction.  Each warp transposes its 32x32 chunk through a
    // private smem tile (row stride 36 floats: conflict-free float4 writes and reads) so that one store
    // instruction writes 4 complete 128-byte row segments: lane -> (row it*4 + lane/8, columns 4*(lane%8)..+3).
    const int q = warp & 3;         // TMEM lane quarter this warp may access
    constexpr int TS = 36;
    // (16-byte aligned by offset arithmetic on the __shared__ pointer: keeps STS/LDS instead of generic ST/LD)
    uint8_t* tb0 = reinterpret_cast<uint8_t*>(tmem_slot + 1);
    float* tbuf = reinterpret_cast<float*>(tb0 + ((16u - (smem_u32(tb0) & 15u)) & 15u)) + (warp - 2) * (32 * TS);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const int part = (warp - 2) >> 2;      // which contiguous chunk range of the tile this warp drains
    int lt = 0;
    const bool ks_atomic = e.ksplit > 1;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++lt) {
      const int tile = unit % num_tiles;
      const bool first_split = unit < num_tiles;     // bias and the C input are added by the K range 0 only
      const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
      const int acc = lt & 1;
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      tc_fence_after();
      const int row0 = m0 + q * 32;
      float best = -INFINITY;     // this thread's row (row0 + lane): running arg-max over the tile's columns
      int best_i = 0x7fffffff;
      if (part >= NPARTS) {         // nothing to drain for this warp in a narrow tile: just release the accumulator
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        continue;
      }
#pragma unroll 1
      for (int c = part * CPP; c < (part + 1) * CPP; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        if (c == (part + 1) * CPP - 1) {     // this warp's share of the accumulator is read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int nb = n0 + c * 32;
        if (row0 >= e.M || nb >= e.N) continue;      // warp-uniform
        if constexpr (PM == 3) {
          // online log-softmax of this thread's row over the chunk's 32 columns: nothing of size [M, N] leaves the SM in fp32
          const long long row = row0 + lane;
          const bool full = nb + 32 <= e.N;
          float x[32];
          float m = -INFINITY;
          if (full && (reinterpret_cast<uintptr_t>(e.bias1) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias1 + nb + j));
              x[j] = __uint_as_float(r[j]) + b4.x; x[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
              x[j + 2] = __uint_as_float(r[j + 2]) + b4.z; x[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
              m = fmaxf(fmaxf(m, fmaxf(x[j], x[j + 1])), fmaxf(x[j + 2], x[j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              x[j] = (nb + j < e.N) ? __uint_as_float(r[j]) + __ldg(e.bias1 + min(nb + j, e.N - 1)) : -INFINITY;
              m = fmaxf(m, x[j]);
            }
          }
          if (row < e.M) {
            const long long t = e.ce_tgt[row];
            if (t >= nb && t < nb + 32) {
              float xt = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) xt = (t == nb + j) ? x[j] : xt;
              e.ce_xt[row] = xt;
            }
          }
          float ssum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            x[j] = __expf(x[j] - m);          // (exp(-inf) = 0 past the last column)
            ssum += x[j];
          }
          if (row < e.M) {
            float* pp = e.ce_part + (row * e.ce_chunks + (nb >> 5)) * 2;
            pp[0] = m;
            pp[1] = ssum;
          }
          // the exponentials leave through the common store path below (smem transpose -> coalesced bf16 rows of D16 = ce_e16):
          // one 16-byte store per lane and 8 columns touched half a sector in 32 different rows per instruction
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(x[j]);
        } else if constexpr (PM == 2) {    // maxima only, one per 16 columns: layout [M, ceil(N / 16)] (the caller recomputes the tiles that
                                    // matter itself: vocab_refine.cu).  The (value, index) scan below costs ~17 instructions per
                                    // element and bounded the single-pass contraction at 6 us per tile
                                    // (profiles/r01_v64_vocab_pass1_hot_lines.txt); this one ~2.5
          float mx[2] = {-INFINITY, -INFINITY};
          if (nb + 32 <= e.N && e.bias1 && (reinterpret_cast<uintptr_t>(e.bias1) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias1 + nb + j));
              mx[j >> 4] = fmaxf(fmaxf(mx[j >> 4], __uint_as_float(r[j]) + b4.x), __uint_as_float(r[j + 1]) + b4.y);
              mx[j >> 4] = fmaxf(fmaxf(mx[j >> 4], __uint_as_float(r[j + 2]) + b4.z), __uint_as_float(r[j + 3]) + b4.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < e.N) mx[j >> 4] = fmaxf(mx[j >> 4], __uint_as_float(r[j]) + (e.bias1 ? __ldg(e.bias1 + nb + j) : 0.f));
          }
          if (row0 + lane < e.M) {
            float* po = e.pmax + (long long)(row0 + lane) * e.tiles_n + (nb >> 4);
            po[0] = mx[0];
            if (nb + 16 < e.N) po[1] = mx[1];
          }
          if (!e.D32 && !e.D16) continue;
        } else if constexpr (PM == 1) {   // ascending column scan with a strict compare: the lowest index wins ties
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nb + j < e.N) {
              const float v = __uint_as_float(r[j]) + (e.bias1 ? __ldg(e.bias1 + nb + j) : 0.f) + (e.bias2 ? __ldg(e.bias2 + nb + j) : 0.f);
              if (v > best) { best = v; best_i = nb + j; }
            }
          }
          if (!e.D32 && !e.D16) continue;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(&tbuf[lane * TS + j]) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        const int n = nb + c4;
        float4 badd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (PM != 3 && (e.bias1 || e.bias2) && first_split) {      // (PM == 3: the bias is already inside the exponentials)
          float bb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            bb[j] = (n + j < e.N) ? (e.bias1 ? __ldg(e.bias1 + n + j) : 0.f) + (e.bias2 ? __ldg(e.bias2 + n + j) : 0.f) : 0.f;
          badd = make_float4(bb[0], bb[1], bb[2], bb[3]);
        }
        const bool vec32 = (n + 3 < e.N) && ((e.ldd32 & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.D32) & 15) == 0);
        const bool vecc = (n + 3 < e.N) && ((e.ldcin & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.Cin) & 15) == 0);
        const bool vec16 = (n + 3 < e.N) && ((e.ldd16 & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.D16) & 7) == 0);
        float4 cin[8];
        const bool use_cin = e.Cin && first_split && !(ks_atomic && e.Cin == e.D32);   // in-place accumulate: the adds land on C itself
        if (use_cin) {
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const long long row = row0 + i8 * 4 + rsub;
            cin[i8] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < e.M && n < e.N) {
              const float* cp = e.Cin + row * e.ldcin + n;
              if (vecc) cin[i8] = *reinterpret_cast<const float4*>(cp);
              else {
                cin[i8].x = cp[0];
                if (n + 1 < e.N) cin[i8].y = cp[1];
                if (n + 2 < e.N) cin[i8].z = cp[2];
                if (n + 3 < e.N) cin[i8].w = cp[3];
              }
            }
          }
        }
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const long long row = row0 + i8 * 4 + rsub;
          float4 v = *reinterpret_cast<const float4*>(&tbuf[(i8 * 4 + rsub) * TS + c4]);
          v.x += badd.x; v.y += badd.y; v.z += badd.z; v.w += badd.w;
          if (use_cin) { v.x += e.beta * cin[i8].x; v.y += e.beta * cin[i8].y; v.z += e.beta * cin[i8].z; v.w += e.beta * cin[i8].w; }
          if (row < e.M && n < e.N) {
            if (e.D32 && ks_atomic) {       // split-K: accumulate this K range's partial tile
              float* dp = e.D32 + row * e.ldd32 + n;
              if (vec32) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
              else {
                atomicAdd(dp, v.x);
                if (n + 1 < e.N) atomicAdd(dp + 1, v.y);
                if (n + 2 < e.N) atomicAdd(dp + 2, v.z);
                if (n + 3 < e.N) atomicAdd(dp + 3, v.w);
              }
            } else if (e.D32) {
              float* dp = e.D32 + row * e.ldd32 + n;
              if (vec32) *reinterpret_cast<float4*>(dp) = v;
              else {
                dp[0] = v.x;
                if (n + 1 < e.N) dp[1] = v.y;
                if (n + 2 < e.N) dp[2] = v.z;
                if (n + 3 < e.N) dp[3] = v.w;
              }
            }
            if (e.D16) {
              __nv_bfloat16* dp = e.D16 + row * e.ldd16 + n;
              if (vec16) {
                __nv_bfloat162 a2 = __floats2bfloat162_rn(v.x, v.y), b2 = __floats2bfloat162_rn(v.z, v.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&a2);
                pk.y = *reinterpret_cast<uint32_t*>(&b2);
                *reinterpret_cast<uint2*>(dp) = pk;
              } else {
                dp[0] = __float2bfloat16(v.x);
                if (n + 1 < e.N) dp[1] = __float2bfloat16(v.y);
                if (n + 2 < e.N) dp[2] = __float2bfloat16(v.z);
                if (n + 3 < e.N) dp[3] = __float2bfloat16(v.w);
              }
            }
          }
        }
        __syncwarp();
      }
      // partial index = first column of this part / (CPP * 32): the layout [M, ceil(N / gemm_tc_argmax_tile_n(N))]
      if constexpr (PM == 1) {
        if (row0 + lane < e.M && n0 + part * CPP * 32 < e.N) {
          const long long o = (long long)(row0 + lane) * e.tiles_n + (long long)(tile / tiles_m) * NPARTS + part;
          e.pmax[o] = best;
          e.pidx[o] = best_i;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

template <int BN, int ES, int STAGES, bool A_MN, bool B_MN, bool SPLIT = false, int PM = 0, bool A_RAW = false>
int launch_cfg(const TcGemmArgs& g, cudaStream_t st) {
  constexpr int BK = 128 / ES;
  CUtensorMap tmA, tmB;
  // K-major operand: array [R, K] (K contiguous), box {BK, tile rows}; MN-major: array [K, R] (R contiguous), box {128B of R, BK rows}
  // split operands: a row holds [hi (>= K columns) ... lo (>= K columns) ...]; the map spans what the row has left
  // from the operand's base column on (columns past it are zero-filled by TMA, never wrapped into the next row)
  const int lo_a = g.lo_a ? g.lo_a : g.K, lo_b = g.lo_b ? g.lo_b : g.K;
  const long long acols = A_RAW ? g.K : SPLIT ? (g.a_cols ? g.a_cols : (long long)lo_a + g.K) : g.K;
  const long long bcols = SPLIT ? (g.b_cols ? g.b_cols : (long long)lo_b + g.K) : g.K;
  if (A_MN) AA_TRY(make_map(&tmA, g.A, ES, g.K, g.M, g.lda, BK));
  else      AA_TRY(make_map(&tmA, g.A, ES, g.M, acols, g.lda, BM));
  if (B_MN) AA_TRY(make_map(&tmB, g.B, ES, g.K, g.N, g.ldb, BK));
  else      AA_TRY(make_map(&tmB, g.B, ES, g.N, bcols, g.ldb, BN));
  const int tiles_m = ceil_div(g.M, BM), tiles_n = ceil_div(g.N, BN);
  TcEpilogue e{};
  e.M = g.M; e.N = g.N; e.K = g.K;
  e.D32 = g.D32; e.ldd32 = g.ldd32; e.D16 = g.D16; e.ldd16 = g.ldd16;
  e.Cin = g.Cin; e.ldcin = g.ldcin; e.beta = g.beta; e.bias1 = g.bias1; e.bias2 = g.bias2;
  constexpr int EW = epi_warps(SPLIT);
  constexpr int NPARTS = (EW / 4) < (BN / 32) ? (EW / 4) : (BN / 32);
  e.kcut_n0 = 0x7fffffff; e.kcut_nkb = 0;
  e.ashift_n0 = 0x7fffffff; e.ashift_cols = 0;
  if (SPLIT && g.ashift_cols > 0) {
    AA_REQUIRE(g.ashift_n0 % BN == 0, "tcgen05 GEMM: the A-window shift must start on a tile boundary (%d %% %d)", g.ashift_n0, BN);
    e.ashift_n0 = g.ashift_n0; e.ashift_cols = g.ashift_cols;
  }
  if (g.kcut_cols > 0 && g.kcut_n0 % BN == 0 && g.kcut_n0 < g.N) {     // (only when the cut falls on a tile boundary of this configuration)
    e.kcut_n0 = g.kcut_n0;
    e.kcut_nkb = ceil_div(g.kcut_cols, 128 / ES);
  }
  if (PM == 3) { e.D16 = g.ce_e16; e.ldd16 = g.ld_ce; }      // the exponentials take the bf16 output path
  e.ce_e16 = g.ce_e16; e.ld_ce = g.ld_ce; e.ce_part = g.ce_part; e.ce_chunks = g.ce_chunks; e.ce_tgt = g.ce_tgt; e.ce_xt = g.ce_xt;
  e.pmax = g.pmax; e.pidx = g.pidx; e.tiles_n = PM == 2 ? ceil_div(g.N, 16) : ceil_div(g.N, BN / NPARTS); e.lo_a = lo_a; e.lo_b = lo_b;
  constexpr size_t smem = (size_t)STAGES * (SPLIT ? 2 : 1) * (BM * 128 + BN * 128) + (3 * STAGES + 4) * 8 + 16 + 32 + EW * 32 * 36 * 4 + 1024;
  static_assert(smem <= 227 * 1024, "tile configuration exceeds shared memory");
  auto kern = gemm_tc_kernel<BN, ES, STAGES, A_MN, B_MN, SPLIT, PM, A_RAW>;
  static bool attr_done = false;
  if (!attr_done) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int num_tiles = tiles_m * tiles_n;
  // Split-K.  One tcgen05.mma with M = 128 costs ~103 cycles for any N <= 128 (tools/mma_probe.cu), so a contraction is
  // cheapest with wide tiles; when wide tiles leave most SMs idle (small M*N, long K: the dX / dW contractions) the K range
  // is cut instead and the partial tiles are summed with vector red.global.add.  Needs a plain fp32 output.
  constexpr int BKk = 128 / ES;
  const int nkb = ceil_div(g.K, BKk);
  int ksplit = 1;
  // (a split needs the output zero-filled first: a memset node and its dependency edge cost ~2-3 us inside a graph, more than the
  //  ~0.9 us that halving an 8-k-block chain saves -- K <= 8 k-blocks is never split)
  if (!SPLIT && g.D32 && !g.D16 && !g.pmax && !g.ce_e16 && g_tc_splitk && nkb > 8) {
    const int sms = num_sms();
    while (ksplit < 16 && num_tiles * (ksplit + 1) <= sms && nkb / (ksplit + 1) >= 4) ++ksplit;
  }
  // fp32-accurate (3xTF32) mode serves decoding, whose results must not vary from run to run: at most TWO K ranges -- the sum of
  // two partials onto a zero-filled output is order-independent (0 + a + b == 0 + b + a bit for bit).  The decode step's
  // [q | r] contraction (4096 x 98 x 1024) has 32 tiles of 384 MMAs each; two ranges halve that chain.
  if (SPLIT && g.D32 && !g.D16 && !g.pmax && !g.Cin && g_tc_splitk && nkb >= 16 && num_tiles * 2 <= num_sms()) ksplit = 2;
  e.kb_per = ceil_div(nkb, ksplit);
  ksplit = ceil_div(nkb, e.kb_per);            // no empty K ranges
  e.ksplit = ksplit;
  if (ksplit > 1) e.kcut_n0 = 0x7fffffff;      // (a K range beyond the cut would be empty: its accumulator never written)
  if (ksplit > 1 && !(g.Cin == g.D32 && g.beta == 1.f))   // (in-place accumulate adds onto C itself: nothing to clear)
    AA_CHECK_CUDA(cudaMemset2DAsync(g.D32, sizeof(float) * (size_t)g.ldd32, 0, sizeof(float) * (size_t)g.N, (size_t)g.M, st));
  const int num_units = num_tiles * ksplit;
  const int grid = num_units < num_sms() ? num_units : num_sms();     // persistent: one CTA per SM
  kern<<<grid, tc_threads(SPLIT, A_RAW), smem, st>>>(tmA, tmB, e, tiles_m, num_tiles, num_units);
  AA_CHECK_LAUNCH("gemm_tc_kernel");
  return AA_OK;
}

// ---- CTA-pair contraction (tcgen05.mma.cta_group::2) ---------------------------------------------------------------------
// Two CTAs on the two SMs of one TPC (cluster of 2) share a 256 x 256 output tile: each stages ITS 128 rows of A and ITS 128
// rows of B per k-block, the leader's elected thread issues one M = 256, N = 256 MMA per k-step for both SMs (each tensor core
// reads its own A half and both B halves), and each CTA's tensor memory receives its 128 rows x 256 columns of the accumulator.
// Per SM and k-block the same 32 KB (bf16) / 64 KB (3xTF32) of operands as a 128 x 128 single-CTA tile -- for twice the output --
// and 256-wide MMAs: an MMA costs 128 cycles at N = 256 against 104 for any N <= 128 (kind::tf32 at half the flops of kind::f16:
// profiles/r02_mma_probe_tf32.jsonl), so only 256-wide tiles reach the tensor peak, and the single-CTA 3xTF32 kernel cannot hold them
// (96 KB stages).  Measured history and what bounds each user now: DESIGN.md 4 item 17, profiles/r02_pair_trace.txt.
//   barriers: full[s] lives in the LEADER (both CTAs' TMA loads complete on it: cp.async.bulk.tensor ... .cta_group::2 with the
//   peer bit of the barrier address cleared); empty[s] and tmem_full[a] exist in both CTAs and are signalled by ONE multicast
//   tcgen05.commit; tmem_empty[a] lives in the leader and counts the epilogue warps of both CTAs (remote mbarrier.arrive).
// K-major operands only; epilogue = the single-CTA kernel's (bias, C input, fp32 / bf16 stores, PM == 2 maxima per 16 columns).
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the even (leader) CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution barrier only: at the end of the kernel nothing is communicated through memory -- the peer must merely be past its last use of
// this CTA's barriers and tensor memory.  (Measured equal to the releasing form: the kernel's exit waits for its output stores either way.)
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(leader_bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc2_commit(uint64_t* bar) {      // arrives on `bar` of BOTH CTAs when the MMAs issued so far are done
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void tc2_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// ARES (A resident; plain operands, K <= ARES_KB k-blocks): the pair walks a CONTIGUOUS range of tiles in n-fastest order, so that its
// 256 rows of A change at most once or twice per launch; they stay in shared memory (one 16 KB box per k-block and CTA) and only B is
// streamed through the stage ring -- half the operand bytes per tile again (336 -> ~175 MB per vocabulary maxima pass, K = 512: 128 KB
// of A per CTA).  With it the pass's tile period is 32 MMAs x 128 cycles: tensor-bound in its steady state (profiles/r02_pair_trace.txt);
// the gain over the streamed form is small (39.3 -> 38.2 us) because set-up, the ninth round and the exit are 40 % of the launch.
constexpr int ARES_KB = 8;

#ifdef AA_PAIR_TRACE      // development build only (tools/trace_pair.py): per-tile wait / work cycles of the MMA warp and of one epilogue warp
__device__ long long g_pair_trace[148 * 16 * 8];
#endif

template <int ES, int STAGES, bool SPLIT, int PM, bool ARES = false>
__global__ void __launch_bounds__(tc_threads(SPLIT), 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcEpilogue e, int tiles_m, int num_tiles) {
  static_assert(PM == 0 || PM == 2, "pair kernel: no epilogue partials other than the maxima");
  static_assert(!ARES || !SPLIT, "the resident-A form takes plain operands");
  constexpr bool TF32 = (ES == 4);
  constexpr int BH = 128;                  // rows of A and of B this CTA stages
  constexpr int BN2 = 256;                 // columns of the pair's tile (= accumulator columns per CTA)
  constexpr int BK = 128 / ES;
  constexpr int UMMA_K = 32 / ES;
  constexpr uint32_t A_BYTES = BH * 128, B_BYTES = BH * 128;
  constexpr uint32_t A_STAGE = SPLIT ? 2 * A_BYTES : A_BYTES, B_STAGE = SPLIT ? 2 * B_BYTES : B_BYTES;
  constexpr uint32_t TCOLS = 512;          // two accumulators of 256 columns
  constexpr int EW = epi_warps(SPLIT);
  constexpr int NCH = BN2 / 32;
  constexpr int NPARTS = EW / 4;
  constexpr int CPP = NCH / NPARTS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                                     // ARES: [ARES_KB] boxes of this CTA's 128 rows, one per k-block
  uint8_t* sB = smem + (ARES ? ARES_KB * A_BYTES : STAGES * A_STAGE);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2] (the leader's are used)
  uint64_t* a_full = tmem_empty + 2;             // ARES: the resident A tile has landed (leader's) / may be overwritten (both CTAs)
  uint64_t* a_empty = a_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef AA_PAIR_TRACE
  if (threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_pair_trace[((long long)blockIdx.x * 16 + 0) * 8 + 7] = (long long)gt;
  }
#endif
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nkb = (e.K + BK - 1) / BK;
  // tiles of this pair: round robin in m-fastest order, or (ARES) a contiguous range in n-fastest order
  const int tiles_n = num_tiles / tiles_m;
  const int u_begin = ARES ? (int)((long long)pair * num_tiles / npairs) : pair;
  const int u_end = ARES ? (int)((long long)(pair + 1) * num_tiles / npairs) : num_tiles;
  const int u_step = ARES ? 1 : npairs;
  auto tile_m = [&](int u) { return ARES ? u / tiles_n : u % tiles_m; };
  auto tile_n = [&](int u) { return ARES ? u % tiles_n : u / tiles_m; };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);        // the leader's producer arms it with the bytes of both CTAs
      mbar_init(&empty_bar[s], 1);       // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EW); // the epilogue warps of both CTAs
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {      // the same warp of both CTAs, the same shared-memory offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef AA_PAIR_TRACE
  if (threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_pair_trace[((long long)blockIdx.x * 16 + 1) * 8 + 7] = (long long)gt;
  }
#endif

  if (warp == 0) {
    // ===== TMA producer (both CTAs: this CTA's halves) =====
    if (lane == 0) {
      int it = 0, cur_m = -1, a_loads = 0;
      for (int tile = u_begin; tile < u_end; tile += u_step) {
        const int ma = tile_m(tile) * 256 + (int)rank * BH, nb = tile_n(tile) * BN2 + (int)rank * BH;
        if constexpr (ARES) {
          if (tile_m(tile) != cur_m) {      // (re)load the resident rows of A -- once every MMA that reads the old ones is done
            if (a_loads > 0) mbar_wait(a_empty, (a_loads - 1) & 1);
            if (rank == 0) mbar_expect_tx(a_full, 2 * nkb * A_BYTES);
            for (int kb = 0; kb < nkb; ++kb) tma2_load_2d(sA + kb * A_BYTES, &tmA, kb * BK, ma, a_full);
            cur_m = tile_m(tile);
            ++a_loads;
          }
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * ((ARES ? 0 : A_STAGE) + B_STAGE));
          uint8_t* a_dst = sA + s * A_STAGE;
          uint8_t* b_dst = sB + s * B_STAGE;
          if constexpr (!ARES) tma2_load_2d(a_dst, &tmA, kb * BK, ma, &full_bar[s]);
          tma2_load_2d(b_dst, &tmB, kb * BK, nb, &full_bar[s]);
          if constexpr (SPLIT) {
            tma2_load_2d(a_dst + A_BYTES, &tmA, e.lo_a + kb * BK, ma, &full_bar[s]);
            tma2_load_2d(b_dst + B_BYTES, &tmB, e.lo_b + kb * BK, nb, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA only =====
    if (rank == 0) {
      constexpr uint32_t fmt = TF32 ? 2u : 1u;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int it = 0, lt = 0, cur_m = -1, a_loads = 0;
      for (int tile = u_begin; tile < u_end; tile += u_step, ++lt) {
        const int acc = lt & 1;
        if constexpr (ARES) {
          if (tile_m(tile) != cur_m) {
            mbar_wait(a_full, a_loads & 1);
            ++a_loads;
            cur_m = tile_m(tile);
          }
        }
#ifdef AA_PAIR_TRACE
        const long long tr0 = clock64();
        long long trf = 0;
#endif
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);     // both epilogues have drained this accumulator
        tc_fence_after();
#ifdef AA_PAIR_TRACE
        const long long tr1 = clock64();
#endif
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN2);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
#ifdef AA_PAIR_TRACE
          const long long tra = clock64();
#endif
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
#ifdef AA_PAIR_TRACE
          trf += clock64() - tra;
#endif
          const uint32_t a_addr = smem_u32(ARES ? sA + kb * A_BYTES : sA + s * A_STAGE);
          const uint32_t b_addr = smem_u32(sB + s * B_STAGE);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if constexpr (SPLIT) {
                const uint64_t ah = make_smem_desc(a_addr + k * 32, 16, 1024), al = make_smem_desc(a_addr + A_BYTES + k * 32, 16, 1024);
                const uint64_t bh = make_smem_desc(b_addr + k * 32, 16, 1024), bl = make_smem_desc(b_addr + B_BYTES + k * 32, 16, 1024);
                tc2_mma<true>(tmem_d, al, bh, idesc, (kb | k) != 0 ? 1u : 0u);
                tc2_mma<true>(tmem_d, ah, bl, idesc, 1u);
                tc2_mma<true>(tmem_d, ah, bh, idesc, 1u);
              } else {
                tc2_mma<TF32>(tmem_d, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(b_addr + k * 32, 16, 1024), idesc,
                              (kb | k) != 0 ? 1u : 0u);
              }
            }
            tc2_commit(&empty_bar[s]);   // frees the stage in both CTAs
          }
          __syncwarp();
        }
#ifdef AA_PAIR_TRACE
        if (lane == 0 && lt < 16) {
          long long* o = g_pair_trace + ((long long)blockIdx.x * 16 + lt) * 8;
          o[0] = tr1 - tr0; o[1] = trf; o[2] = clock64() - tr1; o[3] = tile;
        }
#endif
        if (elect_one()) {
          tc2_commit(&tmem_full[acc]);
          if constexpr (ARES) {      // last tile on these rows of A: the producers of both CTAs may overwrite them once the MMAs so far are done
            if (tile + 1 < u_end && tile_m(tile + 1) != cur_m) tc2_commit(a_empty);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue (both CTAs: this CTA's 128 rows x 256 columns) =====
    // launch_pair() guarantees: N % 16 == 0; 16-byte aligned bias, C input and output with row strides that are multiples of 4
    // floats; no bf16 output -- no scalar fall-back paths here (they made this kernel 60 KB of code, and the MMA warp's
    // instruction fetches missed: 45 -> 60 us per launch).
    // Either form keeps the NEXT chunk's accumulator read (tensor memory) and global loads (bias / C input) in flight while the
    // current chunk is worked on -- two register sets, two copies of the stage.  Done strictly in sequence a chunk cost a warp
    // ~1800 cycles (maxima) / ~2.5 us (stores with a C input from HBM): 7200 cycles of epilogue per tile against 4100 cycles of
    // MMAs in the vocabulary pass, and 20 us of exposed last-tile epilogue in the decode step's gate contraction
    // (profiles/r02_pair_ncu_details.txt).
    const int q = warp & 3;
    constexpr int TS = 36;
    uint8_t* tb0 = reinterpret_cast<uint8_t*>(tmem_slot + 1);
    float* tbuf = reinterpret_cast<float*>(tb0 + ((16u - (smem_u32(tb0) & 15u)) & 15u)) + (warp - 2) * (32 * TS);   // (PM == 0 only)
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    const int cfirst = ((warp - 2) >> 2) * CPP;
    // The C input (the decode step's static gate terms: 34 MB per step, evicted from L2 by the 411 MB of V in between) is pulled into
    // L2 two tiles ahead, one row of the tile per thread.
    auto prefetch_cin = [&](int tile) {
      if (tile >= num_tiles || !e.Cin || threadIdx.x >= 64 + BH) return;
      const int pr = tile_m(tile) * 256 + (int)rank * BH + (int)threadIdx.x - 64, pn = tile_n(tile) * BN2;
      if (pr >= e.M || pn >= e.N) return;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(e.Cin + (long long)pr * e.ldcin + pn), "r"(min(BN2, e.N - pn) * 4) : "memory");
    };
    if constexpr (PM == 0) {
      prefetch_cin(pair);
      prefetch_cin(pair + npairs);
    }
    uint32_t rr[2][32];
    float4 gg[2][8];            // PM == 2: the chunk's 32 bias values;  PM == 0: this lane's 8 float4 of the C input
    int lt = 0;
    for (int tile = u_begin; tile < u_end; tile += u_step, ++lt) {
      const int m0 = tile_m(tile) * 256 + (int)rank * BH, n0 = tile_n(tile) * BN2;
      const int acc = lt & 1;
      const int row0 = m0 + q * 32;
      if constexpr (PM == 0) prefetch_cin(tile + 2 * npairs);
      auto issue = [&](uint32_t (&r)[32], float4 (&g)[8], int c) {
        tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2 + c * 32), r);
        const int nb = n0 + c * 32;
        if constexpr (PM == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j)      // (N % 16 == 0: a chunk that starts inside N holds 16 or 32 columns)
            g[j] = (nb + 4 * j < e.N) ? __ldg(reinterpret_cast<const float4*>(e.bias1 + nb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          const int n = nb + c4;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const long long row = row0 + i8 * 4 + rsub;
            g[i8] = (e.Cin && row < e.M && n < e.N) ? __ldcg(reinterpret_cast<const float4*>(e.Cin + row * e.ldcin + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      auto stage = [&](uint32_t (&r)[32], float4 (&g)[8], uint32_t (&rn)[32], float4 (&gn)[8], int i) {
        tmem_wait_ld(r);
        if (i + 1 < CPP) {
          issue(rn, gn, cfirst + i + 1);
        } else {      // this warp's share of the accumulator is read: hand it back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
        }
        const int nb = n0 + (cfirst + i) * 32;
        if (row0 >= e.M || nb >= e.N) return;      // warp-uniform
        if constexpr (PM == 2) {
          float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = g[j >> 2];
            mx[j >> 4] = fmaxf(fmaxf(mx[j >> 4], __uint_as_float(r[j]) + b4.x), __uint_as_float(r[j + 1]) + b4.y);
            mx[j >> 4] = fmaxf(fmaxf(mx[j >> 4], __uint_as_float(r[j + 2]) + b4.z), __uint_as_float(r[j + 3]) + b4.w);
          }
          if (row0 + lane < e.M) {
            float* po = e.pmax + (long long)(row0 + lane) * e.tiles_n + (nb >> 4);
            po[0] = mx[0];
            if (nb + 16 < e.N) po[1] = mx[1];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(&tbuf[lane * TS + j]) =
                make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          __syncwarp();
          const int n = nb + c4;
          float4 badd = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n < e.N) {
            if (e.bias1) badd = __ldg(reinterpret_cast<const float4*>(e.bias1 + n));
            if (e.bias2) { const float4 b2 = __ldg(reinterpret_cast<const float4*>(e.bias2 + n)); badd.x += b2.x; badd.y += b2.y; badd.z += b2.z; badd.w += b2.w; }
          }
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const long long row = row0 + i8 * 4 + rsub;
            float4 v = *reinterpret_cast<const float4*>(&tbuf[(i8 * 4 + rsub) * TS + c4]);
            v.x += badd.x + e.beta * g[i8].x; v.y += badd.y + e.beta * g[i8].y; v.z += badd.z + e.beta * g[i8].z; v.w += badd.w + e.beta * g[i8].w;
            if (row < e.M && n < e.N) *reinterpret_cast<float4*>(e.D32 + row * e.ldd32 + n) = v;
          }
          __syncwarp();
        }
      };
      static_assert(CPP % 2 == 0, "pair epilogue: an even number of chunks per warp");
#ifdef AA_PAIR_TRACE
      const long long te0 = clock64();
#endif
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      tc_fence_after();
#ifdef AA_PAIR_TRACE
      const long long te1 = clock64();
#endif
      issue(rr[0], gg[0], cfirst);
#pragma unroll 1
      for (int i = 0; i < CPP; i += 2) {
        stage(rr[0], gg[0], rr[1], gg[1], i);
        stage(rr[1], gg[1], rr[0], gg[0], i + 1);
      }
#ifdef AA_PAIR_TRACE
      if (warp == 2 && lane == 0 && lt < 16) {
        long long* o = g_pair_trace + ((long long)blockIdx.x * 16 + lt) * 8;
        o[4] = te1 - te0; o[5] = clock64() - te1; o[6] = tile;
      }
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef AA_PAIR_TRACE
  if (threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_pair_trace[((long long)blockIdx.x * 16 + 2) * 8 + 7] = (long long)gt;
  }
#endif
  cluster_sync_relaxed();      // neither CTA frees tensor memory (or exits: its barriers are the peer's targets) before both are done
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
#ifdef AA_PAIR_TRACE
  if (threadIdx.x == 64) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_pair_trace[((long long)blockIdx.x * 16 + 3) * 8 + 7] = (long long)gt;
  }
#endif
}

// AA_GEMM_PAIR=0 keeps every contraction on the single-CTA kernel; so does a device on which no CTA pair can be made resident
// (found out at the first launch: g_pair_unavailable).
bool g_pair_unavailable = false;
int g_pair_override = -1;      // diagnostics (aa_debug_set_gemm_pair): 0 / 1 overrides the environment
bool pair_enabled() {
  static const bool on = [] { const char* e = getenv("AA_GEMM_PAIR"); return !e || e[0] != '0'; }();
  return (g_pair_override >= 0 ? g_pair_override != 0 : on) && !g_pair_unavailable;
}

// The choice must not depend on M: a row's result may not change with the number of batch mates (decode sharding).
bool pair_applies(const TcGemmArgs& g) {
  if (!pair_enabled() || g.a_mn || g.b_mn || g.a_raw || g.ashift_cols > 0 || g.kcut_cols > 0 || g.ce_e16 || g.pidx) return false;
  if (g.N < 1024 || g.N % 16 != 0) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(g.bias1) || !al16(g.bias2)) return false;
  if (g.split3)        // plain fp32-accurate contraction with an fp32 output (the decode step's gates)
    return g.elem_size == 4 && !g.pmax && g.D32 && !g.D16 && al16(g.D32) && g.ldd32 % 4 == 0 && (!g.Cin || (al16(g.Cin) && g.ldcin % 4 == 0));
  return g.elem_size == 2 && g.pmax && g.bias1 && !g.D32 && !g.D16 && !g.Cin && !g.bias2;      // the bf16 maxima pass of the vocabulary arg-max
}

template <int ES, int STAGES, bool SPLIT, int PM, bool ARES = false>
int launch_pair(const TcGemmArgs& g, cudaStream_t st) {
  constexpr int BK = 128 / ES;
  CUtensorMap tmA, tmB;
  const int lo_a = g.lo_a ? g.lo_a : g.K, lo_b = g.lo_b ? g.lo_b : g.K;
  const long long acols = SPLIT ? (g.a_cols ? g.a_cols : (long long)lo_a + g.K) : g.K;
  const long long bcols = SPLIT ? (g.b_cols ? g.b_cols : (long long)lo_b + g.K) : g.K;
  AA_TRY(make_map(&tmA, g.A, ES, g.M, acols, g.lda, 128));
  AA_TRY(make_map(&tmB, g.B, ES, g.N, bcols, g.ldb, 128));
  const int tiles_m = ceil_div(g.M, 256), tiles_n = ceil_div(g.N, 256);
  TcEpilogue e{};
  e.M = g.M; e.N = g.N; e.K = g.K;
  e.D32 = g.D32; e.ldd32 = g.ldd32; e.D16 = g.D16; e.ldd16 = g.ldd16;
  e.Cin = g.Cin; e.ldcin = g.ldcin; e.beta = g.beta; e.bias1 = g.bias1; e.bias2 = g.bias2;
  e.pmax = g.pmax; e.pidx = nullptr; e.tiles_n = ceil_div(g.N, 16); e.lo_a = lo_a; e.lo_b = lo_b;
  e.ksplit = 1; e.kb_per = ceil_div(g.K, BK);
  constexpr int EW = epi_warps(SPLIT);
  constexpr size_t smem = (ARES ? (size_t)ARES_KB * 128 * 128 + (size_t)STAGES * 128 * 128 : (size_t)STAGES * (SPLIT ? 2 : 1) * (128 * 128 + 128 * 128)) +
                          (2 * STAGES + 6) * 8 + 16 + 32 + (PM == 2 ? 0 : EW * 32 * 36 * 4) + 1024;      // (the maxima-only instantiation has no transpose tiles)
  if (ARES) AA_REQUIRE(ceil_div(g.K, BK) <= ARES_KB, "tcgen05 pair GEMM: the resident-A form holds at most %d k-blocks", ARES_KB);
  static_assert(smem <= 227 * 1024, "pair tile configuration exceeds shared memory");
  auto kern = gemm_pair_kernel<ES, STAGES, SPLIT, PM, ARES>;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(tc_threads(SPLIT));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  static int max_pairs = 0;
  if (!max_pairs) {
    AA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {      // the caller takes the single-CTA kernel, from now on
      cudaGetLastError();
      g_pair_unavailable = true;
      return AA_ERR_UNSUPPORTED;
    }
    max_pairs = n < num_sms() / 2 ? n : num_sms() / 2;
  }
  const int num_tiles = tiles_m * tiles_n;
  cfg.gridDim = dim3(2 * (num_tiles < max_pairs ? num_tiles : max_pairs));
  AA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, e, tiles_m, num_tiles));
  return AA_OK;
}

template <int BN, int ES, int STAGES>
int launch_major(const TcGemmArgs& g, cudaStream_t st) {
  if (!g.a_mn && !g.b_mn) return launch_cfg<BN, ES, STAGES, false, false>(g, st);
  if (!g.a_mn && g.b_mn) return launch_cfg<BN, ES, STAGES, false, true>(g, st);
  if (g.a_mn && !g.b_mn) return launch_cfg<BN, ES, STAGES, true, false>(g, st);
  return launch_cfg<BN, ES, STAGES, true, true>(g, st);
}

template <int ES>
int launch_es(const TcGemmArgs& g, cudaStream_t st) {
  // pick BN so that the grid covers the SMs when the problem allows it
  const long long sms = num_sms();
  const long long mt = ceil_div(g.M, BM);
  const bool can_splitk = g_tc_splitk && g.D32 && !g.D16 && !g.pmax && ceil_div(g.K, 128 / ES) >= 8;
  if (g.ce_e16) {      // cross-entropy epilogue: bf16 only (checked by the caller)
    if constexpr (ES == 2) return launch_cfg<128, 2, 5, false, false, false, 3>(g, st);
    else return AA_ERR_UNSUPPORTED;
  }
  // maxima partials of a plain (single-pass) contraction: one per 16 columns, layout [M, ceil(N / 16)] (gemm_tc_argmax_tile_n_plain)
  if (g.pmax) {
    // (bf16, wide N: 256-column tiles pull 48 KB per k-block for twice the columns of a 32 KB 128-column k-block -- the pass is
    //  bound by the operand feed; AA_ARGMAX_BN256=0 keeps the 128-column tiles)
    if constexpr (ES == 2) {
      static const bool wide = [] { const char* e = getenv("AA_ARGMAX_BN256"); return !e || e[0] != '0'; }();
      if (wide && g.N >= 2048) return launch_cfg<256, 2, 3, false, false, false, 2>(g, st);
    }
    return launch_cfg<128, ES, 5, false, false, false, 2>(g, st);
  }
  // bf16, 256-column tiles: one A k-block feeds twice the columns (48 KB per k-block for 2x the flops of a 128x128 tile's 32 KB)
  // and a 128x256x16 MMA costs 128 cycles against 2 x 103.  Measured (tools/bench_gemm.py, r01_v58): the config-5 vocabulary
  // projection (4608 x 20000 x 1024) 195 -> 168 us (1.12 PFLOP/s), but every K <= 512 contraction of config 2 a few per cent
  // SLOWER (only three 48 KB stages fit next to the epilogue's transpose tiles, and a short K loop exposes the shallower
  // pipeline): taken only for long-K problems with at least two full waves of the wide tiles.
  if constexpr (ES == 2) {
    static const bool bn256 = [] { const char* e = getenv("AA_GEMM_BN256"); return !e || e[0] != '0'; }();
    if (bn256 && g.N >= 256 && !g.pmax && mt * ceil_div(g.N, 256) >= 2 * sms && ceil_div(g.K, 64) >= 16) return launch_major<256, ES, 3>(g, st);
  }
  // widest tile the problem fills: the per-MMA cost does not depend on N, and split-K covers the SMs when it applies
  if (g.N > 64 && (can_splitk || mt * ceil_div(g.N, 128) >= sms || g.N > 2048)) return launch_major<128, ES, 5>(g, st);
  if (g.N > 32 && (can_splitk || mt * ceil_div(g.N, 64) >= sms || g.N > 512 || (g.b_mn && ES == 2))) return launch_major<64, ES, 6>(g, st);
  if (g.b_mn && ES == 2) return launch_major<64, ES, 6>(g, st);   // (an MN-major bf16 B tile needs >= 64 columns: one 128-byte swizzle row)
  return launch_major<32, ES, 6>(g, st);
}

}  // namespace

int set_gemm_pair(int on) {
  g_pair_override = on < 0 ? -1 : (on ? 1 : 0);
  return AA_OK;
}

int set_gemm_splitk(int on) {
  g_tc_splitk = on ? 1 : 0;
  return AA_OK;
}

int gemm_tc_argmax_tile_n(int N) { return N > 64 ? 128 : 64; }
int gemm_tc_argmax_tile_n_plain(int) { return 16; }

int launch_gemm_tc(const TcGemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return AA_OK;
  AA_REQUIRE(g.K > 0 && g.A && g.B && (g.D32 || g.D16 || g.pmax || g.ce_e16), "tcgen05 GEMM: bad arguments");
  AA_REQUIRE(!g.ce_e16 || (g.elem_size == 2 && !g.split3 && !g.a_mn && !g.b_mn && !g.D32 && !g.D16 && !g.pmax && !g.Cin && g.bias1 && !g.bias2 &&
                           g.ce_part && g.ce_tgt && g.ce_xt && g.ld_ce % 8 == 0 && g.ce_chunks == (g.N + 31) / 32 &&
                           (reinterpret_cast<uintptr_t>(g.ce_e16) & 15) == 0),
             "tcgen05 GEMM: the cross-entropy epilogue needs K-major bf16 operands, one bias, no other output and 16-byte aligned bf16 rows");
  AA_REQUIRE(g.elem_size == 2 || g.elem_size == 4, "tcgen05 GEMM: element size must be 2 (bf16) or 4 (tf32)");
  AA_REQUIRE(!g.pmax || (!g.Cin && (g.split3 || (!g.a_mn && !g.b_mn))), "tcgen05 GEMM: arg-max partials need K-major operands and no C input");
  AA_REQUIRE(!g.pmax || (g.split3 ? g.pidx != nullptr : (g.pidx == nullptr && !g.bias2)),
             "tcgen05 GEMM: split mode keeps (maximum, index) partials and needs pidx; a plain contraction keeps maxima only (pidx == NULL, one bias)");
  if (g.split3) {
    AA_REQUIRE(g.elem_size == 4 && !g.a_mn && !g.b_mn, "tcgen05 GEMM: split (3xTF32) mode needs K-major fp32 operands");
    AA_REQUIRE(g.K % 32 == 0, "tcgen05 GEMM: split mode needs K (per half) padded to a multiple of 32 (got %d)", g.K);
    if (pair_applies(g)) {
      const int rc = launch_pair<4, 3, true, 0>(g, st);
      if (rc != AA_ERR_UNSUPPORTED) return rc;
    }
    // the arg-max partial layout [M, ceil(N / tile_n)] is part of the contract: tile_n = gemm_tc_argmax_tile_n(N)
    // (256-column tiles -- 128 cycles per MMA for twice the columns -- were measured SLOWER here: only two 96 KB stages fit
    //  and the TMA feed, ~50 B/clk per SM, becomes the bound: vocabulary GEMM 224 us against 205 us, r01_v33)
    if (g.pmax) return g.N > 64 ? launch_cfg<128, 4, 3, false, false, true, 1>(g, st) : launch_cfg<64, 4, 4, false, false, true, 1>(g, st);
    if (g.a_raw) {      // plain fp32 A split on the fly (narrow outputs only: the decode prologue's P = V W_v^T, N = 49)
      AA_REQUIRE(!g.pmax && !g.ashift_cols && g.N <= 64, "tcgen05 GEMM: the on-the-fly split of A serves outputs of at most 64 columns");
      return launch_cfg<64, 4, 4, false, false, true, 0, true>(g, st);
    }
    if (g.ashift_cols > 0) return launch_cfg<64, 4, 4, false, false, true>(g, st);
    if (g.N > 64) return launch_cfg<128, 4, 3, false, false, true>(g, st);
    return launch_cfg<64, 4, 4, false, false, true>(g, st);
  }
  if (g.elem_size == 4 && (g.a_mn || g.b_mn)) {
    // 32-bit MN-major operands need the SWIZZLE_128B_BASE32B layout; only the forward (K-major) form is built
    set_error("tcgen05 GEMM: tf32 operands must be K-major");
    return AA_ERR_UNSUPPORTED;
  }
  if (pair_applies(g)) {
    static const bool ares = [] { const char* e = getenv("AA_GEMM_PAIR_ARES"); return !e || e[0] != '0'; }();
    const int rc = (ares && g.K <= ARES_KB * 64) ? launch_pair<2, 6, false, 2, true>(g, st) : launch_pair<2, 7, false, 2>(g, st);
    if (rc != AA_ERR_UNSUPPORTED) return rc;
  }
  return g.elem_size == 2 ? launch_es<2>(g, st) : launch_es<4>(g, st);
}

}  // namespace aa

#ifdef AA_PAIR_TRACE
extern "C" int aa_debug_pair_trace(long long* out, int n, int reset) {
  if (cudaMemcpyFromSymbol(out, aa::g_pair_trace, sizeof(long long) * (size_t)n) != cudaSuccess) return 2;
  if (reset) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, aa::g_pair_trace) != cudaSuccess || cudaMemset(p, 0, sizeof(aa::g_pair_trace)) != cudaSuccess) return 2;
  }
  return 0;
}
#endif
