// In-place sum all-reduce of a gradient bucket over the GPUs of one NVSwitch box, written against PEER-MAPPED memory
// instead of NCCL: the exchange step of data-parallel training (the reference's own scheme is
// nn.DataParallel's reduce_add_coalesced, baseline_attention.py:184-187).
//
// Why not NCCL here: every captured ncclAllReduce of the training step costs ~45 us whatever its size
// (profiles/r01_v47_timeline_n2.txt) and its ring kernels hold SMs the backward's cluster kernels need; the three small
// buckets of a 0.37 ms step are latency problems, not bandwidth problems.
//
// Every rank's bucket lives at the same offset of a symmetric allocation (one cuMem allocation per rank, mapped into every
// peer; the host side gets the mappings from torch.distributed._symmetric_memory -- plumbing).  Two kernels:
//   * ar_multimem_kernel (NVLS: a multicast mapping of the allocation exists): rank r owns slice r of the bucket; ONE
//     multimem.ld_reduce per 16 bytes makes the SWITCH add the W ranks' values, ONE multimem.st broadcasts the sum into all W
//     buffers.  Per GPU n/W loads + n/W stores instead of a ring's 2(W-1)/W n.
//   * ar_twoshot_kernel (no multicast): rank r sums slice r from the W peer mappings in fixed rank order (deterministic),
//     writes it to its own buffer, then pulls the other W-1 reduced slices from their owners.
// Cross-GPU ordering: per-(channel, CTA, source rank) flags inside the symmetric allocation, set by the sender with a
// system-scope release CAS 0 -> 1 and consumed by the receiver with an acquire CAS 1 -> 0: reusable without epochs, so a
// kernel can be captured into a CUDA graph and replayed.  CTA b of every rank only ever talks to CTA b of its peers; the grid
// (<= AR_MAX_BLOCKS CTAs, far below the SM count) is co-resident by construction.
#include "kernels.cuh"

namespace aa {
namespace {

constexpr int AR_THREADS = 512;

__device__ __forceinline__ unsigned cas_release_sys(unsigned* p, unsigned cmp, unsigned val) {
  unsigned old;
  asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ unsigned cas_acquire_sys(unsigned* p, unsigned cmp, unsigned val) {
  unsigned old;
  asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}

// CTA-level barrier with the same CTA of every peer.  flags_of[p] = flag words of rank p (peer mapping); word (base + src)
// of rank dst is written by src and consumed by dst.
__device__ __forceinline__ void ar_barrier(unsigned* const* __restrict__ flags_of, int base, int rank, int world) {
  __syncthreads();
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    const int peer = threadIdx.x;
    __threadfence_system();
    long long t0 = 0;
    while (cas_release_sys(flags_of[peer] + base + rank, 0u, 1u) != 0u) {      // tell `peer` that this CTA has arrived
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
    t0 = 0;
    while (cas_acquire_sys(flags_of[rank] + base + peer, 1u, 0u) != 1u) {      // wait for `peer`'s arrival, consume it
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
    __threadfence_system();
  }
  __syncthreads();
}

struct ArArgs {
  float* bufs[AA_AR_MAX_WORLD];          // peer mappings of the symmetric allocation (bufs[rank] = the local one)
  unsigned* flags[AA_AR_MAX_WORLD];      // flag words inside each mapping
  float* mc;                             // multicast mapping (NVLS) or null
  long long off, n;                      // bucket = elements [off, off + n), n % 4 == 0, off % 4 == 0
  int rank, world, flag_base;            // flag_base = channel * AR_MAX_BLOCKS * world
};

__device__ __forceinline__ float4 mm_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(AR_THREADS) ar_multimem_kernel(const ArArgs a) {
  const int W = a.world, r = a.rank;
  const int fb = a.flag_base + blockIdx.x * W;
  ar_barrier(a.flags, fb, r, W);                      // every rank's bucket is final (each kernel runs after its own producer)
  const long long n4 = a.n / 4;
  const long long per = (n4 + W - 1) / W;             // float4 groups per rank slice
  const long long lo = per * r, hi = min(n4, lo + per);
  float* mc = a.mc + a.off;
  const long long stride = (long long)gridDim.x * AR_THREADS;
  long long i = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x;
  for (; i + 3 * stride < hi; i += 4 * stride) {      // four independent 16-byte reductions in flight per thread
    const float4 v0 = mm_ld_reduce(mc + 4 * i), v1 = mm_ld_reduce(mc + 4 * (i + stride));
    const float4 v2 = mm_ld_reduce(mc + 4 * (i + 2 * stride)), v3 = mm_ld_reduce(mc + 4 * (i + 3 * stride));
    mm_st(mc + 4 * i, v0); mm_st(mc + 4 * (i + stride), v1); mm_st(mc + 4 * (i + 2 * stride), v2); mm_st(mc + 4 * (i + 3 * stride), v3);
  }
  for (; i < hi; i += stride) mm_st(mc + 4 * i, mm_ld_reduce(mc + 4 * i));
  ar_barrier(a.flags, fb, r, W);                      // all ranks' broadcast stores have landed in every buffer
}

__global__ void __launch_bounds__(AR_THREADS) ar_twoshot_kernel(const ArArgs a) {
  const int W = a.world, r = a.rank;
  const int fb = a.flag_base + blockIdx.x * W;
  ar_barrier(a.flags, fb, r, W);
  const long long n4 = a.n / 4;
  const long long per = (n4 + W - 1) / W;
  const long long lo = per * r, hi = min(n4, lo + per);
  const long long stride = (long long)gridDim.x * AR_THREADS;
  float* mine = a.bufs[r] + a.off;
  // reduce-scatter: slice r = sum over ranks in rank order (the same order on every replay: deterministic)
  for (long long i = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i < hi; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int p = 0; p < W; ++p) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(a.bufs[p] + a.off) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(mine)[i] = acc;
  }
  ar_barrier(a.flags, fb, r, W);                      // every slice is reduced at its owner
  // all-gather: pull the other ranks' reduced slices
  for (int s = 1; s < W; ++s) {
    const int p = (r + s) % W;
    const long long plo = per * p, phi = min(n4, plo + per);
    const float4* src = reinterpret_cast<const float4*>(a.bufs[p] + a.off);
    for (long long i = plo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i < phi; i += stride)
      reinterpret_cast<float4*>(mine)[i] = __ldcg(src + i);
  }
  ar_barrier(a.flags, fb, r, W);                      // nobody still reads a slice its owner is about to overwrite
}

}  // namespace
}  // namespace aa

using namespace aa;

extern "C" {

size_t aa_allreduce_flag_bytes(void) { return sizeof(unsigned) * AA_AR_CHANNELS * AA_AR_MAX_BLOCKS * AA_AR_MAX_WORLD; }

int aa_allreduce_sum_f32(void* const* peer_bufs, void* multicast_buf, long long flag_offset_bytes, int rank, int world,
                         long long offset_elems, long long n_elems, int channel, int max_blocks, void* stream) {
  AA_REQUIRE(peer_bufs && world >= 2 && world <= AA_AR_MAX_WORLD && rank >= 0 && rank < world, "aa_allreduce_sum_f32: bad rank / world (%d / %d)", rank, world);
  AA_REQUIRE(channel >= 0 && channel < AA_AR_CHANNELS, "aa_allreduce_sum_f32: channel must be in [0, %d)", AA_AR_CHANNELS);
  AA_REQUIRE(offset_elems % 4 == 0 && n_elems % 4 == 0 && n_elems >= 0 && flag_offset_bytes % 16 == 0,
             "aa_allreduce_sum_f32: offset and length must be multiples of 4 elements");
  if (n_elems == 0) return AA_OK;
  ArArgs a{};
  for (int p = 0; p < world; ++p) {
    AA_REQUIRE(peer_bufs[p], "aa_allreduce_sum_f32: null peer mapping %d", p);
    a.bufs[p] = static_cast<float*>(peer_bufs[p]);
    a.flags[p] = reinterpret_cast<unsigned*>(static_cast<char*>(peer_bufs[p]) + flag_offset_bytes);
  }
  a.mc = static_cast<float*>(multicast_buf);
  a.off = offset_elems; a.n = n_elems; a.rank = rank; a.world = world;
  a.flag_base = channel * AA_AR_MAX_BLOCKS * AA_AR_MAX_WORLD;
  // enough CTAs to keep the NVLink ports busy, few enough to leave the SMs to the backward: one CTA per 64 KB of this rank's slice
  const long long slice_bytes = (n_elems * 4 + world - 1) / world;
  int blocks = (int)((slice_bytes + 65535) / 65536);
  const int cap = max_blocks > 0 && max_blocks < AA_AR_MAX_BLOCKS ? max_blocks : AA_AR_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.mc) ar_multimem_kernel<<<blocks, AR_THREADS, 0, st>>>(a);
  else ar_twoshot_kernel<<<blocks, AR_THREADS, 0, st>>>(a);
  AA_CHECK_LAUNCH("allreduce");
  return AA_OK;
}

}  // extern "C"
