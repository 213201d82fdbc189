// In-place sum all-reduce of a gradient bucket over the GPUs of one NVSwitch box, written against PEER-MAPPED memory
// instead of NCCL: the exchange step of data-parallel training (the reference's own scheme is
// nn.DataParallel's reduce_add_coalesced, baseline_attention.py:184-187).
//
// Why not NCCL here: every captured ncclAllReduce of the training step costs ~45 us whatever its size
// (profiles/r01_v47_timeline_n2.txt) and its ring kernels hold SMs the backward's cluster kernels need; the buckets of a
// 0.37 ms step are latency problems first.
//
// Every rank's bucket lives at the same offset of a symmetric allocation (one cuMem allocation per rank, mapped into every
// peer; the host side gets the mappings from torch.distributed._symmetric_memory -- plumbing).  Two kernels:
//   * ar_multimem_kernel (NVLS: a multicast mapping of the allocation exists): rank r owns slice r of the bucket; ONE
//     multimem.ld_reduce per 16 bytes makes the SWITCH add the W ranks' values, ONE multimem.st broadcasts the sum into all W
//     buffers.  Per GPU n/W loads + n/W stores instead of a ring's 2(W-1)/W n.
//   * ar_twoshot_kernel (no multicast): rank r sums slice r from the W peer mappings in fixed rank order (deterministic),
//     writes it to its own buffer, then pulls the other W-1 reduced slices from their owners.
// Cross-GPU ordering: one arrival counter per (channel, CTA) inside every rank's copy of the allocation.  A barrier = every
// rank adds 1 to that counter on ALL ranks -- one multimem.red through the multicast mapping (the switch fans it out) or W
// release-adds through the peer mappings -- and spins on its LOCAL copy until it has seen (barriers so far) * W arrivals.
// The number of barriers a CTA has been through lives next to the counter (device memory, touched only by that CTA), so the
// kernels can be captured into a CUDA graph and replayed: no host-side epoch.  CTA b of every rank only ever talks to CTA b of
// its peers; the grid (<= AA_AR_MAX_BLOCKS CTAs, far below the SM count) is co-resident by construction.
#include <stdlib.h>

#include <algorithm>

#include "kernels.cuh"

namespace aa {
namespace {

constexpr int AR_THREADS = 512;      // upper bound; the launch picks blockDim.x (AA_AR_THREADS)
constexpr int AR_UNROLL = 8;
constexpr int AR_SLOTS = AA_AR_CHANNELS * AA_AR_MAX_BLOCKS;      // flag words: [AR_SLOTS] arrival counters, then [AR_SLOTS] local barrier counts

struct ArArgs {
  float* bufs[AA_AR_MAX_WORLD];          // peer mappings of the symmetric allocation (bufs[rank] = the local one)
  unsigned* flags[AA_AR_MAX_WORLD];      // flag words inside each mapping
  unsigned* mc_flags;                    // the same words through the multicast mapping (or null)
  float* mc;                             // multicast mapping (NVLS) or null
  long long off, n;                      // bucket = elements [off, off + n), n % 4 == 0, off % 4 == 0
  int rank, world, slot_base;            // slot_base = channel * AA_AR_MAX_BLOCKS
};

__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// CTA-level barrier with the same CTA of every peer (see the header comment).
__device__ __forceinline__ void ar_barrier(const ArArgs& a) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const int slot = a.slot_base + blockIdx.x;
    unsigned* mine = a.flags[a.rank];
    const unsigned e = mine[AR_SLOTS + slot] + 1u;      // barriers this CTA has entered, this one included (local word)
    mine[AR_SLOTS + slot] = e;
    __threadfence_system();                             // this CTA's stores (seen through the __syncthreads above) before the arrival
    if (a.mc_flags) {
      asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(a.mc_flags + slot), "r"(1u) : "memory");
    } else {
      for (int p = 0; p < a.world; ++p)
        asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(a.flags[p] + slot), "r"(1u) : "memory");
    }
    const unsigned target = e * (unsigned)a.world;
    long long t0 = 0;
    // relaxed polls with a pause in between, ONE acquire fence at the end: an acquire per poll is a system-scope fence per
    // iteration on an SM that the backward's kernels share with this CTA (measured: the BPTT kernel next to 32 such pollers
    // took 112 us instead of 58, profiles/r02_timeline_n8_v1.txt)
    while ((int)(ld_relaxed_sys(mine + slot) - target) < 0) {
      __nanosleep(100);
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();      // a peer that never arrives must fault, not hang the GPU
    }
    __threadfence_system();
  }
  __syncthreads();
}

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
  ar_barrier(a);                                      // every rank's bucket is final (each kernel runs after its own producer)
  const long long n4 = a.n / 4;
  const long long per = (n4 + W - 1) / W;             // float4 groups per rank slice
  const long long lo = per * r, hi = min(n4, lo + per);
  float* mc = a.mc + a.off;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (AR_UNROLL - 1) * stride < hi; i += AR_UNROLL * stride) {      // AR_UNROLL independent 16-byte reductions in flight per thread
    float4 v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) v[u] = mm_ld_reduce(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) mm_st(mc + 4 * (i + u * stride), v[u]);
  }
  for (; i < hi; i += stride) mm_st(mc + 4 * i, mm_ld_reduce(mc + 4 * i));
  ar_barrier(a);                                      // all ranks' broadcast stores have landed in every buffer
}

__global__ void __launch_bounds__(AR_THREADS) ar_twoshot_kernel(const ArArgs a) {
  const int W = a.world, r = a.rank;
  ar_barrier(a);
  const long long n4 = a.n / 4;
  const long long per = (n4 + W - 1) / W;
  const long long lo = per * r, hi = min(n4, lo + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  float* mine = a.bufs[r] + a.off;
  // reduce-scatter: slice r = sum over ranks in rank order (the same order on every replay: deterministic)
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    float4 v[AA_AR_MAX_WORLD];
#pragma unroll
    for (int p = 0; p < AA_AR_MAX_WORLD; ++p)
      if (p < W) v[p] = __ldcg(reinterpret_cast<const float4*>(a.bufs[p] + a.off) + i);
    float4 acc = v[0];
#pragma unroll
    for (int p = 1; p < AA_AR_MAX_WORLD; ++p)
      if (p < W) { acc.x += v[p].x; acc.y += v[p].y; acc.z += v[p].z; acc.w += v[p].w; }
    reinterpret_cast<float4*>(mine)[i] = acc;
  }
  ar_barrier(a);                                      // every slice is reduced at its owner
  // all-gather: pull the other ranks' reduced slices
  for (int s = 1; s < W; ++s) {
    const int p = (r + s) % W;
    const long long plo = per * p, phi = min(n4, plo + per);
    const float4* src = reinterpret_cast<const float4*>(a.bufs[p] + a.off);
    long long i = plo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < phi; i += 4 * stride) {
      const float4 v0 = __ldcg(src + i), v1 = __ldcg(src + i + stride), v2 = __ldcg(src + i + 2 * stride), v3 = __ldcg(src + i + 3 * stride);
      reinterpret_cast<float4*>(mine)[i] = v0;
      reinterpret_cast<float4*>(mine)[i + stride] = v1;
      reinterpret_cast<float4*>(mine)[i + 2 * stride] = v2;
      reinterpret_cast<float4*>(mine)[i + 3 * stride] = v3;
    }
    for (; i < phi; i += stride) reinterpret_cast<float4*>(mine)[i] = __ldcg(src + i);
  }
  ar_barrier(a);                                      // nobody still reads a slice its owner is about to overwrite
}

// ---- bf16 exchange: the bucket travels as bf16 (half the NVLink bytes), the switch accumulates in fp32 ----
struct uint4x { unsigned x, y, z, w; };
__device__ __forceinline__ uint4x mm_ld_reduce_bf16(const void* p) {
  uint4x v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_bf16(void* p, const uint4x& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// a.off / a.n count 16-byte groups (8 bf16) of the staging region here; a.mc / a.bufs point at the allocation's base
__global__ void __launch_bounds__(AR_THREADS) ar_multimem_bf16_kernel(const ArArgs a) {
  const int W = a.world, r = a.rank;
  ar_barrier(a);
  const long long n4 = a.n;
  const long long per = (n4 + W - 1) / W;
  const long long lo = per * r, hi = min(n4, lo + per);
  uint4* mc = reinterpret_cast<uint4*>(a.mc) + a.off;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (AR_UNROLL - 1) * stride < hi; i += AR_UNROLL * stride) {
    uint4x v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) v[u] = mm_ld_reduce_bf16(mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) mm_st_bf16(mc + i + u * stride, v[u]);
  }
  for (; i < hi; i += stride) mm_st_bf16(mc + i, mm_ld_reduce_bf16(mc + i));
  ar_barrier(a);
}

__device__ __forceinline__ void bf16x8_accum(float (&acc)[8], const uint4& v) {
  const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[2 * j] += __uint_as_float(w[j] << 16);
    acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(AR_THREADS) ar_twoshot_bf16_kernel(const ArArgs a) {
  const int W = a.world, r = a.rank;
  ar_barrier(a);
  const long long n4 = a.n;
  const long long per = (n4 + W - 1) / W;
  const long long lo = per * r, hi = min(n4, lo + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  uint4* mine = reinterpret_cast<uint4*>(a.bufs[r]) + a.off;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < W; ++p) bf16x8_accum(acc, __ldcg(reinterpret_cast<const uint4*>(a.bufs[p]) + a.off + i));
    uint4 o;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(acc[0], acc[1]); o.x = *reinterpret_cast<unsigned*>(&t);
    t = __floats2bfloat162_rn(acc[2], acc[3]); o.y = *reinterpret_cast<unsigned*>(&t);
    t = __floats2bfloat162_rn(acc[4], acc[5]); o.z = *reinterpret_cast<unsigned*>(&t);
    t = __floats2bfloat162_rn(acc[6], acc[7]); o.w = *reinterpret_cast<unsigned*>(&t);
    mine[i] = o;
  }
  ar_barrier(a);
  for (int s = 1; s < W; ++s) {
    const int p = (r + s) % W;
    const long long plo = per * p, phi = min(n4, plo + per);
    const uint4* src = reinterpret_cast<const uint4*>(a.bufs[p]) + a.off;
    for (long long i = plo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < phi; i += stride) mine[i] = __ldcg(src + i);
  }
  ar_barrier(a);
}

// fp32 -> bf16 (round to nearest even) and back, 8 elements per thread and iteration
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, uint4* __restrict__ dst, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(a.x, a.y); o.x = *reinterpret_cast<unsigned*>(&t);
    t = __floats2bfloat162_rn(a.z, a.w); o.y = *reinterpret_cast<unsigned*>(&t);
    t = __floats2bfloat162_rn(b.x, b.y); o.z = *reinterpret_cast<unsigned*>(&t);
    t = __floats2bfloat162_rn(b.z, b.w); o.w = *reinterpret_cast<unsigned*>(&t);
    dst[i] = o;
  }
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const uint4* __restrict__ src, float* __restrict__ dst, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldcg(src + i);
    float4 a, b;
    a.x = __uint_as_float(v.x << 16); a.y = __uint_as_float(v.x & 0xffff0000u);
    a.z = __uint_as_float(v.y << 16); a.w = __uint_as_float(v.y & 0xffff0000u);
    b.x = __uint_as_float(v.z << 16); b.y = __uint_as_float(v.z & 0xffff0000u);
    b.z = __uint_as_float(v.w << 16); b.w = __uint_as_float(v.w & 0xffff0000u);
    reinterpret_cast<float4*>(dst)[2 * i] = a;
    reinterpret_cast<float4*>(dst)[2 * i + 1] = b;
  }
}

int ar_threads() {
  static const int threads = [] {
    const char* e = getenv("AA_AR_THREADS");
    const int t = e ? atoi(e) : 512;
    return (t >= 32 && t <= AR_THREADS && t % 32 == 0) ? t : 512;
  }();
  return threads;
}

}  // namespace
}  // namespace aa

using namespace aa;

extern "C" {

size_t aa_allreduce_flag_bytes(void) { return sizeof(unsigned) * 2 * AR_SLOTS; }

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
  a.mc_flags = multicast_buf ? reinterpret_cast<unsigned*>(static_cast<char*>(multicast_buf) + flag_offset_bytes) : nullptr;
  a.off = offset_elems; a.n = n_elems; a.rank = rank; a.world = world;
  a.slot_base = channel * AA_AR_MAX_BLOCKS;
  // enough CTAs to keep the NVLink ports busy (bytes in flight = CTAs x 512 threads x AR_UNROLL x 16 B against ~3 us of
  // latency), few enough to leave the SMs to the backward: one CTA per 64 KB of this rank's slice
  const long long slice_bytes = (n_elems * 4 + world - 1) / world;
  int blocks = (int)((slice_bytes + 65535) / 65536);
  const int cap = max_blocks > 0 && max_blocks < AA_AR_MAX_BLOCKS ? max_blocks : AA_AR_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = ar_threads();
  if (a.mc) ar_multimem_kernel<<<blocks, threads, 0, st>>>(a);
  else ar_twoshot_kernel<<<blocks, threads, 0, st>>>(a);
  AA_CHECK_LAUNCH("allreduce");
  return AA_OK;
}

// bf16 variant: fp32 bucket [offset_elems, +n_elems) of the allocation -> cast into the bf16 staging region that starts
// stage_offset_bytes into the allocation (same element index) -> summed over the ranks as bf16 with fp32 accumulation (in the
// switch with multicast, in registers without) -> cast back into the fp32 bucket.  n_elems and offset_elems multiples of 8.
int aa_allreduce_sum_bf16(void* const* peer_bufs, void* multicast_buf, long long flag_offset_bytes, long long stage_offset_bytes, int rank,
                          int world, long long offset_elems, long long n_elems, int channel, int max_blocks, void* stream) {
  AA_REQUIRE(peer_bufs && world >= 2 && world <= AA_AR_MAX_WORLD && rank >= 0 && rank < world, "aa_allreduce_sum_bf16: bad rank / world (%d / %d)", rank, world);
  AA_REQUIRE(channel >= 0 && channel < AA_AR_CHANNELS, "aa_allreduce_sum_bf16: channel must be in [0, %d)", AA_AR_CHANNELS);
  AA_REQUIRE(offset_elems % 8 == 0 && n_elems % 8 == 0 && n_elems >= 0 && flag_offset_bytes % 16 == 0 && stage_offset_bytes % 16 == 0,
             "aa_allreduce_sum_bf16: offset and length must be multiples of 8 elements");
  if (n_elems == 0) return AA_OK;
  ArArgs a{};
  for (int p = 0; p < world; ++p) {
    AA_REQUIRE(peer_bufs[p], "aa_allreduce_sum_bf16: null peer mapping %d", p);
    a.bufs[p] = reinterpret_cast<float*>(static_cast<char*>(peer_bufs[p]) + stage_offset_bytes);      // base of the bf16 staging region
    a.flags[p] = reinterpret_cast<unsigned*>(static_cast<char*>(peer_bufs[p]) + flag_offset_bytes);
  }
  a.mc = multicast_buf ? reinterpret_cast<float*>(static_cast<char*>(multicast_buf) + stage_offset_bytes) : nullptr;
  a.mc_flags = multicast_buf ? reinterpret_cast<unsigned*>(static_cast<char*>(multicast_buf) + flag_offset_bytes) : nullptr;
  a.off = offset_elems / 8; a.n = n_elems / 8; a.rank = rank; a.world = world;
  a.slot_base = channel * AA_AR_MAX_BLOCKS;
  const long long slice_bytes = (n_elems * 2 + world - 1) / world;
  int blocks = (int)((slice_bytes + 65535) / 65536);
  const int cap = max_blocks > 0 && max_blocks < AA_AR_MAX_BLOCKS ? max_blocks : AA_AR_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  float* grads = static_cast<float*>(peer_bufs[rank]) + offset_elems;
  uint4* stage = reinterpret_cast<uint4*>(static_cast<char*>(peer_bufs[rank]) + stage_offset_bytes) + offset_elems / 8;
  const long long n8 = n_elems / 8;
  const int cgrid = (int)std::min<long long>((n8 + 255) / 256, 4LL * num_sms());
  cast_f32_bf16_kernel<<<cgrid, 256, 0, st>>>(grads, stage, n8);
  AA_CHECK_LAUNCH("cast_f32_bf16");
  const int threads = ar_threads();
  if (a.mc) ar_multimem_bf16_kernel<<<blocks, threads, 0, st>>>(a);
  else ar_twoshot_bf16_kernel<<<blocks, threads, 0, st>>>(a);
  AA_CHECK_LAUNCH("allreduce_bf16");
  cast_bf16_f32_kernel<<<cgrid, 256, 0, st>>>(stage, grads, n8);
  AA_CHECK_LAUNCH("cast_bf16_f32");
  return AA_OK;
}

}  // extern "C"
