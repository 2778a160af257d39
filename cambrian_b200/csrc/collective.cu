// cambrian_b200 — in-switch gradient all-reduce over NVLink 5 / NVSwitch for the bucketed data-parallel step
// (replaces the gradient reduction the reference leaves to torch_xla FSDP / `xm.all_reduce`,
// cambrian_trainer.py:181-190).
//
// Every rank holds its gradient bucket at the SAME offset of a symmetric allocation that is also mapped through an NVLS
// multicast address.  One kernel per bucket, small (128-thread) CTAs that co-reside with the GEMM CTAs of the backward pass:
//     barrier            every rank has finished writing its copy of the bucket (kernel boundary + release/acquire flags)
//     reduce + publish   rank r owns slice r of the bucket: `multimem.ld_reduce` asks the SWITCH for the fp32-accumulated sum
//                        of that 16-byte chunk over all ranks' copies, `multimem.st` multicasts the bf16 result back into
//                        every rank's copy.  Each GPU therefore issues loads / stores for only 1/world of the bucket — the
//                        reduction arithmetic and the fan-out happen in the NVSwitch, not in SM instructions — so the
//                        SM-side cost is a few thousand threads of loads in flight, not whole SMs.
//     barrier            every rank has published its slice: the whole bucket is final everywhere.
// Without multicast support (no NVSwitch fabric) the same schedule runs on plain peer pointers: the owner of a slice loads it
// from every peer over NVLink, adds in fp32 and stores the result to every peer.
//
// Flags: signal pad word [channel = CTA][source rank] on every rank, written with st.release.sys through the peer
// mapping, polled with ld.acquire.sys; values are monotonically increasing epochs supplied by the host (two per launch),
// so nothing is ever reset and back-to-back launches on one stream cannot confuse each other.  A spin that outlives
// ~10 s traps (a dead peer must become an error, not a hung GPU).
#include "common.cuh"

namespace cb {

constexpr int AR_MAX_RANKS = 8;
constexpr int AR_MAX_CTAS = 160;
constexpr int AR_THREADS = 128;  // 4 warps x 32 registers: fits NEXT TO a resident GEMM CTA (320 threads, 54 K registers, no
                                  // smem of ours) on the same SM — the collective does not take SMs away from the backward pass
constexpr int AR_PAD_WORDS_OFFSET = 0;     // the flag words live in their own small symmetric allocation (comm.py)

struct ArPeers {
  unsigned long long buf[AR_MAX_RANKS];  // peer-mapped base address of each rank's symmetric buffer
  unsigned long long pad[AR_MAX_RANKS];  // peer-mapped base address of each rank's signal pad
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all ranks' CTA `blockIdx.x` meet here; `epoch` strictly increases from launch to launch
__device__ __forceinline__ void cross_rank_barrier(const ArPeers& peers, int rank, int world, unsigned int epoch) {
  __syncthreads();
  if (threadIdx.x < world) {
    const int peer = threadIdx.x;
    unsigned int* theirs = reinterpret_cast<unsigned int*>(peers.pad[peer]) + AR_PAD_WORDS_OFFSET + blockIdx.x * AR_MAX_RANKS + rank;
    st_release_sys(theirs, epoch);
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(peers.pad[rank]) + AR_PAD_WORDS_OFFSET + blockIdx.x * AR_MAX_RANKS + peer;
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s at 2 GHz: a peer never arrived
    }
  }
  __syncthreads();
}

__device__ __forceinline__ uint4 mm_ld_reduce(unsigned long long addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(unsigned long long addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 p2p_reduce(const ArPeers& peers, int world, long long c) {
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = 0; p < world; ++p) {
    uint4 v;  // peer (or own) copy through the NVLink mapping: bypass L1, the data was just written by another GPU's kernels
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(reinterpret_cast<const uint4*>(peers.buf[p]) + c)
                 : "memory");
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  }
  return pack8(acc);
}

template <bool MULTIMEM>
__global__ void __launch_bounds__(AR_THREADS)
allreduce_bf16_kernel(unsigned long long mc_base, ArPeers peers, long long offset_bytes, long long nbytes, int rank, int world,
                      unsigned int epoch) {
  cross_rank_barrier(peers, rank, world, epoch);
  // slice of this rank, in 16-byte chunks (nbytes is a multiple of 16 * world: the engine pads buckets)
  const long long chunks = nbytes / 16 / world;
  const long long first = offset_bytes / 16 + rank * chunks;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (MULTIMEM) {
    const unsigned long long base = mc_base + static_cast<unsigned long long>(first) * 16ull;
    for (; i + 3 * stride < chunks; i += 4 * stride) {  // 4 independent switch round trips in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = mm_ld_reduce(base + static_cast<unsigned long long>(i + u * stride) * 16ull);
#pragma unroll
      for (int u = 0; u < 4; ++u) mm_st(base + static_cast<unsigned long long>(i + u * stride) * 16ull, v[u]);
    }
    for (; i < chunks; i += stride) mm_st(base + static_cast<unsigned long long>(i) * 16ull, mm_ld_reduce(base + static_cast<unsigned long long>(i) * 16ull));
  } else {
    for (; i < chunks; i += stride) {
      const uint4 out = p2p_reduce(peers, world, first + i);
      for (int p = 0; p < world; ++p) *(reinterpret_cast<uint4*>(peers.buf[p]) + first + i) = out;
    }
  }
  __threadfence_system();  // this thread's published chunks are visible system-wide before the flag below is raised
  cross_rank_barrier(peers, rank, world, epoch + 1);
}

int allreduce_symm_launch(unsigned long long mc_base, const unsigned long long* buf_ptrs, const unsigned long long* pad_ptrs,
                          long long offset_bytes, long long nbytes, int rank, int world, unsigned int epoch, int ctas,
                          cudaStream_t st) {
  CB_CHECK_ARG(world >= 2 && world <= AR_MAX_RANKS && rank >= 0 && rank < world, "allreduce: bad rank %d / world %d", rank, world);
  CB_CHECK_ARG(nbytes > 0 && nbytes % (16LL * world) == 0 && offset_bytes % 16 == 0,
               "allreduce: the byte range must be a multiple of 16 * world (got %lld at %lld)", nbytes, offset_bytes);
  CB_CHECK_ARG(ctas >= 1 && ctas <= AR_MAX_CTAS, "allreduce: ctas=%d out of [1,%d]", ctas, AR_MAX_CTAS);
  CB_CHECK_ARG(buf_ptrs && pad_ptrs, "allreduce: null peer tables");
  ArPeers peers;
  for (int i = 0; i < AR_MAX_RANKS; ++i) {
    peers.buf[i] = i < world ? buf_ptrs[i] : 0ull;
    peers.pad[i] = i < world ? pad_ptrs[i] : 0ull;
  }
  if (mc_base)
    allreduce_bf16_kernel<true><<<ctas, AR_THREADS, 0, st>>>(mc_base, peers, offset_bytes, nbytes, rank, world, epoch);
  else
    allreduce_bf16_kernel<false><<<ctas, AR_THREADS, 0, st>>>(0ull, peers, offset_bytes, nbytes, rank, world, epoch);
  CB_CUDA_LAUNCH_CHECK("allreduce_bf16");
  return CB_OK;
}

}  // namespace cb
