// Gradient all-reduce over NVLink 5 / NVSwitch multicast (NVLS), one kernel, for sm_100a.
//
// The reference is single-GPU (no collective anywhere).  The view-parallel
// driver either side of the hot path (BASELINE config 4) sums one flat 14*N-float parameter
// gradient across the ranks of a node every step.  Instead of NCCL's ring / tree schedule this
// kernel uses the switch itself: the gradient buffer of every rank is one symmetric allocation
// mapped at a multicast address; rank r owns slice r, pulls it ALREADY SUMMED across all ranks
// with multimem.ld_reduce (the reduction happens inside the NVSwitch) and pushes the result back
// to every rank with one multimem.st.  Per GPU that is 2 * bytes * (W-1)/W over NVLink -- the
// all-reduce lower bound -- with no intermediate buffers, no extra launches and no host proxy.
//
// Cross-GPU synchronisation is inside the kernel (epoch-valued flags at the tail of the symmetric
// buffer, written peer-to-peer with release stores and polled locally with acquire loads):
//   arrive  : "my gradient is final"   -- before anybody reduces
//   depart  : "my slice is broadcast"  -- before the kernel (hence the stream) completes
#include <algorithm>

#include "../../include/jaxsplat_b200.h"
#include "common.cuh"

namespace jsb {
namespace {

constexpr int kMaxRanks = 16;
constexpr int AR_THREADS = 512;

struct PeerTable {
    unsigned *flags[kMaxRanks];  // flag block of every rank (peer-mapped), [2][kMaxRanks] words
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4 *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4 *mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// epochs only grow; wrap-safe comparison
__device__ __forceinline__ bool reached(unsigned seen, unsigned epoch) { return (int)(seen - epoch) >= 0; }

template <int U>
__global__ void __launch_bounds__(AR_THREADS)
multimem_allreduce_kernel(float4 *__restrict__ mc, unsigned long long n_vec, const PeerTable peers,
                          unsigned *__restrict__ my_flags, unsigned *__restrict__ done_counter,
                          int rank, int world, unsigned epoch, int *__restrict__ err) {
    unsigned *arrive = my_flags, *depart = my_flags + kMaxRanks;
    // arrive: one CTA tells every peer; every CTA waits on its own rank's (local) flag block
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)world) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x < (unsigned)world) {
        int spins = 0;
        while (!reached(ld_acquire_sys(arrive + threadIdx.x), epoch)) {
            if (++spins > (1 << 26)) { *err = 3; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();

    // slice `rank`: reduce in the switch, broadcast through the switch
    const unsigned long long per = (n_vec + world - 1) / world;
    const unsigned long long begin = per * rank, end = min(n_vec, begin + per);
    const unsigned long long stride = (unsigned long long)gridDim.x * AR_THREADS;
    for (unsigned long long i = begin + (unsigned long long)blockIdx.x * AR_THREADS + threadIdx.x; i < end;
         i += stride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < end) v[u] = multimem_ld_reduce_add(mc + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < end) multimem_st(mc + i + u * stride, v[u]);
    }

    // depart: the last CTA of this rank tells every peer, then holds the kernel open until every
    // peer's slice has landed here
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *done_counter = 0u;  // ready for the next call (stream ordered)
    if (threadIdx.x < (unsigned)world) {
        st_release_sys(peers.flags[threadIdx.x] + kMaxRanks + rank, epoch);
        int spins = 0;
        while (!reached(ld_acquire_sys(depart + threadIdx.x), epoch)) {
            if (++spins > (1 << 26)) { *err = 4; break; }
            __nanosleep(64);
        }
    }
}

}  // namespace
}  // namespace jsb

extern "C" int jsb200_multimem_allreduce_f32(jsb200_stream_t stream_, void *multicast_base,
                                             void *const *peer_bases, size_t n_floats,
                                             size_t flags_offset_bytes, int rank, int world,
                                             unsigned epoch) {
    using namespace jsb;
    try {
        cudaStream_t stream = (cudaStream_t)stream_;
        if (!multicast_base || !peer_bases || world < 1 || world > kMaxRanks || rank < 0 ||
            rank >= world || (n_floats & 3) || (flags_offset_bytes & 15) ||
            n_floats * 4 > flags_offset_bytes) {
            set_error("multimem_allreduce: bad arguments (world %d rank %d n %zu flags@%zu)", world,
                      rank, n_floats, flags_offset_bytes);
            return 1;
        }
        PeerTable peers;
        for (int r = 0; r < kMaxRanks; ++r)
            peers.flags[r] = r < world
                                 ? reinterpret_cast<unsigned *>((char *)peer_bases[r] + flags_offset_bytes)
                                 : nullptr;
        unsigned *my_flags = peers.flags[rank];
        // words [0, 2*kMaxRanks): arrive / depart epochs; then the CTA counter and an error word
        unsigned *done_counter = my_flags + 2 * kMaxRanks;
        int *err = reinterpret_cast<int *>(my_flags + 2 * kMaxRanks + 1);
        const unsigned long long n_vec = n_floats / 4;
        const unsigned long long per = (n_vec + world - 1) / world;
        // link-bound, not SM-bound (measured flat from 32 to 592 CTAs): one CTA per SM leaves the
        // other half of every SM's thread slots to whatever else runs on the device
        unsigned blocks = (unsigned)std::min<unsigned long long>(148ull, (per + AR_THREADS - 1) / AR_THREADS);
        if (blocks == 0) blocks = 1;
        multimem_allreduce_kernel<4><<<blocks, AR_THREADS, 0, stream>>>(
            reinterpret_cast<float4 *>(multicast_base), n_vec, peers, my_flags, done_counter, rank,
            world, epoch, err);
        JSB_CUDA_OK(cudaGetLastError());
        note_launch();
        return 0;
    } catch (...) {
        set_error("multimem_allreduce: exception");
        return 1;
    }
}
