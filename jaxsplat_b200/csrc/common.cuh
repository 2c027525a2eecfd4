// jaxsplat_b200 -- shared host/device definitions for the sm_100a splatting hot path.
//
// Replaces (does not copy) the reference's lib/common.h + lib/ops.h:
//   * Descriptor mirrors the 80-byte opaque POD of reference lib/ops.h:9-22 so the reference
//     library and this one are interchangeable under one harness.
//   * No exceptions ever cross the C ABI: errors are recorded in a sticky per-thread string
//     (reference throws std::runtime_error through XLA's C frame, lib/common.h:16-23).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace jsb {

struct alignas(16) Descriptor {
    uint32_t num_points;     // @0
    uint32_t img_w, img_h, img_d;  // @4  dim3 (W,H,1)
    float fx, fy, cx, cy;    // @16 float4 intrins
    float glob_scale;        // @32
    float clip_thresh;       // @36
    uint32_t block_width;    // @40
    uint32_t grid_dim_1d;    // @44 (kept for layout compatibility; unused by our launches)
    uint32_t block_dim_1d;   // @48
    uint32_t tiles_x, tiles_y, tiles_z;  // @52 dim3 grid_dim_2d
    uint32_t bdx, bdy, bdz;  // @64 dim3 block_dim_2d
    uint32_t pad;            // @76 (zeroed; the reference leaves it uninitialised)
};
static_assert(sizeof(Descriptor) == 80, "descriptor must stay 80 bytes");

// sticky error channel (capi.cu)
void set_error(const char *fmt, ...);
// counts kernel launches of this library (bench.py reports it as gpu_launches)
void note_launch(int n = 1);

#define JSB_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            jsb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                           cudaGetErrorString(_e));                                    \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Exact-rounding arithmetic.  Every value that feeds an integer / index decision (radii, tile
// boxes, alpha tests, transmittance tests) is computed with explicit round-to-nearest intrinsics
// in the operation order the reference's compiled kernels use (nvcc default -fmad=true contraction,
// IEEE div/sqrt; order read from the reference's sm_100a PTX, see DESIGN.md "rounding contract").
// Intrinsics are never re-contracted by ptxas, so the order below is the order executed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

// MUFU.EX2 of x (x already multiplied by log2(e)).  The reference's __expf lowers to
// ex2.approx.f32 (non-ftz), whose sm_100a expansion only differs from the .ftz form for
// x < -126, where the result (a denormal vs 0) multiplied by an opacity <= 1 is far below the
// 1/255 alpha cut either way -- so every accept/reject decision and every accepted value is
// identical, and we save the two guard instructions per pair.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr float kNegLog2e = -1.44269502162933349609375f;  // 0fBFB8AA3B
constexpr float kAlphaMin = 0.0039215688593685626983642578125f;  // 0f3B808081 = 1.f/255.f
constexpr float kTmin = 9.9999997473787516355514526367188e-05f;  // 0f38D1B717 = 1e-4f

// sigma of the reference rasterisers (forward.cu:424-426, backward.cu:257-259) in the exact
// operation order of its compiled code.
__device__ __forceinline__ float conic_sigma(float cx, float cy, float cz, float dx, float dy) {
    float a = fmul(cx, dx);
    float b = fmul(cz, dy);
    float c = fmul(dy, b);
    float s = ffma(dx, a, c);
    s = fmul(s, 0.5f);
    float d = fmul(cy, dx);
    return ffma(d, dy, s);
}

// Asynchronous global -> shared copies (LDGSTS): the rasterisers stage the NEXT batch of splat
// records with these while the current batch is being composited.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;" ::: "memory");
}

__host__ __device__ inline unsigned ceil_div(unsigned a, unsigned b) { return (a + b - 1) / b; }

// Programmatic dependent launch (option "pdl"): a kernel launched through launch_chained() may
// become resident while the kernel before it in the stream drains its last wave.  Such a kernel
// calls chain_wait() before it touches global memory (all prerequisite grids complete, their
// writes visible) and only then lets ITS successor in (chain_release()), so at most one waiting
// grid is resident at a time.  Both are no-ops for a kernel launched the plain way.
__device__ __forceinline__ void chain_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
extern bool g_pdl;  // capi.cu
template <typename... KArgs, typename... Args>
inline void launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                           cudaStream_t stream, Args &&...args) {
    if (!g_pdl) {
        kernel<<<grid, block, smem, stream>>>(static_cast<KArgs>(args)...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...) == cudaErrorNotSupported) {
        (void)cudaGetLastError();  // (no programmatic launches here: plain ones from now on)
        g_pdl = false;
        kernel<<<grid, block, smem, stream>>>(static_cast<KArgs>(args)...);
    }
}

// tile box of a projected Gaussian; reference helpers.h:9-32 (get_tile_bbox -> get_bbox).
// float->int conversions are cvt.rzi (truncate, saturating), exactly the reference's casts.
__device__ __forceinline__ void tile_bbox(float px, float py, float radius, float bw,
                                          unsigned tiles_x, unsigned tiles_y, unsigned &x0,
                                          unsigned &x1, unsigned &y0, unsigned &y1) {
    float tcx = fdiv(px, bw), tcy = fdiv(py, bw), tr = fdiv(radius, bw);
    x0 = min((unsigned)max(__float2int_rz(fsub(tcx, tr)), 0), tiles_x);
    x1 = min((unsigned)max(__float2int_rz(fadd(fadd(tcx, tr), 1.0f)), 0), tiles_x);
    y0 = min((unsigned)max(__float2int_rz(fsub(tcy, tr)), 0), tiles_y);
    y1 = min((unsigned)max(__float2int_rz(fadd(fadd(tcy, tr), 1.0f)), 0), tiles_y);
}

// ---------------------------------------------------------------------------------------------
// Coalesced AoS <-> shared staging.  The operator boundary fixes AoS (N,3)/(N,6) float arrays;
// a CTA moves its 256-record slab with 16-byte vector transactions and each thread then picks
// its own record out of shared memory (stride 3 and 6 words: at most 2-way bank conflicts).
// ---------------------------------------------------------------------------------------------
template <int K, int THREADS>
__device__ __forceinline__ void slab_load(float *smem, const float *__restrict__ g,
                                          long long first_elem, long long total_elems) {
    long long remain = total_elems - first_elem;
    int n = remain < (long long)K * THREADS ? (int)remain : K * THREADS;
    const float *src = g + first_elem;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += THREADS)
            reinterpret_cast<float4 *>(smem)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += THREADS) smem[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < n; i += THREADS) smem[i] = __ldg(src + i);
    }
}

template <int K, int THREADS>
__device__ __forceinline__ void slab_store(const float *smem, float *__restrict__ g,
                                           long long first_elem, long long total_elems) {
    long long remain = total_elems - first_elem;
    int n = remain < (long long)K * THREADS ? (int)remain : K * THREADS;
    float *dst = g + first_elem;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += THREADS)
            reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(smem)[i];
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += THREADS) dst[i] = smem[i];
    } else {
        for (int i = threadIdx.x; i < n; i += THREADS) dst[i] = smem[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Single-pass chained scan ("decoupled look-back") status words: flag in the high word, value in
// the low word, written with one 64-bit store so no fence is needed between flag and value.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kFlagAggregate = 1ull << 32;
constexpr unsigned long long kFlagInclusive = 2ull << 32;
constexpr int kSpinLimit = 1 << 22;  // bounded spin: a scheduling bug becomes an error, not a hang

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by the first warp of a CTA that owns chain position `pos`.  Publishes the CTA's
// aggregate, walks predecessors and returns the exclusive prefix (valid in every lane), then
// publishes the inclusive prefix.  `err` is set if the spin bound trips.
//
// Every CTA of a wave reaches this point at about the same time, so the nearest predecessor that
// already holds an INCLUSIVE prefix is roughly "number of resident CTAs" positions back; the walk
// is a chain of L2 round trips, so each round inspects a window of 32 x kChainWindow = 256
// predecessors (8 independent loads per lane in flight) instead of 32.
constexpr int kChainWindow = 1;  // wider windows measured slower on B200 (status polling traffic)
__device__ __forceinline__ unsigned chain_exclusive_prefix(unsigned long long *status, int pos,
                                                           unsigned aggregate, int *err) {
    const int lane = threadIdx.x & 31;
    if (pos == 0) {
        if (lane == 0) st_status(status, kFlagInclusive | aggregate);
        return 0u;
    }
    if (lane == 0) st_status(status + pos, kFlagAggregate | aggregate);
    unsigned exclusive = 0;
    int j = pos - 1;  // nearest predecessor of this round's window
    while (true) {
        // lane owns distances lane*W .. lane*W+W-1 from j (nearest first)
        unsigned long long w[kChainWindow];
        const int first = j - lane * kChainWindow;
#pragma unroll
        for (int u = 0; u < kChainWindow; ++u)
            w[u] = (first - u >= 0) ? ld_status(status + (first - u)) : kFlagInclusive;
        int spins = 0;
        bool pending = true;
        while (pending) {
            pending = false;
#pragma unroll
            for (int u = 0; u < kChainWindow; ++u) {
                if ((w[u] >> 32) == 0) {
                    w[u] = ld_status(status + (first - u));
                    pending = pending || ((w[u] >> 32) == 0);
                }
            }
            if (pending) {
                if (++spins > kSpinLimit) {
                    *err = 1;
#pragma unroll
                    for (int u = 0; u < kChainWindow; ++u)
                        if ((w[u] >> 32) == 0) w[u] = kFlagInclusive;
                    pending = false;
                } else {
                    __nanosleep(32);
                }
            }
        }
        // per-lane: sum up to and including the nearest inclusive entry
        unsigned v = 0;
        bool has_incl = false;
#pragma unroll
        for (int u = 0; u < kChainWindow; ++u) {
            if (!has_incl) {
                v += (unsigned)w[u];
                has_incl = (w[u] >> 32) == 2;
            }
        }
        const unsigned incl = __ballot_sync(0xffffffffu, has_incl);
        if (incl) {
            const int firstl = __ffs(incl) - 1;
            if (lane > firstl) v = 0;
        }
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        exclusive += v;
        if (incl) break;
        j -= 32 * kChainWindow;
    }
    if (lane == 0) st_status(status + pos, kFlagInclusive | (unsigned)(exclusive + aggregate));
    return exclusive;
}

// CTA-wide exclusive scan of one unsigned per thread (THREADS multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the CTA sum (all threads).
template <int THREADS>
__device__ __forceinline__ unsigned cta_exclusive_scan(unsigned v, unsigned *total,
                                                       unsigned *warp_sums /*[THREADS/32 + 1]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned s = lane < THREADS / 32 ? warp_sums[lane] : 0;
        unsigned si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        if (lane < THREADS / 32) warp_sums[lane] = si - s;
        if (lane == 31) warp_sums[THREADS / 32] = si;
    }
    __syncthreads();
    unsigned base = warp_sums[warp];
    *total = warp_sums[THREADS / 32];
    return base + inc - v;
}

}  // namespace jsb
