// Stable LSD "onesweep" radix sort of (key,value) pairs for sm_100a, written from scratch.
//
// Replaces the reference's cub::DeviceRadixSort::SortPairs<int64,int32> call (lib/ops.cu:722-757,
// 6 passes over 12-byte pairs).  Here keys are as narrow as the live bits allow (uint16 tile ids
// or uint32 depth bits), one upfront kernel builds the digit histograms of ALL passes, and each
// pass is a single kernel: load tile -> per-warp match_any ranking -> chained look-back for the
// cross-CTA digit offsets -> shared-memory reorder -> coalesced scatter.  No pass reads its input
// twice and there is no per-pass global scan.
//
// Algorithmic HBM bytes per pair: sizeof(Key) (histogram read) + passes * 2 * (sizeof(Key) + 4).
#include "common.cuh"
#include "kernels.cuh"

namespace jsb {
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
// pairs per thread: 16 for the I-scale tile sort (4096 per CTA), 4 for N-scale sorts so that a
// million keys still spread over ~1000 CTAs instead of 245
constexpr int RADIX = 256;
constexpr int MAX_PASSES = 4;

// status word for the per-digit look-back: 2 flag bits | 30-bit count
constexpr unsigned ST_AGG = 1u << 30, ST_INC = 2u << 30, ST_MASK = (1u << 30) - 1;

__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- upfront histogram of every pass ---------------------------------------------------------
// 16-byte loads + plain shared-memory atomics (warp-aggregating them with match.any was measured
// 4x slower on B200: MATCH.ANY costs more than the bank conflicts it removes).
template <typename KeyT>
__global__ void __launch_bounds__(256)
radix_hist_kernel(const KeyT *__restrict__ keys, unsigned n, int passes,
                  unsigned *__restrict__ ghist /*[passes][256]*/) {
    __shared__ unsigned s_h[MAX_PASSES * RADIX];
    constexpr int PER = 16 / sizeof(KeyT);  // keys per 16-byte load
    for (int i = threadIdx.x; i < passes * RADIX; i += 256) s_h[i] = 0;
    __syncthreads();
    const unsigned nvec = n / PER;
    const unsigned stride = gridDim.x * 256u;
    for (unsigned v = blockIdx.x * 256u + threadIdx.x; v < nvec; v += stride) {
        unsigned k[PER];
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(keys) + v);
        if (sizeof(KeyT) == 2) {
            k[0] = q.x & 0xffffu; k[1] = q.x >> 16; k[2 % PER] = q.y & 0xffffu; k[3 % PER] = q.y >> 16;
            k[4 % PER] = q.z & 0xffffu; k[5 % PER] = q.z >> 16; k[6 % PER] = q.w & 0xffffu; k[7 % PER] = q.w >> 16;
        } else {
            k[0] = q.x; k[1] = q.y; k[2] = q.z; k[3] = q.w;
        }
#pragma unroll
        for (int e = 0; e < PER; ++e)
            for (int ps = 0; ps < passes; ++ps)
                atomicAdd(&s_h[ps * RADIX + ((k[e] >> (8 * ps)) & 255u)], 1u);
    }
    // scalar tail
    if (blockIdx.x == 0) {
        for (unsigned i = nvec * PER + threadIdx.x; i < n; i += 256) {
            const unsigned kk = keys[i];
            for (int ps = 0; ps < passes; ++ps) atomicAdd(&s_h[ps * RADIX + ((kk >> (8 * ps)) & 255u)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += 256) {
        const unsigned c = s_h[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// exclusive scan of each pass's 256-bin histogram, in place; one CTA of 256 threads per pass
__global__ void __launch_bounds__(256) radix_hist_scan_kernel(unsigned *__restrict__ ghist) {
    __shared__ unsigned s_ws[256 / 32 + 1];
    unsigned *h = ghist + blockIdx.x * RADIX;
    const unsigned v = h[threadIdx.x];
    unsigned total;
    const unsigned e = cta_exclusive_scan<256>(v, &total, s_ws);
    h[threadIdx.x] = e;
}

// ---- one onesweep pass ------------------------------------------------------------------------
// CHAIN = true : onesweep (tile order from a ticket, cross-tile digit offsets by chained look-back)
// CHAIN = false: downsweep of the upsweep/scan/downsweep variant: `status` holds, digit-major
//                [256][tiles], the row-exclusive tile offsets and `gbase` the 256 row totals; no
//                inter-CTA dependency, no spinning.
template <typename KeyT, int RS_ITEMS, bool CHAIN>
__global__ void __launch_bounds__(RS_THREADS)
onesweep_pass_kernel(const KeyT *__restrict__ keys_in, const unsigned *__restrict__ vals_in,
                     KeyT *__restrict__ keys_out, unsigned *__restrict__ vals_out, unsigned n,
                     int shift, int live_bits, const unsigned *__restrict__ gbase /*[256]*/,
                     unsigned *__restrict__ status, unsigned *__restrict__ ticket,
                     int *__restrict__ err) {
    __shared__ unsigned s_whist[RS_WARPS * RADIX];  // per-warp digit counts -> exclusive over warps
    __shared__ unsigned s_binstart[RADIX];          // exclusive scan of the CTA histogram
    __shared__ unsigned s_goff[RADIX];              // global offset minus s_binstart (wraps)
    __shared__ unsigned s_ws[RS_THREADS / 32 + 1];
    __shared__ unsigned s_tile;
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    __shared__ KeyT s_keys[RS_TILE];
    __shared__ unsigned s_vals[RS_TILE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (CHAIN && tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < RS_WARPS * RADIX; i += RS_THREADS) s_whist[i] = 0;
    __syncthreads();
    const unsigned tile = CHAIN ? s_tile : blockIdx.x;
    const unsigned tile_base = tile * (unsigned)RS_TILE;
    const unsigned count = min((unsigned)RS_TILE, n - tile_base);

    // warp-striped load: item i of lane l is element warp*ITEMS*32 + i*32 + l of the tile
    KeyT key[RS_ITEMS];
    unsigned val[RS_ITEMS];
    unsigned short rank[RS_ITEMS];
    const unsigned wbase = warp * (RS_ITEMS * 32) + lane;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const unsigned o = wbase + i * 32;
        key[i] = (o < count) ? keys_in[tile_base + o] : (KeyT)~(KeyT)0;
        val[i] = (o < count) ? vals_in[tile_base + o] : 0u;
    }

    // stable ranking inside the warp's chunk: memory order == (item, lane) order.
    // Peers (lanes holding the same digit) come from one ballot per live digit bit: the native
    // MATCH.ANY is data dependent on B200 -- ~2x slower when all 32 lanes differ, which is the
    // common case for the low byte of consecutive tile ids (measured: 184 vs 101 us per pass).
    unsigned *wh = s_whist + warp * RADIX;
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const bool valid = (wbase + i * 32) < count;
        const unsigned d = valid ? (((unsigned)key[i] >> shift) & 255u) : 256u;
        unsigned peers = __ballot_sync(0xffffffffu, valid);
        peers = valid ? peers : ~peers;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b < live_bits) {
                const unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? m : ~m;
            }
        }
        const unsigned lower = peers & lt_mask;
        unsigned pre = 0;
        if (valid) pre = wh[d];
        __syncwarp();
        if (valid && lower == 0) wh[d] = pre + __popc(peers);
        __syncwarp();
        rank[i] = (unsigned short)(pre + __popc(lower));
    }
    __syncthreads();

    // thread d owns digit d: exclusive scan over warps, CTA histogram, look-back
    unsigned cta_count = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const unsigned c = s_whist[w * RADIX + tid];
        s_whist[w * RADIX + tid] = cta_count;
        cta_count += c;
    }
    unsigned total;
    const unsigned binstart = cta_exclusive_scan<RS_THREADS>(cta_count, &total, s_ws);
    s_binstart[tid] = binstart;
    if (!CHAIN) {
        // digit base = exclusive scan of the row totals; tile offset precomputed by scan_rows
        unsigned tot2;
        const unsigned dbase = cta_exclusive_scan<RS_THREADS>(gbase[tid], &tot2, s_ws);
        s_goff[tid] = dbase + status[(size_t)tid * gridDim.x + tile] - binstart;
    } else {
        unsigned *st = status + (size_t)tile * RADIX + tid;
        unsigned excl = 0;
        if (tile == 0) {
            st_relaxed(st, ST_INC | cta_count);
        } else {
            st_relaxed(st, ST_AGG | cta_count);
            // walk predecessors, LB status loads in flight per step (the walk is a chain of L2
            // round trips; speculative loads of older tiles cost nothing and cut it LB-fold)
#ifndef JSB_RADIX_LB
#define JSB_RADIX_LB 4
#endif
            constexpr int LB = JSB_RADIX_LB;
            int j = (int)tile - 1;
            int spins = 0;
            bool done = false;
            while (!done) {
                unsigned w[LB];
#pragma unroll
                for (int u = 0; u < LB; ++u)
                    w[u] = (j - u >= 0) ? ld_relaxed(status + (size_t)(j - u) * RADIX + tid) : ST_INC;
#pragma unroll
                for (int u = 0; u < LB; ++u) {
                    if (done) break;
                    const unsigned f = w[u] & ~ST_MASK;
                    if (f == 0) {  // not published yet: retry from this tile
                        if (++spins > kSpinLimit) { *err = 2; done = true; }
                        else __nanosleep(20);
                        break;
                    }
                    excl += w[u] & ST_MASK;
                    --j;
                    if (f == ST_INC) done = true;
                }
            }
            st_relaxed(st, ST_INC | ((excl + cta_count) & ST_MASK));
        }
        s_goff[tid] = gbase[tid] + excl - binstart;
    }
    __syncthreads();

    // reorder through shared memory
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        if ((wbase + i * 32) < count) {
            const unsigned d = ((unsigned)key[i] >> shift) & 255u;
            const unsigned pos = s_binstart[d] + wh[d] + rank[i];
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
    }
    __syncthreads();
    // coalesced scatter: consecutive positions of one digit go to consecutive addresses
    for (unsigned pos = tid; pos < count; pos += RS_THREADS) {
        const KeyT k = s_keys[pos];
        const unsigned d = ((unsigned)k >> shift) & 255u;
        const unsigned o = pos + s_goff[d];
        keys_out[o] = k;
        vals_out[o] = s_vals[pos];
    }
}

// ---- chain-free variant: per-tile digit histograms + row scan ---------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(256)
upsweep_kernel(const KeyT *__restrict__ keys, unsigned n, int shift,
               unsigned *__restrict__ tile_hist /*[256][tiles]*/) {
    __shared__ unsigned s_h[RADIX];
    constexpr int TILE = RS_THREADS * 16;
    constexpr int PER = 16 / sizeof(KeyT);
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned tile = blockIdx.x;
    const unsigned base = tile * (unsigned)TILE;
    const unsigned count = min((unsigned)TILE, n - base);
    const KeyT *src = keys + base;  // base is a multiple of 4096 keys: 16-byte aligned
    const unsigned nvec = count / PER;
    for (unsigned v = threadIdx.x; v < nvec; v += 256) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src) + v);
        unsigned k[PER];
        if (sizeof(KeyT) == 2) {
            k[0] = q.x & 0xffffu; k[1] = q.x >> 16; k[2 % PER] = q.y & 0xffffu; k[3 % PER] = q.y >> 16;
            k[4 % PER] = q.z & 0xffffu; k[5 % PER] = q.z >> 16; k[6 % PER] = q.w & 0xffffu; k[7 % PER] = q.w >> 16;
        } else {
            k[0] = q.x; k[1] = q.y; k[2] = q.z; k[3] = q.w;
        }
#pragma unroll
        for (int e = 0; e < PER; ++e) atomicAdd(&s_h[(k[e] >> shift) & 255u], 1u);
    }
    for (unsigned i = nvec * PER + threadIdx.x; i < count; i += 256)
        atomicAdd(&s_h[((unsigned)src[i] >> shift) & 255u], 1u);
    __syncthreads();
    tile_hist[(size_t)threadIdx.x * gridDim.x + tile] = s_h[threadIdx.x];
}

// one CTA per digit: exclusive scan of its row of per-tile counts in place, row total out.
// Each thread owns a contiguous chunk (serial sum, one CTA scan of the chunk sums, serial write).
__global__ void __launch_bounds__(256)
scan_rows_kernel(unsigned *__restrict__ tile_hist, unsigned tiles, unsigned *__restrict__ row_totals) {
    __shared__ unsigned s_ws[256 / 32 + 1];
    unsigned *row = tile_hist + (size_t)blockIdx.x * tiles;
    const unsigned per = (tiles + 255u) / 256u;
    const unsigned b0 = min(tiles, threadIdx.x * per), b1 = min(tiles, b0 + per);
    unsigned sum = 0;
    for (unsigned i = b0; i < b1; ++i) sum += row[i];
    unsigned total;
    unsigned run = cta_exclusive_scan<256>(sum, &total, s_ws);
    for (unsigned i = b0; i < b1; ++i) {
        const unsigned v = row[i];
        row[i] = run;
        run += v;
    }
    if (threadIdx.x == 0) row_totals[blockIdx.x] = total;
}

#ifndef JSB_RADIX_ITEMS
#define JSB_RADIX_ITEMS 16  // 4, 8 and 32 per thread measured no faster at N = 1M (tools/gpu_ab_builds.sh)
#endif
inline int items_for(unsigned n) { (void)n; return JSB_RADIX_ITEMS; }
inline unsigned num_tiles_of(unsigned n) { return ceil_div(n, (unsigned)(RS_THREADS * items_for(n))); }
inline int passes_of(int bits) { return (bits + 7) / 8; }

}  // namespace

// scratch layout: [ghist: MAX_PASSES*256 u32][tickets: MAX_PASSES u32 (padded to 16)]
//                 [status: passes * tiles * 256 u32]
// n at or above this uses upsweep/scan/downsweep (no inter-CTA spinning); below it the sort is
// latency bound either way and the single-kernel-per-pass onesweep wins on launch count
constexpr unsigned kChainFreeMin = 1u << 21;

size_t radix_scratch_bytes(unsigned n, int bits) {
    const size_t head = (MAX_PASSES * RADIX + 16) * sizeof(unsigned);
    if (n >= kChainFreeMin) return head + (size_t)num_tiles_of(n) * RADIX * sizeof(unsigned);
    return head + (size_t)passes_of(bits) * num_tiles_of(n) * RADIX * sizeof(unsigned);
}

template <typename KeyT>
int radix_sort_pairs(cudaStream_t stream, KeyT *keys_a, KeyT *keys_b, unsigned *vals_a,
                     unsigned *vals_b, unsigned n, int bits, void *scratch, int *err,
                     bool *result_in_b) {
    *result_in_b = false;
    if (n == 0 || bits <= 0) return 0;
    const int passes = passes_of(bits);
    if (passes > MAX_PASSES || (size_t)bits > sizeof(KeyT) * 8) {
        set_error("radix_sort_pairs: unsupported bit count %d", bits);
        return 1;
    }
    if (n >= (1u << 30)) {
        set_error("radix_sort_pairs: n=%u exceeds the 2^30 look-back counter range", n);
        return 1;
    }
    const unsigned tiles = num_tiles_of(n);
    unsigned *ghist = reinterpret_cast<unsigned *>(scratch);
    unsigned *tickets = ghist + MAX_PASSES * RADIX;
    unsigned *status = tickets + 16;
    if (n >= kChainFreeMin) {
        // upsweep (per-tile digit counts) -> row scan -> downsweep; scratch needs no zeroing
        KeyT *kin = keys_a, *kout = keys_b;
        unsigned *vin = vals_a, *vout = vals_b;
        unsigned *row_totals = ghist;
        for (int ps = 0; ps < passes; ++ps) {
            upsweep_kernel<KeyT><<<tiles, 256, 0, stream>>>(kin, n, 8 * ps, status);
            JSB_CUDA_OK(cudaGetLastError());
            scan_rows_kernel<<<RADIX, 256, 0, stream>>>(status, tiles, row_totals);
            JSB_CUDA_OK(cudaGetLastError());
            onesweep_pass_kernel<KeyT, JSB_RADIX_ITEMS, false><<<tiles, RS_THREADS, 0, stream>>>(
                kin, vin, kout, vout, n, 8 * ps, min(8, bits - 8 * ps), row_totals, status, nullptr,
                err);
            JSB_CUDA_OK(cudaGetLastError());
            note_launch(3);
            KeyT *tk = kin; kin = kout; kout = tk;
            unsigned *tv = vin; vin = vout; vout = tv;
            *result_in_b = !*result_in_b;
        }
        return 0;
    }
    JSB_CUDA_OK(cudaMemsetAsync(scratch, 0, radix_scratch_bytes(n, bits), stream));
    const unsigned hist_blocks = min(ceil_div(n, 256u * 32u), 148u * 8u);
    radix_hist_kernel<KeyT><<<hist_blocks, 256, 0, stream>>>(keys_a, n, passes, ghist);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    radix_hist_scan_kernel<<<passes, 256, 0, stream>>>(ghist);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    KeyT *kin = keys_a, *kout = keys_b;
    unsigned *vin = vals_a, *vout = vals_b;
    for (int ps = 0; ps < passes; ++ps) {
        onesweep_pass_kernel<KeyT, JSB_RADIX_ITEMS, true><<<tiles, RS_THREADS, 0, stream>>>(
            kin, vin, kout, vout, n, 8 * ps, min(8, bits - 8 * ps), ghist + ps * RADIX,
            status + (size_t)ps * tiles * RADIX, tickets + ps, err);
        JSB_CUDA_OK(cudaGetLastError());
    note_launch();
        KeyT *tk = kin; kin = kout; kout = tk;
        unsigned *tv = vin; vin = vout; vout = tv;
        *result_in_b = !*result_in_b;
    }
    return 0;
}

template int radix_sort_pairs<uint16_t>(cudaStream_t, uint16_t *, uint16_t *, unsigned *,
                                        unsigned *, unsigned, int, void *, int *, bool *);
template int radix_sort_pairs<uint32_t>(cudaStream_t, uint32_t *, uint32_t *, unsigned *,
                                        unsigned *, unsigned, int, void *, int *, bool *);

}  // namespace jsb
