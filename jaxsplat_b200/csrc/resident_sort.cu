// Depth order of the Gaussians in ONE kernel launch, for sm_100a.
//
// Replaces (together with rowcol_lists.cu) the reference's cub::DeviceRadixSort::SortPairs over
// 64-bit (tile | depth) keys (lib/ops.cu:722-757): here only the N Gaussians are ordered by their
// 32-bit depth keys -- stable, ties keep index order -- and the per-tile lists are distributed
// from that order.
//
// Why a new kernel: at N = 1M a 4-pass onesweep (ours or CUB's, measured on the same box:
// 91 vs 92 us, profiles/r02a_depth_sort_vs_cub.json) moves 16 MB per pass, ~3 us of HBM time, and
// spends the other ~16 us of every pass on launch ramp and the chained look-back through ~245
// tiles.  This kernel is launched once, cooperatively, with at most one CTA per SM; a pass is
//   A.  rank the CTA's keys by digit (stable, per-warp ballots), publish its 512 digit counts,
//       and group the tile by digit in shared memory
//   -- grid barrier --
//   B1. CTA c scans, over all CTAs, the counts of the few digits it owns (512 / grid of them)
//       and publishes every CTA's exclusive prefix for those digits (each count is read ONCE;
//       letting every CTA sum the rows before it costs grid^2 x 2 KB = 40 MB of L2 reads a pass)
//   -- grid barrier --
//   B2. every CTA reads its own row of prefixes and the digit totals, scans the totals and
//       scatters its tile
//   -- grid barrier --
// so the cross-CTA dependency is three barriers instead of a serial chain, nothing is launched
// between passes, and for N <= 148 x 8192 the keys never leave the SM between A and B2.
// Only the live bits are sorted: the digits come from `key - min` (min / max of the real keys
// arrive from the projection kernel), 9 bits per pass, so a scene whose far/near ratio is below
// 2^8 sorts in THREE passes (27 live bits) and most synthetic ones (C3: 24 bits) too.  The first
// pass makes up the ids (no identity array is read), the last pass writes only the ids.
// Algorithmic HBM bytes per Gaussian: 4 + (passes - 1) * 16 + 4.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace jsb {
namespace {

// threads per CTA: the register file holds one CTA per SM either way (keys, ids and ranks of a tile
// live in registers), so more threads x fewer items each = more warps to hide the ranking chain
#ifndef JSB_DS_THREADS
#define JSB_DS_THREADS 512
#endif
constexpr int DS_THREADS = JSB_DS_THREADS;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_TILE = 8192;                    // keys per tile, at most
constexpr int DS_IPT = DS_TILE / DS_THREADS;     // items per thread and tile, at most
constexpr int DS_BITS = 9;
constexpr int DS_BINS = 1 << DS_BITS;            // thread d < DS_BINS owns digit d
static_assert(DS_THREADS >= DS_BINS && DS_THREADS % DS_BINS == 0, "one thread per digit, at least");
constexpr unsigned kSentinelKey = 0xFFFFFFFFu;   // Gaussians that hit no tile
constexpr int DS_MAXG = 256;                     // CTAs (= SMs used) at most
constexpr int DS_MAXK = 16;                      // digits one CTA scans at most (grid > 32)

struct DsParams {
    const unsigned *keys_in;  // depth keys as the projection wrote them
    unsigned *kA, *vA, *kB, *vB;  // ping-pong (kA may be keys_in itself)
    unsigned *order;          // out: Gaussian ids in (depth, index) order
    unsigned n;
    const unsigned *mm;       // {max(~key), max(key)} over the real keys; {0, 0} = none
    unsigned *hist;           // [grid][DS_BINS] digit counts of every CTA
    unsigned *pre;            // [grid][DS_BINS] their exclusive prefix over the CTAs
    unsigned *tot;            // [DS_BINS] digit totals
    unsigned *bar;            // [0] grid barrier counter, [1] CTAs that are through: both zero at
                              // launch, and zero again when the last CTA leaves (no memset)
    unsigned *ctrl;           // 16 control words the last CTA out clears for the caller, or null
    int *err;
    unsigned tile_items;      // multiple of DS_THREADS, <= DS_TILE
    unsigned subtiles;        // tiles per CTA
};

struct DsSmem {
    unsigned whist[DS_WARPS][DS_BINS];  // per-warp digit counts -> exclusive over the warps
    unsigned keys[DS_TILE];
    unsigned vals[DS_TILE];
    unsigned binstart[DS_BINS];         // exclusive scan of the tile's digit counts
    unsigned goff[DS_BINS];             // global offset of the digit's run minus binstart
    unsigned colcnt[DS_MAXK][DS_MAXG];  // B1: counts of the digits this CTA scans, per CTA
    unsigned ws[DS_WARPS + 1];
};

// (relaxed: an acquire load compiles to LDG + CCTL.IVALL, i.e. every poll of a spin loop would
// invalidate the whole L1; one fence after the loop does it once)
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the (co-resident) grid; `target` = arrivals after which this barrier opens
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned target, int *err) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        int spins = 0;
        while (ld_relaxed_gpu(bar) < target) {
            if (++spins > kSpinLimit) { *err = 3; break; }  // a scheduling bug becomes an error
        }
        __threadfence();
    }
    __syncthreads();
}

// how a pass reads its digit from a re-based key (the sentinel stays 0xFFFFFFFF: all-ones in
// every digit, so it sorts last)
struct Digit {
    int shift;
    unsigned mask;
    int w;      // live bits (ballots)
    __device__ __forceinline__ unsigned operator()(unsigned k) const { return (k >> shift) & mask; }
};

__global__ void __launch_bounds__(DS_THREADS, 1) resident_depth_sort_kernel(const DsParams P) {
    extern __shared__ __align__(16) unsigned char ds_raw[];
    DsSmem &S = *reinterpret_cast<DsSmem *>(ds_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned G = gridDim.x, c = blockIdx.x;

    // the last CTA out (every CTA has read `mm` and left every barrier by then) puts the barrier
    // words back to zero and clears the caller's control words for its next binning
    auto leave = [&]() {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(P.bar + 1, 1u) == G - 1u) {
                P.bar[0] = 0u;
                P.bar[1] = 0u;
                if (P.ctrl)
                    for (int i = 0; i < 16; ++i) P.ctrl[i] = 0u;
            }
        }
    };
    const unsigned inv_min = __ldg(P.mm), kmax = __ldg(P.mm + 1);
    const unsigned kmin = ~inv_min;
    const unsigned span = (kmax >= kmin) ? kmax - kmin : 0u;
    const int bits = span ? 32 - __clz(span) : 0;
    const int passes = (bits + DS_BITS - 1) / DS_BITS;
    const unsigned cta_items = P.tile_items * P.subtiles;
    if (passes == 0) {  // one depth (or no real key at all): index order is the stable order
        const unsigned b0 = c * cta_items;
        for (unsigned i = tid; i < cta_items && b0 + i < P.n; i += DS_THREADS) P.order[b0 + i] = b0 + i;
        leave();
        return;
    }
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool single = P.subtiles == 1;
    unsigned arrivals = 0;

    unsigned key[DS_IPT], val[DS_IPT];
    unsigned short rank[DS_IPT];
    unsigned cta_count = 0, binstart = 0;

    // `count` elements starting at `base` into registers, warp-striped over ipt items per thread
    // (memory order == (warp, item, lane) order).  first: the keys are the projection's (re-base
    // them, make up the ids).  Every load of the thread is in flight before the first use.
    auto load_items = [&](const unsigned *ksrc, const unsigned *vsrc, unsigned base, unsigned count,
                          unsigned ipt, bool first, bool with_vals) {
        const unsigned wbase = warp * (ipt * 32u) + lane;
#pragma unroll
        for (int i = 0; i < DS_IPT; ++i) {
            const unsigned o = wbase + i * 32u;
            const bool have = (unsigned)i < ipt && o < count;
            key[i] = have ? __ldcg(ksrc + base + o) : 0u;
            if (with_vals && !first) val[i] = have ? __ldcg(vsrc + base + o) : 0u;
        }
        if (first) {
#pragma unroll
            for (int i = 0; i < DS_IPT; ++i) {
                key[i] = (key[i] == kSentinelKey) ? kSentinelKey : key[i] - kmin;  // live bits only
                val[i] = base + wbase + i * 32u;
            }
        }
    };
    // stable rank of every item inside its warp's chunk + per-warp digit counts; leaves the
    // exclusive-over-warps counts in S.whist, the tile's digit count of digit `tid` in
    // cta_count and its exclusive scan over the digits in binstart / S.binstart
    auto rank_tile = [&](unsigned count, unsigned ipt, const Digit dg) {
        for (unsigned i = tid; i < DS_WARPS * DS_BINS; i += DS_THREADS) (&S.whist[0][0])[i] = 0u;
        __syncthreads();
        unsigned *wh = S.whist[warp];
        const unsigned wbase = warp * (ipt * 32u) + lane;
#pragma unroll
        for (int i = 0; i < DS_IPT; ++i) {
            if ((unsigned)i >= ipt) break;  // warp-uniform
            const bool valid = (wbase + i * 32u) < count;
            const unsigned d = valid ? dg(key[i]) : 0u;
            // peers = lanes holding the same digit: one ballot per live digit bit (the native
            // MATCH.ANY is data dependent on B200 and slower when most lanes differ)
            unsigned peers = __ballot_sync(0xffffffffu, valid);
            peers = valid ? peers : ~peers;
#pragma unroll
            for (int b = 0; b < DS_BITS; ++b) {
                if (b < dg.w) {
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
        cta_count = 0;
        if (tid < (unsigned)DS_BINS) {
#pragma unroll
            for (int ww = 0; ww < DS_WARPS; ++ww) {
                const unsigned cc = S.whist[ww][tid];
                S.whist[ww][tid] = cta_count;
                cta_count += cc;
            }
        }
        unsigned total;
        binstart = cta_exclusive_scan<DS_THREADS>(cta_count, &total, S.ws);
        if (tid < (unsigned)DS_BINS) S.binstart[tid] = binstart;
        __syncthreads();
    };
    // tile in registers -> shared memory, grouped by digit, in stable order
    auto reorder_tile = [&](unsigned count, unsigned ipt, const Digit dg) {
        const unsigned *wh = S.whist[warp];
        const unsigned wbase = warp * (ipt * 32u) + lane;
#pragma unroll
        for (int i = 0; i < DS_IPT; ++i) {
            if ((unsigned)i < ipt && (wbase + i * 32u) < count) {
                const unsigned d = dg(key[i]);
                const unsigned pos = S.binstart[d] + wh[d] + rank[i];
                S.keys[pos] = key[i];
                S.vals[pos] = val[i];
            }
        }
    };

    // One GLOBAL pass: the keys of (ksrc, vsrc) leave grouped by digit, stable, into (kdst, vdst).
    auto global_pass = [&](const Digit dg, const unsigned *ksrc, const unsigned *vsrc, unsigned *kdst,
                           unsigned *vdst, bool first, bool write_keys) {
        const unsigned ipt = P.tile_items / DS_THREADS;
        auto tile_count = [&](unsigned tile) -> unsigned {
            const unsigned base = tile * P.tile_items;
            return base < P.n ? min(P.tile_items, P.n - base) : 0u;
        };
        // ---- A: digit counts of this CTA's tiles ------------------------------------------------
        unsigned count = 0, cta_total = 0;
        for (unsigned st = 0; st < P.subtiles; ++st) {
            const unsigned tile = c * P.subtiles + st;
            count = tile_count(tile);
            load_items(ksrc, vsrc, tile * P.tile_items, count, ipt, first, single);
            rank_tile(count, ipt, dg);
            cta_total += cta_count;
        }
        if (tid < (unsigned)DS_BINS) P.hist[(size_t)c * DS_BINS + tid] = cta_total;
        // (single tile: the reorder needs nothing from the other CTAs -- done under the barrier)
        if (single) reorder_tile(count, ipt, dg);
        arrivals += G;
        grid_barrier(P.bar, arrivals, P.err);

        // ---- B: offsets, scatter ------------------------------------------------------------------
        unsigned before = 0, all = 0;
        if (G <= 32u) {
            // a small grid: every CTA sums the rows before it itself (at most 32 loads a thread)
            for (unsigned r = 0; r < G && tid < (unsigned)DS_BINS; ++r) {
                const unsigned h = __ldcg(P.hist + (size_t)r * DS_BINS + tid);
                all += h;
                if (r < c) before += h;
            }
        } else {
            // B1: this CTA scans digits [c K, c K + K) over the G CTAs (K = ceil(512 / G) <= 16)
            {
                const unsigned K = (DS_BINS + G - 1) / G;
                const unsigned tpd = DS_THREADS / DS_MAXK;  // threads per digit while loading
                const unsigned k = tid / tpd, r0 = tid % tpd;
                const unsigned d = c * K + k;
                if (k < K && d < DS_BINS)
                    for (unsigned r = r0; r < G; r += tpd) S.colcnt[k][r] = __ldcg(P.hist + (size_t)r * DS_BINS + d);
                __syncthreads();
                if (warp < K && c * K + warp < DS_BINS) {  // one warp per digit: lane owns a run of rows
                    const unsigned dd = c * K + warp;
                    const unsigned per = (G + 31u) / 32u;
                    const unsigned b0 = min(G, lane * per), b1 = min(G, b0 + per);
                    unsigned sum = 0;
                    for (unsigned r = b0; r < b1; ++r) sum += S.colcnt[warp][r];
                    unsigned incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned tt = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= (unsigned)o) incl += tt;
                    }
                    unsigned runp = incl - sum;
                    for (unsigned r = b0; r < b1; ++r) {
                        P.pre[(size_t)r * DS_BINS + dd] = runp;
                        runp += S.colcnt[warp][r];
                    }
                    if (lane == 31) P.tot[dd] = incl;
                }
            }
            arrivals += G;
            grid_barrier(P.bar, arrivals, P.err);
            // B2: own row of prefixes + digit totals
            if (tid < (unsigned)DS_BINS) {
                before = __ldcg(P.pre + (size_t)c * DS_BINS + tid);
                all = __ldcg(P.tot + tid);
            }
        }
        unsigned grand;
        const unsigned dbase = cta_exclusive_scan<DS_THREADS>(all, &grand, S.ws);
        unsigned run = dbase + before;  // next global slot of digit `tid` for this CTA
        for (unsigned st = 0; st < P.subtiles; ++st) {
            if (!single) {
                const unsigned tile = c * P.subtiles + st;
                count = tile_count(tile);
                load_items(ksrc, vsrc, tile * P.tile_items, count, ipt, first, true);
                rank_tile(count, ipt, dg);
                reorder_tile(count, ipt, dg);
            }
            if (tid < (unsigned)DS_BINS) S.goff[tid] = run - binstart;
            run += cta_count;
            __syncthreads();
            // consecutive positions of one digit go to consecutive addresses
            for (unsigned pos = tid; pos < count; pos += DS_THREADS) {
                const unsigned k = S.keys[pos];
                const unsigned o = pos + S.goff[dg(k)];
                if (write_keys) kdst[o] = k;
                vdst[o] = S.vals[pos];
            }
            __syncthreads();
        }
    };

    // ---- LSD: one global pass per 9 live bits ------------------------------------------------------
    const int width = (bits + passes - 1) / passes;  // <= 9 live bits per pass
    for (int ps = 0; ps < passes; ++ps) {
        const int shift = ps * width;
        const int w = min(width, bits - shift);
        const Digit dg{shift, (1u << w) - 1u, w};
        const bool last = ps + 1 == passes;
        const unsigned *ksrc = ps == 0 ? P.keys_in : ((ps & 1) ? P.kB : P.kA);
        const unsigned *vsrc = (ps & 1) ? P.vB : P.vA;
        unsigned *kdst = (ps & 1) ? P.kA : P.kB, *vdst = last ? P.order : ((ps & 1) ? P.vA : P.vB);
        global_pass(dg, ksrc, vsrc, kdst, vdst, ps == 0, !last);
        if (!last) {
            arrivals += G;
            grid_barrier(P.bar, arrivals, P.err);
        }
    }
    leave();
}

struct DeviceFacts {
    bool known = false, cooperative = false;
    int max_ctas = 0;
};
DeviceFacts g_facts[64];

}  // namespace

size_t resident_sort_scratch_bytes(unsigned n) {
    (void)n;
    return (size_t)(2 * DS_MAXG + 2) * DS_BINS * sizeof(unsigned);  // hist + pre rows, totals
}

// returns 0 on success, 1 on error, -1 if this device cannot run it (caller falls back)
int resident_depth_sort(cudaStream_t stream, const unsigned *keys_in, unsigned *kA, unsigned *vA,
                        unsigned *kB, unsigned *vB, unsigned *order, unsigned n,
                        const unsigned *mm, void *scratch, int *err, unsigned *barrier_words,
                        unsigned *clear_words) {
    if (n == 0) return 0;
    int dev = 0;
    JSB_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return -1;
    DeviceFacts &F = g_facts[dev];
    if (!F.known) {
        int coop = 0, sms = 0, per_sm = 0;
        JSB_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        JSB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        JSB_CUDA_OK(cudaFuncSetAttribute(resident_depth_sort_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(DsSmem)));
        JSB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, resident_depth_sort_kernel, DS_THREADS, sizeof(DsSmem)));
        F.cooperative = coop != 0 && per_sm > 0;
        F.max_ctas = std::min(sms, DS_MAXG);  // one CTA per SM is all the parallelism a pass needs
        F.known = true;
    }
    if (!F.cooperative) return -1;
    DsParams P;
    P.keys_in = keys_in; P.kA = kA; P.vA = vA; P.kB = kB; P.vB = vB; P.order = order;
    P.n = n; P.mm = mm; P.err = err;
    P.hist = reinterpret_cast<unsigned *>(scratch);
    P.pre = P.hist + (size_t)DS_MAXG * DS_BINS;
    P.tot = P.pre + (size_t)DS_MAXG * DS_BINS;
    P.bar = barrier_words;
    P.ctrl = clear_words;
    const unsigned G0 = (unsigned)F.max_ctas;
    P.subtiles = std::max(1u, ceil_div(n, G0 * (unsigned)DS_TILE));
    const unsigned per_tile = ceil_div(n, G0 * P.subtiles);
    P.tile_items = std::min((unsigned)DS_TILE, std::max((unsigned)DS_THREADS, ceil_div(per_tile, (unsigned)DS_THREADS) * (unsigned)DS_THREADS));
    const unsigned grid = ceil_div(n, P.tile_items * P.subtiles);
    void *args[] = {(void *)&P};
    JSB_CUDA_OK(cudaLaunchCooperativeKernel((const void *)resident_depth_sort_kernel, dim3(grid),
                                            dim3(DS_THREADS), args, sizeof(DsSmem), stream));
    note_launch();
    return 0;
}

}  // namespace jsb
