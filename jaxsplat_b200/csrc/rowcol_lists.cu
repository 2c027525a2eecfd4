// Per-tile depth-ordered Gaussian lists for sm_100a by a two-level in-order distribution:
// Gaussians -> tile rows -> tiles.  Nothing I-scale is ever written except the final list.
//
// Behavioural spec (same as tile_lists.cu): reference lib/kernels/forward.cu:135-179
// (map_gaussian_to_intersects), lib/ops.cu:722-757 (SortPairs on tile<<32|depth) and
// forward.cu:184-208 (get_tile_bin_edges): per tile, the ids of the Gaussians whose tile box
// covers it in ascending (depth bits, index) order, tiles concatenated in tile-id order, plus
// every tile's (first, one-past-last) range.
//
// Design (not a port).  The Gaussians arrive in (depth, index) order.  A tile's list is the
// subsequence of that order whose boxes cover it, so the list is a *distribution*, not a sort:
//   level 1  each Gaussian drops one 8-byte entry {id, x0 | w << 16} into every tile ROW of its
//            box: R row lists, each still in depth order (sum of box heights entries -- about
//            a quarter of the intersections);
//   level 2  each row entry drops its id into the tiles x0 .. x0+w-1 of that row.
// Both levels are the same single-warp kernel: a chunk of <= 256 items is ranked all at once --
// the rank of an item in each bin it covers comes from per-bin lane masks (popc of the earlier
// groups' masks + popc of the lower lanes of its own group = item order), so every bin's run is
// in item order by construction, with no serial walk over the items.  The ids land in a shared-memory staging area laid
// out bin by bin, and the staged chunk leaves as full coalesced runs -- the previous
// design (tile_lists.cu) issued one 4-byte partial-sector write per intersection and was bound by
// the L2's partial-write rate.  Run starts come from per-(chunk, bin) counts (difference arrays
// in shared memory) and a column scan; the count matrices are (chunks x rows) and
// (row chunks x tile columns): a few MB instead of the (chunks x tiles) matrix of tile_lists.cu.
// HBM traffic per intersection: the 4-byte id, once.  Per row entry: 8 B written, 8 B read twice.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace jsb {
namespace {

// register groups of 32 items per chunk and staged outputs per chunk, level 1 / level 2: smaller
// chunks halve a CTA's shared memory (the single-warp scatter CTAs are limited by it: 13-14 per SM
// at 8 groups) at the price of more fixed work per output
#ifndef JSB_RC_GROUPS1
#define JSB_RC_GROUPS1 8
#endif
#ifndef JSB_RC_GROUPS2
#define JSB_RC_GROUPS2 4
#endif
#ifndef JSB_RC_STAGE1
#define JSB_RC_STAGE1 (128 * JSB_RC_GROUPS1)
#endif
#ifndef JSB_RC_STAGE2
#define JSB_RC_STAGE2 (192 * JSB_RC_GROUPS2)
#endif
// super-column level 2 (fused render): 8-byte entries into a few wide bins per row
#ifndef JSB_RC_GROUPS2S
#define JSB_RC_GROUPS2S 8
#endif
#ifndef JSB_RC_STAGE2S
#define JSB_RC_STAGE2S (80 * JSB_RC_GROUPS2S)
#endif
constexpr int RC_G1 = JSB_RC_GROUPS1, RC_G2 = JSB_RC_GROUPS2, RC_G2S = JSB_RC_GROUPS2S;
constexpr int RC_ITEMS = 256;    // items per chunk of the COUNT kernels (8 register groups of 32)
constexpr int RC_STAGE1 = JSB_RC_STAGE1;  // staged 8-byte row entries per level-1 chunk
constexpr int RC_STAGE2 = JSB_RC_STAGE2;  // staged 4-byte ids per level-2 chunk
constexpr int RC_STAGE2S = JSB_RC_STAGE2S;  // staged 8-byte entries per super-column level-2 chunk

struct RcParams {
    unsigned n;          // Gaussians
    unsigned rows;       // R: tile rows of the band
    unsigned cols;       // X: tiles per row
    unsigned sc_shift;   // level-2 bins are super-columns of 2^sc_shift tiles (0: the tiles themselves)
    unsigned cols2;      // level-2 bins per row = ceil(cols / 2^sc_shift)
    unsigned band_y0;    // first tile row of the band
    unsigned k1;         // depth positions per level-1 chunk (multiple of 32, <= 256)
    unsigned k2;         // row entries per level-2 chunk (multiple of 32, <= 256)
    unsigned n1;         // level-1 chunks = ceil(n / k1)
    unsigned n1s;        // row stride of the level-1 count matrix (n1 rounded up to 4)
    unsigned n2max;      // upper bound of the level-2 chunk count
};

// word offsets inside the scratch buffer
struct RcLayout {
    size_t l1m;         // [rows][n1s]   level-1 counts -> exclusive prefix per row
    size_t row_tot;     // [rows]
    size_t row_start;   // [rows + 1]
    size_t chunk_base;  // [rows + 1]
    size_t chunk_row;   // uint4 [n2max]: {row, first entry, one past last entry, 0} of every level-2 chunk
    size_t tile_cnt;    // [rows * cols2] -> exclusive scan (bin bases)
    size_t m;           // [n2max][cols2] level-2 counts -> exclusive prefix per (row, bin)
    size_t entries;     // uint2 [entries] (8-byte aligned)
    size_t end;
};

__host__ __device__ inline size_t up4(size_t v) { return (v + 3) & ~(size_t)3; }

static RcLayout rc_layout(const RcParams &p, unsigned long long row_entries) {
    RcLayout L;
    size_t o = 0;
    L.l1m = o;        o += up4((size_t)p.rows * p.n1s);
    L.row_tot = o;    o += up4(p.rows);
    L.row_start = o;  o += up4(p.rows + 1);
    L.chunk_base = o; o += up4(p.rows + 1);
    L.chunk_row = o;  o += 4 * (size_t)p.n2max;
    L.tile_cnt = o;   o += up4((size_t)p.rows * p.cols2 + 1);
    L.m = o;          o += up4((size_t)p.n2max * p.cols2);
    L.entries = o;    o += up4(2 * (size_t)row_entries + 4);
    L.end = o;
    return L;
}

__device__ __forceinline__ unsigned warp_inclusive_scan(unsigned v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// band-clamped box of the Gaussian at depth position `pos`: rows relative to the band
__device__ __forceinline__ bool load_box(const RcParams &p, unsigned pos,
                                         const unsigned *__restrict__ order,
                                         const uint2 *__restrict__ box_by_id, unsigned &g,
                                         unsigned &y0, unsigned &h, unsigned &xw) {
    g = 0; y0 = 0; h = 0; xw = 0;
    if (pos >= p.n) return false;
    g = order[pos];
    const uint2 b = __ldg(box_by_id + g);
    const unsigned w = b.y & 0xffffu;
    h = b.y >> 16;
    if (w == 0 || h == 0) { h = 0; return false; }
    y0 = (b.x >> 16) - p.band_y0;
    xw = (b.x & 0xffffu) | (w << 16);
    return true;
}

// ---- level 1 counts: rows covered by each chunk of k1 depth positions ----------------------------
// One warp per chunk (8 chunks per CTA, no CTA-wide barrier), 8 positions per lane.
constexpr int RC_RC_WARPS = 8;
__global__ void __launch_bounds__(32 * RC_RC_WARPS)
rc_row_count_kernel(const RcParams p, const unsigned *__restrict__ order,
                    const uint2 *__restrict__ box_by_id, unsigned *__restrict__ l1m) {
    __shared__ unsigned s_diff_all[RC_RC_WARPS][256 + 32];
    chain_wait();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *s_diff = s_diff_all[warp];
    const unsigned chunk = blockIdx.x * RC_RC_WARPS + warp;
    if (chunk >= p.n1) return;  // warp-uniform; no CTA barrier below
    for (unsigned i = lane; i <= p.rows; i += 32) s_diff[i] = 0u;
    unsigned y0[8], h[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        unsigned g, xw;
        const unsigned item = q * 32 + lane;
        const bool ok = item < p.k1 && load_box(p, chunk * p.k1 + item, order, box_by_id, g, y0[q], h[q], xw);
        if (!ok) h[q] = 0u;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (h[q]) {
            atomicAdd(&s_diff[y0[q]], 1u);
            atomicAdd(&s_diff[y0[q] + h[q]], 0xffffffffu);  // -1 (mod 2^32)
        }
    }
    __syncwarp();
    unsigned carry = 0;
    for (unsigned base = 0; base < p.rows; base += 32) {
        const unsigned i = base + lane;
        const unsigned cnt = warp_inclusive_scan(i < p.rows ? s_diff[i] : 0u, lane) + carry;
        carry = __shfl_sync(0xffffffffu, cnt, 31);
        if (i < p.rows) l1m[(size_t)i * p.n1s + chunk] = cnt;  // entries of this chunk in row i
    }
}

// In-place exclusive scan of n words by one CTA of 32 warps: warp w owns the contiguous slice
// [w * per, (w + 1) * per) and walks it 32 words at a time (coalesced, loads of the next strip
// independent of the running prefix), so the CTA synchronises only once; `each(i, base, v)` sees
// every element's exclusive prefix and value.  Returns the total (all threads).
template <typename F>
__device__ __forceinline__ unsigned cta32_exclusive_scan_inplace(unsigned *data, unsigned n, F each) {
    __shared__ unsigned s_wtot[32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned per = ((n + 31u) / 32u + 31u) & ~31u;  // slice length, multiple of 32
    const unsigned i0 = min(n, warp * per), i1 = min(n, i0 + per);
    unsigned sum = 0;
#pragma unroll 4
    for (unsigned i = i0 + lane; i < i1; i += 32) sum += data[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_wtot[warp] = sum;
    __syncthreads();
    unsigned run = 0, total = 0;
#pragma unroll
    for (unsigned w = 0; w < 32; ++w) {
        const unsigned v = s_wtot[w];
        if (w < warp) run += v;
        total += v;
    }
    for (unsigned b0 = i0; b0 < i1; b0 += 32) {
        const unsigned i = b0 + lane;
        const unsigned v = i < i1 ? data[i] : 0u;
        const unsigned incl = warp_inclusive_scan(v, lane);
        if (i < i1) {
            data[i] = run + incl - v;
            each(i, run + incl - v, v);
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    return total;
}

// ---- in-place exclusive scan of every row of the level-1 matrix (one CTA per row) -----------------
__global__ void __launch_bounds__(1024)
rc_row_scan_kernel(const RcParams p, unsigned *__restrict__ l1m, unsigned *__restrict__ row_tot) {
    chain_wait();
    unsigned *row = l1m + (size_t)blockIdx.x * p.n1s;
    const unsigned total = cta32_exclusive_scan_inplace(row, p.n1, [](unsigned, unsigned, unsigned) {});
    if (threadIdx.x == 0) row_tot[blockIdx.x] = total;
}

// ---- row starts, level-2 chunk bases and {row, entry range} of every level-2 chunk ----------------
// Every CTA rebuilds the two <= 256-entry scans from the row totals (cheaper than a launch of its
// own); CTA 0 publishes them for the later kernels.  The per-chunk record spares the level-2
// kernels a chain of three dependent global loads before they can touch their entries.
__global__ void __launch_bounds__(256)
rc_chunk_rows_kernel(const RcParams p, const unsigned *__restrict__ row_tot,
                     unsigned *__restrict__ row_start, unsigned *__restrict__ chunk_base,
                     uint4 *__restrict__ chunk_info) {
    __shared__ unsigned s_base[256 + 1], s_start[256 + 1];
    __shared__ unsigned s_ws[256 / 32 + 1];
    const unsigned tid = threadIdx.x;
    chain_wait();
    const unsigned n = tid < p.rows ? row_tot[tid] : 0u;
    unsigned total;
    unsigned e = cta_exclusive_scan<256>(n, &total, s_ws);
    if (tid < p.rows) s_start[tid] = e;
    if (tid == 0) s_start[p.rows] = total;
    if (blockIdx.x == 0) {
        if (tid < p.rows) row_start[tid] = e;
        if (tid == 0) row_start[p.rows] = total;
    }
    __syncthreads();
    const unsigned c = (n + p.k2 - 1) / p.k2;
    e = cta_exclusive_scan<256>(c, &total, s_ws);
    if (tid < p.rows) s_base[tid] = e;
    if (tid == 0) s_base[p.rows] = total;
    if (blockIdx.x == 0) {
        if (tid < p.rows) chunk_base[tid] = e;
        if (tid == 0) chunk_base[p.rows] = total;
    }
    __syncthreads();
    const unsigned ch = blockIdx.x * 256u + tid;
    if (ch >= p.n2max) return;
    uint4 info = make_uint4(0u, 0u, 0u, 0u);  // chunks beyond the total: empty range
    if (ch < s_base[p.rows]) {
        unsigned lo = 0, hi = p.rows;  // invariant: s_base[lo] <= ch < s_base[hi]
        while (hi - lo > 1) {
            const unsigned mid = (lo + hi) >> 1;
            if (s_base[mid] <= ch) lo = mid; else hi = mid;
        }
        const unsigned ebeg = s_start[lo] + (ch - s_base[lo]) * p.k2;
        info = make_uint4(lo, ebeg, min(s_start[lo + 1], ebeg + p.k2), 0u);
    }
    chunk_info[ch] = info;
}

// level-2 bins [lo, lo + len) covered by a row entry's tile columns x0 .. x0 + w - 1
__device__ __forceinline__ void entry_bins(unsigned xw, unsigned shift, unsigned &lo, unsigned &len) {
    const unsigned x0 = xw & 0xffffu, w = xw >> 16;
    lo = x0 >> shift;
    len = w ? ((x0 + w - 1u) >> shift) - lo + 1u : 0u;
}

// ---- level 2 counts: tile columns covered by each chunk of k2 row entries -------------------------
// One warp per chunk (8 chunks per CTA, no CTA-wide barrier): difference array in the warp's own
// shared memory, 8 entries per lane, then a prefix over the columns.
constexpr int RC_CC_WARPS = 8;
__global__ void __launch_bounds__(32 * RC_CC_WARPS)
rc_col_count_kernel(const RcParams p, const uint4 *__restrict__ chunk_info,
                    const uint2 *__restrict__ entries, unsigned *__restrict__ m) {
    __shared__ unsigned s_diff_all[RC_CC_WARPS][512 + 32];
    chain_wait();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *s_diff = s_diff_all[warp];
    const unsigned c = blockIdx.x * RC_CC_WARPS + warp;
    if (c >= p.n2max) return;  // warp-uniform; no CTA barrier below
    const uint4 info = __ldg(chunk_info + c);
    const unsigned ebeg = info.y, eend = info.z;
    if (ebeg >= eend) return;
    for (unsigned i = lane; i <= p.cols2; i += 32) s_diff[i] = 0u;
    unsigned xw[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const unsigned e = ebeg + q * 32 + lane;
        xw[q] = (q * 32 + lane < p.k2 && e < eend) ? entries[e].y : 0u;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        unsigned lo, len;
        entry_bins(xw[q], p.sc_shift, lo, len);
        if (len) {
            atomicAdd(&s_diff[lo], 1u);
            atomicAdd(&s_diff[lo + len], 0xffffffffu);
        }
    }
    __syncwarp();
    unsigned *mrow = m + (size_t)c * p.cols2;
    unsigned carry = 0;
    for (unsigned base = 0; base < p.cols2; base += 32) {
        const unsigned i = base + lane;
        const unsigned cnt = warp_inclusive_scan(i < p.cols2 ? s_diff[i] : 0u, lane) + carry;
        carry = __shfl_sync(0xffffffffu, cnt, 31);
        if (i < p.cols2) mrow[i] = cnt;
    }
}

// ---- per (row, column): exclusive scan over the row's chunks in place; tile totals out ------------
// CTA = (row, 32 columns); its 32 warps split the row's chunks, lanes = columns.
constexpr int RC_CS_WARPS = 32;
__global__ void __launch_bounds__(32 * RC_CS_WARPS)
rc_col_scan_kernel(const RcParams p, const unsigned *__restrict__ chunk_base,
                   unsigned *__restrict__ m, unsigned *__restrict__ tile_cnt) {
    __shared__ unsigned s_part[RC_CS_WARPS][32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned r = blockIdx.y, col = blockIdx.x * 32u + lane;
    chain_wait();
    const unsigned cb = chunk_base[r], ce = chunk_base[r + 1];
    const unsigned per = (ce - cb + RC_CS_WARPS - 1) / RC_CS_WARPS;
    const unsigned c0 = min(ce, cb + warp * per), c1 = min(ce, c0 + per);
    const bool live = col < p.cols2;
    unsigned sum = 0;
    if (live) {
#pragma unroll 8
        for (unsigned c = c0; c < c1; ++c) sum += m[(size_t)c * p.cols2 + col];
    }
    s_part[warp][lane] = sum;
    __syncthreads();
    unsigned run = 0, total = 0;
#pragma unroll
    for (unsigned w = 0; w < RC_CS_WARPS; ++w) {
        const unsigned v = s_part[w][lane];
        if (w < warp) run += v;
        total += v;
    }
    if (live) {
        constexpr int U = 8;
        for (unsigned c = c0; c < c1; c += U) {
            unsigned v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = (c + u < c1) ? m[(size_t)(c + u) * p.cols2 + col] : 0u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (c + u < c1) m[(size_t)(c + u) * p.cols2 + col] = run;
                run += v[u];
            }
        }
        if (warp == 0) tile_cnt[(size_t)r * p.cols2 + col] = total;
    }
}

// ---- tile counts -> tile bases (in place) and tile bins (one CTA) ----------------------------------
__global__ void __launch_bounds__(1024)
rc_tile_scan_bins_kernel(const RcParams p, unsigned *__restrict__ tile_cnt,
                         int2 *__restrict__ tile_bins) {
    chain_wait();
    // (super-column bins are a band-local array of the library's own; tile bins index the frame)
    int2 *bins = p.sc_shift ? tile_bins : tile_bins + (size_t)p.band_y0 * p.cols;
    cta32_exclusive_scan_inplace(tile_cnt, p.rows * p.cols2, [bins](unsigned t, unsigned base, unsigned v) {
        // tiles without intersections stay (0,0) like the reference's memset + edge detection
        bins[t] = v ? make_int2((int)base, (int)(base + v)) : make_int2(0, 0);
    });
}

// ---- the in-order staged distribution of one chunk (both levels) -----------------------------------
// Items live in registers: group q (0..7), lane l = item 32 q + l of the chunk, covering bins
// [lo, lo + len) with payload `pay`; len == 0 marks an absent item.  Order-preserving placement
// without walking the items one by one:
//   1. every item ORs its lane bit into mask[q][b] of each bin b it covers;
//   2. per bin: pre[q][b] = items of groups < q covering b, count[b] = all of them; an exclusive
//      scan of count over the bins lays the chunk's output out bin by bin (off[b]);
//   3. item (q, lane) owns slot off[b] + pre[q][b] + popc(mask[q][b] & lower lanes) of bin b --
//      its rank in item order -- and drops its payload there: into the shared staging area, or
//      straight to global memory for slots beyond the staging capacity (correct, just slower);
//   4. the staged slots leave in slot order = run by run, as coalesced stores.
// Shared memory per warp (nbp = bins + 1 rounded up to 32):
//   mask u32[8][nbp] | off u32[nbp] | gbase u32[nbp] | stage Payload[STAGE] | pre u16[8][nbp] |
//   bin u16[STAGE]
template <typename Payload, int STAGE, int GROUPS>
struct DistSmem {
    Payload *stage;
    unsigned *mask, *off, *gbase;
    unsigned short *pre, *bin;
    unsigned nbp;
    __device__ __forceinline__ DistSmem(unsigned char *base, unsigned nb) {
        nbp = (nb + 1 + 31) & ~31u;
        stage = reinterpret_cast<Payload *>(base);
        mask = reinterpret_cast<unsigned *>(base + (size_t)STAGE * sizeof(Payload));
        off = mask + GROUPS * nbp;
        gbase = off + nbp;
        pre = reinterpret_cast<unsigned short *>(gbase + nbp);
        bin = pre + GROUPS * nbp;
    }
    static size_t bytes(unsigned nb) {
        const size_t nbp = (nb + 1 + 31) & ~31u;
        return (size_t)STAGE * sizeof(Payload) + (GROUPS + 2) * nbp * 4 + GROUPS * nbp * 2 + (size_t)STAGE * 2;
    }
};

// steps 1 + 2; returns the number of (item, bin) pairs of the chunk (warp-uniform)
template <typename Payload, int STAGE, int GROUPS>
__device__ __forceinline__ unsigned distribute_rank(const unsigned (&lo)[GROUPS], const unsigned (&len)[GROUPS],
                                                    unsigned nb, const DistSmem<Payload, STAGE, GROUPS> &S) {
    const unsigned lane = threadIdx.x, mine = 1u << lane, nbp = S.nbp;
    uint4 *mz = reinterpret_cast<uint4 *>(S.mask);
    for (unsigned i = lane; i < (GROUPS / 4) * nbp; i += 32) mz[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < GROUPS; ++q)
        for (unsigned k = 0; k < len[q]; ++k) atomicOr(&S.mask[q * nbp + lo[q] + k], mine);
    __syncwarp();
    unsigned carry = 0;
    for (unsigned base = 0; base < nb; base += 32) {
        const unsigned b = base + lane;
        unsigned run = 0;
        if (b < nb) {
#pragma unroll
            for (int q = 0; q < GROUPS; ++q) {
                S.pre[q * nbp + b] = (unsigned short)run;
                run += __popc(S.mask[q * nbp + b]);
            }
        }
        const unsigned incl = warp_inclusive_scan(run, lane);
        if (b < nb) S.off[b] = incl - run + carry;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    return carry;
}

// steps 3 + 4 (S.gbase[b] = global index of the first element of the (chunk, bin) run)
template <typename Payload, int STAGE, int GROUPS>
__device__ __forceinline__ void distribute_place(const unsigned (&lo)[GROUPS], const unsigned (&len)[GROUPS],
                                                 const Payload (&pay)[GROUPS], unsigned pairs,
                                                 const DistSmem<Payload, STAGE, GROUPS> &S,
                                                 Payload *__restrict__ out) {
    const unsigned lane = threadIdx.x, below = (1u << lane) - 1u, nbp = S.nbp;
#pragma unroll
    for (int q = 0; q < GROUPS; ++q) {
        for (unsigned k = 0; k < len[q]; ++k) {
            const unsigned b = lo[q] + k;
            const unsigned rank = S.pre[q * nbp + b] + __popc(S.mask[q * nbp + b] & below);
            const unsigned slot = S.off[b] + rank;
            if (slot < (unsigned)STAGE) {
                S.stage[slot] = pay[q];
                S.bin[slot] = (unsigned short)b;
            } else {
                out[S.gbase[b] + rank] = pay[q];
            }
        }
    }
    __syncwarp();
    const unsigned staged = min(pairs, (unsigned)STAGE);
    for (unsigned s = lane; s < staged; s += 32) {
        const unsigned b = S.bin[s];
        out[S.gbase[b] + (s - S.off[b])] = S.stage[s];
    }
}

// ---- level 1: Gaussians (depth order) -> row lists ------------------------------------------------
__global__ void __launch_bounds__(32)
rc_row_scatter_kernel(const RcParams p, const unsigned *__restrict__ order,
                      const uint2 *__restrict__ box_by_id, const unsigned *__restrict__ l1m,
                      const unsigned *__restrict__ row_start, uint2 *__restrict__ entries) {
    extern __shared__ __align__(16) unsigned char rc_smem[];
    const DistSmem<uint2, RC_STAGE1, RC_G1> S(rc_smem, p.rows);
    const unsigned lane = threadIdx.x, chunk = blockIdx.x;
    // (launched right behind rc_chunk_rows_kernel, which started after the depth order and the
    // boxes were complete: a CTA that is resident while that kernel runs gathers them now and
    // waits for the kernel only in front of the run starts.  ld.global.cg, not the read-only path:
    // the compiler is free to sink invariant loads below the wait, and did)
    unsigned lo[RC_G1], len[RC_G1];
    uint2 pay[RC_G1];
    unsigned gid[RC_G1];
    bool have[RC_G1];
#pragma unroll
    for (int q = 0; q < RC_G1; ++q) {  // independent loads of the depth order ...
        const unsigned item = q * 32 + lane, pos = chunk * p.k1 + item;
        have[q] = item < p.k1 && pos < p.n;
        gid[q] = have[q] ? __ldcg(order + pos) : 0u;
    }
#pragma unroll
    for (int q = 0; q < RC_G1; ++q) {  // ... then independent gathers of the boxes
        uint2 b = make_uint2(0u, 0u);
        if (have[q]) b = __ldcg(box_by_id + gid[q]);
        const unsigned w = b.y & 0xffffu, h = b.y >> 16;
        const bool ok = w != 0u && h != 0u;
        lo[q] = ok ? (b.x >> 16) - p.band_y0 : 0u;
        len[q] = ok ? h : 0u;
        pay[q] = make_uint2(gid[q], (b.x & 0xffffu) | (w << 16));
    }
    chain_wait();
    // run starts of this chunk in every row: loads in flight before the dependent stores
    unsigned gb[8];  // rows <= 256
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned r = lane + 32u * i;
        gb[i] = r < p.rows ? __ldg(row_start + r) + __ldg(l1m + (size_t)r * p.n1s + chunk) : 0u;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned r = lane + 32u * i;
        if (r < p.rows) S.gbase[r] = gb[i];
    }
    const unsigned pairs = distribute_rank<uint2, RC_STAGE1, RC_G1>(lo, len, p.rows, S);
    if (pairs == 0) return;  // warp-uniform
    distribute_place<uint2, RC_STAGE1, RC_G1>(lo, len, pay, pairs, S, entries);
}

// ---- level 2: row entries -> per-tile id lists -----------------------------------------------------
// SUPER: the bins are super-columns of 2^sc_shift tiles and the payload is the whole 8-byte row
// entry {id, x0 | w << 16}: the rasterisers of the fused render walk a super-column's list and
// keep the entries whose tile columns cover their own tile (a third of the pairs of the exact
// lists at 8 tiles per super-column, and 16 bins per row instead of 120 to rank over).
template <bool SUPER>
__global__ void __launch_bounds__(32)
rc_col_scatter_kernel(const RcParams p, const uint4 *__restrict__ chunk_info,
                      const uint2 *__restrict__ entries, const unsigned *__restrict__ m,
                      const unsigned *__restrict__ tile_base, void *__restrict__ out_list) {
    using Payload = typename std::conditional<SUPER, uint2, unsigned>::type;
    constexpr int STAGE = SUPER ? RC_STAGE2S : RC_STAGE2;
    constexpr int G = SUPER ? RC_G2S : RC_G2;
    extern __shared__ __align__(16) unsigned char rc_smem[];
    const DistSmem<Payload, STAGE, G> S(rc_smem, p.cols2);
    const unsigned lane = threadIdx.x, c = blockIdx.x;
    // (launched right behind rc_tile_scan_bins_kernel -- one CTA -- which started after the row
    // entries, the chunk records and the run starts within the rows were complete: a CTA that is
    // resident while it runs loads those now and waits only in front of the scanned bin starts)
    const uint4 info = __ldg(chunk_info + c);
    const unsigned r = info.x, ebeg = info.y, eend = info.z;
    if (ebeg >= eend) {
        chain_wait();
        return;
    }
    const unsigned *mrow = m + (size_t)c * p.cols2;
    const unsigned *tb = tile_base + (size_t)r * p.cols2;
    constexpr int NB = SUPER ? 4 : 16;  // bins <= 512 (exact) / <= 128 (super-columns)
    unsigned gb[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const unsigned x = lane + 32u * i;
        gb[i] = x < p.cols2 ? __ldcg(mrow + x) : 0u;
    }
    unsigned lo[G], len[G];
    Payload pay[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
        const unsigned e = ebeg + q * 32 + lane;
        uint2 v = make_uint2(0u, 0u);
        if (q * 32 + lane < p.k2 && e < eend) v = __ldcg(entries + e);  // (see rc_row_scatter_kernel)
        if constexpr (SUPER) pay[q] = v; else pay[q] = v.x;
        entry_bins(v.y, p.sc_shift, lo[q], len[q]);
    }
    chain_wait();
    // run starts of this chunk: all loads in flight before the first dependent store
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const unsigned x = lane + 32u * i;
        if (x < p.cols2) gb[i] += __ldg(tb + x);
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const unsigned x = lane + 32u * i;
        if (x < p.cols2) S.gbase[x] = gb[i];
    }
    const unsigned pairs = distribute_rank<Payload, STAGE, G>(lo, len, p.cols2, S);
    if (pairs == 0) return;
    distribute_place<Payload, STAGE, G>(lo, len, pay, pairs, S, reinterpret_cast<Payload *>(out_list));
}

static unsigned pick_items(unsigned long long produced, unsigned long long items, unsigned stage,
                           unsigned max_items) {
    // items per chunk such that the expected output of a chunk fills ~3/4 of the staging area
    if (produced == 0 || items == 0) return max_items;
    unsigned long long k = (unsigned long long)stage * 3ull * items / (4ull * produced);
    k &= ~31ull;
    if (k < 32) k = 32;
    if (k > max_items) k = max_items;
    return (unsigned)k;
}

// num_intersects: level-2 outputs (intersections; super-column entries when sc_shift > 0)
static RcParams rc_params(const BinParams &bp, unsigned long long row_entries,
                          unsigned long long num_intersects, unsigned sc_shift) {
    RcParams p;
    p.n = bp.num_points;
    p.rows = bp.band_y1 - bp.band_y0;
    p.cols = bp.tiles_x;
    p.sc_shift = sc_shift;
    p.cols2 = ((bp.tiles_x - 1u) >> sc_shift) + 1u;
    p.band_y0 = bp.band_y0;
    p.k1 = pick_items(row_entries, bp.num_points, RC_STAGE1, 32 * RC_G1);
    p.k2 = sc_shift ? pick_items(num_intersects, row_entries, RC_STAGE2S, 32 * RC_G2S)
                    : pick_items(num_intersects, row_entries, RC_STAGE2, 32 * RC_G2);
    p.n1 = ceil_div(p.n, p.k1);
    p.n1s = (p.n1 + 3u) & ~3u;
    p.n2max = (unsigned)(row_entries / p.k2) + p.rows;
    return p;
}

}  // namespace

bool rowcol_supported(const BinParams &bp) {
    return bp.band_y1 > bp.band_y0 && bp.band_y1 - bp.band_y0 <= 256u && bp.tiles_x <= 512u;
}

bool supercols_supported(const BinParams &bp, unsigned sc_shift) {
    return rowcol_supported(bp) && sc_shift >= 1 && sc_shift <= 6 && (((bp.tiles_x - 1u) >> sc_shift) + 1u) <= 128u;
}

size_t rowcol_scratch_bytes(const BinParams &bp, unsigned long long row_entries,
                            unsigned long long num_intersects, unsigned sc_shift) {
    const RcParams p = rc_params(bp, row_entries, num_intersects, sc_shift);
    return rc_layout(p, row_entries).end * sizeof(unsigned);
}

int launch_rowcol_lists(cudaStream_t stream, const BinParams &bp, const unsigned *order,
                        const uint2 *box_by_id, unsigned long long row_entries,
                        unsigned long long num_intersects, unsigned *scratch, void *ids_sorted,
                        int2 *tile_bins, int phase, unsigned sc_shift) {
    if (bp.num_points == 0 || num_intersects == 0) return 0;
    if (!rowcol_supported(bp)) {
        set_error("rowcol_lists: band of %u rows x %u tiles not supported", bp.band_y1 - bp.band_y0,
                  bp.tiles_x);
        return 1;
    }
    if (row_entries >= (1ull << 31) || num_intersects >= (1ull << 31)) {
        set_error("rowcol_lists: %llu row entries / %llu intersections exceed 2^31", row_entries,
                  num_intersects);
        return 1;
    }
    if (sc_shift && !supercols_supported(bp, sc_shift)) {
        set_error("rowcol_lists: super-columns of 2^%u tiles not supported for %u tiles per row", sc_shift, bp.tiles_x);
        return 1;
    }
    RcParams p = rc_params(bp, row_entries, num_intersects, sc_shift);
    const RcLayout L = rc_layout(p, row_entries);
    unsigned *l1m = scratch + L.l1m, *row_tot = scratch + L.row_tot;
    unsigned *row_start = scratch + L.row_start, *chunk_base = scratch + L.chunk_base;
    uint4 *chunk_info = reinterpret_cast<uint4 *>(scratch + L.chunk_row);
    unsigned *tile_cnt = scratch + L.tile_cnt;
    unsigned *m = scratch + L.m;
    uint2 *entries = reinterpret_cast<uint2 *>(scratch + L.entries);
    const size_t smem1 = DistSmem<uint2, RC_STAGE1, RC_G1>::bytes(p.rows);
    const size_t smem2 = sc_shift ? DistSmem<uint2, RC_STAGE2S, RC_G2S>::bytes(p.cols2)
                                  : DistSmem<unsigned, RC_STAGE2, RC_G2>::bytes(p.cols2);

    if (phase == 0) {  // level 1: row lists
        launch_chained(rc_row_count_kernel, ceil_div(p.n1, (unsigned)RC_RC_WARPS), 32 * RC_RC_WARPS, 0,
                       stream, p, order, box_by_id, l1m);
        JSB_CUDA_OK(cudaGetLastError());
        launch_chained(rc_row_scan_kernel, p.rows, 1024, 0, stream, p, l1m, row_tot);
        JSB_CUDA_OK(cudaGetLastError());
        launch_chained(rc_chunk_rows_kernel, ceil_div(p.n2max, 256u), 256, 0, stream, p, row_tot,
                       row_start, chunk_base, chunk_info);
        JSB_CUDA_OK(cudaGetLastError());
        launch_chained(rc_row_scatter_kernel, p.n1, 32, smem1, stream, p, order, box_by_id, l1m,
                       row_start, entries);
        JSB_CUDA_OK(cudaGetLastError());
        note_launch(4);
    } else if (phase == 1) {  // level 2 run starts and tile bins
        launch_chained(rc_col_count_kernel, ceil_div(p.n2max, (unsigned)RC_CC_WARPS), 32 * RC_CC_WARPS, 0,
                       stream, p, chunk_info, entries, m);
        JSB_CUDA_OK(cudaGetLastError());
        launch_chained(rc_col_scan_kernel, dim3(ceil_div(p.cols2, 32u), p.rows), 32 * RC_CS_WARPS, 0,
                       stream, p, chunk_base, m, tile_cnt);
        JSB_CUDA_OK(cudaGetLastError());
        launch_chained(rc_tile_scan_bins_kernel, 1, 1024, 0, stream, p, tile_cnt, tile_bins);
        JSB_CUDA_OK(cudaGetLastError());
        note_launch(3);
    } else {  // level 2: the lists
        if (sc_shift)
            launch_chained(rc_col_scatter_kernel<true>, p.n2max, 32, smem2, stream, p, chunk_info,
                           entries, m, tile_cnt, ids_sorted);
        else
            launch_chained(rc_col_scatter_kernel<false>, p.n2max, 32, smem2, stream, p, chunk_info,
                           entries, m, tile_cnt, ids_sorted);
        JSB_CUDA_OK(cudaGetLastError());
        note_launch();
    }
    return 0;
}

}  // namespace jsb
