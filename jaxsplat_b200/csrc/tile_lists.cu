// Per-tile depth-ordered Gaussian lists for sm_100a: a single-pass counting distribution.
//
// Behavioural spec: reference lib/kernels/forward.cu:135-179 (map_gaussian_to_intersects),
// lib/ops.cu:722-757 (cub::DeviceRadixSort::SortPairs on tile<<32|depth) and forward.cu:184-208
// (get_tile_bin_edges): the output is, per tile, the ids of the Gaussians whose tile box covers
// it, in ascending (depth bits, index) order, tiles concatenated in tile-id order, plus the
// (first, one-past-last) range of every tile.
//
// Design (not a port).  The Gaussians are already in (depth, index) order (binning.cu step 1), so
// a tile's list is simply the subsequence of that order whose boxes cover the tile: the I-scale
// data never has to be *sorted*, only distributed -- one counting pass over all tiles at once:
//   1. box_gather:  tile boxes in depth order (N-scale);
//   2. emit:        (tile id, gaussian id) pairs in depth order, row-major inside each box --
//                   exactly the reference's map_gaussian_to_intersects order for depth-sorted
//                   input -- with balanced, 16-byte vector stores;
//   3. offsets:     the pairs are cut into G equal chunks; one CTA per chunk counts its pairs per
//                   tile in shared memory (G x T matrix); a two-level column scan (32 chunk
//                   groups) and a scan over tiles turn the counts into the absolute start of
//                   every (chunk, tile) run and yield tile_bins;
//   4. scatter:     one warp per chunk walks its pairs 32 at a time IN ORDER, placing ids at
//                   shared-memory cursors (<= 8192 tiles = 32 KiB per warp; larger tile grids are
//                   cut into row strips).  A chunk is walked by a single warp, so every run is
//                   stable by construction and the concatenation over chunks is the stable order.
// I-scale HBM traffic: 6 B write + 2 B + 6 B read + 4 B write per intersection plus the count
// matrix, one pass; the previous design (two LSD radix passes over 6-byte pairs with their
// upsweeps) moved ~34 B.  The scatter is bound by the rate of 4-byte partial-sector writes
// (~85 G/s measured on B200), not by bandwidth: see DESIGN.md.
#include "common.cuh"
#include "kernels.cuh"

namespace jsb {
namespace {

constexpr int TL_TCAP = 8192;   // tiles per strip (shared-memory cursors of one warp)
constexpr int TL_TAGS = 4096;   // hashed duplicate-detection tags of the scatter rounds
constexpr int TL_GROUPS = 32;   // chunk groups of the two-level column scan

struct WalkParams {
    unsigned n;            // Gaussians
    unsigned chunks;       // G
    unsigned tiles_x;
    unsigned band_y0;      // first tile row of the band
    unsigned band_rows;    // tile rows in the band
    unsigned strip_rows;   // tile rows per strip (strip tiles <= TL_TCAP)
    unsigned band_tiles;   // band_rows * tiles_x
    unsigned stride;       // band_tiles rounded up to 4 (row stride of the count matrices)
};

// ---- step 1: boxes in depth order + per-256 totals ---------------------------------------------
__global__ void __launch_bounds__(256)
box_gather_kernel(unsigned n, const unsigned *__restrict__ order, const uint2 *__restrict__ box_by_id,
                  unsigned row0, unsigned row1, uint4 *__restrict__ boxes,
                  unsigned *__restrict__ block_totals) {
    __shared__ unsigned s_w[8];
    const long long pos = (long long)blockIdx.x * 256 + threadIdx.x;
    unsigned c = 0;
    if (pos < (long long)n) {
        const unsigned g = order[pos];
        const uint2 b = __ldg(box_by_id + g);
        const unsigned w = b.y & 0xffffu, y0 = b.x >> 16;
        const unsigned ya = max(y0, row0), yb = min(y0 + (b.y >> 16), row1);
        const unsigned h = yb > ya ? yb - ya : 0u;
        c = w * h;
        boxes[pos] = make_uint4(g, (b.x & 0xffffu) | (ya << 16), h ? (w | (h << 16)) : 0u, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_w[w];
        block_totals[blockIdx.x] = t;
    }
}

// ---- step 2: (tile, id) pairs in depth order ---------------------------------------------------
// One CTA = 256 depth-ordered boxes.  Every thread owns 8-element chunks of the OUTPUT aligned in
// global index space, so full chunks leave as 16-byte vector stores; inside a chunk the
// (box, row, col) cursor advances incrementally (one binary search + one division per chunk).
// block_prefix = exclusive scan of box_gather's per-256 totals.
template <typename TileKey>
__global__ void __launch_bounds__(256)
emit_kernel(unsigned n, unsigned tiles_x, const uint4 *__restrict__ boxes,
            const unsigned *__restrict__ block_prefix, TileKey *__restrict__ keys,
            unsigned *__restrict__ vals) {
    constexpr int BT = 256;
    __shared__ unsigned s_off[BT + 1];
    __shared__ unsigned s_tile0[BT];   // first tile id of the box
    __shared__ unsigned s_w[BT];       // box width in tiles
    __shared__ unsigned s_g[BT];
    __shared__ unsigned s_ws[BT / 32 + 1];
    const int tid = threadIdx.x;
    const unsigned bid = blockIdx.x;
    const long long pos = (long long)bid * BT + tid;
    unsigned cnt = 0;
    if (pos < (long long)n) {
        const uint4 b = __ldg(boxes + pos);
        cnt = b.w;
        s_g[tid] = b.x;
        s_tile0[tid] = (b.y >> 16) * tiles_x + (b.y & 0xffffu);
        s_w[tid] = b.z & 0xffffu;
    }
    unsigned total;
    const unsigned excl = cta_exclusive_scan<BT>(cnt, &total, s_ws);
    s_off[tid] = excl;
    if (tid == 0) s_off[BT] = total;
    __syncthreads();
    const unsigned long long g_begin = block_prefix[bid], g_end = g_begin + total;
    const unsigned long long first_chunk = g_begin >> 3;
    const unsigned num_chunks = (unsigned)(((g_end + 7) >> 3) - first_chunk);
    for (unsigned c = tid; c < num_chunks; c += BT) {
        const unsigned long long cg = (first_chunk + c) << 3;
        const unsigned kb = cg > g_begin ? (unsigned)(cg - g_begin) : 0u;
        const unsigned ke = (cg + 8 < g_end) ? (unsigned)(cg + 8 - g_begin) : total;
        if (kb >= ke) continue;
        int lo = 0, hi = BT;  // invariant: s_off[lo] <= kb < s_off[hi]
#pragma unroll
        for (int st = 0; st < 8; ++st) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= kb) lo = mid; else hi = mid;
        }
        unsigned next_off = s_off[lo + 1];
        unsigned w = s_w[lo], g = s_g[lo];
        const unsigned j = kb - s_off[lo];
        const unsigned row = j / w;
        unsigned col = j - row * w;
        unsigned tile = s_tile0[lo] + row * tiles_x + col;
        TileKey kk[8];
        unsigned vv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const unsigned k = kb + e;
            if (k < ke) {
                if (k >= next_off) {
                    do { ++lo; next_off = s_off[lo + 1]; } while (k >= next_off);
                    w = s_w[lo]; g = s_g[lo]; tile = s_tile0[lo]; col = 0;
                }
                kk[e] = (TileKey)tile;
                vv[e] = g;
                ++col; ++tile;
                if (col == w) { col = 0; tile += tiles_x - w; }
            }
        }
        const bool full = (cg >= g_begin) && (ke - kb == 8);
        if (full) {
            if (sizeof(TileKey) == 2) {
                uint4 pk;
                pk.x = (unsigned)kk[0] | ((unsigned)kk[1] << 16);
                pk.y = (unsigned)kk[2] | ((unsigned)kk[3] << 16);
                pk.z = (unsigned)kk[4] | ((unsigned)kk[5] << 16);
                pk.w = (unsigned)kk[6] | ((unsigned)kk[7] << 16);
                *reinterpret_cast<uint4 *>(keys + cg) = pk;
            } else {
                *reinterpret_cast<uint4 *>(keys + cg) = make_uint4(kk[0], kk[1], kk[2], kk[3]);
                *reinterpret_cast<uint4 *>(keys + cg + 4) = make_uint4(kk[4], kk[5], kk[6], kk[7]);
            }
            *reinterpret_cast<uint4 *>(vals + cg) = make_uint4(vv[0], vv[1], vv[2], vv[3]);
            *reinterpret_cast<uint4 *>(vals + cg + 4) = make_uint4(vv[4], vv[5], vv[6], vv[7]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (kb + e < ke) {
                    keys[g_begin + kb + e] = kk[e];
                    vals[g_begin + kb + e] = vv[e];
                }
        }
    }
}

// chunk c owns pairs [c * per, min(I, (c + 1) * per)), per a multiple of 32
struct ChunkCut {
    unsigned per;
    unsigned total;  // I
};

// ---- step 3a: per-(chunk, tile) counts ----------------------------------------------------------
template <typename TileKey>
__global__ void __launch_bounds__(256)
chunk_count_kernel(const WalkParams p, const ChunkCut cut, const TileKey *__restrict__ keys,
                   unsigned *__restrict__ hist /*[G][stride]*/) {
    __shared__ __align__(16) unsigned s_cnt[TL_TCAP];
    const unsigned strip = blockIdx.x / p.chunks, chunk = blockIdx.x - strip * p.chunks;
    const unsigned sy0 = p.band_y0 + strip * p.strip_rows;
    const unsigned sy1 = min(sy0 + p.strip_rows, p.band_y0 + p.band_rows);
    const unsigned ntl = (sy1 - sy0) * p.tiles_x;
    const unsigned t0 = sy0 * p.tiles_x;  // first global tile id of the strip
    const size_t row_off = (size_t)(sy0 - p.band_y0) * p.tiles_x;
    unsigned *hrow = hist + (size_t)chunk * p.stride + row_off;
    for (unsigned i = threadIdx.x; i < ntl; i += 256) s_cnt[i] = 0u;
    __syncthreads();
    const unsigned long long begin = (unsigned long long)chunk * cut.per;
    const unsigned end = (unsigned)min((unsigned long long)cut.total, begin + cut.per);
    for (unsigned long long i = begin + threadIdx.x; i < end; i += 256) {
        const unsigned t = (unsigned)keys[i] - t0;
        if (t < ntl) atomicAdd(&s_cnt[t], 1u);
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < ntl; i += 256) hrow[i] = s_cnt[i];
}

// ---- step 4: scatter -----------------------------------------------------------------------------
// One warp per (chunk, strip) takes the chunk's pairs 32 at a time, in order.  Shared memory
// holds the next free slot of every tile's run.  A round whose lanes hold a tile twice (two boxes
// covering it; detected conservatively with hashed lane tags -- MATCH.ANY costs ~1000 cycles on
// 32 distinct values) is replayed one box at a time.  The cursor read -> write is the only
// dependency between rounds; the loads run D rounds ahead.
template <typename TileKey>
__global__ void __launch_bounds__(32)
chunk_scatter_kernel(const WalkParams p, const ChunkCut cut, const TileKey *__restrict__ keys,
                     const unsigned *__restrict__ vals, const unsigned *__restrict__ hist,
                     unsigned *__restrict__ out_ids) {
    __shared__ __align__(16) unsigned s_cur[TL_TCAP];
    __shared__ unsigned char s_tag[TL_TAGS];
    const unsigned lane = threadIdx.x;
    const unsigned strip = blockIdx.x / p.chunks, chunk = blockIdx.x - strip * p.chunks;
    const unsigned long long begin = (unsigned long long)chunk * cut.per;
    if (begin >= cut.total) return;
    const unsigned end = (unsigned)min((unsigned long long)cut.total, begin + cut.per);
    const unsigned sy0 = p.band_y0 + strip * p.strip_rows;
    const unsigned sy1 = min(sy0 + p.strip_rows, p.band_y0 + p.band_rows);
    const unsigned ntl = (sy1 - sy0) * p.tiles_x;
    const unsigned t0 = sy0 * p.tiles_x;
    const size_t row_off = (size_t)(sy0 - p.band_y0) * p.tiles_x;
    const unsigned *hrow = hist + (size_t)chunk * p.stride + row_off;
    // matrix rows are 16-byte aligned (stride multiple of 4); a strip's slice is when its first
    // tile is: then it moves as 16-byte vectors, 8 in flight per lane
    const unsigned nvec = ((row_off & 3u) == 0) ? (ntl >> 2) : 0u;
#pragma unroll 8
    for (unsigned i = lane; i < nvec; i += 32)
        reinterpret_cast<uint4 *>(s_cur)[i] = __ldcs(reinterpret_cast<const uint4 *>(hrow) + i);
#pragma unroll 4
    for (unsigned i = nvec * 4u + lane; i < ntl; i += 32) s_cur[i] = hrow[i];
    __syncwarp();

    constexpr int D = 8;  // rounds of loads in flight
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned last = end - 1u;
    // 32-bit shared-window addresses, computed once (indexing the arrays makes nvcc rebuild the
    // window base in every round)
    const unsigned cur_base = (unsigned)__cvta_generic_to_shared(s_cur);
    const unsigned tag_base = (unsigned)__cvta_generic_to_shared(s_tag);
    unsigned kq[D], vq[D];
    unsigned base = (unsigned)begin;
    // (loads are unconditional on a clamped index and a slot is refilled only after its old
    // value is dead: anything else makes ptxas copy the fresh value one round later, which
    // waits out the whole memory latency every round)
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const unsigned i = min(base + d * 32 + lane, last);
        kq[d] = (unsigned)keys[i];
        vq[d] = vals[i];
    }
    for (; base < end; base += 32 * D) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const unsigned first = base + d * 32;
            if (first >= end) break;  // warp-uniform
            const unsigned t = kq[d] - t0;
            const unsigned id = vq[d];
            const bool mine = (first + lane < end) && (t < ntl);
            const unsigned ca = cur_base + 4u * t, ta = tag_base + (t & (TL_TAGS - 1));
            // a tile held by two lanes (two boxes covering it)?  hashed tags: never misses one
            unsigned seen = lane;
            if (mine) asm volatile("st.shared.u8 [%0], %1;" ::"r"(ta), "r"(lane) : "memory");
            __syncwarp();  // (also keeps ptxas from forwarding the stored lane into the load)
            if (mine) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(seen) : "r"(ta) : "memory");
            if (!__any_sync(0xffffffffu, seen != lane)) {
                if (mine) {
                    unsigned pos;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pos) : "r"(ca) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(ca), "r"(pos + 1u) : "memory");
                    out_ids[pos] = id;
                }
            } else {
                // one box (run of equal ids) at a time: no tile repeats inside a box
                const unsigned prev = __shfl_up_sync(0xffffffffu, id, 1);
                const unsigned starts = __ballot_sync(0xffffffffu, lane == 0 || id != prev);
                const int seg = __popc(starts & (lt_mask | (1u << lane))) - 1;
                const int nseg = __popc(starts);
                for (int sg = 0; sg < nseg; ++sg) {
                    if (mine && seg == sg) {
                        unsigned pos;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pos) : "r"(ca) : "memory");
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ca), "r"(pos + 1u) : "memory");
                        out_ids[pos] = id;
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            {   // refill this slot D rounds ahead
                const unsigned i = min(first + D * 32 + lane, last);
                kq[d] = (unsigned)keys[i];
                vq[d] = vals[i];
            }
        }
    }
}

// ---- step 3: offsets ---------------------------------------------------------------------------
// per (tile, group): sum of the group's chunk counts
__global__ void __launch_bounds__(256)
group_sum_kernel(const WalkParams p, unsigned group_size, const unsigned *__restrict__ hist,
                 unsigned *__restrict__ group_hist) {
    const unsigned t = blockIdx.x * 256u + threadIdx.x;
    if (t >= p.band_tiles) return;
    const unsigned g = blockIdx.y;
    const unsigned c0 = g * group_size, c1 = min(p.chunks, c0 + group_size);
    unsigned sum = 0;
#pragma unroll 8
    for (unsigned c = c0; c < c1; ++c) sum += __ldg(hist + (size_t)c * p.stride + t);
    group_hist[(size_t)g * p.stride + t] = sum;
}

// per tile: exclusive scan over the chunk groups in place, total out
__global__ void __launch_bounds__(256)
group_scan_kernel(unsigned band_tiles, unsigned stride, unsigned *__restrict__ group_hist,
                  unsigned *__restrict__ tile_totals) {
    const unsigned t = blockIdx.x * 256u + threadIdx.x;
    if (t >= band_tiles) return;
    unsigned v[TL_GROUPS];
#pragma unroll
    for (int g = 0; g < TL_GROUPS; ++g) v[g] = group_hist[(size_t)g * stride + t];
    unsigned run = 0;
#pragma unroll
    for (int g = 0; g < TL_GROUPS; ++g) {
        group_hist[(size_t)g * stride + t] = run;
        run += v[g];
    }
    tile_totals[t] = run;
}

// per (tile, group): counts -> absolute run starts, in place; group 0 also writes tile_bins.
// tile_base = exclusive scan of tile_totals (in place, scan_totals_kernel).
__global__ void __launch_bounds__(256)
hist_apply_kernel(const WalkParams p, unsigned group_size, const unsigned *__restrict__ group_hist,
                  const unsigned *__restrict__ tile_base, unsigned total, unsigned list_offset,
                  unsigned *__restrict__ hist, int2 *__restrict__ tile_bins) {
    const unsigned t = blockIdx.x * 256u + threadIdx.x;
    if (t >= p.band_tiles) return;
    const unsigned g = blockIdx.y;
    const unsigned base = list_offset + tile_base[t];
    if (g == 0) {
        const unsigned end = list_offset + ((t + 1 < p.band_tiles) ? tile_base[t + 1] : total);
        // tiles without intersections stay (0,0) like the reference's memset + edge detection
        tile_bins[(size_t)p.band_y0 * p.tiles_x + t] =
            end > base ? make_int2((int)base, (int)end) : make_int2(0, 0);
    }
    unsigned run = base + group_hist[(size_t)g * p.stride + t];
    const unsigned c0 = g * group_size, c1 = min(p.chunks, c0 + group_size);
    constexpr int U = 8;
    for (unsigned c = c0; c < c1; c += U) {
        unsigned v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (c + u < c1) ? hist[(size_t)(c + u) * p.stride + t] : 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (c + u < c1) hist[(size_t)(c + u) * p.stride + t] = run;
            run += v[u];
        }
    }
}

}  // namespace

unsigned tile_lists_chunks(unsigned long long num_intersects) {
    // enough single-warp chains to fill the machine in one wave (6 x 32 KiB of cursors per SM),
    // never less than 2048 intersections per chunk
    const unsigned long long want = (num_intersects + 2047) / 2048;
    const unsigned cap = 148u * 6u;
    return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

static ChunkCut chunk_cut(unsigned long long num_intersects, unsigned chunks) {
    ChunkCut c;
    const unsigned long long per = (num_intersects + chunks - 1) / chunks;
    c.per = (unsigned)((per + 31) & ~31ull);
    c.total = (unsigned)num_intersects;
    return c;
}

static WalkParams walk_params(const BinParams &p, unsigned chunks) {
    WalkParams w;
    w.n = p.num_points;
    w.chunks = chunks;
    w.tiles_x = p.tiles_x;
    w.band_y0 = p.band_y0;
    w.band_rows = p.band_y1 - p.band_y0;
    w.strip_rows = p.tiles_x ? (unsigned)TL_TCAP / p.tiles_x : 1u;
    if (w.strip_rows == 0) w.strip_rows = 1;
    if (w.strip_rows > w.band_rows) w.strip_rows = w.band_rows;
    w.band_tiles = w.band_rows * p.tiles_x;
    w.stride = (w.band_tiles + 3u) & ~3u;
    return w;
}

// scratch layout (words): hist [G][stride] | group_hist [GROUPS][stride] | tile totals [stride]
size_t tile_lists_scratch_words(const BinParams &p, unsigned chunks) {
    const WalkParams w = walk_params(p, chunks);
    return ((size_t)chunks + TL_GROUPS + 1) * w.stride + 16;
}

unsigned tile_strip_rows(unsigned tiles_x) { return tiles_x ? (unsigned)TL_TCAP / tiles_x : 0u; }

int launch_box_gather(cudaStream_t stream, unsigned n, const unsigned *order, const uint2 *box_by_id,
                      unsigned row0, unsigned row1, uint4 *boxes, unsigned *block_totals) {
    if (n == 0) return 0;
    box_gather_kernel<<<ceil_div(n, 256u), 256, 0, stream>>>(n, order, box_by_id, row0, row1, boxes,
                                                            block_totals);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

template <typename TileKey>
int launch_emit(cudaStream_t stream, const BinParams &p, const uint4 *boxes,
                unsigned *block_totals, TileKey *keys, unsigned *vals) {
    if (p.num_points == 0) return 0;
    const unsigned nb = ceil_div(p.num_points, 256u);
    if (launch_scan_totals(stream, block_totals, nb)) return 1;
    emit_kernel<TileKey><<<nb, 256, 0, stream>>>(p.num_points, p.tiles_x, boxes, block_totals, keys, vals);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

template <typename TileKey>
int launch_tile_offsets(cudaStream_t stream, const BinParams &p, unsigned chunks,
                        unsigned long long num_intersects, unsigned long long list_offset,
                        const TileKey *keys, unsigned *scratch, int2 *tile_bins) {
    if (p.num_points == 0 || p.band_y1 <= p.band_y0) return 0;
    if (p.tiles_x > (unsigned)TL_TCAP) {
        set_error("tile grid %u tiles wide exceeds the supported %d", p.tiles_x, TL_TCAP);
        return 1;
    }
    const WalkParams w = walk_params(p, chunks);
    const ChunkCut cut = chunk_cut(num_intersects, chunks);
    unsigned *hist = scratch;
    unsigned *group_hist = hist + (size_t)chunks * w.stride;
    unsigned *tile_totals = group_hist + (size_t)TL_GROUPS * w.stride;
    const unsigned strips = ceil_div(w.band_rows, w.strip_rows);
    const unsigned group_size = ceil_div(chunks, (unsigned)TL_GROUPS);
    const unsigned tile_blocks = ceil_div(w.band_tiles, 256u);

    chunk_count_kernel<TileKey><<<chunks * strips, 256, 0, stream>>>(w, cut, keys, hist);
    JSB_CUDA_OK(cudaGetLastError());
    group_sum_kernel<<<dim3(tile_blocks, TL_GROUPS), 256, 0, stream>>>(w, group_size, hist, group_hist);
    JSB_CUDA_OK(cudaGetLastError());
    group_scan_kernel<<<tile_blocks, 256, 0, stream>>>(w.band_tiles, w.stride, group_hist, tile_totals);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch(3);
    if (launch_scan_totals(stream, tile_totals, w.band_tiles)) return 1;
    hist_apply_kernel<<<dim3(tile_blocks, TL_GROUPS), 256, 0, stream>>>(
        w, group_size, group_hist, tile_totals, (unsigned)num_intersects, (unsigned)list_offset, hist,
        tile_bins);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

template <typename TileKey>
int launch_tile_scatter(cudaStream_t stream, const BinParams &p, unsigned chunks,
                        unsigned long long num_intersects, const TileKey *keys,
                        const unsigned *vals, const unsigned *scratch, unsigned *ids_sorted) {
    if (p.num_points == 0 || p.band_y1 <= p.band_y0) return 0;
    const WalkParams w = walk_params(p, chunks);
    const ChunkCut cut = chunk_cut(num_intersects, chunks);
    const unsigned strips = ceil_div(w.band_rows, w.strip_rows);
    chunk_scatter_kernel<TileKey><<<chunks * strips, 32, 0, stream>>>(w, cut, keys, vals, scratch,
                                                                      ids_sorted);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

#define JSB_INSTANTIATE(K)                                                                        \
    template int launch_emit<K>(cudaStream_t, const BinParams &, const uint4 *, unsigned *, K *,   \
                                unsigned *);                                                       \
    template int launch_tile_offsets<K>(cudaStream_t, const BinParams &, unsigned,                 \
                                        unsigned long long, unsigned long long, const K *,         \
                                        unsigned *, int2 *);                                       \
    template int launch_tile_scatter<K>(cudaStream_t, const BinParams &, unsigned,                 \
                                        unsigned long long, const K *, const unsigned *,           \
                                        const unsigned *, unsigned *);
JSB_INSTANTIATE(uint16_t)
JSB_INSTANTIATE(uint32_t)

}  // namespace jsb
