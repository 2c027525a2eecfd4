// Intersect emission and tile binning for sm_100a.
//
// Behavioural spec: reference lib/kernels/forward.cu:135-179 (map_gaussian_to_intersects),
// :184-208 (get_tile_bin_edges) and lib/ops.cu:645-793 (sort_and_bin): the sorted list is the
// stable ascending order of keys (tile_id << 32 | depth_bits) over all (Gaussian, covered tile)
// pairs, emitted Gaussian-major, row-major inside each tile box.
//
// Design (not a port).  The reference sorts I 64-bit keys with 6 radix passes.  The same total
// order is obtained here with far less traffic by splitting the key:
//   1. sort the N Gaussians once by depth bits (stable, ties keep index order)        [N-scale]
//   2. emit (tile id, gaussian id) pairs in that depth order, offsets from a fused scan
//   3. one stable LSD sort of the I pairs on the ceil(log2 T) live tile bits only      [I-scale]
// Stability of step 3 preserves the depth-then-index order inside every tile, i.e. exactly the
// reference's (tile, depth, index) order, with 2-byte keys and 2 passes instead of 8-byte keys
// and 6 passes.
#include "common.cuh"
#include "kernels.cuh"
#include "raster_cull.cuh"

namespace jsb {
namespace {

constexpr int BT = 256;

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// WRITE=false: fingerprint + intersection count only (rasterize_bwd validating the forward's cache)
template <bool WRITE>
__global__ void __launch_bounds__(BT)
bin_prep_kernel(const BinParams p, const float2 *__restrict__ xys,
                const float *__restrict__ depths, const int *__restrict__ radii,
                const float *__restrict__ conics, const float *__restrict__ opacities,
                const float *__restrict__ colors, SplatRecord *__restrict__ records,
                unsigned *__restrict__ depth_keys, unsigned *__restrict__ ids,
                uint2 *__restrict__ box_by_id, unsigned long long *__restrict__ fingerprint) {
    __shared__ __align__(16) float s_buf[BT * 6];
    __shared__ unsigned s_strip[kMaxStrips];
    const int tid = threadIdx.x;
    const unsigned nstrips = ceil_div(p.band_y1 - p.band_y0, p.strip_rows);
    if (WRITE && nstrips > 1)
        for (unsigned i = tid; i < nstrips; i += BT) s_strip[i] = 0u;
    const long long base = (long long)blockIdx.x * BT;
    const long long g = base + tid;
    const long long n = p.num_points;
    float *s_conic = s_buf, *s_color = s_buf + BT * 3;
    const bool pack = conics != nullptr;  // sort_and_bin-only callers pass no appearance data
    if (pack) {
        slab_load<3, BT>(s_conic, conics, base * 3, n * 3);
        slab_load<3, BT>(s_color, colors, base * 3, n * 3);
    }
    __syncthreads();
    unsigned cnt = 0, cnt_rows = 0, cnt_sup = 0, kinv = 0, kmax = 0;
    if (g < n) {
        const float2 xy = __ldg(xys + g);
        const float depth = __ldg(depths + g);
        const int radius = __ldg(radii + g);
        float opac = 0.f, cx = 0.f, cy = 0.f, cz = 0.f, r = 0.f, gr = 0.f, b = 0.f;
        if (pack) {
            opac = __ldg(opacities + g);
            cx = s_conic[tid * 3]; cy = s_conic[tid * 3 + 1]; cz = s_conic[tid * 3 + 2];
            r = s_color[tid * 3]; gr = s_color[tid * 3 + 1]; b = s_color[tid * 3 + 2];
            if (WRITE) {
                SplatRecord rec;
                rec.q0 = make_float4(xy.x, xy.y, opac, cx);
                rec.q1 = make_float4(cy, cz, r, gr);
                rec.q2 = make_float4(b, cull_record_thresh(opac, cx, cy, cz, (float)radius, p.block_width_f), 0.f, 0.f);
                records[g] = rec;
            }
        }
        unsigned key = 0xFFFFFFFFu;
        uint2 box = make_uint2(0u, 0u);  // {x0 | y0 << 16, width | height << 16}, band-clamped
        if (radius > 0) {
            unsigned x0, x1, y0, y1;
            tile_bbox(xy.x, xy.y, (float)radius, p.block_width_f, p.tiles_x, p.tiles_y, x0, x1,
                      y0, y1);
            y0 = max(y0, p.band_y0);
            y1 = min(y1, p.band_y1);
            if (y1 > y0 && x1 > x0) {
                cnt = (x1 - x0) * (y1 - y0);
                cnt_rows = y1 - y0;
                cnt_sup = cnt_rows * (((x1 - 1u) >> p.sc_shift) - (x0 >> p.sc_shift) + 1u);
                box = make_uint2(x0 | (y0 << 16), (x1 - x0) | ((y1 - y0) << 16));
                if (WRITE && nstrips > 1) {  // tiles per strip (the strips are binned one by one)
                    for (unsigned st = (y0 - p.band_y0) / p.strip_rows; st < nstrips; ++st) {
                        const unsigned r0 = p.band_y0 + st * p.strip_rows;
                        if (r0 >= y1) break;
                        const unsigned rows = min(y1, r0 + p.strip_rows) - max(y0, r0);
                        atomicAdd(&s_strip[st], (x1 - x0) * rows);
                    }
                }
            }
            key = __float_as_uint(depth);
            kinv = ~key;
            kmax = key;
        }
        if (WRITE) {
            depth_keys[g] = key;
            ids[g] = (unsigned)g;
            box_by_id[g] = box;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        cnt_rows += __shfl_xor_sync(0xffffffffu, cnt_rows, o);
        cnt_sup += __shfl_xor_sync(0xffffffffu, cnt_sup, o);
    }
    kinv = __reduce_max_sync(0xffffffffu, kinv);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    // one pair of global atomics per CTA (they all hit the same two words: per-warp atomics
    // serialise in the L2 and were the kernel's critical path)
    __shared__ unsigned s_tot[BT / 32][5];
    if ((tid & 31) == 0) {
        s_tot[tid >> 5][0] = cnt; s_tot[tid >> 5][1] = cnt_rows;
        s_tot[tid >> 5][2] = kinv; s_tot[tid >> 5][3] = kmax;
        s_tot[tid >> 5][4] = cnt_sup;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long t0 = 0, t1 = 0, t2 = 0;
        unsigned m0 = 0, m1 = 0;
#pragma unroll
        for (int w = 0; w < BT / 32; ++w) {
            t0 += s_tot[w][0]; t1 += s_tot[w][1]; t2 += s_tot[w][4];
            m0 = max(m0, s_tot[w][2]); m1 = max(m1, s_tot[w][3]);
        }
        if (WRITE && (m0 | m1)) {  // range of the real depth keys, for the live-bit depth sort
            unsigned *mm = reinterpret_cast<unsigned *>(fingerprint + 5);
            atomicMax(mm, m0);
            atomicMax(mm + 1, m1);
        }
        if (t0) {
            atomicAdd(fingerprint + 1, t0);  // total intersections
            // sum of box heights = row entries of the two-level distribution (rowcol_lists.cu)
            if (WRITE) {
                atomicAdd(fingerprint + 4, t1);
                atomicAdd(fingerprint + 6, t2);  // (row entry, super-column) pairs
            }
        }
    }
    if (WRITE && nstrips > 1) {
        __syncthreads();
        for (unsigned i = tid; i < nstrips; i += BT)
            if (s_strip[i]) atomicAdd(fingerprint + 8 + i, (unsigned long long)s_strip[i]);
    }
}

// Order-independent 64-bit fingerprint of up to 6 word arrays (sum of per-word hashes keyed by
// array id and word index).  Pure streaming read with 16-byte loads: rasterize_bwd uses it to
// prove that its inputs are byte-identical to the forward's before reusing the cached binning.
struct FpArrays {
    const unsigned *ptr[6];
    unsigned long long words[6];
};
__device__ __forceinline__ unsigned long long fp_word(unsigned a, unsigned long long i, unsigned v) {
    return mix64((unsigned long long)v * 0x9e3779b97f4a7c15ull + (i + ((unsigned long long)(a + 1) << 40)));
}
__global__ void __launch_bounds__(256) fingerprint_kernel(const FpArrays A, unsigned long long *out) {
    unsigned long long h = 0;
    const unsigned long long tid = blockIdx.x * 256ull + threadIdx.x, nthr = gridDim.x * 256ull;
#pragma unroll 1
    for (int a = 0; a < 6; ++a) {
        const unsigned *p = A.ptr[a];
        if (!p) continue;
        const unsigned long long nw = A.words[a];
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            const unsigned long long nv = nw >> 2;
            for (unsigned long long v = tid; v < nv; v += nthr) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + v);
                const unsigned long long lo = ((unsigned long long)q.y << 32) | q.x;
                const unsigned long long hi = ((unsigned long long)q.w << 32) | q.z;
                // one avalanche per 16 bytes (a different, equally valid hash than the word path:
                // a fwd/bwd pair only has to agree when both see the same alignment)
                h += mix64((lo * 0x9e3779b97f4a7c15ull) ^ (hi * 0xc2b2ae3d27d4eb4full + v) ^
                           ((unsigned long long)(a + 1) << 56));
            }
            for (unsigned long long i = (nv << 2) + tid; i < nw; i += nthr) h += fp_word(a, i, __ldg(p + i));
        } else {
            for (unsigned long long i = tid; i < nw; i += nthr) h += fp_word(a, i, __ldg(p + i));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    __shared__ unsigned long long s_h[8];
    if ((threadIdx.x & 31) == 0) s_h[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {  // one same-address atomic per CTA
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_h[w];
        if (t) atomicAdd(out, t);
    }
}

// inclusive int32 prefix sum, single pass (reference ops.cu:607-643 uses cub::DeviceScan)
constexpr int SC_ITEMS = 4;
__global__ void __launch_bounds__(BT)
inclusive_scan_kernel(const int *__restrict__ in, int *__restrict__ out, unsigned n,
                      unsigned long long *__restrict__ chain, unsigned *__restrict__ ticket,
                      int *__restrict__ err) {
    __shared__ unsigned s_ws[BT / 32 + 1];
    __shared__ unsigned s_bid, s_prefix;
    const int tid = threadIdx.x;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const unsigned long long first = (unsigned long long)bid * (BT * SC_ITEMS) + (unsigned)tid * SC_ITEMS;
    unsigned v[SC_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        v[i] = (first + i < n) ? (unsigned)in[first + i] : 0u;
        sum += v[i];
    }
    unsigned total;
    unsigned run = cta_exclusive_scan<BT>(sum, &total, s_ws);
    if (tid < 32) {
        unsigned pre = chain_exclusive_prefix(chain, (int)bid, total, err);
        if (tid == 0) s_prefix = pre;
    }
    __syncthreads();
    run += s_prefix;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        run += v[i];
        if (first + i < n) out[first + i] = (int)run;
    }
}

// in-place exclusive scan of the CTA totals: a single CTA walks them 1024 at a time
__global__ void __launch_bounds__(1024) scan_totals_kernel(unsigned *__restrict__ totals, unsigned nb) {
    __shared__ unsigned s_ws[1024 / 32 + 1];
    unsigned carry = 0;
    for (unsigned b = 0; b < nb; b += 1024) {
        const unsigned i = b + threadIdx.x;
        const unsigned v = i < nb ? totals[i] : 0u;
        unsigned total;
        const unsigned e = cta_exclusive_scan<1024>(v, &total, s_ws);
        if (i < nb) totals[i] = carry + e;
        carry += total;
        __syncthreads();
    }
}

// cum[i] += prefix[i / per_block]   (per_block is a multiple of 4: one int4 never straddles)
__global__ void __launch_bounds__(256)
add_block_prefix_kernel(int *__restrict__ cum, const unsigned *__restrict__ prefix, unsigned n,
                        unsigned per_block) {
    const unsigned v = blockIdx.x * 256u + threadIdx.x;
    const unsigned i = v * 4u;
    if (i >= n) return;
    const int add = (int)prefix[i / per_block];
    if (i + 4 <= n && (reinterpret_cast<uintptr_t>(cum) & 15) == 0) {
        int4 q = reinterpret_cast<int4 *>(cum)[v];
        q.x += add; q.y += add; q.z += add; q.w += add;
        reinterpret_cast<int4 *>(cum)[v] = q;
    } else {
        for (unsigned k = i; k < n && k < i + 4; ++k) cum[k] += add;
    }
}

__global__ void rebuild_keys64_kernel(const int2 *__restrict__ bins, unsigned num_tiles,
                                      const unsigned *__restrict__ ids,
                                      const float *__restrict__ depths,
                                      long long *__restrict__ keys64) {
    const unsigned t = blockIdx.x;
    if (t >= num_tiles) return;
    const int2 r = bins[t];
    for (int k = r.x + (int)threadIdx.x; k < r.y; k += (int)blockDim.x) {
        const long long d = (long long)(int)__float_as_uint(depths[ids[k]]);
        keys64[k] = ((long long)t << 32) | d;
    }
}

}  // namespace

int launch_inclusive_scan_i32(cudaStream_t stream, const int *in, int *out, unsigned n,
                              unsigned long long *chain, unsigned *ticket, int *err) {
    if (n == 0) return 0;
    inclusive_scan_kernel<<<ceil_div(n, BT * SC_ITEMS), BT, 0, stream>>>(in, out, n, chain, ticket, err);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_fingerprint(cudaStream_t stream, unsigned n, const float *xys, const float *depths,
                       const int *radii, const float *conics, const float *opacities,
                       const float *colors, unsigned long long *out) {
    if (n == 0) return 0;
    FpArrays A;
    const void *ptrs[6] = {xys, depths, radii, conics, opacities, colors};
    const unsigned long long words[6] = {2ull * n, n, n, 3ull * n, n, 3ull * n};
    for (int a = 0; a < 6; ++a) {
        A.ptr[a] = reinterpret_cast<const unsigned *>(ptrs[a]);
        A.words[a] = words[a];
    }
    fingerprint_kernel<<<148 * 8, 256, 0, stream>>>(A, out);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_bin_prep(cudaStream_t stream, const BinParams &p, const float *xys,
                    const float *depths, const int *radii, const float *conics,
                    const float *opacities, const float *colors, SplatRecord *records,
                    unsigned *depth_keys, unsigned *ids, uint2 *box_by_id,
                    unsigned long long *fingerprint, bool write) {
    if (p.num_points == 0) return 0;
    if (write)
        bin_prep_kernel<true><<<ceil_div(p.num_points, BT), BT, 0, stream>>>(
            p, reinterpret_cast<const float2 *>(xys), depths, radii, conics, opacities, colors,
            records, depth_keys, ids, box_by_id, fingerprint);
    else
        bin_prep_kernel<false><<<ceil_div(p.num_points, BT), BT, 0, stream>>>(
            p, reinterpret_cast<const float2 *>(xys), depths, radii, conics, opacities, colors,
            records, depth_keys, ids, box_by_id, fingerprint);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_scan_totals(cudaStream_t stream, unsigned *totals, unsigned nb) {
    if (nb == 0) return 0;
    scan_totals_kernel<<<1, 1024, 0, stream>>>(totals, nb);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_add_block_prefix(cudaStream_t stream, int *cum, const unsigned *prefix, unsigned n,
                            unsigned per_block) {
    if (n == 0) return 0;
    add_block_prefix_kernel<<<ceil_div(ceil_div(n, 4), 256), 256, 0, stream>>>(cum, prefix, n, per_block);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}
int launch_rebuild_keys64(cudaStream_t stream, const int2 *tile_bins, unsigned num_tiles,
                          const unsigned *ids_sorted, const float *depths, long long *keys64) {
    if (num_tiles == 0) return 0;
    rebuild_keys64_kernel<<<num_tiles, 128, 0, stream>>>(tile_bins, num_tiles, ids_sorted, depths,
                                                         keys64);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace jsb
