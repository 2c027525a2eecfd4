// Launcher declarations for the sm_100a kernels.  Every launcher enqueues on `stream` only and
// returns 0 / non-zero (error text via jsb::set_error).  No launcher allocates or synchronises.
#pragma once
#include "common.cuh"

namespace jsb {

struct ProjectFwdParams {
    unsigned num_points;
    float glob_scale;
    float fx, fy, cx, cy;
    float tan_fovx, tan_fovy;  // float(0.5 * double(W) / double(fx)) evaluated on the host
    float clip_thresh;
    float block_width_f;  // (float)(int)block_width
    unsigned tiles_x, tiles_y;
};

struct ProjectBwdParams {
    unsigned num_points;
    float glob_scale;
    float fx, fy;
};

int launch_project_fwd(cudaStream_t stream, const ProjectFwdParams &p, const float *means3d,
                       const float *scales, const float *quats, const float *viewmat,
                       float *cov3ds, float *xys, float *depths, int *radii, float *conics,
                       float *compensation, int *num_tiles_hit, int *cum_tiles_hit,
                       unsigned *block_totals /* ceil(N/256) scratch words */);

int launch_project_bwd(cudaStream_t stream, const ProjectBwdParams &p, const float *means3d,
                       const float *scales, const float *quats, const float *viewmat,
                       const float *cov3ds, const int *radii, const float *conics,
                       const float *compensation, const float *v_compensation, const float *v_xy,
                       const float *v_depth, const float *v_conic, float *v_mean3d,
                       float *v_scale, float *v_quat, float *v_cov2d, float *v_cov3d);

// ---- binning (binning.cu) -------------------------------------------------------------------
// Per-Gaussian record gathered by the rasterisers: 3 x float4 = 48 B, one 16-B aligned line.
//   q0 = {x, y, opacity, conic.x}   q1 = {conic.y, conic.z, r, g}   q2 = {b, 0, 0, 0}
struct __align__(16) SplatRecord {
    float4 q0, q1, q2;
};

struct BinParams {
    unsigned num_points;
    float block_width_f;
    unsigned tiles_x, tiles_y;
    unsigned band_y0, band_y1;  // tile rows [band_y0, band_y1) this call owns (full frame: 0, tiles_y)
};

// prep: pack records, depth keys (0xFFFFFFFF for radii<=0), per-Gaussian tile counts,
// identity ids; fingerprint[0] += 64-bit hash of every input, fingerprint[1] += tile count
// (= num_intersects).  conics == nullptr skips the record packing (sort_and_bin-only callers).
int launch_bin_prep(cudaStream_t stream, const BinParams &p, const float *xys,
                    const float *depths, const int *radii, const float *conics,
                    const float *opacities, const float *colors, SplatRecord *records,
                    unsigned *depth_keys, unsigned *ids, unsigned *tile_counts,
                    unsigned long long *fingerprint, bool write = true);

// *out += order-independent 64-bit hash of every rasterizer input word (null arrays skipped)
int launch_fingerprint(cudaStream_t stream, unsigned n, const float *xys, const float *depths,
                       const int *radii, const float *conics, const float *opacities,
                       const float *colors, unsigned long long *out);

// inclusive int32 prefix sum (chain: ceil(n/1024) u64 zeroed, ticket zeroed)
int launch_inclusive_scan_i32(cudaStream_t stream, const int *in, int *out, unsigned n,
                              unsigned long long *chain, unsigned *ticket, int *err);

// in-place exclusive scan of nb CTA totals (single CTA), and cum[i] += totals[i / per_block]
int launch_scan_totals(cudaStream_t stream, unsigned *totals, unsigned nb);
int launch_add_block_prefix(cudaStream_t stream, int *cum, const unsigned *prefix, unsigned n,
                            unsigned per_block);

// emit: for Gaussians in depth order, write (tile id, gaussian id) for every covered tile,
// row-major inside the box (reference forward.cu:135-179); the per-CTA output offsets come from
// a count kernel + a scan of the CTA totals (no chained look-back, see project.cu).  TileKey is uint16_t (<= 65536 tiles) or uint32_t.
template <typename TileKey>
int launch_emit(cudaStream_t stream, const BinParams &p, const unsigned *order,
                const unsigned *tile_counts, const float *xys, const int *radii, TileKey *keys,
                unsigned *vals, unsigned *block_totals /* ceil(N/256) scratch words */);

// tile_bins[t] = (first, one-past-last) index of tile t in the sorted list; tiles without
// intersections stay (0,0) (reference forward.cu:184-208 + memset ops.cu:482-488).
template <typename TileKey>
int launch_bin_edges(cudaStream_t stream, const TileKey *sorted_keys, unsigned n, int2 *tile_bins);

// optional debug/parity helper: rebuild the reference's sorted 64-bit keys (tile<<32 | depth bits)
int launch_rebuild_keys64(cudaStream_t stream, const int2 *tile_bins, unsigned num_tiles,
                          const unsigned *ids_sorted, const float *depths, long long *keys64);

// ---- radix sort (radix_sort.cu) -------------------------------------------------------------
// Stable LSD onesweep sort of n (key,value) pairs on key bits [0, bits).  Buffers ping-pong
// a<->b; *result_in_b says where the sorted data ended up.  `scratch` needs
// radix_scratch_bytes(n, bits) bytes and is zeroed inside (one memset).
size_t radix_scratch_bytes(unsigned n, int bits);
template <typename KeyT>
int radix_sort_pairs(cudaStream_t stream, KeyT *keys_a, KeyT *keys_b, unsigned *vals_a,
                     unsigned *vals_b, unsigned n, int bits, void *scratch, int *err,
                     bool *result_in_b);

// ---- rasterisers (raster_fwd.cu / raster_bwd.cu) --------------------------------------------
struct RasterParams {
    unsigned img_w, img_h;
    unsigned tiles_x, tiles_y;
    unsigned block_width;
    unsigned band_y0, band_y1;  // tile rows rasterised by this call (screen-tile-row sharding)
};

int launch_raster_fwd(cudaStream_t stream, const RasterParams &p, const unsigned *ids_sorted,
                      const int2 *tile_bins, const SplatRecord *records, const float *background,
                      float *final_Ts, int *final_idx, float *out_img);

int launch_raster_bwd(cudaStream_t stream, const RasterParams &p, const unsigned *ids_sorted,
                      const int2 *tile_bins, const SplatRecord *records, const float *background,
                      const float *final_Ts, const int *final_idx, const float *v_out,
                      const float *v_out_alpha, float *v_xy, float *v_xy_abs, float *v_conic,
                      float *v_colors, float *v_opacity);

// image = background, final_Ts = 0, final_idx = 0  (reference num_intersects==0 path,
// ops.cu:119-127 + memsets :436-458)
int launch_fill_background(cudaStream_t stream, const RasterParams &p, const float *background,
                           float *final_Ts, int *final_idx, float *out_img);

}  // namespace jsb
