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
    unsigned sc_shift = 0;  // fused prep: log2 of the tiles per super-column counted into totals[6]
};

struct ProjectBwdParams {
    unsigned num_points;
    float glob_scale;
    float fx, fy;
    // the kernel right before this one on the stream is the library's own and everything but the
    // cotangents (parameters, radii, conics, cov3ds, viewmat, opacities) was complete before THAT
    // kernel started: those are loaded ahead of the in-kernel wait on the predecessor
    bool residuals_ready = false;
};

struct SplatRecord;
// Fused render, full-frame band: project_fwd also produces what bin_prep would (binning.cu):
// records, depth keys (0xFFFFFFFF for radii <= 0), identity ids, tile boxes, adds the
// intersection / row-entry totals to totals[1] / totals[4], the (row entry, super-column) pairs to
// totals[6], and folds the range of the real depth keys into the two u32 words of totals[5]:
// {max(~key), max(key)} (both 0 = no real key).
struct FusedPrep {
    const float *colors = nullptr, *opacities = nullptr;
    SplatRecord *records = nullptr;
    unsigned *depth_keys = nullptr, *ids = nullptr;
    uint2 *box_by_id = nullptr;
    unsigned long long *totals = nullptr;
};

int launch_project_fwd(cudaStream_t stream, const ProjectFwdParams &p, const float *means3d,
                       const float *scales, const float *quats, const float *viewmat,
                       float *cov3ds, float *xys, float *depths, int *radii, float *conics,
                       float *compensation, int *num_tiles_hit, int *cum_tiles_hit,
                       unsigned *block_totals /* ceil(N/256) scratch words */,
                       const FusedPrep *fused = nullptr);

// cov3ds == nullptr (lean residuals): the 3D covariance is recomputed from scales and quaternions.
int launch_project_bwd(cudaStream_t stream, const ProjectBwdParams &p, const float *means3d,
                       const float *scales, const float *quats, const float *viewmat,
                       const float *cov3ds, const int *radii, const float *conics,
                       const float *compensation, const float *v_compensation, const float *v_xy,
                       const float *v_depth, const float *v_conic, float *v_mean3d,
                       float *v_scale, float *v_quat, float *v_cov2d, float *v_cov3d,
                       const float *opacities = nullptr, float *moment_opacity = nullptr);
// opacities + moment_opacity != null: v_xy / v_conic / moment_opacity hold the rasteriser's raw
// moments; they are turned into gradients on the fly and moment_opacity[i] becomes v_opacity[i].

// batched cameras, one launch for all views: per-view radii (B,n), conics (B,n,3) and raw moments
// mom_xy (view b at mom_xy + b * mom_stride floats, n float2) / mom_conic (view b at mom_conic +
// b * mom_stride floats, 3n floats); mom_opacity (n) = sum over the views of M0, leaves as v_opacity
int launch_project_bwd_views(cudaStream_t stream, const ProjectBwdParams &p, unsigned num_views,
                            const float *means3d, const float *scales, const float *quats,
                            const float *viewmats, const int *radii, const float *conics,
                            const float *mom_xy, const float *mom_conic, size_t mom_stride,
                            float *v_mean3d, float *v_scale, float *v_quat, const float *opacities,
                            float *mom_opacity);

// batched cameras: v_mean3d / v_scale / v_quat += one view's staged arrays (3n / 3n / 4n floats)
int launch_accumulate_param_grads(cudaStream_t stream, size_t n, const float *s_mean,
                                  const float *s_scale, const float *s_quat, float *v_mean3d,
                                  float *v_scale, float *v_quat);

// ---- binning (binning.cu) -------------------------------------------------------------------
// Per-Gaussian record gathered by the rasterisers: 3 x float4 = 48 B, one 16-B aligned line.
//   q0 = {x, y, opacity, conic.x}   q1 = {conic.y, conic.z, r, g}   q2 = {b, cull threshold, 0, 0}
// (cull threshold: raster_cull.cuh, cull_record_thresh -- the per-Gaussian part of the conservative
// quadrant pre-filter, so the rasterisers do not redo it for every tile the Gaussian is listed in)
struct __align__(16) SplatRecord {
    float4 q0, q1, q2;
};

struct BinParams {
    unsigned num_points;
    float block_width_f;
    unsigned tiles_x, tiles_y;
    unsigned band_y0, band_y1;  // tile rows [band_y0, band_y1) this call owns (full frame: 0, tiles_y)
    unsigned strip_rows;        // the band is binned in strips of this many tile rows (tile_lists.cu)
    unsigned sc_shift = 0;      // prep: log2 of the tiles per super-column counted into fingerprint[6]
};

constexpr unsigned kMaxStrips = 256;
// tile rows per strip such that one strip holds at most 8192 tiles (0 if a single row does not fit)
unsigned tile_strip_rows(unsigned tiles_x);

// prep: pack records, depth keys (0xFFFFFFFF for radii<=0), per-Gaussian band-clamped tile boxes
// {x0 | y0 << 16, width | height << 16}, identity ids; fingerprint[0] += 64-bit hash of every input, fingerprint[1] += tile count
// (= num_intersects).  conics == nullptr skips the record packing (sort_and_bin-only callers).
int launch_bin_prep(cudaStream_t stream, const BinParams &p, const float *xys,
                    const float *depths, const int *radii, const float *conics,
                    const float *opacities, const float *colors, SplatRecord *records,
                    unsigned *depth_keys, unsigned *ids, uint2 *box_by_id,
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

// ---- per-tile lists (tile_lists.cu) ----------------------------------------------------------
// boxes[i] = {id, x0|y0<<16, w|h<<16, tiles} of the i-th Gaussian in depth order; block_totals[b] =
// tiles of depth positions [256 b, 256 b + 256)
// (boxes clamped to tile rows [row0, row1))
int launch_box_gather(cudaStream_t stream, unsigned n, const unsigned *order, const uint2 *box_by_id,
                      unsigned row0, unsigned row1, uint4 *boxes,
                      unsigned *block_totals /* ceil(N/256) words */);
// (tile id, gaussian id) pairs in depth order, row-major inside each box (the reference's
// map_gaussian_to_intersects order, forward.cu:135-179, for depth-sorted Gaussians).
// block_totals is scanned in place.  TileKey is uint16_t (<= 65536 tiles) or uint32_t.
template <typename TileKey>
int launch_emit(cudaStream_t stream, const BinParams &p, const uint4 *boxes,
                unsigned *block_totals, TileKey *keys, unsigned *vals);
// number of single-warp chunks the depth-ordered pairs are cut into
unsigned tile_lists_chunks(unsigned long long num_intersects);
size_t tile_lists_scratch_words(const BinParams &p, unsigned chunks);
// count + offsets: fills the (chunk, tile) run starts in `scratch` (absolute: list_offset is the
// position of the band's first pair in the whole list) and tile_bins of the band's tiles ((0,0) for tiles without intersections, reference forward.cu:184-208 + ops.cu:482-488)
template <typename TileKey>
int launch_tile_offsets(cudaStream_t stream, const BinParams &p, unsigned chunks,
                        unsigned long long num_intersects, unsigned long long list_offset,
                        const TileKey *keys, unsigned *scratch, int2 *tile_bins);
// scatter: ids_sorted = per-tile (depth, index)-ordered Gaussian ids, tiles in tile-id order --
// the value array of the reference's sorted (tile<<32|depth, id) pairs (ops.cu:722-757)
template <typename TileKey>
int launch_tile_scatter(cudaStream_t stream, const BinParams &p, unsigned chunks,
                        unsigned long long num_intersects, const TileKey *keys,
                        const unsigned *vals, const unsigned *scratch, unsigned *ids_sorted);

// ---- per-tile lists, two-level row/column distribution (rowcol_lists.cu) ----------------------
// Same result as box_gather + emit + tile_offsets + tile_scatter for bands of <= 256 tile rows
// and <= 512 tiles per row.  row_entries = sum of the band-clamped box heights (bin_prep counts
// it), phase 0 = row lists, 1 = run starts + tile_bins of the band, 2 = ids_sorted.
// sc_shift > 0 (fused render, super-column lists): level 2 distributes the 8-byte row entries
// {id, x0 | w << 16} into super-columns of 2^sc_shift tiles instead of ids into tiles;
// num_intersects is then the number of (row entry, super-column) pairs, ids_sorted a uint2 array
// of that many entries and tile_bins a band-local (rows x ceil(tiles_x / 2^sc_shift)) array.
bool rowcol_supported(const BinParams &p);
bool supercols_supported(const BinParams &p, unsigned sc_shift);
size_t rowcol_scratch_bytes(const BinParams &p, unsigned long long row_entries,
                            unsigned long long num_intersects, unsigned sc_shift = 0);
int launch_rowcol_lists(cudaStream_t stream, const BinParams &p, const unsigned *order,
                        const uint2 *box_by_id, unsigned long long row_entries,
                        unsigned long long num_intersects, unsigned *scratch, void *ids_sorted,
                        int2 *tile_bins, int phase, unsigned sc_shift = 0);

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

// ---- depth order in one cooperative launch (resident_sort.cu) ----------------------------------
// order[i] = id of the i-th Gaussian in (depth key, index) order; keys_in[g] = depth key of
// Gaussian g (0xFFFFFFFF = hits no tile: placed anywhere at or after the largest real key);
// mm = {max(~key), max(key)} over the real keys.  kA/vA/kB/vB: n-word ping-pong buffers (kA may be
// keys_in).  barrier_words: two words that are zero at the launch (the kernel leaves them zero:
// allocate once, clear once); clear_words: 16 words (may hold mm) that the last CTA out sets to
// zero once every CTA has read mm, or null.  Returns 0 on success, 1 on error, -1 if the device
// cannot launch cooperatively (the caller then uses radix_sort_pairs; nothing was cleared).
size_t resident_sort_scratch_bytes(unsigned n);
int resident_depth_sort(cudaStream_t stream, const unsigned *keys_in, unsigned *kA, unsigned *vA,
                        unsigned *kB, unsigned *vB, unsigned *order, unsigned n,
                        const unsigned *mm, void *scratch, int *err, unsigned *barrier_words,
                        unsigned *clear_words);

// ---- rasterisers (raster_fwd.cu / raster_bwd.cu) --------------------------------------------
struct RasterParams {
    unsigned img_w, img_h;
    unsigned tiles_x, tiles_y;
    unsigned block_width;
    unsigned band_y0, band_y1;  // tile rows rasterised by this call (screen-tile-row sharding)
    // sc_shift > 0 (fused render, block_width 16): `ids_sorted` is a uint2 array of row entries
    // {id, x0 | w << 16} listed per super-column of 2^sc_shift tiles, `tile_bins` the band-local
    // (rows x cols2) ranges of those lists (rowcol_lists.cu), final_idx positions in them, and the
    // hit words follow hit_quadrant_offset_super; a tile keeps the entries whose columns cover it
    unsigned sc_shift = 0, cols2 = 0;
};

// hit_words (may be null): the forward records which list entries each 8x8 quadrant accepted
// (raster_hits.cuh; hit_words_capacity(I, tiles) words); *wrote_hits says whether the kernel that
// ran does (only the default 16x16 two-pixel kernel).
int launch_raster_fwd(cudaStream_t stream, const RasterParams &p, const void *ids_sorted,
                      const int2 *tile_bins, const SplatRecord *records, const float *background,
                      float *final_Ts, int *final_idx, float *out_img,
                      unsigned *hit_words = nullptr, bool *wrote_hits = nullptr);

int launch_raster_bwd(cudaStream_t stream, const RasterParams &p, const void *ids_sorted,
                      const int2 *tile_bins, const SplatRecord *records, const float *background,
                      const float *final_Ts, const int *final_idx, const float *v_out,
                      const float *v_out_alpha, float *v_xy, float *v_xy_abs, float *v_conic,
                      float *v_colors, float *v_opacity, const unsigned *hit_words = nullptr,
                      bool *moments_out = nullptr);
// moments_out (fused render; may be null): if the hit-driven kernel ran without v_xy_abs it left RAW
// MOMENTS in v_opacity / v_xy / v_conic (raster_bwd.cu: MOMENTS) and sets *moments_out; the caller
// must then run launch_project_bwd with finish_moments (which also finalises v_opacity).

// image = background, final_Ts = 0, final_idx = 0  (reference num_intersects==0 path,
// ops.cu:119-127 + memsets :436-458)
int launch_fill_background(cudaStream_t stream, const RasterParams &p, const float *background,
                           float *final_Ts, int *final_idx, float *out_img);

}  // namespace jsb
