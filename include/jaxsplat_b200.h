/* jaxsplat_b200 -- C ABI of the B200-native (sm_100a) splatting hot path.
 *
 * Drop-in boundary for yklcs/jaxsplat's native layer.  Each entry point names the reference
 * interface it replaces (paths relative to the reference repository root).  Plain C types only:
 * no torch / pybind / C++ types cross this boundary, nothing here ever throws, and every call
 * enqueues work on the caller's stream only.
 *
 * The four jsb200_{project,rasterize}_{fwd,bwd} functions are XLA GPU custom-call targets
 * (legacy "API_VERSION_ORIGINAL/STATUS_RETURNING"-compatible signature) and take exactly the
 * operand-then-result buffer lists of the reference:
 *
 *   project_fwd    in : mean3ds(N,3) scales(N,3) quats(N,4) viewmat(4,4)
 *                  out: cov3ds(N,6) xys(N,2) depths(N) radii(N)i32 conics(N,3) compensation(N)
 *                       num_tiles_hit(N)i32 cum_tiles_hit(N)i32
 *   project_bwd    in : mean3ds scales quats viewmat cov3ds xys radii conics compensation
 *                       v_compensation(N) v_xy(N,2) v_depth(N) v_conic(N,3)
 *                  out: v_mean3d(N,3) v_scale(N,3) v_quat(N,4) v_cov2d(N,3) v_cov3d(N,6)
 *   rasterize_fwd  in : colors(N,3) opacities(N) background(3) xys depths radii conics
 *                       cum_tiles_hit
 *                  out: final_Ts(H,W) final_idx(H,W)i32 out_img(H,W,3)
 *   rasterize_bwd  in : colors opacities background xys depths radii conics cum_tiles_hit
 *                       final_Ts final_idx v_out_img(H,W,3) v_out_img_alpha(H,W)
 *                  out: v_colors(N,3) v_opacity(N) v_xy(N,2) v_xy_abs(N,2) v_conic(N,3)
 *
 * All arrays are C-contiguous float32 unless marked i32.  Results arrive uninitialised and are
 * fully written by the call.  `opaque` is the 80-byte descriptor from jsb200_make_descriptor.
 */
#ifndef JAXSPLAT_B200_H_
#define JAXSPLAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *jsb200_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define JSB200_API __attribute__((visibility("default")))
#else
#define JSB200_API
#endif

#define JSB200_DESCRIPTOR_BYTES 80

/* replaces ops::project::fwd::xla   (lib/ops.h:45-53, lib/ops.cu:12-61; registered by
 * lib/ffi.cu:6 as "project_fwd") */
JSB200_API void jsb200_project_fwd(jsb200_stream_t stream, void **buffers, const char *opaque,
                        size_t opaque_len);
/* replaces ops::project::bwd::xla   (lib/ops.h:58-66, lib/ops.cu:63-105; "project_bwd") */
JSB200_API void jsb200_project_bwd(jsb200_stream_t stream, void **buffers, const char *opaque,
                        size_t opaque_len);
/* replaces ops::rasterize::fwd::xla (lib/ops.h:71-79, lib/ops.cu:107-170; "rasterize_fwd") */
JSB200_API void jsb200_rasterize_fwd(jsb200_stream_t stream, void **buffers, const char *opaque,
                          size_t opaque_len);
/* replaces ops::rasterize::bwd::xla (lib/ops.h:84-92, lib/ops.cu:172-235; "rasterize_bwd") */
JSB200_API void jsb200_rasterize_bwd(jsb200_stream_t stream, void **buffers, const char *opaque,
                          size_t opaque_len);

/* The same four targets for XLA's API_VERSION_STATUS_RETURNING custom calls (the default
 * api_version of jaxlib.hlo_helpers.custom_call, which jaxsplat/_project/lowering.py:40-75 and
 * jaxsplat/_rasterize/lowering.py:40-73 use): XLA passes an XlaCustomCallStatus* as a fifth
 * argument.  The reference ignores it and throws std::runtime_error through the C frame instead
 * (lib/common.h:11-23, lib/ffi.h:19-21); these variants report a failure through
 * XlaCustomCallStatusSetFailure (resolved in the host process at the time of the failure) so the
 * error surfaces as an XlaRuntimeError in Python.  Behaviour on success is identical. */
typedef struct XlaCustomCallStatus_ jsb200_xla_status_t;
JSB200_API void jsb200_project_fwd_status(jsb200_stream_t stream, void **buffers,
                                          const char *opaque, size_t opaque_len,
                                          jsb200_xla_status_t *status);
JSB200_API void jsb200_project_bwd_status(jsb200_stream_t stream, void **buffers,
                                          const char *opaque, size_t opaque_len,
                                          jsb200_xla_status_t *status);
JSB200_API void jsb200_rasterize_fwd_status(jsb200_stream_t stream, void **buffers,
                                            const char *opaque, size_t opaque_len,
                                            jsb200_xla_status_t *status);
JSB200_API void jsb200_rasterize_bwd_status(jsb200_stream_t stream, void **buffers,
                                            const char *opaque, size_t opaque_len,
                                            jsb200_xla_status_t *status);

/* Fused render (no single reference entry point: replaces the chain
 * jaxsplat/__init__.py:54-70 = project (lib/ops.cu:12-61) then rasterize (lib/ops.cu:107-170) and,
 * for the backward, rasterize_bwd (lib/ops.cu:172-235) then project_bwd (lib/ops.cu:63-105)).
 *   render_fwd  in : means3d scales quats viewmat colors opacities background
 *               out: cov3ds xys depths radii conics compensation final_Ts final_idx out_img
 *                    (the two ops' results, same layouts, minus num_tiles_hit / cum_tiles_hit,
 *                    which nothing downstream reads)
 *   render_bwd  in : means3d scales quats viewmat colors opacities background cov3ds xys depths
 *                    radii conics compensation final_Ts final_idx v_out_img v_out_img_alpha
 *               out: v_mean3d v_scale v_quat v_colors v_opacity | v_xy v_conic v_cov2d v_cov3d
 *                    (v_xy (N,2) and v_conic (N,3) are accumulation scratch of the library: their
 *                    contents on return are unspecified -- the rasteriser backward may leave raw
 *                    moments there that the projection backward consumes; v_cov2d / v_cov3d are
 *                    the projection VJP's intermediate cotangents; v_xy_abs, which nothing
 *                    downstream of rasterize_bwd reads, is not computed)
 * v_out_img_alpha may be NULL (= zero cotangent; block_width 16 and 8 only); v_cov2d / v_cov3d may
 * be NULL (not written: nothing downstream of a render reads them).
 * Lean residuals: cov3ds and compensation may be NULL in render_fwd (not written) and in the
 * matching render_bwd, which then recomputes the 3D covariance from scales and quaternions -- 28 /
 * 24 bytes per Gaussian less.  Also valid for the per-view arrays of render_views_fwd / _bwd.
 * `opaque` is a full descriptor (f, c, glob_scale, clip_thresh set).  *token_out identifies the
 * binning this forward left in the per-(device, stream) workspace (0 = none); a render_bwd given
 * that token reuses it without any check or host synchronisation if no other binning ran on the
 * stream in between, and rebuilds it otherwise.  Results are bit-identical to the four-call
 * chain.  Return 0 on success. */
JSB200_API int jsb200_render_fwd(jsb200_stream_t stream, void **buffers, const char *opaque,
                                 size_t opaque_len, unsigned long long *token_out);
JSB200_API int jsb200_render_bwd(jsb200_stream_t stream, void **buffers, const char *opaque,
                                 size_t opaque_len, unsigned long long token);
/* jsb200_render_bwd for a caller that has already set the four atomically accumulated outputs
 * (v_colors, v_opacity, v_xy, v_conic: 36 bytes per Gaussian) to zero, ordered before this call on
 * `stream` -- e.g. on a side stream while the forward ran, joined with an event (the torch host
 * layer does: jaxsplat_b200/_render.py).  The library then issues no memset in front of the
 * rasteriser backward (36 MB at 1M Gaussians); everything else as jsb200_render_bwd. */
JSB200_API int jsb200_render_bwd_precleared(jsb200_stream_t stream, void **buffers,
                                            const char *opaque, size_t opaque_len,
                                            unsigned long long token);
/* Sets `bytes` bytes at `ptr` to zero on `stream` (cudaMemsetAsync; no kernel of a host framework):
 * what a host layer issues on its side stream for jsb200_render_bwd_precleared.  The reference
 * clears the same arrays with cudaMemsetAsync in front of its backward (lib/ops.cu:534-572). */
JSB200_API int jsb200_clear_async(jsb200_stream_t stream, void *ptr, size_t bytes);

/* Batched cameras (an API addition: the reference renders one camera per call,
 * jaxsplat/__init__.py:6-70, and a multi-view step loops over it in Python).  num_views cameras of
 * the same Gaussians in one call; per-view arrays are stacked on a leading view axis.
 *   render_views_fwd  in : means3d scales quats viewmats(B,4,4) colors opacities background
 *                     out: cov3ds(B,N,6) xys(B,N,2) depths(B,N) radii(B,N) conics(B,N,3)
 *                          compensation(B,N) final_Ts(B,H,W) final_idx(B,H,W) out_img(B,H,W,3)
 *   render_views_bwd  in : the 7 operands and 8 residual arrays above (final_Ts, final_idx last),
 *                          v_out_img(B,H,W,3), v_out_img_alpha(B,H,W) or NULL
 *                     out: v_mean3d v_scale v_quat v_colors v_opacity -- summed over the views
 * View b is bit-identical to jsb200_render_fwd with viewmats[b]; the gradients equal the sum of
 * the per-view jsb200_render_bwd results up to float summation order.  tokens_out / tokens hold
 * num_views binning tokens (may be NULL: the backward then re-bins every view).  The binnings of
 * the most recent "view_slots" views (jsb200_set_option, default 8) are kept for the backward.
 * Return 0 on success. */
JSB200_API int jsb200_render_views_fwd(jsb200_stream_t stream, void **buffers, const char *opaque,
                                       size_t opaque_len, unsigned num_views,
                                       unsigned long long *tokens_out);
JSB200_API int jsb200_render_views_bwd(jsb200_stream_t stream, void **buffers, const char *opaque,
                                       size_t opaque_len, unsigned num_views,
                                       const unsigned long long *tokens);

/* replaces make_descriptor (lib/ffi.cu:15-52): fills out[80] with the reference-compatible
 * descriptor (img_shape is given as H, W like the Python API).  Returns 0 on success. */
JSB200_API int jsb200_make_descriptor(unsigned num_points, unsigned img_h, unsigned img_w, float fx, float fy,
                           float cx, float cy, float glob_scale, float clip_thresh,
                           unsigned block_width, char *out);

/* Extension (no reference equivalent): restricts the rasterize calls that use this descriptor to
 * the tile rows [tile_row_begin, tile_row_end) -- screen-tile-row sharding of one large frame
 * across GPUs with the Gaussians replicated.  Stored in the descriptor's 4 pad bytes and marked by
 * a magic value in block_dim_2d.z (1 in reference descriptors): without the marker the pad bytes
 * -- which the reference's make_descriptor leaves uninitialised -- are ignored (= full frame).  Only pixels of the band are written;
 * final_idx then indexes the band's own sorted list.  Returns 0 on success. */
JSB200_API int jsb200_set_descriptor_band(char *desc80, unsigned tile_row_begin, unsigned tile_row_end);

/* replaces ops::sort_and_bin (lib/ops.h:31-41, lib/ops.cu:645-793): binning only; writes the
 * depth-sorted gaussian ids (num_intersects i32) and tile_bins (tiles x int2) to device buffers.
 * Returns 0 on success. */
JSB200_API int jsb200_sort_and_bin(jsb200_stream_t stream, const char *opaque, size_t opaque_len,
                        unsigned num_intersects, const float *xys, const float *depths,
                        const int *radii, const int *cum_tiles_hit, int *gaussian_ids_sorted,
                        int *tile_bins);

/* replaces ops::cumsum (lib/ops.h:24-29, lib/ops.cu:607-643): inclusive int32 prefix sum. */
JSB200_API int jsb200_cumsum(jsb200_stream_t stream, const int *input, int *output, int num_items);

/* Error channel (the reference throws std::runtime_error through the C frame,
 * lib/common.h:11-23).  Returns the sticky error text of the calling thread's last failing call,
 * or NULL; jsb200_clear_error resets it.  Error state is also mirrored process-wide so a host
 * framework thread can poll it: jsb200_error_count(). */
JSB200_API const char *jsb200_last_error(void);
JSB200_API void jsb200_clear_error(void);
JSB200_API unsigned long long jsb200_error_count(void);

/* Test / measurement hooks (no reference equivalent). */
/* key: "cull" (1 default; 0 disables the conservative warp-rectangle pre-filter),
 *      "cache" (1 default; 0 forces rasterize_bwd to re-bin instead of reusing the forward's),
 *      "rowcol" (1 default; 0 builds the per-tile lists with the older counting distribution of
 *               tile_lists.cu instead of the two-level row/column distribution),
 *      "fwdwarp" (1 default; 0 runs rasterize_fwd with the per-CTA batch kernel (two CTA barriers
 *               per batch of 128 entries) instead of four independent quadrant warps per tile),
 *      "bwdmom" (1 default; 0 makes the fused render's hit-driven rasteriser backward accumulate
 *               finished v_xy / v_conic / v_opacity instead of raw moments that the projection
 *               backward finishes),
 *      "supercols" (2 default; k > 0: the fused render (block_width 16) lists the 8-byte row
 *               entries per super-column of 2^k tiles and its rasterisers keep the entries whose
 *               tile columns cover their tile; 0: the exact per-tile id lists everywhere.  Must
 *               not change between a render_fwd and its render_bwd),
 *      "bwdhits" (1 default; 0 runs rasterize_bwd with the per-tile batch kernel even when the
 *               forward recorded which list entries every 8x8 quadrant accepted),
 *      "sort1k" (1 default; 0 orders the Gaussians by depth with four onesweep passes over all 32
 *               key bits instead of the single cooperative launch over the live bits),
 *      "pdl" (1 default; the kernels of a step that follow another kernel directly are launched
 *               with programmatic stream serialization: they become resident while their
 *               predecessor drains and wait for it in the kernel (griddepcontrol.wait) before
 *               their first global access; 0: plain launches),
 *      "timing" (0 default; 1 brackets every stage with CUDA events, see jsb200_get_stage_times),
 *      "view_slots" (8 default; binning workspaces a batched-camera call cycles through) */
JSB200_API int jsb200_set_option(const char *key, int value);
/* Copies the binning produced by the most recent rasterize/sort_and_bin call on `stream` into
 * caller device buffers (any may be NULL): ids_sorted[num_intersects] i32, tile_bins[tiles] int2,
 * keys64_sorted[num_intersects] i64 = (tile << 32 | depth bits) as the reference sorts them.
 * Returns num_intersects, or -1 on error. */
JSB200_API long long jsb200_debug_binning(jsb200_stream_t stream, const float *depths, int *ids_sorted,
                               int *tile_bins, long long *keys64_sorted);
/* Copies the hit words the most recent forward rasterisation on `stream` recorded for its backward
 * (layout: csrc/raster_hits.cuh) into a caller device buffer of max_words words.  Returns the
 * number of words that exist, or -1 (no forward with hit words cached). */
JSB200_API long long jsb200_debug_hit_words(jsb200_stream_t stream, unsigned *out_words,
                                            size_t max_words);
/* Per-stage device time, measured with CUDA events on `stream` while option "timing" is 1.
 * Synchronises the stream.  Stage order: project_fwd, bin_prep, depth_sort, emit, tile_scatter,
 * tile_offsets, raster_fwd, raster_bwd, project_bwd.  ms_out9 / calls_out9 receive accumulated
 * milliseconds and span counts; reset != 0 clears the accumulators.  Returns 0 on success. */
JSB200_API int jsb200_get_stage_times(jsb200_stream_t stream, double *ms_out9,
                                      unsigned long long *calls_out9, int reset);
/* counters since load: [0] rasterize_fwd calls, [1] rasterize_bwd calls, [2] binning runs,
 * [3] binning cache hits, [4] last num_intersects, [5] workspace bytes, [6] kernel launches,
 * [7] last row-entry count (sum of tile-box heights) */
JSB200_API void jsb200_get_stats(unsigned long long *out8);
/* the last list build: [0] intersections, [1] row entries, [2] (row entry, super-column) pairs the
 * prep counted, [3] log2 of the tiles per super-column the lists were built with (0 = exact lists) */
JSB200_API void jsb200_get_list_stats(unsigned long long *out4);
/* releases every workspace (all devices/streams). */
JSB200_API void jsb200_release(void);

/* ---- multi-GPU (no reference equivalent: yklcs/jaxsplat is single-device) ----
 * In-place sum of n_floats floats across the `world` ranks of one NVSwitch node, in ONE kernel:
 * every rank's buffer is the same symmetric allocation, mapped peer-to-peer (peer_bases[r] = rank
 * r's buffer as seen from this rank) and at a multicast address (multicast_base).  Rank r reduces
 * slice r inside the switch (multimem.ld_reduce) and broadcasts it (multimem.st).  The 16-byte
 * aligned block of >= 256 bytes at flags_offset_bytes (beyond the data, zero-initialised once)
 * holds the cross-GPU arrive/depart flags; `epoch` must grow by one per call, starting at 1.
 * n_floats must be a multiple of 4.  Returns 0 on success. */
JSB200_API int jsb200_multimem_allreduce_f32(jsb200_stream_t stream, void *multicast_base,
                                             void *const *peer_bases, size_t n_floats,
                                             size_t flags_offset_bytes, int rank, int world,
                                             unsigned epoch);

/* ---- optimiser step either side of the path -----------------------------------
 * One Adam update of a flat float array in one streaming kernel; the semantics of optax.adam(lr)
 * as used by the reference's examples/single_image.py:21-40 (b1 = 0.9, b2 = 0.999, eps = 1e-8,
 * bias-corrected moments, no weight decay).  `step` counts from 1.  m and v are the caller's
 * moment buffers (zero before the first step).  Returns 0 on success. */
JSB200_API int jsb200_adam_step(jsb200_stream_t stream, float *param, const float *grad, float *m,
                                float *v, size_t n, float lr, float b1, float b2, float eps,
                                unsigned long long step);

/* Mean-squared-error loss and its gradient in one streaming kernel (the loss of the reference's
 * examples/single_image.py:21-30 and of the benchmark metric: mean((img - target)^2)):
 * v_img[i] = (img[i] - target[i]) * scale (pass scale = 2 / n for d loss / d img), and
 * *loss = sum_i (img[i] - target[i])^2 / n (device float, cleared inside; summed with one atomic
 * per CTA, so its last bits depend on the order in which the CTAs finish).  Returns 0 on success. */
JSB200_API int jsb200_mse_loss_grad(jsb200_stream_t stream, const float *img, const float *target,
                                    size_t n, float scale, float *v_img, float *loss);
/* The same with an 8-bit target image (what a data loader delivers): element i is decoded as
 * float(target[i]) * (1 / 255) inside the kernel -- the value a separate u8 -> f32 decode pass would
 * have stored -- so the image never exists as floats in memory. */
JSB200_API int jsb200_mse_loss_grad_u8(jsb200_stream_t stream, const float *img,
                                       const unsigned char *target, size_t n, float scale,
                                       float *v_img, float *loss);

#ifdef __cplusplus
}
#endif
#endif /* JAXSPLAT_B200_H_ */
