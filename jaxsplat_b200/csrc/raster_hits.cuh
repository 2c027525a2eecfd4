// Hit words: what the forward rasteriser tells the backward (raster_fwd.cu -> raster_bwd.cu).
//
// For every tile, quadrant warp q (8x8 pixels) and group of 32 consecutive entries of the tile's
// sorted list (groups aligned to the list's first entry), one 32-bit word: bit b is set iff entry
// 32 g + b was visited by the warp and at least one of its 64 pixels -- not yet saturated at that
// point -- passed the accept test (sigma >= 0, alpha >= 1/255).  Every (pixel, Gaussian) pair the
// backward differentiates (entry <= final_idx of the pixel, same accept test: reference
// backward.cu:257-264 vs forward.cu:424-430) lies in such an entry.  A word is written for every
// group the warp entered while not all of its pixels were done, which covers every group up to
// the warp's largest final_idx.
//
// Storage: tile t (index inside the band) with list range [x, y) owns 4 * cap words starting at
// word 4 * ((x >> 5) + t), cap = (y >> 5) - (x >> 5) + 1 >= ceil((y - x) / 32); quadrant q's words
// are contiguous: [q * cap, q * cap + cap).  Ranges are concatenated in tile order, so the slices
// of different tiles never overlap; 4 * (I / 32 + tiles) words in all.
#pragma once
#include "common.cuh"

namespace jsb {

__host__ __device__ inline size_t hit_words_capacity(unsigned long long num_intersects,
                                                     unsigned tiles) {
    return 4 * ((size_t)(num_intersects >> 5) + tiles + 1);
}

__device__ __forceinline__ size_t hit_quadrant_offset(const int2 range, unsigned tile_local,
                                                      unsigned quad) {
    const unsigned g0 = (unsigned)range.x >> 5;
    const unsigned cap = ((unsigned)range.y >> 5) - g0 + 1u;
    return 4 * ((size_t)g0 + tile_local) + (size_t)quad * cap;
}

// block_width 8: a tile is one quadrant; tile t with list range [x, y) owns cap words starting at
// word (x >> 5) + t (inside the capacity above).
__device__ __forceinline__ size_t hit_tile8_offset(const int2 range, unsigned tile_local) {
    return (size_t)((unsigned)range.x >> 5) + tile_local;
}

// Super-column lists (fused render): the list belongs to a super-column of T = 2^shift tiles of
// one tile row, and each of its T tiles keeps its own four quadrant slices of the same length:
// super-tile s (band-local index) with list range [x, y) owns 4 T cap words starting at word
// 4 T ((x >> 5) + s); tile t of the super-column, quadrant q: [(4 t + q) cap, (4 t + q) cap + cap).
__host__ __device__ inline size_t hit_words_capacity_super(unsigned long long entries, unsigned supers,
                                                           unsigned shift) {
    return (size_t)(4u << shift) * ((size_t)(entries >> 5) + supers + 1);
}

__device__ __forceinline__ size_t hit_quadrant_offset_super(const int2 range, unsigned super_local,
                                                            unsigned tile_in_super, unsigned quad,
                                                            unsigned shift) {
    const unsigned g0 = (unsigned)range.x >> 5;
    const unsigned cap = ((unsigned)range.y >> 5) - g0 + 1u;
    return (size_t)(4u << shift) * ((size_t)g0 + super_local) + (size_t)(4u * tile_in_super + quad) * cap;
}

}  // namespace jsb
