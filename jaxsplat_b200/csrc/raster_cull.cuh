// Conservative warp-rectangle culling shared by the forward and backward tile rasterisers.
//
// A (pixel, Gaussian) pair contributes iff  sigma >= 0  and  min(clamp, opac * 2^(-sigma*log2e))
// >= 1/255  (reference forward.cu:424-430, backward.cu:257-264), i.e. iff sigma <= ln(255*opac).
// The exact per-pair test is evaluated by the rasterisers in the reference's rounding order; the
// test below only PRE-FILTERS whole 8x4-pixel warp rectangles: it computes the minimum of the
// (convex) quadratic sigma over the rectangle and discards the rectangle when that minimum
// exceeds the threshold by a safety margin that dominates every float rounding term.  It must
// never discard a pair the exact test accepts; tests/test_cull_exact.py checks bit-identical
// output against the un-culled kernel, including adversarial thin/diagonal conics.
#pragma once
#include "common.cuh"

namespace jsb {

constexpr int kRectW = 8, kRectH = 4;  // pixels per warp rectangle (lane = y*8 + x)
constexpr float kCullAbsMargin = 0.02f;
constexpr float kCullRelMargin = 1e-4f;

struct CullGauss {
    float gx, gy, cx, cy, cz;
    float icx, icz;   // reciprocals (only steer the clamped edge minimiser)
    float thresh;     // ln(255 opac) + margins
    bool never;       // opacity below the alpha cut: no pixel can accept
    bool always;      // conic not safely positive definite: do not cull
};

__device__ __forceinline__ CullGauss cull_setup(float gx, float gy, float opac, float cx,
                                                float cy, float cz, float tile_x0, float tile_y0,
                                                float tile_w, float tile_h) {
    CullGauss c;
    c.gx = gx; c.gy = gy; c.cx = cx; c.cy = cy; c.cz = cz;
    c.never = opac < kAlphaMin * 0.9999f;
    const float det = cx * cz - cy * cy;
    c.always = !(det > 0.f && cx > 0.f && cz > 0.f && det < 3.0e38f);
    c.icx = __fdividef(1.f, cx);
    c.icz = __fdividef(1.f, cz);
    // largest |d| over the whole tile bounds the magnitude of every term of sigma
    const float dxm = fmaxf(fabsf(gx - (tile_x0 + 0.5f)), fabsf(gx - (tile_x0 + tile_w - 0.5f)));
    const float dym = fmaxf(fabsf(gy - (tile_y0 + 0.5f)), fabsf(gy - (tile_y0 + tile_h - 0.5f)));
    const float amax = 0.5f * (cx * dxm * dxm + cz * dym * dym);
    c.thresh = __logf(255.f * fmaxf(opac, 1e-30f)) + kCullAbsMargin + kCullRelMargin * amax;
    if (!(c.thresh == c.thresh) || !(amax < 3.0e38f)) c.always = true;
    return c;
}

// The Gaussian-only part of the set-up, evaluated once per Gaussian by the prep kernels and kept
// in its record (SplatRecord::q2.y): ln(255 opac) + margins, with the relative margin taken over
// every pixel of every tile the Gaussian's box covers (|d| <= radius + tile width per axis, so it
// bounds the per-tile value of cull_setup); -inf when no pixel can accept (`never`), +inf when the
// conic is not safely positive definite (`always`: never cull).
__host__ __device__ __forceinline__ float cull_record_thresh(float opac, float cx, float cy, float cz,
                                                             float radius, float tile_w) {
    if (opac < kAlphaMin * 0.9999f) return -INFINITY;
    const float det = cx * cz - cy * cy;
    if (!(det > 0.f && cx > 0.f && cz > 0.f && det < 3.0e38f)) return INFINITY;
    const float dm = radius + tile_w;
    const float amax = 0.5f * (cx + cz) * dm * dm;
    const float t = logf(255.f * fmaxf(opac, 1e-30f)) + kCullAbsMargin + kCullRelMargin * amax;
    if (!(t == t) || !(amax < 3.0e38f)) return INFINITY;
    return t;
}

// set-up from a record that carries cull_record_thresh
__device__ __forceinline__ CullGauss cull_from_record(float gx, float gy, float cx, float cy, float cz,
                                                      float thresh) {
    CullGauss c;
    c.gx = gx; c.gy = gy; c.cx = cx; c.cy = cy; c.cz = cz;
    c.never = false;   // (both folded into thresh)
    c.always = false;
    // the reciprocals only steer the clamped edge minimiser (any value keeps the test
    // conservative): one MUFU each, without the range fix-ups of a division
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c.icx) : "f"(cx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c.icz) : "f"(cz));
    c.thresh = thresh;
    return c;
}

// may any pixel centre in [x0+0.5, x0+w-0.5] x [y0+0.5, y0+h-0.5] pass the exact test?
// (written as !(min > thresh): a NaN from a degenerate conic never culls)
__device__ __forceinline__ bool cull_rect_may_hit(const CullGauss &c, float x0, float y0, float w,
                                                  float h) {
    // delta = g - p ranges over [dx_lo, dx_hi] x [dy_lo, dy_hi]
    const float dx_hi = c.gx - (x0 + 0.5f), dx_lo = c.gx - (x0 + w - 0.5f);
    const float dy_hi = c.gy - (y0 + 0.5f), dy_lo = c.gy - (y0 + h - 0.5f);
    const float X = fminf(fmaxf(0.f, dx_lo), dx_hi);  // point of the range nearest to 0
    const float Y = fminf(fmaxf(0.f, dy_lo), dy_hi);
    // the minimum of a convex quadratic over a box that excludes its centre lies on an edge
    // facing the centre: x = X (if X != 0) or y = Y (if Y != 0); evaluate both branch-free.
    const float ya = fminf(fmaxf(-c.cy * X * c.icz, dy_lo), dy_hi);
    const float xb = fminf(fmaxf(-c.cy * Y * c.icx, dx_lo), dx_hi);
    const float qa = 0.5f * (c.cx * X * X + c.cz * ya * ya) + c.cy * X * ya;
    const float qb = 0.5f * (c.cx * xb * xb + c.cz * Y * Y) + c.cy * xb * Y;
    return !(fminf(qa, qb) > c.thresh);
}

// 8-bit mask over the 2x4 warp rectangles of a 16x16 tile (bit = wy*2 + wx)
__device__ __forceinline__ unsigned cull_mask16(float gx, float gy, float opac, float cx, float cy,
                                                float cz, float tile_x0, float tile_y0) {
    const CullGauss c = cull_setup(gx, gy, opac, cx, cy, cz, tile_x0, tile_y0, 16.f, 16.f);
    if (c.never) return 0u;
    if (c.always) return 0xFFu;
    if (!cull_rect_may_hit(c, tile_x0, tile_y0, 16.f, 16.f)) return 0u;
    unsigned m = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const float rx = tile_x0 + (float)((r & 1) * kRectW);
        const float ry = tile_y0 + (float)((r >> 1) * kRectH);
        if (cull_rect_may_hit(c, rx, ry, (float)kRectW, (float)kRectH)) m |= 1u << r;
    }
    return m;
}

}  // namespace jsb
