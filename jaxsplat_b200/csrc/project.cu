// Projection forward (+ fused cumulative tile count) and projection backward for sm_100a.
//
// Behavioural spec: reference lib/kernels/forward.cu:13-131 (project_gaussians_fwd),
// :465-568 (project_cov3d_ewa, scale_rot_to_cov3d), helpers.h:9-57,122-153,194-211 and
// lib/ops.cu:607-643 (cumsum); backward: backward.cu:345-570, helpers.h:60-98,131-192.
// Design (not a port): one thread per Gaussian, every load of a thread in flight before its first
// use, AoS results transposed through a warp-private strip of shared memory so they leave as full
// 128-byte rows (no CTA-wide barrier anywhere on the fused path); EVERY output element is written
// (no pre-zeroing memsets); the unfused op also computes the CTA-local part of cum_tiles_hit in the
// same launch and two tiny kernels (scan of the CTA totals, add prefix) finish the scan.  A
// single-launch chained look-back scan was measured slower on B200: ~half of the kernel's stall
// samples sat on the look-back wait, because these CTAs live only a few microseconds.  No CUB, no
// temp allocation.
#include "common.cuh"
#include "kernels.cuh"
#include "raster_cull.cuh"

namespace jsb {

namespace {

#ifndef JSB_PROJECT_THREADS
#define JSB_PROJECT_THREADS 256
#endif
constexpr int PT = JSB_PROJECT_THREADS;  // Gaussians (= threads) per CTA

struct ProjOut {
    float cov3d[6];
    float2 xy;
    float depth;
    int radius;
    float3 conic;
    float comp;
    int tiles;
};

// Exact operation order of the reference's compiled projection (see common.cuh rounding note).
// v = row-major world->camera 3x4; returns everything the reference would have left in its
// (pre-zeroed) outputs, including the partial writes of culled Gaussians.
__device__ __forceinline__ ProjOut project_one(float mx, float my, float mz, float sx, float sy,
                                               float sz, float4 q, const float *__restrict__ v,
                                               const ProjectFwdParams &p) {
    ProjOut o;
#pragma unroll
    for (int i = 0; i < 6; ++i) o.cov3d[i] = 0.f;
    o.xy = make_float2(0.f, 0.f);
    o.depth = 0.f;
    o.radius = 0;
    o.conic = make_float3(0.f, 0.f, 0.f);
    o.comp = 0.f;
    o.tiles = 0;

    const float v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
    const float v4 = v[4], v5 = v[5], v6 = v[6], v7 = v[7];
    const float v8 = v[8], v9 = v[9], v10 = v[10], v11 = v[11];

    // p_view = R p + t   (helpers.h:101-108)
    const float tx = fadd(v3, ffma(mz, v2, ffma(mx, v0, fmul(my, v1))));
    const float ty = fadd(v7, ffma(mz, v6, ffma(mx, v4, fmul(my, v5))));
    const float tz = fadd(v11, ffma(mz, v10, ffma(mx, v8, fmul(my, v9))));
    if (tz <= p.clip_thresh) return o;  // near-plane cull: everything stays zero

    // R(q), q = (w,x,y,z)   (helpers.h:140-153)
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    // ptxas contracts the reference's products into the sums below (read from its sm_100a SASS):
    // the six off-diagonal terms are single FMAs, the diagonal ones keep one separate add.
    const float yz = fmul(y, z), wz = fmul(w, z), wy = fmul(w, y);
    const float yy = fmul(y, y), zz = fmul(z, z);
    float t0 = ffma(w, x, yz);
    const float r12 = fadd(t0, t0);          // 2(wx + yz)
    t0 = ffma(-w, x, yz);
    const float r21 = fadd(t0, t0);          // 2(yz - wx)
    t0 = ffma(x, y, wz);
    const float r01 = fadd(t0, t0);          // 2(xy + wz)   column 0, row 1
    t0 = ffma(x, z, -wy);
    const float r02 = fadd(t0, t0);          // 2(xz - wy)
    t0 = ffma(x, y, -wz);
    const float r10 = fadd(t0, t0);          // 2(xy - wz)
    t0 = ffma(x, z, wy);
    const float r20 = fadd(t0, t0);          // 2(xz + wy)
    t0 = fadd(yy, zz);
    const float r00 = fsub(1.f, fadd(t0, t0));
    t0 = ffma(x, x, zz);
    const float r11 = fsub(1.f, fadd(t0, t0));
    t0 = ffma(x, x, yy);
    const float r22 = fsub(1.f, fadd(t0, t0));

    // M = R diag(glob_scale * s); cov3d = M M^T   (forward.cu:547-568)
    const float s0 = fmul(p.glob_scale, sx), s1 = fmul(p.glob_scale, sy),
                s2 = fmul(p.glob_scale, sz);
    const float a0 = fmul(s0, r00), b0 = fmul(s0, r01), c0 = fmul(s0, r02);
    const float a1 = fmul(s1, r10), b1 = fmul(s1, r11), c1 = fmul(s1, r12);
    const float a2 = fmul(s2, r20), b2 = fmul(s2, r21), c2 = fmul(s2, r22);
    const float k0 = ffma(a2, a2, ffma(a0, a0, fmul(a1, a1)));
    const float k1 = ffma(b2, a2, ffma(b0, a0, fmul(b1, a1)));
    const float k2 = ffma(c2, a2, ffma(c0, a0, fmul(c1, a1)));
    const float k3 = ffma(b2, b2, ffma(b0, b0, fmul(b1, b1)));
    const float k4 = ffma(c2, b2, ffma(c0, b0, fmul(c1, b1)));
    const float k5 = ffma(c2, c2, ffma(c0, c0, fmul(c1, c1)));
    o.cov3d[0] = k0; o.cov3d[1] = k1; o.cov3d[2] = k2;
    o.cov3d[3] = k3; o.cov3d[4] = k4; o.cov3d[5] = k5;

    // EWA projection   (forward.cu:465-544); tan_fov is the reference's FP64 expression,
    // evaluated on the host in double and rounded once to float.
    const float lim_x = fmul(p.tan_fovx, 1.3f), lim_y = fmul(p.tan_fovy, 1.3f);
    const float cxl = fmul(tz, fminf(lim_x, fmaxf(-lim_x, fdiv(tx, tz))));
    const float cyl = fmul(tz, fminf(lim_y, fmaxf(-lim_y, fdiv(ty, tz))));
    const float rz = frcp(tz);
    const float rz2 = fmul(rz, rz);
    const float j00 = fmul(p.fx, rz), j11 = fmul(p.fy, rz);
    const float j02 = fmul(rz2, fmul(cxl, -p.fx));
    const float j12 = fmul(rz2, fmul(cyl, -p.fy));
    // T = J W (rows 0,1)
    const float T0x = ffma(v8, j02, fmul(v0, j00)), T1x = ffma(v8, j12, fmul(v4, j11));
    const float T0y = ffma(v9, j02, fmul(v1, j00)), T1y = ffma(v9, j12, fmul(v5, j11));
    const float T0z = ffma(v10, j02, fmul(v2, j00)), T1z = ffma(v10, j12, fmul(v6, j11));
    // T V
    const float u00 = ffma(k2, T0z, ffma(k0, T0x, fmul(k1, T0y)));
    const float u10 = ffma(k2, T1z, ffma(k0, T1x, fmul(k1, T1y)));
    const float u01 = ffma(k4, T0z, ffma(k1, T0x, fmul(k3, T0y)));
    const float u11 = ffma(k4, T1z, ffma(k1, T1x, fmul(k3, T1y)));
    const float u02 = ffma(k5, T0z, ffma(k2, T0x, fmul(k4, T0y)));
    const float u12 = ffma(k5, T1z, ffma(k2, T1x, fmul(k4, T1y)));
    // (T V) T^T
    const float c00 = ffma(T0z, u02, ffma(T0x, u00, fmul(T0y, u01)));
    const float c01 = ffma(T0z, u12, ffma(T0x, u10, fmul(T0y, u11)));
    const float c11 = ffma(T1z, u12, ffma(T1x, u10, fmul(T1y, u11)));
    const float c01sq = fmul(c01, c01);
    const float det_orig = ffma(c00, c11, -c01sq);  // contracted in the reference SASS
    const float A = fadd(c00, 0.3f), B = fadd(c11, 0.3f);
    const float det = ffma(A, B, -c01sq);           // contracted
    const float comp_sq = fmaxf(fdiv(det_orig, det), 0.f);
    if (det == 0.f) return o;  // compute_cov2d_bounds failed: only cov3d was written

    // conic + 3-sigma radius   (helpers.h:34-57)
    const float inv = frcp(det);
    o.conic.x = fmul(B, inv);
    o.conic.y = fmul(c01, -inv);
    o.conic.z = fmul(A, inv);
    const float bb = fmul(fadd(A, B), 0.5f);
    const float sq = fsqrt(fmaxf(ffma(bb, bb, -det), 0.1f));  // contracted
    const float lam = fmaxf(fadd(bb, sq), fsub(bb, sq));
    const float radius = ceilf(fmul(fsqrt(lam), 3.f));

    // pixel centre (helpers.h:122-128) and tile box (helpers.h:9-32)
    const float rw = frcp(fadd(tz, 1e-6f));
    const float px = ffma(p.fx, fmul(tx, rw), p.cx);
    const float py = ffma(p.fy, fmul(ty, rw), p.cy);
    unsigned x0, x1, y0, y1;
    tile_bbox(px, py, radius, p.block_width_f, p.tiles_x, p.tiles_y, x0, x1, y0, y1);
    const int area = (int)((x1 - x0) * (y1 - y0));
    if (area <= 0) return o;  // off-screen: cov3d and conic were written, rest zero

    o.tiles = area;
    o.depth = tz;
    o.radius = __float2int_rz(radius);
    o.xy = make_float2(px, py);
    o.comp = fsqrt(comp_sq);
    return o;
}

// ---- warp-private AoS staging ------------------------------------------------------------------
// The operator boundary fixes AoS (N,3)/(N,6) float arrays.  Loads of a 12- or 24-byte-stride
// record are issued directly (the three / six scalar loads of a warp hit the same few cache lines,
// L1 serves the repeats); stores would each touch every sector of the warp's span with 4 bytes,
// so they are transposed through a warp-private strip of shared memory and leave as full
// 128-byte rows.  Only __syncwarp is involved: warps of a CTA never wait for each other, which is
// what the earlier CTA-wide slab staging (two to four __syncthreads per CTA) cost -- 54 us against
// the reference's 22 us for the same projection at 1M Gaussians.
constexpr int kWarpStripFloats = 32 * 12;  // one SplatRecord per lane

// lane's K consecutive floats -> out[first + lane*K .. ], K floats per lane, coalesced 4-byte rows
template <int K>
__device__ __forceinline__ void warp_store_aos(float *strip, const float (&v)[K],
                                               float *__restrict__ out, long long first,
                                               long long total, int lane) {
    __syncwarp();
    if (K % 2 == 0) {
#pragma unroll
        for (int i = 0; i < K / 2; ++i)
            reinterpret_cast<float2 *>(strip)[lane * (K / 2) + i] = make_float2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) strip[lane * K + i] = v[i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const long long e = first + i * 32 + lane;
        if (e < total) out[e] = strip[i * 32 + lane];
    }
}

// FUSED (fused render, full-frame band): the kernel also does bin_prep's job (binning.cu) for the
// Gaussian it just projected -- packed rasteriser record, depth key, tile box, the per-CTA totals
// of intersections and row entries and the range of the real depth keys -- so the projected arrays
// are not read back by a second kernel.
template <bool FUSED>
__global__ void __launch_bounds__(PT)
project_fwd_kernel(const ProjectFwdParams p, const float *__restrict__ means3d,
                   const float *__restrict__ scales, const float4 *__restrict__ quats,
                   const float *__restrict__ viewmat, float *__restrict__ cov3ds,
                   float2 *__restrict__ xys, float *__restrict__ depths, int *__restrict__ radii,
                   float *__restrict__ conics, float *__restrict__ compensation,
                   int *__restrict__ num_tiles_hit, int *__restrict__ cum_tiles_hit,
                   unsigned *__restrict__ block_totals, const FusedPrep f) {
    __shared__ __align__(16) float s_strip[PT / 32][kWarpStripFloats];
    __shared__ unsigned s_ws[PT / 32 + 1];
    __shared__ unsigned s_tot[PT / 32][5];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bid = blockIdx.x;
    const long long idx = (long long)bid * PT + tid;
    const long long wfirst = (long long)bid * PT + warp * 32;  // first Gaussian of this warp
    const long long n = p.num_points;
    float *strip = s_strip[warp];

    ProjOut o;
#pragma unroll
    for (int i = 0; i < 6; ++i) o.cov3d[i] = 0.f;
    o.xy = make_float2(0.f, 0.f);
    o.depth = 0.f; o.radius = 0; o.conic = make_float3(0.f, 0.f, 0.f); o.comp = 0.f; o.tiles = 0;
    const bool live = idx < n;
    float cr = 0.f, cg = 0.f, cb = 0.f, opac = 0.f;
    chain_wait();
    if (live) {
        // every load of the thread is in flight before the first use
        const float mx = __ldg(means3d + idx * 3), my = __ldg(means3d + idx * 3 + 1),
                    mz = __ldg(means3d + idx * 3 + 2);
        const float sx = __ldg(scales + idx * 3), sy = __ldg(scales + idx * 3 + 1),
                    sz = __ldg(scales + idx * 3 + 2);
        const float4 q = __ldg(quats + idx);
        if (FUSED) {
            cr = __ldg(f.colors + idx * 3); cg = __ldg(f.colors + idx * 3 + 1);
            cb = __ldg(f.colors + idx * 3 + 2);
            opac = __ldg(f.opacities + idx);
        }
        o = project_one(mx, my, mz, sx, sy, sz, q, viewmat, p);
    }
    unsigned rows = 0, sup = 0, kinv = 0, kmax = 0;
    if (FUSED) {
        unsigned key = 0xFFFFFFFFu;
        uint2 box = make_uint2(0u, 0u);  // {x0 | y0 << 16, width | height << 16}
        if (live && o.radius > 0) {
            // same expression as bin_prep_kernel: box from the stored xy and integer radius
            unsigned x0, x1, y0, y1;
            tile_bbox(o.xy.x, o.xy.y, (float)o.radius, p.block_width_f, p.tiles_x, p.tiles_y,
                      x0, x1, y0, y1);
            if (y1 > y0 && x1 > x0) {
                rows = y1 - y0;
                // (row entry, super-column) pairs: the level-2 output of the super-column lists
                sup = rows * (((x1 - 1u) >> p.sc_shift) - (x0 >> p.sc_shift) + 1u);
                box = make_uint2(x0 | (y0 << 16), (x1 - x0) | ((y1 - y0) << 16));
            }
            key = __float_as_uint(o.depth);
            kinv = ~key;
            kmax = key;
        }
        if (live) {
            f.depth_keys[idx] = key;
            f.ids[idx] = (unsigned)idx;
            f.box_by_id[idx] = box;
        }
        // the 48-byte records of the warp leave as 3 coalesced 16-byte rows per lane
        __syncwarp();
        float4 *s4 = reinterpret_cast<float4 *>(strip);
        s4[lane * 3 + 0] = make_float4(o.xy.x, o.xy.y, opac, o.conic.x);
        s4[lane * 3 + 1] = make_float4(o.conic.y, o.conic.z, cr, cg);
        s4[lane * 3 + 2] = make_float4(cb, cull_record_thresh(opac, o.conic.x, o.conic.y, o.conic.z, (float)o.radius,
                                                            p.block_width_f), 0.f, 0.f);
        __syncwarp();
        float4 *rec4 = reinterpret_cast<float4 *>(f.records) + wfirst * 3;
        const long long rec_left = (n - wfirst) * 3;  // float4s of this warp that exist
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (i * 32 + lane < rec_left) rec4[i * 32 + lane] = s4[i * 32 + lane];
    }
    if (live) {
        xys[idx] = o.xy;
        depths[idx] = o.depth;
        radii[idx] = o.radius;
        if (compensation) compensation[idx] = o.comp;
        // (the fused render passes null: nothing downstream of it reads the tile counts)
        if (num_tiles_hit) num_tiles_hit[idx] = o.tiles;
    }
    // (null = lean residuals of the fused render: the projection backward recomputes them)
    if (cov3ds) warp_store_aos<6>(strip, o.cov3d, cov3ds, wfirst * 6, n * 6, lane);
    const float cn[3] = {o.conic.x, o.conic.y, o.conic.z};
    if (conics) warp_store_aos<3>(strip, cn, conics, wfirst * 3, n * 3, lane);

    if (cum_tiles_hit) {
        // CTA-local inclusive count; add_block_prefix_kernel adds the scanned CTA totals
        unsigned total;
        const unsigned excl = cta_exclusive_scan<PT>((unsigned)o.tiles, &total, s_ws);  // syncs
        if (tid == 0) block_totals[bid] = total;
        if (live) cum_tiles_hit[idx] = (int)(excl + (unsigned)o.tiles);
    }
    if (FUSED) {
        // one set of same-address atomics per CTA (per-warp atomics on these words serialise in
        // the L2 and were measured as the kernel's critical path)
        const unsigned tiles = __reduce_add_sync(0xffffffffu, (unsigned)o.tiles);
        rows = __reduce_add_sync(0xffffffffu, rows);
        sup = __reduce_add_sync(0xffffffffu, sup);
        kinv = __reduce_max_sync(0xffffffffu, kinv);
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if (lane == 0) {
            s_tot[warp][0] = tiles; s_tot[warp][1] = rows; s_tot[warp][2] = kinv; s_tot[warp][3] = kmax;
            s_tot[warp][4] = sup;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long t0 = 0, t1 = 0, t2 = 0;
            unsigned m0 = 0, m1 = 0;
#pragma unroll
            for (int w = 0; w < PT / 32; ++w) {
                t0 += s_tot[w][0]; t1 += s_tot[w][1]; t2 += s_tot[w][4];
                m0 = max(m0, s_tot[w][2]); m1 = max(m1, s_tot[w][3]);
            }
            if (t0) {
                atomicAdd(f.totals + 1, t0);
                atomicAdd(f.totals + 4, t1);
                atomicAdd(f.totals + 6, t2);
            }
            if (m0 | m1) {  // range of the real depth keys: [~max(~key), max(key)]
                unsigned *mm = reinterpret_cast<unsigned *>(f.totals + 5);
                atomicMax(mm, m0);
                atomicMax(mm + 1, m1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Projection backward.  Float outputs only (tolerance-compared).  The maths is that of
// backward.cu:345-570 (project_pix_vjp, cov2d_to_conic_vjp, cov2d_to_compensation_vjp,
// project_cov3d_ewa_vjp with the unclamped t of SURVEY Q2, scale_rot_to_cov3d_vjp,
// quat_to_rotmat_vjp), expanded by hand on the operands' structure instead of generic 3x3
// products: J has four non-zeros, the 2x2 cotangent G and the 3x3 covariances are symmetric, so
//   T   = J W                 two rows T0, T1
//   v_V = T^T G T             = T0 (x) A0 + T1 (x) A1          with A = G T (two rows)
//   v_T = 2 G T V             = 2 A V                          (two rows)
//   v_J = v_T W^T             only the four entries the t-gradient reads
//   v_M = 2 v_V M,  v_s = gs diag(R^T v_M),  v_R = v_M S
// ~300 instructions per Gaussian instead of ~1100, no shared-memory slabs, no CTA barriers.
// v_cov2d / v_cov3d may be null (the fused render has no use for the intermediate cotangents).
// ---------------------------------------------------------------------------------------------
// FINISH_MOMENTS (fused render after the hit-driven rasteriser backward, raster_bwd.cu: MOMENTS):
// v_xy / v_conic / moment_opacity hold raw moments (Mx, My) / (Mxx, Mxy, Myy) / M0 of
// opacity * vis * v_alpha; with the conic and the opacity at hand they become
//   v_xy = -(cx Mx + cy My, cy Mx + cz My), v_conic = -(Mxx / 2, Mxy, Myy / 2), v_opacity = M0 / opacity
// here, and moment_opacity[i] is overwritten with v_opacity.
// RECOMPUTE (lean residuals of the fused render: cov3ds == null): the 3D covariance is re-derived
// from the rotation and the scales this kernel needs anyway (21 multiply-adds) instead of being
// written by the forward and read back here: 24 bytes per Gaussian less in each direction.
// (Recomputing the conic as well -- the whole EWA projection -- was measured: the forward gained
// 4.5 us, this kernel lost 4 us; profiles/r03a_bench_C3_recompute_conic.json.)
template <bool FINISH_MOMENTS, bool RECOMPUTE>
__global__ void __launch_bounds__(PT)
project_bwd_kernel(const ProjectBwdParams p, const float *__restrict__ means3d,
                   const float *__restrict__ scales, const float4 *__restrict__ quats,
                   const float *__restrict__ viewmat, const float *__restrict__ cov3ds,
                   const int *__restrict__ radii, const float *__restrict__ conics,
                   const float *__restrict__ compensation,
                   const float *__restrict__ v_compensation, const float2 *__restrict__ v_xy,
                   const float *__restrict__ v_depth, const float *__restrict__ v_conic,
                   float *__restrict__ v_mean3d, float *__restrict__ v_scale,
                   float *__restrict__ v_quat, float *__restrict__ v_cov2d,
                   float *__restrict__ v_cov3d, const float *__restrict__ opacities,
                   float *__restrict__ moment_opacity) {
    __shared__ __align__(16) float s_strip[PT / 32][32 * 6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long idx = (long long)blockIdx.x * PT + tid;
    const long long wfirst = (long long)blockIdx.x * PT + warp * 32;
    const long long n = p.num_points;
    float *strip = s_strip[warp];

    float o_mean[3] = {0.f, 0.f, 0.f}, o_scale[3] = {0.f, 0.f, 0.f};
    float o_cov2d[3] = {0.f, 0.f, 0.f}, o_cov3d[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float o_quat[4] = {0.f, 0.f, 0.f, 0.f};
    const bool live = idx < n;
    // p.residuals_ready (the fused render's backward, right behind its own rasteriser backward):
    // everything the FORWARD left -- parameters, radii, conics, the camera -- was complete before
    // the predecessor started, so a CTA that became resident while the predecessor drains loads it
    // before it waits; only the cotangents / moments are read behind the wait.
    if (!p.residuals_ready) chain_wait();
    const bool vis = live && __ldg(radii + idx) > 0;
    float mx = 0.f, my = 0.f, mz = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    float cx = 0.f, cy = 0.f, cz = 0.f, opac = 1.f;
    float k0 = 0.f, k1 = 0.f, k2 = 0.f, k3 = 0.f, k4 = 0.f, k5 = 0.f;
    if (vis) {
        mx = __ldg(means3d + idx * 3); my = __ldg(means3d + idx * 3 + 1); mz = __ldg(means3d + idx * 3 + 2);
        sc0 = __ldg(scales + idx * 3); sc1 = __ldg(scales + idx * 3 + 1); sc2 = __ldg(scales + idx * 3 + 2);
        q = __ldg(quats + idx);
        cx = __ldg(conics + idx * 3); cy = __ldg(conics + idx * 3 + 1); cz = __ldg(conics + idx * 3 + 2);
        if (FINISH_MOMENTS) opac = __ldg(opacities + idx);
        if (!RECOMPUTE) {
            k0 = __ldg(cov3ds + idx * 6); k1 = __ldg(cov3ds + idx * 6 + 1);
            k2 = __ldg(cov3ds + idx * 6 + 2); k3 = __ldg(cov3ds + idx * 6 + 3);
            k4 = __ldg(cov3ds + idx * 6 + 4); k5 = __ldg(cov3ds + idx * 6 + 5);
        }
    }
    const float *v = viewmat;
    const float w00 = __ldg(v + 0), w01 = __ldg(v + 1), w02 = __ldg(v + 2), w03 = __ldg(v + 3);
    const float w10 = __ldg(v + 4), w11 = __ldg(v + 5), w12 = __ldg(v + 6), w13 = __ldg(v + 7);
    const float w20 = __ldg(v + 8), w21 = __ldg(v + 9), w22 = __ldg(v + 10), w23 = __ldg(v + 11);
    if (p.residuals_ready) chain_wait();
    if (vis) {
        // R(q), q = (w, x, y, z), and the scaled axes: also what the scale / quaternion VJP needs
        const float w = q.x, x = q.y, y = q.z, z = q.w;
        const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - w * z), R02 = 2.f * (x * z + w * y);
        const float R10 = 2.f * (x * y + w * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - w * x);
        const float R20 = 2.f * (x * z - w * y), R21 = 2.f * (y * z + w * x), R22 = 1.f - 2.f * (x * x + y * y);
        const float gs = p.glob_scale;
        const float s0 = gs * sc0, s1 = gs * sc1, s2 = gs * sc2;
        if (RECOMPUTE) {  // V = R S S R^T
            const float e0 = s0 * s0, e1 = s1 * s1, e2 = s2 * s2;
            k0 = e0 * R00 * R00 + e1 * R01 * R01 + e2 * R02 * R02;
            k1 = e0 * R00 * R10 + e1 * R01 * R11 + e2 * R02 * R12;
            k2 = e0 * R00 * R20 + e1 * R01 * R21 + e2 * R02 * R22;
            k3 = e0 * R10 * R10 + e1 * R11 * R11 + e2 * R12 * R12;
            k4 = e0 * R10 * R20 + e1 * R11 * R21 + e2 * R12 * R22;
            k5 = e0 * R20 * R20 + e1 * R21 * R21 + e2 * R22 * R22;
        }
        float gx = __ldg(v_conic + idx * 3), gy = 0.5f * __ldg(v_conic + idx * 3 + 1),
              gz = __ldg(v_conic + idx * 3 + 2);
        float2 vxy = __ldg(v_xy + idx);
        if (FINISH_MOMENTS) {
            const float mmx = vxy.x, mmy = vxy.y;
            vxy = make_float2(-(cx * mmx + cy * mmy), -(cy * mmx + cz * mmy));
            gx *= -0.5f;   // v_conic.x = -Mxx / 2
            gy = -gy;      // v_conic.y / 2 = -Mxy / 2
            gz *= -0.5f;   // v_conic.z = -Myy / 2
            const float m0 = moment_opacity[idx];
            moment_opacity[idx] = m0 != 0.f ? m0 / opac : 0.f;
        }
        const float vz = v_depth ? __ldg(v_depth + idx) : 0.f;  // null = zero cotangent (fused render)
        const float fx = p.fx, fy = p.fy;
        // p_view
        const float pvx = w00 * mx + w01 * my + w02 * mz + w03;
        const float pvy = w10 * mx + w11 * my + w12 * mz + w13;
        const float pvz = w20 * mx + w21 * my + w22 * mz + w23;
        // project_pix_vjp (helpers.h:131-138) then R^T (helpers.h:90-98), + depth term
        const float rw = 1.f / (pvz + 1e-6f);
        const float vpx = fx * vxy.x, vpy = fy * vxy.y;
        const float vvx = vpx * rw, vvy = vpy * rw;
        const float vvz = -(vpx * pvx + vpy * pvy) * rw * rw + vz;
        float vm_x = w00 * vvx + w10 * vvy + w20 * vvz;
        float vm_y = w01 * vvx + w11 * vvy + w21 * vvz;
        float vm_z = w02 * vvx + w12 * vvy + w22 * vvz;

        // cov2d_to_conic_vjp (helpers.h:60-71): v_Sigma = -X G X, X = [[cx,cy],[cy,cz]]
        const float p00 = cx * gx + cy * gy, p01 = cx * gy + cy * gz;   // (X G) row 0
        const float p10 = cy * gx + cz * gy, p11 = cy * gy + cz * gz;   // (X G) row 1
        float vc_x = -(p00 * cx + p01 * cy);
        float vc_y = -((p00 * cy + p01 * cz) + (p10 * cx + p11 * cy));
        float vc_z = -(p10 * cy + p11 * cz);
        // cov2d_to_compensation_vjp (helpers.h:73-87): the reference evaluates this in FP64
        if (v_compensation) {
            const float vcomp = __ldg(v_compensation + idx);
            if (vcomp != 0.f) {
                const float comp = __ldg(compensation + idx);
                const float inv_det = cx * cz - cy * cy;
                const float om = 1 - comp * comp;
                const float vs = (float)((double)vcomp * 0.5 / ((double)comp + 1e-6));
                vc_x += (float)((double)vs * ((double)(om * cx) - 0.3 * (double)inv_det));
                vc_y += 2 * vs * (om * cy);
                vc_z += (float)((double)vs * ((double)(om * cz) - 0.3 * (double)inv_det));
            }
        }
        o_cov2d[0] = vc_x; o_cov2d[1] = vc_y; o_cov2d[2] = vc_z;

        // project_cov3d_ewa_vjp (backward.cu:421-524); note: unclamped t
        const float rz = 1.f / pvz, rz2 = rz * rz;
        const float j00 = fx * rz, j11 = fy * rz, j02 = -fx * pvx * rz2, j12 = -fy * pvy * rz2;
        const float T00 = j00 * w00 + j02 * w20, T01 = j00 * w01 + j02 * w21, T02 = j00 * w02 + j02 * w22;
        const float T10 = j11 * w10 + j12 * w20, T11 = j11 * w11 + j12 * w21, T12 = j11 * w12 + j12 * w22;
        const float g00 = vc_x, g01 = 0.5f * vc_y, g11 = vc_z;
        // A = G T
        const float A00 = g00 * T00 + g01 * T10, A01 = g00 * T01 + g01 * T11, A02 = g00 * T02 + g01 * T12;
        const float A10 = g01 * T00 + g11 * T10, A11 = g01 * T01 + g11 * T11, A12 = g01 * T02 + g11 * T12;
        // v_V = T^T A (symmetric); packed upper triangle, off-diagonals doubled
        const float V00 = T00 * A00 + T10 * A10, V11 = T01 * A01 + T11 * A11, V22 = T02 * A02 + T12 * A12;
        const float V01 = T00 * A01 + T10 * A11, V02 = T00 * A02 + T10 * A12, V12 = T01 * A02 + T11 * A12;
        o_cov3d[0] = V00; o_cov3d[1] = V01 + V01; o_cov3d[2] = V02 + V02;
        o_cov3d[3] = V11; o_cov3d[4] = V12 + V12; o_cov3d[5] = V22;
        // v_T = 2 A V (V = cov3d, symmetric)
        const float vT00 = 2.f * (A00 * k0 + A01 * k1 + A02 * k2);
        const float vT01 = 2.f * (A00 * k1 + A01 * k3 + A02 * k4);
        const float vT02 = 2.f * (A00 * k2 + A01 * k4 + A02 * k5);
        const float vT10 = 2.f * (A10 * k0 + A11 * k1 + A12 * k2);
        const float vT11 = 2.f * (A10 * k1 + A11 * k3 + A12 * k4);
        const float vT12 = 2.f * (A10 * k2 + A11 * k4 + A12 * k5);
        // v_J = v_T W^T: the four entries the t-gradient reads
        const float vJ00 = vT00 * w00 + vT01 * w01 + vT02 * w02;
        const float vJ02 = vT00 * w20 + vT01 * w21 + vT02 * w22;
        const float vJ11 = vT10 * w10 + vT11 * w11 + vT12 * w12;
        const float vJ12 = vT10 * w20 + vT11 * w21 + vT12 * w22;
        const float rz3 = rz2 * rz;
        const float vt_x = -fx * rz2 * vJ02;
        const float vt_y = -fy * rz2 * vJ12;
        const float vt_z = -fx * rz2 * vJ00 + 2.f * fx * pvx * rz3 * vJ02 - fy * rz2 * vJ11 +
                           2.f * fy * pvy * rz3 * vJ12;
        vm_x += vt_x * w00 + vt_y * w10 + vt_z * w20;
        vm_y += vt_x * w01 + vt_y * w11 + vt_z * w21;
        vm_z += vt_x * w02 + vt_y * w12 + vt_z * w22;
        o_mean[0] = vm_x; o_mean[1] = vm_y; o_mean[2] = vm_z;

        // scale_rot_to_cov3d_vjp (backward.cu:528-570): M = R S, v_M = 2 v_V M
        // B = v_V R (3x3); v_M[r][c] = 2 s_c B[r][c]
        const float B00 = V00 * R00 + V01 * R10 + V02 * R20, B01 = V00 * R01 + V01 * R11 + V02 * R21,
                    B02 = V00 * R02 + V01 * R12 + V02 * R22;
        const float B10 = V01 * R00 + V11 * R10 + V12 * R20, B11 = V01 * R01 + V11 * R11 + V12 * R21,
                    B12 = V01 * R02 + V11 * R12 + V12 * R22;
        const float B20 = V02 * R00 + V12 * R10 + V22 * R20, B21 = V02 * R01 + V12 * R11 + V22 * R21,
                    B22 = V02 * R02 + V12 * R12 + V22 * R22;
        // v_scale_c = gs * sum_r R[r][c] v_M[r][c]
        o_scale[0] = gs * 2.f * s0 * (R00 * B00 + R10 * B10 + R20 * B20);
        o_scale[1] = gs * 2.f * s1 * (R01 * B01 + R11 * B11 + R21 * B21);
        o_scale[2] = gs * 2.f * s2 * (R02 * B02 + R12 * B12 + R22 * B22);
        // v_R[r][c] = v_M[r][c] s_c = 2 s_c^2 B[r][c]
        const float e0 = 2.f * s0 * s0, e1 = 2.f * s1 * s1, e2 = 2.f * s2 * s2;
        const float vR00 = e0 * B00, vR01 = e1 * B01, vR02 = e2 * B02;
        const float vR10 = e0 * B10, vR11 = e1 * B11, vR12 = e2 * B12;
        const float vR20 = e0 * B20, vR21 = e1 * B21, vR22 = e2 * B22;
        // quat_to_rotmat_vjp (helpers.h:155-192)
        o_quat[0] = 2.f * (x * (vR21 - vR12) + y * (vR02 - vR20) + z * (vR10 - vR01));
        o_quat[1] = 2.f * (-2.f * x * (vR11 + vR22) + y * (vR10 + vR01) + z * (vR20 + vR02) + w * (vR21 - vR12));
        o_quat[2] = 2.f * (x * (vR10 + vR01) - 2.f * y * (vR00 + vR22) + z * (vR21 + vR12) + w * (vR02 - vR20));
        o_quat[3] = 2.f * (x * (vR20 + vR02) + y * (vR21 + vR12) - 2.f * z * (vR00 + vR11) + w * (vR10 - vR01));
    }
    warp_store_aos<3>(strip, o_mean, v_mean3d, wfirst * 3, n * 3, lane);
    warp_store_aos<3>(strip, o_scale, v_scale, wfirst * 3, n * 3, lane);
    warp_store_aos<4>(strip, o_quat, v_quat, wfirst * 4, n * 4, lane);
    if (v_cov2d) warp_store_aos<3>(strip, o_cov2d, v_cov2d, wfirst * 3, n * 3, lane);
    if (v_cov3d) warp_store_aos<6>(strip, o_cov3d, v_cov3d, wfirst * 6, n * 6, lane);
}

// ---------------------------------------------------------------------------------------------
// Batched cameras: the projection backward of B views of the same Gaussians in ONE
// launch.  A thread loads its Gaussian's parameters once, builds R, S and the 3D covariance once,
// then walks the views: per view the rasteriser's raw moments (raster_bwd.cu: MOMENTS) become
// v_xy / v_conic, the EWA VJP runs with that view's matrix, v_mean3d and the covariance cotangent
// v_V accumulate in registers; the scale / quaternion VJP -- linear in v_V -- runs once on the sum,
// and every gradient is written once.  Replaces, per view, a full projection backward (parameters
// and gradients through HBM) plus the add kernel that summed the views (120 bytes per Gaussian).
// mom_opacity holds the sum over the views of M0 and leaves as v_opacity.
__global__ void __launch_bounds__(PT)
project_bwd_views_kernel(const ProjectBwdParams p, const unsigned num_views,
                         const float *__restrict__ means3d, const float *__restrict__ scales,
                         const float4 *__restrict__ quats, const float *__restrict__ viewmats,
                         const int *__restrict__ radii, const float *__restrict__ conics,
                         const float2 *__restrict__ mom_xy, const float *__restrict__ mom_conic,
                         const size_t mom_stride /* floats between two views' moment arrays / m */,
                         float *__restrict__ v_mean3d, float *__restrict__ v_scale,
                         float *__restrict__ v_quat, const float *__restrict__ opacities,
                         float *__restrict__ mom_opacity) {
    __shared__ __align__(16) float s_strip[PT / 32][32 * 6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long idx = (long long)blockIdx.x * PT + tid;
    const long long wfirst = (long long)blockIdx.x * PT + warp * 32;
    const long long n = p.num_points;
    float *strip = s_strip[warp];
    float o_mean[3] = {0.f, 0.f, 0.f}, o_scale[3] = {0.f, 0.f, 0.f};
    float o_quat[4] = {0.f, 0.f, 0.f, 0.f};
    if (idx < n) {
        const float mx = __ldg(means3d + idx * 3), my = __ldg(means3d + idx * 3 + 1),
                    mz = __ldg(means3d + idx * 3 + 2);
        const float sc0 = __ldg(scales + idx * 3), sc1 = __ldg(scales + idx * 3 + 1),
                    sc2 = __ldg(scales + idx * 3 + 2);
        const float4 q = __ldg(quats + idx);
        const float w = q.x, x = q.y, y = q.z, z = q.w;
        const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - w * z), R02 = 2.f * (x * z + w * y);
        const float R10 = 2.f * (x * y + w * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - w * x);
        const float R20 = 2.f * (x * z - w * y), R21 = 2.f * (y * z + w * x), R22 = 1.f - 2.f * (x * x + y * y);
        const float gs = p.glob_scale;
        const float s0 = gs * sc0, s1 = gs * sc1, s2 = gs * sc2;
        const float e0 = s0 * s0, e1 = s1 * s1, e2 = s2 * s2;
        const float k0 = e0 * R00 * R00 + e1 * R01 * R01 + e2 * R02 * R02;
        const float k1 = e0 * R00 * R10 + e1 * R01 * R11 + e2 * R02 * R12;
        const float k2 = e0 * R00 * R20 + e1 * R01 * R21 + e2 * R02 * R22;
        const float k3 = e0 * R10 * R10 + e1 * R11 * R11 + e2 * R12 * R12;
        const float k4 = e0 * R10 * R20 + e1 * R11 * R21 + e2 * R12 * R22;
        const float k5 = e0 * R20 * R20 + e1 * R21 * R21 + e2 * R22 * R22;
        float vm_x = 0.f, vm_y = 0.f, vm_z = 0.f;
        float V00 = 0.f, V01 = 0.f, V02 = 0.f, V11 = 0.f, V12 = 0.f, V22 = 0.f;  // sum of v_V
        const float fx = p.fx, fy = p.fy;
        for (unsigned b = 0; b < num_views; ++b) {
            if (__ldg(radii + (size_t)b * n + idx) <= 0) continue;
            const float *cn = conics + ((size_t)b * n + idx) * 3;
            const float cx = __ldg(cn), cy = __ldg(cn + 1), cz = __ldg(cn + 2);
            const float2 mxy = __ldg(mom_xy + (size_t)b * (mom_stride / 2) + idx);
            const float *mc = mom_conic + (size_t)b * mom_stride + idx * 3;
            // raw moments -> v_xy, v_conic (v_conic.y halved for the symmetric product below)
            const float vxy_x = -(cx * mxy.x + cy * mxy.y), vxy_y = -(cy * mxy.x + cz * mxy.y);
            const float gx = -0.5f * __ldg(mc), gy = -0.5f * __ldg(mc + 1), gz = -0.5f * __ldg(mc + 2);
            const float *v = viewmats + 16 * (size_t)b;
            const float w00 = __ldg(v + 0), w01 = __ldg(v + 1), w02 = __ldg(v + 2), w03 = __ldg(v + 3);
            const float w10 = __ldg(v + 4), w11 = __ldg(v + 5), w12 = __ldg(v + 6), w13 = __ldg(v + 7);
            const float w20 = __ldg(v + 8), w21 = __ldg(v + 9), w22 = __ldg(v + 10), w23 = __ldg(v + 11);
            const float pvx = w00 * mx + w01 * my + w02 * mz + w03;
            const float pvy = w10 * mx + w11 * my + w12 * mz + w13;
            const float pvz = w20 * mx + w21 * my + w22 * mz + w23;
            const float rw = 1.f / (pvz + 1e-6f);
            const float vpx = fx * vxy_x, vpy = fy * vxy_y;
            const float vvx = vpx * rw, vvy = vpy * rw;
            const float vvz = -(vpx * pvx + vpy * pvy) * rw * rw;
            vm_x += w00 * vvx + w10 * vvy + w20 * vvz;
            vm_y += w01 * vvx + w11 * vvy + w21 * vvz;
            vm_z += w02 * vvx + w12 * vvy + w22 * vvz;
            // cov2d_to_conic_vjp
            const float p00 = cx * gx + cy * gy, p01 = cx * gy + cy * gz;
            const float p10 = cy * gx + cz * gy, p11 = cy * gy + cz * gz;
            const float vc_x = -(p00 * cx + p01 * cy);
            const float vc_y = -((p00 * cy + p01 * cz) + (p10 * cx + p11 * cy));
            const float vc_z = -(p10 * cy + p11 * cz);
            // project_cov3d_ewa_vjp (unclamped t)
            const float rz = 1.f / pvz, rz2 = rz * rz;
            const float j00 = fx * rz, j11 = fy * rz, j02 = -fx * pvx * rz2, j12 = -fy * pvy * rz2;
            const float T00 = j00 * w00 + j02 * w20, T01 = j00 * w01 + j02 * w21, T02 = j00 * w02 + j02 * w22;
            const float T10 = j11 * w10 + j12 * w20, T11 = j11 * w11 + j12 * w21, T12 = j11 * w12 + j12 * w22;
            const float g00 = vc_x, g01 = 0.5f * vc_y, g11 = vc_z;
            const float A00 = g00 * T00 + g01 * T10, A01 = g00 * T01 + g01 * T11, A02 = g00 * T02 + g01 * T12;
            const float A10 = g01 * T00 + g11 * T10, A11 = g01 * T01 + g11 * T11, A12 = g01 * T02 + g11 * T12;
            V00 += T00 * A00 + T10 * A10; V11 += T01 * A01 + T11 * A11; V22 += T02 * A02 + T12 * A12;
            V01 += T00 * A01 + T10 * A11; V02 += T00 * A02 + T10 * A12; V12 += T01 * A02 + T11 * A12;
            const float vT00 = 2.f * (A00 * k0 + A01 * k1 + A02 * k2);
            const float vT01 = 2.f * (A00 * k1 + A01 * k3 + A02 * k4);
            const float vT02 = 2.f * (A00 * k2 + A01 * k4 + A02 * k5);
            const float vT10 = 2.f * (A10 * k0 + A11 * k1 + A12 * k2);
            const float vT11 = 2.f * (A10 * k1 + A11 * k3 + A12 * k4);
            const float vT12 = 2.f * (A10 * k2 + A11 * k4 + A12 * k5);
            const float vJ00 = vT00 * w00 + vT01 * w01 + vT02 * w02;
            const float vJ02 = vT00 * w20 + vT01 * w21 + vT02 * w22;
            const float vJ11 = vT10 * w10 + vT11 * w11 + vT12 * w12;
            const float vJ12 = vT10 * w20 + vT11 * w21 + vT12 * w22;
            const float rz3 = rz2 * rz;
            const float vt_x = -fx * rz2 * vJ02;
            const float vt_y = -fy * rz2 * vJ12;
            const float vt_z = -fx * rz2 * vJ00 + 2.f * fx * pvx * rz3 * vJ02 - fy * rz2 * vJ11 +
                               2.f * fy * pvy * rz3 * vJ12;
            vm_x += vt_x * w00 + vt_y * w10 + vt_z * w20;
            vm_y += vt_x * w01 + vt_y * w11 + vt_z * w21;
            vm_z += vt_x * w02 + vt_y * w12 + vt_z * w22;
        }
        o_mean[0] = vm_x; o_mean[1] = vm_y; o_mean[2] = vm_z;
        // scale_rot_to_cov3d_vjp and quat_to_rotmat_vjp, once, on the summed v_V
        const float B00 = V00 * R00 + V01 * R10 + V02 * R20, B01 = V00 * R01 + V01 * R11 + V02 * R21,
                    B02 = V00 * R02 + V01 * R12 + V02 * R22;
        const float B10 = V01 * R00 + V11 * R10 + V12 * R20, B11 = V01 * R01 + V11 * R11 + V12 * R21,
                    B12 = V01 * R02 + V11 * R12 + V12 * R22;
        const float B20 = V02 * R00 + V12 * R10 + V22 * R20, B21 = V02 * R01 + V12 * R11 + V22 * R21,
                    B22 = V02 * R02 + V12 * R12 + V22 * R22;
        o_scale[0] = gs * 2.f * s0 * (R00 * B00 + R10 * B10 + R20 * B20);
        o_scale[1] = gs * 2.f * s1 * (R01 * B01 + R11 * B11 + R21 * B21);
        o_scale[2] = gs * 2.f * s2 * (R02 * B02 + R12 * B12 + R22 * B22);
        const float f0 = 2.f * e0, f1 = 2.f * e1, f2 = 2.f * e2;
        const float vR00 = f0 * B00, vR01 = f1 * B01, vR02 = f2 * B02;
        const float vR10 = f0 * B10, vR11 = f1 * B11, vR12 = f2 * B12;
        const float vR20 = f0 * B20, vR21 = f1 * B21, vR22 = f2 * B22;
        o_quat[0] = 2.f * (x * (vR21 - vR12) + y * (vR02 - vR20) + z * (vR10 - vR01));
        o_quat[1] = 2.f * (-2.f * x * (vR11 + vR22) + y * (vR10 + vR01) + z * (vR20 + vR02) + w * (vR21 - vR12));
        o_quat[2] = 2.f * (x * (vR10 + vR01) - 2.f * y * (vR00 + vR22) + z * (vR21 + vR12) + w * (vR02 - vR20));
        o_quat[3] = 2.f * (x * (vR20 + vR02) + y * (vR21 + vR12) - 2.f * z * (vR00 + vR11) + w * (vR10 - vR01));
        const float m0 = mom_opacity[idx];
        mom_opacity[idx] = m0 != 0.f ? m0 / __ldg(opacities + idx) : 0.f;
    }
    warp_store_aos<3>(strip, o_mean, v_mean3d, wfirst * 3, n * 3, lane);
    warp_store_aos<3>(strip, o_scale, v_scale, wfirst * 3, n * 3, lane);
    warp_store_aos<4>(strip, o_quat, v_quat, wfirst * 4, n * 4, lane);
}

// batched cameras: adds one view's staged v_mean3d (3n) / v_scale (3n) / v_quat (4n) to the sums
__global__ void __launch_bounds__(256)
accumulate_param_grads_kernel(const float *__restrict__ s_mean, const float *__restrict__ s_scale,
                              const float *__restrict__ s_quat, float *__restrict__ v_mean3d,
                              float *__restrict__ v_scale, float *__restrict__ v_quat, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n3 = 3 * n, n6 = 6 * n, n10 = 10 * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n10; i += stride) {
        if (i < n3) v_mean3d[i] += __ldg(s_mean + i);
        else if (i < n6) v_scale[i - n3] += __ldg(s_scale + (i - n3));
        else v_quat[i - n6] += __ldg(s_quat + (i - n6));
    }
}

}  // namespace

int launch_accumulate_param_grads(cudaStream_t stream, size_t n, const float *s_mean,
                                  const float *s_scale, const float *s_quat, float *v_mean3d,
                                  float *v_scale, float *v_quat) {
    if (n == 0) return 0;
    const size_t want = (10 * n + 255) / 256;
    const unsigned blocks = (unsigned)(want < 148u * 16u ? want : 148u * 16u);
    accumulate_param_grads_kernel<<<blocks, 256, 0, stream>>>(s_mean, s_scale, s_quat, v_mean3d,
                                                              v_scale, v_quat, n);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_project_fwd(cudaStream_t stream, const ProjectFwdParams &p, const float *means3d,
                       const float *scales, const float *quats, const float *viewmat,
                       float *cov3ds, float *xys, float *depths, int *radii, float *conics,
                       float *compensation, int *num_tiles_hit, int *cum_tiles_hit,
                       unsigned *block_totals, const FusedPrep *fused) {
    if (p.num_points == 0) return 0;
    const unsigned blocks = ceil_div(p.num_points, PT);
    if (fused)
        launch_chained(project_fwd_kernel<true>, blocks, PT, 0, stream,
            p, means3d, scales, reinterpret_cast<const float4 *>(quats), viewmat, cov3ds,
            reinterpret_cast<float2 *>(xys), depths, radii, conics, compensation, num_tiles_hit,
            cum_tiles_hit, block_totals, *fused);
    else
        launch_chained(project_fwd_kernel<false>, blocks, PT, 0, stream,
            p, means3d, scales, reinterpret_cast<const float4 *>(quats), viewmat, cov3ds,
            reinterpret_cast<float2 *>(xys), depths, radii, conics, compensation, num_tiles_hit,
            cum_tiles_hit, block_totals, FusedPrep{});
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    if (!cum_tiles_hit) return 0;
    // cum_tiles_hit = CTA-local inclusive counts + exclusive scan of the CTA totals
    if (launch_scan_totals(stream, block_totals, blocks)) return 1;
    return launch_add_block_prefix(stream, cum_tiles_hit, block_totals, p.num_points, PT);
}

int launch_project_bwd(cudaStream_t stream, const ProjectBwdParams &p, const float *means3d,
                       const float *scales, const float *quats, const float *viewmat,
                       const float *cov3ds, const int *radii, const float *conics,
                       const float *compensation, const float *v_compensation, const float *v_xy,
                       const float *v_depth, const float *v_conic, float *v_mean3d,
                       float *v_scale, float *v_quat, float *v_cov2d, float *v_cov3d,
                       const float *opacities, float *moment_opacity) {
    if (p.num_points == 0) return 0;
    const unsigned blocks = ceil_div(p.num_points, PT);
    const bool recompute = cov3ds == nullptr;
    if (v_compensation && !compensation) {
        set_error("project_bwd: a compensation cotangent needs the forward's compensation");
        return 1;
    }
    if (!conics || !radii) {
        set_error("project_bwd: conics / radii missing");
        return 1;
    }
    const bool finish = opacities && moment_opacity;
    const float4 *q4 = reinterpret_cast<const float4 *>(quats);
    const float2 *vxy2 = reinterpret_cast<const float2 *>(v_xy);
#define JSB_PBWD(F, R)                                                                              \
    launch_chained(project_bwd_kernel<F, R>, blocks, PT, 0, stream,                                 \
        p, means3d, scales, q4, viewmat, cov3ds, radii, conics, compensation, v_compensation,   \
        vxy2, v_depth, v_conic, v_mean3d, v_scale, v_quat, v_cov2d, v_cov3d,                        \
        finish ? opacities : nullptr, finish ? moment_opacity : nullptr)
    if (finish && recompute) JSB_PBWD(true, true);
    else if (finish) JSB_PBWD(true, false);
    else if (recompute) JSB_PBWD(false, true);
    else JSB_PBWD(false, false);
#undef JSB_PBWD
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_project_bwd_views(cudaStream_t stream, const ProjectBwdParams &p, unsigned num_views,
                            const float *means3d, const float *scales, const float *quats,
                            const float *viewmats, const int *radii, const float *conics,
                            const float *mom_xy, const float *mom_conic, size_t mom_stride,
                            float *v_mean3d, float *v_scale, float *v_quat, const float *opacities,
                            float *mom_opacity) {
    if (p.num_points == 0) return 0;
    const unsigned blocks = ceil_div(p.num_points, PT);
    project_bwd_views_kernel<<<blocks, PT, 0, stream>>>(
        p, num_views, means3d, scales, reinterpret_cast<const float4 *>(quats), viewmats, radii, conics,
        reinterpret_cast<const float2 *>(mom_xy), mom_conic, mom_stride, v_mean3d, v_scale, v_quat,
        opacities, mom_opacity);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace jsb
