// Projection forward (+ fused cumulative tile count) and projection backward for sm_100a.
//
// Behavioural spec: reference lib/kernels/forward.cu:13-131 (project_gaussians_fwd),
// :465-568 (project_cov3d_ewa, scale_rot_to_cov3d), helpers.h:9-57,122-153,194-211 and
// lib/ops.cu:607-643 (cumsum); backward: backward.cu:345-570, helpers.h:60-98,131-192.
// Design (not a port): one CTA owns a 256-Gaussian slab, moves the AoS operands with 16-byte
// coalesced transactions through shared memory, writes EVERY output element (no pre-zeroing
// memsets) and computes the CTA-local part of cum_tiles_hit in the same launch; two tiny kernels
// (scan of the CTA totals, add prefix) finish the scan.  A single-launch chained look-back scan
// was measured slower on B200: ~half of the kernel's stall samples sat on the look-back wait,
// because these CTAs live only a few microseconds.  No CUB, no temp allocation.
#include "common.cuh"
#include "kernels.cuh"

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

// FUSED (fused render, full-frame band): the kernel also does bin_prep's job (binning.cu) for the
// Gaussian it just projected -- packed rasteriser record, depth key, identity id, tile box, and
// the per-CTA totals of intersections and row entries -- so the projected arrays are not read
// back by a second kernel.
template <bool FUSED>
__global__ void __launch_bounds__(PT)
project_fwd_kernel(const ProjectFwdParams p, const float *__restrict__ means3d,
                   const float *__restrict__ scales, const float4 *__restrict__ quats,
                   const float *__restrict__ viewmat, float *__restrict__ cov3ds,
                   float2 *__restrict__ xys, float *__restrict__ depths, int *__restrict__ radii,
                   float *__restrict__ conics, float *__restrict__ compensation,
                   int *__restrict__ num_tiles_hit, int *__restrict__ cum_tiles_hit,
                   unsigned *__restrict__ block_totals, const FusedPrep f) {
    __shared__ __align__(16) float s_buf[PT * 9];
    __shared__ __align__(16) float s_color[FUSED ? PT * 3 : 4];
    __shared__ unsigned s_ws[PT / 32 + 1];
    __shared__ unsigned s_rows;
    const int tid = threadIdx.x;
    const unsigned bid = blockIdx.x;
    const long long base = (long long)bid * PT;
    const long long idx = base + tid;
    const long long n = p.num_points;
    float *s_mean = s_buf, *s_scale = s_buf + PT * 3;
    slab_load<3, PT>(s_mean, means3d, base * 3, n * 3);
    slab_load<3, PT>(s_scale, scales, base * 3, n * 3);
    if (FUSED) {
        slab_load<3, PT>(s_color, f.colors, base * 3, n * 3);
        if (tid == 0) s_rows = 0u;
    }
    __syncthreads();

    ProjOut o;
    o.tiles = 0;
    const bool live = idx < n;
    if (live) {
        const float4 q = __ldg(quats + idx);
        o = project_one(s_mean[tid * 3], s_mean[tid * 3 + 1], s_mean[tid * 3 + 2],
                        s_scale[tid * 3], s_scale[tid * 3 + 1], s_scale[tid * 3 + 2], q, viewmat,
                        p);
    }
    if (FUSED) {
        unsigned rows = 0;
        if (live) {
            SplatRecord rec;
            rec.q0 = make_float4(o.xy.x, o.xy.y, __ldg(f.opacities + idx), o.conic.x);
            rec.q1 = make_float4(o.conic.y, o.conic.z, s_color[tid * 3], s_color[tid * 3 + 1]);
            rec.q2 = make_float4(s_color[tid * 3 + 2], 0.f, 0.f, 0.f);
            f.records[idx] = rec;
            unsigned key = 0xFFFFFFFFu;
            uint2 box = make_uint2(0u, 0u);  // {x0 | y0 << 16, width | height << 16}
            if (o.radius > 0) {
                // same expression as bin_prep_kernel: box from the stored xy and integer radius
                unsigned x0, x1, y0, y1;
                tile_bbox(o.xy.x, o.xy.y, (float)o.radius, p.block_width_f, p.tiles_x, p.tiles_y,
                          x0, x1, y0, y1);
                if (y1 > y0 && x1 > x0) {
                    rows = y1 - y0;
                    box = make_uint2(x0 | (y0 << 16), (x1 - x0) | ((y1 - y0) << 16));
                }
                key = __float_as_uint(o.depth);
            }
            f.depth_keys[idx] = key;
            f.ids[idx] = (unsigned)idx;
            f.box_by_id[idx] = box;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, s);
        if ((tid & 31) == 0 && rows) atomicAdd(&s_rows, rows);
    }
    unsigned total;
    const unsigned excl = cta_exclusive_scan<PT>((unsigned)o.tiles, &total, s_ws);  // syncs
    if (tid == 0) {
        if (cum_tiles_hit) block_totals[bid] = total;
        if (FUSED && total) {  // one pair of same-address atomics per CTA
            atomicAdd(f.totals + 1, (unsigned long long)total);
            atomicAdd(f.totals + 4, (unsigned long long)s_rows);
        }
    }
    // stage AoS outputs (inputs are dead after the syncs inside the scan)
    float *s_cov = s_buf, *s_conic = s_buf + PT * 6;
    if (live) {
#pragma unroll
        for (int i = 0; i < 6; ++i) s_cov[tid * 6 + i] = o.cov3d[i];
        s_conic[tid * 3] = o.conic.x;
        s_conic[tid * 3 + 1] = o.conic.y;
        s_conic[tid * 3 + 2] = o.conic.z;
        xys[idx] = o.xy;
        depths[idx] = o.depth;
        radii[idx] = o.radius;
        compensation[idx] = o.comp;
        // (the fused render passes null: nothing downstream of it reads the tile counts)
        if (num_tiles_hit) num_tiles_hit[idx] = o.tiles;
        // CTA-local inclusive count; add_block_prefix_kernel adds the scanned CTA totals
        if (cum_tiles_hit) cum_tiles_hit[idx] = (int)(excl + (unsigned)o.tiles);
    }
    __syncthreads();
    slab_store<6, PT>(s_cov, cov3ds, base * 6, n * 6);
    slab_store<3, PT>(s_conic, conics, base * 3, n * 3);
}

// ---------------------------------------------------------------------------------------------
// Projection backward.  Float outputs only (tolerance-compared), so plain expressions; the
// structure follows the maths of backward.cu:345-570 but keeps every intermediate in registers
// (the reference read-modify-writes v_cov2d / v_mean3d / v_cov3d through global memory).
// ---------------------------------------------------------------------------------------------
struct M3 {  // column-major like the reference's mat3: m[c][r]
    float m[3][3];
};
__device__ __forceinline__ M3 mm(const M3 &a, const M3 &b) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
            r.m[c][rr] = a.m[0][rr] * b.m[c][0] + a.m[1][rr] * b.m[c][1] + a.m[2][rr] * b.m[c][2];
    return r;
}
__device__ __forceinline__ M3 tr(const M3 &a) {
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) r.m[c][rr] = a.m[rr][c];
    return r;
}

__global__ void __launch_bounds__(PT)
project_bwd_kernel(const ProjectBwdParams p, const float *__restrict__ means3d,
                   const float *__restrict__ scales, const float4 *__restrict__ quats,
                   const float *__restrict__ viewmat, const float *__restrict__ cov3ds,
                   const int *__restrict__ radii, const float *__restrict__ conics,
                   const float *__restrict__ compensation,
                   const float *__restrict__ v_compensation, const float2 *__restrict__ v_xy,
                   const float *__restrict__ v_depth, const float *__restrict__ v_conic,
                   float *__restrict__ v_mean3d, float *__restrict__ v_scale,
                   float4 *__restrict__ v_quat, float *__restrict__ v_cov2d,
                   float *__restrict__ v_cov3d) {
    // slabs: mean 3, scale 3, cov3d 6, conic 3, v_conic 3  -> 18 floats/Gaussian in,
    // outputs reuse: v_mean 3, v_scale 3, v_cov2d 3, v_cov3d 6
    __shared__ __align__(16) float s_buf[PT * 18];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * PT;
    const long long idx = base + tid;
    const long long n = p.num_points;
    float *s_mean = s_buf, *s_scale = s_buf + PT * 3, *s_cov = s_buf + PT * 6,
          *s_conic = s_buf + PT * 12, *s_vconic = s_buf + PT * 15;
    slab_load<3, PT>(s_mean, means3d, base * 3, n * 3);
    slab_load<3, PT>(s_scale, scales, base * 3, n * 3);
    slab_load<6, PT>(s_cov, cov3ds, base * 6, n * 6);
    slab_load<3, PT>(s_conic, conics, base * 3, n * 3);
    slab_load<3, PT>(s_vconic, v_conic, base * 3, n * 3);
    __syncthreads();

    float o_mean[3] = {0.f, 0.f, 0.f}, o_scale[3] = {0.f, 0.f, 0.f};
    float o_cov2d[3] = {0.f, 0.f, 0.f}, o_cov3d[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float4 o_quat = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool live = idx < n;
    if (live && radii[idx] > 0) {
        const float mx = s_mean[tid * 3], my = s_mean[tid * 3 + 1], mz = s_mean[tid * 3 + 2];
        const float *v = viewmat;
        const float fx = p.fx, fy = p.fy;
        // p_view
        const float pvx = v[0] * mx + v[1] * my + v[2] * mz + v[3];
        const float pvy = v[4] * mx + v[5] * my + v[6] * mz + v[7];
        const float pvz = v[8] * mx + v[9] * my + v[10] * mz + v[11];
        // project_pix_vjp (helpers.h:131-138) then R^T (helpers.h:90-98), + depth term
        const float2 vxy = v_xy[idx];
        const float rw = 1.f / (pvz + 1e-6f);
        const float vpx = fx * vxy.x, vpy = fy * vxy.y;
        const float vvx = vpx * rw, vvy = vpy * rw, vvz = -(vpx * pvx + vpy * pvy) * rw * rw;
        float vm_x = v[0] * vvx + v[4] * vvy + v[8] * vvz;
        float vm_y = v[1] * vvx + v[5] * vvy + v[9] * vvz;
        float vm_z = v[2] * vvx + v[6] * vvy + v[10] * vvz;
        const float vz = v_depth ? v_depth[idx] : 0.f;  // null = zero cotangent (fused render)
        vm_x += v[8] * vz;
        vm_y += v[9] * vz;
        vm_z += v[10] * vz;

        // cov2d_to_conic_vjp (helpers.h:60-71): v_Sigma = -X G X
        const float cx = s_conic[tid * 3], cy = s_conic[tid * 3 + 1], cz = s_conic[tid * 3 + 2];
        const float gx = s_vconic[tid * 3], gy = s_vconic[tid * 3 + 1] / 2.f,
                    gz = s_vconic[tid * 3 + 2];
        // column-major 2x2: X = [[cx,cy],[cy,cz]] ; nX = -X
        const float n00 = -cx, n01 = -cy, n10 = -cy, n11 = -cz;  // n[c][r]
        // P = nX * G
        const float p00 = n00 * gx + n10 * gy, p01 = n01 * gx + n11 * gy;
        const float p10 = n00 * gy + n10 * gz, p11 = n01 * gy + n11 * gz;
        // S = P * X
        const float S00 = p00 * cx + p10 * cy, S01 = p01 * cx + p11 * cy;
        const float S10 = p00 * cy + p10 * cz, S11 = p01 * cy + p11 * cz;
        float vc_x = S00, vc_y = S10 + S01, vc_z = S11;
        // cov2d_to_compensation_vjp (helpers.h:73-87): the reference evaluates this in FP64
        {
            const float comp = compensation[idx], vcomp = v_compensation ? v_compensation[idx] : 0.f;
            const float inv_det = cx * cz - cy * cy;
            const float om = 1 - comp * comp;
            const float vs = (float)((double)vcomp * 0.5 / ((double)comp + 1e-6));
            vc_x += (float)((double)vs * ((double)(om * cx) - 0.3 * (double)inv_det));
            vc_y += 2 * vs * (om * cy);
            vc_z += (float)((double)vs * ((double)(om * cz) - 0.3 * (double)inv_det));
        }
        o_cov2d[0] = vc_x; o_cov2d[1] = vc_y; o_cov2d[2] = vc_z;

        // project_cov3d_ewa_vjp (backward.cu:421-524); note: unclamped t
        M3 W = {{{v[0], v[4], v[8]}, {v[1], v[5], v[9]}, {v[2], v[6], v[10]}}};
        const float rz = 1.f / pvz, rz2 = rz * rz;
        M3 J = {{{fx * rz, 0.f, 0.f}, {0.f, fy * rz, 0.f}, {-fx * pvx * rz2, -fy * pvy * rz2, 0.f}}};
        const float *k = s_cov + tid * 6;
        M3 V = {{{k[0], k[1], k[2]}, {k[1], k[3], k[4]}, {k[2], k[4], k[5]}}};
        M3 G = {{{vc_x, 0.5f * vc_y, 0.f}, {0.5f * vc_y, vc_z, 0.f}, {0.f, 0.f, 0.f}}};
        M3 T = mm(J, W);
        M3 Tt = tr(T);
        M3 vV = mm(mm(Tt, G), T);
        M3 t1 = mm(mm(G, T), tr(V));
        M3 t2 = mm(mm(tr(G), T), V);
        M3 vT;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) vT.m[c][r] = t1.m[c][r] + t2.m[c][r];
        o_cov3d[0] = vV.m[0][0];
        o_cov3d[1] = vV.m[0][1] + vV.m[1][0];
        o_cov3d[2] = vV.m[0][2] + vV.m[2][0];
        o_cov3d[3] = vV.m[1][1];
        o_cov3d[4] = vV.m[1][2] + vV.m[2][1];
        o_cov3d[5] = vV.m[2][2];
        M3 vJ = mm(vT, tr(W));
        const float rz3 = rz2 * rz;
        const float vt_x = -fx * rz2 * vJ.m[2][0];
        const float vt_y = -fy * rz2 * vJ.m[2][1];
        const float vt_z = -fx * rz2 * vJ.m[0][0] + 2.f * fx * pvx * rz3 * vJ.m[2][0] -
                           fy * rz2 * vJ.m[1][1] + 2.f * fy * pvy * rz3 * vJ.m[2][1];
        vm_x += vt_x * W.m[0][0] + vt_y * W.m[0][1] + vt_z * W.m[0][2];
        vm_y += vt_x * W.m[1][0] + vt_y * W.m[1][1] + vt_z * W.m[1][2];
        vm_z += vt_x * W.m[2][0] + vt_y * W.m[2][1] + vt_z * W.m[2][2];
        o_mean[0] = vm_x; o_mean[1] = vm_y; o_mean[2] = vm_z;

        // scale_rot_to_cov3d_vjp (backward.cu:528-570)
        const float4 q = __ldg(quats + idx);
        const float w = q.x, x = q.y, y = q.z, z = q.w;
        M3 R = {{{1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y)},
                 {2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x)},
                 {2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y)}}};
        const float gs = p.glob_scale;
        const float s0 = gs * s_scale[tid * 3], s1 = gs * s_scale[tid * 3 + 1],
                    s2 = gs * s_scale[tid * 3 + 2];
        M3 S = {{{s0, 0.f, 0.f}, {0.f, s1, 0.f}, {0.f, 0.f, s2}}};
        M3 M = mm(R, S);
        M3 vVs = {{{o_cov3d[0], 0.5f * o_cov3d[1], 0.5f * o_cov3d[2]},
                   {0.5f * o_cov3d[1], o_cov3d[3], 0.5f * o_cov3d[4]},
                   {0.5f * o_cov3d[2], 0.5f * o_cov3d[4], o_cov3d[5]}}};
        M3 two = {{{2.f, 0.f, 0.f}, {0.f, 2.f, 0.f}, {0.f, 0.f, 2.f}}};
        M3 vM = mm(mm(vVs, M), two);
        o_scale[0] = (R.m[0][0] * vM.m[0][0] + R.m[0][1] * vM.m[0][1] + R.m[0][2] * vM.m[0][2]) * gs;
        o_scale[1] = (R.m[1][0] * vM.m[1][0] + R.m[1][1] * vM.m[1][1] + R.m[1][2] * vM.m[1][2]) * gs;
        o_scale[2] = (R.m[2][0] * vM.m[2][0] + R.m[2][1] * vM.m[2][1] + R.m[2][2] * vM.m[2][2]) * gs;
        M3 vR = mm(vM, S);
        // quat_to_rotmat_vjp (helpers.h:155-192)
        o_quat.x = 2.f * (x * (vR.m[1][2] - vR.m[2][1]) + y * (vR.m[2][0] - vR.m[0][2]) +
                          z * (vR.m[0][1] - vR.m[1][0]));
        o_quat.y = 2.f * (-2.f * x * (vR.m[1][1] + vR.m[2][2]) + y * (vR.m[0][1] + vR.m[1][0]) +
                          z * (vR.m[0][2] + vR.m[2][0]) + w * (vR.m[1][2] - vR.m[2][1]));
        o_quat.z = 2.f * (x * (vR.m[0][1] + vR.m[1][0]) - 2.f * y * (vR.m[0][0] + vR.m[2][2]) +
                          z * (vR.m[1][2] + vR.m[2][1]) + w * (vR.m[2][0] - vR.m[0][2]));
        o_quat.w = 2.f * (x * (vR.m[0][2] + vR.m[2][0]) + y * (vR.m[1][2] + vR.m[2][1]) -
                          2.f * z * (vR.m[0][0] + vR.m[1][1]) + w * (vR.m[0][1] - vR.m[1][0]));
    }
    __syncthreads();  // inputs dead; reuse for outputs
    float *s_vmean = s_buf, *s_vscale = s_buf + PT * 3, *s_vcov2d = s_buf + PT * 6,
          *s_vcov3d = s_buf + PT * 9;
    if (live) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            s_vmean[tid * 3 + i] = o_mean[i];
            s_vscale[tid * 3 + i] = o_scale[i];
            s_vcov2d[tid * 3 + i] = o_cov2d[i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) s_vcov3d[tid * 6 + i] = o_cov3d[i];
        v_quat[idx] = o_quat;
    }
    __syncthreads();
    slab_store<3, PT>(s_vmean, v_mean3d, base * 3, n * 3);
    slab_store<3, PT>(s_vscale, v_scale, base * 3, n * 3);
    slab_store<3, PT>(s_vcov2d, v_cov2d, base * 3, n * 3);
    slab_store<6, PT>(s_vcov3d, v_cov3d, base * 6, n * 6);
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
        project_fwd_kernel<true><<<blocks, PT, 0, stream>>>(
            p, means3d, scales, reinterpret_cast<const float4 *>(quats), viewmat, cov3ds,
            reinterpret_cast<float2 *>(xys), depths, radii, conics, compensation, num_tiles_hit,
            cum_tiles_hit, block_totals, *fused);
    else
        project_fwd_kernel<false><<<blocks, PT, 0, stream>>>(
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
                       float *v_scale, float *v_quat, float *v_cov2d, float *v_cov3d) {
    if (p.num_points == 0) return 0;
    const unsigned blocks = ceil_div(p.num_points, PT);
    project_bwd_kernel<<<blocks, PT, 0, stream>>>(
        p, means3d, scales, reinterpret_cast<const float4 *>(quats), viewmat, cov3ds, radii,
        conics, compensation, v_compensation, reinterpret_cast<const float2 *>(v_xy), v_depth,
        v_conic, v_mean3d, v_scale, reinterpret_cast<float4 *>(v_quat), v_cov2d, v_cov3d);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace jsb
