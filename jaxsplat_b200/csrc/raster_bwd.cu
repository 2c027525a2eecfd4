// Tile rasteriser, backward, for sm_100a.
//
// Behavioural spec: reference lib/kernels/backward.cu:153-343 (rasterize_bwd): per tile, replay
// the sorted list back to front starting from each pixel's final_idx / final_Ts, un-composite T,
// and accumulate d(loss)/d{colour, opacity, conic, xy, |xy|} per Gaussian.  Alpha is clamped at
// 0.99 here (0.999 in the forward) and the clamp is treated as identity.
//
// Design (not a port):
//   * the default kernel (raster_bwd_hits_kernel) is driven by the forward's hit words: one
//     independent warp per 8x8 quadrant replays only the list entries the forward marked as
//     accepted, from a dense staged array, with no CTA barrier anywhere;
//   * the batch kernel (raster_bwd16x2p_kernel) is what runs when no hit words exist (the binning
//     was rebuilt by the backward itself): same CTA/quadrant layout, staging and conservative
//     culling as raster_fwd.cu, starting at the batch holding the largest final_idx of the tile;
//   * both: two pixels per lane in packed FP32x2 registers; the per-pixel partial gradients are
//     reduced across the warp through a per-warp shared transpose (conflict-free STS, then lanes
//     summing rows with LDS.128 + FADD2) instead of full butterflies, and the totals end up in
//     different lanes which issue ONE RED.ADD.F32 instruction.  The hit-driven kernel reduces the
//     rows of THREE visits together (one lane per row); the batch kernel reduces every visit on
//     its own (18 lanes, half a row each), deferred into the next visit;
//   * invalid lanes are not branched around: their alpha and visibility are forced to 0, which
//     makes every partial exactly 0 and leaves T and the colour buffer unchanged.
#include "common.cuh"
#include "kernels.cuh"
#include "raster_cull.cuh"
#include "raster_hits.cuh"

namespace jsb {
namespace {

constexpr int RED_STRIDE = 36;  // floats per transposed row: 16-byte phases of LDS.128 conflict-free

// per-pixel gradient of one valid (pixel, Gaussian) pair; updates T and the colour buffer
__device__ __forceinline__ void pair_grad(bool valid, float alpha, float vis, float opac, float dx, float dy,
                                          float cx, float cy, float cz, float r, float g, float b,
                                          const float3 v_out, float v_out_alpha, float T_final,
                                          const float3 bg, float &T, float3 &buf, float *o) {
    // invalid lanes: alpha = vis = 0 and ra = 1 make every partial exactly 0 and keep T / buf
    alpha = valid ? alpha : 0.f;
    vis = valid ? vis : 0.f;
    float ra;  // gradient path: MUFU.RCP (1 ulp) is enough; 1-alpha is in [0.01, 1]
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(1.f - alpha));
    ra = valid ? ra : 1.f;
    T = T * ra;
    const float fac = alpha * T;
    o[0] = v_out.x * fac;
    o[1] = v_out.y * fac;
    o[2] = v_out.z * fac;
    float va = v_out.x * (T * r - buf.x * ra);
    va = fmaf(v_out.y, T * g - buf.y * ra, va);
    va = fmaf(v_out.z, T * b - buf.z * ra, va);
    va = fmaf(v_out_alpha, T_final * ra, va);
    const float nb = ra * (-T_final);
    va = fmaf(v_out.x, nb * bg.x, va);
    va = fmaf(v_out.y, nb * bg.y, va);
    va = fmaf(v_out.z, nb * bg.z, va);
    buf.x = fmaf(fac, r, buf.x);
    buf.y = fmaf(fac, g, buf.y);
    buf.z = fmaf(fac, b, buf.z);
    const float v_sigma = (vis * (-opac)) * va;
    const float hs = v_sigma * 0.5f;
    o[3] = dx * (dx * hs);
    o[4] = dy * (dx * v_sigma);
    o[5] = dy * (dy * hs);
    o[6] = fmaf(dy, cy, dx * cx) * v_sigma;
    o[7] = fmaf(dy, cz, dx * cy) * v_sigma;
    o[8] = fabsf(o[6]);
    o[9] = fabsf(o[7]);
    o[10] = vis * va;
}

// ---- two pixels per thread ---------------------------------------------------------------------
// A CTA is 4 warps and each warp owns an 8x8 quadrant of the tile: lane l replays pixels (x, y) and
// (x, y + 4).  The two pixels' partial gradients are added in registers before the cross-lane
// reduction, so the transpose / half-row sums / RED (a third of the instructions of a visit) and
// the candidate bookkeeping are paid once per 64 pixels.  ABS = false (fused render: nothing
// consumes v_xy_abs) drops two of the eleven rows.  Row order of the reduction: rgb(3) conic(3)
// xy(2) opacity(1) xy_abs(2).
constexpr int RB2 = 128;
// tuning knobs of the packed kernel (A/B builds: make EXTRA=-DJSB_BWD_CTAS=5 ...)
#ifndef JSB_BWD_CTAS
#define JSB_BWD_CTAS 6   // resident CTAs per SM the register allocation is bounded for
#endif
#ifndef JSB_BWD_EPT
#define JSB_BWD_EPT 2    // staged entries per thread and batch (batch = 128 * EPT entries)
#endif

// ---- two pixels per thread, packed FP32x2 math -------------------------------------------------
// sm_100 adds FFMA2 / FMUL2 / FADD2: one issue slot, two FP32 lanes of a 64-bit register pair
// (measured on B200, tools/microbench/fp32x2_rate.cu: the packed forms issue at half the scalar
// rate, i.e. the same lane throughput for half the issue slots).  The replay loop is bound by issue
// slots, most of them FP32, and a thread's two pixels (x, y) / (x, y+4) run the same arithmetic on
// the same Gaussian -- so every per-pixel quantity lives in a {pixel 0, pixel 1} pair, and the
// per-Gaussian operands enter as scalar broadcasts (the SASS operand form `R.F32` of the packed
// instructions: no move is needed to duplicate them).  sigma / alpha keep the reference's
// operation order lane by lane (FMUL2 / FFMA2 round like FMUL / FFMA), so the accept decisions
// are those of the scalar kernels; the gradient arithmetic is regrouped (tolerance-tested).
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }

template <bool CULL, bool ABS>
__global__ void __launch_bounds__(RB2, JSB_BWD_CTAS)
raster_bwd16x2p_kernel(const RasterParams p, const unsigned *__restrict__ ids_sorted,
                       const int2 *__restrict__ tile_bins, const SplatRecord *__restrict__ records,
                       const float *__restrict__ background, const float *__restrict__ final_Ts,
                       const int *__restrict__ final_idx, const float *__restrict__ v_out_img,
                       const float *__restrict__ v_out_alpha_img, float *__restrict__ v_xy,
                       float *__restrict__ v_xy_abs, float *__restrict__ v_conic,
                       float *__restrict__ v_colors, float *__restrict__ v_opacity) {
    constexpr int NV = ABS ? 11 : 9;
    constexpr int EPT = JSB_BWD_EPT;      // staged entries per thread and batch
    constexpr int BATCH = RB2 * EPT;
    __shared__ float4 sA2[2][BATCH];  // double-buffered batches (cp.async staging)
    __shared__ float4 sB2[2][BATCH];
    __shared__ float sC2[2][BATCH];
    __shared__ unsigned sId2[2][BATCH];
    __shared__ unsigned char sMask[BATCH];
    __shared__ int s_max[RB2 / 32];
    __shared__ __align__(16) float sRed[(RB2 / 32) * NV * RED_STRIDE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *red = sRed + warp * NV * RED_STRIDE;
    const unsigned tile = p.band_y0 * p.tiles_x + blockIdx.x;
    const unsigned tile_y = tile / p.tiles_x, tile_x = tile - tile_y * p.tiles_x;
    const unsigned j = tile_x * 16 + (warp & 1) * 8 + (lane & 7);
    const unsigned i0 = tile_y * 16 + (warp >> 1) * 8 + (lane >> 3), i1 = i0 + 4;
    const float px = (float)j + 0.5f, py0 = (float)i0 + 0.5f, py1 = (float)i1 + 0.5f;
    const bool in0 = (i0 < p.img_h) && (j < p.img_w), in1 = (i1 < p.img_h) && (j < p.img_w);
    const int2 range = tile_bins[tile];
    const int qbit = (warp & 1) + 4 * (warp >> 1);

    float Tf0 = 1.f, Tf1 = 1.f, voa0 = 0.f, voa1 = 0.f;
    float3 vo0 = make_float3(0.f, 0.f, 0.f), vo1 = vo0;
    int bf0 = -1, bf1 = -1;
    if (in0) {
        const size_t pix = (size_t)i0 * p.img_w + j;
        Tf0 = final_Ts[pix];
        bf0 = final_idx[pix];
        vo0 = make_float3(v_out_img[pix * 3], v_out_img[pix * 3 + 1], v_out_img[pix * 3 + 2]);
        voa0 = v_out_alpha_img ? v_out_alpha_img[pix] : 0.f;
    }
    if (in1) {
        const size_t pix = (size_t)i1 * p.img_w + j;
        Tf1 = final_Ts[pix];
        bf1 = final_idx[pix];
        vo1 = make_float3(v_out_img[pix * 3], v_out_img[pix * 3 + 1], v_out_img[pix * 3 + 2]);
        voa1 = v_out_alpha_img ? v_out_alpha_img[pix] : 0.f;
    }
    const float3 bg = make_float3(__ldg(background), __ldg(background + 1), __ldg(background + 2));
    // per-pixel constants as {pixel 0, pixel 1} pairs
    const float2 tfc2 = f2(Tf0 * (voa0 - (vo0.x * bg.x + vo0.y * bg.y + vo0.z * bg.z)),
                           Tf1 * (voa1 - (vo1.x * bg.x + vo1.y * bg.y + vo1.z * bg.z)));
    const float2 vox2 = f2(vo0.x, vo1.x), voy2 = f2(vo0.y, vo1.y), voz2 = f2(vo0.z, vo1.z);
    const float2 nvox2 = f2(-vo0.x, -vo1.x), nvoy2 = f2(-vo0.y, -vo1.y), nvoz2 = f2(-vo0.z, -vo1.z);
    const float2 npy2 = f2(-py0, -py1);
    float2 T2 = f2(Tf0, Tf1);
    float2 bufx2 = f2(0.f, 0.f), bufy2 = bufx2, bufz2 = bufx2;

    int wmax = max(bf0, bf1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_max[warp] = wmax;
    __syncthreads();
    int tmax = s_max[0];
#pragma unroll
    for (int w = 1; w < RB2 / 32; ++w) tmax = max(tmax, s_max[w]);
    const int top = min(range.y - 1, tmax);
    if (top < range.x) return;  // nothing in this tile was composited (uniform across the CTA)
    const int num_batches = (top - range.x + BATCH) / BATCH;
    const float tile_x0 = (float)(tile_x * 16), tile_y0 = (float)(tile_y * 16);

    const int slot = lane >> 1;
    float *out_base;
    int out_stride;
    if (slot < 3) { out_base = v_colors + slot; out_stride = 3; }
    else if (slot < 6) { out_base = v_conic + (slot - 3); out_stride = 3; }
    else if (slot < 8) { out_base = v_xy + (slot - 6); out_stride = 2; }
    else if (slot < 9) { out_base = v_opacity; out_stride = 1; }
    else { out_base = v_xy_abs + (slot - 9); out_stride = 2; }
    const bool writer = ((lane & 1) == 0) && slot < NV;

    auto load_id = [&](int bb, int h) -> unsigned {
        const int idx = top - BATCH * bb - (tid + h * RB2);
        return (bb < num_batches && idx >= range.x) ? ids_sorted[idx] : 0u;
    };
    auto issue = [&](int bb, int h, unsigned g) {
        const int t = tid + h * RB2;
        if (top - BATCH * bb - t >= range.x) {
            const SplatRecord *r = records + g;
            cp_async16(&sA2[bb & 1][t], &r->q0);
            cp_async16(&sB2[bb & 1][t], &r->q1);
            cp_async4(&sC2[bb & 1][t], &r->q2);
            sId2[bb & 1][t] = g;
        }
    };
    unsigned gn[EPT];
#pragma unroll
    for (int h = 0; h < EPT; ++h) issue(0, h, load_id(0, h));
#pragma unroll
    for (int h = 0; h < EPT; ++h) gn[h] = load_id(1, h);

    // The cross-lane reduction of a visit is deferred into the NEXT visit: its shared-memory reads,
    // adds, shuffle and RED have nothing to do with that visit's sigma / alpha arithmetic, so in
    // one straight-line block the scheduler overlaps the two latency chains (the reduction tail
    // held 27 % of the kernel's stall samples when it ran at the end of its own visit).
    bool pend = false;       // `red` holds the partials of a visit that is not reduced yet
    unsigned pend_id = 0;    // its Gaussian
    const float4 *red_row = reinterpret_cast<const float4 *>(
        red + min(lane >> 1, NV - 1) * RED_STRIDE + (lane & 1) * 16);
    auto reduce_rows = [&]() -> float {
        const float4 p0 = red_row[0], p1 = red_row[1], p2 = red_row[2], p3 = red_row[3];
        float2 u0 = __fadd2_rn(f2(p0.x, p0.y), f2(p0.z, p0.w));
        float2 u1 = __fadd2_rn(f2(p1.x, p1.y), f2(p1.z, p1.w));
        float2 u2 = __fadd2_rn(f2(p2.x, p2.y), f2(p2.z, p2.w));
        float2 u3 = __fadd2_rn(f2(p3.x, p3.y), f2(p3.z, p3.w));
        u0 = __fadd2_rn(u0, u1);
        u2 = __fadd2_rn(u2, u3);
        u0 = __fadd2_rn(u0, u2);
        return u0.x + u0.y;
    };

    for (int b = 0; b < num_batches; ++b) {
        cp_async_wait_all();
        __syncthreads();  // batch b landed; every warp is done with batch b-1
        const int batch_end = top - BATCH * b;  // slot t holds sorted index batch_end - t
        const int buf = b & 1;
        const float4 *sA = sA2[buf];
        const float4 *sB = sB2[buf];
        const float *sC = sC2[buf];
        const unsigned *sId = sId2[buf];
        if (b + 1 < num_batches) {
#pragma unroll
            for (int h = 0; h < EPT; ++h) issue(b + 1, h, gn[h]);
        }
#pragma unroll
        for (int h = 0; h < EPT; ++h) gn[h] = load_id(b + 2, h);
#pragma unroll
        for (int h = 0; h < EPT; ++h) {
            const int t = tid + h * RB2;
            unsigned mask = 0;
            if (batch_end - t >= range.x) {
                const float4 q0 = sA[t];
                const float2 q1 = *reinterpret_cast<const float2 *>(&sB[t]);
                mask = CULL ? cull_mask16(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tile_x0, tile_y0) : 0xFFu;
            }
            sMask[t] = (unsigned char)(mask | (mask >> 2));
        }
        __syncthreads();

#pragma unroll 1
        for (int k = 0; k < BATCH / 32; ++k) {
            // entries newer than the warp's largest final index cannot be valid for any lane
            if (batch_end - (k * 32 + 31) > wmax) continue;
            const unsigned m = sMask[k * 32 + lane];
            unsigned bits = __ballot_sync(0xffffffffu, (m >> qbit) & 1u);
            while (bits) {
                const int t = k * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                const float4 a = sA[t];    // {x, y, opacity, cx}
                const float4 bq = sB[t];   // {cy, cz, r, g}
                float total = reduce_rows();  // previous visit's partials (stale rows if !pend)
                // conic_sigma (common.cuh) lane by lane, in the reference's operation order
                const float dx = fsub(a.x, px);
                const float2 dy2 = __fadd2_rn(bc(a.y), npy2);
                const float ca = fmul(a.w, dx);                        // cx dx
                const float ce = fmul(bq.x, dx);                       // cy dx
                const float2 b2 = __fmul2_rn(bc(bq.y), dy2);           // cz dy
                const float2 c2 = __fmul2_rn(dy2, b2);
                float2 s2 = __ffma2_rn(bc(dx), bc(ca), c2);
                s2 = __fmul2_rn(s2, bc(0.5f));
                const float2 sg2 = __ffma2_rn(bc(ce), dy2, s2);
                const float2 x2 = __fmul2_rn(sg2, bc(kNegLog2e));
                const float vis0 = ex2_approx(x2.x), vis1 = ex2_approx(x2.y);
                const float2 ar2 = __fmul2_rn(bc(a.z), f2(vis0, vis1));  // opacity * vis
                const float al0 = fminf(ar2.x, 0.99f), al1 = fminf(ar2.y, 0.99f);
                const int sidx = batch_end - t;
                const bool ok0 = in0 && (sidx <= bf0) && !(sg2.x < 0.f) && !(al0 < kAlphaMin);
                const bool ok1 = in1 && (sidx <= bf1) && !(sg2.y < 0.f) && !(al1 < kAlphaMin);
                total += __shfl_xor_sync(0xffffffffu, total, 1);
                if (writer && pend) atomicAdd(out_base + pend_id * (unsigned)out_stride, total);
                pend = false;
                if (!__any_sync(0xffffffffu, ok0 || ok1)) continue;

                // invalid pixels: alpha = opacity*vis = 0 and ra = 1 make every partial exactly 0
                // and leave T and the colour buffer unchanged
                const float2 al2 = f2(ok0 ? al0 : 0.f, ok1 ? al1 : 0.f);
                const float2 av2 = f2(ok0 ? ar2.x : 0.f, ok1 ? ar2.y : 0.f);   // opacity * vis
                const float2 vis2 = f2(ok0 ? vis0 : 0.f, ok1 ? vis1 : 0.f);
                const float2 om2 = __ffma2_rn(al2, bc(-1.f), bc(1.f));         // 1 - alpha
                float r0, r1;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(om2.x));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(om2.y));
                const float2 ra2 = f2(ok0 ? r0 : 1.f, ok1 ? r1 : 1.f);
                const float cb = sC[t];
                T2 = __fmul2_rn(T2, ra2);
                const float2 fac2 = __fmul2_rn(al2, T2);
                // v_alpha = sum_c v_out_c (T c_c - buf_c ra) + tfc ra
                //         = T <v_out, c> + ra (tfc - <v_out, buf>)
                float2 P2 = __fmul2_rn(vox2, bc(bq.z));
                P2 = __ffma2_rn(voy2, bc(bq.w), P2);
                P2 = __ffma2_rn(voz2, bc(cb), P2);
                float2 Q2 = __ffma2_rn(nvox2, bufx2, tfc2);
                Q2 = __ffma2_rn(nvoy2, bufy2, Q2);
                Q2 = __ffma2_rn(nvoz2, bufz2, Q2);
                const float2 va2 = __ffma2_rn(ra2, Q2, __fmul2_rn(T2, P2));
                bufx2 = __ffma2_rn(fac2, bc(bq.z), bufx2);
                bufy2 = __ffma2_rn(fac2, bc(bq.w), bufy2);
                bufz2 = __ffma2_rn(fac2, bc(cb), bufz2);
                // v_sigma = -opacity vis v_alpha
                const float2 pv2 = __fmul2_rn(av2, va2);
                const float2 vs2 = __fmul2_rn(pv2, bc(-1.f));
                const float2 hs2 = __fmul2_rn(pv2, bc(-0.5f));
                const float2 dxv2 = __fmul2_rn(bc(dx), vs2);    // dx v_sigma
                float2 o2[9];
                o2[0] = __fmul2_rn(vox2, fac2);
                o2[1] = __fmul2_rn(voy2, fac2);
                o2[2] = __fmul2_rn(voz2, fac2);
                o2[3] = __fmul2_rn(bc(dx * dx), hs2);
                o2[4] = __fmul2_rn(dy2, dxv2);
                o2[5] = __fmul2_rn(dy2, __fmul2_rn(dy2, hs2));
                o2[6] = __ffma2_rn(__fmul2_rn(dy2, bc(bq.x)), vs2, __fmul2_rn(bc(ca), vs2));  // (cx dx + cy dy) v_sigma
                o2[7] = __fmul2_rn(__fadd2_rn(bc(ce), b2), vs2);                               // (cy dx + cz dy) v_sigma
                o2[8] = __fmul2_rn(vis2, va2);
                float v[NV];
#pragma unroll
                for (int r = 0; r < 9; ++r) v[r] = o2[r].x + o2[r].y;
                if (ABS) {
                    v[9] = fabsf(o2[6].x) + fabsf(o2[6].y);
                    v[10] = fabsf(o2[7].x) + fabsf(o2[7].y);
                }
                // transpose through shared memory: row r = value r of the 32 lanes
                __syncwarp();
#pragma unroll
                for (int r = 0; r < NV; ++r) red[r * RED_STRIDE + lane] = v[r];
                __syncwarp();
                pend = true;
                pend_id = sId[t];
            }
        }
    }
    if (pend) {  // warp-uniform: the last visit of this warp
        float total = reduce_rows();
        total += __shfl_xor_sync(0xffffffffu, total, 1);
        if (writer) atomicAdd(out_base + pend_id * (unsigned)out_stride, total);
    }
}

// ---- hit-driven backward: one independent warp per 8x8 quadrant ---------------------------------
// The forward rasteriser (raster_fwd16x2_kernel) records, per quadrant warp and per group of 32
// list entries, a bitmask of the entries for which at least one of the warp's 64 pixels passed the
// accept test (hit_words layout: raster_hits.cuh).  Every (pixel, Gaussian) pair the backward has
// to differentiate lies in such an entry, so this kernel
//   * never stages or cull-tests an entry its warp does not visit (the batch kernels above stage
//     and mask every entry of the tile from the top down: ~500 per tile at C3, of which a warp
//     visits ~40),
//   * needs no CTA: a CTA IS one warp = one quadrant (4 per tile), there is no barrier anywhere,
//     and a quadrant with few visits frees its slot instead of waiting for the tile's slowest,
//   * walks a DENSE array of visits (no find-first-set walk over candidate masks).
// Per span of 32 hit words (1024 list entries) the lanes expand their word into a list of visit
// positions in shared memory (newest first); visits are then staged 32 at a time -- lane j gathers
// the sorted id and the 48-byte record of visit j (cp.async, double buffered, ids one sub-chunk
// further ahead) -- and replayed with exactly the arithmetic of raster_bwd16x2p_kernel, including
// the cross-lane reduction (batched: see JSB_BWDH_BATCH).
#ifndef JSB_BWDH_SPAN
#define JSB_BWDH_SPAN 32
#endif
constexpr int HB_SPAN = JSB_BWDH_SPAN;  // hit words expanded at a time (<= 32)
// Visits whose partial rows are reduced together.  Three visits x nine rows = 27 rows: ONE lane
// per row sums its 32 partials (8 LDS.128) and all 27 atomics leave in one instruction, behind two
// __syncwarp per three visits instead of two per visit; the visits of a batch are unrolled, so the
// batch slot is a compile-time constant.  More than 32 rows take another round of the same lanes
// (measured slower: 2 -> 207, 3 -> 196, 5 -> 209, 7 -> 211 us at C3).
#ifndef JSB_BWDH_BATCH
#define JSB_BWDH_BATCH 3
#endif
#ifndef JSB_BWDH_CTAS
#define JSB_BWDH_CTAS 24     // resident one-warp CTAs per SM the register allocation is bounded for
                             // (80 registers; with the batched, unrolled reduction 24 beats 20: 190 vs 196 us)
#endif

struct HbStage {      // one sub-chunk of <= 32 visits
    float4 A[32];     // x, y, opac, conic.x
    float4 B[32];     // conic.y, conic.z, r, g
    float4 C[32];     // b, Gaussian id (bits), sorted index (bits), -
};

// MOMENTS (fused render only, no v_xy_abs): the nine reduced rows are rgb(3) and the raw moments
//   M0 = sum a v, Mx = sum dx a v, My = sum dy a v, Mxx = sum dx^2 a v, Mxy = sum dx dy a v,
//   Myy = sum dy^2 a v          (a = opacity * vis, v = v_alpha, dx / dy = mean - pixel)
// accumulated into v_opacity / v_xy / v_conic in that order; the projection backward, which has
// the conic and the opacity of the Gaussian in registers anyway, turns them into the gradients
//   v_opacity = M0 / opacity, v_xy = -(cx Mx + cy My, cy Mx + cz My), v_conic = -(Mxx/2, Mxy, Myy/2)
// (project.cu: FINISH_MOMENTS) -- five packed multiplies per visit instead of eleven.
// SUPER: super-column lists (RasterParams::sc_shift); with all_entries the entries whose tile
// columns do not cover the tile are skipped like the forward skips them.
// T8: block_width 8 -- the tile is the quadrant (one CTA per tile).
// ALL: no hit words -- every entry of the quadrant's range is replayed (the backward rebuilt the
// lists itself, or the test hook asked for it).
template <bool ABS, bool MOMENTS, bool SUPER, bool T8 = false, bool ALL = false>
__global__ void __launch_bounds__(32, JSB_BWDH_CTAS)
raster_bwd_hits_kernel(const RasterParams p, const void *__restrict__ list,
                       const int2 *__restrict__ tile_bins, const SplatRecord *__restrict__ records,
                       const unsigned *__restrict__ hit_words,
                       const float *__restrict__ background, const float *__restrict__ final_Ts,
                       const int *__restrict__ final_idx, const float *__restrict__ v_out_img,
                       const float *__restrict__ v_out_alpha_img, float *__restrict__ v_xy,
                       float *__restrict__ v_xy_abs, float *__restrict__ v_conic,
                       float *__restrict__ v_colors, float *__restrict__ v_opacity) {
    constexpr int NV = ABS ? 11 : 9;
    constexpr int VB = ABS ? 2 : JSB_BWDH_BATCH;  // visits per reduction batch
    static_assert(VB >= 1 && VB <= 8, "batch ids live in eight shared words");
    __shared__ unsigned s_bid[8];  // the Gaussians of the batch's visits
    __shared__ HbStage st[2];
    __shared__ unsigned short s_list[HB_SPAN * 32];
    __shared__ __align__(16) float red[VB * NV * RED_STRIDE];

    const int lane = threadIdx.x;
    constexpr unsigned TS = T8 ? 8 : 16;  // tile size in pixels
    const unsigned quad = T8 ? 0u : blockIdx.x & 3u, tile_local = T8 ? blockIdx.x : blockIdx.x >> 2;
    const unsigned tile = p.band_y0 * p.tiles_x + tile_local;
    const unsigned tile_y = tile / p.tiles_x, tile_x = tile - tile_y * p.tiles_x;
    const unsigned j = tile_x * TS + (quad & 1) * 8 + (lane & 7);
    const unsigned i0 = tile_y * TS + (quad >> 1) * 8 + (lane >> 3), i1 = i0 + 4;
    const float px = (float)j + 0.5f, py0 = (float)i0 + 0.5f, py1 = (float)i1 + 0.5f;
    const bool in0 = (i0 < p.img_h) && (j < p.img_w), in1 = (i1 < p.img_h) && (j < p.img_w);
    const unsigned *ids_sorted = reinterpret_cast<const unsigned *>(list);
    const uint2 *ents = reinterpret_cast<const uint2 *>(list);
    const unsigned super_local = (tile_y - p.band_y0) * p.cols2 + (tile_x >> p.sc_shift);
    chain_wait();
    const int2 range = tile_bins[SUPER ? super_local : tile];

    float Tf0 = 1.f, Tf1 = 1.f, voa0 = 0.f, voa1 = 0.f;
    float3 vo0 = make_float3(0.f, 0.f, 0.f), vo1 = vo0;
    int bf0 = -1, bf1 = -1;
    if (in0) {
        const size_t pix = (size_t)i0 * p.img_w + j;
        Tf0 = final_Ts[pix];
        bf0 = final_idx[pix];
        vo0 = make_float3(v_out_img[pix * 3], v_out_img[pix * 3 + 1], v_out_img[pix * 3 + 2]);
        voa0 = v_out_alpha_img ? v_out_alpha_img[pix] : 0.f;
    }
    if (in1) {
        const size_t pix = (size_t)i1 * p.img_w + j;
        Tf1 = final_Ts[pix];
        bf1 = final_idx[pix];
        vo1 = make_float3(v_out_img[pix * 3], v_out_img[pix * 3 + 1], v_out_img[pix * 3 + 2]);
        voa1 = v_out_alpha_img ? v_out_alpha_img[pix] : 0.f;
    }
    const int wmax = __reduce_max_sync(0xffffffffu, max(bf0, bf1));
    const int top = min(range.y - 1, wmax);
    if (top < range.x) return;  // nothing in this quadrant was composited (warp-uniform)

    const float3 bg = make_float3(__ldg(background), __ldg(background + 1), __ldg(background + 2));
    const float2 tfc2 = f2(Tf0 * (voa0 - (vo0.x * bg.x + vo0.y * bg.y + vo0.z * bg.z)),
                           Tf1 * (voa1 - (vo1.x * bg.x + vo1.y * bg.y + vo1.z * bg.z)));
    const float2 vox2 = f2(vo0.x, vo1.x), voy2 = f2(vo0.y, vo1.y), voz2 = f2(vo0.z, vo1.z);
    const float2 npy2 = f2(-py0, -py1);
    float2 T2 = f2(Tf0, Tf1);
    // <v_out, colour composited behind the current entry>: the only way the colour buffer of
    // the reference (backward.cu:283-291) enters the gradient, so it is carried as ONE scalar per
    // pixel instead of three channels
    float2 S2 = f2(0.f, 0.f);

    // where row `s` (0 .. NV-1) of a visit's reduced gradient accumulates: base pointer, stride
    auto row_target = [&](int s, float *&base, int &stride) {
        if (s < 3) { base = v_colors + s; stride = 3; }
        else if (MOMENTS) {
            if (s < 4) { base = v_opacity; stride = 1; }
            else if (s < 6) { base = v_xy + (s - 4); stride = 2; }
            else { base = v_conic + min(s - 6, 2); stride = 3; }
        }
        else if (s < 6) { base = v_conic + (s - 3); stride = 3; }
        else if (s < 8) { base = v_xy + (s - 6); stride = 2; }
        else if (s < 9) { base = v_opacity; stride = 1; }
        else { base = v_xy_abs + (s - 9); stride = 2; }
    };
    // batched reduction: in round k lane L owns row 32 k + L of the batch = row (32 k + L) % NV of
    // visit (32 k + L) / NV
    constexpr int ROUNDS = (VB * NV + 31) / 32;
    float *b_base[ROUNDS];
    int b_stride[ROUNDS];
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) row_target((32 * k + lane) % NV, b_base[k], b_stride[k]);
    auto flush_batch = [&](const int nb) {  // warp-uniform; rows [0, nb * NV) hold partials
        __syncwarp();
#pragma unroll
        for (int k = 0; k < ROUNDS; ++k) {
            const int rowi = 32 * k + lane;
            if (rowi < nb * NV) {
                const float4 *row = reinterpret_cast<const float4 *>(red + rowi * RED_STRIDE);
                float2 u[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 q = row[i];
                    u[i] = __fadd2_rn(f2(q.x, q.y), f2(q.z, q.w));
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) u[i] = __fadd2_rn(u[i], u[i + 4]);
                u[0] = __fadd2_rn(u[0], u[2]);
                u[1] = __fadd2_rn(u[1], u[3]);
                u[0] = __fadd2_rn(u[0], u[1]);
                const unsigned id = s_bid[rowi / NV];
                atomicAdd(b_base[k] + id * (unsigned)b_stride[k], u[0].x + u[0].y);
            }
        }
        __syncwarp();
    };

    const unsigned *hw = hit_words + (SUPER
        ? hit_quadrant_offset_super(range, super_local, tile_x & ((1u << p.sc_shift) - 1u), quad, p.sc_shift)
        : T8 ? hit_tile8_offset(range, tile_local) : hit_quadrant_offset(range, tile_local, quad));
    const int rel_top = top - range.x;
    for (int gs = rel_top >> 5; gs >= 0; gs -= HB_SPAN) {
        // ---- expand this span's hit words into visit positions, newest first -------------------
        const int g = gs - lane;  // lane's group (32 list entries) of this span
        unsigned w = (g >= 0 && lane < HB_SPAN) ? (ALL ? 0xFFFFFFFFu : __ldg(hw + g)) : 0u;
        if (g == (rel_top >> 5)) w &= (2u << (rel_top & 31)) - 1u;  // entries above `top`
        const int cnt = __popc(w);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tt = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += tt;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;  // warp-uniform
        __syncwarp();              // every lane is done reading the previous span's list
        {
            int pos = incl - cnt;
            while (w) {
                const int b = 31 - __clz(w);
                w &= ~(1u << b);
                s_list[pos++] = (unsigned short)((lane << 5) | b);  // (group offset in span, bit)
            }
        }
        __syncwarp();
        // list entry e -> sorted index: group gs - (e >> 5), bit e & 31
        auto sorted_index = [&](int o) -> int {
            const unsigned e = s_list[o];
            return range.x + ((gs - (int)(e >> 5)) << 5) + (int)(e & 31u);
        };
        auto issue = [&](int s, int sidx, unsigned gid) {  // lane's visit of sub-chunk s
            if (s * 32 + lane < total) {
                HbStage &S = st[s & 1];
                const SplatRecord *r = records + gid;
                cp_async16(&S.A[lane], &r->q0);
                cp_async16(&S.B[lane], &r->q1);
                cp_async4(&S.C[lane].x, &r->q2);
                S.C[lane].y = __uint_as_float(gid);
                S.C[lane].z = __int_as_float(sidx);
            }
        };
        const int nsub = (total + 31) >> 5;
        // visit o of the span: its sorted index and Gaussian id (a super-column entry that does not
        // cover this tile -- only met when every entry is replayed -- gets an index no pixel reaches)
        int sidx_n = 0;
        unsigned gid_n = 0u;
        auto fetch = [&](int o) {
            sidx_n = 0;
            gid_n = 0u;
            if (o < total) {
                sidx_n = sorted_index(o);
                if (SUPER) {
                    const uint2 e = __ldg(ents + sidx_n);
                    gid_n = e.x;
                    if (ALL && !((tile_x - (e.y & 0xffffu)) < (e.y >> 16))) sidx_n = 0x7fffffff;
                } else {
                    gid_n = ids_sorted[sidx_n];
                }
            }
        };
        fetch(lane);
        issue(0, sidx_n, gid_n);
        fetch(32 + lane);

        for (int s = 0; s < nsub; ++s) {
            cp_async_wait_all();
            __syncwarp();  // sub-chunk s landed; every lane is done with sub-chunk s - 1
            if (s + 1 < nsub) issue(s + 1, sidx_n, gid_n);
            fetch((s + 2) * 32 + lane);
            const HbStage &S = st[s & 1];
            const int cs = min(32, total - s * 32);
            // visit t of the sub-chunk; its partial rows go to slot k of the reduction batch (k is
            // a compile-time constant in the unrolled callers: row addresses are immediates)
            auto visit = [&](const int t, const int k) {
                const float4 a = S.A[t];    // {x, y, opacity, cx}
                const float4 bq = S.B[t];   // {cy, cz, r, g}
                const float4 cq = S.C[t];   // {b, id, sorted index, -}
                // conic_sigma (common.cuh) lane by lane, in the reference's operation order
                const float dx = fsub(a.x, px);
                const float2 dy2 = __fadd2_rn(bc(a.y), npy2);
                const float ca = fmul(a.w, dx);                        // cx dx
                const float ce = fmul(bq.x, dx);                       // cy dx
                const float2 b2 = __fmul2_rn(bc(bq.y), dy2);           // cz dy
                const float2 c2 = __fmul2_rn(dy2, b2);
                float2 s2 = __ffma2_rn(bc(dx), bc(ca), c2);
                s2 = __fmul2_rn(s2, bc(0.5f));
                const float2 sg2 = __ffma2_rn(bc(ce), dy2, s2);
                const float2 x2 = __fmul2_rn(sg2, bc(kNegLog2e));
                const float vis0 = ex2_approx(x2.x), vis1 = ex2_approx(x2.y);
                const float2 ar2 = __fmul2_rn(bc(a.z), f2(vis0, vis1));  // opacity * vis
                const float al0 = fminf(ar2.x, 0.99f), al1 = fminf(ar2.y, 0.99f);
                const int sidx = __float_as_int(cq.z);
                const bool ok0 = in0 && (sidx <= bf0) && !(sg2.x < 0.f) && !(al0 < kAlphaMin);
                const bool ok1 = in1 && (sidx <= bf1) && !(sg2.y < 0.f) && !(al1 < kAlphaMin);
                if (lane == 0) s_bid[k] = __float_as_uint(cq.y);
                // (a hit-driven visit has an accepting pixel by construction; only the replay of
                // EVERY entry meets visits without one, worth a vote to skip -- leaving zero rows)
                if (ALL && !__any_sync(0xffffffffu, ok0 || ok1)) {
#pragma unroll
                    for (int r = 0; r < NV; ++r) red[(k * NV + r) * RED_STRIDE + lane] = 0.f;
                    return;
                }

                // invalid pixels: vis = 0 makes alpha = 0, ra = rcp(1) = 1 and every partial
                // exactly 0, and leaves T and the colour buffer unchanged
                const float2 vis2 = f2(ok0 ? vis0 : 0.f, ok1 ? vis1 : 0.f);
                // (opacity * vis and its clamp were evaluated for the accept test: selected, not redone)
                const float2 av2 = f2(ok0 ? ar2.x : 0.f, ok1 ? ar2.y : 0.f);   // opacity * vis
                const float2 al2 = f2(ok0 ? al0 : 0.f, ok1 ? al1 : 0.f);
                const float2 om2 = __ffma2_rn(al2, bc(-1.f), bc(1.f));         // 1 - alpha
                float r0, r1;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(om2.x));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(om2.y));
                const float2 ra2 = f2(r0, r1);
                const float cb = cq.x;
                T2 = __fmul2_rn(T2, ra2);
                const float2 fac2 = __fmul2_rn(al2, T2);
                // v_alpha = sum_c v_out_c (T c_c - buf_c ra) + tfc ra
                //         = T <v_out, c> + ra (tfc - <v_out, buf>)
                float2 P2 = __fmul2_rn(vox2, bc(bq.z));
                P2 = __ffma2_rn(voy2, bc(bq.w), P2);
                P2 = __ffma2_rn(voz2, bc(cb), P2);
                const float2 Q2 = __fadd2_rn(tfc2, f2(-S2.x, -S2.y));
                const float2 va2 = __ffma2_rn(ra2, Q2, __fmul2_rn(T2, P2));
                S2 = __ffma2_rn(fac2, P2, S2);
                const float2 pv2 = __fmul2_rn(av2, va2);        // -v_sigma = opacity vis v_alpha
                float v[NV];
                {
                    const float2 c0 = __fmul2_rn(vox2, fac2), c1 = __fmul2_rn(voy2, fac2),
                                 c2 = __fmul2_rn(voz2, fac2);
                    v[0] = c0.x + c0.y;
                    v[1] = c1.x + c1.y;
                    v[2] = c2.x + c2.y;
                }
                if (MOMENTS) {
                    // the lane's two pixels share dx: the moments in dx come from the pair SUM
                    // (scalar multiplies) instead of a packed multiply and a pair add each
                    const float m0 = pv2.x + pv2.y;                // M0  -> v_opacity
                    const float mx = dx * m0;                      // Mx  -> v_xy.x
                    const float2 my2 = __fmul2_rn(dy2, pv2);
                    const float my = my2.x + my2.y;                // My  -> v_xy.y
                    const float2 myy2 = __fmul2_rn(dy2, my2);
                    v[3] = m0;
                    v[4] = mx;
                    v[5] = my;
                    v[6] = dx * mx;                                // Mxx -> v_conic.x
                    v[7] = dx * my;                                // Mxy -> v_conic.y
                    v[8] = myy2.x + myy2.y;                        // Myy -> v_conic.z
                } else {
                    const float2 vs2 = __fmul2_rn(pv2, bc(-1.f));
                    const float2 hs2 = __fmul2_rn(pv2, bc(-0.5f));
                    const float2 dxv2 = __fmul2_rn(bc(dx), vs2);    // dx v_sigma
                    const float2 o3 = __fmul2_rn(bc(dx * dx), hs2);
                    const float2 o4 = __fmul2_rn(dy2, dxv2);
                    const float2 o5 = __fmul2_rn(dy2, __fmul2_rn(dy2, hs2));
                    const float2 o6 = __ffma2_rn(__fmul2_rn(dy2, bc(bq.x)), vs2, __fmul2_rn(bc(ca), vs2));  // (cx dx + cy dy) v_sigma
                    const float2 o7 = __fmul2_rn(__fadd2_rn(bc(ce), b2), vs2);                               // (cy dx + cz dy) v_sigma
                    const float2 o8 = __fmul2_rn(vis2, va2);
                    v[3] = o3.x + o3.y; v[4] = o4.x + o4.y; v[5] = o5.x + o5.y;
                    v[6] = o6.x + o6.y; v[7] = o7.x + o7.y; v[8] = o8.x + o8.y;
                    if (ABS) {
                        v[9] = fabsf(o6.x) + fabsf(o6.y);
                        v[10] = fabsf(o7.x) + fabsf(o7.y);
                    }
                }
                // transpose through shared memory: row r = value r of the 32 lanes
#pragma unroll
                for (int r = 0; r < NV; ++r) red[(k * NV + r) * RED_STRIDE + lane] = v[r];
            };
            int t0 = 0;
#pragma unroll 1
            for (; t0 + VB <= cs; t0 += VB) {  // full batches: VB visits, one reduction
#pragma unroll
                for (int k = 0; k < VB; ++k) visit(t0 + k, k);
                flush_batch(VB);
            }
            if (t0 < cs) {  // the sub-chunk's last, partial batch
#pragma unroll
                for (int k = 0; k < VB - 1; ++k)
                    if (t0 + k < cs) visit(t0 + k, k);
                flush_batch(cs - t0);
            }
        }
    }
}

// any block_width: one thread per pixel, per-thread atomics (semantics only, not tuned)
__global__ void raster_bwd_generic_kernel(
    const RasterParams p, const unsigned *__restrict__ ids_sorted,
    const int2 *__restrict__ tile_bins, const SplatRecord *__restrict__ records,
    const float *__restrict__ background, const float *__restrict__ final_Ts,
    const int *__restrict__ final_idx, const float *__restrict__ v_out_img,
    const float *__restrict__ v_out_alpha_img, float *__restrict__ v_xy,
    float *__restrict__ v_xy_abs, float *__restrict__ v_conic, float *__restrict__ v_colors,
    float *__restrict__ v_opacity) {
    const unsigned tile = (blockIdx.y + p.band_y0) * p.tiles_x + blockIdx.x;
    const unsigned i = (blockIdx.y + p.band_y0) * blockDim.y + threadIdx.y;
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (!(i < p.img_h && j < p.img_w)) return;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const int2 range = tile_bins[tile];
    const size_t pix = (size_t)i * p.img_w + j;
    const float T_final = final_Ts[pix];
    const int bin_final = final_idx[pix];
    const float3 v_out = make_float3(v_out_img[pix * 3], v_out_img[pix * 3 + 1], v_out_img[pix * 3 + 2]);
    const float v_oa = v_out_alpha_img[pix];
    const float3 bg = make_float3(__ldg(background), __ldg(background + 1), __ldg(background + 2));
    float T = T_final;
    float3 buf = make_float3(0.f, 0.f, 0.f);
    for (int idx = min(range.y - 1, bin_final); idx >= range.x; --idx) {
        const unsigned g = ids_sorted[idx];
        const float4 a = __ldg(&records[g].q0);
        const float4 bq = __ldg(&records[g].q1);
        const float cb = __ldg(&records[g].q2).x;
        const float dx = fsub(a.x, px), dy = fsub(a.y, py);
        const float sigma = conic_sigma(a.w, bq.x, bq.y, dx, dy);
        const float vis = ex2_approx(fmul(sigma, kNegLog2e));
        const float alpha = fminf(fmul(a.z, vis), 0.99f);
        if (sigma < 0.f || alpha < kAlphaMin) continue;
        float o[11];
        pair_grad(true, alpha, vis, a.z, dx, dy, a.w, bq.x, bq.y, bq.z, bq.w, cb, v_out, v_oa,
                  T_final, bg, T, buf, o);
        atomicAdd(v_colors + 3 * (size_t)g + 0, o[0]);
        atomicAdd(v_colors + 3 * (size_t)g + 1, o[1]);
        atomicAdd(v_colors + 3 * (size_t)g + 2, o[2]);
        atomicAdd(v_conic + 3 * (size_t)g + 0, o[3]);
        atomicAdd(v_conic + 3 * (size_t)g + 1, o[4]);
        atomicAdd(v_conic + 3 * (size_t)g + 2, o[5]);
        atomicAdd(v_xy + 2 * (size_t)g + 0, o[6]);
        atomicAdd(v_xy + 2 * (size_t)g + 1, o[7]);
        atomicAdd(v_xy_abs + 2 * (size_t)g + 0, o[8]);
        atomicAdd(v_xy_abs + 2 * (size_t)g + 1, o[9]);
        atomicAdd(v_opacity + g, o[10]);
    }
}

}  // namespace

extern bool g_disable_cull;
bool g_bwd_no_hits = false;    // option "bwdhits" = 0: the batch kernels even when hit words exist
bool g_bwd_no_moments = false;  // option "bwdmom" = 0: the fused backward emits finished gradients, not moments
bool g_bwd_all_entries = false;  // option "bwdhits" = 2 (test hook): the hit-driven kernel replays EVERY entry

template <bool CULL, bool ABS>
static void launch_x2(cudaStream_t stream, unsigned tiles, const RasterParams &p,
                      const unsigned *ids_sorted, const int2 *tile_bins, const SplatRecord *records,
                      const float *background, const float *final_Ts, const int *final_idx,
                      const float *v_out, const float *v_out_alpha, float *v_xy, float *v_xy_abs,
                      float *v_conic, float *v_colors, float *v_opacity) {
    raster_bwd16x2p_kernel<CULL, ABS><<<tiles, RB2, 0, stream>>>(
        p, ids_sorted, tile_bins, records, background, final_Ts, final_idx, v_out, v_out_alpha,
        v_xy, v_xy_abs, v_conic, v_colors, v_opacity);
}

// v_xy_abs == nullptr: not computed (fused render; block_width 16 only)
int launch_raster_bwd(cudaStream_t stream, const RasterParams &p, const void *list,
                      const int2 *tile_bins, const SplatRecord *records, const float *background,
                      const float *final_Ts, const int *final_idx, const float *v_out,
                      const float *v_out_alpha, float *v_xy, float *v_xy_abs, float *v_conic,
                      float *v_colors, float *v_opacity, const unsigned *hit_words,
                      bool *moments_out) {
    const unsigned tiles = p.tiles_x * (p.band_y1 - p.band_y0);
    const unsigned *ids_sorted = reinterpret_cast<const unsigned *>(list);
    if (moments_out) *moments_out = false;
    if (tiles == 0) return 0;
    if (g_bwd_no_moments) moments_out = nullptr;
    if (p.sc_shift && p.block_width != 16) {
        set_error("raster_bwd: super-column lists need block_width 16");
        return 1;
    }
#define JSB_HITS(ABS_, MOM_, SUP_, T8_, ALL_, GRID_, HW_)                                              \
    launch_chained(raster_bwd_hits_kernel<ABS_, MOM_, SUP_, T8_, ALL_>, GRID_, 32, 0, stream, p, list,  \
                   tile_bins, records, HW_, background, final_Ts, final_idx, v_out, v_out_alpha, v_xy, \
                   v_xy_abs, v_conic, v_colors, v_opacity)
#define JSB_HITS_MODE(SUP_, T8_, ALL_, GRID_, HW_)                                                     \
    do {                                                                                               \
        if (v_xy_abs) JSB_HITS(true, false, SUP_, T8_, ALL_, GRID_, HW_);                              \
        else if (moments_out) JSB_HITS(false, true, SUP_, T8_, ALL_, GRID_, HW_);                      \
        else JSB_HITS(false, false, SUP_, T8_, ALL_, GRID_, HW_);                                      \
    } while (0)
    if (p.sc_shift) {
        // super-column lists (fused render): always the hit-driven kernel; without hit words (the
        // backward rebuilt the lists itself) it replays every entry of the quadrant's range
        const bool all = g_bwd_all_entries || hit_words == nullptr;
        const unsigned *hw = hit_words ? hit_words : reinterpret_cast<const unsigned *>(list);  // (not read)
        if (all) JSB_HITS_MODE(true, false, true, tiles * 4, hw);
        else JSB_HITS_MODE(true, false, false, tiles * 4, hw);
        if (moments_out) *moments_out = !v_xy_abs;
    } else if (p.block_width == 8 && !(g_bwd_no_hits && v_xy_abs)) {
        // 8x8 tiles: the hit-driven kernel with the tile as its quadrant; without hit words (the
        // backward rebuilt the lists itself) it replays every entry of the tile's range
        const bool all = g_bwd_all_entries || hit_words == nullptr || g_bwd_no_hits;
        const unsigned *hw = hit_words ? hit_words : reinterpret_cast<const unsigned *>(list);  // (not read)
        if (all) JSB_HITS_MODE(false, true, true, tiles, hw);
        else JSB_HITS_MODE(false, true, false, tiles, hw);
        if (moments_out) *moments_out = !v_xy_abs;
    } else if (p.block_width == 16 && hit_words && !g_bwd_no_hits) {
        // one independent warp per quadrant, driven by the forward's hit words
        if (g_bwd_all_entries) JSB_HITS_MODE(false, false, true, tiles * 4, hit_words);
        else JSB_HITS_MODE(false, false, false, tiles * 4, hit_words);
        if (moments_out) *moments_out = !v_xy_abs;
#undef JSB_HITS_MODE
#undef JSB_HITS
    } else if (p.block_width == 16) {
        if (v_xy_abs) {
            if (g_disable_cull)
                launch_x2<false, true>(stream, tiles, p, ids_sorted, tile_bins, records, background, final_Ts, final_idx, v_out, v_out_alpha, v_xy, v_xy_abs, v_conic, v_colors, v_opacity);
            else
                launch_x2<true, true>(stream, tiles, p, ids_sorted, tile_bins, records, background, final_Ts, final_idx, v_out, v_out_alpha, v_xy, v_xy_abs, v_conic, v_colors, v_opacity);
        } else {
            if (g_disable_cull)
                launch_x2<false, false>(stream, tiles, p, ids_sorted, tile_bins, records, background, final_Ts, final_idx, v_out, v_out_alpha, v_xy, v_xy_abs, v_conic, v_colors, v_opacity);
            else
                launch_x2<true, false>(stream, tiles, p, ids_sorted, tile_bins, records, background, final_Ts, final_idx, v_out, v_out_alpha, v_xy, v_xy_abs, v_conic, v_colors, v_opacity);
        }
    } else {
        if (!v_xy_abs) {
            set_error("raster_bwd: block_width %u needs a v_xy_abs buffer", p.block_width);
            return 1;
        }
        dim3 grid(p.tiles_x, p.band_y1 - p.band_y0), block(p.block_width, p.block_width);
        raster_bwd_generic_kernel<<<grid, block, 0, stream>>>(
            p, ids_sorted, tile_bins, records, background, final_Ts, final_idx, v_out, v_out_alpha,
            v_xy, v_xy_abs, v_conic, v_colors, v_opacity);
    }
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace jsb
