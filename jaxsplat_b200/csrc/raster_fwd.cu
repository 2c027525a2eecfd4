// Tile rasteriser, forward, for sm_100a.
//
// Behavioural spec: reference lib/kernels/forward.cu:339-462 (rasterize_fwd): per 16x16-pixel
// tile, walk the tile's depth-sorted Gaussians front to back; per pixel evaluate sigma, alpha,
// skip if sigma<0 or alpha<1/255, stop the pixel when next_T <= 1e-4 (that Gaussian is NOT
// composited), accumulate colour, record final T and the global sorted index of the last
// composited Gaussian.  Every decision is evaluated in the reference's compiled rounding order
// (common.cuh), so final_idx is bit-exact and the image is bit-identical.
//
// Design (not a port):
//   * one CTA = one tile = 4 warps; each warp owns an 8x8 quadrant and each lane two pixels,
//     (x, y) and (x, y + 4), held as packed FP32x2 register pairs (FMUL2 / FFMA2 / FADD2: one issue
//     slot for both pixels, each lane rounding like the scalar instruction);
//   * batches of 128 Gaussians are staged once per CTA from the packed 48-byte SplatRecord table;
//     the staging thread also computes a conservative mask saying which quadrants the Gaussian can
//     reach (raster_cull.cuh);
//   * each warp turns the masks into candidate bitmaps with ballots and only visits its own
//     candidates (find-first-set walk): pairs that cannot pass the alpha test are never evaluated,
//     which is where most of the reference's issue slots go;
//   * per candidate the warp does 2 broadcast LDS.128, and colours come from shared memory instead
//     of three dependent global loads;
//   * warps retire independently (ballot on `done`), the CTA retires on __syncthreads_count;
//   * the staging is software-pipelined: the records of batch b+1 arrive by cp.async (LDGSTS,
//     double-buffered shared memory) and the sorted ids of batch b+2 by a register prefetch while
//     batch b is composited;
//   * every warp records, per group of 32 list entries, which entries some pixel of its quadrant
//     accepted (raster_hits.cuh): the backward replays exactly those.
#include "common.cuh"
#include "kernels.cuh"
#include "raster_cull.cuh"
#include "raster_hits.cuh"

namespace jsb {
namespace {

constexpr int RB = 256;  // staging slots of the generic (any block_width) kernel

// ---- two pixels per thread -------------------------------------------------------------------
// A CTA is 4 warps and each warp owns an 8x8 quadrant of the tile: lane l composites pixels (x, y)
// and (x, y + 4) of the quadrant.  The candidate walk, the record loads, dx and the two conic
// products that depend only on dx are paid once per two pixels (about a third of the issue slots
// of a visit); every per-pixel operation keeps the reference's rounding order, so the outputs
// stay bit-identical.
constexpr int RF2 = 128;
// tuning knobs (A/B builds: make EXTRA=-DJSB_FWD_CTAS=6 ...)
#ifndef JSB_FWD_CTAS
#define JSB_FWD_CTAS 8   // resident CTAs per SM the register allocation is bounded for
#endif
#ifndef JSB_FWD_EPT
#define JSB_FWD_EPT 1    // staged entries per thread and batch (batch = 128 * EPT entries)
#endif

template <bool CULL>
__global__ void __launch_bounds__(RF2, JSB_FWD_CTAS)
raster_fwd16x2_kernel(const RasterParams p, const unsigned *__restrict__ ids_sorted,
                      const int2 *__restrict__ tile_bins, const SplatRecord *__restrict__ records,
                      const float *__restrict__ background, float *__restrict__ final_Ts,
                      int *__restrict__ final_idx, float *__restrict__ out_img,
                      unsigned *__restrict__ hit_words) {
    constexpr int EPT = JSB_FWD_EPT;      // staged entries per thread and batch
    constexpr int BATCH = RF2 * EPT;
    __shared__ float4 sA[2][BATCH];
    __shared__ float4 sB[2][BATCH];
    __shared__ float sC[2][BATCH];
    __shared__ unsigned char sMask[BATCH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tile = p.band_y0 * p.tiles_x + blockIdx.x;
    const unsigned tile_y = tile / p.tiles_x, tile_x = tile - tile_y * p.tiles_x;
    const unsigned j = tile_x * 16 + (warp & 1) * 8 + (lane & 7);
    const unsigned i0 = tile_y * 16 + (warp >> 1) * 8 + (lane >> 3), i1 = i0 + 4;
    const float px = (float)j + 0.5f, py0 = (float)i0 + 0.5f, py1 = (float)i1 + 0.5f;
    const bool in0 = (i0 < p.img_h) && (j < p.img_w), in1 = (i1 < p.img_h) && (j < p.img_w);
    bool done0 = !in0, done1 = !in1;
    const int qbit = (warp & 1) + 4 * (warp >> 1);  // quadrant bit of the row-folded cull mask

    const int2 range = tile_bins[tile];
    const int num_batches = (range.y - range.x + BATCH - 1) / BATCH;
    const float tile_x0 = (float)(tile_x * 16), tile_y0 = (float)(tile_y * 16);

    float T0 = 1.f, T1 = 1.f;
    int cur0 = 0, cur1 = 0;
    float r0 = 0.f, g0 = 0.f, b0 = 0.f, r1 = 0.f, g1 = 0.f, b1 = 0.f;
    // this quadrant's hit words (raster_hits.cuh): which entries the backward has to replay
    // (kept as a 32-bit word index: the kernel runs at its register cap)
    const unsigned hw_off = (unsigned)hit_quadrant_offset(range, blockIdx.x, (unsigned)warp);

    auto load_id = [&](int b, int h) -> unsigned {
        const int idx = range.x + BATCH * b + tid + h * RF2;
        return (b < num_batches && idx < range.y) ? ids_sorted[idx] : 0u;
    };
    auto issue = [&](int b, int h, unsigned g) {
        const int t = tid + h * RF2;
        if (range.x + BATCH * b + t < range.y) {
            const SplatRecord *r = records + g;
            cp_async16(&sA[b & 1][t], &r->q0);
            cp_async16(&sB[b & 1][t], &r->q1);
            cp_async4(&sC[b & 1][t], &r->q2);
        }
    };
    unsigned gn[EPT];
    if (num_batches > 0) {
#pragma unroll
        for (int h = 0; h < EPT; ++h) issue(0, h, load_id(0, h));
    }
#pragma unroll
    for (int h = 0; h < EPT; ++h) gn[h] = load_id(1, h);

    for (int b = 0; b < num_batches; ++b) {
        cp_async_wait_all();
        if (__syncthreads_count(done0 && done1) >= RF2) break;
        const int batch_start = range.x + BATCH * b;
        const int buf = b & 1;
        if (b + 1 < num_batches) {
#pragma unroll
            for (int h = 0; h < EPT; ++h) issue(b + 1, h, gn[h]);
        }
#pragma unroll
        for (int h = 0; h < EPT; ++h) gn[h] = load_id(b + 2, h);
#pragma unroll
        for (int h = 0; h < EPT; ++h) {
            const int t = tid + h * RF2;
            unsigned mask = 0;
            if (batch_start + t < range.y) {
                const float4 q0 = sA[buf][t];
                const float2 q1 = *reinterpret_cast<const float2 *>(&sB[buf][t]);
                mask = CULL ? cull_mask16(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, tile_x0, tile_y0) : 0xFFu;
            }
            sMask[t] = (unsigned char)(mask | (mask >> 2));
        }
        __syncthreads();

        bool warp_done = __all_sync(0xffffffffu, done0 && done1);
#pragma unroll
        for (int k = 0; k < BATCH / 32; ++k) {
            if (warp_done) continue;
            const unsigned m = sMask[k * 32 + lane];
            unsigned bits = __ballot_sync(0xffffffffu, (m >> qbit) & 1u);
            unsigned hit = 0;  // entries of this group some pixel of this lane accepted
            while (bits) {
                const int t = k * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                const float4 a = sA[buf][t];
                const float4 bq = sB[buf][t];
                const float dx = fsub(a.x, px);
                // conic_sigma (common.cuh) for both pixels at once with FMUL2 / FFMA2 / FADD2 (one
                // issue slot, two FP32 lanes, each rounding like the scalar form), the per-Gaussian
                // operands as scalar broadcasts: same operation order per pixel as the reference's
                // compiled code, so the outputs stay bit-identical
                const float2 dy2 = __fadd2_rn(make_float2(a.y, a.y), make_float2(-py0, -py1));
                const float ca = fmul(a.w, dx);                                   // cx dx
                const float ce = fmul(bq.x, dx);                                  // cy dx
                const float2 b2 = __fmul2_rn(make_float2(bq.y, bq.y), dy2);       // cz dy
                const float2 c2 = __fmul2_rn(dy2, b2);
                float2 s2 = __ffma2_rn(make_float2(dx, dx), make_float2(ca, ca), c2);
                s2 = __fmul2_rn(s2, make_float2(0.5f, 0.5f));
                const float2 sg2 = __ffma2_rn(make_float2(ce, ce), dy2, s2);
                const float2 x2 = __fmul2_rn(sg2, make_float2(kNegLog2e, kNegLog2e));
                const float2 ar2 = __fmul2_rn(make_float2(a.z, a.z),
                                              make_float2(ex2_approx(x2.x), ex2_approx(x2.y)));
                const float sg0 = sg2.x, sg1 = sg2.y;
                const float al0 = fminf(ar2.x, 0.999f), al1 = fminf(ar2.y, 0.999f);
                const bool ok0 = !done0 && !(sg0 < 0.f) && !(al0 < kAlphaMin);
                const bool ok1 = !done1 && !(sg1 < 0.f) && !(al1 < kAlphaMin);
                if (ok0 || ok1) {
                    hit |= __funnelshift_l(0u, 1u, (unsigned)t);  // 1 << (t & 31), one SHF
                    const float cb = sC[buf][t];
                    if (ok0) {
                        const float next_T = fmul(T0, fsub(1.f, al0));
                        if (next_T <= kTmin) {
                            done0 = true;  // this Gaussian is not composited
                        } else {
                            const float vis = fmul(T0, al0);
                            r0 = ffma(vis, bq.z, r0);
                            g0 = ffma(vis, bq.w, g0);
                            b0 = ffma(vis, cb, b0);
                            T0 = next_T;
                            cur0 = batch_start + t;
                        }
                    }
                    if (ok1) {
                        const float next_T = fmul(T1, fsub(1.f, al1));
                        if (next_T <= kTmin) {
                            done1 = true;
                        } else {
                            const float vis = fmul(T1, al1);
                            r1 = ffma(vis, bq.z, r1);
                            g1 = ffma(vis, bq.w, g1);
                            b1 = ffma(vis, cb, b1);
                            T1 = next_T;
                            cur1 = batch_start + t;
                        }
                    }
                }
            }
            // one word per (quadrant, group of 32 entries) the warp entered -- only for groups
            // that exist: the batch's trailing groups beyond the end of the list belong to the
            // next quadrant's (or tile's) slice
            if (hit_words && batch_start + 32 * k < range.y) {
                hit = __reduce_or_sync(0xffffffffu, hit);
                if (lane == 0) hit_words[hw_off + b * (BATCH / 32) + k] = hit;
            }
            warp_done = __all_sync(0xffffffffu, done0 && done1);
        }
    }
    cp_async_wait_all();

    const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
    if (in0) {
        const size_t pix = (size_t)i0 * p.img_w + j;
        final_Ts[pix] = T0;
        final_idx[pix] = cur0;
        out_img[pix * 3 + 0] = ffma(T0, bg0, r0);
        out_img[pix * 3 + 1] = ffma(T0, bg1, g0);
        out_img[pix * 3 + 2] = ffma(T0, bg2, b0);
    }
    if (in1) {
        const size_t pix = (size_t)i1 * p.img_w + j;
        final_Ts[pix] = T1;
        final_idx[pix] = cur1;
        out_img[pix * 3 + 0] = ffma(T1, bg0, r1);
        out_img[pix * 3 + 1] = ffma(T1, bg1, g1);
        out_img[pix * 3 + 2] = ffma(T1, bg2, b1);
    }
}

// ---- barrier-free forward: four independent quadrant warps per tile -----------------------------
// The counterpart of the hit-driven backward.  raster_fwd16x2_kernel above stages every batch once
// per CTA behind two __syncthreads and lets one thread compute the cull masks of all eight 8x4
// rectangles of its entry; its warps then wait for each other at every batch.  Here a warp walks
// the tile's list on its own, 32 entries at a time:
//   * lane l gathers entry 32 g + l (sorted id, then the 48-byte record; the next group's records
//     and the ids of the group after are in flight while this group is composited -- the four
//     warps of a tile share an SM, so three of the four gathers hit L1),
//   * tests it against the warp's OWN 8x8 quadrant only (one conservative rectangle test),
//   * the candidates are compacted into a dense array in the warp's shared memory (ballot prefix)
//     and replayed in order: no find-first-set walk, no CTA barrier, and a quadrant whose pixels
//     are all saturated simply stops.
// Per-visit arithmetic is the text of raster_fwd16x2_kernel: same rounding order, bit-identical.
// Shared-memory accesses of the warp kernel by 32-bit shared address: with generic pointers the
// compiler, at its register cap, re-derives the shared window base (S2R SR_CgaCtaId, LEA, ...) for
// every group -- ~20 issue slots of a ~120-slot group; an opaque 32-bit address has to stay in a
// register.
__device__ __forceinline__ void sts128(unsigned a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void sts64(unsigned a, const float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y));
}
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}

struct FwSlot {          // one compacted candidate
    float4 A;            // x, y, opac, conic.x
    float4 B;            // conic.y, conic.z, r, g
};

// SUPER: the list is a super-column's (RasterParams::sc_shift): 8-byte row entries, of which the
// tile keeps those whose tile columns cover it -- exactly the tile's own list, in the same order.
// T8: block_width 8 -- a tile IS one 8x8 quadrant, so each of the CTA's four warps takes a whole
// tile of its own (tile 4 blockIdx + warp) and everything else is the same code.
template <bool CULL, bool SUPER, bool T8>
__global__ void __launch_bounds__(RF2, JSB_FWD_CTAS)
raster_fwd_warp_kernel(const RasterParams p, const void *__restrict__ list,
                       const int2 *__restrict__ tile_bins, const SplatRecord *__restrict__ records,
                       const float *__restrict__ background, float *__restrict__ final_Ts,
                       int *__restrict__ final_idx, float *__restrict__ out_img,
                       unsigned *__restrict__ hit_words) {
    __shared__ FwSlot s_slot[RF2 / 32][32];
    __shared__ float2 s_c[RF2 / 32][32];  // b, sorted index (bits)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned TS = T8 ? 8 : 16;  // tile size in pixels
    const unsigned tile_local = T8 ? blockIdx.x * (RF2 / 32) + warp : blockIdx.x;
    if (T8 && tile_local >= p.tiles_x * (p.band_y1 - p.band_y0)) return;  // (no CTA barrier below)
    const unsigned tile = p.band_y0 * p.tiles_x + tile_local;
    const unsigned tile_y = tile / p.tiles_x, tile_x = tile - tile_y * p.tiles_x;
    const unsigned qx = T8 ? 0u : (warp & 1) * 8, qy = T8 ? 0u : (warp >> 1) * 8;  // quadrant in the tile
    const unsigned j = tile_x * TS + qx + (lane & 7);
    const unsigned i0 = tile_y * TS + qy + (lane >> 3), i1 = i0 + 4;
    const float px = (float)j + 0.5f, py0 = (float)i0 + 0.5f, py1 = (float)i1 + 0.5f;
    const bool in0 = (i0 < p.img_h) && (j < p.img_w), in1 = (i1 < p.img_h) && (j < p.img_w);
    bool done0 = !in0, done1 = !in1;
    const unsigned *ids_sorted = reinterpret_cast<const unsigned *>(list);
    const uint2 *ents = reinterpret_cast<const uint2 *>(list);
    const unsigned super_local = (tile_y - p.band_y0) * p.cols2 + (tile_x >> p.sc_shift);
    chain_wait();
    const int2 range = tile_bins[SUPER ? super_local : tile];
    const int ngroups = (range.y - range.x + 31) >> 5;
    const float quad_x0 = (float)(tile_x * TS + qx), quad_y0 = (float)(tile_y * TS + qy);
    const unsigned hw_off = SUPER
        ? (unsigned)hit_quadrant_offset_super(range, super_local, tile_x & ((1u << p.sc_shift) - 1u),
                                              (unsigned)warp, p.sc_shift)
        : T8 ? (unsigned)hit_tile8_offset(range, tile_local)
             : (unsigned)hit_quadrant_offset(range, tile_local, (unsigned)warp);
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned slot_sa = (unsigned)__cvta_generic_to_shared(s_slot[warp]);  // 32 bytes per candidate
    unsigned sc_sa = (unsigned)__cvta_generic_to_shared(s_c[warp]);       // 8 bytes per candidate
    asm volatile("" : "+r"(slot_sa), "+r"(sc_sa));  // opaque: kept in registers, not re-derived

    // {pixel 0, pixel 1} pairs: transmittance and the three colour sums
    float2 T2 = make_float2(1.f, 1.f);
    float2 R2 = make_float2(0.f, 0.f), G2 = R2, B2 = R2;
    int cur0 = 0, cur1 = 0;

    // software pipeline: records of group g in registers, records of g + 1 and ids of g + 2 in flight
    auto load_id = [&](int g) -> unsigned {
        const int idx = range.x + 32 * g + lane;
        if (!(g < ngroups && idx < range.y)) return 0xFFFFFFFFu;
        if (SUPER) {  // the entry's tile columns [x0, x0 + w) must cover this tile
            const uint2 e = __ldg(ents + idx);
            return (tile_x - (e.y & 0xffffu)) < (e.y >> 16) ? e.x : 0xFFFFFFFFu;
        }
        return ids_sorted[idx];
    };
    float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
    float2 qc = make_float2(0.f, 0.f);  // b, cull threshold
    auto load_rec = [&](unsigned id) {
        if (id != 0xFFFFFFFFu) {
            const SplatRecord *r = records + id;
            qa = __ldg(&r->q0);
            qb = __ldg(&r->q1);
            qc = __ldg(reinterpret_cast<const float2 *>(&r->q2));
        }
    };
    unsigned id_cur = load_id(0);
    unsigned id_next = load_id(1);
    load_rec(id_cur);

    bool warp_done = __all_sync(0xffffffffu, done0 && done1);
    for (int g = 0; g < ngroups && !warp_done; ++g) {
        // this group's record (arrived) -> candidate test -> compacted into the warp's slots
        bool may = id_cur != 0xFFFFFFFFu;
        if (CULL && may) {  // the record carries the per-Gaussian part of the test (cull_record_thresh)
            const CullGauss c = cull_from_record(qa.x, qa.y, qa.w, qb.x, qb.y, qc.y);
            may = cull_rect_may_hit(c, quad_x0, quad_y0, 8.f, 8.f);
        }
        const unsigned bits = __ballot_sync(0xffffffffu, may);
        if (may) {
            const unsigned pos = __popc(bits & lt_mask);
            sts128(slot_sa + pos * 32u, qa);
            sts128(slot_sa + pos * 32u + 16u, qb);
            sts64(sc_sa + pos * 8u, make_float2(qc.x, __int_as_float(range.x + 32 * g + lane)));
        }
        // the record registers are free again: the next group's gathers go out and fly while this
        // group is composited
        id_cur = id_next;
        id_next = load_id(g + 2);
        load_rec(id_cur);

        unsigned hit = 0;
        if (bits) {
            __syncwarp();
            const int cnt = __popc(bits);
#pragma unroll 1
            unsigned pa = slot_sa, pc = sc_sa;
            for (int t = 0; t < cnt; ++t, pa += 32u, pc += 8u) {
                const float4 a = lds128(pa);
                const float4 bq = lds128(pa + 16u);
                const float dx = fsub(a.x, px);
                // conic_sigma (common.cuh) for both pixels at once: same operation order per pixel
                // as the reference's compiled code, so the outputs stay bit-identical
                const float2 dy2 = __fadd2_rn(make_float2(a.y, a.y), make_float2(-py0, -py1));
                const float ca = fmul(a.w, dx);                                   // cx dx
                const float ce = fmul(bq.x, dx);                                  // cy dx
                const float2 b2 = __fmul2_rn(make_float2(bq.y, bq.y), dy2);       // cz dy
                const float2 c2 = __fmul2_rn(dy2, b2);
                float2 s2 = __ffma2_rn(make_float2(dx, dx), make_float2(ca, ca), c2);
                s2 = __fmul2_rn(s2, make_float2(0.5f, 0.5f));
                const float2 sg2 = __ffma2_rn(make_float2(ce, ce), dy2, s2);
                const float2 x2 = __fmul2_rn(sg2, make_float2(kNegLog2e, kNegLog2e));
                const float2 ar2 = __fmul2_rn(make_float2(a.z, a.z),
                                              make_float2(ex2_approx(x2.x), ex2_approx(x2.y)));
                const float2 al2 = make_float2(fminf(ar2.x, 0.999f), fminf(ar2.y, 0.999f));
                const bool ok0 = !done0 && !(sg2.x < 0.f) && !(al2.x < kAlphaMin);
                const bool ok1 = !done1 && !(sg2.y < 0.f) && !(al2.y < kAlphaMin);
                // Branch-free compositing of both pixels (the four nested data-dependent branches
                // of the scalar form cost more issue slots in BSSY / BSYNC / BRA than the work they
                // skip): next_T = T (1 - alpha) and vis = T alpha are evaluated for every lane with
                // the reference's operations (1 - alpha as fma(alpha, -1, 1): one rounding of the
                // exact difference, like the subtraction), and a pixel that does not composite
                // this Gaussian -- rejected, saturated, or stopped by it -- gets
                // vis = 0, which leaves its sums unchanged to the bit, and keeps its T.
                const float2 cq = lds64(pc);
                const int sidx = __float_as_int(cq.y);
                const float2 om2 = __ffma2_rn(al2, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
                const float2 nT2 = __fmul2_rn(T2, om2);
                float2 vis2 = __fmul2_rn(T2, al2);
                const bool stop0 = ok0 && (nT2.x <= kTmin), stop1 = ok1 && (nT2.y <= kTmin);
                const bool upd0 = ok0 && !(nT2.x <= kTmin), upd1 = ok1 && !(nT2.y <= kTmin);
                vis2.x = upd0 ? vis2.x : 0.f;
                vis2.y = upd1 ? vis2.y : 0.f;
                R2 = __ffma2_rn(vis2, make_float2(bq.z, bq.z), R2);
                G2 = __ffma2_rn(vis2, make_float2(bq.w, bq.w), G2);
                B2 = __ffma2_rn(vis2, make_float2(cq.x, cq.x), B2);
                T2.x = upd0 ? nT2.x : T2.x;
                T2.y = upd1 ? nT2.y : T2.y;
                cur0 = upd0 ? sidx : cur0;
                cur1 = upd1 ? sidx : cur1;
                done0 = done0 || stop0;
                done1 = done1 || stop1;
                hit |= (ok0 || ok1) ? __funnelshift_l(0u, 1u, (unsigned)(sidx - range.x)) : 0u;  // 1 << (rel & 31)
            }
            __syncwarp();  // the slots are free again
        }
        if (hit_words) {
            hit = __reduce_or_sync(0xffffffffu, hit);
            if (lane == 0) hit_words[hw_off + g] = hit;
        }
        warp_done = __all_sync(0xffffffffu, done0 && done1);
    }

    const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
    if (in0) {
        const size_t pix = (size_t)i0 * p.img_w + j;
        final_Ts[pix] = T2.x;
        final_idx[pix] = cur0;
        out_img[pix * 3 + 0] = ffma(T2.x, bg0, R2.x);
        out_img[pix * 3 + 1] = ffma(T2.x, bg1, G2.x);
        out_img[pix * 3 + 2] = ffma(T2.x, bg2, B2.x);
    }
    if (in1) {
        const size_t pix = (size_t)i1 * p.img_w + j;
        final_Ts[pix] = T2.y;
        final_idx[pix] = cur1;
        out_img[pix * 3 + 0] = ffma(T2.y, bg0, R2.y);
        out_img[pix * 3 + 1] = ffma(T2.y, bg1, G2.y);
        out_img[pix * 3 + 2] = ffma(T2.y, bg2, B2.y);
    }
}

// Any block_width in [1,16]: one CTA of bw*bw threads per tile, thread (ty,tx) = pixel, same
// staging, no culling.  Semantics only; block_size 16 is the tuned case.
__global__ void raster_fwd_generic_kernel(const RasterParams p,
                                          const unsigned *__restrict__ ids_sorted,
                                          const int2 *__restrict__ tile_bins,
                                          const SplatRecord *__restrict__ records,
                                          const float *__restrict__ background,
                                          float *__restrict__ final_Ts,
                                          int *__restrict__ final_idx,
                                          float *__restrict__ out_img) {
    __shared__ float4 sA[RB];
    __shared__ float4 sB[RB];
    __shared__ float sC[RB];
    const int bs = blockDim.x * blockDim.y;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const unsigned tile = (blockIdx.y + p.band_y0) * p.tiles_x + blockIdx.x;
    const unsigned i = (blockIdx.y + p.band_y0) * blockDim.y + threadIdx.y;
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < p.img_h) && (j < p.img_w);
    bool done = !inside;
    const int2 range = tile_bins[tile];
    const int num_batches = (range.y - range.x + bs - 1) / bs;
    float T = 1.f;
    int cur_idx = 0;
    float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f;
    for (int b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(done) >= bs) break;
        const int batch_start = range.x + bs * b;
        const int idx = batch_start + tid;
        if (idx < range.y) {
            const unsigned g = ids_sorted[idx];
            sA[tid] = __ldg(&records[g].q0);
            sB[tid] = __ldg(&records[g].q1);
            sC[tid] = __ldg(&records[g].q2).x;
        }
        __syncthreads();
        const int batch_size = min(bs, range.y - batch_start);
        for (int t = 0; t < batch_size && !done; ++t) {
            const float4 a = sA[t];
            const float4 bq = sB[t];
            const float dx = fsub(a.x, px), dy = fsub(a.y, py);
            const float sigma = conic_sigma(a.w, bq.x, bq.y, dx, dy);
            const float alpha = fminf(fmul(a.z, ex2_approx(fmul(sigma, kNegLog2e))), 0.999f);
            if (sigma < 0.f || alpha < kAlphaMin) continue;
            const float next_T = fmul(T, fsub(1.f, alpha));
            if (next_T <= kTmin) { done = true; break; }
            const float vis = fmul(T, alpha);
            acc_r = ffma(vis, bq.z, acc_r);
            acc_g = ffma(vis, bq.w, acc_g);
            acc_b = ffma(vis, sC[t], acc_b);
            T = next_T;
            cur_idx = batch_start + t;
        }
    }
    if (inside) {
        const size_t pix = (size_t)i * p.img_w + j;
        final_Ts[pix] = T;
        final_idx[pix] = cur_idx;
        out_img[pix * 3 + 0] = ffma(T, __ldg(background), acc_r);
        out_img[pix * 3 + 1] = ffma(T, __ldg(background + 1), acc_g);
        out_img[pix * 3 + 2] = ffma(T, __ldg(background + 2), acc_b);
    }
}

__global__ void fill_background_kernel(const float *__restrict__ background, size_t first,
                                       size_t pixels, float *__restrict__ final_Ts,
                                       int *__restrict__ final_idx, float *__restrict__ out_img) {
    const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
    for (size_t i = first + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < first + pixels;
         i += (size_t)gridDim.x * blockDim.x) {
        final_Ts[i] = 0.f;
        final_idx[i] = 0;
        out_img[i * 3] = bg0;
        out_img[i * 3 + 1] = bg1;
        out_img[i * 3 + 2] = bg2;
    }
}

}  // namespace

bool g_disable_cull = false;  // test hook (capi: jsb200_set_option)
bool g_fwd_batch = false;     // option "fwdwarp" = 0: the per-CTA batch kernel instead of the barrier-free one

int launch_raster_fwd(cudaStream_t stream, const RasterParams &p, const void *list,
                      const int2 *tile_bins, const SplatRecord *records, const float *background,
                      float *final_Ts, int *final_idx, float *out_img, unsigned *hit_words,
                      bool *wrote_hits) {
    const unsigned tiles = p.tiles_x * (p.band_y1 - p.band_y0);
    const unsigned *ids_sorted = reinterpret_cast<const unsigned *>(list);
    if (wrote_hits) *wrote_hits = false;
    if (tiles == 0) return 0;
    if (p.sc_shift && (p.block_width != 16 || g_fwd_batch)) {
        set_error("raster_fwd: super-column lists need the block_width 16 warp kernel");
        return 1;
    }
    if (p.sc_shift) {
        if (wrote_hits) *wrote_hits = hit_words != nullptr;
        if (g_disable_cull)
            launch_chained(raster_fwd_warp_kernel<false, true, false>, tiles, RF2, 0, stream, p, list,
                           tile_bins, records, background, final_Ts, final_idx, out_img, hit_words);
        else
            launch_chained(raster_fwd_warp_kernel<true, true, false>, tiles, RF2, 0, stream, p, list,
                           tile_bins, records, background, final_Ts, final_idx, out_img, hit_words);
    } else if (p.block_width == 16 && !g_fwd_batch) {
        if (wrote_hits) *wrote_hits = hit_words != nullptr;
        if (g_disable_cull)
            launch_chained(raster_fwd_warp_kernel<false, false, false>, tiles, RF2, 0, stream, p, list,
                           tile_bins, records, background, final_Ts, final_idx, out_img, hit_words);
        else
            launch_chained(raster_fwd_warp_kernel<true, false, false>, tiles, RF2, 0, stream, p, list,
                           tile_bins, records, background, final_Ts, final_idx, out_img, hit_words);
    } else if (p.block_width == 8 && !g_fwd_batch) {
        // 8x8 tiles: one warp per tile, four tiles per CTA
        if (wrote_hits) *wrote_hits = hit_words != nullptr;
        const unsigned ctas = (tiles + RF2 / 32 - 1) / (RF2 / 32);
        if (g_disable_cull)
            launch_chained(raster_fwd_warp_kernel<false, false, true>, ctas, RF2, 0, stream, p, list,
                           tile_bins, records, background, final_Ts, final_idx, out_img, hit_words);
        else
            launch_chained(raster_fwd_warp_kernel<true, false, true>, ctas, RF2, 0, stream, p, list,
                           tile_bins, records, background, final_Ts, final_idx, out_img, hit_words);
    } else if (p.block_width == 16) {
        if (wrote_hits) *wrote_hits = hit_words != nullptr;
        if (g_disable_cull)
            raster_fwd16x2_kernel<false><<<tiles, RF2, 0, stream>>>(
                p, ids_sorted, tile_bins, records, background, final_Ts, final_idx, out_img, hit_words);
        else
            raster_fwd16x2_kernel<true><<<tiles, RF2, 0, stream>>>(
                p, ids_sorted, tile_bins, records, background, final_Ts, final_idx, out_img, hit_words);
    } else {
        dim3 grid(p.tiles_x, p.band_y1 - p.band_y0), block(p.block_width, p.block_width);
        raster_fwd_generic_kernel<<<grid, block, 0, stream>>>(p, ids_sorted, tile_bins, records,
                                                              background, final_Ts, final_idx,
                                                              out_img);
    }
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

int launch_fill_background(cudaStream_t stream, const RasterParams &p, const float *background,
                           float *final_Ts, int *final_idx, float *out_img) {
    const size_t row0 = (size_t)p.band_y0 * p.block_width;
    const size_t row1 = min((size_t)p.img_h, (size_t)p.band_y1 * p.block_width);
    if (row1 <= row0 || p.img_w == 0) return 0;
    const size_t first = row0 * p.img_w, pixels = (row1 - row0) * p.img_w;
    const unsigned blocks = (unsigned)((pixels + 255) / 256 < 148 * 8 ? (pixels + 255) / 256 : 148 * 8);
    fill_background_kernel<<<blocks, 256, 0, stream>>>(background, first, pixels, final_Ts,
                                                       final_idx, out_img);
    JSB_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}

}  // namespace jsb
