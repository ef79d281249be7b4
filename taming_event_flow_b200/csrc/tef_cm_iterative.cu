// tef_cm_iterative.cu -- fused Iterative contrast-maximization loss for sm_100a.
//
// Replaces Iterative.forward + autograd (upstream loss/flow.py:492-746, utils/iwe.py)
// with three phases separated by the two global dependencies of the loss
// (SURVEY.md §7 hard part 3):
//
//   iter_fwd_kernel   one thread per (event, flow scale): walks the whole warping
//                     chain on chip (positions at every reference time), forms
//                     the shared border mask, and splats count + time-weighted
//                     images straight into L2-resident slot images with native
//                     16-byte vector reductions (REDG.E.ADD.F32x4, one per image row).
//                     Indices and weights never reach HBM.
//   iwe_reduce/finalize   per-pixel normalisation, sum of squares, non-zero count.
//   iwe_grad_kernel + iter_bwd_kernel   gradient images in place, then one thread
//                     per gradient-carrying event: read the chain positions the
//                     forward kernel kept (88 B/event, coalesced), gather the
//                     gradient images at the corners, walk the chain in reverse and
//                     reduce into the packed flow-gradient maps (REDG F32x4).
//
// Shared-memory tiles are deliberately NOT used: on sm_100a fp32 atomicAdd on
// shared memory compiles to an ATOMS.CAST.SPIN compare-and-swap loop, whereas
// global fp32 reductions are native fire-and-forget REDG ops served by the L2
// (126 MB, holds all slot images of every benchmark configuration).  See DESIGN.md.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

#ifndef TEF_FWD_MIN_BLOCKS
#define TEF_FWD_MIN_BLOCKS 5
#endif
#ifndef TEF_BWD_MIN_BLOCKS
#define TEF_BWD_MIN_BLOCKS 4
#endif

namespace tef {

// ---------------------------------------------------------------------------------
// per-event chain (loss/flow.py:521-586): node positions for tref = 0..P.  The positions
// live in shared memory as pos[tref][thread] (conflict-free, 8 B per lane), which keeps the
// loops rolled and the register count low for any P; they never travel to HBM.
// The pass index t is uniform per CTA, so the loop bounds do not diverge.
// ---------------------------------------------------------------------------------
// One chain step from position q = (x, y): sample the map, move by dt * flow (utils/iwe.py:14).  The move stays scalar:
// ptxas contracts a packed product feeding a packed sum into FFMA2 (see tef_device.cuh), the reference rounds twice.
// QUAD: `map` walks the quad-cell copies (one 256-bit gather per sample); only the event's own location can lie outside the
// sensor, and that first step of either chain samples the map of the event's own pass, `own` (dual-phase rows, generic sample).
template <bool QUAD>
__device__ __forceinline__ float2 chain_step(const float2 *__restrict__ map, const float2 *__restrict__ own, const Res &r, float2 q, float dt, bool safe) {
    float2 v;                                          // (x-flow, y-flow)
    if (safe) v = sample_flow_inside_xy<false, QUAD>(map, r, q, nullptr);
    else v = sample_flow<false>(QUAD ? own : map, r, q.y, q.x, nullptr);
    return make_float2(q.x + dt * v.x, q.y + dt * v.y);
}
template <bool QUAD>
__device__ __forceinline__ uint32_t warp_chain(const CmParams &p, const float2 *__restrict__ flow_fb /* dual-phase maps of (f, sample b), pass 0 */,
                                               const float2 *__restrict__ walk_fb /* what the chain walks: flow_fb, or the quad-cell copies of (f, b), pass 0 */,
                                               int t, float ts, float y0, float x0, float2 *__restrict__ pos /* [P+1][kThreads], (x, y) */,
                                               float2 *__restrict__ pb /* this row of posbuf (stride rows_grad) or nullptr */,
                                               int keep_lo, int keep_hi /* nodes some scale splats for this pass; the others only feed the mask */) {
    const long stride = QUAD ? (long)p.B * 16 * p.res.cplane : (long)p.B * 2 * p.res.fplane;   // one pass further, in float2 units
    const float2 *own = flow_fb + (long)t * p.B * 2 * p.res.fplane;
    uint32_t alive = 0;
    // the event's own location may lie outside the sensor (generic sample); every later position is inside
    const bool in0 = inside(y0, x0, p.res);
#if TEF_CHAIN_INTERLEAVE
    // Variant (DESIGN.md decision 21): the forward and the backward chain of an event are independent, so they advance in the
    // same iteration -- two gathers in flight per thread instead of one.  Dead lanes sample position (0, 0) (a broadcast load)
    // instead of branching around the sample, so the two samples of an iteration sit in one basic block.
    if (!QUAD && in0) {
        float2 qf = make_float2(x0, y0), qb = qf;
        bool alf = true, alb = true;
        const float2 *mf = walk_fb + (long)t * stride, *mb = mf;
        float2 *pwf = pos + (t + 1) * kThreads + threadIdx.x, *pwb = pos + t * kThreads + threadIdx.x;
        const int nf = p.P - t, nb = t + 1, nc = min(nf, nb);
        float dtf = (float)(t + 1) - ts, dtb = (float)t - ts;
        const float2 zero = make_float2(0.f, 0.f);
        int k = 0;
        for (; k < nc; ++k, mf += stride, mb -= stride, pwf += kThreads, pwb -= kThreads) {
            const int trf = t + 1 + k, trb = t - k;
            const float2 vf = sample_flow_inside_xy<false, false>(mf, p.res, alf ? qf : zero, nullptr);
            const float2 vb = sample_flow_inside_xy<false, false>(mb, p.res, alb ? qb : zero, nullptr);
            if (alf) { qf = make_float2(qf.x + dtf * vf.x, qf.y + dtf * vf.y); alf = inside(qf.y, qf.x, p.res); if (alf) alive |= (1u << trf); }
            if (alb) { qb = make_float2(qb.x + dtb * vb.x, qb.y + dtb * vb.y); alb = inside(qb.y, qb.x, p.res); if (alb) alive |= (1u << trb); }
            dtf = 1.0f; dtb = -1.0f;
            if (trf <= keep_hi) { *pwf = qf; if (pb) __stcs(pb + (long)trf * p.rows_grad, qf); }
            if (trb >= keep_lo) { *pwb = qb; if (pb) __stcs(pb + (long)trb * p.rows_grad, qb); }
        }
        for (int kk = k; kk < nf; ++kk, mf += stride, pwf += kThreads) {
            const int trf = t + 1 + kk;
            if (alf) {
                const float2 vf = sample_flow_inside_xy<false, false>(mf, p.res, qf, nullptr);
                qf = make_float2(qf.x + dtf * vf.x, qf.y + dtf * vf.y); alf = inside(qf.y, qf.x, p.res); if (alf) alive |= (1u << trf);
            }
            dtf = 1.0f;
            if (trf <= keep_hi) { *pwf = qf; if (pb) __stcs(pb + (long)trf * p.rows_grad, qf); }
        }
        for (int kk = k; kk < nb; ++kk, mb -= stride, pwb -= kThreads) {
            const int trb = t - kk;
            if (alb) {
                const float2 vb = sample_flow_inside_xy<false, false>(mb, p.res, qb, nullptr);
                qb = make_float2(qb.x + dtb * vb.x, qb.y + dtb * vb.y); alb = inside(qb.y, qb.x, p.res); if (alb) alive |= (1u << trb);
            }
            dtb = -1.0f;
            if (trb >= keep_lo) { *pwb = qb; if (pb) __stcs(pb + (long)trb * p.rows_grad, qb); }
        }
        return alive;
    }
#endif
    float2 q = make_float2(x0, y0);
    float tprev = ts;
    bool al = true, safe = in0;
    const float2 *map = walk_fb + (long)t * stride;
    float2 *pw = pos + (t + 1) * kThreads + threadIdx.x;
    for (int tr = t + 1; tr <= p.P; ++tr, map += stride, pw += kThreads) {   // forward: sample map tr-1, land on node tr
        if (al) {
            q = chain_step<QUAD>(map, own, p.res, q, (float)tr - tprev, safe);
            al = inside(q.y, q.x, p.res);              // utils/iwe.py:52-59
            if (al) alive |= (1u << tr);
            safe = true;
        }
        tprev = (float)tr;
        if (tr <= keep_hi) {
            *pw = q;
            if (pb) __stcs(pb + (long)tr * p.rows_grad, q);   // coalesced streaming store (evict-first), kept for the backward kernel
        }
    }
    q = make_float2(x0, y0); tprev = ts; al = true; safe = in0;
    map = walk_fb + (long)t * stride;
    pw = pos + t * kThreads + threadIdx.x;
    for (int tr = t; tr >= 0; --tr, map -= stride, pw -= kThreads) {         // backward: sample map tr, land on node tr
        if (al) {
            q = chain_step<QUAD>(map, own, p.res, q, (float)tr - tprev, safe);
            al = inside(q.y, q.x, p.res);
            if (al) alive |= (1u << tr);
            safe = true;
        }
        tprev = (float)tr;
        if (tr >= keep_lo) {
            *pw = q;
            if (pb) __stcs(pb + (long)tr * p.rows_grad, q);
        }
    }
    return alive;
}

// Sub-window bookkeeping of one temporal scale for the events of pass t (loss/flow.py:657-686).  The pass index is
// uniform per CTA, so the table is built once per CTA in shared memory (no per-thread integer division):
// tr0..tr1 = reference times this pass feeds at the scale (max(lo, tr-delta) <= t < min(hi, tr+delta), :685-686),
// mask = reference times whose in-image bits form the shared border mask (:671-681).
struct WinS {
    int valid, tr0, tr1, slot0;
    uint32_t mask;
    float fdelta, rdelta;
};
__device__ __forceinline__ void build_windows(const CmParams &p, int t, WinS *sw) {
    const int s = threadIdx.x;
    if (s < p.sc.S) {
        WinS w;
        const int L = p.sc.L[s];
        w.valid = t < (L << s);                    // passes beyond 2^s windows are unused at this scale
        const int wi = t / L, delta = p.sc.delta[s];
        const int lo = wi * L, hi = lo + L;
        int low_tref = lo, high_tref = hi + 1;
        if (p.mode == 4) { low_tref = lo + delta; high_tref = lo + 3 * delta + 1; }
        w.mask = (uint32_t)((1ull << high_tref) - 1ull) & ~((1u << low_tref) - 1u);
        w.tr0 = max(low_tref, t - delta + 1); w.tr1 = min(high_tref - 1, t + delta);
        w.slot0 = p.sc.slot_base[s] + wi * p.sc.ntau[s] - low_tref;
        w.fdelta = (float)delta; w.rdelta = 1.0f / w.fdelta;
        sw[s] = w;
    }
    __syncthreads();
}
// scales at which this event is splatted: inside a sub-window and, with border compensation, alive at all its reference times
__device__ __forceinline__ uint32_t active_scales(const CmParams &p, const WinS *sw, uint32_t alive) {
    uint32_t has = 0;
    for (int s = 0; s < p.sc.S; ++s)
        if (sw[s].valid && (!p.border || (alive & sw[s].mask) == sw[s].mask)) has |= 1u << s;
    return has;
}

template <bool DET, bool QUAD>
__global__ void __launch_bounds__(kThreads, TEF_FWD_MIN_BLOCKS) iter_fwd_kernel(const __grid_constant__ CmParams p) {
    extern __shared__ float2 pos[];
    __shared__ WinS sw[TEF_MAX_SCALES];
    int t, b, row, set; float4 e; float2 m;
    const bool live = locate_sorted(p, t, b, e, m, row, set);
    build_windows(p, t, sw);
    const int f = blockIdx.y;
    const long slot_stride = (DET ? 8 : 4) * p.ig.plane;          // float2 elements per slot (int64 pairs in deterministic mode)
    float2 *img_fb = nullptr;
    uint32_t alive = 0, has = 0;
    if (live) {                                                   // threads without an event stay for the warp-level merge below
        // gradient-carrying rows keep their chain for the backward kernel
        float2 *pb = (set == 0 && p.posbuf) ? p.posbuf + (long)f * (p.P + 1) * p.rows_grad + row : nullptr;
        // nodes that any scale splats for this pass (uniform per CTA): only those are kept, on chip and for the backward kernel
        int keep_lo = p.P + 1, keep_hi = -1;
        for (int s = 0; s < p.sc.S; ++s)
            if (sw[s].valid) { keep_lo = min(keep_lo, sw[s].tr0); keep_hi = max(keep_hi, sw[s].tr1); }
        const float2 *flow_fb = p.flow + ((long)f * p.P * p.B + b) * 2 * p.res.fplane;
        const float2 *walk_fb = QUAD ? reinterpret_cast<const float2 *>(p.flowq + ((long)f * p.P * p.B + b) * 8 * p.res.cplane) : flow_fb;
        alive = warp_chain<QUAD>(p, flow_fb, walk_fb, t, e.x, e.y, e.z, pos, pb, keep_lo, keep_hi);
        if (pb) p.alivebuf[(long)f * p.rows_grad + row] = alive;
        has = active_scales(p, sw, alive);
        img_fb = p.img + ((long)f * p.B + b) * p.nslots * slot_stride;
    }
    // The loader's masks are one-hot {0,1}: one polarity plane, weights not scaled, every splatted node inside the image.
    // If that holds for every event of the warp, the warp splats together and merges equal slots of neighbouring lanes.
    const bool fast = !DET && p.border && (!has || (m.x == 1.0f && m.y == 0.0f) || (m.x == 0.0f && m.y == 1.0f));
    if (__all_sync(0xffffffffu, fast)) {
        const unsigned lane = threadIdx.x & 31u;
        const int pol = (has && m.x == 0.0f) ? 1 : 0;
        const float ts = has ? e.x : 0.0f;
        const unsigned key_base = has ? (unsigned)b * (unsigned)(4 * p.ig.plane) : 0u;
        for (int s = 0; s < p.sc.S; ++s) {
            const WinS w = sw[s];
            if (!w.valid) continue;                               // uniform per CTA, like the range of reference times
            const bool on = ((has >> s) & 1u) != 0u;
            float2 *slot = img_fb + (long)(w.slot0 + w.tr0) * slot_stride;
            const float2 *pq = pos + w.tr0 * kThreads + threadIdx.x;
            for (int tr = w.tr0; tr <= w.tr1; ++tr, slot += slot_stride, pq += kThreads) {
                const float nts = 1.0f - div_const(fabsf((float)tr - ts), w.fdelta, w.rdelta);   // loss/flow.py:94-95
                splat_inside_1hot_warp(slot, p.ig, *pq, nts, pol, on, lane, key_base);
            }
        }
        return;
    }
    if (!has) return;
    for (int s = 0; s < p.sc.S; ++s) {
        if (!((has >> s) & 1u)) continue;
        const WinS w = sw[s];
        for (int tr = w.tr0; tr <= w.tr1; ++tr) {
            if (!p.border && !((alive >> tr) & 1u)) continue;
            const float nts = 1.0f - div_const(fabsf((float)tr - e.x), w.fdelta, w.rdelta);   // loss/flow.py:94-95
            const float2 q = pos[tr * kThreads + threadIdx.x];
            splat<true, DET>(img_fb + (long)(w.slot0 + tr) * slot_stride, p.res, p.ig, q.y, q.x, nts, m);
        }
    }
}

// one reverse chain step: reduce dL/dmap, return dL/d(source position)
template <bool DET>
__device__ __forceinline__ void step_bwd(const float2 *__restrict__ map, float2 *__restrict__ gmap, const Res &r, const ImgGeom &g, float sy,
                                         float sx, float dt, float gpy, float gpx, float &cy_, float &cx_, float ginv = 1.0f) {
    Taps tp;
    if (inside(sy, sx, r)) sample_flow_inside<true>(map, r, sy, sx, &tp);
    else sample_flow<true>(map, r, sy, sx, &tp);
    taps_red<DET>(gmap, g, tp, dt, gpy, gpx, ginv);
    const float dvy_dy = (1.0f - tp.ax) * (tp.v[2].y - tp.v[0].y) + tp.ax * (tp.v[3].y - tp.v[1].y);
    const float dvy_dx = (1.0f - tp.ay) * (tp.v[1].y - tp.v[0].y) + tp.ay * (tp.v[3].y - tp.v[2].y);
    const float dvx_dy = (1.0f - tp.ax) * (tp.v[2].x - tp.v[0].x) + tp.ax * (tp.v[3].x - tp.v[1].x);
    const float dvx_dx = (1.0f - tp.ay) * (tp.v[1].x - tp.v[0].x) + tp.ay * (tp.v[3].x - tp.v[2].x);
    cy_ = gpy + dt * (dvy_dy * gpy + dvx_dy * gpx);
    cx_ = gpx + dt * (dvy_dx * gpy + dvx_dx * gpx);
}

// step_bwd for a whole warp: the two tap-row reductions of neighbouring lanes that hit the same slot of the same
// flow-gradient map are merged before they leave the SM (merge_equal_neighbours).  `on` = this lane has a gradient to
// push through the step; the others only take part in the shuffles.  A row is reduced when its merged values are not all
// zero: a row outside the map (or with zero weights) carries zeros in every lane of the cell, and adding zeros is a no-op.
// Must be called by all 32 lanes.  Non-deterministic mode only.
template <bool QUAD>
__device__ __forceinline__ void step_bwd_warp(const float2 *__restrict__ map, const CmParams &p, long own_index /* ((f * P + t) * B + b): the event's own pass */,
                                              float2 *__restrict__ gmap, const Res &r, const ImgGeom &g, float sy,
                                              float sx, float dt, float gpy, float gpx, float &cy_, float &cx_, bool on, unsigned lane,
                                              unsigned key_base /* sample * map size */) {
    if (on && !inside(sy, sx, r)) {                 // an event whose own location lies outside the sensor (first step only): generic sample, not merged
        // (with QUAD `map` walks the quad-cell copies; the generic sample reads the dual-phase rows of the event's own pass)
        step_bwd<false>(QUAD ? p.flow + own_index * 2 * r.fplane : map, gmap, r, g, sy, sx, dt, gpy, gpx, cy_, cx_);
        on = false;
    }
    float v[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    unsigned off = 0;
    if (on) {
        Taps tp;
        sample_flow_inside<true, QUAD>(map, r, sy, sx, &tp);
        const int phase = tp.x0 & 1;
        off = (unsigned)(phase * (int)g.plane + tp.y0 * g.Wp + tp.x0 + phase);
        // taps_red's coefficients and values, and step_bwd's Jacobian, on packed fp32x2 arithmetic: same products, rounded
        // one by one (sums of two products stay scalar: ptxas would contract a packed product into a packed sum)
        const float2 c01 = mul2(bc(dt), make_float2(tp.w[0], tp.w[1])), c23 = mul2(bc(dt), make_float2(tp.w[2], tp.w[3]));
        const float c0 = c01.x, c1 = tp.ok[1] ? c01.y : 0.0f, c2 = tp.ok[2] ? c23.x : 0.0f, c3 = tp.ok[3] ? c23.y : 0.0f;
        const float2 gp = make_float2(gpx, gpy);
        const float2 r0 = mul2(bc(c0), gp), r1 = mul2(bc(c1), gp), r2 = mul2(bc(c2), gp), r3 = mul2(bc(c3), gp);
        v[0] = r0.x; v[1] = r0.y; v[2] = r1.x; v[3] = r1.y; v[4] = r2.x; v[5] = r2.y; v[6] = r3.x; v[7] = r3.y;
        // d(flow)/dy and d(flow)/dx as (x-flow, y-flow) pairs
        const float2 a20 = sub2(tp.v[2], tp.v[0]), a31 = sub2(tp.v[3], tp.v[1]), a10 = sub2(tp.v[1], tp.v[0]), a32 = sub2(tp.v[3], tp.v[2]);
        const float2 py0 = mul2(bc(1.0f - tp.ax), a20), py1 = mul2(bc(tp.ax), a31);
        const float2 px0 = mul2(bc(1.0f - tp.ay), a10), px1 = mul2(bc(tp.ay), a32);
        const float dvx_dy = py0.x + py1.x, dvy_dy = py0.y + py1.y, dvx_dx = px0.x + px1.x, dvy_dx = px0.y + px1.y;
        cy_ = gpy + dt * (dvy_dy * gpy + dvx_dy * gpx);
        cx_ = gpx + dt * (dvy_dx * gpy + dvx_dx * gpx);
    }
    const bool gave = merge_equal_neighbours<8>(on ? key_base + off : (0x80000000u | lane), lane, v);
    if (on && !gave) {
        if (v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f || v[3] != 0.0f) red_add_v4(gmap + off, v[0], v[1], v[2], v[3]);
        if (v[4] != 0.0f || v[5] != 0.0f || v[6] != 0.0f || v[7] != 0.0f) red_add_v4(gmap + (off + (unsigned)g.Wp), v[4], v[5], v[6], v[7]);
    }
}

// One thread per gradient-carrying event.  Chain positions come from the forward kernel's posbuf
// (coalesced loads); per node: gather the gradient images at the corners, add what flows back from the
// next node, reduce into the packed flow-gradient map and step towards the event's own window.
// The node loops run over the CTA-uniform range of nodes any scale splats for this pass, and (outside the deterministic
// mode) threads without an event or without a gradient stay in them for the warp-level merge of the reductions.
template <bool DET, bool QUAD>
__global__ void __launch_bounds__(kThreads, TEF_BWD_MIN_BLOCKS) iter_bwd_kernel(const __grid_constant__ CmParams p) {
    __shared__ WinS sw[TEF_MAX_SCALES];
    int t, b, row, set; float4 e; float2 m;
    const bool live = locate_sorted(p, t, b, e, m, row, set);
    build_windows(p, t, sw);
    if (DET && !live) return;
    const int f = blockIdx.y;
    const long HW = 2 * p.res.fplane;                              // float2 elements per (pass, sample) dual-phase flow map
    const long HWq = QUAD ? 16 * (long)p.res.cplane : HW;          // ... of what the steps sample (quad-cell copy or the same)
    const float2 *flow_f = QUAD ? reinterpret_cast<const float2 *>(p.flowq) + (long)f * p.P * p.B * HWq : p.flow + (long)f * p.P * p.B * HW;
    const long gmap_sz = (DET ? 4 : 2) * p.ig.plane;               // float2 elements per (pass, sample) gradient map
    float2 *gflow_f = p.gflow + (long)f * p.P * p.B * gmap_sz;
    uint32_t alive = 0, has = 0;
    const float2 *pb = nullptr;
    float ts = 0.f, y0 = 0.f, x0 = 0.f;
    if (live) {
        ts = e.x; y0 = e.y; x0 = e.z;
        alive = p.alivebuf[(long)f * p.rows_grad + row];
        pb = p.posbuf + (long)f * (p.P + 1) * p.rows_grad + row;
        // scales whose sub-window takes this event
        has = active_scales(p, sw, alive);
    } else {
        b = 0; m = make_float2(0.f, 0.f);
    }
    if (DET && !has) return;
    const bool act = has != 0u;
    // nodes that receive an image gradient at some scale: the CTA-uniform range the forward kernel kept
    int lo_node = p.P + 1, hi_node = -1;
    for (int s = 0; s < p.sc.S; ++s)
        if (sw[s].valid) { lo_node = min(lo_node, sw[s].tr0); hi_node = max(hi_node, sw[s].tr1); }
    // gradient images: [phase][pol][H][Wp] float2 per slot -- in place in img, or gimg in deterministic mode
    const long gslot = 4 * p.ig.plane;
    const float2 *img_fb = (DET ? p.gimg : p.img) + ((long)f * p.B + b) * p.nslots * gslot;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned key_base = (unsigned)b * (unsigned)gmap_sz;

    // scale 0 (the only one in every shipped config) is kept in registers: the window table is otherwise re-read from
    // shared memory at every node (14 M broadcast loads per launch, 11 % of the kernel's L1TEX wavefronts)
    const int w0_tr0 = (has & 1u) ? sw[0].tr0 : 1, w0_tr1 = (has & 1u) ? sw[0].tr1 : 0, w0_slot0 = sw[0].slot0;
    const float w0_fdelta = sw[0].fdelta, w0_rdelta = sw[0].rdelta;
    // the loader's one-hot masks read one polarity plane and skip the mask products
    const bool onehot = (m.x == 1.0f && m.y == 0.0f) || (m.x == 0.0f && m.y == 1.0f);
    const float2 *img_pol = img_fb + (m.x != 0.0f ? 0 : p.ig.plane);
    // quad-cell gradient images of this sample: [slot][pol][4][cplane] cells of two float4
    const float4 *gq_pol = QUAD ? p.gimgq + ((((long)f * p.B + b) * p.nslots) * 2 + (m.x != 0.0f ? 0 : 1)) * 8 * p.res.cplane : nullptr;
    const long own_index = ((long)f * p.P + t) * p.B + b;
    auto node_grad = [&](int tr, float2 q, float &gy, float &gx) {
        if (tr >= w0_tr0 && tr <= w0_tr1) {
            const float nts = 1.0f - div_const(fabsf((float)tr - ts), w0_fdelta, w0_rdelta);
            if (onehot) {
                if constexpr (QUAD) iwe_grad_inside_1hot_quad(gq_pol + (long)(w0_slot0 + tr) * 16 * p.res.cplane, p.res, q, nts, gy, gx);
                else iwe_grad_inside_1hot(img_pol + (long)(w0_slot0 + tr) * gslot, p.res, p.ig, q, nts, gy, gx);
            }
            else iwe_grad<true>(img_fb + (long)(w0_slot0 + tr) * gslot, p.res, p.ig, q.y, q.x, nts, m, gy, gx);
        }
        for (int s = 1; s < p.sc.S; ++s) {
            if (!((has >> s) & 1u) || tr < sw[s].tr0 || tr > sw[s].tr1) continue;
            const float nts = 1.0f - div_const(fabsf((float)tr - ts), sw[s].fdelta, sw[s].rdelta);
            iwe_grad<true>(img_fb + (long)(sw[s].slot0 + tr) * gslot, p.res, p.ig, q.y, q.x, nts, m, gy, gx);
        }
    };

    const long map_stride = (long)p.B * HWq, gmap_stride = (long)p.B * gmap_sz;
    const float ginv = DET ? __ldg(p.den + p.F * p.B * p.nslots + 1) : 1.0f;     // 1 / scale of the fixed-point gradient words
    // reverse of the forward chain: nodes hi_node .. t+1 (nodes beyond carry no gradient)
    float cy_ = 0.f, cx_ = 0.f;
    {
        int tr = min(hi_node, p.P);
        const float2 *pq = pb + (long)tr * p.rows_grad;                  // position of node tr; one row_grad back: node tr-1
        const float2 *map = flow_f + (long)(tr - 1) * map_stride + (long)b * HWq;
        float2 *gmap = gflow_f + (long)(tr - 1) * gmap_stride + (long)b * gmap_sz;
        float2 q = (act && tr >= t + 1) ? __ldcs(pq) : make_float2(0.f, 0.f);
        for (; tr >= t + 1; --tr, pq -= p.rows_grad, map -= map_stride, gmap -= gmap_stride) {
            const bool first = (tr - 1 == t);
            const float2 src = first ? make_float2(x0, y0) : (act ? __ldcs(pq - p.rows_grad) : make_float2(0.f, 0.f));
            const bool al = ((alive >> tr) & 1u) != 0;                   // alive == 0 without an event
            float gy = 0.f, gx = 0.f;
            if (al) node_grad(tr, q, gy, gx);
            const float gpy = al ? gy + cy_ : 0.f, gpx = al ? gx + cx_ : 0.f;
            cy_ = 0.f; cx_ = 0.f;
            const bool on = gpy != 0.f || gpx != 0.f;
            const float dt = first ? ((float)tr - ts) : 1.0f;
            if (!DET) step_bwd_warp<QUAD>(map, p, own_index, gmap, p.res, p.ig, src.y, src.x, dt, gpy, gpx, cy_, cx_, on, lane, key_base);
            else if (on) step_bwd<DET>(map, gmap, p.res, p.ig, src.y, src.x, dt, gpy, gpx, cy_, cx_, ginv);
            q = src;
        }
    }
    // reverse of the backward chain: nodes lo_node .. t
    cy_ = 0.f; cx_ = 0.f;
    {
        int tr = max(lo_node, 0);
        const float2 *pq = pb + (long)tr * p.rows_grad;
        const float2 *map = flow_f + (long)tr * map_stride + (long)b * HWq;
        float2 *gmap = gflow_f + (long)tr * gmap_stride + (long)b * gmap_sz;
        float2 q = (act && tr <= t) ? __ldcs(pq) : make_float2(0.f, 0.f);
        for (; tr <= t; ++tr, pq += p.rows_grad, map += map_stride, gmap += gmap_stride) {
            const bool first = (tr == t);
            const float2 src = first ? make_float2(x0, y0) : (act ? __ldcs(pq + p.rows_grad) : make_float2(0.f, 0.f));
            const bool al = ((alive >> tr) & 1u) != 0;
            float gy = 0.f, gx = 0.f;
            if (al) node_grad(tr, q, gy, gx);
            const float gpy = al ? gy + cy_ : 0.f, gpx = al ? gx + cx_ : 0.f;
            cy_ = 0.f; cx_ = 0.f;
            const bool on = gpy != 0.f || gpx != 0.f;
            const float dt = first ? ((float)tr - ts) : -1.0f;
            if (!DET) step_bwd_warp<QUAD>(map, p, own_index, gmap, p.res, p.ig, src.y, src.x, dt, gpy, gpx, cy_, cx_, on, lane, key_base);
            else if (on) step_bwd<DET>(map, gmap, p.res, p.ig, src.y, src.x, dt, gpy, gpx, cy_, cx_, ginv);
            q = src;
        }
    }
}

}  // namespace tef

// ---------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------
using namespace tef;

int tef_sort_events(const CmParams &p, cudaStream_t st);           // tef_cm_sort.cu
int tef_reduce_and_finalize(const CmParams &p, cudaStream_t st);   // tef_cm_reduce.cu
int tef_grad_images(const CmParams &p, cudaStream_t st);           // tef_cm_reduce.cu

static size_t chain_smem(const CmParams &p) { return sizeof(float2) * (size_t)(p.P + 1) * kThreads; }

static int launch_fwd(const CmParams &p, cudaStream_t st) {
    if (p.seg.blk_off[p.seg.nseg] > 0) {
        // the opt-in to more than 48 KB of dynamic shared memory (P >= 23) is a per-device attribute of the function
        static unsigned long long attr_done = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !((attr_done >> dev) & 1ull)) {
            const int mx = (int)(sizeof(float2) * (TEF_MAX_PASSES + 1) * kThreads);
            cudaFuncSetAttribute(iter_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            cudaFuncSetAttribute(iter_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            cudaFuncSetAttribute(iter_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
            if (dev >= 0 && dev < 64) attr_done |= 1ull << dev;
        }
        dim3 grid(p.seg.blk_off[p.seg.nseg], p.F);
        ProfScope ps(K_ITER_FWD, st);
        if (p.det) iter_fwd_kernel<true, false><<<grid, kThreads, chain_smem(p), st>>>(p);
        else if (p.flowq) iter_fwd_kernel<false, true><<<grid, kThreads, chain_smem(p), st>>>(p);
        else iter_fwd_kernel<false, false><<<grid, kThreads, chain_smem(p), st>>>(p);
    }
    return (int)cudaGetLastError();
}
static int launch_bwd(const CmParams &p, cudaStream_t st) {
    if (p.seg.blk_off[p.seg.nseg] > 0) {
        dim3 grid(p.seg.blk_off[p.seg.nseg], p.F);
        ProfScope ps(K_ITER_BWD, st);
        if (p.det) iter_bwd_kernel<true, false><<<grid, kThreads, 0, st>>>(p);
        else if (p.flowq && p.gimgq) iter_bwd_kernel<false, true><<<grid, kThreads, 0, st>>>(p);
        else iter_bwd_kernel<false, false><<<grid, kThreads, 0, st>>>(p);
    }
    return (int)cudaGetLastError();
}

extern "C" int tef_iterative_forward(const tef_cm_desc *d, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CmParams p;
    int rc = fill_params(d, 0, p);
    if (rc) return rc;
    if (!p.flow || !p.img || !p.acc_sum || !p.acc_nnz || !p.den || !p.loss) return TEF_EINVAL;
    const long nimg = (long)p.F * p.B * p.nslots;
    cudaMemsetAsync(p.img, 0, sizeof(float2) * nimg * (p.det ? 16 : 4) * p.ig.plane, st);     // deterministic: high and low words
    rc = tef_sort_events(p, st);
    if (rc) return rc;
    rc = launch_fwd(p, st);
    if (rc) return rc;
    return tef_reduce_and_finalize(p, st);
}

extern "C" int tef_iterative_backward(const tef_cm_desc *d, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CmParams p;
    int rc = fill_params(d, 0, p);                 // same segment / bin layout as the forward call
    if (rc) return rc;
    if (!p.flow || !p.gflow || !p.img || !p.den || !p.grad_out || !p.sort.bins || !p.sort.rec) return TEF_EINVAL;
    if (p.rows_grad > 0 && (!p.posbuf || !p.alivebuf)) return TEF_EINVAL;
    grad_segments_only(p, (long)p.B * p.nslots * 4 * p.ig.plane * 8);
    if (p.det && !p.gimg) return TEF_EINVAL;
    cudaMemsetAsync(p.gflow, 0, sizeof(float2) * (long)p.F * p.P * p.B * (p.det ? 4 : 2) * p.ig.plane, st);
    rc = tef_grad_images(p, st);
    if (rc) return rc;
    rc = launch_bwd(p, st);
    return rc;
}
