// tef_cm_linear.cu -- fused Linear contrast-maximization loss (upstream loss/flow.py:216-412).
//
// The per-event flow vector that upstream samples in `update` (:266-285) is sampled inside the kernels
// from the packed map of the event's own pass (same values, no per-event buffer in HBM);
// `forward` warps every event linearly to both ends of its sub-window, applies the shared
// border mask and splats into two image slots per sub-window; the backward gathers the
// gradient images at both ends and reduces the per-event flow gradient into the packed
// flow-gradient map through the bilinear taps of the original location.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

// positions at both window ends (:337-343); returns the shared mask bit
__device__ __forceinline__ bool linear_ends(const CmParams &p, float lo, float hi, float ts, float y0, float x0, float2 v /* (y,x) */,
                                            float2 &fw, float2 &bw, float &dth, float &dtl) {
    dth = hi - ts; dtl = lo - ts;
    fw = make_float2(y0 + dth * v.x, x0 + dth * v.y);
    bw = make_float2(y0 + dtl * v.x, x0 + dtl * v.y);
    if (!p.border) return true;
    return inside(fw.x, fw.y, p.res) && inside(bw.x, bw.y, p.res);
}

// the event's own flow vector (loss/flow.py:266-285): generic sample, or the bounds-free packed path when the location is inside
// the sensor (bit-identical there)
template <bool KEEP>
__device__ __forceinline__ float2 own_flow(const float2 *__restrict__ map, const Res &r, float y, float x, Taps *tp) {
    if (inside(y, x, r)) return sample_flow_inside_xy<KEEP>(map, r, make_float2(x, y), tp);
    return sample_flow<KEEP>(map, r, y, x, tp);
}

// One thread per event.  With border compensation and the loader's one-hot masks (every shipped configuration) the warp
// splats together: packed fp32x2 corner weights and equal slots of neighbouring lanes merged before they leave the SM
// (splat_inside_1hot_warp, the forward path of the Iterative kernel); threads without an event stay for the shuffles.
template <bool DET>
__global__ void __launch_bounds__(kThreads) linear_fwd_kernel(const __grid_constant__ CmParams p) {
    int t, b, row, set; float4 e; float2 m;
    const bool live = locate_sorted(p, t, b, e, m, row, set);
    if (DET && !live) return;
    if (!live) { b = 0; e = make_float4(0.f, 0.f, 0.f, 0.f); m = make_float2(0.f, 0.f); }
    const int f = blockIdx.y;
    float2 v = make_float2(0.f, 0.f);
    if (live) {
        const float2 vxy = own_flow<false>(p.flow + (((long)f * p.P + t) * p.B + b) * 2 * p.res.fplane, p.res, e.y, e.z, nullptr);
        v = make_float2(vxy.y, vxy.x);                                 // (y, x), utils/iwe.py:38
    }
    const long slot_stride = (DET ? 8 : 4) * p.ig.plane;
    float2 *img_fb = p.img + ((long)f * p.B + b) * p.nslots * slot_stride;
    const bool onehot = (m.x == 1.0f && m.y == 0.0f) || (m.x == 0.0f && m.y == 1.0f);
    const bool fast = !DET && p.border && (!live || onehot);
    if (__all_sync(0xffffffffu, fast)) {
        const unsigned lane = threadIdx.x & 31u;
        const int pol = m.x == 0.0f ? 1 : 0;
        const unsigned key_base = (unsigned)b * (unsigned)(4 * p.ig.plane);
        for (int s = 0; s < p.sc.S; ++s) {
            const int L = p.sc.L[s];
            if (t >= (L << s)) continue;                               // uniform per CTA (the pass index is)
            const int wi = t / L, lo = wi * L, hi = lo + L;
            float2 fw = make_float2(0.f, 0.f), bw = fw; float dth, dtl;
            const bool on = live && linear_ends(p, (float)lo, (float)hi, e.x, e.y, e.z, v, fw, bw, dth, dtl);
            const int slot = p.sc.slot_base[s] + wi * 2;
            const float nf = 1.0f - fabsf((float)hi - e.x) / (float)L; // iwe_formatting(.., high_pass, scale) (:345-351)
            const float nb = 1.0f - fabsf((float)lo - e.x) / (float)L;
            splat_inside_1hot_warp(img_fb + (long)slot * slot_stride, p.ig, make_float2(fw.y, fw.x), nf, pol, on, lane, key_base);
            splat_inside_1hot_warp(img_fb + (long)(slot + 1) * slot_stride, p.ig, make_float2(bw.y, bw.x), nb, pol, on, lane, key_base);
        }
        return;
    }
    if (!live) return;
    for (int s = 0; s < p.sc.S; ++s) {
        const int L = p.sc.L[s];
        if (t >= (L << s)) continue;
        const int wi = t / L, lo = wi * L, hi = lo + L;
        float2 fw, bw; float dth, dtl;
        if (!linear_ends(p, (float)lo, (float)hi, e.x, e.y, e.z, v, fw, bw, dth, dtl)) continue;
        const int slot = p.sc.slot_base[s] + wi * 2;
        const float nf = 1.0f - fabsf((float)hi - e.x) / (float)L;
        const float nb = 1.0f - fabsf((float)lo - e.x) / (float)L;
        if (p.border) {      // both ends passed purge_unfeasible: in-image fast path
            splat<true, DET>(img_fb + (long)slot * slot_stride, p.res, p.ig, fw.x, fw.y, nf, m);
            splat<true, DET>(img_fb + (long)(slot + 1) * slot_stride, p.res, p.ig, bw.x, bw.y, nb, m);
        } else {
            splat<false, DET>(img_fb + (long)slot * slot_stride, p.res, p.ig, fw.x, fw.y, nf, m);
            splat<false, DET>(img_fb + (long)(slot + 1) * slot_stride, p.res, p.ig, bw.x, bw.y, nb, m);
        }
    }
}

// taps_red<false> for the events of a whole warp whose sample lies inside the map: the two tap-row reductions of neighbouring
// lanes that hit the same slot of the same flow-gradient map are merged first.  Events sit on integer pixels and are sorted
// by pixel, so the events of one pixel (3 on average at 1 M events per window) share one reduction.  `on` = this lane has a
// gradient to reduce; must be called by all 32 lanes.
__device__ __forceinline__ void taps_red_inside_warp(float2 *__restrict__ gmap, const ImgGeom &g, const Taps &tp, float gpy, float gpx, bool on,
                                                     unsigned lane, unsigned key_base) {
    float v[8] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    unsigned off = 0;
    if (on) {
        const int phase = tp.x0 & 1;
        off = (unsigned)(phase * (int)g.plane + tp.y0 * g.Wp + tp.x0 + phase);
        // taps_red's coefficients with dt = 1 (1.0f * w == w) and a zero for the taps outside the map
        const float c0 = tp.w[0], c1 = tp.ok[1] ? tp.w[1] : 0.0f, c2 = tp.ok[2] ? tp.w[2] : 0.0f, c3 = tp.ok[3] ? tp.w[3] : 0.0f;
        v[0] = c0 * gpx; v[1] = c0 * gpy; v[2] = c1 * gpx; v[3] = c1 * gpy;
        v[4] = c2 * gpx; v[5] = c2 * gpy; v[6] = c3 * gpx; v[7] = c3 * gpy;
    }
    const bool gave = merge_equal_neighbours<8>(on ? key_base + off : (0x80000000u | lane), lane, v);
    if (on && !gave) {
        if (v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f || v[3] != 0.0f) red_add_v4(gmap + off, v[0], v[1], v[2], v[3]);
        if (v[4] != 0.0f || v[5] != 0.0f || v[6] != 0.0f || v[7] != 0.0f) red_add_v4(gmap + (off + (unsigned)g.Wp), v[4], v[5], v[6], v[7]);
    }
}

template <bool DET>
__global__ void __launch_bounds__(kThreads, 4) linear_bwd_kernel(const __grid_constant__ CmParams p) {
    int t, b, row, set; float4 e; float2 m;
    const bool live = locate_sorted(p, t, b, e, m, row, set);
    if (DET && !live) return;
    if (!live) { b = 0; e = make_float4(0.f, 0.f, 0.f, 0.f); m = make_float2(0.f, 0.f); }
    const int f = blockIdx.y;
    const long mo = (((long)f * p.P + t) * p.B + b) * 2 * p.res.fplane;
    Taps tp;
    float2 v = make_float2(0.f, 0.f);
    const bool own_inside = live && inside(e.y, e.z, p.res);
    if (live) {
        const float2 vxy = own_flow<true>(p.flow + mo, p.res, e.y, e.z, &tp);
        v = make_float2(vxy.y, vxy.x);
    }
    const long gslot = 4 * p.ig.plane;
    const float2 *img_fb = (DET ? p.gimg : p.img) + ((long)f * p.B + b) * p.nslots * gslot;
    const bool onehot = (m.x == 1.0f && m.y == 0.0f) || (m.x == 0.0f && m.y == 1.0f);
    const bool fast = !DET && p.border && (!live || onehot);               // one-hot masks, both ends inside the image
    const bool warp_fast = __all_sync(0xffffffffu, fast);
    const float2 *img_pol = img_fb + (m.x != 0.0f ? 0 : p.ig.plane);
    float gvy = 0.f, gvx = 0.f;
    if (live) {
        for (int s = 0; s < p.sc.S; ++s) {
            const int L = p.sc.L[s];
            if (t >= (L << s)) continue;
            const int wi = t / L, lo = wi * L, hi = lo + L;
            float2 fw, bw; float dth, dtl;
            if (!linear_ends(p, (float)lo, (float)hi, e.x, e.y, e.z, v, fw, bw, dth, dtl)) continue;
            const int slot = p.sc.slot_base[s] + wi * 2;
            const float nf = 1.0f - fabsf((float)hi - e.x) / (float)L;
            const float nb = 1.0f - fabsf((float)lo - e.x) / (float)L;
            float gy = 0.f, gx = 0.f;
            if (warp_fast) iwe_grad_inside_1hot(img_pol + (long)slot * gslot, p.res, p.ig, make_float2(fw.y, fw.x), nf, gy, gx);
            else iwe_grad<false>(img_fb + (long)slot * gslot, p.res, p.ig, fw.x, fw.y, nf, m, gy, gx);
            gvy += dth * gy; gvx += dth * gx;
            gy = 0.f; gx = 0.f;
            if (warp_fast) iwe_grad_inside_1hot(img_pol + (long)(slot + 1) * gslot, p.res, p.ig, make_float2(bw.y, bw.x), nb, gy, gx);
            else iwe_grad<false>(img_fb + (long)(slot + 1) * gslot, p.res, p.ig, bw.x, bw.y, nb, m, gy, gx);
            gvy += dtl * gy; gvx += dtl * gx;
        }
    }
    const bool has_grad = live && !(gvy == 0.f && gvx == 0.f);
    float2 *gmap = p.gflow + (((long)f * p.P + t) * p.B + b) * (DET ? 4 : 2) * p.ig.plane;
    if (!DET && warp_fast) {
        // samples inside the map are reduced together (merged); an event whose own location lies outside the sensor takes the generic path
        if (has_grad && !own_inside) taps_red<false>(gmap, p.ig, tp, 1.0f, gvy, gvx);
        taps_red_inside_warp(gmap, p.ig, tp, gvy, gvx, has_grad && own_inside, threadIdx.x & 31u, (unsigned)b * (unsigned)(2 * p.ig.plane));
        return;
    }
    if (!has_grad) return;
    taps_red<DET>(gmap, p.ig, tp, 1.0f, gvy, gvx, DET ? __ldg(p.den + p.F * p.B * p.nslots + 1) : 1.0f);
}

}  // namespace tef

using namespace tef;

int tef_sort_events(const CmParams &p, cudaStream_t st);           // tef_cm_sort.cu
int tef_reduce_and_finalize(const CmParams &p, cudaStream_t st);   // tef_cm_reduce.cu
int tef_grad_images(const CmParams &p, cudaStream_t st);

extern "C" int tef_linear_forward(const tef_cm_desc *d, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CmParams p;
    int rc = fill_params(d, 1, p);
    if (rc) return rc;
    if (!p.flow || !p.img || !p.acc_sum || !p.acc_nnz || !p.den || !p.loss) return TEF_EINVAL;
    cudaMemsetAsync(p.img, 0, sizeof(float2) * (long)p.F * p.B * p.nslots * (p.det ? 16 : 4) * p.ig.plane, st);   // deterministic: high and low words
    rc = tef_sort_events(p, st);
    if (rc) return rc;
    if (p.seg.blk_off[p.seg.nseg] > 0) {
        ProfScope ps(K_LIN_FWD, st);
        if (p.det) linear_fwd_kernel<true><<<dim3(p.seg.blk_off[p.seg.nseg], p.F), kThreads, 0, st>>>(p);
        else linear_fwd_kernel<false><<<dim3(p.seg.blk_off[p.seg.nseg], p.F), kThreads, 0, st>>>(p);
    }
    rc = (int)cudaGetLastError();
    if (rc) return rc;
    return tef_reduce_and_finalize(p, st);
}

extern "C" int tef_linear_backward(const tef_cm_desc *d, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CmParams p;
    int rc = fill_params(d, 1, p);
    if (rc) return rc;
    if (!p.flow || !p.gflow || !p.img || !p.den || !p.grad_out || !p.sort.bins || !p.sort.rec) return TEF_EINVAL;
    grad_segments_only(p, (long)p.B * p.nslots * 4 * p.ig.plane * 8);
    if (p.det && !p.gimg) return TEF_EINVAL;
    cudaMemsetAsync(p.gflow, 0, sizeof(float2) * (long)p.F * p.P * p.B * (p.det ? 4 : 2) * p.ig.plane, st);
    rc = tef_grad_images(p, st);
    if (rc) return rc;
    if (p.seg.blk_off[p.seg.nseg] > 0) {
        ProfScope ps(K_LIN_BWD, st);
        if (p.det) linear_bwd_kernel<true><<<dim3(p.seg.blk_off[p.seg.nseg], p.F), kThreads, 0, st>>>(p);
        else linear_bwd_kernel<false><<<dim3(p.seg.blk_off[p.seg.nseg], p.F), kThreads, 0, st>>>(p);
    }
    return (int)cudaGetLastError();
}
