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

template <bool DET>
__global__ void __launch_bounds__(kThreads) linear_fwd_kernel(const __grid_constant__ CmParams p) {
    int t, b, row, set; float4 e; float2 m;
    if (!locate_sorted(p, t, b, e, m, row, set)) return;
    const int f = blockIdx.y;
    const float2 vxy = sample_flow<false>(p.flow + (((long)f * p.P + t) * p.B + b) * 2 * p.res.fplane, p.res, e.y, e.z, nullptr);
    const float2 v = make_float2(vxy.y, vxy.x);                        // (y, x), utils/iwe.py:38
    const long slot_stride = (DET ? 8 : 4) * p.ig.plane;
    float2 *img_fb = p.img + ((long)f * p.B + b) * p.nslots * slot_stride;
    for (int s = 0; s < p.sc.S; ++s) {
        const int L = p.sc.L[s];
        if (t >= (L << s)) continue;
        const int wi = t / L, lo = wi * L, hi = lo + L;
        float2 fw, bw; float dth, dtl;
        if (!linear_ends(p, (float)lo, (float)hi, e.x, e.y, e.z, v, fw, bw, dth, dtl)) continue;
        const int slot = p.sc.slot_base[s] + wi * 2;
        const float nf = 1.0f - fabsf((float)hi - e.x) / (float)L;    // iwe_formatting(.., high_pass, scale) (:345-351)
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

template <bool DET>
__global__ void __launch_bounds__(kThreads, 4) linear_bwd_kernel(const __grid_constant__ CmParams p) {
    int t, b, row, set; float4 e; float2 m;
    if (!locate_sorted(p, t, b, e, m, row, set)) return;
    const int f = blockIdx.y;
    const long mo = (((long)f * p.P + t) * p.B + b) * 2 * p.res.fplane;
    Taps tp;
    const float2 vxy = sample_flow<true>(p.flow + mo, p.res, e.y, e.z, &tp);
    const float2 v = make_float2(vxy.y, vxy.x);
    const long gslot = 4 * p.ig.plane;
    const float2 *img_fb = (DET ? p.gimg : p.img) + ((long)f * p.B + b) * p.nslots * gslot;
    float gvy = 0.f, gvx = 0.f;
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
        iwe_grad<false>(img_fb + (long)slot * gslot, p.res, p.ig, fw.x, fw.y, nf, m, gy, gx);
        gvy += dth * gy; gvx += dth * gx;
        gy = 0.f; gx = 0.f;
        iwe_grad<false>(img_fb + (long)(slot + 1) * gslot, p.res, p.ig, bw.x, bw.y, nb, m, gy, gx);
        gvy += dtl * gy; gvx += dtl * gx;
    }
    if (gvy == 0.f && gvx == 0.f) return;
    taps_red<DET>(p.gflow + (((long)f * p.P + t) * p.B + b) * (DET ? 4 : 2) * p.ig.plane, p.ig, tp, 1.0f, gvy, gvx,
                  DET ? __ldg(p.den + p.F * p.B * p.nslots + 1) : 1.0f);
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
