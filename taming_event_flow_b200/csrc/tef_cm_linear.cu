// tef_cm_linear.cu -- fused Linear contrast-maximization loss (upstream loss/flow.py:216-412).
//
// `update` samples one flow vector per event from that pass' maps (linear_sample_kernel);
// `forward` warps every event linearly to both ends of its sub-window, applies the shared
// border mask and splats into two image slots per sub-window; the backward gathers the
// gradient images at both ends and reduces the per-event flow gradient into the packed
// flow-gradient map through the bilinear taps of the original location.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

__device__ __forceinline__ bool locate(const CmParams &p, int &sg, long &row, int &n, int &t) {
    sg = 0;
    const int blk = blockIdx.x;
    while (blk >= p.seg.blk_off[sg + 1]) ++sg;
    n = p.seg.n[sg]; t = p.seg.pass[sg];
    row = (long)(blk - p.seg.blk_off[sg]) * kThreads + threadIdx.x;
    return row < (long)p.B * n;
}

// Linear.update (:266-285): event flow (y, x) = get_event_flow(maps of this pass, raw locations)
__global__ void __launch_bounds__(kThreads) linear_sample_kernel(const __grid_constant__ CmParams p) {
    int sg, n, t; long row;
    if (!locate(p, sg, row, n, t)) return;
    const float4 e = __ldg(p.seg.ev[sg] + row);
    const int b = (int)(row / n), f = blockIdx.y;
    const long HW = (long)p.H * p.W;
    const float2 *map = p.flow + (((long)f * p.P + t) * p.B + b) * HW;
    const float2 v = sample_flow<false>(map, p.res, e.y, e.z, nullptr);
    p.seg.evflow[sg][(long)f * p.B * n + row] = make_float2(v.y, v.x);
}

// positions at both window ends (:337-343); returns the shared mask bit
__device__ __forceinline__ bool linear_ends(const CmParams &p, float lo, float hi, float ts, float y0, float x0, float2 v /* (y,x) */,
                                            float2 &fw, float2 &bw, float &dth, float &dtl) {
    dth = hi - ts; dtl = lo - ts;
    fw = make_float2(y0 + dth * v.x, x0 + dth * v.y);
    bw = make_float2(y0 + dtl * v.x, x0 + dtl * v.y);
    if (!p.border) return true;
    return inside(fw.x, fw.y, p.res) && inside(bw.x, bw.y, p.res);
}

__global__ void __launch_bounds__(kThreads) linear_fwd_kernel(const __grid_constant__ CmParams p) {
    int sg, n, t; long row;
    if (!locate(p, sg, row, n, t)) return;
    const float2 m = __ldg(p.seg.mk[sg] + row);
    if (m.x == 0.0f && m.y == 0.0f) return;
    const float4 e = __ldg(p.seg.ev[sg] + row);
    const int b = (int)(row / n), f = blockIdx.y;
    const long HW = (long)p.H * p.W;
    const float2 v = __ldg(p.seg.evflow[sg] + (long)f * p.B * n + row);
    float4 *img_fb = p.img + ((long)f * p.B + b) * p.nslots * HW;
    for (int s = 0; s < p.sc.S; ++s) {
        const int L = p.sc.L[s];
        if (t >= (L << s)) continue;
        const int wi = t / L, lo = wi * L, hi = lo + L;
        float2 fw, bw; float dth, dtl;
        if (!linear_ends(p, (float)lo, (float)hi, e.x, e.y, e.z, v, fw, bw, dth, dtl)) continue;
        const int slot = p.sc.slot_base[s] + wi * 2;
        const float nf = 1.0f - fabsf((float)hi - e.x) / (float)L;    // iwe_formatting(.., high_pass, scale) (:345-351)
        const float nb = 1.0f - fabsf((float)lo - e.x) / (float)L;
        splat(img_fb + (long)slot * HW, p.res, fw.x, fw.y, nf, m);
        splat(img_fb + (long)(slot + 1) * HW, p.res, bw.x, bw.y, nb, m);
    }
}

__global__ void __launch_bounds__(kThreads) linear_bwd_kernel(const __grid_constant__ CmParams p) {
    int sg, n, t; long row;
    if (!locate(p, sg, row, n, t)) return;
    const float2 m = __ldg(p.seg.mk[sg] + row);
    if (m.x == 0.0f && m.y == 0.0f) return;
    const float4 e = __ldg(p.seg.ev[sg] + row);
    const int b = (int)(row / n), f = blockIdx.y;
    const long HW = (long)p.H * p.W;
    const float2 v = __ldg(p.seg.evflow[sg] + (long)f * p.B * n + row);
    const float4 *img_fb = p.img + ((long)f * p.B + b) * p.nslots * HW;
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
        iwe_grad(img_fb + (long)slot * HW, p.res, fw.x, fw.y, nf, m, gy, gx);
        gvy += dth * gy; gvx += dth * gx;
        gy = 0.f; gx = 0.f;
        iwe_grad(img_fb + (long)(slot + 1) * HW, p.res, bw.x, bw.y, nb, m, gy, gx);
        gvy += dtl * gy; gvx += dtl * gx;
    }
    if (gvy == 0.f && gvx == 0.f) return;
    const long mo = (((long)f * p.P + t) * p.B + b) * HW;
    Taps tp;
    sample_flow<true>(p.flow + mo, p.res, e.y, e.z, &tp);
    float2 *g = p.gflow + mo + (long)tp.y0 * p.W + tp.x0;
    const int off[4] = { 0, 1, p.W, p.W + 1 };
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (tp.ok[k]) red_add_v2(g + off[k], tp.w[k] * gvx, tp.w[k] * gvy);
}

}  // namespace tef

using namespace tef;

int tef_reduce_and_finalize(const CmParams &p, cudaStream_t st);   // tef_cm_reduce.cu
int tef_grad_images(const CmParams &p, cudaStream_t st);

extern "C" int tef_linear_sample(const tef_cm_desc *d, int t, void *stream) {
    if (!d || t < 0 || t >= d->P) return TEF_EINVAL;
    // a one-pass view of the descriptor: only pass t has rows
    tef_cm_desc one = *d;
    one.mode = 2; one.S = 1;
    for (int k = 0; k < 2; ++k)
        for (int q = 0; q < TEF_MAX_PASSES; ++q)
            if (q != t) one.n[k][q] = 0;
    CmParams p;
    int rc = fill_params(&one, 1, false, p);
    if (rc) return rc;
    if (!p.flow) return TEF_EINVAL;
    for (int s = 0; s < p.seg.nseg; ++s) if (!p.seg.evflow[s]) return TEF_EINVAL;
    if (p.seg.blk_off[p.seg.nseg] > 0) {
        ProfScope ps(K_LIN_SAMPLE, (cudaStream_t)stream);
        linear_sample_kernel<<<dim3(p.seg.blk_off[p.seg.nseg], p.F), kThreads, 0, (cudaStream_t)stream>>>(p);
    }
    return (int)cudaGetLastError();
}

extern "C" int tef_linear_forward(const tef_cm_desc *d, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CmParams p;
    int rc = fill_params(d, 1, false, p);
    if (rc) return rc;
    if (!p.flow || !p.img || !p.acc_sum || !p.acc_nnz || !p.den || !p.loss) return TEF_EINVAL;
    for (int s = 0; s < p.seg.nseg; ++s) if (!p.seg.evflow[s]) return TEF_EINVAL;
    const long HW = (long)p.H * p.W;
    cudaMemsetAsync(p.img, 0, sizeof(float4) * (long)p.F * p.B * p.nslots * HW, st);
    if (p.seg.blk_off[p.seg.nseg] > 0) {
        ProfScope ps(K_LIN_FWD, st);
        linear_fwd_kernel<<<dim3(p.seg.blk_off[p.seg.nseg], p.F), kThreads, 0, st>>>(p);
    }
    rc = (int)cudaGetLastError();
    if (rc) return rc;
    return tef_reduce_and_finalize(p, st);
}

extern "C" int tef_linear_backward(const tef_cm_desc *d, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CmParams p;
    int rc = fill_params(d, 1, true, p);
    if (rc) return rc;
    if (!p.flow || !p.gflow || !p.img || !p.den || !p.grad_out) return TEF_EINVAL;
    for (int s = 0; s < p.seg.nseg; ++s) if (!p.seg.evflow[s]) return TEF_EINVAL;
    const long HW = (long)p.H * p.W;
    cudaMemsetAsync(p.gflow, 0, sizeof(float2) * (long)p.F * p.P * p.B * HW, st);
    rc = tef_grad_images(p, st);
    if (rc) return rc;
    if (p.seg.blk_off[p.seg.nseg] > 0) {
        ProfScope ps(K_LIN_BWD, st);
        linear_bwd_kernel<<<dim3(p.seg.blk_off[p.seg.nseg], p.F), kThreads, 0, st>>>(p);
    }
    return (int)cudaGetLastError();
}
