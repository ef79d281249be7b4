// tef_cm_reduce.cu -- image-side kernels of the fused CM loss (shared by Iterative and Linear):
// per-pixel normalisation + sum of squares + non-zero count (focus_loss, upstream
// loss/flow.py:112-129 with the division of :727 / :387,:393 fused in), the scalar
// finalisation (:730-736 / :396-402) and the in-place gradient images for the backward.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

constexpr int kPixPerBlock = 2048;

__device__ __forceinline__ int scale_of_slot(const ScaleTable &sc, int q) {
    int s = 0;
    while (s + 1 < sc.S && q >= sc.slot_base[s + 1]) ++s;
    return s;
}

// sum of the two phases of one pixel: (count, time-weighted count) of one polarity
template <bool DET>
__device__ __forceinline__ float2 pixel_sum(const float2 *__restrict__ slot_base, const ImgGeom &g, int pol, long o) {
    if (!DET) {
        const float2 a = slot_base[(long)pol * g.plane + o];                 // phase 0: column x
        const float2 b = slot_base[(long)(2 + pol) * g.plane + o + 1];       // phase 1: column x + 1
        return make_float2(a.x + b.x, a.y + b.y);
    }
    const longlong2 *q = reinterpret_cast<const longlong2 *>(slot_base), *ql = q + g.lo_off / 2;
    const longlong2 a = q[(long)pol * g.plane + o], b = q[(long)(2 + pol) * g.plane + o + 1];
    const longlong2 al = ql[(long)pol * g.plane + o], bl = ql[(long)(2 + pol) * g.plane + o + 1];
    return make_float2(from_fix2(a.x + b.x, al.x + bl.x), from_fix2(a.y + b.y, al.y + bl.y));     // exact integer sums, one rounding
}

// per-CTA partial sums in a fixed slot (no atomics): the final per-image sum has a fixed order in both modes
template <bool DET>
__global__ void __launch_bounds__(kThreads) iwe_reduce_kernel(const float2 *__restrict__ img, double *__restrict__ part_sum,
                                                              int *__restrict__ part_nnz, int W, long HW, ImgGeom g) {
    const long image = blockIdx.y;
    const float2 *im = img + image * (DET ? 8 : 4) * g.plane;
    const long p0 = (long)blockIdx.x * kPixPerBlock;
    const long p1 = min(p0 + kPixPerBlock, HW);
    double acc = 0.0;
    int cnt = 0;
    for (long i = p0 + threadIdx.x; i < p1; i += kThreads) {
        const long o = (i / W) * g.Wp + (i % W);
        const float2 vp = pixel_sum<DET>(im, g, 0, o), vn = pixel_sum<DET>(im, g, 1, o);
        const float ap = vp.y / (vp.x + 1e-9f);       // loss/flow.py:727
        const float an = vn.y / (vn.x + 1e-9f);
        acc += (double)(ap * ap) + (double)(an * an); // :123
        cnt += ((vp.x + vn.x) != 0.0f);               // :125
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ double s_acc[kThreads / 32];
    __shared__ int s_cnt[kThreads / 32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_acc[wid] = acc; s_cnt[wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0; int c = 0;
        for (int k = 0; k < kThreads / 32; ++k) { a += s_acc[k]; c += s_cnt[k]; }
        part_sum[image * gridDim.x + blockIdx.x] = a;
        part_nnz[image * gridDim.x + blockIdx.x] = c;
    }
}

// One CTA.  den = nnz + 1e-9 (:127), loss = sum over flow maps, scales, windows, trefs, samples.  A group of G lanes
// (G = 32 for large images with many partial sums, 1 for small ones) owns one image; every sum has a fixed order
// (lane-strided partials, shuffle tree, groups in order): bit-reproducible.
template <int G>
__global__ void __launch_bounds__(kThreads) finalize_kernel(const __grid_constant__ CmParams p) {
    const int nimg = p.F * p.B * p.nslots;
    const int lane = threadIdx.x % G, grp = threadIdx.x / G;
    double acc = 0.0;                                   // lane 0 of each group accumulates its images
    for (int i = grp; i < nimg; i += kThreads / G) {
        double sum = 0.0; int nnz = 0;
        for (int c = lane; c < p.nchunks; c += G) { sum += p.acc_sum[(long)i * p.nchunks + c]; nnz += p.acc_nnz[(long)i * p.nchunks + c]; }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); nnz += __shfl_xor_sync(0xffffffffu, nnz, o); }
        if (lane == 0) {
            const int q = i % p.nslots;
            const int s = scale_of_slot(p.sc, q);
            const float den = p.loss_scaling ? ((float)nnz + 1e-9f) : 1.0f;
            p.den[i] = den;
            const double div_a = p.linear ? 2.0 : (double)(2 * p.sc.delta[s] + 1);
            acc += sum / (double)den / (double)(1 << s) / div_a / (double)p.sc.S / (double)p.F;
        }
    }
    __shared__ double s_acc[kThreads];
    s_acc[threadIdx.x] = (lane == 0) ? acc : 0.0;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {        // fixed-order tree
        if (threadIdx.x < o) s_acc[threadIdx.x] += s_acc[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *p.loss = (float)s_acc[0];
}

// in place: (cnt, ts) -> (dL/dcnt, dL/dts) per polarity, in autograd's operation order
// (pow: grad*(2*A); div: grad/D and -grad*((T/D)/D)), see oracle/cm_oracle_impl.h phase 3.
template <bool DET>
__global__ void __launch_bounds__(kThreads) iwe_grad_kernel(const __grid_constant__ CmParams p, long HW) {
    const long image = blockIdx.y;
    const int q = (int)(image % p.nslots);
    const int s = scale_of_slot(p.sc, q);
    const float div_a = p.linear ? 2.0f : (float)(2 * p.sc.delta[s] + 1);
    const float cf = upstream(__ldg(p.grad_out), p.F, p.sc.S, div_a, s) / p.den[image];
    float2 *im = p.img + image * (DET ? 8 : 4) * p.ig.plane;
    float2 *out = DET ? p.gimg + image * 4 * p.ig.plane : im;     // [phase][pol][H][Wp]; in place unless deterministic
    float2 *outq = (!DET && p.gimgq) ? reinterpret_cast<float2 *>(p.gimgq) + image * 2 * 16 * p.res.cplane : nullptr;   // [pol][4][cplane] cells of 4 float2
    const long p0 = (long)blockIdx.x * kPixPerBlock;
    const long p1 = min(p0 + kPixPerBlock, HW);
    for (long i = p0 + threadIdx.x; i < p1; i += kThreads) {
        const long o = (i / p.W) * p.ig.Wp + (i % p.W);
#pragma unroll
        for (int pol = 0; pol < 2; ++pol) {
            const float2 v = pixel_sum<DET>(im, p.ig, pol, o);
            const float d = v.x + 1e-9f;
            const float a = v.y / d;
            const float ga = cf * (2.0f * a);
            const float2 gv = make_float2(-(ga * (a / d)), ga / d);                    // (dL/dcount, dL/dtime-weighted)
            out[(long)pol * p.ig.plane + o] = gv;                                      // phase 0: column x
            out[(long)(2 + pol) * p.ig.plane + o + 1] = gv;                            // phase 1: column x + 1
            if (!DET && outq) {
                // quad-cell copy: the pixel is the (ry, rx) entry of one cell per parity (tef_device.cuh, quad_cell)
                const int y = (int)(i / p.W), x = (int)(i % p.W);
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) {
                    const int yy = y + (ph >> 1), xx = x + (ph & 1);
                    const long cell = (long)(pol * 4 + ph) * p.res.cplane + (yy >> 1) * p.res.CX + (xx >> 1);
                    outq[cell * 4 + (yy & 1) * 2 + (xx & 1)] = gv;
                }
            }
        }
    }
}

// Deterministic mode: power-of-two scale of the fixed-point flow-gradient words = the largest per-image factor
// |upstream gradient / normaliser| rounded up, so that the normalised addends are O(weights x events per pixel) whatever the
// caller multiplies the loss by.  den[nimg] = scale, den[nimg + 1] = 1 / scale (both exact powers of two).
__global__ void det_scale_kernel(const __grid_constant__ CmParams p) {
    const int nimg = p.F * p.B * p.nslots;
    float m = 0.0f;
    for (int i = threadIdx.x; i < nimg; i += 32) {
        const int s = scale_of_slot(p.sc, i % p.nslots);
        const float div_a = p.linear ? 2.0f : (float)(2 * p.sc.delta[s] + 1);
        m = fmaxf(m, fabsf(upstream(__ldg(p.grad_out), p.F, p.sc.S, div_a, s) / p.den[i]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) {
        int e = 0;
        if (m > 0.0f && m < 3.0e38f) frexpf(m, &e);                 // m = f * 2^e, f in [0.5, 1)
        e = max(-100, min(100, e));
        p.den[nimg] = ldexpf(1.0f, e);
        p.den[nimg + 1] = ldexpf(1.0f, -e);
    }
}

}  // namespace tef

using namespace tef;

int tef_reduce_and_finalize(const CmParams &p, cudaStream_t st) {
    const long HW = (long)p.H * p.W;
    const int nimg = p.F * p.B * p.nslots;
    dim3 grid((unsigned)p.nchunks, nimg);
    {
        ProfScope ps(K_IWE_REDUCE, st);
        if (p.det) iwe_reduce_kernel<true><<<grid, kThreads, 0, st>>>(p.img, p.acc_sum, p.acc_nnz, p.W, HW, p.ig);
        else iwe_reduce_kernel<false><<<grid, kThreads, 0, st>>>(p.img, p.acc_sum, p.acc_nnz, p.W, HW, p.ig);
    }
    {
        ProfScope ps(K_FINALIZE, st);
        if (p.nchunks >= 32) finalize_kernel<32><<<1, kThreads, 0, st>>>(p);
        else finalize_kernel<1><<<1, kThreads, 0, st>>>(p);
    }
    return (int)cudaGetLastError();
}

int tef_grad_images(const CmParams &p, cudaStream_t st) {
    const long HW = (long)p.H * p.W;
    const int nimg = p.F * p.B * p.nslots;
    dim3 grid((unsigned)p.nchunks, nimg);
    ProfScope ps(K_IWE_GRAD, st);
    if (p.det) det_scale_kernel<<<1, 32, 0, st>>>(p);
    if (p.det) iwe_grad_kernel<true><<<grid, kThreads, 0, st>>>(p, HW);
    else iwe_grad_kernel<false><<<grid, kThreads, 0, st>>>(p, HW);
    return (int)cudaGetLastError();
}

extern "C" int tef_cm_num_slots(const tef_cm_desc *d, int linear) {
    int rc = check_desc(d, linear);
    if (rc) return rc;
    ScaleTable sc;
    return build_scales(d, linear, sc);
}

extern "C" int tef_cm_sizes(const tef_cm_desc *d, int linear, long *out) {
    CmParams p;
    int rc = fill_params(d, linear, p);
    if (rc) return rc;
    if (!out) return TEF_EINVAL;
    long r = 0;
    for (int s = 0; s < p.seg.nseg; ++s) r += (long)p.B * p.seg.n[s];
    out[0] = p.nslots;
    const long wide = p.det ? 2 : 1;                              // deterministic mode: int64 instead of float
    out[1] = (long)p.F * p.B * p.nslots * 4 * p.ig.plane * 2 * wide * wide;   // floats in img (deterministic: a high and a low int64 word per value)
    out[2] = (long)p.F * p.P * p.B * 2 * p.ig.plane * 2 * wide;        // floats in gflow
    out[3] = p.sort.nbins + 1;                                    // ints in sort_bins
    out[4] = p.sort.nbins / 2048 + 2;                             // ints in sort_sums
    out[5] = r;                                                   // rows of sorted_ev / sorted_mk
    out[6] = p.rows_grad;                                         // gradient-carrying rows
    out[7] = linear ? 0 : (long)p.F * (p.P + 1) * p.rows_grad * 2; // floats in posbuf (alivebuf: F * rows_grad u64)
    out[8] = p.ig.Wp;
    out[9] = p.nchunks;                                           // partial sums per image: acc_sum / acc_nnz hold F*B*slots*nchunks entries
    out[10] = p.det ? (long)p.F * p.B * p.nslots * 4 * p.ig.plane * 2 : 0;   // floats in gimg (deterministic mode)
    const bool quad_ok = !linear && !p.det;
    out[11] = quad_ok ? (long)p.F * p.B * p.nslots * 2 * 4 * p.res.cplane * 8 : 0;   // floats in gimgq
    out[12] = quad_ok ? (long)p.F * p.P * p.B * 4 * p.res.cplane * 8 : 0;            // floats in flowq
    return 0;
}
