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
__device__ __forceinline__ float2 pixel_sum(const float2 *__restrict__ slot_base, const ImgGeom &g, int pol, long o) {
    const float2 a = slot_base[(long)pol * g.plane + o];                 // phase 0: column x
    const float2 b = slot_base[(long)(2 + pol) * g.plane + o + 1];       // phase 1: column x + 1
    return make_float2(a.x + b.x, a.y + b.y);
}

__global__ void __launch_bounds__(kThreads) iwe_reduce_kernel(const float2 *__restrict__ img, double *__restrict__ acc_sum,
                                                              int *__restrict__ acc_nnz, int W, long HW, ImgGeom g) {
    const long image = blockIdx.y;
    const float2 *im = img + image * 4 * g.plane;
    const long p0 = (long)blockIdx.x * kPixPerBlock;
    const long p1 = min(p0 + kPixPerBlock, HW);
    double acc = 0.0;
    int cnt = 0;
    for (long i = p0 + threadIdx.x; i < p1; i += kThreads) {
        const long o = (i / W) * g.Wp + (i % W);
        const float2 vp = pixel_sum(im, g, 0, o), vn = pixel_sum(im, g, 1, o);
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
        atomicAdd(acc_sum + image, a);
        atomicAdd(acc_nnz + image, c);
    }
}

// one CTA: den = nnz + 1e-9 (:127), loss = sum over flow maps, scales, windows, trefs, samples
__global__ void __launch_bounds__(kThreads) finalize_kernel(const __grid_constant__ CmParams p) {
    const int nimg = p.F * p.B * p.nslots;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nimg; i += kThreads) {
        const int q = i % p.nslots;
        const int s = scale_of_slot(p.sc, q);
        const float den = p.loss_scaling ? ((float)p.acc_nnz[i] + 1e-9f) : 1.0f;
        p.den[i] = den;
        const double div_a = p.linear ? 2.0 : (double)(2 * p.sc.delta[s] + 1);
        acc += p.acc_sum[i] / (double)den / (double)(1 << s) / div_a / (double)p.sc.S / (double)p.F;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double s_acc[kThreads / 32];
    if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) a += s_acc[k];
        *p.loss = (float)a;
    }
}

// in place: (cnt, ts) -> (dL/dcnt, dL/dts) per polarity, in autograd's operation order
// (pow: grad*(2*A); div: grad/D and -grad*((T/D)/D)), see oracle/cm_oracle_impl.h phase 3.
__global__ void __launch_bounds__(kThreads) iwe_grad_kernel(const __grid_constant__ CmParams p, long HW) {
    const long image = blockIdx.y;
    const int q = (int)(image % p.nslots);
    const int s = scale_of_slot(p.sc, q);
    const float div_a = p.linear ? 2.0f : (float)(2 * p.sc.delta[s] + 1);
    const float cf = upstream(__ldg(p.grad_out), p.F, p.sc.S, div_a, s) / p.den[image];
    float2 *im = p.img + image * 4 * p.ig.plane;
    const long p0 = (long)blockIdx.x * kPixPerBlock;
    const long p1 = min(p0 + kPixPerBlock, HW);
    for (long i = p0 + threadIdx.x; i < p1; i += kThreads) {
        const long o = (i / p.W) * p.ig.Wp + (i % p.W);
#pragma unroll
        for (int pol = 0; pol < 2; ++pol) {
            const float2 v = pixel_sum(im, p.ig, pol, o);
            const float d = v.x + 1e-9f;
            const float a = v.y / d;
            const float ga = cf * (2.0f * a);
            im[(long)pol * p.ig.plane + o] = make_float2(-(ga * (a / d)), ga / d);    // (dL/dcount, dL/dtime-weighted), phase-0 plane
        }
    }
}

}  // namespace tef

using namespace tef;

int tef_reduce_and_finalize(const CmParams &p, cudaStream_t st) {
    const long HW = (long)p.H * p.W;
    const int nimg = p.F * p.B * p.nslots;
    cudaMemsetAsync(p.acc_sum, 0, sizeof(double) * nimg, st);
    cudaMemsetAsync(p.acc_nnz, 0, sizeof(int) * nimg, st);
    dim3 grid((unsigned)((HW + kPixPerBlock - 1) / kPixPerBlock), nimg);
    { ProfScope ps(K_IWE_REDUCE, st); iwe_reduce_kernel<<<grid, kThreads, 0, st>>>(p.img, p.acc_sum, p.acc_nnz, p.W, HW, p.ig); }
    { ProfScope ps(K_FINALIZE, st); finalize_kernel<<<1, kThreads, 0, st>>>(p); }
    return (int)cudaGetLastError();
}

int tef_grad_images(const CmParams &p, cudaStream_t st) {
    const long HW = (long)p.H * p.W;
    const int nimg = p.F * p.B * p.nslots;
    dim3 grid((unsigned)((HW + kPixPerBlock - 1) / kPixPerBlock), nimg);
    ProfScope ps(K_IWE_GRAD, st);
    iwe_grad_kernel<<<grid, kThreads, 0, st>>>(p, HW);
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
    out[1] = (long)p.F * p.B * p.nslots * 4 * p.ig.plane * 2;     // floats in img
    out[2] = (long)p.F * p.P * p.B * 2 * p.ig.plane * 2;          // floats in gflow
    out[3] = p.sort.nbins + 1;                                    // ints in sort_bins
    out[4] = p.sort.nbins / 2048 + 2;                             // ints in sort_sums
    out[5] = r;                                                   // rows of sorted_ev / sorted_mk
    out[6] = p.rows_grad;                                         // gradient-carrying rows
    out[7] = linear ? 0 : (long)p.F * (p.P + 1) * p.rows_grad * 2; // floats in posbuf (alivebuf: F * rows_grad u64)
    out[8] = p.ig.Wp;
    return 0;
}
