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

__global__ void __launch_bounds__(kThreads) iwe_reduce_kernel(const float4 *__restrict__ img, double *__restrict__ acc_sum,
                                                              int *__restrict__ acc_nnz, long HW) {
    const long image = blockIdx.y;
    const float4 *im = img + image * HW;
    const long p0 = (long)blockIdx.x * kPixPerBlock;
    const long p1 = min(p0 + kPixPerBlock, HW);
    double acc = 0.0;
    int cnt = 0;
    for (long i = p0 + threadIdx.x; i < p1; i += kThreads) {
        const float4 v = im[i];                       // cnt+, ts+, cnt-, ts-
        const float ap = v.y / (v.x + 1e-9f);         // loss/flow.py:727
        const float an = v.w / (v.z + 1e-9f);
        acc += (double)(ap * ap) + (double)(an * an); // :123
        cnt += ((v.x + v.z) != 0.0f);                 // :125
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
    float4 *im = p.img + image * HW;
    const long p0 = (long)blockIdx.x * kPixPerBlock;
    const long p1 = min(p0 + kPixPerBlock, HW);
    for (long i = p0 + threadIdx.x; i < p1; i += kThreads) {
        const float4 v = im[i];
        const float dp = v.x + 1e-9f, dn = v.z + 1e-9f;
        const float ap = v.y / dp, an = v.w / dn;
        const float gap = cf * (2.0f * ap), gan = cf * (2.0f * an);
        float4 g;
        g.y = gap / dp; g.x = -(gap * (ap / dp));
        g.w = gan / dn; g.z = -(gan * (an / dn));
        im[i] = g;
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
    { ProfScope ps(K_IWE_REDUCE, st); iwe_reduce_kernel<<<grid, kThreads, 0, st>>>(p.img, p.acc_sum, p.acc_nnz, HW); }
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

extern "C" int tef_cm_sort_workspace(const tef_cm_desc *d, int linear, long *nbins, long *nsums, long *rows) {
    CmParams p;
    int rc = fill_params(d, linear, p);
    if (rc) return rc;
    long r = 0;
    for (int s = 0; s < p.seg.nseg; ++s) r += (long)p.B * p.seg.n[s];
    if (nbins) *nbins = p.sort.nbins + 1;
    if (nsums) *nsums = p.sort.nbins / 2048 + 2;
    if (rows) *rows = r;
    return 0;
}
