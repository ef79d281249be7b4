// tef_cm_common.cuh -- host/device tables shared by the fused CM-loss kernels.
#pragma once
#include "tef_device.cuh"
#include "../../include/tef_b200.h"

namespace tef {

constexpr int kThreads = 256;
constexpr int kMaxSeg = 2 * TEF_MAX_PASSES;

// One segment = one (event set, pass): [B][n] rows, handled by whole CTAs so that the
// pass index (and with it every chain predicate) is uniform inside a CTA.
struct SegTable {
    int nseg;
    int set[kMaxSeg], pass[kMaxSeg], n[kMaxSeg];
    int blk_off[kMaxSeg + 1];
    const float4 *ev[kMaxSeg];
    const float2 *mk[kMaxSeg];
    int first_bin[kMaxSeg + 1];      // first sort bin of the segment (bins are segment-major)
};

// geometry and buffers of the tile sort (tef_cm_sort.cu)
struct SortGeom {
    int B, H, W, tiles_x, tiles;     // 16x8-pixel tiles per sample
    long nbins;                      // nseg * B * tiles * 128
    int *bins;                       // [nbins + 1] histogram -> offsets -> bin ends
    int *sums;                       // scan scratch, one int per 2048 bins
    float4 *ev;                      // sorted rows (ts, y, x, sample index as int bits)
    float2 *mk;
};

// One entry per temporal scale (loss/flow.py:42-44, :434-441, :657-668)
struct ScaleTable {
    int S;
    int L[TEF_MAX_SCALES];          // window length passes_loss[s]
    int delta[TEF_MAX_SCALES];      // delta_passes[s] (Linear: = L)
    int slot_base[TEF_MAX_SCALES];  // first image slot of the scale
    int ntau[TEF_MAX_SCALES];       // reference times per sub-window (Linear: 2)
};

struct CmParams {
    int B, H, W, P, F, mode, border, loss_scaling, nslots, linear;
    Res res;
    const float2 *flow;
    float2 *gflow;
    float4 *img;
    double *acc_sum;
    int *acc_nnz;
    float *den;
    float *loss;
    const float *grad_out;
    SegTable seg;
    ScaleTable sc;
    SortGeom sort;
};

// slot -> scale / divisor tables for the reduction kernels (few entries, in constant param space)
struct SlotInfo {
    int s;         // temporal scale index
    float div_a;   // 2*delta+1 (Iterative, loss/flow.py:731) or 2 (Linear, :397)
};

inline int build_scales(const tef_cm_desc *d, int linear, ScaleTable &sc) {
    sc.S = d->S;
    int base = 0;
    for (int s = 0; s < d->S; ++s) {
        int L = d->P >> s;
        int delta = linear ? L : (d->mode == 1 ? L : (d->mode == 2 ? L / 2 : L / 4));
        int ntau = linear ? 2 : (d->mode == 4 ? 2 * delta + 1 : L + 1);
        sc.L[s] = L; sc.delta[s] = delta; sc.slot_base[s] = base; sc.ntau[s] = ntau;
        base += (1 << s) * ntau;
    }
    return base;
}

inline int check_desc(const tef_cm_desc *d, int linear) {
    if (!d) return TEF_EINVAL;
    if (d->B < 1 || d->H < 2 || d->W < 2 || d->P < 1 || d->F < 1 || d->S < 1) return TEF_EINVAL;
    if (d->P > TEF_MAX_PASSES || d->S > TEF_MAX_SCALES || d->F > TEF_MAX_FLOWS) return TEF_ELIMIT;
    if (!linear && d->mode != 1 && d->mode != 2 && d->mode != 4) return TEF_EINVAL;
    for (int s = 0; s < d->S; ++s) {
        int L = d->P >> s;
        int delta = linear ? L : (d->mode == 1 ? L : (d->mode == 2 ? L / 2 : L / 4));
        if (L == 0 || delta == 0) return TEF_EEMPTY;       // torch.cat([]) in the reference
    }
    if (!linear && d->mode == 4 && d->border_comp) return TEF_EMODE4;
    return 0;
}

inline int fill_params(const tef_cm_desc *d, int linear, CmParams &p) {
    int rc = check_desc(d, linear);
    if (rc) return rc;
    p.B = d->B; p.H = d->H; p.W = d->W; p.P = d->P; p.F = d->F; p.mode = d->mode;
    p.border = d->border_comp; p.loss_scaling = d->loss_scaling; p.linear = linear;
    p.res = Res::make(d->H, d->W);
    p.flow = (const float2 *)d->flow; p.gflow = (float2 *)d->gflow; p.img = (float4 *)d->img;
    p.acc_sum = d->acc_sum; p.acc_nnz = d->acc_nnz; p.den = d->den; p.loss = d->loss; p.grad_out = d->grad_out;
    p.nslots = build_scales(d, linear, p.sc);
    int ns = 0, blk = 0;
    for (int set = 0; set < 2; ++set)
        for (int t = 0; t < d->P; ++t) {
            if (d->n[set][t] <= 0) continue;
            if (!d->ev[set][t] || !d->mk[set][t]) return TEF_EINVAL;
            p.seg.set[ns] = set; p.seg.pass[ns] = t; p.seg.n[ns] = d->n[set][t];
            p.seg.ev[ns] = (const float4 *)d->ev[set][t]; p.seg.mk[ns] = (const float2 *)d->mk[set][t];
            p.seg.blk_off[ns] = blk;
            long rows = (long)d->B * d->n[set][t];
            blk += (int)((rows + kThreads - 1) / kThreads);
            ++ns;
        }
    p.seg.nseg = ns; p.seg.blk_off[ns] = blk;
    SortGeom &g = p.sort;
    g.B = d->B; g.H = d->H; g.W = d->W;
    g.tiles_x = (d->W + 15) / 16; g.tiles = g.tiles_x * ((d->H + 7) / 8);
    for (int s = 0; s <= ns; ++s) p.seg.first_bin[s] = s * d->B * g.tiles * 128;
    g.nbins = (long)ns * d->B * g.tiles * 128;
    g.bins = (int *)d->sort_bins; g.sums = (int *)d->sort_sums;
    g.ev = (float4 *)d->sorted_ev; g.mk = (float2 *)d->sorted_mk;
    return 0;
}

// the backward only visits the gradient-carrying set; its segments come first
inline void grad_segments_only(CmParams &p) {
    int ng = 0;
    while (ng < p.seg.nseg && p.seg.set[ng] == 0) ++ng;
    p.seg.nseg = ng;
}

// rows of segment sg in the sorted arrays: [lo, hi)
__device__ __forceinline__ void seg_rows(const CmParams &p, int sg, int &lo, int &hi) {
    const int fb = p.seg.first_bin[sg];
    lo = fb ? __ldg(p.sort.bins + fb - 1) : 0;
    hi = __ldg(p.sort.bins + p.seg.first_bin[sg + 1] - 1);
}

// CTA -> segment, thread -> sorted row; false when the thread has no event
__device__ __forceinline__ bool locate_sorted(const CmParams &p, int &t, int &b, float4 &e, float2 &m) {
    int sg = 0;
    const int blk = blockIdx.x;
    while (blk >= p.seg.blk_off[sg + 1]) ++sg;
    t = p.seg.pass[sg];
    int lo, hi;
    seg_rows(p, sg, lo, hi);
    const int row = lo + (blk - p.seg.blk_off[sg]) * kThreads + threadIdx.x;
    if (row >= hi) return false;
    e = __ldg(p.sort.ev + row);
    m = __ldg(p.sort.mk + row);
    b = __float_as_int(e.w);
    return true;
}

// upstream gradient of one focus_loss value, divided in the order autograd unwinds
// loss/flow.py:730-736 (Linear :396-402)
__host__ __device__ inline float upstream(float gout, int F, int S, float div_a, int s) {
    float g = gout;
    g = g / (float)F;
    g = g / (float)S;
    g = g / div_a;
    g = g / (float)(1 << s);
    return g;
}

// iwe_formatting (loss/flow.py:81-110) for one event: 4 corners x (count, time-weighted) of its polarity
__device__ __forceinline__ void splat(float4 *__restrict__ im, const Res &r, float y, float x, float nts, float2 m) {
    Corners c;
    corners(y, x, r, c);
#pragma unroll
    for (int ky = 0; ky < 2; ++ky)
#pragma unroll
        for (int kx = 0; kx < 2; ++kx) {
            if (!(c.oky[ky] && c.okx[kx])) continue;
            const float w = c.wy[ky] * c.wx[kx];
            if (w == 0.0f) continue;
            const float wt = w * nts;
            float2 *dst = reinterpret_cast<float2 *>(im + ((long)c.cy[ky] * r.W + (long)c.cx[kx]));
            if (m.x != 0.0f) red_add_v2(dst, w * m.x, wt * m.x);
            if (m.y != 0.0f) red_add_v2(dst + 1, w * m.y, wt * m.y);
        }
}

// gradient of the loss w.r.t. the position at one reference time, through the bilinear
// splat weights; im holds the gradient images.
__device__ __forceinline__ void iwe_grad(const float4 *__restrict__ im, const Res &r, float y, float x, float nts, float2 m,
                                         float &gy, float &gx) {
    Corners c;
    corners(y, x, r, c);
#pragma unroll
    for (int ky = 0; ky < 2; ++ky)
#pragma unroll
        for (int kx = 0; kx < 2; ++kx) {
            if (!(c.oky[ky] && c.okx[kx])) continue;
            const float4 g = __ldg(im + ((long)c.cy[ky] * r.W + (long)c.cx[kx]));
            const float gw = m.x * (g.x + nts * g.y) + m.y * (g.z + nts * g.w);
            gy += gw * d1(y, c.cy[ky]) * c.wx[kx];
            gx += gw * c.wy[ky] * d1(x, c.cx[kx]);
        }
}

}  // namespace tef
