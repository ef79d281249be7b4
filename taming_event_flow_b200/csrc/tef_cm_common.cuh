// tef_cm_common.cuh -- host/device tables shared by the fused CM-loss kernels.
#pragma once
#include <cstdlib>
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
    int nbands, band_ctas[kMaxSeg];  // band-major CTA order of the event kernels (see locate_sorted)
    int band_off[33];
    const float4 *ev[kMaxSeg];
    const float2 *mk[kMaxSeg];
    int first_bin[kMaxSeg + 1];      // first sort bin of the segment (bins are segment-major)
};

// geometry and buffers of the tile sort (tef_cm_sort.cu)
struct SortGeom {
    int blk_off[kMaxSeg + 1];        // CTA ranges of the sort kernels (4 rows per thread)
    int B, H, W, tiles_x, tiles;     // 16x8-pixel tiles per sample
    long nbins;                      // nseg * B * tiles * kBinsPerTile
    int *bins;                       // [nbins + 1] histogram -> offsets -> bin ends
    int *sums;                       // scan scratch, one int per scan chunk (4096 bins; sized for 2048 by tef_cm_workspace)
    float4 *rec;                     // sorted rows, 32 B each: (ts, y, x, sample index bits | mask+, mask-, 0, 0)
};

// Sort key of one event inside its segment: (sample, 16x8-pixel tile, pixel, polarity).  `first_bin` is the segment's first bin.
// Polarity is part of the key because the images are polarity-planar: two events of one pixel reduce into the same slots (and
// gather the same gradient-image sectors) only if they have the same polarity, so grouping them doubles what the
// neighbour-lane merge of the event kernels finds (DESIGN.md decision 20).
constexpr int kBinsPerTile = 256;    // 16 x 8 pixels x 2 polarities
__device__ __forceinline__ int sort_bin(int first_bin, int tiles_x, int tiles, int H, int W, int b, float y, float x, int neg) {
    const int iy = (int)fminf(fmaxf(floorf(y), 0.0f), (float)(H - 1));
    const int ix = (int)fminf(fmaxf(floorf(x), 0.0f), (float)(W - 1));
    const int tile = (iy >> 3) * tiles_x + (ix >> 4);
#ifdef TEF_SORT_NO_POL
    neg = 0;                          // A/B build: round 1's key (sample, tile, pixel)
#endif
    return first_bin + (b * tiles + tile) * kBinsPerTile + ((((iy & 7) << 4) + (ix & 15)) << 1) + neg;
}

// One entry per temporal scale (loss/flow.py:42-44, :434-441, :657-668)
struct ScaleTable {
    int S;
    int L[TEF_MAX_SCALES];          // window length passes_loss[s]
    int delta[TEF_MAX_SCALES];      // delta_passes[s] (Linear: = L)
    int slot_base[TEF_MAX_SCALES];  // first image slot of the scale
    int ntau[TEF_MAX_SCALES];       // reference times per sub-window (Linear: 2)
};

// Accumulation images are polarity-planar float2 (count, time-weighted count) with rows padded to Wp and
// stored twice ("phases"): phase 0 holds pixel x at column x, phase 1 at column x+1.  A horizontally adjacent
// corner pair (x0, x0+1) is therefore always one 16-byte aligned float4 in the phase x0&1, and goes out as a
// single red.global.add.v4.f32 -- half the reduction lane-ops of per-corner updates.  The two phases are
// summed when the image is read.  Layout: [F][B][slot][phase][pol][H][Wp] float2.
// The packed flow gradient uses the same trick: [F][P][B][phase][H][Wp] float2 (d/dx-flow, d/dy-flow).
struct ImgGeom {
    int Wp;          // padded row length (even, >= W + 2)
    long plane;      // H * Wp
    long lo_off;     // deterministic mode: distance (in int64 words) from a value's high word to its low word (see to_fix2)
};

// Deterministic mode: every accumulation (images, flow-gradient maps) is done in 64-bit fixed point with integer reductions,
// which are associative, so results are bit-reproducible run to run and independent of the event order; the per-image
// reductions use fixed-order partial sums in both modes.
//  * Images: TWO words per value.  high = v rounded to 2^-40, low = the remainder (|.| <= 2^-41, exact in double) at 2^-88.
//    Every fp32 weight (>= 2^-48, 24-bit mantissa) is represented EXACTLY, so the sum is the exact sum of the reference's
//    addends, rounded to fp32 once: pixels that only ever receive tiny corner weights (1e-10 happens: the diagonal neighbour
//    of an event displaced by 1e-5 px) keep full relative precision, which a single 2^-40 word does not give them -- and
//    those are the pixels with the largest gradients (dL/dIWE ~ 1/IWE).  The low word is non-zero only for weights below
//    2^-16, so it costs a reduction only then.
//  * Flow-gradient maps: one word at 2^-40 RELATIVE to a power-of-two scale derived from the upstream gradient and the
//    per-image normalisers (det_scale_kernel), so the resolution does not depend on how the caller scales the loss.
constexpr double kFixScale = 1099511627776.0;      // 2^40
constexpr double kFixLoScale = 309485009821345068724781056.0;      // 2^88
__device__ __forceinline__ long long to_fix(float v) { return __double2ll_rn((double)v * kFixScale); }
__device__ __forceinline__ float from_fix(long long v) { return (float)((double)v * (1.0 / kFixScale)); }
__device__ __forceinline__ void to_fix2(float v, long long &hi, long long &lo) {
    hi = __double2ll_rn((double)v * kFixScale);
    lo = __double2ll_rn(((double)v - (double)hi * (1.0 / kFixScale)) * kFixLoScale);
}
__device__ __forceinline__ float from_fix2(long long hi, long long lo) {
    return (float)((double)hi * (1.0 / kFixScale) + (double)lo * (1.0 / kFixLoScale));
}
__device__ __forceinline__ void red_add_i64(long long *addr, long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

struct CmParams {
    int B, H, W, P, F, mode, border, loss_scaling, nslots, linear, det, nchunks, hist_done;
    Res res;
    ImgGeom ig;
    const float2 *flow;
    float2 *gflow;
    float2 *img;             // deterministic mode: same layout with every float replaced by an int64 (twice the bytes)
    float2 *gimg;            // deterministic mode: gradient images [F][B][slot][phase][pol][H][Wp] float2 (otherwise in place in img)
    const float4 *flowq;     // quad-cell copy of the flow maps [F][P][B][4][cplane] cells (2 float4 each) or nullptr
    float4 *gimgq;           // quad-cell copy of the gradient images [F][B][slot][pol][4][cplane] cells or nullptr (written by iwe_grad_kernel)
    float2 *posbuf;          // [(P+1)][rows_grad] chain positions (x, y) of the gradient-carrying rows (Iterative)
    uint32_t *alivebuf;      // [F][rows_grad] cumulative in-image bits (bit tref)
    long rows_grad;
    double *acc_sum;
    int *acc_nnz;
    float *den;
    float *loss;
    const float *grad_out;
    SegTable seg;
    ScaleTable sc;
    SortGeom sort;
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

// Band-major CTA order.  Every segment is tile-sorted, so the k-th fraction of each segment covers about the same
// image region.  Running band k of ALL segments before band k+1 of any keeps the slot images, gradient images and
// flow-gradient maps of that region in L2 while every pass that touches it goes by (the segment-after-segment
// order re-fetched each image ~8 times: profiles/r1_g, 3.2 GB of DRAM traffic in the forward kernel).
inline void build_bands(CmParams &p, long image_bytes) {
    SegTable &g = p.seg;
    static const long target = [] {                      // tuning knob; 8 MB per band measured best on B200 (profiles/r1_band_sweep.txt)
        const char *e = getenv("TEF_BAND_BYTES");
        const long v = e ? atol(e) : 0;
        return v > 0 ? v : (8l << 20);
    }();
    int nb = (int)(image_bytes / target);
    g.nbands = nb < 1 ? 1 : (nb > 32 ? 32 : nb);
    for (int s = 0; s < g.nseg; ++s) {
        const int c = g.blk_off[s + 1] - g.blk_off[s];
        g.band_ctas[s] = (c + g.nbands - 1) / g.nbands;
    }
    g.band_off[0] = 0;
    for (int k = 0; k < g.nbands; ++k) {
        int n = 0;
        for (int s = 0; s < g.nseg; ++s) {
            const int c = g.blk_off[s + 1] - g.blk_off[s];
            const int left = c - k * g.band_ctas[s];
            n += left < 0 ? 0 : (left > g.band_ctas[s] ? g.band_ctas[s] : left);
        }
        g.band_off[k + 1] = g.band_off[k] + n;
    }
}

inline int fill_params(const tef_cm_desc *d, int linear, CmParams &p) {
    int rc = check_desc(d, linear);
    if (rc) return rc;
    p.B = d->B; p.H = d->H; p.W = d->W; p.P = d->P; p.F = d->F; p.mode = d->mode;
    p.border = d->border_comp; p.loss_scaling = d->loss_scaling; p.linear = linear; p.det = d->deterministic ? 1 : 0;
    p.nchunks = (int)(((long)d->H * d->W + 2047) / 2048);
    p.gimg = (float2 *)d->gimg;
    p.res = Res::make(d->H, d->W);
    p.flow = (const float2 *)d->flow; p.gflow = (float2 *)d->gflow; p.img = (float2 *)d->img;
    p.ig.Wp = (d->W + 3) & ~1; p.ig.plane = (long)d->H * p.ig.Wp;
    p.posbuf = (float2 *)d->posbuf; p.alivebuf = (uint32_t *)d->alivebuf;
    // the quad-cell copies serve the one-hot fast paths of the Iterative kernels (not the deterministic mode, not Linear)
    const bool quad_ok = !linear && !d->deterministic;
    p.flowq = quad_ok ? (const float4 *)d->flowq : nullptr;
    p.gimgq = (quad_ok && d->flowq) ? (float4 *)d->gimgq : nullptr;
    p.acc_sum = d->acc_sum; p.acc_nnz = d->acc_nnz; p.den = d->den; p.loss = d->loss; p.grad_out = d->grad_out;
    p.nslots = build_scales(d, linear, p.sc);
    p.ig.lo_off = (long)p.F * p.B * p.nslots * 8 * p.ig.plane;      // the low words follow the high words of all images
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
    p.rows_grad = 0;
    for (int t = 0; t < d->P; ++t) p.rows_grad += (long)d->B * (d->n[0][t] > 0 ? d->n[0][t] : 0);
    SortGeom &g = p.sort;
    g.B = d->B; g.H = d->H; g.W = d->W;
    g.tiles_x = (d->W + 15) / 16; g.tiles = g.tiles_x * ((d->H + 7) / 8);
    // bins are laid out by FIXED segment slots (set * P + pass), empty passes included, so that tef_update_pass can count
    // a pass before the later ones are known; the detached set's slots exist only if it has any rows
    if (2l * d->P * d->B * g.tiles * kBinsPerTile > 0x7fffffffl) return TEF_ELIMIT;  // bin indices are 32-bit
    if ((long)d->B * 4 * p.ig.plane > 0x7fffffffl) return TEF_ELIMIT;               // merge keys (sample * slot size + offset) are 31-bit
    const int per_seg = d->B * g.tiles * kBinsPerTile;
    bool any_detached = false;
    for (int s = 0; s < ns; ++s) {
        p.seg.first_bin[s] = (p.seg.set[s] * d->P + p.seg.pass[s]) * per_seg;
        any_detached |= p.seg.set[s] == 1;
    }
    const int nslots_sort = d->P * (any_detached ? 2 : 1);
    p.seg.first_bin[ns] = nslots_sort * per_seg;
    p.hist_done = d->hist_done ? 1 : 0;
    g.blk_off[0] = 0;
    for (int s = 0; s < ns; ++s) g.blk_off[s + 1] = g.blk_off[s] + (int)(((long)d->B * p.seg.n[s] + kThreads * 4 - 1) / (kThreads * 4));
    g.nbins = (long)nslots_sort * per_seg;
    g.bins = (int *)d->sort_bins; g.sums = (int *)d->sort_sums;
    g.rec = (float4 *)d->sorted_ev;
    build_bands(p, (long)d->B * p.nslots * 4 * p.ig.plane * 8);
    return 0;
}

// the backward only visits the gradient-carrying set; its segments come first
inline void grad_segments_only(CmParams &p, long image_bytes) {
    int ng = 0;
    while (ng < p.seg.nseg && p.seg.set[ng] == 0) ++ng;
    p.seg.nseg = ng;
    build_bands(p, image_bytes);
}

// rows of segment sg in the sorted arrays: [lo, hi)
__device__ __forceinline__ void seg_rows(const CmParams &p, int sg, int &lo, int &hi) {
    const int fb = p.seg.first_bin[sg];
    lo = fb ? __ldg(p.sort.bins + fb - 1) : 0;
    hi = __ldg(p.sort.bins + p.seg.first_bin[sg + 1] - 1);
}

// CTA -> segment, thread -> sorted row; false when the thread has no event
__device__ __forceinline__ bool locate_sorted(const CmParams &p, int &t, int &b, float4 &e, float2 &m, int &row, int &set) {
    // blockIdx.x -> (band, segment, CTA inside the segment); all of it uniform per CTA
    int k = 0;
    while ((int)blockIdx.x >= p.seg.band_off[k + 1]) ++k;
    int rem = blockIdx.x - p.seg.band_off[k], sg = 0, cta = 0;
    for (;; ++sg) {
        const int c = p.seg.blk_off[sg + 1] - p.seg.blk_off[sg];
        const int left = c - k * p.seg.band_ctas[sg];
        const int n = left < 0 ? 0 : min(left, p.seg.band_ctas[sg]);
        if (rem < n) { cta = k * p.seg.band_ctas[sg] + rem; break; }
        rem -= n;
    }
    t = p.seg.pass[sg]; set = p.seg.set[sg];
    int lo, hi;
    seg_rows(p, sg, lo, hi);
    row = lo + cta * kThreads + threadIdx.x;
    if (row >= hi) return false;
    __attribute__((unused)) float m2, m3;            // record padding
    // one 256-bit load per event (LDG.E.ENL2.256): the whole 32-byte sorted record
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w), "=f"(m.x), "=f"(m.y), "=f"(m2), "=f"(m3)
                 : "l"(p.sort.rec + 2 * (long)row));
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

__device__ __forceinline__ void red_add_v4(float2 *addr, float a, float b, float c, float d) {
#ifdef TEF_EXP_NO_RED
    if (a == 123.456f) *addr = make_float2(b, c + d);     // experiment: keep the operands live, issue nothing
    return;
#endif
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// planes of one image slot: [phase][pol]
__device__ __forceinline__ float2 *img_plane(float2 *slot_base, const ImgGeom &g, int phase, int pol) {
    return slot_base + (long)(phase * 2 + pol) * g.plane;
}

// iwe_formatting (loss/flow.py:81-110) for one event: per image row one 16-byte reduction carrying
// (w_left, w_left*n, w_right, w_right*n) into the plane of the event's polarity.
// Corner coordinates, weights and in-image tests are exactly get_interpolation's (utils/iwe.py:85-107);
// a corner outside the image (or with weight 0) contributes an exact +0.
// Masks with both polarities non-zero never come out of the reference's loader ({0,1} masks); the extra work for them
// sits in non-inlined functions, so that the common case does not even issue predicated-off instructions.
static __device__ __noinline__ void splat_second_polarity(float2 *plane1, float wl, float tl, float wr, float tr, float my) {
    red_add_v4(plane1, wl * my, tl * my, wr * my, tr * my);
}
// (returns by value: reference parameters of an out-of-line function live on the stack, and the caller then stores and
// reloads them around every -- almost never taken -- call: profiles/r1_h showed 23 M local-memory wavefronts from that)
static __device__ __noinline__ float2 grad_second_polarity(const float2 *plane1, float nts, float my) {
    const float4 u = __ldg(reinterpret_cast<const float4 *>(plane1));
    return make_float2(my * (u.x + nts * u.y), my * (u.z + nts * u.w));
}

template <bool INSIDE, bool DET>
__device__ __forceinline__ void splat(float2 *__restrict__ slot_base, const Res &r, const ImgGeom &g, float y, float x, float nts, float2 m) {
    Corners c;
    corners<INSIDE>(y, x, r, c);
    if (!INSIDE && !(c.okx[0] || c.okx[1])) return;
    const int xl = (INSIDE || c.okx[0]) ? (int)c.cx[0] : (int)c.cx[1] - 1;      // column of the left corner (-1 .. W-1)
    const int phase = xl & 1;
    const int col = xl + phase;                                     // even -> 16-byte aligned pair
    const int pol = (m.x != 0.0f) ? 0 : 1;
    const float mv = pol ? m.y : m.x;
    const bool both = (m.x != 0.0f) && (m.y != 0.0f);               // non-binary masks only
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
        if (!c.oky[ky]) continue;
        const float wl = c.okx[0] ? c.wy[ky] * c.wx[0] : 0.0f;
        const float wr = c.okx[1] ? c.wy[ky] * c.wx[1] : 0.0f;
        if (wl == 0.0f && wr == 0.0f) continue;
        const int off = (int)c.cy[ky] * g.Wp + col;
        const float tl = wl * nts, tr = wr * nts;
        if (!DET) {
            if (mv == 1.0f) red_add_v4(img_plane(slot_base, g, phase, pol) + off, wl, tl, wr, tr);      // x * 1 == x: skip the products
            else red_add_v4(img_plane(slot_base, g, phase, pol) + off, wl * mv, tl * mv, wr * mv, tr * mv);
            if (both) splat_second_polarity(img_plane(slot_base, g, phase, 1) + off, wl, tl, wr, tr, m.y);
        } else {
            for (int q = pol; q < (both ? 2 : pol + 1); ++q) {
                const float mq = q ? m.y : m.x;
                long long *dst = reinterpret_cast<long long *>(slot_base) + ((long)(phase * 2 + q) * g.plane + off) * 2;
                const float v[4] = { wl * mq, tl * mq, wr * mq, tr * mq };
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (v[k] == 0.0f) continue;
                    long long hi, lo;
                    to_fix2(v[k], hi, lo);                         // exact: a non-zero contribution stays non-zero (nnz of focus_loss)
                    if (hi != 0) red_add_i64(dst + k, hi);
                    if (lo != 0) red_add_i64(dst + g.lo_off + k, lo);
                }
            }
        }
    }
}

// Warp-level merge of equal reduction addresses (DESIGN.md decision 13: the reduction path charges one sector per lane
// whatever the addresses, and tile-sorted events put equal addresses on neighbouring lanes).  `key` identifies the 16-byte
// slot a lane is about to reduce into (lanes without work pass a key no other lane has).  Inside every run of equal keys
// the lanes of odd rank hand their values to the lane before them; further rounds (TEF_MERGE_ROUNDS) repeat that among the
// survivors at distance 2, 4, ...  Returns true for a lane that gave its values away; the receivers' v[] hold the sums.
// Must be called by all 32 lanes.
#ifndef TEF_MERGE_ROUNDS
#define TEF_MERGE_ROUNDS 1
#endif
#ifndef TEF_MERGE_MODE
#define TEF_MERGE_MODE 0        // 0: runs of neighbouring lanes (the shipped kernels); 1: every lane with the same key (__match_any_sync), experiment
#endif
#if TEF_MERGE_MODE == 1
// Experiment (DESIGN.md decision 14): merge ALL lanes of the warp that reduce into the same slot, neighbours or not.  The lowest
// lane of every group of equal keys collects its followers one per round (as many rounds as the largest group has followers).
template <int NV>
__device__ __forceinline__ bool merge_equal_neighbours(unsigned key, unsigned lane, float (&v)[NV]) {
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const unsigned leader = (unsigned)__ffs((int)peers) - 1u;
    unsigned rest = (lane == leader) ? (peers & (peers - 1u)) : 0u;        // the leader's followers, lowest lane first
    while (__any_sync(0xffffffffu, rest != 0u)) {
        const unsigned src = rest ? (unsigned)__ffs((int)rest) - 1u : lane;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const float n = __shfl_sync(0xffffffffu, v[k], src);
            if (rest) v[k] += n;
        }
        rest &= rest - 1u;
    }
    return lane != leader;
}
#else
template <int NV>
__device__ __forceinline__ bool merge_equal_neighbours(unsigned key, unsigned lane, float (&v)[NV]) {
    const unsigned prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool same = lane > 0u && key == prev;
    const unsigned heads = __ballot_sync(0xffffffffu, !same);               // first lane of every run
    if (heads == 0xffffffffu) return false;                                // nothing to merge in this warp (uniform branch)
    const unsigned rank = lane - (31u - (unsigned)__clz(heads & (0xffffffffu >> (31u - lane))));
    bool gave = false;
#pragma unroll
    for (int r = 0; r < TEF_MERGE_ROUNDS; ++r) {                           // round r: rank d (mod 2d) hands over to rank 0 (mod 2d), d = 2^r
        const unsigned d = 1u << r;
        const bool give = (rank & (2u * d - 1u)) == d;
        const unsigned givers = __ballot_sync(0xffffffffu, give);
        if (r > 0 && givers == 0u) break;                                  // no run longer than d (uniform branch)
        const bool recv = lane + d < 32u && ((givers >> (lane + d)) & 1u) != 0u;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const float n = __shfl_down_sync(0xffffffffu, v[k], d);
            if (recv) v[k] += n;
        }
        gave |= give;
    }
    return gave;
}
#endif

// splat<true, false> for the events of a whole warp with one-hot {0,1} polarity masks -- the only kind the reference's loader
// produces (dataloader/base.py:264-278) -- on packed fp32x2 arithmetic, for positions (x, y) that satisfy inside(), with equal
// slots of neighbouring lanes merged before they leave the SM.  Same values as splat():
//  * the in-image tests of the bottom / right corners are implied by the weights: floor(v + 1) > size - 1 needs
//    v + 1 >= size, i.e. |v - corner| >= 1, whose clamped weight is an exact 0;
//  * v - floor(v) is exact and < 1, so the top / left weights are >= 2^-24: the top row always has a non-zero weight
//    (never skipped), and the bottom row has two zero weights exactly when its row weight is 0 (products of weights
//    >= 2^-24 do not underflow);
//  * a bottom row with a non-zero weight has floor(y + 1) == floor(y) + 1.
// `on` = this lane has an event to splat (the others only take part in the shuffles); `plane_sel` = polarity (0 / 1).  Both
// image rows of a splat share one key (equal top slots imply equal bottom slots); a lane whose own bottom row has weight 0
// carries zeros there, and the bottom row is reduced when the merged row weight is non-zero (then the row exists: the
// contributor that supplied it shares the cell).  Lanes of one warp may belong to different samples (same slot offsets,
// different images), hence `key_base`.  Must be called by all 32 lanes.
__device__ __forceinline__ void splat_inside_1hot_warp(float2 *__restrict__ slot_base, const ImgGeom &g, float2 p /* (x, y) */, float nts,
                                                       int plane_sel, bool on, unsigned lane, unsigned key_base /* sample * slot size */) {
    const float2 c0 = make_float2(floorf(p.x), floorf(p.y));
    const float2 p1 = add2(p, bc(1.0f));
    const float2 c1 = make_float2(floorf(p1.x), floorf(p1.y));
    const float2 d0 = sub2(p, c0), d1_ = sub2(p, c1);
    const float2 u0 = sub2(bc(1.0f), make_float2(fabsf(d0.x), fabsf(d0.y)));      // utils/iwe.py:96-99
    const float2 u1 = sub2(bc(1.0f), make_float2(fabsf(d1_.x), fabsf(d1_.y)));
    const float2 wx = make_float2(fmaxf(0.0f, u0.x), fmaxf(0.0f, u1.x));          // (left, right)
    const float wy0 = fmaxf(0.0f, u0.y), wy1 = fmaxf(0.0f, u1.y);
    const int xl = (int)c0.x, phase = xl & 1;
    const unsigned off = (unsigned)((phase * 2 + plane_sel) * (int)g.plane + (int)c0.y * g.Wp + xl + phase);   // a slot is far below 2^31 elements
    const float2 a = mul2(bc(wy0), wx), b = mul2(bc(wy1), wx);                    // (w_left, w_right) of the top / bottom row
    float v[8] = { a.x, a.x * nts, a.y, a.y * nts, b.x, b.x * nts, b.y, b.y * nts };
    const bool gave = merge_equal_neighbours<8>(on ? key_base + off : (0x80000000u | lane), lane, v);   // real keys stay below 2^31 (fill_params)
    if (on && !gave) {
        char *base = reinterpret_cast<char *>(slot_base);
        red_add_v4(reinterpret_cast<float2 *>(base + (size_t)off * 8u), v[0], v[1], v[2], v[3]);
        if (v[4] != 0.0f) red_add_v4(reinterpret_cast<float2 *>(base + (size_t)(off + (unsigned)g.Wp) * 8u), v[4], v[5], v[6], v[7]);
    }
}

// gradient of the loss w.r.t. the position at one reference time, through the bilinear splat weights
// (SURVEY.md Appendix A.4).  The gradient images (dL/dcount, dL/dtime-weighted) are stored in both phases
// ([phase][pol][H][Wp] float2), so each image row is one 16-byte gather of the (left, right) corner pair.
template <bool INSIDE>
__device__ __forceinline__ void iwe_grad(const float2 *__restrict__ slot_base, const Res &r, const ImgGeom &g, float y, float x, float nts,
                                         float2 m, float &gy, float &gx) {
    Corners c;
    corners<INSIDE>(y, x, r, c);
    if (!INSIDE && !(c.okx[0] || c.okx[1])) return;
    const int xl = (INSIDE || c.okx[0]) ? (int)c.cx[0] : (int)c.cx[1] - 1;
    const int phase = xl & 1;
    const int col = xl + phase;
    const float dy[2] = { d1(y, c.cy[0]), d1(y, c.cy[1]) };
    const float dx[2] = { d1(x, c.cx[0]), d1(x, c.cx[1]) };
    const bool binary = (m.y == 0.0f) || (m.x == 0.0f);            // {0,1} masks: one polarity plane is read
    const int pol0 = (m.x != 0.0f) ? 0 : 1;
    const float m0 = pol0 ? m.y : m.x;
    const float2 *g0 = slot_base + (long)(phase * 2 + pol0) * g.plane;
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
        if (!c.oky[ky]) continue;
        const int off = (int)c.cy[ky] * g.Wp + col;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(g0 + off));
        float gl = m0 * (v.x + nts * v.y), gr = m0 * (v.z + nts * v.w);
        if (!binary) {
            const float2 add = grad_second_polarity(slot_base + (long)(phase * 2 + 1) * g.plane + off, nts, m.y);
            gl = gl + add.x; gr = gr + add.y;
        }
        if (c.okx[0]) { gy += gl * dy[ky] * c.wx[0]; gx += gl * c.wy[ky] * dx[0]; }
        if (c.okx[1]) { gy += gr * dy[ky] * c.wx[1]; gx += gr * c.wy[ky] * dx[1]; }
    }
}

// iwe_grad<true> for a one-hot {0,1} polarity mask and a position (x, y) that satisfies inside(): same values, fewer
// instructions (the backward kernel is the one that is bound by instruction issue).
//  * the mask product is x * 1 == x and only the event's polarity plane is read (`pol_plane` = slot base + polarity * plane);
//  * d1() with the signs known: v - floor(v) is in [0, 1), so the top / left weight is positive and its derivative is
//    -1 for v > floor(v) and -0 on the pixel centre; floor(v + 1) >= floor(v) + 1 > v, so the bottom / right difference is
//    negative and the derivative is +1, +1/2 on the tie (weight exactly 0) and 0 beyond it;
//  * a bottom row inside the image is read at row floor(y) + 1: when floor(y + 1) is a row further (y + 1 rounded up across
//    an integer) the row weight and its derivative are both 0 and any finite row gives the same +-0 contributions.
__device__ __forceinline__ void iwe_grad_inside_1hot(const float2 *__restrict__ pol_plane, const Res &r, const ImgGeom &g, float2 p /* (x, y) */,
                                                     float nts, float &gy, float &gx) {
    const float2 c0 = make_float2(floorf(p.x), floorf(p.y));
    const float2 p1 = add2(p, bc(1.0f));
    const float2 c1 = make_float2(floorf(p1.x), floorf(p1.y));
    const float2 d0 = sub2(p, c0), e1 = sub2(p, c1);
    const float2 u0 = sub2(bc(1.0f), make_float2(fabsf(d0.x), fabsf(d0.y)));
    const float2 u1 = sub2(bc(1.0f), make_float2(fabsf(e1.x), fabsf(e1.y)));
    const float wx0 = fmaxf(0.0f, u0.x), wy0 = fmaxf(0.0f, u0.y), wx1 = fmaxf(0.0f, u1.x), wy1 = fmaxf(0.0f, u1.y);
    const float dx0 = d0.x > 0.0f ? -1.0f : -0.0f, dy0 = d0.y > 0.0f ? -1.0f : -0.0f;
    const float dx1 = u1.x > 0.0f ? 1.0f : (u1.x == 0.0f ? 0.5f : 0.0f), dy1 = u1.y > 0.0f ? 1.0f : (u1.y == 0.0f ? 0.5f : 0.0f);
    const int xl = (int)c0.x, phase = xl & 1;
    const float4 *row = reinterpret_cast<const float4 *>(pol_plane + (phase * 2 * (int)g.plane + (int)c0.y * g.Wp + xl + phase));
    const bool okx1 = c1.x <= r.wm1;
    {
        const float4 v = __ldg(row);
        const float gl = v.x + nts * v.y, gr = v.z + nts * v.w;
        gy += gl * dy0 * wx0; gx += gl * wy0 * dx0;
        if (okx1) { gy += gr * dy0 * wx1; gx += gr * wy0 * dx1; }
    }
    if (c1.y <= r.hm1) {
        const float4 v = __ldg(row + (g.Wp >> 1));
        const float gl = v.x + nts * v.y, gr = v.z + nts * v.w;
        gy += gl * dy1 * wx0; gx += gl * wy1 * dx0;
        if (okx1) { gy += gr * dy1 * wx1; gx += gr * wy1 * dx1; }
    }
}

// iwe_grad_inside_1hot on the quad-cell copy of the gradient images: `cells` = cells of (slot, polarity); both image rows of the
// corner quad arrive in one 256-bit gather.  Entries of pixels outside the image are never written and never used (the right
// column / bottom row tests below are the ones of iwe_grad_inside_1hot).
__device__ __forceinline__ void iwe_grad_inside_1hot_quad(const float4 *__restrict__ cells, const Res &r, float2 p /* (x, y) */, float nts, float &gy,
                                                          float &gx) {
    const float2 c0 = make_float2(floorf(p.x), floorf(p.y));
    const float2 p1 = add2(p, bc(1.0f));
    const float2 c1 = make_float2(floorf(p1.x), floorf(p1.y));
    const float2 d0 = sub2(p, c0), e1 = sub2(p, c1);
    const float2 u0 = sub2(bc(1.0f), make_float2(fabsf(d0.x), fabsf(d0.y)));
    const float2 u1 = sub2(bc(1.0f), make_float2(fabsf(e1.x), fabsf(e1.y)));
    const float wx0 = fmaxf(0.0f, u0.x), wy0 = fmaxf(0.0f, u0.y), wx1 = fmaxf(0.0f, u1.x), wy1 = fmaxf(0.0f, u1.y);
    const float dx0 = d0.x > 0.0f ? -1.0f : -0.0f, dy0 = d0.y > 0.0f ? -1.0f : -0.0f;
    const float dx1 = u1.x > 0.0f ? 1.0f : (u1.x == 0.0f ? 0.5f : 0.0f), dy1 = u1.y > 0.0f ? 1.0f : (u1.y == 0.0f ? 0.5f : 0.0f);
    float4 v, w;
    load_quad(cells, quad_cell(r, (int)c0.y, (int)c0.x), v, w);
    const bool okx1 = c1.x <= r.wm1;
    {
        const float gl = v.x + nts * v.y, gr = v.z + nts * v.w;
        gy += gl * dy0 * wx0; gx += gl * wy0 * dx0;
        if (okx1) { gy += gr * dy0 * wx1; gx += gr * wy0 * dx1; }
    }
    if (c1.y <= r.hm1) {
        const float gl = w.x + nts * w.y, gr = w.z + nts * w.w;
        gy += gl * dy1 * wx0; gx += gl * wy1 * dx0;
        if (okx1) { gy += gr * dy1 * wx1; gx += gr * wy1 * dx1; }
    }
}

// dL/dmap of one bilinear flow sample: two 16-byte reductions (one per tap row)
// into the dual-phase packed gradient map; c_k = dt * w_k, value = c_k * (g_x, g_y).
// DET: `ginv` = 1 / (power-of-two scale of the flow-gradient words), see det_scale_kernel
template <bool DET>
__device__ __forceinline__ void taps_red(float2 *__restrict__ gmap_phase0, const ImgGeom &g, const Taps &tp, float dt, float gpy, float gpx,
                                         float ginv = 1.0f) {
    if (tp.x0 < -1) return;                                         // sample outside the map: all taps invalid
    const int phase = tp.x0 & 1;
    const int col = tp.x0 + phase;
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
        if (!(tp.ok[2 * ky] || tp.ok[2 * ky + 1])) continue;
        const float cl = tp.ok[2 * ky] ? dt * tp.w[2 * ky] : 0.0f;
        const float cr = tp.ok[2 * ky + 1] ? dt * tp.w[2 * ky + 1] : 0.0f;
        if (cl == 0.0f && cr == 0.0f) continue;                     // e.g. the bottom row of a sample at an integer position (an event's own pixel)
        const long off = (long)phase * g.plane + (long)(tp.y0 + ky) * g.Wp + col;
        if (!DET) {
            red_add_v4(gmap_phase0 + off, cl * gpx, cl * gpy, cr * gpx, cr * gpy);
        } else {
            long long *dst = reinterpret_cast<long long *>(gmap_phase0) + off * 2;
            const float v[4] = { cl * gpx, cl * gpy, cr * gpx, cr * gpy };
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (v[k] != 0.0f) red_add_i64(dst + k, __double2ll_rn((double)v[k] * (double)ginv * kFixScale));
        }
    }
}

}  // namespace tef
