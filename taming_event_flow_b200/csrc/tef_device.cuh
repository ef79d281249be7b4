// tef_device.cuh -- per-event device arithmetic shared by all kernels.
//
// The library is compiled with -fmad=false: every fp32 operation below rounds
// exactly once, in the order written, which is the order of the reference's
// eager PyTorch CPU path (one ATen kernel per operation).  The only fused
// multiply-adds are the explicit __fmaf_rn calls in sample_flow(): ATen's
// grid_sampler_2d accumulates its four taps as  nw*w0, fma(ne,w1,.), fma(sw,w2,.),
// fma(se,w3,.)  (established bit-exactly against torch 2.11, see
// tests/test_oracle_golden.py).  Division is IEEE (nvcc default -prec-div=true).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tef {

struct Res {
    int H, W;
    float hm1, wm1;   // (float)(H-1), (float)(W-1)
    float sh, sw;     // (H-1)/2, (W-1)/2  (ATen's align_corners=True scaling factor)
    float rhm1, rwm1; // RN(1/hm1), RN(1/wm1) for div_const()
    int Wp, fplane;   // packed flow maps: padded row length (even, >= W+2) and plane size (H+1)*Wp, see sample_flow()
    __host__ __device__ static Res make(int H, int W) {
        Res r; r.H = H; r.W = W; r.hm1 = (float)(H - 1); r.wm1 = (float)(W - 1);
        r.sh = r.hm1 / 2.0f; r.sw = r.wm1 / 2.0f; r.rhm1 = 1.0f / r.hm1; r.rwm1 = 1.0f / r.wm1;
        r.Wp = (W + 3) & ~1; r.fplane = (H + 1) * r.Wp; return r;
    }
};

// IEEE-correct a / c for a divisor known in advance, rc = RN(1/c): one Newton correction of the product
// (Markstein).  Three instructions instead of the generic division's rcp + refinement + range check.
// Verified exhaustively against a / c (every fp32 a in [0, 2c+4], c in {127, 479, 639, 255, 259, 345, 1023, 1279, 1..40, ...});
// it only deviates when the quotient is denormal, hence the guard.
__device__ __forceinline__ float div_const(float a, float c, float rc) {
    if (fabsf(a) >= 1e-30f) {
        const float q0 = a * rc;
        const float r0 = __fmaf_rn(-q0, c, a);
        return __fmaf_rn(r0, rc, q0);
    }
    return a / c;
}

// purge_unfeasible (utils/iwe.py:52-57): inclusive bounds [0, res-1]; (float)H - 1.0f == (float)(H-1) for any image size
__device__ __forceinline__ bool inside(float y, float x, const Res &r) {
    return (y >= 0.0f) && (y <= r.hm1) && (x >= 0.0f) && (x <= r.wm1);
}

// bilinear sampling set-up = utils/iwe.py:17-40 (normalisation) + ATen grid_sampler_2d
// (zeros padding, align_corners=True): tap indices, weights and fractional offsets.
struct Bil {
    float w[4];       // nw, ne, sw, se
    float ax, ay;     // fractional offsets along x / y
    int y0, x0;
    bool ok[4];
};
__device__ __forceinline__ void bilinear_setup(const Res &r, float y, float x, Bil &b) {
    float gy = div_const(2.0f * y, r.hm1, r.rhm1) - 1.0f;      // utils/iwe.py:30
    float gx = div_const(2.0f * x, r.wm1, r.rwm1) - 1.0f;      // utils/iwe.py:31
    float iy = (gy + 1.0f) * r.sh;                 // ATen unnormalize, align_corners=True
    float ix = (gx + 1.0f) * r.sw;
    float fy0 = floorf(iy), fx0 = floorf(ix);
    float w_ = ix - fx0, e_ = 1.0f - w_;
    float n_ = iy - fy0, s_ = 1.0f - n_;
    b.w[0] = s_ * e_; b.w[1] = s_ * w_; b.w[2] = n_ * e_; b.w[3] = n_ * w_;
    b.ax = w_; b.ay = n_;
    if (!(fx0 >= -2.0f && fx0 <= (float)(r.W + 1) && fy0 >= -2.0f && fy0 <= (float)(r.H + 1))) { b.y0 = -2; b.x0 = -2; }
    else { b.y0 = (int)fy0; b.x0 = (int)fx0; }
    const bool oy0 = (b.y0 >= 0) && (b.y0 < r.H), oy1 = (b.y0 + 1 >= 0) && (b.y0 + 1 < r.H);
    const bool ox0 = (b.x0 >= 0) && (b.x0 < r.W), ox1 = (b.x0 + 1 >= 0) && (b.x0 + 1 < r.W);
    b.ok[0] = oy0 && ox0; b.ok[1] = oy0 && ox1; b.ok[2] = oy1 && ox0; b.ok[3] = oy1 && ox1;
}

// Packed flow maps are float2 (x-flow, y-flow), stored TWICE per (scale, pass, sample): phase 0 holds pixel x at
// column x, phase 1 at column x+1, rows padded to Wp with zeros and one extra zero row H.  The two taps of an image
// row (x0, x0+1) are then always one 16-byte aligned float4 in phase x0&1, out-of-map taps read the zero padding,
// and a bilinear sample is two 16-byte gathers instead of four 8-byte ones (half the L2 gather lane-ops).
// ATen accumulates nw*w0, fma(ne,w1,.), fma(sw,w2,.), fma(se,w3,.).
struct Taps {
    float w[4];
    float2 v[4];      // tap values (.x = x-flow, .y = y-flow), 0 outside the map
    float ax, ay;
    int y0, x0;
    bool ok[4];
};

__device__ __forceinline__ const float4 *tap_row(const float2 *__restrict__ map, const Res &r, int y, int x0) {
    const int phase = x0 & 1;
    return reinterpret_cast<const float4 *>(map + (phase * r.fplane + y * r.Wp + x0 + phase));
}

template <bool KEEP>
__device__ __forceinline__ float2 sample_flow(const float2 *__restrict__ map, const Res &r, float y, float x, Taps *tp) {
    Bil b;
    bilinear_setup(r, y, x, b);
    float4 top = make_float4(0.f, 0.f, 0.f, 0.f), bot = top;
    if (b.x0 >= -1 && b.x0 <= r.W - 1 && b.y0 >= -1 && b.y0 <= r.H - 1) {
        if (b.y0 >= 0) top = __ldg(tap_row(map, r, b.y0, b.x0));
        bot = __ldg(tap_row(map, r, b.y0 + 1, b.x0));            // rows 0..H exist (row H is zero)
    }
    float ox = top.x * b.w[0], oy = top.y * b.w[0];
    ox = __fmaf_rn(top.z, b.w[1], ox); oy = __fmaf_rn(top.w, b.w[1], oy);
    ox = __fmaf_rn(bot.x, b.w[2], ox); oy = __fmaf_rn(bot.y, b.w[2], oy);
    ox = __fmaf_rn(bot.z, b.w[3], ox); oy = __fmaf_rn(bot.w, b.w[3], oy);
    if (KEEP) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { tp->w[k] = b.w[k]; tp->ok[k] = b.ok[k]; }
        tp->v[0] = make_float2(top.x, top.y); tp->v[1] = make_float2(top.z, top.w);
        tp->v[2] = make_float2(bot.x, bot.y); tp->v[3] = make_float2(bot.z, bot.w);
        tp->ax = b.ax; tp->ay = b.ay; tp->y0 = b.y0; tp->x0 = b.x0;
    }
    return make_float2(ox, oy);   // (.x = x-flow, .y = y-flow)
}

// Same sample for a position known to satisfy inside(): then 0 <= iy <= H-1 and 0 <= ix <= W-1 exactly
// (gy + 1 is in [0, 2]), so both tap rows exist and no bounds test is needed at all (zero padding).
// Bit-identical to sample_flow() on such positions.
template <bool KEEP>
__device__ __forceinline__ float2 sample_flow_inside(const float2 *__restrict__ map, const Res &r, float y, float x, Taps *tp) {
    const float gy = div_const(2.0f * y, r.hm1, r.rhm1) - 1.0f;
    const float gx = div_const(2.0f * x, r.wm1, r.rwm1) - 1.0f;
    const float iy = (gy + 1.0f) * r.sh, ix = (gx + 1.0f) * r.sw;
    const float fy0 = floorf(iy), fx0 = floorf(ix);
    const float w_ = ix - fx0, e_ = 1.0f - w_;
    const float n_ = iy - fy0, s_ = 1.0f - n_;
    const int y0 = (int)fy0, x0 = (int)fx0;
    const float4 *p = tap_row(map, r, y0, x0);
#ifdef TEF_EXP_NO_GATHER
    const float4 top = make_float4(0.3f * (float)(x0 & 7), -0.2f, 0.1f, 0.4f), bot = make_float4(0.2f, 0.1f * (float)(y0 & 3), -0.3f, 0.2f);
    if (p == nullptr) return make_float2(0.f, 0.f);
#else
    const float4 top = __ldg(p), bot = __ldg(p + (r.Wp >> 1));
#endif
    const float w0 = s_ * e_, w1 = s_ * w_, w2 = n_ * e_, w3 = n_ * w_;
    float ox = top.x * w0, oy = top.y * w0;
    ox = __fmaf_rn(top.z, w1, ox); oy = __fmaf_rn(top.w, w1, oy);
    ox = __fmaf_rn(bot.x, w2, ox); oy = __fmaf_rn(bot.y, w2, oy);
    ox = __fmaf_rn(bot.z, w3, ox); oy = __fmaf_rn(bot.w, w3, oy);
    if (KEEP) {
        const bool oy1 = y0 + 1 < r.H, ox1 = x0 + 1 < r.W;
        tp->w[0] = w0; tp->w[1] = w1; tp->w[2] = w2; tp->w[3] = w3;
        tp->v[0] = make_float2(top.x, top.y); tp->v[1] = make_float2(top.z, top.w);
        tp->v[2] = make_float2(bot.x, bot.y); tp->v[3] = make_float2(bot.z, bot.w);
        tp->ok[0] = true; tp->ok[1] = ox1; tp->ok[2] = oy1; tp->ok[3] = oy1 && ox1;
        tp->ax = w_; tp->ay = n_; tp->y0 = y0; tp->x0 = x0;
    }
    return make_float2(ox, oy);
}

// get_event_flow (utils/iwe.py:17-40) of one location (y, x) on planar maps [H][W]: ATen's bilinear grid_sample
// arithmetic (nw*w0, then three FMAs), zero padding; returns (flow_y, flow_x) like utils/iwe.py:38
__device__ __forceinline__ float2 event_flow_planar(const float *__restrict__ mapx, const float *__restrict__ mapy, float y, float x, const Res &r) {
    Bil bl;
    bilinear_setup(r, y, x, bl);
    const long base = (long)bl.y0 * r.W + bl.x0;
    const int off[4] = { 0, 1, r.W, r.W + 1 };
    float ox = 0.f, oy = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float vx = bl.ok[k] ? __ldg(mapx + base + off[k]) : 0.f;
        const float vy = bl.ok[k] ? __ldg(mapy + base + off[k]) : 0.f;
        if (k == 0) { ox = vx * bl.w[0]; oy = vy * bl.w[0]; }
        else { ox = __fmaf_rn(vx, bl.w[k], ox); oy = __fmaf_rn(vy, bl.w[k], oy); }
    }
    return make_float2(oy, ox);
}

// get_interpolation, bilinear branch (utils/iwe.py:85-107), one event
struct Corners {
    float cy[2], cx[2];   // top/bottom, left/right corner coordinates (as fp32)
    float wy[2], wx[2];   // clamped 1-D weights
    bool oky[2], okx[2];  // strict in-image tests (utils/iwe.py:103)
};
// INSIDE: the position satisfies inside(), so the top/left corners are in the image and only the
// bottom/right ones (floor(v + 1) <= size) need the strict test.
template <bool INSIDE = false>
__device__ __forceinline__ void corners(float y, float x, const Res &r, Corners &c) {
    c.cy[0] = floorf(y); c.cy[1] = floorf(y + 1.0f);
    c.cx[0] = floorf(x); c.cx[1] = floorf(x + 1.0f);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        c.wy[k] = fmaxf(0.0f, 1.0f - fabsf(y - c.cy[k]));
        c.wx[k] = fmaxf(0.0f, 1.0f - fabsf(x - c.cx[k]));
        if (INSIDE) {
            c.oky[k] = (k == 0) || (c.cy[k] <= r.hm1);
            c.okx[k] = (k == 0) || (c.cx[k] <= r.wm1);
        } else {
            c.oky[k] = (c.cy[k] >= 0.0f) && (c.cy[k] < (float)r.H);
            c.okx[k] = (c.cx[k] >= 0.0f) && (c.cx[k] < (float)r.W);
        }
    }
}

// d/dv max(0, 1-|v-c|) with autograd's conventions: abs'(0)=0, max() ties split 1/2
__device__ __forceinline__ float d1(float v, float c) {
    float d = v - c;
    float u = 1.0f - fabsf(d);
    float sg = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
    return (u > 0.0f) ? -sg : ((u == 0.0f) ? -0.5f * sg : 0.0f);
}

// native fp32 reduction (REDG.E.ADD.F32, no return value requested); the 16-byte variant is in tef_cm_common.cuh
__device__ __forceinline__ void red_add_f32(float *addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

}  // namespace tef
