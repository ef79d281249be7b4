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
    int CX, cplane;   // quad-cell copies (flow maps, gradient images): cells per row W/2+1 and per phase plane (H/2+1)*CX, see quad_cell()
    float2 m1_xy, rm1_xy, s_xy;   // the same constants as (x, y) pairs for the packed fp32x2 path: (wm1, hm1), (rwm1, rhm1), (sw, sh)
    __host__ __device__ static Res make(int H, int W) {
        Res r; r.H = H; r.W = W; r.hm1 = (float)(H - 1); r.wm1 = (float)(W - 1);
        r.sh = r.hm1 / 2.0f; r.sw = r.wm1 / 2.0f; r.rhm1 = 1.0f / r.hm1; r.rwm1 = 1.0f / r.wm1;
        r.Wp = (W + 3) & ~1; r.fplane = (H + 1) * r.Wp;
        r.CX = W / 2 + 1; r.cplane = (H / 2 + 1) * r.CX;
        r.m1_xy.x = r.wm1; r.m1_xy.y = r.hm1; r.rm1_xy.x = r.rwm1; r.rm1_xy.y = r.rhm1; r.s_xy.x = r.sw; r.s_xy.y = r.sh;
        return r;
    }
};

// Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 of sm_100a): two independent IEEE round-to-nearest fp32 operations
// per issue slot, each half rounded exactly like the scalar instruction (no FTZ), so results stay bit-identical to the
// scalar code.  The event kernels are bound by instruction issue, and their per-event arithmetic
// comes in (x, y) pairs.  ptxas folds negation, |.| and scalar broadcast (`bc`) into the operands.
// CAUTION: unlike the scalar mul.rn / add.rn, ptxas (12.9) CONTRACTS a mul2 whose only use is an add2 into one FFMA2,
// even with --fmad=false.  Never feed a mul2 result into an add2 / sub2: keep such a step scalar.  tests/test_cabi.py
// counts the FFMA2 instructions of the event kernels against the explicit fma2() calls.
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 bc(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return add2(a, neg2(b)); }    // a - b == a + (-b) in IEEE arithmetic

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

// Quad-cell copy of a map (DESIGN.md decision 15).  The 2x2 neighbourhood (y0..y0+1) x (x0..x0+1) of an in-image position is
// ONE 32-byte record: the map is stored four times, once per parity (x0 & 1, y0 & 1), as rows of cells
// {row y0: (left, right), row y0+1: (left, right)} of float2 pixels, zero where the pixel lies outside the map.  A bilinear
// sample (or the corner quad of a gradient image) is then a single 256-bit gather touching one sector, where the dual-phase
// rows need two 16-byte gathers in two sectors.  Layout per map: [py*2+px][H/2+1][W/2+1] cells of 2 float4.
__device__ __forceinline__ int quad_cell(const Res &r, int y0, int x0) {
    const int px = x0 & 1, py = y0 & 1;
    return (py * 2 + px) * r.cplane + ((y0 + py) >> 1) * r.CX + ((x0 + px) >> 1);
}
__device__ __forceinline__ void load_quad(const float4 *__restrict__ cells, int cell, float4 &top, float4 &bot) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(top.x), "=f"(top.y), "=f"(top.z), "=f"(top.w), "=f"(bot.x), "=f"(bot.y), "=f"(bot.z), "=f"(bot.w)
                 : "l"(cells + 2 * (long)cell));
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
// The arithmetic is packed fp32x2 on (x, y) pairs: 19 floating-point issue slots instead of 32.
// div_const()'s guard becomes: exact division only for a non-zero numerator below 1e-30; a zero numerator may take
// the corrected product, which returns +0 where a / c returns the numerator's signed zero -- the "- 1.0f" that
// follows gives -1 either way.
// QUAD: `map` points at the quad-cell copy of the map (float4 cells) instead of the dual-phase rows: same eight values from one
// 256-bit gather.
template <bool KEEP, bool QUAD = false>
__device__ __forceinline__ float2 sample_flow_inside_xy(const float2 *__restrict__ map, const Res &r, float2 p /* (x, y) */, Taps *tp) {
    const float2 a = add2(p, p);                                   // 2.0f * v (exact either way)
    float2 g;
    if ((fabsf(a.x) >= 1e-30f || a.x == 0.0f) && (fabsf(a.y) >= 1e-30f || a.y == 0.0f)) {
        const float2 q0 = mul2(a, r.rm1_xy);
        const float2 r0 = fma2(neg2(q0), r.m1_xy, a);
        g = fma2(r0, r.rm1_xy, q0);
    } else {
        g = make_float2(a.x / r.wm1, a.y / r.hm1);
    }
    g = add2(g, bc(-1.0f));                                        // utils/iwe.py:30-31
    const float2 i = mul2(add2(g, bc(1.0f)), r.s_xy);              // ATen unnormalize: (ix, iy)
    const float2 fl = make_float2(floorf(i.x), floorf(i.y));
    const float2 fr = sub2(i, fl);                                 // (w_, n_)
    const float2 om = sub2(bc(1.0f), fr);                          // (e_, s_)
    const int x0 = (int)fl.x, y0 = (int)fl.y;
    float4 top, bot;
    if (QUAD) {
        load_quad(reinterpret_cast<const float4 *>(map), quad_cell(r, y0, x0), top, bot);
    } else {
        const float4 *q = tap_row(map, r, y0, x0);
        top = __ldg(q); bot = __ldg(q + (r.Wp >> 1));
    }
    const float2 ew = make_float2(om.x, fr.x);                     // (e_, w_)
    const float2 w01 = mul2(bc(om.y), ew);                         // s_*e_, s_*w_
    const float2 w23 = mul2(bc(fr.y), ew);                         // n_*e_, n_*w_
    float2 o = mul2(make_float2(top.x, top.y), bc(w01.x));         // (x-flow, y-flow): nw*w0, then three FMAs like ATen
    o = fma2(make_float2(top.z, top.w), bc(w01.y), o);
    o = fma2(make_float2(bot.x, bot.y), bc(w23.x), o);
    o = fma2(make_float2(bot.z, bot.w), bc(w23.y), o);
    if (KEEP) {
        const bool oy1 = y0 + 1 < r.H, ox1 = x0 + 1 < r.W;
        tp->w[0] = w01.x; tp->w[1] = w01.y; tp->w[2] = w23.x; tp->w[3] = w23.y;
        tp->v[0] = make_float2(top.x, top.y); tp->v[1] = make_float2(top.z, top.w);
        tp->v[2] = make_float2(bot.x, bot.y); tp->v[3] = make_float2(bot.z, bot.w);
        tp->ok[0] = true; tp->ok[1] = ox1; tp->ok[2] = oy1; tp->ok[3] = oy1 && ox1;
        tp->ax = fr.x; tp->ay = fr.y; tp->y0 = y0; tp->x0 = x0;
    }
    return o;
}
template <bool KEEP, bool QUAD = false>
__device__ __forceinline__ float2 sample_flow_inside(const float2 *__restrict__ map, const Res &r, float y, float x, Taps *tp) {
    return sample_flow_inside_xy<KEEP, QUAD>(map, r, make_float2(x, y), tp);
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
