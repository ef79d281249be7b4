// tef_primitives.cu -- the stand-alone operators of the reference's utils/iwe.py as sm_100a kernels
// (forward and the backward autograd needs).  One thread per event (or per corner row); inputs are read
// with coalesced vector loads, image accumulation uses native REDG reductions.  The fused loss kernels
// (tef_cm_*.cu) do not call these; they exist so that `from utils.iwe import ...` users keep working.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

#define TEF_GRID(n) (unsigned)(((n) + kThreads - 1) / kThreads)

// ---- event_propagation (utils/iwe.py:5-14) ------------------------------------------------------
__global__ void __launch_bounds__(kThreads) prop_fwd_kernel(const float *__restrict__ ts, const float2 *__restrict__ loc,
                                                            const float2 *__restrict__ flow, float tref, float2 *__restrict__ out, long n) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float dt = tref - ts[i];
    const float2 l = loc[i], f = flow[i];
    out[i] = make_float2(l.x + dt * f.x, l.y + dt * f.y);
}
__global__ void __launch_bounds__(kThreads) prop_bwd_kernel(const float2 *__restrict__ g, const float *__restrict__ ts,
                                                            const float2 *__restrict__ flow, float tref, float *__restrict__ g_ts,
                                                            float2 *__restrict__ g_loc, float2 *__restrict__ g_flow, long n) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float dt = tref - ts[i];
    const float2 go = g[i], f = flow[i];
    if (g_loc) g_loc[i] = go;
    if (g_flow) g_flow[i] = make_float2(go.x * dt, go.y * dt);
    if (g_ts) g_ts[i] = -(go.x * f.x + go.y * f.y);
}

// ---- get_event_flow (utils/iwe.py:17-40), planar maps --------------------------------------------
__global__ void __launch_bounds__(kThreads) gef_fwd_kernel(const float *__restrict__ mapx, const float *__restrict__ mapy,
                                                           const float2 *__restrict__ loc, float2 *__restrict__ out, int B, int N, Res r) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * N) return;
    const long HW = (long)r.H * r.W;
    const int b = (int)(i / N);
    const float2 l = loc[i];
    out[i] = event_flow_planar(mapx + b * HW, mapy + b * HW, l.x, l.y, r);
}
__global__ void __launch_bounds__(kThreads) gef_bwd_kernel(const float2 *__restrict__ gout, const float *__restrict__ mapx,
                                                           const float *__restrict__ mapy, const float2 *__restrict__ loc,
                                                           float *__restrict__ g_mapx, float *__restrict__ g_mapy, float2 *__restrict__ g_loc,
                                                           int B, int N, Res r) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * N) return;
    const long HW = (long)r.H * r.W;
    const int b = (int)(i / N);
    const float2 l = loc[i];
    const float2 go = gout[i];         // (d/d flow_y, d/d flow_x)
    Bil bl;
    bilinear_setup(r, l.x, l.y, bl);
    const long base = b * HW + (long)bl.y0 * r.W + bl.x0;
    const int off[4] = { 0, 1, r.W, r.W + 1 };
    float vx[4], vy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        vx[k] = bl.ok[k] ? __ldg(mapx + base + off[k]) : 0.f;
        vy[k] = bl.ok[k] ? __ldg(mapy + base + off[k]) : 0.f;
        if (bl.ok[k]) {
            if (g_mapx) red_add_f32(g_mapx + base + off[k], bl.w[k] * go.y);
            if (g_mapy) red_add_f32(g_mapy + base + off[k], bl.w[k] * go.x);
        }
    }
    if (g_loc) {
        // ATen grid_sampler_2d_backward: d/dix = (ne-nw)*s + (se-sw)*n, d/diy = (sw-nw)*e + (se-ne)*w, scaled by
        // (size-1)/2, then the chain rule of 2*v/(size-1)-1 (utils/iwe.py:30-31)
        const float s_ = 1.0f - bl.ay, e_ = 1.0f - bl.ax;
        float gix = ((vx[1] - vx[0]) * s_ + (vx[3] - vx[2]) * bl.ay) * go.y + ((vy[1] - vy[0]) * s_ + (vy[3] - vy[2]) * bl.ay) * go.x;
        float giy = ((vx[2] - vx[0]) * e_ + (vx[3] - vx[1]) * bl.ax) * go.y + ((vy[2] - vy[0]) * e_ + (vy[3] - vy[1]) * bl.ax) * go.x;
        gix = ((gix * r.sw) / r.wm1) * 2.0f;
        giy = ((giy * r.sh) / r.hm1) * 2.0f;
        g_loc[i] = make_float2(giy, gix);
    }
}

// ---- purge_unfeasible (utils/iwe.py:43-60) --------------------------------------------------------
__global__ void __launch_bounds__(kThreads) purge_kernel(const float2 *__restrict__ loc, const float2 *__restrict__ a, const float2 *__restrict__ b,
                                                         float2 *__restrict__ oa, float2 *__restrict__ ob, long n, Res r) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float2 l = loc[i];
    const float in = inside(l.x, l.y, r) ? 1.0f : 0.0f;
    if (a && oa) { const float2 v = a[i]; oa[i] = make_float2(v.x * in, v.y * in); }
    if (b && ob) { const float2 v = b[i]; ob[i] = make_float2(v.x * in, v.y * in); }
}

// ---- get_interpolation (utils/iwe.py:63-113) ------------------------------------------------------
__global__ void __launch_bounds__(kThreads) interp_idx_kernel(const float2 *__restrict__ warped, float *__restrict__ idx, float *__restrict__ w,
                                                              int B, int N, Res r, int round_idx) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * N) return;
    const int b = (int)(i / N), e = (int)(i % N);
    const float2 l = warped[i];
    if (round_idx) {
        const float ry = rintf(l.x), rx = rintf(l.y);              // torch.round: half to even (:79)
        const float ok = (ry >= 0.f && ry < (float)r.H && rx >= 0.f && rx < (float)r.W) ? 1.0f : 0.0f;
        idx[i] = (ry * ok) * (float)r.W + rx * ok;
        w[i] = (1.0f * 1.0f) * ok;
        return;
    }
    Corners c;
    corners(l.x, l.y, r, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ky = k >> 1, kx = k & 1;                          // TL, TR, BL, BR (:90-94)
        const float ok = (c.oky[ky] && c.okx[kx]) ? 1.0f : 0.0f;
        const long o = (long)b * 4 * N + (long)k * N + e;
        idx[o] = (c.cy[ky] * ok) * (float)r.W + c.cx[kx] * ok;      // :104,:110-111
        w[o] = (c.wy[ky] * c.wx[kx]) * ok;                          // :107
    }
}
__global__ void __launch_bounds__(kThreads) interp_idx_bwd_kernel(const float2 *__restrict__ warped, const float *__restrict__ g_w,
                                                                  float2 *__restrict__ g_warped, int B, int N, Res r) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * N) return;
    const int b = (int)(i / N), e = (int)(i % N);
    const float2 l = warped[i];
    Corners c;
    corners(l.x, l.y, r, c);
    float gy = 0.f, gx = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ky = k >> 1, kx = k & 1;
        if (!(c.oky[ky] && c.okx[kx])) continue;
        const float g = g_w[(long)b * 4 * N + (long)k * N + e];
        gy += g * d1(l.x, c.cy[ky]) * c.wx[kx];
        gx += g * c.wy[ky] * d1(l.y, c.cx[kx]);
    }
    g_warped[i] = make_float2(gy, gx);
}

// ---- interpolate (utils/iwe.py:116-136) ------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) scatter_kernel(const float *__restrict__ idx, const float *__restrict__ w, const float *__restrict__ pol,
                                                           float *__restrict__ iwe, int B, long M, long HW) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * M) return;
    const int b = (int)(i / M);
    float v = w[i];
    if (pol) v = v * pol[i];
    const long px = (long)idx[i];                                   // idx.long() (:134)
    if (px < 0 || px >= HW) return;                                 // the reference would raise; never produced by get_interpolation
    if (v != 0.0f) red_add_f32(iwe + b * HW + px, v);
}
__global__ void __launch_bounds__(kThreads) gather_kernel(const float *__restrict__ idx, const float *__restrict__ pol, const float *__restrict__ w,
                                                          const float *__restrict__ g_iwe, float *__restrict__ g_w, float *__restrict__ g_pol,
                                                          int B, long M, long HW) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * M) return;
    const int b = (int)(i / M);
    const long px = (long)idx[i];
    const float g = (px >= 0 && px < HW) ? __ldg(g_iwe + b * HW + px) : 0.f;
    if (g_w) g_w[i] = pol ? g * pol[i] : g;
    if (g_pol && pol) g_pol[i] = g * w[i];
}

// ---- deblur_events (utils/iwe.py:139-224), fused ----------------------------------------------------
__global__ void __launch_bounds__(kThreads) deblur_kernel(const float *__restrict__ flow, const float4 *__restrict__ ev, const float *__restrict__ pol,
                                                          long pol_stride, float *__restrict__ iwe, long iwe_bstride, int B, int N, Res r,
                                                          int round_idx, int round_flow) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * N) return;
    const long HW = (long)r.H * r.W;
    const int b = (int)(i / N);
    const float4 e = ev[i];
    const float *fx = flow + (long)b * 2 * HW, *fy = fx + HW;
    const float y = e.y, x = e.z;
    const float feas = (y >= 0.f && y < (float)r.H && x >= 0.f && x < (float)r.W) ? 1.0f : 0.0f;   // :154-160
    const float qy = y * feas, qx = x * feas;
    float vy, vx;
    if (round_flow) {
        const long id = (long)(qy * (float)r.W + qx);               // :185-191, .long() truncates
        const bool ok = id >= 0 && id < HW;
        vy = ok ? __ldg(fy + id) : 0.f; vx = ok ? __ldg(fx + id) : 0.f;
    } else {
        Corners c;                                                  // :164-209 manual 4-tap gather, plain multiply-adds
        corners(qy, qx, r, c);
        vy = 0.f; vx = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ky = k >> 1, kx = k & 1;
            const float ok = (c.oky[ky] && c.okx[kx]) ? 1.0f : 0.0f;
            const long id = (long)((c.cy[ky] * ok) * (float)r.W + c.cx[kx] * ok);
            const float wk = (c.wy[ky] * c.wx[kx]) * ok;
            const float ty = wk * __ldg(fy + id), tx = wk * __ldg(fx + id);
            vy = (k == 0) ? ty : vy + ty; vx = (k == 0) ? tx : vx + tx;
        }
    }
    const float dt = 1.0f - e.x;
    const float wy_ = y + dt * vy, wx_ = x + dt * vx;               // :214 warps the unmasked location
    const float pm = pol ? pol[i * pol_stride] : 1.0f;
    float *im = iwe + b * iwe_bstride;
    if (round_idx) {
        const float ry = rintf(wy_), rx = rintf(wx_);
        if (ry >= 0.f && ry < (float)r.H && rx >= 0.f && rx < (float)r.W) {
            const float v = (1.0f * feas) * pm;
            if (v != 0.0f) red_add_f32(im + (long)ry * r.W + (long)rx, v);
        }
    } else {
        Corners c;
        corners(wy_, wx_, r, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ky = k >> 1, kx = k & 1;
            if (!(c.oky[ky] && c.okx[kx])) continue;
            const float v = ((c.wy[ky] * c.wx[kx]) * feas) * pm;
            if (v != 0.0f) red_add_f32(im + (long)c.cy[ky] * r.W + (long)c.cx[kx], v);
        }
    }
}

}  // namespace tef

using namespace tef;
#define ST ((cudaStream_t)stream)

extern "C" int tef_event_propagation(const float *ts, const float *loc, const float *flow, float tref, float *out, long n, void *stream) {
    if (n < 0) return TEF_EINVAL; if (n == 0) return 0;
    if (!ts || !loc || !flow || !out) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    prop_fwd_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>(ts, (const float2 *)loc, (const float2 *)flow, tref, (float2 *)out, n);
    return (int)cudaGetLastError();
}
extern "C" int tef_event_propagation_bwd(const float *gout, const float *ts, const float *flow, float tref, float *g_ts, float *g_loc,
                                         float *g_flow, long n, void *stream) {
    if (n < 0) return TEF_EINVAL; if (n == 0) return 0;
    if (!gout || !ts || !flow) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    prop_bwd_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>((const float2 *)gout, ts, (const float2 *)flow, tref, g_ts, (float2 *)g_loc, (float2 *)g_flow, n);
    return (int)cudaGetLastError();
}
extern "C" int tef_get_event_flow(const float *mapx, const float *mapy, const float *loc, float *out, int B, int N, int H, int W, void *stream) {
    if (B < 0 || N < 0 || H < 2 || W < 2) return TEF_EINVAL; if ((long)B * N == 0) return 0;
    if (!mapx || !mapy || !loc || !out) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    gef_fwd_kernel<<<TEF_GRID((long)B * N), kThreads, 0, ST>>>(mapx, mapy, (const float2 *)loc, (float2 *)out, B, N, Res::make(H, W));
    return (int)cudaGetLastError();
}
extern "C" int tef_get_event_flow_bwd(const float *gout, const float *mapx, const float *mapy, const float *loc, float *g_mapx, float *g_mapy,
                                      float *g_loc, int B, int N, int H, int W, void *stream) {
    if (B < 0 || N < 0 || H < 2 || W < 2) return TEF_EINVAL; if ((long)B * N == 0) return 0;
    if (!gout || !mapx || !mapy || !loc) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    gef_bwd_kernel<<<TEF_GRID((long)B * N), kThreads, 0, ST>>>((const float2 *)gout, mapx, mapy, (const float2 *)loc, g_mapx, g_mapy, (float2 *)g_loc,
                                                              B, N, Res::make(H, W));
    return (int)cudaGetLastError();
}
extern "C" int tef_purge_unfeasible(const float *loc, const float *mask, float *out_loc, float *out_mask, long n, int H, int W, void *stream) {
    if (n < 0) return TEF_EINVAL; if (n == 0) return 0;
    if (!loc) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    purge_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>((const float2 *)loc, (const float2 *)loc, (const float2 *)mask, (float2 *)out_loc, (float2 *)out_mask, n,
                                                  Res::make(H, W));
    return (int)cudaGetLastError();
}
extern "C" int tef_purge_unfeasible_bwd(const float *loc, const float *g_loc_out, const float *g_mask_out, float *g_loc, float *g_mask, long n, int H,
                                        int W, void *stream) {
    if (n < 0) return TEF_EINVAL; if (n == 0) return 0;
    if (!loc) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    purge_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>((const float2 *)loc, (const float2 *)g_loc_out, (const float2 *)g_mask_out, (float2 *)g_loc,
                                                  (float2 *)g_mask, n, Res::make(H, W));
    return (int)cudaGetLastError();
}
extern "C" int tef_get_interpolation(const float *warped, float *idx, float *w, int B, int N, int H, int W, int round_idx, void *stream) {
    if (B < 0 || N < 0) return TEF_EINVAL; if ((long)B * N == 0) return 0;
    if (!warped || !idx || !w) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    interp_idx_kernel<<<TEF_GRID((long)B * N), kThreads, 0, ST>>>((const float2 *)warped, idx, w, B, N, Res::make(H, W), round_idx);
    return (int)cudaGetLastError();
}
extern "C" int tef_get_interpolation_bwd(const float *warped, const float *g_w, float *g_warped, int B, int N, int H, int W, void *stream) {
    if (B < 0 || N < 0) return TEF_EINVAL; if ((long)B * N == 0) return 0;
    if (!warped || !g_w || !g_warped) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    interp_idx_bwd_kernel<<<TEF_GRID((long)B * N), kThreads, 0, ST>>>((const float2 *)warped, g_w, (float2 *)g_warped, B, N, Res::make(H, W));
    return (int)cudaGetLastError();
}
extern "C" int tef_interpolate(const float *idx, const float *w, const float *pol, float *iwe, int B, long M, int H, int W, void *stream) {
    if (B < 0 || M < 0) return TEF_EINVAL; if ((long)B * M == 0) return 0;
    if (!idx || !w || !iwe) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    scatter_kernel<<<TEF_GRID((long)B * M), kThreads, 0, ST>>>(idx, w, pol, iwe, B, M, (long)H * W);
    return (int)cudaGetLastError();
}
extern "C" int tef_interpolate_bwd(const float *idx, const float *pol, const float *w, const float *g_iwe, float *g_w, float *g_pol, int B, long M,
                                   int H, int W, void *stream) {
    if (B < 0 || M < 0) return TEF_EINVAL; if ((long)B * M == 0) return 0;
    if (!idx || !g_iwe) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    gather_kernel<<<TEF_GRID((long)B * M), kThreads, 0, ST>>>(idx, pol, w, g_iwe, g_w, g_pol, B, M, (long)H * W);
    return (int)cudaGetLastError();
}
extern "C" int tef_deblur_events(const float *flow, const float *events, const float *pol, long pol_stride, float *iwe, long iwe_batch_stride, int B,
                                 int N, int H, int W, int round_idx, int round_flow, void *stream) {
    if (B < 0 || N < 0 || H < 1 || W < 1) return TEF_EINVAL;
    if (!iwe) return TEF_EINVAL;
    for (int b = 0; b < B; ++b) cudaMemsetAsync(iwe + (long)b * iwe_batch_stride, 0, sizeof(float) * (long)H * W, ST);
    if ((long)B * N == 0) return 0;
    if (!flow || !events) return TEF_EINVAL;
    ProfScope ps(K_PRIMITIVE, ST);
    deblur_kernel<<<TEF_GRID((long)B * N), kThreads, 0, ST>>>(flow, (const float4 *)events, pol, pol_stride, iwe, iwe_batch_stride, B, N, Res::make(H, W),
                                                             round_idx, round_flow);
    return (int)cudaGetLastError();
}
