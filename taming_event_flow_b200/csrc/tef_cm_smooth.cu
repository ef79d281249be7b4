// tef_cm_smooth.cu -- the two smoothness priors of the loss (loss/flow.py:131-209), forward and backward,
// on the packed flow maps `update` already built (float2 (x, y), dual-phase, zero-padded: tef_device.cuh).
//
//   spatial:  Charbonnier sqrt(d^2 + 1e-6) of the horizontal, vertical and both diagonal differences of both flow
//             components, mean over pixels, mean over passes, mean of the four directions, mean over flow scales.
//   temporal: each map is compared with the NEXT map sampled (bilinear, zeros outside) where its own flow points:
//             sum_c sqrt((v_j - v_{j+1}(p + v_j))^2 + 1e-9) averaged over the pixels whose target stays in the image.
//
// Stencil / one-gather kernels, HBM-streaming (8 B per pixel per map read, 8 B written in the backward); sums are
// per-block partials added in a fixed order in double, so the values are reproducible.  Gradients go to a packed
// dual-phase gradient buffer (the layout tef_unpack_flow_grad reads): plain stores for the stencil, the loss kernels'
// 16-byte reductions for the bilinear taps of the temporal prior.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

struct SmoothGeom {
    int B, H, W, P, F, n;        // P: passes of the packed layout, n <= P: passes given to update()
    int nb;                      // blocks per map
    Res r;
    ImgGeom g;
};

__device__ __forceinline__ float block_sum(float v, float *sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0)
        for (int k = 0; k < kThreads / 32; ++k) t += sh[k];
    __syncthreads();
    return t;                    // valid in thread 0
}

__device__ __forceinline__ float charb(float d, float eps) { return sqrtf(d * d + eps); }
__device__ __forceinline__ float charb2(float2 a, float2 b, float eps) { return charb(a.x - b.x, eps) + charb(a.y - b.y, eps); }
__device__ __forceinline__ float2 dcharb2(float2 a, float2 b, float eps, float c) {      // c * d/da of charb2(a, b)
    const float dx = a.x - b.x, dy = a.y - b.y;
    return make_float2(c * (dx / sqrtf(dx * dx + eps)), c * (dy / sqrtf(dy * dy + eps)));
}

// map index (blockIdx.y) -> (f, t, b) over the first n passes
__device__ __forceinline__ void map_of(const SmoothGeom &s, int n_t, int &f, int &t, int &b) {
    const int m = blockIdx.y;
    b = m % s.B; t = (m / s.B) % n_t; f = m / (s.B * n_t);
}
__device__ __forceinline__ const float2 *flow_map(const float2 *packed, const SmoothGeom &s, int f, int t, int b) {
    return packed + (((long)f * s.P + t) * s.B + b) * 2 * s.r.fplane;
}
__device__ __forceinline__ float2 *grad_map(float2 *packed, const SmoothGeom &s, int f, int t, int b) {
    return packed + (((long)f * s.P + t) * s.B + b) * 2 * s.g.plane;
}

// ---- spatial prior (loss/flow.py:170-209) ---------------------------------------------------------
__global__ void __launch_bounds__(kThreads) spat_fwd_kernel(const float2 *__restrict__ packed, float *__restrict__ part, SmoothGeom s) {
    __shared__ float sh[kThreads / 32];
    int f, t, b;
    map_of(s, s.n, f, t, b);
    const float2 *m = flow_map(packed, s, f, t, b);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float d[4] = { 0.f, 0.f, 0.f, 0.f };
    if (i < s.H * s.W) {
        const int y = i / s.W, x = i - y * s.W;
        const float2 v = m[y * s.r.Wp + x];
        const bool xr = x + 1 < s.W, yd = y + 1 < s.H;
        if (xr) d[0] = charb2(v, m[y * s.r.Wp + x + 1], 1e-6f);                              // [:, :-1] - [:, 1:]
        if (yd) d[1] = charb2(v, m[(y + 1) * s.r.Wp + x], 1e-6f);                            // [:-1, :] - [1:, :]
        if (xr && yd) {
            d[2] = charb2(v, m[(y + 1) * s.r.Wp + x + 1], 1e-6f);                            // [:-1, :-1] - [1:, 1:]
            d[3] = charb2(m[(y + 1) * s.r.Wp + x], m[y * s.r.Wp + x + 1], 1e-6f);            // [1:, :-1] - [:-1, 1:]
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float tot = block_sum(d[k], sh);
        if (threadIdx.x == 0) part[((long)blockIdx.y * s.nb + blockIdx.x) * 4 + k] = tot;
    }
}
__global__ void spat_finalize_kernel(const float *__restrict__ part, float *__restrict__ out, SmoothGeom s) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    const double cnt[4] = { (double)s.H * (s.W - 1), (double)(s.H - 1) * s.W, (double)(s.H - 1) * (s.W - 1), (double)(s.H - 1) * (s.W - 1) };
    double loss = 0.0;
    for (int f = 0; f < s.F; ++f) {
        double dir[4] = { 0.0, 0.0, 0.0, 0.0 };
        for (int t = 0; t < s.n; ++t) {
            const float *p = part + ((((long)f * s.n + t) * s.B + b) * s.nb) * 4;
            double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
            for (int k = 0; k < s.nb; ++k)
                for (int c = 0; c < 4; ++c) acc[c] += (double)p[k * 4 + c];
            for (int c = 0; c < 4; ++c) dir[c] += acc[c] / cnt[c];                           // .mean(2)
        }
        loss += (dir[0] + dir[1] + dir[2] + dir[3]) / s.n / 4.0;                             // .mean(1), /4
    }
    out[b] = (float)(loss / s.F);
}
// gather form: every pixel collects d/dv of the (up to) eight pairs it belongs to; plain stores into phase 0
__global__ void __launch_bounds__(kThreads) spat_bwd_kernel(const float2 *__restrict__ packed, const float *__restrict__ gout, float2 *__restrict__ gpacked,
                                                            SmoothGeom s) {
    int f, t, b;
    map_of(s, s.n, f, t, b);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= s.H * s.W) return;
    const float2 *m = flow_map(packed, s, f, t, b);
    const int y = i / s.W, x = i - y * s.W;
    const float base = gout[b] / (float)s.F / 4.0f / (float)s.n;
    const float c_dx = base / ((float)s.H * (float)(s.W - 1)), c_dy = base / ((float)(s.H - 1) * (float)s.W);
    const float c_dg = base / ((float)(s.H - 1) * (float)(s.W - 1));
    const float2 v = m[y * s.r.Wp + x];
    float2 g = make_float2(0.f, 0.f);
    auto pair = [&](int yy, int xx, float c) {
        if (yy < 0 || yy >= s.H || xx < 0 || xx >= s.W) return;
        const float2 q = dcharb2(v, m[yy * s.r.Wp + xx], 1e-6f, c);
        g.x += q.x; g.y += q.y;
    };
    pair(y, x + 1, c_dx); pair(y, x - 1, c_dx);
    pair(y + 1, x, c_dy); pair(y - 1, x, c_dy);
    pair(y + 1, x + 1, c_dg); pair(y - 1, x - 1, c_dg);
    pair(y - 1, x + 1, c_dg); pair(y + 1, x - 1, c_dg);
    grad_map(gpacked, s, f, t, b)[y * s.g.Wp + x] = g;
}

// ---- temporal prior (loss/flow.py:131-168) --------------------------------------------------------
__global__ void __launch_bounds__(kThreads) temp_fwd_kernel(const float2 *__restrict__ packed, float *__restrict__ part, SmoothGeom s) {
    __shared__ float sh[kThreads / 32];
    int f, j, b;
    map_of(s, s.n - 1, f, j, b);
    const float2 *m = flow_map(packed, s, f, j, b), *mn = flow_map(packed, s, f, j + 1, b);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    if (i < s.H * s.W) {
        const int y = i / s.W, x = i - y * s.W;
        const float2 v = m[y * s.r.Wp + x];                                                   // (x, y)
        const float ty = (float)y + v.y, tx = (float)x + v.x;
        const float2 nx = sample_flow<false>(mn, s.r, ty, tx, nullptr);
        const float d = charb(v.y - nx.y, 1e-9f) + charb(v.x - nx.x, 1e-9f);
        const float in = inside(ty, tx, s.r) ? 1.0f : 0.0f;
        s1 = d * in; s2 = in;
    }
    const float t1 = block_sum(s1, sh), t2 = block_sum(s2, sh);
    if (threadIdx.x == 0) {
        part[((long)blockIdx.y * s.nb + blockIdx.x) * 2] = t1;
        part[((long)blockIdx.y * s.nb + blockIdx.x) * 2 + 1] = t2;
    }
}
// sums [F][n-1][B][2] = (sum of masked differences, number of valid pixels), kept for the backward
__global__ void temp_finalize_kernel(const float *__restrict__ part, float *__restrict__ sums, float *__restrict__ out, SmoothGeom s) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    double loss = 0.0;
    for (int f = 0; f < s.F; ++f)
        for (int j = 0; j < s.n - 1; ++j) {
            const long mi = ((long)f * (s.n - 1) + j) * s.B + b;
            const float *p = part + mi * s.nb * 2;
            double a1 = 0.0, a2 = 0.0;
            for (int k = 0; k < s.nb; ++k) { a1 += (double)p[2 * k]; a2 += (double)p[2 * k + 1]; }
            sums[mi * 2] = (float)a1; sums[mi * 2 + 1] = (float)a2;
            loss += (double)((float)a1 / ((float)a2 + 1e-9f));                                // loss/flow.py:163
        }
    out[b] = (float)(loss / s.F / (s.n - 1));
}
__global__ void __launch_bounds__(kThreads) temp_bwd_kernel(const float2 *__restrict__ packed, const float *__restrict__ sums, const float *__restrict__ gout,
                                                            float2 *__restrict__ gpacked, SmoothGeom s) {
    int f, j, b;
    map_of(s, s.n - 1, f, j, b);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= s.H * s.W) return;
    const float2 *m = flow_map(packed, s, f, j, b), *mn = flow_map(packed, s, f, j + 1, b);
    const int y = i / s.W, x = i - y * s.W;
    const float2 v = m[y * s.r.Wp + x];
    const float ty = (float)y + v.y, tx = (float)x + v.x;
    if (!inside(ty, tx, s.r)) return;                                                        // masked out: no gradient
    Taps tp;
    const float2 nx = sample_flow_inside<true>(mn, s.r, ty, tx, &tp);
    const float coef = gout[b] / (float)s.F / (float)(s.n - 1) / (sums[(((long)f * (s.n - 1) + j) * s.B + b) * 2 + 1] + 1e-9f);
    const float dy = v.y - nx.y, dx = v.x - nx.x;
    const float gy = coef * (dy / sqrtf(dy * dy + 1e-9f)), gx = coef * (dx / sqrtf(dx * dx + 1e-9f));     // d/d v_j (direct); d/d nxt = -g
    taps_red<false>(grad_map(gpacked, s, f, j + 1, b), s.g, tp, 1.0f, -gy, -gx);             // into map j+1, bilinear weights
    // through the sampling position (tgt = pixel + v_j): same derivative as a reverse chain step
    const float dvy_dy = (1.0f - tp.ax) * (tp.v[2].y - tp.v[0].y) + tp.ax * (tp.v[3].y - tp.v[1].y);
    const float dvy_dx = (1.0f - tp.ay) * (tp.v[1].y - tp.v[0].y) + tp.ay * (tp.v[3].y - tp.v[2].y);
    const float dvx_dy = (1.0f - tp.ax) * (tp.v[2].x - tp.v[0].x) + tp.ax * (tp.v[3].x - tp.v[1].x);
    const float dvx_dx = (1.0f - tp.ay) * (tp.v[1].x - tp.v[0].x) + tp.ay * (tp.v[3].x - tp.v[2].x);
    const float cy = gy - (dvy_dy * gy + dvx_dy * gx), cx = gx - (dvy_dx * gy + dvx_dx * gx);
    float *dst = reinterpret_cast<float *>(grad_map(gpacked, s, f, j, b) + y * s.g.Wp + x);
    red_add_f32(dst, cx);
    red_add_f32(dst + 1, cy);
}

}  // namespace tef

using namespace tef;
#define ST ((cudaStream_t)stream)

static int smooth_geom(int B, int H, int W, int P, int F, int n, SmoothGeom &s) {
    if (B < 1 || H < 2 || W < 2 || P < 1 || F < 1 || n < 1 || n > P) return TEF_EINVAL;
    s.B = B; s.H = H; s.W = W; s.P = P; s.F = F; s.n = n;
    s.nb = (H * W + kThreads - 1) / kThreads;
    s.r = Res::make(H, W);
    s.g.Wp = (W + 3) & ~1; s.g.plane = (long)H * s.g.Wp;
    return 0;
}

extern "C" long tef_flow_smoothing_scratch(int B, int H, int W, int n, int F) {
    return (long)F * n * B * ((H * W + kThreads - 1) / kThreads) * 4;
}

extern "C" int tef_flow_spatial_smoothing(const float *packed_flow, int B, int H, int W, int P, int F, int n, float *scratch, float *out, void *stream) {
    SmoothGeom s;
    int rc = smooth_geom(B, H, W, P, F, n, s);
    if (rc) return rc;
    if (!packed_flow || !scratch || !out) return TEF_EINVAL;
    ProfScope pr(K_SMOOTH, ST);
    spat_fwd_kernel<<<dim3(s.nb, F * n * B), kThreads, 0, ST>>>((const float2 *)packed_flow, scratch, s);
    spat_finalize_kernel<<<(B + 63) / 64, 64, 0, ST>>>(scratch, out, s);
    return (int)cudaGetLastError();
}

extern "C" int tef_flow_spatial_smoothing_bwd(const float *packed_flow, const float *gout, float *packed_grad, int B, int H, int W, int P, int F, int n,
                                              void *stream) {
    SmoothGeom s;
    int rc = smooth_geom(B, H, W, P, F, n, s);
    if (rc) return rc;
    if (!packed_flow || !gout || !packed_grad) return TEF_EINVAL;
    cudaMemsetAsync(packed_grad, 0, sizeof(float2) * (long)F * P * B * 2 * s.g.plane, ST);
    ProfScope pr(K_SMOOTH, ST);
    spat_bwd_kernel<<<dim3(s.nb, F * n * B), kThreads, 0, ST>>>((const float2 *)packed_flow, gout, (float2 *)packed_grad, s);
    return (int)cudaGetLastError();
}

extern "C" int tef_flow_temporal_smoothing(const float *packed_flow, int B, int H, int W, int P, int F, int n, float *scratch, float *sums, float *out,
                                           void *stream) {
    SmoothGeom s;
    int rc = smooth_geom(B, H, W, P, F, n, s);
    if (rc) return rc;
    if (n < 2) return TEF_EINVAL;                                   // upstream divides by passes - 1
    if (!packed_flow || !scratch || !sums || !out) return TEF_EINVAL;
    ProfScope pr(K_SMOOTH, ST);
    temp_fwd_kernel<<<dim3(s.nb, F * (n - 1) * B), kThreads, 0, ST>>>((const float2 *)packed_flow, scratch, s);
    temp_finalize_kernel<<<(B + 63) / 64, 64, 0, ST>>>(scratch, sums, out, s);
    return (int)cudaGetLastError();
}

extern "C" int tef_flow_temporal_smoothing_bwd(const float *packed_flow, const float *sums, const float *gout, float *packed_grad, int B, int H, int W,
                                               int P, int F, int n, void *stream) {
    SmoothGeom s;
    int rc = smooth_geom(B, H, W, P, F, n, s);
    if (rc) return rc;
    if (n < 2) return TEF_EINVAL;
    if (!packed_flow || !sums || !gout || !packed_grad) return TEF_EINVAL;
    cudaMemsetAsync(packed_grad, 0, sizeof(float2) * (long)F * P * B * 2 * s.g.plane, ST);
    ProfScope pr(K_SMOOTH, ST);
    temp_bwd_kernel<<<dim3(s.nb, F * (n - 1) * B), kThreads, 0, ST>>>((const float2 *)packed_flow, sums, gout, (float2 *)packed_grad, s);
    return (int)cudaGetLastError();
}
