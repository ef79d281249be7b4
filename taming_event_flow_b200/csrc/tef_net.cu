// tef_net.cu -- the element-wise stages of the recurrent flow network's training step, fused.
//
// The network itself stays on PyTorch/cuDNN convolutions (north_star).  What surrounds them in upstream's step is a swarm
// of tiny element-wise kernels -- profiles/r2_f_train_kernels_f32.txt: of the 4 913 launches of a graphed training step,
// 1 284 are element-wise adds, 240 are bias-gradient reductions (12 us each), 160 are sigmoid / tanh and their derivatives,
// plus the multiplications and concatenations between them.  The kernels here replace them:
//
//   ConvGRU (models/submodules.py:134-152): between its two convolutions (update|reset gates merged into one, and the
//       candidate) the cell is two forward and three backward kernels, bias gradients included;
//   conv + bias + activation (models/submodules.py ConvLayer / ResidualBlock): bias, optional residual and ReLU / tanh in one
//       kernel, the activation derivative and the bias gradient in another;
//   flow head up-sampling (models/model.py:65-85, train_flow.py:106-108): bilinear up-sampling of the 2-channel
//       prediction to the input size and both flow scalings in one kernel each way, straight into the planar [B][2][H][W]
//       layout tef_update_pass packs from.
//
// Layout: NHWC ("channels_last") fp32, M = B*H*W pixel rows of C channels, C a multiple of 4 (16-byte accesses).
// Column sums (bias gradients) are accumulated per thread over a strided set of rows, combined per CTA in shared memory and
// added to the output with one fp32 reduction per column and CTA (order-free up to fp32 rounding, like any atomic sum).
#include "../../include/tef_b200.h"
#include "tef_prof.cuh"
#include <cuda_runtime.h>

namespace tef {

constexpr int kNetThreads = 256;

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }
#define TEF_MAP4(out, expr) do { float4 o_; { const int k = 0; o_.x = (expr); } { const int k = 1; o_.y = (expr); } \
                                 { const int k = 2; o_.z = (expr); } { const int k = 3; o_.w = (expr); } (out) = o_; } while (0)
__device__ __forceinline__ float el(const float4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// Geometry of a pass over [M][C4] float4 columns with per-column sums: cpb column groups x rpb rows per CTA iteration.
struct ColGeom { int cpb, rpb, nblk; };
static ColGeom col_geom(long M, int C4) {
    ColGeom g;
    g.cpb = 1;
    while (g.cpb < C4 && g.cpb < kNetThreads) g.cpb <<= 1;
    g.rpb = kNetThreads / g.cpb;
    // Every CTA ends with one reduction per column group into the same few cache lines of the bias gradient, and reductions
    // into one line serialise in the L2 (measured: 256 CTAs x 512 scalar atomics = 14 us for a 1 MB tensor).  So a thread
    // walks 8-16 rows before it flushes, and the flush is one 16-byte reduction per column group.
    const long rows_per_thread = (M * C4 < (1l << 20)) ? 8 : 16;
    long nb = (M + g.rpb * rows_per_thread - 1) / (g.rpb * rows_per_thread);
    if (nb > 148 * 4) nb = 148 * 4;
    g.nblk = (int)(nb < 1 ? 1 : nb);
    return g;
}
// combine the per-thread partial sums of one column group over the rpb row lanes of the CTA; one reduction per column
template <int NS>
__device__ __forceinline__ void col_flush(float4 (&s)[NS], float *const (&dst)[NS], int cg, int C4, int cpb, int rpb) {
    __shared__ float4 sm[kNetThreads];
    const int rl = threadIdx.x / cpb;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (!dst[i]) continue;                                   // uniform
        __syncthreads();
        sm[threadIdx.x] = s[i];
        __syncthreads();
        if (rl == 0 && cg < C4) {
            float4 a = s[i];
            for (int r = 1; r < rpb; ++r) { const float4 b = sm[r * cpb + (threadIdx.x % cpb)]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst[i] + 4 * cg), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
        }
    }
}

// ---- ConvGRU ---------------------------------------------------------------------------------------------
// zr [M][2C]: conv output -> (z, r) = sigmoid(. + bias), in place;  xh [M][Cx+C] = (x, h);  xrh [M][Cx+C] = (x, h*r)
__global__ void __launch_bounds__(kNetThreads) gru_gates_kernel(float *__restrict__ zr, const float *__restrict__ bias, const float *__restrict__ xh,
                                                                float *__restrict__ xrh, long M, int Cx, int C) {
    const int Cx4 = Cx >> 2, C4 = C >> 2, W4 = Cx4 + C4;
    const long n = M * W4;
    for (long i = (long)blockIdx.x * kNetThreads + threadIdx.x; i < n; i += (long)gridDim.x * kNetThreads) {
        const long m = i / W4;
        const int c = (int)(i - m * W4);
        if (c < Cx4) { st4(xrh + m * (Cx + C) + 4 * c, ld4(xh + m * (Cx + C) + 4 * c)); continue; }
        const int h4 = 4 * (c - Cx4);
        float4 z = ld4(zr + m * 2 * C + h4), r = ld4(zr + m * 2 * C + C + h4);
        if (bias) {
            const float4 bz = ld4(bias + h4), br = ld4(bias + C + h4);
            z.x += bz.x; z.y += bz.y; z.z += bz.z; z.w += bz.w; r.x += br.x; r.y += br.y; r.z += br.z; r.w += br.w;
        }
        TEF_MAP4(z, sigm(el(z, k)));
        TEF_MAP4(r, sigm(el(r, k)));
        const float4 h = ld4(xh + m * (Cx + C) + Cx + h4);
        st4(zr + m * 2 * C + h4, z);
        st4(zr + m * 2 * C + C + h4, r);
        st4(xrh + m * (Cx + C) + Cx + h4, make_float4(h.x * r.x, h.y * r.y, h.z * r.z, h.w * r.w));
    }
}
// c [M][C]: conv output -> cand = tanh(. + bias), in place;  out = h * (1 - z) + cand * z   (models/submodules.py:149-150)
__global__ void __launch_bounds__(kNetThreads) gru_output_kernel(float *__restrict__ c, const float *__restrict__ bias, const float *__restrict__ xh,
                                                                 const float *__restrict__ zr, float *__restrict__ out, long M, int Cx, int C) {
    const int C4 = C >> 2;
    const long n = M * C4;
    for (long i = (long)blockIdx.x * kNetThreads + threadIdx.x; i < n; i += (long)gridDim.x * kNetThreads) {
        const long m = i / C4;
        const int h4 = 4 * (int)(i - m * C4);
        float4 a = ld4(c + m * C + h4);
        if (bias) { const float4 b = ld4(bias + h4); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        TEF_MAP4(a, tanhf(el(a, k)));
        const float4 z = ld4(zr + m * 2 * C + h4), h = ld4(xh + m * (Cx + C) + Cx + h4);
        float4 o;
        TEF_MAP4(o, el(h, k) * (1.0f - el(z, k)) + el(a, k) * el(z, k));
        st4(c + m * C + h4, a);
        st4(out + m * C + h4, o);
    }
}
// reverse of gru_output_kernel: gc = g z (1 - cand^2), gzr[:, :C] = g (cand - h) z (1 - z), gh = g (1 - z); column sums of the first two
__global__ void __launch_bounds__(kNetThreads) gru_output_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ cand, const float *__restrict__ xh,
                                                                     const float *__restrict__ zr, float *__restrict__ gc, float *__restrict__ gzr,
                                                                     float *__restrict__ gh, float *gbias_c, float *gbias_zr, long M, int Cx, int C, int cpb, int rpb) {
    const int C4 = C >> 2;
    const int rl = threadIdx.x / cpb;
    for (int cg = threadIdx.x % cpb; cg - (int)(threadIdx.x % cpb) < C4; cg += cpb) {          // uniform trip count
        float4 s[2] = { make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f) };
        if (cg < C4)
            for (long m = (long)blockIdx.x * rpb + rl; m < M; m += (long)gridDim.x * rpb) {
                const int h4 = 4 * cg;
                const float4 g = ld4(gout + m * C + h4), a = ld4(cand + m * C + h4), z = ld4(zr + m * 2 * C + h4), h = ld4(xh + m * (Cx + C) + Cx + h4);
                float4 vc, vz, vh;
                TEF_MAP4(vc, el(g, k) * el(z, k) * (1.0f - el(a, k) * el(a, k)));
                TEF_MAP4(vz, el(g, k) * (el(a, k) - el(h, k)) * (el(z, k) * (1.0f - el(z, k))));
                TEF_MAP4(vh, el(g, k) * (1.0f - el(z, k)));
                st4(gc + m * C + h4, vc);
                st4(gzr + m * 2 * C + h4, vz);
                st4(gh + m * C + h4, vh);
                s[0].x += vc.x; s[0].y += vc.y; s[0].z += vc.z; s[0].w += vc.w;
                s[1].x += vz.x; s[1].y += vz.y; s[1].z += vz.z; s[1].w += vz.w;
            }
        float *const dst[2] = { gbias_c, gbias_zr };
        col_flush<2>(s, dst, cg, C4, cpb, rpb);
    }
}
// reverse of gru_gates_kernel: ghr = gxrh[:, Cx:];  gh += ghr r;  gzr[:, C:] = ghr h r (1 - r), with its column sums
__global__ void __launch_bounds__(kNetThreads) gru_gates_bwd_kernel(const float *__restrict__ gxrh, const float *__restrict__ xh, const float *__restrict__ zr,
                                                                    float *__restrict__ gzr, float *__restrict__ gh, float *gbias_zr, long M, int Cx, int C,
                                                                    int cpb, int rpb) {
    const int C4 = C >> 2;
    const int rl = threadIdx.x / cpb;
    for (int cg = threadIdx.x % cpb; cg - (int)(threadIdx.x % cpb) < C4; cg += cpb) {
        float4 s[1] = { make_float4(0.f, 0.f, 0.f, 0.f) };
        if (cg < C4)
            for (long m = (long)blockIdx.x * rpb + rl; m < M; m += (long)gridDim.x * rpb) {
                const int h4 = 4 * cg;
                const float4 g = ld4(gxrh + m * (Cx + C) + Cx + h4), r = ld4(zr + m * 2 * C + C + h4), h = ld4(xh + m * (Cx + C) + Cx + h4);
                float4 acc = ld4(gh + m * C + h4), vr;
                TEF_MAP4(acc, el(acc, k) + el(g, k) * el(r, k));
                TEF_MAP4(vr, el(g, k) * el(h, k) * (el(r, k) * (1.0f - el(r, k))));
                st4(gh + m * C + h4, acc);
                st4(gzr + m * 2 * C + C + h4, vr);
                s[0].x += vr.x; s[0].y += vr.y; s[0].z += vr.z; s[0].w += vr.w;
            }
        float *const dst[1] = { gbias_zr ? gbias_zr + C : nullptr };
        col_flush<1>(s, dst, cg, C4, cpb, rpb);
    }
}
// gx = gxrh[:, :Cx] + gxh[:, :Cx];  gh += gxh[:, Cx:]
__global__ void __launch_bounds__(kNetThreads) gru_input_grads_kernel(const float *__restrict__ gxrh, const float *__restrict__ gxh, float *__restrict__ gx,
                                                                      float *__restrict__ gh, long M, int Cx, int C) {
    const int Cx4 = Cx >> 2, C4 = C >> 2, W4 = Cx4 + C4;
    const long n = M * W4;
    for (long i = (long)blockIdx.x * kNetThreads + threadIdx.x; i < n; i += (long)gridDim.x * kNetThreads) {
        const long m = i / W4;
        const int c = (int)(i - m * W4);
        const float4 a = ld4(gxh + m * (Cx + C) + 4 * c);
        if (c < Cx4) {
            const float4 b = ld4(gxrh + m * (Cx + C) + 4 * c);
            st4(gx + m * Cx + 4 * c, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
        } else {
            float *p = gh + m * C + 4 * (c - Cx4);
            const float4 b = ld4(p);
            st4(p, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
        }
    }
}

// ---- conv + bias (+ residual) + activation ----------------------------------------------------------------
// act: 0 none, 1 ReLU, 2 tanh.  y [M][C] in place.
__global__ void __launch_bounds__(kNetThreads) bias_act_kernel(float *__restrict__ y, const float *__restrict__ bias, const float *__restrict__ res, int act,
                                                               long M, int C) {
    const int C4 = C >> 2;
    const long n = M * C4;
    for (long i = (long)blockIdx.x * kNetThreads + threadIdx.x; i < n; i += (long)gridDim.x * kNetThreads) {
        const int h4 = 4 * (int)(i % C4);
        float4 a = ld4(y + 4 * i);
        if (bias) { const float4 b = ld4(bias + h4); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        if (res) { const float4 b = ld4(res + 4 * i); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        if (act == 1) TEF_MAP4(a, fmaxf(el(a, k), 0.0f));
        else if (act == 2) TEF_MAP4(a, tanhf(el(a, k)));
        st4(y + 4 * i, a);
    }
}
// gpre = gy * act'(y) (from the activated output), column sums -> gbias
__global__ void __launch_bounds__(kNetThreads) bias_act_bwd_kernel(const float *__restrict__ gy, const float *__restrict__ y, float *__restrict__ gpre,
                                                                   float *gbias, int act, long M, int C, int cpb, int rpb) {
    const int C4 = C >> 2;
    const int rl = threadIdx.x / cpb;
    for (int cg = threadIdx.x % cpb; cg - (int)(threadIdx.x % cpb) < C4; cg += cpb) {
        float4 s[1] = { make_float4(0.f, 0.f, 0.f, 0.f) };
        if (cg < C4)
            for (long m = (long)blockIdx.x * rpb + rl; m < M; m += (long)gridDim.x * rpb) {
                const long o = m * C + 4 * cg;
                float4 g = ld4(gy + o);
                if (act) {
                    const float4 v = ld4(y + o);
                    if (act == 1) TEF_MAP4(g, el(v, k) > 0.0f ? el(g, k) : 0.0f);
                    else TEF_MAP4(g, el(g, k) * (1.0f - el(v, k) * el(v, k)));
                }
                if (gpre != gy || act) st4(gpre + o, g);
                s[0].x += g.x; s[0].y += g.y; s[0].z += g.z; s[0].w += g.w;
            }
        float *const dst[1] = { gbias };
        col_flush<1>(s, dst, cg, C4, cpb, rpb);
    }
}

// ---- flow head: bilinear up-sampling (align_corners = False, like F.interpolate) x scale ----------------------------
// ATen's source index: max(0, (dst + 0.5) * in / out - 0.5), neighbour clamped at the border (UpSample.cuh)
struct UpIdx { int i0, i1; float l0, l1; };
__device__ __forceinline__ UpIdx up_index(int d, float ratio, int in) {
    float s = ratio * ((float)d + 0.5f) - 0.5f;
    if (s < 0.0f) s = 0.0f;
    UpIdx u;
    u.i0 = (int)s; if (u.i0 > in - 1) u.i0 = in - 1;
    u.i1 = u.i0 + (u.i0 < in - 1 ? 1 : 0);
    u.l1 = s - (float)u.i0; u.l0 = 1.0f - u.l1;
    return u;
}
// pred: [B][2][h][w] through element strides (sb, sc, sy, sx);  out: [B][2][H][W] contiguous
__global__ void __launch_bounds__(kNetThreads) upsample_scale_kernel(const float *__restrict__ pred, long sb, long sc, long sy, long sx, int h, int w,
                                                                     float *__restrict__ out, int B, int H, int W, float scale) {
    const long n = (long)B * 2 * H * W;
    const float rh = (float)h / (float)H, rw = (float)w / (float)W;
    for (long i = (long)blockIdx.x * kNetThreads + threadIdx.x; i < n; i += (long)gridDim.x * kNetThreads) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const long bc = i / ((long)W * H);
        const float *p = pred + (bc >> 1) * sb + (bc & 1) * sc;
        const UpIdx uy = up_index(y, rh, h), ux = up_index(x, rw, w);
        const float v00 = __ldg(p + uy.i0 * sy + ux.i0 * sx), v01 = __ldg(p + uy.i0 * sy + ux.i1 * sx);
        const float v10 = __ldg(p + uy.i1 * sy + ux.i0 * sx), v11 = __ldg(p + uy.i1 * sy + ux.i1 * sx);
        out[i] = (uy.l0 * (ux.l0 * v00 + ux.l1 * v01) + uy.l1 * (ux.l0 * v10 + ux.l1 * v11)) * scale;
    }
}
// its adjoint, as a gather (deterministic): LPE lanes per element of the low-resolution gradient walk the window of output
// pixels that can reference it and re-derive their weights (LPE = 1 for an identity-sized head, 32 for the x8 one)
template <int LPE>
__global__ void __launch_bounds__(kNetThreads) upsample_scale_bwd_kernel(const float *__restrict__ gout, int B, int H, int W, float scale, float *__restrict__ gpred,
                                                                         long sb, long sc, long sy, long sx, int h, int w) {
    const long el_ = ((long)blockIdx.x * kNetThreads + threadIdx.x) / LPE;
    const int lane = threadIdx.x % LPE;
    const long n = (long)B * 2 * h * w;
    const bool live = el_ < n;
    const long e = live ? el_ : 0;
    const int j = (int)(e % w), i = (int)((e / w) % h);
    const long bc = e / ((long)w * h);
    const float rh = (float)h / (float)H, rw = (float)w / (float)W;
    // output rows whose source index lies in (i - 1, i + 1): (y + 0.5) * rh - 0.5 in that range, widened by one for rounding
    const int y_lo = max(0, (int)floorf(((float)i - 0.5f) / rh - 0.5f) - 1), y_hi = min(H - 1, (int)ceilf(((float)i + 1.5f) / rh - 0.5f) + 1);
    const int x_lo = max(0, (int)floorf(((float)j - 0.5f) / rw - 0.5f) - 1), x_hi = min(W - 1, (int)ceilf(((float)j + 1.5f) / rw - 0.5f) + 1);
    const int nx = x_hi - x_lo + 1, ny = y_hi - y_lo + 1;
    const float *g = gout + bc * (long)H * W;
    float acc = 0.0f;
    if (live)
        for (int k = lane; k < nx * ny; k += LPE) {
            const int y = y_lo + k / nx, x = x_lo + k % nx;
            const UpIdx uy = up_index(y, rh, h), ux = up_index(x, rw, w);
            const float wy = (uy.i0 == i ? uy.l0 : 0.0f) + (uy.i1 == i ? uy.l1 : 0.0f);
            const float wx = (ux.i0 == j ? ux.l0 : 0.0f) + (ux.i1 == j ? ux.l1 : 0.0f);
            if (wy != 0.0f && wx != 0.0f) acc += wy * wx * __ldg(g + (long)y * W + x);
        }
#pragma unroll
    for (int o = LPE / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && lane == 0) gpred[(bc >> 1) * sb + (bc & 1) * sc + i * sy + j * sx] = acc * scale;
}

// ---- decoder stage input (models/arch.py decoder loop): (x + skip), the previous flow prediction concatenated in front, and the
// bilinear x2 up-sampling that precedes the decoder's convolution -- one kernel instead of add, cat and up-sample ----------------
// x, skip: [B][h][w][C] (NHWC);  pred: [B][2][h][w] through element strides, or NULL;  out: [B][H][W][Cp], Cp = C + (pred ? 2 : 0).
// Channels travel in PAIRS (C is even; Cp = 66, 130, 258 are not multiples of four).  Generic scale: one CTA per output row
// (b, Y), 32 lanes over the channel pairs of a pixel (coalesced 256-byte runs), 8 pixel lanes along the row, the two prediction
// channels by one thread per pixel.  Every decoder stage is exactly x2 and takes decoder_up2_kernel below instead.
__global__ void __launch_bounds__(kNetThreads) decoder_up_kernel(const float *__restrict__ x, const float *__restrict__ skip, const float *__restrict__ pred,
                                                                 long sb, long sc, long sy, long sx, float *__restrict__ out, int B, int h, int w, int C,
                                                                 int H, int W) {
    const int np = pred ? 1 : 0, Cp2 = (C >> 1) + np;
    const float rh = (float)h / (float)H, rw = (float)w / (float)W;
    const int lane = threadIdx.x & 31, pl = threadIdx.x >> 5;
    for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
        const int b = row / H, Y = row - b * H;
        const UpIdx uy = up_index(Y, rh, h);
        const float *x0 = x + ((long)b * h + uy.i0) * w * C, *x1 = x + ((long)b * h + uy.i1) * w * C;
        const float *s0 = skip ? skip + ((long)b * h + uy.i0) * w * C : nullptr, *s1 = skip ? skip + ((long)b * h + uy.i1) * w * C : nullptr;
        float *orow = out + (long)row * W * 2 * Cp2;
        for (int X = pl; X < W; X += kNetThreads / 32) {
            const UpIdx ux = up_index(X, rw, w);
            float *o = orow + (long)X * 2 * Cp2 + 2 * np;
            for (int c = 2 * lane; c < C; c += 64) {                 // C / 2 pairs: 32, 64, 128, 256 -- whole warps
                const int a0 = ux.i0 * C + c, a1 = ux.i1 * C + c;
                float2 v00 = *reinterpret_cast<const float2 *>(x0 + a0), v01 = *reinterpret_cast<const float2 *>(x0 + a1);
                float2 v10 = *reinterpret_cast<const float2 *>(x1 + a0), v11 = *reinterpret_cast<const float2 *>(x1 + a1);
                if (skip) {
                    const float2 t00 = *reinterpret_cast<const float2 *>(s0 + a0), t01 = *reinterpret_cast<const float2 *>(s0 + a1);
                    const float2 t10 = *reinterpret_cast<const float2 *>(s1 + a0), t11 = *reinterpret_cast<const float2 *>(s1 + a1);
                    v00.x += t00.x; v00.y += t00.y; v01.x += t01.x; v01.y += t01.y; v10.x += t10.x; v10.y += t10.y; v11.x += t11.x; v11.y += t11.y;
                }
                // same association as ATen: l0h * (l0w * a + l1w * b) + l1h * (l0w * c + l1w * d)
                float2 r;
                r.x = uy.l0 * (ux.l0 * v00.x + ux.l1 * v01.x) + uy.l1 * (ux.l0 * v10.x + ux.l1 * v11.x);
                r.y = uy.l0 * (ux.l0 * v00.y + ux.l1 * v01.y) + uy.l1 * (ux.l0 * v10.y + ux.l1 * v11.y);
                *reinterpret_cast<float2 *>(o + c) = r;
            }
        }
        if (np)                                                      // the two prediction channels in front: one thread per pixel
            for (int X = threadIdx.x; X < W; X += kNetThreads) {
                const UpIdx ux = up_index(X, rw, w);
                const float *p = pred + b * sb;
                const long o00 = uy.i0 * sy + ux.i0 * sx, o01 = uy.i0 * sy + ux.i1 * sx, o10 = uy.i1 * sy + ux.i0 * sx, o11 = uy.i1 * sy + ux.i1 * sx;
                float2 r;
                r.x = uy.l0 * (ux.l0 * __ldg(p + o00) + ux.l1 * __ldg(p + o01)) + uy.l1 * (ux.l0 * __ldg(p + o10) + ux.l1 * __ldg(p + o11));
                r.y = uy.l0 * (ux.l0 * __ldg(p + o00 + sc) + ux.l1 * __ldg(p + o01 + sc)) + uy.l1 * (ux.l0 * __ldg(p + o10 + sc) + ux.l1 * __ldg(p + o11 + sc));
                *reinterpret_cast<float2 *>(orow + (long)X * 2 * Cp2) = r;
            }
    }
}
// Exactly x2 (H = 2h, W = 2w; every decoder stage): output pixels (2i+1, 2i+2) x (2j+1, 2j+2) interpolate the same four input
// pixels (i, i+1) x (j, j+1), so one thread per input CELL (i, j in -1 .. size-1, clamped at the border) and channel pair loads
// the taps once and writes up to four outputs: 2 loads per output instead of 8.  Weights come from up_index like the generic kernel.
__global__ void __launch_bounds__(kNetThreads) decoder_up2_kernel(const float *__restrict__ x, const float *__restrict__ skip, const float *__restrict__ pred,
                                                                  long sb, long sc, long sy, long sx, float *__restrict__ out, int B, int h, int w, int C) {
    const int np = pred ? 1 : 0, Cp2 = (C >> 1) + np, Cp = 2 * Cp2, H = 2 * h, W = 2 * w;
    const unsigned n_row = (unsigned)(w + 1) * (unsigned)Cp2;             // work items of one cell row
    const int b = blockIdx.z, i = (int)blockIdx.y - 1;
    const int r0 = max(i, 0), r1 = min(i + 1, h - 1);
    for (unsigned k = blockIdx.x * kNetThreads + threadIdx.x; k < n_row; k += gridDim.x * kNetThreads) {
        const int j = (int)(k / (unsigned)Cp2) - 1, c2 = (int)(k % (unsigned)Cp2);
        const int q0 = max(j, 0), q1 = min(j + 1, w - 1);
        float2 v00, v01, v10, v11;
        if (c2 < np) {
            const float *p = pred + b * sb;
            const long o00 = r0 * sy + q0 * sx, o01 = r0 * sy + q1 * sx, o10 = r1 * sy + q0 * sx, o11 = r1 * sy + q1 * sx;
            v00 = make_float2(__ldg(p + o00), __ldg(p + o00 + sc)); v01 = make_float2(__ldg(p + o01), __ldg(p + o01 + sc));
            v10 = make_float2(__ldg(p + o10), __ldg(p + o10 + sc)); v11 = make_float2(__ldg(p + o11), __ldg(p + o11 + sc));
        } else {
            const int c = 2 * (c2 - np);
            const long a00 = (((long)b * h + r0) * w + q0) * C + c, a01 = (((long)b * h + r0) * w + q1) * C + c;
            const long a10 = (((long)b * h + r1) * w + q0) * C + c, a11 = (((long)b * h + r1) * w + q1) * C + c;
            v00 = *reinterpret_cast<const float2 *>(x + a00); v01 = *reinterpret_cast<const float2 *>(x + a01);
            v10 = *reinterpret_cast<const float2 *>(x + a10); v11 = *reinterpret_cast<const float2 *>(x + a11);
            if (skip) {
                const float2 t00 = *reinterpret_cast<const float2 *>(skip + a00), t01 = *reinterpret_cast<const float2 *>(skip + a01);
                const float2 t10 = *reinterpret_cast<const float2 *>(skip + a10), t11 = *reinterpret_cast<const float2 *>(skip + a11);
                v00.x += t00.x; v00.y += t00.y; v01.x += t01.x; v01.y += t01.y; v10.x += t10.x; v10.y += t10.y; v11.x += t11.x; v11.y += t11.y;
            }
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int Y = 2 * i + 1 + a;
            if (Y < 0 || Y >= H) continue;
            const UpIdx uy = up_index(Y, 0.5f, h);                        // rows (uy.i0, uy.i1) == (r0, r1) up to equal clamped rows
            const float2 t0 = uy.i0 == r0 ? v00 : v10, t1 = uy.i0 == r0 ? v01 : v11;      // row i0
            const float2 u0 = uy.i1 == r1 ? v10 : v00, u1 = uy.i1 == r1 ? v11 : v01;      // row i1
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int X = 2 * j + 1 + e;
                if (X < 0 || X >= W) continue;
                const UpIdx ux = up_index(X, 0.5f, w);
                const bool l = ux.i0 == q0, rr = ux.i1 == q1;
                const float2 a0 = l ? t0 : t1, a1 = rr ? t1 : t0, b0 = l ? u0 : u1, b1 = rr ? u1 : u0;
                float2 r;
                r.x = uy.l0 * (ux.l0 * a0.x + ux.l1 * a1.x) + uy.l1 * (ux.l0 * b0.x + ux.l1 * b1.x);
                r.y = uy.l0 * (ux.l0 * a0.y + ux.l1 * a1.y) + uy.l1 * (ux.l0 * b0.y + ux.l1 * b1.y);
                *reinterpret_cast<float2 *>(out + (((long)b * H + Y) * W + X) * Cp + 2 * c2) = r;
            }
        }
    }
}

// adjoint (deterministic gather): one thread per low-resolution pixel and channel pair; the 1-D weights of the candidate output
// rows / columns are derived once, then the non-zero ones are walked.  gx receives the gradient of x and of skip alike.
constexpr int kUpWin = 8;          // candidate output rows per input row: 2 * scale + 3 <= 8 covers scale factors up to 2.5
__global__ void __launch_bounds__(kNetThreads) decoder_up_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gx, float *__restrict__ gpred,
                                                                     long sb, long sc, long sy, long sx, int B, int h, int w, int C, int H, int W) {
    const int np = gpred ? 1 : 0, Cp2 = (C >> 1) + np, Cp = 2 * Cp2;
    const long n = (long)B * h * w * Cp2;
    const float rh = (float)h / (float)H, rw = (float)w / (float)W;
    for (long idx = (long)blockIdx.x * kNetThreads + threadIdx.x; idx < n; idx += (long)gridDim.x * kNetThreads) {
        const int c2 = (int)(idx % Cp2);
        const long pix = idx / Cp2;
        const int j = (int)(pix % w), i = (int)((pix / w) % h);
        const long b = pix / ((long)w * h);
        const int y_lo = max(0, (int)floorf(((float)i - 0.5f) / rh - 0.5f) - 1), x_lo = max(0, (int)floorf(((float)j - 0.5f) / rw - 0.5f) - 1);
        float wy[kUpWin], wx[kUpWin];
#pragma unroll
        for (int k = 0; k < kUpWin; ++k) {
            const int y = y_lo + k, xx = x_lo + k;
            wy[k] = 0.0f; wx[k] = 0.0f;
            if (y < H) { const UpIdx u = up_index(y, rh, h); wy[k] = (u.i0 == i ? u.l0 : 0.0f) + (u.i1 == i ? u.l1 : 0.0f); }
            if (xx < W) { const UpIdx u = up_index(xx, rw, w); wx[k] = (u.i0 == j ? u.l0 : 0.0f) + (u.i1 == j ? u.l1 : 0.0f); }
        }
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < kUpWin; ++ky) {
            if (wy[ky] == 0.0f) continue;
#pragma unroll
            for (int kx = 0; kx < kUpWin; ++kx) {
                if (wx[kx] == 0.0f) continue;
                const float2 g = *reinterpret_cast<const float2 *>(gout + ((b * H + y_lo + ky) * W + x_lo + kx) * Cp + 2 * c2);
                const float wgt = wy[ky] * wx[kx];
                acc.x += wgt * g.x; acc.y += wgt * g.y;
            }
        }
        if (c2 < np) {
            float *p = gpred + b * sb + (long)i * sy + (long)j * sx;
            p[0] = acc.x; p[sc] = acc.y;
        } else {
            *reinterpret_cast<float2 *>(gx + ((b * h + i) * w + j) * C + 2 * (c2 - np)) = acc;
        }
    }
}

__global__ void __launch_bounds__(kNetThreads) scale_copy_strided_kernel(const float *__restrict__ g, float scale, float *__restrict__ out, long sb, long sc, long sy,
                                                                         long sx, int B, int h, int w) {
    const long n = (long)B * 2 * h * w;
    for (long i = (long)blockIdx.x * kNetThreads + threadIdx.x; i < n; i += (long)gridDim.x * kNetThreads) {
        const int j = (int)(i % w), r = (int)((i / w) % h);
        const long bc = i / ((long)w * h);
        out[(bc >> 1) * sb + (bc & 1) * sc + r * sy + j * sx] = g[i] * scale;
    }
}

static int grid_for(long n) {
    long nb = (n + kNetThreads - 1) / kNetThreads;
    if (nb > 148 * 8) nb = 148 * 8;
    return (int)(nb < 1 ? 1 : nb);
}
static bool bad_c(int C) { return C < 4 || (C & 3); }

}  // namespace tef

using namespace tef;

extern "C" int tef_gru_gates(float *zr, const float *bias_zr, const float *xh, float *xrh, long M, int Cx, int C, void *stream) {
    if (!zr || !xh || !xrh || M < 0 || bad_c(C) || bad_c(Cx)) return TEF_EINVAL;
    if (M == 0) return 0;
    ProfScope ps(K_NETWORK, (cudaStream_t)stream);
    gru_gates_kernel<<<grid_for(M * ((Cx + C) >> 2)), kNetThreads, 0, (cudaStream_t)stream>>>(zr, bias_zr, xh, xrh, M, Cx, C);
    return (int)cudaGetLastError();
}
extern "C" int tef_gru_output(float *c, const float *bias_c, const float *xh, const float *zr, float *out, long M, int Cx, int C, void *stream) {
    if (!c || !xh || !zr || !out || M < 0 || bad_c(C) || bad_c(Cx)) return TEF_EINVAL;
    if (M == 0) return 0;
    ProfScope ps(K_NETWORK, (cudaStream_t)stream);
    gru_output_kernel<<<grid_for(M * (C >> 2)), kNetThreads, 0, (cudaStream_t)stream>>>(c, bias_c, xh, zr, out, M, Cx, C);
    return (int)cudaGetLastError();
}
extern "C" int tef_gru_output_bwd(const float *gout, const float *cand, const float *xh, const float *zr, float *gc, float *gzr, float *gh, float *gbias_c,
                                  float *gbias_zr, long M, int Cx, int C, void *stream) {
    if (!gout || !cand || !xh || !zr || !gc || !gzr || !gh || M < 0 || bad_c(C) || bad_c(Cx)) return TEF_EINVAL;
    if (M == 0) return 0;
    const ColGeom g = col_geom(M, C >> 2);
    ProfScope ps(K_NETWORK, (cudaStream_t)stream);
    gru_output_bwd_kernel<<<g.nblk, kNetThreads, 0, (cudaStream_t)stream>>>(gout, cand, xh, zr, gc, gzr, gh, gbias_c, gbias_zr, M, Cx, C, g.cpb, g.rpb);
    return (int)cudaGetLastError();
}
extern "C" int tef_gru_gates_bwd(const float *gxrh, const float *xh, const float *zr, float *gzr, float *gh, float *gbias_zr, long M, int Cx, int C,
                                 void *stream) {
    if (!gxrh || !xh || !zr || !gzr || !gh || M < 0 || bad_c(C) || bad_c(Cx)) return TEF_EINVAL;
    if (M == 0) return 0;
    const ColGeom g = col_geom(M, C >> 2);
    ProfScope ps(K_NETWORK, (cudaStream_t)stream);
    gru_gates_bwd_kernel<<<g.nblk, kNetThreads, 0, (cudaStream_t)stream>>>(gxrh, xh, zr, gzr, gh, gbias_zr, M, Cx, C, g.cpb, g.rpb);
    return (int)cudaGetLastError();
}
extern "C" int tef_gru_input_grads(const float *gxrh, const float *gxh, float *gx, float *gh, long M, int Cx, int C, void *stream) {
    if (!gxrh || !gxh || !gx || !gh || M < 0 || bad_c(C) || bad_c(Cx)) return TEF_EINVAL;
    if (M == 0) return 0;
    ProfScope ps(K_NETWORK, (cudaStream_t)stream);
    gru_input_grads_kernel<<<grid_for(M * ((Cx + C) >> 2)), kNetThreads, 0, (cudaStream_t)stream>>>(gxrh, gxh, gx, gh, M, Cx, C);
    return (int)cudaGetLastError();
}
extern "C" int tef_bias_act(float *y, const float *bias, const float *residual, int act, long M, int C, void *stream) {
    if (!y || M < 0 || bad_c(C) || act < 0 || act > 2) return TEF_EINVAL;
    if (M == 0) return 0;
    ProfScope ps(K_NETWORK, (cudaStream_t)stream);
    bias_act_kernel<<<grid_for(M * (C >> 2)), kNetThreads, 0, (cudaStream_t)stream>>>(y, bias, residual, act, M, C);
    return (int)cudaGetLastError();
}
extern "C" int tef_bias_act_bwd(const float *gy, const float *y, float *gpre, float *gbias, int act, long M, int C, void *stream) {
    if (!gy || !gpre || (act && !y) || M < 0 || bad_c(C) || act < 0 || act > 2) return TEF_EINVAL;
    if (M == 0) return 0;
    const ColGeom g = col_geom(M, C >> 2);
    ProfScope ps(K_NETWORK, (cudaStream_t)stream);
    bias_act_bwd_kernel<<<g.nblk, kNetThreads, 0, (cudaStream_t)stream>>>(gy, y, gpre, gbias, act, M, C, g.cpb, g.rpb);
    return (int)cudaGetLastError();
}
extern "C" int tef_upsample_scale(const float *pred, const long *strides, int h, int w, float *out, int B, int H, int W, float scale, void *stream) {
    if (!pred || !strides || !out || B < 1 || h < 1 || w < 1 || H < 1 || W < 1) return TEF_EINVAL;
    ProfScope ps(K_NETWORK, (cudaStream_t)stream);
    upsample_scale_kernel<<<grid_for((long)B * 2 * H * W), kNetThreads, 0, (cudaStream_t)stream>>>(pred, strides[0], strides[1], strides[2], strides[3], h, w,
                                                                                                  out, B, H, W, scale);
    return (int)cudaGetLastError();
}
extern "C" int tef_upsample_scale_bwd(const float *gout, int B, int H, int W, float scale, float *gpred, const long *strides, int h, int w, void *stream) {
    if (!gout || !gpred || !strides || B < 1 || h < 1 || w < 1 || H < 1 || W < 1) return TEF_EINVAL;
    const long n = (long)B * 2 * h * w;
    if (h == H && w == W) {                                 // identity-sized head: source index == output index, weights (1, 0)
        ProfScope ps(K_NETWORK, (cudaStream_t)stream);
        scale_copy_strided_kernel<<<grid_for(n), kNetThreads, 0, (cudaStream_t)stream>>>(gout, scale, gpred, strides[0], strides[1], strides[2], strides[3], B, h, w);
        return (int)cudaGetLastError();
    }
    const float win = ((float)H / (float)h * 2.0f + 3.0f) * ((float)W / (float)w * 2.0f + 3.0f);      // output pixels an element looks at
    const int lpe = win <= 40.0f ? 1 : (win <= 160.0f ? 8 : 32);
    const unsigned nb = (unsigned)((n * lpe + kNetThreads - 1) / kNetThreads);
    const cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(K_NETWORK, st);
    if (lpe == 1) upsample_scale_bwd_kernel<1><<<nb, kNetThreads, 0, st>>>(gout, B, H, W, scale, gpred, strides[0], strides[1], strides[2], strides[3], h, w);
    else if (lpe == 8) upsample_scale_bwd_kernel<8><<<nb, kNetThreads, 0, st>>>(gout, B, H, W, scale, gpred, strides[0], strides[1], strides[2], strides[3], h, w);
    else upsample_scale_bwd_kernel<32><<<nb, kNetThreads, 0, st>>>(gout, B, H, W, scale, gpred, strides[0], strides[1], strides[2], strides[3], h, w);
    return (int)cudaGetLastError();
}
extern "C" int tef_decoder_up(const float *x, const float *skip, const float *pred, const long *pred_strides, float *out, int B, int h, int w, int C,
                              int H, int W, void *stream) {
    if (!x || !out || B < 1 || h < 1 || w < 1 || H < 1 || W < 1 || C < 2 || (C & 1) || (pred && !pred_strides)) return TEF_EINVAL;
    const long z[4] = { 0, 0, 0, 0 };
    const long *ps = pred ? pred_strides : z;
    ProfScope ps_(K_NETWORK, (cudaStream_t)stream);
    if (H == 2 * h && W == 2 * w && B <= 65535 && h + 1 <= 65535) {
        const unsigned n_row = (unsigned)(w + 1) * (unsigned)((C >> 1) + (pred ? 1 : 0));
        decoder_up2_kernel<<<dim3((n_row + kNetThreads - 1) / kNetThreads, h + 1, B), kNetThreads, 0, (cudaStream_t)stream>>>(x, skip, pred, ps[0], ps[1], ps[2],
                                                                                                                        ps[3], out, B, h, w, C);
    } else {
        decoder_up_kernel<<<B * H, kNetThreads, 0, (cudaStream_t)stream>>>(x, skip, pred, ps[0], ps[1], ps[2], ps[3], out, B, h, w, C, H, W);
    }
    return (int)cudaGetLastError();
}
extern "C" int tef_decoder_up_bwd(const float *gout, float *gx, float *gpred, const long *pred_strides, int B, int h, int w, int C, int H, int W,
                                  void *stream) {
    if (!gout || !gx || B < 1 || h < 1 || w < 1 || H < 1 || W < 1 || C < 2 || (C & 1) || (gpred && !pred_strides)) return TEF_EINVAL;
    if ((float)H / (float)h * 2.0f + 3.0f > (float)kUpWin || (float)W / (float)w * 2.0f + 3.0f > (float)kUpWin) return TEF_ELIMIT;
    const long z[4] = { 0, 0, 0, 0 };
    const long *ps = gpred ? pred_strides : z;
    ProfScope ps_(K_NETWORK, (cudaStream_t)stream);
    decoder_up_bwd_kernel<<<grid_for((long)B * h * w * ((C >> 1) + (gpred ? 1 : 0))), kNetThreads, 0, (cudaStream_t)stream>>>(gout, gx, gpred, ps[0], ps[1], ps[2],
                                                                                                                          ps[3], B, h, w, C, H, W);
    return (int)cudaGetLastError();
}
