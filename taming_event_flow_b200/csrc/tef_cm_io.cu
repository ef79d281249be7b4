// tef_cm_io.cu -- staging kernels of the loss modules' `update` / gradient hand-back:
// event staging (loss/flow.py:456-473), flow-map packing (update_base, :46-66) and
// unpacking of the packed flow gradient.  All three are pure HBM streaming kernels
// (vectorised 16 B accesses, grid sized in multiples of the SM count by the caller).
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

__global__ void __launch_bounds__(kThreads) stage_events_kernel(float4 *__restrict__ ev_inout, const float2 *__restrict__ mk_in,
                                                                float4 *__restrict__ ev_out, float2 *__restrict__ mk_out, long rows,
                                                                float pass_index, const float *__restrict__ ts_override) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= rows) return;
    float4 e = ev_inout[i];
    e.x = e.x + pass_index;                 // event_list[:, :, 0:1] += self._passes (in place, :457)
    ev_inout[i] = e;
    if (ts_override) e.x = __ldg(ts_override);   // round_ts (:461-463)
    ev_out[i] = e;
    mk_out[i] = mk_in[i];
}

// both event sets of one pass in one launch: CTAs [0, nb0) stage set 0, the rest set 1.  With `bins` the kernel also
// counts every non-padding row into the tile-sort histogram of its segment (what sort_hist_kernel would do in the
// forward call, there at the price of another pass over the events).
struct StageTwo {
    float4 *ev_io[2]; const float2 *mk_in[2]; float4 *ev_out[2]; float2 *mk_out[2];
    long rows[2]; float pass_index[2]; const float *ts_override[2]; int nb0;
    int *bins; int first_bin[2]; int n[2]; int tiles_x, tiles, H, W;
    // element strides (sample, row, column) of the caller's event / mask tensors; strided[k] = 0: contiguous [B][n][4] / [B][n][2]
    // rows (one 16-byte / 8-byte access).  The reference's custom_collate hands out transposed views of [B][C][n] storage
    // (dataloader/base.py:414-431), which stay strided after .to(device).
    int strided[2]; long es[2][3], ms[2][3];
};
__device__ __forceinline__ void stage_two_body(const StageTwo &s, int blk) {
    const int k = blk >= s.nb0 ? 1 : 0;
    const long i = (long)(blk - (k ? s.nb0 : 0)) * kThreads + threadIdx.x;
    if (i >= s.rows[k]) return;
    float4 e; float2 m;
    if (!s.strided[k]) {
        e = s.ev_io[k][i];
        e.x = e.x + s.pass_index[k];
        s.ev_io[k][i] = e;
        m = s.mk_in[k][i];
    } else {
        const long b = i / s.n[k], r = i - b * s.n[k];
        float *pe = reinterpret_cast<float *>(s.ev_io[k]) + b * s.es[k][0] + r * s.es[k][1];
        const float *pm = reinterpret_cast<const float *>(s.mk_in[k]) + b * s.ms[k][0] + r * s.ms[k][1];
        e = make_float4(pe[0] + s.pass_index[k], pe[s.es[k][2]], pe[2 * s.es[k][2]], pe[3 * s.es[k][2]]);
        pe[0] = e.x;                        // only the timestamp column changes (:457)
        m = make_float2(pm[0], pm[s.ms[k][2]]);
    }
    if (s.ts_override[k]) e.x = __ldg(s.ts_override[k]);
    s.ev_out[k][i] = e;
    s.mk_out[k][i] = m;
    if (s.bins && !(m.x == 0.0f && m.y == 0.0f))
        atomicAdd(s.bins + sort_bin(s.first_bin[k], s.tiles_x, s.tiles, s.H, s.W, (int)(i / s.n[k]), e.y, e.z, m.x == 0.0f), 1);
}
struct FlowPtrs { const float *p[TEF_MAX_FLOWS]; };

// [B][2][H][W] planar (ch0 = x, ch1 = y) -> dual-phase float2 [B][phase][H+1][Wp] with zero padding (tef_device.cuh);
// bx CTAs walk one (flow scale f, sample b) map
__device__ __forceinline__ void pack_flow_body(const FlowPtrs &src, float2 *__restrict__ packed, float4 *__restrict__ packedq, int t, int P, int B,
                                               const Res &r, int f, int b, int xblk, int bx) {
    const long HW = (long)r.H * r.W;
    const float *sx = src.p[f] + (long)b * 2 * HW, *sy = sx + HW;
    float2 *dst = packed + (((long)f * P + t) * B + b) * 2 * r.fplane;
    const int n = 2 * r.fplane;
    for (int i = xblk * kThreads + threadIdx.x; i < n; i += bx * kThreads) {
        const int phase = i >= r.fplane;
        const int q = i - phase * r.fplane;
        const int y = q / r.Wp, x = q % r.Wp - phase;
        const bool in = (y < r.H) && (x >= 0) && (x < r.W);
        dst[i] = in ? make_float2(sx[y * r.W + x], sy[y * r.W + x]) : make_float2(0.f, 0.f);
    }
    if (!packedq) return;
    // quad-cell copy (tef_device.cuh, quad_cell): cell (Y, X) of parity (px, py) holds rows 2Y-py, 2Y-py+1 x columns 2X-px, 2X-px+1
    float4 *dq = packedq + (((long)f * P + t) * B + b) * 8 * r.cplane;
    const int ncell = 4 * r.cplane;
    for (int i = xblk * kThreads + threadIdx.x; i < ncell; i += bx * kThreads) {
        const int ph = i / r.cplane, c = i - ph * r.cplane;
        const int Y = c / r.CX, X = c - Y * r.CX;
        const int y = 2 * Y - (ph >> 1), x = 2 * X - (ph & 1);
        float2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int yy = y + (k >> 1), xx = x + (k & 1);
            const bool in = (yy >= 0) && (yy < r.H) && (xx >= 0) && (xx < r.W);
            v[k] = in ? make_float2(sx[yy * r.W + xx], sy[yy * r.W + xx]) : make_float2(0.f, 0.f);
        }
        dq[2 * (long)i] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        dq[2 * (long)i + 1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
    }
}
__global__ void __launch_bounds__(kThreads) pack_flow_kernel(const __grid_constant__ FlowPtrs src, float2 *__restrict__ packed, int t, int P,
                                                             int B, Res r) {
    pack_flow_body(src, packed, nullptr, t, P, B, r, blockIdx.z, blockIdx.y, blockIdx.x, gridDim.x);
}

// Iterative.update / Linear.update in ONE launch: CTAs [0, nb_stage) stage (and count) the events, the rest pack the flow maps
struct PackArgs { FlowPtrs src; float2 *packed; float4 *packedq; int t, P, B, F, bx; Res r; };
__global__ void __launch_bounds__(kThreads) update_pass_kernel(const __grid_constant__ StageTwo s, const __grid_constant__ PackArgs k, int nb_stage) {
    if ((int)blockIdx.x < nb_stage) { stage_two_body(s, blockIdx.x); return; }
    const int pb = blockIdx.x - nb_stage;
    pack_flow_body(k.src, k.packed, k.packedq, k.t, k.P, k.B, k.r, pb / (k.bx * k.B), (pb / k.bx) % k.B, pb % k.bx, k.bx);
}

template <bool DET>
__global__ void __launch_bounds__(kThreads) unpack_grad_kernel(const float2 *__restrict__ packed, float *__restrict__ out, int F, int P, int B,
                                                               int W, long HW, ImgGeom g, const float *__restrict__ det_scale) {
    const int fp = blockIdx.z, b = blockIdx.y;
    const int f = fp / P, t = fp % P;
    const float2 *src = packed + (((long)f * P + t) * B + b) * (DET ? 4 : 2) * g.plane;
    float *ox = out + ((((long)t * F + f) * B + b) * 2) * HW, *oy = ox + HW;
    for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < HW; i += (long)gridDim.x * kThreads) {
        const long o = (i / W) * g.Wp + (i % W);
        if (!DET) {
            const float2 a = src[o], c = src[g.plane + o + 1];
            ox[i] = a.x + c.x; oy[i] = a.y + c.y;
        } else {
            const longlong2 *q = reinterpret_cast<const longlong2 *>(src);
            const longlong2 a = q[o], c = q[g.plane + o + 1];
            const double sc = (det_scale ? (double)__ldg(det_scale) : 1.0) * (1.0 / kFixScale);      // power of two: exact
            ox[i] = (float)((double)(a.x + c.x) * sc); oy[i] = (float)((double)(a.y + c.y) * sc);
        }
    }
}

}  // namespace tef

using namespace tef;

extern "C" int tef_stage_events(void *events_inout, const void *pol_mask, void *ev_out, void *mk_out, long rows, float pass_index,
                                const float *ts_override, void *stream) {
    if (rows < 0) return TEF_EINVAL;
    if (rows == 0) return 0;
    if (!events_inout || !pol_mask || !ev_out || !mk_out) return TEF_EINVAL;
    ProfScope ps(K_STAGE_EVENTS, (cudaStream_t)stream);
    stage_events_kernel<<<(unsigned)((rows + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        (float4 *)events_inout, (const float2 *)pol_mask, (float4 *)ev_out, (float2 *)mk_out, rows, pass_index, ts_override);
    return (int)cudaGetLastError();
}

extern "C" int tef_pack_flow(const void *const *flow_maps_host, int F, int t, int P, int B, int H, int W, void *packed, void *stream) {
    if (!flow_maps_host || !packed || F < 1 || F > TEF_MAX_FLOWS || t < 0 || t >= P || B < 1) return TEF_EINVAL;
    FlowPtrs src;
    for (int f = 0; f < F; ++f) { if (!flow_maps_host[f]) return TEF_EINVAL; src.p[f] = (const float *)flow_maps_host[f]; }
    const Res r = Res::make(H, W);
    int bx = (2 * r.fplane + kThreads - 1) / kThreads;
    if (bx > 148 * 4) bx = 148 * 4;
    ProfScope ps(K_PACK_FLOW, (cudaStream_t)stream);
    pack_flow_kernel<<<dim3(bx, B, F), kThreads, 0, (cudaStream_t)stream>>>(src, (float2 *)packed, t, P, B, r);
    return (int)cudaGetLastError();
}

extern "C" int tef_unpack_flow_grad(const void *packed, void *out, int F, int P, int B, int H, int W, int deterministic, const float *det_scale,
                                    void *stream) {
    if (!packed || !out || F < 1 || P < 1 || B < 1) return TEF_EINVAL;
    const long HW = (long)H * W;
    int bx = (int)((HW + kThreads - 1) / kThreads);
    if (bx > 148 * 4) bx = 148 * 4;
    ProfScope ps(K_UNPACK_GRAD, (cudaStream_t)stream);
    ImgGeom g; g.Wp = (W + 3) & ~1; g.plane = (long)H * g.Wp; g.lo_off = 0;
    if (deterministic) unpack_grad_kernel<true><<<dim3(bx, B, F * P), kThreads, 0, (cudaStream_t)stream>>>((const float2 *)packed, (float *)out, F, P, B, W, HW, g, det_scale);
    else unpack_grad_kernel<false><<<dim3(bx, B, F * P), kThreads, 0, (cudaStream_t)stream>>>((const float2 *)packed, (float *)out, F, P, B, W, HW, g, nullptr);
    return (int)cudaGetLastError();
}

extern "C" int tef_update_pass(const tef_update_desc *u, void *stream) {
    if (!u) return TEF_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    PackArgs k;
    k.F = 0; k.bx = 0; k.B = u->B; k.packedq = nullptr;
    if (u->F > 0) {
        if (!u->packed || u->F > TEF_MAX_FLOWS || u->t < 0 || u->t >= u->P || u->B < 1) return TEF_EINVAL;
        for (int f = 0; f < u->F; ++f) { if (!u->flow_maps[f]) return TEF_EINVAL; k.src.p[f] = (const float *)u->flow_maps[f]; }
        k.packed = (float2 *)u->packed; k.packedq = (float4 *)u->packedq; k.t = u->t; k.P = u->P; k.F = u->F; k.r = Res::make(u->H, u->W);
        k.bx = (2 * k.r.fplane + kThreads - 1) / kThreads;
        if (k.bx > 148 * 4) k.bx = 148 * 4;
    }
    StageTwo s;
    for (int i = 0; i < 2; ++i) {
        if (u->rows[i] < 0) return TEF_EINVAL;
        if (u->rows[i] > 0 && (!u->events[i] || !u->masks[i] || !u->ev_out[i] || !u->mk_out[i])) return TEF_EINVAL;
        s.ev_io[i] = (float4 *)u->events[i]; s.mk_in[i] = (const float2 *)u->masks[i];
        s.ev_out[i] = (float4 *)u->ev_out[i]; s.mk_out[i] = (float2 *)u->mk_out[i];
        s.rows[i] = u->rows[i]; s.pass_index[i] = u->pass_index[i]; s.ts_override[i] = u->ts_override[i];
        s.strided[i] = u->strided[i] ? 1 : 0;
        s.n[i] = 1;
        if (s.strided[i]) {
            if (u->B < 1 || u->rows[i] % u->B) return TEF_EINVAL;
            s.n[i] = (int)(u->rows[i] / u->B) > 0 ? (int)(u->rows[i] / u->B) : 1;
            for (int j = 0; j < 3; ++j) { s.es[i][j] = u->ev_strides[i][j]; s.ms[i][j] = u->mk_strides[i][j]; }
        }
    }
    s.bins = nullptr;
    if (u->hist) {
        if (!u->sort_bins || u->t < 0 || u->t >= u->P || u->B < 1) return TEF_EINVAL;
        s.tiles_x = (u->W + 15) / 16; s.tiles = s.tiles_x * ((u->H + 7) / 8); s.H = u->H; s.W = u->W;
        const long per_seg = (long)u->B * s.tiles * kBinsPerTile;
        if (2 * u->P * per_seg > 0x7fffffffl) return TEF_ELIMIT;
        if (u->zero_bins) cudaMemsetAsync(u->sort_bins, 0, sizeof(int) * (2 * u->P * per_seg + 1), st);
        s.bins = (int *)u->sort_bins;
        for (int i = 0; i < 2; ++i) {
            if (u->rows[i] % u->B) return TEF_EINVAL;
            s.first_bin[i] = (int)((i * u->P + u->t) * per_seg);
            s.n[i] = (int)(u->rows[i] / u->B) > 0 ? (int)(u->rows[i] / u->B) : 1;
        }
    }
    s.nb0 = (int)((u->rows[0] + kThreads - 1) / kThreads);
    const int nb = s.nb0 + (int)((u->rows[1] + kThreads - 1) / kThreads);
    const int npk = k.bx * k.B * k.F;
    if (nb + npk > 0) {
        ProfScope ps(K_STAGE_EVENTS, st);                  // one launch: staging (+ histogram) and flow packing
        update_pass_kernel<<<nb + npk, kThreads, 0, st>>>(s, k, nb);
    }
    return (int)cudaGetLastError();
}

extern "C" int tef_version(void) { return 100; }

extern "C" const char *tef_strerror(int code) {
    switch (code) {
        case 0: return "success";
        case TEF_EINVAL: return "invalid argument (size or null pointer)";
        case TEF_ELIMIT: return "configuration beyond the library's static limits (TEF_MAX_*)";
        case TEF_EEMPTY: return "empty window list (the reference raises RuntimeError: torch.cat of an empty list)";
        case TEF_EMODE4: return "iterative_mode 'four' with border compensation (the reference raises TypeError)";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}
