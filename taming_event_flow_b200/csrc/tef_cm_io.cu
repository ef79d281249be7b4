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

// both event sets of one pass in one launch: CTAs [0, nb0) stage set 0, the rest set 1
struct StageTwo {
    float4 *ev_io[2]; const float2 *mk_in[2]; float4 *ev_out[2]; float2 *mk_out[2];
    long rows[2]; float pass_index[2]; const float *ts_override[2]; int nb0;
};
__global__ void __launch_bounds__(kThreads) stage_two_kernel(const __grid_constant__ StageTwo s) {
    const int k = blockIdx.x >= s.nb0 ? 1 : 0;
    const long i = (long)(blockIdx.x - (k ? s.nb0 : 0)) * kThreads + threadIdx.x;
    if (i >= s.rows[k]) return;
    float4 e = s.ev_io[k][i];
    e.x = e.x + s.pass_index[k];
    s.ev_io[k][i] = e;
    if (s.ts_override[k]) e.x = __ldg(s.ts_override[k]);
    s.ev_out[k][i] = e;
    s.mk_out[k][i] = s.mk_in[k][i];
}

struct FlowPtrs { const float *p[TEF_MAX_FLOWS]; };

// [B][2][H][W] planar (ch0 = x, ch1 = y) -> dual-phase float2 [B][phase][H+1][Wp] with zero padding (tef_device.cuh)
__global__ void __launch_bounds__(kThreads) pack_flow_kernel(const __grid_constant__ FlowPtrs src, float2 *__restrict__ packed, int t, int P,
                                                             int B, Res r) {
    const int f = blockIdx.z, b = blockIdx.y;
    const long HW = (long)r.H * r.W;
    const float *sx = src.p[f] + (long)b * 2 * HW, *sy = sx + HW;
    float2 *dst = packed + (((long)f * P + t) * B + b) * 2 * r.fplane;
    const int n = 2 * r.fplane;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const int phase = i >= r.fplane;
        const int q = i - phase * r.fplane;
        const int y = q / r.Wp, x = q % r.Wp - phase;
        const bool in = (y < r.H) && (x >= 0) && (x < r.W);
        dst[i] = in ? make_float2(sx[y * r.W + x], sy[y * r.W + x]) : make_float2(0.f, 0.f);
    }
}

template <bool DET>
__global__ void __launch_bounds__(kThreads) unpack_grad_kernel(const float2 *__restrict__ packed, float *__restrict__ out, int F, int P, int B,
                                                               int W, long HW, ImgGeom g) {
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
            ox[i] = from_fix(a.x + c.x); oy[i] = from_fix(a.y + c.y);
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

extern "C" int tef_unpack_flow_grad(const void *packed, void *out, int F, int P, int B, int H, int W, int deterministic, void *stream) {
    if (!packed || !out || F < 1 || P < 1 || B < 1) return TEF_EINVAL;
    const long HW = (long)H * W;
    int bx = (int)((HW + kThreads - 1) / kThreads);
    if (bx > 148 * 4) bx = 148 * 4;
    ProfScope ps(K_UNPACK_GRAD, (cudaStream_t)stream);
    ImgGeom g; g.Wp = (W + 3) & ~1; g.plane = (long)H * g.Wp;
    if (deterministic) unpack_grad_kernel<true><<<dim3(bx, B, F * P), kThreads, 0, (cudaStream_t)stream>>>((const float2 *)packed, (float *)out, F, P, B, W, HW, g);
    else unpack_grad_kernel<false><<<dim3(bx, B, F * P), kThreads, 0, (cudaStream_t)stream>>>((const float2 *)packed, (float *)out, F, P, B, W, HW, g);
    return (int)cudaGetLastError();
}

extern "C" int tef_update_pass(const tef_update_desc *u, void *stream) {
    if (!u) return TEF_EINVAL;
    if (u->F > 0) {
        int rc = tef_pack_flow(u->flow_maps, u->F, u->t, u->P, u->B, u->H, u->W, u->packed, stream);
        if (rc) return rc;
    }
    StageTwo s;
    for (int k = 0; k < 2; ++k) {
        if (u->rows[k] < 0) return TEF_EINVAL;
        if (u->rows[k] > 0 && (!u->events[k] || !u->masks[k] || !u->ev_out[k] || !u->mk_out[k])) return TEF_EINVAL;
        s.ev_io[k] = (float4 *)u->events[k]; s.mk_in[k] = (const float2 *)u->masks[k];
        s.ev_out[k] = (float4 *)u->ev_out[k]; s.mk_out[k] = (float2 *)u->mk_out[k];
        s.rows[k] = u->rows[k]; s.pass_index[k] = u->pass_index[k]; s.ts_override[k] = u->ts_override[k];
    }
    s.nb0 = (int)((u->rows[0] + kThreads - 1) / kThreads);
    const int nb = s.nb0 + (int)((u->rows[1] + kThreads - 1) / kThreads);
    if (nb > 0) {
        ProfScope ps(K_STAGE_EVENTS, (cudaStream_t)stream);
        stage_two_kernel<<<nb, kThreads, 0, (cudaStream_t)stream>>>(s);
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
