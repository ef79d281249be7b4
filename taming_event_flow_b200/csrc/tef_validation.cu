// tef_validation.cu -- fused steps of the reference's validation criteria (loss/flow_val.py).
//
// `Iterative.update` of the reference re-warps every accumulated event, pulls the new window back through every flow
// map, and carries every older flow map one window forward by splatting it along itself: ~150 small eager ops per
// window at ten windows, all launch-bound.  The four kernels below do one whole stage each, with the per-element
// arithmetic of the stand-alone operators (tef_primitives.cu) in the same order, so results are bit-identical to the
// operator-by-operator route except for the order of the image sums of the flow splat.  Validation runs with batch
// size 1 like upstream (loss/flow_val.py:30-38 builds batch-1 index grids).
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

// event_propagation + purge_unfeasible of one event (utils/iwe.py:5-14, :43-60)
__device__ __forceinline__ float step_event(float2 &loc, float ts, float tref, float2 flow, const Res &r) {
    const float dt = tref - ts;
    const float y = loc.x + dt * flow.x, x = loc.y + dt * flow.y;
    const float in = inside(y, x, r) ? 1.0f : 0.0f;
    loc = make_float2(y * in, x * in);
    return in;
}

// loss/flow_val.py:483-517: every event seen so far, one window forward with the newest flow map, in place
__global__ void __launch_bounds__(kThreads) val_fw_step_kernel(const float *__restrict__ mapx, const float *__restrict__ mapy, float2 *__restrict__ loc,
                                                               float *__restrict__ ts, float2 *__restrict__ mask, float tref, long n, Res r) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float2 l = loc[i];
    const float2 m = mask[i];
    const float in = step_event(l, ts[i], tref, event_flow_planar(mapx, mapy, l.x, l.y, r), r);
    loc[i] = l;
    mask[i] = make_float2(m.x * in, m.y * in);
    ts[i] = tref;
}

// loss/flow_val.py:519-556: the new window back to time 0 through maps n_maps-1 ... 0 (maps [n_maps][H][W]), in place
__global__ void __launch_bounds__(kThreads) val_bw_chain_kernel(const float *__restrict__ mapsx, const float *__restrict__ mapsy, int n_maps,
                                                                float2 *__restrict__ loc, const float *__restrict__ ts0, float2 *__restrict__ mask,
                                                                long n, Res r) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const long HW = (long)r.H * r.W;
    float2 l = loc[i], m = mask[i];
    float ts = ts0[i];
    for (int k = n_maps - 1; k >= 0; --k) {
        const float in = step_event(l, ts, (float)k, event_flow_planar(mapsx + k * HW, mapsy + k * HW, l.x, l.y, r), r);
        m = make_float2(m.x * in, m.y * in);
        ts = (float)k;
    }
    loc[i] = l;
    mask[i] = m;
}

// forward_prop_flow (loss/flow_val.py:43-74), splat half: pixel p of map i moves by dt_i * flow and deposits
// (weight, weight*flow_y, weight*flow_x) on its four corners; acc [n_maps][3][H][W]
__global__ void __launch_bounds__(kThreads) val_prop_splat_kernel(const float *__restrict__ mapsx, const float *__restrict__ mapsy, int first, int n_maps,
                                                                  int tref_is_next, float tref, float *__restrict__ acc, Res r) {
    const long HW = (long)r.H * r.W;
    const long j = (long)blockIdx.x * kThreads + threadIdx.x;
    if (j >= HW * n_maps) return;
    const int i = (int)(j / HW);
    const long p = j - (long)i * HW;
    const float py = (float)(p / r.W), px = (float)(p % r.W);
    const float2 f = event_flow_planar(mapsx + i * HW, mapsy + i * HW, py, px, r);          // (y, x)
    float2 l = make_float2(py, px);
    const float in = step_event(l, (float)(first + i), tref_is_next ? (float)(first + i + 1) : tref, f, r);
    Corners c;
    corners(l.x, l.y, r, c);
    float *a = acc + (long)i * 3 * HW;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ky = k >> 1, kx = k & 1;
        const float ok = (c.oky[ky] && c.okx[kx]) ? 1.0f : 0.0f;
        const long id = (long)((c.cy[ky] * ok) * (float)r.W + c.cx[kx] * ok);               // utils/iwe.py:104,:110-111
        const float w = (c.wy[ky] * c.wx[kx]) * ok;                                         // :107
        const float vn = w * in, vy = (w * f.x) * in, vx = (w * f.y) * in;                  // interpolate(): weights * mask
        if (vn != 0.0f) red_add_f32(a + id, vn);
        if (vy != 0.0f) red_add_f32(a + HW + id, vy);
        if (vx != 0.0f) red_add_f32(a + 2 * HW + id, vx);
    }
}
// ... and the normalisation half: map = splat / (norm + 1e-9)
__global__ void __launch_bounds__(kThreads) val_prop_norm_kernel(const float *__restrict__ acc, float *__restrict__ outx, float *__restrict__ outy,
                                                                 int n_maps, long HW) {
    const long j = (long)blockIdx.x * kThreads + threadIdx.x;
    if (j >= HW * n_maps) return;
    const long i = j / HW, p = j - i * HW;
    const float *a = acc + i * 3 * HW;
    const float d = a[p] + 1e-9f;
    outy[j] = a[HW + p] / d;
    outx[j] = a[2 * HW + p] / d;
}

// loss/flow_val.py:579-605: pixel trajectories through the newest map; idx [2][H][W] = (y, x) planes, in place
__global__ void __launch_bounds__(kThreads) val_trajectory_kernel(const float *__restrict__ mapx, const float *__restrict__ mapy, float *__restrict__ idx,
                                                                  float *__restrict__ out_mask, float *__restrict__ accx, float *__restrict__ accy, Res r) {
    const long HW = (long)r.H * r.W;
    const long p = (long)blockIdx.x * kThreads + threadIdx.x;
    if (p >= HW) return;
    const float y = idx[p], x = idx[HW + p];
    const float valid = (y >= 0.0f && y <= r.hm1 && x >= 0.0f && x <= r.wm1) ? 1.0f : 0.0f;
    out_mask[p] = out_mask[p] + valid;
    const float2 f = event_flow_planar(mapx, mapy, y, x, r);
    const float ny = y + f.x * valid, nx = x + f.y * valid;
    idx[p] = ny; idx[HW + p] = nx;
    accx[p] = nx - (float)(p % r.W);
    accy[p] = ny - (float)(p / r.W);
}

// _pol_images of the validation criteria (loss/flow_val.py:116-131, :189-274): get_interpolation, the optional per-event
// weight (`weights * extra`), and one interpolate per polarity -- the operator route's arithmetic, event by event:
// value = ((corner weight) * extra) * mask_c reduced into out[c][pixel]; out [2][H][W] is zeroed by the caller.
__global__ void __launch_bounds__(kThreads) val_pol_images_kernel(const float2 *__restrict__ loc, const float2 *__restrict__ mask,
                                                                  const float *__restrict__ extra, float *__restrict__ out, long n, Res r, int round_idx) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const long HW = (long)r.H * r.W;
    const float2 l = loc[i], m = mask[i];
    const bool has_ex = extra != nullptr;
    const float ex = has_ex ? extra[i] : 1.0f;
    if (round_idx) {
        const float ry = rintf(l.x), rx = rintf(l.y);              // torch.round: half to even (utils/iwe.py:79)
        const float ok = (ry >= 0.f && ry < (float)r.H && rx >= 0.f && rx < (float)r.W) ? 1.0f : 0.0f;
        const long px = (long)((ry * ok) * (float)r.W + rx * ok);
        float w = (1.0f * 1.0f) * ok;
        if (has_ex) w = w * ex;
        const float vp = w * m.x, vn = w * m.y;
        if (vp != 0.0f) red_add_f32(out + px, vp);
        if (vn != 0.0f) red_add_f32(out + HW + px, vn);
        return;
    }
    Corners c;
    corners(l.x, l.y, r, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ky = k >> 1, kx = k & 1;
        const float ok = (c.oky[ky] && c.okx[kx]) ? 1.0f : 0.0f;
        const long px = (long)((c.cy[ky] * ok) * (float)r.W + c.cx[kx] * ok);               // utils/iwe.py:104,:110-111
        float w = (c.wy[ky] * c.wx[kx]) * ok;                                               // :107
        if (has_ex) w = w * ex;
        const float vp = w * m.x, vn = w * m.y;
        if (vp != 0.0f) red_add_f32(out + px, vp);
        if (vn != 0.0f) red_add_f32(out + HW + px, vn);
    }
}

// One launch for everything `update` appends (loss/flow_val.py:75-114, :483-487, :519-528, :558-562): CTAs [0, nb_ev) copy the
// window's events into the row arrays (raw list, forward list, backward list), adding the pass index to the caller's
// timestamps in place; the other CTAs copy the newest flow map and event mask into the per-window map stacks.
struct ValAppend {
    float4 *events; const float2 *mask; long n; float pass_index; const float *ts_override;
    float *ev_ts; float2 *ev_loc, *ev_mask;                  // raw lists (BaseValidation)
    float *fw_ts; float2 *fw_loc, *fw_mask;                  // Iterative: forward-warped lists (nullptr for Linear)
    float2 *bw_loc, *bw_mask;                                // Iterative: this window's rows of the backward-warped lists
    const float *flow, *emask; long HW;                      // newest flow [2][H][W] and event mask [H][W]
    float *map_x, *map_y, *map_e, *prop_x, *prop_y;          // slot `now` of the stacks (prop_*: Iterative, else nullptr)
    int nb_ev;
};
__global__ void __launch_bounds__(kThreads) val_append_kernel(const __grid_constant__ ValAppend a) {
    if ((int)blockIdx.x < a.nb_ev) {
        const long i = (long)blockIdx.x * kThreads + threadIdx.x;
        if (i >= a.n) return;
        float4 e = a.events[i];
        e.x = e.x + a.pass_index;                            // event_list[:, :, 0:1] += self._passes (in place, :86)
        a.events[i] = e;
        const float ts = a.ts_override ? __ldg(a.ts_override) : e.x;      // round_ts (:87-88)
        const float2 loc = make_float2(e.y, e.z), m = a.mask[i];
        a.ev_ts[i] = ts; a.ev_loc[i] = loc; a.ev_mask[i] = m;
        if (a.fw_ts) { a.fw_ts[i] = ts; a.fw_loc[i] = loc; a.fw_mask[i] = m; a.bw_loc[i] = loc; a.bw_mask[i] = m; }
        return;
    }
    const long p = (long)(blockIdx.x - a.nb_ev) * kThreads + threadIdx.x;
    if (p >= a.HW) return;
    const float fx = a.flow[p], fy = a.flow[a.HW + p];
    a.map_x[p] = fx; a.map_y[p] = fy; a.map_e[p] = a.emask[p];
    if (a.prop_x) { a.prop_x[p] = fx; a.prop_y[p] = fy; }
}

}  // namespace tef

using namespace tef;
#define ST ((cudaStream_t)stream)
#define TEF_GRID(n) (unsigned)(((n) + kThreads - 1) / kThreads)

extern "C" int tef_val_forward_step(const float *mapx, const float *mapy, float *loc, float *ts, float *mask, float tref, long n, int H, int W,
                                    void *stream) {
    if (n < 0 || H < 2 || W < 2) return TEF_EINVAL;
    if (n == 0) return 0;
    if (!mapx || !mapy || !loc || !ts || !mask) return TEF_EINVAL;
    ProfScope pr(K_VALIDATION, ST);
    val_fw_step_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>(mapx, mapy, (float2 *)loc, ts, (float2 *)mask, tref, n, Res::make(H, W));
    return (int)cudaGetLastError();
}

extern "C" int tef_val_backward_chain(const float *mapsx, const float *mapsy, int n_maps, float *loc, const float *ts, float *mask, long n,
                                      int H, int W, void *stream) {
    if (n < 0 || n_maps < 1 || H < 2 || W < 2) return TEF_EINVAL;
    if (n == 0) return 0;
    if (!mapsx || !mapsy || !loc || !ts || !mask) return TEF_EINVAL;
    ProfScope pr(K_VALIDATION, ST);
    val_bw_chain_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>(mapsx, mapsy, n_maps, (float2 *)loc, ts, (float2 *)mask, n, Res::make(H, W));
    return (int)cudaGetLastError();
}

extern "C" int tef_val_forward_prop_flow(const float *mapsx, const float *mapsy, int first, int n_maps, int tref_is_next, float tref, float *acc,
                                         float *outx, float *outy, int H, int W, void *stream) {
    if (first < 0 || n_maps < 0 || H < 2 || W < 2) return TEF_EINVAL;
    if (n_maps == 0) return 0;
    if (!mapsx || !mapsy || !acc || !outx || !outy) return TEF_EINVAL;
    const long HW = (long)H * W;
    cudaMemsetAsync(acc, 0, sizeof(float) * 3 * HW * n_maps, ST);
    ProfScope pr(K_VALIDATION, ST);
    val_prop_splat_kernel<<<TEF_GRID(HW * n_maps), kThreads, 0, ST>>>(mapsx, mapsy, first, n_maps, tref_is_next, tref, acc, Res::make(H, W));
    val_prop_norm_kernel<<<TEF_GRID(HW * n_maps), kThreads, 0, ST>>>(acc, outx, outy, n_maps, HW);
    return (int)cudaGetLastError();
}

extern "C" int tef_val_trajectory_step(const float *mapx, const float *mapy, float *idx, float *out_mask, float *accx, float *accy, int H, int W,
                                       void *stream) {
    if (H < 2 || W < 2 || !mapx || !mapy || !idx || !out_mask || !accx || !accy) return TEF_EINVAL;
    ProfScope pr(K_VALIDATION, ST);
    val_trajectory_kernel<<<TEF_GRID((long)H * W), kThreads, 0, ST>>>(mapx, mapy, idx, out_mask, accx, accy, Res::make(H, W));
    return (int)cudaGetLastError();
}

extern "C" int tef_val_append_window(const tef_val_append *d, void *stream) {
    if (!d || d->n < 0 || d->H < 2 || d->W < 2) return TEF_EINVAL;
    if (!d->flow || !d->event_mask || !d->map_x || !d->map_y || !d->map_e) return TEF_EINVAL;
    if (d->n > 0 && (!d->events || !d->pol_mask || !d->ev_ts || !d->ev_loc || !d->ev_mask)) return TEF_EINVAL;
    if (d->fw_ts && (!d->fw_loc || !d->fw_mask || !d->bw_loc || !d->bw_mask || !d->prop_x || !d->prop_y)) return TEF_EINVAL;
    ValAppend a;
    a.events = (float4 *)d->events; a.mask = (const float2 *)d->pol_mask; a.n = d->n; a.pass_index = d->pass_index; a.ts_override = d->ts_override;
    a.ev_ts = d->ev_ts; a.ev_loc = (float2 *)d->ev_loc; a.ev_mask = (float2 *)d->ev_mask;
    a.fw_ts = d->fw_ts; a.fw_loc = (float2 *)d->fw_loc; a.fw_mask = (float2 *)d->fw_mask;
    a.bw_loc = (float2 *)d->bw_loc; a.bw_mask = (float2 *)d->bw_mask;
    a.flow = d->flow; a.emask = d->event_mask; a.HW = (long)d->H * d->W;
    a.map_x = d->map_x; a.map_y = d->map_y; a.map_e = d->map_e; a.prop_x = d->prop_x; a.prop_y = d->prop_y;
    a.nb_ev = (int)TEF_GRID(d->n);
    ProfScope pr(K_VALIDATION, ST);
    val_append_kernel<<<a.nb_ev + TEF_GRID(a.HW), kThreads, 0, ST>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int tef_val_pol_images(const float *loc, const float *mask, const float *extra, float *out, long n, int H, int W, int round_idx, void *stream) {
    if (n < 0 || H < 2 || W < 2 || !out) return TEF_EINVAL;
    cudaMemsetAsync(out, 0, sizeof(float) * 2 * (long)H * W, ST);
    if (n == 0) return 0;
    if (!loc || !mask) return TEF_EINVAL;
    ProfScope pr(K_VALIDATION, ST);
    val_pol_images_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>((const float2 *)loc, (const float2 *)mask, extra, out, n, Res::make(H, W), round_idx);
    return (int)cudaGetLastError();
}
