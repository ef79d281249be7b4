// tef_encodings.cu -- event encodings of the reference's dataloader/encodings.py as sm_100a kernels.
// HBM-streaming kernels: 12-16 B read per event (coalesced), one or two native fp32 reductions into an
// image that stays in L2.  Counting encodings are exact in fp32 for any order (integers < 2^24);
// the voxel grid sums real weights, so only its summation order differs from the reference.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

// xs.long()/ys.long() truncate toward zero; negative indices wrap like Python indexing (index_put_).  An index that is
// still outside the sensor makes the reference raise IndexError (dataloader/encodings.py:23-27): the event is skipped
// and *oob is set, which the host mirror turns into that IndexError.
__device__ __forceinline__ bool pixel_of(float xf, float yf, int H, int W, long &px, int *oob) {
    long x = (long)xf, y = (long)yf;
    if (x < 0) x += W;
    if (y < 0) y += H;
    if (x < 0 || x >= W || y < 0 || y >= H) {
        if (oob) *oob = 1;                                           // benign race: every writer stores the same value
        return false;
    }
    px = y * W + x;
    return true;
}

// Four consecutive events per thread, read with one 16-byte load per input array when the arrays are 16-byte aligned
// (`vec`, decided on the host) and the thread's four events exist; otherwise element by element.
constexpr int kEvPerThread = 4;
struct Quad { float v[kEvPerThread]; };
__device__ __forceinline__ Quad load_quad(const float *__restrict__ a, long i0, long n, bool vec) {
    Quad q;
    if (vec && i0 + kEvPerThread <= n) {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(a + i0));     // streamed once: evict-first
        q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
    } else {
#pragma unroll
        for (int k = 0; k < kEvPerThread; ++k) q.v[k] = (i0 + k < n) ? a[i0 + k] : 0.0f;
    }
    return q;
}

__global__ void __launch_bounds__(kThreads) to_image_kernel(const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ ps,
                                                            float *__restrict__ img, long n, int H, int W, int vec, int *oob) {
    const long i0 = ((long)blockIdx.x * kThreads + threadIdx.x) * kEvPerThread;
    if (i0 >= n) return;
    const Quad x = load_quad(xs, i0, n, vec), y = load_quad(ys, i0, n, vec), p = load_quad(ps, i0, n, vec);
#pragma unroll
    for (int k = 0; k < kEvPerThread; ++k) {
        long px;
        if (i0 + k < n && pixel_of(x.v[k], y.v[k], H, W, px, oob)) red_add_f32(img + px, p.v[k]);     // dataloader/encodings.py:27
    }
}

// accumulate=False (index_put_ without accumulation): the reference's CPU kernel walks the events in order, so the LAST
// event of a pixel wins.  Two passes make that deterministic on the GPU: the image first collects, as integers, the
// highest event index per pixel (atomicMax; -1 = no event), then every pixel replaces its index by that event's value.
__global__ void __launch_bounds__(kThreads) last_writer_kernel(const float *__restrict__ xs, const float *__restrict__ ys, int *__restrict__ winner,
                                                               long n, int H, int W, int *oob) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    long px;
    if (pixel_of(xs[i], ys[i], H, W, px, oob)) atomicMax(winner + px, (int)i);
}
__global__ void __launch_bounds__(kThreads) put_winner_kernel(const float *__restrict__ ps, float *__restrict__ img, long npx) {
    const long px = (long)blockIdx.x * kThreads + threadIdx.x;
    if (px >= npx) return;
    const int i = reinterpret_cast<const int *>(img)[px];
    img[px] = i < 0 ? 0.0f : ps[i];
}

__global__ void __launch_bounds__(kThreads) to_channels_kernel(const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ ps,
                                                               float *__restrict__ out, long n, int H, int W, int vec, int *oob) {
    const long i0 = ((long)blockIdx.x * kThreads + threadIdx.x) * kEvPerThread;
    if (i0 >= n) return;
    const Quad x = load_quad(xs, i0, n, vec), y = load_quad(ys, i0, n, vec), pq = load_quad(ps, i0, n, vec);
#pragma unroll
    for (int k = 0; k < kEvPerThread; ++k) {
        long px;
        if (i0 + k >= n || !pixel_of(x.v[k], y.v[k], H, W, px, oob)) continue;
        const float p = pq.v[k];
        const float mpos = p < 0.f ? 0.f : (p > 0.f ? 1.f : p);        // :71-76
        const float mneg = p > 0.f ? 0.f : (p < 0.f ? -1.f : p);
        const float vp = p * mpos, vn = p * mneg;                       // :78-79
        if (vp != 0.f) red_add_f32(out + px, vp);
        if (vn != 0.f) red_add_f32(out + (long)H * W + px, vn);
    }
}

__global__ void __launch_bounds__(kThreads) to_voxel_kernel(const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ ts,
                                                            const float *__restrict__ ps, float *__restrict__ out, long n, int bins, int H, int W,
                                                            int vec, int *oob) {
    const long i0 = ((long)blockIdx.x * kThreads + threadIdx.x) * kEvPerThread;
    if (i0 >= n) return;
    const Quad x = load_quad(xs, i0, n, vec), y = load_quad(ys, i0, n, vec), tq = load_quad(ts, i0, n, vec), pq = load_quad(ps, i0, n, vec);
    const float scale = (float)(bins - 1);
    const long plane = (long)H * W;
#pragma unroll
    for (int k = 0; k < kEvPerThread; ++k) {
        long px;
        if (i0 + k >= n || !pixel_of(x.v[k], y.v[k], H, W, px, oob)) continue;
        const float t = tq.v[k] * scale;                                // :47
        const float p = pq.v[k];
        // only the bins within distance 1 of t have a non-zero weight (:52); the others add +0
        const int b0 = (int)floorf(t) - 1;
        for (int b = max(b0, 0); b <= min(b0 + 3, bins - 1); ++b) {
            const float wgt = fmaxf(0.0f, 1.0f - fabsf(t - (float)b));
            const float v = p * wgt;                                    // :53
            if (v != 0.f) red_add_f32(out + (long)b * plane + px, v);
        }
    }
}

// Batched variant of events_to_channels for the loader->network hand-off: rows are the
// zero-padded [B][N][4] event tensor (ts, y, x, p) that custom_collate produces; padding rows have p = 0 and add nothing.
__global__ void __launch_bounds__(kThreads) to_channels_batched_kernel(const float4 *__restrict__ ev, float *__restrict__ out, int B, int N, int H, int W) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * N) return;
    const float4 e = ev[i];
    const float p = e.w;
    if (p == 0.f) return;
    long px;
    if (!pixel_of(e.z, e.y, H, W, px, nullptr)) return;
    const float mpos = p < 0.f ? 0.f : 1.f, mneg = p > 0.f ? 0.f : -1.f;
    float *o = out + (long)(i / N) * 2 * H * W;
    if (mpos != 0.f) red_add_f32(o + px, p * mpos);
    if (mneg != 0.f) red_add_f32(o + (long)H * W + px, p * mneg);
}

// Hot-pixel mask.  NOT in the reference ("parity unpinned"); the specification implemented here is the
// published routine of the sibling code base (tudelft/event_flow, dataloader/encodings.py: get_hot_event_mask):
//   mask = 1 everywhere; if idx > min_obvs: up to max_px times { take the arg-max of event_rate (lowest flat index on
//   ties); if its rate > max_rate: zero the rate there and clear the mask bit, else stop }.
// One CTA: every round is a block-wide arg-max over the (small) rate image.
__global__ void __launch_bounds__(1024) hot_mask_kernel(float *__restrict__ rate, float *__restrict__ mask, int n, int max_px, float max_rate, int active) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    __shared__ int s_stop;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mask[i] = 1.0f;
    if (!active) return;
    __syncthreads();
    for (int round = 0; round < max_px; ++round) {
        float bv = -INFINITY; int bi = n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float v = rate[i];
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
                if (s_v[k] > bv || (s_v[k] == bv && s_i[k] < bi)) { bv = s_v[k]; bi = s_i[k]; }
            s_stop = !(bi < n && bv > max_rate);
            if (!s_stop) { rate[bi] = 0.0f; mask[bi] = 0.0f; }
        }
        __syncthreads();
        if (s_stop) break;
    }
}

}  // namespace tef

using namespace tef;
#define ST ((cudaStream_t)stream)
#define TEF_GRID(n) (unsigned)(((n) + kThreads - 1) / kThreads)

#define TEF_GRID4(n) (unsigned)(((n) + (long)kThreads * kEvPerThread - 1) / ((long)kThreads * kEvPerThread))
static int aligned16(const void *a, const void *b, const void *c, const void *d = nullptr) {
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15u) == 0;
}

extern "C" int tef_events_to_image(const float *xs, const float *ys, const float *ps, float *img, long n, int H, int W, int accumulate, int *oob,
                                   void *stream) {
    if (n < 0 || H < 1 || W < 1 || !img) return TEF_EINVAL;
    if (!accumulate && n > 0x7fffffffl) return TEF_ELIMIT;                // event indices travel through an int image
    cudaMemsetAsync(img, accumulate ? 0 : 0xff, sizeof(float) * (long)H * W, ST);     // 0xffffffff = -1: no event yet
    if (n > 0 && (!xs || !ys || !ps)) return TEF_EINVAL;
    ProfScope pr(K_ENCODING, ST);
    if (accumulate) {
        if (n == 0) return 0;
        to_image_kernel<<<TEF_GRID4(n), kThreads, 0, ST>>>(xs, ys, ps, img, n, H, W, aligned16(xs, ys, ps), oob);
    } else {
        if (n > 0) last_writer_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>(xs, ys, (int *)img, n, H, W, oob);
        put_winner_kernel<<<TEF_GRID((long)H * W), kThreads, 0, ST>>>(ps, img, (long)H * W);
    }
    return (int)cudaGetLastError();
}
extern "C" int tef_events_to_channels(const float *xs, const float *ys, const float *ps, float *out, long n, int H, int W, int *oob, void *stream) {
    if (n < 0 || H < 1 || W < 1 || !out) return TEF_EINVAL;
    cudaMemsetAsync(out, 0, sizeof(float) * 2 * (long)H * W, ST);
    if (n == 0) return 0;
    if (!xs || !ys || !ps) return TEF_EINVAL;
    ProfScope pr(K_ENCODING, ST);
    to_channels_kernel<<<TEF_GRID4(n), kThreads, 0, ST>>>(xs, ys, ps, out, n, H, W, aligned16(xs, ys, ps), oob);
    return (int)cudaGetLastError();
}
extern "C" int tef_events_to_voxel(const float *xs, const float *ys, const float *ts, const float *ps, float *out, long n, int bins, int H, int W,
                                   int *oob, void *stream) {
    if (n < 0 || H < 1 || W < 1 || bins < 1 || !out) return TEF_EINVAL;
    cudaMemsetAsync(out, 0, sizeof(float) * (long)bins * H * W, ST);
    if (n == 0) return 0;
    if (!xs || !ys || !ts || !ps) return TEF_EINVAL;
    ProfScope pr(K_ENCODING, ST);
    to_voxel_kernel<<<TEF_GRID4(n), kThreads, 0, ST>>>(xs, ys, ts, ps, out, n, bins, H, W, aligned16(xs, ys, ts, ps), oob);
    return (int)cudaGetLastError();
}
extern "C" int tef_events_to_channels_batched(const float *events, float *out, int B, int N, int H, int W, void *stream) {
    if (B < 1 || N < 0 || H < 1 || W < 1 || !out) return TEF_EINVAL;
    cudaMemsetAsync(out, 0, sizeof(float) * (long)B * 2 * H * W, ST);
    if (N == 0) return 0;
    if (!events) return TEF_EINVAL;
    ProfScope pr(K_ENCODING, ST);
    to_channels_batched_kernel<<<TEF_GRID((long)B * N), kThreads, 0, ST>>>((const float4 *)events, out, B, N, H, W);
    return (int)cudaGetLastError();
}
extern "C" int tef_get_hot_event_mask(float *event_rate, float *mask, int H, int W, int idx, int max_px, int min_obvs, float max_rate, void *stream) {
    if (!event_rate || !mask || H < 1 || W < 1 || max_px < 0) return TEF_EINVAL;
    ProfScope pr(K_ENCODING, ST);
    hot_mask_kernel<<<1, 1024, 0, ST>>>(event_rate, mask, H * W, max_px, max_rate, idx > min_obvs ? 1 : 0);
    return (int)cudaGetLastError();
}
