// tef_encodings.cu -- event encodings of the reference's dataloader/encodings.py as sm_100a kernels.
// HBM-streaming kernels: 12-16 B read per event (coalesced), one or two native fp32 reductions into an
// image that stays in L2.  Counting encodings are exact in fp32 for any order (integers < 2^24);
// the voxel grid sums real weights, so only its summation order differs from the reference.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

// xs.long()/ys.long() truncate toward zero; negative indices wrap like Python indexing (index_put_)
__device__ __forceinline__ bool pixel_of(float xf, float yf, int H, int W, long &px) {
    long x = (long)xf, y = (long)yf;
    if (x < 0) x += W;
    if (y < 0) y += H;
    if (x < 0 || x >= W || y < 0 || y >= H) return false;          // the reference raises IndexError here
    px = y * W + x;
    return true;
}

__global__ void __launch_bounds__(kThreads) to_image_kernel(const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ ps,
                                                            float *__restrict__ img, long n, int H, int W, int accumulate) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    long px;
    if (!pixel_of(xs[i], ys[i], H, W, px)) return;
    if (accumulate) red_add_f32(img + px, ps[i]);                   // dataloader/encodings.py:27
    else img[px] = ps[i];
}

__global__ void __launch_bounds__(kThreads) to_channels_kernel(const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ ps,
                                                               float *__restrict__ out, long n, int H, int W) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    long px;
    if (!pixel_of(xs[i], ys[i], H, W, px)) return;
    const float p = ps[i];
    const float mpos = p < 0.f ? 0.f : (p > 0.f ? 1.f : p);        // :71-76
    const float mneg = p > 0.f ? 0.f : (p < 0.f ? -1.f : p);
    const float vp = p * mpos, vn = p * mneg;                       // :78-79
    if (vp != 0.f) red_add_f32(out + px, vp);
    if (vn != 0.f) red_add_f32(out + (long)H * W + px, vn);
}

__global__ void __launch_bounds__(kThreads) to_voxel_kernel(const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ ts,
                                                            const float *__restrict__ ps, float *__restrict__ out, long n, int bins, int H, int W) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    long px;
    if (!pixel_of(xs[i], ys[i], H, W, px)) return;
    const float t = ts[i] * (float)(bins - 1);                      // :47
    const float p = ps[i];
    // only the bins within distance 1 of t have a non-zero weight (:52); the others add +0
    int b0 = (int)floorf(t) - 1;
    for (int b = max(b0, 0); b <= min(b0 + 3, bins - 1); ++b) {
        const float wgt = fmaxf(0.0f, 1.0f - fabsf(t - (float)b));
        const float v = p * wgt;                                    // :53
        if (v != 0.f) red_add_f32(out + (long)b * H * W + px, v);
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
    if (!pixel_of(e.z, e.y, H, W, px)) return;
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

extern "C" int tef_events_to_image(const float *xs, const float *ys, const float *ps, float *img, long n, int H, int W, int accumulate, void *stream) {
    if (n < 0 || H < 1 || W < 1 || !img) return TEF_EINVAL;
    cudaMemsetAsync(img, 0, sizeof(float) * (long)H * W, ST);
    if (n == 0) return 0;
    if (!xs || !ys || !ps) return TEF_EINVAL;
    ProfScope pr(K_ENCODING, ST);
    to_image_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>(xs, ys, ps, img, n, H, W, accumulate);
    return (int)cudaGetLastError();
}
extern "C" int tef_events_to_channels(const float *xs, const float *ys, const float *ps, float *out, long n, int H, int W, void *stream) {
    if (n < 0 || H < 1 || W < 1 || !out) return TEF_EINVAL;
    cudaMemsetAsync(out, 0, sizeof(float) * 2 * (long)H * W, ST);
    if (n == 0) return 0;
    if (!xs || !ys || !ps) return TEF_EINVAL;
    ProfScope pr(K_ENCODING, ST);
    to_channels_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>(xs, ys, ps, out, n, H, W);
    return (int)cudaGetLastError();
}
extern "C" int tef_events_to_voxel(const float *xs, const float *ys, const float *ts, const float *ps, float *out, long n, int bins, int H, int W,
                                   void *stream) {
    if (n < 0 || H < 1 || W < 1 || bins < 1 || !out) return TEF_EINVAL;
    cudaMemsetAsync(out, 0, sizeof(float) * (long)bins * H * W, ST);
    if (n == 0) return 0;
    if (!xs || !ys || !ts || !ps) return TEF_EINVAL;
    ProfScope pr(K_ENCODING, ST);
    to_voxel_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>(xs, ys, ts, ps, out, n, bins, H, W);
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
