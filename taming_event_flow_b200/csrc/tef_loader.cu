// tef_loader.cu -- the loader -> loss contract of the reference (dataloader/base.py) as device kernels.
//
// Upstream formats every window on `self.device`, moves every tensor back to the host (dataloader/h5.py:409-421),
// zero-pads and stacks on the host (custom_collate, base.py:391-434) and uploads 24 B/event again in the training loop
// (train_flow.py:106-116).  Here a window crosses PCIe once, as 8-byte packed events (fp32 raw timestamp + x/y/polarity
// bit field); one kernel turns a ragged batch of them into the padded [B][N][4] event list, the [B][N][2] polarity mask
// and (optionally) the 2-channel count image the network eats.  All arithmetic is the reference's, in fp32:
// ts = (ts - ts[0]) / (ts[-1] - ts[0]) (base.py:168-169), ps = pol*2 - 1 (:167), mask rows (:264-278).
// HBM streaming: 8 B read, 24 B written per event, plus two fp32 reductions per event into the L2-resident count image.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

// create_polarity_mask (base.py:264-278) on one polarity value: row 0 = (ps<0 -> 0, ps>0 -> 1), row 1 = (ps<0 -> -1,
// ps>0 -> 0) * -1; the -0.0 the reference produces for positive events is kept
__device__ __forceinline__ float2 polarity_mask(float p) {
    const float r0 = p < 0.f ? 0.f : (p > 0.f ? 1.f : p);
    const float r1 = p < 0.f ? -1.f : (p > 0.f ? 0.f : p);
    return make_float2(r0, r1 * -1.0f);
}

__global__ void __launch_bounds__(kThreads) polarity_mask_kernel(const float *__restrict__ ps, float *__restrict__ mask, long n) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float2 m = polarity_mask(ps[i]);
    mask[i] = m.x;
    mask[n + i] = m.y;
}

// create_mask_encoding (base.py:302-314): sum of the two count channels, positive -> 1
__global__ void __launch_bounds__(kThreads) mask_encoding_kernel(const float *__restrict__ cnt, float *__restrict__ out, long hw, long total) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= total) return;
    const long b = i / hw, px = i - b * hw;
    const float s = cnt[(2 * b) * hw + px] + cnt[(2 * b + 1) * hw + px];
    out[i] = s > 0.0f ? 1.0f : s;
}

// custom_collate (base.py:416-428) for one sample: [C][n] channel-major -> rows [N][C], zero rows from n to N
__global__ void __launch_bounds__(kThreads) collate_kernel(const float *__restrict__ src, float *__restrict__ dst, long n, long N, int C) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= N * C) return;
    const long row = i / C;
    const int c = (int)(i - row * C);
    dst[i] = row < n ? src[(long)c * n + row] : 0.0f;
}

struct Unpacked { float ts, y, x, p; };
__device__ __forceinline__ Unpacked unpack_event(uint2 w) {
    Unpacked e;
    e.ts = __uint_as_float(w.x);
    e.x = (float)(w.y & 0x3fffu);
    e.y = (float)((w.y >> 14) & 0x3fffu);
    e.p = (float)((w.y >> 28) & 1u) * 2.0f - 1.0f;                 // base.py:167
    return e;
}

// event_formatting + create_list_encoding + create_polarity_mask + custom_collate for a ragged batch of packed windows.
// One thread per padded row; cnt (nullable) is events_to_channels of the same events (dataloader/encodings.py:59-81).
__global__ void __launch_bounds__(kThreads) format_events_kernel(const uint2 *__restrict__ packed, const long *__restrict__ offsets, int B, long n_pad,
                                                                 float4 *__restrict__ event_list, float2 *__restrict__ pol_mask,
                                                                 float *__restrict__ cnt, int H, int W) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * n_pad) return;
    const int b = (int)(i / n_pad);
    const long r = i - (long)b * n_pad;
    const long lo = offsets[b], hi = offsets[b + 1];
    float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 mk = make_float2(0.f, 0.f);
    if (r < hi - lo) {
        const Unpacked e = unpack_event(packed[lo + r]);
        const float t0 = __uint_as_float(packed[lo].x), t1 = __uint_as_float(packed[hi - 1].x);
        ev = make_float4((e.ts - t0) / (t1 - t0), e.y, e.x, e.p);   // base.py:168-169, :262 (ts, y, x, p)
        mk = polarity_mask(e.p);
        if (cnt) {
            const int x = (int)e.x, y = (int)e.y;
            if (x < W && y < H) red_add_f32(cnt + ((long)(2 * b + (e.p > 0.f ? 0 : 1)) * H + y) * W + x, 1.0f);
        }
    }
    event_list[i] = ev;
    pol_mask[i] = mk;
}

// A keyed pseudo-random permutation of [0, n): 4-round balanced Feistel network on 2h >= log2(n) bits, cycle-walked
// back into the domain (a bijection for every key, < 4 walks on average).  Replaces a sort of random keys.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t permute_index(uint32_t i, uint32_t n, uint64_t key) {
    int h = 1;
    while ((1ull << (2 * h)) < n) ++h;
    const uint32_t half = (1u << h) - 1u;
    uint32_t x = i;
    do {
        uint32_t l = x >> h, r = x & half;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint32_t f = mix32(r ^ (uint32_t)(key >> (16 * round)) ^ (0x9e3779b9u * (round + 1))) & half;
            const uint32_t nl = r;
            r = l ^ f;
            l = nl;
        }
        x = (l << h) | r;
    } while (x >= n);
    return x;
}

// split_event_list (base.py:347-377) for a padded batch: events whose rank in a random permutation of the sample's
// events is below k carry gradients, the others are detached.  Samples with at most k events keep all of them.
__global__ void __launch_bounds__(kThreads) split_events_kernel(const float4 *__restrict__ ev, const float2 *__restrict__ mk, const long *__restrict__ offsets,
                                                                uint64_t seed, int B, long N, long k,
                                                                float4 *__restrict__ g_ev, float2 *__restrict__ g_mk, long Ng,
                                                                float4 *__restrict__ d_ev, float2 *__restrict__ d_mk, long Nd) {
    const long i = (long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= (long)B * N) return;
    const int b = (int)(i / N);
    const long r = i - (long)b * N;
    const long n = offsets[b + 1] - offsets[b];
    if (r >= n) return;
    if (n <= k) { g_ev[(long)b * Ng + r] = ev[i]; g_mk[(long)b * Ng + r] = mk[i]; return; }
    const uint64_t key = seed ^ ((uint64_t)mix32((uint32_t)b + 0x51ed270bu) << 32 | mix32((uint32_t)b * 0x2545f491u + 1u));
    const long q = permute_index((uint32_t)r, (uint32_t)n, key);
    if (q < k) { g_ev[(long)b * Ng + q] = ev[i]; g_mk[(long)b * Ng + q] = mk[i]; }
    else       { d_ev[(long)b * Nd + q - k] = ev[i]; d_mk[(long)b * Nd + q - k] = mk[i]; }
}

}  // namespace tef

using namespace tef;
#define ST ((cudaStream_t)stream)
#define TEF_GRID(n) (unsigned)(((n) + kThreads - 1) / kThreads)

extern "C" int tef_create_polarity_mask(const float *ps, float *mask, long n, void *stream) {
    if (n < 0) return TEF_EINVAL;
    if (n == 0) return 0;
    if (!ps || !mask) return TEF_EINVAL;
    ProfScope pr(K_LOADER, ST);
    polarity_mask_kernel<<<TEF_GRID(n), kThreads, 0, ST>>>(ps, mask, n);
    return (int)cudaGetLastError();
}

extern "C" int tef_create_mask_encoding(const float *cnt, float *out, int B, int H, int W, void *stream) {
    if (B < 1 || H < 1 || W < 1 || !cnt || !out) return TEF_EINVAL;
    ProfScope pr(K_LOADER, ST);
    const long total = (long)B * H * W;
    mask_encoding_kernel<<<TEF_GRID(total), kThreads, 0, ST>>>(cnt, out, (long)H * W, total);
    return (int)cudaGetLastError();
}

extern "C" int tef_collate_events(const float *src, float *dst, long n, long N, int C, void *stream) {
    if (n < 0 || N < n || C < 1) return TEF_EINVAL;
    if (N == 0) return 0;
    if (!dst || (n > 0 && !src)) return TEF_EINVAL;
    ProfScope pr(K_LOADER, ST);
    collate_kernel<<<TEF_GRID(N * C), kThreads, 0, ST>>>(src, dst, n, N, C);
    return (int)cudaGetLastError();
}

extern "C" int tef_format_events(const void *packed, const long *offsets, int B, long n_pad, float *event_list, float *pol_mask,
                                 float *cnt, int H, int W, void *stream) {
    if (B < 1 || n_pad < 0 || !offsets) return TEF_EINVAL;
    if (cnt) {
        if (H < 1 || W < 1 || H > 16384 || W > 16384) return TEF_EINVAL;
        cudaMemsetAsync(cnt, 0, sizeof(float) * (long)B * 2 * H * W, ST);
    }
    if (n_pad == 0) return 0;
    if (!packed || !event_list || !pol_mask) return TEF_EINVAL;
    ProfScope pr(K_LOADER, ST);
    format_events_kernel<<<TEF_GRID((long)B * n_pad), kThreads, 0, ST>>>((const uint2 *)packed, offsets, B, n_pad, (float4 *)event_list,
                                                                        (float2 *)pol_mask, cnt, H, W);
    return (int)cudaGetLastError();
}

extern "C" int tef_split_events(const float *event_list, const float *pol_mask, const long *offsets, unsigned long long seed, int B, long N, long k,
                                float *g_events, float *g_mask, long Ng, float *d_events, float *d_mask, long Nd, void *stream) {
    if (B < 1 || N < 0 || N > 0x7fffffffl || k < 0 || Ng < 0 || Nd < 0 || !offsets) return TEF_EINVAL;
    if (Ng > 0) {
        if (!g_events || !g_mask) return TEF_EINVAL;
        cudaMemsetAsync(g_events, 0, sizeof(float4) * (long)B * Ng, ST);
        cudaMemsetAsync(g_mask, 0, sizeof(float2) * (long)B * Ng, ST);
    }
    if (Nd > 0) {
        if (!d_events || !d_mask) return TEF_EINVAL;
        cudaMemsetAsync(d_events, 0, sizeof(float4) * (long)B * Nd, ST);
        cudaMemsetAsync(d_mask, 0, sizeof(float2) * (long)B * Nd, ST);
    }
    if (N == 0) return 0;
    if (!event_list || !pol_mask) return TEF_EINVAL;
    ProfScope pr(K_LOADER, ST);
    split_events_kernel<<<TEF_GRID((long)B * N), kThreads, 0, ST>>>((const float4 *)event_list, (const float2 *)pol_mask, offsets, (uint64_t)seed, B, N, k,
                                                                    (float4 *)g_events, (float2 *)g_mask, Ng, (float4 *)d_events, (float2 *)d_mask, Nd);
    return (int)cudaGetLastError();
}
