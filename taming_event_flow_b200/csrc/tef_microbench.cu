// tef_microbench.cu -- measures the two rates that actually bound the CM kernels on this GPU (SURVEY.md §8d asks for a
// measured atomic peak): 16-byte vector reductions (red.global.add.v4.f32) and 8-byte gathers, both on an L2-resident
// buffer, with uniformly random addresses ("spread") or addresses confined to a 4 KB window per warp ("local", what
// tile-sorted events produce).  bench.py runs them once, outside the timed region, and reports the CM kernels' achieved
// lane-op rates against them.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }

// mode 0: spread, 1: local.  slots = number of float4 slots in buf (power of two)
__global__ void __launch_bounds__(kThreads) mb_red_kernel(float4 *buf, uint32_t slots, int iters, int mode) {
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t warp_base = ((blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 40503u * 256u) & (slots - 1);
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = lcg(s) >> 8;
        uint32_t a = mode == 0 ? (r & (slots - 1)) : ((warp_base + (r & 255u) + (uint32_t)i * 64u) & (slots - 1));
        // access-pattern probes (scripts/red_patterns.py): 2 = 32 consecutive 16-byte slots per warp, 3 = one slot for the whole warp,
        // 4 = lane pairs share a slot, 5 = lane pairs share a 32-byte sector (different halves), 6 = lane quads share a slot
        const uint32_t lane = threadIdx.x & 31u;
        if (mode == 2) a = (warp_base + lane + (uint32_t)i * 32u) & (slots - 1);
        else if (mode == 3) a = (warp_base + (uint32_t)i) & (slots - 1);
        else if (mode >= 4) {
            const uint32_t g = mode == 6 ? (lane >> 2) : (lane >> 1);       // modes 4, 5, 7: pairs
            uint32_t h = (warp_base * 31u + g * 2654435761u + (uint32_t)i * 40503u);      // same value for the lanes of a group
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            const uint32_t slot = (warp_base + (h & 255u) + (uint32_t)i * 64u) & (slots - 1);
            a = mode == 5 ? ((slot & ~1u) | (lane & 1u)) : slot;
        }
        if (mode >= 7) {
            // software merge of adjacent equal addresses, as a kernel would do it: mode 7 = pairs always equal (upper bound of the
            // benefit), mode 8 = random slots, nothing to merge (pure cost of the shuffles)
            if (mode == 8) a = (warp_base + (r & 255u) + (uint32_t)i * 64u) & (slots - 1);
            float v0 = 1.0f + (float)lane, v1 = 0.5f, v2 = 0.25f * (float)i, v3 = 0.125f;
            const uint32_t nk = __shfl_down_sync(0xffffffffu, a, 1);
            const float n0 = __shfl_down_sync(0xffffffffu, v0, 1), n1 = __shfl_down_sync(0xffffffffu, v1, 1);
            const float n2 = __shfl_down_sync(0xffffffffu, v2, 1), n3 = __shfl_down_sync(0xffffffffu, v3, 1);
            const bool take = !(lane & 1u) && nk == a;                  // even lane absorbs its odd neighbour
            const uint32_t taken = __ballot_sync(0xffffffffu, take);
            const bool gone = (lane & 1u) && ((taken >> (lane - 1u)) & 1u);
            if (take) { v0 += n0; v1 += n1; v2 += n2; v3 += n3; }
            if (!gone) red_add_v4(reinterpret_cast<float2 *>(buf + a), v0, v1, v2, v3);
            continue;
        }
        red_add_v4(reinterpret_cast<float2 *>(buf + a), 1.0f, 0.5f, 0.25f, 0.125f);
    }
}
__global__ void __launch_bounds__(kThreads) mb_gather_kernel(const float2 *__restrict__ buf, uint32_t slots, int iters, int mode, float *sink) {
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 777u;
    const uint32_t warp_base = ((blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 40503u * 512u) & (slots - 1);
    float acc = 0.f;
    for (int i = 0; i < iters; i += 4) {
        float2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t r = lcg(s) >> 8;
            const uint32_t a = mode == 0 ? (r & (slots - 1)) : ((warp_base + (r & 511u) + (uint32_t)(i + k) * 128u) & (slots - 1));
            v[k] = __ldg(buf + a);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += v[k].x + v[k].y;
    }
    if (acc == 123.456f) *sink = acc;
}

// 16-byte gathers: what the event kernels issue since the dual-phase layouts (one tap row / corner pair per load)
__global__ void __launch_bounds__(kThreads) mb_gather16_kernel(const float4 *__restrict__ buf, uint32_t slots, int iters, int mode, float *sink) {
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 777u;
    const uint32_t warp_base = ((blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 40503u * 256u) & (slots - 1);
    float acc = 0.f;
    for (int i = 0; i < iters; i += 4) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t r = lcg(s) >> 8;
            const uint32_t a = mode == 0 ? (r & (slots - 1)) : ((warp_base + (r & 255u) + (uint32_t)(i + k) * 64u) & (slots - 1));
            v[k] = __ldg(buf + a);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    if (acc == 123.456f) *sink = acc;
}

}  // namespace tef

using namespace tef;

// kind 0: red.v4, kind 1: 8-byte gather, kind 2: 16-byte gather.  buf: at least `bytes` (power of two) of device memory.  Launches
// 148*8 CTAs x 256 threads x iters operations; returns 0 and the operation count through *ops.
extern "C" int tef_microbench(int kind, int mode, void *buf, long bytes, int iters, long *ops, void *stream) {
    if (!buf || bytes < (1 << 20) || (bytes & (bytes - 1)) || iters < 4) return TEF_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = 148 * 8;
    ProfScope ps(K_MICROBENCH, st);
    if (kind == 0) mb_red_kernel<<<ctas, kThreads, 0, st>>>((float4 *)buf, (uint32_t)(bytes / 16), iters, mode);
    else if (kind == 1) mb_gather_kernel<<<ctas, kThreads, 0, st>>>((const float2 *)buf, (uint32_t)(bytes / 8), iters & ~3, mode, (float *)buf);
    else if (kind == 2) mb_gather16_kernel<<<ctas, kThreads, 0, st>>>((const float4 *)buf, (uint32_t)(bytes / 16), iters & ~3, mode, (float *)buf);
    else return TEF_EINVAL;
    if (ops) *ops = (long)ctas * kThreads * (kind == 0 ? iters : (iters & ~3));
    return (int)cudaGetLastError();
}
