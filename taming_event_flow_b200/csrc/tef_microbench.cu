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
        if (mode == 9 || mode == 10) {
            // tile-sorted events, ~3 per pixel: neighbouring lanes reduce into the same or the next pixel pair (xl, xl+1) of one
            // image row.  mode 9 = every pair xl has its own 16-byte slot in ONE plane (pairs xl and xl+1 are adjacent: a 32-byte
            // sector holds two of them); mode 10 = the dual-phase layout of the CM kernels (odd xl lives in a second plane, so
            // lanes on xl and xl+1 never share a sector)
            const uint32_t xl = (lane / 3u) + ((r >> 4) & 1u) + (((uint32_t)i * 7u) & 127u);
            const uint32_t row = (warp_base + (uint32_t)i * 1024u) & (slots / 2u - 1u) & ~1023u;
            if (mode == 9) a = row + xl;
            else a = ((xl & 1u) ? slots / 2u : 0u) + row + ((xl + (xl & 1u)) >> 1);
            a &= slots - 1u;
        }
        if (mode >= 7 && mode <= 8) {
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

// 2x2 neighbourhood fetch (bilinear taps of a flow map / corner quad of a gradient image) for tile-sorted-like positions:
// mode 0 = the dual-phase layout: two 16-byte gathers, rows y and y+1 of the phase plane (x0 & 1);
// mode 1 = a quad-phase layout: the four pixels of the cell (y0, x0) are one 32-byte record in the plane (x0 & 1, y0 & 1): one
//          256-bit gather, one sector.
// One op = one 2x2 fetch.  Positions: three neighbouring lanes share a pixel, +-1 pixel jitter in x and y, rows of 640 pixels.
__global__ void __launch_bounds__(kThreads) mb_quad_kernel(const float4 *__restrict__ buf, uint32_t slots, int iters, int mode, float *sink) {
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 99u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wid = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const uint32_t Wp = 642u, plane = 481u * Wp;                     // float2 pixels per phase plane (dual-phase)
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = lcg(s) >> 8;
        const uint32_t x0 = (wid * 11u + (uint32_t)i * 5u) % 600u + lane / 3u + (r & 1u);
        const uint32_t y0 = (wid * 7u + (uint32_t)i * 3u) % 470u + ((r >> 1) & 3u);
        const uint32_t img = ((wid >> 3) * 40503u + (uint32_t)i) & 7u;            // eight maps in flight (passes)
        if (mode == 0) {
            const uint32_t px = x0 & 1u;
            const uint32_t e = img * 2u * plane + px * plane + y0 * Wp + x0 + px;          // float2 index, even
            const float4 a = __ldg(buf + ((e >> 1) & (slots - 1)));
            const float4 b = __ldg(buf + (((e + Wp) >> 1) & (slots - 1)));
            acc += (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
        } else {
            const uint32_t px = x0 & 1u, py = y0 & 1u;
            const uint32_t cells_x = Wp / 2u, qplane = 241u * cells_x;                      // 32-byte cells per quad plane
            const uint32_t c = img * 4u * qplane + (py * 2u + px) * qplane + ((y0 + py) >> 1) * cells_x + ((x0 + px) >> 1);
            const float4 *q = buf + ((2u * c) & (slots - 1) & ~1u);
            float v0, v1, v2, v3, v4, v5, v6, v7;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3), "=f"(v4), "=f"(v5), "=f"(v6), "=f"(v7) : "l"(q));
            acc += (v0 + v1) + (v2 + v3) + (v4 + v5) + (v6 + v7);
        }
    }
    if (acc == 123.456f) *sink = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-privatised accumulation in shared memory (the design BASELINE.json's north_star sketches): every thread adds a
// 16-byte update (4 values) to a slot of a CTA-private patch with shared-memory atomics; after `k` updates per thread the
// patch is flushed to the global image with COALESCED red.v4 (consecutive slots) and cleared.  ATOM 0 = fp32 atomicAdd
// (sm_100a: ATOMS.CAST.SPIN compare-and-swap loop), 1 = native 32-bit integer ATOMS.ADD (fixed point 2^-23, which is not
// within the reference's fp32 tolerance for small weights: shown as the upper bound of what shared memory could give),
// 2 = 64-bit integer (fixed point 2^-40 like the deterministic mode; sm_100a: ATOMS.CAST.SPIN.64).
// pattern 0: uniformly random slots of the patch; 1: tile-sorted-like (three neighbouring lanes per pixel, +-1 jitter).
// The rate returned is ORIGINAL 16-byte updates per second, directly comparable with the red.v4 lane rates above.
// ---------------------------------------------------------------------------------------------------------------------
template <int ATOM>
__global__ void __launch_bounds__(kThreads) mb_smem_kernel(float4 *buf, uint32_t slots, int iters, int patch_slots, int k, int pattern) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nval = patch_slots * 4;
    float *pf = reinterpret_cast<float *>(smem_raw);
    unsigned *pu = reinterpret_cast<unsigned *>(smem_raw);
    unsigned long long *pl = reinterpret_cast<unsigned long long *>(smem_raw);
    for (int q = threadIdx.x; q < nval; q += kThreads) { if (ATOM == 2) pl[q] = 0ull; else pu[q] = 0u; }
    __syncthreads();
    uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 4242u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t cta_base = (blockIdx.x * 40503u * (uint32_t)patch_slots) & (slots - 1);
    for (int i = 0; i < iters; i += k) {
        for (int j = 0; j < k; ++j) {
            const uint32_t r = lcg(s) >> 8;
            uint32_t a = r;
            if (pattern == 1) a = warp * 40u + (uint32_t)(i + j) * 13u + lane / 3u + ((r >> 4) & 1u);
            a &= (uint32_t)patch_slots - 1u;
            const float v[4] = { 1.0f, 0.5f, 0.25f + (float)lane * 0.001f, 0.125f };
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (ATOM == 0) atomicAdd(pf + a * 4 + c, v[c]);
                else if (ATOM == 1) atomicAdd(pu + a * 4 + c, (unsigned)__float2uint_rn(v[c] * 8388608.0f));
                else atomicAdd(pl + a * 4 + c, (unsigned long long)__double2ll_rn((double)v[c] * kFixScale));
            }
        }
        __syncthreads();
        for (int q = threadIdx.x; q < patch_slots; q += kThreads) {          // coalesced flush: consecutive 16-byte slots
            float o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (ATOM == 0) { o[c] = pf[q * 4 + c]; pf[q * 4 + c] = 0.0f; }
                else if (ATOM == 1) { o[c] = (float)pu[q * 4 + c] * (1.0f / 8388608.0f); pu[q * 4 + c] = 0u; }
                else { o[c] = from_fix((long long)pl[q * 4 + c]); pl[q * 4 + c] = 0ull; }
            }
            if (o[0] != 0.0f || o[2] != 0.0f)
                red_add_v4(reinterpret_cast<float2 *>(buf + ((cta_base + (uint32_t)i * 64u + (uint32_t)q) & (slots - 1))), o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
    }
}

}  // namespace tef

using namespace tef;

// kind 0: red.v4, kind 1: 8-byte gather, kind 2: 16-byte gather, kind 4: 2x2 fetch (mode 0: two 16-byte gathers, mode 1: one 32-byte gather).  buf: at least `bytes` (power of two) of device memory.  Launches
// 148*8 CTAs x 256 threads x iters operations; returns 0 and the operation count through *ops.
extern "C" int tef_microbench(int kind, int mode, void *buf, long bytes, int iters, long *ops, void *stream) {
    if (!buf || bytes < (1 << 20) || (bytes & (bytes - 1)) || iters < 4) return TEF_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = 148 * 8;
    ProfScope ps(K_MICROBENCH, st);
    if (kind == 0) mb_red_kernel<<<ctas, kThreads, 0, st>>>((float4 *)buf, (uint32_t)(bytes / 16), iters, mode);
    else if (kind == 1) mb_gather_kernel<<<ctas, kThreads, 0, st>>>((const float2 *)buf, (uint32_t)(bytes / 8), iters & ~3, mode, (float *)buf);
    else if (kind == 2) mb_gather16_kernel<<<ctas, kThreads, 0, st>>>((const float4 *)buf, (uint32_t)(bytes / 16), iters & ~3, mode, (float *)buf);
    else if (kind == 4) mb_quad_kernel<<<ctas, kThreads, 0, st>>>((const float4 *)buf, (uint32_t)(bytes / 16), iters, mode, (float *)buf);
    else return TEF_EINVAL;
    if (ops) *ops = (long)ctas * kThreads * ((kind == 0 || kind == 4) ? iters : (iters & ~3));
    return (int)cudaGetLastError();
}

// CTA-privatised shared-memory accumulation + coalesced flush (see mb_smem_kernel).  atom: 0 fp32 CAS loop, 1 native u32, 2 u64;
// patch_slots: 16-byte slots per CTA patch (power of two, <= 4096); k: updates per thread between flushes; pattern: 0 random, 1 sorted-like.
extern "C" int tef_microbench_smem(int atom, int patch_slots, int k, int pattern, void *buf, long bytes, int iters, long *ops, void *stream) {
    if (!buf || bytes < (1 << 20) || (bytes & (bytes - 1)) || iters < 1 || k < 1 || patch_slots < 32 || patch_slots > 4096 ||
        (patch_slots & (patch_slots - 1)) || atom < 0 || atom > 2)
        return TEF_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = 148 * 8;
    iters = (iters + k - 1) / k * k;
    const size_t smem = (size_t)patch_slots * 4 * (atom == 2 ? 8 : 4);
    ProfScope ps(K_MICROBENCH, st);
    if (atom == 0) {
        cudaFuncSetAttribute(mb_smem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mb_smem_kernel<0><<<ctas, kThreads, smem, st>>>((float4 *)buf, (uint32_t)(bytes / 16), iters, patch_slots, k, pattern);
    } else if (atom == 1) {
        cudaFuncSetAttribute(mb_smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mb_smem_kernel<1><<<ctas, kThreads, smem, st>>>((float4 *)buf, (uint32_t)(bytes / 16), iters, patch_slots, k, pattern);
    } else {
        cudaFuncSetAttribute(mb_smem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mb_smem_kernel<2><<<ctas, kThreads, smem, st>>>((float4 *)buf, (uint32_t)(bytes / 16), iters, patch_slots, k, pattern);
    }
    if (ops) *ops = (long)ctas * kThreads * iters;
    return (int)cudaGetLastError();
}
