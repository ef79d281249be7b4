// tef_cm_sort.cu -- device counting sort of the staged events by pixel tile.
//
// Why: with events in arrival (time) order every bilinear gather and every image reduction of a warp
// touches ~32 different L2 sectors, and both loss kernels saturate the L1/L2 transaction rate
// (profiles/r1_a_*: 30.7 sectors per load request, 15.4 per RED request, L1 hit 21 %).  The loss is a sum
// over events, so their order is free: sorting each (set, pass) segment by (sample, 16x8-pixel tile, pixel, polarity)
// makes the lanes of a warp start from the same few pixels and, the flow being smooth, stay neighbours
// along the whole warping chain.  Padding rows (mask 0,0) are dropped on the way.
//
// Three steps: histogram over per-pixel bins (1 L2 atomic per event; fused into tef_update_pass when the descriptor says
// hist_done, otherwise done here), exclusive scan of the bins, scatter (1 returning atomic per event).  Output rows are 32-byte records
// (ts, y, x, sample index | mask+, mask-, 0, 0): the polarity column is not needed by the loss, the mask carries it.
#include "tef_cm_common.cuh"
#include "tef_prof.cuh"

namespace tef {

__device__ __forceinline__ int bin_of(const CmParams &p, int seg, int b, float y, float x, float2 m) {
    return sort_bin(p.seg.first_bin[seg], p.sort.tiles_x, p.sort.tiles, p.sort.H, p.sort.W, b, y, x, m.x == 0.0f);
}

// Both kernels walk a segment with kSortIlp rows per thread (strided by the CTA size, so every access stays
// coalesced): the returning atomic of the scatter has ~1 us latency and one row per thread leaves the
// kernel latency-bound (profiles/r1_c: 11 % issue, long-scoreboard 103).
constexpr int kSortIlp = 4;

__device__ __forceinline__ bool sort_rows(const CmParams &p, int &sg, int &n, long (&row)[kSortIlp], long &nrows) {
    sg = 0;
    const int blk = blockIdx.x;
    while (blk >= p.sort.blk_off[sg + 1]) ++sg;
    n = p.seg.n[sg];
    nrows = (long)p.B * n;
    const long base = (long)(blk - p.sort.blk_off[sg]) * (kThreads * kSortIlp) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kSortIlp; ++k) row[k] = base + (long)k * kThreads;
    return base < nrows;
}

__global__ void __launch_bounds__(kThreads) sort_hist_kernel(const __grid_constant__ CmParams p) {
    int sg, n; long row[kSortIlp], nrows;
    if (!sort_rows(p, sg, n, row, nrows)) return;
    float2 m[kSortIlp]; float4 e[kSortIlp];
#pragma unroll
    for (int k = 0; k < kSortIlp; ++k) {
        const bool ok = row[k] < nrows;
        m[k] = ok ? __ldg(p.seg.mk[sg] + row[k]) : make_float2(0.f, 0.f);
        e[k] = ok ? __ldg(p.seg.ev[sg] + row[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kSortIlp; ++k) {
        if (m[k].x == 0.0f && m[k].y == 0.0f) continue;            // padding rows are dropped
        atomicAdd(p.sort.bins + bin_of(p, sg, (int)(row[k] / n), e[k].y, e[k].z, m[k]), 1);
    }
}

__global__ void __launch_bounds__(kThreads) sort_scatter_kernel(const __grid_constant__ CmParams p) {
    int sg, n; long row[kSortIlp], nrows;
    if (!sort_rows(p, sg, n, row, nrows)) return;
    float2 m[kSortIlp]; float4 e[kSortIlp]; int dst[kSortIlp];
#pragma unroll
    for (int k = 0; k < kSortIlp; ++k) {
        const bool ok = row[k] < nrows;
        m[k] = ok ? __ldg(p.seg.mk[sg] + row[k]) : make_float2(0.f, 0.f);
        e[k] = ok ? __ldg(p.seg.ev[sg] + row[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kSortIlp; ++k) {
        dst[k] = -1;
        if (m[k].x == 0.0f && m[k].y == 0.0f) continue;
        const int b = (int)(row[k] / n);
        dst[k] = atomicAdd(p.sort.bins + bin_of(p, sg, b, e[k].y, e[k].z, m[k]), 1);
        e[k].w = __int_as_float(b);
    }
    // one 256-bit store per event (STG.E.ENL2.256): a full, aligned 32-byte sector, so the scattered writes never
    // leave partially written sectors behind in L2 (profiles/r1_e: 16 B + 8 B stores cost 150 MB of DRAM fill reads)
#pragma unroll
    for (int k = 0; k < kSortIlp; ++k)
        if (dst[k] >= 0)
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p.sort.rec + 2 * (long)dst[k]), "f"(e[k].x), "f"(e[k].y),
                         "f"(e[k].z), "f"(e[k].w), "f"(m[k].x), "f"(m[k].y), "f"(0.0f), "f"(0.0f)
                         : "memory");
}

// ---- exclusive scan of the bins (int), three small kernels ------------------------------------------
// 16 bins per thread as four 16-byte accesses (the bin array is as large as the event data at 1 M events per window: 6.1 M bins,
// 24.6 MB, read twice and written once per step).
constexpr int kScanPerThread = 16;
constexpr int kScanChunk = kThreads * kScanPerThread;     // elements per CTA

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_sums[kThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int s = lane < kThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        if (lane < kThreads / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const int base = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[kThreads / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// the thread's 16 consecutive bins (zero beyond nbins); `bins` is 16-byte aligned (start of an allocation)
__device__ __forceinline__ void load_bins16(const int *__restrict__ bins, long base, long nbins, int (&v)[kScanPerThread]) {
    if (base + kScanPerThread <= nbins) {
#pragma unroll
        for (int q = 0; q < kScanPerThread / 4; ++q) {
            const int4 t = *reinterpret_cast<const int4 *>(bins + base + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanPerThread; ++k) v[k] = (base + k < nbins) ? bins[base + k] : 0;
    }
}

__global__ void __launch_bounds__(kThreads) scan_sums_kernel(const int *__restrict__ bins, int *__restrict__ sums, long nbins) {
    const long base = (long)blockIdx.x * kScanChunk + threadIdx.x * kScanPerThread;
    int v[kScanPerThread], s = 0;
    load_bins16(bins, base, nbins, v);
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) s += v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ int ws[kThreads / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < kThreads / 32; ++k) t += ws[k]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(kThreads) scan_top_kernel(int *__restrict__ sums, int nsums) {
    int carry = 0;
    for (int base = 0; base < nsums; base += kThreads) {
        const int i = base + threadIdx.x;
        const int v = i < nsums ? sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        if (i < nsums) sums[i] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(kThreads) scan_apply_kernel(int *__restrict__ bins, const int *__restrict__ sums, long nbins) {
    const long base = (long)blockIdx.x * kScanChunk + threadIdx.x * kScanPerThread;
    int v[kScanPerThread], s = 0;
    load_bins16(bins, base, nbins, v);
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) s += v[k];
    int run = block_exclusive_scan(s, nullptr) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) { const int c = v[k]; v[k] = run; run += c; }
    if (base + kScanPerThread <= nbins) {
#pragma unroll
        for (int q = 0; q < kScanPerThread / 4; ++q) *reinterpret_cast<int4 *>(bins + base + 4 * q) = make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanPerThread; ++k) if (base + k < nbins) bins[base + k] = v[k];
    }
}

}  // namespace tef

using namespace tef;

// Sort every segment of `p` (as laid out by fill_params) into p.sort.ev / p.sort.mk.  Afterwards
// bins[i] holds the END of bin i, so segment s occupies rows [bins[first_bin(s)-1] (or 0), bins[last_bin(s)]).
int tef_sort_events(const CmParams &p, cudaStream_t st) {
    const int nblk = p.sort.blk_off[p.seg.nseg];
    const long nbins = p.sort.nbins;
    if (!p.sort.bins || !p.sort.sums || (nblk > 0 && !p.sort.rec)) return TEF_EINVAL;
    if (!p.hist_done) cudaMemsetAsync(p.sort.bins, 0, sizeof(int) * (nbins + 1), st);     // else counted by tef_update_pass
    if (nblk == 0) return (int)cudaGetLastError();
    const int nchunks = (int)((nbins + kScanChunk - 1) / kScanChunk);
    if (!p.hist_done) { ProfScope ps(K_SORT_HIST, st); sort_hist_kernel<<<nblk, kThreads, 0, st>>>(p); }
    { ProfScope ps(K_SORT_SCAN, st); scan_sums_kernel<<<nchunks, kThreads, 0, st>>>(p.sort.bins, p.sort.sums, nbins); }
    { ProfScope ps(K_SORT_SCAN, st); scan_top_kernel<<<1, kThreads, 0, st>>>(p.sort.sums, nchunks); }
    { ProfScope ps(K_SORT_SCAN, st); scan_apply_kernel<<<nchunks, kThreads, 0, st>>>(p.sort.bins, p.sort.sums, nbins); }
    { ProfScope ps(K_SORT_SCATTER, st); sort_scatter_kernel<<<nblk, kThreads, 0, st>>>(p); }
    return (int)cudaGetLastError();
}
