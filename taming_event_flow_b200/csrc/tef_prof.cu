#include "tef_prof.cuh"
#include "../../include/tef_b200.h"
#include <atomic>
#include <mutex>
#include <vector>

namespace tef {

static const char *kNames[K_COUNT] = {
    "stage_events_kernel", "pack_flow_kernel", "unpack_grad_kernel", "iter_fwd_kernel", "iwe_reduce_kernel", "finalize_kernel",
    "iwe_grad_kernel", "iter_bwd_kernel", "sort_hist_kernel", "sort_scan_kernels", "sort_scatter_kernel", "linear_fwd_kernel", "linear_bwd_kernel", "primitive_kernels",
    "encoding_kernels", "microbench_kernels", "loader_kernels", "validation_kernels", "smoothing_kernels", "network_kernels"
};

struct Pair { cudaEvent_t a, b; int id, dev; };
constexpr int kMaxDev = 64;
static std::mutex g_mu;                               // guards the pools, the pending list and the timing totals
static std::atomic<bool> g_on{false};
static std::atomic<long> g_launches[K_COUNT];
static double g_ms[K_COUNT] = {0};
static long g_timed[K_COUNT] = {0};
static std::vector<Pair> g_pending;
static std::vector<Pair> g_free[kMaxDev];             // events belong to the device they were created on

ProfScope::ProfScope(int id_, cudaStream_t st_) : id(id_), st(st_), a(nullptr), b(nullptr), dev(-1) {
    g_launches[id].fetch_add(1, std::memory_order_relaxed);
    if (!g_on.load(std::memory_order_acquire)) return;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;   // no timing inside a graph capture
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDev) return;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_free[d].empty()) { a = g_free[d].back().a; b = g_free[d].back().b; g_free[d].pop_back(); }
    }
    if (!a && (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess)) { a = b = nullptr; return; }
    dev = d;
    cudaEventRecord(a, st);
}
ProfScope::~ProfScope() {
    if (dev < 0) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_mu);
    g_pending.push_back(Pair{a, b, id, dev});
}

static void drain() {                                 // caller holds g_mu
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &p : g_pending) {
        if (p.dev != cur) cudaSetDevice(p.dev);
        cudaEventSynchronize(p.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { g_ms[p.id] += ms; ++g_timed[p.id]; }
        g_free[p.dev].push_back(p);
        if (p.dev != cur) cudaSetDevice(cur);
    }
    g_pending.clear();
}

}  // namespace tef

using namespace tef;

extern "C" void tef_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!on) { g_on.store(false, std::memory_order_release); drain(); return; }
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < kMaxDev)        // create the event pool outside any timed region
        while (g_free[d].size() < 2048) { Pair p; cudaEventCreate(&p.a); cudaEventCreate(&p.b); p.id = 0; p.dev = d; g_free[d].push_back(p); }
    g_on.store(true, std::memory_order_release);
}
extern "C" void tef_prof_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    drain();
    for (int i = 0; i < K_COUNT; ++i) { g_launches[i].store(0); g_ms[i] = 0; g_timed[i] = 0; }
}
extern "C" int tef_prof_num_kernels(void) { return K_COUNT; }
extern "C" const char *tef_prof_name(int id) { return (id >= 0 && id < K_COUNT) ? kNames[id] : ""; }
extern "C" int tef_prof_read(int id, double *ms_total, long *timed_launches, long *launches) {
    if (id < 0 || id >= K_COUNT) return TEF_EINVAL;
    std::lock_guard<std::mutex> lk(g_mu);
    drain();
    if (ms_total) *ms_total = g_ms[id];
    if (timed_launches) *timed_launches = g_timed[id];
    if (launches) *launches = g_launches[id].load();
    return 0;
}
extern "C" long tef_launch_count(void) {
    long n = 0;
    for (int i = 0; i < K_COUNT; ++i) n += g_launches[i].load(std::memory_order_relaxed);
    return n;
}
