#include "tef_prof.cuh"
#include "../../include/tef_b200.h"
#include <mutex>
#include <vector>

namespace tef {

static const char *kNames[K_COUNT] = {
    "stage_events_kernel", "pack_flow_kernel", "unpack_grad_kernel", "iter_fwd_kernel", "iwe_reduce_kernel", "finalize_kernel",
    "iwe_grad_kernel", "iter_bwd_kernel", "sort_hist_kernel", "sort_scan_kernels", "sort_scatter_kernel", "linear_fwd_kernel", "linear_bwd_kernel", "primitive_kernels",
    "encoding_kernels", "microbench_kernels", "loader_kernels", "validation_kernels", "smoothing_kernels"
};

struct Pair { cudaEvent_t a, b; int id; };
static std::mutex g_mu;
static bool g_on = false;
static long g_launches[K_COUNT] = {0};
static double g_ms[K_COUNT] = {0};
static long g_timed[K_COUNT] = {0};
static std::vector<Pair> g_pending;
static std::vector<Pair> g_free;

ProfScope::ProfScope(int id_, cudaStream_t st_) : id(id_), st(st_), slot(-1) {
    std::lock_guard<std::mutex> lk(g_mu);
    ++g_launches[id];
    if (!g_on) return;
    Pair p;
    if (!g_free.empty()) { p = g_free.back(); g_free.pop_back(); }
    else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
    p.id = id;
    cudaEventRecord(p.a, st);
    g_pending.push_back(p);
    slot = (int)g_pending.size() - 1;
}
ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEventRecord(g_pending[slot].b, st);
}

static void drain() {
    for (auto &p : g_pending) {
        cudaEventSynchronize(p.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { g_ms[p.id] += ms; ++g_timed[p.id]; }
        g_free.push_back(p);
    }
    g_pending.clear();
}

}  // namespace tef

using namespace tef;

extern "C" void tef_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!on) drain();
    if (on && g_free.size() < 2048) {               // create the event pool outside any timed region
        while (g_free.size() < 2048) { Pair p; cudaEventCreate(&p.a); cudaEventCreate(&p.b); p.id = 0; g_free.push_back(p); }
    }
    g_on = on != 0;
}
extern "C" void tef_prof_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    drain();
    for (int i = 0; i < K_COUNT; ++i) { g_launches[i] = 0; g_ms[i] = 0; g_timed[i] = 0; }
}
extern "C" int tef_prof_num_kernels(void) { return K_COUNT; }
extern "C" const char *tef_prof_name(int id) { return (id >= 0 && id < K_COUNT) ? kNames[id] : ""; }
extern "C" int tef_prof_read(int id, double *ms_total, long *timed_launches, long *launches) {
    if (id < 0 || id >= K_COUNT) return TEF_EINVAL;
    std::lock_guard<std::mutex> lk(g_mu);
    drain();
    if (ms_total) *ms_total = g_ms[id];
    if (timed_launches) *timed_launches = g_timed[id];
    if (launches) *launches = g_launches[id];
    return 0;
}
extern "C" long tef_launch_count(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    long n = 0;
    for (int i = 0; i < K_COUNT; ++i) n += g_launches[i];
    return n;
}
