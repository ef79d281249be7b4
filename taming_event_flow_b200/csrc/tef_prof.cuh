// tef_prof.cuh -- launch counting and optional per-kernel CUDA-event timing.
// Every kernel launch of the library sits inside a ProfScope: the launch counter is always
// maintained (bench.py reports it as gpu_launches); when timing is enabled the scope records
// a CUDA event pair on the launching stream, which bench.py reads back for the roofline.
#pragma once
#include <cuda_runtime.h>

namespace tef {

enum KernelId {
    K_STAGE_EVENTS = 0, K_PACK_FLOW, K_UNPACK_GRAD, K_ITER_FWD, K_IWE_REDUCE, K_FINALIZE, K_IWE_GRAD, K_ITER_BWD,
    K_SORT_HIST, K_SORT_SCAN, K_SORT_SCATTER, K_LIN_FWD, K_LIN_BWD, K_PRIMITIVE, K_ENCODING, K_MICROBENCH, K_LOADER, K_VALIDATION, K_SMOOTH, K_NETWORK, K_COUNT
};

// The event pair lives in the scope itself and joins the pending list only in the destructor, so a concurrent
// tef_prof_read / tef_prof_reset (the autograd backward runs on a worker thread) can never invalidate it.  With timing
// off a scope costs one relaxed atomic increment: no lock on the launch path.
struct ProfScope {
    int id; cudaStream_t st; cudaEvent_t a, b; int dev;     // dev < 0: not timed
    ProfScope(int id, cudaStream_t st);
    ~ProfScope();
    ProfScope(const ProfScope &) = delete;
    ProfScope &operator=(const ProfScope &) = delete;
};

}  // namespace tef
