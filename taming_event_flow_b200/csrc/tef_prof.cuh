// tef_prof.cuh -- launch counting and optional per-kernel CUDA-event timing.
// Every kernel launch of the library sits inside a ProfScope: the launch counter is always
// maintained (bench.py reports it as gpu_launches); when timing is enabled the scope records
// a CUDA event pair on the launching stream, which bench.py reads back for the roofline.
#pragma once
#include <cuda_runtime.h>

namespace tef {

enum KernelId {
    K_STAGE_EVENTS = 0, K_PACK_FLOW, K_UNPACK_GRAD, K_ITER_FWD, K_IWE_REDUCE, K_FINALIZE, K_IWE_GRAD, K_ITER_BWD,
    K_SORT_HIST, K_SORT_SCAN, K_SORT_SCATTER, K_LIN_FWD, K_LIN_BWD, K_PRIMITIVE, K_ENCODING, K_MICROBENCH, K_LOADER, K_VALIDATION, K_SMOOTH, K_COUNT
};

struct ProfScope {
    int id; cudaStream_t st; int slot;
    ProfScope(int id, cudaStream_t st);
    ~ProfScope();
};

}  // namespace tef
