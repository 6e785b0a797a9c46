// registry.h -- host-side table of compiled kernel instantiations.
#pragma once
#include <cuda_runtime.h>
#include "cyrk_defs.h"

namespace cyrk {

enum FpMode : int { FP_FAST = 0, FP_STRICT = 1, FP_COUNT = 2 };
constexpr int RHS_TABLE_SIZE = 16;   // built-in ids 0..15
constexpr int EV_TABLE_SIZE = 10;    // built-in event ids 0..9 (problems_builtin.cuh)

struct KernelEntry {
    cudaError_t (*launch)(const BatchParams&, int grid, int block, cudaStream_t);
    cudaError_t (*occupancy)(int* ctas_per_sm, int block);
    const char* name;
};

struct ProblemEntry {
    bool present;
    int ny, nextra, nargs_dbl;
    const char* struct_name;             // the problem's struct in problems_builtin.cuh (feature builds are compiled on demand)
    KernelEntry solve[3][VAR_COUNT];     // [method - 3][variant]
    cudaError_t (*dense_eval[3])(const DenseEvalParams&, int grid, int block, cudaStream_t);
    cudaError_t (*prep[3])(const BatchParams&, double* prep, int grid, int block, cudaStream_t);
    cudaError_t (*step_dense[3])(const StepDenseParams&, int grid, int block, cudaStream_t);
    cudaError_t (*event_roots[3])(const BatchParams&, int grid, int block, cudaStream_t);
};

// table[fp_mode][rhs_id]
ProblemEntry* registry_slot(int fp_mode, int rhs_id);
const ProblemEntry* registry_lookup(int fp_mode, int rhs_id);
void registry_init();   // idempotent

}  // namespace cyrk
