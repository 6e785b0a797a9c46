// nvrtc_rhs.h -- user right-hand sides compiled at run time (defined in nvrtc_rhs.cpp).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "cyrk_defs.h"

namespace cyrk {
// Registers (and validates by compiling one kernel) a user right-hand side; returns its id (>= 1000) or -1.
int user_rhs_register(const char* source, const char* entry, size_t ny, size_t nx, size_t args_bytes, size_t n_events,
                      char* log_out, size_t log_len);
int user_rhs_info(int rhs_id, int* ny, int* nx, int* nargs_dbl, int* n_events);
cudaError_t user_rhs_launch_solve(int rhs_id, int method, int variant, bool strict, int device, const BatchParams& p, int grid,
                                  int block, cudaStream_t s);
cudaError_t user_rhs_occupancy(int rhs_id, int method, int variant, bool strict, int device, int* ctas, int block);
cudaError_t user_rhs_launch_prep(int rhs_id, int method, bool strict, int device, const BatchParams& p, double* prep, int grid,
                                 int block, cudaStream_t s);
cudaError_t user_rhs_launch_dense(int rhs_id, int method, bool strict, int device, const DenseEvalParams& p, int grid, int block,
                                  cudaStream_t s);
}  // namespace cyrk
