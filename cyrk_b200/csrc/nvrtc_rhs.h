// nvrtc_rhs.h -- user right-hand sides compiled at run time (defined in nvrtc_rhs.cpp).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "cyrk_defs.h"
#include "cta_core_params.h"

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
cudaError_t user_rhs_launch_step_dense(int rhs_id, int method, bool strict, int device, const StepDenseParams& p, int grid, int block,
                                       cudaStream_t s);
// Large systems (registered with n_y == 0): the source defines
//     __device__ double <entry>(unsigned i, double t, const double* y, unsigned n, const char* args);   -> dy_i
// and, when registered with n_events > 0,
//     __device__ double <entry>_event(int id, double t, const double* y, unsigned n, const char* args); -> g_id(t, y)
// cta2_C > 0: the register-resident kernel (cta2_core.cuh) with that many components per thread; 0: cta_core.cuh
cudaError_t user_rhs_launch_cta(int rhs_id, int method, int cta2_C, bool strict, int device, const CtaParams& p, int grid, int block,
                                size_t smem, cudaStream_t s);
}  // namespace cyrk
