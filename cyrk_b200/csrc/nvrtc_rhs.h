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
// A BUILT-IN problem (struct name in problems_builtin.cuh) as a run-time compiled right-hand side: rarely used feature
// builds of the precompiled kernels (`defs`, e.g. "-DCYRK_TOL_PER_TRAJ=1") are compiled on demand instead of being
// instantiated for every problem x method x variant in the library.  Returns an id usable with the calls below, or -1.
int builtin_jit_id(const char* prob_struct, int ny, int nx, int nargs_dbl);
// `defs`: extra -D options (space separated) that select a feature build of the kernel; part of the cache key.
cudaError_t user_rhs_launch_solve(int rhs_id, int method, int variant, bool strict, int device, const BatchParams& p, int grid,
                                  int block, cudaStream_t s, const char* defs = "");
cudaError_t user_rhs_occupancy(int rhs_id, int method, int variant, bool strict, int device, int* ctas, int block, const char* defs = "");
cudaError_t user_rhs_launch_prep(int rhs_id, int method, bool strict, int device, const BatchParams& p, double* prep, int grid,
                                 int block, cudaStream_t s, const char* defs = "");
// Compiled cubins are also kept on disk ($CYRK_B200_CACHE_DIR, else $XDG_CACHE_HOME/cyrk_b200, else ~/.cache/cyrk_b200;
// CYRK_B200_CACHE=0 disables), keyed by a hash of the full source, the options and the NVRTC version.
void jit_cache_stats(unsigned long long* compiled, unsigned long long* disk_hits);
// Compile (or fetch from the disk cache) without loading the module: works on a box without a GPU.  kind: "solve" / "prep".
int jit_compile_only(int rhs_id, const char* kind, int method, int variant, bool strict, const char* defs);
cudaError_t user_rhs_launch_dense(int rhs_id, int method, bool strict, int device, const DenseEvalParams& p, int grid, int block,
                                  cudaStream_t s);
cudaError_t user_rhs_launch_event_roots(int rhs_id, int method, bool strict, int device, const BatchParams& p, int grid, int block,
                                        cudaStream_t s, const char* defs = "");
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
