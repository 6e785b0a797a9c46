// cta_solver.h -- host entry of the CTA-per-trajectory path (defined in inst_cta.cu, once per FP mode).
#pragma once
#include <cuda_runtime.h>
#include "cta_core_params.h"

namespace cyrk {
// rhs_id: RHS_LARGEY / RHS_LARGEY_SIMPLE.  Returns cudaErrorInvalidValue for an unknown combination.
cudaError_t cta_launch_fast(int rhs_id, int method, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s);
cudaError_t cta_launch_strict(int rhs_id, int method, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s);
// register-resident variant (cta2_core.cuh): C components per thread, n_y <= C * block
cudaError_t cta2_launch_fast(int rhs_id, int method, int C, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s);
cudaError_t cta2_launch_strict(int rhs_id, int method, int C, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s);
cudaError_t cta_dense_eval_fast(int method, const CtaDenseEvalParams& p, int grid, int block, cudaStream_t s);
cudaError_t cta_dense_eval_strict(int method, const CtaDenseEvalParams& p, int grid, int block, cudaStream_t s);
int cta_num_vectors(int method);      // working vectors of one trajectory: y_now, y_old, K_0..K_{s}(+3 for DOP853)
}  // namespace cyrk
