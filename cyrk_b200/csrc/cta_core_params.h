// cta_core_params.h -- launch parameters of the CTA-per-trajectory kernels (host + device).
#pragma once

namespace cyrk {

struct CtaParams {
    unsigned long long n_traj;
    unsigned n;                  // n_y
    unsigned n_pad;              // n rounded up to a multiple of 2 (128-bit accesses)
    const double* y0;            // [n_traj][n]
    const char*   args;          // per-trajectory blob of args_stride bytes, or shared (stride 0)
    unsigned      args_stride;
    const double* t_span;
    unsigned      t_span_stride;
    const double* rtol;          // [n]  (floored at 100 eps on the host)
    const double* atol;          // [n]
    double max_step, first_step;
    unsigned max_num_steps;
    int    user_max_steps;
    const double* t_eval;        // ascending (reversed copy for backward runs)
    unsigned n_t_eval;
    int*      status;
    unsigned* size;
    unsigned* steps_taken;
    unsigned* n_rejected;
    unsigned* nfev;
    double*   t_last;
    double*   y_last;            // [n_traj][n]
    double*   t_rows;            // [n_traj][row_cap]
    double*   y_rows;            // [n_traj][row_cap][n]
    unsigned  row_cap;
    double*   scratch;           // [gridDim.x][n_vec_global + 1][n_pad]: spilled vectors + the problem's aux table
    unsigned  n_smem_vec;        // how many of the working vectors live in shared memory
    unsigned long long* queue;
};

}  // namespace cyrk
