/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * One flat C interface implemented twice:
 *   orc_*  oracle/cyrk_oracle.c   -- CPU restatement ("port") of the reference
 *                                    algorithm, plain C.
 *   ref_*  oracle/ref_driver.cpp  -- thin glue over the UNMODIFIED reference
 *                                    sources compiled where they lie under
 *                                    /root/reference (oracle/_ref/libcyrk_ref.so).
 * The shape follows the reference's own flat API, CyRK/nb/numba_c_api_.cpp:10-222
 * (solve -> opaque pointer, getters, free).
 */
#ifndef CYRK_ORACLE_API_H
#define CYRK_ORACLE_API_H

#include <stddef.h>
#include "rhs_lib.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_problem {
    orc_diffeq_fn   diffeq;
    orc_pre_eval_fn pre_eval;
    int             method;        /* ODEMethod ints, cysolver.hpp:34-41: 3 RK23, 4 RK45, 5 DOP853 */
    double          t_start, t_end;
    const double*   y0;
    size_t          num_y;
    size_t          num_extra;
    const char*     args;
    size_t          args_len;
    size_t          max_num_steps;
    size_t          max_ram_MB;
    size_t          expected_size;
    int             dense_output;
    const double*   t_eval;
    size_t          n_t_eval;
    const orc_event_fn* event_fns;
    const int*      event_dirs;
    const size_t*   event_max;     /* (size_t)-1 == unlimited, c_events.hpp:33 */
    size_t          n_events;
    const double*   rtols;
    size_t          n_rtols;
    const double*   atols;
    size_t          n_atols;
    double          max_step;      /* INF = none */
    double          first_step;    /* 0 = automatic */
    int             fix_quirk11;   /* port only: 1 = do NOT emulate the reference's
                                      termination+extras overrun */
} orc_problem;

#define ORC_DECLARE(P)                                                              \
    void*  P##_solve(const orc_problem* prob);                                      \
    void   P##_free(void* h);                                                       \
    int    P##_status(void* h);                                                     \
    int    P##_success(void* h);                                                    \
    size_t P##_size(void* h);                                                       \
    size_t P##_steps_taken(void* h);                                                \
    size_t P##_num_interpolates(void* h);                                           \
    size_t P##_num_dy(void* h);                                                     \
    size_t P##_nfev(void* h);                                                       \
    const double* P##_t_ptr(void* h);                                               \
    const double* P##_y_ptr(void* h);          /* row-major [size][num_dy] */       \
    int    P##_event_terminated(void* h);                                           \
    size_t P##_event_terminate_index(void* h);                                      \
    size_t P##_event_count(void* h, size_t e);                                      \
    const double* P##_event_t_ptr(void* h, size_t e);                               \
    const double* P##_event_y_ptr(void* h, size_t e);  /* [count][num_dy] */        \
    int    P##_call(void* h, const double* t, size_t n, double* out /*[n][num_dy]*/);\
    /* Batch of independent solves (the prange pattern, prange_test.pyx:59-97):     \
       one solver object per thread, reused.  Per-trajectory y0/args/t_span;        \
       t_spans may be NULL (shared).  teval_out may be NULL, else [B][n_t_eval][num_dy] \
       (rows beyond `size` untouched).  Returns wall seconds of the parallel region. */ \
    double P##_solve_batch(const orc_problem* base, size_t B, const double* y0s,    \
                           const char* argss, size_t args_stride,                   \
                           const double* t_spans, int n_threads,                    \
                           int* status, size_t* size, size_t* steps, size_t* nfev,  \
                           double* t_last, double* y_last, double* teval_out);

ORC_DECLARE(orc)
ORC_DECLARE(ref)

/* root finder exposed for the brentq known-answer test
 * (Tests/F_Optimize_Tests/test_brentq.py; CyRK/optimize/optimize_test.pyx) */
double orc_brentq_poly3(const double a[4], double xa, double xb, double xtol, double rtol,
                        size_t mitr, int* err, size_t* funcalls, size_t* iterations);
double ref_brentq_poly3(const double a[4], double xa, double xb, double xtol, double rtol,
                        size_t mitr, int* err, size_t* funcalls, size_t* iterations);

#ifdef __cplusplus
}
#endif
#endif
