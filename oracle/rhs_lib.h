/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Plain-C right-hand-side and event callbacks in the reference's callback ABI
 * (CyRK/cy/c_common.hpp:218-220, CyRK/cy/c_events.hpp:9).  They restate the
 * built-in problems of CyRK/cy/cysolver_test.pyx:22-244 (and the events of
 * :309-364 and CyRK/cy/events_test.pyx:87-121) so that BOTH the CPU oracle
 * (oracle/cyrk_oracle.c) and the compiled reference (oracle/_ref) can be
 * driven with bit-identical callbacks.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 */
#ifndef CYRK_ORACLE_RHS_LIB_H
#define CYRK_ORACLE_RHS_LIB_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*orc_pre_eval_fn)(char*, double, double*, char*);
typedef void (*orc_diffeq_fn)(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre);
typedef double (*orc_event_fn)(double t, double* y, char* args);

/* Problem ids -- the same numbering is used by the CUDA built-ins
 * (include/cyrk_b200.h, CYRK_RHS_*).  0-10 follow cytester's diffeq_number
 * (cysolver_test.pyx:419-445). */
enum {
    ORC_RHS_BASELINE = 0,        /* cysolver_test.pyx:22-29   */
    ORC_RHS_ACCURACY = 1,        /* :32-39                    */
    ORC_RHS_EXTRAOUTPUT = 2,     /* :42-57   (2 extras)       */
    ORC_RHS_LORENZ = 3,          /* :60-75                    */
    ORC_RHS_LORENZ_EXTRA = 4,    /* :78-101  (3 extras)       */
    ORC_RHS_LOTKAVOLTERRA = 5,   /* :104-118                  */
    ORC_RHS_PENDULUM = 6,        /* :122-141                  */
    ORC_RHS_ARBITRARY_ARGS = 7,  /* :144-178 (struct args)    */
    ORC_RHS_PENDULUM_PREEVAL = 8,/* :182-222                  */
    ORC_RHS_LARGEY = 9,          /* :224-237, n from args[1]  */
    ORC_RHS_LARGEY_SIMPLE = 10,  /* :239-244, n from args[1]  */
    ORC_RHS_PENDULUM_DRAG = 11,  /* Documentation/Advanced_CySolver.md:127-152 + torque as extra */
    ORC_RHS_DENSE_KAT = 12,      /* Tests/D_PySolver_Tests/test_b_pysolve_dense.py:4-11 (2 extras) */
    ORC_RHS_COUNT = 13
};

/* Event-function ids (a flat table shared by all problems). */
enum {
    ORC_EV_Y0_GT_50 = 0,     /* cysolver_test.pyx:310-313 / :336-339 */
    ORC_EV_Y1_LT_50 = 1,     /* :315-318 / :341-344 */
    ORC_EV_T_GT_5 = 2,       /* :320-324 / :346-350 */
    ORC_EV_LORENZ_T_GE_5 = 3,    /* events_test.pyx:87-93  */
    ORC_EV_LORENZ_Y_RANGE = 4,   /* events_test.pyx:95-105 */
    ORC_EV_LORENZ_ARGS = 5,      /* events_test.pyx:107-121 */
    ORC_EV_Y0 = 6,           /* g = y[0]                (SURVEY 8c C3 KAT, E0) */
    ORC_EV_Y1 = 7,           /* g = y[1]                (E1) */
    ORC_EV_T_MINUS_ARG_LAST = 8, /* g = t - args_dbl[4] (E2: per-trajectory stop time) */
    ORC_EV_ABS_Y0_MINUS_ARG_LAST = 9, /* g = |y[0]| - args_dbl[4] */
    ORC_EV_COUNT = 10
};

orc_diffeq_fn  orc_rhs_get(int rhs_id);
orc_pre_eval_fn orc_rhs_pre_eval(int rhs_id);
int            orc_rhs_num_extra(int rhs_id);
orc_event_fn   orc_event_get(int event_id);

/* nfev counter: every callback above increments a thread-local counter. */
void   orc_nfev_reset(void);
size_t orc_nfev_get(void);

#ifdef __cplusplus
}
#endif
#endif
