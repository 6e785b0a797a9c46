/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See rhs_lib.h.
 *
 * Compile WITHOUT fused multiply-add (gcc -O2 -ffp-contract=off, no -march):
 * the reference's Linux wheels contain no FMA instructions, so
 * every expression below is evaluated as separately rounded IEEE-754 binary64
 * operations, left to right, exactly as Cython emits them.
 */
#include "rhs_lib.h"

#include <math.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static _Thread_local size_t g_nfev = 0;
void   orc_nfev_reset(void) { g_nfev = 0; }
size_t orc_nfev_get(void)   { return g_nfev; }

/* --- cysolver_test.pyx:22-29 ------------------------------------------- */
static void rhs_baseline(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)t; (void)args; (void)pre; g_nfev++;
    const double y0 = y[0], y1 = y[1];
    dy[0] = (1. - 0.01 * y1) * y0;
    dy[1] = (0.02 * y0 - 1.) * y1;
}

/* --- cysolver_test.pyx:32-39 ------------------------------------------- */
static void rhs_accuracy(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)args; (void)pre; g_nfev++;
    const double y0 = y[0], y1 = y[1];
    dy[0] = sin(t) - y1;
    dy[1] = cos(t) + y0;
}

/* --- cysolver_test.pyx:42-57 ------------------------------------------- */
static void rhs_extraoutput(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)t; (void)args; (void)pre; g_nfev++;
    const double y0 = y[0], y1 = y[1];
    const double e0 = (1. - 0.01 * y1);
    const double e1 = (0.02 * y0 - 1.);
    dy[0] = e0 * y0;
    dy[1] = e1 * y1;
    dy[2] = e0;
    dy[3] = e1;
}

/* --- cysolver_test.pyx:60-75 ------------------------------------------- */
static void rhs_lorenz(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)t; (void)pre; g_nfev++;
    const double* p = (const double*)args;
    const double a = p[0], b = p[1], c = p[2];
    const double y0 = y[0], y1 = y[1], y2 = y[2];
    dy[0] = a * (y1 - y0);
    dy[1] = y0 * (b - y2) - y1;
    dy[2] = y0 * y1 - c * y2;
}

/* --- cysolver_test.pyx:78-101 ------------------------------------------ */
static void rhs_lorenz_extra(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)t; (void)pre; g_nfev++;
    const double* p = (const double*)args;
    const double a = p[0], b = p[1], c = p[2];
    const double y0 = y[0], y1 = y[1], y2 = y[2];
    const double e1 = a;
    const double e2 = (b - y2);
    const double e3 = c * y2;
    dy[0] = e1 * (y1 - y0);
    dy[1] = y0 * e2 - y1;
    dy[2] = y0 * y1 - e3;
    dy[3] = e1;
    dy[4] = e2;
    dy[5] = e3;
}

/* --- cysolver_test.pyx:104-118 ----------------------------------------- */
static void rhs_lotkavolterra(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)t; (void)pre; g_nfev++;
    const double* p = (const double*)args;
    const double a = p[0], b = p[1], c = p[2], d = p[3];
    const double y0 = y[0], y1 = y[1];
    dy[0] = a * y0 - b * y0 * y1;
    dy[1] = -c * y1 + d * y0 * y1;
}

/* --- cysolver_test.pyx:122-141 ----------------------------------------- */
static void rhs_pendulum(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)pre; g_nfev++;
    const double* p = (const double*)args;
    const double l = p[0], m = p[1], g = p[2];
    const double coeff_1 = (-3. * g / (2. * l));
    const double coeff_2 = (3. / (m * pow(l, 2.0)));
    const double y0 = y[0], y1 = y[1];
    const double torque = 0.1 * sin(t);
    dy[0] = y1;
    dy[1] = coeff_1 * sin(y0) + coeff_2 * torque;
}

/* --- cysolver_test.pyx:144-178 ----------------------------------------- */
typedef struct {
    double l;
    _Bool cause_fail;
    _Bool checker;
    double m;
    double g;
} orc_arbitrary_args;

static void rhs_arbitrary_args(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)pre; g_nfev++;
    const orc_arbitrary_args* p = (const orc_arbitrary_args*)args;
    const double l = p->l, m = p->m, g = p->g;
    const double coeff_1 = (-3. * g / (2. * l));
    const double coeff_2 = (3. / (m * pow(l, 2.0)));
    const double y0 = y[0], y1 = y[1];
    const double torque = 0.1 * sin(t);
    if (p->cause_fail) {
        dy[0] = pow(100 * t, 1000.0);
        dy[1] = pow(-100 * t, 1000.0);
    } else {
        dy[0] = y1;
        dy[1] = coeff_1 * sin(y0) + coeff_2 * torque;
    }
}

/* --- cysolver_test.pyx:182-222 ----------------------------------------- */
static void pre_eval_pendulum(char* out, double t, double* y, char* args)
{
    (void)y;
    const double* p = (const double*)args;
    const double l = p[0], m = p[1], g = p[2];
    double* o = (double*)out;
    o[0] = 0.1 * sin(t);
    o[1] = (-3. * g / (2. * l));
    o[2] = (3. / (m * pow(l, 2.0)));
}

static void rhs_pendulum_preeval(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    g_nfev++;
    double s[3];
    pre((char*)s, t, y, args);
    const double y0 = y[0], y1 = y[1];
    dy[0] = y1;
    dy[1] = s[1] * sin(y0) + s[2] * s[0];
}

/* --- cysolver_test.pyx:224-237 -----------------------------------------
 * The reference hard-codes num_y = 10 000.  Here n = (size_t)args_dbl[1] so
 * the same function serves BASELINE.json's n_y = 4096 case; with n = 10 000
 * it is the reference function. */
static void rhs_largey(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)pre; g_nfev++;
    const double* p = (const double*)args;
    double decay_rate = p[0];
    const size_t n = (size_t)p[1];
    for (size_t i = 0; i < n; i++) {
        decay_rate *= 0.9999;
        if (i < (n - 1))
            dy[i] = decay_rate * y[i] * sin(2 * M_PI * t / 5.0 + y[i + 1] / 50.0);
        else
            dy[i] = decay_rate * y[i] * sin(2 * M_PI * t / 5.0);
    }
}

/* --- cysolver_test.pyx:239-244 (memset length quirk not reproduced: the
 * reference's buffers are already zero, so the result is identical). ------ */
static void rhs_largey_simple(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)y; (void)pre; g_nfev++;
    const double* p = (const double*)args;
    const size_t n = (size_t)p[1];
    dy[0] = sin(2.0 * M_PI * t / 10.0);
    if (n > 1) memset(&dy[1], 0, sizeof(double) * (n - 1));
}

/* --- Documentation/Advanced_CySolver.md:127-152 with drag coefficient from
 * args (l, m, g, c[, stop]) and the torque exported as one extra output
 * (SURVEY.md 8c "C3-style KAT"). -------------------------------------------*/
static void rhs_pendulum_drag(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)pre; g_nfev++;
    const double* p = (const double*)args;
    const double l = p[0], m = p[1], g = p[2], c = p[3];
    const double coeff_1 = (-3. * g / (2. * l));
    const double coeff_2 = (3. / (m * pow(l, 2.0)));
    const double y0 = y[0], y1 = y[1];
    const double torque = 0.1 * sin(t);
    dy[0] = y1;
    dy[1] = coeff_1 * sin(y0) + coeff_2 * torque;
    dy[1] -= c * y1;
    dy[2] = torque;
}

/* --- Tests/D_PySolver_Tests/test_b_pysolve_dense.py:4-11 ---------------- */
static void rhs_dense_kat(double* dy, double t, double* y, char* args, orc_pre_eval_fn pre)
{
    (void)args; (void)pre; g_nfev++;
    dy[0] = 3.1 * t - y[1];
    dy[1] = y[0] * (0.3 * t * y[1]);
    dy[2] = 0.25;
    dy[3] = t / 2.;
}

orc_diffeq_fn orc_rhs_get(int id)
{
    switch (id) {
    case ORC_RHS_BASELINE:         return rhs_baseline;
    case ORC_RHS_ACCURACY:         return rhs_accuracy;
    case ORC_RHS_EXTRAOUTPUT:      return rhs_extraoutput;
    case ORC_RHS_LORENZ:           return rhs_lorenz;
    case ORC_RHS_LORENZ_EXTRA:     return rhs_lorenz_extra;
    case ORC_RHS_LOTKAVOLTERRA:    return rhs_lotkavolterra;
    case ORC_RHS_PENDULUM:         return rhs_pendulum;
    case ORC_RHS_ARBITRARY_ARGS:   return rhs_arbitrary_args;
    case ORC_RHS_PENDULUM_PREEVAL: return rhs_pendulum_preeval;
    case ORC_RHS_LARGEY:           return rhs_largey;
    case ORC_RHS_LARGEY_SIMPLE:    return rhs_largey_simple;
    case ORC_RHS_PENDULUM_DRAG:    return rhs_pendulum_drag;
    case ORC_RHS_DENSE_KAT:        return rhs_dense_kat;
    default: return (orc_diffeq_fn)0;
    }
}

orc_pre_eval_fn orc_rhs_pre_eval(int id)
{
    return id == ORC_RHS_PENDULUM_PREEVAL ? pre_eval_pendulum : (orc_pre_eval_fn)0;
}

int orc_rhs_num_extra(int id)
{
    switch (id) {
    case ORC_RHS_EXTRAOUTPUT:   return 2;
    case ORC_RHS_LORENZ_EXTRA:  return 3;
    case ORC_RHS_PENDULUM_DRAG: return 1;
    case ORC_RHS_DENSE_KAT:     return 2;
    default: return 0;
    }
}

/* --- events ------------------------------------------------------------- */
static double ev_y0_gt_50(double t, double* y, char* a) { (void)t; (void)a; return (y[0] > 50.0) ? 0.0 : 1.0; }
static double ev_y1_lt_50(double t, double* y, char* a) { (void)t; (void)a; return (y[1] < 50.0) ? 0.0 : 1.0; }
static double ev_t_gt_5(double t, double* y, char* a)   { (void)y; (void)a; return (t > 5.0) ? 0.0 : 1.0; }
static double ev_lorenz_t_ge_5(double t, double* y, char* a) { (void)y; (void)a; return (t >= 5.0) ? 0.0 : 1.0; }
static double ev_lorenz_y_range(double t, double* y, char* a)
{
    (void)t; (void)a;
    if (y[0] < -10.0) return 0.0;
    else if (y[2] > 30.0) return 0.0;
    return 1.0;
}
static double ev_lorenz_args(double t, double* y, char* a)
{
    (void)t; (void)y;
    const double* p = (const double*)a;
    return (p[0] == 10.0 && p[1] == 28.0 && p[2] == 8.0 / 3.0) ? 0.0 : 1.0;
}
static double ev_y0(double t, double* y, char* a) { (void)t; (void)a; return y[0]; }
static double ev_y1(double t, double* y, char* a) { (void)t; (void)a; return y[1]; }
static double ev_t_minus_arg(double t, double* y, char* a) { (void)y; return t - ((const double*)a)[4]; }
static double ev_absy0_minus_arg(double t, double* y, char* a) { (void)t; return fabs(y[0]) - ((const double*)a)[4]; }

orc_event_fn orc_event_get(int id)
{
    switch (id) {
    case ORC_EV_Y0_GT_50: return ev_y0_gt_50;
    case ORC_EV_Y1_LT_50: return ev_y1_lt_50;
    case ORC_EV_T_GT_5:   return ev_t_gt_5;
    case ORC_EV_LORENZ_T_GE_5:  return ev_lorenz_t_ge_5;
    case ORC_EV_LORENZ_Y_RANGE: return ev_lorenz_y_range;
    case ORC_EV_LORENZ_ARGS:    return ev_lorenz_args;
    case ORC_EV_Y0: return ev_y0;
    case ORC_EV_Y1: return ev_y1;
    case ORC_EV_T_MINUS_ARG_LAST: return ev_t_minus_arg;
    case ORC_EV_ABS_Y0_MINUS_ARG_LAST: return ev_absy0_minus_arg;
    default: return (orc_event_fn)0;
    }
}
