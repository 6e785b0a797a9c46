// problems_builtin.cuh -- built-in right-hand sides and event functions as inlined device functors.
//
// These are the reference's benchmark / test problems (CyRK/cy/cysolver_test.pyx:22-244,
// Performance/*.py, Documentation/Advanced_CySolver.md:127-152) with the reference's callback
// contract (CyRK/cy/c_common.hpp:218-220, c_events.hpp:9):
//     rhs(dy[NY+NEXTRA], t, y[NY], args)        event(t, y_and_extras[NY+NEXTRA], args) -> double
// `args` is the per-trajectory opaque blob (copied into registers by the solver).
// Expression order follows the Cython source so the -fmad=false build rounds identically.
#pragma once
#include "cyrk_defs.h"

namespace cyrk {

// ids shared with include/cyrk_b200.h (CYRK_RHS_*) and oracle/rhs_lib.h
enum BuiltinRhs : int {
    RHS_BASELINE = 0, RHS_ACCURACY = 1, RHS_EXTRAOUTPUT = 2, RHS_LORENZ = 3, RHS_LORENZ_EXTRA = 4,
    RHS_LOTKAVOLTERRA = 5, RHS_PENDULUM = 6, RHS_ARBITRARY_ARGS = 7, RHS_PENDULUM_PREEVAL = 8,
    RHS_LARGEY = 9, RHS_LARGEY_SIMPLE = 10, RHS_PENDULUM_DRAG = 11, RHS_DENSE_KAT = 12,
    RHS_BUILTIN_COUNT = 13
};

// flat event-function table (same ids for every problem; oracle/rhs_lib.h ORC_EV_*)
enum BuiltinEvent : int {
    EV_Y0_GT_50 = 0, EV_Y1_LT_50 = 1, EV_T_GT_5 = 2, EV_LORENZ_T_GE_5 = 3, EV_LORENZ_Y_RANGE = 4,
    EV_LORENZ_ARGS = 5, EV_Y0 = 6, EV_Y1 = 7, EV_T_MINUS_ARG4 = 8, EV_ABS_Y0_MINUS_ARG4 = 9,
    EV_BUILTIN_COUNT = 10
};

// Events that only touch y[0], y[1], t and args: valid for every problem with NY >= 2.
// (cysolver_test.pyx:309-350, events_test.pyx:87-121, SURVEY.md 8c C3 KAT)
template <int NY, int NARGS_DBL>
__device__ __forceinline__ double builtin_event(int id, double t, const double* y, const char* args)
{
    const double* a = reinterpret_cast<const double*>(args);
    switch (id) {
    case EV_Y0_GT_50: return (y[0] > 50.0) ? 0.0 : 1.0;
    case EV_Y1_LT_50: return (y[NY > 1 ? 1 : 0] < 50.0) ? 0.0 : 1.0;
    case EV_T_GT_5:   return (t > 5.0) ? 0.0 : 1.0;
    case EV_LORENZ_T_GE_5: return (t >= 5.0) ? 0.0 : 1.0;
    case EV_LORENZ_Y_RANGE:
        if (y[0] < -10.0) return 0.0;
        else if (y[NY > 2 ? 2 : 0] > 30.0) return 0.0;
        return 1.0;
    case EV_LORENZ_ARGS:
        if constexpr (NARGS_DBL >= 3) return (a[0] == 10.0 && a[1] == 28.0 && a[2] == 8.0 / 3.0) ? 0.0 : 1.0;
        else return 1.0;
    case EV_Y0: return y[0];
    case EV_Y1: return y[NY > 1 ? 1 : 0];
    case EV_T_MINUS_ARG4:
        if constexpr (NARGS_DBL >= 5) return t - a[4];
        else return 1.0;
    case EV_ABS_Y0_MINUS_ARG4:
        if constexpr (NARGS_DBL >= 5) return fabs(y[0]) - a[4];
        else return 1.0;
    default: return 1.0;
    }
}

// Problems may append NDERIVED doubles to their per-trajectory argument registers: derive(a) runs once when a
// trajectory is loaded and fills a[NARGS_DBL ..] from a[0 .. NARGS_DBL) -- coefficients the reference recomputes in
// every right-hand-side call (two FP64 divisions per call for the pendulum: ~70 instructions per RK23 attempt).  The
// operations are the reference's, only hoisted, so the strict build stays bit-identical.
#define CYRK_PROBLEM_COMMON(ny, nx, nargs)                                                        \
    CYRK_PROBLEM_HEAD(ny, nx, nargs)                                                              \
    static constexpr int NDERIVED = 0;                                                            \
    __device__ __forceinline__ static void derive(double*) {}

#define CYRK_PROBLEM_HEAD(ny, nx, nargs)                                                          \
    static constexpr int NY = ny;                                                                 \
    static constexpr int NEXTRA = nx;                                                             \
    static constexpr int NARGS_DBL = nargs;                                                       \
    __device__ __forceinline__ static double event(int id, double t, const double* y, const char* args) \
    {                                                                                             \
        return builtin_event<ny, nargs>(id, t, y, args);                                          \
    }

struct ProbBaseline {           // cysolver_test.pyx:22-29
    CYRK_PROBLEM_COMMON(2, 0, 0)
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char*)
    {
        const double y0 = y[0], y1 = y[1];
        dy[0] = (1. - 0.01 * y1) * y0;
        dy[1] = (0.02 * y0 - 1.) * y1;
    }
};

struct ProbAccuracy {           // cysolver_test.pyx:32-39
    CYRK_PROBLEM_COMMON(2, 0, 0)
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char*)
    {
        const double y0 = y[0], y1 = y[1];
        dy[0] = CYRK_FP_NS::psin(t) - y1;
        dy[1] = CYRK_FP_NS::pcos(t) + y0;
    }
};

struct ProbExtraOutput {        // cysolver_test.pyx:42-57
    CYRK_PROBLEM_COMMON(2, 2, 0)
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char*)
    {
        const double y0 = y[0], y1 = y[1];
        const double e0 = (1. - 0.01 * y1);
        const double e1 = (0.02 * y0 - 1.);
        dy[0] = e0 * y0;
        dy[1] = e1 * y1;
        dy[2] = e0;
        dy[3] = e1;
    }
};

struct ProbLorenz {             // cysolver_test.pyx:60-75, Performance/lorenz.py:8-24
    CYRK_PROBLEM_COMMON(3, 0, 3)
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char* args)
    {
        const double* p = reinterpret_cast<const double*>(args);
        const double a = p[0], b = p[1], c = p[2];
        const double y0 = y[0], y1 = y[1], y2 = y[2];
        dy[0] = a * (y1 - y0);
        dy[1] = y0 * (b - y2) - y1;
        dy[2] = y0 * y1 - c * y2;
    }
};

struct ProbLorenzExtra {        // cysolver_test.pyx:78-101
    CYRK_PROBLEM_COMMON(3, 3, 3)
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char* args)
    {
        const double* p = reinterpret_cast<const double*>(args);
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
};

struct ProbLotkaVolterra {      // cysolver_test.pyx:104-118, Performance/lotkavolterra.py:7-21
    CYRK_PROBLEM_COMMON(2, 0, 4)
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char* args)
    {
        const double* p = reinterpret_cast<const double*>(args);
        const double a = p[0], b = p[1], c = p[2], d = p[3];
        const double y0 = y[0], y1 = y[1];
#if CYRK_FP_STRICT
        dy[0] = a * y0 - b * y0 * y1;
        dy[1] = -c * y1 + d * y0 * y1;
#else
        // production build: the same polynomial factored (2 DFMA + 2 DMUL instead of 2 DFMA + 4 DMUL); the
        // re-association changes the result by <= 2 ulp, the same order as the FMA contraction itself
        dy[0] = y0 * fma(-b, y1, a);
        dy[1] = y1 * fma(d, y0, -c);
#endif
    }
};

struct ProbPendulum {           // cysolver_test.pyx:122-141, Performance/pendulum.py:8-25
    // args (l, m, g) plus two spare doubles so the per-trajectory stop-time events
    // (EV_T_MINUS_ARG4 / EV_ABS_Y0_MINUS_ARG4) can be used with the undamped pendulum too.
    CYRK_PROBLEM_HEAD(2, 0, 5)
    static constexpr int NDERIVED = 2;
    __device__ __forceinline__ static void derive(double* a)
    {
        const double l = a[0], m = a[1], g = a[2];
        a[5] = (-3. * g / (2. * l));
        a[6] = (3. / (m * (l * l)));
    }
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char* args)
    {
        const double* p = reinterpret_cast<const double*>(args);
        const double coeff_1 = p[5], coeff_2 = p[6];
        const double y0 = y[0], y1 = y[1];
        double st, sy;
        CYRK_FP_NS::psin2(t, y0, st, sy);
        const double torque = CYRK_LIT_TENTH * st;
        dy[0] = y1;
        dy[1] = coeff_1 * sy + coeff_2 * torque;
    }
};

struct ProbArbitraryArgs {      // cysolver_test.pyx:144-178: heterogeneous struct behind the char* blob
    struct Args { double l; bool cause_fail; bool checker; double m; double g; };
    CYRK_PROBLEM_COMMON(2, 0, 4)
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char* args)
    {
        const Args* p = reinterpret_cast<const Args*>(args);
        const double l = p->l, m = p->m, g = p->g;
        const double coeff_1 = (-3. * g / (2. * l));
        const double coeff_2 = (3. / (m * (l * l)));
        const double y0 = y[0], y1 = y[1];
        const double torque = 0.1 * CYRK_FP_NS::psin(t);
        if (p->cause_fail) {
            dy[0] = pow(100 * t, 1000.0);
            dy[1] = pow(-100 * t, 1000.0);
        } else {
            dy[0] = y1;
            dy[1] = coeff_1 * CYRK_FP_NS::psin(y0) + coeff_2 * torque;
        }
    }
};

struct ProbPendulumPreEval {    // cysolver_test.pyx:182-222: the pre-eval hook is just another inlined functor
    CYRK_PROBLEM_COMMON(2, 0, 3)
    __device__ __forceinline__ static void pre_eval(double* out, double t, const double* y, const char* args)
    {
        const double* p = reinterpret_cast<const double*>(args);
        const double l = p[0], m = p[1], g = p[2];
        out[0] = 0.1 * CYRK_FP_NS::psin(t);
        out[1] = (-3. * g / (2. * l));
        out[2] = (3. / (m * (l * l)));
    }
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char* args)
    {
        double s[3];
        pre_eval(s, t, y, args);
        const double y0 = y[0], y1 = y[1];
        dy[0] = y1;
        dy[1] = s[1] * CYRK_FP_NS::psin(y0) + s[2] * s[0];
    }
};

struct ProbPendulumDrag {       // Documentation/Advanced_CySolver.md:127-152; args (l, m, g, c, stop)
    CYRK_PROBLEM_HEAD(2, 1, 5)
    static constexpr int NDERIVED = 2;
    __device__ __forceinline__ static void derive(double* a)
    {
        const double l = a[0], m = a[1], g = a[2];
        a[5] = (-3. * g / (2. * l));
        a[6] = (3. / (m * (l * l)));
    }
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char* args)
    {
        const double* p = reinterpret_cast<const double*>(args);
        const double c = p[3];
        const double coeff_1 = p[5], coeff_2 = p[6];
        const double y0 = y[0], y1 = y[1];
        double st, sy;
        CYRK_FP_NS::psin2(t, y0, st, sy);
        const double torque = CYRK_LIT_TENTH * st;
        dy[0] = y1;
        double d1 = coeff_1 * sy + coeff_2 * torque;
        d1 -= c * y1;
        dy[1] = d1;
        dy[2] = torque;
    }
};

struct ProbDenseKat {           // Tests/D_PySolver_Tests/test_b_pysolve_dense.py:4-11
    CYRK_PROBLEM_COMMON(2, 2, 0)
    __device__ __forceinline__ static void rhs(double* dy, double t, const double* y, const char*)
    {
        dy[0] = 3.1 * t - y[1];
        dy[1] = y[0] * (0.3 * t * y[1]);
        dy[2] = 0.25;
        dy[3] = t / 2.;
    }
};

}  // namespace cyrk
