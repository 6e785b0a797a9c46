/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Glue over the UNMODIFIED reference.  This file contains no solver logic: it
 * includes the reference's own headers from /root/reference/CyRK (compiled
 * where they lie by oracle/Makefile, output only in oracle/_ref/) and calls
 *   baseline_cysolve_ivp_noreturn   CyRK/cy/cysolve.cpp:6-59
 *   CySolverResult::call_vectorize  CyRK/cy/cysolution.cpp:598-616
 *   c_brentq                        CyRK/optimize/c_brentq.cpp:41-204
 * exposing them through the flat interface of oracle_api.h (ref_* functions).
 * The batch entry follows CyRK/cy/prange_test.pyx:59-97 with one
 * CySolverResult per OpenMP thread, reused (Documentation/CySolverResult_Reuses.md).
 */
#include <vector>
#include <memory>
#include <cstring>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

/* The reference is a unity build: its Cython .pxd files pull the .cpp files in
 * with `cdef extern from "x.cpp"` (CyRK/cy/cysolver_api.pxd:16-29,76,107,237,310;
 * CyRK/cy/common.pxd:9) after Python.h.  The separate translation units do not
 * compile on their own (incomplete types, missing <cstdint>), so this driver
 * reproduces the same single-TU include order.  Nothing is copied: the files
 * are included from where they lie (-I$(REF)/CyRK/cy etc. in oracle/Makefile). */
#include <Python.h>
#include <cstdint>
#include <cstddef>
#include <cmath>
#include "c_common.cpp"
#include "c_brentq.cpp"
#include "c_events.cpp"
#include "cy_array.cpp"
#include "cysolution.cpp"
#include "dense.cpp"
#include "cysolver.cpp"
#include "rk.cpp"
#include "cysolve.cpp"

#include "oracle_api.h"

namespace {

struct RefHandle {
    std::unique_ptr<CySolverResult> res;
    size_t nfev = 0;
};

struct Inputs {
    std::vector<double> y0, t_eval, rtols, atols;
    std::vector<char> args;
    std::vector<Event> events;
};

void fill_inputs(const orc_problem* p, Inputs& in)
{
    in.y0.assign(p->y0, p->y0 + p->num_y);
    in.args.assign(p->args, p->args + p->args_len);
    in.t_eval.assign(p->t_eval, p->t_eval + p->n_t_eval);
    in.rtols.assign(p->rtols, p->rtols + p->n_rtols);
    in.atols.assign(p->atols, p->atols + p->n_atols);
    in.events.clear();
    for (size_t e = 0; e < p->n_events; e++)
        in.events.emplace_back((EventFunc)p->event_fns[e], p->event_max[e], p->event_dirs[e]);
}

void run(CySolverResult* r, const orc_problem* p, Inputs& in)
{
    baseline_cysolve_ivp_noreturn(
        r, (DiffeqFuncType)p->diffeq, p->t_start, p->t_end, in.y0,
        p->expected_size, p->num_extra, in.args, p->max_num_steps, p->max_ram_MB,
        p->dense_output != 0, in.t_eval, (PreEvalFunc)p->pre_eval, in.events,
        in.rtols, in.atols, p->max_step, p->first_step, /*force_retain_solver=*/true);
}

}  // namespace

extern "C" {

void* ref_solve(const orc_problem* p)
{
    RefHandle* h = new RefHandle();
    Inputs in;
    fill_inputs(p, in);
    h->res = std::make_unique<CySolverResult>(static_cast<ODEMethod>(p->method));
    orc_nfev_reset();
    run(h->res.get(), p, in);
    h->nfev = orc_nfev_get();
    return h;
}

void ref_free(void* h) { delete static_cast<RefHandle*>(h); }

#define R(h) (static_cast<RefHandle*>(h)->res.get())
int    ref_status(void* h) { return static_cast<int>(R(h)->status); }
int    ref_success(void* h) { return R(h)->success ? 1 : 0; }
size_t ref_size(void* h) { return R(h)->size; }
size_t ref_steps_taken(void* h) { return R(h)->steps_taken; }
size_t ref_num_interpolates(void* h) { return R(h)->num_interpolates; }
size_t ref_num_dy(void* h) { return R(h)->num_dy; }
size_t ref_nfev(void* h) { return static_cast<RefHandle*>(h)->nfev; }
const double* ref_t_ptr(void* h) { return R(h)->time_domain_vec.data(); }
const double* ref_y_ptr(void* h) { return R(h)->solution.data(); }
int    ref_event_terminated(void* h) { return R(h)->event_terminated ? 1 : 0; }
size_t ref_event_terminate_index(void* h) { return R(h)->event_terminate_index; }
size_t ref_event_count(void* h, size_t e) { return e < R(h)->event_times.size() ? R(h)->event_times[e].size() : 0; }
const double* ref_event_t_ptr(void* h, size_t e) { return R(h)->event_times[e].data(); }
const double* ref_event_y_ptr(void* h, size_t e) { return R(h)->event_states[e].data(); }
int ref_call(void* h, const double* t, size_t n, double* out)
{
    return static_cast<int>(R(h)->call_vectorize(t, n, out));
}

double ref_solve_batch(const orc_problem* base, size_t B, const double* y0s,
                       const char* argss, size_t args_stride,
                       const double* t_spans, int n_threads,
                       int* status, size_t* size, size_t* steps, size_t* nfev,
                       double* t_last, double* y_last, double* teval_out)
{
    const size_t ny = base->num_y, ndy = base->num_y + base->num_extra;
    if (n_threads < 1) n_threads = 1;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(n_threads)
    {
        CySolverResult res(static_cast<ODEMethod>(base->method));
        orc_problem p = *base;
        Inputs in;
        fill_inputs(&p, in);
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)B; i++) {
            std::memcpy(in.y0.data(), y0s + (size_t)i * ny, sizeof(double) * ny);
            if (args_stride) std::memcpy(in.args.data(), argss + (size_t)i * args_stride, base->args_len);
            if (t_spans) { p.t_start = t_spans[2 * i]; p.t_end = t_spans[2 * i + 1]; }
            /* the reference mutates rtols in place (rk.cpp:600-608) and events carry
               counters; both are re-copied from the caller's vectors by update_properties. */
            orc_nfev_reset();
            run(&res, &p, in);
            if (status) status[i] = static_cast<int>(res.status);
            if (size) size[i] = res.size;
            if (steps) steps[i] = res.steps_taken;
            if (nfev) nfev[i] = orc_nfev_get();
            if (res.size > 0) {
                if (t_last) t_last[i] = res.time_domain_vec[res.size - 1];
                if (y_last) std::memcpy(y_last + (size_t)i * ndy, &res.solution[(res.size - 1) * ndy], sizeof(double) * ndy);
                if (teval_out && base->n_t_eval) {
                    size_t rows = res.size < base->n_t_eval ? res.size : base->n_t_eval;
                    std::memcpy(teval_out + (size_t)i * base->n_t_eval * ndy, res.solution.data(), sizeof(double) * rows * ndy);
                }
            }
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

/* cubic used by Tests/F_Optimize_Tests/test_brentq.py via optimize_test.pyx:17-22 */
static double poly3_event(Event*, double x, double*, char* a)
{
    const double* c = reinterpret_cast<const double*>(a);
    return ((c[3] * x + c[2]) * x + c[1]) * x + c[0];
}

double ref_brentq_poly3(const double a[4], double xa, double xb, double xtol, double rtol,
                        size_t mitr, int* err, size_t* funcalls, size_t* iterations)
{
    std::vector<char> args(sizeof(double) * 4);
    std::memcpy(args.data(), a, sizeof(double) * 4);
    OptimizeInfo info;
    double r = c_brentq(poly3_event, xa, xb, xtol, rtol, mitr, args, &info, nullptr, nullptr);
    if (err) *err = static_cast<int>(info.error_num);
    if (funcalls) *funcalls = info.funcalls;
    if (iterations) *iterations = info.iterations;
    return r;
}

}  // extern "C"
