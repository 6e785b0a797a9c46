// cta_core.cuh -- adaptive explicit Runge-Kutta solver for LARGE systems: one CTA per trajectory.
//
// When n_y is in the thousands (BASELINE.json config 4: "largey", n_y = 4096) one thread cannot hold the
// stage matrix, and one trajectory already offers thousands-wide data parallelism.  So:
//   * a CTA owns one trajectory at a time (persistent CTAs, two 512-thread CTAs per SM, pull trajectories from a
//     global counter); its threads stride over the components;
//   * the first working vectors (y_now, y_old, K_0, ...) live in SHARED MEMORY -- as many as leave room for two
//     CTAs per SM (3 vectors of 4096 doubles each) -- the rest in a per-CTA slab of global memory that stays
//     L2-resident (296 CTAs x a few 32 KB vectors << 126 MB).  Measured: placement changes the time by < 2 %, the
//     second CTA per SM (barrier waits overlap arithmetic) by 10 %;
//   * every step-size decision is made from CTA-wide reductions (warp shuffles, then one shared-memory hop),
//     so all threads of the CTA take the same branch: no divergence at all;
//   * accepted rows are streamed to HBM with 128-bit stores, one fully coalesced row of n_y doubles per step.
//
// Reference behaviour restated: identical to rk_core.cuh (same file:line map); the scalar step controller
// uses the reference formulas verbatim in both FP builds -- only the per-component loops are contracted to
// FMAs in the production build.  The error norm is summed by a parallel tree instead of the reference's
// sequential loop (rk.cpp:1047-1072), which changes it by O(eps*sqrt(n_y)) -- documented in DESIGN.md.
//
// Not available on this path yet (the host shim rejects them): extra outputs, events, stored dense output.
#pragma once
#include "cyrk_defs.h"
#include "cta_core_params.h"
#include "tableaux.cuh"

#ifndef CYRK_FP_NS
#define CYRK_FP_NS fast
#endif

namespace cyrk {

namespace CYRK_FP_NS {

#define CYRK_DI __device__ __forceinline__

// Sum over the CTA; every thread receives the same value.  `red` holds 33 doubles of shared memory.
CYRK_DI double cta_sum(double v, double* red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = (blockDim.x + 31) >> 5;
    __syncthreads();                       // protect `red` from the previous reduction's readers
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double w = lane < nwarp ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0) red[32] = w;
    }
    __syncthreads();
    return red[32];
}

template <class Prob, int METHOD>
struct CtaSolver {
    using Tab = typename TabFor<METHOD>::type;
    using Ld = TabLd<Tab>;
    using LdDop = TabLd<TabDOP853>;
    static constexpr int NS = Tab::NS;
    static constexpr int NK = Tab::NK;
    static constexpr int NKD = Tab::NK_DENSE;         // with the interpolant's extra stages
    static constexpr int V_YOLD = 1, V_YNOW = 0, V_K0 = 2;
    static constexpr int NVEC = 2 + NKD;              // y_now, y_old, K_0 .. K_{NKD-1}
    static constexpr int QO = Tab::Q_ORDER;

    double* smem_base;           // first n_smem_vec working vectors
    double* slab_base;           // the others (per-CTA, L2-resident)
    unsigned n_pad, n_smem_vec;
    bool prepared;
    double* aux;
    double* red;
    unsigned n;
    const char* args;
    const double *rtolp, *atolp;          // this trajectory's tolerance rows
    // scalars, replicated in every thread (all derived from CTA-wide reductions, hence identical)
    double t_old, t_now, t_end, t_start, step, step_size, min_step, dir_inf, t_last;
    unsigned size, steps, nrej, nfev, len_t, te_old;
    int status;
    bool fwd, te_finished, error_flag;

    CYRK_DI double* vecp(int v) const
    {
        return (unsigned)v < n_smem_vec ? smem_base + (unsigned long long)v * n_pad : slab_base + (unsigned long long)((unsigned)v - n_smem_vec) * n_pad;
    }
    CYRK_DI double* K(int j) const { return vecp(V_K0 + j); }

    // dy[i] = f_i(t, y) for the whole vector; y must be complete (caller syncs before and after)
    // A problem given in the reference's WHOLE-VECTOR form -- void f(dy, t, y, args), c_common.hpp:218-220 -- cannot be
    // split over threads: one thread evaluates it into the derivative vector (the stage arithmetic around it stays
    // parallel).  This is how user systems with n_y > CYRK_MAX_NY keep the signature they have for small n_y.
    CYRK_DI void eval_rhs(double t, const double* y, double* dy)
    {
        if constexpr (Prob::WHOLE_VECTOR) { if (threadIdx.x == 0) Prob::eval_all(dy, t, y, args); }
        else for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dy[i] = Prob::eval(i, t, y, n, args, aux);
        nfev++;
        __syncthreads();
    }

    template <class CoefFn, int NTERMS>
    CYRK_DI double dot_k(unsigned i, CoefFn)
    {
        double acc = 0.0;
        bool first = true;
        static_for<0, NTERMS>([&](auto jj) {
            constexpr int J = decltype(jj)::value;
            constexpr double c = CoefFn::get(J);
            if constexpr (c != 0.0) {
                if (first) { acc = CoefFn::ld(J) * K(J)[i]; first = false; }
                else       { acc += CoefFn::ld(J) * K(J)[i]; }
            }
        });
        return acc;
    }
    template <int S> struct ARow { CYRK_HD static constexpr double get(int j) { return Tab::a(S, j); } CYRK_DI static double ld(int j) { return Ld::a(S, j); } };
    struct BRow { CYRK_HD static constexpr double get(int j) { return Tab::b(j); } CYRK_DI static double ld(int j) { return Ld::b(j); } };
    struct ERow { CYRK_HD static constexpr double get(int j) { return Tab::e(j); } CYRK_DI static double ld(int j) { return Ld::e(j); } };
    struct E3Row { CYRK_HD static constexpr double get(int j) { return TabDOP853::e3(j); } CYRK_DI static double ld(int j) { return LdDop::e3(j); } };
    struct E5Row { CYRK_HD static constexpr double get(int j) { return TabDOP853::e5(j); } CYRK_DI static double ld(int j) { return LdDop::e5(j); } };
    template <int C> struct PRow { CYRK_HD static constexpr double get(int j) { return Tab::p(C, j); } CYRK_DI static double ld(int j) { return Ld::p(C, j); } };
    template <int R> struct XRow { CYRK_HD static constexpr double get(int j) { return TabDOP853::ax(R, j); } CYRK_DI static double ld(int j) { return LdDop::ax(R, j); } };
    template <int R> struct DRow { CYRK_HD static constexpr double get(int j) { return TabDOP853::d(R, j); } CYRK_DI static double ld(int j) { return LdDop::d(R, j); } };

    // stages of one attempt (rk.cpp:695-776, 886-1025, 1154-1412)
    CYRK_DI void compute_stages()
    {
        double* y_old = vecp(V_YOLD);
        double* y_now = vecp(V_YNOW);
        {
            const double a10 = Ld::a(1, 0);
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
                const double k0 = K(0)[i];
                if constexpr (Tab::STAGE1_STEP_FIRST) y_now[i] = y_old[i] + (step * k0 * a10);
                else                                  y_now[i] = y_old[i] + step * (a10 * k0);
            }
            __syncthreads();
            eval_rhs(t_old + Ld::c(1) * step, y_now, K(1));
        }
        static_for<2, NS>([&](auto ss) {
            constexpr int S = decltype(ss)::value;
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) y_now[i] = y_old[i] + (step * dot_k<ARow<S>, S>(i, ARow<S>{}));
            __syncthreads();
            eval_rhs(t_old + Ld::c(S) * step, y_now, K(S));
        });
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) y_now[i] = y_old[i] + (step * dot_k<BRow, NS>(i, BRow{}));
        __syncthreads();
        eval_rhs(t_now, y_now, K(NS));
    }

    // rk.cpp:778-823, 1027-1075, 1414-1489 (parallel-tree sum instead of the sequential loop)
    CYRK_DI double estimate_error(const CtaParams& p)
    {
        const double* y_old = vecp(V_YOLD);
        const double* y_now = vecp(V_YNOW);
        if constexpr (!Tab::IS_DOP) {
            double part = 0.0;
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
                const double e = dot_k<ERow, NK>(i, ERow{});
                const double sc = e / (atolp[i] + fmax(fabs(y_old[i]), fabs(y_now[i])) * rtolp[i]);
                part += (sc * sc);
            }
            const double norm = cta_sum(part, red);
            return step_size * sqrt(norm) / sqrt((double)n);
        } else {
            double p3 = 0.0, p5 = 0.0;
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
                const double scale_inv = 1.0 / (atolp[i] + fmax(fabs(y_old[i]), fabs(y_now[i])) * rtolp[i]);
                const double e3 = scale_inv * dot_k<E3Row, NS>(i, E3Row{});
                const double e5 = scale_inv * dot_k<E5Row, NS>(i, E5Row{});
                p3 += (e3 * e3);
                p5 += (e5 * e5);
            }
            const double n3 = cta_sum(p3, red);
            const double n5 = cta_sum(p5, red);
            if ((n5 == 0.0) && (n3 == 0.0)) return 0.0;
            const double denom = n5 + 0.01 * n3;
            return step_size * n5 / sqrt(denom * (double)n);
        }
    }

    CYRK_DI double min_step_at(double t)
    {
        const long long b = __double_as_longlong(t);
        double nxt;
        if (t == 0.0) nxt = __longlong_as_double(fwd ? 1LL : (long long)0x8000000000000001ULL);
        else nxt = __longlong_as_double(b + (((t > 0.0) == fwd) ? 1LL : -1LL));
        return 10. * fabs(nxt - t);
    }

    // rk.cpp:183-282
    CYRK_DI void first_step_size(const CtaParams& p)
    {
        const double* y_old = vecp(V_YOLD);
        double* y_now = vecp(V_YNOW);
        double a = 0.0, b = 0.0;
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
            const double scale = atolp[i] + fabs(y_old[i]) * rtolp[i];
            const double a0 = y_old[i] / scale;
            const double a1 = K(0)[i] / scale;
            a += (a0 * a0);
            b += (a1 * a1);
        }
        const double ny_sqrt = sqrt((double)n);
        const double d0 = sqrt(cta_sum(a, red)) / ny_sqrt;
        const double d1 = sqrt(cta_sum(b, red)) / ny_sqrt;
        double h0 = 1.0e-6;
        if (!((d0 < 1.0e-5) || (d1 < 1.0e-5))) h0 = 0.01 * d0 / d1;
        const double h0_dir = fwd ? h0 : -h0;
        t_now = t_old + h0_dir;
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) y_now[i] = y_old[i] + h0_dir * K(0)[i];
        __syncthreads();
        eval_rhs(t_now, y_now, K(1));           // scratch use of K_1: overwritten by the first attempt
        double c = 0.0;
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
            const double scale = atolp[i] + fabs(y_old[i]) * rtolp[i];
            const double a2 = (K(1)[i] - K(0)[i]) / scale;
            c += (a2 * a2);
        }
        const double d2 = sqrt(cta_sum(c, red)) / (h0 * ny_sqrt);
        double h1;
        if ((d1 <= 1.0e-15) && (d2 <= 1.0e-15)) h1 = fmax(1.0e-6, h0 * 1.0e-3);
        else h1 = pow((0.01 / fmax(d1, d2)), Tab::err_exponent());
        step_size = fmax(min_step_at(t_old), fmin(100.0 * h0, h1));
    }

    // interpolated component i at x = (t - t_old)/h for the step just taken (dense.cpp:71-170, rk.cpp set_Q_array)
    CYRK_DI double interp_component(unsigned i, double x, double h)
    {
        const double yo = vecp(V_YOLD)[i];
        if constexpr (!Tab::IS_DOP) {
            double cp = x;
            double acc = K(0)[i] * cp;               // Q_0 = k_0
            static_for<1, QO>([&](auto cc) {
                constexpr int C = decltype(cc)::value;
                cp *= x;
                acc += dot_k<PRow<C>, NK>(i, PRow<C>{}) * cp;
            });
            acc *= h;
            return yo + acc;
        } else {
            const double q0 = h * dot_k<DRow<3>, 16>(i, DRow<3>{});
            const double q1 = h * dot_k<DRow<2>, 16>(i, DRow<2>{});
            const double q2 = h * dot_k<DRow<1>, 16>(i, DRow<1>{});
            const double q3 = h * dot_k<DRow<0>, 16>(i, DRow<0>{});
            const double delta_y = vecp(V_YNOW)[i] - yo;
            const double sum_dy = K(NS)[i] + K(0)[i];
            const double q4 = 2.0 * delta_y - h * sum_dy;
            const double q5 = h * K(0)[i] - delta_y;
            const double q6 = delta_y;
            double acc = q0;
            acc *= x;
            acc += q1; acc *= (1.0 - x);
            acc += q2; acc *= x;
            acc += q3; acc *= (1.0 - x);
            acc += q4; acc *= x;
            acc += q5; acc *= (1.0 - x);
            acc += q6; acc *= x;
            return yo + acc;
        }
    }

    CYRK_DI void save_row(const CtaParams& p, unsigned long long traj, double t, const double* y)
    {
        const unsigned ndy = Prob::HAS_EXTRAS ? n + p.n_extra : n;
        const bool row = p.t_rows != nullptr && size < p.row_cap;
        if constexpr (Prob::HAS_EXTRAS) {               // extra outputs of this row: the functor's components n .. (see cta2_core.cuh)
            for (unsigned j = threadIdx.x; j < p.n_extra; j += blockDim.x) {
                const double v = Prob::eval(n + j, t, y, n, args, aux);
                if (row) p.y_rows[(traj * p.row_cap + size) * ndy + n + j] = v;
                p.y_last[traj * ndy + n + j] = v;
            }
            __syncthreads();         // (the functor read other threads' components of y)
        }
        if (row) {
            const unsigned long long r = traj * p.row_cap + size;
            if (threadIdx.x == 0) p.t_rows[r] = t;
            double* dst = p.y_rows + r * ndy;
            if (((n | ndy) & 1u) == 0u) {  // 128-bit stores: rows are 16-byte aligned when the row width is even
                double2* d2 = reinterpret_cast<double2*>(dst);
                for (unsigned i = threadIdx.x; i < (n >> 1); i += blockDim.x) d2[i] = make_double2(y[2 * i], y[2 * i + 1]);
            } else {
                for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[i] = y[i];
            }
        }
        if (p.n_t_eval > 0) {      // raw steps saved after t_eval has finished (quirk 6): y_last follows t_last
            double* yl = p.y_last + traj * ndy;
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) yl[i] = y[i];
        }
        t_last = t;
        size++;
    }

    // whole trajectory (cysolver.cpp:303-543 setup, 752-1000 loop)
    CYRK_DI void solve(const CtaParams& p, unsigned long long traj)
    {
        double* y_old = vecp(V_YOLD);
        double* y_now = vecp(V_YNOW);
        const double* ts = p.t_span + traj * p.t_span_stride;
        t_start = ts[0]; t_end = ts[1];
        fwd = (t_end - t_start) >= 0.0;
        dir_inf = fwd ? __longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double((long long)0xfff0000000000000ULL);
        args = p.args + traj * p.args_stride;
        rtolp = p.rtol + traj * p.tol_stride;              // (stride 0: one tolerance set for the batch)
        atolp = p.atol + traj * p.tol_stride;
        size = steps = nrej = nfev = len_t = 0;
        status = ST_NO_ERROR; error_flag = false; te_finished = false;
        t_old = t_now = t_start; step = 0.0; step_size = 0.0;
        t_last = __longlong_as_double(0x7ff8000000000000LL);
        const double* y0 = p.y0 + traj * n;
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) { const double v = y0[i]; y_old[i] = v; y_now[i] = v; }
        Prob::prepare(args, n, aux, prepared && p.args_stride == 0);   // cooperative; ends with __syncthreads()
        prepared = true;
        __syncthreads();
        eval_rhs(t_now, y_now, K(0));           // RHS evaluation #1
        if (p.n_t_eval == 0) { save_row(p, traj, t_now, y_now); }
        te_old = fwd ? 0u : p.n_t_eval;
        bool bad_first = false;
        if (p.first_step != 0.0) {
            if (p.first_step < 0.0 || p.first_step > (fabs(t_end - t_start) * 0.5)) { status = ST_BAD_INITIAL_STEP_SIZE; bad_first = true; }
            step_size = p.first_step;
        } else {
            first_step_size(p);
        }
        // ---- main loop: one reference take_step per iteration
        while (!bad_first) {
            if (status != ST_NO_ERROR || error_flag) break;
            if (t_now == t_end) { t_old = t_end; status = ST_SUCCESSFUL_INTEGRATION; break; }
            if (len_t >= p.max_num_steps) {
                error_flag = true;
                status = p.user_max_steps ? ST_MAX_STEPS_USER_EXCEEDED : ST_MAX_STEPS_SYSARCH_EXCEEDED;
                break;
            }
            min_step = min_step_at(t_old);
            if (step_size < min_step) step_size = min_step;
            else if (p.max_step < step_size) step_size = p.max_step;
            bool rejected = false;
            for (;;) {                              // rk.cpp:447-523
                if (!(step_size >= min_step)) { status = ST_STEP_SIZE_ERROR_SPACING; break; }
                double over;
                if (fwd) { step = step_size;  t_now = t_old + step; over = t_now - t_end; }
                else     { step = -step_size; t_now = t_old + step; over = t_end - t_now; }
                if (over > 0.0) {
                    t_now = t_end;
                    step = t_now - t_old;
                    step_size = fwd ? step : -step;
                }
                compute_stages();
                const double err = estimate_error(p);
                const double err_safe = 0.9 / pow(err, Tab::err_exponent());
                if (err < 1.0) {
                    double f = 10.0;
                    if (err != 0.0) f = fmin(f, err_safe);
                    if (rejected) f = fmin(f, 1.0);
                    step_size *= f;
                    break;
                }
                step_size *= fmax(0.2, err_safe);
                rejected = true;
                nrej++;
            }
            len_t++;
            steps++;
            // ---- t_eval emission (cysolver.cpp:838-967) or plain save (:968-972)
            bool do_save = true;
            if (p.n_t_eval > 0 && !te_finished) {
                do_save = false;
                unsigned lo = 0, hi = p.n_t_eval;
                const double* tev = p.t_eval + traj * p.t_eval_stride;
                while (lo < hi) { const unsigned m = (lo + hi) >> 1; if (tev[m] < t_now) lo = m + 1; else hi = m; }
                unsigned up = lo;
                while (up < p.n_t_eval && !(t_now < tev[up])) up++;
                unsigned idx_new;
                if (lo == up) idx_new = lo;
                else if (fwd) { idx_new = up; if (idx_new == p.n_t_eval) te_finished = true; }
                else          { idx_new = lo; if (idx_new == 0) te_finished = true; }
                const bool any = fwd ? (idx_new > te_old) : (te_old > idx_new);
                const unsigned delta = fwd ? (idx_new - te_old) : (te_old - idx_new);
                if (any) {
                    const double h = t_now - t_old;
                    if constexpr (Tab::IS_DOP) {
                        // extra stages 13..15; the stage input is staged through `aux`-independent K slots
                        const double* yo = vecp(V_YOLD);
                        double* in = K(15);
                        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) in[i] = yo[i] + dot_k<XRow<0>, 13>(i, XRow<0>{}) * step;
                        __syncthreads();
                        eval_rhs(t_old + (step * LdDop::cx(0)), in, K(13));
                        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) in[i] = yo[i] + dot_k<XRow<1>, 14>(i, XRow<1>{}) * step;
                        __syncthreads();
                        eval_rhs(t_old + (step * LdDop::cx(1)), in, K(14));
                        // third input needs K_13, K_14 only (ax(2,15) does not exist), so K_15 may hold it; the
                        // derivative is written to the problem-independent spare slot and copied back.
                        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) in[i] = yo[i] + dot_k<XRow<2>, 15>(i, XRow<2>{}) * step;
                        __syncthreads();
                        double* spare = aux + (unsigned long long)p.n_pad;
                        eval_rhs(t_old + (step * LdDop::cx(2)), in, spare);
                        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) in[i] = spare[i];
                        __syncthreads();
                    }
                    for (unsigned q = 0; q < delta; q++) {
                        const double ti = fwd ? tev[te_old + q] : tev[te_old - q - 1];
                        const double x = (ti - t_old) / h;
                        // interpolate straight into the output row
                        if (p.t_rows != nullptr && size < p.row_cap) {
                            const unsigned long long r = traj * p.row_cap + size;
                            if (threadIdx.x == 0) p.t_rows[r] = ti;
                            double* dst = p.y_rows + r * n;
                            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[i] = interp_component(i, x, h);
                        }
                        if (q + 1 == delta) {
                            double* yl = p.y_last + traj * n;
                            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) yl[i] = interp_component(i, x, h);
                        }
                        t_last = ti;
                        size++;
                    }
                }
                te_old = idx_new;
            }
            if (do_save && !error_flag) save_row(p, traj, t_now, y_now);
            // roll: y_old <- y_now, K_0 <- K_NS (FSAL)
            t_old = t_now;
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) { y_old[i] = y_now[i]; K(0)[i] = K(NS)[i]; }
            __syncthreads();
        }
        // ---- results
        if (p.n_t_eval == 0) {
            double* yl = p.y_last + traj * (Prob::HAS_EXTRAS ? n + p.n_extra : n);
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) yl[i] = y_now[i];
        } else if (size == 0) {
            double* yl = p.y_last + traj * n;
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x) yl[i] = t_last;   // NaN
        }
        if (threadIdx.x == 0) {
            p.status[traj] = status;
            p.size[traj] = size;
            p.steps_taken[traj] = steps;
            p.n_rejected[traj] = nrej;
            p.nfev[traj] = nfev;
            p.t_last[traj] = t_last;
        }
        __syncthreads();
    }
};

template <class Prob, int METHOD>
CYRK_DI void rk_cta_body(const CtaParams& p)
{
    using Solver = CtaSolver<Prob, METHOD>;
    extern __shared__ double smem[];
    __shared__ double red[33];
    __shared__ unsigned long long next_traj;
    Solver S;
    S.n = p.n;
    S.red = red;
    S.prepared = false;
    // working vectors: the first n_smem_vec in shared memory, the rest in this CTA's L2-resident slab
    const unsigned n_glob = Solver::NVEC > p.n_smem_vec ? Solver::NVEC - p.n_smem_vec : 0u;
    double* slab = p.scratch + (unsigned long long)blockIdx.x * (unsigned long long)(n_glob + 2) * p.n_pad;
    S.smem_base = smem;
    S.slab_base = slab;
    S.n_pad = p.n_pad;
    S.n_smem_vec = p.n_smem_vec;
    S.aux = slab + (unsigned long long)n_glob * p.n_pad;          // aux table [n_pad] followed by one spare vector [n_pad]
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) next_traj = atomicAdd(p.queue, 1ULL);
        __syncthreads();
        const unsigned long long traj = next_traj;
        if (traj >= p.n_traj) break;
        S.solve(p, traj);
    }
}

template <class Prob, int METHOD>
__global__ void __launch_bounds__(1024, 1) rk_cta_kernel(const CtaParams p)
{
    rk_cta_body<Prob, METHOD>(p);
}

}  // namespace CYRK_FP_NS
}  // namespace cyrk
