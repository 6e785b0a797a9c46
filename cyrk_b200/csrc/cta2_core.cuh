// cta2_core.cuh -- adaptive explicit Runge-Kutta solver for LARGE systems, one CTA per trajectory,
// REGISTER-RESIDENT variant: n_y <= 512 * C components, C in {1, 2, 4, 8} with every working vector on chip
// (n_y <= 4096), C in {12, 16, 20, 24} with the stage derivatives in an L2-resident slab (n_y <= 12 288; the reference's
// own largey has 10 000).  cta_core_params.h holds the placement of every vector as a function of (method, C).
//
// The first CTA kernel (cta_core.cuh) keeps every working vector in shared memory / an L2 slab behind a run-time
// address selection and brackets every loop with __syncthreads(); it measured 11.7 SM-cycles per component-attempt
// on BASELINE config 4 (largey, n_y = 4096, RK45) against an FP64-pipe bound of ~4.5; this kernel measures 4.5 with the
// pipe 55 % busy (profiles/r02_notes.md, section 7).  Here
//   * thread t OWNS the components i = t + 512*c, c < C, for the whole trajectory (C = 1: one thread per component):
//     y_old and the most frequently read stage derivatives K_0..K_{NREG-1} of those components live in REGISTERS, the
//     next NSM derivatives in shared memory, the remainder in a per-CTA L2-resident slab -- placements, strides and
//     vector lengths are compile-time and the solver holds ONE shared-memory base pointer, so every access is a plain
//     register / LDS / LDG with a constant offset;
//   * the only data other threads ever read is the stage INPUT vector (the right-hand side may couple components),
//     which is double-buffered in shared memory: ONE __syncthreads() per right-hand-side evaluation, plus one for
//     the (double-buffered) error-norm reduction -- 7 barriers per RK45 attempt instead of 16;
//   * the right-hand side is inlined into a rolled loop over the thread's components (one copy of its code per stage)
//     whose results pass through the thread's own slots of a memory vector -- see eval_into / eval_rhs;
//   * the problem's auxiliary table (largey: the per-component rates) sits in shared memory as well (slab from C = 20);
//   * every step-size decision comes from a CTA-wide reduction, so all threads take the same branch;
//   * accepted rows are streamed to HBM straight from the owning threads, one coalesced row per step.
//
// Reference behaviour restated: the same file:line map as rk_core.cuh / cta_core.cuh.  The production build uses the
// thread path's controller arithmetic (reciprocal instead of division, squared norm, one-Newton-step root); the strict
// build uses the reference formulas verbatim.  The error norm is summed by a tree (reference: sequential loop,
// rk.cpp:1047-1072): an O(eps*sqrt(n_y)) difference, documented in DESIGN.md.
//
// Stored dense output: one coalesced record [y_old | Q_0 | .. | Q_{QO-1}] per accepted step, evaluated afterwards by
// cta_dense_eval_kernel.  Events: the values g_e(t, y) are evaluated by thread 0 on the complete vector and published
// through shared memory; every thread then runs the reference's trigger logic and Brent iteration redundantly (identical
// inputs -> identical branches), interpolating its own components of the state for each evaluation.  Extra outputs
// (problems that declare HAS_EXTRAS): the functor's components n_y .., evaluated at every saved row; not together with
// t_eval, events or a dense store (the host shim says so).
#pragma once

#ifndef CYRK_EVAL_WIDTH
#define CYRK_EVAL_WIDTH 2      // components per call of a problem's phase-interleaved evaluation (eval_uw), where it has one
#endif
#ifndef CYRK_EVAL_UNROLL
#define CYRK_EVAL_UNROLL 8     // unroll factor of the per-thread component loop of a right-hand-side evaluation
#endif
#ifndef CYRK_EVAL_UNROLL_LARGE
#define CYRK_EVAL_UNROLL_LARGE 4     // the same for C >= 12 (stage derivatives in the slab)
#endif
#include "rk_core.cuh"
#include "cta_core_params.h"

namespace cyrk {
namespace CYRK_FP_NS {

template <class...> struct VoidT { typedef void type; };
template <class P, class = void> struct HasTprep { static constexpr bool value = false; };
template <class P> struct HasTprep<P, typename VoidT<decltype(P::tprep(0.0, (const char*)nullptr))>::type> { static constexpr bool value = true; };

// EVENTS: compiled with the event machinery (a separate instantiation: its per-event state and extra pointers would
// otherwise push the event-free kernel over its 128-register budget).
template <class Prob, int METHOD, int C, bool EVENTS>
struct Cta2Solver {
    using Tab = typename TabFor<METHOD>::type;
    using Ld = TabLd<Tab>;
    using LdDop = TabLd<TabDOP853>;
    static constexpr int NS = Tab::NS;
    static constexpr int NK = Tab::NK;
    static constexpr int NKD = Tab::NK_DENSE;
    static constexpr int QO = Tab::Q_ORDER;
    // Placement of K_0 .. K_{NKD-1}.  Registers: y_old plus NREG derivative vectors within a budget of 40 doubles per
    // thread.  Shared memory: two stage-input buffers, the aux table, then NSM derivative vectors -- sized for the
    // largest block (512 threads): 227 KB / (C*512*8 B) slots.
    static constexpr int NREG = cta2_nreg(NKD, C);
    static constexpr int NSM = cta2_nsm(NKD, C);
    static constexpr int NGL = cta2_ngl(NKD, C);
    static_assert(NKD == cta2_nkd(METHOD), "cta_core_params.h out of sync with tableaux.cuh");

    double yo[C];
    double kr[NREG > 0 ? NREG : 1][C];
    // Shared memory, V doubles per vector: [buf 0 | buf 1 | aux | K_NREG .. K_{NREG+NSM-1}] (+ one staging vector when
    // NSM == 0).  The solver keeps ONE base pointer and small indices instead of a pointer per vector: with C >= 2 every
    // address is base + thread offset + immediate, and the 64-bit pointers do not compete with the state for registers.
    double* sm;
    int yb;                      // stage-input buffer that holds y_now (last y_new computed, or y0; the terminal state after an event)
    int ys;                      // buffer with y_new of the step just taken (the interpolant's end point; == yb except after termination)
    double* gsh;                 // shared memory [1 + CYRK_MAX_EVENTS]: event values published by thread 0
    double* xbuf;                // global scratch vector [V]: the interpolated state at xb during a root search
    double* kgl;                 // [NGL][V] global slab
    double (*red)[32];
    int red_phase;
    unsigned V, T, tid, n;
    bool prepared;
    const char* args;
    double t_old, t_now, t_end, t_start, step, step_size, min_step, t_last, inv_n;
    unsigned size, steps, nrej, nfev, te_old, n_dense;
    int status;
    bool fwd, te_finished, error_flag;
    bool uni_tol;
    double rt0, at0;

    // With more than one component per thread the CTA always has cta2_threads(C) threads, so the component stride and
    // the vector length are compile-time: every shared-memory access is [thread base + immediate].
    static constexpr unsigned TB = (unsigned)cta2_threads(C);
    CYRK_DI unsigned comp(int c) const { if constexpr (C >= 2) return tid + (unsigned)c * TB; else return tid; }
    CYRK_DI unsigned vlen() const { if constexpr (C >= 2) return (unsigned)C * TB; else return V; }
    CYRK_DI double* buf(int b) const { return sm + (b ? vlen() : 0u); }
    CYRK_DI double* ybuf() const { return buf(yb); }
    CYRK_DI const double* ystep() const { return buf(ys); }
    static constexpr bool AUX_GLOBAL = cta2_aux_global(C);
    static constexpr int NGL_ALLOC = NGL > 0 ? NGL : 1;
    static constexpr int SLAB_VECTORS = cta2_slab_vectors(NKD, C);
    CYRK_DI double* aux() const
    {
        if constexpr (AUX_GLOBAL) return kgl + (unsigned long long)(NGL_ALLOC + 1) * vlen();
        else return sm + 2u * vlen();
    }
    CYRK_DI double* ksm() const { return sm + 3u * vlen(); }
    // Staging vector of eval_rhs (and of first_step_size's trial evaluation): the vector of the first memory-resident
    // K_j with j >= 1 -- dead whenever a register-resident K_J, J < j, or the trial derivative is computed -- or a
    // vector of its own when every K is in registers.
    static constexpr int STG_J = NREG > 1 ? NREG : 1;
    CYRK_DI double* stg() const
    {
        if constexpr (STG_J >= NKD) return sm + 3u * vlen();
        else return kvec<STG_J>();
    }
    template <int J> CYRK_DI double* kvec() const        // where K_J lives when it is not in registers
    {
        static_assert(J >= NREG, "K_J is register-resident");
        if constexpr (J < NREG + NSM) return ksm() + (unsigned)(J - NREG) * vlen();
        else return kgl + (unsigned long long)(J - NREG - NSM) * vlen();
    }
    template <int J> CYRK_DI double kget(int c) const
    {
        if constexpr (J < NREG) return kr[J][c];
        else return kvec<J>()[comp(c)];
    }
    template <int J> CYRK_DI void kset(int c, double v)
    {
        if constexpr (J < NREG) kr[J][c] = v;
        else kvec<J>()[comp(c)] = v;
    }
    CYRK_DI double rtol_at(const CtaParams& p, unsigned i) const { return uni_tol ? rt0 : p.rtol[i]; }
    CYRK_DI double atol_at(const CtaParams& p, unsigned i) const { return uni_tol ? at0 : p.atol[i]; }

    // Sum over the CTA with ONE barrier: warp shuffles, one shared-memory hop into the phase's row, then every warp
    // reduces the (<= 32) warp partials redundantly.  The next call uses the other row, and the barrier of this call
    // orders it after every read of that row.  Summation order is fixed (tree), so results do not depend on timing.
    CYRK_DI double cta_sum(double v)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const unsigned warp = tid >> 5, lane = tid & 31u, nwarp = (T + 31u) >> 5;
        if (lane == 0) red[red_phase][warp] = v;
        __syncthreads();
        double w = lane < nwarp ? red[red_phase][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        red_phase ^= 1;
        return w;
    }

    template <class CoefFn, int NTERMS>
    CYRK_DI double dot_k(int c, CoefFn) const
    {
        double acc = 0.0;
        bool first = true;
        static_for<0, NTERMS>([&](auto jj) {
            constexpr int J = decltype(jj)::value;
            constexpr double cf = CoefFn::get(J);
            if constexpr (cf != 0.0) {
                if (first) { acc = CoefFn::ld(J) * kget<J>(c); first = false; }
                else       { acc += CoefFn::ld(J) * kget<J>(c); }
            }
        });
        return acc;
    }
    template <int S> struct ARow { CYRK_HD static constexpr double get(int j) { return Tab::a(S, j); } CYRK_DI static double ld(int j) { return Ld::a(S, j); } };
    struct BRow { CYRK_HD static constexpr double get(int j) { return Tab::b(j); } CYRK_DI static double ld(int j) { return Ld::b(j); } };
    struct ERow { CYRK_HD static constexpr double get(int j) { return Tab::e(j); } CYRK_DI static double ld(int j) { return Ld::e(j); } };
    struct E3Row { CYRK_HD static constexpr double get(int j) { return TabDOP853::e3(j); } CYRK_DI static double ld(int j) { return LdDop::e3(j); } };
    struct E5Row { CYRK_HD static constexpr double get(int j) { return TabDOP853::e5(j); } CYRK_DI static double ld(int j) { return LdDop::e5(j); } };
    template <int Q> struct PRow { CYRK_HD static constexpr double get(int j) { return Tab::p(Q, j); } CYRK_DI static double ld(int j) { return Ld::p(Q, j); } };
    template <int R> struct XRow { CYRK_HD static constexpr double get(int j) { return TabDOP853::ax(R, j); } CYRK_DI static double ld(int j) { return LdDop::ax(R, j); } };
    template <int R> struct DRow { CYRK_HD static constexpr double get(int j) { return TabDOP853::d(R, j); } CYRK_DI static double ld(int j) { return LdDop::d(R, j); } };

    // K_J = f(t, in) for the components this thread owns; `in` must be complete (the caller has synchronised).
    // The functor is inlined into a ROLLED loop over the thread's components (unrolled CYRK_EVAL_UNROLL times so two
    // dependent chains interleave): one copy of its code per stage instead of C (fully unrolled it was 19 456
    // instructions for largey/RK45/C=8 and 30 % of the stall samples were instruction fetch; as one call per component
    // it paid its literals and ~25 marshalling instructions per component, profiles/r01_notes.md), loop-invariant
    // work (constants, anything that depends on t only) hoisted by the compiler.  A rolled loop cannot index a register
    // array, so the values go to the thread's own slots of a shared-memory vector: K_J's own vector when K_J lives in
    // memory, else `stg` -- the vector of the first memory-resident K (dead while a register-resident K_J, J < NREG, is
    // computed: stages run in order and every K_j, j >= 1, dies with the attempt) or a vector of its own -- and are
    // read back with immediate offsets.  Own slots only: no barrier.
    // A problem may split its functor into tprep(t, args) -- whatever depends on t and args only, computed once per
    // evaluation and thread -- and eval_u(i, u, ..., slow) (ProbLargeY: the phase 2*pi*t/5).  eval_u has no slow path of
    // its own (a branch per component would keep the unrolled chains from interleaving): it raises `slow` when an
    // argument left the range of its fast arithmetic, and the thread then redoes its components with eval().
    // Slots past n_y (the CTA's padding) evaluate component n_y - 1: no branch, and nobody reads those slots.
    CYRK_DI void eval_into(double* dst, double t, const double* in) const
    {
        // (slab-resident range, C >= 12: unrolled 4 times -- n_y = 10 000: 18.1 ms against 19.3 with 8 and 29.7 fully
        // unrolled, profiles/r02_ab_c4ref_unroll.log)
        constexpr int UNR_CFG = C >= 12 ? CYRK_EVAL_UNROLL_LARGE : CYRK_EVAL_UNROLL;
        constexpr int UNR = UNR_CFG < C ? UNR_CFG : C;
        const unsigned last = n - 1u;
        if constexpr (HasTprep<Prob>::value) {
            auto u = Prob::tprep(t, args);
            asm volatile("" : "+d"(u));             // opaque: not recomputed inside the loop
            unsigned slow = 0u;
            constexpr int W = CYRK_EVAL_WIDTH < C ? CYRK_EVAL_WIDTH : C;
            if constexpr (W > 1 && C % W == 0) {
                constexpr int UNG = (UNR / W) > 0 ? (UNR / W) : 1;
#pragma unroll(UNG)
                for (int c0 = 0; c0 < C; c0 += W) {
                    unsigned ii[W];
                    double out[W];
#pragma unroll
                    for (int w = 0; w < W; w++) { const unsigned i = comp(c0 + w); ii[w] = i < last ? i : last; }
                    Prob::template eval_uw<W>(ii, u, in, n, args, aux(), slow, out);
#pragma unroll
                    for (int w = 0; w < W; w++) dst[comp(c0 + w)] = out[w];
                }
            } else {
#pragma unroll(UNR)
                for (int c = 0; c < C; c++) {
                    const unsigned i = comp(c);
                    dst[i] = Prob::eval_u(i < last ? i : last, u, in, n, args, aux(), slow);
                }
            }
            if (Prob::is_slow(slow)) {
#pragma unroll 1
                for (int c = 0; c < C; c++) {
                    const unsigned i = comp(c);
                    dst[i] = Prob::eval(i < last ? i : last, t, in, n, args, aux());
                }
            }
        } else {
#pragma unroll(UNR)
            for (int c = 0; c < C; c++) {
                const unsigned i = comp(c);
                dst[i] = Prob::eval(i < last ? i : last, t, in, n, args, aux());
            }
        }
    }
    template <int J>
    CYRK_DI void eval_rhs(double t, const double* in)
    {
        if constexpr (Prob::INLINE_EVAL) {           // a handful of instructions: straight into K_J
#pragma unroll
            for (int c = 0; c < C; c++) {
                const unsigned i = comp(c);
                kset<J>(c, i < n ? Prob::eval(i, t, in, n, args, aux()) : 0.0);
            }
        } else if constexpr (J < NREG) {
            eval_into(stg(), t, in);
#pragma unroll
            for (int c = 0; c < C; c++) kr[J][c] = stg()[comp(c)];
        } else {
            eval_into(kvec<J>(), t, in);
        }
        nfev++;
    }

    // stages of one attempt (rk.cpp:695-776, 886-1025, 1154-1412); leaves y_new in `ybuf`
    CYRK_DI void compute_stages()
    {
        // never start in the buffer that holds y_now (it is still read below); then alternate
        int ib = yb ^ 1;
        double* in = buf(ib);
        {
            const double a10 = Ld::a(1, 0);
#pragma unroll
            for (int c = 0; c < C; c++) {
                const double k0 = kget<0>(c);
                if constexpr (Tab::STAGE1_STEP_FIRST) in[comp(c)] = yo[c] + (step * k0 * a10);
                else                                  in[comp(c)] = yo[c] + step * (a10 * k0);
            }
            __syncthreads();
            eval_rhs<1>(t_old + Ld::c(1) * step, in);
            ib ^= 1; in = buf(ib);
        }
        static_for<2, NS>([&](auto ss) {
            constexpr int S = decltype(ss)::value;
#pragma unroll
            for (int c = 0; c < C; c++) in[comp(c)] = yo[c] + (step * dot_k<ARow<S>, S>(c, ARow<S>{}));
            __syncthreads();
            eval_rhs<S>(t_old + Ld::c(S) * step, in);
            ib ^= 1; in = buf(ib);
        });
#pragma unroll
        for (int c = 0; c < C; c++) in[comp(c)] = yo[c] + (step * dot_k<BRow, NS>(c, BRow{}));
        __syncthreads();
        eval_rhs<NS>(t_now, in);
        yb = ib;
        ys = ib;
    }

    // rk.cpp:778-823, 1027-1075, 1414-1489.  Strict build: returns err (reference formulas).  Production build:
    // returns err^2 (reciprocal + squared norm, like the thread path).
    CYRK_DI double estimate_error(const CtaParams& p)
    {
        if constexpr (!Tab::IS_DOP) {
            double part = 0.0;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const unsigned i = comp(c);
                if (i < n) {
                    const double e = dot_k<ERow, NK>(c, ERow{});
                    const double yn = ybuf()[i];
#if CYRK_FP_STRICT
                    const double sc = e / (atol_at(p, i) + fmax(fabs(yo[c]), fabs(yn)) * rtol_at(p, i));
                    part += (sc * sc);
#else
                    const double sc = e * fast_rcp(fma(fabs(larger_mag(yo[c], yn)), rtol_at(p, i), atol_at(p, i)));
                    part = fma(sc, sc, part);
#endif
                }
            }
            const double norm = cta_sum(part);
#if CYRK_FP_STRICT
            return step_size * sqrt(norm) / sqrt((double)n);
#else
            return (step_size * step_size) * norm * inv_n;
#endif
        } else {
            double p3 = 0.0, p5 = 0.0;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const unsigned i = comp(c);
                if (i < n) {
                    const double yn = ybuf()[i];
#if CYRK_FP_STRICT
                    const double scale_inv = 1.0 / (atol_at(p, i) + fmax(fabs(yo[c]), fabs(yn)) * rtol_at(p, i));
#else
                    const double scale_inv = fast_rcp(fma(fabs(larger_mag(yo[c], yn)), rtol_at(p, i), atol_at(p, i)));
#endif
                    const double e3 = scale_inv * dot_k<E3Row, NS>(c, E3Row{});
                    const double e5 = scale_inv * dot_k<E5Row, NS>(c, E5Row{});
                    p3 += (e3 * e3);
                    p5 += (e5 * e5);
                }
            }
            const double n3 = cta_sum(p3);
            const double n5 = cta_sum(p5);
            if ((n5 == 0.0) && (n3 == 0.0)) return 0.0;
            const double denom = n5 + 0.01 * n3;
#if CYRK_FP_STRICT
            return step_size * n5 / sqrt(denom * (double)n);
#else
            return (step_size * step_size) * (n5 * n5) * inv_n / denom;
#endif
        }
    }

    // rk.cpp:183-282 (scalar part verbatim; runs once per trajectory)
    CYRK_DI void first_step_size(const CtaParams& p)
    {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const unsigned i = comp(c);
            if (i < n) {
                const double scale = atol_at(p, i) + fabs(yo[c]) * rtol_at(p, i);
                const double a0 = yo[c] / scale;
                const double a1 = kget<0>(c) / scale;
                a += (a0 * a0);
                b += (a1 * a1);
            }
        }
        const double ny_sqrt = sqrt((double)n);
        const double d0 = sqrt(cta_sum(a)) / ny_sqrt;
        const double d1 = sqrt(cta_sum(b)) / ny_sqrt;
        double h0 = 1.0e-6;
        if (!((d0 < 1.0e-5) || (d1 < 1.0e-5))) h0 = 0.01 * d0 / d1;
        const double h0_dir = fwd ? h0 : -h0;
        t_now = t_old + h0_dir;
        double* in = buf(1);                      // buffer 0 keeps y0 == y_now
#pragma unroll
        for (int c = 0; c < C; c++) in[comp(c)] = yo[c] + h0_dir * kget<0>(c);
        __syncthreads();
        double cc = 0.0;
        if constexpr (!Prob::INLINE_EVAL) eval_into(stg(), t_now, in);      // every K_j, j >= 1, is dead here
#pragma unroll
        for (int c = 0; c < C; c++) {
            const unsigned i = comp(c);
            if (i < n) {
                double f1;
                if constexpr (Prob::INLINE_EVAL) f1 = Prob::eval(i, t_now, in, n, args, aux());
                else f1 = stg()[i];
                const double scale = atol_at(p, i) + fabs(yo[c]) * rtol_at(p, i);
                const double a2 = (f1 - kget<0>(c)) / scale;
                cc += (a2 * a2);
            }
        }
        nfev++;
        const double d2 = sqrt(cta_sum(cc)) / (h0 * ny_sqrt);
        double h1;
        if ((d1 <= 1.0e-15) && (d2 <= 1.0e-15)) h1 = fmax(1.0e-6, h0 * 1.0e-3);
        else h1 = pow((0.01 / fmax(d1, d2)), Tab::err_exponent());
        step_size = fmax(min_step_at(t_old, fwd), fmin(100.0 * h0, h1));
    }

    // interpolated own component c at x = (t - t_old)/h for the step just taken (dense.cpp:71-170, rk.cpp set_Q_array)
    CYRK_DI double interp_component(int c, double x, double h) const
    {
        if constexpr (!Tab::IS_DOP) {
            double cp = x;
            double acc = kget<0>(c) * cp;               // Q_0 = k_0
            static_for<1, QO>([&](auto qq) {
                constexpr int Q = decltype(qq)::value;
                cp *= x;
                acc += dot_k<PRow<Q>, NK>(c, PRow<Q>{}) * cp;
            });
            acc *= h;
            return yo[c] + acc;
        } else {
            const double q0 = h * dot_k<DRow<3>, 16>(c, DRow<3>{});
            const double q1 = h * dot_k<DRow<2>, 16>(c, DRow<2>{});
            const double q2 = h * dot_k<DRow<1>, 16>(c, DRow<1>{});
            const double q3 = h * dot_k<DRow<0>, 16>(c, DRow<0>{});
            const double delta_y = (EVENTS ? ystep() : ybuf())[comp(c)] - yo[c];
            const double sum_dy = kget<NS>(c) + kget<0>(c);
            const double q4 = 2.0 * delta_y - h * sum_dy;
            const double q5 = h * kget<0>(c) - delta_y;
            const double q6 = delta_y;
            double acc = q0;
            acc *= x;
            acc += q1; acc *= (1.0 - x);
            acc += q2; acc *= x;
            acc += q3; acc *= (1.0 - x);
            acc += q4; acc *= x;
            acc += q5; acc *= (1.0 - x);
            acc += q6; acc *= x;
            return yo[c] + acc;
        }
    }

    // DOP853: the three extra stages of the 7th-order interpolant (rk.cpp:1491-1676).  Their inputs go through the
    // stage buffer that does not hold y_now.
    CYRK_DI void dense_extra_stages()
    {
        if constexpr (Tab::IS_DOP) {
            double* in = buf(yb ^ 1);
#pragma unroll
            for (int c = 0; c < C; c++) in[comp(c)] = yo[c] + dot_k<XRow<0>, 13>(c, XRow<0>{}) * step;
            __syncthreads();
            eval_rhs<13>(t_old + (step * LdDop::cx(0)), in);
            __syncthreads();                       // every thread is done reading `in` before it is rewritten
#pragma unroll
            for (int c = 0; c < C; c++) in[comp(c)] = yo[c] + dot_k<XRow<1>, 14>(c, XRow<1>{}) * step;
            __syncthreads();
            eval_rhs<14>(t_old + (step * LdDop::cx(1)), in);
            __syncthreads();
#pragma unroll
            for (int c = 0; c < C; c++) in[comp(c)] = yo[c] + dot_k<XRow<2>, 15>(c, XRow<2>{}) * step;
            __syncthreads();
            eval_rhs<15>(t_old + (step * LdDop::cx(2)), in);
            __syncthreads();
        }
    }

    CYRK_DI void save_row_from_ybuf(const CtaParams& p, unsigned long long traj, double t)
    {
        const unsigned ndy = Prob::HAS_EXTRAS ? n + p.n_extra : n;
        const bool row = p.t_rows != nullptr && size < p.row_cap;
        if (row) {
            const unsigned long long r = traj * p.row_cap + size;
            if (tid == 0) p.t_rows[r] = t;
            double* dst = p.y_rows + r * ndy;
#pragma unroll
            for (int c = 0; c < C; c++) { const unsigned i = comp(c); if (i < n) dst[i] = ybuf()[i]; }
        }
        if constexpr (Prob::HAS_EXTRAS) {
            // extra outputs of this row = components n.. of the functor at (t, y): what the right-hand-side call that
            // produced dy_now wrote to dy[n_y:] (same inputs -> same values).  ybuf is complete and visible here: the last
            // stage evaluation read it after a barrier.  (No t_eval / events / dense store with extras on this path.)
            for (unsigned j = tid; j < p.n_extra; j += T) {
                const double v = Prob::eval(n + j, t, ybuf(), n, args, aux());
                if (row) p.y_rows[(traj * p.row_cap + size) * ndy + n + j] = v;
                p.y_last[traj * ndy + n + j] = v;
            }
            __syncthreads();         // the functor read other threads' components of ybuf: nobody overwrites them before this
        }
        if (p.n_t_eval > 0) {
            // raw steps saved after t_eval has finished: keep y_last on the row t_last belongs to
            // (without t_eval the final write-out stores it once)
            double* yl = p.y_last + traj * ndy;
#pragma unroll
            for (int c = 0; c < C; c++) { const unsigned i = comp(c); if (i < n) yl[i] = ybuf()[i]; }
        }
        t_last = t;
        size++;
    }

    // One interpolant record (dense.cpp:27-37, rk.cpp set_Q_array: 825-862, 1077-1131, 1491-1676) for the step just
    // taken; `degenerate` = interpolant #0 at t0, whose K is still all zero.  DOP853: the caller
    // has run dense_extra_stages().
    CYRK_DI void store_dense(const CtaParams& p, unsigned long long traj, bool degenerate)
    {
        if (n_dense < p.dense_cap) {
            const unsigned long long r = traj * p.dense_cap + n_dense;
            if (tid == 0) p.dense_t[r] = t_now;
            double* rec = p.dense_state + r * (unsigned long long)n * (QO + 1);
#pragma unroll
            for (int c = 0; c < C; c++) {
                const unsigned i = comp(c);
                if (i < n) {
                    rec[i] = yo[c];
                    double q[QO];
                    if (degenerate) {
#pragma unroll
                        for (int k = 0; k < QO; k++) q[k] = 0.0;
                    } else if constexpr (!Tab::IS_DOP) {
                        q[0] = kget<0>(c);
                        static_for<1, QO>([&](auto qq) {
                            constexpr int Q = decltype(qq)::value;
                            q[Q] = dot_k<PRow<Q>, NK>(c, PRow<Q>{});
                        });
                    } else {
                        q[0] = step * dot_k<DRow<3>, 16>(c, DRow<3>{});
                        q[1] = step * dot_k<DRow<2>, 16>(c, DRow<2>{});
                        q[2] = step * dot_k<DRow<1>, 16>(c, DRow<1>{});
                        q[3] = step * dot_k<DRow<0>, 16>(c, DRow<0>{});
                        const double delta_y = (EVENTS ? ystep() : ybuf())[i] - yo[c];
                        const double sum_dy = kget<NS>(c) + kget<0>(c);
                        q[4] = 2.0 * delta_y - step * sum_dy;
                        q[5] = step * kget<0>(c) - delta_y;
                        q[6] = delta_y;
                    }
#pragma unroll
                    for (int k = 0; k < QO; k++) rec[(unsigned long long)n * (1 + k) + i] = q[k];
                }
            }
        }
        n_dense++;
    }

    // ---------------------------------------------------------------- events
    // g_e(t, y) for a state `y` that every thread has finished writing its components of; all threads get the value.
    CYRK_DI double event_value(const CtaParams& p, int e, double t, const double* y)
    {
        __syncthreads();                     // y is complete; everybody has read the previous gsh[0]
        if (tid == 0) gsh[0] = Prob::event(p.ev_id[e], t, y, n, args);
        __syncthreads();
        return gsh[0];
    }
    // own components of the step's interpolant at time t (dense.cpp:71-170) into dst
    CYRK_DI void interp_into(double* dst, double t, double told, double h) const
    {
        const double x = (t - told) / h;
#pragma unroll
        for (int c = 0; c < C; c++) dst[comp(c)] = interp_component(c, x, h);
    }

    // c_brentq.cpp:41-204 over [xa, xb] for event e, restated like TrajSolver::root_step / brent_advance (rk_core.cuh).
    // `scr` ends up holding the state of the LAST evaluation that went through it (the reference reports that buffer,
    // SURVEY 8a quirk 5).  Returns the error code (ST_CONVERGED = found) and the root.
    CYRK_DI int brentq(const CtaParams& p, int e, double xa, double xb, double told, double h, double* scr, double& root)
    {
        const double xtol = 4.0 * CYRK_EPS, rtol = 4.0 * CYRK_EPS;
        double xpre = xa, xcur = xb, xblk = 0., fblk = 0., spre = 0., scur = 0.;
        interp_into(scr, xpre, told, h);
        double fpre = event_value(p, e, xpre, scr);
        if (fpre == 0) { root = xpre; return ST_CONVERGED; }
        interp_into(xbuf, xcur, told, h);
        double fcur = event_value(p, e, xcur, xbuf);
        if (fcur == 0) {
#pragma unroll
            for (int c = 0; c < C; c++) scr[comp(c)] = xbuf[comp(c)];
            root = xcur;
            return ST_CONVERGED;
        }
        if (sign_bit(fpre) == sign_bit(fcur)) { root = 0.0; return ST_OPTIMIZE_SIGN_ERROR; }
        for (int iter = 0; iter < 100; iter++) {
            if ((fpre != 0) && (fcur != 0) && (sign_bit(fpre) != sign_bit(fcur))) {
                xblk = xpre; fblk = fpre;
                spre = scur = xcur - xpre;
            }
            if (fabs(fblk) < fabs(fcur)) {
                xpre = xcur; xcur = xblk; xblk = xpre;
                fpre = fcur; fcur = fblk; fblk = fpre;
            }
            const double delta = (xtol + rtol * fabs(xcur)) / 2.0;
            const double sbis = (xblk - xcur) / 2.0;
            if ((fcur == 0) || (fabs(sbis) < delta)) { root = xcur; return ST_CONVERGED; }
            if ((fabs(spre) > delta) && (fabs(fcur) < fabs(fpre))) {
                double stry;
                if (xpre == xblk) {
                    stry = -fcur * (xcur - xpre) / (fcur - fpre);
                } else {
                    const double dpre = (fpre - fcur) / (xpre - xcur);
                    const double dblk = (fblk - fcur) / (xblk - xcur);
                    stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
                }
                const double m_a = fabs(spre), m_b = 3.0 * fabs(sbis) - delta;
                if (2.0 * fabs(stry) < (m_a < m_b ? m_a : m_b)) { spre = scur; scur = stry; }
                else { spre = sbis; scur = sbis; }
            } else {
                spre = sbis; scur = sbis;
            }
            xpre = xcur; fpre = fcur;
            if (fabs(scur) > delta) xcur += scur;
            else xcur += (sbis > 0 ? delta : -delta);
            interp_into(scr, xcur, told, h);
            fcur = event_value(p, e, xcur, scr);
        }
        root = xcur;
        return ST_OPTIMIZE_CONVERGENCE_ERROR;
    }

    // whole trajectory (cysolver.cpp:303-543 setup, 752-1000 loop)
    CYRK_DI void solve(const CtaParams& p, unsigned long long traj)
    {
        const double* ts = p.t_span + traj * p.t_span_stride;
        t_start = ts[0]; t_end = ts[1];
        fwd = (t_end - t_start) >= 0.0;
        args = p.args + traj * p.args_stride;
        size = steps = nrej = nfev = n_dense = 0;
        status = ST_NO_ERROR; error_flag = false; te_finished = false;
        t_old = t_now = t_start; step = 0.0; step_size = 0.0;
        t_last = d_nan();
        inv_n = 1.0 / (double)n;
        const double* y0 = p.y0 + traj * n;
        yb = 0;
        ys = 0;
        __syncthreads();                                   // the previous trajectory's readers are done with the buffers
#pragma unroll
        for (int c = 0; c < C; c++) {
            const unsigned i = comp(c);
            const double v = i < n ? y0[i] : 0.0;
            yo[c] = v;
            buf(0)[i] = v;
        }
        Prob::prepare(args, n, aux(), prepared && p.args_stride == 0);   // cooperative; ends with __syncthreads()
        prepared = true;
        __syncthreads();
        eval_rhs<0>(t_now, ybuf());                         // RHS evaluation #1
        if (p.n_t_eval == 0) save_row_from_ybuf(p, traj, t_now);
        te_old = fwd ? 0u : p.n_t_eval;
        // per-event state, replicated in every thread (all copies evolve identically)
        constexpr int NEV = EVENTS ? CYRK_MAX_EVENTS : 1;
        double g_old[NEV], ev_root[NEV];
        unsigned long long ev_seen[NEV];
        unsigned ev_hits[NEV];
        bool event_terminated = false;
        int term_index = -1;
        if constexpr (EVENTS) {
            if (tid == 0)
                for (int e = 0; e < p.n_events; e++) gsh[1 + e] = Prob::event(p.ev_id[e], t_now, ybuf(), n, args);
            __syncthreads();
            for (int e = 0; e < p.n_events; e++) { g_old[e] = gsh[1 + e]; ev_seen[e] = 0; ev_hits[e] = 0; ev_root[e] = 0.0; }
        }
        if (p.dense_output) {
            // interpolant #0 at t0: Q == 0 (quirk 9); DOP853 spends its three extra evaluations on it in the reference
            store_dense(p, traj, true);
            if constexpr (Tab::IS_DOP) nfev += 3;
        }
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
            if (steps >= p.max_num_steps) {
                error_flag = true;
                status = p.user_max_steps ? ST_MAX_STEPS_USER_EXCEEDED : ST_MAX_STEPS_SYSARCH_EXCEEDED;
                break;
            }
            min_step = min_step_at(t_old, fwd);
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
#if CYRK_FP_STRICT
                const double err_safe = 0.9 / pow(err, Tab::err_exponent());
                if (err < 1.0) {
                    double f = 10.0;
                    if (err != 0.0) f = fmin(f, err_safe);
                    if (rejected) f = fmin(f, 1.0);
                    step_size *= f;
                    break;
                }
                step_size *= fmax(0.2, err_safe);
#else
                // err holds err^2 here (see estimate_error); same decisions, no pow()/sqrt()
                const double err_safe = safe_factor<Tab::METHOD - METHOD_RK23, 1>(err);
                if (err < 1.0) {
                    double f = fmin(10.0, err_safe);
                    if (rejected) f = fmin(f, 1.0);
                    step_size *= f;
                    break;
                }
                step_size *= (err_safe > 0.2) ? err_safe : 0.2;
#endif
                rejected = true;
                nrej++;
            }
            steps++;
            bool extra_done = false;
            if (p.dense_output) {                          // cysolver.cpp:785-806
                dense_extra_stages();
                extra_done = true;
                store_dense(p, traj, false);
            }
            const double h_step = t_now - t_old;          // the interpolant's (t_old, step): t_now may move to an event root
            const double t_step_old = t_old;
            bool extra_counted = false;
            if constexpr (EVENTS) {                        // p_check_events, cysolver.cpp:563-750
                if (tid == 0)
                    for (int e = 0; e < p.n_events; e++) gsh[1 + e] = Prob::event(p.ev_id[e], t_now, ybuf(), n, args);
                __syncthreads();
                unsigned mask = 0;
                for (int e = 0; e < p.n_events; e++) {
                    const double g_now = gsh[1 + e];
                    const double go = g_old[e];
                    const bool up = (go <= 0.0) && (g_now >= 0.0);
                    const bool down = (go >= 0.0) && (g_now <= 0.0);         // 0 -> 0 triggers (quirk 4)
                    const int dir = p.ev_dir[e];
                    const bool trig = (dir == 0 && (up || down)) || (dir > 0 && up) || (dir < 0 && down);
                    g_old[e] = g_now;
                    if (trig) { ev_seen[e]++; mask |= (1u << e); }
                }
                if constexpr (Tab::IS_DOP) {
                    // the reference builds the interpolant on every step with events: three more evaluations
                    if (!extra_done) {
                        if (mask != 0u) { dense_extra_stages(); extra_done = true; }
                        else { nfev += 3; extra_counted = true; }
                    }
                }
                if (mask != 0u) {
                    const int scr_b = yb ^ 1;
                    double* scr = buf(scr_b);
                    double term_root = fwd ? d_inf() : -d_inf();
                    for (int e = 0; e < p.n_events && !error_flag; e++) {
                        if (!(mask & (1u << e))) continue;
                        double root = 0.0;
                        const int err = brentq(p, e, t_old, t_now, t_step_old, h_step, scr, root);
                        if (err != ST_CONVERGED) { error_flag = true; status = err; break; }
                        ev_root[e] = root;
                        if (p.ev_t != nullptr && ev_hits[e] < p.ev_cap) {      // slot ev_hits[e]; committed below
                            const unsigned long long r = (traj * p.n_events + e) * (unsigned long long)p.ev_cap + ev_hits[e];
                            if (tid == 0) p.ev_t[r] = root;
                            double* dst = p.ev_y + r * n;
#pragma unroll
                            for (int c = 0; c < C; c++) { const unsigned i = comp(c); if (i < n) dst[i] = scr[i]; }
                        }
                        if (ev_seen[e] >= p.ev_max[e]) {
                            event_terminated = true;
                            if (fwd) { if (root < term_root) term_root = root; }
                            else     { if (root > term_root) term_root = root; }
                        }
                    }
                    if (!error_flag) {
                        // keep the hits that are not beyond the termination root (cysolver.cpp:706-748)
                        for (int e = 0; e < p.n_events; e++) {
                            if (!(mask & (1u << e))) continue;
                            if (event_terminated) {
                                if (ev_root[e] == term_root) term_index = e;
                                const bool beyond = fwd ? (ev_root[e] > term_root) : (ev_root[e] < term_root);
                                if (beyond) continue;
                            }
                            ev_hits[e]++;
                        }
                        status = ST_NO_ERROR;
                        if (event_terminated) {
                            status = ST_EVENT_TERMINATED;
                            if (t_now != term_root) {
                                // the trajectory ends at the root: its state is the interpolant there (cysolver.cpp:812-836)
                                __syncthreads();               // thread 0 may still be reading scr
                                t_now = term_root;
                                interp_into(scr, t_now, t_step_old, h_step);
                                yb = scr_b;
                            }
                        }
                    }
                } else {
                    status = ST_NO_ERROR;
                }
            }
            // ---- t_eval emission (cysolver.cpp:838-967) or plain save (:968-972)
            bool do_save = true;
            if (p.n_t_eval > 0 && !te_finished && !error_flag) {
                do_save = false;
                unsigned lo = 0, hi = p.n_t_eval;
                while (lo < hi) { const unsigned m = (lo + hi) >> 1; if (p.t_eval[m] < t_now) lo = m + 1; else hi = m; }
                unsigned up = lo;
                while (up < p.n_t_eval && !(t_now < p.t_eval[up])) up++;
                unsigned idx_new;
                if (lo == up) idx_new = lo;
                else if (fwd) { idx_new = up; if (idx_new == p.n_t_eval) te_finished = true; }
                else          { idx_new = lo; if (idx_new == 0) te_finished = true; }
                const bool any = fwd ? (idx_new > te_old) : (te_old > idx_new);
                const unsigned delta = fwd ? (idx_new - te_old) : (te_old - idx_new);
                if (any) {
                    const double h = h_step;
                    if (!extra_done) { dense_extra_stages(); if (extra_counted) nfev -= 3; }
                    for (unsigned q = 0; q < delta; q++) {
                        const double ti = fwd ? p.t_eval[te_old + q] : p.t_eval[te_old - q - 1];
                        const double x = (ti - t_old) / h;
                        const bool row = (p.t_rows != nullptr && size < p.row_cap);
                        const bool last = (q + 1 == delta);
                        if (row || last) {
                            const unsigned long long r = traj * p.row_cap + size;
                            if (row && tid == 0) p.t_rows[r] = ti;
#pragma unroll
                            for (int c = 0; c < C; c++) {
                                const unsigned i = comp(c);
                                if (i < n) {
                                    const double v = interp_component(c, x, h);     // interpolate straight into the output row
                                    if (row) p.y_rows[r * n + i] = v;
                                    if (last) p.y_last[traj * n + i] = v;
                                }
                            }
                        }
                        t_last = ti;
                        size++;
                    }
                }
                te_old = idx_new;
            }
            if (do_save && !error_flag) save_row_from_ybuf(p, traj, t_now);
            // roll: y_old <- y_now, K_0 <- K_NS (FSAL); own components only, no barrier needed
            t_old = t_now;
#pragma unroll
            for (int c = 0; c < C; c++) { yo[c] = ybuf()[comp(c)]; kset<0>(c, kget<NS>(c)); }
        }
        // ---- results
        if (p.n_t_eval == 0) {
            double* yl = p.y_last + traj * (Prob::HAS_EXTRAS ? n + p.n_extra : n);   // (the extras of the last saved row are already there)
#pragma unroll
            for (int c = 0; c < C; c++) { const unsigned i = comp(c); if (i < n) yl[i] = ybuf()[i]; }
        } else if (size == 0) {
            double* yl = p.y_last + traj * n;
#pragma unroll
            for (int c = 0; c < C; c++) { const unsigned i = comp(c); if (i < n) yl[i] = t_last; }   // NaN
        }
        if (tid == 0) {
            p.status[traj] = status;
            p.size[traj] = size;
            p.steps_taken[traj] = steps;
            p.n_rejected[traj] = nrej;
            p.nfev[traj] = nfev;
            p.t_last[traj] = t_last;
            if (p.dense_count != nullptr) p.dense_count[traj] = n_dense;
            if constexpr (EVENTS) {
                p.ev_term_index[traj] = term_index;
                for (int e = 0; e < p.n_events; e++) p.ev_count[traj * p.n_events + e] = ev_hits[e];
            }
        }
    }
};

template <class Prob, int METHOD, int C, bool EVENTS = false>
CYRK_DI void rk_cta2_body(const CtaParams& p)
{
    using Solver = Cta2Solver<Prob, METHOD, C, EVENTS>;
    extern __shared__ double smem[];
    __shared__ double red[2][32];
    __shared__ unsigned long long next_traj;
    __shared__ double ev_vals[1 + CYRK_MAX_EVENTS];
    Solver S;
    S.gsh = ev_vals;
    S.n = p.n;
    S.T = blockDim.x;
    S.tid = threadIdx.x;
    S.V = blockDim.x * (unsigned)C;
    S.red = red;
    S.red_phase = 0;
    S.prepared = false;
    S.sm = smem;
    // per-CTA global slab: the K vectors that fit neither registers nor shared memory, then one vector for root searches
    S.kgl = p.scratch + (unsigned long long)blockIdx.x * (unsigned long long)Solver::SLAB_VECTORS * S.V;
    S.xbuf = S.kgl + (unsigned long long)Solver::NGL_ALLOC * S.V;
    S.uni_tol = p.uniform_tol != 0;
    S.rt0 = p.rtol0;
    S.at0 = p.atol0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) next_traj = atomicAdd(p.queue, 1ULL);
        __syncthreads();
        const unsigned long long traj = next_traj;
        if (traj >= p.n_traj) break;
        S.solve(p, traj);
    }
}

template <class Prob, int METHOD, int C, bool EVENTS>
__global__ void __launch_bounds__(cta2_threads(C), 1) rk_cta2_kernel(const __grid_constant__ CtaParams p)
{
    rk_cta2_body<Prob, METHOD, C, EVENTS>(p);
}

// ------------------------------------------------------------------------------------------------
// Post-hoc dense evaluation for CTA-path solutions: one CTA per (trajectory, query time); semantics of
// CySolverResult::call / call_vectorize (cy/cysolution.cpp:511-616) and CySolverDense::call (cy/dense.cpp:71-170),
// the same lookup rules as dense_eval_body in rk_core.cuh.
// ------------------------------------------------------------------------------------------------
template <int METHOD>
CYRK_DI void cta_dense_eval_body(const CtaDenseEvalParams& p)
{
    using Tab = typename TabFor<METHOD>::type;
    constexpr int QO = Tab::Q_ORDER;
    const unsigned long long pair = blockIdx.x;
    const unsigned long long traj = pair / p.n_query;
    const unsigned q = (unsigned)(pair - traj * p.n_query);
    const int st = p.status[traj];
    int code = ST_NO_ERROR;
    const unsigned nrec = p.dense_count[traj] < p.dense_cap ? p.dense_count[traj] : p.dense_cap;
    const double* ts = p.t_span + traj * p.t_span_stride;
    const double t = p.t_query[q];
    if (!(st == ST_SUCCESSFUL_INTEGRATION || st == ST_EVENT_TERMINATED)) code = ST_INTEGRATION_NOT_SUCCESSFUL;
    else if (nrec == 0) code = ST_DENSE_OUTPUT_NOT_SAVED;
    else if ((t < fmin(ts[0], ts[1])) || (t > fmax(ts[0], ts[1]))) code = ST_ARGUMENT_ERROR;
    if (code != ST_NO_ERROR) {                       // uniform over the CTA
        if (threadIdx.x == 0) atomicCAS(p.code + traj, 0, code);
        return;
    }
    const double* times = p.dense_t + traj * p.dense_cap;
    const bool fwd = (ts[1] - ts[0]) >= 0.0;
    unsigned idx = 0;
    if (nrec > 1) {
        unsigned lo = 0, hi = nrec;
        if (fwd) {
            while (lo < hi) { const unsigned m = (lo + hi) >> 1; if (!(t < times[m])) lo = m + 1; else hi = m; }
            idx = lo;
        } else {
            while (lo < hi) { const unsigned m = (lo + hi) >> 1; if (times[nrec - 1 - m] < t) lo = m + 1; else hi = m; }
            idx = nrec - lo - 1;
        }
    }
    if (idx >= nrec) idx = nrec - 1;
    const double t_now_i = times[idx];
    const double t_old_i = idx == 0 ? times[0] : times[idx - 1];
    const double dstep = t_now_i - t_old_i;
    const double x = (t - t_old_i) / dstep;
    const unsigned long long n = p.n;
    const double* rec = p.dense_state + (traj * p.dense_cap + idx) * n * (QO + 1);
    double* out = p.out + pair * n;
    for (unsigned i = threadIdx.x; i < p.n; i += blockDim.x) {
        double qv[QO];
#pragma unroll
        for (int k = 0; k < QO; k++) qv[k] = rec[n * (1 + k) + i];
        double y;
        if constexpr (!Tab::IS_DOP) {
            double cp = x;
            double acc = qv[0] * cp;
#pragma unroll
            for (int pw = 1; pw < QO; pw++) { cp *= x; acc += qv[pw] * cp; }
            acc *= dstep;
            y = rec[i] + acc;
        } else {
            double acc = qv[0];
            acc *= x;
            acc += qv[1]; acc *= (1.0 - x);
            acc += qv[2]; acc *= x;
            acc += qv[3]; acc *= (1.0 - x);
            acc += qv[4]; acc *= x;
            acc += qv[5]; acc *= (1.0 - x);
            acc += qv[6]; acc *= x;
            y = rec[i] + acc;
        }
        out[i] = y;
    }
}

template <int METHOD>
__global__ void cta_dense_eval_kernel(const __grid_constant__ CtaDenseEvalParams p)
{
    cta_dense_eval_body<METHOD>(p);
}

}  // namespace CYRK_FP_NS
}  // namespace cyrk
