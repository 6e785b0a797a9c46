// cta_core_params.h -- launch parameters of the CTA-per-trajectory kernels (host + device).
#pragma once

namespace cyrk {

struct CtaParams {
    unsigned long long n_traj;
    unsigned n;                  // n_y
    unsigned n_pad;              // n rounded up to a multiple of 2 (128-bit accesses)
    unsigned n_extra;            // extra outputs: the functor's components n .. n + n_extra - 1, evaluated at every SAVED row
                                 // (the reference captures dy[n_y:], cysolver.cpp:945-961); rows and y_last are n + n_extra wide
    const double* y0;            // [n_traj][n]
    const char*   args;          // per-trajectory blob of args_stride bytes, or shared (stride 0)
    unsigned      args_stride;
    const double* t_span;
    unsigned      t_span_stride;
    const double* rtol;          // [n]  (floored at 100 eps on the host)
    const double* atol;          // [n]
    unsigned tol_stride;         // 0: one tolerance set for the batch; n: rtol / atol hold one row of n entries per trajectory
    int    uniform_tol;          // every component has the same tolerances (rtol0, atol0): no array reads
    double rtol0, atol0;
    double max_step, first_step;
    unsigned max_num_steps;
    int    user_max_steps;
    const double* t_eval;        // ascending (reversed copy for backward runs)
    unsigned n_t_eval;
    unsigned t_eval_stride;      // 0: one t_eval for the batch; n_t_eval: one row per trajectory
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
    // ---- dense store (register-resident kernel only): interpolant k covers (dense_t[k-1], dense_t[k]]; one record is
    //      [y_old[n] | Q_0[n] | ... | Q_{QO-1}[n]] (coefficient-major, so every store and load is coalesced)
    int       dense_output;
    double*   dense_t;           // [n_traj][dense_cap]
    double*   dense_state;       // [n_traj][dense_cap][n * (Q_ORDER + 1)]
    unsigned* dense_count;       // [n_traj]
    unsigned  dense_cap;
    // ---- events (register-resident kernel only): same records as the thread path (BatchParams)
    int       n_events;
    int       ev_id[8];
    int       ev_dir[8];
    unsigned long long ev_max[8];         // ~0ull = never terminate
    double*   ev_t;              // [n_traj][n_events][ev_cap]
    double*   ev_y;              // [n_traj][n_events][ev_cap][n]
    unsigned* ev_count;          // [n_traj][n_events]
    int*      ev_term_index;     // [n_traj]   (-1 = none)
    unsigned  ev_cap;
    double*   scratch;           // [gridDim.x][n_vec_global + 1][n_pad]: spilled vectors + the problem's aux table
    unsigned  n_smem_vec;        // how many of the working vectors live in shared memory
    unsigned long long* queue;
};

// Post-hoc dense evaluation of a CTA-path solution (CySolverResult::call_vectorize, cysolution.cpp:511-616): one CTA per
// (trajectory, query time), threads stride over the components.
struct CtaDenseEvalParams {
    unsigned long long n_traj;
    unsigned n;
    const double* t_query;       // [n_query] shared
    unsigned      n_query;
    const double* t_span;
    unsigned      t_span_stride;
    const int*      status;
    const double*   dense_t;
    const double*   dense_state;
    const unsigned* dense_count;
    unsigned        dense_cap;
    double*       out;           // [n_traj][n_query][n]
    int*          code;          // [n_traj]  first non-zero CySolverResult::call status, else 0
};

// ---- register-resident kernel (cta2_core.cuh): compile-time placement of the stage derivatives K_0..K_{nkd-1}
// for C components per thread.  Shared by the kernel and the host launcher.
// C components per thread, 512 threads (C = 1: one thread per component).  C in {1, 2, 4, 8}: n_y <= 4096, every vector
// on chip (registers + shared memory; DOP853 at C = 8 puts nine of its 16 stage vectors into the L2-resident slab).
// C in {12, 16, 20, 24}: n_y <= 12 288 (the reference's largey has 10 000): the two stage-input buffers stay in shared
// memory, y_old (and K_0 while it fits) in registers, the other stage derivatives -- and from C = 20 the problem's aux
// table -- in the per-CTA slab of global memory, which stays L2-resident (148 CTAs x <= 9 vectors x <= 96 KB << 126 MB).
constexpr int cta2_max_block = 512;
constexpr int cta2_max_c_onchip = 8;
constexpr int cta2_max_c = 24;
constexpr int cta2_threads(int) { return cta2_max_block; }
constexpr int cta2_choose_c(unsigned long long n)
{
    return n <= 512 ? 1 : n <= 1024 ? 2 : n <= 2048 ? 4 : n <= 4096 ? 8 : n <= 6144 ? 12 : n <= 8192 ? 16 : n <= 10240 ? 20 : n <= 12288 ? 24 : 0;
}
// doubles of state per thread kept in registers: y_old + NREG derivative vectors
#ifndef CYRK_CTA2_REGD_LARGE      // experiment knob: multiple of C doubles kept in registers for C >= 20
#define CYRK_CTA2_REGD_LARGE 1
#endif
constexpr int cta2_reg_doubles(int C) { return C >= 20 ? CYRK_CTA2_REGD_LARGE * C : C >= 12 ? 2 * C : C >= 8 ? 32 : 40; }
constexpr int cta2_nkd(int method) { return method == 3 ? 4 : method == 4 ? 7 : 16; }      // RK23 / RK45 / DOP853 (+3 dense)
constexpr int cta2_nreg(int nkd, int C) { return (cta2_reg_doubles(C) / C - 1) < nkd ? (cta2_reg_doubles(C) / C - 1) : nkd; }
constexpr int cta2_sm_slots(int C) { return (227 * 1024 - 1024) / (C * cta2_threads(C) * 8); }
// shared memory: [stage-input buffer 0 | buffer 1 | aux (when a third vector fits) | NSM derivative vectors]
constexpr bool cta2_aux_global(int C) { return cta2_sm_slots(C) < 3; }
constexpr int cta2_nsm(int nkd, int C)
{
    return cta2_sm_slots(C) <= 3 ? 0
         : (nkd - cta2_nreg(nkd, C)) < (cta2_sm_slots(C) - 3) ? (nkd - cta2_nreg(nkd, C)) : (cta2_sm_slots(C) - 3);
}
constexpr int cta2_ngl(int nkd, int C) { return nkd - cta2_nreg(nkd, C) - cta2_nsm(nkd, C); }
// The staging vector of a right-hand-side evaluation whose K_J is register-resident is the vector of the first
// memory-resident K_j, j >= 1 (shared memory or slab; dead at that point); only when every derivative (from K_1 on) is
// in registers does the kernel need one more shared-memory vector.
constexpr int cta2_smem_vectors(int nkd, int C)
{
    return (cta2_aux_global(C) ? 2 : 3) + cta2_nsm(nkd, C) + ((cta2_nreg(nkd, C) > 1 ? cta2_nreg(nkd, C) : 1) >= nkd ? 1 : 0);
}
// per-CTA slab of global memory, in vectors: [max(NGL, 1) derivatives | one vector for event root searches | aux if not in smem]
constexpr int cta2_slab_vectors(int nkd, int C) { return (cta2_ngl(nkd, C) > 0 ? cta2_ngl(nkd, C) : 1) + 1 + (cta2_aux_global(C) ? 1 : 0); }
// threads per CTA: with C >= 2 always cta2_threads(C) (compile-time component stride), with C == 1 one per component
constexpr int cta2_block(unsigned long long n, int C) { return C >= 2 ? cta2_threads(C) : (int)((n + 31) / 32 * 32); }

}  // namespace cyrk
