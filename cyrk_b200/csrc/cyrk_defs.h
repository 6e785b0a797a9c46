// cyrk_defs.h -- constants and plain-old-data shared by host code, device code and
// NVRTC-compiled user kernels.  Must stay free of host-only includes (NVRTC).
//
// Values mirror the reference so that per-trajectory results are interchangeable:
//   status codes       CyRK/cy/c_common.hpp:25-82
//   step constants     CyRK/cy/c_common.hpp:189-196   (SAFETY, MIN_FACTOR, MAX_FACTOR, EPS_100)
//   method ints        CyRK/cy/cysolver.hpp:34-41     (RK23=3, RK45=4, DOP853=5)
//   brentq tolerances  CyRK/cy/cysolver.hpp:44-46     (4*eps, 4*eps, 100 iterations)
#pragma once

namespace cyrk {

enum Status : int {
    ST_CONVERGED = 20,
    ST_INITIALIZING = 10,
    ST_EVENT_TERMINATED = 2,
    ST_SUCCESSFUL_INTEGRATION = 1,
    ST_NO_ERROR = 0,
    ST_GENERAL_ERROR = -1,
    ST_PROPERTY_NOT_SET = -2,
    ST_UNSUPPORTED_UNKNOWN_MODEL = -3,
    ST_UNINITIALIZED_CLASS = -4,
    ST_ATTRIBUTE_ERROR = -7,
    ST_ARGUMENT_ERROR = -10,
    ST_DENSE_OUTPUT_NOT_SAVED = -21,
    ST_BAD_CONFIG_DATA = -22,
    ST_OPTIMIZE_SIGN_ERROR = -30,
    ST_OPTIMIZE_CONVERGENCE_ERROR = -31,
    ST_MEMORY_ALLOCATION_ERROR = -40,
    ST_MAX_STEPS_USER_EXCEEDED = -52,
    ST_MAX_STEPS_SYSARCH_EXCEEDED = -53,
    ST_STEP_SIZE_ERROR_SPACING = -54,
    ST_STEP_SIZE_ERROR_ACCEPTANCE = -55,
    ST_INTEGRATION_NOT_SUCCESSFUL = -57,
    ST_BAD_INITIAL_STEP_SIZE = -80,
    ST_UNSET_ERROR_CODE = -100
};

enum Method : int { METHOD_RK23 = 3, METHOD_RK45 = 4, METHOD_DOP853 = 5 };

// Kernel feature variants (compile-time): what one trajectory may produce besides its last row.
enum Variant : int {
    VAR_LEAN = 0,   // final row + counters only: no t_eval, no dense store, no events, no row store
    VAR_ROWS = 1,   // + row store (every accepted step, or the t_eval samples) + dense-interpolant store
    VAR_FULL = 2,   // + events
    VAR_STEPS = 3,  // VAR_LEAN + one record (t, y, signed step) per booked step: first half of the step-parallel
                    // t_eval / dense-output path (step_dense_kernel rebuilds the interpolants from these records)
    VAR_EVENTS = 4, // VAR_FULL without the row store, the t_eval emission and the dense store (store='final' + events):
                    // the same decisions from ~25 % less code -- the events kernels are instruction-fetch bound
    VAR_LEAN_SHARED = 5, // VAR_LEAN for a batch whose args are shared: the right-hand side reads them from the kernel
                    // parameter block (constant bank / uniform registers) instead of per-lane registers
    VAR_EVDETECT = 6, // events with store='final', first half of the two-pass events path: VAR_LEAN + per booked step the event
                    // functions and the trigger test; a step that triggers is pushed as a work item (t_old, t_now, step, y_old,
                    // hit slots) for event_roots_kernel, which redoes the step, builds its interpolant and runs the Brent
                    // searches -- one thread per triggered step instead of one lane of a stepping warp
    VAR_COUNT = 7
};
// one work item of the two-pass events path, in doubles: trajectory index, t_old, t_now, signed step, flags
// (bit 31: the step terminates the trajectory; low bits: mask of triggered events), y_old[NY], hit slot of every event (uint32 pairs)
#define CYRK_EVQ_ITEM_DOUBLES(ny) (5 + (ny) + CYRK_MAX_EVENTS / 2)
#define CYRK_SHARED_ARGS_MAX 8     // doubles

#define CYRK_MAX_NY 32        // thread-per-trajectory kernels: compile-time n_y up to this
#define CYRK_MAX_EVENTS 8
// one record of the VAR_STEPS solve: t, signed step, y[NY], padded to whole 16-byte pairs (128-bit stores / loads)
#define CYRK_STEP_REC_DOUBLES(ny) (((ny) + 3) & ~1)
#ifndef CYRK_STEP_DENSE_BLOCK
#define CYRK_STEP_DENSE_BLOCK 128   // threads per CTA of step_dense_kernel (its shared-memory staging is sized for it)
#endif

// One launch of the thread-per-trajectory solver.  Passed by value (constant bank).
struct BatchParams {
    // ---- batch inputs (device pointers)
    unsigned long long n_traj;
    const double* y0;            // [n_traj][NY]
    const char*   args;          // shared blob, or [n_traj][args_stride]
    unsigned      args_stride;   // bytes; 0 = shared
    double        shared_args[CYRK_SHARED_ARGS_MAX];   // VAR_LEAN_SHARED: the shared args themselves
    const double* t_span;        // [2] shared, or [n_traj][2]
    unsigned      t_span_stride; // 0 = shared, 2 = per trajectory
    double rtol[CYRK_MAX_NY];    // already floored at 100*eps on the host (rk.cpp:600-608)
    double atol[CYRK_MAX_NY];
    const double* rtol_pt;       // CYRK_TOL_PER_TRAJ kernels: [n_traj][NY] (floored like rtol[]); NULL otherwise
    const double* atol_pt;       //                            [n_traj][NY]
    double max_step;             // INF = none
    int    has_max_step;         // max_step < INF (lets the kernel test it with a uniform integer compare)
    double first_step;           // 0 = automatic
    unsigned max_num_steps;      // cysolver.cpp:403-410 (RAM cap folded in on the host)
    int    user_max_steps;       // which of -52 / -53 to report
    // ---- t_eval (shared by the batch; ascending; for backward runs the host passes the reversed copy)
    const double* t_eval;
    unsigned      n_t_eval;
    unsigned      t_eval_stride; // 0 = shared; n_t_eval = one row per trajectory (CYRK_FLAG_T_EVAL_PER_TRAJ)
    // ---- per-trajectory results
    int*      status;            // [n_traj]
    unsigned* size;              // rows the reference would have saved
    unsigned* steps_taken;       // accepted steps
    unsigned* n_rejected;        // rejected attempts
    unsigned* nfev;              // RHS evaluations (counted the reference's way)
    double*   t_last;            // [n_traj]        last saved row
    double*   y_last;            // [n_traj][NDY]
    // ---- row store (VAR_ROWS / VAR_FULL): layout follows CySolverResult::solution, row-major by step
    double*   t_rows;            // [n_traj][row_cap]
    double*   y_rows;            // [n_traj][row_cap][NDY]
    unsigned  row_cap;
    double*   step_recs;         // VAR_STEPS only: [n_traj][row_cap][CYRK_STEP_REC_DOUBLES(NY)] = (t, signed step, y[NY], pad)
    // ---- dense store (VAR_FULL, dense_output): interpolant i covers (dense_t[i-1], dense_t[i]]
    int       dense_output;
    double*   dense_t;           // [n_traj][dense_cap]      t_now of interpolant i (entry 0: t_start)
    double*   dense_state;       // [n_traj][dense_cap][NY*(Q_ORDER+1)]   [y_old | Q], dense.cpp:27-37
    unsigned* dense_count;       // [n_traj]
    unsigned  dense_cap;
    // ---- events (VAR_FULL)
    int       n_events;
    int       ev_id[CYRK_MAX_EVENTS];     // index into the problem's event table
    int       ev_dir[CYRK_MAX_EVENTS];
    unsigned long long ev_max[CYRK_MAX_EVENTS];   // ~0ull = never terminate
    double*   ev_t;              // [n_traj][n_events][ev_cap]
    double*   ev_y;              // [n_traj][n_events][ev_cap][NDY]
    unsigned* ev_count;          // [n_traj][n_events]
    int*      ev_term_index;     // [n_traj]   (-1 = none)
    unsigned  ev_cap;
    // ---- two-pass events (VAR_EVDETECT + event_roots_kernel)
    double*   evq_items;         // [evq_cap][CYRK_EVQ_ITEM_DOUBLES(NY)]
    unsigned long long* evq_count;   // [0] items pushed (may exceed evq_cap: the host re-runs with that capacity);
                                 // [1] != 0: a case only the one-pass kernel reproduces occurred (the host falls back to it)
    unsigned long long evq_cap;
    // ---- scheduling
    const double* prep;          // [n_traj][NDY+2] start-up state from rk_prep_kernel, or NULL (set up inline)
    unsigned long long* queue;   // global work counter (zeroed before launch)
    int       refill_min;        // refill a warp when at least this many lanes are idle (1..32)
    int       root_batch;        // run the event root search when this many lanes of a warp wait for it (1..32)
};

// Post-hoc dense evaluation (CySolverResult::call_vectorize, cysolution.cpp:511-616).
struct DenseEvalParams {
    unsigned long long n_traj;
    const double* t_query;       // [n_query] shared
    unsigned      n_query;
    const double* t_span;        // as above
    unsigned      t_span_stride;
    const int*      status;
    const double*   dense_t;
    const double*   dense_state;
    const unsigned* dense_count;
    unsigned        dense_cap;
    // interpolants kept as STEP RECORDS (two-pass path, see StepDenseParams): [n_traj][dense_cap][CYRK_STEP_REC_DOUBLES(NY)];
    // the queried step is recomputed from (t_old, y_old, step) -- 40 bytes per step for Lorenz instead of a 216-byte
    // [y_old | Q] record.  NULL = dense_t / dense_state hold materialised interpolants.
    const double*   recs;
    const unsigned* steps_taken;
    const char*   args;
    unsigned      args_stride;
    double*       out;           // [n_traj][n_query][NDY]
    int*          code;          // [n_traj]  first non-zero CySolverResult::call status, else 0
};

// Step-parallel interpolants (step_dense_kernel): one thread per (trajectory, booked step) recomputes the stages of
// that step from the VAR_STEPS records, builds its interpolant, stores it and emits the t_eval samples the step covers.
struct StepDenseParams {
    unsigned long long n_traj;
    unsigned      step_cap;      // records per trajectory (record 0 = the initial row)
    unsigned      step_threads;  // threads per trajectory: 1 + the largest steps_taken of the batch (<= step_cap)
    const unsigned* steps_taken; // [n_traj]
    const double* recs;          // [n_traj][step_cap][CYRK_STEP_REC_DOUBLES(NY)]  (t, signed step, y[NY], pad)
    int           rows_are_steps;// no t_eval and the caller wants rows: t_rows / y_rows receive the step records (row s = step s)
    const char*   args;
    unsigned      args_stride;
    const double* t_span;        // shared or per trajectory (direction only)
    unsigned      t_span_stride;
    const double* t_eval;        // ascending (reversed copy for backward runs), shared
    unsigned      n_t_eval;
    double*       t_rows;        // [n_traj][row_cap]       t_eval samples (row j = the j-th sample the run passes)
    double*       y_rows;        // [n_traj][row_cap][NY]
    unsigned      row_cap;
    int           dense_output;  // count interpolants / their right-hand-side evaluations like the reference does
    double*       dense_t;       // as in BatchParams; NULL = do not materialise the interpolants: the step records
    double*       dense_state;   // themselves are the dense store and sol(t) recomputes the step it needs
    unsigned*     dense_count;
    unsigned      dense_cap;
    unsigned*     nfev;          // += the interpolant's extra right-hand-side evaluations (DOP853), counted the reference's way
};

}  // namespace cyrk
