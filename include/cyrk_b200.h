/* cyrk_b200.h -- C ABI of the B200-native batched IVP solver (libcyrk_b200.so).
 *
 * Drop-in boundary for ONE path of jrenaud90/CyRK: the adaptive explicit Runge-Kutta step loop
 * behind `cysolve_ivp` / `pysolve_ivp`.  A loop of reference calls
 *
 *     for i in prange(B):  cysolve_ivp_noreturn(results[i], diffeq, t0, t1, y0[i], ..., args[i])
 *                                              (CyRK/cy/prange_test.pyx:87-97)
 *
 * becomes one `cyrk_b200_solve_batch` call.  The shape of this API follows the reference's own
 * flat C interface, CyRK/nb/numba_c_api_.cpp:
 *     numba_cysolve_ivp(...) -> void*          :10-65    ->  cyrk_b200_solve_batch(...) -> void*
 *     cysolver_get_{status,size,...}(void*)    :91-213   ->  cyrk_b200_get_*(void*) (per-trajectory arrays)
 *     cysolver_call_call_vectorize             (:~190)   ->  cyrk_b200_call_vectorize
 *     cysolver_get_status_message_buffer       :70-86    ->  cyrk_b200_status_message
 *     cysolver_free(void*)                     :218-222  ->  cyrk_b200_free
 * and the argument list is that of cysolve_ivp (CyRK/cy/cysolver_api.pyx:416-469) with the
 * function-pointer callbacks replaced by ids of device functors (built-in, or compiled from CUDA
 * source with NVRTC), because a host function pointer cannot run on the GPU.
 *
 * Plain C types only; no torch types.  All entry points are callable without the Python GIL and
 * are thread-safe for distinct handles.  Per-trajectory status codes are the reference's
 * CyrkErrorCodes verbatim (CyRK/cy/c_common.hpp:25-82).  Call-level errors (CUDA, NVRTC, bad
 * arguments) are reported through cyrk_b200_last_error().
 *
 * There is NO CPU fallback: without a CUDA device every solve entry fails loudly.
 */
#ifndef CYRK_B200_H
#define CYRK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The library is built with -fvisibility=hidden; only this API is exported. */
#pragma GCC visibility push(default)

/* ODEMethod ints, CyRK/cy/cysolver.hpp:34-41 */
#define CYRK_METHOD_RK23   3
#define CYRK_METHOD_RK45   4
#define CYRK_METHOD_DOP853 5

/* Built-in right-hand sides: 0-10 are cytester's diffeq_number (cysolver_test.pyx:419-445). */
#define CYRK_RHS_BASELINE          0
#define CYRK_RHS_ACCURACY          1
#define CYRK_RHS_EXTRAOUTPUT       2
#define CYRK_RHS_LORENZ            3
#define CYRK_RHS_LORENZ_EXTRA      4
#define CYRK_RHS_LOTKAVOLTERRA     5
#define CYRK_RHS_PENDULUM          6
#define CYRK_RHS_ARBITRARY_ARGS    7
#define CYRK_RHS_PENDULUM_PREEVAL  8
#define CYRK_RHS_LARGEY            9   /* CTA-per-trajectory path, any n_y (events / dense output: n_y <= 12288, DOP853: <= 4096) */
#define CYRK_RHS_LARGEY_SIMPLE     10  /* CTA-per-trajectory path, any n_y */
#define CYRK_RHS_PENDULUM_DRAG     11
#define CYRK_RHS_DENSE_KAT         12
#define CYRK_RHS_USER_BASE         1000 /* ids returned by cyrk_b200_rhs_compile start here */

/* Built-in event functions (valid with any right-hand side that has n_y >= 2). */
#define CYRK_EV_Y0_GT_50            0
#define CYRK_EV_Y1_LT_50            1
#define CYRK_EV_T_GT_5              2
#define CYRK_EV_LORENZ_T_GE_5       3
#define CYRK_EV_LORENZ_Y_RANGE      4
#define CYRK_EV_LORENZ_ARGS         5
#define CYRK_EV_Y0                  6
#define CYRK_EV_Y1                  7
#define CYRK_EV_T_MINUS_ARG4        8
#define CYRK_EV_ABS_Y0_MINUS_ARG4   9

/* What to keep per trajectory. */
#define CYRK_STORE_FINAL 0   /* last saved row + counters only                      */
#define CYRK_STORE_ROWS  1   /* every row the reference would save (steps, or t_eval samples) */

/* Flags. */
#define CYRK_FLAG_STRICT_FP   1u  /* kernels compiled with -fmad=false: reference's rounding (no FMA) */
#define CYRK_FLAG_DEVICE_IO   2u  /* y0/args/t_span are DEVICE pointers; results stay on the device   */
#define CYRK_FLAG_T_EVAL_PER_TRAJ 4u /* t_eval holds n_traj rows of n_t_eval sample times (one row per trajectory) */
#define CYRK_FLAG_TOL_PER_TRAJ    8u /* rtols / atols hold n_traj rows of n_y entries (n_rtols == n_atols == n_traj * n_y).
                                        Thread-per-trajectory path: the kernels of this option are compiled on first use;
                                        CTA-per-trajectory path (both per-trajectory flags): raw step rows / t_eval / final
                                        state, not together with events or stored dense output */

#define CYRK_EVENT_NO_LIMIT ((size_t)-1)   /* Event::max_allowed default, c_events.hpp:33 */

/* ---- library / device ---------------------------------------------------------------------- */
int         cyrk_b200_version(void);
int         cyrk_b200_device_count(void);                 /* 0 without a usable CUDA device */
const char* cyrk_b200_last_error(void);                   /* thread-local, never NULL */
void        cyrk_b200_status_message(int status_code, char* buffer, size_t max_len); /* numba_c_api_.cpp:70-86 */

/* ---- right-hand sides ---------------------------------------------------------------------- */
/* Look up a built-in by name ("lotkavolterra", "lorenz", ...): returns its id or -1. */
int cyrk_b200_rhs_builtin(const char* name);
/* Sizes of a right-hand side (built-in or compiled): returns 0, or -1 for an unknown id.
 * args_bytes is the size of the per-trajectory args blob the functor reads (shorter blobs are
 * zero-padded); n_y == 0 means "any n_y" (CTA-per-trajectory problems). */
int cyrk_b200_rhs_info(int rhs_id, size_t* n_y, size_t* n_extra, size_t* args_bytes);
/* Compile a user right-hand side (and optional events) from CUDA C++ source with NVRTC for
 * sm_100a.  `source` must define
 *     __device__ void  <entry>(double* dy, double t, const double* y, const char* args);
 * and, when n_event_funcs > 0,
 *     __device__ double <entry>_event(int id, double t, const double* y_and_extras, const char* args);
 * mirroring DiffeqFuncType / EventFunc (c_common.hpp:218-220, c_events.hpp:9); a pre-eval hook is
 * just another __device__ function the source calls.  LARGE systems (n_y == 0 here; the size comes with each solve
 * call; CTA-per-trajectory path) are written component-wise instead:
 *     __device__ double <entry>(unsigned i, double t, const double* y, unsigned n, const char* args);      -> dy_i
 *         (i >= n: extra output i - n, CyRK's dy[n_y:]; asked for at saved rows only; extra outputs on this path
 *          come with raw step rows / the final state, not with t_eval, events or dense_output)
 * A whole-vector <entry>(dy, t, y, args) with n_y > 32 is accepted too (same signature as for small systems; no extras /
 * events): it runs on the CTA-per-trajectory path with ONE thread evaluating the function per stage -- correct, and fine
 * for cheap right-hand sides; the component-wise form is the one that scales (n_y = 64 / 128, DOP853: 12x faster; it
 * also beats the thread-per-trajectory kernel for DOP853 at n_y = 32, whose 16 x n_y stage derivatives spill: 1.9x).
 *     __device__ double <entry>_event(int id, double t, const double* y, unsigned n, const char* args);     -> g_id
 * Returns an id >= CYRK_RHS_USER_BASE, or -1
 * (compile log in `log`).  Compiled modules are cached by source hash per process and on disk (see below). */
int cyrk_b200_rhs_compile(const char* source, const char* entry, size_t n_y, size_t n_extra,
                          size_t args_bytes, size_t n_event_funcs, char* log, size_t log_len);
/* Run-time compiled kernels (user right-hand sides; feature builds of the built-ins such as CYRK_FLAG_TOL_PER_TRAJ) are
 * cached per process AND on disk: $CYRK_B200_CACHE_DIR, else $XDG_CACHE_HOME/cyrk_b200, else ~/.cache/cyrk_b200
 * (CYRK_B200_CACHE=0 disables); the key hashes the complete source, every embedded solver header, the options and the NVRTC
 * version.  cyrk_b200_jit_compile compiles one solve kernel ahead of time (variant: 0 final state, 1 rows / t_eval / dense,
 * 2 + events, 4 events with final state only; flags: CYRK_FLAG_STRICT_FP / CYRK_FLAG_TOL_PER_TRAJ) -- it needs NVRTC but
 * no GPU, so a build box can warm the cache.  Returns 0 / -1.  jit_cache_stats: kernels compiled / found on disk so far. */
int  cyrk_b200_jit_compile(int rhs_id, int method, int variant, unsigned flags);
void cyrk_b200_jit_cache_stats(unsigned long long* compiled, unsigned long long* disk_hits);

/* ---- solve --------------------------------------------------------------------------------- */
/* Arguments mirror numba_cysolve_ivp (numba_c_api_.cpp:10-30), batched:
 *   rhs_id, method            device functor id, CYRK_METHOD_*
 *   n_traj                    number of independent initial-value problems
 *   t_span, n_t_span          (t_start, t_end) pairs: n_t_span == 1 shared, == n_traj per trajectory
 *   y0, n_y                   [n_traj][n_y] row-major
 *   num_extra                 extra outputs captured from dy[n_y:] (must match the functor)
 *   args, args_len, n_args    opaque blobs, args_len bytes each: n_args == 1 shared (or 0: none),
 *                             == n_traj per trajectory
 *   max_num_steps, max_ram_MB as the reference (cy/c_common.cpp:20-56); 0 / 2000 are its defaults
 *   dense_output              keep every step's interpolant on the device for call_vectorize
 *   t_eval, n_t_eval          sample times (ordered in the direction of integration), always a HOST pointer: n_t_eval shared
 *                             values, or with CYRK_FLAG_T_EVAL_PER_TRAJ [n_traj][n_t_eval]
 *   event_ids/dirs/max        n_events entries; max == CYRK_EVENT_NO_LIMIT never terminates
 *   rtols/atols               1 or n_y entries (HOST pointers), or with CYRK_FLAG_TOL_PER_TRAJ [n_traj][n_y] each
 *   max_step, first_step      INFINITY / 0.0 are the reference defaults
 *   expected_size             rows (and interpolants) reserved per trajectory in CYRK_STORE_ROWS /
 *                             dense mode; 0 = automatic.  If a trajectory needs more, the batch is
 *                             re-run once with the exact capacity (results are deterministic).
 *   store_mode, flags         CYRK_STORE_*, CYRK_FLAG_*
 *   device                    CUDA device ordinal (one process per GPU drives its own device)
 * Returns an opaque handle (never reuse after cyrk_b200_free), or NULL with cyrk_b200_last_error().
 */
void* cyrk_b200_solve_batch(
    int rhs_id, int method, size_t n_traj,
    const double* t_span, size_t n_t_span,
    const double* y0, size_t n_y,
    size_t num_extra,
    const char* args, size_t args_len, size_t n_args,
    size_t max_num_steps, size_t max_ram_MB,
    int dense_output,
    const double* t_eval, size_t n_t_eval,
    const int* event_ids, const int* event_dirs, const size_t* event_max, size_t n_events,
    const double* rtols, size_t n_rtols,
    const double* atols, size_t n_atols,
    double max_step, double first_step,
    size_t expected_size,
    int store_mode, unsigned flags, int device);

/* The same call over several GPUs of ONE process (Documentation/Parallelization.md: trajectories are independent):
 * contiguous shards, one host thread and one stream per device, no communication.  Shard r owns the trajectories
 * [bounds_out[r], bounds_out[r+1]) (the first n_traj % n_dev shards get one extra) and comes back as its own handle in
 * handles_out[r]; per-trajectory inputs (y0, per-trajectory t_span / args) are sliced accordingly.  Host buffers only.
 * Returns 0, or -1 (no handle is left allocated).  The one-process-per-GPU alternative is bench.py / torchrun. */
int cyrk_b200_solve_batch_sharded(
    const int* device_ids, size_t n_dev, void** handles_out, size_t* bounds_out /* [n_dev + 1] or NULL */,
    int rhs_id, int method, size_t n_traj,
    const double* t_span, size_t n_t_span,
    const double* y0, size_t n_y,
    size_t num_extra,
    const char* args, size_t args_len, size_t n_args,
    size_t max_num_steps, size_t max_ram_MB,
    int dense_output,
    const double* t_eval, size_t n_t_eval,
    const int* event_ids, const int* event_dirs, const size_t* event_max, size_t n_events,
    const double* rtols, size_t n_rtols,
    const double* atols, size_t n_atols,
    double max_step, double first_step,
    size_t expected_size,
    int store_mode, unsigned flags);

/* Same, re-using the device arena / pinned host buffers of an existing handle
 * (cysolve_ivp_noreturn / solution_reuse, CyRK/cy/cysolve.cpp:6-59).  Returns 0 or -1. */
int cyrk_b200_solve_batch_noreturn(
    void* handle,
    int rhs_id, int method, size_t n_traj,
    const double* t_span, size_t n_t_span,
    const double* y0, size_t n_y,
    size_t num_extra,
    const char* args, size_t args_len, size_t n_args,
    size_t max_num_steps, size_t max_ram_MB,
    int dense_output,
    const double* t_eval, size_t n_t_eval,
    const int* event_ids, const int* event_dirs, const size_t* event_max, size_t n_events,
    const double* rtols, size_t n_rtols,
    const double* atols, size_t n_atols,
    double max_step, double first_step,
    size_t expected_size,
    int store_mode, unsigned flags, int device);

/* Empty handle for the _noreturn entry. */
void* cyrk_b200_result_create(void);
void  cyrk_b200_free(void* handle);                                   /* cysolver_free */

/* ---- getters (host pointers valid until the next solve on the handle or cyrk_b200_free;
 *      with CYRK_FLAG_DEVICE_IO they are DEVICE pointers).  Pull-based like the reference's cysolver_get_*
 *      (numba_c_api_.cpp:91-213): a solve copies back status, steps_taken and y_last (8 + 8*num_dy bytes per trajectory,
 *      plus whatever rows / events were asked for); size, n_rejected, nfev and t_last stay in the device arena until
 *      their getter is first called, which then performs one device-to-host copy (NULL + cyrk_b200_last_error() if
 *      that copy fails). ------------------------------------------------------------------------ */
size_t          cyrk_b200_get_n_traj(void* h);
size_t          cyrk_b200_get_num_y(void* h);
size_t          cyrk_b200_get_num_dy(void* h);
const int32_t*  cyrk_b200_get_status_ptr(void* h);        /* [n_traj] CyrkErrorCodes */
const uint32_t* cyrk_b200_get_size_ptr(void* h);          /* [n_traj] rows saved (reference `size`) */
const uint32_t* cyrk_b200_get_steps_taken_ptr(void* h);   /* [n_traj] accepted steps */
const uint32_t* cyrk_b200_get_n_rejected_ptr(void* h);    /* [n_traj] rejected attempts */
const uint32_t* cyrk_b200_get_nfev_ptr(void* h);          /* [n_traj] RHS evaluations */
const double*   cyrk_b200_get_t_last_ptr(void* h);        /* [n_traj] t of the last saved row */
const double*   cyrk_b200_get_y_last_ptr(void* h);        /* [n_traj][num_dy] last saved row */
size_t          cyrk_b200_get_row_capacity(void* h);      /* rows reserved per trajectory (0 in STORE_FINAL) */
const double*   cyrk_b200_get_t_ptr(void* h);             /* [n_traj][row_capacity] */
const double*   cyrk_b200_get_y_ptr(void* h);             /* [n_traj][row_capacity][num_dy] (row-major by step,
                                                             like CySolverResult::solution) */
size_t          cyrk_b200_get_num_events(void* h);
/* Event hits.  The reference keeps one ragged list per event (CySolverResult::event_times / event_states,
 * cysolution.cpp:423-448).  Two layouts:
 *   packed (host-buffer solves on the thread-per-trajectory path; event_capacity == 0): the hits of list
 *     l = traj * n_events + e are event_t[offsets[l] .. offsets[l+1]) and event_y[...][num_dy] -- lists packed back to
 *     back on the device before they cross PCIe; offsets has n_traj * n_events + 1 entries, event_total hits in all;
 *   padded (CYRK_FLAG_DEVICE_IO and the CTA-per-trajectory path; event_capacity > 0, offsets == NULL): every list owns
 *     event_capacity slots of which event_count[traj][e] are filled. */
size_t          cyrk_b200_get_event_capacity(void* h);
const uint32_t* cyrk_b200_get_event_count_ptr(void* h);   /* [n_traj][n_events] */
const uint64_t* cyrk_b200_get_event_offsets_ptr(void* h); /* [n_traj * n_events + 1], packed layout only */
size_t          cyrk_b200_get_event_total(void* h);       /* hits in the packed arrays */
const double*   cyrk_b200_get_event_t_ptr(void* h);       /* packed: [event_total]; padded: [n_traj][n_events][event_capacity] */
const double*   cyrk_b200_get_event_y_ptr(void* h);       /* packed: [event_total][num_dy]; padded: [...][event_capacity][num_dy] */
const int32_t*  cyrk_b200_get_event_terminate_index_ptr(void* h); /* [n_traj], -1 = not terminated */
size_t          cyrk_b200_get_num_interpolates_capacity(void* h);
const uint32_t* cyrk_b200_get_num_interpolates_ptr(void* h);      /* [n_traj] (dense_output only) */

/* sol(t) for every trajectory: out[n_traj][n_t][num_dy]; codes[n_traj] receives the reference's
 * CySolverResult::call status (0 ok, -57 not successful, -21 no dense output, -10 out of range).
 * `out`/`codes` are host pointers (device pointers with CYRK_FLAG_DEVICE_IO).  Returns 0 or -1. */
int cyrk_b200_call_vectorize(void* h, const double* t, size_t n_t, double* out, int32_t* codes);

/* Ensemble sums over the successful trajectories (status 1 or 2) of the last solve, computed on the device with a
 * fixed-order two-stage reduction: out[0..num_dy) = sum of y_last, out[num_dy..2*num_dy) = sum of y_last^2,
 * out[2*num_dy] = number of trajectories counted.  This is the local half of the optional ensemble reduction; the
 * caller all-reduces the 2*num_dy+1 doubles across ranks (NCCL / torch.distributed) -- the only collective this
 * path ever needs.  `out` is a host pointer (device pointer with CYRK_FLAG_DEVICE_IO).  Returns 0 or -1. */
int cyrk_b200_ensemble_sums(void* h, double* out);

/* ---- measurement hooks (bench.py) ----------------------------------------------------------- */
double   cyrk_b200_get_kernel_ms(void* h);      /* CUDA-event time of the solve kernel(s) of the last call */
double   cyrk_b200_get_total_ms(void* h);       /* CUDA-event time H2D -> kernels -> D2H of the last call   */
unsigned cyrk_b200_get_launch_count(void* h);   /* kernels launched by the last call                        */
size_t   cyrk_b200_get_h2d_bytes(void* h);
size_t   cyrk_b200_get_d2h_bytes(void* h);      /* copied inside the solve call                             */
size_t   cyrk_b200_get_lazy_d2h_bytes(void* h); /* copied afterwards by getters (see above)                 */
/* Tuning knobs (process-wide): "refill_min" (1..32), "block" (threads per CTA of the thread-per-trajectory kernels,
 * 0 = by occupancy), "ctas_per_sm" (0 = occupancy API), "prep" (start-up kernel: -1 auto / 0 / 1), "root_batch" (1..32),
 * "cta_kernel" (0 = register-resident CTA kernel where it applies, 1 = shared-memory/slab kernel), "cta_block",
 * "cta_smem_vecs" (slab kernel only), "cta2_large" (1 = register-resident kernel also for 4096 < n_y <= 12288), "pipe_chunk" (largest body chunk of the pipelined host-I/O path, trajectories), "pipe_sched" (1 = head / body / tail schedule, 0 = equal chunks),
 * "pipe_taper" (short first/last chunks, 0/1), "pipe_min" (smallest host-buffer final-state batch that takes the pipelined path;
 * default 256 Ki, 0 = never), "two_pass" (dense_output / t_eval without events or extras: 1 = step-parallel
 * interpolants when dense_output is set [default], 2 = also for t_eval alone, 0 = always the per-trajectory kernel; the
 * results are bitwise identical either way), "events_two_pass" (events with store='final': 1 [default] = detection solve + one thread per triggered step, 0 = roots found by
 * the stepping lanes; same results), "dense_store" (step-parallel path with dense_output: 0 [default] = the step records
 * are the dense store and sol(t) recomputes the step it needs, 1 = every interpolant is materialised as well), "shared_args" (final-state solves whose args are shared read them from the
 * kernel parameter block instead of per-lane registers: 1 [default] / 0; bitwise identical with strict FP, within FMA-contraction
 * noise in the production build). */
int      cyrk_b200_set_option(const char* name, int value);
/* FP64 FMA throughput probe (dependent-free DFMA streams on every SM): returns TFLOP/s. */
double   cyrk_b200_measure_fp64_tflops(int device);
/* Diagnostic: the production build's sin/cos for the built-in right-hand sides, evaluated on the device for n host
 * values (accuracy test against libm).  Returns 0 / -1. */
int      cyrk_b200_debug_sincos(const double* x, double* sin_out, double* cos_out, size_t n, int device);

#pragma GCC visibility pop

#ifdef __cplusplus
}
#endif
#endif /* CYRK_B200_H */
