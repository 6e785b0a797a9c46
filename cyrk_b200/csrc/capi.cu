// capi.cu -- the extern "C" boundary (include/cyrk_b200.h): argument validation, device arena,
// host<->device transfers, kernel dispatch, result handle.  No solver arithmetic lives here.
//
// Reference counterparts: CyRK/nb/numba_c_api_.cpp (flat API), CyRK/cy/cysolve.cpp:6-114
// (config update -> setup -> solve), CyRK/cy/c_common.cpp:20-56 (max_num_steps from the RAM cap),
// CyRK/cy/rk.cpp:578-608 (tolerance validation + rtol floor).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cyrk_b200.h"
#include "cyrk_defs.h"
#include "registry.h"
#include "rk_core.cuh"
#include "problems_builtin.cuh"
#include "cta_solver.h"
#include "nvrtc_rhs.h"

namespace cyrk {

// ------------------------------------------------------------------------------------------------
// error reporting
// ------------------------------------------------------------------------------------------------
thread_local std::string g_last_error;

void set_error(const char* fmt, ...)
{
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

#define CU_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                      cudaGetErrorString(e__));                                              \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

// ------------------------------------------------------------------------------------------------
// registry
// ------------------------------------------------------------------------------------------------
static ProblemEntry g_table[FP_COUNT][RHS_TABLE_SIZE];
ProblemEntry* registry_slot(int fp, int id) { return &g_table[fp][id]; }
const ProblemEntry* registry_lookup(int fp, int id)
{
    if (fp < 0 || fp >= FP_COUNT || id < 0 || id >= RHS_TABLE_SIZE) return nullptr;
    return g_table[fp][id].present ? &g_table[fp][id] : nullptr;
}

#define CYRK_DECL_REG(P) void cyrk_register_##P##_fast(); void cyrk_register_##P##_strict();
#define CYRK_CALL_REG(P) cyrk_register_##P##_fast(); cyrk_register_##P##_strict();
#define CYRK_ALL_PROBLEMS(X)                                                                       \
    X(ProbBaseline) X(ProbAccuracy) X(ProbExtraOutput) X(ProbLorenz) X(ProbLorenzExtra)           \
    X(ProbLotkaVolterra) X(ProbPendulum) X(ProbArbitraryArgs) X(ProbPendulumPreEval)              \
    X(ProbPendulumDrag) X(ProbDenseKat)
CYRK_ALL_PROBLEMS(CYRK_DECL_REG)

void registry_init()
{
    static std::once_flag once;
    std::call_once(once, [] { CYRK_ALL_PROBLEMS(CYRK_CALL_REG) });
}

// ------------------------------------------------------------------------------------------------
// options
// ------------------------------------------------------------------------------------------------
// Measured on B200: with the start-up state precomputed by rk_prep_kernel a refill is a handful of loads, so lanes
// refill immediately (refill_min 1).  Round 1 (profiles/r01_notes.md) found that to win from ~2 Mi trajectories up; with
// the round-2 kernels (shared args in the parameter block) it wins from 256 Ki: 0.773 vs 0.785 ms there, 1.34 vs 1.42 at
// 512 Ki, 2.46 vs 2.65 at 1 Mi, 4.71 vs 5.10 at 2 Mi (profiles/r02_c1_prep_by_batch.log).  Below that the extra launch
// costs what it saves and inline set-up with refill_min 4 stays.
static int g_opt_refill_min = 0;      // 0 = automatic (1 with prep, 4 without)
static int g_opt_prep = -1;           // -1 = automatic (n_traj >= 256 Ki), 1 = always, 0 = never
static int g_opt_block = 0;            // threads per CTA of the thread-per-trajectory kernels; 0 = pick by occupancy (128 or 96)
static int g_opt_cta2_per_sm = 0;     // register-resident CTA kernel: cap on CTAs per SM (0 = the hardware's 32; registers allow 16
                                      // 32-thread CTAs: n_y = 32, DOP853: 3.8 vs 6.4 ms with the former cap of 8, profiles/r02_medium_ny_component_wise.log)
static int g_opt_small_cta = 0;       // 1: DOP853 user systems with 16 < n_y <= 32 run a warp per trajectory (cta_core.cuh) instead of a thread (measured 4x slower)
static int g_opt_cta2_large = 1;      // register-resident CTA kernel for 4096 < n_y <= 12 288 (derivatives in the L2 slab); 0: cta_core.cuh
static int g_opt_cta_smem_vecs = -1;  // CTA path: working vectors kept in shared memory (-1 = as many as fit)
static int g_opt_cta_block = 0;       // CTA path: threads per CTA (0 = by n_y)
static int g_opt_pipe_taper = 1;       // pipelined host-I/O path: short first / last chunks
static int g_opt_pipe_chunk = 3 << 20; // pipelined host-I/O path: largest body chunk (trajectories).  Larger chunks run the persistent grid better,
                                       // smaller ones overlap more of the copies: 8 Mi LV trajectories end to end 20.02 ms (1 Mi) / 19.35 (2 Mi) /
                                       // 19.20 (3 - 4 Mi) (profiles/r02_e2e_pipe_chunk.log, r02_s9_e2e_sched.log)
static int g_opt_pipe_sched = 1;       // pipelined host-I/O path: 1 = head / body / tail schedule, 0 = equal chunks with tapered ends (round 2, first half)
static int g_opt_pipe_min = 1 << 18;   // pipelined host-I/O path: smallest batch that takes it (0 = never)
static int g_opt_cta_kernel = 0;      // CTA path: 0 = pick (register-resident kernel where it is faster), 1 = always the
                                      // shared-memory/slab kernel, 2 = the register-resident kernel whenever n_y <= 4096
static int g_opt_shared_args = 1;     // shared args through the kernel parameter block (VAR_LEAN_SHARED); 0 = per-lane registers
static int g_opt_two_pass = 1;        // t_eval / dense_output without events or extras: step-parallel interpolants (0 = the one-pass kernel)
static int g_opt_dense_store = 0;     // two-pass path with dense_output: 0 = the 40-byte step records are the dense store (sol(t)
                                      // recomputes the step it needs), 1 = materialise every [y_old | Q] interpolant as well
static int g_opt_events_variant = 1;  // events with store='final' and no t_eval / dense store: the VAR_EVENTS kernels (0 = VAR_FULL)
static int g_opt_events_two_pass = 1; // ... as detection solve + one thread per triggered step (event_roots_kernel); 0 = one-pass VAR_EVENTS
static int g_opt_events_static = 1;   // ... with the detection solve compiled for this run's event list (NVRTC, cached); 0 = generic kernel
static int g_opt_root_batch = 12;     // measured: 8..16 is the flat optimum on the C3 workload (profiles/r01_notes.md)
static int g_opt_ctas_per_sm = 0;

// ------------------------------------------------------------------------------------------------
// grow-only buffers
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            size_t fr = 0, tot = 0;
            cudaMemGetInfo(&fr, &tot);
            set_error("device allocation of %zu bytes failed (%s); %zu of %zu bytes free -- split the batch or "
                      "lower expected_size / t_eval / dense_output storage", bytes, cudaGetErrorString(e), fr, tot);
            cudaGetLastError();
            return -1;
        }
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return static_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) {
            set_error("pinned host allocation of %zu bytes failed (%s)", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return -1;
        }
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return static_cast<T*>(p); }
};

// ------------------------------------------------------------------------------------------------
// result handle
// ------------------------------------------------------------------------------------------------
struct Handle {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_comp2 = nullptr;      // pipelined host-I/O path
    std::vector<cudaEvent_t> ev_pool;                                        // per-chunk events, grown on demand
    cudaEvent_t ev_begin = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_end = nullptr;
    // problem of the last solve
    int rhs_id = -1, method = 0, fp_mode = 0;
    size_t n_traj = 0, ny = 0, nx = 0, ndy = 0, q_order = 0;
    size_t n_events = 0, n_t_eval = 0;
    size_t row_cap = 0, dense_cap = 0, ev_cap = 0;
    bool dense_output = false, device_io = false, solved = false, cta_path = false;
    bool dense_from_steps = false;         // sol(t) rebuilds interpolants from the step records in d_st (two-pass path)
    size_t args_stride = 0;
    // device arena
    DevBuf d_y0, d_args, d_tspan, d_teval, d_queue;
    DevBuf d_status, d_size, d_steps, d_nrej, d_nfev, d_tlast, d_ylast;
    DevBuf d_trows, d_yrows, d_dense_t, d_dense_state, d_dense_count;
    DevBuf d_ev_t, d_ev_y, d_ev_count, d_ev_term;
    DevBuf d_q_t, d_q_out, d_q_code, d_scratch, d_prep, d_rtol, d_atol;
    DevBuf d_st;                            // step records of the two-pass t_eval / dense path
    DevBuf d_ev_t2, d_ev_y2, d_ev_off, d_scan;   // packed (CSR) event lists: hits back to back + offsets
    DevBuf d_evq, d_evq_count;              // work items of the two-pass events path
    DevBuf d_max;                           // [3] largest size / dense_count / ev_count of the last launch (max_counts_kernel)
    size_t evq_seen = 0;                    // items the last two-pass events solve pushed (sizes the next queue of a reused handle)
    PinBuf h_ev_off;
    bool ev_packed = false;                 // host mirrors of the event lists are CSR-packed (host-buffer solves on the thread path)
    size_t ev_total = 0;                    // hits in the packed arrays
    // pinned host mirrors
    PinBuf h_status, h_size, h_steps, h_nrej, h_nfev, h_tlast, h_ylast;
    PinBuf h_trows, h_yrows, h_dense_count, h_ev_t, h_ev_y, h_ev_count, h_ev_term;
    PinBuf h_stage;
    // device-pointer inputs in DEVICE_IO mode
    const double* in_y0 = nullptr;
    const char* in_args = nullptr;
    const double* in_tspan = nullptr;
    size_t tspan_stride = 0;
    // stats
    float kernel_ms = 0.f, total_ms = 0.f;
    unsigned launches = 0;
    size_t h2d_bytes = 0, d2h_bytes = 0;
    // Per-trajectory result arrays whose host mirror is current.  A solve copies back what every caller reads (status,
    // steps_taken, y_last; `size` where it decides about a re-run); the others stay in the device arena until their getter
    // is called (cyrk_b200_get_*_ptr are pull-based like the reference's cysolver_get_*, numba_c_api_.cpp:91-213).
    unsigned host_valid = 0;
    size_t lazy_d2h_bytes = 0;
    // ---- streamed (chunked) solves: a batch whose rows / event lists / step records are produced chunk by chunk.  The
    // parent handle owns the whole-batch inputs, per-trajectory results, step records and pinned host arrays; two
    // sub-handles own one chunk-sized arena each (rows, event slots) and alternate, so chunk k+1 solves while chunk k's
    // rows cross PCIe.  `force` is set on a sub-handle while it solves a chunk.
    Handle* sub[2] = {nullptr, nullptr};
    struct Force {
        bool active = false;
        size_t row_cap = 0;                // rows per trajectory of the parent's host layout (0 = not fixed yet: probe chunk)
        size_t scap = 0;                   // step records per trajectory of the parent's record store (0 = not fixed yet)
        double* recs = nullptr;            // where this chunk's step records go (parent's store); NULL = the sub's own buffer
        unsigned long long ev_base = 0;    // offset of this chunk's first hit in the parent's packed event arrays
        size_t need_row_cap = 0, need_scap = 0;    // out: capacity this chunk would have needed (set when it returns 1)
    } force;
    bool took_two_pass = false;            // out: the last solve ran the step-parallel path
    size_t step_cap = 0;                   // out: step records per trajectory of that path

    int bind(int dev)
    {
        if (device == dev && stream) { CU_TRY(cudaSetDevice(dev)); return 0; }
        if (stream) destroy_cuda();
        CU_TRY(cudaSetDevice(dev));
        CU_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&s_comp2, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreate(&ev_begin));
        CU_TRY(cudaEventCreate(&ev_k0));
        CU_TRY(cudaEventCreate(&ev_k1));
        CU_TRY(cudaEventCreate(&ev_end));
        device = dev;
        return 0;
    }
    void destroy_cuda()
    {
        int prev_dev = -1;
        if (cudaGetDevice(&prev_dev) != cudaSuccess) { cudaGetLastError(); prev_dev = -1; }
        if (device >= 0) cudaSetDevice(device);
        DevBuf* dbs[] = {&d_y0, &d_args, &d_tspan, &d_teval, &d_queue, &d_status, &d_size, &d_steps, &d_nrej, &d_nfev,
                         &d_tlast, &d_ylast, &d_trows, &d_yrows, &d_dense_t, &d_dense_state, &d_dense_count, &d_ev_t,
                         &d_ev_y, &d_ev_count, &d_ev_term, &d_q_t, &d_q_out, &d_q_code, &d_scratch, &d_prep, &d_rtol, &d_atol,
                         &d_st, &d_ev_t2, &d_ev_y2, &d_ev_off, &d_scan, &d_evq, &d_evq_count, &d_max};
        for (DevBuf* b : dbs) b->release();
        PinBuf* pbs[] = {&h_status, &h_size, &h_steps, &h_nrej, &h_nfev, &h_tlast, &h_ylast, &h_trows, &h_yrows,
                         &h_dense_count, &h_ev_t, &h_ev_y, &h_ev_count, &h_ev_term, &h_stage, &h_ev_off};
        for (PinBuf* b : pbs) b->release();
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_k0) cudaEventDestroy(ev_k0);
        if (ev_k1) cudaEventDestroy(ev_k1);
        if (ev_end) cudaEventDestroy(ev_end);
        if (stream) cudaStreamDestroy(stream);
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_d2h) cudaStreamDestroy(s_d2h);
        if (s_comp2) cudaStreamDestroy(s_comp2);
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        ev_pool.clear();
        ev_begin = ev_k0 = ev_k1 = ev_end = nullptr;
        stream = s_h2d = s_d2h = s_comp2 = nullptr;
        device = -1;
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
    }
    int events(size_t n)
    {
        while (ev_pool.size() < n) {
            cudaEvent_t e;
            CU_TRY(cudaEventCreate(&e));
            ev_pool.push_back(e);
        }
        return 0;
    }
    ~Handle() { delete sub[0]; delete sub[1]; sub[0] = sub[1] = nullptr; destroy_cuda(); }
};

// Every entry point that binds a device restores the caller's current device on the way out (a handle may be freed by a
// garbage collector at any time; a multi-GPU process must not find its current device changed behind its back).
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

enum HostArr : int { HA_STATUS = 0, HA_SIZE, HA_STEPS, HA_NREJ, HA_NFEV, HA_TLAST, HA_YLAST, HA_COUNT };

// Host (pinned) copy of one per-trajectory result array, fetched from the device arena on first use.
static const void* host_array(Handle* h, int which)
{
    DevBuf* d[HA_COUNT] = {&h->d_status, &h->d_size, &h->d_steps, &h->d_nrej, &h->d_nfev, &h->d_tlast, &h->d_ylast};
    PinBuf* p[HA_COUNT] = {&h->h_status, &h->h_size, &h->h_steps, &h->h_nrej, &h->h_nfev, &h->h_tlast, &h->h_ylast};
    if (h->device_io) return d[which]->p;
    if (!(h->host_valid & (1u << which)) && h->solved) {
        const size_t per = which == HA_YLAST ? h->ndy * 8 : which == HA_TLAST ? 8 : 4;
        const size_t bytes = h->n_traj * per;
        DeviceGuard g;
        if (cudaSetDevice(h->device) != cudaSuccess || p[which]->ensure(bytes) ||
            cudaMemcpy(p[which]->p, d[which]->p, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            if (g_last_error.empty()) set_error("lazy device-to-host copy of a result array failed");
            return nullptr;
        }
        h->lazy_d2h_bytes += bytes;
        h->host_valid |= 1u << which;
    }
    return p[which]->p;
}

// Built-in (precompiled) or user (NVRTC) kernels behind one interface.
struct Dispatch {
    int rhs_id = -1;
    bool strict = false;
    int device = 0;
    const ProblemEntry* pe = nullptr;       // null => user right-hand side
    int ny = 0, nextra = 0, nargs_dbl = 0;
    // Feature builds of the solve / prep kernels (per-trajectory tolerances): compiled on demand with NVRTC -- for the
    // built-in problems too (jit_id: the problem registered as a run-time right-hand side) -- instead of doubling the
    // number of precompiled instantiations for a rarely used option.
    std::string jit_defs;
    std::string ev_defs;        // "-DCYRK_EV_STATIC=1 ...": the detection solve (VAR_EVDETECT) specialised for this run's event list
    int jit_id = -1;
    std::string defs_for(int variant) const
    {
        if (variant != VAR_EVDETECT || ev_defs.empty()) return jit_defs;
        return jit_defs.empty() ? ev_defs : jit_defs + " " + ev_defs;
    }
    const char* name(int method, int variant) const { return (pe && jit_defs.empty()) ? pe->solve[method - 3][variant].name : (pe ? pe->struct_name : "user"); }
    cudaError_t occupancy(int method, int variant, int* ctas, int block) const
    {
        const std::string d = defs_for(variant);
        if (pe && d.empty()) return pe->solve[method - 3][variant].occupancy(ctas, block);
        return user_rhs_occupancy(pe ? jit_id : rhs_id, method, variant, strict, device, ctas, block, d.c_str());
    }
    cudaError_t launch(int method, int variant, const BatchParams& p, int grid, int block, cudaStream_t s) const
    {
        const std::string d = defs_for(variant);
        if (pe && d.empty()) return pe->solve[method - 3][variant].launch(p, grid, block, s);
        return user_rhs_launch_solve(pe ? jit_id : rhs_id, method, variant, strict, device, p, grid, block, s, d.c_str());
    }
    cudaError_t prep(int method, const BatchParams& p, double* out, int grid, int block, cudaStream_t s) const
    {
        if (pe && jit_defs.empty()) return pe->prep[method - 3](p, out, grid, block, s);
        return user_rhs_launch_prep(pe ? jit_id : rhs_id, method, strict, device, p, out, grid, block, s, jit_defs.c_str());
    }
    int need_jit_id()
    {
        if (pe && jit_id < 0) {
            jit_id = builtin_jit_id(pe->struct_name, pe->ny, pe->nextra, pe->nargs_dbl);
            if (jit_id < 0) { set_error("cannot register built-in problem %s for run-time compilation", pe->struct_name); return -1; }
        }
        return 0;
    }
    int use_feature(const char* defs)
    {
        jit_defs = defs;
        if (pe) {
            jit_id = builtin_jit_id(pe->struct_name, pe->ny, pe->nextra, pe->nargs_dbl);
            if (jit_id < 0) { set_error("cannot register built-in problem %s for run-time compilation", pe->struct_name); return -1; }
        }
        return 0;
    }
    cudaError_t event_roots(int method, const BatchParams& p, int grid, int block, cudaStream_t s) const
    {
        if (pe && jit_defs.empty()) return pe->event_roots[method - 3](p, grid, block, s);
        return user_rhs_launch_event_roots(pe ? jit_id : rhs_id, method, strict, device, p, grid, block, s, jit_defs.c_str());
    }
    cudaError_t step_dense(int method, const StepDenseParams& p, int grid, int block, cudaStream_t s) const
    {
        if (pe) return pe->step_dense[method - 3](p, grid, block, s);
        return user_rhs_launch_step_dense(rhs_id, method, strict, device, p, grid, block, s);
    }
    cudaError_t dense(int method, const DenseEvalParams& p, int grid, int block, cudaStream_t s) const
    {
        if (pe) return pe->dense_eval[method - 3](p, grid, block, s);
        return user_rhs_launch_dense(rhs_id, method, strict, device, p, grid, block, s);
    }
};

static bool make_dispatch(int rhs_id, int fp, int device, Dispatch* d)
{
    d->rhs_id = rhs_id; d->strict = fp == FP_STRICT; d->device = device;
    if (rhs_id >= CYRK_RHS_USER_BASE) {
        int ne = 0;
        if (user_rhs_info(rhs_id, &d->ny, &d->nextra, &d->nargs_dbl, &ne)) return false;
        d->pe = nullptr;
        return true;
    }
    d->pe = registry_lookup(fp, rhs_id);
    if (!d->pe) return false;
    d->ny = d->pe->ny; d->nextra = d->pe->nextra; d->nargs_dbl = d->pe->nargs_dbl;
    return true;
}

// cy/c_common.cpp:20-56
static void find_max_num_steps(size_t ny, size_t nx, size_t user, size_t ram_MB, int* user_flag, size_t* out)
{
    const double ram = ram_MB * (1000.0 * 1000.0) / (sizeof(double) * (1.0 + ny + nx));
    const size_t ram_steps = (size_t)std::floor(ram);
    *user_flag = 0;
    *out = ram_steps;
    if (user > 0) {
        if (user > ram_steps) *out = ram_steps;
        else { *user_flag = 1; *out = user; }
    }
    if (*out > (SIZE_MAX / 10)) *out = SIZE_MAX / 10;
}

struct SolveArgs {
    int rhs_id, method;
    size_t n_traj;
    const double* t_span; size_t n_t_span;
    const double* y0; size_t n_y;
    size_t num_extra;
    const char* args; size_t args_len, n_args;
    size_t max_num_steps, max_ram_MB;
    int dense_output;
    const double* t_eval; size_t n_t_eval;
    const int* event_ids; const int* event_dirs; const size_t* event_max; size_t n_events;
    const double* rtols; size_t n_rtols;
    const double* atols; size_t n_atols;
    double max_step, first_step;
    size_t expected_size;
    int store_mode; unsigned flags; int device;
};

static int upload(Handle* h, DevBuf& dst, const void* src, size_t bytes)
{
    if (dst.ensure(bytes ? bytes : 8)) return -1;
    if (bytes) {
        CU_TRY(cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
        h->h2d_bytes += bytes;
    }
    return 0;
}

static int download(Handle* h, PinBuf& dst, const DevBuf& src, size_t bytes)
{
    if (dst.ensure(bytes)) return -1;
    if (bytes) {
        CU_TRY(cudaMemcpyAsync(dst.p, src.p, bytes, cudaMemcpyDeviceToHost, h->stream));
        h->d2h_bytes += bytes;
    }
    return 0;
}

static int device_sm_count(int dev)
{
    static int cache[64];
    if (dev < 64 && cache[dev]) return cache[dev];
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 64) cache[dev] = n;
    return n;
}

// Host-I/O fast path for the final-state variant: the batch is cut into chunks that flow through
//   copy-in stream  ->  two alternating compute streams  ->  copy-out stream
// so PCIe traffic hides behind the solve kernels, and the tail of one chunk's persistent kernel overlaps the head of the
// next (they run on different streams).  Results are identical to the single-launch path: trajectories are independent.
static int solve_lean_pipelined(Handle* h, const SolveArgs& a, const Dispatch* pe, BatchParams p, int grid, int block, bool use_prep, int lean_variant)
{
    const size_t B = a.n_traj, NY = pe->ny, NDY = pe->ny + pe->nextra;
    const size_t args_stride = (size_t)pe->nargs_dbl * sizeof(double);
    // Chunk schedule: full chunks in the middle, a short one in front (the first copy-in is not overlapped with any
    // compute) and two short ones at the end (the last copy-out is not overlapped either).
    // chunk size: `pipe_chunk` (2 Mi: measured best at 8 Mi) for large batches, a quarter of the batch below 4 Mi so that
    // a 256 Ki ... 2 Mi batch still overlaps its copies with the solve; never more than 32 chunks
    const size_t chunk = std::max<size_t>(std::min<size_t>((size_t)g_opt_pipe_chunk, ((B + 3) / 4 + 1023) / 1024 * 1024), (B + 31) / 32);
    std::vector<std::pair<size_t, size_t>> sched;          // (offset, count)
    if (g_opt_pipe_sched == 1) {
        // Head / body / tail: what the copies want (a SHORT first chunk: its copy-in overlaps nothing; short last chunks: the
        // last copy-out overlaps nothing) is decoupled from what the solve wants (FEW, LARGE chunks: a chunk of 256 Ki runs
        // the persistent grid at 1.7 trajectories per lane).  head = B/16 in [32 Ki, 512 Ki]; tail = 2 head + head; the body is
        // cut into equal parts of at most pipe_chunk.
        size_t off = 0;
        auto push = [&](size_t n) { n = std::min(n, B - off); if (n) { sched.emplace_back(off, n); off += n; } };
        const size_t head_min = std::min<size_t>((size_t)32 << 10, (size_t)g_opt_pipe_chunk);
        const size_t head = std::min<size_t>(std::max<size_t>(((B / 16) + 1023) / 1024 * 1024, head_min), (size_t)512 << 10);
        if (B <= 4 * head) push(B);
        else {
            const size_t body = B - 4 * head;
            const size_t parts = (body + (size_t)g_opt_pipe_chunk - 1) / (size_t)g_opt_pipe_chunk;
            const size_t part = ((body + parts - 1) / parts + 1023) / 1024 * 1024;
            push(head);
            while (B - off > 3 * head + part) push(part);
            push(B - off - 3 * head);
            push(2 * head);
            push(head);
        }
    } else {
        size_t off = 0;
        auto push = [&](size_t n) { n = std::min(n, B - off); if (n) { sched.emplace_back(off, n); off += n; } };
        const bool taper = g_opt_pipe_taper && B >= 4 * chunk;
        const size_t tail = taper ? chunk / 2 + chunk / 4 : 0;
        if (taper) push(chunk / 4);
        while (B - off > tail + chunk / 2) push(chunk);
        push(B - off - std::min(tail, B - off));
        if (taper) { push(chunk / 2); push(chunk / 4); }
    }
    const size_t nchunk = sched.size();
    if (h->events(4 * nchunk)) return -1;
    if (h->d_y0.ensure(B * NY * 8) || h->d_queue.ensure(8 * nchunk)) return -1;
    if (use_prep && h->d_prep.ensure(B * (NDY + 2) * 8)) return -1;
    if (h->h_status.ensure(B * 4) || h->h_steps.ensure(B * 4) || h->h_ylast.ensure(B * NDY * 8)) return -1;
    const bool per_traj_args = a.n_args > 1 && pe->nargs_dbl > 0;
    const bool per_traj_span = a.n_t_span > 1;
    // shared inputs once, on the copy-in stream
    CU_TRY(cudaEventRecord(h->ev_begin, h->s_h2d));
    const char* args_src = a.args;
    if (pe->nargs_dbl > 0) {
        const size_t n_blobs = per_traj_args ? B : 1;
        if (h->d_args.ensure(n_blobs * args_stride)) return -1;
        if (a.n_args == 0 || a.args_len != args_stride) {          // repack (zero padded) through pinned staging
            if (h->h_stage.ensure(n_blobs * args_stride)) return -1;
            char* st = h->h_stage.as<char>();
            memset(st, 0, n_blobs * args_stride);
            if (a.n_args >= 1)
                for (size_t i = 0; i < n_blobs; i++) memcpy(st + i * args_stride, a.args + i * a.args_len, a.args_len);
            args_src = st;
        }
        if (!per_traj_args) {
            CU_TRY(cudaMemcpyAsync(h->d_args.p, args_src, args_stride, cudaMemcpyHostToDevice, h->s_h2d));
            h->h2d_bytes += args_stride;
        }
    } else if (h->d_args.ensure(8)) return -1;
    if (h->d_tspan.ensure(a.n_t_span * 16)) return -1;
    if (!per_traj_span) {
        CU_TRY(cudaMemcpyAsync(h->d_tspan.p, a.t_span, 16, cudaMemcpyHostToDevice, h->s_h2d));
        h->h2d_bytes += 16;
    }
    CU_TRY(cudaMemsetAsync(h->d_queue.p, 0, 8 * nchunk, h->s_h2d));
    p.args_stride = per_traj_args ? (unsigned)args_stride : 0u;
    p.t_span_stride = per_traj_span ? 2u : 0u;
    h->in_y0 = h->d_y0.as<double>(); h->in_args = h->d_args.as<char>(); h->in_tspan = h->d_tspan.as<double>();
    h->tspan_stride = p.t_span_stride; h->args_stride = p.args_stride;

    for (size_t c = 0; c < nchunk; c++) {
        const size_t off = sched[c].first, n = sched[c].second;
        cudaEvent_t e_in = h->ev_pool[4 * c], e_k0 = h->ev_pool[4 * c + 1], e_k1 = h->ev_pool[4 * c + 2];
        cudaStream_t sc = (c & 1) ? h->s_comp2 : h->stream;
        // ---- copy in
        CU_TRY(cudaMemcpyAsync(h->d_y0.as<double>() + off * NY, a.y0 + off * NY, n * NY * 8, cudaMemcpyHostToDevice, h->s_h2d));
        h->h2d_bytes += n * NY * 8;
        if (per_traj_args) {
            CU_TRY(cudaMemcpyAsync(h->d_args.as<char>() + off * args_stride, args_src + off * args_stride, n * args_stride,
                                   cudaMemcpyHostToDevice, h->s_h2d));
            h->h2d_bytes += n * args_stride;
        }
        if (per_traj_span) {
            CU_TRY(cudaMemcpyAsync(h->d_tspan.as<double>() + off * 2, a.t_span + off * 2, n * 16, cudaMemcpyHostToDevice, h->s_h2d));
            h->h2d_bytes += n * 16;
        }
        CU_TRY(cudaEventRecord(e_in, h->s_h2d));
        // ---- solve
        BatchParams q = p;
        q.n_traj = n;
        q.y0 = h->d_y0.as<double>() + off * NY;
        q.args = h->d_args.as<char>() + (per_traj_args ? off * args_stride : 0);
        q.t_span = h->d_tspan.as<double>() + (per_traj_span ? off * 2 : 0);
        q.status = p.status + off; q.size = p.size + off; q.steps_taken = p.steps_taken + off;
        q.n_rejected = p.n_rejected + off; q.nfev = p.nfev + off; q.t_last = p.t_last + off; q.y_last = p.y_last + off * NDY;
        if (p.rtol_pt) { q.rtol_pt = p.rtol_pt + off * NY; q.atol_pt = p.atol_pt + off * NY; }
        q.queue = h->d_queue.as<unsigned long long>() + c;
        const int cgrid = (int)std::min<unsigned long long>((unsigned long long)grid, (n + block - 1) / block);
        CU_TRY(cudaStreamWaitEvent(sc, e_in, 0));
        CU_TRY(cudaEventRecord(e_k0, sc));
        if (use_prep) {
            double* pr = h->d_prep.as<double>() + off * (NDY + 2);
            const int pgrid = (int)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)device_sm_count(h->device) * 16);
            if (pe->prep(a.method, q, pr, pgrid, 256, sc) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("prep kernel launch failed"); return -1; }
            q.prep = pr;
            h->launches++;
        }
        if (pe->launch(a.method, lean_variant, q, cgrid, block, sc) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("solve kernel launch failed"); return -1; }
        h->launches++;
        CU_TRY(cudaEventRecord(e_k1, sc));
        // ---- copy out
        CU_TRY(cudaStreamWaitEvent(h->s_d2h, e_k1, 0));
#define D2H(host, dev, T, per) CU_TRY(cudaMemcpyAsync(h->host.as<T>() + off * (per), h->dev.as<T>() + off * (per), n * (per) * sizeof(T), cudaMemcpyDeviceToHost, h->s_d2h)); h->d2h_bytes += n * (per) * sizeof(T)
        // what every caller reads: 8 + 8*NDY bytes per trajectory.  size (= steps + 1 here), n_rejected, nfev and t_last
        // stay in the device arena until their getter asks for them (host_array).
        D2H(h_ylast, d_ylast, double, NDY); D2H(h_status, d_status, int, 1); D2H(h_steps, d_steps, unsigned, 1);
#undef D2H
    }
    CU_TRY(cudaEventRecord(h->ev_end, h->s_d2h));
    CU_TRY(cudaStreamSynchronize(h->s_d2h));
    CU_TRY(cudaStreamSynchronize(h->stream));
    CU_TRY(cudaStreamSynchronize(h->s_comp2));
    // kernel time = span from the first chunk's first launch to the last kernel's end (chunks overlap across the two
    // compute streams, so summing per-chunk spans would double count)
    float ms = 0.f, k = 0.f;
    for (size_t c = 0; c < nchunk; c++) { cudaEventElapsedTime(&ms, h->ev_pool[1], h->ev_pool[4 * c + 2]); k = std::max(k, ms); }
    h->kernel_ms = k;
    cudaEventElapsedTime(&h->total_ms, h->ev_begin, h->ev_end);
    h->row_cap = h->dense_cap = h->ev_cap = 0;
    h->host_valid = (1u << HA_STATUS) | (1u << HA_STEPS) | (1u << HA_YLAST);
    return 0;
}

// The thread-per-trajectory path.  Returns 0 / -1.
// Last pass of the step-parallel t_eval path: rows saved (= position of the t_eval cursor after the last booked step,
// cysolver.cpp:848-878) and the last saved row, per trajectory.
__global__ void teval_last_kernel(unsigned long long n_traj, const unsigned* __restrict__ steps, const double* __restrict__ t_steps,
                                  unsigned step_cap, const double* __restrict__ t_span, unsigned t_span_stride,
                                  unsigned rec_doubles, const double* __restrict__ t_eval, unsigned n_t_eval, const double* __restrict__ t_rows,
                                  const double* __restrict__ y_rows, unsigned row_cap, unsigned ndy, unsigned* __restrict__ size,
                                  double* __restrict__ t_last, double* __restrict__ y_last)
{
    const unsigned long long traj = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (traj >= n_traj) return;
    const unsigned ns = steps[traj];
    const double* ts = t_span + traj * t_span_stride;
    const bool fwd = (ts[1] - ts[0]) >= 0.0;
    unsigned sz = 0;
    if (ns > 0) {
        const double x = t_steps[(traj * step_cap + ns) * rec_doubles];
        unsigned a = 0, b = n_t_eval;
        while (a < b) {
            const unsigned m = (a + b) >> 1;
            const double v = t_eval[m];
            if (fwd ? !(x < v) : (v < x)) a = m + 1; else b = m;
        }
        sz = fwd ? a : n_t_eval - a;
    }
    size[traj] = sz;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (sz > 0 && sz <= row_cap) {
        const unsigned long long r = traj * row_cap + (sz - 1);
        t_last[traj] = t_rows[r];
        for (unsigned i = 0; i < ndy; i++) y_last[traj * ndy + i] = y_rows[r * ndy + i];
    } else {
        t_last[traj] = nan;
        for (unsigned i = 0; i < ndy; i++) y_last[traj * ndy + i] = nan;
    }
}

// Largest entry of up to three per-trajectory counter arrays (rows saved, interpolants stored, event hits): what decides
// whether an arena overflowed.  Reduced on the device: walking 12 Mi counters on one host core took 10 ms of an 85 ms step.
__global__ void max_counts_kernel(const unsigned* __restrict__ a, unsigned long long na, const unsigned* __restrict__ b, unsigned long long nb,
                                  const unsigned* __restrict__ c, unsigned long long nc, unsigned* __restrict__ out /*[3]*/)
{
    const unsigned* arr[3] = {a, b, c};
    const unsigned long long len[3] = {na, nb, nc};
    for (int k = 0; k < 3; k++) {
        if (arr[k] == nullptr) continue;
        unsigned m = 0;
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < len[k]; i += (unsigned long long)gridDim.x * blockDim.x)
            m = max(m, arr[k][i]);
        m = __reduce_max_sync(0xffffffffu, m);
        if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out + k, m);
    }
}

// Event lists are ragged (the reference keeps one std::vector per event, cysolution.cpp:423-448).  On the device every
// (trajectory, event) list owns `ev_cap` slots (default 64) but typically holds a handful of hits, so before they cross PCIe
// the lists are packed back to back (CSR): offsets[l] .. offsets[l+1] index the hits of list l = traj * n_events + e.
// Exclusive prefix sum of the per-list counts: block sums -> serial scan of the (few thousand) block sums -> offsets.
constexpr int SCAN_BLOCK = 256, SCAN_ITEMS = 16;
__global__ void scan_block_sums_kernel(const unsigned* __restrict__ cnt, unsigned long long n, unsigned long long* __restrict__ bsum)
{
    __shared__ unsigned long long sh[SCAN_BLOCK / 32];
    const unsigned long long base = ((unsigned long long)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) s += cnt[base + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < SCAN_BLOCK / 32; w++) t += sh[w];
        bsum[blockIdx.x] = t;
    }
}
__global__ void scan_block_offsets_kernel(unsigned long long* __restrict__ bsum, unsigned nb, unsigned long long* __restrict__ total_out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long run = 0;
        for (unsigned b = 0; b < nb; b++) { const unsigned long long v = bsum[b]; bsum[b] = run; run += v; }
        *total_out = run;
    }
}
__global__ void scan_write_kernel(const unsigned* __restrict__ cnt, unsigned long long n, const unsigned long long* __restrict__ bsum,
                                  unsigned long long* __restrict__ off, unsigned long long first)
{
    __shared__ unsigned long long sh[SCAN_BLOCK];
    const unsigned long long base = ((unsigned long long)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    unsigned c[SCAN_ITEMS];
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { c[i] = (base + i < n) ? cnt[base + i] : 0u; s += c[i]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {             // inclusive Hillis-Steele scan of the thread totals
        const unsigned long long v = (threadIdx.x >= (unsigned)o) ? sh[threadIdx.x - o] : 0ull;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = first + bsum[blockIdx.x] + sh[threadIdx.x] - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) off[base + i] = run; run += c[i]; }
    if (base <= n && n < base + SCAN_ITEMS) off[n] = first + bsum[blockIdx.x] + sh[threadIdx.x];     // the thread whose range holds index n
}
__global__ void pack_events_kernel(unsigned long long n_lists, unsigned cap, unsigned used_cap, unsigned ndy,
                                   const unsigned* __restrict__ cnt, const unsigned long long* __restrict__ off, unsigned long long first,
                                   const double* __restrict__ t_in, const double* __restrict__ y_in,
                                   double* __restrict__ t_out, double* __restrict__ y_out)
{
    const unsigned long long total = n_lists * used_cap;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long l = idx / used_cap;
        const unsigned j = (unsigned)(idx - l * used_cap);
        if (j >= cnt[l]) continue;
        const unsigned long long src = l * cap + j, dst = off[l] - first + j;      // (offsets are global, the output buffer is this chunk's)
        t_out[dst] = t_in[src];
        for (unsigned i = 0; i < ndy; i++) y_out[dst * ndy + i] = y_in[src * ndy + i];
    }
}

// max of up to three device counter arrays into out[3]; returns with the stream synchronised
static int counts_max(Handle* h, const unsigned* a, size_t na, const unsigned* b, size_t nb, const unsigned* c, size_t nc, unsigned* out)
{
    if (h->d_max.ensure(16)) return -1;
    CU_TRY(cudaMemsetAsync(h->d_max.p, 0, 16, h->stream));
    const size_t n = std::max(std::max(a ? na : 0, b ? nb : 0), c ? nc : 0);
    if (n > 0) {
        const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)device_sm_count(h->device) * 16);
        max_counts_kernel<<<grid, 256, 0, h->stream>>>(a, na, b, nb, c, nc, h->d_max.as<unsigned>());
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaMemcpyAsync(out, h->d_max.p, 12, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

static int solve_thread_path(Handle* h, const SolveArgs& a, const Dispatch* pe)
{
    const size_t B = a.n_traj, NY = pe->ny, NX = pe->nextra, NDY = NY + NX;
    const bool device_io = (a.flags & CYRK_FLAG_DEVICE_IO) != 0;
    const bool want_rows = a.store_mode == CYRK_STORE_ROWS;
    const bool dense = a.dense_output != 0;
    const bool stores = dense || want_rows || a.n_t_eval > 0;
    int variant = a.n_events > 0 ? ((stores || !g_opt_events_variant) ? VAR_FULL : VAR_EVENTS) : (stores ? VAR_ROWS : VAR_LEAN);
    const int q_order = a.method == CYRK_METHOD_RK23 ? 3 : a.method == CYRK_METHOD_RK45 ? 4 : 7;

    // Step-parallel interpolants (rk_core.cuh, step_dense_body): a final-state solve that records its booked steps, then one
    // thread per (trajectory, step).  Taken when the interpolant is only needed for output (no events), the right-hand side has
    // no extra outputs (those change how the reference rolls its state inside the t_eval loop, cysolver.cpp:915-961), and t_eval
    // cannot be exhausted before the last step (afterwards the reference saves raw steps again).
    // Measured (profiles/r01_notes.md): worth it when the interpolants are STORED (dense_output); for t_eval alone the
    // per-trajectory kernel is as fast and needs no per-step records (option two_pass = 2 forces it there as well).
    bool two_pass = g_opt_two_pass != 0 && variant == VAR_ROWS && NX == 0 && (dense || (a.n_t_eval > 0 && g_opt_two_pass == 2)) &&
                    (a.n_t_eval == 0 || want_rows);
    const bool te_pt = a.n_t_eval > 0 && (a.flags & CYRK_FLAG_T_EVAL_PER_TRAJ) != 0;
    const bool tol_pt = (a.flags & CYRK_FLAG_TOL_PER_TRAJ) != 0;
    if (two_pass && a.n_t_eval > 0) {
        if (a.n_t_span > 1 || te_pt) two_pass = false;      // (the step-parallel path ranks samples in ONE shared grid)
        else {
            double ts0[2];
            if (device_io) {
                if (cudaMemcpy(ts0, a.t_span, sizeof(ts0), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); set_error("cannot read t_span from the device"); return -1; }
            } else { ts0[0] = a.t_span[0]; ts0[1] = a.t_span[1]; }
            const double last = a.t_eval[a.n_t_eval - 1];         // t_eval arrives in the order the run passes it
            two_pass = ((ts0[1] - ts0[0]) >= 0.0) ? (last >= ts0[1]) : (last <= ts0[1]);
        }
    }

    h->ny = NY; h->nx = NX; h->ndy = NDY; h->q_order = q_order;
    h->n_traj = B; h->n_events = a.n_events; h->n_t_eval = a.n_t_eval;
    h->dense_output = dense; h->device_io = device_io; h->dense_from_steps = false;
    h->ev_packed = false; h->ev_total = 0;
    h->h2d_bytes = h->d2h_bytes = h->lazy_d2h_bytes = 0; h->launches = 0; h->kernel_ms = 0.f;
    h->host_valid = 0;

    BatchParams p;
    memset(&p, 0, sizeof(p));
    p.n_traj = B;
    // Final-state solve of a batch whose args are shared: the values travel in the kernel parameter block and the right-hand
    // side reads them from the constant bank (VAR_LEAN_SHARED) -- 4 registers per lane fewer for Lotka-Volterra / RK45, which
    // is the difference between 7 and 8 resident CTAs per SM.
    if (variant == VAR_LEAN && g_opt_shared_args && pe->nargs_dbl > 0 && pe->nargs_dbl <= CYRK_SHARED_ARGS_MAX && a.n_args <= 1) {
        const size_t want = (size_t)pe->nargs_dbl * sizeof(double);
        bool ok = true;
        if (a.n_args == 1 && a.args != nullptr) {
            if (device_io) {
                ok = a.args_len == want && cudaMemcpy(p.shared_args, a.args, want, cudaMemcpyDeviceToHost) == cudaSuccess;
                if (!ok) cudaGetLastError();
            } else {
                memcpy(p.shared_args, a.args, std::min(want, a.args_len));      // (zero padded like the device copy)
            }
        }
        if (ok) variant = VAR_LEAN_SHARED;
    }
    for (size_t i = 0; i < NY && !tol_pt; i++) {
        double r = a.rtols[a.n_rtols > 1 ? i : 0];
        if (r < 100.0 * std::numeric_limits<double>::epsilon()) r = 100.0 * std::numeric_limits<double>::epsilon();
        p.rtol[i] = r;
        p.atol[i] = a.atols[a.n_atols > 1 ? i : 0];
    }
    if (tol_pt) {
        // one (rtol, atol) row per trajectory, held in per-lane registers by the CYRK_TOL_PER_TRAJ build of the kernels;
        // rtol is floored like the reference does in place (rk.cpp:600-608)
        std::vector<double> rt(a.rtols, a.rtols + B * NY);
        for (double& r : rt) if (r < 100.0 * std::numeric_limits<double>::epsilon()) r = 100.0 * std::numeric_limits<double>::epsilon();
        if (h->d_rtol.ensure(B * NY * 8) || h->d_atol.ensure(B * NY * 8)) return -1;
        CU_TRY(cudaMemcpyAsync(h->d_rtol.p, rt.data(), B * NY * 8, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaMemcpyAsync(h->d_atol.p, a.atols, B * NY * 8, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));          // (`rt` is a temporary)
        p.rtol_pt = h->d_rtol.as<double>(); p.atol_pt = h->d_atol.as<double>();
    }
    p.max_step = a.max_step;
    p.has_max_step = (a.max_step < std::numeric_limits<double>::infinity()) ? 1 : 0;
    p.first_step = a.first_step;
    {
        int uf; size_t mx;
        find_max_num_steps(NY, NX, a.max_num_steps, a.max_ram_MB, &uf, &mx);
        p.user_max_steps = uf;
        p.max_num_steps = (unsigned)std::min<size_t>(mx, 0xffffffffu);
    }
    p.dense_output = dense ? 1 : 0;
    p.n_events = (int)a.n_events;
    for (size_t e = 0; e < a.n_events; e++) {
        p.ev_id[e] = a.event_ids[e];
        p.ev_dir[e] = a.event_dirs ? a.event_dirs[e] : 0;
        p.ev_max[e] = a.event_max ? (unsigned long long)a.event_max[e] : ~0ull;
    }
    p.root_batch = std::min(32, std::max(1, g_opt_root_batch));
    const bool use_prep = g_opt_prep < 0 ? (B >= (size_t(1) << 18)) : (g_opt_prep != 0);
    p.refill_min = g_opt_refill_min > 0 ? std::min(32, g_opt_refill_min) : (use_prep ? 1 : 4);

    CU_TRY(cudaEventRecord(h->ev_begin, h->stream));

    // ---- per-trajectory outputs
    if (h->d_status.ensure(B * 4) || h->d_size.ensure(B * 4) || h->d_steps.ensure(B * 4) || h->d_nrej.ensure(B * 4) ||
        h->d_nfev.ensure(B * 4) || h->d_tlast.ensure(B * 8) || h->d_ylast.ensure(B * NDY * 8) || h->d_queue.ensure(8))
        return -1;
    p.status = h->d_status.as<int>();
    p.size = h->d_size.as<unsigned>();
    p.steps_taken = h->d_steps.as<unsigned>();
    p.n_rejected = h->d_nrej.as<unsigned>();
    p.nfev = h->d_nfev.as<unsigned>();
    p.t_last = h->d_tlast.as<double>();
    p.y_last = h->d_ylast.as<double>();
    p.queue = h->d_queue.as<unsigned long long>();

    // CTA size: 128 threads unless 96 keeps more warps resident (register-heavy variants: Lorenz/DOP853 with rows + dense
    // needs 186 registers -> 2 x 128 or 3 x 96 threads per SM; measured 5.31 vs 4.92 ms on the C2 t_eval+dense workload,
    // while 64-thread CTAs, although they fit 10 warps, were no faster than 128).
    int grid = 0, block = 0;
    auto pick_grid = [&](int solve_variant) -> int {
        block = g_opt_block > 0 ? g_opt_block : 128;
        int ctas = 0;
        if (pe->occupancy(a.method, solve_variant, &ctas, block) != cudaSuccess) {
            if (g_last_error.empty()) set_error("occupancy query failed for %s", pe->name(a.method, solve_variant));
            cudaGetLastError();
            return -1;
        }
        if (g_opt_block <= 0) {
            int ctas96 = 0;
            if (pe->occupancy(a.method, solve_variant, &ctas96, 96) == cudaSuccess && ctas96 * 96 > ctas * 128) { block = 96; ctas = ctas96; }
            else cudaGetLastError();
        }
        if (g_opt_ctas_per_sm > 0) ctas = g_opt_ctas_per_sm;
        if (ctas <= 0) { set_error("kernel %s cannot be resident with %d threads per CTA", pe->name(a.method, solve_variant), block); return -1; }
        grid = device_sm_count(h->device) * ctas;
        const unsigned long long need = (B + block - 1) / block;
        if ((unsigned long long)grid > need) grid = (int)std::max<unsigned long long>(need, 1);
        if (getenv("CYRK_B200_DEBUG")) fprintf(stderr, "[cyrk_b200] %s: block %d, %d CTAs/SM, grid %d\n", pe->name(a.method, solve_variant), block, ctas, grid);
        return 0;
    };
    // Events with store='final': detection solve + one thread per triggered step (rk_core.cuh, event_roots_body)
    bool ev2 = variant == VAR_EVENTS && g_opt_events_two_pass != 0;
    Dispatch disp_ev;                      // (a copy: the caller's dispatch stays as it is)
    if (ev2 && g_opt_events_static) {
        // The detection solve evaluates every event function on every booked step.  Compiled for THIS event list (ids and
        // directions as compile-time constants) the event table's switch folds away and g_old / the hit counters live in
        // registers; with the list as launch parameters they are local-memory arrays.  One NVRTC compilation per
        // (problem, method, FP build, event list), cached in the process and on disk; any failure keeps the generic kernel.
        disp_ev = *pe;
        std::string d = "-DCYRK_EV_STATIC=1 -DCYRK_EV_N=" + std::to_string(a.n_events);
        for (size_t e = 0; e < CYRK_MAX_EVENTS; e++) {
            d += " -DCYRK_EV_ID" + std::to_string(e) + "=" + std::to_string(e < a.n_events ? a.event_ids[e] : 0);
            const int dir = (e < a.n_events && a.event_dirs) ? a.event_dirs[e] : 0;
            d += " -DCYRK_EV_DIR" + std::to_string(e) + "=" + std::to_string(dir > 0 ? 1 : dir < 0 ? -1 : 0);
        }
        // RK23 with a small state: the detection solve sits just above 96 registers (4 CTAs of 128 threads per SM); capped
        // for 6 CTAs (80 registers, a few bytes of spills) it measured 17.7 vs 19.6 ms on C3full at 1 Mi trajectories, 5 CTAs
        // 18.1, 8 CTAs 18.8 (profiles/r02_c3full_minctas.log)
        if (a.method == CYRK_METHOD_RK23 && pe->ny <= 4 && g_opt_block <= 0) d += " -DCYRK_USER_MINCTAS=6";
        disp_ev.ev_defs = d;
        int ctas_probe = 0;
        if (disp_ev.need_jit_id() == 0 && disp_ev.occupancy(a.method, VAR_EVDETECT, &ctas_probe, 128) == cudaSuccess) pe = &disp_ev;
        else { cudaGetLastError(); g_last_error.clear(); }
    }
    if (pick_grid(two_pass ? VAR_STEPS : (ev2 ? VAR_EVDETECT : variant))) return -1;

    // ---- host buffers + final-state variant + a batch worth cutting up: pipelined copy/compute path
    if (!device_io && (variant == VAR_LEAN || variant == VAR_LEAN_SHARED) && g_opt_pipe_min > 0 && B >= (size_t)g_opt_pipe_min)
        return solve_lean_pipelined(h, a, pe, p, grid, block, use_prep, variant);

    // ---- inputs
    const size_t args_stride = (size_t)pe->nargs_dbl * sizeof(double);
    if (device_io) {
        p.y0 = a.y0;
        p.args = a.args;
        p.args_stride = a.n_args > 1 ? (unsigned)args_stride : 0u;
        if (a.args_len != args_stride && pe->nargs_dbl > 0) {
            set_error("CYRK_FLAG_DEVICE_IO requires args blobs of exactly %zu bytes for this right-hand side (got %zu)",
                      args_stride, a.args_len);
            return -1;
        }
        p.t_span = a.t_span;
        p.t_span_stride = a.n_t_span > 1 ? 2u : 0u;
    } else {
        if (upload(h, h->d_y0, a.y0, B * NY * sizeof(double))) return -1;
        p.y0 = h->d_y0.as<double>();
        // args: zero-pad every blob to the functor's size
        const size_t n_blobs = a.n_args > 1 ? B : 1;
        if (pe->nargs_dbl > 0) {
            if (a.args_len == args_stride && a.n_args >= 1) {
                if (upload(h, h->d_args, a.args, n_blobs * args_stride)) return -1;
            } else {
                if (h->h_stage.ensure(n_blobs * args_stride)) return -1;
                char* st = h->h_stage.as<char>();
                memset(st, 0, n_blobs * args_stride);
                if (a.n_args >= 1)
                    for (size_t i = 0; i < n_blobs; i++) memcpy(st + i * args_stride, a.args + i * a.args_len, a.args_len);
                if (upload(h, h->d_args, st, n_blobs * args_stride)) return -1;
            }
        } else {
            if (h->d_args.ensure(8)) return -1;
        }
        p.args = h->d_args.as<char>();
        p.args_stride = a.n_args > 1 ? (unsigned)args_stride : 0u;
        if (upload(h, h->d_tspan, a.t_span, a.n_t_span * 2 * sizeof(double))) return -1;
        p.t_span = h->d_tspan.as<double>();
        p.t_span_stride = a.n_t_span > 1 ? 2u : 0u;
    }
    h->in_y0 = p.y0; h->in_args = p.args; h->in_tspan = p.t_span;
    h->tspan_stride = p.t_span_stride; h->args_stride = p.args_stride;

    // t_eval: always a host pointer (small, shared).  Backward runs search the reversed copy
    // (cysolver.cpp:381-400).
    std::vector<double> te;
    if (a.n_t_eval) {
        te.assign(a.t_eval, a.t_eval + a.n_t_eval);
        bool fwd0;
        if (device_io) {
            double ts0[2];
            CU_TRY(cudaMemcpyAsync(ts0, a.t_span, sizeof(ts0), cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
            fwd0 = (ts0[1] - ts0[0]) >= 0.0;
        } else {
            fwd0 = (a.t_span[1] - a.t_span[0]) >= 0.0;
            if (a.n_t_span > 1) {
                for (size_t i = 1; i < B; i++)
                    if (((a.t_span[2 * i + 1] - a.t_span[2 * i]) >= 0.0) != fwd0) {
                        set_error("t_eval is shared by the batch: every t_span must integrate in the same direction");
                        return -1;
                    }
            }
        }
        if (te_pt) te.assign(a.t_eval, a.t_eval + B * a.n_t_eval);
        if (!fwd0) for (size_t r = 0; r < (te_pt ? B : 1); r++) std::reverse(te.begin() + r * a.n_t_eval, te.begin() + (r + 1) * a.n_t_eval);
        p.t_eval_stride = te_pt ? (unsigned)a.n_t_eval : 0u;
        if (h->d_teval.ensure(te.size() * sizeof(double))) return -1;
        // synchronous on purpose: `te` is a stack-owned temporary
        CU_TRY(cudaMemcpyAsync(h->d_teval.p, te.data(), te.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        h->h2d_bytes += te.size() * sizeof(double);
        p.t_eval = h->d_teval.as<double>();
        p.n_t_eval = (unsigned)a.n_t_eval;
    }

    size_t row_cap = 0, dense_cap = 0, ev_cap = 0, ev_used = 0;
    Handle::Force& force = h->force;            // (inactive unless this handle is solving one chunk of a streamed batch)
    if (want_rows) row_cap = a.n_t_eval ? a.n_t_eval : (a.expected_size ? a.expected_size : 256);
    if (want_rows && force.active && force.row_cap) row_cap = force.row_cap;
    if (dense) dense_cap = a.expected_size ? a.expected_size + 1 : 257;
    if (a.n_events) ev_cap = a.expected_size ? a.expected_size : 64;

    size_t p_scap = a.expected_size ? a.expected_size + 1 : 257, p_mx = 0;
    if (force.active && force.scap) p_scap = force.scap;
    h->took_two_pass = two_pass;
    BatchParams p_steps;
    if (two_pass) {
        // ---- pass A: final-state solve that records (t, y, step) of every booked step
        size_t& scap = p_scap;
        const bool rows_are_steps = want_rows && a.n_t_eval == 0;     // the step records ARE the requested rows
        BatchParams& pa = p_steps;
        pa = p;
        pa.t_eval = nullptr; pa.n_t_eval = 0; pa.dense_output = 0;
        // (the recording solve runs with inline set-up and refill_min 4 unless the options say otherwise: what costs is
        // the immediate refill that comes with the start-up kernel -- Lorenz/DOP853, both passes, 2 Mi trajectories:
        // refill_min 1 / 2 / 4 / 8 = 82.2 / 70.4 / 68.2 / 69.0 ms, start-up kernel + refill_min 4: 67.8; 256 Ki 9.0 vs 10.0,
        // 1 Mi 34.4 vs 40.4 -- profiles/r02_c2full_refill_min.log, r02_prep_c2full.log.  Round 1's automatic rule had
        // switched to the start-up kernel + refill_min 1 from 2 Mi.)
        const bool use_prep = g_opt_prep > 0;
        pa.refill_min = g_opt_refill_min > 0 ? std::min(32, g_opt_refill_min) : (use_prep ? 1 : 4);
        const size_t RA = CYRK_STEP_REC_DOUBLES(NY);
        for (int pass = 0; pass < 2; pass++) {
            scap = (scap + 31) / 32 * 32;       // whole warps per trajectory in step_dense_kernel
            if (force.active && force.recs) pa.step_recs = force.recs;
            else { if (h->d_st.ensure(B * scap * RA * 8)) return -1; pa.step_recs = h->d_st.as<double>(); }
            pa.t_rows = nullptr; pa.y_rows = nullptr;
            pa.row_cap = (unsigned)scap;
            CU_TRY(cudaMemsetAsync(h->d_queue.p, 0, 8, h->stream));
            CU_TRY(cudaEventRecord(h->ev_k0, h->stream));
            if (use_prep) {
                if (h->d_prep.ensure(B * (NDY + 2) * 8)) return -1;
                const int pgrid = (int)std::min<unsigned long long>((B + 255) / 256, (unsigned long long)device_sm_count(h->device) * 16);
                if (pe->prep(a.method, pa, h->d_prep.as<double>(), pgrid, 256, h->stream) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("prep kernel launch failed"); return -1; }
                h->launches++;
                pa.prep = h->d_prep.as<double>();
            }
            if (pe->launch(a.method, VAR_STEPS, pa, grid, block, h->stream) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("solve kernel launch failed"); return -1; }
            h->launches++;
            CU_TRY(cudaEventRecord(h->ev_k1, h->stream));
            unsigned mxs[3] = {0, 0, 0};
            if (counts_max(h, pa.size, B, nullptr, 0, nullptr, 0, mxs)) return -1;      // (synchronises the stream)
            { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1); h->kernel_ms += ms; if (getenv("CYRK_B200_DEBUG")) fprintf(stderr, "[cyrk_b200] step-recording solve: %.3f ms\n", ms); }
            const size_t mx = mxs[0];
            p_mx = mx;
            if (mx <= scap) break;
            if (force.active && force.scap) { force.need_scap = mx; return 1; }      // the parent re-plans the whole batch
            if (pass == 1) { set_error("internal: arena overflow after the exact-capacity re-run"); return -1; }
            if (!dense && !rows_are_steps) {
                // t_eval only: the per-trajectory kernel needs no per-step storage at all.  When some trajectory takes so
                // many steps that the records would eat the device (stiff / diverging problems), go back to it.
                size_t fr = 0, tot = 0;
                cudaMemGetInfo(&fr, &tot);
                if ((double)B * (double)mx * (8.0 * RA) > 0.5 * (double)(fr + h->d_st.cap)) { two_pass = false; break; }
            }
            // deterministic re-run with the exact capacity (the probe chunk of a streamed batch leaves headroom: its
            // capacity becomes the layout of every later chunk)
            scap = (force.active && !force.scap) ? mx + mx / 16 + 8 : mx;
        }
        if (!two_pass) { h->d_st.release(); h->kernel_ms = 0.f; h->took_two_pass = false; if (pick_grid(variant)) return -1; }
    }
    if (two_pass) {
        const size_t scap = p_scap;
        h->step_cap = scap;
        const bool rows_are_steps = want_rows && a.n_t_eval == 0;
        // ---- pass B: one thread per (trajectory, booked step): interpolant, dense record, t_eval samples
        row_cap = rows_are_steps ? scap : (want_rows ? a.n_t_eval : 0);
        dense_cap = dense ? scap : 0;
        StepDenseParams q;
        memset(&q, 0, sizeof(q));
        const bool materialise = dense && g_opt_dense_store != 0;
        q.n_traj = B; q.step_cap = (unsigned)scap; q.step_threads = (unsigned)std::min((std::max<size_t>(1, p_mx) + 31) / 32 * 32, scap);
        if (!materialise && !a.n_t_eval && !rows_are_steps) q.step_threads = 32;      // only the per-trajectory counters are left to write
        q.steps_taken = p.steps_taken;
        q.recs = p_steps.step_recs; q.rows_are_steps = rows_are_steps ? 1 : 0;
        q.args = p.args; q.args_stride = p.args_stride;
        q.t_span = p.t_span; q.t_span_stride = p.t_span_stride;
        q.t_eval = p.t_eval; q.n_t_eval = p.n_t_eval;
        if (a.n_t_eval || rows_are_steps) {
            if (h->d_trows.ensure(B * row_cap * 8) || h->d_yrows.ensure(B * row_cap * NDY * 8)) return -1;
            q.t_rows = h->d_trows.as<double>(); q.y_rows = h->d_yrows.as<double>(); q.row_cap = (unsigned)row_cap;
        }
        q.dense_output = dense ? 1 : 0;
        if (dense) {
            // The step records pass A wrote are all sol(t) needs: the interpolant of a step is a pure function of
            // (t_old, y_old, step), so it is rebuilt when it is queried (dense_eval_body) instead of being stored --
            // Lorenz/DOP853: 40 bytes per step kept instead of 216 written.  Option dense_store = 1 materialises them.
            if (h->d_dense_count.ensure(B * 4)) return -1;
            q.dense_count = h->d_dense_count.as<unsigned>(); q.dense_cap = (unsigned)dense_cap;
            if (materialise) {
                if (h->d_dense_t.ensure(B * dense_cap * 8) || h->d_dense_state.ensure(B * dense_cap * NY * (q_order + 1) * 8)) return -1;
                q.dense_t = h->d_dense_t.as<double>(); q.dense_state = h->d_dense_state.as<double>();
            } else {
                h->d_dense_t.release(); h->d_dense_state.release();       // (a reused handle may still hold them)
                h->dense_from_steps = true;
            }
        }
        q.nfev = p.nfev;
        CU_TRY(cudaEventRecord(h->ev_k0, h->stream));
        {
            const unsigned long long threads = (unsigned long long)B * q.step_threads;
            const int sblock = CYRK_STEP_DENSE_BLOCK;
            const unsigned long long sgrid = (threads + sblock - 1) / sblock;
            if (sgrid > 0x7fffffffull) { set_error("batch x step capacity too large for one launch: split the batch"); return -1; }
            if (pe->step_dense(a.method, q, (int)sgrid, sblock, h->stream) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("step_dense kernel launch failed"); return -1; }
            h->launches++;
        }
        if (a.n_t_eval) {
            teval_last_kernel<<<(unsigned)((B + 255) / 256), 256, 0, h->stream>>>(B, p.steps_taken, q.recs, q.step_cap, p.t_span, p.t_span_stride,
                                                                                (unsigned)CYRK_STEP_REC_DOUBLES(NY), p.t_eval, p.n_t_eval, q.t_rows, q.y_rows, q.row_cap, (unsigned)NDY,
                                                                                p.size, p.t_last, p.y_last);
            CU_TRY(cudaGetLastError());
            h->launches++;
        }
        CU_TRY(cudaEventRecord(h->ev_k1, h->stream));
        if (!device_io) {
            if (download(h, h->h_size, h->d_size, B * 4)) return -1;
            h->host_valid |= 1u << HA_SIZE;
            if (dense && download(h, h->h_dense_count, h->d_dense_count, B * 4)) return -1;
        }
        CU_TRY(cudaStreamSynchronize(h->stream));
        { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1); h->kernel_ms += ms; if (getenv("CYRK_B200_DEBUG")) fprintf(stderr, "[cyrk_b200] step_dense (+ last row): %.3f ms, %u threads per trajectory\n", ms, q.step_threads); }
    } else {
    size_t evq_cap = std::max<size_t>(B * 16, h->evq_seen + h->evq_seen / 8);
    int guard = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (++guard > 8) { set_error("internal: the events solve did not settle"); return -1; }
        if (want_rows) {
            if (h->d_trows.ensure(B * row_cap * 8) || h->d_yrows.ensure(B * row_cap * NDY * 8)) return -1;
            p.t_rows = h->d_trows.as<double>();
            p.y_rows = h->d_yrows.as<double>();
            p.row_cap = (unsigned)row_cap;
        }
        if (dense) {
            if (h->d_dense_t.ensure(B * dense_cap * 8) ||
                h->d_dense_state.ensure(B * dense_cap * NY * (q_order + 1) * 8) || h->d_dense_count.ensure(B * 4))
                return -1;
            p.dense_t = h->d_dense_t.as<double>();
            p.dense_state = h->d_dense_state.as<double>();
            p.dense_count = h->d_dense_count.as<unsigned>();
            p.dense_cap = (unsigned)dense_cap;
        }
        if (a.n_events) {
            if (h->d_ev_t.ensure(B * a.n_events * ev_cap * 8) || h->d_ev_y.ensure(B * a.n_events * ev_cap * NDY * 8) ||
                h->d_ev_count.ensure(B * a.n_events * 4) || h->d_ev_term.ensure(B * 4))
                return -1;
            p.ev_t = h->d_ev_t.as<double>();
            p.ev_y = h->d_ev_y.as<double>();
            p.ev_count = h->d_ev_count.as<unsigned>();
            p.ev_term_index = h->d_ev_term.as<int>();
            p.ev_cap = (unsigned)ev_cap;
        }
        CU_TRY(cudaMemsetAsync(h->d_queue.p, 0, 8, h->stream));
        CU_TRY(cudaEventRecord(h->ev_k0, h->stream));
        if (use_prep) {
            if (h->d_prep.ensure(B * (NDY + 2) * 8)) return -1;
            const int pgrid = (int)std::min<unsigned long long>((B + 255) / 256, (unsigned long long)device_sm_count(h->device) * 16);
            if (pe->prep(a.method, p, h->d_prep.as<double>(), pgrid, 256, h->stream) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("prep kernel launch failed"); return -1; }
            h->launches++;
            p.prep = h->d_prep.as<double>();
        }
        if (ev2) {
            if (h->d_evq.ensure(evq_cap * CYRK_EVQ_ITEM_DOUBLES(NY) * 8) || h->d_evq_count.ensure(16)) return -1;
            CU_TRY(cudaMemsetAsync(h->d_evq_count.p, 0, 16, h->stream));
            p.evq_items = h->d_evq.as<double>(); p.evq_count = h->d_evq_count.as<unsigned long long>(); p.evq_cap = evq_cap;
        }
        if (pe->launch(a.method, ev2 ? VAR_EVDETECT : variant, p, grid, block, h->stream) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("solve kernel launch failed"); return -1; }
        h->launches++;
        if (ev2) {
            // the number of parked steps sizes the second half; more than the queue holds -> deterministic re-run
            unsigned long long cnt[2] = {0, 0};
            CU_TRY(cudaMemcpyAsync(cnt, h->d_evq_count.p, 16, cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaStreamSynchronize(h->stream));
            if (getenv("CYRK_B200_DEBUG")) fprintf(stderr, "[cyrk_b200] event detection: %llu triggered steps (queue %zu), fallback flag %llu\n", cnt[0], evq_cap, cnt[1]);
            if (cnt[1] != 0) {             // a triggered step that was booked without stages: only the one-pass kernel reproduces it
                ev2 = false;
                if (pick_grid(variant)) return -1;
                pass--; continue;
            }
            if (cnt[0] > evq_cap) { evq_cap = (size_t)cnt[0]; pass--; continue; }
            h->evq_seen = (size_t)cnt[0];
            if (cnt[0] > 0) {
                const int rgrid = (int)std::min<unsigned long long>((cnt[0] + 127) / 128, (unsigned long long)device_sm_count(h->device) * 32);
                if (pe->event_roots(a.method, p, rgrid, 128, h->stream) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("event_roots kernel launch failed"); return -1; }
                h->launches++;
            }
        }
        CU_TRY(cudaEventRecord(h->ev_k1, h->stream));

        // counters come back first: they decide whether any arena overflowed
        const bool need_check = want_rows || dense || a.n_events;
        if (!need_check) break;
        unsigned long long evq_flag = 0;
        if (ev2) CU_TRY(cudaMemcpyAsync(&evq_flag, h->d_evq_count.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, h->stream));
        unsigned mxs[3] = {0, 0, 0};
        if (counts_max(h, want_rows ? p.size : nullptr, B, dense ? p.dense_count : nullptr, B, a.n_events ? p.ev_count : nullptr, B * a.n_events, mxs)) return -1;
        if (ev2 && evq_flag != 0) {        // a Brent search failed (-30 / -31): the one-pass kernel reproduces what the reference does next
            ev2 = false; h->kernel_ms = 0.f;
            if (pick_grid(variant)) return -1;
            pass--; continue;
        }
        {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1);
            h->kernel_ms += ms;
        }
        const size_t mx_rows = mxs[0], mx_dense = mxs[1], mx_ev = mxs[2];
        const bool overflow = mx_rows > row_cap || mx_dense > dense_cap || mx_ev > ev_cap;
        ev_used = mx_ev;
        if (!overflow) break;
        if (force.active && force.row_cap && mx_rows > row_cap) { force.need_row_cap = mx_rows; return 1; }
        if (pass == 1) { set_error("internal: arena overflow after the exact-capacity re-run"); return -1; }
        // deterministic re-run with exact capacities
        row_cap = std::max(row_cap, (force.active && !force.row_cap && mx_rows > row_cap) ? mx_rows + mx_rows / 16 + 8 : mx_rows);
        dense_cap = std::max(dense_cap, mx_dense);
        ev_cap = std::max(ev_cap, mx_ev);
    }
    }
    h->ev_packed = false; h->ev_total = 0;
    if (a.n_events && (!device_io || force.active)) {
        // pack the ragged hit lists back to back before they cross PCIe (C3: ~13 hits per trajectory in 3 x 64 slots)
        const unsigned long long NL = (unsigned long long)B * a.n_events;
        const unsigned nb = (unsigned)((NL + (unsigned long long)SCAN_BLOCK * SCAN_ITEMS - 1) / ((unsigned long long)SCAN_BLOCK * SCAN_ITEMS)) + 1u;   // (+1: the block that owns index NL)
        if (h->d_ev_off.ensure((NL + 1) * 8) || h->d_scan.ensure(((size_t)nb + 1) * 8) || h->h_ev_off.ensure((NL + 1) * 8)) return -1;
        unsigned long long* bsum = h->d_scan.as<unsigned long long>();
        scan_block_sums_kernel<<<nb, SCAN_BLOCK, 0, h->stream>>>(p.ev_count, NL, bsum);
        scan_block_offsets_kernel<<<1, 32, 0, h->stream>>>(bsum, nb, bsum + nb);
        const unsigned long long ev_first = force.active ? force.ev_base : 0ull;
        scan_write_kernel<<<nb, SCAN_BLOCK, 0, h->stream>>>(p.ev_count, NL, bsum, h->d_ev_off.as<unsigned long long>(), ev_first);
        CU_TRY(cudaGetLastError());
        unsigned long long total = 0;
        CU_TRY(cudaMemcpyAsync(&total, bsum + nb, 8, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        if (h->d_ev_t2.ensure(std::max<size_t>(total, 1) * 8) || h->d_ev_y2.ensure(std::max<size_t>(total, 1) * NDY * 8)) return -1;
        if (total > 0) {
            const unsigned used = (unsigned)std::max<size_t>(ev_used, 1);
            const unsigned long long work = NL * used;
            const int rgrid = (int)std::min<unsigned long long>((work + 255) / 256, (unsigned long long)device_sm_count(h->device) * 32);
            pack_events_kernel<<<rgrid, 256, 0, h->stream>>>(NL, (unsigned)ev_cap, used, (unsigned)NDY, p.ev_count, h->d_ev_off.as<unsigned long long>(), ev_first,
                                                             h->d_ev_t.as<double>(), h->d_ev_y.as<double>(),
                                                             h->d_ev_t2.as<double>(), h->d_ev_y2.as<double>());
            CU_TRY(cudaGetLastError());
        }
        h->launches += 4;
        h->ev_packed = true; h->ev_total = (size_t)total;
    }
    h->row_cap = row_cap; h->dense_cap = dense_cap; h->ev_cap = ev_cap;

    // ---- results to the host
    if (!device_io) {
        if (download(h, h->h_status, h->d_status, B * 4) || download(h, h->h_steps, h->d_steps, B * 4) ||
            download(h, h->h_ylast, h->d_ylast, B * NDY * 8))
            return -1;
        h->host_valid |= (1u << HA_STATUS) | (1u << HA_STEPS) | (1u << HA_YLAST);
        if ((want_rows || dense || a.n_events) && !two_pass) {
            if (download(h, h->h_size, h->d_size, B * 4)) return -1;
            h->host_valid |= 1u << HA_SIZE;
        }
        if (dense && !two_pass && download(h, h->h_dense_count, h->d_dense_count, B * 4)) return -1;
        if (a.n_events && download(h, h->h_ev_count, h->d_ev_count, B * a.n_events * 4)) return -1;
        if (want_rows) {
            if (download(h, h->h_trows, h->d_trows, B * row_cap * 8) || download(h, h->h_yrows, h->d_yrows, B * row_cap * NDY * 8))
                return -1;
        }
        if (a.n_events) {
            if (download(h, h->h_ev_off, h->d_ev_off, (B * a.n_events + 1) * 8) ||
                download(h, h->h_ev_t, h->d_ev_t2, h->ev_total * 8) ||
                download(h, h->h_ev_y, h->d_ev_y2, h->ev_total * NDY * 8) ||
                download(h, h->h_ev_term, h->d_ev_term, B * 4))
                return -1;
        }
    }
    CU_TRY(cudaEventRecord(h->ev_end, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    if (!(want_rows || dense || a.n_events)) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1);
        h->kernel_ms = ms;
    }
    cudaEventElapsedTime(&h->total_ms, h->ev_begin, h->ev_end);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Streamed solves: rows / t_eval samples / event lists / step records chunk by chunk.
//
// CySolverResult grows its storage as the integration proceeds and hands everything to the caller at the end
// (cysolution.cpp:396-448); a batch of millions of trajectories cannot keep every row in HBM at once (BASELINE config 2:
// 4 Mi x 1000 samples = 134 GB).  So the batch is cut into chunks: whole-batch inputs, the 44-byte per-trajectory results
// and -- with dense_output -- the 40-byte step records stay on the device in the parent handle; two sub-handles own one
// chunk-sized arena each and alternate, so while chunk k's rows cross PCIe into the parent's pinned arrays, chunk k+1 is
// being solved.  Every chunk runs the unmodified single-batch code (solve_thread_path on device pointers), so results
// are identical to the unchunked call by construction (tests/test_gpu_streaming.py).  The first chunk is a probe: it
// fixes the row / record capacity of the host layout; a later chunk that needs more makes the whole batch re-plan.
// Returns 0, -1, or 2 = "not streamable after all" (the caller falls back to the single-arena path).
// ------------------------------------------------------------------------------------------------
static int g_opt_stream_chunk = 0;       // trajectories per chunk (0 = by arena size)
static int g_opt_stream_taper = 1;       // two automatic halves become chunks of 1/2, 1/3, 1/6 (see solve_thread_streamed)
static int g_opt_stream = 1;             // 0 = never stream

static size_t stream_chunk_for(const SolveArgs& a, const Dispatch* pe)
{
    if (!g_opt_stream || (a.flags & CYRK_FLAG_DEVICE_IO)) return 0;
    const size_t B = a.n_traj, NY = pe->ny, NX = pe->nextra, NDY = NY + NX;
    const bool want_rows = a.store_mode == CYRK_STORE_ROWS, dense = a.dense_output != 0;
    if (!(want_rows || dense || a.n_events || a.n_t_eval)) return 0;         // final-state solves have their own pipeline
    // a dense store must outlive its chunk: only the step-record form (two-pass path) does
    if (dense && (a.n_events || NX || !g_opt_two_pass || (a.n_t_eval && (!want_rows || a.n_t_span > 1)))) return 0;
    const size_t rows = want_rows ? (a.n_t_eval ? a.n_t_eval : (a.expected_size ? a.expected_size : 256)) : 0;
    const size_t per_traj = rows * (1 + NDY) * 8 + a.n_events * (a.expected_size ? a.expected_size : 64) * (1 + NDY) * 8 +
                            ((a.n_t_eval || dense) ? (size_t)(a.expected_size ? a.expected_size + 32 : 288) * CYRK_STEP_REC_DOUBLES(NY) * 8 : 0) + 64;
    size_t chunk;
    if (g_opt_stream_chunk > 0) chunk = (size_t)g_opt_stream_chunk;
    else {
        if ((double)B * (double)per_traj < 1073741824.0) return 0;           // < 1 GiB of arena: one launch
        // Chunks as large as an eighth of the free device memory allows (two sub-handles hold one arena each; between 3 and
        // 16 GiB), at most half the batch: every chunk is a persistent kernel with its own drain phase, so fewer chunks
        // solve faster (C3full, 4 Mi trajectories end to end: 8 chunks 91.5 ms, 4: 87.0, 2: 85.5;
        // profiles/r02_c3full_e2e_chunks.log) while half the batch still overlaps its copy-out with the other half's solve.
        size_t free_b = 0, total_b = 0;
        double cap = 3221225472.0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) cap = std::min(17179869184.0, std::max(cap, (double)free_b / 8.0));
        else cudaGetLastError();
        const size_t max_chunk = std::max<size_t>((size_t)(cap / (double)per_traj), 16384);
        const size_t nch = std::max<size_t>(2, (B + max_chunk - 1) / max_chunk);          // equal chunks, at least two
        chunk = std::max<size_t>(((B + nch - 1) / nch + 1023) / 1024 * 1024, 16384);
        return B > chunk ? chunk : 0;
    }
    return (B >= 2 * chunk) ? chunk : 0;
}

static int solve_thread_streamed(Handle* h, const SolveArgs& a, const Dispatch* pe, size_t chunk)
{
    const size_t B = a.n_traj, NY = pe->ny, NX = pe->nextra, NDY = NY + NX, E = a.n_events;
    const bool want_rows = a.store_mode == CYRK_STORE_ROWS, dense = a.dense_output != 0;
    const size_t args_stride = (size_t)pe->nargs_dbl * 8;
    const bool per_args = a.n_args > 1 && pe->nargs_dbl > 0, per_span = a.n_t_span > 1;
    const size_t RA = CYRK_STEP_REC_DOUBLES(NY);
    const int q_order = a.method == CYRK_METHOD_RK23 ? 3 : a.method == CYRK_METHOD_RK45 ? 4 : 7;

    h->ny = NY; h->nx = NX; h->ndy = NDY; h->q_order = q_order;
    h->n_traj = B; h->n_events = E; h->n_t_eval = a.n_t_eval;
    h->dense_output = dense; h->device_io = false; h->dense_from_steps = false;
    h->ev_packed = false; h->ev_total = 0;
    h->h2d_bytes = h->d2h_bytes = h->lazy_d2h_bytes = 0; h->launches = 0; h->kernel_ms = 0.f;
    h->host_valid = 0;
    for (int s = 0; s < 2; s++) {
        if (!h->sub[s]) h->sub[s] = new Handle();
        if (h->sub[s]->bind(a.device)) return -1;
    }
    auto sync_subs = [&]() -> int {
        CU_TRY(cudaStreamSynchronize(h->sub[0]->stream));
        CU_TRY(cudaStreamSynchronize(h->sub[1]->stream));
        CU_TRY(cudaStreamSynchronize(h->s_h2d));          // (the copy-in of the later chunks' inputs: the caller owns that memory)
        return 0;
    };
    CU_TRY(cudaEventRecord(h->ev_begin, h->stream));

    // Chunk schedule.  Equal chunks in general; when the automatic rule settled on two halves (memory is not the limit),
    // three chunks of 1/2, 1/3 and 1/6 of the batch instead: the copy-out of the LAST chunk is the part no solve overlaps,
    // every chunk costs one drain phase, and a chunk's copy-out has to fit under the next chunk's solve (C3full, 4 Mi:
    // kernels 61.3 ms in one launch + ~2.7 ms per further chunk; 33 ms of copy-out in total; profiles/r02_notes.md)
    std::vector<std::pair<size_t, size_t>> sched;          // (offset, count)
    if (g_opt_stream_chunk <= 0 && g_opt_stream_taper && chunk < B && B <= 2 * chunk && B >= 6 * 16384) {
        const size_t n0 = (B / 2 + 1023) / 1024 * 1024, n1 = (B / 3 + 1023) / 1024 * 1024;
        sched.emplace_back(0, n0);
        sched.emplace_back(n0, n1);
        sched.emplace_back(n0 + n1, B - n0 - n1);
    } else {
        for (size_t off = 0; off < B; off += chunk) sched.emplace_back(off, std::min(chunk, B - off));
    }
    const size_t nchunk = sched.size();

    // ---- whole-batch inputs, uploaded once; the chunk solves read them in place.  Only the FIRST chunk's share is waited
    // for: the rest goes through the copy-in stream while the first chunk is solved (later chunks wait on `ev_rest`).
    const size_t n_first = sched[0].second;
    if (h->events(1)) return -1;
    cudaEvent_t ev_rest = h->ev_pool[0];
    if (h->d_y0.ensure(B * NY * 8)) return -1;
    CU_TRY(cudaMemcpyAsync(h->d_y0.p, a.y0, n_first * NY * 8, cudaMemcpyHostToDevice, h->stream));
    if (B > n_first)
        CU_TRY(cudaMemcpyAsync(h->d_y0.as<double>() + n_first * NY, a.y0 + n_first * NY, (B - n_first) * NY * 8, cudaMemcpyHostToDevice, h->s_h2d));
    h->h2d_bytes += B * NY * 8;
    if (pe->nargs_dbl > 0) {
        const size_t n_blobs = per_args ? B : 1;
        if (a.args_len == args_stride && a.n_args >= 1) {
            if (h->d_args.ensure(n_blobs * args_stride)) return -1;
            const size_t nb_first = per_args ? n_first : 1;
            CU_TRY(cudaMemcpyAsync(h->d_args.p, a.args, nb_first * args_stride, cudaMemcpyHostToDevice, h->stream));
            if (n_blobs > nb_first)
                CU_TRY(cudaMemcpyAsync(h->d_args.as<char>() + nb_first * args_stride, a.args + nb_first * args_stride,
                                       (n_blobs - nb_first) * args_stride, cudaMemcpyHostToDevice, h->s_h2d));
            h->h2d_bytes += n_blobs * args_stride;
        } else {                                                    // zero-pad every blob to the functor's size
            if (h->h_stage.ensure(n_blobs * args_stride)) return -1;
            char* st = h->h_stage.as<char>();
            memset(st, 0, n_blobs * args_stride);
            if (a.n_args >= 1)
                for (size_t i = 0; i < n_blobs; i++) memcpy(st + i * args_stride, a.args + i * a.args_len, a.args_len);
            if (upload(h, h->d_args, st, n_blobs * args_stride)) return -1;
        }
    } else if (h->d_args.ensure(8)) return -1;
    if (upload(h, h->d_tspan, a.t_span, a.n_t_span * 16)) return -1;
    CU_TRY(cudaEventRecord(ev_rest, h->s_h2d));
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->in_y0 = h->d_y0.as<double>(); h->in_args = h->d_args.as<char>(); h->in_tspan = h->d_tspan.as<double>();
    h->tspan_stride = per_span ? 2u : 0u; h->args_stride = per_args ? args_stride : 0;

    // ---- whole-batch per-trajectory results (device) and their eager host mirrors
    if (h->d_status.ensure(B * 4) || h->d_size.ensure(B * 4) || h->d_steps.ensure(B * 4) || h->d_nrej.ensure(B * 4) ||
        h->d_nfev.ensure(B * 4) || h->d_tlast.ensure(B * 8) || h->d_ylast.ensure(B * NDY * 8))
        return -1;
    if (h->h_status.ensure(B * 4) || h->h_steps.ensure(B * 4) || h->h_ylast.ensure(B * NDY * 8) || h->h_size.ensure(B * 4)) return -1;
    if (dense && (h->d_dense_count.ensure(B * 4) || h->h_dense_count.ensure(B * 4))) return -1;
    if (E && (h->d_ev_count.ensure(B * E * 4) || h->d_ev_term.ensure(B * 4) || h->h_ev_count.ensure(B * E * 4) ||
              h->h_ev_term.ensure(B * 4) || h->h_ev_off.ensure((B * E + 1) * 8)))
        return -1;

    // (pinned buffers that other chunks' copies may still be writing into are only replaced after both streams drained)
    auto grow_pinned = [&](PinBuf& b, size_t need, size_t keep) -> int {
        if (need <= b.cap) return 0;
        if (sync_subs()) return -1;
        PinBuf nb;
        if (nb.ensure(need)) return -1;
        if (keep && b.p) memcpy(nb.p, b.p, keep);
        b.release();
        b = nb;
        return 0;
    };

    size_t row_cap = 0, scap = 0;
    bool planned = false, use_recs = false;
    for (int plan = 0;; plan++) {
        if (plan >= 8) { set_error("streamed solve: capacity re-planning did not converge"); return -1; }
        bool replan = false;
        unsigned long long ev_base = 0;
        h->kernel_ms = 0.f; h->launches = 0; h->d2h_bytes = 0;
        if (planned) {                                             // (a re-plan: the layout grew)
            if (want_rows && (h->h_trows.ensure(B * row_cap * 8) || h->h_yrows.ensure(B * row_cap * NDY * 8))) return -1;
            if (use_recs && h->d_st.ensure(B * scap * RA * 8)) return -1;
        }
        for (size_t c = 0; c < nchunk && !replan; c++) {
            const size_t off = sched[c].first, n = sched[c].second;
            Handle* s = h->sub[c & 1];
            SolveArgs ac = a;
            ac.n_traj = n;
            ac.y0 = h->d_y0.as<double>() + off * NY;
            ac.args = h->d_args.as<char>() + (per_args ? off * args_stride : 0);
            ac.args_len = pe->nargs_dbl > 0 ? args_stride : 0;
            ac.n_args = pe->nargs_dbl > 0 ? (per_args ? n : 1) : 0;
            ac.t_span = h->d_tspan.as<double>() + (per_span ? 2 * off : 0);
            ac.n_t_span = per_span ? n : 1;
            ac.flags |= CYRK_FLAG_DEVICE_IO;
            if (a.flags & CYRK_FLAG_TOL_PER_TRAJ) { ac.rtols = a.rtols + off * NY; ac.atols = a.atols + off * NY; ac.n_rtols = ac.n_atols = n * NY; }
            if (a.n_t_eval && (a.flags & CYRK_FLAG_T_EVAL_PER_TRAJ)) ac.t_eval = a.t_eval + off * a.n_t_eval;
            if (c > 0) CU_TRY(cudaStreamWaitEvent(s->stream, ev_rest, 0));      // (the inputs beyond the first chunk)
            s->force = Handle::Force();
            s->force.active = true;
            s->force.ev_base = ev_base;
            if (planned) {
                s->force.row_cap = row_cap; s->force.scap = scap;
                if (use_recs) s->force.recs = h->d_st.as<double>() + off * scap * RA;
            }
            const int rc = solve_thread_path(s, ac, pe);          // (returns with this chunk's kernels finished)
            s->force.active = false;
            if (rc < 0) { sync_subs(); return -1; }
            if (rc == 1) {                                         // needs more rows / records than the plan has: re-plan
                if (s->force.need_row_cap) row_cap = std::max(row_cap, s->force.need_row_cap + s->force.need_row_cap / 8 + 1);
                if (s->force.need_scap) scap = std::max(scap, (s->force.need_scap + s->force.need_scap / 8 + 32) / 32 * 32);
                if (want_rows && !a.n_t_eval && scap) row_cap = scap;      // (two-pass path without t_eval: the rows ARE the records)
                replan = true;
                break;
            }
            if (!planned) {
                // probe chunk: its layout becomes the batch's
                if (dense && !s->dense_from_steps) { sync_subs(); return 2; }     // dense store would not outlive the chunk
                row_cap = s->row_cap;
                scap = s->took_two_pass ? s->step_cap : 0;
                use_recs = s->took_two_pass && dense;
                if (want_rows && (h->h_trows.ensure(B * row_cap * 8) || h->h_yrows.ensure(B * row_cap * NDY * 8))) return -1;
                if (use_recs) {
                    if (h->d_st.ensure(B * scap * RA * 8)) return -1;
                    CU_TRY(cudaMemcpyAsync(h->d_st.p, s->d_st.p, n * scap * RA * 8, cudaMemcpyDeviceToDevice, s->stream));
                }
                planned = true;
            }
            // ---- deliver this chunk (on ITS stream: the next solve on this sub-handle is ordered behind the copies)
#define D2D(name, T, per) CU_TRY(cudaMemcpyAsync(h->name.as<T>() + off * (per), s->name.as<T>(), n * (per) * sizeof(T), cudaMemcpyDeviceToDevice, s->stream))
#define D2H(host, dev, T, per) do { CU_TRY(cudaMemcpyAsync(h->host.as<T>() + off * (per), s->dev.as<T>(), n * (per) * sizeof(T), cudaMemcpyDeviceToHost, s->stream)); h->d2h_bytes += n * (per) * sizeof(T); } while (0)
            D2D(d_status, int, 1); D2D(d_size, unsigned, 1); D2D(d_steps, unsigned, 1); D2D(d_nrej, unsigned, 1);
            D2D(d_nfev, unsigned, 1); D2D(d_tlast, double, 1); D2D(d_ylast, double, NDY);
            D2H(h_status, d_status, int, 1); D2H(h_steps, d_steps, unsigned, 1); D2H(h_ylast, d_ylast, double, NDY);
            D2H(h_size, d_size, unsigned, 1);
            if (dense) { D2D(d_dense_count, unsigned, 1); D2H(h_dense_count, d_dense_count, unsigned, 1); }
            if (want_rows) { D2H(h_trows, d_trows, double, row_cap); D2H(h_yrows, d_yrows, double, row_cap * NDY); }
            if (E) {
                const size_t tot = s->ev_total;
                if (ev_base + tot > h->h_ev_t.cap / 8) {            // grow from what the chunks so far suggest
                    const double per_traj = (double)(ev_base + tot) / (double)(off + n);
                    const size_t want = std::max<size_t>((size_t)(per_traj * 1.25 * (double)B) + 1024, ev_base + tot);
                    if (grow_pinned(h->h_ev_t, want * 8, ev_base * 8) || grow_pinned(h->h_ev_y, want * NDY * 8, ev_base * NDY * 8)) return -1;
                }
                D2D(d_ev_count, unsigned, E); D2D(d_ev_term, int, 1);
                D2H(h_ev_count, d_ev_count, unsigned, E); D2H(h_ev_term, d_ev_term, int, 1);
                CU_TRY(cudaMemcpyAsync(h->h_ev_off.as<unsigned long long>() + off * E, s->d_ev_off.p, (n * E + 1) * 8, cudaMemcpyDeviceToHost, s->stream));
                if (tot) {
                    CU_TRY(cudaMemcpyAsync(h->h_ev_t.as<double>() + ev_base, s->d_ev_t2.p, tot * 8, cudaMemcpyDeviceToHost, s->stream));
                    CU_TRY(cudaMemcpyAsync(h->h_ev_y.as<double>() + ev_base * NDY, s->d_ev_y2.p, tot * NDY * 8, cudaMemcpyDeviceToHost, s->stream));
                }
                h->d2h_bytes += (n * E + 1) * 8 + tot * (1 + NDY) * 8;
                ev_base += tot;
            }
#undef D2D
#undef D2H
            h->kernel_ms += s->kernel_ms; h->launches += s->launches;
        }
        if (sync_subs()) return -1;
        if (!replan) { h->ev_total = (size_t)ev_base; break; }
    }
    h->row_cap = want_rows ? row_cap : 0;
    h->dense_cap = dense ? scap : 0;
    h->ev_cap = 0; h->ev_packed = E > 0;
    h->dense_from_steps = use_recs;
    h->step_cap = scap;
    h->host_valid = (1u << HA_STATUS) | (1u << HA_STEPS) | (1u << HA_YLAST) | (1u << HA_SIZE);
    CU_TRY(cudaEventRecord(h->ev_end, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->total_ms, h->ev_begin, h->ev_end);
    return 0;
}

// The CTA-per-trajectory path (large n_y; built-ins RHS_LARGEY / RHS_LARGEY_SIMPLE).  Returns 0 / -1.
static int solve_cta_path(Handle* h, const SolveArgs& a, int fp)
{
    const size_t B = a.n_traj, N = a.n_y;
    const bool device_io = (a.flags & CYRK_FLAG_DEVICE_IO) != 0;
    const bool want_rows = a.store_mode == CYRK_STORE_ROWS;
    const size_t NXL = a.num_extra;
    if (NXL != 0) {
        // Extra outputs on this path: the functor's components n_y .. n_y + num_extra - 1, evaluated at every saved row.
        // With t_eval, events or a dense store the reference re-evaluates the right-hand side at interpolated states and
        // lets that leak into its stepping state; those combinations stay on the thread path.
        if (a.rhs_id < CYRK_RHS_USER_BASE) { set_error("the built-in large problems have no extra outputs"); return -1; }
        if (a.n_t_eval || a.n_events || a.dense_output) { set_error("extra outputs on the CTA-per-trajectory path: raw step rows / final state only (no t_eval, events or dense_output)"); return -1; }
    }
    const bool tol_pt = (a.flags & CYRK_FLAG_TOL_PER_TRAJ) != 0;                       // rtols / atols: [n_traj][n_y]
    const bool te_pt = a.n_t_eval > 0 && (a.flags & CYRK_FLAG_T_EVAL_PER_TRAJ) != 0;   // t_eval: [n_traj][n_t_eval]
    const bool dense = a.dense_output != 0;
    const bool user_rhs = a.rhs_id >= CYRK_RHS_USER_BASE;
    if (!user_rhs && (a.n_args == 0 || a.args_len < 8)) { set_error("largey problems need args = (decay_rate, ...)"); return -1; }
    if (N > 0x7fffffffu) { set_error("n_y too large"); return -1; }
    h->ny = N; h->nx = NXL; h->ndy = N + NXL; h->q_order = 0;
    h->n_traj = B; h->n_events = a.n_events; h->n_t_eval = a.n_t_eval;
    h->dense_output = dense; h->device_io = device_io; h->cta_path = true; h->dense_from_steps = false;
    h->ev_packed = false; h->ev_total = 0;
    h->q_order = a.method == CYRK_METHOD_RK23 ? 3 : a.method == CYRK_METHOD_RK45 ? 4 : 7;
    h->h2d_bytes = h->d2h_bytes = h->lazy_d2h_bytes = 0; h->launches = 0; h->kernel_ms = 0.f;
    h->host_valid = 0;

    CtaParams p;
    memset(&p, 0, sizeof(p));
    p.n_traj = B;
    p.n = (unsigned)N;
    p.n_pad = (unsigned)((N + 1) & ~size_t(1));
    p.n_extra = (unsigned)NXL;
    p.max_step = a.max_step;
    p.first_step = a.first_step;
    {
        int uf; size_t mx;
        find_max_num_steps(N, 0, a.max_num_steps, a.max_ram_MB, &uf, &mx);
        p.user_max_steps = uf;
        p.max_num_steps = (unsigned)std::min<size_t>(mx, 0xffffffffu);
    }
    CU_TRY(cudaEventRecord(h->ev_begin, h->stream));
    // tolerances, expanded to n_y entries (one row, or with CYRK_FLAG_TOL_PER_TRAJ one row per trajectory); rtol floored
    // like the reference does in place (rk.cpp:600-608)
    const size_t n_tol = tol_pt ? B * N : N;
    std::vector<double> rt(n_tol), at(n_tol);
    for (size_t i = 0; i < n_tol; i++) {
        double r = tol_pt ? a.rtols[i] : a.rtols[a.n_rtols > 1 ? i : 0];
        if (r < 100.0 * std::numeric_limits<double>::epsilon()) r = 100.0 * std::numeric_limits<double>::epsilon();
        rt[i] = r;
        at[i] = tol_pt ? a.atols[i] : a.atols[a.n_atols > 1 ? i : 0];
    }
    if (h->d_rtol.ensure(n_tol * 8) || h->d_atol.ensure(n_tol * 8)) return -1;
    CU_TRY(cudaMemcpyAsync(h->d_rtol.p, rt.data(), n_tol * 8, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(h->d_atol.p, at.data(), n_tol * 8, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    p.rtol = h->d_rtol.as<double>();
    p.atol = h->d_atol.as<double>();
    p.tol_stride = tol_pt ? (unsigned)N : 0u;
    if (device_io) {
        p.y0 = a.y0; p.args = a.args; p.t_span = a.t_span;
    } else {
        if (upload(h, h->d_y0, a.y0, B * N * 8)) return -1;
        if (upload(h, h->d_args, a.args, a.n_args ? (a.n_args > 1 ? B : 1) * a.args_len : 0)) return -1;
        if (upload(h, h->d_tspan, a.t_span, a.n_t_span * 16)) return -1;
        p.y0 = h->d_y0.as<double>(); p.args = h->d_args.as<char>(); p.t_span = h->d_tspan.as<double>();
    }
    p.args_stride = a.n_args > 1 ? (unsigned)a.args_len : 0u;
    p.t_span_stride = a.n_t_span > 1 ? 2u : 0u;
    h->in_y0 = p.y0; h->in_args = p.args; h->in_tspan = p.t_span; h->tspan_stride = p.t_span_stride; h->args_stride = p.args_stride;
    std::vector<double> te;
    if (a.n_t_eval) {
        const size_t n_rows = te_pt ? B : 1;
        te.assign(a.t_eval, a.t_eval + n_rows * a.n_t_eval);
        double ts0[2];
        if (device_io) { CU_TRY(cudaMemcpy(ts0, a.t_span, 16, cudaMemcpyDeviceToHost)); }
        else { ts0[0] = a.t_span[0]; ts0[1] = a.t_span[1]; }
        // the kernels read an ascending table (reversed copy for backward runs, cysolver.cpp:520-527); validate() has
        // checked that every trajectory runs in the direction of the first one
        if (!((ts0[1] - ts0[0]) >= 0.0))
            for (size_t r = 0; r < n_rows; r++) std::reverse(te.begin() + r * a.n_t_eval, te.begin() + (r + 1) * a.n_t_eval);
        p.t_eval_stride = te_pt ? (unsigned)a.n_t_eval : 0u;
        if (h->d_teval.ensure(te.size() * 8)) return -1;
        CU_TRY(cudaMemcpyAsync(h->d_teval.p, te.data(), te.size() * 8, cudaMemcpyHostToDevice, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        p.t_eval = h->d_teval.as<double>();
        p.n_t_eval = (unsigned)a.n_t_eval;
    }
    if (h->d_status.ensure(B * 4) || h->d_size.ensure(B * 4) || h->d_steps.ensure(B * 4) || h->d_nrej.ensure(B * 4) ||
        h->d_nfev.ensure(B * 4) || h->d_tlast.ensure(B * 8) || h->d_ylast.ensure(B * (N + NXL) * 8) || h->d_queue.ensure(8))
        return -1;
    p.status = h->d_status.as<int>(); p.size = h->d_size.as<unsigned>(); p.steps_taken = h->d_steps.as<unsigned>();
    p.n_rejected = h->d_nrej.as<unsigned>(); p.nfev = h->d_nfev.as<unsigned>();
    p.t_last = h->d_tlast.as<double>(); p.y_last = h->d_ylast.as<double>();
    p.queue = h->d_queue.as<unsigned long long>();

    // Register-resident kernel (cta2_core.cuh) for built-in problems with n_y <= 8 * 512: thread t owns C components.
    int optin = 0;
    CU_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    int smem_per_sm = 0;
    CU_TRY(cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device));
    // Measured at n_y = 4096 (scripts/quick_c4_methods.py, round 2): the register-resident kernel wins for every method
    // (RK23 9.3 vs 18.4 ms, RK45 2.7 vs 5.4 ms, DOP853 3.6 vs 6.1 ms per 1184 trajectories), DOP853's nine slab-resident
    // stage vectors included.  Option cta_kernel = 1 forces the shared-memory/slab kernel.
    bool whole_vector = false;           // user system in the whole-vector form: one thread evaluates it (slab kernel only)
    if (user_rhs) { int uny = 0, unx = 0, una = 0, une = 0; if (user_rhs_info(a.rhs_id, &uny, &unx, &una, &une) == 0) whole_vector = uny > 0; }     // (registered with its n_y: the whole-vector form; component-wise systems register n_y = 0)
    // n_y in (4096, 12 288]: the same kernel with the stage derivatives in the L2-resident slab (RK23 / RK45); beyond
    // that, and for DOP853, the shared-memory/slab kernel of cta_core.cuh
    const bool cta2_slab_ok = N > (size_t)cta2_max_c_onchip * cta2_max_block && N <= (size_t)cta2_max_c * cta2_max_block &&
                              a.method != CYRK_METHOD_DOP853 && g_opt_cta2_large;
    // per-trajectory tolerance rows / t_eval rows: the shared-memory/slab kernel (any n_y) reads them through per-trajectory
    // pointers; the register-capped kernel is left alone (the row offsets cost its uniform-tolerance instantiations
    // 1.5 ... 4.7 % at n_y = 4096 even when unused, profiles/r02_notes.md)
    const bool use_cta2 = !whole_vector && g_opt_cta_kernel != 1 && !tol_pt && !te_pt &&
                          (N <= (size_t)cta2_max_c_onchip * cta2_max_block || cta2_slab_ok);
    int cta2_C = 0;
    int block, grid;
    size_t smem;
    {
        bool uni = !tol_pt;              // (per-trajectory rows are always read from the arrays)
        for (size_t i = 1; i < N && uni; i++) uni = (rt[i] == rt[0]) && (at[i] == at[0]);
        p.uniform_tol = uni ? 1 : 0;
        p.rtol0 = rt[0];
        p.atol0 = at[0];
    }
    if ((a.n_events || dense) && (tol_pt || te_pt)) {
        set_error("CTA-per-trajectory path: per-trajectory t_eval / tolerances cannot be combined with events or stored dense output");
        return -1;
    }
    if (a.n_events && !use_cta2) {
        set_error("events on the CTA-per-trajectory path need the register-resident kernel (n_y <= %d; RK23 / RK45: n_y <= %d)",
                  cta2_max_c_onchip * cta2_max_block, cta2_max_c * cta2_max_block);
        return -1;
    }
    p.n_events = (int)a.n_events;
    for (size_t e = 0; e < a.n_events; e++) {
        p.ev_id[e] = a.event_ids[e];
        p.ev_dir[e] = a.event_dirs ? a.event_dirs[e] : 0;
        p.ev_max[e] = a.event_max ? (unsigned long long)a.event_max[e] : ~0ull;
    }
    if (dense && !use_cta2) {
        set_error("stored dense output on the CTA-per-trajectory path needs the register-resident kernel (n_y <= %d; RK23 / RK45: n_y <= %d); use t_eval",
                  cta2_max_c_onchip * cta2_max_block, cta2_max_c * cta2_max_block);
        return -1;
    }
    if (use_cta2) {
        cta2_C = cta2_choose_c(N);
        block = cta2_block(N, cta2_C);
        const size_t V = (size_t)block * cta2_C;
        const int nkd = cta2_nkd(a.method);
        smem = (size_t)cta2_smem_vectors(nkd, cta2_C) * V * 8;
        if (smem + 1024 > (size_t)optin) { set_error("internal: CTA kernel shared-memory layout exceeds the device limit"); return -1; }
        // registers: the kernel is compiled for <= 128 per thread (__launch_bounds__(512, 1))
        const size_t by_regs = std::max<size_t>(1, (size_t)65536 / ((size_t)128 * block));
        const size_t by_smem = std::max<size_t>(1, (size_t)smem_per_sm / (smem + 2048));
        const int per_sm = (int)std::min<size_t>({(size_t)(g_opt_cta2_per_sm > 0 ? g_opt_cta2_per_sm : 32), by_regs, by_smem});
        grid = (int)std::min<unsigned long long>((unsigned long long)device_sm_count(h->device) * per_sm, B);
        if (h->d_scratch.ensure((size_t)grid * (size_t)cta2_slab_vectors(nkd, cta2_C) * V * 8)) return -1;
        p.scratch = h->d_scratch.as<double>();
    } else {
        // launch geometry: one CTA per SM when the working set fills shared memory, more when it is small
        const int nvec = cta_num_vectors(a.method);
        const size_t vec_bytes = (size_t)p.n_pad * 8;
        size_t fit = ((size_t)optin - 1024) / vec_bytes;
        // Measured (profiles/r01_notes.md, n_y = 4096): where the stage vectors live (shared memory vs the L2-resident
        // slab) changes the kernel time by < 2 %, but TWO 512-thread CTAs per SM beat one 1024-thread CTA by 10 %
        // (one CTA's barrier waits hide behind the other's arithmetic).  So: 512 threads, and only as many vectors in
        // shared memory as still leave room for two CTAs per SM.
        block = N <= 32 ? 32 : N <= 64 ? 64 : N <= 256 ? 128 : N <= 1024 ? 256 : 512;      // (<= 64: a warp or two per trajectory)
        if (g_opt_cta_block > 0) block = g_opt_cta_block;
        const size_t fit2 = ((size_t)smem_per_sm / 2 > 2048 + vec_bytes) ? ((size_t)smem_per_sm / 2 - 2048) / vec_bytes : 0;
        p.n_smem_vec = (unsigned)std::min<size_t>({fit, fit2, (size_t)nvec});
        if (g_opt_cta_smem_vecs >= 0) p.n_smem_vec = std::min<unsigned>((unsigned)std::min<size_t>(fit, (size_t)nvec), (unsigned)g_opt_cta_smem_vecs);
        smem = (size_t)p.n_smem_vec * vec_bytes;
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>({(size_t)(block <= 64 ? 32 : 8), (size_t)(2048 / block), ((size_t)smem_per_sm) / (smem + 2048)}));
        grid = (int)std::min<unsigned long long>((unsigned long long)device_sm_count(h->device) * per_sm, B);
        const size_t n_glob = (size_t)nvec - p.n_smem_vec;
        if (h->d_scratch.ensure((size_t)grid * (n_glob + 2) * vec_bytes)) return -1;
        p.scratch = h->d_scratch.as<double>();
    }

    size_t row_cap = 0, dense_cap = 0, ev_cap = 0;
    if (want_rows) row_cap = a.n_t_eval ? a.n_t_eval : (a.expected_size ? a.expected_size : 64);
    if (dense) dense_cap = a.expected_size ? a.expected_size : 64;
    if (a.n_events) ev_cap = a.expected_size ? a.expected_size : 8;      // one hit record is n_y + 1 doubles: start small
    for (int pass = 0; pass < 2; pass++) {
        if (want_rows) {
            if (h->d_trows.ensure(B * row_cap * 8) || h->d_yrows.ensure(B * row_cap * (N + NXL) * 8)) return -1;
            p.t_rows = h->d_trows.as<double>(); p.y_rows = h->d_yrows.as<double>(); p.row_cap = (unsigned)row_cap;
        }
        if (dense) {
            // one record per accepted step: [y_old | Q] = n_y * (Q_ORDER + 1) doubles (n_y = 4096, RK45: 164 KB)
            if (h->d_dense_t.ensure(B * dense_cap * 8) || h->d_dense_state.ensure(B * dense_cap * N * (h->q_order + 1) * 8) ||
                h->d_dense_count.ensure(B * 4) || h->h_dense_count.ensure(B * 4))
                return -1;
            p.dense_output = 1;
            p.dense_t = h->d_dense_t.as<double>(); p.dense_state = h->d_dense_state.as<double>();
            p.dense_count = h->d_dense_count.as<unsigned>(); p.dense_cap = (unsigned)dense_cap;
        }
        if (a.n_events) {
            if (h->d_ev_t.ensure(B * a.n_events * ev_cap * 8) || h->d_ev_y.ensure(B * a.n_events * ev_cap * N * 8) ||
                h->d_ev_count.ensure(B * a.n_events * 4) || h->d_ev_term.ensure(B * 4) || h->h_ev_count.ensure(B * a.n_events * 4))
                return -1;
            p.ev_t = h->d_ev_t.as<double>(); p.ev_y = h->d_ev_y.as<double>();
            p.ev_count = h->d_ev_count.as<unsigned>(); p.ev_term_index = h->d_ev_term.as<int>();
            p.ev_cap = (unsigned)ev_cap;
        }
        CU_TRY(cudaMemsetAsync(h->d_queue.p, 0, 8, h->stream));
        CU_TRY(cudaEventRecord(h->ev_k0, h->stream));
        if (user_rhs) {
            if (user_rhs_launch_cta(a.rhs_id, a.method, use_cta2 ? cta2_C : 0, fp == FP_STRICT, h->device, p, grid, block, smem, h->stream) != cudaSuccess) {
                cudaGetLastError();
                if (g_last_error.empty()) set_error("user CTA kernel launch failed");
                return -1;
            }
        } else {
            if (use_cta2) {
                CU_TRY((fp == FP_STRICT ? cta2_launch_strict : cta2_launch_fast)(a.rhs_id, a.method, cta2_C, p, grid, block, smem, h->stream));
            } else {
                CU_TRY((fp == FP_STRICT ? cta_launch_strict : cta_launch_fast)(a.rhs_id, a.method, p, grid, block, smem, h->stream));
            }
        }
        h->launches++;
        CU_TRY(cudaEventRecord(h->ev_k1, h->stream));
        if (want_rows) { if (download(h, h->h_size, h->d_size, B * 4)) return -1; h->host_valid |= 1u << HA_SIZE; }
        if (dense && download(h, h->h_dense_count, h->d_dense_count, B * 4)) return -1;
        if (a.n_events && download(h, h->h_ev_count, h->d_ev_count, B * a.n_events * 4)) return -1;
        CU_TRY(cudaStreamSynchronize(h->stream));
        { float ms = 0.f; cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1); h->kernel_ms += ms; }
        if (!want_rows && !dense && !a.n_events) break;
        size_t mx = 0, mxd = 0, mxe = 0;
        if (a.n_events) { const unsigned* ec = h->h_ev_count.as<unsigned>(); for (size_t i = 0; i < B * a.n_events; i++) mxe = std::max<size_t>(mxe, ec[i]); }
        if (want_rows) { const unsigned* sz = h->h_size.as<unsigned>(); for (size_t i = 0; i < B; i++) mx = std::max<size_t>(mx, sz[i]); }
        if (dense) { const unsigned* dc = h->h_dense_count.as<unsigned>(); for (size_t i = 0; i < B; i++) mxd = std::max<size_t>(mxd, dc[i]); }
        if (mx <= row_cap && mxd <= dense_cap && mxe <= ev_cap) break;
        if (pass == 1) { set_error("internal: arena overflow after the exact-capacity re-run"); return -1; }
        row_cap = std::max(row_cap, mx);
        dense_cap = std::max(dense_cap, mxd);
        ev_cap = std::max(ev_cap, mxe);
    }
    h->row_cap = row_cap; h->dense_cap = dense_cap; h->ev_cap = ev_cap;
    if (!device_io) {
        if (download(h, h->h_status, h->d_status, B * 4) || download(h, h->h_steps, h->d_steps, B * 4) ||
            download(h, h->h_ylast, h->d_ylast, B * (N + NXL) * 8))
            return -1;
        h->host_valid |= (1u << HA_STATUS) | (1u << HA_STEPS) | (1u << HA_YLAST);
        if (want_rows && (download(h, h->h_trows, h->d_trows, B * row_cap * 8) || download(h, h->h_yrows, h->d_yrows, B * row_cap * (N + NXL) * 8)))
            return -1;
        if (a.n_events && (download(h, h->h_ev_t, h->d_ev_t, B * a.n_events * ev_cap * 8) ||
                           download(h, h->h_ev_y, h->d_ev_y, B * a.n_events * ev_cap * N * 8) ||
                           download(h, h->h_ev_term, h->d_ev_term, B * 4)))
            return -1;
    }
    CU_TRY(cudaEventRecord(h->ev_end, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->total_ms, h->ev_begin, h->ev_end);
    return 0;
}

static int validate(const SolveArgs& a)
{
    if (a.n_traj == 0) { set_error("n_traj must be > 0"); return -1; }
    if (a.method < CYRK_METHOD_RK23 || a.method > CYRK_METHOD_DOP853) {
        set_error("unknown or unsupported integration method %d (supported: RK23=3, RK45=4, DOP853=5)", a.method);
        return -1;
    }
    if (!a.y0 || a.n_y == 0) { set_error("y0 is required and n_y must be > 0"); return -1; }
    if (!a.t_span || !(a.n_t_span == 1 || a.n_t_span == a.n_traj)) { set_error("t_span must hold 1 or n_traj (t_start, t_end) pairs"); return -1; }
    if (a.flags & CYRK_FLAG_TOL_PER_TRAJ) {
        if (!a.rtols || !a.atols || a.n_rtols != a.n_traj * a.n_y || a.n_atols != a.n_traj * a.n_y) {
            set_error("CYRK_FLAG_TOL_PER_TRAJ: rtols and atols must each hold n_traj * n_y entries");
            return -1;
        }
    } else if (!a.rtols || !a.atols || a.n_rtols == 0 || a.n_atols == 0 || (a.n_rtols > 1 && a.n_rtols != a.n_y) ||
        (a.n_atols > 1 && a.n_atols != a.n_y)) {
        set_error("rtols/atols must have 1 or n_y entries (CyrkErrorCodes::BAD_CONFIG_DATA, -22)");
        return -1;
    }
    if (!(a.n_args == 0 || a.n_args == 1 || a.n_args == a.n_traj)) { set_error("n_args must be 0, 1 or n_traj"); return -1; }
    if (a.n_args > 0 && a.args_len > 0 && !a.args) { set_error("args pointer is NULL"); return -1; }
    if (a.n_events > CYRK_MAX_EVENTS) { set_error("at most %d events are supported", CYRK_MAX_EVENTS); return -1; }
    if (a.n_events && !a.event_ids) { set_error("event_ids is NULL"); return -1; }
    for (size_t e = 0; e < a.n_events; e++)        // (both solver paths: the CTA path dispatches before the thread path's checks)
        if (a.event_ids[e] < 0 || (a.rhs_id < CYRK_RHS_USER_BASE && a.event_ids[e] >= EV_TABLE_SIZE)) { set_error("unknown event id %d", a.event_ids[e]); return -1; }
    if (a.n_t_eval && !a.t_eval) { set_error("t_eval is NULL"); return -1; }
    if (a.store_mode != CYRK_STORE_FINAL && a.store_mode != CYRK_STORE_ROWS) { set_error("unknown store_mode %d", a.store_mode); return -1; }
    if (a.max_step < 0.0) { set_error("max_step must be positive"); return -1; }
    const bool device_io = (a.flags & CYRK_FLAG_DEVICE_IO) != 0;
    // t_span: finite, and -- because t_eval is shared -- one direction for the whole batch.  (Device-resident t_span is
    // checked where the shim reads its first pair; NaN there gives every trajectory status -54 on its first attempt.)
    bool fwd0 = true;
    if (!device_io) {
        for (size_t i = 0; i < a.n_t_span; i++) {
            if (!std::isfinite(a.t_span[2 * i]) || !std::isfinite(a.t_span[2 * i + 1])) {
                set_error("t_span[%zu] = (%g, %g) is not finite", i, a.t_span[2 * i], a.t_span[2 * i + 1]);
                return -1;
            }
            const bool fwd = (a.t_span[2 * i + 1] - a.t_span[2 * i]) >= 0.0;
            if (i == 0) fwd0 = fwd;
            else if (a.n_t_eval && fwd != fwd0) {
                set_error("t_eval is shared by the batch: every t_span must integrate in the same direction");
                return -1;
            }
        }
    }
    const size_t n_te_rows = (a.flags & CYRK_FLAG_T_EVAL_PER_TRAJ) ? a.n_traj : 1;
    for (size_t r = 0; r < n_te_rows; r++) {
        const double* te = a.t_eval + r * a.n_t_eval;
        for (size_t i = 0; i < a.n_t_eval; i++) {
            if (!std::isfinite(te[i])) { set_error("t_eval[%zu] = %g is not finite", r * a.n_t_eval + i, te[i]); return -1; }
            if (i > 0 && !device_io && (fwd0 ? (te[i] < te[i - 1]) : (te[i] > te[i - 1]))) {
                set_error("t_eval must be ordered in the direction of integration (t_eval[%zu] = %g follows %g)", r * a.n_t_eval + i, te[i], te[i - 1]);
                return -1;
            }
        }
    }
    return 0;
}

static int solve_impl(Handle* h, const SolveArgs& a)
{
    h->solved = false;
    if (validate(a)) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: cyrk_b200 has no CPU fallback");
        return -1;
    }
    if (a.device < 0 || a.device >= ndev) { set_error("device %d out of range (%d devices)", a.device, ndev); return -1; }
    DeviceGuard guard;
    if (h->bind(a.device)) return -1;
    registry_init();
    const int fp = (a.flags & CYRK_FLAG_STRICT_FP) ? FP_STRICT : FP_FAST;
    h->rhs_id = a.rhs_id; h->method = a.method; h->fp_mode = fp; h->cta_path = false;

    if (a.rhs_id == CYRK_RHS_LARGEY || a.rhs_id == CYRK_RHS_LARGEY_SIMPLE) {
        if (solve_cta_path(h, a, fp)) return -1;
        h->solved = true;
        return 0;
    }
    Dispatch disp;
    g_last_error.clear();
    if (!make_dispatch(a.rhs_id, fp, a.device, &disp)) { set_error("unknown right-hand side id %d", a.rhs_id); return -1; }
    if (a.rhs_id >= CYRK_RHS_USER_BASE && (disp.ny == 0 || disp.ny > CYRK_MAX_NY)) {     // user right-hand side registered for large systems
        if (disp.ny != 0 && (size_t)disp.ny != a.n_y) { set_error("right-hand side %d has n_y = %d but y0 has n_y = %zu", a.rhs_id, disp.ny, a.n_y); return -1; }
        if ((size_t)disp.nextra != a.num_extra) {
            set_error("right-hand side %d produces %d extra outputs but num_extra = %zu", a.rhs_id, disp.nextra, a.num_extra);
            return -1;
        }
        if (solve_cta_path(h, a, fp)) return -1;
        h->solved = true;
        return 0;
    }
    // DOP853 with 16 < n_y <= 32 state components: one thread holds 16 x n_y stage derivatives (up to 4.3 KB of local
    // memory; 8.7 % of the FP64 peak at n_y = 32).  Option small_cta = 1 runs a WARP per trajectory instead (the
    // CTA-per-trajectory kernel of cta_core.cuh with 32-thread CTAs, the whole-vector function evaluated by lane 0, the stage
    // arithmetic spread over the lanes).  Measured (profiles/r02_medium_ny_small_cta.log): 28.6 vs 7.2 ms at n_y = 32,
    // 29.8 vs 3.4 ms at n_y = 24 -- the serial right-hand side on one lane and that kernel's barrier per loop cost far more
    // than the spills -- so the default stays the thread-per-trajectory kernel.
    if (a.rhs_id >= CYRK_RHS_USER_BASE && g_opt_small_cta && a.method == CYRK_METHOD_DOP853 && disp.ny > 16 && disp.ny <= CYRK_MAX_NY &&
        (size_t)disp.ny == a.n_y && a.n_events == 0 && !a.dense_output && a.num_extra == 0 && disp.nextra == 0 &&
        !(a.flags & (CYRK_FLAG_T_EVAL_PER_TRAJ | CYRK_FLAG_TOL_PER_TRAJ)) &&
        (a.n_args == 0 ? disp.nargs_dbl == 0 : a.args_len == (size_t)disp.nargs_dbl * 8)) {
        if (solve_cta_path(h, a, fp)) return -1;
        h->solved = true;
        return 0;
    }
    if ((a.flags & CYRK_FLAG_TOL_PER_TRAJ) && disp.use_feature("-DCYRK_TOL_PER_TRAJ=1")) return -1;
    const Dispatch* pe = &disp;
    if ((size_t)pe->ny != a.n_y) { set_error("right-hand side %d has n_y = %d but y0 has n_y = %zu", a.rhs_id, pe->ny, a.n_y); return -1; }
    if ((size_t)pe->nextra != a.num_extra) {
        set_error("right-hand side %d produces %d extra outputs but num_extra = %zu", a.rhs_id, pe->nextra, a.num_extra);
        return -1;
    }
    if (a.n_args > 0 && a.args_len > (size_t)pe->nargs_dbl * 8) {
        set_error("args blob of %zu bytes exceeds the %d bytes right-hand side %d reads", a.args_len, pe->nargs_dbl * 8, a.rhs_id);
        return -1;
    }
    if (const size_t chunk = stream_chunk_for(a, pe)) {
        const int rc = solve_thread_streamed(h, a, pe, chunk);
        if (rc < 0) return -1;
        if (rc == 0) { h->solved = true; return 0; }
    }
    if (solve_thread_path(h, a, pe)) return -1;
    h->solved = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Ensemble sums: per-CTA partials in a fixed order, then one CTA folds the partials (deterministic).
// ------------------------------------------------------------------------------------------------
__global__ void ensemble_partial_kernel(const int* __restrict__ status, const double* __restrict__ y_last,
                                        unsigned long long n, int ndy, double* __restrict__ partial)
{
    extern __shared__ double sh[];          // [blockDim.x / 32][2 * ndy + 1]
    const int nv = 2 * ndy + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const unsigned long long per = (n + gridDim.x - 1) / gridDim.x;
    const unsigned long long lo = (unsigned long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    for (int v = 0; v < nv; v++) {
        double acc = 0.0;
        for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const int st = status[i];
            if (st == 1 || st == 2) {
                if (v == 2 * ndy) acc += 1.0;
                else { const double y = y_last[i * ndy + (v % ndy)]; acc += v < ndy ? y : y * y; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) sh[warp * nv + v] = acc;
    }
    __syncthreads();
    if (threadIdx.x < nv) {
        double acc = 0.0;
        for (int w = 0; w < nwarp; w++) acc += sh[w * nv + threadIdx.x];
        partial[(size_t)blockIdx.x * nv + threadIdx.x] = acc;
    }
}

__global__ void ensemble_final_kernel(const double* __restrict__ partial, int nblocks, int nv, double* __restrict__ out)
{
    if ((int)threadIdx.x < nv) {
        double acc = 0.0;
        for (int b = 0; b < nblocks; b++) acc += partial[(size_t)b * nv + threadIdx.x];
        out[threadIdx.x] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// FP64 throughput probe: 8 independent DFMA chains per thread
// ------------------------------------------------------------------------------------------------
// the production build's sin/cos (rk_core.cuh psin/pcos), one thread per element, for the accuracy test
__global__ void sincos_probe_kernel(const double* __restrict__ x, double* __restrict__ s, double* __restrict__ c, size_t n, int mode)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mode == 0) { s[i] = cyrk::fast::psin(x[i]); c[i] = cyrk::fast::pcos(x[i]); }
    else { s[i] = cyrk::fast::psin_pi(x[i]); c[i] = cyrk::fast::psin_fast(x[i]); }     // the check-free variants (|x| <= 1e5)
}

__global__ void fp64_probe_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace cyrk

using namespace cyrk;

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

int cyrk_b200_version(void) { return 100; }

int cyrk_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* cyrk_b200_last_error(void) { return g_last_error.c_str(); }

void cyrk_b200_status_message(int code, char* buffer, size_t max_len)
{
    // message table of CyRK/cy/c_common.hpp:84-177
    const char* m = "UNKNOWN_ERROR";
    switch (code) {
    case 20: m = "An optimization routine has successfully converged."; break;
    case 10: m = "Initializing. If you see this message then it was likely interrupted."; break;
    case 2: m = "Integration ended early: An event has been triggered the maximum allowed times. No issues detected."; break;
    case 1: m = "Integration completed without issue."; break;
    case 0: m = "No errors were encountered."; break;
    case -1: m = "An unspecified error has occurred."; break;
    case -2: m = "A required property was not set."; break;
    case -3: m = "The provided model is currently unsupported or not known."; break;
    case -4: m = "A class object was not fully initialized."; break;
    case -5: m = "Error in CySolver initialization."; break;
    case -6: m = "The provided input is incompatible with other input or the current object(s) state."; break;
    case -7: m = "Error when accessing or modifying one or more attributes/properties."; break;
    case -8: m = "A bounds error was encountered."; break;
    case -9: m = "A required function or method argument was not set."; break;
    // -10, -30, -31 and -60 have no entry in the reference's table (its flat API answers UNKNOWN_ERROR, its
    // Python layer's map::at throws); this library does emit -10 (call_vectorize), -30 and -31 (event root search)
    case -10: m = "A function or method argument is out of range or otherwise invalid."; break;
    case -30: m = "Root finding failed: the function has the same sign at both ends of the bracket."; break;
    case -31: m = "Root finding failed: the maximum number of iterations was reached without convergence."; break;
    case -60: m = "Event setup failed."; break;
    case -20: m = "An object's additional setup function was not called before its methods were used."; break;
    case -21: m = "Can't call solution: Dense output was not saved."; break;
    case -22: m = "Error during setup: Provided configuration data does not make sense or is missing required parameters."; break;
    case -41: m = "A C++ vector object's size exceeds limits imposed by user or architecture."; break;
    case -40: m = "There was an error while allocating memory."; break;
    case -50: m = "The number of dependent `y` values is 0. There are no differential equations to solve."; break;
    case -51: m = "The maximum number of iterations was hit."; break;
    case -52: m = "Maximum number of steps (set by user) exceeded during integration."; break;
    case -53: m = "Maximum number of steps (set by system architecture) exceeded during integration."; break;
    case -54: m = "Error in step size calculation: Required step size is less than spacing between numbers."; break;
    case -55: m = "Error in step size calculation: Error in step size acceptance."; break;
    case -56: m = "Error during dense output build."; break;
    case -57: m = "Can't call solution: Integration has not been completed or it was unsuccessful."; break;
    case -70: m = "There was an error in the C++ backend when trying to import the required Python module."; break;
    case -80: m = "User-provided initial step size must be a positive number."; break;
    case -99: m = "An unknown error occurred."; break;
    case -100: m = "The error code was never set."; break;
    default: break;
    }
    if (!buffer || max_len == 0) return;
    strncpy(buffer, m, max_len - 1);
    buffer[max_len - 1] = '\0';
}

static const struct { const char* name; int id; } k_builtin_names[] = {
    {"baseline", CYRK_RHS_BASELINE}, {"accuracy", CYRK_RHS_ACCURACY}, {"extraoutput", CYRK_RHS_EXTRAOUTPUT},
    {"lorenz", CYRK_RHS_LORENZ}, {"lorenz_extraoutput", CYRK_RHS_LORENZ_EXTRA}, {"lotkavolterra", CYRK_RHS_LOTKAVOLTERRA},
    {"pendulum", CYRK_RHS_PENDULUM}, {"arbitrary_args", CYRK_RHS_ARBITRARY_ARGS}, {"pendulum_preeval", CYRK_RHS_PENDULUM_PREEVAL},
    {"largey", CYRK_RHS_LARGEY}, {"largey_simple", CYRK_RHS_LARGEY_SIMPLE}, {"pendulum_drag", CYRK_RHS_PENDULUM_DRAG},
    {"dense_kat", CYRK_RHS_DENSE_KAT}};

int cyrk_b200_rhs_builtin(const char* name)
{
    if (!name) return -1;
    for (const auto& e : k_builtin_names)
        if (strcmp(e.name, name) == 0) return e.id;
    set_error("unknown built-in right-hand side '%s'", name);
    return -1;
}

int cyrk_b200_rhs_info(int rhs_id, size_t* n_y, size_t* n_extra, size_t* args_bytes)
{
    registry_init();
    if (rhs_id == CYRK_RHS_LARGEY || rhs_id == CYRK_RHS_LARGEY_SIMPLE) {
        if (n_y) *n_y = 0;
        if (n_extra) *n_extra = 0;
        if (args_bytes) *args_bytes = 16;
        return 0;
    }
    if (rhs_id >= CYRK_RHS_USER_BASE) {
        int a = 0, b = 0, c = 0, d = 0;
        if (user_rhs_info(rhs_id, &a, &b, &c, &d)) { set_error("unknown right-hand side id %d", rhs_id); return -1; }
        if (n_y) *n_y = (size_t)a;
        if (n_extra) *n_extra = (size_t)b;
        if (args_bytes) *args_bytes = (size_t)c * 8;
        return 0;
    }
    const ProblemEntry* pe = registry_lookup(FP_FAST, rhs_id);
    if (!pe) { set_error("unknown right-hand side id %d", rhs_id); return -1; }
    if (n_y) *n_y = pe->ny;
    if (n_extra) *n_extra = pe->nextra;
    if (args_bytes) *args_bytes = (size_t)pe->nargs_dbl * 8;
    return 0;
}

int cyrk_b200_rhs_compile(const char* source, const char* entry, size_t n_y, size_t n_extra, size_t args_bytes,
                          size_t n_event_funcs, char* log, size_t log_len)
{
    return user_rhs_register(source, entry, n_y, n_extra, args_bytes, n_event_funcs, log, log_len);
}

int cyrk_b200_jit_compile(int rhs_id, int method, int variant, unsigned flags)
{
    // Compile-only: warms the process and disk caches for a run-time compiled kernel (user right-hand side, or the
    // per-trajectory-tolerance build of a built-in problem).  Needs NVRTC, not a GPU.
    registry_init();
    if (method < CYRK_METHOD_RK23 || method > CYRK_METHOD_DOP853 || variant < 0 || variant >= VAR_COUNT) { set_error("bad method / variant"); return -1; }
    const bool strict = (flags & CYRK_FLAG_STRICT_FP) != 0;
    const char* defs = (flags & CYRK_FLAG_TOL_PER_TRAJ) ? "-DCYRK_TOL_PER_TRAJ=1" : "";
    int id = rhs_id;
    if (rhs_id < CYRK_RHS_USER_BASE) {
        const ProblemEntry* pe = registry_lookup(strict ? FP_STRICT : FP_FAST, rhs_id);
        if (!pe) { set_error("right-hand side %d has no thread-per-trajectory kernels", rhs_id); return -1; }
        id = builtin_jit_id(pe->struct_name, pe->ny, pe->nextra, pe->nargs_dbl);
    }
    g_last_error.clear();
    return jit_compile_only(id, "solve", method, variant, strict, defs);
}

void cyrk_b200_jit_cache_stats(unsigned long long* compiled, unsigned long long* disk_hits) { jit_cache_stats(compiled, disk_hits); }

void* cyrk_b200_result_create(void) { return new Handle(); }

void cyrk_b200_free(void* h) { delete static_cast<Handle*>(h); }

int cyrk_b200_solve_batch_noreturn(
    void* handle, int rhs_id, int method, size_t n_traj, const double* t_span, size_t n_t_span, const double* y0,
    size_t n_y, size_t num_extra, const char* args, size_t args_len, size_t n_args, size_t max_num_steps,
    size_t max_ram_MB, int dense_output, const double* t_eval, size_t n_t_eval, const int* event_ids,
    const int* event_dirs, const size_t* event_max, size_t n_events, const double* rtols, size_t n_rtols,
    const double* atols, size_t n_atols, double max_step, double first_step, size_t expected_size, int store_mode,
    unsigned flags, int device)
{
    if (!handle) { set_error("handle is NULL"); return -1; }
    SolveArgs a{rhs_id, method, n_traj, t_span, n_t_span, y0, n_y, num_extra, args, args_len, n_args, max_num_steps,
                max_ram_MB, dense_output, t_eval, n_t_eval, event_ids, event_dirs, event_max, n_events, rtols, n_rtols,
                atols, n_atols, max_step, first_step, expected_size, store_mode, flags, device};
    return solve_impl(static_cast<Handle*>(handle), a);
}

void* cyrk_b200_solve_batch(
    int rhs_id, int method, size_t n_traj, const double* t_span, size_t n_t_span, const double* y0, size_t n_y,
    size_t num_extra, const char* args, size_t args_len, size_t n_args, size_t max_num_steps, size_t max_ram_MB,
    int dense_output, const double* t_eval, size_t n_t_eval, const int* event_ids, const int* event_dirs,
    const size_t* event_max, size_t n_events, const double* rtols, size_t n_rtols, const double* atols, size_t n_atols,
    double max_step, double first_step, size_t expected_size, int store_mode, unsigned flags, int device)
{
    Handle* h = new Handle();
    if (cyrk_b200_solve_batch_noreturn(h, rhs_id, method, n_traj, t_span, n_t_span, y0, n_y, num_extra, args, args_len,
                                       n_args, max_num_steps, max_ram_MB, dense_output, t_eval, n_t_eval, event_ids,
                                       event_dirs, event_max, n_events, rtols, n_rtols, atols, n_atols, max_step,
                                       first_step, expected_size, store_mode, flags, device)) {
        delete h;
        return nullptr;
    }
    return h;
}

int cyrk_b200_solve_batch_sharded(
    const int* device_ids, size_t n_dev, void** handles_out, size_t* bounds_out,
    int rhs_id, int method, size_t n_traj, const double* t_span, size_t n_t_span, const double* y0, size_t n_y,
    size_t num_extra, const char* args, size_t args_len, size_t n_args, size_t max_num_steps, size_t max_ram_MB,
    int dense_output, const double* t_eval, size_t n_t_eval, const int* event_ids, const int* event_dirs,
    const size_t* event_max, size_t n_events, const double* rtols, size_t n_rtols, const double* atols, size_t n_atols,
    double max_step, double first_step, size_t expected_size, int store_mode, unsigned flags)
{
    if (!device_ids || !handles_out || n_dev == 0) { set_error("device_ids, handles_out and n_dev > 0 are required"); return -1; }
    if (flags & CYRK_FLAG_DEVICE_IO) { set_error("the sharded call takes host buffers (each shard is copied to its own device)"); return -1; }
    if (n_traj < n_dev) { set_error("fewer trajectories (%zu) than devices (%zu)", n_traj, n_dev); return -1; }
    // contiguous shards; the first n_traj % n_dev get one extra (== cyrk_b200.sharding.shard_bounds)
    std::vector<size_t> lo(n_dev + 1, 0);
    for (size_t r = 0; r < n_dev; r++) lo[r + 1] = lo[r] + n_traj / n_dev + (r < n_traj % n_dev ? 1 : 0);
    std::vector<Handle*> hs(n_dev, nullptr);
    std::vector<int> rc(n_dev, 0);
    std::vector<std::string> err(n_dev);
    std::vector<std::thread> th;
    const bool per_span = n_t_span > 1, per_args = n_args > 1;
    for (size_t r = 0; r < n_dev; r++) {
        hs[r] = new Handle();
        th.emplace_back([&, r]() {
            const size_t o = lo[r], n = lo[r + 1] - lo[r];
            SolveArgs a{rhs_id, method, n, per_span ? t_span + 2 * o : t_span, per_span ? n : n_t_span, y0 + o * n_y, n_y,
                        num_extra, per_args ? args + o * args_len : args, args_len, per_args ? n : n_args, max_num_steps,
                        max_ram_MB, dense_output, t_eval, n_t_eval, event_ids, event_dirs, event_max, n_events, rtols,
                        n_rtols, atols, n_atols, max_step, first_step, expected_size, store_mode, flags, device_ids[r]};
            rc[r] = solve_impl(hs[r], a);
            if (rc[r]) err[r] = g_last_error;          // thread-local: carry it to the caller's thread
        });
    }
    for (auto& t : th) t.join();
    for (size_t r = 0; r < n_dev; r++) {
        if (rc[r]) {
            const std::string msg = err[r];
            for (size_t k = 0; k < n_dev; k++) delete hs[k];
            set_error("shard %zu (device %d): %s", r, device_ids[r], msg.c_str());
            return -1;
        }
    }
    for (size_t r = 0; r < n_dev; r++) handles_out[r] = hs[r];
    if (bounds_out) for (size_t r = 0; r <= n_dev; r++) bounds_out[r] = lo[r];
    return 0;
}

#define H(h) (static_cast<Handle*>(h))
#define PICK(h, dev, host, T) (H(h)->device_io ? static_cast<const T*>(H(h)->dev.p) : static_cast<const T*>(H(h)->host.p))
size_t cyrk_b200_get_n_traj(void* h) { return H(h)->n_traj; }
size_t cyrk_b200_get_num_y(void* h) { return H(h)->ny; }
size_t cyrk_b200_get_num_dy(void* h) { return H(h)->ndy; }
const int32_t* cyrk_b200_get_status_ptr(void* h) { return static_cast<const int32_t*>(host_array(H(h), HA_STATUS)); }
const uint32_t* cyrk_b200_get_size_ptr(void* h) { return static_cast<const uint32_t*>(host_array(H(h), HA_SIZE)); }
const uint32_t* cyrk_b200_get_steps_taken_ptr(void* h) { return static_cast<const uint32_t*>(host_array(H(h), HA_STEPS)); }
const uint32_t* cyrk_b200_get_n_rejected_ptr(void* h) { return static_cast<const uint32_t*>(host_array(H(h), HA_NREJ)); }
const uint32_t* cyrk_b200_get_nfev_ptr(void* h) { return static_cast<const uint32_t*>(host_array(H(h), HA_NFEV)); }
const double* cyrk_b200_get_t_last_ptr(void* h) { return static_cast<const double*>(host_array(H(h), HA_TLAST)); }
const double* cyrk_b200_get_y_last_ptr(void* h) { return static_cast<const double*>(host_array(H(h), HA_YLAST)); }
size_t cyrk_b200_get_row_capacity(void* h) { return H(h)->row_cap; }
const double* cyrk_b200_get_t_ptr(void* h) { return H(h)->row_cap ? PICK(h, d_trows, h_trows, double) : nullptr; }
const double* cyrk_b200_get_y_ptr(void* h) { return H(h)->row_cap ? PICK(h, d_yrows, h_yrows, double) : nullptr; }
size_t cyrk_b200_get_num_events(void* h) { return H(h)->n_events; }
size_t cyrk_b200_get_event_capacity(void* h) { return H(h)->ev_packed ? 0 : H(h)->ev_cap; }
const uint64_t* cyrk_b200_get_event_offsets_ptr(void* h) { return (H(h)->n_events && H(h)->ev_packed) ? static_cast<const uint64_t*>(H(h)->h_ev_off.p) : nullptr; }
size_t cyrk_b200_get_event_total(void* h) { return H(h)->ev_total; }
const uint32_t* cyrk_b200_get_event_count_ptr(void* h) { return H(h)->n_events ? PICK(h, d_ev_count, h_ev_count, uint32_t) : nullptr; }
const double* cyrk_b200_get_event_t_ptr(void* h) { return H(h)->n_events ? PICK(h, d_ev_t, h_ev_t, double) : nullptr; }
const double* cyrk_b200_get_event_y_ptr(void* h) { return H(h)->n_events ? PICK(h, d_ev_y, h_ev_y, double) : nullptr; }
const int32_t* cyrk_b200_get_event_terminate_index_ptr(void* h) { return H(h)->n_events ? PICK(h, d_ev_term, h_ev_term, int32_t) : nullptr; }
size_t cyrk_b200_get_num_interpolates_capacity(void* h) { return H(h)->dense_cap; }
const uint32_t* cyrk_b200_get_num_interpolates_ptr(void* h) { return H(h)->dense_output ? PICK(h, d_dense_count, h_dense_count, uint32_t) : nullptr; }

int cyrk_b200_call_vectorize(void* hv, const double* t, size_t n_t, double* out, int32_t* codes)
{
    Handle* h = H(hv);
    if (!h || !h->solved) { set_error("handle holds no solution"); return -1; }
    if (!t || !out || !codes || n_t == 0) { set_error("t, out and codes are required"); return -1; }
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(h->device));
    const size_t B = h->n_traj, NDY = h->ndy;
    if (!h->dense_output) {
        // CySolverResult::call: -57 when the trajectory did not succeed, else -21 (cysolution.cpp:513-520)
        std::vector<int32_t> st(B);
        if (h->device_io) {
            CU_TRY(cudaMemcpy(st.data(), h->d_status.p, B * 4, cudaMemcpyDeviceToHost));
        } else {
            memcpy(st.data(), h->h_status.p, B * 4);
        }
        for (size_t i = 0; i < B; i++) st[i] = (st[i] == 1 || st[i] == 2) ? ST_DENSE_OUTPUT_NOT_SAVED : ST_INTEGRATION_NOT_SUCCESSFUL;
        if (h->device_io) { CU_TRY(cudaMemcpy(codes, st.data(), B * 4, cudaMemcpyHostToDevice)); }
        else memcpy(codes, st.data(), B * 4);
        return 0;
    }
    if (h->cta_path) {
        // CTA-path solution: coefficient-major records, one CTA per (trajectory, query time)
        CtaDenseEvalParams q;
        memset(&q, 0, sizeof(q));
        q.n_traj = B; q.n = (unsigned)h->ny; q.n_query = (unsigned)n_t;
        if (h->d_q_t.ensure(n_t * 8) || h->d_q_code.ensure(B * 4)) return -1;
        CU_TRY(cudaMemcpyAsync(h->d_q_t.p, t, n_t * 8, h->device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        q.t_query = h->d_q_t.as<double>();
        q.t_span = h->in_tspan; q.t_span_stride = (unsigned)h->tspan_stride;
        q.status = h->d_status.as<int>();
        q.dense_t = h->d_dense_t.as<double>(); q.dense_state = h->d_dense_state.as<double>();
        q.dense_count = h->d_dense_count.as<unsigned>(); q.dense_cap = (unsigned)h->dense_cap;
        double* d_out = out;
        if (!h->device_io) {
            if (h->d_q_out.ensure(B * n_t * NDY * 8)) return -1;
            d_out = h->d_q_out.as<double>();
        }
        q.out = d_out;
        q.code = h->device_io ? codes : h->d_q_code.as<int>();
        CU_TRY(cudaMemsetAsync(q.code, 0, B * 4, h->stream));
        const unsigned long long pairs = (unsigned long long)B * n_t;
        if (pairs > 0x7fffffffull) { set_error("too many query points"); return -1; }
        const int blk = h->ny >= 256 ? 256 : 64;
        CU_TRY((h->fp_mode == FP_STRICT ? cta_dense_eval_strict : cta_dense_eval_fast)(h->method, q, (int)pairs, blk, h->stream));
        if (!h->device_io) {
            CU_TRY(cudaMemcpyAsync(out, d_out, B * n_t * NDY * 8, cudaMemcpyDeviceToHost, h->stream));
            CU_TRY(cudaMemcpyAsync(codes, q.code, B * 4, cudaMemcpyDeviceToHost, h->stream));
        }
        CU_TRY(cudaStreamSynchronize(h->stream));
        return 0;
    }
    Dispatch disp;
    if (!make_dispatch(h->rhs_id, h->fp_mode, h->device, &disp)) { set_error("dense evaluation is not available for right-hand side %d", h->rhs_id); return -1; }
    DenseEvalParams q;
    memset(&q, 0, sizeof(q));
    q.n_traj = B;
    q.n_query = (unsigned)n_t;
    if (h->d_q_t.ensure(n_t * 8) || h->d_q_code.ensure(B * 4)) return -1;
    CU_TRY(cudaMemcpyAsync(h->d_q_t.p, t, n_t * 8, h->device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    q.t_query = h->d_q_t.as<double>();
    q.t_span = h->in_tspan;
    q.t_span_stride = (unsigned)h->tspan_stride;
    q.status = h->d_status.as<int>();
    q.dense_t = h->d_dense_t.as<double>();
    q.dense_state = h->d_dense_state.as<double>();
    q.dense_count = h->d_dense_count.as<unsigned>();
    q.dense_cap = (unsigned)h->dense_cap;
    if (h->dense_from_steps) { q.recs = h->d_st.as<double>(); q.steps_taken = h->d_steps.as<unsigned>(); }
    q.args = h->in_args;
    q.args_stride = (unsigned)h->args_stride;
    double* d_out;
    if (h->device_io) d_out = out;
    else {
        if (h->d_q_out.ensure(B * n_t * NDY * 8)) return -1;
        d_out = h->d_q_out.as<double>();
    }
    q.out = d_out;
    q.code = h->device_io ? codes : h->d_q_code.as<int>();
    CU_TRY(cudaMemsetAsync(q.code, 0, B * 4, h->stream));
    const unsigned long long total = (unsigned long long)B * n_t;
    const int block = 256;
    const unsigned long long grid = (total + block - 1) / block;
    if (grid > 0x7fffffffull) { set_error("too many query points"); return -1; }
    if (disp.dense(h->method, q, (int)grid, block, h->stream) != cudaSuccess) { cudaGetLastError(); if (g_last_error.empty()) set_error("dense evaluation launch failed"); return -1; }
    if (!h->device_io) {
        CU_TRY(cudaMemcpyAsync(out, d_out, B * n_t * NDY * 8, cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaMemcpyAsync(codes, q.code, B * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

int cyrk_b200_ensemble_sums(void* hv, double* out)
{
    Handle* h = H(hv);
    if (!h || !h->solved) { set_error("handle holds no solution"); return -1; }
    if (!out) { set_error("out is NULL"); return -1; }
    const int ndy = (int)h->ndy, nv = 2 * ndy + 1;
    if (nv > 256) { set_error("ensemble sums support num_dy <= 127"); return -1; }
    DeviceGuard guard;
    CU_TRY(cudaSetDevice(h->device));
    const int block = 256, grid = std::max(1, std::min(device_sm_count(h->device) * 4, (int)((h->n_traj + block - 1) / block)));
    if (h->d_scratch.ensure((size_t)(grid + 1) * nv * 8)) return -1;
    double* partial = h->d_scratch.as<double>();
    double* dout = h->device_io ? out : partial + (size_t)grid * nv;
    ensemble_partial_kernel<<<grid, block, (block / 32) * nv * sizeof(double), h->stream>>>(
        h->d_status.as<int>(), h->d_ylast.as<double>(), h->n_traj, ndy, partial);
    ensemble_final_kernel<<<1, 256, 0, h->stream>>>(partial, grid, nv, dout);
    CU_TRY(cudaGetLastError());
    if (!h->device_io) CU_TRY(cudaMemcpyAsync(out, dout, nv * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

double cyrk_b200_get_kernel_ms(void* h) { return H(h)->kernel_ms; }
double cyrk_b200_get_total_ms(void* h) { return H(h)->total_ms; }
unsigned cyrk_b200_get_launch_count(void* h) { return H(h)->launches; }
size_t cyrk_b200_get_h2d_bytes(void* h) { return H(h)->h2d_bytes; }
size_t cyrk_b200_get_d2h_bytes(void* h) { return H(h)->d2h_bytes; }
size_t cyrk_b200_get_lazy_d2h_bytes(void* h) { return H(h)->lazy_d2h_bytes; }

int cyrk_b200_set_option(const char* name, int value)
{
    if (!name) return -1;
    if (!strcmp(name, "refill_min")) { g_opt_refill_min = value; return 0; }
    if (!strcmp(name, "block")) { if (value == 0) { g_opt_block = 0; return 0; } if (value < 32 || value > 1024 || value % 32) { set_error("block must be a multiple of 32 in [32,1024]"); return -1; } g_opt_block = value; return 0; }
    if (!strcmp(name, "ctas_per_sm")) { g_opt_ctas_per_sm = value; return 0; }
    if (!strcmp(name, "prep")) { g_opt_prep = value; return 0; }
    if (!strcmp(name, "root_batch")) { g_opt_root_batch = value; return 0; }
    if (!strcmp(name, "two_pass")) { g_opt_two_pass = value; return 0; }
    if (!strcmp(name, "dense_store")) { g_opt_dense_store = value; return 0; }
    if (!strcmp(name, "stream")) { g_opt_stream = value; return 0; }
    if (!strcmp(name, "stream_chunk")) { g_opt_stream_chunk = value; return 0; }
    if (!strcmp(name, "stream_taper")) { g_opt_stream_taper = value; return 0; }
    if (!strcmp(name, "shared_args")) { g_opt_shared_args = value; return 0; }
    if (!strcmp(name, "events_variant")) { g_opt_events_variant = value; return 0; }
    if (!strcmp(name, "events_two_pass")) { g_opt_events_two_pass = value; return 0; }
    if (!strcmp(name, "events_static")) { g_opt_events_static = value; return 0; }
    if (!strcmp(name, "cta_smem_vecs")) { g_opt_cta_smem_vecs = value; return 0; }
    if (!strcmp(name, "cta2_large")) { g_opt_cta2_large = value; return 0; }
    if (!strcmp(name, "small_cta")) { g_opt_small_cta = value; return 0; }
    if (!strcmp(name, "cta2_per_sm")) { g_opt_cta2_per_sm = value; return 0; }
    if (!strcmp(name, "cta_block")) { g_opt_cta_block = value; return 0; }
    if (!strcmp(name, "cta_kernel")) { g_opt_cta_kernel = value; return 0; }
    if (!strcmp(name, "pipe_taper")) { g_opt_pipe_taper = value; return 0; }
    if (!strcmp(name, "pipe_sched")) { g_opt_pipe_sched = value; return 0; }
    if (!strcmp(name, "pipe_min")) { g_opt_pipe_min = value; return 0; }
    if (!strcmp(name, "pipe_chunk")) { if (value < 1024) { set_error("pipe_chunk must be >= 1024"); return -1; } g_opt_pipe_chunk = value; return 0; }
    set_error("unknown option '%s'", name);
    return -1;
}

int cyrk_b200_debug_sincos(const double* x, double* sin_out, double* cos_out, size_t n, int device)
{
    DeviceGuard guard;
    const int mode = device >> 8;            // debug entry point: bits 8.. of `device` select the function pair (see the kernel)
    device &= 0xff;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); set_error("no CUDA device %d", device); return -1; }
    double *dx = nullptr, *ds = nullptr, *dc = nullptr;
    int rc = -1;
    if (cudaMalloc(&dx, n * 8) == cudaSuccess && cudaMalloc(&ds, n * 8) == cudaSuccess && cudaMalloc(&dc, n * 8) == cudaSuccess &&
        cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice) == cudaSuccess) {
        sincos_probe_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dx, ds, dc, n, mode);
        if (cudaMemcpy(sin_out, ds, n * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
            cudaMemcpy(cos_out, dc, n * 8, cudaMemcpyDeviceToHost) == cudaSuccess)
            rc = 0;
    }
    if (rc != 0) { set_error("debug_sincos: %s", cudaGetErrorString(cudaGetLastError())); }
    cudaFree(dx); cudaFree(ds); cudaFree(dc);
    return rc;
}

double cyrk_b200_measure_fp64_tflops(int device)
{
    DeviceGuard guard;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); set_error("no CUDA device %d", device); return -1.0; }
    const int sms = device_sm_count(device);
    const int block = 256, grid = sms * 8, iters = 1 << 15;
    double* d = nullptr;
    if (cudaMalloc(&d, (size_t)grid * block * 8) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        fp64_probe_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double)iters * (double)grid * block;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

}  // extern "C"
