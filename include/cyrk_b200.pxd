# cyrk_b200.pxd -- Cython declarations of include/cyrk_b200.h (generated from the header; see INTEGRATION.md).
# `cimport` users call these `nogil`, like cysolve_ivp (CyRK/cy/cysolver_api.pxd:385-407).
from libc.stdint cimport int32_t, uint32_t, uint64_t

cdef extern from "cyrk_b200.h" nogil:
    enum: CYRK_METHOD_RK23
    enum: CYRK_METHOD_RK45
    enum: CYRK_METHOD_DOP853
    enum: CYRK_RHS_BASELINE
    enum: CYRK_RHS_ACCURACY
    enum: CYRK_RHS_EXTRAOUTPUT
    enum: CYRK_RHS_LORENZ
    enum: CYRK_RHS_LORENZ_EXTRA
    enum: CYRK_RHS_LOTKAVOLTERRA
    enum: CYRK_RHS_PENDULUM
    enum: CYRK_RHS_ARBITRARY_ARGS
    enum: CYRK_RHS_PENDULUM_PREEVAL
    enum: CYRK_RHS_LARGEY
    enum: CYRK_RHS_LARGEY_SIMPLE
    enum: CYRK_RHS_PENDULUM_DRAG
    enum: CYRK_RHS_DENSE_KAT
    enum: CYRK_RHS_USER_BASE
    enum: CYRK_EV_Y0_GT_50
    enum: CYRK_EV_Y1_LT_50
    enum: CYRK_EV_T_GT_5
    enum: CYRK_EV_LORENZ_T_GE_5
    enum: CYRK_EV_LORENZ_Y_RANGE
    enum: CYRK_EV_LORENZ_ARGS
    enum: CYRK_EV_Y0
    enum: CYRK_EV_Y1
    enum: CYRK_EV_T_MINUS_ARG4
    enum: CYRK_EV_ABS_Y0_MINUS_ARG4
    enum: CYRK_STORE_FINAL
    enum: CYRK_STORE_ROWS
    enum: CYRK_FLAG_STRICT_FP
    enum: CYRK_FLAG_DEVICE_IO
    int cyrk_b200_version()
    int cyrk_b200_device_count()
    const char* cyrk_b200_last_error()
    void cyrk_b200_status_message(int status_code, char* buffer, size_t max_len)
    int cyrk_b200_rhs_builtin(const char* name)
    int cyrk_b200_rhs_info(int rhs_id, size_t* n_y, size_t* n_extra, size_t* args_bytes)
    int cyrk_b200_rhs_compile(const char* source, const char* entry, size_t n_y, size_t n_extra, size_t args_bytes, size_t n_event_funcs, char* log, size_t log_len)
    void* cyrk_b200_solve_batch( int rhs_id, int method, size_t n_traj, const double* t_span, size_t n_t_span, const double* y0, size_t n_y, size_t num_extra, const char* args, size_t args_len, size_t n_args, size_t max_num_steps, size_t max_ram_MB, int dense_output, const double* t_eval, size_t n_t_eval, const int* event_ids, const int* event_dirs, const size_t* event_max, size_t n_events, const double* rtols, size_t n_rtols, const double* atols, size_t n_atols, double max_step, double first_step, size_t expected_size, int store_mode, unsigned int flags, int device)
    int cyrk_b200_solve_batch_sharded(const int* device_ids, size_t n_dev, void** handles_out, size_t* bounds_out,
                                int rhs_id, int method, size_t n_traj,
                                const double* t_span, size_t n_t_span, const double* y0, size_t n_y, size_t num_extra,
                                const char* args, size_t args_len, size_t n_args, size_t max_num_steps, size_t max_ram_MB,
                                int dense_output, const double* t_eval, size_t n_t_eval,
                                const int* event_ids, const int* event_dirs, const size_t* event_max, size_t n_events,
                                const double* rtols, size_t n_rtols, const double* atols, size_t n_atols,
                                double max_step, double first_step, size_t expected_size,
                                int store_mode, unsigned int flags)
    int cyrk_b200_solve_batch_noreturn( void* handle, int rhs_id, int method, size_t n_traj, const double* t_span, size_t n_t_span, const double* y0, size_t n_y, size_t num_extra, const char* args, size_t args_len, size_t n_args, size_t max_num_steps, size_t max_ram_MB, int dense_output, const double* t_eval, size_t n_t_eval, const int* event_ids, const int* event_dirs, const size_t* event_max, size_t n_events, const double* rtols, size_t n_rtols, const double* atols, size_t n_atols, double max_step, double first_step, size_t expected_size, int store_mode, unsigned int flags, int device)
    void* cyrk_b200_result_create()
    void cyrk_b200_free(void* handle)
    size_t cyrk_b200_get_n_traj(void* h)
    size_t cyrk_b200_get_num_y(void* h)
    size_t cyrk_b200_get_num_dy(void* h)
    const int32_t* cyrk_b200_get_status_ptr(void* h)
    const uint32_t* cyrk_b200_get_size_ptr(void* h)
    const uint32_t* cyrk_b200_get_steps_taken_ptr(void* h)
    const uint32_t* cyrk_b200_get_n_rejected_ptr(void* h)
    const uint32_t* cyrk_b200_get_nfev_ptr(void* h)
    const double* cyrk_b200_get_t_last_ptr(void* h)
    const double* cyrk_b200_get_y_last_ptr(void* h)
    size_t cyrk_b200_get_row_capacity(void* h)
    const double* cyrk_b200_get_t_ptr(void* h)
    const double* cyrk_b200_get_y_ptr(void* h)
    size_t cyrk_b200_get_num_events(void* h)
    size_t cyrk_b200_get_event_capacity(void* h)
    const uint64_t* cyrk_b200_get_event_offsets_ptr(void* h)
    size_t cyrk_b200_get_event_total(void* h)
    const uint32_t* cyrk_b200_get_event_count_ptr(void* h)
    const double* cyrk_b200_get_event_t_ptr(void* h)
    const double* cyrk_b200_get_event_y_ptr(void* h)
    const int32_t* cyrk_b200_get_event_terminate_index_ptr(void* h)
    size_t cyrk_b200_get_num_interpolates_capacity(void* h)
    const uint32_t* cyrk_b200_get_num_interpolates_ptr(void* h)
    int cyrk_b200_call_vectorize(void* h, const double* t, size_t n_t, double* out, int32_t* codes)
    int cyrk_b200_ensemble_sums(void* h, double* out)
    double cyrk_b200_get_kernel_ms(void* h)
    double cyrk_b200_get_total_ms(void* h)
    unsigned cyrk_b200_get_launch_count(void* h)
    size_t cyrk_b200_get_h2d_bytes(void* h)
    size_t cyrk_b200_get_d2h_bytes(void* h)
    size_t cyrk_b200_get_lazy_d2h_bytes(void* h)
    int cyrk_b200_jit_compile(int rhs_id, int method, int variant, unsigned flags)
    void cyrk_b200_jit_cache_stats(unsigned long long* compiled, unsigned long long* disk_hits)
    int cyrk_b200_set_option(const char* name, int value)
    double cyrk_b200_measure_fp64_tflops(int device)
    int cyrk_b200_debug_sincos(const double* x, double* sin_out, double* cos_out, size_t n, int device)
