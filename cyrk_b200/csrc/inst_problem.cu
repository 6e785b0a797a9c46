// inst_problem.cu -- instantiates the thread-per-trajectory kernels for ONE built-in problem and
// ONE floating-point mode.  Compiled once per (problem, mode) by cyrk_b200/csrc/Makefile with
//   -DCYRK_INST_PROB=<Prob struct> -DCYRK_INST_ID=<RHS_* id> -DCYRK_FP_NS=fast|strict
//   -DCYRK_FP_MODE=FP_FAST|FP_STRICT   and   -fmad=true|false
#include "rk_core.cuh"
#include "problems_builtin.cuh"
#include "registry.h"

#define CYRK_CAT_(a, b, c) a##_##b##_##c
#define CYRK_CAT(a, b, c) CYRK_CAT_(a, b, c)
#define CYRK_STR_(x) #x
#define CYRK_STR(x) CYRK_STR_(x)

namespace cyrk {
namespace CYRK_FP_NS {
namespace {

// events + storage, RK23 / RK45: the register-capped definition (rk_core.cuh); DOP853 would spill 500 bytes under the cap
#ifndef CYRK_FULL_CAP
#define CYRK_FULL_CAP 1
#endif
template <int METHOD, int VAR>
constexpr bool capped() { return CYRK_FULL_CAP && VAR == VAR_FULL && METHOD != METHOD_DOP853; }

template <int METHOD, int VAR>
cudaError_t launch(const BatchParams& p, int grid, int block, cudaStream_t s)
{
    if constexpr (capped<METHOD, VAR>()) rk_thread_kernel_capped<CYRK_INST_PROB, METHOD, VAR><<<grid, block, 0, s>>>(p);
    else rk_thread_kernel<CYRK_INST_PROB, METHOD, VAR><<<grid, block, 0, s>>>(p);
    return cudaGetLastError();
}
template <int METHOD, int VAR>
cudaError_t occupancy(int* ctas, int block)
{
    if constexpr (capped<METHOD, VAR>())
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, rk_thread_kernel_capped<CYRK_INST_PROB, METHOD, VAR>, block, 0);
    else
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, rk_thread_kernel<CYRK_INST_PROB, METHOD, VAR>, block, 0);
}
template <int METHOD>
cudaError_t dense_eval(const DenseEvalParams& p, int grid, int block, cudaStream_t s)
{
    dense_eval_kernel<CYRK_INST_PROB, METHOD><<<grid, block, 0, s>>>(p);
    return cudaGetLastError();
}
template <int METHOD>
cudaError_t prep(const BatchParams& p, double* out, int grid, int block, cudaStream_t s)
{
    rk_prep_kernel<CYRK_INST_PROB, METHOD><<<grid, block, 0, s>>>(p, out);
    return cudaGetLastError();
}
template <int METHOD>
cudaError_t step_dense(const StepDenseParams& p, int grid, int block, cudaStream_t s)
{
    step_dense_kernel<CYRK_INST_PROB, METHOD><<<grid, block, 0, s>>>(p);
    return cudaGetLastError();
}
template <int METHOD>
cudaError_t event_roots(const BatchParams& p, int grid, int block, cudaStream_t s)
{
    event_roots_kernel<CYRK_INST_PROB, METHOD><<<grid, block, 0, s>>>(p);
    return cudaGetLastError();
}
template <int METHOD>
void fill(ProblemEntry* e)
{
    e->solve[METHOD - 3][VAR_LEAN] = {launch<METHOD, VAR_LEAN>, occupancy<METHOD, VAR_LEAN>, CYRK_STR(CYRK_INST_PROB) "/lean"};
    e->solve[METHOD - 3][VAR_LEAN_SHARED] = {launch<METHOD, VAR_LEAN_SHARED>, occupancy<METHOD, VAR_LEAN_SHARED>, CYRK_STR(CYRK_INST_PROB) "/lean-shared-args"};
    e->solve[METHOD - 3][VAR_ROWS] = {launch<METHOD, VAR_ROWS>, occupancy<METHOD, VAR_ROWS>, CYRK_STR(CYRK_INST_PROB) "/rows"};
    e->solve[METHOD - 3][VAR_FULL] = {launch<METHOD, VAR_FULL>, occupancy<METHOD, VAR_FULL>, CYRK_STR(CYRK_INST_PROB) "/full"};
    e->solve[METHOD - 3][VAR_STEPS] = {launch<METHOD, VAR_STEPS>, occupancy<METHOD, VAR_STEPS>, CYRK_STR(CYRK_INST_PROB) "/steps"};
    e->solve[METHOD - 3][VAR_EVENTS] = {launch<METHOD, VAR_EVENTS>, occupancy<METHOD, VAR_EVENTS>, CYRK_STR(CYRK_INST_PROB) "/events"};
    e->solve[METHOD - 3][VAR_EVDETECT] = {launch<METHOD, VAR_EVDETECT>, occupancy<METHOD, VAR_EVDETECT>, CYRK_STR(CYRK_INST_PROB) "/event-detect"};
    e->event_roots[METHOD - 3] = event_roots<METHOD>;
    e->dense_eval[METHOD - 3] = dense_eval<METHOD>;
    e->step_dense[METHOD - 3] = step_dense<METHOD>;
    e->prep[METHOD - 3] = prep<METHOD>;
}

}  // namespace
}  // namespace CYRK_FP_NS

void CYRK_CAT(cyrk_register, CYRK_INST_PROB, CYRK_FP_NS)()
{
    ProblemEntry* e = registry_slot(CYRK_FP_MODE, CYRK_INST_ID);
    e->present = true;
    e->struct_name = CYRK_STR(CYRK_INST_PROB);
    e->ny = CYRK_INST_PROB::NY;
    e->nextra = CYRK_INST_PROB::NEXTRA;
    e->nargs_dbl = CYRK_INST_PROB::NARGS_DBL;
    CYRK_FP_NS::fill<METHOD_RK23>(e);
    CYRK_FP_NS::fill<METHOD_RK45>(e);
    CYRK_FP_NS::fill<METHOD_DOP853>(e);
}

}  // namespace cyrk
