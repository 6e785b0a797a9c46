// inst_cta.cu -- instantiates the CTA-per-trajectory kernels (one large built-in problem per translation unit, three
// methods) for ONE floating-point mode: compiled four times by the Makefile (-DCYRK_CTA_PART=0|1 x
// -DCYRK_FP_NS=fast -fmad=true / strict -fmad=false).
#include "cta_core.cuh"
#include "cta2_core.cuh"
#include "problems_builtin.cuh"
#include "problems_large.cuh"
#include "cta_solver.h"

#define CYRK_CAT2_(a, b) a##b
#define CYRK_CAT2(a, b) CYRK_CAT2_(a, b)

namespace cyrk {
namespace CYRK_FP_NS {
namespace {
template <class Prob, int METHOD>
cudaError_t go(const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(rk_cta_kernel<Prob, METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rk_cta_kernel<Prob, METHOD><<<grid, block, smem, s>>>(p);
    return cudaGetLastError();
}
template <class Prob>
cudaError_t by_method(int method, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    switch (method) {
    case METHOD_RK23: return go<Prob, METHOD_RK23>(p, grid, block, smem, s);
    case METHOD_RK45: return go<Prob, METHOD_RK45>(p, grid, block, smem, s);
    case METHOD_DOP853: return go<Prob, METHOD_DOP853>(p, grid, block, smem, s);
    default: return cudaErrorInvalidValue;
    }
}
// register-resident kernels (cta2_core.cuh): C components per thread
template <class Prob, int METHOD, int C, bool EVENTS>
cudaError_t go2e(const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(rk_cta2_kernel<Prob, METHOD, C, EVENTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rk_cta2_kernel<Prob, METHOD, C, EVENTS><<<grid, block, smem, s>>>(p);
    return cudaGetLastError();
}
template <class Prob, int METHOD, int C>
cudaError_t go2(const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    return p.n_events > 0 ? go2e<Prob, METHOD, C, true>(p, grid, block, smem, s) : go2e<Prob, METHOD, C, false>(p, grid, block, smem, s);
}
template <class Prob, int METHOD>
cudaError_t by_c(int C, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    switch (C) {
    case 1: return go2<Prob, METHOD, 1>(p, grid, block, smem, s);
    case 2: return go2<Prob, METHOD, 2>(p, grid, block, smem, s);
    case 4: return go2<Prob, METHOD, 4>(p, grid, block, smem, s);
    case 8: return go2<Prob, METHOD, 8>(p, grid, block, smem, s);
    // n_y in (4096, 12 288]: derivatives in the L2-resident slab; RK23 / RK45 (DOP853's 16 vectors of 96 KB x 148 CTAs
    // would not stay in L2; cta_core.cuh takes it)
    case 12: if constexpr (METHOD != METHOD_DOP853) return go2<Prob, METHOD, 12>(p, grid, block, smem, s); else return cudaErrorInvalidValue;
    case 16: if constexpr (METHOD != METHOD_DOP853) return go2<Prob, METHOD, 16>(p, grid, block, smem, s); else return cudaErrorInvalidValue;
    case 20: if constexpr (METHOD != METHOD_DOP853) return go2<Prob, METHOD, 20>(p, grid, block, smem, s); else return cudaErrorInvalidValue;
    case 24: if constexpr (METHOD != METHOD_DOP853) return go2<Prob, METHOD, 24>(p, grid, block, smem, s); else return cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
    }
}
template <class Prob>
cudaError_t by_method2(int method, int C, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    switch (method) {
    case METHOD_RK23: return by_c<Prob, METHOD_RK23>(C, p, grid, block, smem, s);
    case METHOD_RK45: return by_c<Prob, METHOD_RK45>(C, p, grid, block, smem, s);
    case METHOD_DOP853: return by_c<Prob, METHOD_DOP853>(C, p, grid, block, smem, s);
    default: return cudaErrorInvalidValue;
    }
}
}  // namespace
}  // namespace CYRK_FP_NS

// One translation unit per built-in large problem (CYRK_CTA_PART: 0 = largey, 1 = largey_simple; they compile in
// parallel); part 0 also holds the dispatch by rhs id and the problem-independent kernels.
#ifndef CYRK_CTA_PART
#error "compile with -DCYRK_CTA_PART=0|1"
#endif
#define CYRK_CAT3_(a, b, c) a##b##c
#define CYRK_CAT3(a, b, c) CYRK_CAT3_(a, b, c)
#if CYRK_CTA_PART == 0
using CtaPartProb = ProbLargeY;
#else
using CtaPartProb = ProbLargeYSimple;
#endif
cudaError_t CYRK_CAT3(cta_launch_part_, CYRK_FP_NS, CYRK_CTA_PART)(int method, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    return CYRK_FP_NS::by_method<CtaPartProb>(method, p, grid, block, smem, s);
}
cudaError_t CYRK_CAT3(cta2_launch_part_, CYRK_FP_NS, CYRK_CTA_PART)(int method, int C, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    return CYRK_FP_NS::by_method2<CtaPartProb>(method, C, p, grid, block, smem, s);
}

#if CYRK_CTA_PART == 0
cudaError_t CYRK_CAT3(cta_launch_part_, CYRK_FP_NS, 1)(int method, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s);
cudaError_t CYRK_CAT3(cta2_launch_part_, CYRK_FP_NS, 1)(int method, int C, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s);

cudaError_t CYRK_CAT2(cta_launch_, CYRK_FP_NS)(int rhs_id, int method, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    if (rhs_id == RHS_LARGEY) return CYRK_CAT3(cta_launch_part_, CYRK_FP_NS, 0)(method, p, grid, block, smem, s);
    if (rhs_id == RHS_LARGEY_SIMPLE) return CYRK_CAT3(cta_launch_part_, CYRK_FP_NS, 1)(method, p, grid, block, smem, s);
    return cudaErrorInvalidValue;
}

cudaError_t CYRK_CAT2(cta2_launch_, CYRK_FP_NS)(int rhs_id, int method, int C, const CtaParams& p, int grid, int block, size_t smem, cudaStream_t s)
{
    if (rhs_id == RHS_LARGEY) return CYRK_CAT3(cta2_launch_part_, CYRK_FP_NS, 0)(method, C, p, grid, block, smem, s);
    if (rhs_id == RHS_LARGEY_SIMPLE) return CYRK_CAT3(cta2_launch_part_, CYRK_FP_NS, 1)(method, C, p, grid, block, smem, s);
    return cudaErrorInvalidValue;
}

cudaError_t CYRK_CAT2(cta_dense_eval_, CYRK_FP_NS)(int method, const CtaDenseEvalParams& p, int grid, int block, cudaStream_t s)
{
    switch (method) {
    case METHOD_RK23: CYRK_FP_NS::cta_dense_eval_kernel<METHOD_RK23><<<grid, block, 0, s>>>(p); break;
    case METHOD_RK45: CYRK_FP_NS::cta_dense_eval_kernel<METHOD_RK45><<<grid, block, 0, s>>>(p); break;
    case METHOD_DOP853: CYRK_FP_NS::cta_dense_eval_kernel<METHOD_DOP853><<<grid, block, 0, s>>>(p); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

#if !CYRK_FP_STRICT
int cta_num_vectors(int method)
{
    return method == METHOD_RK23 ? 2 + TabRK23::NK_DENSE : method == METHOD_RK45 ? 2 + TabRK45::NK_DENSE : 2 + TabDOP853::NK_DENSE;
}
#endif
#endif  // CYRK_CTA_PART == 0

}  // namespace cyrk
