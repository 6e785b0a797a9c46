// FP64 pipe microbenchmark: warp-instructions per cycle per SM for DFMA / DMUL / DADD / DSETP+FSEL mixes, with register,
// constant-bank and mixed operand forms, 8 independent chains per thread.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_k[4] = {1.0000001, 0.9999999, 1e-9, -1e-9};
template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double a0, double b0)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a0 + 1e-3 * (threadIdx.x + i);
    double b = b0, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) x[i] = fma(x[i], b, c);                                   // DFMA reg,reg,reg
                if (MODE == 1) x[i] = x[i] * b;                                          // DMUL
                if (MODE == 2) x[i] = x[i] + c;                                          // DADD
                if (MODE == 3) x[i] = (u & 1) ? x[i] * b : fma(x[i], b, c);              // DFMA / DMUL alternating
                if (MODE == 4) x[i] = fma(x[i], c_k[0], c_k[2]);                         // DFMA with constant operands
                if (MODE == 5) x[i] = (x[i] > b) ? x[i] * b : x[i] + c;                  // DSETP + select + both
                if (MODE == 6) x[i] = fma(x[i], b, x[(i + 1) & 7]);                      // DFMA, 3 distinct registers
                if (MODE == 7) x[i] = fabs(x[i]) * c_k[1];                               // DMUL |x|, const
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}
template <int MODE> void run(const char* name, int fp_per_elem, double* d, int sms)
{
    const int iters = 4096, blocks = sms * 4, threads = 256;
    probe<MODE><<<blocks, threads>>>(d, 16, 1.0, 1.0000001);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(d, iters, 1.0, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d + blocks * threads, 8, cudaMemcpyDeviceToHost);
    // per SM: 4 CTAs x 8 warps = 32 warps; per warp: iters*4*8 elems
    const double warp_elems = (double)iters * 32.0 * 32.0;   // warp-level "element updates" per SM
    // (clock64 of one CTA is printed for reference only: the compiler is free to move the reads; the rate uses the event time
    //  and the SM clock under load, 1.965 GHz on this pool's B200s)
    const double clk = ms * 1.965e6;
    printf("%-34s %8.3f ms  (cta0 clock64 %9.0f)  %.3f FP64 warp-instr/clk/SM of 2.0  (%d FP64 instr per update)\n",
           name, ms, cyc, fp_per_elem * warp_elems / clk, fp_per_elem);
}
int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, (size_t)(sms * 4 * 256 + 1) * 8);
    run<0>("DFMA r,r,r", 1, d, sms);
    run<1>("DMUL", 1, d, sms);
    run<2>("DADD", 1, d, sms);
    run<3>("DFMA/DMUL alternating", 1, d, sms);
    run<4>("DFMA r,c,c (constant bank)", 1, d, sms);
    run<5>("DSETP + select(DMUL, DADD)", 3, d, sms);
    run<6>("DFMA r,r,r2 (3 distinct)", 1, d, sms);
    run<7>("DMUL |r|,c", 1, d, sms);
    return 0;
}
