// fp64 pipe microbenchmark (development tool): DFMA throughput per SM vs resident warps and operand pattern.
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS, int MODE>
__global__ void k(double *sink, int iters, double s0, double s1)
{
    double a[CHAINS], b[CHAINS], c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = 1.0 + 1e-9 * (threadIdx.x + i); b[i] = s0 + 1e-12 * i; c[i] = s1 + 1e-13 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) a[i] = fma(a[i], b[0], c[0]);            // shared multiplier/addend (probe pattern)
            else if (MODE == 1) a[i] = fma(a[i], b[i], c[i]);       // three distinct register operands
            else { a[i] = fma(a[i], b[i], c[i]); b[i] = b[i] * c[i]; }   // DFMA + DMUL mix
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] + b[i];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS, int MODE>
void run(double *sink, int warps_per_sm)
{
    int iters = 4096;
    dim3 grid(148), block(32 * warps_per_sm);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CHAINS, MODE><<<grid, block>>>(sink, 16, 1.0000001, 1e-15);
    cudaEventRecord(e0);
    k<CHAINS, MODE><<<grid, block>>>(sink, iters, 1.0000001, 1e-15);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)iters * CHAINS * (MODE == 2 ? 2 : 1) * 148 * warps_per_sm;   // warp instructions
    double cyc = ms * 1e-3 * 1.92e9;
    printf("chains %2d mode %d warps/SM %2d : %.3f ms  %.2f fp64 warp-inst/clk/SM  (%.1f TFLOP/s DFMA-equiv)\n", CHAINS, MODE,
           warps_per_sm, ms, inst / cyc / 148, inst * 32 * 2 / (ms * 1e-3) / 1e12);
}
int main()
{
    double *sink; cudaMalloc(&sink, 148 * 1024 * sizeof(double));
    for (int w : {4, 8, 16, 32}) { run<1, 1>(sink, w); run<2, 1>(sink, w); run<4, 1>(sink, w); run<8, 1>(sink, w); }
    for (int w : {4, 8, 16}) { run<8, 0>(sink, w); run<8, 2>(sink, w); run<16, 0>(sink, w);}
    return 0;
}
