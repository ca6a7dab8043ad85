// Micro-benchmark: what the FP64 pipe of one B200 SM sustains (DFMA per clock per SM), alone and with integer work interleaved --
// the roof the fp64 check rule is measured against. nvcc -arch=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS, int INT_PER_FMA>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b, int s)
{
    double x[CHAINS];
    int y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
    {
        x[c] = threadIdx.x * 1e-3 + c;
        y[c] = threadIdx.x + c;
    }
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
        {
            x[c] = fma(x[c], a, b);
#pragma unroll
            for (int q = 0; q < INT_PER_FMA; ++q)
                y[c] = y[c] * s + q; // IMAD
        }
    double acc = 0;
    int iacc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
    {
        acc += x[c];
        iacc += y[c];
    }
    if (acc == 123.456 || iacc == 77)
        out[threadIdx.x] = acc + iacc;
}
template <int CHAINS, int INT_PER_FMA>
void run(int sms, double mhz, int blocks_per_sm)
{
    double *out;
    cudaMalloc(&out, 4096);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<CHAINS, INT_PER_FMA><<<sms * blocks_per_sm, 256>>>(out, 100, 1.0000001, 1e-9, 3);
    cudaEventRecord(e0);
    k<CHAINS, INT_PER_FMA><<<sms * blocks_per_sm, 256>>>(out, iters, 1.0000001, 1e-9, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)sms * blocks_per_sm * 256 * iters * CHAINS;
    printf("chains %d, %d IMAD per DFMA, %d warps/SM: %.3f ms, %.2f T DFMA/s = %.2f TFLOP/s, %.1f DFMA lanes/clk/SM at %.0f MHz; total warp-instr/clk/SM %.2f\n", CHAINS, INT_PER_FMA,
           blocks_per_sm * 8, ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / (mhz * 1e6), mhz,
           fmas * (1 + INT_PER_FMA) / 32 / (ms * 1e-3) / sms / (mhz * 1e6));
    cudaFree(out);
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    printf("%s, %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    run<8, 0>(p.multiProcessorCount, mhz, 1);
    run<8, 0>(p.multiProcessorCount, mhz, 2);
    run<8, 0>(p.multiProcessorCount, mhz, 4);
    run<4, 0>(p.multiProcessorCount, mhz, 8);
    run<8, 1>(p.multiProcessorCount, mhz, 4);
    run<8, 2>(p.multiProcessorCount, mhz, 4);
    run<8, 3>(p.multiProcessorCount, mhz, 4);
    run<2, 0>(p.multiProcessorCount, mhz, 2);
    run<1, 0>(p.multiProcessorCount, mhz, 1);
    return 0;
}
