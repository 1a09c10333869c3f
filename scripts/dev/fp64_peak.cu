// FP64 issue rate of one B200 SM, measured: independent DFMA / DADD chains per thread, W warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak scripts/dev/fp64_peak.cu && /tmp/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool FMA>
__global__ void k(double* out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = FMA ? fma(v[i], a, b) : v[i] + a;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP, bool FMA>
void run(int warps, const char* name)
{
    double* d;
    cudaMalloc(&d, 8);
    const int iters = 20000, sms = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, FMA><<<sms, warps * 32>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP, FMA><<<sms, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)iters * ILP * warps;                // warp instructions per SM
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%s ILP %d warps/SM %2d: %.3f warp-inst/cycle/SM  (%.1f TFLOP/s if all 148 SMs, %s)\n", name, ILP, warps, inst / cyc,
           inst / cyc * 32 * (FMA ? 2 : 1) * 148 * 1.965e9 / 1e12, FMA ? "FMA = 2 flop" : "ADD = 1 flop");
    cudaFree(d);
}

int main()
{
    for (int w : {1, 4, 8, 16, 32}) run<8, true>(w, "DFMA");
    for (int w : {4, 16}) run<8, false>(w, "DADD");
    for (int w : {4, 16}) run<2, true>(w, "DFMA");
    return 0;
}
