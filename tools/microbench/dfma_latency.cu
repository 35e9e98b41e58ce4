// DFMA latency / throughput probe: C independent chains per thread, W warps per SM sub-partition.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, double a, double b, int iters) {
    double v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = a + c + threadIdx.x * 1e-9;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = fma(v[c], b, a);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += v[c];
    if (s == 123.456) out[0] = s;
}
template <int C>
void run(int warps_per_smsp, int sms) {
    double* d; cudaMalloc(&d, 8);
    int threads = warps_per_smsp * 4 * 32;   // one block per SM
    int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<C><<<sms, threads>>>(d, 1.0000001, 0.9999999, iters);
    cudaEventRecord(e0);
    k<C><<<sms, threads>>>(d, 1.0000001, 0.9999999, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double per_warp_instr = (double)iters * 8 * C;
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    double rate = per_warp_instr * warps_per_smsp / cycles;   // warp-instr per cycle per SMSP
    printf("chains %d warps/SMSP %d: %.2f cycles per dependent DFMA step; %.3f warp-DFMA/cycle/SMSP (peak 0.5)\n", C, warps_per_smsp,
           cycles / (iters * 8.0), rate);
    cudaFree(d);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int w : {1, 2, 3, 4}) { run<1>(w, sms); run<2>(w, sms); run<3>(w, sms); run<4>(w, sms); run<6>(w, sms); run<8>(w, sms); }
    return 0;
}
