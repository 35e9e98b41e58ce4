// dmma_pipe.cu -- does the FP64 tensor-core path (mma.sync m8n8k4 f64, SASS DMMA) run BESIDE the DFMA pipe on sm_100a?
//
// Three kernels with the same grid (all SMs, 8 warps each): (1) DFMA only, 8 independent chains per thread; (2) DMMA only,
// 4 independent accumulator chains per warp; (3) both streams interleaved in one loop.  If the two share one pipe the mixed
// kernel takes the SUM of the two times; if they are separate units it takes about the MAX.  Rates are printed as FP64
// FMAs per clock per SM (DFMA: 32 per warp instruction, DMMA m8n8k4: 256 per warp instruction).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_pipe dmma_pipe.cu && ./dmma_pipe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 131072, NF = 8, NM = 4;

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <bool FMA, bool MMA>
__global__ void __launch_bounds__(256) pipe_kernel(double* out, double a, double b) {
    double f[NF];
    double m[NM][2];
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = a + i + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NM; ++i) { m[i][0] = a + i; m[i][1] = b + i; }
    const double am = a * 1e-3, bm = b * 1e-3;
#pragma unroll 1
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {           // 8 rounds per trip: the loop-carried counter chain must not be the limit
            if (FMA) {
#pragma unroll
                for (int i = 0; i < NF; ++i) f[i] = fma(f[i], b, a);
            }
            if (MMA) {
#pragma unroll
                for (int i = 0; i < NM; ++i) dmma(m[i], am, bm);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) s += m[i][0] + m[i][1];
    if (s == 123.456) out[0] = s;
}

template <bool FMA, bool MMA>
float run(int blocks, double* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    pipe_kernel<FMA, MMA><<<blocks, 256>>>(d_out, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        pipe_kernel<FMA, MMA><<<blocks, 256>>>(d_out, 1.0000001, 0.9999999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double* d_out; cudaMalloc(&d_out, 8);
    const int blocks = p.multiProcessorCount;          // one 8-warp block per SM: 2 warps per scheduler
    const double warps = (double)blocks * 8;
    const float t_f = run<true, false>(blocks, d_out), t_m = run<false, true>(blocks, d_out), t_b = run<true, true>(blocks, d_out);
    const double fma_f = warps * ITERS * NF * 32.0, fma_m = warps * ITERS * NM * 256.0;
    auto per_clk_sm = [&](double fmas, float ms) { return fmas / (ms * 1e-3) / (clk_khz * 1e3) / blocks; };
    printf("%s, %d SMs, %d MHz (nominal)\n", p.name, blocks, clk_khz / 1000);
    printf("DFMA only : %.3f ms  %.1f FMA/clk/SM (nominal clock)\n", t_f, per_clk_sm(fma_f, t_f));
    printf("DMMA only : %.3f ms  %.1f FMA/clk/SM\n", t_m, per_clk_sm(fma_m, t_m));
    printf("both      : %.3f ms  (sum of the two alone %.3f ms, max %.3f ms)  %.1f FMA/clk/SM combined\n", t_b, t_f + t_m, t_f > t_m ? t_f : t_m,
           per_clk_sm(fma_f + fma_m, t_b));
    printf("verdict   : %s\n", t_b < 0.75f * (t_f + t_m) ? "the two streams overlap: DMMA runs beside the DFMA pipe" : "no overlap: one FP64 pipe serves both");
    return 0;
}
