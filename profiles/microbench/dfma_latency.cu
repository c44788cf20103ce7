// dfma_latency.cu -- how many independent FP64 chains per SM sub-partition saturate the B200 FP64 pipe.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_latency dfma_latency.cu && ./dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void k(double* out, int iters, double m, long long* cyc) {
    double a[C];
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) a[i] = fma(a[i], m, 1e-9);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += a[i];
    if (s == 12345.678) out[threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int C>
void run(int warps, double* d, long long* dc) {
    const int iters = 4096;
    k<C><<<148, warps * 32>>>(d, iters, 1.0000001, dc);
    k<C><<<148, warps * 32>>>(d, iters, 1.0000001, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double per_warp_instr = (double)c / (iters * C);                 // cycles per DFMA as seen by one warp
    const double smsp_rate = (double)iters * C * (warps / 4.0) / c;         // warp-DFMAs per cycle per SMSP
    printf("chains/thread %2d  warps/SM %2d (%.0f per SMSP): %6.2f cycles per DFMA per warp, %5.3f warp-DFMA/cycle/SMSP (peak 0.5)\n",
           C, warps, warps / 4.0, per_warp_instr, smsp_rate);
}

int main() {
    double* d; long long* dc;
    cudaMalloc(&d, 1 << 16); cudaMalloc(&dc, 8);
    for (int w : {4, 8, 16, 24, 32}) { run<1>(w, d, dc); run<2>(w, d, dc); run<4>(w, d, dc); }
    return 0;
}
