// fp64_pipes.cu -- B200 FP64 microbenchmark: vector DFMA, tensor DMMA (mma.sync m8n8k4 f64), both mixed in
// one warp and both in different warps of one CTA.  Answers whether the response fold should use DMMA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: DFMA only, 1: DMMA only, 2: both interleaved in every warp, 3: even warps DFMA / odd warps DMMA
__global__ void __launch_bounds__(256) k(double* out, int iters, int mode, double m) {
    double a[8], c[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    const bool do_f = mode == 0 || mode == 2 || (mode == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_m = mode == 1 || mode == 2 || (mode == 3 && ((threadIdx.x >> 5) & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (do_f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, 1e-9);
        }
        if (do_m) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[2 * i], c[2 * i + 1], m, m);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[threadIdx.x] = s;
}

int main() {
    double* d;
    cudaMalloc(&d, 4096);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, iters = 1 << 13;
    const char* names[4] = {"DFMA only", "DMMA only", "DFMA+DMMA same warp", "DFMA | DMMA split warps"};
    for (int mode = 0; mode < 4; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            k<<<blocks, 256>>>(d, iters, mode, 1.0000001);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        const double thr = (double)blocks * 256 * iters;
        double f_dfma = 0, f_dmma = 0;      // flops issued
        if (mode == 0 || mode == 2) f_dfma = thr * 8 * 2;
        if (mode == 1 || mode == 2) f_dmma = thr / 32 * 8 * 512;   // 8 mma per warp-iter, 256 FMA each
        if (mode == 3) { f_dfma = thr / 2 * 8 * 2; f_dmma = thr / 2 / 32 * 8 * 512; }
        printf("%-26s %8.3f ms  DFMA %6.2f TFLOP/s  DMMA %6.2f TFLOP/s  total %6.2f\n", names[mode], best,
               f_dfma / best * 1e-9, f_dmma / best * 1e-9, (f_dfma + f_dmma) / best * 1e-9);
    }
    return 0;
}
