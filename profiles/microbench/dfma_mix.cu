// dfma_mix.cu -- what the B200 FP64 pipe sustains when the DFMA stream shares its warps with other instructions.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_mix dfma_mix.cu && ./dfma_mix
// Per iteration every thread issues C independent DFMAs (C chains) and K integer / select / shared-memory
// instructions (independent chains of their own); 148 CTAs, W warps each.  Reports warp-DFMAs per cycle per SM
// sub-partition (pipe peak 0.5) -- the number the likelihood kernel's 36 % DFMA / 64 % other mix has to live with.
#include <cstdio>
#include <cuda_runtime.h>

template <int C, int KI, int KS, int KL>
__global__ void k(double* out, int iters, double m, long long* cyc) {
    __shared__ double sh[1024];
    double a[C];
    int b[KI > 0 ? KI : 1];
    float f[KS > 0 ? KS : 1];
    double l[KL > 0 ? KL : 1];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = i;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < KI; ++i) b[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < KS; ++i) f[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < KL; ++i) l[i] = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            a[i] = fma(a[i], m, 1e-9);
            if (i < KI) b[i] = b[i] * 3 + it;                       // IMAD
            if (i < KS) f[i] = (b[0] & 1) ? f[i] : -f[i];           // select-like
            if (i < KL) l[i] += sh[(threadIdx.x + 80 * i + it) & 1023];   // LDS + DADD (counts as FP64 too)
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < KI; ++i) s += b[i];
#pragma unroll
    for (int i = 0; i < KS; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < KL; ++i) s += l[i];
    if (s == 12345.678) out[threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int C, int KI, int KS, int KL>
void run(int warps, double* d, long long* dc) {
    const int iters = 4096;
    k<C, KI, KS, KL><<<148, warps * 32>>>(d, iters, 1.0000001, dc);
    k<C, KI, KS, KL><<<148, warps * 32>>>(d, iters, 1.0000001, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double rate = (double)iters * (C + KL) * (warps / 4.0) / c;
    printf("chains %2d  +IMAD %2d +SEL %2d +LDS/DADD %2d  warps/SMSP %.0f: %5.3f warp-FP64/cycle/SMSP (peak 0.5), %5.2f cycles per DFMA per warp\n",
           C, KI, KS, KL, warps / 4.0, rate, (double)c / (iters * C));
}

int main() {
    double* d; long long* dc;
    cudaMalloc(&d, 1 << 16); cudaMalloc(&dc, 8);
    for (int w : {4, 8, 16}) {
        run<1, 0, 0, 0>(w, d, dc); run<2, 0, 0, 0>(w, d, dc); run<4, 0, 0, 0>(w, d, dc); run<8, 0, 0, 0>(w, d, dc); run<16, 0, 0, 0>(w, d, dc);
    }
    for (int w : {8, 16}) {
        run<8, 4, 0, 0>(w, d, dc); run<8, 8, 0, 0>(w, d, dc); run<8, 8, 4, 0>(w, d, dc); run<8, 8, 8, 0>(w, d, dc);
        run<8, 4, 2, 1>(w, d, dc); run<8, 4, 4, 2>(w, d, dc); run<4, 4, 4, 0>(w, d, dc); run<2, 2, 2, 0>(w, d, dc);
    }
    return 0;
}
