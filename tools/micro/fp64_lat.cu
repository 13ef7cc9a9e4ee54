// DFMA latency / issue rate on one SM: CHAINS independent dependent-chains per thread, W warps on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-3 + c;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
void run(int warps, double* out, long long* cyc) {
    int iters = 2000;
    k<CH><<<1, warps * 32>>>(out, cyc, iters, 0.999, 1e-3);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / (iters * 8.0 * CH);   // cycles per DFMA of one warp
    printf("chains %2d warps %2d (%d/sched): %.2f cycles per DFMA per warp, round of %d chains = %.1f cycles, SMSP DFMA issue interval %.2f\n", CH, warps, warps / 4, per, CH,
           per * CH, per / (warps / 4.0));
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 64);
    for (int w : {4, 8, 12, 16}) {
        run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); run<12>(w, out, cyc); run<16>(w, out, cyc);
    }
    return 0;
}
