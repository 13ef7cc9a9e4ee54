// Do DMMA (mma.sync.m8n8k4.f64) and DFMA share a datapath on sm_100a?  Per loop iteration: ND independent DMMA and NF independent
// DFMA per warp, two warps per scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_vs_dfma dmma_vs_dfma.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int ND, int NF>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double c[8][2];
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = c[i][1] = 0.0; x[i] = threadIdx.x * 1e-3 + i; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int d = 0; d < ND; d++) dmma(c[d][0], c[d][1], a, x[d & 7]);
#pragma unroll
            for (int f = 0; f < NF; f++) x[f] = fma(x[f], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ND, int NF>
void run(double* out, long long* cyc) {
    const int iters = 2000;
    k<ND, NF><<<1, 256>>>(out, cyc, iters, 0.999, 1e-3);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("two warps per scheduler, %d DMMA + %d DFMA per warp and group: %.1f cycles per group (DMMA alone would need %d, DFMA alone %d)\n", ND, NF,
           (double)c / (iters * 4.0), 2 * ND * 16, 2 * NF * 2);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 64);
    run<4, 0>(out, cyc); run<0, 8>(out, cyc); run<4, 8>(out, cyc); run<2, 8>(out, cyc); run<1, 8>(out, cyc); run<4, 4>(out, cyc); run<1, 4>(out, cyc);
    return 0;
}
