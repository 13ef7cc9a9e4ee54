// Does an fp64 instruction hold the scheduler's issue port for two cycles?  Four independent DFMA chains per thread plus NI
// independent integer (LOP3/IADD3) or fp32 (FFMA) instructions per DFMA group, one or two warps per scheduler.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, bool FP32>
__global__ void k(double* out, long long* cyc, int iters, double a, double b, int ia, float fa) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
    float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            if (!FP32) {
                if (NI > 0) i0 = (i0 ^ ia) + r; if (NI > 1) i1 = (i1 ^ ia) + r; if (NI > 2) i2 = (i2 ^ ia) + r; if (NI > 3) i3 = (i3 ^ ia) + r;
                if (NI > 4) i4 = (i4 ^ ia) + r; if (NI > 5) i5 = (i5 ^ ia) + r; if (NI > 6) i6 = (i6 ^ ia) + r; if (NI > 7) i7 = (i7 ^ ia) + r;
            } else {
                if (NI > 0) f0 = fmaf(f0, fa, 1.f); if (NI > 1) f1 = fmaf(f1, fa, 1.f); if (NI > 2) f2 = fmaf(f2, fa, 1.f); if (NI > 3) f3 = fmaf(f3, fa, 1.f);
                if (NI > 4) f4 = fmaf(f4, fa, 1.f); if (NI > 5) f5 = fmaf(f5, fa, 1.f); if (NI > 6) f6 = fmaf(f6, fa, 1.f); if (NI > 7) f7 = fmaf(f7, fa, 1.f);
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + (i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7) + (f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NI, bool FP32>
void run(int warps, double* out, long long* cyc) {
    const int iters = 2000;
    k<NI, FP32><<<1, warps * 32>>>(out, cyc, iters, 0.999, 1e-3, 12345, 0.999f);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%d warp(s) per scheduler, 4 DFMA + %d %s per group: %.2f cycles per group (fp64 pipe alone needs %d)\n", warps / 4, NI,
           FP32 ? "FFMA" : "integer ops (2 instructions each: LOP3 + IADD)", (double)c / (iters * 8.0), 8 * (warps / 4));
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 64);
    for (int w : {4, 8}) {
        run<0, false>(w, out, cyc); run<2, false>(w, out, cyc); run<4, false>(w, out, cyc); run<8, false>(w, out, cyc);
        run<2, true>(w, out, cyc); run<4, true>(w, out, cyc); run<8, true>(w, out, cyc);
    }
    return 0;
}
