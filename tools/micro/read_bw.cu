// Read-only HBM stream ceiling of one GPU: every thread sums 16-byte loads of a buffer far larger than L2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_bw read_bw.cu ; ./read_bw [GiB]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int UNROLL>
__global__ void __launch_bounds__(512) read_kernel(const double2* __restrict__ p, size_t n, double* out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        double2 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) s += v[u].x + v[u].y;
    }
    for (; i < n; i += stride) { const double2 v = __ldcs(p + i); s += v.x + v.y; }
    if (s == 1.2345e-300) out[0] = s;   // never true: keeps the loads
}
int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 16.0;
    const size_t bytes = (size_t)(gib * (1ull << 30));
    double2* p; double* out;
    cudaMalloc(&p, bytes); cudaMalloc(&out, 8);
    cudaMemset(p, 0, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas_per_sm : {1, 2, 3, 4}) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            for (int it = 0; it < 5; it++) read_kernel<8><<<sms * ctas_per_sm, 512>>>(p, bytes / 16, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("read-only stream, %.0f GiB, %d x %d CTAs of 512 threads, 8 x 16 B in flight per thread: %.1f GB/s\n", gib, sms, ctas_per_sm,
                            5.0 * bytes / (ms * 1e-3) / 1e9);
        }
    }
    // a device-to-device copy of half the buffer for comparison (read + write bytes)
    cudaEventRecord(e0);
    for (int it = 0; it < 5; it++) cudaMemcpyAsync((char*)p + bytes / 2, p, bytes / 2, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy device to device, read + write bytes: %.1f GB/s\n", 5.0 * bytes / (ms * 1e-3) / 1e9);
    return 0;
}
