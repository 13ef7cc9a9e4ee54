// moments.cu -- per-frame centre of mass and moment tensor, batched over resident frames.
//
// chem.CenterOfMass (/root/reference/geometric.go:609-631): c = (sum_i m_i x_i) / sum_i m_i.
// chem.MomentTensor (geometric.go:705-733): M = sum_i m_i (x_i - c)(x_i - c)^T (the reference scales the centred
// rows by sqrt(m_i) and forms center^T * center).  One sweep per frame: sum m, sum m x (3), sum m x x^T (6), then
// M = S2 - (sum m) c c^T.  HBM-bound: 24 bytes per atom per frame, the mass plane stays in L2.
#include <vector>

#include "common.cuh"

namespace gcb {

namespace {
constexpr int MT = 256;

__global__ void __launch_bounds__(MT, 4) moments_kernel(const double* __restrict__ frames, long stride, int np, long nframes,
                                                       const double* __restrict__ mass, double msum,
                                                       double* __restrict__ center, double* __restrict__ moment) {
    __shared__ double red[MT / 32][9];
    const int np2 = np >> 1;
    const double2* __restrict__ M2 = reinterpret_cast<const double2*>(mass);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long f = blockIdx.x; f < nframes; f += gridDim.x) {
        const double* fr = frames + f * stride;
        const double2* __restrict__ X = reinterpret_cast<const double2*>(fr);
        const double2* __restrict__ Y = reinterpret_cast<const double2*>(fr + np);
        const double2* __restrict__ Z = reinterpret_cast<const double2*>(fr + 2L * np);
        double a[9];
#pragma unroll
        for (int k = 0; k < 9; k++) a[k] = 0.0;
#pragma unroll 2
        for (int j = threadIdx.x; j < np2; j += MT) {
            const double2 x = X[j], y = Y[j], z = Z[j], m = M2[j];
            {
                const double mx = m.x * x.x, my = m.x * y.x, mz = m.x * z.x;
                a[0] += mx; a[1] += my; a[2] += mz;
                a[3] = fma(mx, x.x, a[3]); a[4] = fma(mx, y.x, a[4]); a[5] = fma(mx, z.x, a[5]);
                a[6] = fma(my, y.x, a[6]); a[7] = fma(my, z.x, a[7]); a[8] = fma(mz, z.x, a[8]);
            }
            {
                const double mx = m.y * x.y, my = m.y * y.y, mz = m.y * z.y;
                a[0] += mx; a[1] += my; a[2] += mz;
                a[3] = fma(mx, x.y, a[3]); a[4] = fma(mx, y.y, a[4]); a[5] = fma(mx, z.y, a[5]);
                a[6] = fma(my, y.y, a[6]); a[7] = fma(my, z.y, a[7]); a[8] = fma(mz, z.y, a[8]);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; k++) {
            double v = a[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s[9];
            for (int k = 0; k < 9; k++) {
                double t = 0.0;
                for (int w = 0; w < MT / 32; w++) t += red[w][k];
                s[k] = t;
            }
            const double inv = 1.0 / msum;
            const double cx = s[0] * inv, cy = s[1] * inv, cz = s[2] * inv;
            if (center) { center[3 * f] = cx; center[3 * f + 1] = cy; center[3 * f + 2] = cz; }
            if (moment) {
                double* m = moment + 9 * f;
                const double xx = s[3] - msum * cx * cx, xy = s[4] - msum * cx * cy, xz = s[5] - msum * cx * cz;
                const double yy = s[6] - msum * cy * cy, yz = s[7] - msum * cy * cz, zz = s[8] - msum * cz * cz;
                m[0] = xx; m[1] = xy; m[2] = xz; m[3] = xy; m[4] = yy; m[5] = yz; m[6] = xz; m[7] = yz; m[8] = zz;
            }
        }
        __syncthreads();
    }
}
}  // namespace

}  // namespace gcb

extern "C" int gcb_moments(gcb_traj* t, long first, long nframes, const double* mass, const int* idx, int nidx, double* center,
                           double* moment) {
    using namespace gcb;
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    if (first < 0 || nframes < 0 || first + nframes > t->nframes) { set_error("frame window outside the trajectory"); return GCB_ERR_ARG; }
    if (nframes == 0) return GCB_OK;
    GCB_CUDA(cudaSetDevice(t->dev));
    // mass plane: m_i for the atoms considered, 0 elsewhere (padding included)
    std::vector<double> w((size_t)t->np, 0.0);
    double msum = 0.0;
    if (idx && nidx > 0) {
        for (int k = 0; k < nidx; k++) {
            if (idx[k] < 0 || idx[k] >= t->natoms) { set_error("atom index %d out of range", idx[k]); return GCB_ERR_INDEXES; }
            w[(size_t)idx[k]] += mass ? mass[idx[k]] : 1.0;
        }
    } else {
        for (int i = 0; i < t->natoms; i++) w[(size_t)i] = mass ? mass[i] : 1.0;
    }
    for (int i = 0; i < t->natoms; i++) msum += w[(size_t)i];   // mat.Sum(mass), geometric.go:629
    if (!(msum > 0.0)) { set_error("goChem: the masses add up to %g", msum); return GCB_ERR_SHAPE; }
    double *d_w = nullptr, *d_c = nullptr, *d_m = nullptr;
    int rc = GCB_OK;
    cudaError_t e;
#define MO_TRY(call) do { e = (call); if (e != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e)); rc = GCB_ERR_CUDA; goto done; } } while (0)
    MO_TRY(cudaMalloc(&d_w, sizeof(double) * (size_t)t->np));
    MO_TRY(cudaMalloc(&d_c, sizeof(double) * 3 * (size_t)nframes));
    MO_TRY(cudaMalloc(&d_m, sizeof(double) * 9 * (size_t)nframes));
    MO_TRY(cudaMemcpyAsync(d_w, w.data(), sizeof(double) * (size_t)t->np, cudaMemcpyHostToDevice, t->stream));
    {
        long grid = (long)t->sm_count * 4;
        if (grid > nframes) grid = nframes;
        moments_kernel<<<(unsigned)grid, 256, 0, t->stream>>>(t->d_frames + first * 3L * t->np, 3L * t->np, t->np, nframes, d_w, msum, d_c, d_m);
        GCB_LAUNCHED();
    }
    if (center) MO_TRY(cudaMemcpyAsync(center, d_c, sizeof(double) * 3 * (size_t)nframes, cudaMemcpyDeviceToHost, t->stream));
    if (moment) MO_TRY(cudaMemcpyAsync(moment, d_m, sizeof(double) * 9 * (size_t)nframes, cudaMemcpyDeviceToHost, t->stream));
    MO_TRY(cudaStreamSynchronize(t->stream));
    MO_TRY(cudaGetLastError());
done:
#undef MO_TRY
    cudaFree(d_w); cudaFree(d_c); cudaFree(d_m);
    return rc;
}
