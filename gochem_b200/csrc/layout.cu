// layout.cu -- host-layout <-> device SoA float64 conversion kernels.
//
// Device layout: frame f = three planes x[np], y[np], z[np] of float64 (np = natoms padded to
// 16).  Host layouts: float64 AoS (v3.Matrix, /root/reference/v3/gonum.go:54-58), float32 AoS
// (libxdrfile output, traj/xtc/libgoxtc.c:87-104, widened and scaled by 10 in
// traj/xtc/xtc.go:129-134), float32 SoA planes (DCD records, traj/dcd/dcd.go:380-413, widened
// in :302-307), int32 AoS fixed point (STF text lines, traj/stf/stf.go:324-345, divided by 10^prec).  The widening / scaling / transposition the reference does on the CPU per
// element happens here, after the raw bytes crossed PCIe.
#include "common.cuh"

namespace gcb {

// DIV: fixed-point sources (STF) divide by the scale after widening, float64(f) / p as traj/stf/stf.go:341 does.
template <typename T, bool SOA, bool DIV = false>
__global__ void __launch_bounds__(256) upload_convert_kernel(const T* __restrict__ src, double* __restrict__ dst,
                                                            int natoms, int np, long nframes, double scale) {
    const long total = nframes * (long)np;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const long f = g / np;
        const int i = (int)(g - f * np);
        double x = 0.0, y = 0.0, z = 0.0;  // padding atoms are zero
        if (i < natoms) {
            const T* s = src + f * 3L * natoms;
            if (SOA) {
                x = (double)s[i];
                y = (double)s[natoms + i];
                z = (double)s[2L * natoms + i];
            } else {
                x = (double)s[3L * i];
                y = (double)s[3L * i + 1];
                z = (double)s[3L * i + 2];
            }
            // the reference multiplies after widening: 10*float64(c) (xtc.go:131)
            if (DIV) { x = x / scale; y = y / scale; z = z / scale; }
            else if (scale != 1.0) { x = scale * x; y = scale * y; z = scale * z; }
        }
        double* d = dst + f * 3L * np;
        d[i] = x;
        d[np + i] = y;
        d[2L * np + i] = z;
    }
}

// float32 planes inside fixed-size records (raw DCD frames with their Fortran record markers, optional unit-cell block included):
// record f starts at src + f * record_bytes and holds x[], y[], z[] at byte offsets ox, oy, oz (traj/dcd/dcd.go:380-413).
__global__ void __launch_bounds__(256) upload_records_kernel(const unsigned char* __restrict__ src, double* __restrict__ dst, int natoms,
                                                            int np, long nframes, long record_bytes, long ox, long oy, long oz,
                                                            double scale) {
    const long total = nframes * (long)np;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const long f = g / np;
        const int i = (int)(g - f * np);
        double x = 0.0, y = 0.0, z = 0.0;
        if (i < natoms) {
            const unsigned char* r = src + f * record_bytes;
            x = (double)reinterpret_cast<const float*>(r + ox)[i];
            y = (double)reinterpret_cast<const float*>(r + oy)[i];
            z = (double)reinterpret_cast<const float*>(r + oz)[i];
            if (scale != 1.0) { x = scale * x; y = scale * y; z = scale * z; }
        }
        double* d = dst + f * 3L * np;
        d[i] = x;
        d[np + i] = y;
        d[2L * np + i] = z;
    }
}

__global__ void __launch_bounds__(256) download_convert_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                              int natoms, int np, long nframes) {
    const long total = nframes * (long)natoms;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const long f = g / natoms;
        const int i = (int)(g - f * natoms);
        const double* s = src + f * 3L * np;
        double* d = dst + f * 3L * natoms + 3L * i;
        d[0] = s[i];
        d[1] = s[np + i];
        d[2] = s[2L * np + i];
    }
}

static int grid_for(long total, int sm) {
    long b = (total + 255) / 256;
    long cap = (long)sm * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int launch_upload_convert(gcb_traj* t, cudaStream_t s, const void* d_src, long first, long nframes, int layout,
                          double scale) {
    if (nframes <= 0) return GCB_OK;
    double* dst = t->d_frames + first * 3L * t->np;
    const int grid = grid_for(nframes * (long)t->np, t->sm_count);
    switch (layout) {
        case GCB_F64_AOS:
            upload_convert_kernel<double, false><<<grid, 256, 0, s>>>((const double*)d_src, dst, t->natoms, t->np, nframes, scale);
            break;
        case GCB_F32_AOS:
            upload_convert_kernel<float, false><<<grid, 256, 0, s>>>((const float*)d_src, dst, t->natoms, t->np, nframes, scale);
            break;
        case GCB_F32_SOA:
            upload_convert_kernel<float, true><<<grid, 256, 0, s>>>((const float*)d_src, dst, t->natoms, t->np, nframes, scale);
            break;
        case GCB_I32_AOS:
            if (scale == 0.0) { set_error("fixed-point layout needs a non-zero divisor"); return GCB_ERR_ARG; }
            upload_convert_kernel<int, false, true><<<grid, 256, 0, s>>>((const int*)d_src, dst, t->natoms, t->np, nframes, scale);
            break;
        default:
            set_error("unknown host layout %d", layout);
            return GCB_ERR_ARG;
    }
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int launch_upload_records(gcb_traj* t, cudaStream_t s, const void* d_src, long first, long nframes, long record_bytes, long ox,
                          long oy, long oz, double scale) {
    if (nframes <= 0) return GCB_OK;
    double* dst = t->d_frames + first * 3L * t->np;
    const int grid = grid_for(nframes * (long)t->np, t->sm_count);
    upload_records_kernel<<<grid, 256, 0, s>>>((const unsigned char*)d_src, dst, t->natoms, t->np, nframes, record_bytes, ox, oy, oz, scale);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int launch_download_convert(gcb_traj* t, cudaStream_t s, double* d_dst, long first, long nframes) {
    if (nframes <= 0) return GCB_OK;
    const double* src = t->d_frames + first * 3L * t->np;
    const int grid = grid_for(nframes * (long)t->natoms, t->sm_count);
    download_convert_kernel<<<grid, 256, 0, s>>>(src, d_dst, t->natoms, t->np, nframes);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

}  // namespace gcb
