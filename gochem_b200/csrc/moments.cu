// moments.cu -- per-frame centre of mass and moment tensor, batched over resident frames.
//
// chem.CenterOfMass (/root/reference/geometric.go:609-631): c = (sum_i m_i x_i) / sum_i m_i.
// chem.MomentTensor (geometric.go:705-733): M = sum_i m_i (x_i - c)(x_i - c)^T (the reference scales the centred
// rows by sqrt(m_i) and forms center^T * center).  One sweep per frame: sum m x (3), sum m x x^T (6), then
// M = S2 - (sum m) c c^T.  HBM-bound: 24 bytes per atom per frame; the mass plane (8 bytes per atom) stays in L2.
//
// moments_stream_kernel: one persistent CTA per SM takes frames b, b + G, ...  A producer thread streams each frame
// in chunks of `chunk` atoms -- the three coordinate segments plus the matching segment of the mass plane -- into a
// ring of shared-memory stages with TMA bulk copies (cp.async.bulk + mbarrier complete_tx); 8 consumer warps
// accumulate the 9 sums over the chunks of a frame, reduce them in a fixed order (warp butterfly, then warp order)
// and the first thread writes centre and tensor.  Every byte of a frame crosses HBM -> SM once; the copies of the
// next frames run under the reduction of the current one.
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace gcb {

namespace {
using namespace ptx;

constexpr int kMomWarps = 8;
constexpr int kMomConsumers = 32 * kMomWarps;
constexpr int kMomThreads = kMomConsumers + 32;   // + producer warp
constexpr int kMomMaxStages = 8;
constexpr int kMomSmemBudget = 200 * 1024;

struct MomShared {
    uint64_t full[kMomMaxStages];
    uint64_t empty[kMomMaxStages];
    double red[2][kMomWarps][10];
};

// chunk: atoms per stage (multiple of 256 unless the whole padded frame is one chunk); nchunks = ceil(np / chunk).
__global__ void __launch_bounds__(kMomThreads, 1) moments_stream_kernel(const double* __restrict__ frames, long stride, int np,
                                                                      long nframes, const double* __restrict__ mass, double msum,
                                                                      int chunk, int nchunks, int stages, int origin_atom,
                                                                      double* __restrict__ center, double* __restrict__ moment) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MomShared& sh = *reinterpret_cast<MomShared*>(smem_raw);
    double* stage_base = reinterpret_cast<double*>(smem_raw + ((sizeof(MomShared) + 127) / 128) * 128);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long stage_doubles = 4L * chunk;
    const long nloc = (blockIdx.x < nframes) ? (nframes - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], kMomWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kMomWarps) {
        if (lane == 0) {
            const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_normal();
            int s = 0;
            uint32_t ph = 1;
            long it = 0;
            for (long j = 0; j < nloc; j++) {
                const double* fr = frames + (blockIdx.x + j * (long)gridDim.x) * stride;
                for (int c = 0; c < nchunks; c++, it++) {
                    const int a0 = c * chunk;
                    const int len = (np - a0 < chunk) ? (np - a0) : chunk;   // multiple of 16 (np is)
                    const uint32_t bytes = (uint32_t)len * 8u;
                    if (it >= stages) mbar_wait(&sh.empty[s], ph);
                    double* dst = stage_base + (long)s * stage_doubles;
                    mbar_arrive_expect_tx(&sh.full[s], 4u * bytes);
                    tma_load_1d(dst, fr + a0, bytes, &sh.full[s], pol_stream);
                    tma_load_1d(dst + chunk, fr + np + a0, bytes, &sh.full[s], pol_stream);
                    tma_load_1d(dst + 2 * chunk, fr + 2L * np + a0, bytes, &sh.full[s], pol_stream);
                    tma_load_1d(dst + 3 * chunk, mass + a0, bytes, &sh.full[s], pol_keep);
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }

    // ---- consumers ----
    int s = 0;
    uint32_t ph = 0;
    for (long j = 0; j < nloc; j++) {
        double a[9];
#pragma unroll
        for (int k = 0; k < 9; k++) a[k] = 0.0;
        // The second moments are taken about a point INSIDE the selection (its first atom), not about the coordinate origin: the
        // reference centres before it multiplies (geometric.go:719-731), and S2 - M c c^T from raw coordinates would lose
        // (|c| / extent)^2 * eps for a small selection far from the origin (a ligand in a large box).  The shift is undone
        // exactly below.
        const double* fr0 = frames + (blockIdx.x + j * (long)gridDim.x) * stride;
        const double ox = __ldg(fr0 + origin_atom), oy = __ldg(fr0 + np + origin_atom), oz = __ldg(fr0 + 2L * np + origin_atom);
        for (int c = 0; c < nchunks; c++) {
            const int len = (np - c * chunk < chunk) ? (np - c * chunk) : chunk;
            mbar_wait(&sh.full[s], ph);
            const double* X = stage_base + (long)s * stage_doubles;
            const double* Y = X + chunk;
            const double* Z = Y + chunk;
            const double* M = Z + chunk;
#pragma unroll 4
            for (int i = tid; i < len; i += kMomConsumers) {
                const double x = X[i] - ox, y = Y[i] - oy, z = Z[i] - oz, m = M[i];
                const double mx = m * x, my = m * y, mz = m * z;
                a[0] += mx; a[1] += my; a[2] += mz;
                a[3] = fma(mx, x, a[3]); a[4] = fma(mx, y, a[4]); a[5] = fma(mx, z, a[5]);
                a[6] = fma(my, y, a[6]); a[7] = fma(my, z, a[7]); a[8] = fma(mz, z, a[8]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty[s]);
            if (++s == stages) { s = 0; ph ^= 1u; }
        }
        // fixed-order reduction: xor butterfly inside the warp, then the 8 warps in order
#pragma unroll
        for (int k = 0; k < 9; k++) {
            double v = a[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            a[k] = v;
        }
        double (*red)[10] = sh.red[j & 1];
        if (lane < 9) {
            double v = a[0];
#pragma unroll
            for (int k = 1; k < 9; k++) v = (lane == k) ? a[k] : v;
            red[warp][lane] = v;
        }
        named_bar_sync(1, kMomConsumers);   // red[j & 1] complete; the buffer of frame j-1 is free again after this barrier
        if (warp == 0) {
            double t = 0.0;
            if (lane < 9) {
#pragma unroll
                for (int w = 0; w < kMomWarps; w++) t += red[w][lane];
            }
            double sum[9];
#pragma unroll
            for (int k = 0; k < 9; k++) sum[k] = __shfl_sync(0xffffffffu, t, k);
            if (lane == 0) {
                const long f = blockIdx.x + j * (long)gridDim.x;
                const double inv = 1.0 / msum;
                const double cx = sum[0] * inv, cy = sum[1] * inv, cz = sum[2] * inv;   // centre relative to the frame's origin atom
                if (center) { center[3 * f] = ox + cx; center[3 * f + 1] = oy + cy; center[3 * f + 2] = oz + cz; }
                if (moment) {
                    double* m = moment + 9 * f;
                    const double xx = sum[3] - msum * cx * cx, xy = sum[4] - msum * cx * cy, xz = sum[5] - msum * cx * cz;
                    const double yy = sum[6] - msum * cy * cy, yz = sum[7] - msum * cy * cz, zz = sum[8] - msum * cz * cz;
                    m[0] = xx; m[1] = xy; m[2] = xz; m[3] = xy; m[4] = yy; m[5] = yz; m[6] = xz; m[7] = yz; m[8] = zz;
                }
            }
        }
    }
}

int ensure_moment_buffers(gcb_traj* t, long nframes) {
    if (!t->d_mom_w) GCB_CUDA(cudaMalloc(&t->d_mom_w, sizeof(double) * (size_t)t->np));
    if (nframes > t->mom_cap) {
        cudaFree(t->d_mom_c); cudaFree(t->d_mom_m);
        t->d_mom_c = t->d_mom_m = nullptr;
        t->mom_cap = 0;
        GCB_CUDA(cudaMalloc(&t->d_mom_c, sizeof(double) * 3 * (size_t)nframes));
        GCB_CUDA(cudaMalloc(&t->d_mom_m, sizeof(double) * 9 * (size_t)nframes));
        t->mom_cap = nframes;
    }
    return GCB_OK;
}

}  // namespace

}  // namespace gcb

extern "C" int gcb_moments_async(gcb_traj* t, long first, long nframes, const double* mass, const int* idx, int nidx) {
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
    int origin_atom = 0;
    for (int i = t->natoms - 1; i >= 0; i--) if (w[(size_t)i] != 0.0) origin_atom = i;   // first atom that counts
    for (int i = 0; i < t->natoms; i++) msum += w[(size_t)i];   // mat.Sum(mass), geometric.go:629
    if (!(msum > 0.0)) { set_error("goChem: the masses add up to %g", msum); return GCB_ERR_SHAPE; }
    int rc = ensure_moment_buffers(t, nframes);
    if (rc) return rc;
    GCB_CUDA(cudaStreamSynchronize(t->stream));   // the previous call may still read the mass plane
    GCB_CUDA(cudaMemcpyAsync(t->d_mom_w, w.data(), sizeof(double) * (size_t)t->np, cudaMemcpyHostToDevice, t->stream));
    GCB_CUDA(cudaStreamSynchronize(t->stream));   // w is a local
    // stage shape: chunks of up to 2048 atoms (64 KB with the mass segment), as many stages as fit
    int chunk = t->np <= 2048 ? t->np : 2048;
    const int nchunks = (t->np + chunk - 1) / chunk;
    const size_t fixed = ((sizeof(MomShared) + 127) / 128) * 128;
    int stages = (int)((kMomSmemBudget - fixed) / (sizeof(double) * 4 * (size_t)chunk));
    if (stages > kMomMaxStages) stages = kMomMaxStages;
    if (stages < 2) { set_error("moments: no stage plan for %d atoms", t->natoms); return GCB_ERR_ARG; }
    const size_t smem = fixed + sizeof(double) * 4 * (size_t)chunk * stages;
    GCB_CUDA(opt_in_smem(moments_stream_kernel, t->dev));
    rc = deps_before_compute(t, first, nframes);
    if (rc) return rc;
    long grid = t->sm_count;
    if (grid > nframes) grid = nframes;
    moments_stream_kernel<<<(unsigned)grid, kMomThreads, smem, t->stream>>>(t->d_frames + first * 3L * t->np, 3L * t->np, t->np, nframes,
                                                                            t->d_mom_w, msum, chunk, nchunks, stages, origin_atom, t->d_mom_c, t->d_mom_m);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    t->mom_frames = nframes;
    return deps_after_compute(t, first, nframes);
}

extern "C" int gcb_moments_fetch(gcb_traj* t, long nframes, double* center, double* moment) {
    using namespace gcb;
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    if (nframes < 0 || nframes > t->mom_frames) { set_error("gcb_moments_fetch: %ld frames asked, %ld computed", nframes, t->mom_frames); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    if (center && nframes) GCB_CUDA(cudaMemcpyAsync(center, t->d_mom_c, sizeof(double) * 3 * (size_t)nframes, cudaMemcpyDeviceToHost, t->stream));
    if (moment && nframes) GCB_CUDA(cudaMemcpyAsync(moment, t->d_mom_m, sizeof(double) * 9 * (size_t)nframes, cudaMemcpyDeviceToHost, t->stream));
    GCB_CUDA(cudaStreamSynchronize(t->stream));
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

extern "C" int gcb_moments(gcb_traj* t, long first, long nframes, const double* mass, const int* idx, int nidx, double* center,
                           double* moment) {
    int rc = gcb_moments_async(t, first, nframes, mass, idx, nidx);
    if (rc) return rc;
    return gcb_moments_fetch(t, nframes, center, moment);
}
