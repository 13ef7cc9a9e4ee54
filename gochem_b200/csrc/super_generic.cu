// super_generic.cu -- shape-agnostic fused superposition + RMSD kernels (one CTA per frame).
//
// What one frame goes through replaces, per frame, chem.Super (/root/reference/handy.go:175-214)
// -> chem.RotatorTranslatorToSuper (geometric.go:127-208) -> chem.MemRMSD (geometric.go:258-288)
// and, in the LOVO pass, chem.MemPerAtomRMSD (geometric.go:294-321) as called from
// align.concproc (align/lovo.go:265-281):
//
//   pass 1  one sweep over the frame: S = sum w_i a_i and H = sum (w_i a_i) b'_i^T, where
//           b' = templa - centroid(templa fit subset) is prepared once per call.  Because
//           sum w_i b'_i = 0, H equals the centred correlation ctest^T ctempla of
//           geometric.go:149-153 without knowing the frame's centroid in advance.
//   solve   thread 0: centroid = S/n, R from H (kabsch3.cuh), trans1 = -centroid,
//           trans2 = centroid(templa subset)   (geometric.go:191-206; the reference adds
//           mean(ctempla - ctest R), which is rounding residue of order 1e-15 A).
//   pass 2  (a_i - centroid) R - b'_i is the explicit residual (geometric.go:284-286 computes the
//           RMSD from the transformed coordinates, not from the singular values); its squared
//           norm feeds the per-frame RMSD, its norm the LOVO per-atom sums, and
//           (a_i - centroid) R + trans2 is written back when the caller wants the reference's
//           in-place Super semantics (handy.go:207-211).
//
// This variant re-reads the frame for pass 2 (L2 hit when the resident CTAs' frames fit in L2).
// It works for any atom count and any index lists and is the fallback of the cluster kernel
// (super_cluster.cu), which keeps the frame in shared memory instead.
#include "common.cuh"
#include "kabsch3.cuh"

namespace gcb {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reduction of NV doubles per thread: shuffle tree inside each warp, then a
// fixed-order sum over the warps.  Result in out[0..NV) (shared), visible to all threads on return.
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) red[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) s += red[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// thread 0: turn the reduced sums into the frame's transform; publish R and the centroid in xf[12].
// one_pass: sums[12] = sum w|a|^2 is there and the RMSD may come from the inner products (E = Ga + Gb - 2 tr(R^T H)) when the
// rounding bound of that form moves it by < 1e-10 A -- the rule of super_cluster.cu; xf[12] = 1 then, and the second sweep is skipped.
__device__ __forceinline__ void solve_frame(const SuperParams& p, long f, const double* sums, double* xf, bool one_pass = false) {
    double cen[3] = {sums[0] * p.inv_n, sums[1] * p.inv_n, sums[2] * p.inv_n};
    double h[9], r[9];
    bool finite = true;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        h[k] = sums[3 + k];
        finite = finite && (fabs(h[k]) < 1.79e308);
    }
    int rc = finite ? k3::kabsch3(h, r) : -4;
    if (rc != 0) {
        if (p.status) atomicExch(p.status, (int)GCB_ERR_SVD);
#pragma unroll
        for (int k = 0; k < 9; k++) r[k] = (k % 4 == 0) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) xf[k] = r[k];
    xf[9] = cen[0]; xf[10] = cen[1]; xf[11] = cen[2];
    if (p.rot) {
#pragma unroll
        for (int k = 0; k < 9; k++) p.rot[f * 9 + k] = r[k];
    }
    if (p.trans1) { p.trans1[f * 3] = -cen[0]; p.trans1[f * 3 + 1] = -cen[1]; p.trans1[f * 3 + 2] = -cen[2]; }
    if (p.trans2) { p.trans2[f * 3] = p.bbar[0]; p.trans2[f * 3 + 1] = p.bbar[1]; p.trans2[f * 3 + 2] = p.bbar[2]; }
    bool formula_ok = false;
    if (one_pass) {
        const double ga = sums[12] - (sums[0] * sums[0] + sums[1] * sums[1] + sums[2] * sums[2]) * p.inv_n;
        const double trrh = fma(r[8], h[8], fma(r[7], h[7], fma(r[6], h[6], fma(r[5], h[5], fma(r[4], h[4],
                            fma(r[3], h[3], fma(r[2], h[2], fma(r[1], h[1], r[0] * h[0]))))))));
        const double e = ga + p.gb - 2.0 * trrh;
        const double smax = sums[12] + p.gb;
        formula_ok = rc == 0 && e > 0.0 && (64.0 * 2.220446049250313e-16) * smax <= 2e-10 * sqrt(p.n_fit * e);
        if (formula_ok && p.rmsd) p.rmsd[f] = sqrt(e * p.inv_n);
        xf[12] = formula_ok ? 1.0 : 0.0;
    }
    if (p.need_explicit) p.need_explicit[f] = formula_ok ? 0 : 1;
}

template <bool WEIGHTED, bool WRITE_BACK, bool DEV>
__global__ void __launch_bounds__(kThreads, 3) super_generic_kernel(const SuperParams p, int one_pass_arg) {
    __shared__ double s_red[kWarps * 13];
    __shared__ double s_sum[13];
    __shared__ double s_xf[13];
    const bool one_pass = !WRITE_BACK && !DEV && one_pass_arg != 0;   // the second sweep only when the RMSD needs it
    const int np = p.np, np2 = np >> 1;
    const double2* __restrict__ BX = reinterpret_cast<const double2*>(p.refc);
    const double2* __restrict__ BY = reinterpret_cast<const double2*>(p.refc + np);
    const double2* __restrict__ BZ = reinterpret_cast<const double2*>(p.refc + 2L * np);
    const double2* __restrict__ W = reinterpret_cast<const double2*>(p.weight);

    for (long f = blockIdx.x; f < p.nframes; f += gridDim.x) {
        const double* fr = p.frames + f * p.frame_stride;
        const double2* __restrict__ X = reinterpret_cast<const double2*>(fr);
        const double2* __restrict__ Y = reinterpret_cast<const double2*>(fr + np);
        const double2* __restrict__ Z = reinterpret_cast<const double2*>(fr + 2L * np);
        double acc[13];
#pragma unroll
        for (int k = 0; k < 13; k++) acc[k] = 0.0;
#pragma unroll 2
        for (int j = threadIdx.x; j < np2; j += kThreads) {
            double2 x = X[j], y = Y[j], z = Z[j];
            const double2 bx = BX[j], by = BY[j], bz = BZ[j];
            if (!WRITE_BACK && !DEV) {   // sum w |a|^2 for the one-pass RMSD
                const double n0 = fma(z.x, z.x, fma(y.x, y.x, x.x * x.x)), n1 = fma(z.y, z.y, fma(y.y, y.y, x.y * x.y));
                if (WEIGHTED) { const double2 w = W[j]; acc[12] = fma(w.y, n1, fma(w.x, n0, acc[12])); }
                else acc[12] += n0 + n1;
            }
            if (WEIGHTED) {
                const double2 w = W[j];
                x.x *= w.x; y.x *= w.x; z.x *= w.x;
                x.y *= w.y; y.y *= w.y; z.y *= w.y;
            }
            acc[0] += x.x; acc[1] += y.x; acc[2] += z.x;
            acc[3] = fma(x.x, bx.x, acc[3]); acc[4] = fma(x.x, by.x, acc[4]); acc[5] = fma(x.x, bz.x, acc[5]);
            acc[6] = fma(y.x, bx.x, acc[6]); acc[7] = fma(y.x, by.x, acc[7]); acc[8] = fma(y.x, bz.x, acc[8]);
            acc[9] = fma(z.x, bx.x, acc[9]); acc[10] = fma(z.x, by.x, acc[10]); acc[11] = fma(z.x, bz.x, acc[11]);
            acc[0] += x.y; acc[1] += y.y; acc[2] += z.y;
            acc[3] = fma(x.y, bx.y, acc[3]); acc[4] = fma(x.y, by.y, acc[4]); acc[5] = fma(x.y, bz.y, acc[5]);
            acc[6] = fma(y.y, bx.y, acc[6]); acc[7] = fma(y.y, by.y, acc[7]); acc[8] = fma(y.y, bz.y, acc[8]);
            acc[9] = fma(z.y, bx.y, acc[9]); acc[10] = fma(z.y, by.y, acc[10]); acc[11] = fma(z.y, bz.y, acc[11]);
        }
        block_reduce<13>(acc, s_red, s_sum);
        if (threadIdx.x == 0) solve_frame(p, f, s_sum, s_xf, one_pass);
        __syncthreads();
        if (one_pass && s_xf[12] != 0.0) continue;   // block-uniform: the RMSD is already written
        const double r0 = s_xf[0], r1 = s_xf[1], r2 = s_xf[2], r3 = s_xf[3], r4 = s_xf[4], r5 = s_xf[5], r6 = s_xf[6],
                     r7 = s_xf[7], r8 = s_xf[8];
        const double cx = s_xf[9], cy = s_xf[10], cz = s_xf[11];
        double* frw = WRITE_BACK ? (p.frames_rw + f * p.frame_stride) : nullptr;
        double* devrow = DEV ? (p.devpart + (long)blockIdx.x * np) : nullptr;
        double e[1] = {0.0};
        for (int j = threadIdx.x; j < np2; j += kThreads) {
            const double2 x = X[j], y = Y[j], z = Z[j];
            const double2 bx = BX[j], by = BY[j], bz = BZ[j];
            double2 w = make_double2(1.0, 1.0);
            if (WEIGHTED) w = W[j];
            const bool v0 = 2 * j < p.natoms, v1 = 2 * j + 1 < p.natoms;
            double2 qx, qy, qz, d;
            {
                const double a0 = x.x - cx, a1 = y.x - cy, a2 = z.x - cz;
                qx.x = fma(a2, r6, fma(a1, r3, a0 * r0));
                qy.x = fma(a2, r7, fma(a1, r4, a0 * r1));
                qz.x = fma(a2, r8, fma(a1, r5, a0 * r2));
                const double dx = qx.x - bx.x, dy = qy.x - by.x, dz = qz.x - bz.x;
                d.x = fma(dz, dz, fma(dy, dy, dx * dx));
            }
            {
                const double a0 = x.y - cx, a1 = y.y - cy, a2 = z.y - cz;
                qx.y = fma(a2, r6, fma(a1, r3, a0 * r0));
                qy.y = fma(a2, r7, fma(a1, r4, a0 * r1));
                qz.y = fma(a2, r8, fma(a1, r5, a0 * r2));
                const double dx = qx.y - bx.y, dy = qy.y - by.y, dz = qz.y - bz.y;
                d.y = fma(dz, dz, fma(dy, dy, dx * dx));
            }
            if (WEIGHTED) {
                e[0] = fma(w.x, d.x, e[0]);
                e[0] = fma(w.y, d.y, e[0]);
            } else {
                if (v0) e[0] += d.x;
                if (v1) e[0] += d.y;
            }
            if (DEV) {
                double2* dr = reinterpret_cast<double2*>(devrow) + j;
                double2 cur = *dr;
                if (v0) cur.x += sqrt(d.x);
                if (v1) cur.y += sqrt(d.y);
                *dr = cur;
            }
            if (WRITE_BACK) {
                qx.x += p.bbar[0]; qy.x += p.bbar[1]; qz.x += p.bbar[2];
                qx.y += p.bbar[0]; qy.y += p.bbar[1]; qz.y += p.bbar[2];
                if (!v0) { qx.x = 0.0; qy.x = 0.0; qz.x = 0.0; }
                if (!v1) { qx.y = 0.0; qy.y = 0.0; qz.y = 0.0; }
                reinterpret_cast<double2*>(frw)[j] = qx;
                reinterpret_cast<double2*>(frw + np)[j] = qy;
                reinterpret_cast<double2*>(frw + 2L * np)[j] = qz;
            }
        }
        block_reduce<1>(e, s_red, s_sum);
        if (threadIdx.x == 0 && p.rmsd) p.rmsd[f] = sqrt(s_sum[0] * p.inv_n);
        // s_sum / s_xf are rewritten only after the next frame's first __syncthreads
    }
}

// General index lists (idx_test[k] pairs with idx_ref[k]): gather formulation.
// refpairs: compact centred template planes bx'[k], by'[k], bz'[k] (k < nidx).
template <bool WRITE_BACK>
__global__ void __launch_bounds__(kThreads, 3)
super_gather_kernel(const SuperParams p, const double* __restrict__ refpairs, const int* __restrict__ idx, int nidx) {
    __shared__ double s_red[kWarps * 12];
    __shared__ double s_sum[12];
    __shared__ double s_xf[12];
    const int np = p.np;
    for (long f = blockIdx.x; f < p.nframes; f += gridDim.x) {
        const double* fr = p.frames + f * p.frame_stride;
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = 0.0;
        for (int k = threadIdx.x; k < nidx; k += kThreads) {
            const int i = idx[k];
            const double x = fr[i], y = fr[np + i], z = fr[2L * np + i];
            const double bx = refpairs[k], by = refpairs[nidx + k], bz = refpairs[2L * nidx + k];
            acc[0] += x; acc[1] += y; acc[2] += z;
            acc[3] = fma(x, bx, acc[3]); acc[4] = fma(x, by, acc[4]); acc[5] = fma(x, bz, acc[5]);
            acc[6] = fma(y, bx, acc[6]); acc[7] = fma(y, by, acc[7]); acc[8] = fma(y, bz, acc[8]);
            acc[9] = fma(z, bx, acc[9]); acc[10] = fma(z, by, acc[10]); acc[11] = fma(z, bz, acc[11]);
        }
        block_reduce<12>(acc, s_red, s_sum);
        if (threadIdx.x == 0) solve_frame(p, f, s_sum, s_xf);
        __syncthreads();
        const double r0 = s_xf[0], r1 = s_xf[1], r2 = s_xf[2], r3 = s_xf[3], r4 = s_xf[4], r5 = s_xf[5], r6 = s_xf[6],
                     r7 = s_xf[7], r8 = s_xf[8];
        const double cx = s_xf[9], cy = s_xf[10], cz = s_xf[11];
        double e[1] = {0.0};
        for (int k = threadIdx.x; k < nidx; k += kThreads) {
            const int i = idx[k];
            const double a0 = fr[i] - cx, a1 = fr[np + i] - cy, a2 = fr[2L * np + i] - cz;
            const double dx = fma(a2, r6, fma(a1, r3, a0 * r0)) - refpairs[k];
            const double dy = fma(a2, r7, fma(a1, r4, a0 * r1)) - refpairs[nidx + k];
            const double dz = fma(a2, r8, fma(a1, r5, a0 * r2)) - refpairs[2L * nidx + k];
            e[0] += fma(dz, dz, fma(dy, dy, dx * dx));
        }
        block_reduce<1>(e, s_red, s_sum);  // also orders pass 2a before the in-place rewrite below
        if (threadIdx.x == 0 && p.rmsd) p.rmsd[f] = sqrt(s_sum[0] * p.inv_n);
        if (WRITE_BACK) {
            double* frw = p.frames_rw + f * p.frame_stride;
            for (int i = threadIdx.x; i < p.natoms; i += kThreads) {
                const double a0 = frw[i] - cx, a1 = frw[np + i] - cy, a2 = frw[2L * np + i] - cz;
                frw[i] = fma(a2, r6, fma(a1, r3, a0 * r0)) + p.bbar[0];
                frw[np + i] = fma(a2, r7, fma(a1, r4, a0 * r1)) + p.bbar[1];
                frw[2L * np + i] = fma(a2, r8, fma(a1, r5, a0 * r2)) + p.bbar[2];
            }
        }
        __syncthreads();
    }
}

// chem.RMSD / chem.MemRMSD for every frame (geometric.go:258-288): no fit, ||a - b||_F / sqrt(n).
// ref: raw (uncentred) template planes, 3 * np.
template <bool WEIGHTED>
__global__ void __launch_bounds__(kThreads, 4)
rmsd_nofit_kernel(const double* __restrict__ frames, long frame_stride, int np, long nframes,
                  const double* __restrict__ ref, const double* __restrict__ weight, double inv_n, int natoms,
                  double* __restrict__ rmsd) {
    __shared__ double s_red[kWarps];
    __shared__ double s_sum[1];
    for (long f = blockIdx.x; f < nframes; f += gridDim.x) {
        const double* fr = frames + f * frame_stride;
        double e[1] = {0.0};
        for (int i = threadIdx.x; i < natoms; i += kThreads) {
            const double dx = fr[i] - ref[i], dy = fr[np + i] - ref[np + i], dz = fr[2L * np + i] - ref[2L * np + i];
            const double d = fma(dz, dz, fma(dy, dy, dx * dx));
            e[0] = WEIGHTED ? fma(weight[i], d, e[0]) : e[0] + d;
        }
        block_reduce<1>(e, s_red, s_sum);
        if (threadIdx.x == 0) rmsd[f] = sqrt(s_sum[0] * inv_n);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads, 4)
rmsd_nofit_gather_kernel(const double* __restrict__ frames, long frame_stride, int np, long nframes,
                         const double* __restrict__ refpairs, const int* __restrict__ idx, int nidx,
                         double* __restrict__ rmsd) {
    __shared__ double s_red[kWarps];
    __shared__ double s_sum[1];
    const double inv_n = 1.0 / (double)nidx;
    for (long f = blockIdx.x; f < nframes; f += gridDim.x) {
        const double* fr = frames + f * frame_stride;
        double e[1] = {0.0};
        for (int k = threadIdx.x; k < nidx; k += kThreads) {
            const int i = idx[k];
            const double dx = fr[i] - refpairs[k], dy = fr[np + i] - refpairs[nidx + k],
                         dz = fr[2L * np + i] - refpairs[2L * nidx + k];
            e[0] += fma(dz, dz, fma(dy, dy, dx * dx));
        }
        block_reduce<1>(e, s_red, s_sum);
        if (threadIdx.x == 0) rmsd[f] = sqrt(s_sum[0] * inv_n);
        __syncthreads();
    }
}

// chem.MemPerAtomRMSD without superposition (geometric.go:294-321): out[k] = |test[idx[k]] - templa_k|
__global__ void __launch_bounds__(256) peratom_dist_kernel(const double* __restrict__ frame, int np,
                                                          const double* __restrict__ refpairs, const int* __restrict__ idx,
                                                          int nidx, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nidx) return;
    const int i = idx[k];
    const double dx = frame[i] - refpairs[k], dy = frame[np + i] - refpairs[nidx + k],
                 dz = frame[2L * np + i] - refpairs[2L * nidx + k];
    out[k] = sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
}

int launch_peratom_dist(gcb_traj* t, const double* frame, const double* refpairs, const int* idx, int nidx, double* out) {
    peratom_dist_kernel<<<(nidx + 255) / 256, 256, 0, t->stream>>>(frame, t->np, refpairs, idx, nidx, out);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

// Fixed-order sum of the per-CTA per-atom rows (deterministic for a fixed launch shape).
__global__ void __launch_bounds__(256) dev_reduce_kernel(const double* __restrict__ part, int rows, int np, int natoms,
                                                        double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= natoms) return;
    double s = 0.0;
    for (int r = 0; r < rows; r++) s += part[(long)r * np + i];
    out[i] = s;
}

static int frames_grid(const gcb_traj* t, long nframes, int ctas_per_sm) {
    long g = (long)t->sm_count * ctas_per_sm;
    if (g > nframes) g = nframes;
    if (g < 1) g = 1;
    return (int)g;
}

int launch_super_generic(gcb_traj* t, const SuperParams& p, bool weighted, bool want_dev, int* grid_out) {
    const int grid = frames_grid(t, p.nframes, 3);
    if (grid_out) *grid_out = grid;
    const bool wb = p.frames_rw != nullptr;
#define GCB_GEN(W, B, D) super_generic_kernel<W, B, D><<<grid, kThreads, 0, t->stream>>>(p, t->tune.force_explicit ? 0 : 1)
    if (weighted) {
        if (wb) { if (want_dev) GCB_GEN(true, true, true); else GCB_GEN(true, true, false); }
        else    { if (want_dev) GCB_GEN(true, false, true); else GCB_GEN(true, false, false); }
    } else {
        if (wb) { if (want_dev) GCB_GEN(false, true, true); else GCB_GEN(false, true, false); }
        else    { if (want_dev) GCB_GEN(false, false, true); else GCB_GEN(false, false, false); }
    }
#undef GCB_GEN
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int launch_super_gather(gcb_traj* t, const SuperParams& p, const double* refpairs, const int* idx, int nidx) {
    const int grid = frames_grid(t, p.nframes, 3);
    if (p.frames_rw) super_gather_kernel<true><<<grid, kThreads, 0, t->stream>>>(p, refpairs, idx, nidx);
    else super_gather_kernel<false><<<grid, kThreads, 0, t->stream>>>(p, refpairs, idx, nidx);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int launch_rmsd_nofit(gcb_traj* t, const double* frames, long nframes, const double* ref_planes, const double* weight,
                      double n_fit, double* rmsd) {
    const int grid = frames_grid(t, nframes, 4);
    if (weight)
        rmsd_nofit_kernel<true><<<grid, kThreads, 0, t->stream>>>(frames, 3L * t->np, t->np, nframes, ref_planes, weight, 1.0 / n_fit, t->natoms, rmsd);
    else
        rmsd_nofit_kernel<false><<<grid, kThreads, 0, t->stream>>>(frames, 3L * t->np, t->np, nframes, ref_planes, nullptr, 1.0 / n_fit, t->natoms, rmsd);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int launch_rmsd_nofit_gather(gcb_traj* t, const double* frames, long nframes, const double* refpairs_raw, const int* idx,
                             int nidx, double* rmsd) {
    const int grid = frames_grid(t, nframes, 4);
    rmsd_nofit_gather_kernel<<<grid, kThreads, 0, t->stream>>>(frames, 3L * t->np, t->np, nframes, refpairs_raw, idx, nidx, rmsd);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int launch_dev_reduce_n(gcb_traj* t, int rows, int stride, int n, double* out) {
    dev_reduce_kernel<<<(n + 255) / 256, 256, 0, t->stream>>>(t->d_devpart, rows, stride, n, out);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int launch_dev_reduce(gcb_traj* t, int rows, double* out) {
    dev_reduce_kernel<<<(t->natoms + 255) / 256, 256, 0, t->stream>>>(t->d_devpart, rows, t->np, t->natoms, out);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

}  // namespace gcb
