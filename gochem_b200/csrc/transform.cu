// transform.cu -- the second half of a superposition whose frames are too large for the cluster kernel.
//
// A frame of more than 28 672 atoms (a solvated system) does not fit the shared-memory stages of
// super_cluster.cu, and the one-CTA-per-frame kernel re-reads it from HBM for its second sweep.  Such calls are
// split instead (api.cu: launch_super):
//
//   1. the fit itself -- rotation, centroid and RMSD of every frame -- by whichever fit-only kernel suits the
//      index lists (the gather kernels when the fit uses a small subset of the frame, which is the usual case:
//      a protein inside its solvent box; the one-pass generic kernel otherwise), and then
//   2. one of the elementwise sweeps below over the whole frames, driven by the per-frame transforms of step 1:
//        transform_frames_kernel  test = (test + trans1) R + trans2 for every atom, in place -- the in-place
//                                 semantics of chem.Super (/root/reference/handy.go:207-211); 48 B per atom
//        peratom_accum_kernel     sum over frames of |(test_i + trans1) R + trans2 - templa_i| per atom --
//                                 chem.MemPerAtomRMSD after Super as align.concproc calls them
//                                 (/root/reference/align/lovo.go:265-281, geometric.go:294-321); 24 B per atom
//
// Both are HBM-bound streams with no reuse (streaming loads and stores).  Sums are formed in a fixed order:
// results are deterministic for a fixed launch shape.
#include "common.cuh"
#include "ptx.cuh"

namespace gcb {

namespace {

constexpr int kTfThreads = 256;
constexpr int kTfCols = 2;                          // double2 columns per thread and tile
constexpr int kTfTile = kTfThreads * kTfCols;       // double2 columns per tile = 1024 atoms

// One tile = kTfTile double2 columns of one frame.  Tiles are dealt to a persistent grid in order, so neighbouring CTAs
// work on neighbouring pieces of the same frame.
__global__ void __launch_bounds__(kTfThreads) transform_frames_kernel(double* __restrict__ frames, long frame_stride, int np, int natoms,
                                                                     long nframes, const double* __restrict__ rot,
                                                                     const double* __restrict__ trans1, double b0, double b1, double b2) {
    const int np2 = np >> 1;
    const int tiles_per_frame = (np2 + kTfTile - 1) / kTfTile;
    const long ntiles = nframes * tiles_per_frame;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long f = tile / tiles_per_frame;
        const int c0 = (int)(tile - f * tiles_per_frame) * kTfTile + threadIdx.x;
        double* fr = frames + f * frame_stride;
        double2* X = reinterpret_cast<double2*>(fr);
        double2* Y = reinterpret_cast<double2*>(fr + np);
        double2* Z = reinterpret_cast<double2*>(fr + 2L * np);
        double2 x[kTfCols], y[kTfCols], z[kTfCols];
#pragma unroll
        for (int c = 0; c < kTfCols; c++) {
            const int j = c0 + c * kTfThreads;
            if (j < np2) { x[c] = __ldcs(X + j); y[c] = __ldcs(Y + j); z[c] = __ldcs(Z + j); }
        }
        const double* r = rot + 9 * f;
        const double r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3), r4 = __ldg(r + 4), r5 = __ldg(r + 5),
                     r6 = __ldg(r + 6), r7 = __ldg(r + 7), r8 = __ldg(r + 8);
        const double t0 = __ldg(trans1 + 3 * f), t1 = __ldg(trans1 + 3 * f + 1), t2 = __ldg(trans1 + 3 * f + 2);
#pragma unroll
        for (int c = 0; c < kTfCols; c++) {
            const int j = c0 + c * kTfThreads;
            if (j >= np2) continue;
            double2 qx, qy, qz;
            {
                const double a0 = x[c].x + t0, a1 = y[c].x + t1, a2 = z[c].x + t2;
                qx.x = fma(a2, r6, fma(a1, r3, a0 * r0)) + b0;
                qy.x = fma(a2, r7, fma(a1, r4, a0 * r1)) + b1;
                qz.x = fma(a2, r8, fma(a1, r5, a0 * r2)) + b2;
            }
            {
                const double a0 = x[c].y + t0, a1 = y[c].y + t1, a2 = z[c].y + t2;
                qx.y = fma(a2, r6, fma(a1, r3, a0 * r0)) + b0;
                qy.y = fma(a2, r7, fma(a1, r4, a0 * r1)) + b1;
                qz.y = fma(a2, r8, fma(a1, r5, a0 * r2)) + b2;
            }
            // padding atoms stay zero
            if (2 * j >= natoms) { qx.x = 0.0; qy.x = 0.0; qz.x = 0.0; }
            if (2 * j + 1 >= natoms) { qx.y = 0.0; qy.y = 0.0; qz.y = 0.0; }
            __stcs(X + j, qx); __stcs(Y + j, qy); __stcs(Z + j, qz);
        }
    }
}

#ifndef GCB_ACC_UNROLL
#define GCB_ACC_UNROLL 4   // frames in flight per thread in the column-owning sweeps
#endif
constexpr int kAccUnroll = GCB_ACC_UNROLL;
constexpr int kAccFrames = 32;   // per-frame transforms staged in shared memory at a time

// CTA (tile, split): double2 columns [tile*256, tile*256+256) of the frames [f0, f1) of its split; the two per-atom sums of a
// thread stay in registers over the whole range.  part[split][np].  bp = centred template planes (templa - trans2).
__global__ void __launch_bounds__(kTfThreads) peratom_accum_kernel(const double* __restrict__ frames, long frame_stride, int np, int natoms,
                                                                  long nframes, int nsplits, const double* __restrict__ rot,
                                                                  const double* __restrict__ trans1, const double* __restrict__ bp,
                                                                  double* __restrict__ part) {
    __shared__ double s_xf[kAccFrames][12];
    const int np2 = np >> 1;
    const int tile = blockIdx.x, split = blockIdx.y;
    const long per = (nframes + nsplits - 1) / nsplits;
    const long f0 = split * per;
    long f1 = f0 + per;
    if (f1 > nframes) f1 = nframes;
    const int j = tile * (int)blockDim.x + threadIdx.x;
    const bool in = j < np2;
    double2 bx = make_double2(0.0, 0.0), by = bx, bz = bx;
    if (in) {
        bx = reinterpret_cast<const double2*>(bp)[j];
        by = reinterpret_cast<const double2*>(bp + np)[j];
        bz = reinterpret_cast<const double2*>(bp + 2L * np)[j];
    }
    double s0 = 0.0, s1 = 0.0;
    for (long fb = f0; fb < f1; fb += kAccFrames) {
        const int nb = (int)((f1 - fb) < kAccFrames ? (f1 - fb) : kAccFrames);
        __syncthreads();   // the previous batch's transforms are no longer read
        for (int k = threadIdx.x; k < nb * 12; k += (int)blockDim.x) {
            const int ff = k / 12, e = k - ff * 12;
            // R, then centroid R (so that (a - c) R = a R - c R costs nothing per atom); the warp that stores element 9..11
            // recomputes it from R and trans1 = -c
            double v;
            if (e < 9) v = rot[9 * (fb + ff) + e];
            else {
                const double* r = rot + 9 * (fb + ff);
                const double* t = trans1 + 3 * (fb + ff);
                const int col = e - 9;
                v = fma(t[2], r[6 + col], fma(t[1], r[3 + col], t[0] * r[col]));   // = -(c R)[col]
            }
            s_xf[ff][e] = v;
        }
        __syncthreads();
        if (!in) continue;
        const double* fr = frames + fb * frame_stride;
#pragma unroll kAccUnroll
        for (int ff = 0; ff < nb; ff++, fr += frame_stride) {
            const double2 x = __ldcs(reinterpret_cast<const double2*>(fr) + j);
            const double2 y = __ldcs(reinterpret_cast<const double2*>(fr + np) + j);
            const double2 z = __ldcs(reinterpret_cast<const double2*>(fr + 2L * np) + j);
            const double* xf = s_xf[ff];
            const double r0 = xf[0], r1 = xf[1], r2 = xf[2], r3 = xf[3], r4 = xf[4], r5 = xf[5], r6 = xf[6], r7 = xf[7], r8 = xf[8];
            const double g0 = xf[9], g1 = xf[10], g2 = xf[11];
            {
                const double dx = fma(z.x, r6, fma(y.x, r3, fma(x.x, r0, g0))) - bx.x;
                const double dy = fma(z.x, r7, fma(y.x, r4, fma(x.x, r1, g1))) - by.x;
                const double dz = fma(z.x, r8, fma(y.x, r5, fma(x.x, r2, g2))) - bz.x;
                s0 += ptx::sqrt_cubic(fma(dz, dz, fma(dy, dy, dx * dx)));
            }
            {
                const double dx = fma(z.y, r6, fma(y.y, r3, fma(x.y, r0, g0))) - bx.y;
                const double dy = fma(z.y, r7, fma(y.y, r4, fma(x.y, r1, g1))) - by.y;
                const double dz = fma(z.y, r8, fma(y.y, r5, fma(x.y, r2, g2))) - bz.y;
                s1 += ptx::sqrt_cubic(fma(dz, dz, fma(dy, dy, dx * dx)));
            }
        }
    }
    if (in) {
        if (2 * j >= natoms) s0 = 0.0;
        if (2 * j + 1 >= natoms) s1 = 0.0;
        reinterpret_cast<double2*>(part + (long)split * np)[j] = make_double2(s0, s1);
    }
}

// Write-back with the accumulation kernel's shape, for frames of a few hundred atoms (the tile kernel above would leave most
// of its threads without a column): a thread owns one double2 column and walks the frames of its range.
__global__ void __launch_bounds__(kTfThreads) transform_cols_kernel(double* __restrict__ frames, long frame_stride, int np, int natoms,
                                                                   long nframes, int nsplits, const double* __restrict__ rot,
                                                                   const double* __restrict__ trans1, double b0, double b1, double b2) {
    __shared__ double s_xf[kAccFrames][12];
    const int np2 = np >> 1;
    const int tile = blockIdx.x, split = blockIdx.y;
    const long per = (nframes + nsplits - 1) / nsplits;
    const long f0 = split * per;
    long f1 = f0 + per;
    if (f1 > nframes) f1 = nframes;
    const int j = tile * (int)blockDim.x + threadIdx.x;
    const bool in = j < np2;
    const bool v0 = 2 * j < natoms, v1 = 2 * j + 1 < natoms;   // padding atoms stay zero
    for (long fb = f0; fb < f1; fb += kAccFrames) {
        const int nb = (int)((f1 - fb) < kAccFrames ? (f1 - fb) : kAccFrames);
        __syncthreads();
        for (int k = threadIdx.x; k < nb * 12; k += (int)blockDim.x) {
            const int ff = k / 12, e = k - ff * 12;
            double v;
            if (e < 9) v = rot[9 * (fb + ff) + e];
            else {
                const double* r = rot + 9 * (fb + ff);
                const double* t = trans1 + 3 * (fb + ff);
                const int col = e - 9;
                v = fma(t[2], r[6 + col], fma(t[1], r[3 + col], t[0] * r[col]));   // (trans1 R)[col] = -(centroid R)[col]
            }
            s_xf[ff][e] = v;
        }
        __syncthreads();
        if (!in) continue;
        double* fr = frames + fb * frame_stride;
#pragma unroll kAccUnroll
        for (int ff = 0; ff < nb; ff++, fr += frame_stride) {
            double2* X = reinterpret_cast<double2*>(fr) + j;
            double2* Y = reinterpret_cast<double2*>(fr + np) + j;
            double2* Z = reinterpret_cast<double2*>(fr + 2L * np) + j;
            const double2 x = __ldcs(X), y = __ldcs(Y), z = __ldcs(Z);
            const double* xf = s_xf[ff];
            const double g0 = xf[9] + b0, g1 = xf[10] + b1, g2 = xf[11] + b2;
            double2 qx, qy, qz;
            // (a + trans1) R + trans2 = a R + (trans1 R + trans2)
            qx.x = fma(z.x, xf[6], fma(y.x, xf[3], x.x * xf[0])) + g0;
            qy.x = fma(z.x, xf[7], fma(y.x, xf[4], x.x * xf[1])) + g1;
            qz.x = fma(z.x, xf[8], fma(y.x, xf[5], x.x * xf[2])) + g2;
            qx.y = fma(z.y, xf[6], fma(y.y, xf[3], x.y * xf[0])) + g0;
            qy.y = fma(z.y, xf[7], fma(y.y, xf[4], x.y * xf[1])) + g1;
            qz.y = fma(z.y, xf[8], fma(y.y, xf[5], x.y * xf[2])) + g2;
            if (!v0) { qx.x = 0.0; qy.x = 0.0; qz.x = 0.0; }
            if (!v1) { qx.y = 0.0; qy.y = 0.0; qz.y = 0.0; }
            __stcs(X, qx); __stcs(Y, qy); __stcs(Z, qz);
        }
    }
}

// The same sums for a chosen list of atoms only (align.LOVO reads the means of its candidate atoms only, rmsdFilter,
// align/lovo.go:519-528): thread k of a CTA owns atom want[k], reads its three sectors per frame and nothing else.
// bw = compact centred template of those atoms (3 x nwant); part[split][stride].
__global__ void __launch_bounds__(kTfThreads) peratom_accum_gather_kernel(const double* __restrict__ frames, long frame_stride, int np,
                                                                         long nframes, int nsplits, const double* __restrict__ rot,
                                                                         const double* __restrict__ trans1, const int* __restrict__ want,
                                                                         int nwant, const double* __restrict__ bw,
                                                                         double* __restrict__ part, int stride) {
    __shared__ double s_xf[kAccFrames][12];
    const int split = blockIdx.y;
    const long per = (nframes + nsplits - 1) / nsplits;
    const long f0 = split * per;
    long f1 = f0 + per;
    if (f1 > nframes) f1 = nframes;
    const int k = blockIdx.x * kTfThreads + threadIdx.x;
    const bool in = k < nwant;
    const int at = in ? want[k] : 0;
    const double bx = in ? bw[k] : 0.0, by = in ? bw[nwant + k] : 0.0, bz = in ? bw[2L * nwant + k] : 0.0;
    double s0 = 0.0, s1 = 0.0;   // two chains: even and odd frames of a batch
    for (long fb = f0; fb < f1; fb += kAccFrames) {
        const int nb = (int)((f1 - fb) < kAccFrames ? (f1 - fb) : kAccFrames);
        __syncthreads();
        for (int e12 = threadIdx.x; e12 < nb * 12; e12 += kTfThreads) {
            const int ff = e12 / 12, e = e12 - ff * 12;
            double v;
            if (e < 9) v = rot[9 * (fb + ff) + e];
            else {
                const double* r = rot + 9 * (fb + ff);
                const double* t = trans1 + 3 * (fb + ff);
                const int col = e - 9;
                v = fma(t[2], r[6 + col], fma(t[1], r[3 + col], t[0] * r[col]));
            }
            s_xf[ff][e] = v;
        }
        __syncthreads();
        if (!in) continue;
        const double* fr = frames + fb * frame_stride + at;
#pragma unroll kAccUnroll
        for (int ff = 0; ff < nb; ff++, fr += frame_stride) {
            const double x = __ldg(fr), y = __ldg(fr + np), z = __ldg(fr + 2L * np);
            const double* xf = s_xf[ff];
            const double dx = fma(z, xf[6], fma(y, xf[3], fma(x, xf[0], xf[9]))) - bx;
            const double dy = fma(z, xf[7], fma(y, xf[4], fma(x, xf[1], xf[10]))) - by;
            const double dz = fma(z, xf[8], fma(y, xf[5], fma(x, xf[2], xf[11]))) - bz;
            const double d = ptx::sqrt_cubic(fma(dz, dz, fma(dy, dy, dx * dx)));
            if (ff & 1) s1 += d; else s0 += d;
        }
    }
    if (in) part[(long)split * stride + k] = s0 + s1;
}

}  // namespace

int peratom_accum_gather_rows(const gcb_traj* t, long nframes, int nwant) {
    const int tiles = (nwant + kTfThreads - 1) / kTfThreads;
    long want = ((long)t->sm_count * 8 + tiles - 1) / tiles;
    if (want > 512) want = 512;
    const long by_batch = (nframes + kAccFrames - 1) / kAccFrames;   // a split of less than one batch of transforms is wasted staging
    if (want > by_batch) want = by_batch;
    if (want < 1) want = 1;
    return (int)want;
}

int launch_peratom_accum_gather(gcb_traj* t, const SuperParams& p, const int* d_want, int nwant, const double* d_wantref, int stride,
                                int* rows_out) {
    const int tiles = (nwant + kTfThreads - 1) / kTfThreads;
    const int splits = peratom_accum_gather_rows(t, p.nframes, nwant);
    peratom_accum_gather_kernel<<<dim3((unsigned)tiles, (unsigned)splits), kTfThreads, 0, t->stream>>>(p.frames, p.frame_stride, p.np, p.nframes, splits,
                                                                                                       p.rot, p.trans1, d_want, nwant, d_wantref,
                                                                                                       p.devpart, stride);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    if (rows_out) *rows_out = splits;
    return GCB_OK;
}

// threads per CTA and frame splits of the column-owning kernels: whole warps covering the frame's double2 columns (at most 256),
// about 8 CTAs of 256 threads per SM in all, never less than one batch of transforms per split
static void cols_shape(const gcb_traj* t, long nframes, int* threads, int* tiles, int* splits) {
    const int np2 = t->np >> 1;
    *threads = np2 >= kTfThreads ? kTfThreads : round_up(np2, 32);
    *tiles = (np2 + *threads - 1) / *threads;
    long want = ((long)t->sm_count * 8 * kTfThreads / *threads + *tiles - 1) / *tiles;
    const long cap = t->np >= 4096 ? 64 : 1024;
    if (want > cap) want = cap;
    const long by_batch = (nframes + kAccFrames - 1) / kAccFrames;
    if (want > by_batch) want = by_batch;
    if (want < 1) want = 1;
    *splits = (int)want;
}

int launch_transform_frames(gcb_traj* t, const SuperParams& p) {
    const int np2 = t->np >> 1;
    if (np2 < kTfTile) {   // small frames: a thread per column
        int threads, tiles, splits;
        cols_shape(t, p.nframes, &threads, &tiles, &splits);
        transform_cols_kernel<<<dim3((unsigned)tiles, (unsigned)splits), threads, 0, t->stream>>>(p.frames_rw, p.frame_stride, p.np, p.natoms, p.nframes,
                                                                                                 splits, p.rot, p.trans1, p.bbar[0], p.bbar[1], p.bbar[2]);
        GCB_LAUNCHED();
        GCB_CUDA(cudaGetLastError());
        return GCB_OK;
    }
    const long tiles = p.nframes * ((np2 + kTfTile - 1) / kTfTile);
    long grid = (long)t->sm_count * 8;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    transform_frames_kernel<<<(unsigned)grid, kTfThreads, 0, t->stream>>>(p.frames_rw, p.frame_stride, p.np, p.natoms, p.nframes, p.rot,
                                                                          p.trans1, p.bbar[0], p.bbar[1], p.bbar[2]);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

// rows of per-atom partial sums the accumulation writes for nframes frames (one per frame split)
int peratom_accum_rows(const gcb_traj* t, long nframes) {
    int threads, tiles, splits;
    cols_shape(t, nframes, &threads, &tiles, &splits);
    return splits;
}

int launch_peratom_accum(gcb_traj* t, const SuperParams& p, const double* ref_planes, int* rows_out) {
    int threads, tiles, splits;
    cols_shape(t, p.nframes, &threads, &tiles, &splits);
    peratom_accum_kernel<<<dim3((unsigned)tiles, (unsigned)splits), threads, 0, t->stream>>>(p.frames, p.frame_stride, p.np, p.natoms, p.nframes,
                                                                                            splits, p.rot, p.trans1, ref_planes, p.devpart);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    if (rows_out) *rows_out = splits;
    return GCB_OK;
}

}  // namespace gcb
