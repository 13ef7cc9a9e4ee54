// super_small.cu -- the fused superposition + RMSD pass for small structures (at most 512 atoms per frame).
//
// Same arithmetic as super_cluster.cu (chem.Super -> RotatorTranslatorToSuper -> MemRMSD / MemPerAtomRMSD,
// /root/reference/handy.go:175-214, geometric.go:127-208, :258-321).  A frame of a few hundred atoms is a few kilobytes:
// with one frame per CTA at a time the per-frame costs that do not shrink with the frame -- the cross-warp reduction,
// the hand-off to the solver, a 3x3 solve on one lane -- dominate (200 M frames/s whatever the size below ~500 atoms).
// Here every WARP owns frames:
//
//  * a warp takes batches of 32 consecutive frames.  Sweep 1 walks the batch frame by frame: lane l holds atoms
//    l, l+32, ... (K per lane, coalesced streaming loads, two frames in flight), the centred template comes from
//    shared memory, and a shuffle butterfly leaves the frame's 13 sums in shared memory;
//  * then the 32 lanes solve the 32 frames of the batch AT ONCE -- lane l runs the 3x3 solve of frame l
//    (kabsch3.cuh), decides between the one-pass RMSD and the explicit residual exactly as the cluster kernel does, and
//    writes R, the translations and the RMSD.  No lane idles through somebody else's solve;
//  * sweep 2, only for the frames that need it (write-back, per-atom distances, explicit residual): the frames are
//    read again (they were read microseconds ago: L2), rotated, written back / accumulated.
//
// Deterministic: every sum has a fixed order.  HBM-bound above ~300 atoms, instruction-bound below.
#include "common.cuh"
#include "kabsch3.cuh"
#include "ptx.cuh"

namespace gcb {

namespace {
using ptx::warp_reduce12;

constexpr int kSmallWarps = 8;
constexpr int kSmallThreads = 32 * kSmallWarps;
constexpr int kSmallMaxAtoms = 512;
constexpr int kMinRun = 4;   // no more warps than frames / 4: a warp's fixed costs (template, solve) want a few frames to spread over

__device__ __forceinline__ double shx2(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// CTAs per SM the instances are built for (measured, tools/small_sweep.py): two (128 registers) for K <= 4 and K = 16; the K = 8
// instance keeps two frames of 8 atoms per lane in registers and is faster unbounded (183 registers) than squeezed into 128.
#ifndef GCB_SMALL_MINB
#define GCB_SMALL_MINB 2
#endif
#ifndef GCB_MID_MINB
#define GCB_MID_MINB 2
#endif
// One 13-double record per frame of the batch, used twice: first the frame's sums -- S (3), H (9), sum w|a|^2 -- written by sweep 1,
// then, in place, what the lane that solved the frame leaves for sweep 2 -- R (9), centroid*R (3), [12] != 0: form the explicit
// residual.  (Lane l reads record l completely before it writes it.)
struct SmallShared {
    double rec[kSmallWarps][32][13];
};

// sqrt with a short dependency chain (see super_cluster.cu: cubic-order correction of the rsqrt seed)
__device__ __forceinline__ double small_sqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double g = x * y;
    const double e = fma(-g, y, 1.0);
    const double out = fma(g, e * fma(e, 0.375, 0.5), g);
    return (__double2hiint(x) < 0x00200000) ? 0.0 : out;
}

// The 3x3 problem of one frame from its 13 sums (one lane per frame): rotation, centroid, the one-pass RMSD when its
// rounding bound allows (same rule as super_cluster.cu) and the per-frame outputs.  Returns true when sweep 2 must form the
// explicit residual for this frame.
__device__ __forceinline__ bool solve_lane(const SuperParams& p, const double* sums, double* xf, long f, bool sweep2_all) {
    double s[13];   // xf may alias sums: take a copy first
#pragma unroll
    for (int k = 0; k < 13; k++) s[k] = sums[k];
    const double cen[3] = {s[0] * p.inv_n, s[1] * p.inv_n, s[2] * p.inv_n};
    double h[9], r[9];
    bool finite = true;
#pragma unroll
    for (int k = 0; k < 9; k++) { h[k] = s[3 + k]; finite = finite && (fabs(h[k]) < 1.79e308); }
    int rc = finite ? k3::kabsch3(h, r) : -4;
    if (rc != 0) {
        if (p.status) atomicExch(p.status, (int)GCB_ERR_SVD);
#pragma unroll
        for (int k = 0; k < 9; k++) r[k] = (k % 4 == 0) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) xf[k] = r[k];
    xf[9] = fma(cen[2], r[6], fma(cen[1], r[3], cen[0] * r[0]));
    xf[10] = fma(cen[2], r[7], fma(cen[1], r[4], cen[0] * r[1]));
    xf[11] = fma(cen[2], r[8], fma(cen[1], r[5], cen[0] * r[2]));
    const double ga = s[12] - (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) * p.inv_n;
    const double trrh = fma(r[8], h[8], fma(r[7], h[7], fma(r[6], h[6], fma(r[5], h[5], fma(r[4], h[4],
                        fma(r[3], h[3], fma(r[2], h[2], fma(r[1], h[1], r[0] * h[0]))))))));
    const double e = ga + p.gb - 2.0 * trrh;
    const double smax = s[12] + p.gb;
    const bool formula_ok = !sweep2_all && rc == 0 && e > 0.0 &&
                            (64.0 * 2.220446049250313e-16) * smax <= 2e-10 * sqrt(p.n_fit * e);
    xf[12] = formula_ok ? 0.0 : 1.0;
    if (p.need_explicit) p.need_explicit[f] = formula_ok ? 0 : 1;
    if (formula_ok && p.rmsd) p.rmsd[f] = sqrt(e * p.inv_n);
    if (p.rot) {
#pragma unroll
        for (int k = 0; k < 9; k++) p.rot[f * 9 + k] = r[k];
    }
    if (p.trans1) { p.trans1[f * 3] = -cen[0]; p.trans1[f * 3 + 1] = -cen[1]; p.trans1[f * 3 + 2] = -cen[2]; }
    if (p.trans2) { p.trans2[f * 3] = p.bbar[0]; p.trans2[f * 3 + 1] = p.bbar[1]; p.trans2[f * 3 + 2] = p.bbar[2]; }
    return !formula_ok;
}

// GATHER: the fit uses nidx atoms picked out of a (possibly much larger) frame -- slot q of the warp is atom idx[q] and the
// template is the compact list refpairs (3 x nidx).  Only those atoms are read: 3 sectors per atom instead of the whole frame.
template <int K, bool MASKED, bool DEV, bool GATHER = false>
__global__ void __launch_bounds__(kSmallThreads, (K == 8 ? 1 : GCB_SMALL_MINB)) super_small_kernel(const SuperParams p, int always_explicit,
                                                                  const double* __restrict__ refpairs = nullptr,
                                                                  const int* __restrict__ idx = nullptr, int nidx = 0) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    SmallShared& sh = *reinterpret_cast<SmallShared*>(small_smem);
    double* tb = reinterpret_cast<double*>(small_smem + sizeof(SmallShared));   // template planes bx | by | bz, 32*K each
    constexpr int NP = 32 * K;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nslots = GATHER ? nidx : p.np;       // slots that hold data
    const int nreal = GATHER ? nidx : p.natoms;    // slots that hold real atoms
    for (int i = tid; i < 3 * NP; i += kSmallThreads) {
        const int c = i / NP, a = i - c * NP;
        tb[i] = a < nslots ? (GATHER ? refpairs[(long)c * nidx + a] : p.refc[(long)c * p.np + a]) : 0.0;
    }
    uint32_t real_mask = 0, fit_mask = 0;
    int at[K];                                      // where slot lane + 32 k sits in a coordinate plane
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int a = lane + 32 * k;
        at[k] = a < nslots ? (GATHER ? idx[a] : a) : 0;
        if (a < nreal) {
            real_mask |= 1u << k;
            if (!MASKED || p.weight[a] != 0.0) fit_mask |= 1u << k;
        }
    }
    __syncthreads();
    const double* bx = tb + lane;
    const double* by = bx + NP;
    const double* bz = by + NP;
    double dev[DEV ? K : 1];
#pragma unroll
    for (int k = 0; k < (DEV ? K : 1); k++) dev[k] = 0.0;
    const bool wb = p.frames_rw != nullptr;
    const bool sweep2_all = always_explicit || wb || DEV;
    const long gwarp = (long)blockIdx.x * kSmallWarps + warp, nwarps = (long)gridDim.x * kSmallWarps;

    // Every warp owns one contiguous run of frames, all runs of (almost) equal length, and walks it in batches of up to 32: the
    // warps finish together.  (Dealing whole batches of 32 left the last round partly empty: 100 000 frames over 2368 warps are
    // 1.3 batches per warp -- two rounds, the second a third full.)
    const long run = p.nframes / nwarps, extra = p.nframes % nwarps;
    const long lo = gwarp * run + (gwarp < extra ? gwarp : extra), hi = lo + run + (gwarp < extra ? 1 : 0);
    for (long f0 = lo; f0 < hi; f0 += 32) {
        const int nf = (int)((hi - f0 < 32) ? (hi - f0) : 32);
        // ---------------- sweep 1: the 13 sums of every frame of the batch ----------------
        double x[K], y[K], z[K];
        auto load = [&](int i, auto& lx, auto& ly, auto& lz) {
            const double* fr = p.frames + (f0 + i) * p.frame_stride;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const bool in = lane + 32 * k < nslots;   // planes are padded to 16: the padding atoms read as zeros
                lx[k] = in ? __ldcs(fr + at[k]) : 0.0;
                ly[k] = in ? __ldcs(fr + p.np + at[k]) : 0.0;
                lz[k] = in ? __ldcs(fr + 2L * p.np + at[k]) : 0.0;
            }
        };
        constexpr bool PF = K <= 8;   // two frames in flight while the registers allow it
        constexpr int KN = PF ? K : 1;
        if (PF) load(0, x, y, z);
        for (int i = 0; i < nf; i++) {
            double nx[KN], ny[KN], nz[KN];
            if constexpr (PF) {
                if (i + 1 < nf) load(i + 1, nx, ny, nz);   // the next frame is in flight while this one is reduced
            } else {
                load(i, x, y, z);
            }
            double acc[12], aa = 0.0;
#pragma unroll
            for (int q = 0; q < 12; q++) acc[q] = 0.0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                double ax = x[k], ay = y[k], az = z[k];
                if (MASKED) {
                    const bool on = (fit_mask >> k) & 1u;
                    ax = on ? ax : 0.0; ay = on ? ay : 0.0; az = on ? az : 0.0;
                }
                const double tx = bx[32 * k], ty = by[32 * k], tz = bz[32 * k];
                aa = fma(az, az, fma(ay, ay, fma(ax, ax, aa)));
                acc[0] += ax; acc[1] += ay; acc[2] += az;
                acc[3] = fma(ax, tx, acc[3]); acc[4] = fma(ax, ty, acc[4]); acc[5] = fma(ax, tz, acc[5]);
                acc[6] = fma(ay, tx, acc[6]); acc[7] = fma(ay, ty, acc[7]); acc[8] = fma(ay, tz, acc[8]);
                acc[9] = fma(az, tx, acc[9]); acc[10] = fma(az, ty, acc[10]); acc[11] = fma(az, tz, acc[11]);
            }
            // halving butterfly (15 shuffles for the 12 sums, 5 for sum |a|^2), fixed order
            double t0, t1;
            warp_reduce12(acc, lane, t0, t1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) aa += shx2(aa, o);
            double* sm = sh.rec[warp][i];
            if (lane == 0) sm[12] = aa;
            if ((lane & 3) == 0) {
                const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1;
                sm[6 * b4 + 3 * b3 + 2 * b2] = t0;
                if (b2 == 0) sm[6 * b4 + 3 * b3 + 1] = t1;
            }
            if constexpr (PF) {
                if (i + 1 < nf) {
#pragma unroll
                    for (int k = 0; k < K; k++) { x[k] = nx[k]; y[k] = ny[k]; z[k] = nz[k]; }
                }
            }
        }
        __syncwarp();
        // ---------------- the 32 solves of the batch, one per lane ----------------
        bool any_sweep2 = sweep2_all;
        if (lane < nf) any_sweep2 = solve_lane(p, sh.rec[warp][lane], sh.rec[warp][lane], f0 + lane, sweep2_all) || any_sweep2;
        any_sweep2 = __any_sync(0xffffffffu, any_sweep2);
        __syncwarp();
        if (!any_sweep2) continue;
        // ---------------- sweep 2: rotate, explicit residual, per-atom distances, write-back ----------------
        // (the frames were read microseconds ago: L2.  When every frame takes part the next one is in flight meanwhile.)
        const bool pf2 = PF && sweep2_all;
        if (pf2) load(0, x, y, z);
        for (int i = 0; i < nf; i++) {
            const double* xf = sh.rec[warp][i];
            if (!sweep2_all && xf[12] == 0.0) continue;   // warp-uniform
            double nx[KN], ny[KN], nz[KN];
            if constexpr (PF) {
                if (pf2) { if (i + 1 < nf) load(i + 1, nx, ny, nz); }
                else load(i, x, y, z);
            } else {
                load(i, x, y, z);
            }
            const double r0 = xf[0], r1 = xf[1], r2 = xf[2], r3 = xf[3], r4 = xf[4], r5 = xf[5], r6 = xf[6], r7 = xf[7], r8 = xf[8];
            const double g0 = -xf[9], g1 = -xf[10], g2 = -xf[11];
            const long fo = (f0 + i) * p.frame_stride;
            double e = 0.0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const double ax = x[k], ay = y[k], az = z[k];
                const double qx = fma(az, r6, fma(ay, r3, fma(ax, r0, g0)));
                const double qy = fma(az, r7, fma(ay, r4, fma(ax, r1, g1)));
                const double qz = fma(az, r8, fma(ay, r5, fma(ax, r2, g2)));
                const double dx = qx - bx[32 * k], dy = qy - by[32 * k], dz = qz - bz[32 * k];
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                if ((fit_mask >> k) & 1u) e += d2;
                if (DEV) { if (lane + 32 * k < nslots) dev[k] += small_sqrt(d2); }
                if (wb && ((real_mask >> k) & 1u)) {
                    p.frames_rw[fo + at[k]] = qx + p.bbar[0];
                    p.frames_rw[fo + p.np + at[k]] = qy + p.bbar[1];
                    p.frames_rw[fo + 2L * p.np + at[k]] = qz + p.bbar[2];
                }
            }
            if (p.rmsd) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) e += shx2(e, o);
                if (lane == 0) p.rmsd[f0 + i] = sqrt(e * p.inv_n);
            }
            if constexpr (PF) {
                if (pf2 && i + 1 < nf) {
#pragma unroll
                    for (int k = 0; k < K; k++) { x[k] = nx[k]; y[k] = ny[k]; z[k] = nz[k]; }
                }
            }
        }
        __syncwarp();
    }
    if (DEV) {
        double* row = p.devpart + gwarp * (long)p.np;
#pragma unroll
        for (int k = 0; k < K; k++)
            if (lane + 32 * k < p.np) row[lane + 32 * k] = dev[k];
    }
}


// ---- 512 < atoms <= 2048: the same scheme with the frame walked in chunks of 128 atoms (4 per lane), the next chunk in
// flight while the current one is accumulated.  Fit, subset fit, explicit residual and write-back; per-atom distance sums of
// structures this size stay with the cluster kernel.
constexpr int kMidChunk = 4;                 // atoms per lane per chunk
constexpr int kMidMaxAtoms = 2048;
constexpr int kGatherMaxAtoms = 4096;        // gathered subsets: template planes of 96 KB in shared memory

template <bool MASKED, bool GATHER = false>
__global__ void __launch_bounds__(kSmallThreads, GATHER ? 1 : GCB_MID_MINB) super_mid_kernel(const SuperParams p, int always_explicit, int nch,
                                                                   const double* __restrict__ refpairs = nullptr,
                                                                   const int* __restrict__ idx = nullptr, int nidx = 0) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    SmallShared& sh = *reinterpret_cast<SmallShared*>(small_smem);
    double* tb = reinterpret_cast<double*>(small_smem + sizeof(SmallShared));   // template planes, NP each
    constexpr int CH = kMidChunk;
    const int NP = 32 * CH * nch;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nslots = GATHER ? nidx : p.np, nreal = GATHER ? nidx : p.natoms;
    for (int i = tid; i < 3 * NP; i += kSmallThreads) {
        const int c = i / NP, a = i - c * NP;
        tb[i] = a < nslots ? (GATHER ? refpairs[(long)c * nidx + a] : p.refc[(long)c * p.np + a]) : 0.0;
    }
    // streaming form: bit c*CH + k of the masks = atom lane + 32*(c*CH + k) (at most 16 chunks); the gather form has no subset
    // inside its list, a slot is real when it is below nidx
    unsigned long long real_mask = 0, fit_mask = 0;
    if (!GATHER) {
        for (int q = 0; q < CH * nch; q++) {
            const int a = lane + 32 * q;
            if (a < p.natoms) {
                real_mask |= 1ull << q;
                if (!MASKED || p.weight[a] != 0.0) fit_mask |= 1ull << q;
            }
        }
    }
    __syncthreads();
    const double* bx = tb + lane;
    const double* by = bx + NP;
    const double* bz = by + NP;
    const bool wb = p.frames_rw != nullptr;
    const bool sweep2_all = always_explicit || wb;
    const long gwarp = (long)blockIdx.x * kSmallWarps + warp, nwarps = (long)gridDim.x * kSmallWarps;
    const long run = p.nframes / nwarps, extra = p.nframes % nwarps;   // contiguous runs of equal length, see super_small_kernel
    const long lo = gwarp * run + (gwarp < extra ? gwarp : extra), hi = lo + run + (gwarp < extra ? 1 : 0);
    for (long f0 = lo; f0 < hi; f0 += 32) {
        const int nf = (int)((hi - f0 < 32) ? (hi - f0) : 32);
        double x[CH], y[CH], z[CH];
        auto load = [&](int i, int c, double (&lx)[CH], double (&ly)[CH], double (&lz)[CH]) {
            const double* fr = p.frames + (f0 + i) * p.frame_stride;
#pragma unroll
            for (int k = 0; k < CH; k++) {
                const int slot = lane + 32 * (CH * c + k);
                const bool in = slot < nslots;
                const int a = GATHER ? (in ? __ldg(idx + slot) : 0) : slot;
                lx[k] = in ? __ldcs(fr + a) : 0.0;
                ly[k] = in ? __ldcs(fr + p.np + a) : 0.0;
                lz[k] = in ? __ldcs(fr + 2L * p.np + a) : 0.0;
            }
        };
        load(0, 0, x, y, z);
        for (int i = 0; i < nf; i++) {
            double acc[12], aa = 0.0;
#pragma unroll
            for (int q = 0; q < 12; q++) acc[q] = 0.0;
            for (int c = 0; c < nch; c++) {
                double nx[CH], ny[CH], nz[CH];
                const bool last = c + 1 == nch;
                if (!last) load(i, c + 1, nx, ny, nz);
                else if (i + 1 < nf) load(i + 1, 0, nx, ny, nz);
                const uint32_t fm = GATHER ? 0u : (uint32_t)(fit_mask >> (CH * c));
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    double ax = x[k], ay = y[k], az = z[k];
                    if (MASKED) {
                        const bool on = (fm >> k) & 1u;
                        ax = on ? ax : 0.0; ay = on ? ay : 0.0; az = on ? az : 0.0;
                    }
                    const int o = 32 * (CH * c + k);
                    const double tx = bx[o], ty = by[o], tz = bz[o];
                    aa = fma(az, az, fma(ay, ay, fma(ax, ax, aa)));
                    acc[0] += ax; acc[1] += ay; acc[2] += az;
                    acc[3] = fma(ax, tx, acc[3]); acc[4] = fma(ax, ty, acc[4]); acc[5] = fma(ax, tz, acc[5]);
                    acc[6] = fma(ay, tx, acc[6]); acc[7] = fma(ay, ty, acc[7]); acc[8] = fma(ay, tz, acc[8]);
                    acc[9] = fma(az, tx, acc[9]); acc[10] = fma(az, ty, acc[10]); acc[11] = fma(az, tz, acc[11]);
                }
                if (!last || i + 1 < nf) {
#pragma unroll
                    for (int k = 0; k < CH; k++) { x[k] = nx[k]; y[k] = ny[k]; z[k] = nz[k]; }
                }
            }
            double t0, t1;
            warp_reduce12(acc, lane, t0, t1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) aa += shx2(aa, o);
            double* sm = sh.rec[warp][i];
            if (lane == 0) sm[12] = aa;
            if ((lane & 3) == 0) {
                const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1;
                sm[6 * b4 + 3 * b3 + 2 * b2] = t0;
                if (b2 == 0) sm[6 * b4 + 3 * b3 + 1] = t1;
            }
        }
        __syncwarp();
        bool any_sweep2 = sweep2_all;
        if (lane < nf) any_sweep2 = solve_lane(p, sh.rec[warp][lane], sh.rec[warp][lane], f0 + lane, sweep2_all) || any_sweep2;
        any_sweep2 = __any_sync(0xffffffffu, any_sweep2);
        __syncwarp();
        if (!any_sweep2) continue;
        for (int i = 0; i < nf; i++) {
            const double* xf = sh.rec[warp][i];
            if (!sweep2_all && xf[12] == 0.0) continue;   // warp-uniform
            const double r0 = xf[0], r1 = xf[1], r2 = xf[2], r3 = xf[3], r4 = xf[4], r5 = xf[5], r6 = xf[6], r7 = xf[7], r8 = xf[8];
            const double g0 = -xf[9], g1 = -xf[10], g2 = -xf[11];
            const long fo = (f0 + i) * p.frame_stride + lane;
            double e = 0.0;
            for (int c = 0; c < nch; c++) {
                uint32_t fm = (uint32_t)(fit_mask >> (CH * c)), rm = (uint32_t)(real_mask >> (CH * c));
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    const int o = 32 * (CH * c + k);
                    if (lane + o >= nslots) continue;
                    if (GATHER) { fm |= 1u << k; rm |= 1u << k; }   // every listed atom is a real atom of the fit
                    const long at = GATHER ? (long)__ldg(idx + lane + o) - lane : (long)o;   // offset from fo
                    const double ax = p.frames[fo + at], ay = p.frames[fo + p.np + at], az = p.frames[fo + 2L * p.np + at];
                    const double qx = fma(az, r6, fma(ay, r3, fma(ax, r0, g0)));
                    const double qy = fma(az, r7, fma(ay, r4, fma(ax, r1, g1)));
                    const double qz = fma(az, r8, fma(ay, r5, fma(ax, r2, g2)));
                    const double dx = qx - bx[o], dy = qy - by[o], dz = qz - bz[o];
                    if ((fm >> k) & 1u) e += fma(dz, dz, fma(dy, dy, dx * dx));
                    if (wb && ((rm >> k) & 1u)) {
                        p.frames_rw[fo + at] = qx + p.bbar[0];
                        p.frames_rw[fo + p.np + at] = qy + p.bbar[1];
                        p.frames_rw[fo + 2L * p.np + at] = qz + p.bbar[2];
                    }
                }
            }
            if (p.rmsd) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) e += shx2(e, o);
                if (lane == 0) p.rmsd[f0 + i] = sqrt(e * p.inv_n);
            }
        }
        __syncwarp();
    }
}

template <int K>
int launch_small_k(gcb_traj* t, const SuperParams& p, bool masked, bool want_dev, int always_explicit, int& grid, size_t smem) {
#define GCB_SMALL_LAUNCH(M, D)                                                                                   \
    do {                                                                                                         \
        GCB_CUDA(opt_in_smem(super_small_kernel<K, M, D>, t->dev)); \
        int per_sm = 1;                                                                                          \
        GCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, super_small_kernel<K, M, D>, kSmallThreads, smem)); \
        if (per_sm < 1) per_sm = 1;                                                                              \
        if (per_sm > 4) per_sm = 4;                                                                              \
        if (grid > t->sm_count * per_sm) grid = t->sm_count * per_sm;                                            \
        super_small_kernel<K, M, D><<<grid, kSmallThreads, smem, t->stream>>>(p, always_explicit);        \
    } while (0)
    if (masked) { if (want_dev) GCB_SMALL_LAUNCH(true, true); else GCB_SMALL_LAUNCH(true, false); }
    else { if (want_dev) GCB_SMALL_LAUNCH(false, true); else GCB_SMALL_LAUNCH(false, false); }
#undef GCB_SMALL_LAUNCH
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

}  // namespace

// Measured on B200 (tools/small_sweep.py, profiles/r1_s3_small_sweep.jsonl): for fit-only passes the warp-per-frame kernels win
// up to ~1250 atoms (6.1 vs 3.8 TB/s at 1000 atoms, 1.37 G vs 0.19 G frames/s at 100 atoms); passes whose second sweep touches
// every frame (write-back, per-atom sums, explicit residual) only up to 256 atoms -- above that the cluster kernel's
// shared-memory-resident second sweep is faster.
// Fit-only pass over nidx <= 512 atoms gathered out of every frame (RMSD, rotation, translations; frames untouched).
int launch_super_small_gather(gcb_traj* t, const SuperParams& p, const double* refpairs, const int* idx, int nidx, int force_explicit) {
    if (nidx > kSmallMaxAtoms) {   // chunked form, up to kGatherMaxAtoms
        const int nch = (nidx + 32 * kMidChunk - 1) / (32 * kMidChunk);
        const size_t smem = sizeof(SmallShared) + sizeof(double) * 3 * 32 * kMidChunk * (size_t)nch;
        // one CTA per SM: with two, a subset of 3000 atoms of a 100 000-atom frame takes 6.6 ms per 20 000 frames instead of 2.8
        // (the gathered sectors live in what the shared-memory carve-out leaves of L1)
        long grid = (long)t->sm_count;
        const long need = (p.nframes + (long)kSmallWarps * kMinRun - 1) / ((long)kSmallWarps * kMinRun);
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        GCB_CUDA(opt_in_smem(super_mid_kernel<false, true>, t->dev));
        super_mid_kernel<false, true><<<(int)grid, kSmallThreads, smem, t->stream>>>(p, force_explicit ? 1 : 0, nch, refpairs, idx, nidx);
        GCB_LAUNCHED();
        GCB_CUDA(cudaGetLastError());
        return GCB_OK;
    }
    const int k_needed = (nidx + 31) / 32;
    const int K = k_needed <= 1 ? 1 : k_needed <= 2 ? 2 : k_needed <= 4 ? 4 : k_needed <= 8 ? 8 : 16;
    const size_t smem = sizeof(SmallShared) + sizeof(double) * 3 * 32 * (size_t)K;
    long grid = (long)t->sm_count * GCB_SMALL_MINB;
    const long need = (p.nframes + (long)kSmallWarps * kMinRun - 1) / ((long)kSmallWarps * kMinRun);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
#define GCB_GATHER_LAUNCH(KK)                                                                                                  \
    do {                                                                                                                       \
        GCB_CUDA(opt_in_smem(super_small_kernel<KK, false, false, true>, t->dev)); \
        super_small_kernel<KK, false, false, true><<<(int)grid, kSmallThreads, smem, t->stream>>>(p, force_explicit ? 1 : 0, refpairs, idx, nidx); \
    } while (0)
    switch (K) {
        case 1: GCB_GATHER_LAUNCH(1); break;
        case 2: GCB_GATHER_LAUNCH(2); break;
        case 4: GCB_GATHER_LAUNCH(4); break;
        case 8: GCB_GATHER_LAUNCH(8); break;
        default: GCB_GATHER_LAUNCH(16); break;
    }
#undef GCB_GATHER_LAUNCH
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int small_gather_max_atoms() { return kGatherMaxAtoms; }

// Passes whose second sweep touches every frame (write-back, per-atom sums): warp-per-frame up to 448 atoms, the cluster kernel above
// (tools/heavy_small_probe.py, profiles/r1_s5_heavy_small_threshold.jsonl: 300 atoms 3.49 -> 2.64 ms write-back, 3.27 -> 2.63 ms
// per-atom sums; 400 atoms 2.63 -> 2.51 / 2.47 -> 2.08; at 512 atoms the cluster kernel wins the write-back, 2.09 against 2.32 ms)
#ifndef GCB_SMALL_HEAVY_MAX
#define GCB_SMALL_HEAVY_MAX 448
#endif
// Fit-only passes: the chunked warp-per-frame kernel up to 1664 atoms (13 chunks), the cluster kernel above (100 000 frames,
// tools/mid_threshold_probe.py, profiles/r1_s5_mid_threshold.jsonl: 1400 atoms 634 -> 526 us, 1500 atoms 634 -> 573 us, a tie at
// 1700, the cluster kernel ahead at 2000: 715 against 776 us)
#ifndef GCB_MID_FIT_MAX
#define GCB_MID_FIT_MAX 1664
#endif
bool small_kernel_supported(const gcb_traj* t, bool heavy) { return t->np <= (heavy ? GCB_SMALL_HEAVY_MAX : GCB_MID_FIT_MAX); }

// rows of the per-atom distance partials this kernel writes (one per warp of the grid)
int small_kernel_rows(const gcb_traj* t) { return t->sm_count * 4 * kSmallWarps; }

int launch_super_small(gcb_traj* t, const SuperParams& p, bool masked, bool want_dev, int force_explicit, int* rows_out) {
    if (t->np > kSmallMaxAtoms) {   // chunked form
        const int nch = (t->np + 32 * kMidChunk - 1) / (32 * kMidChunk);
        const size_t smem = sizeof(SmallShared) + sizeof(double) * 3 * 32 * kMidChunk * (size_t)nch;
        long grid = (long)t->sm_count * GCB_MID_MINB;
        const long need = (p.nframes + (long)kSmallWarps * kMinRun - 1) / ((long)kSmallWarps * kMinRun);
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        const int always = (force_explicit || p.frames_rw != nullptr) ? 1 : 0;
        if (masked) {
            GCB_CUDA(opt_in_smem(super_mid_kernel<true>, t->dev));
            super_mid_kernel<true><<<(int)grid, kSmallThreads, smem, t->stream>>>(p, always, nch);
        } else {
            GCB_CUDA(opt_in_smem(super_mid_kernel<false>, t->dev));
            super_mid_kernel<false><<<(int)grid, kSmallThreads, smem, t->stream>>>(p, always, nch);
        }
        GCB_LAUNCHED();
        GCB_CUDA(cudaGetLastError());
        if (rows_out) *rows_out = 0;
        return GCB_OK;
    }
    const int k_needed = (t->np + 31) / 32;
    const int K = k_needed <= 1 ? 1 : k_needed <= 2 ? 2 : k_needed <= 4 ? 4 : k_needed <= 8 ? 8 : 16;
    const size_t smem = sizeof(SmallShared) + sizeof(double) * 3 * 32 * (size_t)K;
    long grid = (long)t->sm_count * 4;   // trimmed to what an SM can hold (occupancy query at launch)
    const long need = (p.nframes + (long)kSmallWarps * kMinRun - 1) / ((long)kSmallWarps * kMinRun);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    int g = (int)grid;
    const int always = (force_explicit || want_dev || p.frames_rw != nullptr) ? 1 : 0;
    if (want_dev) GCB_CUDA(cudaMemsetAsync(p.devpart, 0, sizeof(double) * (size_t)small_kernel_rows(t) * (size_t)t->np, t->stream));
    int rc;
    switch (K) {
        case 1: rc = launch_small_k<1>(t, p, masked, want_dev, always, g, smem); break;
        case 2: rc = launch_small_k<2>(t, p, masked, want_dev, always, g, smem); break;
        case 4: rc = launch_small_k<4>(t, p, masked, want_dev, always, g, smem); break;
        case 8: rc = launch_small_k<8>(t, p, masked, want_dev, always, g, smem); break;
        default: rc = launch_small_k<16>(t, p, masked, want_dev, always, g, smem); break;
    }
    if (rows_out) *rows_out = g * kSmallWarps;
    return rc;
}

}  // namespace gcb
