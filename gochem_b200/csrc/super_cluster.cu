// super_cluster.cu -- the B200 path of the fused superposition + RMSD pass.
//
// Same arithmetic as super_generic.cu (chem.Super -> RotatorTranslatorToSuper -> MemRMSD /
// MemPerAtomRMSD, /root/reference/handy.go:175-214, geometric.go:127-208, :258-321), organised so
// that every byte of a frame crosses HBM -> SM exactly once and nothing else is re-read per frame:
//
//  * A thread-block cluster of C CTAs owns one frame at a time; CTA rank r always owns atoms
//    [r*P, (r+1)*P).  The kernel is persistent: cluster g handles frames g, g+G, g+2G, ...
//  * One producer thread per CTA streams its part of each frame (three plane segments) into a ring
//    of S shared-memory stages with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) and an
//    L2 cache policy; full/empty mbarriers hand stages to the consumers and back.
//  * 256 consumer threads (8 warps that never wait for one another) each own K fixed atom slots of
//    the part.  The centred template b' and the weights of those slots live in REGISTERS for the
//    whole kernel, so the template costs no shared-memory or L2 bandwidth per frame.
//  * Pass 1 (S = sum w a, H = sum (w a) b'^T, sum w|a|^2) reads the stage once; each warp transposes its 32 x 13 per-lane
//    sums through a padded shared-memory scratch (13 stores, 16 loads, one shuffle per lane), leaves its totals in red[slot] and
//    arrives on the frame's red_full mbarrier; the solver warp that owns the frame adds the 8 warps in a fixed tree and pushes
//    the CTA's 13 partial sums to every CTA of the cluster with st.async (data + mbarrier complete_tx in one DSMEM transaction).
//  * Solver warps add the C partials in rank order (bit-identical in every CTA), solve the 3x3
//    problem (kabsch3.cuh), publish R and centroid*R, and decide where the RMSD comes from: the
//    inner-product form E = Ga + Gb - 2 tr(R^T H) when its rounding bound moves the RMSD by less than
//    1e-10 A, the explicit residual of pass 2 otherwise (frames close to the template).
//  * Pass 2 runs only for flagged frames in fit mode, for every frame when the superposed coordinates
//    are written back or LOVO's per-atom distances are accumulated.  It trails pass 1 by D frames so the
//    exchange + solve latency hides behind later frames, and reads the part either from its
//    shared-memory stage (RESIDENT) or again from L2 (the stage is then released right after pass 1,
//    and two stages of a fat part are enough).
//
// Per-warp squared-error partials of pass 2 go to a small global array and a trailing kernel finishes
// rmsd[f] = sqrt(sum E[f][..] / n) in fixed order for the flagged frames.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "kabsch3.cuh"
#include "ptx.cuh"

namespace gcb {

namespace {

#ifndef GCB_CONSUMER_WARPS
#define GCB_CONSUMER_WARPS 8
#endif
// Consumer warps per CTA (template parameter W of the kernel): 8 in every instance of the product build.
// A build with -DGCB_DEV_WARPS=12 gives the per-atom-distance instances a third consumer warp per scheduler: 16 warps start
// with 128 registers, the producer / solver warpgroup gives up 48 with setmaxnreg and the three consumer warpgroups take 144
// each (no spills up to 8 atoms per thread), the per-frame reduction runs in two rounds of 6 sums so that its scratch still
// fits beside three resident stages.  Measured slower on every frame size (50 k x 5 k: 1.384 against 1.136 ms, 100 k x 10 k:
// 6.13 against 5.09, profiles/r2_tuning.md): the pass is bound by the chain last-warp-delivered -> cluster sums -> 3x3 solve ->
// pass 2 as much as by the consumers' issue rate, and twelve warps deliver later and reduce 1.5 x as much.  Kept as a tuning
// build only.
constexpr int kDefaultWarps = GCB_CONSUMER_WARPS;
#ifndef GCB_DEV_WARPS
#define GCB_DEV_WARPS GCB_CONSUMER_WARPS
#endif
constexpr int kDevWarps = GCB_DEV_WARPS;
#ifndef GCB_DEV_WARPS_MIN_K
#define GCB_DEV_WARPS_MIN_K 6     // atoms per thread of the 8-warp plan from which the 12-warp instance takes over
#endif
constexpr int kSolverWarps = 3;
constexpr int threads_cta(int W) { return 32 * W + 32 + 32 * kSolverWarps; }  // consumers + producer warp + solver warps
constexpr int kRegsHelper = 80, kRegsConsumer = 144;   // setmaxnreg targets of the 16-warp block: 3 * 128 * 144 + 128 * 80 = 64 K
constexpr int kMaxStages = 8;
// Exchange slots.  Slot j mod NS is rewritten by the peers' pushes of frame j + NS.  A peer can start pass 1 of frame j + NS once
// it has finished pass 2 of frame j + NS - D - 1, which needs every CTA's partial sums of that frame, i.e. this CTA's consumers
// are past pass 2 of frame j + NS - 2D - 2.  With NS >= 2D + 2 that frame is >= j, whose pass 2 waited for this CTA's solver to
// publish frame j -- after it had read the slot.  (NS = 2D + 1 left a window in which a delayed solver warp could read sums of
// frame j + NS.)
constexpr int kMaxSlots = 14;            // >= 2*D+2 with D <= 6, and > kSolverWarps
constexpr int kResidentSlots = 8;        // plans that keep the part resident run at most D = 3 frames ahead
#ifndef GCB_RED_ROW
#define GCB_RED_ROW 34
#endif
constexpr int kRedRow = GCB_RED_ROW;     // transposition scratch of the per-frame reduction: [13 sums][32 lanes], rows padded by two:
                                         // 16-byte aligned rows (the tree reads them with LDS.128: 8 loads instead of 16) whose
                                         // starts fall on different bank groups
// The transposition runs in R rounds of QR sums: one round of 13 with 8 warps; with 12 warps (12 sums: those instances never
// carry sum w|a|^2) two rounds of 6, so that the scratch of 12 warps still fits beside three resident stages.
constexpr int red_rounds(int W) { return W > 8 ? 2 : 1; }
constexpr int red_per_round(int W) { return W > 8 ? 6 : 13; }
constexpr int red_warp_doubles(int W) { return red_per_round(W) * kRedRow; }
constexpr size_t red_scratch_bytes(int W) { return ((sizeof(double) * red_warp_doubles(W) * W + 127) / 128) * 128; }
constexpr int kMaxCluster = 8;
constexpr int kSmemBudget = 227 * 1024;

struct ClusterParams {
    SuperParams p;
    int cluster;        // C
    int part;           // P: atoms per CTA (multiple of 16)
    int stages;         // S
    int depth;          // D
    int slots;          // NS >= 2*D+2
    int nclusters;      // G
    int write_back;
    int always_explicit;  // pass 2 runs for every frame (write-back / per-atom sums / caller's choice)
    double* epart;      // [nframes][C][8 consumer warps] squared-error partials
};

using namespace ptx;

#ifdef GCB_TRACE
__device__ long long g_trace[8][512];
#define TRACE(row, j) do { if (blockIdx.x == 0 && (j) < 512) g_trace[row][j] = clock64(); } while (0)
#ifndef GCB_TRACE_WARP
#define GCB_TRACE_WARP 0      // the consumer warp whose times rows 4..7 hold
#endif
#define TRACE_TID (32 * GCB_TRACE_WARP)
#else
#define TRACE(row, j) do {} while (0)
#define TRACE_TID 0
#endif

template <int NSLOTS, int W>
struct __align__(16) SharedT {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t part_full[NSLOTS];           // cluster -> solver: every CTA's partial sums of a frame have landed
    uint64_t solved[NSLOTS];              // solver -> consumers: xf[slot] holds the frame's transform
    double partials[NSLOTS][kMaxCluster][14];   // 12 sums + sum w|a|^2
    double xf[NSLOTS][16];                // R[9], centroid*R [3], [12] = 1.0 when pass 2 must run for the RMSD
    double red[NSLOTS][W][14];            // per-warp totals of a frame's sums
    uint64_t red_full[NSLOTS];            // consumers -> solver: red[slot] holds every consumer warp's totals of the frame
};
// bytes in front of the stage ring: barriers and exchange slots, then the reduction scratch
template <int W>
constexpr size_t fixed_smem_w(bool resident) {
    return ((resident ? sizeof(SharedT<kResidentSlots, W>) : sizeof(SharedT<kMaxSlots, W>)) + 127) / 128 * 128 + red_scratch_bytes(W);
}
constexpr size_t fixed_smem(bool resident, int W) { return W == kDevWarps ? fixed_smem_w<kDevWarps>(resident) : fixed_smem_w<kDefaultWarps>(resident); }

// sqrt of a positive double in five fp64 operations after the hardware seed.  y = reciprocal-square-root seed (relative error
// eps, |eps| < 2^-20), g = x y = sqrt(x) (1 + eps), e = 1 - g y = -(2 eps + eps^2); since (1 - e)^(-1/2) = 1 + e/2 + 3 e^2/8 + O(e^3),
// sqrt(x) = g (1 + e (1/2 + 3/8 e)) with relative error ~2.5 |eps|^3 < 2^-58 before rounding -- one correction of cubic order.
// (Round 1 wrote the same correction in six operations; this pass is bound by the schedulers' issue ports, two cycles per fp64
// instruction, so the sixth cost 2 % of it: profiles/r2_tuning.md, section 12.)  The argument must be bounded away from zero (a
// zero or denormal gives seed = inf): the caller folds 1e-290 into the sum of squares, which moves the square root by less than
// 1e-145 and saves the three integer instructions of a guard.
__device__ __forceinline__ double fast_sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double g = x * y;
    const double e = fma(-g, y, 1.0);
    const double c = fma(e, 0.375, 0.5);
    const double ec = e * c;
    return fma(g, ec, g);
}

// Slot k of a consumer thread <-> atom of the CTA's part.  With an even K a thread owns PAIRS of neighbouring atoms (2 tid and
// 2 tid + 1 of every block of 512 atoms), so its loads and stores are 16 bytes wide (LDS.128 / LDG.128 / STG.128): half the
// memory instructions of the one-atom-per-256 layout, which remains for odd K.  Parts hold a multiple of 16 atoms, so a pair is
// inside the part or outside it as a whole.
// Measured (profiles/r2_tuning.md): the pairs pay in the per-atom-distance pass (50 k x 5 k: 1.141 -> 1.126 ms); the wider
// loads cost registers, and the fit-only K = 14 instance (3.27 -> 3.33 ms) and the write-back K = 10 instance (7.7 -> 10.6 ms,
// spills inside the store loop) lose, so only the DEV instances pair.
template <int K, bool DEV, int W>
struct Slots {
    static constexpr int kConsumers = 32 * W;
    static constexpr bool PAIR = DEV && (K % 2 == 0);
    static __device__ __forceinline__ int base(int tid) { return PAIR ? 2 * tid : tid; }
    static __device__ __forceinline__ constexpr int off(int k) { return PAIR ? (k >> 1) * 2 * kConsumers + (k & 1) : k * kConsumers; }
};

// pass 2 for one thread's K atom slots.  RESIDENT: the frame part is still in its shared-memory stage
// (plane stride K*256); otherwise it is re-read from global memory, where it still sits in L2 (plane
// stride np, streaming loads).  WB = also store the superposed coordinates.
// MASKED: the fit uses a subset of the atoms (fit_mask bit k = slot k is in it); otherwise every real atom is.
// DEV: per-atom distances are accumulated and the frame's squared error is not (align.RMSDTraj never asks for it).
template <int K, bool MASKED, bool DEV, bool WB, bool RESIDENT, int W>
__device__ __forceinline__ double pass2_atoms(const double* __restrict__ src, long plane, const double* xf,
                                              const double (&bx)[K], const double (&by)[K], const double (&bz)[K],
                                              uint32_t fit_mask, double (&dev)[DEV ? K : 1], double* frw, int np,
                                              uint32_t real_mask, uint32_t in_mask, double b0, double b1, double b2) {
    using SL = Slots<K, DEV, W>;
    const double r0 = xf[0], r1 = xf[1], r2 = xf[2], r3 = xf[3], r4 = xf[4], r5 = xf[5], r6 = xf[6], r7 = xf[7], r8 = xf[8];
    const double g0 = -xf[9], g1 = -xf[10], g2 = -xf[11];
    double e0 = 0.0, e1 = 0.0;
#ifdef GCB_PASS2_CH
    constexpr int CH = GCB_PASS2_CH < K ? GCB_PASS2_CH : K;   // tuning builds (even for even K)
#else
    // atoms in flight per thread: enough independent chains, bounded register use; whole pairs when the slots are paired
    constexpr int CH = SL::PAIR ? 4 : ((K % 5 == 0) ? 5 : 4);
#endif
#pragma unroll
    for (int kb = 0; kb < K; kb += CH) {
        double x[CH], y[CH], z[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int k = kb + c;
            if (k >= K) { x[c] = y[c] = z[c] = 0.0; continue; }
            if (SL::PAIR) {
                if (c & 1) continue;   // loaded together with its even neighbour
                const double* q = src + SL::off(k);
                double2 vx, vy, vz;
                if (RESIDENT) {
                    vx = *reinterpret_cast<const double2*>(q); vy = *reinterpret_cast<const double2*>(q + plane);
                    vz = *reinterpret_cast<const double2*>(q + 2 * plane);
                } else if ((in_mask >> k) & 1u) {
                    vx = __ldcs(reinterpret_cast<const double2*>(q)); vy = __ldcs(reinterpret_cast<const double2*>(q + plane));
                    vz = __ldcs(reinterpret_cast<const double2*>(q + 2 * plane));
                } else {
                    vx = vy = vz = make_double2(0.0, 0.0);
                }
                x[c] = vx.x; y[c] = vy.x; z[c] = vz.x;
                if (c + 1 < CH) { x[c + 1] = vx.y; y[c + 1] = vy.y; z[c + 1] = vz.y; }
            } else if (RESIDENT) {
                x[c] = src[SL::off(k)]; y[c] = src[plane + SL::off(k)]; z[c] = src[2 * plane + SL::off(k)];
            } else if ((in_mask >> k) & 1u) {
                x[c] = __ldcs(src + SL::off(k)); y[c] = __ldcs(src + plane + SL::off(k)); z[c] = __ldcs(src + 2 * plane + SL::off(k));
            } else {
                x[c] = 0.0; y[c] = 0.0; z[c] = 0.0;
            }
        }
        // one atom: rotate, distance to the template, sums; returns the superposed coordinates for the write-back
        auto atom = [&](int c, double& ox, double& oy, double& oz) {
            const int k = kb + c;
            // (a - centroid) R  computed as  a R - centroid R
            const double qx = fma(z[c], r6, fma(y[c], r3, fma(x[c], r0, g0)));
            const double qy = fma(z[c], r7, fma(y[c], r4, fma(x[c], r1, g1)));
            const double qz = fma(z[c], r8, fma(y[c], r5, fma(x[c], r2, g2)));
            const double dx = qx - bx[k], dy = qy - by[k], dz = qz - bz[k];
            if (DEV) {
                // 1e-290 keeps the argument of the guard-free square root away from zero (it moves the distance by < 1e-145)
                dev[k] += fast_sqrt_pos(fma(dz, dz, fma(dy, dy, fma(dx, dx, 1e-290))));
            } else if (((MASKED ? fit_mask : real_mask) >> k) & 1u) {   // atoms of the fit count once; padding slots never
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                if (k & 1) e1 += d2; else e0 += d2;
            }
            if (WB) { ox = qx + b0; oy = qy + b1; oz = qz + b2; }
        };
        if (SL::PAIR) {
#pragma unroll
            for (int c = 0; c < CH; c += 2) {
                const int k = kb + c;
                if (k >= K) continue;
                double ax = 0.0, ay = 0.0, az = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
                atom(c, ax, ay, az);
                atom(c + 1, cx, cy, cz);
                if (WB) {
                    double* q = frw + SL::off(k);
                    const uint32_t both = (real_mask >> k) & 3u;
                    if (both == 3u) {
                        *reinterpret_cast<double2*>(q) = make_double2(ax, cx);
                        *reinterpret_cast<double2*>(q + np) = make_double2(ay, cy);
                        *reinterpret_cast<double2*>(q + 2L * np) = make_double2(az, cz);
                    } else if (both == 1u) {   // the last real atom of an odd count
                        q[0] = ax; q[np] = ay; q[2L * np] = az;
                    }
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int k = kb + c;
                if (k >= K) continue;
                double ax = 0.0, ay = 0.0, az = 0.0;
                atom(c, ax, ay, az);
                if (WB && ((real_mask >> k) & 1u)) {
                    double* q = frw + SL::off(k);
                    q[0] = ax; q[np] = ay; q[2L * np] = az;
                }
            }
        }
    }
    return e0 + e1;
}

// NOWB: this instance never writes frames back (the per-atom-distance pass of align.RMSDTraj without an aligned trajectory): the
// write pointer, its per-frame update and the write-back variants of pass 2 are not compiled in -- ~20 instructions per warp and
// frame of a pass that is bound by issue slots (profiles/r2_tuning.md, section 12).
template <int K, bool MASKED, bool DEV, bool RESIDENT, int W, bool NOWB = false>
__global__ void __launch_bounds__(threads_cta(W), 1) super_cluster_kernel(const ClusterParams cp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NS = RESIDENT ? kResidentSlots : kMaxSlots;
    constexpr int kConsumerWarps = W, kConsumers = 32 * W, kThreadsCta = threads_cta(W);
    constexpr bool REDEAL = threads_cta(W) == 512;   // four warpgroups: registers re-dealt between the roles (setmaxnreg)
    static_assert(DEV || red_rounds(W) == 1, "two-round reduction carries 12 sums only");
    using Shared = SharedT<NS, W>;
    Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
    constexpr int kStageDoubles = 3 * K * kConsumers;
    double* red_scratch = reinterpret_cast<double*>(smem_raw + fixed_smem_w<W>(RESIDENT) - red_scratch_bytes(W));
    double* stage_base = reinterpret_cast<double*>(smem_raw + fixed_smem_w<W>(RESIDENT));

    const SuperParams& p = cp.p;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;   // warps: consumers 0..kConsumerWarps-1, producer, solvers
    const int C = cp.cluster, S = cp.stages, D = cp.depth, G = cp.nclusters;
    const uint32_t rank = cluster_ctarank();
    const long cid = cluster_id_x();
    const int part0 = (int)rank * cp.part;                       // first atom of this CTA's part
    int part_len = p.np - part0;                                 // atoms (incl. padding) actually loaded
    if (part_len > cp.part) part_len = cp.part;
    if (part_len < 0) part_len = 0;
    const int nloc = (cid < p.nframes) ? (int)((p.nframes - cid + G - 1) / G) : 0;   // frames of this cluster

    // zero the stage ring once: slots beyond part_len are never written by TMA and must read as 0
    for (int i = tid; i < S * kStageDoubles; i += kThreadsCta) stage_base[i] = 0.0;
    if (red_rounds(W) > 1) {   // entry 12 of red[][] (sum w|a|^2) is not written by the two-round reduction and travels as 0
        for (int i = tid; i < NS * W * 14; i += kThreadsCta) (&sh.red[0][0][0])[i] = 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], kConsumerWarps); }
        for (int s = 0; s < NS; s++) {
            mbar_init(&sh.part_full[s], 1);
            mbar_init(&sh.solved[s], 1);
            mbar_init(&sh.red_full[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // make the zero-fill (generic proxy) visible to the async proxy before any TMA write lands next to it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == kConsumerWarps + 1 && lane == 0) {
        for (int s = 0; s < NS; s++) mbar_arrive_expect_tx(&sh.part_full[s], (uint32_t)(C * 13 * sizeof(double)));
    }
    cluster_sync_all();   // every CTA's barriers are initialised and armed before any remote store

    // REDEAL: all four warps of a warpgroup execute the same setmaxnreg, first thing in their role; the consumers' request waits
    // for the helpers' release
    if (warp >= kConsumerWarps) {
    if (REDEAL) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsHelper));
    if (warp == kConsumerWarps) {
        // ===================== producer =====================
        if (lane == 0 && part_len > 0) {
            const uint64_t pol = RESIDENT ? policy_evict_first() : policy_evict_normal();
            const uint32_t seg_bytes = (uint32_t)part_len * 8u;
            const double* src = p.frames + cid * p.frame_stride + part0;
            const long src_step = (long)G * p.frame_stride;
            int s = 0;
            uint32_t ph = 1;   // parity of the previous use of the stage
            for (int j = 0; j < nloc; j++, src += src_step) {
                if (j >= S) mbar_wait(&sh.empty[s], ph);
                double* dst = stage_base + (long)s * kStageDoubles;
                mbar_arrive_expect_tx(&sh.full[s], 3u * seg_bytes);
                tma_load_1d(dst, src, seg_bytes, &sh.full[s], pol);
                tma_load_1d(dst + K * kConsumers, src + p.np, seg_bytes, &sh.full[s], pol);
                tma_load_1d(dst + 2 * K * kConsumers, src + 2L * p.np, seg_bytes, &sh.full[s], pol);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        } else if (lane == 0) {
            // an empty part still has to complete the full barriers its consumers wait on
            int s = 0;
            uint32_t ph = 1;
            for (int j = 0; j < nloc; j++) {
                if (j >= S) mbar_wait(&sh.empty[s], ph);
                mbar_arrive(&sh.full[s]);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ===================== solvers: warp w takes local frames w, w + kSolverWarps, ... =====================
        // Per frame: rank-ordered sum of the cluster's partial sums, 3x3 solve, publish R and centroid R.
        const int sw = warp - (kConsumerWarps + 1);
        int slot = sw;                                  // NS > kSolverWarps
        uint32_t sph = 0;
        for (int j = sw; j < nloc; j += kSolverWarps) {
            {
                // the CTA's own sums: every consumer warp has delivered its totals of frame j; add them in a fixed tree and push
                // the result to every CTA of the cluster (data + complete_tx on the receiver's barrier in one DSMEM transaction)
                mbar_wait(&sh.red_full[slot], sph);
                if (lane < 13) {
                    double wv[kConsumerWarps];
#pragma unroll
                    for (int w = 0; w < kConsumerWarps; w++) wv[w] = sh.red[slot][w][lane];
#pragma unroll
                    for (int span = 1; span < kConsumerWarps; span *= 2) {
#pragma unroll
                        for (int w = 0; w + span < kConsumerWarps; w += 2 * span) wv[w] += wv[w + span];
                    }
                    const uint32_t dst = smem_u32(&sh.partials[slot][rank][lane]);
                    const uint32_t bar = smem_u32(&sh.part_full[slot]);
                    for (int c = 0; c < C; c++) st_async_f64(mapa(dst, (uint32_t)c), wv[0], mapa(bar, (uint32_t)c));
                }
                if (lane == 0) TRACE(1, j);
                mbar_wait(&sh.part_full[slot], sph);
                if (lane == 0) TRACE(2, j);
                double s = 0.0;
                if (lane < 13) {
                    for (int r = 0; r < C; r++) s += sh.partials[slot][r][lane];   // rank order: identical in every CTA
                }
                double sums[13];
#pragma unroll
                for (int k = 0; k < 13; k++) sums[k] = __shfl_sync(0xffffffffu, s, k);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&sh.part_full[slot], (uint32_t)(C * 13 * sizeof(double)));  // re-arm for frame j+NS
                    const long f = cid + (long)j * G;
                    const double cen[3] = {sums[0] * p.inv_n, sums[1] * p.inv_n, sums[2] * p.inv_n};
                    double h[9], r[9];
#pragma unroll
                    for (int k = 0; k < 9; k++) h[k] = sums[3 + k];
                    int rc = k3::kabsch3(h, r);   // non-finite sums fail the norm guards of both solvers
                    if (rc != 0) {
                        if (p.status) atomicExch(p.status, (int)GCB_ERR_SVD);
#pragma unroll
                        for (int k = 0; k < 9; k++) r[k] = (k % 4 == 0) ? 1.0 : 0.0;
                    }
                    double* xf = sh.xf[slot];
#pragma unroll
                    for (int k = 0; k < 9; k++) xf[k] = r[k];
                    xf[9] = fma(cen[2], r[6], fma(cen[1], r[3], cen[0] * r[0]));
                    xf[10] = fma(cen[2], r[7], fma(cen[1], r[4], cen[0] * r[1]));
                    xf[11] = fma(cen[2], r[8], fma(cen[1], r[5], cen[0] * r[2]));
                    bool formula_ok = false;
                    double e = 0.0;
                    if (!cp.always_explicit) {
                        const double ga = sums[12] - (sums[0] * sums[0] + sums[1] * sums[1] + sums[2] * sums[2]) * p.inv_n;
                        const double tr0 = fma(r[2], h[2], fma(r[1], h[1], r[0] * h[0]));
                        const double tr1 = fma(r[5], h[5], fma(r[4], h[4], r[3] * h[3]));
                        const double tr2 = fma(r[8], h[8], fma(r[7], h[7], r[6] * h[6]));
                        e = ga + p.gb - 2.0 * ((tr0 + tr1) + tr2);
                        const double lhs = (64.0 * 2.220446049250313e-16) * (sums[12] + p.gb);
                        formula_ok = rc == 0 && e > 0.0 && lhs * lhs <= 4e-20 * (p.n_fit * e);
                    }
                    xf[12] = formula_ok ? 0.0 : 1.0;
                    TRACE(3, j);
                    mbar_arrive(&sh.solved[slot]);
                    if (rank == 0) {
                        if (p.need_explicit) p.need_explicit[f] = formula_ok ? 0 : 1;
                        if (formula_ok && p.rmsd) p.rmsd[f] = sqrt(e * p.inv_n);
                        if (p.rot) {
#pragma unroll
                            for (int k = 0; k < 9; k++) p.rot[f * 9 + k] = r[k];
                        }
                        if (p.trans1) { p.trans1[f * 3] = -cen[0]; p.trans1[f * 3 + 1] = -cen[1]; p.trans1[f * 3 + 2] = -cen[2]; }
                        if (p.trans2) { p.trans2[f * 3] = p.bbar[0]; p.trans2[f * 3 + 1] = p.bbar[1]; p.trans2[f * 3 + 2] = p.bbar[2]; }
                    }
                }
            }
            __syncwarp();
            slot += kSolverWarps;
            if (slot >= NS) { slot -= NS; sph ^= 1u; }
        }
    }
    } else {
        if (REDEAL) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsConsumer));
        // ===================== consumers (warps run free of one another) =====================
        // fixed atom slots (Slots<K>: pairs of neighbouring atoms when K is even); template and weights stay in registers
        using SL = Slots<K, DEV, W>;
        const int tbase = SL::base(tid);
        double bx[K], by[K], bz[K], dev[DEV ? K : 1];
        uint32_t real_mask = 0, fit_mask = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int li = tbase + SL::off(k);
            const int gi = part0 + li;
            const bool in = li < part_len;
            bx[k] = in ? p.refc[gi] : 0.0;
            by[k] = in ? p.refc[p.np + gi] : 0.0;
            bz[k] = in ? p.refc[2L * p.np + gi] : 0.0;
            if (MASKED && in && gi < p.natoms && p.weight[gi] != 0.0) fit_mask |= 1u << k;   // weights are 0 or 1 here
            if (DEV) dev[k] = 0.0;
            if (in && gi < p.natoms) real_mask |= 1u << k;   // slots that are real atoms of THIS part
        }
        uint32_t in_mask = 0;
#pragma unroll
        for (int k = 0; k < K; k++) if (tbase + SL::off(k) < part_len) in_mask |= 1u << k;
        int s1 = 0, s2 = 0, slot1 = 0, slot2 = 0;      // stage / exchange slot of pass 1 (frame j) and pass 2 (frame j-D)
        uint32_t ph1 = 0, ph2 = 0;                     // parities: full[s1], solved[slot2]
        const long fstep = (long)G * p.frame_stride;
        const double* fr2 = p.frames + cid * p.frame_stride + part0 + tbase;     // frame of pass 2 (global / L2 copy)
        double* frw = (!NOWB && cp.write_back) ? (p.frames_rw + cid * p.frame_stride + part0 + tbase) : nullptr;
        double* ep = cp.epart + (cid * C + rank) * kConsumerWarps + warp;
        const long ep_step = (long)G * C * kConsumerWarps;
        for (int j = 0; j < nloc + D; j++) {
            if (j < nloc) {
                // ---- pass 1 of local frame j ----
                double acc[12];
#pragma unroll
                for (int k = 0; k < 12; k++) acc[k] = 0.0;
                double aa0 = 0.0, aa1 = 0.0;   // sum w |a|^2 for the one-pass RMSD
                if (tid == TRACE_TID) TRACE(4, j);
                mbar_wait(&sh.full[s1], ph1);
                if (tid == TRACE_TID) TRACE(5, j);
                const double* sx = stage_base + (long)s1 * kStageDoubles + tbase;
                const double* sy = sx + K * kConsumers;
                const double* sz = sy + K * kConsumers;
                // one atom of pass 1: S += w a, H += (w a) b'^T, sum w|a|^2
                auto atom = [&](double x, double y, double z, int k) {
                    if (MASKED) {
                        // Atoms outside the fit must contribute nothing.  Clearing the HIGH word alone leaves a denormal below
                        // 2^32 * 2^-1074 = 2e-314 -- nothing a sum of coordinates can see -- for one select per coordinate instead
                        // of two (and no fp64 multiply by a weight).
                        const bool on = (fit_mask >> k) & 1u;
                        x = __hiloint2double(on ? __double2hiint(x) : 0, __double2loint(x));
                        y = __hiloint2double(on ? __double2hiint(y) : 0, __double2loint(y));
                        z = __hiloint2double(on ? __double2hiint(z) : 0, __double2loint(z));
                    }
                    if (!DEV) {   // only the one-pass RMSD needs sum w|a|^2; the per-atom-distance pass always runs pass 2
                        const double n2 = fma(z, z, fma(y, y, x * x));
                        if (k & 1) aa1 += n2; else aa0 += n2;
                    }
                    acc[0] += x; acc[1] += y; acc[2] += z;
                    acc[3] = fma(x, bx[k], acc[3]); acc[4] = fma(x, by[k], acc[4]); acc[5] = fma(x, bz[k], acc[5]);
                    acc[6] = fma(y, bx[k], acc[6]); acc[7] = fma(y, by[k], acc[7]); acc[8] = fma(y, bz[k], acc[8]);
                    acc[9] = fma(z, bx[k], acc[9]); acc[10] = fma(z, by[k], acc[10]); acc[11] = fma(z, bz[k], acc[11]);
                };
                if (SL::PAIR) {
#pragma unroll
                    for (int k = 0; k < K; k += 2) {
                        const double2 vx = *reinterpret_cast<const double2*>(sx + SL::off(k));
                        const double2 vy = *reinterpret_cast<const double2*>(sy + SL::off(k));
                        const double2 vz = *reinterpret_cast<const double2*>(sz + SL::off(k));
                        atom(vx.x, vy.x, vz.x, k);
                        atom(vx.y, vy.y, vz.y, k + 1);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < K; k++) atom(sx[SL::off(k)], sy[SL::off(k)], sz[SL::off(k)], k);
                }
                if (!RESIDENT) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.empty[s1]);   // pass 2 re-reads the frame from L2: free the stage now
                }
                // ---- warp totals -> red[slot]; the last warp to deliver sums the 8 warps in fixed order and pushes
                //      the CTA's partial sums to every CTA of the cluster ----
                // The 32 x 13 per-lane sums are transposed through the warp's own scratch (rows padded to 33: stores and loads
                // are conflict-free): lane q + 16 h adds half h of sum q in a fixed tree and one shuffle joins the halves --
                // 13 stores, 16 loads and one shuffle per lane instead of a 15-shuffle butterfly with its selects.
                {
                    constexpr int R = red_rounds(W), QR = red_per_round(W);
                    constexpr int NQ = DEV ? 12 : 13;
                    double* tr = red_scratch + warp * red_warp_doubles(W);
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        if (r > 0) __syncwarp();   // the previous round's loads are done with the scratch
#pragma unroll
                        for (int qq = 0; qq < QR; qq++) {
                            const int q = r * QR + qq;
                            if (q < 12) tr[qq * kRedRow + lane] = acc[q];
                            else if (q == 12 && !DEV) tr[qq * kRedRow + lane] = aa0 + aa1;
                        }
                        __syncwarp();
                        const int ql = lane & 15;
                        double v = 0.0;
                        if (ql < QR && r * QR + ql < NQ) {
                            const double* row = tr + ql * kRedRow + (lane & 16);
                            double s0, s1, s2, s3, s4, s5, s6, s7;
                            if constexpr (kRedRow % 2 == 0) {
                                const double2* r2 = reinterpret_cast<const double2*>(row);
                                const double2 u0 = r2[0], u1 = r2[1], u2 = r2[2], u3 = r2[3], u4 = r2[4], u5 = r2[5], u6 = r2[6], u7 = r2[7];
                                s0 = u0.x + u0.y; s1 = u1.x + u1.y; s2 = u2.x + u2.y; s3 = u3.x + u3.y;
                                s4 = u4.x + u4.y; s5 = u5.x + u5.y; s6 = u6.x + u6.y; s7 = u7.x + u7.y;
                            } else {
                                s0 = row[0] + row[1]; s1 = row[2] + row[3]; s2 = row[4] + row[5]; s3 = row[6] + row[7];
                                s4 = row[8] + row[9]; s5 = row[10] + row[11]; s6 = row[12] + row[13]; s7 = row[14] + row[15];
                            }
                            v = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
                        }
                        v += __shfl_down_sync(0xffffffffu, v, 16);
                        if (lane < QR && r * QR + lane < 13) sh.red[slot1][warp][r * QR + lane] = v;   // DEV: entry 12 is 0 (no sum w|a|^2 in that mode)
                    }
                }
                __syncwarp();   // the lanes' stores to red[] are ordered before lane 0's release below; the scratch is free again
                if (tid == TRACE_TID) TRACE(7, j);
                if (lane == 0) mbar_arrive(&sh.red_full[slot1]);   // release: the warp's stores to red[] (ordered by the __syncwarp above)
                if (++s1 == S) { s1 = 0; ph1 ^= 1u; }
                if (++slot1 == NS) slot1 = 0;
            }
            if (j >= D) {
                // ---- pass 2 of local frame j - D ----
                mbar_wait(&sh.solved[slot2], ph2);
                if (tid == TRACE_TID) TRACE(6, j - D);
                double e = 0.0;
                const bool run_pass2 = DEV || cp.always_explicit || sh.xf[slot2][12] != 0.0;   // warp-uniform; DEV: known at compile time
                if (!run_pass2) {
                    if (RESIDENT) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sh.empty[s2]);
                        if (++s2 == S) s2 = 0;
                    }
                } else if (RESIDENT) {
                    const double* sx = stage_base + (long)s2 * kStageDoubles + tbase;
                    if (!NOWB && frw) e = pass2_atoms<K, MASKED, DEV, true, true, W>(sx, K * kConsumers, sh.xf[slot2], bx, by, bz, fit_mask, dev, frw, p.np, real_mask, in_mask, p.bbar[0], p.bbar[1], p.bbar[2]);
                    else e = pass2_atoms<K, MASKED, DEV, false, true, W>(sx, K * kConsumers, sh.xf[slot2], bx, by, bz, fit_mask, dev, nullptr, p.np, real_mask, in_mask, 0.0, 0.0, 0.0);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.empty[s2]);   // this warp is done reading the stage
                    if (++s2 == S) s2 = 0;
                } else {
                    if (!NOWB && frw) e = pass2_atoms<K, MASKED, DEV, true, false, W>(fr2, p.np, sh.xf[slot2], bx, by, bz, fit_mask, dev, frw, p.np, real_mask, in_mask, p.bbar[0], p.bbar[1], p.bbar[2]);
                    else e = pass2_atoms<K, MASKED, DEV, false, false, W>(fr2, p.np, sh.xf[slot2], bx, by, bz, fit_mask, dev, nullptr, p.np, real_mask, in_mask, 0.0, 0.0, 0.0);
                }
                if (run_pass2 && !DEV) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) e += shx(e, o);
                    if (lane == 0) *ep = e;          // per-warp squared error; finish_rmsd_kernel adds them in fixed order
                }
                ep += ep_step;
                fr2 += fstep;
                if (!NOWB && frw) frw += fstep;
                if (++slot2 == NS) { slot2 = 0; ph2 ^= 1u; }
            }
        }
        if (DEV) {
            double* row = p.devpart + cid * (long)p.np;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int li = tbase + SL::off(k);
                if (li < part_len) row[part0 + li] = dev[k];
            }
        }
    }
    // nobody leaves while a peer may still store into this CTA's shared memory
    __syncthreads();
    cluster_sync_all();
}

__global__ void __launch_bounds__(256) finish_rmsd_kernel(const double* __restrict__ epart, int C, int kConsumerWarps, long nframes, double inv_n,
                                                         const unsigned char* __restrict__ need_explicit, double* __restrict__ rmsd) {
    const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    if (need_explicit && !need_explicit[f]) return;   // the solver already wrote the one-pass RMSD
    double s = 0.0;
    for (int r = 0; r < C * kConsumerWarps; r++) s += epart[f * C * kConsumerWarps + r];
    rmsd[f] = sqrt(s * inv_n);
}

struct Plan {
    int cluster, part, kinst, stages, depth, resident, warps;
    size_t smem;
};

constexpr int kInst[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14};

// Choose CTAs per frame, atoms per thread, stage count, pipeline depth.
//  resident = 1: pass 2 reads the frame part from its shared-memory stage (needs depth <= stages - 2)
//  resident = 0: pass 2 re-reads the part from L2; stages only cover the HBM latency, depth <= 6
bool make_plan_w(const Tuning& tu, int np, bool heavy, int W, Plan* pl) {
    const int kConsumers = 32 * W;
    const int tune_cluster = tu.cluster, tune_depth = tu.depth, tune_stages = tu.stages;
    const int tune_resident = tu.resident < 0 ? -1 : (tu.resident & 1);
    // Measured on B200 (profiles/r1_tuning.md).  Fit-only passes skip pass 2 for almost every frame, so the stage can be
    // released right after pass 1 (pass 2, when needed, re-reads the part from L2): two stages of the fattest part that
    // fits win (up to 14 atoms per thread, fewest CTAs per frame, clusters of <= 3 tile the GPCs almost exactly).
    // Passes that always run pass 2 (write-back, per-atom sums) keep the part resident in shared memory: <= 10 atoms per
    // thread, >= 3 stages.
    const int resident = tune_resident >= 0 ? tune_resident : (heavy ? 1 : 0);
    const int kmax = resident ? (heavy ? 10 : 10) : 14;
    const size_t fixed = fixed_smem(resident != 0, W);
    for (int c = 1; c <= kMaxCluster; c++) {
        if (tune_cluster && c != tune_cluster) continue;
        const int part = round_up((np + c - 1) / c, 16);
        const int kneed = (part + kConsumers - 1) / kConsumers;
        int kinst = 0;
        for (int k : kInst) if (k >= kneed) { kinst = k; break; }
        if (!kinst) continue;
        if (!tune_cluster && kinst > kmax && c < kMaxCluster) continue;
        const size_t stage = (size_t)3 * kinst * kConsumers * sizeof(double);
        int stages = (int)((kSmemBudget - fixed) / stage);
        if (stages > kMaxStages) stages = kMaxStages;
        if (!resident && stages > (kinst <= 6 ? 4 : 2)) stages = (kinst <= 6 ? 4 : 2);
        if (tune_stages && tune_stages < stages) stages = tune_stages;
        if (stages < (resident ? 3 : 2)) continue;
        int depth;
        if (resident) {
            depth = stages - 2 > 2 ? 2 : stages - 2;
            if (tune_depth >= 1 && tune_depth <= stages - 2 && 2 * tune_depth + 2 <= kResidentSlots) depth = tune_depth;
        } else {
            depth = kinst <= 6 ? 6 : 3;   // costs no shared memory here: only the exchange slots
            if (tune_depth >= 1 && tune_depth <= 6) depth = tune_depth;
            static_assert(2 * 6 + 2 <= kMaxSlots, "exchange slots must cover 2 D + 2 frames");
        }
        pl->cluster = c; pl->part = part; pl->kinst = kinst; pl->stages = stages; pl->depth = depth;
        pl->resident = resident;
        pl->warps = W;
        pl->smem = fixed + stage * stages;
        return true;
    }
    return false;
}

// dev = the pass accumulates per-atom distances.  In a -DGCB_DEV_WARPS=12 build that pass runs the 12-warp instance on fat parts
// (>= GCB_DEV_WARPS_MIN_K atoms per thread of the 8-warp plan); tuning: resident = 1 | (warps << 4) forces the warp count.
bool make_plan(const Tuning& tu, int np, bool heavy, bool dev, Plan* pl) {
    if (!make_plan_w(tu, np, heavy, kDefaultWarps, pl)) return false;
    const int forced = tu.resident > 0 ? (tu.resident >> 4) : 0;
    if (!dev || !pl->resident || kDevWarps == kDefaultWarps || forced == kDefaultWarps) return true;
    if (forced != kDevWarps && pl->kinst < GCB_DEV_WARPS_MIN_K) return true;
    Plan p12;
    if (make_plan_w(tu, np, heavy, kDevWarps, &p12) && p12.resident && p12.kinst <= 8) *pl = p12;
    return true;
}

template <int K, bool WT, bool DV, bool RES, int W, bool NOWB = false>
int launch_inst(gcb_traj* t, const ClusterParams& cp, const Plan& pl, int* nclusters_out) {
#ifndef GCB_NO_NOWB
    if constexpr (DV && RES && !NOWB) {
        if (!cp.write_back) return launch_inst<K, WT, DV, RES, W, true>(t, cp, pl, nclusters_out);   // the instance without write-back code
    }
#endif
    auto kern = super_cluster_kernel<K, WT, DV, RES, W, NOWB>;
    // the occupancy query costs tens of microseconds: once per (instance, device, shape).  The cache is shared by every handle
    // of the process and guarded; a miss only repeats the query.
    static std::mutex cache_mu;
    static int cached_dev = -1, cached_cluster = 0, cached_max = 0;
    static size_t cached_smem = 0;
    GCB_CUDA(opt_in_smem(kern, t->dev));
    bool hit;
    int max_clusters;
    {
        std::lock_guard<std::mutex> lock(cache_mu);
        hit = cached_dev == t->dev && cached_cluster == pl.cluster && cached_smem == pl.smem;
        max_clusters = cached_max;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(threads_cta(W));
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = t->stream;
    cfg.gridDim = dim3(pl.cluster * (t->sm_count / pl.cluster));
    if (!hit) {
        GCB_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
        std::lock_guard<std::mutex> lock(cache_mu);
        cached_dev = t->dev; cached_cluster = pl.cluster; cached_smem = pl.smem; cached_max = max_clusters;
    }
    if (max_clusters < 1) { set_error("cluster of %d CTAs with %zu bytes of shared memory cannot be scheduled", pl.cluster, pl.smem); return GCB_ERR_CUDA; }
    long g = max_clusters;
    if (g > cp.p.nframes) g = cp.p.nframes;
    ClusterParams c2 = cp;
    c2.nclusters = (int)g;
    cfg.gridDim = dim3((unsigned)(g * pl.cluster));
    GCB_CUDA(cudaLaunchKernelEx(&cfg, kern, c2));
    GCB_LAUNCHED();
    *nclusters_out = (int)g;
    t->last_plan[0] = pl.cluster; t->last_plan[1] = pl.part; t->last_plan[2] = pl.kinst; t->last_plan[3] = pl.stages;
    t->last_plan[4] = pl.depth; t->last_plan[5] = pl.resident | (pl.warps << 4); t->last_plan[6] = (int)g; t->last_plan[7] = (int)pl.smem;
    return GCB_OK;
}

template <bool WT, bool DV, bool RES>
int launch_k(gcb_traj* t, const ClusterParams& cp, const Plan& pl, int* g) {
    constexpr int W8 = kDefaultWarps;
    if constexpr (DV && RES && kDevWarps != kDefaultWarps) {
        if (pl.warps == kDevWarps) {
            switch (pl.kinst) {   // 144 registers hold the template and the sums of up to 8 atoms
                case 1: return launch_inst<1, WT, DV, RES, kDevWarps>(t, cp, pl, g);
                case 2: return launch_inst<2, WT, DV, RES, kDevWarps>(t, cp, pl, g);
                case 3: return launch_inst<3, WT, DV, RES, kDevWarps>(t, cp, pl, g);
                case 4: return launch_inst<4, WT, DV, RES, kDevWarps>(t, cp, pl, g);
                case 5: return launch_inst<5, WT, DV, RES, kDevWarps>(t, cp, pl, g);
                case 6: return launch_inst<6, WT, DV, RES, kDevWarps>(t, cp, pl, g);
                case 7: return launch_inst<7, WT, DV, RES, kDevWarps>(t, cp, pl, g);
                case 8: return launch_inst<8, WT, DV, RES, kDevWarps>(t, cp, pl, g);
            }
            set_error("no 12-warp kernel instance for %d atoms per thread", pl.kinst);
            return GCB_ERR_ARG;
        }
    }
    switch (pl.kinst) {
        case 1: return launch_inst<1, WT, DV, RES, W8>(t, cp, pl, g);
        case 2: return launch_inst<2, WT, DV, RES, W8>(t, cp, pl, g);
        case 3: return launch_inst<3, WT, DV, RES, W8>(t, cp, pl, g);
        case 4: return launch_inst<4, WT, DV, RES, W8>(t, cp, pl, g);
        case 5: return launch_inst<5, WT, DV, RES, W8>(t, cp, pl, g);
        case 6: return launch_inst<6, WT, DV, RES, W8>(t, cp, pl, g);
        case 7: return launch_inst<7, WT, DV, RES, W8>(t, cp, pl, g);
        case 8: return launch_inst<8, WT, DV, RES, W8>(t, cp, pl, g);
        case 10: return launch_inst<10, WT, DV, RES, W8>(t, cp, pl, g);
        case 12: return launch_inst<12, WT, DV, RES, W8>(t, cp, pl, g);
        case 14: return launch_inst<14, WT, DV, RES, W8>(t, cp, pl, g);
    }
    set_error("no kernel instance for %d atoms per thread", pl.kinst);
    return GCB_ERR_ARG;
}

template <bool RES>
int launch_r(gcb_traj* t, const ClusterParams& cp, const Plan& pl, bool weighted, bool want_dev, int* g) {
    if (weighted) return want_dev ? launch_k<true, true, RES>(t, cp, pl, g) : launch_k<true, false, RES>(t, cp, pl, g);
    return want_dev ? launch_k<false, true, RES>(t, cp, pl, g) : launch_k<false, false, RES>(t, cp, pl, g);
}

}  // namespace

#ifdef GCB_TRACE
extern "C" int gcb_trace_read(long long* out) {
    return cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 8 * 512) == cudaSuccess ? 0 : -7;
}
#endif

// heavy: the pass visits every atom twice (write-back, per-atom sums, explicit residuals on request) and keeps the part resident
// in shared memory, which holds fewer atoms per CTA (20 480 per frame against 28 672 for fit-only passes)
bool cluster_kernel_supported(const gcb_traj* t, bool heavy) {
    Plan pl;
    return make_plan(t->tune, t->np, heavy || t->tune.force_explicit, false, &pl);
}

int launch_super_cluster(gcb_traj* t, const SuperParams& p, bool weighted, bool want_dev, int* rows_out) {
    Plan pl;
    const int force_explicit = t->tune.force_explicit;
    if (!make_plan(t->tune, t->np, want_dev || p.frames_rw != nullptr || force_explicit, want_dev, &pl)) { set_error("no cluster plan for %d atoms", t->natoms); return GCB_ERR_ARG; }
    // pass 2's squared-error partials: scratch of THIS handle (its stream orders every use)
    const size_t need = sizeof(double) * (size_t)p.nframes * (size_t)pl.cluster * pl.warps;
    if (need > t->epart_bytes) {
        GCB_CUDA(cudaStreamSynchronize(t->stream));
        cudaFree(t->d_epart);
        t->d_epart = nullptr;
        t->epart_bytes = 0;
        GCB_CUDA(cudaMalloc(&t->d_epart, need));
        t->epart_bytes = need;
    }
    ClusterParams cp;
    cp.p = p;
    cp.cluster = pl.cluster;
    cp.part = pl.part;
    cp.stages = pl.stages;
    cp.depth = pl.depth;
    cp.slots = kMaxSlots;
    cp.nclusters = 0;
    cp.write_back = p.frames_rw != nullptr;
    cp.always_explicit = (cp.write_back || want_dev || force_explicit) ? 1 : 0;
    cp.epart = t->d_epart;
    int g = 0;
    int rc = pl.resident ? launch_r<true>(t, cp, pl, weighted, want_dev, &g) : launch_r<false>(t, cp, pl, weighted, want_dev, &g);
    if (rc) return rc;
    if (rows_out) *rows_out = g;
    if (p.rmsd) {
        finish_rmsd_kernel<<<(unsigned)((p.nframes + 255) / 256), 256, 0, t->stream>>>(cp.epart, pl.cluster, pl.warps, p.nframes, p.inv_n, p.need_explicit, p.rmsd);
        GCB_LAUNCHED();
    }
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

}  // namespace gcb
