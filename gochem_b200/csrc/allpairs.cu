// allpairs.cu -- all-pairs frame RMSD matrix (BASELINE.json configs[3]) as an fp64 tensor-core GEMM.
//
// D_ij = chem.RMSD(X_i, X_j) = ||X_i - X_j||_F / sqrt(n) (/root/reference/geometric.go:232-288: no superposition
// inside RMSD; frames are superposed beforehand with gcb_super_rmsd(write_back = 1)).  Expanding the square,
//     D_ij^2 = (G_ii + G_jj - 2 G_ij) / n,   G = Y Y^T,
// where Y is the F x 3n matrix of frame coordinates MINUS one common structure (the mean of the frames): a common
// offset cancels in every difference, and removing it shrinks G_ii from ~3n*|x|^2 to ~3n*fluctuation^2, which is
// what keeps the cancellation in G_ii + G_jj - 2 G_ij harmless.  Pairs whose D^2 is still below the rounding floor
// of that expression are recomputed from the explicit differences.
//
// The contraction is the one dense GEMM of the path, so it runs on the tensor cores: tcgen05 has no fp64 kind, the
// fp64 tensor path on sm_100a is mma.sync.m8n8k4.f64 (DMMA).  128 x 128 CTA tiles, 8 MMA warps (64 x 32 each) fed by one
// producer warp: 2-D TMA (cp.async.bulk.tensor, 128-byte swizzle) into a ring of shared-memory stages handed over with
// full / empty mbarriers -- no CTA-wide barrier and no address arithmetic in the MMA warps.  Only tiles on or above the
// diagonal are computed and mirrored in the epilogue.
#include <cuda.h>

#include <vector>

#include <stdlib.h>

#include "common.cuh"
#include "kabsch3.cuh"
#include "ptx.cuh"

namespace gcb {

namespace {

constexpr int BM = 128, BN = 128;
constexpr int GEMM_THREADS = 256;
#ifndef GCB_AP_WARPS
#define GCB_AP_WARPS 8
#endif
constexpr int AP_WARPS = GCB_AP_WARPS;            // warps of the no-fit GEMM CTA: (AP_WARPS / 4) x 4 warp tiles
constexpr int AP_THREADS = 32 * AP_WARPS;
constexpr int AP_WR = BM / (AP_WARPS / 4);        // rows of a warp tile: 64 (8 warps) or 32 (16 warps)
constexpr int AP_MI = AP_WR / 8;
constexpr int TLD = BN + 1;           // staging row stride: transposed reads touch every bank once per half warp
constexpr int kMaxPeers = 8;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Y[f][c] = X[f][c] - mu[c] on the packed row layout (ld doubles per frame); norms[f] = sum_c Y[f][c]^2.
__global__ void __launch_bounds__(256) centre_rows_kernel(const double* __restrict__ X, long ldx, const double* __restrict__ mu,
                                                         double* __restrict__ Y, long ldy, int ncols, long nframes,
                                                         double* __restrict__ norms) {
    __shared__ double red[8];
    for (long f = blockIdx.x; f < nframes; f += gridDim.x) {
        const double* x = X + f * ldx;
        double* y = Y + f * ldy;
        double s = 0.0;
        for (int c = threadIdx.x; c < ncols; c += 256) {
            const double v = x[c] - mu[c];
            y[c] = v;
            s = fma(v, v, s);
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += red[w];
            norms[f] = t;
        }
        __syncthreads();
    }
}

// column means of X (nframes x ncols) in two fixed-order steps: sums over kColChunks row chunks (grid.y), then the chunks in order
constexpr int kColChunks = 64;
__global__ void __launch_bounds__(256) col_partial_kernel(const double* __restrict__ X, long ldx, int ncols, long nframes,
                                                         double* __restrict__ partial) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    const long per = (nframes + kColChunks - 1) / kColChunks;
    const long f0 = (long)blockIdx.y * per, f1 = (f0 + per < nframes) ? f0 + per : nframes;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    long f = f0;
    for (; f + 3 < f1; f += 4) {
        s0 += X[f * ldx + c]; s1 += X[(f + 1) * ldx + c]; s2 += X[(f + 2) * ldx + c]; s3 += X[(f + 3) * ldx + c];
    }
    for (; f < f1; f++) s0 += X[f * ldx + c];
    partial[(long)blockIdx.y * ncols + c] = (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(256) col_mean_kernel(const double* __restrict__ partial, int ncols, long nframes, double* __restrict__ mu) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double s = 0.0;
    for (int k = 0; k < kColChunks; k++) s += partial[(long)k * ncols + c];
    mu[c] = s / (double)nframes;
}

// gather a subset of atoms into packed rows [x(idx) | y(idx) | z(idx)], padded to ncols_pad with zeros
__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ frames, long stride, int np,
                                                         const int* __restrict__ idx, int nidx, int ncols_pad, long nframes,
                                                         double* __restrict__ out) {
    const long total = nframes * (long)ncols_pad;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const long f = g / ncols_pad;
        const int c = (int)(g - f * ncols_pad);
        double v = 0.0;
        if (c < 3 * nidx) {
            const int plane = c / nidx, k = c - plane * nidx;
            v = frames[f * stride + (long)plane * np + idx[k]];
        }
        out[g] = v;
    }
}

// gather a subset of atoms keeping the plane layout: out[f][c][k] = frame f plane c atom idx[k], plane stride ncols
__global__ void __launch_bounds__(256) gather_planes_kernel(const double* __restrict__ frames, long stride, int np,
                                                           const int* __restrict__ idx, int nidx, int ncols, long nframes,
                                                           double* __restrict__ out) {
    const long total = nframes * 3L * ncols;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        const long f = g / (3L * ncols);
        const int rem = (int)(g - f * 3L * ncols);
        const int plane = rem / ncols, k = rem - plane * ncols;
        out[g] = k < nidx ? frames[f * stride + (long)plane * np + idx[k]] : 0.0;
    }
}

struct PairList {
    unsigned int* count;
    uint2* pairs;
    unsigned int cap;
};

// Where the rows of the matrix live.  Row i belongs to owner o when row0[o] <= i < row0[o + 1] and is stored at
// blk[o] + (i - row0[o]) * ld.  Single GPU: one owner (the caller's row block).  Multi-GPU: blk[o] is rank o's row block,
// mapped into this process with cudaIpcOpenMemHandle, so a finished tile goes straight to its owners over NVLink.
// Rows outside [row0[0], row0[nowners]) are not stored.
struct PeerOut {
    double* blk[kMaxPeers];
    long row0[kMaxPeers + 1];
    int nowners;
};

__device__ __forceinline__ double* peer_row(const PeerOut& po, long i, long ld) {
    if (i < po.row0[0] || i >= po.row0[po.nowners]) return nullptr;
    int o = 0;
    while (o + 1 < po.nowners && i >= po.row0[o + 1]) o++;
    return po.blk[o] + (i - po.row0[o]) * ld;
}

// ---- the GEMM main loop, TMA-fed ----
// Stage = 16 k of both operands: two boxes of 128 rows x 16 doubles (128 bytes per row, the span of the 128-byte swizzle), 32 KB.
// Inside a stage the 16-byte chunk c of row r sits at chunk position c ^ (r & 7).  The order of k inside a stage is free as long
// as both operands use the same one, so lane t of an MMA quad takes the chunks 2t (k = 4t, 4t + 1: DMMAs 0 and 1 of the stage)
// and 2t + 1 (k = 4t + 2, 4t + 3: DMMAs 2 and 3): one 16-byte load feeds two DMMAs, and the eight lanes of a quarter warp
// (two rows g, four t) land on eight different chunk positions -- conflict-free.
constexpr int TK = 16;                                    // k per stage
constexpr int TSTAGES = 6;
constexpr int kTmaThreads = AP_THREADS;                   // eight MMA warps; lane 0 of warp 0 also feeds the ring (a ninth warp would
                                                          // cost the register file a whole allocation unit: 168 registers per thread)
constexpr size_t kStageBytes = (size_t)(BM + BN) * TK * sizeof(double);   // 32 KB
constexpr size_t kTmaSmem = TSTAGES * kStageBytes + 1024;  // + room to align the ring to 1024 bytes (swizzle atom)
static_assert(64 * TLD * sizeof(double) <= TSTAGES * kStageBytes, "the epilogue staging buffer reuses the stage ring");

// The main loop shared by the GEMM kernels: C (TMA x TMB, this warp's 8 MI x 8 NI part in acc) += A_rows . B_rows^T over kdim,
// operands fetched with the tensor maps (boxes of TMA resp. TMB rows x TK doubles) into ring (1024-byte aligned, TSTAGES stages
// of the two boxes).
// Eight warps as 2 x 4; thread 0 is also the producer.  full / empty: TSTAGES mbarriers each, initialised by the caller
// (full: 1 arrival + bytes, empty: 8 arrivals).  Returns when every warp has consumed every stage (no trailing barrier).
template <int TMA, int TMB, int MI, int NI>
__device__ __forceinline__ void tma_dmma_mainloop(const CUtensorMap* tmapA, const CUtensorMap* tmapB, unsigned char* ring, uint64_t* full,
                                                  uint64_t* empty, int rowA, int rowB, int kdim, double (&acc)[MI][NI][2]) {
    using namespace ptx;
    constexpr uint32_t kBox = (uint32_t)TMA * TK * sizeof(double);      // the A box; B follows it in a stage
    constexpr uint32_t kBoxB = (uint32_t)TMB * TK * sizeof(double);
    constexpr uint32_t kStage = kBox + kBoxB;
    static_assert(kBox % 1024 == 0 && kStage % 1024 == 0, "boxes start on swizzle atoms");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    const int nk = (kdim + TK - 1) / TK;
    // byte offsets of this lane's chunks inside an operand box: row r = base + 8 mi + g (base a multiple of 8, so r & 7 = g)
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t a_row = (uint32_t)(wm * 8 * MI + g) * (TK * 8u);
    const uint32_t b_row = kBox + (uint32_t)(wn * 8 * NI + g) * (TK * 8u);
    const uint32_t ch0 = (uint32_t)((2 * t) ^ g) * 16u, ch1 = (uint32_t)((2 * t + 1) ^ g) * 16u;
    // Not volatile: the scheduler may start a stage's loads early and spread them between the DMMAs.  They cannot move above the
    // wait for the stage (their address comes from the token taken after it) nor below the release (the DMMAs that consume them
    // are volatile and stay in front of the arrive).
    auto lds2 = [](uint32_t addr, double& x, double& y) {
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
    };
    // producer state (thread 0 only): next k-step to load, its stage and the parity of that stage's previous release
    const bool producer = tid == 0;
    int pk = 0, ps = 0;
    uint32_t pph = 1;
    auto issue = [&]() {
        unsigned char* st = ring + (size_t)ps * kStage;
        mbar_arrive_expect_tx(&full[ps], kStage);
        tma_load_2d(st, tmapA, pk * TK, rowA, &full[ps]);
        tma_load_2d(st + kBox, tmapB, pk * TK, rowB, &full[ps]);
        pk++;
        if (++ps == TSTAGES) { ps = 0; pph ^= 1u; }
    };
    if (producer) {
        prefetch_tensormap(tmapA);
        if (tmapB != tmapA) prefetch_tensormap(tmapB);
        while (pk < nk && pk < TSTAGES) issue();     // the ring starts empty
    }
    int s = 0;
    uint32_t ph = 0;
    for (int kt = 0; kt < nk; kt++) {
        if (producer) {
            // refill every stage the eight warps have released (non-blocking test); a stage this warp is about to wait for
            // must have been requested, so that one is waited for
            while (pk < nk && mbar_test(&empty[ps], pph)) issue();
            if (pk <= kt) { mbar_wait(&empty[ps], pph); issue(); }
        }
        mbar_wait(&full[s], ph);
        uint32_t st;
        asm volatile("mov.u32 %0, %1;" : "=r"(st) : "r"(ring_u32 + (uint32_t)s * kStage) : "memory");
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t ch = h ? ch1 : ch0;
            double ax[MI], ay[MI], bx[NI], by[NI];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) lds2(st + b_row + (uint32_t)ni * (8u * TK * 8u) + ch, bx[ni], by[ni]);
#pragma unroll
            for (int mi = 0; mi < MI; mi++) lds2(st + a_row + (uint32_t)mi * (8u * TK * 8u) + ch, ax[mi], ay[mi]);
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], ax[mi], bx[ni]);
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], ay[mi], by[ni]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);   // this warp is done with the stage
        if (++s == TSTAGES) { s = 0; ph ^= 1u; }
    }
}

// One CTA computes the BM x BN block (bi, bj), bj >= bi, of G = Y Y^T and delivers D for the block and for its mirror
// image to the owners of those rows (PeerOut): the block is staged in shared memory, 64 rows at a time, and leaves as
// whole-row segments (256-byte stores per warp), the mirror image read transposed from the same staging buffer.
__global__ void __launch_bounds__(kTmaThreads, 1)
allpairs_dmma_kernel(const __grid_constant__ CUtensorMap tmap, int kdim, long nframes, const double* __restrict__ norms, double inv_n,
                     const int2* __restrict__ tiles, const PeerOut po, long ldd, double floor_rel, PairList redo) {
    using namespace ptx;
    extern __shared__ unsigned char smem_raw_ap[];
    __shared__ uint64_t full[TSTAGES], empty[TSTAGES];
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw_ap) + 1023) & ~(uintptr_t)1023);
    const int2 tile = tiles[blockIdx.x];
    const long i0 = (long)tile.x * BM, j0 = (long)tile.y * BN;   // global frame indexes
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;        // (AP_WARPS / 4) x 4 warps -> AP_WR x 32 warp tiles
    const int g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], AP_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc[AP_MI][4][2];
#pragma unroll
    for (int a = 0; a < AP_MI; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    tma_dmma_mainloop<BM, BN, AP_MI, 4>(&tmap, &tmap, ring, full, empty, (int)i0, (int)j0, kdim, acc);
    __syncthreads();   // every warp is done with the ring (all TMA writes were waited for): reuse it as the staging buffer
    // epilogue: D^2 = (n_i + n_j - 2 g_ij) / n, both triangles
    double* sT = reinterpret_cast<double*>(ring);   // [64][TLD]
    for (int half = 0; half < 2; half++) {
        if ((wm * AP_WR) / 64 == half) {
            const int rbase = wm * AP_WR - half * 64;   // first staged row of this warp
#pragma unroll
            for (int mi = 0; mi < AP_MI; mi++) {
                const long i = i0 + wm * AP_WR + mi * 8 + g;
                const double ni_ = i < nframes ? norms[i] : 0.0;
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int c = wn * 32 + ni * 8 + 2 * t + e;
                        const long j = j0 + c;
                        double d = 0.0;
                        if (i < nframes && j < nframes && i != j) {
                            const double nj_ = norms[j];
                            const double d2 = (ni_ + nj_ - 2.0 * acc[mi][ni][e]) * inv_n;
                            if (d2 < floor_rel * (ni_ + nj_) * inv_n && j > i) {
                                // below the rounding floor of the expansion: recompute from explicit differences later
                                const unsigned int slot = atomicAdd(redo.count, 1u);
                                if (slot < redo.cap) redo.pairs[slot] = make_uint2((unsigned int)i, (unsigned int)j);
                            }
                            d = sqrt(d2 > 0.0 ? d2 : 0.0);
                        }
                        sT[(rbase + mi * 8 + g) * TLD + c] = d;
                    }
                }
            }
        }
        __syncthreads();
        const long ib = i0 + half * 64;   // first row staged
        {
            for (int r = warp; r < 64; r += AP_WARPS) {          // rows of the block: 128 contiguous doubles each
                const long i = ib + r;
                if (i >= nframes) break;
                double* dst = peer_row(po, i, ldd);
                if (!dst) continue;
                dst += j0;
#pragma unroll
                for (int c = lane; c < BN; c += 32)
                    if (j0 + c < nframes) dst[c] = sT[r * TLD + c];
            }
            if (tile.x != tile.y) {                       // mirror image: rows j of the matrix, 64 contiguous doubles each
                for (int c = warp; c < BN; c += AP_WARPS) {
                    const long j = j0 + c;
                    if (j >= nframes) break;
                    double* dst = peer_row(po, j, ldd);
                    if (!dst) continue;
                    dst += ib;
#pragma unroll
                    for (int r = lane; r < 64; r += 32)
                        if (ib + r < nframes) dst[r] = sT[r * TLD + c];
                }
            }
        }
        __syncthreads();
    }
}

// The same block computation for HALF a block: rows [64 hx, 64 hx + 64) x columns [128 by, 128 by + 128), eight warps of 32 x 32.
// A rank's share of the blocks rarely fills its last wave of CTAs (1550 blocks on 148 SMs: the eleventh wave is 70 blocks); cutting
// those blocks in two makes it one wave of half the length.  Every element still adds the same products in the same order
// (the k order inside a stage depends on the lane's place in its MMA quad only), so the matrix does not change by a bit.
constexpr int HM = 64;
constexpr size_t kHalfStageBytes = (size_t)(HM + BN) * TK * sizeof(double);   // 24 KB
constexpr size_t kHalfSmem = TSTAGES * kHalfStageBytes + 1024;
static_assert(HM * TLD * sizeof(double) <= TSTAGES * kHalfStageBytes, "the epilogue staging buffer reuses the stage ring");

__global__ void __launch_bounds__(kTmaThreads, 1)
allpairs_dmma_half_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, int kdim, long nframes,
                          const double* __restrict__ norms, double inv_n, const int2* __restrict__ tiles, const PeerOut po, long ldd,
                          double floor_rel, PairList redo) {
    using namespace ptx;
    extern __shared__ unsigned char smem_raw_ap[];
    __shared__ uint64_t full[TSTAGES], empty[TSTAGES];
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw_ap) + 1023) & ~(uintptr_t)1023);
    if (AP_WARPS != 8) return;                      // written for eight warps of 32 x 32; other builds never launch it
    const int2 tile = tiles[blockIdx.x];            // x: half-block row (64 frames), y: block column (128 frames)
    const long i0 = (long)tile.x * HM, j0 = (long)tile.y * BN;
    const bool diagonal = (tile.x >> 1) == tile.y;  // the two halves of a diagonal block cover each other's mirror image
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;        // 2 x 4 warps -> 32 x 32 warp tiles
    const int g = lane >> 2, t = lane & 3;
    if (tid == 0) {
        for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], AP_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    tma_dmma_mainloop<HM, BN, 4, 4>(&tmapA, &tmapB, ring, full, empty, (int)i0, (int)j0, kdim, acc);
    __syncthreads();
    double* sT = reinterpret_cast<double*>(ring);   // [64][TLD]
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const long i = i0 + wm * 32 + mi * 8 + g;
        const double ni_ = i < nframes ? norms[i] : 0.0;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int c = wn * 32 + ni * 8 + 2 * t + e;
                const long j = j0 + c;
                double d = 0.0;
                if (i < nframes && j < nframes && i != j) {
                    const double nj_ = norms[j];
                    const double d2 = (ni_ + nj_ - 2.0 * acc[mi][ni][e]) * inv_n;
                    if (d2 < floor_rel * (ni_ + nj_) * inv_n && j > i) {
                        const unsigned int slot = atomicAdd(redo.count, 1u);
                        if (slot < redo.cap) redo.pairs[slot] = make_uint2((unsigned int)i, (unsigned int)j);
                    }
                    d = sqrt(d2 > 0.0 ? d2 : 0.0);
                }
                sT[(wm * 32 + mi * 8 + g) * TLD + c] = d;
            }
        }
    }
    __syncthreads();
    for (int r = warp; r < HM; r += AP_WARPS) {          // rows of the half block: 128 contiguous doubles each
        const long i = i0 + r;
        if (i >= nframes) break;
        double* dst = peer_row(po, i, ldd);
        if (!dst) continue;
        dst += j0;
#pragma unroll
        for (int c = lane; c < BN; c += 32)
            if (j0 + c < nframes) dst[c] = sT[r * TLD + c];
    }
    if (!diagonal) {                                      // mirror image: rows j of the matrix, 64 contiguous doubles each
        for (int c = warp; c < BN; c += AP_WARPS) {
            const long j = j0 + c;
            if (j >= nframes) break;
            double* dst = peer_row(po, j, ldd);
            if (!dst) continue;
            dst += i0;
#pragma unroll
            for (int r = lane; r < HM; r += 32)
                if (i0 + r < nframes) dst[r] = sT[r * TLD + c];
        }
    }
}

// explicit ||Y_i - Y_j|| for the flagged pairs: one warp per pair
__global__ void __launch_bounds__(256) redo_pairs_kernel(const double* __restrict__ Y, long ld, int kdim, double inv_n, PairList redo,
                                                        const PeerOut po, long ldd) {
    const unsigned int n = min(*redo.count, redo.cap);
    const int lane = threadIdx.x & 31;
    for (unsigned int p = blockIdx.x * 8 + (threadIdx.x >> 5); p < n; p += gridDim.x * 8) {
        const uint2 ij = redo.pairs[p];
        const double* a = Y + (long)ij.x * ld;
        const double* b = Y + (long)ij.y * ld;
        double s = 0.0;
        for (int c = lane; c < kdim; c += 32) {
            const double v = a[c] - b[c];
            s = fma(v, v, s);
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const double d = sqrt(s * inv_n);
            double* ri = peer_row(po, (long)ij.x, ldd);
            double* rj = peer_row(po, (long)ij.y, ldd);
            if (ri) ri[ij.y] = d;
            if (rj) rj[ij.x] = d;
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Variant (ii): RMSD after the optimal superposition of every pair.  With every frame centred on its own centroid,
// the 3x3 correlation of frames i and j is H_ij = Y_i^T Y_j, i.e. the (3i..3i+2, 3j..3j+2) block of Z Z^T where Z is
// the (3F) x n matrix whose rows are the coordinate planes -- exactly how frames are stored.  The minimum over
// rotations of ||Y_i R - Y_j||^2 is G_i + G_j - 2 tr(R^T H_ij) with R from the same 3x3 solve as everywhere else
// (kabsch3.cuh; geometric.go:154-190), evaluated per pair in the epilogue.
// ------------------------------------------------------------------------------------------------------------------
constexpr int FT = 32;              // frames per tile side
constexpr int FM = 3 * FT;          // 96 rows
constexpr int FLDC = FM + 1;        // padded row of the staged C tile

// rows = planes: Zc[(3f+c)][k] = frame f plane c minus that plane's mean over the used atoms; norms[f] = sum |.|^2
__global__ void __launch_bounds__(256) centre_frames_kernel(const double* __restrict__ X, long ldrow, int ncols_used, int ncols,
                                                           long nframes, double* __restrict__ Z, double* __restrict__ norms) {
    __shared__ double red[8][4];
    __shared__ double cen[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long f = blockIdx.x; f < nframes; f += gridDim.x) {
        const double* x = X + f * 3 * ldrow;
        double s[3] = {0, 0, 0};
        for (int k = threadIdx.x; k < ncols_used; k += 256) { s[0] += x[k]; s[1] += x[ldrow + k]; s[2] += x[2 * ldrow + k]; }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_down_sync(0xffffffffu, s[c], o);
            if (lane == 0) red[warp][c] = s[c];
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += red[w][threadIdx.x];
            cen[threadIdx.x] = t / (double)ncols_used;
        }
        __syncthreads();
        double* z = Z + f * 3 * (long)ncols;
        double g = 0.0;
        for (int k = threadIdx.x; k < ncols; k += 256) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double v = k < ncols_used ? x[c * ldrow + k] - cen[c] : 0.0;
                z[(long)c * ncols + k] = v;
                g = fma(v, v, g);
            }
        }
        for (int o = 16; o > 0; o >>= 1) g += __shfl_down_sync(0xffffffffu, g, o);
        if (lane == 0) red[warp][3] = g;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += red[w][3];
            norms[f] = t;
        }
        __syncthreads();
    }
}

constexpr size_t kFitStageBytes = (size_t)2 * FM * TK * sizeof(double);   // two boxes of 96 rows x 16 doubles
constexpr size_t kFitSmem = TSTAGES * kFitStageBytes + 1024;
static_assert((size_t)FM * FLDC * sizeof(double) <= TSTAGES * kFitStageBytes, "the staged C tile reuses the stage ring");

__global__ void __launch_bounds__(GEMM_THREADS, 1)
allpairs_fit_dmma_kernel(const __grid_constant__ CUtensorMap tmap, int kdim, long nframes, const double* __restrict__ norms, double inv_n,
                         const int2* __restrict__ tiles, const PeerOut po, long ldd, double floor_rel, PairList redo) {
    using namespace ptx;
    extern __shared__ unsigned char smem_raw_fit[];
    __shared__ uint64_t full[TSTAGES], empty[TSTAGES];
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw_fit) + 1023) & ~(uintptr_t)1023);
    const int2 tile = tiles[blockIdx.x];
    const long fi0 = (long)tile.x * FT, fj0 = (long)tile.y * FT;   // first frames of the tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;         // 2 x 4 warps -> 48 x 24 warp tiles
    const int g = lane >> 2, t = lane & 3;
    if (tid == 0) {
        for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], GEMM_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc[6][3][2];
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    // rows of Z are coordinate planes: frame f owns rows 3f .. 3f+2; rows beyond 3 * nframes arrive as zeros
    tma_dmma_mainloop<FM, FM, 6, 3>(&tmap, &tmap, ring, full, empty, (int)(3 * fi0), (int)(3 * fj0), kdim, acc);
    __syncthreads();   // everyone is done with the ring: reuse it for the C tile
    double* smem = reinterpret_cast<double*>(ring);
    double* sC = smem;  // [FM][FLDC]
#pragma unroll
    for (int mi = 0; mi < 6; mi++)
#pragma unroll
        for (int ni = 0; ni < 3; ni++) {
            const int r = wm * 48 + mi * 8 + g, c = wn * 24 + ni * 8 + 2 * t;
            sC[r * FLDC + c] = acc[mi][ni][0];
            sC[r * FLDC + c + 1] = acc[mi][ni][1];
        }
    __syncthreads();
    // one 3x3 problem per frame pair of the tile: 1024 pairs, 4 per thread
    for (int pidx = tid; pidx < FT * FT; pidx += GEMM_THREADS) {
        const int pi = pidx / FT, pj = pidx % FT;
        const long i = fi0 + pi, j = fj0 + pj;
        if (i >= nframes || j >= nframes) continue;
        if (tile.x == tile.y && j < i) continue;   // the diagonal tile holds both orders; keep one
        double d;
        if (i == j) {
            d = 0.0;
        } else {
            double h[9], r[9];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) h[3 * a + b] = sC[(3 * pi + a) * FLDC + 3 * pj + b];
            if (k3::kabsch3(h, r) != 0) {
#pragma unroll
                for (int k = 0; k < 9; k++) r[k] = (k % 4 == 0) ? 1.0 : 0.0;
            }
            double tr = 0.0;
#pragma unroll
            for (int k = 0; k < 9; k++) tr = fma(r[k], h[k], tr);
            const double gi = norms[i], gj = norms[j];
            const double d2 = (gi + gj - 2.0 * tr) * inv_n;
            if (d2 < floor_rel * (gi + gj) * inv_n) {
                const unsigned int slot = atomicAdd(redo.count, 1u);
                if (slot < redo.cap) redo.pairs[slot] = make_uint2((unsigned int)i, (unsigned int)j);
            }
            d = sqrt(d2 > 0.0 ? d2 : 0.0);
        }
        double* ri = peer_row(po, i, ldd);
        if (ri) ri[j] = d;
        if (j != i) {
            double* rj = peer_row(po, j, ldd);
            if (rj) rj[i] = d;
        }
    }
}

// flagged pairs again, from explicit sums: H by direct dot products, R, then sum |y_i R - y_j|^2 -- one warp per pair
__global__ void __launch_bounds__(256) redo_fit_pairs_kernel(const double* __restrict__ Z, int kdim, double inv_n, PairList redo,
                                                            const PeerOut po, long ldd) {
    const unsigned int n = min(*redo.count, redo.cap);
    const int lane = threadIdx.x & 31;
    for (unsigned int p = blockIdx.x * 8 + (threadIdx.x >> 5); p < n; p += gridDim.x * 8) {
        const uint2 ij = redo.pairs[p];
        const double* a = Z + (long)ij.x * 3 * kdim;
        const double* b = Z + (long)ij.y * 3 * kdim;
        double h[9];
#pragma unroll
        for (int k = 0; k < 9; k++) h[k] = 0.0;
        for (int c = lane; c < kdim; c += 32) {
            const double ax = a[c], ay = a[kdim + c], az = a[2L * kdim + c];
            const double bx = b[c], by = b[kdim + c], bz = b[2L * kdim + c];
            h[0] = fma(ax, bx, h[0]); h[1] = fma(ax, by, h[1]); h[2] = fma(ax, bz, h[2]);
            h[3] = fma(ay, bx, h[3]); h[4] = fma(ay, by, h[4]); h[5] = fma(ay, bz, h[5]);
            h[6] = fma(az, bx, h[6]); h[7] = fma(az, by, h[7]); h[8] = fma(az, bz, h[8]);
        }
#pragma unroll
        for (int k = 0; k < 9; k++)
            for (int o = 16; o > 0; o >>= 1) h[k] += __shfl_xor_sync(0xffffffffu, h[k], o);
        double r[9];
        if (k3::kabsch3(h, r) != 0) {
#pragma unroll
            for (int k = 0; k < 9; k++) r[k] = (k % 4 == 0) ? 1.0 : 0.0;
        }
        double s = 0.0;
        for (int c = lane; c < kdim; c += 32) {
            const double ax = a[c], ay = a[kdim + c], az = a[2L * kdim + c];
            const double dx = fma(az, r[6], fma(ay, r[3], ax * r[0])) - b[c];
            const double dy = fma(az, r[7], fma(ay, r[4], ax * r[1])) - b[kdim + c];
            const double dz = fma(az, r[8], fma(ay, r[5], ax * r[2])) - b[2L * kdim + c];
            s = fma(dz, dz, fma(dy, dy, fma(dx, dx, s)));
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const double d = sqrt(s * inv_n);
            double* ri = peer_row(po, (long)ij.x, ldd);
            double* rj = peer_row(po, (long)ij.y, ldd);
            if (ri) ri[ij.y] = d;
            if (rj) rj[ij.x] = d;
        }
    }
}

}  // namespace

// Device-side driver.  Y source: the trajectory planes themselves (idx == nullptr: rows of 3*np doubles, padding
// zeros) or a gathered subset.  The F x F matrix (ld = nframes) goes to the row owners in `po`.
//   deal_n == 1: single-GPU call -- every tile on or above the diagonal that touches the owner's rows, directly or
//                through its mirror image, is computed here.
//   deal_n  > 1: the tiles on or above the diagonal are dealt round-robin to deal_n ranks (this one is deal_rank) and
//                each finished tile is stored into the row blocks of its owners, local or peer.
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links the runtime only)
static int make_tensor_map(CUtensorMap* out, const double* base, unsigned long long cols, unsigned long long rows, unsigned long long pitch_bytes,
                           int box_rows = BM) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return GCB_ERR_CUDA;
        }
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for %llu x %llu doubles, pitch %llu", (int)r, rows, cols, pitch_bytes);
        return GCB_ERR_CUDA;
    }
    return GCB_OK;
}

static int allpairs_device(gcb_traj* t, long first, long nframes, const int* d_idx, int nidx, const PeerOut& po, int deal_rank,
                           int deal_n, float* gemm_ms, long* redone, bool fit) {
    cudaStream_t s = t->stream;
    const int natoms_used = d_idx ? nidx : t->natoms;
    // (i): one row per frame, 3n columns.  (ii): three rows (planes) per frame, n columns.
    const int kdim = fit ? (d_idx ? round_up(nidx, 2) : t->np) : (d_idx ? round_up(3 * nidx, 2) : 3 * t->np);
    const long ld = kdim;
    double *Y = nullptr, *mu = nullptr, *norms = nullptr, *X = nullptr, *colpart = nullptr;
    int2* d_tiles = nullptr;
    PairList redo{nullptr, nullptr, 1u << 22};
    int rc = GCB_OK;
    std::vector<int2> tiles;
    const int tile_frames = fit ? FT : BM;
    const int nt = (int)((nframes + tile_frames - 1) / tile_frames);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    size_t nfull = 0, nhalf = 0;   // blocks of the GEMM launch, half blocks of its split last wave
#define AP_TRY(call)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            set_error("%s failed: %s", #call, cudaGetErrorString(e__));                           \
            rc = (e__ == cudaErrorMemoryAllocation) ? GCB_ERR_ALLOC : GCB_ERR_CUDA;               \
            goto done;                                                                            \
        }                                                                                         \
    } while (0)
    {
        // one workspace per trajectory handle, grown on demand and kept: no cudaMalloc / cudaFree (both synchronise the
        // device) inside repeated calls
        const size_t ybytes = sizeof(double) * (size_t)nframes * (size_t)(fit ? 3 * ld : ld);
        const size_t ntile_max = (size_t)nt * ((size_t)nt + 1) / 2 + 1 + (size_t)t->sm_count;   // + the halves of a split last wave
        size_t off = 0;
        auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
        const size_t oY = carve(ybytes), oX = carve(d_idx ? ybytes : 0), oMu = carve(sizeof(double) * (size_t)kdim),
                     oPart = carve(sizeof(double) * (size_t)kdim * kColChunks), oNorm = carve(sizeof(double) * (size_t)nframes),
                     oCnt = carve(sizeof(unsigned int)), oPairs = carve(sizeof(uint2) * (size_t)redo.cap),
                     oTiles = carve(sizeof(int2) * ntile_max);
        if (off > t->ap_ws_bytes) {
            AP_TRY(cudaStreamSynchronize(s));
            cudaFree(t->d_ap_ws);
            t->d_ap_ws = nullptr;
            t->ap_ws_bytes = 0;
            AP_TRY(cudaMalloc(&t->d_ap_ws, off));
            t->ap_ws_bytes = off;
        }
        unsigned char* w = (unsigned char*)t->d_ap_ws;
        Y = (double*)(w + oY); X = (double*)(w + oX); mu = (double*)(w + oMu); colpart = (double*)(w + oPart);
        norms = (double*)(w + oNorm); redo.count = (unsigned int*)(w + oCnt); redo.pairs = (uint2*)(w + oPairs);
        d_tiles = (int2*)(w + oTiles);
    }
    AP_TRY(cudaMemsetAsync(redo.count, 0, sizeof(unsigned int), s));
    if ((rc = deps_before_compute(t, first, nframes)) != GCB_OK) goto done;
    {
        const double* src = t->d_frames + first * 3L * t->np;
        long ldx = 3L * t->np;
        if (fit) {
            long ldrow = t->np;   // plane stride inside a frame
            int used = t->natoms;
            if (d_idx) {
                // gather the subset into packed planes [x(idx) | y(idx) | z(idx)] with plane stride kdim
                gather_planes_kernel<<<t->sm_count * 8, 256, 0, s>>>(src, 3L * t->np, t->np, d_idx, nidx, kdim, nframes, X);
                GCB_LAUNCHED();
                src = X;
                ldrow = kdim;
                used = nidx;
            }
            centre_frames_kernel<<<t->sm_count * 4, 256, 0, s>>>(src, ldrow, used, kdim, nframes, Y, norms);
            GCB_LAUNCHED();
        } else {
            if (d_idx) {
                gather_rows_kernel<<<t->sm_count * 8, 256, 0, s>>>(src, 3L * t->np, t->np, d_idx, nidx, kdim, nframes, X);
                GCB_LAUNCHED();
                src = X;
                ldx = ld;
            }
            col_partial_kernel<<<dim3((kdim + 255) / 256, kColChunks), 256, 0, s>>>(src, ldx, kdim, nframes, colpart);
            GCB_LAUNCHED();
            col_mean_kernel<<<(kdim + 255) / 256, 256, 0, s>>>(colpart, kdim, nframes, mu);
            GCB_LAUNCHED();
            centre_rows_kernel<<<t->sm_count * 4, 256, 0, s>>>(src, ldx, mu, Y, ld, kdim, nframes, norms);
            GCB_LAUNCHED();
        }
    }
    {
        const long row0 = po.row0[0], row1 = po.row0[po.nowners];
        long k = 0;
        for (int bi = 0; bi < nt; bi++)
            for (int bj = bi; bj < nt; bj++, k++) {
                if (deal_n > 1) {
                    if (k % deal_n == deal_rank) tiles.push_back(make_int2(bi, bj));
                    continue;
                }
                const long ia = (long)bi * tile_frames, ib = ia + tile_frames, ja = (long)bj * tile_frames, jb = ja + tile_frames;
                const bool rows_hit = ia < row1 && ib > row0;
                const bool mirror_hit = ja < row1 && jb > row0;
                if (rows_hit || mirror_hit) tiles.push_back(make_int2(bi, bj));
            }
    }
    if (tiles.empty()) tiles.push_back(make_int2(nt, nt));   // a block outside the matrix: nothing is stored
    // A last wave that fills at most half of the SMs is computed as half blocks (allpairs_dmma_half_kernel): one wave of half the
    // length instead of a whole one.  The full blocks come first in the list, the halves after them.
    nfull = tiles.size();
    nhalf = 0;
    if (!fit && AP_WARPS == 8 && !getenv("GCB_AP_NO_SPLIT")) {   // the variable is for tests: both ways give the same bits
        const size_t sms = (size_t)t->sm_count, rem = tiles.size() % sms;
        if (tiles.size() > sms && rem > 0 && 2 * rem <= sms) {
            nfull = tiles.size() - rem;
            std::vector<int2> halves;
            for (size_t k = nfull; k < tiles.size(); k++)
                for (int h = 0; h < 2; h++)
                    if ((long)(2 * tiles[k].x + h) * HM < nframes) halves.push_back(make_int2(2 * tiles[k].x + h, tiles[k].y));
            tiles.resize(nfull);
            tiles.insert(tiles.end(), halves.begin(), halves.end());
            nhalf = halves.size();
        }
    }
    AP_TRY(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, s));
    AP_TRY(cudaEventCreate(&e0));
    AP_TRY(cudaEventCreate(&e1));
    if (fit) {
        // tensor map of Z: 3 * nframes rows (coordinate planes) of kdim doubles, boxes of 96 rows x 16 doubles
        CUtensorMap tmap;
        rc = make_tensor_map(&tmap, Y, (unsigned long long)kdim, 3ull * (unsigned long long)nframes, (unsigned long long)kdim * sizeof(double), FM);
        if (rc) goto done;
        AP_TRY(opt_in_smem(allpairs_fit_dmma_kernel, t->dev));
        AP_TRY(cudaEventRecord(e0, s));
        allpairs_fit_dmma_kernel<<<(unsigned)tiles.size(), GEMM_THREADS, kFitSmem, s>>>(tmap, kdim, nframes, norms, 1.0 / (double)natoms_used,
                                                                                      d_tiles, po, nframes, 1e-7, redo);
        GCB_LAUNCHED();
        AP_TRY(cudaEventRecord(e1, s));
        redo_fit_pairs_kernel<<<t->sm_count * 4, 256, 0, s>>>(Y, kdim, 1.0 / (double)natoms_used, redo, po, nframes);
        GCB_LAUNCHED();
    } else {
        // tensor map of Y: nframes rows of kdim doubles (row pitch ld), boxes of 128 rows x 16 doubles, 128-byte swizzle,
        // zeros beyond either extent (the tails of k and of the frames need no predicates in the kernel)
        CUtensorMap tmap;
        rc = make_tensor_map(&tmap, Y, (unsigned long long)kdim, (unsigned long long)nframes, (unsigned long long)ld * sizeof(double));
        if (rc) goto done;
        AP_TRY(opt_in_smem(allpairs_dmma_kernel, t->dev));
        CUtensorMap tmap_half;
        if (nhalf) {
            rc = make_tensor_map(&tmap_half, Y, (unsigned long long)kdim, (unsigned long long)nframes, (unsigned long long)ld * sizeof(double), HM);
            if (rc) goto done;
            AP_TRY(opt_in_smem(allpairs_dmma_half_kernel, t->dev));
        }
        AP_TRY(cudaEventRecord(e0, s));
        allpairs_dmma_kernel<<<(unsigned)nfull, kTmaThreads, kTmaSmem, s>>>(tmap, kdim, nframes, norms, 1.0 / (double)natoms_used,
                                                                            d_tiles, po, nframes, 1e-7, redo);
        GCB_LAUNCHED();
        if (nhalf) {
            allpairs_dmma_half_kernel<<<(unsigned)nhalf, kTmaThreads, kHalfSmem, s>>>(tmap_half, tmap, kdim, nframes, norms, 1.0 / (double)natoms_used,
                                                                                      d_tiles + nfull, po, nframes, 1e-7, redo);
            GCB_LAUNCHED();
        }
        AP_TRY(cudaEventRecord(e1, s));
        redo_pairs_kernel<<<t->sm_count * 4, 256, 0, s>>>(Y, ld, kdim, 1.0 / (double)natoms_used, redo, po, nframes);
        GCB_LAUNCHED();
    }
    AP_TRY(cudaStreamSynchronize(s));
    AP_TRY(cudaGetLastError());
    if (gemm_ms) AP_TRY(cudaEventElapsedTime(gemm_ms, e0, e1));
    if (redone) {
        unsigned int c = 0;
        AP_TRY(cudaMemcpy(&c, redo.count, sizeof c, cudaMemcpyDeviceToHost));
        *redone = c;
        if (c > redo.cap) {
            set_error("all-pairs: %u near-duplicate pairs exceed the explicit-recompute list (%u)", c, redo.cap);
            rc = GCB_ERR_ARG;
        }
    }
done:
#undef AP_TRY
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

}  // namespace gcb

static int upload_idx(gcb_traj* t, const int* idx, int nidx, int** d_idx) {
    using namespace gcb;
    *d_idx = nullptr;
    if (!(idx && nidx > 0)) return GCB_OK;
    for (int k = 0; k < nidx; k++)
        if (idx[k] < 0 || idx[k] >= t->natoms) { set_error("atom index %d out of range", idx[k]); return GCB_ERR_INDEXES; }
    GCB_CUDA(cudaMalloc(d_idx, sizeof(int) * (size_t)nidx));
    GCB_CUDA(cudaMemcpy(*d_idx, idx, sizeof(int) * (size_t)nidx, cudaMemcpyHostToDevice));
    return GCB_OK;
}

static int allpairs_entry(gcb_traj* t, long first, long nframes, const int* idx, int nidx, long row0, long nrows, double* out_host,
                          float* gemm_ms, long* redone, bool fit) {
    using namespace gcb;
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    if (first < 0 || nframes <= 0 || first + nframes > t->nframes) { set_error("frame window outside the trajectory"); return GCB_ERR_ARG; }
    if (row0 < 0 || nrows <= 0 || row0 + nrows > nframes || !out_host) { set_error("bad row block"); return GCB_ERR_ARG; }
    if (nframes > 0x3fffffffL) { set_error("too many frames"); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    int* d_idx = nullptr;
    int rc = upload_idx(t, idx, nidx, &d_idx);
    if (rc) return rc;
    double* d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_out, sizeof(double) * (size_t)nrows * (size_t)nframes);
    if (e != cudaSuccess) { cudaFree(d_idx); set_error("cudaMalloc of the %ld x %ld output block failed", nrows, nframes); return GCB_ERR_ALLOC; }
    PeerOut po = {};
    po.blk[0] = d_out; po.row0[0] = row0; po.row0[1] = row0 + nrows; po.nowners = 1;
    rc = allpairs_device(t, first, nframes, d_idx, nidx, po, 0, 1, gemm_ms, redone, fit);
    if (rc == GCB_OK) {
        e = cudaMemcpy(out_host, d_out, sizeof(double) * (size_t)nrows * (size_t)nframes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("copying the RMSD matrix back failed: %s", cudaGetErrorString(e)); rc = GCB_ERR_CUDA; }
    }
    cudaFree(d_out);
    cudaFree(d_idx);
    return rc;
}

extern "C" int gcb_allpairs_rmsd(gcb_traj* t, long first, long nframes, const int* idx, int nidx, long row0, long nrows,
                                 double* out_host, float* gemm_ms, long* redone) {
    return allpairs_entry(t, first, nframes, idx, nidx, row0, nrows, out_host, gemm_ms, redone, false);
}

extern "C" int gcb_allpairs_rmsd_fit(gcb_traj* t, long first, long nframes, const int* idx, int nidx, long row0, long nrows,
                                     double* out_host, float* gemm_ms, long* redone) {
    return allpairs_entry(t, first, nframes, idx, nidx, row0, nrows, out_host, gemm_ms, redone, true);
}

// Multi-GPU matrix: see include/gochem_b200.h.  Rank r owns rows [r*base + min(r, rem), ...) (base = F / nranks); the
// upper-triangle tiles are dealt round-robin, so every rank does the same share of the GEMM whatever rows it owns, and
// the epilogue of each tile stores it -- and its mirror image -- into the owners' row blocks through the peer mappings
// of comm (cudaIpc over NVLink).  A stream synchronisation plus a barrier across the ranks completes every block.
extern "C" int gcb_allpairs_rmsd_sharded(gcb_traj* t, long first, long nframes, const int* idx, int nidx, int fit, gcb_comm* comm,
                                         double* out_host_rows, long* row0_out, long* nrows_out, float* gemm_ms, float* total_ms,
                                         long* redone) {
    using namespace gcb;
    if (!t || !comm) { set_error("NULL trajectory or communicator"); return GCB_ERR_ARG; }
    if (first < 0 || nframes <= 0 || first + nframes > t->nframes) { set_error("frame window outside the trajectory"); return GCB_ERR_ARG; }
    if (nframes > 0x3fffffffL) { set_error("too many frames"); return GCB_ERR_ARG; }
    if (comm->dev != t->dev) { set_error("communicator and trajectory live on different devices"); return GCB_ERR_ARG; }
    if (comm->nranks > kMaxPeers) { set_error("at most %d ranks", kMaxPeers); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    const int n = comm->nranks;
    const long base = nframes / n, rem = nframes % n;
    PeerOut po = {};
    po.nowners = n;
    for (int r = 0; r <= n; r++) po.row0[r] = (long)r * base + (r < rem ? r : rem);
    const size_t blk_bytes = sizeof(double) * (size_t)(base + (rem ? 1 : 0)) * (size_t)nframes;
    double* blk[kMaxPeers] = {nullptr};
    int rc = comm_peer_blocks(comm, blk_bytes, blk);   // collective
    if (rc) return rc;
    for (int r = 0; r < n; r++) po.blk[r] = blk[r];
    int* d_idx = nullptr;
    rc = upload_idx(t, idx, nidx, &d_idx);
    if (rc) return rc;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    GCB_CUDA(cudaEventCreate(&e0));
    GCB_CUDA(cudaEventCreate(&e1));
    rc = comm_barrier(comm, t->stream);                  // nobody stores into a block its owner may still be reading
    if (rc == GCB_OK) {
        cudaEventRecord(e0, t->stream);
        rc = allpairs_device(t, first, nframes, d_idx, nidx, po, comm->rank, n, gemm_ms, redone, fit != 0);
    }
    if (rc == GCB_OK) rc = comm_barrier(comm, t->stream);   // every rank's kernels have completed: all blocks are whole
    if (rc == GCB_OK) {
        cudaEventRecord(e1, t->stream);
        if (cudaStreamSynchronize(t->stream) != cudaSuccess) { set_error("all-pairs: stream synchronisation failed"); rc = GCB_ERR_CUDA; }
    }
    if (rc == GCB_OK && total_ms) cudaEventElapsedTime(total_ms, e0, e1);
    const long my0 = po.row0[comm->rank], myn = po.row0[comm->rank + 1] - my0;
    if (row0_out) *row0_out = my0;
    if (nrows_out) *nrows_out = myn;
    if (rc == GCB_OK && out_host_rows && myn > 0) {
        cudaError_t e = cudaMemcpy(out_host_rows, blk[comm->rank], sizeof(double) * (size_t)myn * (size_t)nframes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("copying the RMSD row block back failed: %s", cudaGetErrorString(e)); rc = GCB_ERR_CUDA; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_idx);
    return rc;
}
