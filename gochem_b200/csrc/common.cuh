// common.cuh -- shared declarations of libgochem_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/gochem_b200.h"

namespace gcb {

void set_error(const char* fmt, ...);
extern std::atomic<long> g_launches;

// Everything that steers kernel choice and shape lives in the handle (gcb_traj::tune), never in process globals: two handles
// driven from two threads (goroutines call chem.Super concurrently on distinct matrices, align/lovo.go:290-295, 329) must not see
// each other's settings.  The gcb_set_* entry points only change the defaults a handle copies when it is created.
struct Tuning {
    int variant = GCB_KERNEL_AUTO;
    int split_wb_max = -1, split_dev_max = -1;          // -1 = built-in thresholds (api.cu)
    int cluster = 0, stages = 0, depth = -1, resident = -1;   // cluster-kernel overrides (0 / -1 = built-in choice)
    int force_explicit = 0;                              // 1 = the RMSD always comes from the explicit residual
};
Tuning default_tuning();

#define GCB_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            gcb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return GCB_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define GCB_LAUNCHED() (gcb::g_launches.fetch_add(1, std::memory_order_relaxed))

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Dynamic shared memory above 48 KB is an opt-in per (kernel, device).  It is raised to the device limit once and never lowered,
// so concurrent callers with different launch shapes cannot shrink it under one another.  The set is keyed by the kernel's
// ADDRESS (all instances of a kernel template share one function-pointer type).
constexpr int kMaxOptinSmem = 227 * 1024;
bool smem_opted_in(const void* kern, int dev, bool mark);
template <typename Kern>
static inline cudaError_t opt_in_smem(Kern kern, int dev) {
    const void* key = reinterpret_cast<const void*>(kern);
    if (smem_opted_in(key, dev, false)) return cudaSuccess;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);    // statically declared shared memory counts against the same limit
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxOptinSmem - (int)fa.sharedSizeBytes);
    if (e == cudaSuccess) smem_opted_in(key, dev, true);
    return e;
}

// Atoms per plane are padded to a multiple of 16 doubles (128 B) so every plane of every frame
// starts on a 128-byte boundary: 16-byte TMA bulk copies and double2 loads are always legal.
constexpr int kPlaneAlign = 16;

// Modes of the fused superposition kernels
struct SuperParams {
    const double* frames;   // first frame of the window, SoA planes
    double* frames_rw;      // same pointer when write_back, else nullptr
    long frame_stride;      // doubles between frames = 3 * np
    int np;                 // padded atoms per plane
    int natoms;             // real atoms
    long nframes;
    const double* refc;     // centred reference planes (b - bbar), 3 * np
    const double* weight;   // np multiplicities (nullptr = all ones)
    double inv_n;           // 1 / (sum of weights)
    double n_fit;           // sum of weights
    double bbar[3];         // centroid of the reference fit subset
    double gb;              // sum w |b'|^2 (inner product of the centred template with itself)
    unsigned char* need_explicit;  // per frame: 1 = the RMSD must come from the explicit residual (pass 2)
    double* rmsd;           // per frame (window-relative), may be nullptr
    double* rot;            // 9 per frame
    double* trans1;         // 3 per frame
    double* trans2;         // 3 per frame
    double* devpart;        // [gridDim.x][np] per-CTA per-atom distance sums, nullptr = off
    int* status;            // set non-zero when a frame's correlation matrix is not finite
};

}  // namespace gcb

struct gcb_traj {
    int dev = 0;
    int natoms = 0;
    int np = 0;
    long cap = 0;
    long nframes = 0;
    int sm_count = 0;
    double* d_frames = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    // reference workspace (device)
    double* d_refc = nullptr;    // 3 * np
    double* d_weight = nullptr;  // np
    double* d_refpairs = nullptr;  // gather path: 3 * nidx compact centred template
    int* d_idx = nullptr;          // gather path: test-side indexes
    int idx_cap = 0;
    // per-frame results (cap entries)
    double* d_rmsd = nullptr;
    double* d_rot = nullptr;
    double* d_t1 = nullptr;
    double* d_t2 = nullptr;
    int* d_status = nullptr;
    unsigned char* d_flag = nullptr;  // cap entries, see SuperParams::need_explicit
    // LOVO accumulation
    double* d_devpart = nullptr;
    size_t devpart_rows = 0;
    double* d_devsum = nullptr;
    double* d_refplanes = nullptr;  // split path of large frames (transform.cu): centred template planes of ALL atoms, 3 * np
    int* d_want = nullptr;          // gcb_peratom_dev_sum_atoms: the atoms asked for, and their compact centred template (3 * want_cap)
    double* d_wantref = nullptr;
    int want_cap = 0;
    // per-frame centre of mass / moment tensor (moments.cu)
    double* d_mom_w = nullptr;   // np mass plane
    double* d_mom_c = nullptr;   // 3 per frame
    double* d_mom_m = nullptr;   // 9 per frame
    long mom_cap = 0, mom_frames = 0;
    // all-pairs workspace (allpairs.cu): centred copy, norms, tile list, near-duplicate list
    void* d_ap_ws = nullptr;
    size_t ap_ws_bytes = 0;
    // host pinned scratch for small transfers
    double* h_scratch = nullptr;
    size_t h_scratch_bytes = 0;
    // pinned bounce buffers for the synchronous upload / download of pageable host memory (api.cu), allocated on first use
    void* h_bounce[2] = {nullptr, nullptr};
    size_t bounce_bytes = 0;
    cudaEvent_t bounce_done[2] = {nullptr, nullptr};
    // Ordering between the two copy streams and the compute stream, by frame range (api.cu: deps_*): a kernel waits only for the
    // uploads whose frames it reads, an upload only for the kernels that still use the frames it overwrites -- so the copy of
    // window k + 1 runs under the kernels of window k.
    long up_lo[2] = {0, 0}, up_hi[2] = {0, 0};     // frames written by the last asynchronous upload through each slot
    bool up_pending[2] = {false, false};           // the compute stream has not been made to wait for copy_done[slot] yet
    cudaEvent_t use_done[2] = {nullptr, nullptr};  // recorded on the compute stream after the last call that touched the slot's frames
    bool use_valid[2] = {false, false};
    cudaEvent_t any_done = nullptr;                // ... after calls on frames outside both slots' ranges: [any_lo, any_hi)
    long any_lo = 0, any_hi = 0;
    // per-handle settings and scratch (nothing of a call's state is process-global)
    gcb::Tuning tune;
    int last_plan[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // shape of this handle's last cluster-kernel launch
    double* d_epart = nullptr;                     // cluster kernel: per-warp squared-error partials of pass 2
    size_t epart_bytes = 0;
    double* d_pdist = nullptr;                     // gcb_peratom_dist output scratch
    size_t pdist_cap = 0;
    void* d_lovo_ws = nullptr;                     // gcb_lovo_resident: candidates, sorted ids / values, template copy, state
    size_t lovo_ws_bytes = 0;
    void* h_lovo = nullptr;                        // ... and the 64 bytes the host reads back per iteration (mapped pinned memory)
};

// NCCL communicator wrapper (comm.cu) plus the peer-mapped row blocks of the multi-GPU all-pairs matrix
struct ncclComm;
struct gcb_comm {
    int dev = 0, nranks = 1, rank = 0;
    ncclComm* comm = nullptr;
    cudaStream_t stream = nullptr;
    double* d_send = nullptr;
    double* d_gather = nullptr;
    long cap = 0;
    unsigned char* d_flags = nullptr;     // barrier scratch
    double* peer_blk[8] = {nullptr};      // [rank] own block (cudaMalloc), others: cudaIpcOpenMemHandle mappings
    size_t blk_bytes = 0;
    // mailboxes of the small all-reduce (LOVO's per-atom sums): rank r's box holds [2 parities][nranks][mail_cap] doubles that the
    // peers store into over NVLink, then [2][nranks] sequence flags; mail[r] = rank r's box as seen from here
    double* mail[8] = {nullptr};
    long mail_cap = 0;                    // doubles per (parity, rank) slot
    unsigned long long mail_seq = 0;      // exchanges done so far (the same on every rank: the call is collective)
    int mail_state = 0;                   // 0 = not tried, 1 = usable, -1 = peer mapping failed: NCCL all-gather instead
};

namespace gcb {
// ---- align.LOVO's bookkeeping on the device (lovo_ctl.cu) ----
enum { LOVO_ERR_VALUES = 1,       // a mean that is negative or not a number, or a fit on no atoms: the host mirror takes over
       LOVO_ERR_MINIMUM_N = 2,    // mobileLimit cannot reach MinimumN (the reference loops forever, lovo.go:177-186)
       LOVO_ERR_BOUNDS = 3 };     // n outside the candidate list (a slice panic in the reference)
struct LovoDevState {
    int n, itercount, disagreement, converged, err, dev_status;
    double prevlim, newlim;
    long pad[3];
};
static_assert(sizeof(LovoDevState) == 64, "the host reads this back whole");
struct LovoCtl {
    const double* devsum;   // per-atom sums over all selected frames (all ranks)
    double total;           // frames behind those sums
    const int* cand;        // candidate atoms (ascending), ncand of them
    int ncand;
    int* old;               // indexesold as candidate positions
    int* ids;               // sorted candidate positions of this iteration
    double* vals;           // their means
    LovoDevState* st;       // device copy of the state
    LovoDevState* host;     // mapped pinned copy the host reads
    double less_than;
    int minimum_n;
    const double* ref_aos;  // template, natoms x 3
    int natoms, np;
    double* refc;           // next pass: centred template planes, 3 * np
    double* weight;         // next pass: 0 / 1 per atom
    const int* status;      // the trajectory's status word (SVD failure, exchange timeout)
};
int lovo_ctl_slots();   // the most candidates the device loop takes
int launch_lovo_controller(gcb_traj* t, const LovoCtl& c);       // rank kernel + controller: after a pass
int launch_lovo_first_template(gcb_traj* t, const LovoCtl& c);   // before the first pass
// api.cu: one per-atom-distance pass over the template the controller left on the device (cluster kernel only)
bool peratom_prepared_supported(gcb_traj* t);
int peratom_prepared(gcb_traj* t, long first, long nframes, int n_fit);
int peratom_local(gcb_traj* t, long first, long nframes, const double* ref_aos, const int* idx, int nidx, int write_back);
int peratom_finish(gcb_traj* t);
// frame-range dependencies between uploads and kernels (api.cu)
int deps_before_compute(gcb_traj* t, long first, long nframes);   // the compute stream waits for pending uploads into [first, first+n)
int deps_after_compute(gcb_traj* t, long first, long nframes);    // remembers that queued work uses [first, first+n)
int comm_barrier(gcb_comm* c, cudaStream_t s);
int comm_peer_blocks(gcb_comm* c, size_t bytes, double** blk);
// kernels / launchers implemented across the .cu files
int launch_upload_convert(gcb_traj* t, cudaStream_t s, const void* d_src, long first, long nframes, int layout, double scale);
int launch_download_convert(gcb_traj* t, cudaStream_t s, double* d_dst, long first, long nframes);
int launch_upload_records(gcb_traj* t, cudaStream_t s, const void* d_src, long first, long nframes, long record_bytes, long ox,
                          long oy, long oz, double scale);
int launch_super_generic(gcb_traj* t, const SuperParams& p, bool weighted, bool want_dev, int* grid_out);
int launch_super_gather(gcb_traj* t, const SuperParams& p, const double* refpairs, const int* idx, int nidx);
int launch_rmsd_nofit(gcb_traj* t, const double* frames, long nframes, const double* ref_planes, const double* weight,
                      double n_fit, double* rmsd);
int launch_rmsd_nofit_gather(gcb_traj* t, const double* frames, long nframes, const double* refpairs_raw, const int* idx,
                             int nidx, double* rmsd);
int launch_dev_reduce(gcb_traj* t, int rows, double* out);
int launch_peratom_dist(gcb_traj* t, const double* frame, const double* refpairs, const int* idx, int nidx, double* out);
int launch_synth(gcb_traj* t, long first, long nframes, const double* d_ref_aos, const double* d_sigma,
                 unsigned long long seed, long frame_id0, int frame0_is_ref, int through_f32);
bool cluster_kernel_supported(const gcb_traj* t, bool heavy = false);
int launch_super_cluster(gcb_traj* t, const SuperParams& p, bool weighted, bool want_dev, int* rows_out);
bool small_kernel_supported(const gcb_traj* t, bool heavy);
int small_kernel_rows(const gcb_traj* t);
int launch_super_small(gcb_traj* t, const SuperParams& p, bool masked, bool want_dev, int force_explicit, int* rows_out);
int small_gather_max_atoms();
int launch_transform_frames(gcb_traj* t, const SuperParams& p);
int peratom_accum_rows(const gcb_traj* t, long nframes);
int launch_peratom_accum(gcb_traj* t, const SuperParams& p, const double* ref_planes, int* rows_out);
int peratom_accum_gather_rows(const gcb_traj* t, long nframes, int nwant);
int launch_peratom_accum_gather(gcb_traj* t, const SuperParams& p, const int* d_want, int nwant, const double* d_wantref, int stride,
                                int* rows_out);
int launch_dev_reduce_n(gcb_traj* t, int rows, int stride, int n, double* out);
int launch_super_small_gather(gcb_traj* t, const SuperParams& p, const double* refpairs, const int* idx, int nidx, int force_explicit);
}  // namespace gcb
