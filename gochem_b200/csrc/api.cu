// api.cu -- the extern "C" boundary declared in include/gochem_b200.h.
//
// Host-side work done here is only what has to happen once per call: argument checks that mirror
// the reference's error returns, preparing the centred template (the reference centres templa in
// chem.MassCenter, /root/reference/geometric.go:670-703, on every call; here it is centred once per
// call and reused by every frame), staging copies, launches.  All arithmetic on frames runs in the
// CUDA kernels; there is no CPU fallback.
#include <sched.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <set>
#include <thread>
#include <vector>

#include "common.cuh"

namespace gcb {

static thread_local std::string t_error;
std::atomic<long> g_launches{0};
// Small structures, passes that visit every atom twice: up to these atom counts the fit-only launch plus one elementwise sweep
// (transform.cu) beats the fused kernels, whose second sweep misses L2 there (gcb_set_split_thresholds; measured defaults)
// 100 000 frames, tools/split_small_probe.py, profiles/r1_s5_split_small_probe.jsonl: write-back 300 atoms 510 -> 392 us, 512 atoms
// 720 -> 627 us, a tie at 600, the cluster kernel ahead above (and the fused warp-per-frame kernel below 64 atoms); per-atom sums
// 300 atoms 644 -> 416 us, 800 atoms 827 -> 773 us, the cluster kernel ahead from 1000 atoms
// Round 2 (profiles/r2_dispatch_probe.jsonl): with the mbarrier hand-over the cluster kernel needs ~530 us per 100 000 frames
// whatever the frame holds below 500 atoms, so it is ahead from ~480 atoms for both passes (write-back 512 atoms: 576 vs 630 us;
// per-atom sums 640 / 800 / 900 atoms: 602 / 634 / 653 vs 655 / 764 / 851 us).
constexpr int kSplitWbMinAtoms = 64, kSplitWbMaxDefault = 480, kSplitDevMaxDefault = 480;

// defaults a new handle copies (gcb_set_*); a handle's own settings are changed with gcb_traj_set_*
static std::mutex g_defaults_mu;
static Tuning g_defaults;
Tuning default_tuning() {
    std::lock_guard<std::mutex> lock(g_defaults_mu);
    return g_defaults;
}

bool smem_opted_in(const void* kern, int dev, bool mark) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    std::lock_guard<std::mutex> lock(mu);
    if (mark) { done.insert({kern, dev}); return true; }
    return done.count({kern, dev}) != 0;
}

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}

constexpr size_t kStageBytes = 64ull << 20;

struct RefCache {
    std::vector<double> ref;
    std::vector<int> idx_t, idx_r;
    int natoms_ref = -1;
    int mode = -1;  // 0 all, 1 weighted, 2 gather
    bool centred = false;
    bool valid = false;
    bool repeated = false;  // weighted mode: some index occurs more than once (multiplicities > 1)
    bool paired = false;    // weighted mode built from two different lists (template planes laid out by test atom)
    double n_fit = 0.0;
    double gb = 0.0;
    double bbar[3] = {0, 0, 0};
};

}  // namespace gcb

struct gcb_traj_ext : gcb_traj {
    gcb::RefCache cache;
};

using namespace gcb;

static gcb_traj_ext* X(gcb_traj* t) { return static_cast<gcb_traj_ext*>(t); }

static int check_window(const gcb_traj* t, long first, long nframes, bool within_nframes) {
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    const long lim = within_nframes ? t->nframes : t->cap;
    if (first < 0 || nframes < 0 || first + nframes > lim) {
        set_error("frame window [%ld, %ld) outside [0, %ld)", first, first + nframes, lim);
        return GCB_ERR_ARG;
    }
    return GCB_OK;
}

namespace gcb {

static inline bool ranges_meet(long a0, long a1, long b0, long b1) { return a0 < b1 && b0 < a1; }

int deps_before_compute(gcb_traj* t, long first, long nframes) {
    for (int s = 0; s < 2; s++)
        if (t->up_pending[s] && ranges_meet(first, first + nframes, t->up_lo[s], t->up_hi[s])) {
            GCB_CUDA(cudaStreamWaitEvent(t->stream, t->copy_done[s], 0));
            t->up_pending[s] = false;
        }
    return GCB_OK;
}

int deps_after_compute(gcb_traj* t, long first, long nframes) {
    if (nframes <= 0) return GCB_OK;
    bool covered = false;
    for (int s = 0; s < 2; s++)
        if (t->up_hi[s] > t->up_lo[s] && ranges_meet(first, first + nframes, t->up_lo[s], t->up_hi[s])) {
            GCB_CUDA(cudaEventRecord(t->use_done[s], t->stream));
            t->use_valid[s] = true;
            covered = covered || (first >= t->up_lo[s] && first + nframes <= t->up_hi[s]);
        }
    if (!covered) {   // frames outside the slots' windows (a resident trajectory): one event for their union
        GCB_CUDA(cudaEventRecord(t->any_done, t->stream));
        if (t->any_hi <= t->any_lo) { t->any_lo = first; t->any_hi = first + nframes; }
        else { if (first < t->any_lo) t->any_lo = first; if (first + nframes > t->any_hi) t->any_hi = first + nframes; }
    }
    return GCB_OK;
}

// an asynchronous upload into [first, first+n) through `slot`: its copy stream waits for the kernels that use those frames (and,
// through the slot's own stream order, for the slot's previous copy); a range the compute stream never waited for is closed first
static int deps_before_upload(gcb_traj* t, long first, long nframes, int slot) {
    cudaStream_t cs = t->copy_stream[slot];
    if (t->up_pending[slot]) {   // the previous upload through this slot was never consumed: order it before later kernels now
        GCB_CUDA(cudaStreamWaitEvent(t->stream, t->copy_done[slot], 0));
        t->up_pending[slot] = false;
    }
    for (int s = 0; s < 2; s++)
        if (t->use_valid[s] && ranges_meet(first, first + nframes, t->up_lo[s], t->up_hi[s]))
            GCB_CUDA(cudaStreamWaitEvent(cs, t->use_done[s], 0));
    if (t->any_hi > t->any_lo && ranges_meet(first, first + nframes, t->any_lo, t->any_hi)) GCB_CUDA(cudaStreamWaitEvent(cs, t->any_done, 0));
    // the other slot's upload may still be writing frames of this range (overlapping windows): keep the copies in order
    const int o = slot ^ 1;
    if (t->up_hi[o] > t->up_lo[o] && ranges_meet(first, first + nframes, t->up_lo[o], t->up_hi[o]))
        GCB_CUDA(cudaStreamWaitEvent(cs, t->copy_done[o], 0));
    return GCB_OK;
}

static void deps_after_upload(gcb_traj* t, long first, long nframes, int slot) {
    t->up_lo[slot] = first;
    t->up_hi[slot] = first + nframes;
    t->up_pending[slot] = true;
    t->use_valid[slot] = false;   // the new range has no users yet
}

}  // namespace gcb

extern "C" {

int gcb_version(void) { return 100; }

size_t gcb_last_error(char* buf, size_t cap) {
    if (buf && cap) {
        size_t n = t_error.size() < cap - 1 ? t_error.size() : cap - 1;
        memcpy(buf, t_error.data(), n);
        buf[n] = 0;
    }
    return t_error.size();
}

int gcb_device_count(int* n) {
    if (!n) return GCB_ERR_ARG;
    *n = 0;
    GCB_CUDA(cudaGetDeviceCount(n));
    return GCB_OK;
}

int gcb_device_info(int dev, int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor) {
    cudaDeviceProp pr;
    GCB_CUDA(cudaGetDeviceProperties(&pr, dev));
    if (sm_count) *sm_count = pr.multiProcessorCount;
    if (total_mem) *total_mem = pr.totalGlobalMem;
    if (cc_major) *cc_major = pr.major;
    if (cc_minor) *cc_minor = pr.minor;
    return GCB_OK;
}

static int check_variant(int variant) {
    if (variant < GCB_KERNEL_AUTO || variant > GCB_KERNEL_CLUSTER) { set_error("bad kernel variant %d", variant); return GCB_ERR_ARG; }
    return GCB_OK;
}

static int check_cluster_tuning(int cluster, int stages, int depth, int resident) {
    if (cluster < 0 || cluster > 8 || stages < 0 || stages > 8 || depth < 0 || depth > 6 || resident < -1 ||
        (resident > 1 && resident != (1 | (8 << 4)) && resident != (1 | (12 << 4)))) {
        set_error("bad cluster tuning (%d, %d, %d, %d)", cluster, stages, depth, resident);
        return GCB_ERR_ARG;
    }
    return GCB_OK;
}

static void apply_cluster_tuning(Tuning* tu, int cluster, int stages, int depth, int resident) {
    tu->cluster = cluster; tu->stages = stages; tu->depth = depth == 0 ? -1 : depth; tu->resident = resident;
}

/* process-wide DEFAULTS: copied by handles created afterwards, never read by a running call */
int gcb_set_kernel_variant(int variant) {
    int rc = check_variant(variant);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_defaults_mu);
    g_defaults.variant = variant;
    return GCB_OK;
}

int gcb_set_cluster_tuning(int cluster, int stages, int depth, int resident) {
    int rc = check_cluster_tuning(cluster, stages, depth, resident);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_defaults_mu);
    apply_cluster_tuning(&g_defaults, cluster, stages, depth, resident);
    return GCB_OK;
}

int gcb_set_split_thresholds(int write_back_max_atoms, int per_atom_max_atoms) {
    std::lock_guard<std::mutex> lock(g_defaults_mu);
    g_defaults.split_wb_max = write_back_max_atoms < 0 ? -1 : write_back_max_atoms;
    g_defaults.split_dev_max = per_atom_max_atoms < 0 ? -1 : per_atom_max_atoms;
    return GCB_OK;
}

int gcb_set_rmsd_mode(int explicit_only) {
    std::lock_guard<std::mutex> lock(g_defaults_mu);
    g_defaults.force_explicit = explicit_only ? 1 : 0;
    return GCB_OK;
}

/* the same settings of ONE handle; the caller orders them against its own calls on that handle */
int gcb_traj_set_kernel_variant(gcb_traj* t, int variant) {
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    int rc = check_variant(variant);
    if (rc) return rc;
    t->tune.variant = variant;
    return GCB_OK;
}

int gcb_traj_set_cluster_tuning(gcb_traj* t, int cluster, int stages, int depth, int resident) {
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    int rc = check_cluster_tuning(cluster, stages, depth, resident);
    if (rc) return rc;
    apply_cluster_tuning(&t->tune, cluster, stages, depth, resident);
    return GCB_OK;
}

int gcb_traj_set_split_thresholds(gcb_traj* t, int write_back_max_atoms, int per_atom_max_atoms) {
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    t->tune.split_wb_max = write_back_max_atoms < 0 ? -1 : write_back_max_atoms;
    t->tune.split_dev_max = per_atom_max_atoms < 0 ? -1 : per_atom_max_atoms;
    return GCB_OK;
}

int gcb_traj_set_rmsd_mode(gcb_traj* t, int explicit_only) {
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    t->tune.force_explicit = explicit_only ? 1 : 0;
    return GCB_OK;
}

int gcb_traj_last_cluster_plan(const gcb_traj* t, int* out8) {
    if (!t || !out8) return GCB_ERR_ARG;
    for (int i = 0; i < 8; i++) out8[i] = t->last_plan[i];
    return GCB_OK;
}

long gcb_launch_count(void) { return g_launches.load(); }

int gcb_pinned_alloc(size_t bytes, void** p) {
    if (!p) return GCB_ERR_ARG;
    *p = nullptr;
    GCB_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
    return GCB_OK;
}

// Run the calling thread on the CPUs next to device `dev` (its PCIe root's NUMA node, sysfs local_cpulist), so that the pinned
// buffers it allocates afterwards and the reader threads filling them are NUMA-local to the GPU's link.
int gcb_bind_host_thread(int dev) {
    char bus[32] = {0};
    GCB_CUDA(cudaDeviceGetPCIBusId(bus, sizeof bus, dev));
    for (char* c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) { set_error("gcb_bind_host_thread: %s is not readable", path); return GCB_ERR_ARG; }
    char line[4096] = {0};
    const bool got = fgets(line, sizeof line, f) != nullptr;
    fclose(f);
    cpu_set_t set;
    CPU_ZERO(&set);
    int ncpu = 0;
    if (got) {
        for (char* p = line; *p;) {   // "0-31,64-95"
            char* end = nullptr;
            long a = strtol(p, &end, 10);
            if (end == p) break;
            long b = a;
            p = end;
            if (*p == '-') { b = strtol(p + 1, &end, 10); p = end; }
            for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, &set); ncpu++; }
            if (*p == ',') p++; else break;
        }
    }
    if (ncpu == 0) { set_error("gcb_bind_host_thread: no CPU list for device %d (%s)", dev, path); return GCB_ERR_ARG; }
    if (sched_setaffinity(0, sizeof set, &set) != 0) { set_error("gcb_bind_host_thread: sched_setaffinity failed"); return GCB_ERR_ARG; }
    return GCB_OK;
}

int gcb_pinned_free(void* p) {
    if (!p) return GCB_OK;
    GCB_CUDA(cudaFreeHost(p));
    return GCB_OK;
}

int gcb_traj_destroy(gcb_traj* tb) {
    if (!tb) return GCB_OK;
    gcb_traj_ext* t = X(tb);
    cudaSetDevice(t->dev);
    if (t->stream) cudaStreamSynchronize(t->stream);
    for (int s = 0; s < 2; s++) {
        if (t->copy_stream[s]) { cudaStreamSynchronize(t->copy_stream[s]); cudaStreamDestroy(t->copy_stream[s]); }
        if (t->copy_done[s]) cudaEventDestroy(t->copy_done[s]);
        if (t->d_stage[s]) cudaFree(t->d_stage[s]);
    }
    if (t->ev0) cudaEventDestroy(t->ev0);
    if (t->ev1) cudaEventDestroy(t->ev1);
    for (int s = 0; s < 2; s++) if (t->use_done[s]) cudaEventDestroy(t->use_done[s]);
    if (t->any_done) cudaEventDestroy(t->any_done);
    if (t->stream) cudaStreamDestroy(t->stream);
    cudaFree(t->d_frames); cudaFree(t->d_refc); cudaFree(t->d_weight); cudaFree(t->d_refpairs); cudaFree(t->d_idx);
    cudaFree(t->d_rmsd); cudaFree(t->d_rot); cudaFree(t->d_t1); cudaFree(t->d_t2); cudaFree(t->d_status); cudaFree(t->d_flag);
    cudaFree(t->d_devpart); cudaFree(t->d_devsum); cudaFree(t->d_refplanes); cudaFree(t->d_want); cudaFree(t->d_wantref);
    cudaFree(t->d_mom_w); cudaFree(t->d_mom_c); cudaFree(t->d_mom_m);
    cudaFree(t->d_ap_ws);
    cudaFree(t->d_epart); cudaFree(t->d_pdist); cudaFree(t->d_lovo_ws);
    if (t->h_lovo) cudaFreeHost(t->h_lovo);
    if (t->h_scratch) cudaFreeHost(t->h_scratch);
    for (int s = 0; s < 2; s++) {
        if (t->h_bounce[s]) cudaFreeHost(t->h_bounce[s]);
        if (t->bounce_done[s]) cudaEventDestroy(t->bounce_done[s]);
    }
    delete t;
    return GCB_OK;
}

int gcb_traj_create(int dev, int natoms, long cap_frames, gcb_traj** out) {
    if (!out) return GCB_ERR_ARG;
    *out = nullptr;
    if (natoms <= 0 || cap_frames <= 0) { set_error("natoms and cap_frames must be positive"); return GCB_ERR_SHAPE; }
    GCB_CUDA(cudaSetDevice(dev));
    cudaDeviceProp pr;
    GCB_CUDA(cudaGetDeviceProperties(&pr, dev));
    if (pr.major != 10) {
        set_error("device %d is sm_%d%d; this library carries sm_100a code only", dev, pr.major, pr.minor);
        return GCB_ERR_CUDA;
    }
    gcb_traj_ext* t = new gcb_traj_ext();
    t->dev = dev;
    t->natoms = natoms;
    t->np = round_up(natoms, kPlaneAlign);
    t->cap = cap_frames;
    t->nframes = 0;
    t->sm_count = pr.multiProcessorCount;
    t->tune = default_tuning();
    int rc = GCB_OK;
#define TRY(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            set_error("%s failed: %s", #call, cudaGetErrorString(e__));                               \
            rc = (e__ == cudaErrorMemoryAllocation) ? GCB_ERR_ALLOC : GCB_ERR_CUDA;                   \
            goto fail;                                                                                \
        }                                                                                             \
    } while (0)
    {
        const size_t frame_bytes = sizeof(double) * 3 * (size_t)t->np;
        TRY(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
        for (int s = 0; s < 2; s++) {
            TRY(cudaStreamCreateWithFlags(&t->copy_stream[s], cudaStreamNonBlocking));
            TRY(cudaEventCreateWithFlags(&t->copy_done[s], cudaEventDisableTiming));
        }
        TRY(cudaEventCreate(&t->ev0));
        TRY(cudaEventCreate(&t->ev1));
        for (int s = 0; s < 2; s++) TRY(cudaEventCreateWithFlags(&t->use_done[s], cudaEventDisableTiming));
        TRY(cudaEventCreateWithFlags(&t->any_done, cudaEventDisableTiming));
        TRY(cudaMalloc(&t->d_frames, frame_bytes * (size_t)cap_frames));
        t->stage_bytes = kStageBytes;
        if (t->stage_bytes < frame_bytes) t->stage_bytes = frame_bytes;
        for (int s = 0; s < 2; s++) TRY(cudaMalloc(&t->d_stage[s], t->stage_bytes));
        TRY(cudaMalloc(&t->d_refc, frame_bytes));
        TRY(cudaMalloc(&t->d_weight, sizeof(double) * (size_t)t->np));
        TRY(cudaMalloc(&t->d_rmsd, sizeof(double) * (size_t)cap_frames));
        TRY(cudaMalloc(&t->d_rot, sizeof(double) * 9 * (size_t)cap_frames));
        TRY(cudaMalloc(&t->d_t1, sizeof(double) * 3 * (size_t)cap_frames));
        TRY(cudaMalloc(&t->d_t2, sizeof(double) * 3 * (size_t)cap_frames));
        TRY(cudaMalloc(&t->d_flag, (size_t)cap_frames));
        TRY(cudaMalloc(&t->d_status, sizeof(int)));
        TRY(cudaMemset(t->d_status, 0, sizeof(int)));
        TRY(cudaMalloc(&t->d_devsum, sizeof(double) * (size_t)t->np));
        t->h_scratch_bytes = sizeof(double) * 4 * (size_t)t->np + 64;
        TRY(cudaHostAlloc((void**)&t->h_scratch, t->h_scratch_bytes, cudaHostAllocPortable));
    }
#undef TRY
    *out = t;
    return GCB_OK;
fail:
    gcb_traj_destroy(t);
    return rc;
}

int gcb_traj_natoms(const gcb_traj* t) { return t ? t->natoms : -1; }
size_t gcb_traj_stage_bytes(const gcb_traj* t) { return t ? t->stage_bytes : 0; }
long gcb_traj_capacity(const gcb_traj* t) { return t ? t->cap : -1; }
long gcb_traj_nframes(const gcb_traj* t) { return t ? t->nframes : -1; }

int gcb_traj_set_nframes(gcb_traj* t, long nframes) {
    if (!t || nframes < 0 || nframes > t->cap) { set_error("nframes out of range"); return GCB_ERR_ARG; }
    t->nframes = nframes;
    return GCB_OK;
}

static size_t host_frame_bytes(const gcb_traj* t, int layout) {
    return (layout == GCB_F64_AOS ? sizeof(double) : sizeof(float)) * 3 * (size_t)t->natoms;
}

// ---- pageable host memory ----
// A Go slice (or any malloc'ed buffer) is pageable: handed to cudaMemcpyAsync it goes through the driver's own bounce buffer, one
// thread, ~10 GB/s.  The synchronous entry points copy it through two library-owned pinned buffers instead, several threads per
// copy, the DMA of one chunk under the memcpy of the next.
constexpr size_t kBounceBytes = 16ull << 20;

static int copy_threads() {
    static const int n = [] {
        const char* e = getenv("GCB_COPY_THREADS");
        int v = e ? atoi(e) : 0;
        if (v <= 0) { v = (int)std::thread::hardware_concurrency() / 2; if (v > 8) v = 8; }
        return v < 1 ? 1 : v;
    }();
    return n;
}

static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const int nt = bytes >= (2u << 20) ? copy_threads() : 1;
    if (nt == 1) { memcpy(dst, src, bytes); return; }
    const size_t part = ((bytes / nt) + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    size_t mine = part < bytes ? part : bytes;   // [0, mine) is copied by the calling thread, and so is whatever no helper took
    size_t handed = mine;
    for (int i = 1; i < nt && handed < bytes; i++) {
        const size_t off = handed;
        const size_t n = (off + part < bytes) ? part : bytes - off;
        try {
            th.emplace_back([=] { memcpy((char*)dst + off, (const char*)src + off, n); });
        } catch (...) {   // no more threads to be had: the caller copies the rest
            break;
        }
        handed += n;
    }
    memcpy(dst, src, mine);
    if (handed < bytes) memcpy((char*)dst + handed, (const char*)src + handed, bytes - handed);
    for (auto& x : th) x.join();
}

static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

static int ensure_bounce(gcb_traj* t) {
    if (t->h_bounce[0]) return GCB_OK;
    size_t bytes = kBounceBytes < t->stage_bytes ? kBounceBytes : t->stage_bytes;
    const size_t frame = sizeof(double) * 3 * (size_t)t->natoms;
    if (bytes < frame) bytes = frame;
    for (int s = 0; s < 2; s++) {
        GCB_CUDA(cudaHostAlloc(&t->h_bounce[s], bytes, cudaHostAllocPortable));
        GCB_CUDA(cudaEventCreateWithFlags(&t->bounce_done[s], cudaEventDisableTiming));
    }
    t->bounce_bytes = bytes;
    return GCB_OK;
}

int gcb_traj_upload(gcb_traj* t, long first, const void* host, long nframes, int layout, double scale) {
    int rc = check_window(t, first, nframes, false);
    if (rc) return rc;
    if (!host && nframes) { set_error("NULL host buffer"); return GCB_ERR_ARG; }
    if (layout < GCB_F64_AOS || layout > GCB_I32_AOS) { set_error("unknown host layout %d", layout); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    GCB_CUDA(cudaStreamSynchronize(t->stream));   // synchronous call: queued kernels may still read the frames being replaced
    t->use_valid[0] = t->use_valid[1] = false;
    t->any_lo = t->any_hi = 0;
    const size_t fb = host_frame_bytes(t, layout);
    if (nframes > 0 && (size_t)nframes * fb >= (1u << 20) && is_pageable(host)) {
        rc = ensure_bounce(t);
        if (rc) return rc;
        const long per = (long)(t->bounce_bytes / fb);
        int slot = 0;
        for (long done = 0; done < nframes; done += per, slot ^= 1) {
            const long n = (nframes - done < per) ? nframes - done : per;
            cudaStream_t s = t->copy_stream[slot];
            if (done >= 2 * per) GCB_CUDA(cudaEventSynchronize(t->bounce_done[slot]));   // the DMA out of this buffer has finished
            parallel_memcpy(t->h_bounce[slot], (const char*)host + (size_t)done * fb, (size_t)n * fb);
            GCB_CUDA(cudaMemcpyAsync(t->d_stage[slot], t->h_bounce[slot], (size_t)n * fb, cudaMemcpyHostToDevice, s));
            GCB_CUDA(cudaEventRecord(t->bounce_done[slot], s));
            rc = launch_upload_convert(t, s, t->d_stage[slot], first + done, n, layout, scale);
            if (rc) return rc;
        }
        GCB_CUDA(cudaStreamSynchronize(t->copy_stream[0]));
        GCB_CUDA(cudaStreamSynchronize(t->copy_stream[1]));
        t->up_pending[0] = t->up_pending[1] = false;   // everything uploaded so far is resident
        if (first + nframes > t->nframes) t->nframes = first + nframes;
        return GCB_OK;
    }
    const long per = (long)(t->stage_bytes / fb);
    int slot = 0;
    for (long done = 0; done < nframes; done += per, slot ^= 1) {
        const long n = (nframes - done < per) ? nframes - done : per;
        cudaStream_t s = t->copy_stream[slot];
        GCB_CUDA(cudaMemcpyAsync(t->d_stage[slot], (const char*)host + (size_t)done * fb, (size_t)n * fb, cudaMemcpyHostToDevice, s));
        rc = launch_upload_convert(t, s, t->d_stage[slot], first + done, n, layout, scale);
        if (rc) return rc;
    }
    GCB_CUDA(cudaStreamSynchronize(t->copy_stream[0]));
    GCB_CUDA(cudaStreamSynchronize(t->copy_stream[1]));
    t->up_pending[0] = t->up_pending[1] = false;   // everything uploaded so far is resident
    if (first + nframes > t->nframes) t->nframes = first + nframes;
    return GCB_OK;
}

int gcb_traj_upload_async(gcb_traj* t, long first, const void* pinned, long nframes, int layout, double scale, int slot) {
    int rc = check_window(t, first, nframes, false);
    if (rc) return rc;
    if (slot < 0 || slot > 1) { set_error("slot must be 0 or 1"); return GCB_ERR_ARG; }
    if (layout < GCB_F64_AOS || layout > GCB_I32_AOS) { set_error("unknown host layout %d", layout); return GCB_ERR_ARG; }
    const size_t bytes = host_frame_bytes(t, layout) * (size_t)nframes;
    if (bytes > t->stage_bytes) { set_error("async upload of %zu bytes exceeds the %zu-byte staging slot", bytes, t->stage_bytes); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    cudaStream_t s = t->copy_stream[slot];
    rc = deps_before_upload(t, first, nframes, slot);   // do not overwrite frames a queued kernel still reads
    if (rc) return rc;
    GCB_CUDA(cudaMemcpyAsync(t->d_stage[slot], pinned, bytes, cudaMemcpyHostToDevice, s));
    rc = launch_upload_convert(t, s, t->d_stage[slot], first, nframes, layout, scale);
    if (rc) return rc;
    GCB_CUDA(cudaEventRecord(t->copy_done[slot], s));
    deps_after_upload(t, first, nframes, slot);          // kernels that read these frames wait for copy_done[slot], others do not
    if (first + nframes > t->nframes) t->nframes = first + nframes;
    return GCB_OK;
}

int gcb_traj_upload_records_async(gcb_traj* t, long first, const void* pinned, long nframes, long record_bytes, long off_x,
                                  long off_y, long off_z, double scale, int slot) {
    int rc = check_window(t, first, nframes, false);
    if (rc) return rc;
    if (slot < 0 || slot > 1) { set_error("slot must be 0 or 1"); return GCB_ERR_ARG; }
    const long plane = 4L * t->natoms;
    if (record_bytes <= 0 || (record_bytes & 3) || ((off_x | off_y | off_z) & 3) || off_x < 0 || off_y < 0 || off_z < 0 ||
        off_x + plane > record_bytes || off_y + plane > record_bytes || off_z + plane > record_bytes) {
        set_error("record layout (%ld bytes, planes at %ld / %ld / %ld) does not hold three float32 planes of %d atoms", record_bytes,
                  off_x, off_y, off_z, t->natoms);
        return GCB_ERR_ARG;
    }
    const size_t bytes = (size_t)record_bytes * (size_t)nframes;
    if (bytes > t->stage_bytes) { set_error("async upload of %zu bytes exceeds the %zu-byte staging slot", bytes, t->stage_bytes); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    cudaStream_t s = t->copy_stream[slot];
    rc = deps_before_upload(t, first, nframes, slot);   // do not overwrite frames a queued kernel still reads
    if (rc) return rc;
    GCB_CUDA(cudaMemcpyAsync(t->d_stage[slot], pinned, bytes, cudaMemcpyHostToDevice, s));
    rc = launch_upload_records(t, s, t->d_stage[slot], first, nframes, record_bytes, off_x, off_y, off_z, scale);
    if (rc) return rc;
    GCB_CUDA(cudaEventRecord(t->copy_done[slot], s));
    deps_after_upload(t, first, nframes, slot);
    if (first + nframes > t->nframes) t->nframes = first + nframes;
    return GCB_OK;
}

int gcb_traj_wait(gcb_traj* t, int slot) {
    if (!t || slot < 0 || slot > 1) return GCB_ERR_ARG;
    GCB_CUDA(cudaSetDevice(t->dev));
    GCB_CUDA(cudaStreamSynchronize(t->copy_stream[slot]));
    return GCB_OK;
}

int gcb_traj_download(gcb_traj* t, long first, long nframes, double* host_aos) {
    int rc = check_window(t, first, nframes, true);
    if (rc) return rc;
    if (!host_aos && nframes) { set_error("NULL host buffer"); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    GCB_CUDA(cudaStreamSynchronize(t->copy_stream[0]));   // the staging slots may still feed an upload; uploads are resident after this
    GCB_CUDA(cudaStreamSynchronize(t->copy_stream[1]));
    t->up_pending[0] = t->up_pending[1] = false;
    const size_t fb = sizeof(double) * 3 * (size_t)t->natoms;
    if (nframes > 0 && (size_t)nframes * fb >= (1u << 20) && is_pageable(host_aos)) {
        // chunk c: layout kernel -> device staging slot -> pinned buffer (DMA); the memcpy of chunk c - 1 into the caller's
        // memory runs on the host meanwhile
        rc = ensure_bounce(t);
        if (rc) return rc;
        const long per = (long)(t->bounce_bytes / fb);
        long prev_done = -1, prev_n = 0;
        int slot = 0;
        for (long done = 0; done < nframes; done += per, slot ^= 1) {
            const long n = (nframes - done < per) ? nframes - done : per;
            rc = launch_download_convert(t, t->stream, (double*)t->d_stage[slot], first + done, n);
            if (rc) return rc;
            GCB_CUDA(cudaMemcpyAsync(t->h_bounce[slot], t->d_stage[slot], (size_t)n * fb, cudaMemcpyDeviceToHost, t->stream));
            GCB_CUDA(cudaEventRecord(t->bounce_done[slot], t->stream));
            if (prev_done >= 0) {
                GCB_CUDA(cudaEventSynchronize(t->bounce_done[slot ^ 1]));
                parallel_memcpy((char*)host_aos + (size_t)prev_done * fb, t->h_bounce[slot ^ 1], (size_t)prev_n * fb);
            }
            prev_done = done; prev_n = n;
        }
        GCB_CUDA(cudaEventSynchronize(t->bounce_done[slot ^ 1]));
        parallel_memcpy((char*)host_aos + (size_t)prev_done * fb, t->h_bounce[slot ^ 1], (size_t)prev_n * fb);
        return GCB_OK;
    }
    const long per = (long)(t->stage_bytes / fb);
    for (long done = 0; done < nframes; done += per) {
        const long n = (nframes - done < per) ? nframes - done : per;
        rc = launch_download_convert(t, t->stream, (double*)t->d_stage[0], first + done, n);
        if (rc) return rc;
        GCB_CUDA(cudaMemcpyAsync((char*)host_aos + (size_t)done * fb, t->d_stage[0], (size_t)n * fb, cudaMemcpyDeviceToHost, t->stream));
        GCB_CUDA(cudaStreamSynchronize(t->stream));
    }
    return GCB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// template preparation
// ---------------------------------------------------------------------------
namespace {

enum { MODE_ALL = 0, MODE_WEIGHTED = 1, MODE_GATHER = 2 };

// Decide the formulation, validate like the reference, centre the template (centred == true) or
// keep it raw (no-fit RMSD), and make it resident.  Cached on content so back-to-back calls with the
// same template and lists (bench loops, LOVO iterations that did not change the set) skip it.
int prepare_template(gcb_traj_ext* t, const double* ref_aos, int natoms_ref, const int* idx_test, const int* idx_ref,
                     int nidx, bool centred, int* mode_out, bool prefer_gather = false) {
    if (!ref_aos) { set_error("NULL template"); return GCB_ERR_ARG; }
    if (nidx < 0) { set_error("negative index count"); return GCB_ERR_ARG; }
    bool all = (idx_test == nullptr || nidx == 0);
    if (!all && nidx == t->natoms && natoms_ref == t->natoms) {
        // both lists name every atom once, in order (align.LOVO's first iteration when every atom is a candidate): the fit on
        // all atoms -- the same arithmetic without a mask or weights
        const int* r0 = idx_ref ? idx_ref : idx_test;
        bool identity = true;
        for (int k = 0; k < nidx && identity; k++) identity = idx_test[k] == k && r0[k] == k;
        if (identity) all = true;
    }
    if (all && natoms_ref != t->natoms) {
        set_error("chem.Super: Ill formed coordinates for Superposition (test %d atoms, templa %d)", t->natoms, natoms_ref);
        return GCB_ERR_SHAPE;
    }
    if (natoms_ref <= 0) { set_error("empty template"); return GCB_ERR_SHAPE; }
    const int* ir = idx_ref ? idx_ref : idx_test;
    int mode = MODE_ALL;
    bool paired = false;   // different lists, no test atom twice: the pairing is folded into the template planes
    if (!all) {
        bool same = (natoms_ref == t->natoms);
        for (int k = 0; k < nidx; k++) {
            if (idx_test[k] < 0 || idx_test[k] >= t->natoms || ir[k] < 0 || ir[k] >= natoms_ref) {
                set_error("index %d of the fit lists is out of range (test %d, templa %d)", k, idx_test[k], ir[k]);
                return GCB_ERR_INDEXES;
            }
            same = same && (idx_test[k] == ir[k]);
        }
        if (!same) {
            // Test atom idx_test[k] is fitted onto template atom idx_ref[k].  When no test atom occurs twice the template can be
            // laid out by TEST atom (plane entry idx_test[k] = templa[idx_ref[k]]) and the fit becomes a subset fit of the
            // streaming kernels -- no gather of the frame.  Lists that repeat a test atom keep the gather kernels.
            std::vector<unsigned char> used((size_t)t->natoms, 0);
            paired = true;
            for (int k = 0; k < nidx && paired; k++) {
                if (used[(size_t)idx_test[k]]) paired = false;
                used[(size_t)idx_test[k]] = 1;
            }
        }
        mode = (same || paired) ? MODE_WEIGHTED : MODE_GATHER;
        // a small subset of a large frame, frames left untouched: reading just those atoms (3 sectors each) beats streaming
        // the whole frame
        if (prefer_gather) mode = MODE_GATHER;
    }
    *mode_out = mode;
    gcb::RefCache& c = t->cache;
    const size_t nref = 3 * (size_t)natoms_ref;
    if (c.valid && c.mode == mode && c.centred == centred && c.natoms_ref == natoms_ref && c.ref.size() == nref &&
        memcmp(c.ref.data(), ref_aos, nref * sizeof(double)) == 0 && (int)c.idx_t.size() == (all ? 0 : nidx) &&
        (all || (memcmp(c.idx_t.data(), idx_test, sizeof(int) * nidx) == 0 && memcmp(c.idx_r.data(), ir, sizeof(int) * nidx) == 0)))
        return GCB_OK;

    c.valid = false;
    c.paired = false;
    GCB_CUDA(cudaStreamSynchronize(t->stream));  // h_scratch and the device template may still be in use
    const int np = t->np;
    const int n = all ? t->natoms : nidx;
    // centroid of the template fit subset, rows summed in list order like the 1 x n * n x 3 product of
    // chem.MassCenter (geometric.go:684-690)
    double s0 = 0, s1 = 0, s2 = 0;
    for (int k = 0; k < n; k++) {
        const int i = all ? k : ir[k];
        s0 += ref_aos[3 * (size_t)i]; s1 += ref_aos[3 * (size_t)i + 1]; s2 += ref_aos[3 * (size_t)i + 2];
    }
    const double inv = 1.0 / (double)n;
    c.bbar[0] = centred ? s0 * inv : 0.0;
    c.bbar[1] = centred ? s1 * inv : 0.0;
    c.bbar[2] = centred ? s2 * inv : 0.0;
    c.n_fit = (double)n;
    if (mode == MODE_GATHER) {
        if (nidx > t->idx_cap) {
            cudaFree(t->d_refpairs); cudaFree(t->d_idx);
            t->d_refpairs = nullptr; t->d_idx = nullptr; t->idx_cap = 0;
            GCB_CUDA(cudaMalloc(&t->d_refpairs, sizeof(double) * 3 * (size_t)nidx));
            GCB_CUDA(cudaMalloc(&t->d_idx, sizeof(int) * (size_t)nidx));
            t->idx_cap = nidx;
        }
        std::vector<double> pairs(3 * (size_t)nidx);
        double gb = 0.0;
        for (int k = 0; k < nidx; k++) {
            const size_t i = (size_t)ir[k];
            pairs[k] = ref_aos[3 * i] - c.bbar[0];
            pairs[(size_t)nidx + k] = ref_aos[3 * i + 1] - c.bbar[1];
            pairs[2 * (size_t)nidx + k] = ref_aos[3 * i + 2] - c.bbar[2];
            gb += pairs[k] * pairs[k] + pairs[(size_t)nidx + k] * pairs[(size_t)nidx + k] + pairs[2 * (size_t)nidx + k] * pairs[2 * (size_t)nidx + k];
        }
        c.gb = gb;
        GCB_CUDA(cudaMemcpyAsync(t->d_refpairs, pairs.data(), sizeof(double) * pairs.size(), cudaMemcpyHostToDevice, t->stream));
        GCB_CUDA(cudaMemcpyAsync(t->d_idx, idx_test, sizeof(int) * (size_t)nidx, cudaMemcpyHostToDevice, t->stream));
        GCB_CUDA(cudaStreamSynchronize(t->stream));
    } else {
        double* h = t->h_scratch;  // [3*np planes | np weights]
        memset(h, 0, sizeof(double) * 4 * (size_t)np);
        if (paired) {
            for (int k = 0; k < nidx; k++) {
                const size_t i = (size_t)idx_test[k], j = (size_t)ir[k];
                h[i] = ref_aos[3 * j] - c.bbar[0];
                h[np + i] = ref_aos[3 * j + 1] - c.bbar[1];
                h[2 * (size_t)np + i] = ref_aos[3 * j + 2] - c.bbar[2];
            }
        } else {
            for (int i = 0; i < t->natoms; i++) {
                h[i] = ref_aos[3 * (size_t)i] - c.bbar[0];
                h[np + i] = ref_aos[3 * (size_t)i + 1] - c.bbar[1];
                h[2 * (size_t)np + i] = ref_aos[3 * (size_t)i + 2] - c.bbar[2];
            }
        }
        c.paired = paired;
        GCB_CUDA(cudaMemcpyAsync(t->d_refc, h, sizeof(double) * 3 * (size_t)np, cudaMemcpyHostToDevice, t->stream));
        double* w = h + 3 * (size_t)np;
        if (mode == MODE_WEIGHTED) {
            c.repeated = false;
            for (int k = 0; k < nidx; k++) {   // multiplicity: a repeated index counts twice, as in SomeVecs
                w[idx_test[k]] += 1.0;
                if (w[idx_test[k]] > 1.0) c.repeated = true;
            }
            GCB_CUDA(cudaMemcpyAsync(t->d_weight, w, sizeof(double) * (size_t)np, cudaMemcpyHostToDevice, t->stream));
        }
        double gb = 0.0;
        for (int i = 0; i < t->natoms; i++) {
            const double wi = (mode == MODE_WEIGHTED) ? w[i] : 1.0;
            gb += wi * (h[i] * h[i] + h[np + i] * h[np + i] + h[2 * (size_t)np + i] * h[2 * (size_t)np + i]);
        }
        c.gb = gb;
        GCB_CUDA(cudaStreamSynchronize(t->stream));
    }
    c.ref.assign(ref_aos, ref_aos + nref);
    c.idx_t.clear(); c.idx_r.clear();
    if (!all) { c.idx_t.assign(idx_test, idx_test + nidx); c.idx_r.assign(ir, ir + nidx); }
    c.natoms_ref = natoms_ref;
    c.mode = mode;
    c.centred = centred;
    c.valid = true;
    return GCB_OK;
}

void fill_params(gcb_traj_ext* t, long first, long nframes, int write_back, SuperParams* p) {
    memset(p, 0, sizeof *p);
    p->frames = t->d_frames + first * 3L * t->np;
    p->frames_rw = write_back ? t->d_frames + first * 3L * t->np : nullptr;
    p->frame_stride = 3L * t->np;
    p->np = t->np;
    p->natoms = t->natoms;
    p->nframes = nframes;
    p->refc = t->d_refc;
    p->weight = nullptr;
    p->n_fit = t->cache.n_fit;
    p->inv_n = 1.0 / t->cache.n_fit;
    p->bbar[0] = t->cache.bbar[0]; p->bbar[1] = t->cache.bbar[1]; p->bbar[2] = t->cache.bbar[2];
    p->gb = t->cache.gb;
    p->need_explicit = t->d_flag + first;
    p->rmsd = t->d_rmsd + first;
    p->rot = t->d_rot + 9 * first;
    p->trans1 = t->d_t1 + 3 * first;
    p->trans2 = t->d_t2 + 3 * first;
    p->devpart = nullptr;
    p->status = t->d_status;
}

int ensure_devpart(gcb_traj_ext* t, size_t rows) {
    if (rows <= t->devpart_rows) return GCB_OK;
    cudaFree(t->d_devpart);
    t->d_devpart = nullptr;
    t->devpart_rows = 0;
    GCB_CUDA(cudaMalloc(&t->d_devpart, sizeof(double) * rows * (size_t)t->np));
    t->devpart_rows = rows;
    return GCB_OK;
}

int check_status(gcb_traj_ext* t) {
    int st = 0;
    GCB_CUDA(cudaMemcpyAsync(&st, t->d_status, sizeof(int), cudaMemcpyDeviceToHost, t->stream));
    GCB_CUDA(cudaStreamSynchronize(t->stream));
    if (st != 0) {
        cudaMemsetAsync(t->d_status, 0, sizeof(int), t->stream);
        if (st == (int)GCB_ERR_CUDA) {
            set_error("the exchange of the per-atom sums timed out: a peer rank never delivered its vector (GCB_EXCHANGE_TIMEOUT_S)");
            return GCB_ERR_CUDA;
        }
        set_error("RotatorTranslatorToSuper: mat.SVD failed (non-finite correlation matrix in at least one frame)");
        return GCB_ERR_SVD;
    }
    return GCB_OK;
}

bool use_cluster(const gcb_traj_ext* t, bool heavy) {
    if (t->tune.variant == GCB_KERNEL_GENERIC) return false;
    return cluster_kernel_supported(t, heavy);
}

int launch_super_split(gcb_traj_ext* t, long first, long nframes, const double* ref_aos, int natoms_ref, const int* idx_test,
                       const int* idx_ref, int nidx, int write_back, bool want_dev, int* dev_rows);

int launch_super_impl(gcb_traj_ext* t, long first, long nframes, const double* ref_aos, int natoms_ref, const int* idx_test,
                      const int* idx_ref, int nidx, int write_back, bool want_dev, int* dev_rows);

// every fused pass goes through here: it waits for the uploads of the frames it touches (and for no others) and leaves a mark
// for later uploads into the same frames
int launch_super(gcb_traj_ext* t, long first, long nframes, const double* ref_aos, int natoms_ref, const int* idx_test,
                 const int* idx_ref, int nidx, int write_back, bool want_dev, int* dev_rows) {
    int rc = deps_before_compute(t, first, nframes);
    if (rc) return rc;
    rc = launch_super_impl(t, first, nframes, ref_aos, natoms_ref, idx_test, idx_ref, nidx, write_back, want_dev, dev_rows);
    if (rc) return rc;
    return deps_after_compute(t, first, nframes);
}

int launch_super_impl(gcb_traj_ext* t, long first, long nframes, const double* ref_aos, int natoms_ref, const int* idx_test,
                      const int* idx_ref, int nidx, int write_back, bool want_dev, int* dev_rows) {
    // Frames beyond the cluster kernel's capacity whose every atom must be visited a second time (write-back, per-atom sums):
    // the fit and the elementwise sweep are separate launches (transform.cu) instead of the generic kernel's re-read.
    const int variant = t->tune.variant;
    const int split_wb_max = t->tune.split_wb_max < 0 ? kSplitWbMaxDefault : t->tune.split_wb_max;
    const int split_dev_max = t->tune.split_dev_max < 0 ? kSplitDevMaxDefault : t->tune.split_dev_max;
    auto force_explicit_mode = [t] { return t->tune.force_explicit; };
    const bool small_split = (write_back && !want_dev && t->natoms >= kSplitWbMinAtoms && t->natoms <= split_wb_max) ||
                             (want_dev && !write_back && t->natoms <= split_dev_max);
    if (variant == GCB_KERNEL_AUTO && (write_back || want_dev) && nframes > 0 && !force_explicit_mode() &&
        (small_split || !cluster_kernel_supported(t, true)))
        return launch_super_split(t, first, nframes, ref_aos, natoms_ref, idx_test, idx_ref, nidx, write_back, want_dev, dev_rows);
    int mode = MODE_ALL;
    // Subset fits that leave the frames untouched need only the subset's atoms.  Gathering them costs 3 sectors (96 bytes) per atom
    // plus a per-frame cost, streaming the frame 24 bytes per atom of the WHOLE frame.  Measured (tools/misc_bench.py): the
    // warp-per-frame gather (<= 512 atoms, ~20 ns per frame at 300 atoms) wins once the frame is 16 times the subset; the
    // CTA-per-frame gather kernel (~0.2 us per frame) only against frames the cluster kernel cannot take (> 28 672 atoms).
    const bool prefer_gather = variant == GCB_KERNEL_AUTO && !write_back && !want_dev && idx_test && nidx > 0 &&
                               (nidx <= 512 ? 16L * nidx <= t->natoms
                                : nidx <= small_gather_max_atoms() ? 8L * nidx <= t->natoms
                                                                   : (t->natoms > 28672 && 12L * nidx <= t->natoms));
    static const bool timing = getenv("GCB_TIMING") != nullptr;
    const auto tp0 = std::chrono::steady_clock::now();
    int rc = prepare_template(t, ref_aos, natoms_ref, idx_test, idx_ref, nidx, true, &mode, prefer_gather);
    if (timing) fprintf(stderr, "[gcb] prepare_template %.1f us\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tp0).count());
    if (rc) return rc;
    if (nframes == 0) return GCB_OK;
    SuperParams p;
    fill_params(t, first, nframes, write_back, &p);
    if (mode == MODE_GATHER) {
        if (want_dev) { set_error("per-atom deviation sums need identical test/templa index lists"); return GCB_ERR_INDEXES; }
        if (variant == GCB_KERNEL_AUTO && !write_back && nidx <= small_gather_max_atoms())
            return launch_super_small_gather(t, p, t->d_refpairs, t->d_idx, nidx, force_explicit_mode());
        return launch_super_gather(t, p, t->d_refpairs, t->d_idx, nidx);
    }
    if (mode == MODE_WEIGHTED) p.weight = t->d_weight;
    if (want_dev && mode == MODE_WEIGHTED && t->cache.paired) { set_error("per-atom deviation sums need identical test/templa index lists"); return GCB_ERR_INDEXES; }
    // the cluster kernel takes the fit subset as a 0/1 mask; index lists with repeats (multiplicities) go to the generic kernel
    const bool cluster = use_cluster(t, write_back != 0 || want_dev) && !(mode == MODE_WEIGHTED && t->cache.repeated);
    if (variant == GCB_KERNEL_CLUSTER && !cluster && !(mode == MODE_WEIGHTED && t->cache.repeated)) {
        set_error("cluster kernel requested but not available for %d atoms", t->natoms);
        return GCB_ERR_ARG;
    }
    // structures of at most 512 atoms: every warp owns frames (super_small.cu); the explicit variants keep their meaning
    const bool small = variant == GCB_KERNEL_AUTO && small_kernel_supported(t, want_dev || write_back != 0 || force_explicit_mode() != 0) && !(mode == MODE_WEIGHTED && t->cache.repeated);
    if (want_dev) {
        const size_t rows = small ? (size_t)small_kernel_rows(t) : (size_t)t->sm_count * 8;
        rc = ensure_devpart(t, rows);
        if (rc) return rc;
        GCB_CUDA(cudaMemsetAsync(t->d_devpart, 0, sizeof(double) * rows * (size_t)t->np, t->stream));
        p.devpart = t->d_devpart;
        p.rmsd = nullptr;   // align.RMSDTraj (align/lovo.go:286-353) never looks at the per-frame RMSD
    }
    int rows_used = 0;
    rc = small ? launch_super_small(t, p, mode == MODE_WEIGHTED, want_dev, force_explicit_mode(), &rows_used)
         : cluster ? launch_super_cluster(t, p, mode == MODE_WEIGHTED, want_dev, &rows_used)
                   : launch_super_generic(t, p, mode == MODE_WEIGHTED, want_dev, &rows_used);
    if (dev_rows) *dev_rows = rows_used;
    return rc;
}

int launch_super_split(gcb_traj_ext* t, long first, long nframes, const double* ref_aos, int natoms_ref, const int* idx_test,
                       const int* idx_ref, int nidx, int write_back, bool want_dev, int* dev_rows) {
    if (want_dev) {
        // align.concproc (align/lovo.go:270-274) passes the same list for both structures, which have the same atoms
        const bool all = (idx_test == nullptr || nidx == 0);
        bool same = natoms_ref == t->natoms;
        const int* ir = idx_ref ? idx_ref : idx_test;
        for (int k = 0; same && !all && k < nidx; k++) same = idx_test[k] == ir[k];
        if (!same) { set_error("per-atom deviation sums need identical test/templa index lists"); return GCB_ERR_INDEXES; }
    }
    // 1. the fit: R, trans1, trans2 and the RMSD of every frame, frames untouched
    int rc = launch_super(t, first, nframes, ref_aos, natoms_ref, idx_test, idx_ref, nidx, 0, false, nullptr);
    if (rc) return rc;
    SuperParams p;
    fill_params(t, first, nframes, write_back, &p);
    if (want_dev) {
        // 2a. per-atom distances to the template after the fit, summed over the frames (before any write-back: they are
        //     formed from the original coordinates and the transform)
        const int np = t->np;
        if (!t->d_refplanes) GCB_CUDA(cudaMalloc(&t->d_refplanes, sizeof(double) * 3 * (size_t)np));
        GCB_CUDA(cudaStreamSynchronize(t->stream));   // h_scratch may still feed an earlier copy
        double* h = t->h_scratch;
        memset(h, 0, sizeof(double) * 3 * (size_t)np);
        for (int i = 0; i < t->natoms; i++) {
            h[i] = ref_aos[3 * (size_t)i] - p.bbar[0];
            h[np + i] = ref_aos[3 * (size_t)i + 1] - p.bbar[1];
            h[2 * (size_t)np + i] = ref_aos[3 * (size_t)i + 2] - p.bbar[2];
        }
        GCB_CUDA(cudaMemcpyAsync(t->d_refplanes, h, sizeof(double) * 3 * (size_t)np, cudaMemcpyHostToDevice, t->stream));
        rc = ensure_devpart(t, (size_t)peratom_accum_rows(t, nframes));
        if (rc) return rc;
        p.devpart = t->d_devpart;
        int rows = 0;
        rc = launch_peratom_accum(t, p, t->d_refplanes, &rows);
        if (rc) return rc;
        if (dev_rows) *dev_rows = rows;
    }
    // 2b. test = (test + trans1) R + trans2 for every atom, in place (handy.go:207-211)
    if (write_back) rc = launch_transform_frames(t, p);
    return rc;
}

}  // namespace

extern "C" {

int gcb_super_rmsd_async(gcb_traj* tb, long first, long nframes, const double* ref_aos, int natoms_ref,
                         const int* idx_test, const int* idx_ref, int nidx, int write_back) {
    int rc = check_window(tb, first, nframes, true);
    if (rc) return rc;
    gcb_traj_ext* t = X(tb);
    GCB_CUDA(cudaSetDevice(t->dev));
    return launch_super(t, first, nframes, ref_aos, natoms_ref, idx_test, idx_ref, nidx, write_back, false, nullptr);
}

int gcb_fetch_results(gcb_traj* tb, long first, long nframes, double* rmsd, double* rot, double* trans1, double* trans2) {
    int rc = check_window(tb, first, nframes, false);
    if (rc) return rc;
    gcb_traj_ext* t = X(tb);
    GCB_CUDA(cudaSetDevice(t->dev));
    cudaStream_t s = t->stream;
    const size_t n = (size_t)nframes;
    if (rmsd) GCB_CUDA(cudaMemcpyAsync(rmsd, t->d_rmsd + first, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    if (rot) GCB_CUDA(cudaMemcpyAsync(rot, t->d_rot + 9 * first, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, s));
    if (trans1) GCB_CUDA(cudaMemcpyAsync(trans1, t->d_t1 + 3 * first, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s));
    if (trans2) GCB_CUDA(cudaMemcpyAsync(trans2, t->d_t2 + 3 * first, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s));
    return check_status(t);
}

int gcb_super_rmsd(gcb_traj* t, long first, long nframes, const double* ref_aos, int natoms_ref, const int* idx_test,
                   const int* idx_ref, int nidx, int write_back, double* rmsd, double* rot, double* trans1, double* trans2) {
    int rc = gcb_super_rmsd_async(t, first, nframes, ref_aos, natoms_ref, idx_test, idx_ref, nidx, write_back);
    if (rc) return rc;
    return gcb_fetch_results(t, first, nframes, rmsd, rot, trans1, trans2);
}

int gcb_rmsd(gcb_traj* tb, long first, long nframes, const double* ref_aos, int natoms_ref, const int* idx_test,
             const int* idx_ref, int nidx, double* rmsd) {
    int rc = check_window(tb, first, nframes, true);
    if (rc) return rc;
    gcb_traj_ext* t = X(tb);
    GCB_CUDA(cudaSetDevice(t->dev));
    int mode = MODE_ALL;
    // a small subset of a large frame: read only its atoms (one CTA per frame, ~0.1 us per frame) instead of streaming the frame
    const bool prefer_gather = t->tune.variant == GCB_KERNEL_AUTO && idx_test && nidx > 0 && 32L * nidx <= t->natoms && t->natoms >= 16384;
    rc = prepare_template(t, ref_aos, natoms_ref, idx_test, idx_ref, nidx, false, &mode, prefer_gather);
    if (rc) return rc;
    if (nframes == 0) return GCB_OK;
    rc = deps_before_compute(t, first, nframes);
    if (rc) return rc;
    const double* frames = t->d_frames + first * 3L * t->np;
    if (mode == MODE_GATHER) rc = launch_rmsd_nofit_gather(t, frames, nframes, t->d_refpairs, t->d_idx, nidx, t->d_rmsd + first);
    else rc = launch_rmsd_nofit(t, frames, nframes, t->d_refc, mode == MODE_WEIGHTED ? t->d_weight : nullptr, t->cache.n_fit, t->d_rmsd + first);
    if (rc) return rc;
    if (rmsd) GCB_CUDA(cudaMemcpyAsync(rmsd, t->d_rmsd + first, sizeof(double) * (size_t)nframes, cudaMemcpyDeviceToHost, t->stream));
    GCB_CUDA(cudaStreamSynchronize(t->stream));
    return deps_after_compute(t, first, nframes);
}

int gcb_peratom_dist(gcb_traj* tb, long frame, const double* ref_aos, int natoms_ref, const int* idx_test,
                     const int* idx_ref, int nidx, double* out) {
    int rc = check_window(tb, frame, 1, true);
    if (rc) return rc;
    gcb_traj_ext* t = X(tb);
    if (!ref_aos || !out) { set_error("NULL template or output"); return GCB_ERR_ARG; }
    const bool all = (idx_test == nullptr || nidx == 0);
    if (all && natoms_ref != t->natoms) { set_error("chem.memMSD: Ill formed matrices for memMSD calculation"); return GCB_ERR_SHAPE; }
    const int n = all ? t->natoms : nidx;
    std::vector<int> it((size_t)n);
    std::vector<double> pairs(3 * (size_t)n);
    for (int k = 0; k < n; k++) {
        const int a = all ? k : idx_test[k];
        const int b = all ? k : (idx_ref ? idx_ref[k] : idx_test[k]);
        if (a < 0 || a >= t->natoms || b < 0 || b >= natoms_ref) { set_error("index %d of the lists is out of range", k); return GCB_ERR_INDEXES; }
        it[(size_t)k] = a;
        pairs[(size_t)k] = ref_aos[3 * (size_t)b];
        pairs[(size_t)n + k] = ref_aos[3 * (size_t)b + 1];
        pairs[2 * (size_t)n + k] = ref_aos[3 * (size_t)b + 2];
    }
    GCB_CUDA(cudaSetDevice(t->dev));
    t->cache.valid = false;  // the gather buffers are shared with the template cache
    GCB_CUDA(cudaStreamSynchronize(t->stream));
    if (n > t->idx_cap) {
        cudaFree(t->d_refpairs); cudaFree(t->d_idx);
        t->d_refpairs = nullptr; t->d_idx = nullptr; t->idx_cap = 0;
        GCB_CUDA(cudaMalloc(&t->d_refpairs, sizeof(double) * 3 * (size_t)n));
        GCB_CUDA(cudaMalloc(&t->d_idx, sizeof(int) * (size_t)n));
        t->idx_cap = n;
    }
    if ((size_t)n > t->pdist_cap) {   // output scratch kept in the handle: no allocation per call
        cudaFree(t->d_pdist);
        t->d_pdist = nullptr; t->pdist_cap = 0;
        GCB_CUDA(cudaMalloc(&t->d_pdist, sizeof(double) * (size_t)n));
        t->pdist_cap = (size_t)n;
    }
    double* d_out = t->d_pdist;
    cudaMemcpyAsync(t->d_refpairs, pairs.data(), sizeof(double) * pairs.size(), cudaMemcpyHostToDevice, t->stream);
    cudaMemcpyAsync(t->d_idx, it.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, t->stream);
    rc = deps_before_compute(t, frame, 1);
    if (rc) return rc;
    rc = launch_peratom_dist(t, t->d_frames + frame * 3L * t->np, t->d_refpairs, t->d_idx, n, d_out);
    if (rc == GCB_OK) cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, t->stream);
    cudaError_t e = cudaStreamSynchronize(t->stream);
    if (rc) return rc;
    if (e != cudaSuccess) { set_error("per-atom distance kernel failed: %s", cudaGetErrorString(e)); return GCB_ERR_CUDA; }
    return GCB_OK;
}

int gcb_timer_start(gcb_traj* t) {
    if (!t) return GCB_ERR_ARG;
    GCB_CUDA(cudaSetDevice(t->dev));
    GCB_CUDA(cudaEventRecord(t->ev0, t->stream));
    return GCB_OK;
}

int gcb_timer_stop(gcb_traj* t, float* ms) {
    if (!t || !ms) return GCB_ERR_ARG;
    GCB_CUDA(cudaSetDevice(t->dev));
    GCB_CUDA(cudaEventRecord(t->ev1, t->stream));
    GCB_CUDA(cudaEventSynchronize(t->ev1));
    GCB_CUDA(cudaEventElapsedTime(ms, t->ev0, t->ev1));
    return GCB_OK;
}

int gcb_sync(gcb_traj* t) {
    if (!t) return GCB_ERR_ARG;
    GCB_CUDA(cudaSetDevice(t->dev));
    GCB_CUDA(cudaStreamSynchronize(t->stream));
    GCB_CUDA(cudaStreamSynchronize(t->copy_stream[0]));
    GCB_CUDA(cudaStreamSynchronize(t->copy_stream[1]));
    t->up_pending[0] = t->up_pending[1] = false;
    t->use_valid[0] = t->use_valid[1] = false;
    t->any_lo = t->any_hi = 0;
    return GCB_OK;
}

int gcb_synth_fill(gcb_traj* tb, long first, long nframes, const double* ref_aos, const double* sigma_atom,
                   unsigned long long seed, long frame_id0, int frame0_is_ref, int through_f32) {
    int rc = check_window(tb, first, nframes, false);
    if (rc) return rc;
    if (!ref_aos || !sigma_atom) { set_error("NULL reference or sigma"); return GCB_ERR_ARG; }
    gcb_traj_ext* t = X(tb);
    GCB_CUDA(cudaSetDevice(t->dev));
    double *d_ref = nullptr, *d_sig = nullptr;
    GCB_CUDA(cudaMalloc(&d_ref, sizeof(double) * 3 * (size_t)t->natoms));
    GCB_CUDA(cudaMalloc(&d_sig, sizeof(double) * (size_t)t->natoms));
    GCB_CUDA(cudaMemcpyAsync(d_ref, ref_aos, sizeof(double) * 3 * (size_t)t->natoms, cudaMemcpyHostToDevice, t->stream));
    GCB_CUDA(cudaMemcpyAsync(d_sig, sigma_atom, sizeof(double) * (size_t)t->natoms, cudaMemcpyHostToDevice, t->stream));
    rc = launch_synth(t, first, nframes, d_ref, d_sig, seed, frame_id0, frame0_is_ref, through_f32);
    cudaStreamSynchronize(t->stream);
    cudaFree(d_ref);
    cudaFree(d_sig);
    if (rc) return rc;
    GCB_CUDA(cudaGetLastError());
    if (first + nframes > t->nframes) t->nframes = first + nframes;
    return GCB_OK;
}

}  // extern "C"

// exported for comm.cu
namespace gcb {
int peratom_local(gcb_traj* tb, long first, long nframes, const double* ref_aos, const int* idx, int nidx, int write_back) {
    gcb_traj_ext* t = X(tb);
    int rows = 0;
    int rc = launch_super(t, first, nframes, ref_aos, t->natoms, idx, idx, nidx, write_back, true, &rows);
    if (rc) return rc;
    if (nframes == 0) {
        GCB_CUDA(cudaMemsetAsync(t->d_devsum, 0, sizeof(double) * (size_t)t->np, t->stream));
        return GCB_OK;
    }
    return launch_dev_reduce(t, rows, t->d_devsum);
}
int peratom_finish(gcb_traj* tb) { return check_status(X(tb)); }

// Would a per-atom-distance pass with a subset fit go to the cluster kernel (launch_super_impl's choice)?  Only that path can run
// on a template prepared on the device (gcb_lovo_resident).
bool peratom_prepared_supported(gcb_traj* tb) {
    gcb_traj_ext* t = X(tb);
    const int variant = t->tune.variant;
    if (variant == GCB_KERNEL_GENERIC) return false;
    const int split_dev_max = t->tune.split_dev_max < 0 ? kSplitDevMaxDefault : t->tune.split_dev_max;
    if (variant == GCB_KERNEL_AUTO && (t->natoms <= split_dev_max || small_kernel_supported(t, true))) return false;
    return cluster_kernel_supported(t, true);
}

// The pass of peratom_local over the template lovo_controller_kernel left in d_refc / d_weight (n_fit atoms in the fit): no host
// preparation, no upload.  The template cache no longer describes the device buffers afterwards.
int peratom_prepared(gcb_traj* tb, long first, long nframes, int n_fit) {
    gcb_traj_ext* t = X(tb);
    if (n_fit <= 0) { set_error("a fit on no atoms"); return GCB_ERR_ARG; }
    t->cache.valid = false;
    t->cache.n_fit = (double)n_fit;
    t->cache.bbar[0] = t->cache.bbar[1] = t->cache.bbar[2] = 0.0;   // only the write-back adds it; this pass never writes back
    t->cache.gb = 0.0;                                              // only the one-pass RMSD reads it; this pass never asks for one
    if (nframes == 0) {
        GCB_CUDA(cudaMemsetAsync(t->d_devsum, 0, sizeof(double) * (size_t)t->np, t->stream));
        return GCB_OK;
    }
    int rc = deps_before_compute(t, first, nframes);
    if (rc) return rc;
    SuperParams p;
    fill_params(t, first, nframes, 0, &p);
    // A fit on every atom (the first LOVO iteration when all atoms are candidates) needs no mask: the unmasked instance does the
    // same arithmetic (a mask with every bit set selects nothing away) without the per-atom selects.
    const bool weighted = n_fit != t->natoms;
    p.weight = weighted ? t->d_weight : nullptr;
    const size_t rows = (size_t)t->sm_count * 8;
    rc = ensure_devpart(t, rows);
    if (rc) return rc;
    GCB_CUDA(cudaMemsetAsync(t->d_devpart, 0, sizeof(double) * rows * (size_t)t->np, t->stream));
    p.devpart = t->d_devpart;
    p.rmsd = nullptr;
    int rows_used = 0;
    rc = launch_super_cluster(t, p, weighted, true, &rows_used);
    if (rc) return rc;
    rc = deps_after_compute(t, first, nframes);
    if (rc) return rc;
    return launch_dev_reduce(t, rows_used, t->d_devsum);
}

// per-atom sums for the atoms of `want` only: d_devsum[k] belongs to want[k]
int peratom_local_atoms(gcb_traj* tb, long first, long nframes, const double* ref_aos, const int* idx, int nidx, const int* want, int nwant) {
    gcb_traj_ext* t = X(tb);
    if (!want || nwant <= 0 || nwant > t->np) { set_error("the list of atoms asked for is empty or longer than a frame"); return GCB_ERR_ARG; }
    for (int k = 0; k < nwant; k++)
        if (want[k] < 0 || want[k] >= t->natoms) { set_error("atom %d of the list asked for is out of range (%d)", k, want[k]); return GCB_ERR_INDEXES; }
    // the fit: R and trans1 of every frame (frames untouched)
    int rc = launch_super(t, first, nframes, ref_aos, t->natoms, idx, idx, nidx, 0, false, nullptr);
    if (rc) return rc;
    if (nframes == 0) {
        GCB_CUDA(cudaMemsetAsync(t->d_devsum, 0, sizeof(double) * (size_t)t->np, t->stream));
        return GCB_OK;
    }
    if (nwant > t->want_cap) {
        GCB_CUDA(cudaStreamSynchronize(t->stream));
        cudaFree(t->d_want); cudaFree(t->d_wantref);
        t->d_want = nullptr; t->d_wantref = nullptr; t->want_cap = 0;
        GCB_CUDA(cudaMalloc(&t->d_want, sizeof(int) * (size_t)nwant));
        GCB_CUDA(cudaMalloc(&t->d_wantref, sizeof(double) * 3 * (size_t)nwant));
        t->want_cap = nwant;
    }
    SuperParams p;
    fill_params(t, first, nframes, 0, &p);
    std::vector<double> bw(3 * (size_t)nwant);
    for (int k = 0; k < nwant; k++) {
        const size_t i = (size_t)want[k];
        bw[(size_t)k] = ref_aos[3 * i] - p.bbar[0];
        bw[(size_t)nwant + k] = ref_aos[3 * i + 1] - p.bbar[1];
        bw[2 * (size_t)nwant + k] = ref_aos[3 * i + 2] - p.bbar[2];
    }
    GCB_CUDA(cudaMemcpyAsync(t->d_want, want, sizeof(int) * (size_t)nwant, cudaMemcpyHostToDevice, t->stream));
    GCB_CUDA(cudaMemcpyAsync(t->d_wantref, bw.data(), sizeof(double) * bw.size(), cudaMemcpyHostToDevice, t->stream));
    GCB_CUDA(cudaStreamSynchronize(t->stream));   // bw and want are pageable host memory of this call
    // partial rows are nwant long (padded to 16), not a whole frame: the workspace is counted in frame-sized rows
    const int stride = round_up(nwant, 16);
    const size_t part_doubles = (size_t)peratom_accum_gather_rows(t, nframes, nwant) * (size_t)stride;
    rc = ensure_devpart(t, (part_doubles + (size_t)t->np - 1) / (size_t)t->np);
    if (rc) return rc;
    p.devpart = t->d_devpart;
    int rows = 0;
    rc = launch_peratom_accum_gather(t, p, t->d_want, nwant, t->d_wantref, stride, &rows);
    if (rc) return rc;
    return launch_dev_reduce_n(t, rows, stride, nwant, t->d_devsum);
}
}  // namespace gcb
