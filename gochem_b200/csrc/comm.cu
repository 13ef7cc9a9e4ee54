// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink.
//
// The reference has no collectives; frames shard across ranks with no data-path
// exchange.  The one real exchange on this path is LOVO's per-iteration sum of the per-atom
// deviation vector (what align.RMSDTraj accumulates over all frames, /root/reference/align/lovo.go:
// 332-336): every rank contributes the sums of its frame shard.  It is done as an all-gather of the
// per-rank vectors followed by a rank-ordered sum, so every rank holds bit-identical totals (the
// index sets LOVO derives from them must agree exactly across ranks).
//
// NCCL is bound with dlopen at first use so the library itself loads on hosts without NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace gcb {
int peratom_local(gcb_traj* t, long first, long nframes, const double* ref_aos, const int* idx, int nidx, int write_back);
int peratom_finish(gcb_traj* t);
int peratom_local_atoms(gcb_traj* t, long first, long nframes, const double* ref_aos, const int* idx, int nidx, const int* want, int nwant);

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))dlsym(api.h, "ncclAllGather");
            api.Broadcast = (decltype(api.Broadcast))dlsym(api.h, "ncclBroadcast");
            api.GroupStart = (decltype(api.GroupStart))dlsym(api.h, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.h, "ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
        }
    });
    if (!api.h || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.Broadcast || !api.GroupStart ||
        !api.GroupEnd) {
        set_error("NCCL (libnccl.so.2) is not loadable: %s", api.h ? "missing symbols" : dlerror());
        return nullptr;
    }
    return &api;
}

#define GCB_NCCL(api, call)                                                                   \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            gcb::set_error("%s failed: %s", #call, (api)->GetErrorString ? (api)->GetErrorString(r__) : "?"); \
            return GCB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

// out[i] = sum over ranks (ascending) of gathered[r][i]
__global__ void __launch_bounds__(256) rank_ordered_sum_kernel(const double* __restrict__ gathered, int nranks, long n,
                                                              double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int r = 0; r < nranks; r++) s += gathered[(long)r * n + i];
    out[i] = s;
}

}  // namespace gcb

using namespace gcb;

static int comm_reserve(gcb_comm* c, long n) {
    if (n <= c->cap) return GCB_OK;
    cudaFree(c->d_send); cudaFree(c->d_gather);
    c->d_send = c->d_gather = nullptr; c->cap = 0;
    GCB_CUDA(cudaMalloc(&c->d_send, sizeof(double) * (size_t)n));
    GCB_CUDA(cudaMalloc(&c->d_gather, sizeof(double) * (size_t)n * (size_t)c->nranks));
    c->cap = n;
    return GCB_OK;
}

namespace gcb {

// ---- the small all-reduce as ONE kernel over peer-mapped memory ----
// LOVO's exchange is 8 bytes per atom and rank (40 KB at 5000 atoms): far below the size at which a library collective's
// bandwidth matters, so what counts is latency.  Every rank stores its vector straight into a mailbox slot of every peer
// (plain stores through the cudaIpc mapping, i.e. NVLink writes), publishes a sequence number with a system-scope release,
// waits for the peers' numbers with system-scope acquires, and adds the nranks vectors of its own mailbox in rank order --
// bit-identical totals on every rank, no NCCL call, no second kernel.  Two parities: a rank that has started exchange s + 1
// has finished reading exchange s, and nobody can reach s + 2 before everybody has started s + 1.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct MailPtrs { double* box[8]; };

// A peer that never arrives (its process died before the call) must not hang this GPU: the wait gives up after timeout_ns,
// fills the vector with NaN and leaves GCB_ERR_CUDA in *status (the trajectory's status word, read at the end of every pass).
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(1024) peer_exchange_sum_kernel(MailPtrs mp, int nranks, int rank, long cap, long n, unsigned long long seq,
                                                                  double* __restrict__ vec, unsigned long long timeout_ns, int* status) {
    __shared__ int lost;
    if (threadIdx.x == 0) lost = 0;
    const int par = (int)(seq & 1ull);
    const size_t slot = ((size_t)par * nranks + rank) * (size_t)cap;       // where this rank's vector goes in every box
    for (int r = 0; r < nranks; r++) {
        double* dst = mp.box[r] + slot;
        for (long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vec[i];
    }
    __threadfence_system();
    __syncthreads();
    const size_t flags_off = (size_t)2 * nranks * (size_t)cap;              // doubles in front of the flags
    if (threadIdx.x < nranks) {
        unsigned long long* peer_flags = reinterpret_cast<unsigned long long*>(mp.box[threadIdx.x] + flags_off);
        st_release_sys(peer_flags + (size_t)par * nranks + rank, seq);
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(mp.box[rank] + flags_off) + (size_t)par * nranks + threadIdx.x;
        const unsigned long long t0 = global_ns();
        unsigned int spins = 0;
        while (ld_acquire_sys(mine) < seq) {
            if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) { lost = 1; break; }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (lost) {
        for (long i = threadIdx.x; i < n; i += blockDim.x) vec[i] = __longlong_as_double(0x7ff8000000000000LL);
        if (threadIdx.x == 0 && status) atomicExch(status, (int)GCB_ERR_CUDA);
        return;
    }
    const double* box = mp.box[rank] + (size_t)par * nranks * (size_t)cap;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < nranks; r++) s += __ldcg(box + (size_t)r * cap + i);   // rank order; L2, never a stale L1 line
        vec[i] = s;
    }
}

}  // namespace gcb
using namespace gcb;

// Collective: (re)allocate the mailboxes for vectors of up to n doubles and map every peer's.  Returns GCB_OK with
// c->mail_state = 1, or leaves mail_state = -1 when the peers' memory cannot be mapped (the caller uses NCCL then).
static int mail_setup(gcb_comm* c, long n) {
    NcclApi* api = nccl();
    if (!api) return GCB_ERR_CUDA;
    GCB_CUDA(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < c->nranks; r++)
        if (r != c->rank && c->mail[r]) { cudaIpcCloseMemHandle(c->mail[r]); c->mail[r] = nullptr; }
    int rc = comm_barrier(c, c->stream);       // every rank has dropped its mappings before anybody frees a box
    if (rc) return rc;
    GCB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->mail[c->rank]);
    c->mail[c->rank] = nullptr;
    const long cap = (n + 511) / 512 * 512;
    const size_t bytes = sizeof(double) * 2 * (size_t)c->nranks * (size_t)cap + sizeof(unsigned long long) * 2 * (size_t)c->nranks;
    void* mine = nullptr;
    GCB_CUDA(cudaMalloc(&mine, bytes));
    GCB_CUDA(cudaMemset(mine, 0, bytes));
    c->mail[c->rank] = (double*)mine;
    cudaIpcMemHandle_t h;
    GCB_CUDA(cudaIpcGetMemHandle(&h, mine));
    unsigned char* d_h = nullptr;
    GCB_CUDA(cudaMalloc(&d_h, 64 * (size_t)(c->nranks + 1)));
    GCB_CUDA(cudaMemcpyAsync(d_h, &h, 64, cudaMemcpyHostToDevice, c->stream));
    GCB_NCCL(api, api->AllGather(d_h, d_h + 64, 64, ncclChar, c->comm, c->stream));
    std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
    GCB_CUDA(cudaMemcpyAsync(all.data(), d_h + 64, 64 * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->stream));
    GCB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_h);
    int ok = 1;
    for (int r = 0; r < c->nranks && ok; r++) {
        if (r == c->rank) continue;
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        c->mail[r] = (double*)p;
    }
    // all ranks must agree on the path: one rank without peer access sends everybody to NCCL
    int* d_ok = nullptr;
    GCB_CUDA(cudaMalloc(&d_ok, sizeof(int) * (size_t)(c->nranks + 1)));
    GCB_CUDA(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    GCB_NCCL(api, api->AllGather(d_ok, d_ok + 1, sizeof(int), ncclChar, c->comm, c->stream));
    std::vector<int> oks((size_t)c->nranks);
    GCB_CUDA(cudaMemcpyAsync(oks.data(), d_ok + 1, sizeof(int) * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->stream));
    GCB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_ok);
    for (int v : oks) ok = ok && v;
    c->mail_cap = cap;
    c->mail_seq = 0;
    c->mail_state = ok ? 1 : -1;
    return GCB_OK;
}

// device vector d_vec[n] (on stream s) -> same vector summed over ranks in rank order
static int allreduce_device(gcb_comm* c, double* d_vec, long n, cudaStream_t s, int* d_status = nullptr) {
    static const unsigned long long timeout_ns = [] {
        const char* e = getenv("GCB_EXCHANGE_TIMEOUT_S");
        const double sec = e ? atof(e) : 30.0;
        return (unsigned long long)((sec > 0.0 ? sec : 30.0) * 1e9);
    }();
    static const bool force_nccl = getenv("GCB_ALLREDUCE_NCCL") != nullptr;   // the library all-gather path, for comparison
    if (!force_nccl && c->nranks <= 8 && (c->mail_state == 0 || (c->mail_state == 1 && n > c->mail_cap))) {
        GCB_CUDA(cudaStreamSynchronize(s));   // set-up is collective and rare; the exchange itself never synchronises
        int rc = mail_setup(c, n);
        if (rc) return rc;
    }
    if (!force_nccl && c->mail_state == 1) {
        MailPtrs mp;
        for (int r = 0; r < 8; r++) mp.box[r] = r < c->nranks ? c->mail[r] : nullptr;
        c->mail_seq++;
        peer_exchange_sum_kernel<<<1, 1024, 0, s>>>(mp, c->nranks, c->rank, c->mail_cap, n, c->mail_seq, d_vec, timeout_ns, d_status);
        GCB_LAUNCHED();
        GCB_CUDA(cudaGetLastError());
        return GCB_OK;
    }
    NcclApi* api = nccl();
    if (!api) return GCB_ERR_CUDA;
    int rc = comm_reserve(c, n);
    if (rc) return rc;
    GCB_NCCL(api, api->AllGather(d_vec, c->d_gather, (size_t)n, ncclDouble, c->comm, s));
    rank_ordered_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->d_gather, c->nranks, n, d_vec);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

namespace gcb {

// Completes when every rank has reached this point of its stream s (a one-byte-per-rank all-gather).
int comm_barrier(gcb_comm* c, cudaStream_t s) {
    if (c->nranks == 1) return GCB_OK;
    NcclApi* api = nccl();
    if (!api) return GCB_ERR_CUDA;
    if (!c->d_flags) {
        GCB_CUDA(cudaMalloc(&c->d_flags, 16 + 16 * (size_t)c->nranks));
        GCB_CUDA(cudaMemset(c->d_flags, 0, 16 + 16 * (size_t)c->nranks));
    }
    GCB_NCCL(api, api->AllGather(c->d_flags, c->d_flags + 16, 16, ncclChar, c->comm, s));
    return GCB_OK;
}

// Collective: every rank ends up with a device block of at least `bytes` bytes of its own and a mapping of every
// other rank's block (cudaIpcOpenMemHandle: peer access over NVLink).  blk[r] = rank r's block as seen from here.
// The blocks are kept in the communicator and reused while they are large enough.
int comm_peer_blocks(gcb_comm* c, size_t bytes, double** blk) {
    if (bytes > c->blk_bytes) {
        NcclApi* api = c->nranks > 1 ? nccl() : nullptr;
        if (c->nranks > 1 && !api) return GCB_ERR_CUDA;
        GCB_CUDA(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < c->nranks; r++)
            if (r != c->rank && c->peer_blk[r]) { cudaIpcCloseMemHandle(c->peer_blk[r]); c->peer_blk[r] = nullptr; }
        if (c->nranks > 1) {   // every rank has dropped its mappings of the old blocks before anybody frees one
            int rc = comm_barrier(c, c->stream);
            if (rc) return rc;
            GCB_CUDA(cudaStreamSynchronize(c->stream));
        }
        cudaFree(c->peer_blk[c->rank]);
        c->peer_blk[c->rank] = nullptr;
        c->blk_bytes = 0;
        void* mine = nullptr;
        GCB_CUDA(cudaMalloc(&mine, bytes));
        c->peer_blk[c->rank] = (double*)mine;
        if (c->nranks > 1) {
            cudaIpcMemHandle_t h;
            GCB_CUDA(cudaIpcGetMemHandle(&h, mine));
            static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
            unsigned char* d_h = nullptr;
            GCB_CUDA(cudaMalloc(&d_h, 64 * (size_t)(c->nranks + 1)));
            GCB_CUDA(cudaMemcpyAsync(d_h, &h, 64, cudaMemcpyHostToDevice, c->stream));
            GCB_NCCL(api, api->AllGather(d_h, d_h + 64, 64, ncclChar, c->comm, c->stream));
            std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
            GCB_CUDA(cudaMemcpyAsync(all.data(), d_h + 64, 64 * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->stream));
            GCB_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(d_h);
            for (int r = 0; r < c->nranks; r++) {
                if (r == c->rank) continue;
                void* p = nullptr;
                cudaError_t e = cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    set_error("cudaIpcOpenMemHandle of rank %d's block failed: %s (peer access over NVLink is required)", r, cudaGetErrorString(e));
                    return GCB_ERR_CUDA;
                }
                c->peer_blk[r] = (double*)p;
            }
        }
        c->blk_bytes = bytes;
    }
    for (int r = 0; r < c->nranks; r++) blk[r] = c->peer_blk[r];
    return GCB_OK;
}

}  // namespace gcb

extern "C" {

// Frames sharded over the ranks -> every rank holds all of them: rank r's counts[r] frames sit at frame offset
// counts[0] + ... + counts[r-1] of its own trajectory buffer; one grouped broadcast per rank fills the rest over NVLink.
int gcb_traj_allgather(gcb_traj* t, gcb_comm* c, const long* counts) {
    if (!t || !c || !counts) { set_error("bad all-gather arguments"); return GCB_ERR_ARG; }
    if (c->dev != t->dev) { set_error("communicator and trajectory live on different devices"); return GCB_ERR_ARG; }
    long total = 0;
    for (int r = 0; r < c->nranks; r++) { if (counts[r] < 0) { set_error("negative frame count"); return GCB_ERR_ARG; } total += counts[r]; }
    if (total > t->cap) { set_error("all-gather of %ld frames into a trajectory of capacity %ld", total, t->cap); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    { int rc = deps_before_compute(t, 0, total); if (rc) return rc; }
    if (c->nranks > 1) {
        NcclApi* api = nccl();
        if (!api) return GCB_ERR_CUDA;
        const size_t fd = 3 * (size_t)t->np;
        GCB_NCCL(api, api->GroupStart());
        long off = 0;
        for (int r = 0; r < c->nranks; r++) {
            double* p = t->d_frames + (size_t)off * fd;
            if (counts[r] > 0) GCB_NCCL(api, api->Broadcast(p, p, (size_t)counts[r] * fd, ncclDouble, r, c->comm, t->stream));
            off += counts[r];
        }
        GCB_NCCL(api, api->GroupEnd());
        GCB_CUDA(cudaStreamSynchronize(t->stream));
    }
    t->nframes = total;
    return GCB_OK;
}

int gcb_comm_unique_id(char* id128) {
    if (!id128) return GCB_ERR_ARG;
    NcclApi* api = nccl();
    if (!api) return GCB_ERR_CUDA;
    ncclUniqueId id;
    GCB_NCCL(api, api->GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return GCB_OK;
}

int gcb_comm_init(int dev, int nranks, int rank, const char* id128, gcb_comm** out) {
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad communicator arguments"); return GCB_ERR_ARG; }
    *out = nullptr;
    NcclApi* api = nccl();
    if (!api) return GCB_ERR_CUDA;
    GCB_CUDA(cudaSetDevice(dev));
    gcb_comm* c = new gcb_comm();
    c->dev = dev; c->nranks = nranks; c->rank = rank;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = api->CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
        delete c;
        return GCB_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        api->CommDestroy(c->comm);
        delete c;
        return GCB_ERR_CUDA;
    }
    *out = c;
    return GCB_OK;
}

int gcb_comm_destroy(gcb_comm* c) {
    if (!c) return GCB_OK;
    cudaSetDevice(c->dev);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    cudaFree(c->d_send); cudaFree(c->d_gather); cudaFree(c->d_flags);
    for (int r = 0; r < c->nranks; r++) {
        if (!c->peer_blk[r]) continue;
        if (r == c->rank) cudaFree(c->peer_blk[r]); else cudaIpcCloseMemHandle(c->peer_blk[r]);
    }
    for (int r = 0; r < c->nranks; r++) {
        if (!c->mail[r]) continue;
        if (r == c->rank) cudaFree(c->mail[r]); else cudaIpcCloseMemHandle(c->mail[r]);
    }
    NcclApi* api = nccl();
    if (api && c->comm) api->CommDestroy(c->comm);
    delete c;
    return GCB_OK;
}

int gcb_allreduce_sum_f64(gcb_comm* c, double* host_vec, long n) {
    if (!c || !host_vec || n <= 0) { set_error("bad allreduce arguments"); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(c->dev));
    int rc = comm_reserve(c, n);
    if (rc) return rc;
    GCB_CUDA(cudaMemcpyAsync(c->d_send, host_vec, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    rc = allreduce_device(c, c->d_send, n, c->stream);
    if (rc) return rc;
    GCB_CUDA(cudaMemcpyAsync(host_vec, c->d_send, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    GCB_CUDA(cudaStreamSynchronize(c->stream));
    return GCB_OK;
}

int gcb_peratom_dev_sum(gcb_traj* t, long first, long nframes, const double* ref_aos, const int* idx, int nidx,
                        int write_back, gcb_comm* comm, double* sum) {
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    if (first < 0 || nframes < 0 || first + nframes > t->nframes) { set_error("frame window outside the trajectory"); return GCB_ERR_ARG; }
    if (!sum) { set_error("NULL output"); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    static const bool timing = getenv("GCB_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = peratom_local(t, first, nframes, ref_aos, idx, nidx, write_back);
    if (rc) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    if (comm && comm->nranks > 1) {
        if (comm->dev != t->dev) { set_error("communicator and trajectory live on different devices"); return GCB_ERR_ARG; }
        rc = allreduce_device(comm, t->d_devsum, t->natoms, t->stream, t->d_status);
        if (rc) return rc;
    }
    GCB_CUDA(cudaMemcpyAsync(sum, t->d_devsum, sizeof(double) * (size_t)t->natoms, cudaMemcpyDeviceToHost, t->stream));
    rc = peratom_finish(t);
    if (timing) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gcb] peratom_dev_sum: launch side %.1f us, until results on the host %.1f us\n",
                std::chrono::duration<double, std::micro>(t1 - t0).count(), std::chrono::duration<double, std::micro>(t2 - t0).count());
    }
    return rc;
}

// align.LOVO's whole fixed-point iteration (align/lovo.go:212-252) over resident frames with the bookkeeping on the device
// (lovo_ctl.cu): per iteration one fused pass, the cross-rank sum, one controller CTA, and 64 bytes back to the host.
// Returns GCB_OK, an error, or 1 = "not taken" (shape or tuning that the device loop does not cover, or values the host mirror
// must sort): the caller then runs the loop on the host with gcb_peratom_dev_sum -- same answers, bit for bit.
int gcb_lovo_resident(gcb_traj* t, long first, long nframes, const double* ref_aos, const int* cand, int ncand, double total_frames,
                      double less_than_rmsd, int minimum_n, int max_iter, gcb_comm* comm, int* n_out, int* iterations_out,
                      int* disagreement_out, int* indexesold_out, int* ids_out, double* vals_out) {
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    if (first < 0 || nframes < 0 || first + nframes > t->nframes) { set_error("frame window outside the trajectory"); return GCB_ERR_ARG; }
    if (!ref_aos || !cand || ncand <= 0 || !n_out || !iterations_out || !indexesold_out || !ids_out || !vals_out) {
        set_error("bad LOVO arguments");
        return GCB_ERR_ARG;
    }
    if (ncand > lovo_ctl_slots() || 8L * ncand <= t->natoms || !(total_frames > 0.0)) return 1;   // big candidate lists sort on the host; sparse ones read their sectors only
    for (int k = 0; k < ncand; k++) {
        if (cand[k] < 0 || cand[k] >= t->natoms) { set_error("candidate atom %d is out of range (%d)", k, cand[k]); return GCB_ERR_INDEXES; }
        if (k > 0 && cand[k] <= cand[k - 1]) return 1;                                // the mirror pairs values and candidates by position
    }
    GCB_CUDA(cudaSetDevice(t->dev));
    const auto t_enter = std::chrono::steady_clock::now();
    if (!peratom_prepared_supported(t)) return 1;
    if (comm && comm->nranks > 1 && comm->dev != t->dev) { set_error("communicator and trajectory live on different devices"); return GCB_ERR_ARG; }
    // workspace of this handle: [state | cand | old | ids | vals | template copy]
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t oSt = carve(sizeof(LovoDevState)), oCand = carve(sizeof(int) * (size_t)ncand), oOld = carve(sizeof(int) * (size_t)ncand),
                 oIds = carve(sizeof(int) * (size_t)ncand), oVals = carve(sizeof(double) * (size_t)ncand),
                 oRef = carve(sizeof(double) * 3 * (size_t)t->natoms);
    if (off > t->lovo_ws_bytes) {
        GCB_CUDA(cudaStreamSynchronize(t->stream));
        cudaFree(t->d_lovo_ws);
        t->d_lovo_ws = nullptr;
        t->lovo_ws_bytes = 0;
        GCB_CUDA(cudaMalloc(&t->d_lovo_ws, off));
        t->lovo_ws_bytes = off;
    }
    if (!t->h_lovo) GCB_CUDA(cudaHostAlloc(&t->h_lovo, 2 * sizeof(LovoDevState) + 768 + (3 * sizeof(int) + sizeof(double)) * (size_t)lovo_ctl_slots(), cudaHostAllocMapped));
    unsigned char* w = (unsigned char*)t->d_lovo_ws;
    LovoDevState* hst = (LovoDevState*)t->h_lovo;            // [0]: what the controller reports; [1]: the initial state, staged
    int* h_iota = (int*)(hst + 2);
    LovoCtl c = {};
    c.devsum = t->d_devsum;
    c.total = total_frames;
    c.cand = (const int*)(w + oCand);
    c.ncand = ncand;
    c.old = (int*)(w + oOld);
    c.ids = (int*)(w + oIds);
    c.vals = (double*)(w + oVals);
    c.st = (LovoDevState*)(w + oSt);
    c.host = hst;
    c.less_than = less_than_rmsd;
    c.minimum_n = minimum_n;
    c.ref_aos = (const double*)(w + oRef);
    c.natoms = t->natoms;
    c.np = t->np;
    c.refc = t->d_refc;
    c.weight = t->d_weight;
    c.status = t->d_status;
    GCB_CUDA(cudaStreamSynchronize(t->stream));              // the staging words below may still feed an earlier run's copies
    memset(&hst[1], 0, sizeof(LovoDevState));
    hst[1].n = ncand;                                        // lovo.go:212-213: indexesold = fullindexes, n = len
    for (int k = 0; k < ncand; k++) h_iota[k] = k;
    GCB_CUDA(cudaMemcpyAsync(c.st, &hst[1], sizeof(LovoDevState), cudaMemcpyHostToDevice, t->stream));
    GCB_CUDA(cudaMemcpyAsync(c.old, h_iota, sizeof(int) * (size_t)ncand, cudaMemcpyHostToDevice, t->stream));
    GCB_CUDA(cudaMemcpyAsync((void*)c.cand, cand, sizeof(int) * (size_t)ncand, cudaMemcpyHostToDevice, t->stream));
    GCB_CUDA(cudaMemcpyAsync((void*)c.ref_aos, ref_aos, sizeof(double) * 3 * (size_t)t->natoms, cudaMemcpyHostToDevice, t->stream));
    int rc = GCB_OK;
    int n_fit = ncand;
    static const bool timing = getenv("GCB_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    const auto tl0 = now();
    for (int it = 0;; it++) {
        const auto ti0 = now();
        // every pass fits on the template the controller left on the device: before the first one that of the whole candidate list
        if (it == 0) {
            rc = launch_lovo_first_template(t, c);
            if (rc) return rc;
        }
        rc = peratom_prepared(t, first, nframes, n_fit);
        if (rc) return rc;
        if (comm && comm->nranks > 1) {
            rc = allreduce_device(comm, t->d_devsum, t->natoms, t->stream, t->d_status);
            if (rc) return rc;
        }
        const auto ti1 = now();
        rc = launch_lovo_controller(t, c);
        if (rc) return rc;
        GCB_CUDA(cudaStreamSynchronize(t->stream));
        const LovoDevState st = hst[0];
        if (timing) fprintf(stderr, "[gcb] lovo_resident: iteration %d: launches %.1f us, until the state is back %.1f us (setup before the loop %.1f us)\n",
                            it, us(ti0, ti1), us(ti0, now()), it == 0 ? us(t_enter, tl0) : 0.0);
        if (st.dev_status != 0) return peratom_finish(t);     // SVD failure or exchange timeout: the usual message, status word cleared
        if (st.err == LOVO_ERR_VALUES) return 1;   // nothing of the template was touched
        if (st.err == LOVO_ERR_MINIMUM_N) {
            set_error("align.mobileLimit: fewer candidate atoms than MinimumN; the reference would loop forever here (lovo.go:177-186)");
            return GCB_ERR_ARG;
        }
        if (st.err == LOVO_ERR_BOUNDS) { set_error("align.LOVO: slice bounds out of range (n outside the candidate list)"); return GCB_ERR_ARG; }
        n_fit = st.n;
        if (st.converged) break;
        if (st.itercount >= max_iter) { set_error("align.LOVO: no convergence after %d iterations", st.itercount); return GCB_ERR_ARG; }
    }
    const LovoDevState st = hst[0];
    *n_out = st.n;
    *iterations_out = st.itercount;
    if (disagreement_out) *disagreement_out = st.disagreement;
    // candidate positions -> atoms; the three arrays sit next to one another in the workspace: one copy
    {
        const size_t span = oVals + sizeof(double) * (size_t)ncand - oOld;   // <= (4 + 4 + 8) * ncand + 2 * 255 bytes of padding
        unsigned char* back = (unsigned char*)(h_iota + lovo_ctl_slots());     // pinned: the copy is one short DMA
        GCB_CUDA(cudaMemcpyAsync(back, w + oOld, span, cudaMemcpyDeviceToHost, t->stream));
        GCB_CUDA(cudaStreamSynchronize(t->stream));
        const int* h_old = (const int*)back;
        const int* h_ids = (const int*)(back + (oIds - oOld));
        for (int k = 0; k < ncand; k++) { indexesold_out[k] = cand[h_old[k]]; ids_out[k] = cand[h_ids[k]]; }
        memcpy(vals_out, back + (oVals - oOld), sizeof(double) * (size_t)ncand);
    }
    if (timing) fprintf(stderr, "[gcb] lovo_resident: whole call %.1f us\n", us(t_enter, now()));
    return GCB_OK;
}

int gcb_peratom_dev_sum_atoms(gcb_traj* t, long first, long nframes, const double* ref_aos, const int* idx, int nidx, const int* want,
                              int nwant, gcb_comm* comm, double* sum_want) {
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    if (first < 0 || nframes < 0 || first + nframes > t->nframes) { set_error("frame window outside the trajectory"); return GCB_ERR_ARG; }
    if (!sum_want) { set_error("NULL output"); return GCB_ERR_ARG; }
    GCB_CUDA(cudaSetDevice(t->dev));
    int rc = peratom_local_atoms(t, first, nframes, ref_aos, idx, nidx, want, nwant);
    if (rc) return rc;
    if (comm && comm->nranks > 1) {
        if (comm->dev != t->dev) { set_error("communicator and trajectory live on different devices"); return GCB_ERR_ARG; }
        rc = allreduce_device(comm, t->d_devsum, nwant, t->stream, t->d_status);
        if (rc) return rc;
    }
    GCB_CUDA(cudaMemcpyAsync(sum_want, t->d_devsum, sizeof(double) * (size_t)nwant, cudaMemcpyDeviceToHost, t->stream));
    return peratom_finish(t);
}

}  // extern "C"
