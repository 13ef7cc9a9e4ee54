// lovo_ctl.cu -- align.LOVO's bookkeeping between two RMSDTraj passes, on the device.
//
// The reference decides the next fit set on the host (/root/reference/align/lovo.go:229-252): per-atom mean deviations ->
// rmsdFilter (:519-528) -> sort.Stable by value (:418-432) -> mobileLimit (:175-188) -> sameElementsInt / approxEq
// (:246-248, :436-457) -> disagreementInt (:459-471) -> indexesold = sorted ids.  The host mirror (host/gochem.cpp LovoState)
// does the same, which costs every iteration a 40 KB device->host copy, ~85 us of host work, a template upload and two stream
// synchronisations while the GPU idles -- a third of an iteration at 8 GPUs.  Here two small kernels do it where the sums
// already are (the order by counting ranks on every SM, the rest in one CTA) and leave the next pass's template (centred planes
// + 0/1 weights) in place; the host reads back 64 bytes per iteration.
//
// Everything is bit-identical to the host mirror: the same divisions, the same total order (value, then position: what a stable
// sort leaves), the same multiplications by 1.1, the centroid of the template subset summed sequentially in list order
// (chem.MassCenter's 1 x n times n x 3 product, geometric.go:684-690) and scaled by 1.0 / n.
// tests/test_gpu_lovo_device.py compares the two loops bit for bit.
#include "common.cuh"

namespace gcb {

namespace {

constexpr int kCtlThreads = 1024;
constexpr int kRankThreads = 256;
constexpr int kRankPerCta = 32;                // elements ranked by one CTA: four per warp
constexpr int kCtlSlots = 16384;               // the most candidates the device loop takes (128 KB of means in shared memory)
constexpr int kSumChunk = 2048;                // template rows gathered at a time for the sequential centroid sum

// Non-negative doubles order like their bit patterns: (value, position) compared on integers.
__device__ __forceinline__ bool key_less(unsigned long long ka, int ia, unsigned long long kb, int ib) { return ka < kb || (ka == kb && ia < ib); }

// Step 1, on every SM: mean k = sum[cand[k]] / total, and its place in the ascending (value, position) order -- what sort.Stable
// by value leaves (lovo.go:418-432) -- by counting: rank = number of candidates that come before.  ncand^2 integer comparisons
// spread over ncand / 32 CTAs take a few microseconds; a bitonic sort of the same list on one SM took 100 (shuffle- and
// issue-bound: one SM has to execute all of it).
__global__ void __launch_bounds__(kRankThreads) lovo_rank_kernel(const LovoCtl c) {
    extern __shared__ __align__(16) unsigned char rank_smem[];
    unsigned long long* key = reinterpret_cast<unsigned long long*>(rank_smem);   // [ncand]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncand = c.ncand;
    bool bad = false;
    for (int k = tid; k < ncand; k += kRankThreads) {
        double v = c.devsum[c.cand[k]] / c.total;
        if (!(v >= 0.0 && v <= 1.79e308)) bad = true;   // negative or NaN: the host mirror sorts those by another route
        v += 0.0;                                        // -0.0 -> +0.0
        key[k] = (unsigned long long)__double_as_longlong(v);
    }
    if (bad) atomicExch(&c.st->err, (int)LOVO_ERR_VALUES);
    __syncthreads();
    constexpr int kPerWarp = kRankPerCta / (kRankThreads / 32);
    const int i0 = blockIdx.x * kRankPerCta + warp * kPerWarp;
    unsigned long long ki[kPerWarp];
    int cnt[kPerWarp];
#pragma unroll
    for (int e = 0; e < kPerWarp; e++) { ki[e] = (i0 + e < ncand) ? key[i0 + e] : 0ull; cnt[e] = 0; }
    for (int j = lane; j < ncand; j += 32) {
        const unsigned long long kj = key[j];
#pragma unroll
        for (int e = 0; e < kPerWarp; e++) cnt[e] += key_less(kj, j, ki[e], i0 + e) ? 1 : 0;
    }
#pragma unroll
    for (int e = 0; e < kPerWarp; e++) {
        int r = cnt[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0 && i0 + e < ncand) { c.ids[r] = i0 + e; c.vals[r] = __longlong_as_double((long long)ki[e]); }
    }
}

// The template of a pass that fits on the candidates pos[0:n] (positions in cand), written where the fused kernel reads it:
// centroid of those template rows summed in list order (chem.MassCenter's 1 x n times n x 3 product, geometric.go:684-690: three
// threads, one coordinate each, sequential like the host) times 1.0 / n, centred planes of ALL atoms, weights 1 on the fit atoms.
// Called by every thread of a 1024-thread CTA; rows = 3 * kSumChunk doubles of shared memory.
__device__ void write_template(const LovoCtl& c, const int* pos, int n, double* rows, double* s_bbar) {
    const int tid = threadIdx.x;
    double run = 0.0;
    for (int k0 = 0; k0 < n; k0 += kSumChunk) {
        const int m = min(kSumChunk, n - k0);
        __syncthreads();
        for (int k = tid; k < m; k += kCtlThreads) {
            const long a = c.cand[pos[k0 + k]];
            rows[k] = c.ref_aos[3 * a];
            rows[kSumChunk + k] = c.ref_aos[3 * a + 1];
            rows[2 * kSumChunk + k] = c.ref_aos[3 * a + 2];
        }
        __syncthreads();
        if (tid < 3) {
            const double* v = rows + tid * kSumChunk;
            int k = 0;
            for (; k + 8 <= m; k += 8) {   // eight loads in flight, the additions in list order
                const double a0 = v[k], a1 = v[k + 1], a2 = v[k + 2], a3 = v[k + 3], a4 = v[k + 4], a5 = v[k + 5], a6 = v[k + 6], a7 = v[k + 7];
                run = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(run, a0), a1), a2), a3), a4), a5), a6), a7);
            }
            for (; k < m; k++) run = __dadd_rn(run, v[k]);
        }
    }
    if (tid < 3) s_bbar[tid] = __dmul_rn(run, 1.0 / (double)n);
    __syncthreads();
    const double b0 = s_bbar[0], b1 = s_bbar[1], b2 = s_bbar[2];
    for (int i = tid; i < c.np; i += kCtlThreads) {
        const bool in = i < c.natoms;
        c.refc[i] = in ? __dsub_rn(c.ref_aos[3L * i], b0) : 0.0;
        c.refc[c.np + i] = in ? __dsub_rn(c.ref_aos[3L * i + 1], b1) : 0.0;
        c.refc[2L * c.np + i] = in ? __dsub_rn(c.ref_aos[3L * i + 2], b2) : 0.0;
        c.weight[i] = 0.0;
    }
    __syncthreads();
    for (int k = tid; k < n; k += kCtlThreads) c.weight[c.cand[pos[k]]] = 1.0;
}

// Before the first pass: the fit is on the candidates as given (indexesold = fullindexes, lovo.go:212-213; old[] holds 0, 1, 2 ...).
__global__ void __launch_bounds__(kCtlThreads) lovo_first_template_kernel(const LovoCtl c) {
    extern __shared__ __align__(16) unsigned char ctl_smem[];
    __shared__ double s_bbar[3];
    write_template(c, c.old, c.ncand, reinterpret_cast<double*>(ctl_smem), s_bbar);
}

// Step 2, one CTA: everything else of lovo.go:238-252 on the sorted list, and the template of the next pass.
__global__ void __launch_bounds__(kCtlThreads) lovo_controller_kernel(const LovoCtl c) {
    extern __shared__ __align__(16) unsigned char ctl_smem[];
    double* rows = reinterpret_cast<double*>(ctl_smem);                           // [3][kSumChunk] gathered template rows
    double* key = rows + 3 * kSumChunk;                                           // [ncand] the sorted means (step 1), for the scans
    unsigned int* inold = reinterpret_cast<unsigned int*>(key + c.ncand);         // bitmap over candidate positions: indexesold[0:n]
    __shared__ int s_n, s_disagree, s_err, s_converged;
    __shared__ double s_newlim, s_bbar[3];
    const int tid = threadIdx.x;
    const int ncand = c.ncand;
    LovoDevState& st = *c.st;
    const int* pos = c.ids;            // sorted candidate positions (step 1)
    if (st.err == LOVO_ERR_VALUES) {   // step 1 met a value the host must sort
        if (tid == 0) { st.dev_status = *c.status; __threadfence_system(); *c.host = st; }
        return;
    }
    for (int k = tid; k < ncand; k += kCtlThreads) key[k] = c.vals[k];
    for (int k = tid; k < (ncand + 31) / 32; k += kCtlThreads) inold[k] = 0u;
    if (tid == 0) { s_disagree = 0; s_err = 0; s_converged = 0; }
    __syncthreads();
    // mobileLimit (lovo.go:175-188): n = (first position with value >= limit) - 1, or the length; limit *= 1.1 after every scan
    if (tid == 0) {
        int n = st.n;
        double limit = c.less_than;
        double newlim = 0.0;
        if (c.less_than > 0.0) {
            n = 0;
            long guard = 0;
            while (n < c.minimum_n) {
                int lo = 0, hi = ncand;             // lower bound: the linear scan's first hit on an ascending list
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (key[mid] >= limit) hi = mid; else lo = mid + 1;
                }
                n = lo < ncand ? lo - 1 : ncand;
                limit = limit * 1.1;
                if (++guard > 100000 || !(fabs(limit) <= 1.79e308)) { s_err = LOVO_ERR_MINIMUM_N; break; }
            }
            newlim = limit;
        }
        if (!s_err && (n < 0 || n > ncand)) s_err = LOVO_ERR_BOUNDS;
        s_n = n;
        s_newlim = newlim;
    }
    __syncthreads();
    if (s_err) {
        if (tid == 0) { st.itercount++; st.err = s_err; st.dev_status = *c.status; __threadfence_system(); *c.host = st; }
        return;
    }
    const int n = s_n;
    // sameElementsInt / disagreementInt of ids[0:n] against indexesold[0:n] (each candidate occurs once in either list)
    for (int k = tid; k < n; k += kCtlThreads) atomicOr(&inold[c.old[k] >> 5], 1u << (c.old[k] & 31));
    __syncthreads();
    int mine = 0;
    for (int k = tid; k < n; k += kCtlThreads) mine += ((inold[pos[k] >> 5] >> (pos[k] & 31)) & 1u) ? 0 : 1;
    if (mine) atomicAdd(&s_disagree, mine);
    __syncthreads();
    if (tid == 0) {
        const double newlim = s_newlim;
        const bool same = s_disagree == 0;
        const bool lim_ok = newlim == c.less_than || (fabs(newlim) - fabs(st.prevlim) <= 0.01);   // approxEq as written (:436-442)
        s_converged = (same && lim_ok) ? 1 : 0;
    }
    __syncthreads();
    const bool converged = s_converged != 0;
    if (!converged) {
        if (n == 0) {   // a fit on no atoms: the host mirror's dispatch decides what that means
            if (tid == 0) { st.itercount++; st.n = 0; st.err = LOVO_ERR_VALUES; st.dev_status = *c.status; __threadfence_system(); *c.host = st; }
            return;
        }
        // indexesold = the sorted ids; the next pass fits on its first n (when converged the OLD order stays, lovo.go:255)
        for (int k = tid; k < ncand; k += kCtlThreads) c.old[k] = pos[k];
        write_template(c, pos, n, rows, s_bbar);
    }
    if (tid == 0) {
        st.itercount++;
        st.n = n;
        st.converged = converged ? 1 : 0;
        if (!converged) { st.prevlim = s_newlim; st.disagreement = s_disagree; }
        st.newlim = s_newlim;
        st.err = 0;
        st.dev_status = *c.status;
        __threadfence_system();
        *c.host = st;
    }
}

}  // namespace

int lovo_ctl_slots() { return kCtlSlots; }

int launch_lovo_first_template(gcb_traj* t, const LovoCtl& c) {
    GCB_CUDA(opt_in_smem(lovo_first_template_kernel, t->dev));
    lovo_first_template_kernel<<<1, kCtlThreads, sizeof(double) * 3 * kSumChunk, t->stream>>>(c);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

int launch_lovo_controller(gcb_traj* t, const LovoCtl& c) {
    const size_t rank_smem = sizeof(unsigned long long) * (size_t)c.ncand;
    const size_t ctl_smem = sizeof(double) * (3 * kSumChunk + (size_t)c.ncand) + sizeof(unsigned int) * (size_t)((c.ncand + 31) / 32 + 1);
    GCB_CUDA(opt_in_smem(lovo_rank_kernel, t->dev));
    GCB_CUDA(opt_in_smem(lovo_controller_kernel, t->dev));
    lovo_rank_kernel<<<(c.ncand + kRankPerCta - 1) / kRankPerCta, kRankThreads, rank_smem, t->stream>>>(c);
    GCB_LAUNCHED();
    lovo_controller_kernel<<<1, kCtlThreads, ctl_smem, t->stream>>>(c);
    GCB_LAUNCHED();
    GCB_CUDA(cudaGetLastError());
    return GCB_OK;
}

}  // namespace gcb
