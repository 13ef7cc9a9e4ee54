// rdf.cu -- solvent shells around a solute for every resident frame.
//
// Restates the frame loop of solv.ConcMolRDF (/root/reference/solv/solvation.go:140-218): for each frame
// solv.FrameUMolCRDF (:380-410) ranks the solvent residues by their distance to the solute (solv.DistRank, :447-541:
// the smallest atom-atom distance, solv.MolShortestDist :669-690, or the distance of the residue's centre when
// Options.COM is set) and counts, for every limit i*step, the residues closer than the limit -- minus one when the
// count is not zero (:402-404).  The per-frame counts are summed over the frames; MDFFromCDF (:317-352) is host work.
//
// The topology is static, so the host turns it into "runs" (maximal stretches of consecutive atoms with one MolID:
// the reference extends a residue from its first atom while the MolID stays the same, :490-496).  What depends on the
// frame is (a) where a water run starts -- atoms of SOL/WAT/HOH further than end+3 from the first solute atom are passed
// over before a residue begins (:480-483) -- and (b) the distances.  One CTA per frame, the solute staged in shared memory
// (whole when it has at most 4096 atoms, else in tiles).  Runs are examined in chunks: the atoms of residues that begin and
// lie within reach of the solute are compacted into a dense work list (most of a solvent box is out of reach), one thread
// per listed atom ranks it against the solute, and the residues' minima are merged in shared memory.  An atom
// further from the solute's bounding sphere than `end` cannot come within `end` of any solute atom, so it is skipped without changing any kept distance.  Counts are integers: the
// accumulation over frames (64-bit atomics) is exact and order-independent.
//
// Compute-bound (pairs x 7 fp64 operations), not HBM-bound: frames are read once, the solute once per tile.
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace gcb {

namespace {

constexpr int RT = 256;            // threads per CTA
constexpr int kRefTileMax = 4096;  // solute atoms held in shared memory at a time (96 KB); larger solutes stream through in tiles
constexpr int kMaxSteps = 2048;
constexpr int kListCap = 4096;     // work-list entries (atoms within reach) per chunk of runs

struct Run {
    int s, e;   // atoms [s, e)
};

struct RdfParams {
    const double* frames;
    long stride;
    int np;
    long nframes;
    const Run* runs;
    int nruns;
    const int* chunk_first;       // [nchunks + 1] run ranges whose atoms fit one work list
    int nchunks;
    const int* run_key;           // per run: id of its (MolID, chain) pair; NULL when no pair names two runs (no skip rule needed)
    const unsigned char* flags;   // per atom: bit 0 = may start a residue, bit 1 = pre-screened water
    const int* refidx;
    int nref;
    int tile;                     // solute atoms per shared-memory tile (even)
    int com;
    double step, end;
    int nsteps;
    unsigned long long* cdf;      // [nsteps] summed over frames
};

// smallest squared distance from (ax, ay, az) to the tn solute atoms staged in shared memory; four independent chains
__device__ __forceinline__ double tile_min_d2(const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                                              int tn, double ax, double ay, double az) {
    double b0 = 1e300, b1 = 1e300, b2 = 1e300, b3 = 1e300;
    int j = 0;
    for (; j + 3 < tn; j += 4) {
        const double2 x0 = *reinterpret_cast<const double2*>(rx + j), x1 = *reinterpret_cast<const double2*>(rx + j + 2);
        const double2 y0 = *reinterpret_cast<const double2*>(ry + j), y1 = *reinterpret_cast<const double2*>(ry + j + 2);
        const double2 z0 = *reinterpret_cast<const double2*>(rz + j), z1 = *reinterpret_cast<const double2*>(rz + j + 2);
        const double d0x = x0.x - ax, d0y = y0.x - ay, d0z = z0.x - az;
        const double d1x = x0.y - ax, d1y = y0.y - ay, d1z = z0.y - az;
        const double d2x = x1.x - ax, d2y = y1.x - ay, d2z = z1.x - az;
        const double d3x = x1.y - ax, d3y = y1.y - ay, d3z = z1.y - az;
        b0 = fmin(b0, fma(d0z, d0z, fma(d0y, d0y, d0x * d0x)));
        b1 = fmin(b1, fma(d1z, d1z, fma(d1y, d1y, d1x * d1x)));
        b2 = fmin(b2, fma(d2z, d2z, fma(d2y, d2y, d2x * d2x)));
        b3 = fmin(b3, fma(d3z, d3z, fma(d3y, d3y, d3x * d3x)));
    }
    for (; j < tn; j++) {
        const double dx = rx[j] - ax, dy = ry[j] - ay, dz = rz[j] - az;
        b0 = fmin(b0, fma(dz, dz, fma(dy, dy, dx * dx)));
    }
    return fmin(fmin(b0, b1), fmin(b2, b3));
}

// One CTA per frame.  Per chunk of runs: (A) the atoms that can be within the cutoff are appended to a dense work list,
// each tagged with the slot of its residue; (B) one thread per listed atom finds its smallest squared distance to the solute
// (identical trip counts: the lanes of a warp stay together) and merges it into the residue's slot with a shared-memory
// atomic minimum on the bit pattern (non-negative doubles order like unsigned integers); (C) one thread per slot bins the
// residue's distance.  With Options.COM a list entry is a residue and its point is the residue's centre.
__global__ void __launch_bounds__(RT) rdf_kernel(const RdfParams p) {
    extern __shared__ __align__(16) unsigned char rdf_smem[];
    double* rx = reinterpret_cast<double*>(rdf_smem);
    double* ry = rx + p.tile;
    double* rz = ry + p.tile;
    unsigned long long* best = reinterpret_cast<unsigned long long*>(rz + p.tile);   // [kListCap] min squared distance per slot
    int2* ent = reinterpret_cast<int2*>(best + kListCap);                             // [kListCap] (atom, slot) or (start, stop)
    unsigned int* hist = reinterpret_cast<unsigned int*>(ent + kListCap);             // [nsteps + 1]
    int* rstart = reinterpret_cast<int*>(hist + p.nsteps + 1);                        // [kListCap] (skip rule only): first atom of a run's residue, -1 = none
    __shared__ int carry_key;      // (MolID, chain) id of the last run that passed in the earlier chunks of the frame
    __shared__ double red[RT / 32][4];
    __shared__ double sphere[4];   // centre of the solute's bounding sphere, radius
    __shared__ int nent, nslot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double cutoff = p.end, cutoffplus = p.end + 3.0;
    const unsigned long long far_bits = (unsigned long long)__double_as_longlong(1e10);   // MolShortestDist starts from 100000 (:674)
    const bool one_tile = p.nref <= p.tile;
    for (long f = blockIdx.x; f < p.nframes; f += gridDim.x) {
        const double* X = p.frames + f * p.stride;
        const double* Y = X + p.np;
        const double* Z = Y + p.np;
        __syncthreads();   // the previous frame's tile and histogram are no longer read
        for (int k = tid; k <= p.nsteps; k += RT) hist[k] = 0u;
        if (tid == 0) carry_key = -1;
        if (one_tile)
            for (int j = tid; j < p.nref; j += RT) { const int a = p.refidx[j]; rx[j] = X[a]; ry[j] = Y[a]; rz[j] = Z[a]; }
        // bounding sphere of the solute: centre = mean position, radius = largest distance from it
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int j = tid; j < p.nref; j += RT) { const int a = p.refidx[j]; s0 += X[a]; s1 += Y[a]; s2 += Z[a]; }
        for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (lane == 0) { red[warp][0] = s0; red[warp][1] = s1; red[warp][2] = s2; }
        __syncthreads();
        if (tid == 0) {
            double a = 0, b = 0, c = 0;
            for (int w = 0; w < RT / 32; w++) { a += red[w][0]; b += red[w][1]; c += red[w][2]; }
            sphere[0] = a / p.nref; sphere[1] = b / p.nref; sphere[2] = c / p.nref;
        }
        __syncthreads();
        const double cx = sphere[0], cy = sphere[1], cz = sphere[2];
        double r2 = 0.0;
        for (int j = tid; j < p.nref; j += RT) {
            const int a = p.refidx[j];
            const double dx = X[a] - cx, dy = Y[a] - cy, dz = Z[a] - cz;
            r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
        }
        for (int o = 16; o > 0; o >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
        if (lane == 0) red[warp][3] = r2;
        __syncthreads();
        if (tid == 0) {
            double m = 0.0;
            for (int w = 0; w < RT / 32; w++) m = fmax(m, red[w][3]);
            sphere[3] = sqrt(m);
        }
        __syncthreads();
        // an atom can be within `cutoff` of the solute only inside this radius (a relative margin covers the rounding)
        const double reach = (sphere[3] + cutoff) * (1.0 + 1e-9) + 1e-9;
        const double reach2 = reach * reach;
        const int a0 = p.refidx[0];
        const double fx = X[a0], fy = Y[a0], fz = Z[a0];   // ref.VecView(0)

        for (int c = 0; c < p.nchunks; c++) {
            // ---- step A: work list of this chunk ----
            __syncthreads();
            if (tid == 0) { nent = 0; nslot = 0; }
            __syncthreads();
            // first atom that begins the residue of run r in this frame (solvation.go:478-486), -1 when none does
            auto residue_start = [&](const Run& run) {
                for (int a = run.s; a < run.e; a++) {
                    const unsigned char fl = p.flags[a];
                    if (fl & 2) {
                        const double dx = fx - X[a], dy = fy - Y[a], dz = fz - Z[a];
                        if (sqrt(dx * dx + dy * dy + dz * dz) > cutoffplus) continue;
                    }
                    if (fl & 1) return a;
                }
                return -1;
            };
            const int r0 = p.chunk_first[c], r1 = p.chunk_first[c + 1];
            if (p.run_key) {
                // The reference remembers (MolID, chain) of the residue it started last and passes over atoms that carry the same
                // pair (molid_skip / chain_skip, :486, :537-538).  Inside a run that skips the rest of the residue just read; across
                // runs it drops a residue whose pair equals that of the last one started -- which happens in real topologies, where
                // residue numbers wrap (9999 in PDB files) and the chain is blank.  Whether or not the previous passing run was
                // itself dropped, the pair remembered after it equals its own, so: a run that passes starts a residue iff its pair
                // differs from that of the PREVIOUS PASSING run of the frame.
                for (int r = r0 + tid; r < r1; r += RT) rstart[r - r0] = residue_start(p.runs[r]);
                __syncthreads();
            }
            for (int r = r0 + tid; r < r1; r += RT) {
                const Run run = p.runs[r];
                int start;
                if (p.run_key) {
                    start = rstart[r - r0];
                    if (start < 0) continue;
                    int prev = carry_key;
                    for (int q = r - r0 - 1; q >= 0; q--)
                        if (rstart[q] >= 0) { prev = p.run_key[r0 + q]; break; }
                    if (prev == p.run_key[r]) continue;
                } else {
                    start = residue_start(run);
                }
                if (start < 0) continue;
                if (p.com) {   // chem.CenterOfMass with nil masses: column sums in row order times 1/n (geometric.go:615-629)
                    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
                    for (int a = start; a < run.e; a++) { c0 += X[a]; c1 += Y[a]; c2 += Z[a]; }
                    const double inv = 1.0 / (double)(run.e - start);
                    const double dx = c0 * inv - cx, dy = c1 * inv - cy, dz = c2 * inv - cz;
                    if (dx * dx + dy * dy + dz * dz <= reach2) {
                        const int e = atomicAdd(&nent, 1);
                        ent[e] = make_int2(start, run.e);
                        best[e] = far_bits;
                    }
                } else {
                    int slot = -1;
                    for (int a = start; a < run.e; a++) {
                        const double dx = X[a] - cx, dy = Y[a] - cy, dz = Z[a] - cz;
                        if (dx * dx + dy * dy + dz * dz > reach2) continue;
                        if (slot < 0) { slot = atomicAdd(&nslot, 1); best[slot] = far_bits; }
                        ent[atomicAdd(&nent, 1)] = make_int2(a, slot);
                    }
                }
            }
            __syncthreads();
            if (p.run_key && tid == 0) {
                for (int q = r1 - r0 - 1; q >= 0; q--)
                    if (rstart[q] >= 0) { carry_key = p.run_key[r0 + q]; break; }
            }
            // ---- step B: every listed point against the solute ----
            const int ne = nent;
            for (int t0 = 0; t0 < p.nref; t0 += p.tile) {
                const int tn = min(p.tile, p.nref - t0);
                if (!one_tile) {
                    __syncthreads();
                    for (int j = tid; j < tn; j += RT) { const int a = p.refidx[t0 + j]; rx[j] = X[a]; ry[j] = Y[a]; rz[j] = Z[a]; }
                    __syncthreads();
                }
                for (int e = tid; e < ne; e += RT) {
                    const int2 en = ent[e];
                    double ax, ay, az;
                    int slot;
                    if (p.com) {
                        double c0 = 0.0, c1 = 0.0, c2 = 0.0;
                        for (int a = en.x; a < en.y; a++) { c0 += X[a]; c1 += Y[a]; c2 += Z[a]; }
                        const double inv = 1.0 / (double)(en.y - en.x);
                        ax = c0 * inv; ay = c1 * inv; az = c2 * inv;
                        slot = e;
                    } else {
                        ax = X[en.x]; ay = Y[en.x]; az = Z[en.x];
                        slot = en.y;
                    }
                    const double d2 = tile_min_d2(rx, ry, rz, tn, ax, ay, az);
                    atomicMin(&best[slot], (unsigned long long)__double_as_longlong(d2));
                }
            }
            __syncthreads();
            // ---- step C: one distance per residue -> shell ----
            const int ns = p.com ? ne : nslot;
            for (int sl = tid; sl < ns; sl += RT) {
                const double d = sqrt(__longlong_as_double((long long)best[sl]));
                if (d <= cutoff) {
                    // smallest i >= 1 with d < i*step (the limits of solvation.go:398-399, same product)
                    int i = (int)(d / p.step);
                    if (i < 1) i = 1;
                    while (i > 1 && d < (double)(i - 1) * p.step) i--;
                    while (i <= p.nsteps && !(d < (double)i * p.step)) i++;
                    if (i <= p.nsteps) atomicAdd(&hist[i], 1u);
                }
            }
        }
        __syncthreads();
        // cumulative counts, each reduced by one when it is not zero (:402-404); nsteps is a few hundred at most
        if (tid == 0) {
            unsigned int c = 0;
            for (int i = 1; i <= p.nsteps; i++) {
                c += hist[i];
                if (c > 1u) atomicAdd(&p.cdf[i - 1], (unsigned long long)(c - 1u));
            }
        }
    }
}

}  // namespace
}  // namespace gcb

extern "C" int gcb_rdf_cdf_sum(gcb_traj* t, long first, long nframes, const int* molid, const int* chain, const unsigned char* flags,
                               const int* refidx, int nref, int com, double step, double end, int cpus, gcb_comm* comm,
                               double* cdf_sum, long* frames_read) {
    using namespace gcb;
    if (!t) { set_error("NULL trajectory handle"); return GCB_ERR_ARG; }
    if (first < 0 || nframes < 0 || first + nframes > t->nframes) { set_error("frame window outside the trajectory"); return GCB_ERR_ARG; }
    if (!molid || !chain || !flags || !refidx || nref <= 0 || !cdf_sum) { set_error("gcb_rdf_cdf_sum: NULL topology, solute or output"); return GCB_ERR_ARG; }
    if (!(step > 0.0) || !(end > 0.0)) { set_error("gcb_rdf_cdf_sum: step and end must be positive"); return GCB_ERR_ARG; }
    const int nsteps = (int)(end / step);   // totalsteps, solvation.go:390
    if (nsteps < 1 || nsteps > kMaxSteps) { set_error("gcb_rdf_cdf_sum: %d shells (1..%d supported)", nsteps, kMaxSteps); return GCB_ERR_ARG; }
    if (cpus < 1) cpus = 1;
    const int natoms = t->natoms;
    for (int k = 0; k < nref; k++)
        if (refidx[k] < 0 || refidx[k] >= natoms) { set_error("solute index %d out of range", refidx[k]); return GCB_ERR_INDEXES; }
    GCB_CUDA(cudaSetDevice(t->dev));
    // the solute's own residues never count (allResIDandChains / repeated, solvation.go:418-436)
    std::map<std::pair<int, int>, int> own;
    for (int k = 0; k < nref; k++) own[{molid[refidx[k]], chain[refidx[k]]}] = 1;
    // runs of one MolID; a run can start a residue when it holds a candidate atom outside the solute's residues
    std::vector<Run> runs;
    std::vector<unsigned char> fl((size_t)t->np, 0);
    std::map<std::pair<int, int>, int> seen;   // (MolID, chain) -> id
    std::vector<int> run_key;
    bool repeated_pairs = false;
    for (int s = 0; s < natoms;) {
        int e = s + 1;
        while (e < natoms && molid[e] == molid[s]) e++;
        bool cand = false, chain_const = true;
        for (int a = s; a < e; a++) {
            chain_const = chain_const && chain[a] == chain[s];
            if ((flags[a] & 1) && !own.count({molid[a], chain[a]})) { cand = true; fl[(size_t)a] = (unsigned char)(flags[a] & 3); }
            else fl[(size_t)a] = (unsigned char)(flags[a] & 2);
        }
        if (cand) {
            // The reference carries (MolID, chain) of the last residue it started and passes over atoms that match it
            // (molid_skip / chain_skip, :486, :537-538).  With one chain per run that rule skips the rest of the run just handled,
            // which is what a run is here; when a pair names several runs the kernel applies it across runs too (run_key).
            if (!chain_const) { set_error("gcb_rdf_cdf_sum: atoms %d..%d share MolID %d but not the chain; not supported", s, e - 1, molid[s]); return GCB_ERR_ARG; }
            auto ins = seen.emplace(std::make_pair(molid[s], chain[s]), (int)seen.size());
            if (!ins.second) repeated_pairs = true;   // e.g. residue numbers that wrap at 9999 with a blank chain
            run_key.push_back(ins.first->second);
            runs.push_back({s, e});
        }
        s = e;
    }
    // chunks of runs whose atoms fit one work list
    std::vector<int> chunk_first{0};
    {
        int atoms_in = 0, runs_in = 0;
        for (size_t r = 0; r < runs.size(); r++) {
            const int len = runs[r].e - runs[r].s;
            if (len > kListCap) { set_error("gcb_rdf_cdf_sum: a solvent residue of %d atoms (at most %d supported)", len, kListCap); return GCB_ERR_ARG; }
            if (atoms_in + len > kListCap || runs_in + 1 > kListCap) { chunk_first.push_back((int)r); atoms_in = 0; runs_in = 0; }
            atoms_in += len;
            runs_in++;
        }
        chunk_first.push_back((int)runs.size());
    }
    const int totalsteps = nsteps;
    for (int k = 0; k < totalsteps; k++) cdf_sum[k] = 0.0;
    const long used = (nframes / cpus) * cpus;   // whole chunks only (solvation.go:166-214 with dcd.go:520-523)
    if (frames_read) *frames_read = used;
    Run* d_runs = nullptr;
    int* d_chunks = nullptr;
    unsigned char* d_fl = nullptr;
    int* d_ref = nullptr;
    int* d_keys = nullptr;
    unsigned long long* d_cdf = nullptr;
    std::vector<unsigned long long> h_cdf((size_t)totalsteps, 0ull);
    int rc = GCB_OK;
    cudaError_t e;
#define RDF_TRY(call) do { e = (call); if (e != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e)); rc = GCB_ERR_CUDA; goto done; } } while (0)
    RDF_TRY(cudaMalloc(&d_runs, sizeof(Run) * std::max<size_t>(runs.size(), 1)));
    RDF_TRY(cudaMalloc(&d_fl, (size_t)t->np));
    RDF_TRY(cudaMalloc(&d_chunks, sizeof(int) * chunk_first.size()));
    RDF_TRY(cudaMemcpyAsync(d_chunks, chunk_first.data(), sizeof(int) * chunk_first.size(), cudaMemcpyHostToDevice, t->stream));
    RDF_TRY(cudaMalloc(&d_ref, sizeof(int) * (size_t)nref));
    if (repeated_pairs) {
        RDF_TRY(cudaMalloc(&d_keys, sizeof(int) * run_key.size()));
        RDF_TRY(cudaMemcpyAsync(d_keys, run_key.data(), sizeof(int) * run_key.size(), cudaMemcpyHostToDevice, t->stream));
    }
    RDF_TRY(cudaMalloc(&d_cdf, sizeof(unsigned long long) * (size_t)totalsteps));
    RDF_TRY(cudaMemcpyAsync(d_runs, runs.data(), sizeof(Run) * runs.size(), cudaMemcpyHostToDevice, t->stream));
    RDF_TRY(cudaMemcpyAsync(d_fl, fl.data(), (size_t)t->np, cudaMemcpyHostToDevice, t->stream));
    RDF_TRY(cudaMemcpyAsync(d_ref, refidx, sizeof(int) * (size_t)nref, cudaMemcpyHostToDevice, t->stream));
    RDF_TRY(cudaMemsetAsync(d_cdf, 0, sizeof(unsigned long long) * (size_t)totalsteps, t->stream));
    if (used > 0 && !runs.empty()) {
        if ((rc = deps_before_compute(t, first, used)) != GCB_OK) goto done;
        RdfParams p;
        p.frames = t->d_frames + first * 3L * t->np;
        p.stride = 3L * t->np;
        p.np = t->np;
        p.nframes = used;
        p.runs = d_runs;
        p.nruns = (int)runs.size();
        p.chunk_first = d_chunks;
        p.nchunks = (int)chunk_first.size() - 1;
        p.tile = nref <= kRefTileMax ? (nref + 1) / 2 * 2 : kRefTileMax;
        p.flags = d_fl;
        p.run_key = d_keys;
        p.refidx = d_ref;
        p.nref = nref;
        p.com = com ? 1 : 0;
        p.step = step;
        p.end = end;
        p.nsteps = totalsteps;
        p.cdf = d_cdf;
        const size_t smem = sizeof(double) * 3 * (size_t)p.tile + (sizeof(unsigned long long) + sizeof(int2)) * kListCap +
                            sizeof(unsigned int) * (size_t)(totalsteps + 1) + (d_keys ? sizeof(int) * kListCap : 0);
        RDF_TRY(opt_in_smem(rdf_kernel, t->dev));
        int per_sm = 1;
        RDF_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rdf_kernel, RT, smem));
        long grid = (long)t->sm_count * (per_sm > 0 ? per_sm : 1);
        if (grid > used) grid = used;
        rdf_kernel<<<(unsigned)grid, RT, smem, t->stream>>>(p);
        GCB_LAUNCHED();
        if ((rc = deps_after_compute(t, first, used)) != GCB_OK) goto done;
    }
    RDF_TRY(cudaMemcpyAsync(h_cdf.data(), d_cdf, sizeof(unsigned long long) * (size_t)totalsteps, cudaMemcpyDeviceToHost, t->stream));
    RDF_TRY(cudaStreamSynchronize(t->stream));
    RDF_TRY(cudaGetLastError());
    for (int k = 0; k < totalsteps; k++) cdf_sum[k] = (double)h_cdf[(size_t)k];
done:
#undef RDF_TRY
    cudaFree(d_runs); cudaFree(d_chunks); cudaFree(d_fl); cudaFree(d_ref); cudaFree(d_keys); cudaFree(d_cdf);
    if (rc) return rc;
    if (comm) {   // frames sharded over ranks: the counts add up (integers below 2^53: exact in any order)
        rc = gcb_allreduce_sum_f64(comm, cdf_sum, totalsteps);
        if (rc) return rc;
        if (frames_read) {
            double fr = (double)used;
            rc = gcb_allreduce_sum_f64(comm, &fr, 1);
            if (rc) return rc;
            *frames_read = (long)fr;
        }
    }
    return GCB_OK;
}
