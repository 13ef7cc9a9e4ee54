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
// over before a residue begins (:480-483) -- and (b) the distances.  One CTA per frame.  Runs are examined 1024 at a time:
// those that begin a residue within reach of the solute are compacted into a dense list (most of a solvent box is not), then
// one thread per listed residue ranks it against the solute, which streams through shared memory in tiles.  An atom
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

constexpr int RT = 256;          // threads per CTA = runs per block of runs
constexpr int kRefTile = 1024;   // solute atoms per shared-memory tile
constexpr int kMaxSteps = 2048;
constexpr int kRunChunk = 1024;  // runs examined per compaction round

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
    const unsigned char* flags;   // per atom: bit 0 = may start a residue, bit 1 = pre-screened water
    const int* refidx;
    int nref;
    int com;
    double step, end;
    int nsteps;
    unsigned long long* cdf;      // [nsteps] summed over frames
};

__global__ void __launch_bounds__(RT) rdf_kernel(const RdfParams p) {
    __shared__ double rx[kRefTile], ry[kRefTile], rz[kRefTile];
    __shared__ unsigned int hist[kMaxSteps + 1];
    __shared__ double red[RT / 32][4];
    __shared__ double sphere[4];   // centre of the solute's bounding sphere, radius
    __shared__ int list_s[kRunChunk], list_e[kRunChunk];   // residues of the current chunk that can be within the cutoff: atoms [s, e)
    __shared__ int nlist;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double cutoff = p.end, cutoffplus = p.end + 3.0;
    for (long f = blockIdx.x; f < p.nframes; f += gridDim.x) {
        const double* X = p.frames + f * p.stride;
        const double* Y = X + p.np;
        const double* Z = Y + p.np;
        for (int k = tid; k <= p.nsteps; k += RT) hist[k] = 0u;
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

        bool tile_ready = false;   // CTA-uniform
        for (int rb = 0; rb < p.nruns; rb += kRunChunk) {
            // ---- step A: which runs of this chunk begin a residue within reach of the solute?  (compacted into a dense list) ----
            __syncthreads();
            if (tid == 0) nlist = 0;
            __syncthreads();
            const int rend = min(rb + kRunChunk, p.nruns);
            for (int r = rb + tid; r < rend; r += RT) {
                const Run run = p.runs[r];
                int start = -1;
                for (int a = run.s; a < run.e; a++) {   // first atom that begins the residue (solvation.go:478-486)
                    const unsigned char fl = p.flags[a];
                    if (fl & 2) {
                        const double dx = fx - X[a], dy = fy - Y[a], dz = fz - Z[a];
                        if (sqrt(dx * dx + dy * dy + dz * dz) > cutoffplus) continue;
                    }
                    if (fl & 1) { start = a; break; }
                }
                if (start < 0) continue;
                bool active = false;
                if (p.com) {
                    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
                    for (int a = start; a < run.e; a++) { c0 += X[a]; c1 += Y[a]; c2 += Z[a]; }
                    const double inv = 1.0 / (double)(run.e - start);
                    const double dx = c0 * inv - cx, dy = c1 * inv - cy, dz = c2 * inv - cz;
                    active = dx * dx + dy * dy + dz * dz <= reach2;
                } else {
                    for (int a = start; a < run.e && !active; a++) {
                        const double dx = X[a] - cx, dy = Y[a] - cy, dz = Z[a] - cz;
                        active = dx * dx + dy * dy + dz * dz <= reach2;
                    }
                }
                if (active) {
                    const int slot = atomicAdd(&nlist, 1);
                    list_s[slot] = start;
                    list_e[slot] = run.e;
                }
            }
            __syncthreads();
            // ---- step B: every listed residue against the whole solute, one thread per residue ----
            const int nl = nlist;
            for (int base = 0; base < nl; base += RT) {
                const bool mine = base + tid < nl;
                const int start = mine ? list_s[base + tid] : 0, stop = mine ? list_e[base + tid] : 0;
                double px = 0.0, py = 0.0, pz = 0.0;   // COM mode: the residue's centre
                if (mine && p.com) {   // chem.CenterOfMass with nil masses: column sums in row order times 1/n (geometric.go:615-629)
                    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
                    for (int a = start; a < stop; a++) { c0 += X[a]; c1 += Y[a]; c2 += Z[a]; }
                    const double inv = 1.0 / (double)(stop - start);
                    px = c0 * inv; py = c1 * inv; pz = c2 * inv;
                }
                double best = 1e10;   // squared; MolShortestDist starts from 100000 (:674)
                for (int t0 = 0; t0 < p.nref; t0 += kRefTile) {
                    const int tn = min(kRefTile, p.nref - t0);
                    if (p.nref > kRefTile || !tile_ready) {   // a solute of one tile is loaded once per frame
                        __syncthreads();
                        for (int j = tid; j < tn; j += RT) { const int a = p.refidx[t0 + j]; rx[j] = X[a]; ry[j] = Y[a]; rz[j] = Z[a]; }
                        __syncthreads();
                        tile_ready = true;
                    }
                    if (!mine) continue;
                    if (p.com) {
                        double b0 = best, b1 = best;
                        int j = 0;
                        for (; j + 1 < tn; j += 2) {
                            const double dx = rx[j] - px, dy = ry[j] - py, dz = rz[j] - pz;
                            const double ex = rx[j + 1] - px, ey = ry[j + 1] - py, ez = rz[j + 1] - pz;
                            b0 = fmin(b0, fma(dz, dz, fma(dy, dy, dx * dx)));
                            b1 = fmin(b1, fma(ez, ez, fma(ey, ey, ex * ex)));
                        }
                        if (j < tn) { const double dx = rx[j] - px, dy = ry[j] - py, dz = rz[j] - pz; b0 = fmin(b0, fma(dz, dz, fma(dy, dy, dx * dx))); }
                        best = fmin(b0, b1);
                    } else {
                        for (int a = start; a < stop; a++) {
                            const double ax = X[a], ay = Y[a], az = Z[a];
                            {
                                const double dx = ax - cx, dy = ay - cy, dz = az - cz;
                                if (dx * dx + dy * dy + dz * dz > reach2) continue;
                            }
                            double b0 = best, b1 = best;
                            int j = 0;
                            for (; j + 1 < tn; j += 2) {
                                const double dx = rx[j] - ax, dy = ry[j] - ay, dz = rz[j] - az;
                                const double ex = rx[j + 1] - ax, ey = ry[j + 1] - ay, ez = rz[j + 1] - az;
                                b0 = fmin(b0, fma(dz, dz, fma(dy, dy, dx * dx)));
                                b1 = fmin(b1, fma(ez, ez, fma(ey, ey, ex * ex)));
                            }
                            if (j < tn) { const double dx = rx[j] - ax, dy = ry[j] - ay, dz = rz[j] - az; b0 = fmin(b0, fma(dz, dz, fma(dy, dy, dx * dx))); }
                            best = fmin(b0, b1);
                        }
                    }
                }
                if (mine) {
                    const double d = sqrt(best);
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
        __syncthreads();
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
    std::map<std::pair<int, int>, int> seen;
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
            // (molid_skip / chain_skip, :486, :537-538).  With one chain per run and no (MolID, chain) pair repeated among the
            // candidate runs that rule only ever skips the rest of the run just handled, which is what a run is here.
            if (!chain_const) { set_error("gcb_rdf_cdf_sum: atoms %d..%d share MolID %d but not the chain; not supported", s, e - 1, molid[s]); return GCB_ERR_ARG; }
            if (seen.count({molid[s], chain[s]})) {
                set_error("gcb_rdf_cdf_sum: MolID %d / chain id %d names two separate solvent residues; renumber the topology", molid[s], chain[s]);
                return GCB_ERR_ARG;
            }
            seen[{molid[s], chain[s]}] = 1;
            runs.push_back({s, e});
        }
        s = e;
    }
    const int totalsteps = nsteps;
    for (int k = 0; k < totalsteps; k++) cdf_sum[k] = 0.0;
    const long used = (nframes / cpus) * cpus;   // whole chunks only (solvation.go:166-214 with dcd.go:520-523)
    if (frames_read) *frames_read = used;
    Run* d_runs = nullptr;
    unsigned char* d_fl = nullptr;
    int* d_ref = nullptr;
    unsigned long long* d_cdf = nullptr;
    std::vector<unsigned long long> h_cdf((size_t)totalsteps, 0ull);
    int rc = GCB_OK;
    cudaError_t e;
#define RDF_TRY(call) do { e = (call); if (e != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e)); rc = GCB_ERR_CUDA; goto done; } } while (0)
    RDF_TRY(cudaMalloc(&d_runs, sizeof(Run) * std::max<size_t>(runs.size(), 1)));
    RDF_TRY(cudaMalloc(&d_fl, (size_t)t->np));
    RDF_TRY(cudaMalloc(&d_ref, sizeof(int) * (size_t)nref));
    RDF_TRY(cudaMalloc(&d_cdf, sizeof(unsigned long long) * (size_t)totalsteps));
    RDF_TRY(cudaMemcpyAsync(d_runs, runs.data(), sizeof(Run) * runs.size(), cudaMemcpyHostToDevice, t->stream));
    RDF_TRY(cudaMemcpyAsync(d_fl, fl.data(), (size_t)t->np, cudaMemcpyHostToDevice, t->stream));
    RDF_TRY(cudaMemcpyAsync(d_ref, refidx, sizeof(int) * (size_t)nref, cudaMemcpyHostToDevice, t->stream));
    RDF_TRY(cudaMemsetAsync(d_cdf, 0, sizeof(unsigned long long) * (size_t)totalsteps, t->stream));
    if (used > 0 && !runs.empty()) {
        RdfParams p;
        p.frames = t->d_frames + first * 3L * t->np;
        p.stride = 3L * t->np;
        p.np = t->np;
        p.nframes = used;
        p.runs = d_runs;
        p.nruns = (int)runs.size();
        p.flags = d_fl;
        p.refidx = d_ref;
        p.nref = nref;
        p.com = com ? 1 : 0;
        p.step = step;
        p.end = end;
        p.nsteps = totalsteps;
        p.cdf = d_cdf;
        long grid = (long)t->sm_count * 4;
        if (grid > used) grid = used;
        rdf_kernel<<<(unsigned)grid, RT, 0, t->stream>>>(p);
        GCB_LAUNCHED();
    }
    RDF_TRY(cudaMemcpyAsync(h_cdf.data(), d_cdf, sizeof(unsigned long long) * (size_t)totalsteps, cudaMemcpyDeviceToHost, t->stream));
    RDF_TRY(cudaStreamSynchronize(t->stream));
    RDF_TRY(cudaGetLastError());
    for (int k = 0; k < totalsteps; k++) cdf_sum[k] = (double)h_cdf[(size_t)k];
done:
#undef RDF_TRY
    cudaFree(d_runs); cudaFree(d_fl); cudaFree(d_ref); cudaFree(d_cdf);
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
