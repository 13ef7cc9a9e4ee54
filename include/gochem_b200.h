/*
 * gochem_b200.h -- C ABI of the B200 superposition / RMSD / LOVO device path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int status returns
 * (0 = ok, like the reference's only FFI precedent, traj/xtc/libgoxtc.c:87-104), opaque
 * handles closed explicitly (traj/xtc/xtc.go:173-183).  A cgo binding for the reference's Go
 * packages is shown in INTEGRATION.md.  Every entry point sets the CUDA device of the handle it
 * is given (cgo calls may hop OS threads) and is synchronous unless its name ends in _async.
 *
 * Coordinates cross the boundary as the reference stores them: float64, row-major N x 3 (AoS
 * xyz), v3.Matrix (v3/gonum.go:54-58).  Reader staging may also hand over float32 AoS
 * (xtc, traj/xtc/xtc.go:129-134) or float32 SoA planes (dcd, traj/dcd/dcd.go:380-413).
 * On the device frames live as float64 SoA planes (x[], y[], z[] per frame).
 *
 * There is no CPU fallback: without a usable CUDA device every compute call fails with
 * GCB_ERR_CUDA and gcb_last_error() says why.
 */
#ifndef GOCHEM_B200_H
#define GOCHEM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define GCB_OK 0
#define GCB_ERR_SHAPE -1     /* "Ill-formed matrices" (geometric.go:130-132), "Ill formed coordinates" (handy.go:199-201) */
#define GCB_ERR_INDEXES -2   /* "Indexes don't match molecules" (handy.go:190) */
#define GCB_ERR_SVD -4       /* "mat.SVD failed" (geometric.go:155-157): non-finite correlation matrix */
#define GCB_ERR_ALLOC -5
#define GCB_ERR_CUDA -7      /* CUDA runtime / driver / NCCL failure; see gcb_last_error */
#define GCB_ERR_ARG -8       /* NULL handle, out-of-range frame window, bad enum */
#define GCB_EOF 11           /* end of trajectory, same value libgoxtc returns (xtc.go:204-213) */

/* host layouts accepted by gcb_traj_upload */
#define GCB_F64_AOS 0 /* float64 [frame][atom][xyz]  -- v3.Matrix.RawSlice (v3/gonum.go:93-95) */
#define GCB_F32_AOS 1 /* float32 [frame][atom][xyz]  -- libxdrfile output (libgoxtc.c:87-104) */
#define GCB_F32_SOA 2 /* float32 [frame][xyz][atom]  -- DCD records (dcd.go:380-413) */
#define GCB_I32_AOS 3 /* int32 [frame][atom][xyz], fixed point -- STF lines (traj/stf/stf.go:324-345); coordinate = float64(value) / scale */

/* kernel variants (gcb_set_kernel_variant): AUTO picks the fastest one the shape allows */
#define GCB_KERNEL_AUTO 0
#define GCB_KERNEL_GENERIC 1 /* one CTA per frame, second pass re-reads the frame from L2 */
#define GCB_KERNEL_CLUSTER 2 /* TMA bulk-staged frame parts resident in shared memory, CTA cluster per frame */

typedef struct gcb_traj gcb_traj; /* device-resident trajectory: cap_frames x natoms, float64 SoA */
typedef struct gcb_comm gcb_comm; /* NCCL communicator wrapper */

/* ---- library ---- */
int gcb_version(void);
/* Copies the calling thread's last error message; returns its length. */
size_t gcb_last_error(char *buf, size_t cap);
int gcb_device_count(int *n);
/* Properties the host side sizes its work by: SM count, bytes of device memory, compute capability. */
int gcb_device_info(int dev, int *sm_count, size_t *total_mem, int *cc_major, int *cc_minor);
/* ---- settings ----
 * Every setting that steers kernel choice or shape belongs to ONE trajectory handle (gcb_traj_set_*): calls on distinct handles
 * are reentrant and may run concurrently from different threads, as the reference's functions are on distinct matrices
 * (align/lovo.go:290-295, 329).  The gcb_set_* forms change the process-wide DEFAULTS that a handle copies when it is created;
 * they never affect an existing handle. */
int gcb_set_kernel_variant(int variant);
int gcb_traj_set_kernel_variant(gcb_traj *t, int variant);
/* Tuning hook for the cluster kernel: CTAs per frame (1..8), shared-memory stages, how many frames pass 2
 * trails pass 1 (1..6), and where pass 2 reads the frame from (1 = its shared-memory stage, 0 = L2,
 * -1 = built-in choice).  0 / 0 / 0 / -1 restores the built-in choice.  resident = 1 | (warps << 4) with warps = 8 or 12 also
 * forces the number of consumer warps of the per-atom-distance instances (12 exists in -DGCB_DEV_WARPS=12 tuning builds only). */
int gcb_set_cluster_tuning(int cluster, int stages, int depth, int resident);
int gcb_traj_set_cluster_tuning(gcb_traj *t, int cluster, int stages, int depth, int resident);
/* Tuning hook: structures of at most this many atoms take the split form (a fit-only launch, then one elementwise sweep) for
 * write-back / per-atom-sum passes instead of the fused kernels.  A negative value restores the built-in threshold. */
int gcb_set_split_thresholds(int write_back_max_atoms, int per_atom_max_atoms);
int gcb_traj_set_split_thresholds(gcb_traj *t, int write_back_max_atoms, int per_atom_max_atoms);
/* How the per-frame RMSD of a fit-only call (no write-back) is formed.  0 (default): from the inner products of
 * the single sweep, E = Ga + Gb - 2 tr(R^T H), whenever its rounding bound moves the RMSD by < 1e-10 A, and from
 * the explicit residual of a second sweep otherwise (frames within ~0.3 A of the template).  1: always the explicit
 * residual, exactly as the reference computes it (geometric.go:284-286). */
int gcb_set_rmsd_mode(int explicit_only);
int gcb_traj_set_rmsd_mode(gcb_traj *t, int explicit_only);
/* Shape of the handle's last cluster-kernel launch: {CTAs per frame, atoms per CTA, atoms per thread, stages, depth,
 * resident | consumer warps << 4, clusters launched, dynamic shared memory bytes}. */
int gcb_traj_last_cluster_plan(const gcb_traj *t, int *out8);

/* ---- C-owned pinned staging (Go memory may not be retained across an async copy) ---- */
int gcb_pinned_alloc(size_t bytes, void **p);
int gcb_pinned_free(void *p);
/* Pin the calling thread to the CPUs local to device dev's PCIe link (sysfs local_cpulist) before allocating staging buffers,
 * so the pinned memory is NUMA-local to the GPU.  Optional; an error leaves the affinity unchanged. */
int gcb_bind_host_thread(int dev);

/* ---- device trajectory: the v3 device-matrix batch ---- */
int gcb_traj_create(int dev, int natoms, long cap_frames, gcb_traj **out);
int gcb_traj_destroy(gcb_traj *t);
int gcb_traj_natoms(const gcb_traj *t);
long gcb_traj_capacity(const gcb_traj *t);
long gcb_traj_nframes(const gcb_traj *t);
int gcb_traj_set_nframes(gcb_traj *t, long nframes);
/* Copy nframes host frames into slots [first, first+nframes); every coordinate is multiplied by
 * scale while widening (xtc: 10.0, nm -> A, xtc.go:129-134; otherwise 1.0).  Blocks until the
 * data is resident. */
int gcb_traj_upload(gcb_traj *t, long first, const void *host, long nframes, int layout, double scale);
/* Same, from a gcb_pinned_alloc buffer, on copy stream `slot` (0 or 1); returns once queued.
 * The buffer may be reused after gcb_traj_wait(t, slot). */
int gcb_traj_upload_async(gcb_traj *t, long first, const void *pinned, long nframes, int layout, double scale, int slot);
/* Same for float32 planes that sit inside fixed-size records: frame f is the record_bytes bytes at pinned + f * record_bytes with
 * x[], y[], z[] at byte offsets off_x / off_y / off_z.  This is a raw DCD frame -- Fortran record markers and the optional
 * unit-cell block included (traj/dcd/dcd.go:349-434) -- so a reader can read() whole runs of frames straight into the pinned
 * buffer with one call and no host-side pass over the data.  How many bytes one staging slot takes: gcb_traj_stage_bytes. */
int gcb_traj_upload_records_async(gcb_traj *t, long first, const void *pinned, long nframes, long record_bytes, long off_x,
                                  long off_y, long off_z, double scale, int slot);
size_t gcb_traj_stage_bytes(const gcb_traj *t);
int gcb_traj_wait(gcb_traj *t, int slot);
/* Frames back to float64 AoS (what chem.Super leaves in `test`, handy.go:207-213). */
int gcb_traj_download(gcb_traj *t, long first, long nframes, double *host_aos);

/* ---- chem.Super + chem.RMSD over frames [first, first+nframes) ----
 * Replaces the per-frame loop  chem.Super(frame, ref, idx_test, idx_ref)  (handy.go:175-214,
 * geometric.go:127-208) followed by  chem.RMSD(frame, ref, idx_test, idx_ref)
 * (geometric.go:232-288).
 *   ref_aos   : template, natoms_ref x 3 float64 row-major (host).
 *   idx_test / idx_ref, nidx : fit subset; NULL/0 = all atoms (handy.go:178-180).  Both lists must
 *               have nidx entries; idx_ref == NULL with idx_test given means "same list".
 *   write_back: 1 = frames are replaced by (frame + trans1) * R + trans2 for ALL atoms, the
 *               reference's in-place semantics (handy.go:207-211); 0 = frames untouched.
 *   rmsd[F]   : RMSD over the fit subset after the fit (explicit residual, geometric.go:284-286).
 *   rot[9F]   : rotation, row-major, right-multiplying row vectors (geometric.go:180-188).
 *   trans1[3F]: -centroid(test subset) (geometric.go:206).  trans2[3F]: Translation (:191-200).
 *   Any output pointer may be NULL.  All outputs are host memory. */
int gcb_super_rmsd(gcb_traj *t, long first, long nframes, const double *ref_aos, int natoms_ref,
                   const int *idx_test, const int *idx_ref, int nidx, int write_back,
                   double *rmsd, double *rot, double *trans1, double *trans2);
/* Launch only: results stay in the handle's device buffers until gcb_fetch_results.  Used to time the
 * kernel with inputs resident (bench.py `value`). */
int gcb_super_rmsd_async(gcb_traj *t, long first, long nframes, const double *ref_aos, int natoms_ref,
                         const int *idx_test, const int *idx_ref, int nidx, int write_back);
int gcb_fetch_results(gcb_traj *t, long first, long nframes, double *rmsd, double *rot, double *trans1, double *trans2);

/* chem.RMSD / chem.MemRMSD without superposition (geometric.go:232-288) for every frame. */
int gcb_rmsd(gcb_traj *t, long first, long nframes, const double *ref_aos, int natoms_ref,
             const int *idx_test, const int *idx_ref, int nidx, double *rmsd);

/* chem.MemPerAtomRMSD (geometric.go:294-321) for one resident frame, no superposition:
 * out[k] = |frame[idx_test[k]] - ref[idx_ref[k]]|, k < nidx (all atoms when idx_test is NULL). */
int gcb_peratom_dist(gcb_traj *t, long frame, const double *ref_aos, int natoms_ref,
                     const int *idx_test, const int *idx_ref, int nidx, double *out);

/* ---- align.RMSDTraj inner pass (align/lovo.go:265-281, 328-340) ----
 * For every frame: Super on the idx subset (both sides), then MemPerAtomRMSD against ref for ALL
 * atoms; sum[i] = sum over frames of |frame_i - ref_i| (not divided).  write_back as above.
 * sum is host memory, natoms doubles.  With comm != NULL the per-rank sums are combined across
 * ranks (all-gather + rank-ordered sum, bit-identical on every rank). */
int gcb_peratom_dev_sum(gcb_traj *t, long first, long nframes, const double *ref_aos,
                        const int *idx, int nidx, int write_back, gcb_comm *comm, double *sum);
/* The same sums for the atoms of `want` only: sum_want[k] belongs to atom want[k] (nwant doubles of host memory).
 * align.LOVO looks at the means of its candidate atoms only (rmsdFilter, align/lovo.go:519-528); when those are a small
 * part of the frame -- a protein in its solvent box -- this reads their sectors instead of streaming whole frames.
 * Frames are left untouched. */
int gcb_peratom_dev_sum_atoms(gcb_traj *t, long first, long nframes, const double *ref_aos, const int *idx, int nidx,
                              const int *want, int nwant, gcb_comm *comm, double *sum_want);

/* ---- align.LOVO's fixed-point iteration over resident frames (align/lovo.go:212-252), bookkeeping on the device ----
 * Runs RMSDTraj passes (fit on the current subset of the candidate atoms `cand`, ascending, per-atom mean deviations of all
 * atoms over total_frames frames, sums combined across the ranks of comm when given) until the subset below less_than_rmsd
 * repeats.  Between two passes one CTA does what lovo.go:229-252 does on the host -- rmsdFilter, sort.Stable by value,
 * mobileLimit, sameElementsInt / approxEq, disagreementInt -- and prepares the next template in place; the host reads 64 bytes
 * per iteration.  Outputs (ncand entries each but the scalars): n and iterations (LOVOReturn.N, .Iterations), the last
 * disagreement, indexesold (atoms; its first n entries are LOVOReturn.Natoms), ids / vals = the candidates and their mean
 * deviations in the order of the last sort (LOVOReturn.RMSD is vals re-ordered by atom).
 * Returns 0, a GCB_ERR_*, or 1 = not taken (more than 16384 candidates, candidates not ascending or at most an eighth of the
 * frame, frames that do not go to the cluster kernel, values the host must sort): run the loop on the host with
 * gcb_peratom_dev_sum then -- the answers are the same bit for bit. */
int gcb_lovo_resident(gcb_traj *t, long first, long nframes, const double *ref_aos, const int *cand, int ncand,
                      double total_frames, double less_than_rmsd, int minimum_n, int max_iter, gcb_comm *comm, int *n_out,
                      int *iterations_out, int *disagreement_out, int *indexesold_out, int *ids_out, double *vals_out);

/* ---- all-pairs frame RMSD matrix (BASELINE configs[3]) ----
 * D[i][j] = chem.RMSD(frame_i, frame_j) over the atoms in idx (NULL = all), no superposition inside (geometric.go:
 * 232-288; superpose first with gcb_super_rmsd(write_back = 1)).  Computed as one fp64 tensor-core GEMM (DMMA) on
 * mean-centred coordinates; near-duplicate pairs are recomputed from explicit differences.  Rows [row0, row0+nrows)
 * of the nframes x nframes matrix are written to out_host (row-major, ld = nframes): ranks of a multi-GPU run take
 * disjoint row blocks.  gemm_ms (optional) = device time of the GEMM kernel, redone = pairs recomputed explicitly. */
int gcb_allpairs_rmsd(gcb_traj *t, long first, long nframes, const int *idx, int nidx, long row0, long nrows,
                      double *out_host, float *gemm_ms, long *redone);

/* Same matrix with the optimal superposition of every pair: D[i][j] = chem.RMSD(chem.Super(frame_i, frame_j), frame_j)
 * (handy.go:175-214 + geometric.go:232-288 for each of the F^2 pairs).  The 3x3 correlation blocks of all pairs come out
 * of one DMMA GEMM over the coordinate planes; the rotation solve and the RMSD run per pair in its epilogue. */
int gcb_allpairs_rmsd_fit(gcb_traj *t, long first, long nframes, const int *idx, int nidx, long row0, long nrows,
                          double *out_host, float *gemm_ms, long *redone);

/* Multi-GPU form (one process per GPU, all on one NVLink/NVSwitch node).  Every rank holds ALL nframes frames (see
 * gcb_traj_allgather).  The tiles on or above the diagonal are dealt round-robin to the ranks, so the GEMM work is split
 * evenly; rank r owns the rows [r*base + min(r, rem), (r+1)*base + min(r+1, rem)) (base = nframes / nranks) and the
 * epilogue of every tile stores the tile and its mirror image straight into the owners' row blocks through peer mappings
 * of their device memory (cudaIpc over NVLink) -- compute and gather are one kernel, there is no separate exchange.
 * Collective: every rank of comm must call it with the same arguments.  out_host_rows (may be NULL: the block stays in
 * the communicator's device buffer) receives this rank's nrows x nframes block; row0_out / nrows_out say which rows.
 * gemm_ms = this rank's GEMM kernel, total_ms = first barrier to last barrier on the device (both optional). */
int gcb_allpairs_rmsd_sharded(gcb_traj *t, long first, long nframes, const int *idx, int nidx, int fit, gcb_comm *comm,
                              double *out_host_rows, long *row0_out, long *nrows_out, float *gemm_ms, float *total_ms,
                              long *redone);

/* ---- per-frame centre of mass and moment tensor ----
 * chem.CenterOfMass (geometric.go:609-631) and chem.MomentTensor (geometric.go:705-733) for every frame of the window:
 * center[3F], moment[9F] (row-major symmetric 3x3).  mass: natoms entries or NULL (geometric centre, unit masses);
 * idx/nidx: atoms considered (NULL = all).  Either output may be NULL. */
int gcb_moments(gcb_traj *t, long first, long nframes, const double *mass, const int *idx, int nidx, double *center,
                double *moment);
/* Launch only (results stay on the device) and the matching fetch of the first nframes results of the last launch. */
int gcb_moments_async(gcb_traj *t, long first, long nframes, const double *mass, const int *idx, int nidx);
int gcb_moments_fetch(gcb_traj *t, long nframes, double *center, double *moment);

/* ---- solvent shells around a solute ----
 * The frame loop of solv.ConcMolRDF (solv/solvation.go:140-218): cdf_sum[k], k < (int)(end / step), is the sum over the frames
 * of solv.FrameUMolCRDF (:380-410) -- per frame, the number of solvent residues whose distance to the solute (solv.DistRank,
 * :447-541) is below (k+1)*step, minus one when that number is not zero.  MDFFromCDF (:317-352) is left to the host.
 *   molid[natoms], chain[natoms]: Atom.MolID and an integer id of Atom.Chain (0 = the empty string).
 *   flags[natoms]: bit 0 = Atom.MolName is in the `residues` argument, bit 1 = it is SOL / WAT / HOH (atoms of those are
 *                  passed over when further than end + 3 from the first solute atom, :480-483).
 *   refidx[nref]: the solute atoms; com: solv.Options.COM (distance from the residue's centre).
 *   cpus: frames are consumed in chunks of cpus; a trailing partial chunk is not read (as with a DCD source, dcd.go:520-523).
 *   comm (may be NULL): ranks hold disjoint frame shards; the counts and frames_read are summed over the ranks.
 * A (MolID, chain) pair may name several separate solvent residues (residue numbers wrap at 9999 in PDB topologies): the
 * reference's molid_skip / chain_skip rule (:486, :537-538) is applied across them as it would be atom by atom. */
int gcb_rdf_cdf_sum(gcb_traj *t, long first, long nframes, const int *molid, const int *chain, const unsigned char *flags,
                    const int *refidx, int nref, int com, double step, double end, int cpus, gcb_comm *comm,
                    double *cdf_sum, long *frames_read);

/* ---- stream timing on the stream the kernels run on ---- */
int gcb_timer_start(gcb_traj *t);
int gcb_timer_stop(gcb_traj *t, float *ms);
int gcb_sync(gcb_traj *t);
/* Number of kernels this library has launched since load (bench.py gpu_launches). */
long gcb_launch_count(void);

/* ---- synthetic trajectories generated on the device (bench / tests only) ---- */
int gcb_synth_fill(gcb_traj *t, long first, long nframes, const double *ref_aos, const double *sigma_atom,
                   unsigned long long seed, long frame_id0, int frame0_is_ref, int through_f32);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ---- */
int gcb_comm_unique_id(char *id128);                       /* rank 0 creates, host side broadcasts the 128 bytes */
int gcb_comm_init(int dev, int nranks, int rank, const char *id128, gcb_comm **out);
int gcb_comm_destroy(gcb_comm *c);
int gcb_allreduce_sum_f64(gcb_comm *c, double *host_vec, long n); /* all-gather + rank-ordered sum */
/* Frames sharded over the ranks -> every rank holds all of them: rank r's counts[r] frames already sit at frame offset
 * counts[0] + ... + counts[r-1] of its own trajectory; grouped ncclBroadcast calls fill in the other ranks' frames. */
int gcb_traj_allgather(gcb_traj *t, gcb_comm *c, const long *counts);

#ifdef __cplusplus
}
#endif
#endif /* GOCHEM_B200_H */
