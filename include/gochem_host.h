/*
 * gochem_host.h -- C shim over the C++ host-side mirror of the reference's Go API
 * (gochem_b200/host/gochem.hpp).  It exists so tests written in Python can call chem.Super, chem.RMSD,
 * align.RMSDTraj, align.LOVO, the DCD reader / writer and the LOVO bookkeeping the way the reference's own
 * tests call the Go functions.  Status 0 = ok (nil error); otherwise gch_last_error() holds err.Error().
 * Arithmetic on coordinates happens on the GPU through include/gochem_b200.h; there is no CPU fallback.
 */
#ifndef GOCHEM_HOST_H
#define GOCHEM_HOST_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

size_t gch_last_error(char *buf, size_t cap);
int gch_last_error_is_last_frame(void);
int gch_set_device(int dev);

/* The 3x3 rotation solve of the kernels (gochem_b200/csrc/kabsch3.cuh, host build): h and rot row-major, rot
 * right-multiplies row vectors.  which: 0 = kernels' choice, 1 = one-sided Jacobi SVD, 2 = polar Newton (returns 1
 * when it declines: det h <= 0 or near-singular). */
int gch_kabsch3(const double *h, double *rot, int which);

/* chem.Super (handy.go:175-214): nlists = number of index slices the Go caller passes (0, 1 or 2).
 * test (ntest x 3, row-major) is modified in place. */
int gch_super(double *test, int ntest, const double *templa, int ntempla, const int *idx0, int n0,
              const int *idx1, int n1, int nlists);
/* chem.RMSD (geometric.go:232-243); *out = -1 on error like the reference. */
int gch_rmsd(const double *test, int ntest, const double *templa, int ntempla, const int *idx0, int n0,
             const int *idx1, int n1, int nlists, double *out);
/* chem.MemPerAtomRMSD (geometric.go:294-321); out has n0 entries (ntest when no lists). */
int gch_mem_per_atom_rmsd(const double *test, int ntest, const double *templa, int ntempla, const int *idx0, int n0,
                          const int *idx1, int n1, int nlists, double *out);
/* chem.RotatorTranslatorToSuper (geometric.go:127-208).  Different row counts give the reference's
 * "goChem: Ill-formed matrices" (geometric.go:130-132); transformed holds ntest x 3. */
int gch_rotator_translator_to_super(const double *test, int ntest, const double *templa, int ntempla, double *transformed,
                                    double *rot, double *trans1, double *trans2);

/* chem.CenterOfMass (geometric.go:609-631) + chem.MomentTensor (geometric.go:705-733); mass NULL = unit masses. */
int gch_center_and_moment(const double *a, int n, const double *mass, double *center, double *moment);

/* align bookkeeping (align/lovo.go:175-188, 306-326) */
int gch_mobile_limit(double limit, int tolerance, const double *sorted_rmsd, int n, int *n_out, double *limit_out);
int gch_approx_eq(double i, double j, double eps);
int gch_same_elements(const int *a, int na, const int *b, int nb);
int gch_disagreement(const int *a, int na, const int *b, int nb);
long gch_selected_frames(long F, int cpus, int begin, int skip, long *out, long cap);
void gch_shard_range(long nselected, int rank, int nranks, long *first, long *count);

/* LOVO iteration state (lovo.go:198-258 without the trajectory pass) */
void *gch_lovo_state_new(const int *fullindexes, int nfull, double less_than_rmsd, int minimum_n);
void gch_lovo_state_free(void *st);
int gch_lovo_state_fit_indexes(void *st, int *out, int cap);
int gch_lovo_state_step(void *st, const double *rmsd, int natoms, int *converged);
int gch_lovo_state_disagreement(void *st);
/* topology for Finish: every atom "CA", chain "A", MolID = i + 1 */
int gch_lovo_state_finish(void *st, int natoms, int frames_read, int *N, int *natoms_out, double *rmsd_out, int *iterations);

/* traj/dcd */
int gch_dcd_write(const char *path, const double *frames_aos, long nframes, int natoms);
int gch_dcd_read(const char *path, double *out_aos, long cap_frames, long *nframes, int *natoms, int use_nextconc_chunk);

/* align.RMSDTraj / align.LOVO on a DCD file.  names / chains / molids describe the topology (natoms entries);
 * NULL names = all "CA", chain "A", MolID i + 1.  comm may be NULL. */
int gch_rmsd_traj(const char *dcd_path, const double *ref, int natoms, const int *indexes, int nidx, int cpus, int begin,
                  int skip, const char *write_traj, int dev, int rank, int nranks, void *comm, double *rmsd, int *frames);
int gch_lovo(const char *dcd_path, const double *ref, int natoms, const char *const *names, const char *const *chains,
             const int *molids, const char *atom_name, int cpus, double less_than_rmsd, int minimum_n, const char *write_traj,
             int dev, int rank, int nranks, void *comm, int *N, int *natoms_out, double *rmsd_out, int *nrmsd, int *frames,
             int *iterations, char *text, size_t text_cap);

/* Same two calls with the trajectory streamed through a device window of 2 x window_frames frames instead of made resident
 * (bounded memory; LOVO then re-reads the file on every iteration like the reference, lovo.go:222). */
int gch_rmsd_traj_streamed(const char *dcd_path, const double *ref, int natoms, const int *indexes, int nidx, int cpus, int begin,
                           int skip, const char *write_traj, long window_frames, int dev, int rank, int nranks, void *comm,
                           double *rmsd, int *frames);
int gch_lovo_streamed(const char *dcd_path, const double *ref, int natoms, int cpus, double less_than_rmsd, int minimum_n,
                      long window_frames, int dev, int *N, int *natoms_out, double *rmsd_out, int *nrmsd, int *frames, int *iterations);

/* The LOVO iteration over a trajectory already resident on the device (a gcb_traj handle holding nlocal of the
 * nselected_total frames all ranks hold together); topology: all "CA", chain "A". */
int gch_lovo_resident(void *traj_handle, long nlocal, long nselected_total, const double *ref, int natoms, int cpus,
                      double less_than_rmsd, int minimum_n, void *comm, int rank, int nranks, int *N, int *natoms_out,
                      double *rmsd_out, int *iterations);

/* traj/xtc over libxdrfile (dlopen; GOCHEM_XDRFILE names the library).  gch_xtc_available: 1 when the codec could be loaded.
 * gch_xtc_write: frames in Angstrom -> .xtc (libxdrfile's writer, precision 1000); gch_xtc_read: frames back in Angstrom through
 * XTCObj.Next (use_raw = 0) or through the raw float32 feed widened and scaled by 10 on the host exactly as the device kernel does. */
int gch_xtc_available(void);
int gch_xtc_write(const char *path, const double *frames_aos, long nframes, int natoms);
int gch_xtc_read(const char *path, double *out_aos, long cap_frames, long *nframes, int *natoms, int use_raw);

/* traj/stf (goChem's own text-in-a-compressed-stream format, traj/stf/stf.go).  The codec follows the last letter of the
 * file name as in the reference: 'z' gzip, 'r' raw deflate (zlib), 'l' LZW, anything else zstd (libzstd through dlopen, GOCHEM_ZSTD
 * names it); gch_stf_zstd_available: 1 when that library could be loaded.  gch_stf_write: frames -> file through StfW
 * (level: flate / gzip level; the reference's default 11 is rejected by those codecs there too).  gch_stf_read: frames back
 * through StfR.Next (use_raw = 0) or through the raw int32 feed divided by its scale on the host exactly as the device
 * kernel does (GCB_I32_AOS). */
int gch_stf_zstd_available(void);
int gch_stf_write(const char *path, const double *frames_aos, long nframes, int natoms, int level, const char *header_text,
                  const double *box9);
int gch_stf_read(const char *path, double *out_aos, long cap_frames, long *nframes, int *natoms, int use_raw, char *header_text,
                 size_t header_cap, double *box9);

/* solv (solv/solvation.go): MDFFromCDF (:317-352), EllipsoidAxes (:39-63; masses NULL = unit masses), and ConcMolRDF
 * (:140-218) on a DCD file.  Topology: molnames / chains (natoms C strings), molids, masses (0 = unknown -> error, as
 * Topology.Masses); coords0 = mol.Coords[0].  rdf / cumulative: (int)(end / step) doubles each.  window_frames > 0 streams
 * the file through a device window instead of making it resident.  comm may be NULL. */
int gch_mdf_from_cdf(double *ret, int n, int framesread, double A, double B, double step, double *ret2);
int gch_ellipsoid_axes(const double *coords, int n, const double *masses, double *rhos3);
/* chem.Rhos + chem.RhoShapeIndexes (geometric.go:511-545) on a 3 x 3 moment tensor; chem.EasyShape (handy.go:103-129) of one
 * structure; and EasyShape for every frame of a window of a device trajectory (moment tensors from gcb_moments, eigenvalues on
 * the host; SURVEY section 8 f rank 3).  A collapsed structure gives -1, -1 and an error. */
int gch_rhos_shape(const double *moment9, double *rhos3, double *linear, double *circular);
int gch_easy_shape(const double *coords, int n, const double *masses, double *linear, double *circular);
int gch_easy_shape_traj(void *traj, long first, long nframes, const double *masses, int nmasses, double *linear, double *circular);
int gch_conc_mol_rdf(const char *dcd_path, const double *coords0, int natoms, const char *const *molnames, const int *molids,
                     const char *const *chains, const double *masses, const int *refidx, int nref, const char *const *residues,
                     int nres, int com, int cpus, double step, double end, long window_frames, int dev, int rank, int nranks,
                     void *comm, double *rdf, double *cumulative, int *frames_read);

#ifdef __cplusplus
}
#endif
#endif
