// capi.cpp -- extern "C" shim of include/gochem_host.h over gochem.hpp (test plumbing for Python).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "../../include/gochem_b200.h"
#include "../../include/gochem_host.h"
#include "../csrc/kabsch3.cuh"
#include "gochem.hpp"

using namespace gochem;

static thread_local std::string t_err;
static thread_local bool t_last_frame = false;

static int fail(const Error& e) {
    t_err = e.String();
    t_last_frame = e.last_frame;
    return e.code ? e.code : -1;
}
static int ret(const Error& e) { return e ? fail(e) : 0; }

static v3::Matrix mat(const double* p, int n) {
    v3::Matrix m = v3::Matrix::Zeros(n);
    if (n > 0) memcpy(m.data(), p, sizeof(double) * 3 * (size_t)n);
    return m;
}

static chem::IndexLists lists(const int* i0, int n0, const int* i1, int n1, int nlists) {
    chem::IndexLists L;
    if (nlists >= 1) L.emplace_back(i0 ? std::vector<int>(i0, i0 + n0) : std::vector<int>());
    if (nlists >= 2) L.emplace_back(i1 ? std::vector<int>(i1, i1 + n1) : std::vector<int>());
    return L;
}

namespace {
struct FlatTopology : chem::Atomer {
    std::vector<chem::Atom> atoms;
    const chem::Atom& Atom(int i) const override { return atoms.at((size_t)i); }
    int Len() const override { return (int)atoms.size(); }
};
FlatTopology topology(int natoms, const char* const* names, const char* const* chains, const int* molids) {
    FlatTopology t;
    t.atoms.resize((size_t)natoms);
    for (int i = 0; i < natoms; i++) {
        t.atoms[(size_t)i].Name = names ? names[i] : "CA";
        t.atoms[(size_t)i].Chain = chains ? chains[i] : "A";
        t.atoms[(size_t)i].MolID = molids ? molids[i] : i + 1;
    }
    return t;
}
}  // namespace

extern "C" {

size_t gch_last_error(char* buf, size_t cap) {
    if (buf && cap) {
        size_t n = t_err.size() < cap - 1 ? t_err.size() : cap - 1;
        memcpy(buf, t_err.data(), n);
        buf[n] = 0;
    }
    return t_err.size();
}
int gch_last_error_is_last_frame(void) { return t_last_frame ? 1 : 0; }
int gch_set_device(int dev) { return ret(chem::SetDevice(dev)); }

// The 3x3 rotation solve the kernels use (csrc/kabsch3.cuh is host + device code): exposed so the CPU tests can
// check it against LAPACK.  which: 0 = the kernels' choice (polar Newton, Jacobi fallback), 1 = Jacobi, 2 = polar.
int gch_kabsch3(const double* h, double* rot, int which) {
    if (which == 1) return k3::kabsch3_jacobi(h, rot);
    if (which == 2) return k3::kabsch3_polar(h, rot) ? 0 : 1;
    return k3::kabsch3(h, rot);
}

int gch_super(double* test, int ntest, const double* templa, int ntempla, const int* i0, int n0, const int* i1, int n1, int nlists) {
    v3::Matrix t = mat(test, ntest), m = mat(templa, ntempla);
    Error e = chem::Super(t, m, lists(i0, n0, i1, n1, nlists));
    if (e) return fail(e);
    memcpy(test, t.data(), sizeof(double) * 3 * (size_t)ntest);
    return 0;
}

int gch_rmsd(const double* test, int ntest, const double* templa, int ntempla, const int* i0, int n0, const int* i1, int n1,
             int nlists, double* out) {
    return ret(chem::RMSD(mat(test, ntest), mat(templa, ntempla), out, lists(i0, n0, i1, n1, nlists)));
}

int gch_mem_per_atom_rmsd(const double* test, int ntest, const double* templa, int ntempla, const int* i0, int n0, const int* i1,
                          int n1, int nlists, double* out) {
    const int n = (nlists == 0 || n0 == 0) ? ntest : n0;
    v3::Matrix tmp = v3::Matrix::Zeros(n), ct = v3::Matrix::Zeros(n), cm = v3::Matrix::Zeros(n);
    std::vector<double> r;
    Error e = chem::MemPerAtomRMSD(mat(test, ntest), mat(templa, ntempla), &ct, &cm, tmp, &r, lists(i0, n0, i1, n1, nlists));
    if (e) return fail(e);
    memcpy(out, r.data(), sizeof(double) * r.size());
    return 0;
}

int gch_rotator_translator_to_super(const double* test, int ntest, const double* templa, int ntempla, double* transformed, double* rot,
                                    double* t1, double* t2) {
    v3::Matrix tr, r, a, b;
    Error e = chem::RotatorTranslatorToSuper(mat(test, ntest), mat(templa, ntempla), transformed ? &tr : nullptr, &r, &a, &b);
    if (e) return fail(e);
    if (transformed) memcpy(transformed, tr.data(), sizeof(double) * 3 * (size_t)ntest);
    if (rot) memcpy(rot, r.data(), sizeof(double) * 9);
    if (t1) memcpy(t1, a.data(), sizeof(double) * 3);
    if (t2) memcpy(t2, b.data(), sizeof(double) * 3);
    return 0;
}

int gch_center_and_moment(const double* a, int n, const double* mass, double* center, double* moment) {
    std::vector<double> m;
    if (mass) m.assign(mass, mass + n);
    v3::Matrix c, mo;
    if (Error e = chem::CenterOfMass(mat(a, n), &c, m)) return fail(e);
    if (Error e = chem::MomentTensor(mat(a, n), &mo, m)) return fail(e);
    if (center) memcpy(center, c.data(), sizeof(double) * 3);
    if (moment) memcpy(moment, mo.data(), sizeof(double) * 9);
    return 0;
}

int gch_mobile_limit(double limit, int tolerance, const double* v, int n, int* n_out, double* limit_out) {
    return ret(align::mobileLimit(limit, tolerance, std::vector<double>(v, v + n), n_out, limit_out));
}
int gch_approx_eq(double i, double j, double eps) { return align::approxEq(i, j, eps) ? 1 : 0; }
int gch_same_elements(const int* a, int na, const int* b, int nb) {
    return align::sameElementsInt(std::vector<int>(a, a + na), std::vector<int>(b, b + nb)) ? 1 : 0;
}
int gch_disagreement(const int* a, int na, const int* b, int nb) {
    return align::disagreementInt(std::vector<int>(a, a + na), std::vector<int>(b, b + nb));
}
long gch_selected_frames(long F, int cpus, int begin, int skip, long* out, long cap) {
    std::vector<long> s = align::selectedFrames(F, cpus, begin, skip);
    for (long i = 0; i < (long)s.size() && i < cap; i++) out[i] = s[(size_t)i];
    return (long)s.size();
}
void gch_shard_range(long nsel, int rank, int nranks, long* first, long* count) { align::shardRange(nsel, rank, nranks, first, count); }

void* gch_lovo_state_new(const int* full, int nfull, double less_than, int minimum_n) {
    return new align::LovoState(std::vector<int>(full, full + nfull), less_than, minimum_n);
}
void gch_lovo_state_free(void* st) { delete static_cast<align::LovoState*>(st); }
int gch_lovo_state_fit_indexes(void* st, int* out, int cap) {
    const std::vector<int>& f = static_cast<align::LovoState*>(st)->FitIndexes();
    for (int i = 0; i < (int)f.size() && i < cap; i++) out[i] = f[(size_t)i];
    return (int)f.size();
}
int gch_lovo_state_step(void* st, const double* rmsd, int natoms, int* converged) {
    Error e;
    bool c = static_cast<align::LovoState*>(st)->Step(std::vector<double>(rmsd, rmsd + natoms), &e);
    if (e) return fail(e);
    *converged = c ? 1 : 0;
    return 0;
}
int gch_lovo_state_disagreement(void* st) { return static_cast<align::LovoState*>(st)->Disagreement(); }
int gch_lovo_state_finish(void* st, int natoms, int frames_read, int* N, int* natoms_out, double* rmsd_out, int* iterations) {
    FlatTopology top = topology(natoms, nullptr, nullptr, nullptr);
    align::LOVOReturn r;
    static_cast<align::LovoState*>(st)->Finish(top, frames_read, &r);
    *N = r.N;
    *iterations = r.Iterations;
    for (int i = 0; i < r.N; i++) natoms_out[i] = r.Natoms[(size_t)i];
    for (size_t i = 0; i < r.RMSD.size(); i++) rmsd_out[i] = r.RMSD[i];
    return 0;
}

int gch_dcd_write(const char* path, const double* frames, long nframes, int natoms) {
    std::unique_ptr<dcd::DCDWObj> w;
    if (Error e = dcd::DCDWObj::NewWriter(path, natoms, &w)) return fail(e);
    for (long f = 0; f < nframes; f++) {
        v3::Matrix m = mat(frames + (size_t)f * 3 * (size_t)natoms, natoms);
        if (Error e = w->WNext(m)) return fail(e);
    }
    w->Close();
    return 0;
}

int gch_dcd_read(const char* path, double* out, long cap, long* nframes, int* natoms, int chunk) {
    std::unique_ptr<dcd::DCDObj> d;
    if (Error e = dcd::DCDObj::New(path, &d)) return fail(e);
    *natoms = d->Len();
    long n = 0;
    const size_t fsz = 3 * (size_t)d->Len();
    if (chunk < 0) {
        // the bulk record feed (NextRecords) in windows of -chunk frames, unpacked on the host: what the device kernel
        // upload_records_kernel does with the same bytes
        if (!d->HasRecordFeed()) return fail(Error("record feed not available for this file"));
        const long win = -(long)chunk;
        std::vector<char> buf;
        for (;;) {
            long got = 0, rb = 0, off[3] = {0, 0, 0};
            double sc = 1.0;
            if (buf.empty()) {
                if (Error e = d->NextRecords(nullptr, 0, &got, &rb, off, &sc)) { if (e.last_frame) break; return fail(e); }
                buf.resize((size_t)rb * (size_t)win);
            }
            Error e = d->NextRecords(buf.data(), win, &got, &rb, off, &sc);
            if (e) {
                if (!e.last_frame) return fail(e);
                break;
            }
            for (long f = 0; f < got; f++, n++) {
                if (n >= cap) continue;
                const char* r = buf.data() + (size_t)f * (size_t)rb;
                for (int i = 0; i < d->Len(); i++)
                    for (int c = 0; c < 3; c++) {
                        float v;
                        memcpy(&v, r + off[c] + 4 * (size_t)i, 4);
                        out[(size_t)n * fsz + 3 * (size_t)i + c] = sc * (double)v;
                    }
            }
        }
    } else if (chunk == 0) {
        v3::Matrix m = v3::Matrix::Zeros(d->Len());
        for (;;) {
            Error e = d->Next(&m);
            if (e) {
                if (!e.last_frame) return fail(e);
                break;
            }
            if (n < cap) memcpy(out + (size_t)n * fsz, m.data(), sizeof(double) * fsz);
            n++;
        }
    } else {
        std::vector<v3::Matrix> ms((size_t)chunk);
        for (auto& m : ms) m = v3::Matrix::Zeros(d->Len());
        for (;;) {
            std::vector<v3::Matrix*> slots;
            for (auto& m : ms) slots.push_back(&m);
            Error e = d->NextConc(slots);
            if (e) {
                if (!e.last_frame) return fail(e);
                break;  // a partial chunk is dropped, as in RMSDTraj
            }
            for (auto* m : slots) {
                if (n < cap) memcpy(out + (size_t)n * fsz, m->data(), sizeof(double) * fsz);
                n++;
            }
        }
    }
    *nframes = n;
    return 0;
}

int gch_rmsd_traj(const char* path, const double* ref, int natoms, const int* indexes, int nidx, int cpus, int begin, int skip,
                  const char* write_traj, int dev, int rank, int nranks, void* comm, double* rmsd, int* frames) {
    std::unique_ptr<chem::ConcTraj> traj;
    if (Error e = align::opentraj(path, &traj)) return fail(e);
    FlatTopology top = topology(natoms, nullptr, nullptr, nullptr);
    align::Options o;
    o.Cpus(&cpus);
    o.Begin(&begin);
    o.Skip(&skip);
    o.DeviceID(&dev);
    o.Shard(rank, nranks, static_cast<gcb_comm*>(comm));
    if (write_traj && *write_traj) { std::string w(write_traj); o.TrajName(&w); }
    std::vector<double> r;
    Error e = align::RMSDTraj(top, mat(ref, natoms), *traj, std::vector<int>(indexes, indexes + nidx), &o, &r, frames);
    if (e) return fail(e);
    memcpy(rmsd, r.data(), sizeof(double) * r.size());
    return 0;
}

int gch_rmsd_traj_streamed(const char* path, const double* ref, int natoms, const int* indexes, int nidx, int cpus, int begin, int skip,
                           const char* write_traj, long window_frames, int dev, int rank, int nranks, void* comm, double* rmsd,
                           int* frames) {
    std::unique_ptr<chem::ConcTraj> traj;
    if (Error e = align::opentraj(path, &traj)) return fail(e);
    FlatTopology top = topology(natoms, nullptr, nullptr, nullptr);
    align::Options o;
    o.Cpus(&cpus);
    o.Begin(&begin);
    o.Skip(&skip);
    o.DeviceID(&dev);
    o.StreamWindow(&window_frames);
    o.Shard(rank, nranks, static_cast<gcb_comm*>(comm));
    if (write_traj && *write_traj) { std::string w(write_traj); o.TrajName(&w); }
    std::vector<double> r;
    Error e = align::RMSDTrajStreamed(top, mat(ref, natoms), *traj, std::vector<int>(indexes, indexes + nidx), &o, &r, frames);
    if (e) return fail(e);
    memcpy(rmsd, r.data(), sizeof(double) * r.size());
    return 0;
}

int gch_lovo_streamed(const char* path, const double* ref, int natoms, int cpus, double less_than, int minimum_n, long window_frames,
                      int dev, int* N, int* natoms_out, double* rmsd_out, int* nrmsd, int* frames, int* iterations) {
    FlatTopology top = topology(natoms, nullptr, nullptr, nullptr);
    std::unique_ptr<align::Options> o(align::DefaultOptions());
    o->Cpus(&cpus);
    o->LessThanRMSD(&less_than);
    o->MinimumN(&minimum_n);
    o->DeviceID(&dev);
    o->StreamWindow(&window_frames);
    align::LOVOReturn r;
    Error e = align::LOVO(top, mat(ref, natoms), path, &r, o.get());
    if (e) return fail(e);
    *N = r.N;
    *frames = r.FramesRead;
    *iterations = r.Iterations;
    *nrmsd = (int)r.RMSD.size();
    for (int i = 0; i < r.N; i++) natoms_out[i] = r.Natoms[(size_t)i];
    for (size_t i = 0; i < r.RMSD.size(); i++) rmsd_out[i] = r.RMSD[i];
    return 0;
}

int gch_lovo(const char* path, const double* ref, int natoms, const char* const* names, const char* const* chains, const int* molids,
             const char* atom_name, int cpus, double less_than, int minimum_n, const char* write_traj, int dev, int rank, int nranks,
             void* comm, int* N, int* natoms_out, double* rmsd_out, int* nrmsd, int* frames, int* iterations, char* text,
             size_t text_cap) {
    FlatTopology top = topology(natoms, names, chains, molids);
    std::unique_ptr<align::Options> o(align::DefaultOptions());
    o->Cpus(&cpus);
    o->LessThanRMSD(&less_than);
    o->MinimumN(&minimum_n);
    o->DeviceID(&dev);
    o->Shard(rank, nranks, static_cast<gcb_comm*>(comm));
    if (atom_name && *atom_name) { std::vector<std::string> n{atom_name}; o->AtomNames(&n); }
    if (write_traj && *write_traj) { std::string w(write_traj); o->TrajName(&w); }
    align::LOVOReturn r;
    Error e = align::LOVO(top, mat(ref, natoms), path, &r, o.get());
    if (e) return fail(e);
    *N = r.N;
    *frames = r.FramesRead;
    *iterations = r.Iterations;
    *nrmsd = (int)r.RMSD.size();
    for (int i = 0; i < r.N; i++) natoms_out[i] = r.Natoms[(size_t)i];
    for (size_t i = 0; i < r.RMSD.size(); i++) rmsd_out[i] = r.RMSD[i];
    if (text && text_cap) {
        std::string s = r.PyMOLSel() + "\n" + r.GMX("A") + "\n" + r.String();
        size_t n = s.size() < text_cap - 1 ? s.size() : text_cap - 1;
        memcpy(text, s.data(), n);
        text[n] = 0;
    }
    return 0;
}

int gch_lovo_resident(void* traj_handle, long nlocal, long nselected_total, const double* ref, int natoms, int cpus,
                      double less_than, int minimum_n, void* comm, int rank, int nranks, int* N, int* natoms_out, double* rmsd_out,
                      int* iterations) {
    // the all-CA topology of the synthetic configs is the caller's in real use: built once per atom count, not per call
    static thread_local FlatTopology top;
    if (top.Len() != natoms) top = topology(natoms, nullptr, nullptr, nullptr);
    std::unique_ptr<align::Options> o(align::DefaultOptions());
    o->Cpus(&cpus);
    o->LessThanRMSD(&less_than);
    o->MinimumN(&minimum_n);
    o->Shard(rank, nranks, static_cast<gcb_comm*>(comm));
    std::unique_ptr<align::DeviceTrajectory> d = align::DeviceTrajectory::Adopt(static_cast<gcb_traj*>(traj_handle), natoms, nlocal, nselected_total);
    align::LOVOReturn r;
    Error e = align::LOVOResident(top, mat(ref, natoms), d, &r, o.get());
    if (e) return fail(e);
    *N = r.N;
    *iterations = r.Iterations;
    for (int i = 0; i < r.N; i++) natoms_out[i] = r.Natoms[(size_t)i];
    for (size_t i = 0; i < r.RMSD.size(); i++) rmsd_out[i] = r.RMSD[i];
    return 0;
}


int gch_mdf_from_cdf(double* r, int n, int framesread, double A, double B, double step, double* ret2) {
    std::vector<double> v(r, r + n), v2;
    if (Error e = solv::MDFFromCDF(v, framesread, A, B, step, &v2)) return fail(e);
    for (int i = 0; i < n; i++) { r[i] = v[(size_t)i]; ret2[i] = v2[(size_t)i]; }
    return 0;
}

int gch_ellipsoid_axes(const double* coords, int n, const double* masses, double* rhos3) {
    chem::Topology top;
    top.Atoms.resize((size_t)n);
    for (int i = 0; i < n; i++) top.Atoms[(size_t)i].Mass = masses ? masses[i] : 1.0;
    std::vector<double> rhos;
    if (Error e = solv::EllipsoidAxes(mat(coords, n), 0.0001, &top, &rhos)) return fail(e);
    for (int k = 0; k < 3; k++) rhos3[k] = rhos[(size_t)k];
    return 0;
}

int gch_conc_mol_rdf(const char* dcd_path, const double* coords0, int natoms, const char* const* molnames, const int* molids,
                     const char* const* chains, const double* masses, const int* refidx, int nref, const char* const* residues,
                     int nres, int com, int cpus, double step, double end, long window_frames, int dev, int rank, int nranks, void* comm,
                     double* rdf, double* cumulative, int* frames_read) {
    chem::Molecule mol;
    mol.Atoms.resize((size_t)natoms);
    for (int i = 0; i < natoms; i++) {
        chem::Atom& a = mol.Atoms[(size_t)i];
        a.MolName = molnames[i];
        a.MolID = molids[i];
        a.Chain = chains[i];
        a.Mass = masses ? masses[i] : 0.0;
    }
    mol.Coords.push_back(mat(coords0, natoms));
    std::unique_ptr<solv::Options> o(solv::DefaultOptions());
    const bool c = com != 0;
    o->COM(&c);
    o->Cpus(&cpus);
    o->Step(&step);
    o->End(&end);
    o->DeviceID(&dev);
    o->StreamWindow(&window_frames);
    o->Shard(rank, nranks, static_cast<gcb_comm*>(comm));
    std::unique_ptr<chem::ConcTraj> traj;
    if (Error e = align::opentraj(dcd_path, &traj)) return fail(e);
    std::vector<std::string> res;
    for (int i = 0; i < nres; i++) res.emplace_back(residues[i]);
    std::vector<double> r1, r2;
    int fr = 0;
    Error e;
    try {
        e = solv::ConcMolRDF(*traj, mol, std::vector<int>(refidx, refidx + nref), res, &r1, &r2, o.get(), &fr);
    } catch (const v3::PanicMsg& p) {
        e = Error(std::string("panic: ") + p.what());
    }
    traj->Close();
    if (e) return fail(e);
    for (size_t i = 0; i < r1.size(); i++) { rdf[i] = r1[i]; cumulative[i] = r2[i]; }
    if (frames_read) *frames_read = fr;
    return 0;
}


int gch_xtc_available(void) { return xtc::XTCObj::Available() ? 1 : 0; }

int gch_xtc_write(const char* path, const double* frames, long nframes, int natoms) {
    return ret(xtc::XTCObj::WriteFile(path, frames, nframes, natoms));
}

int gch_xtc_read(const char* path, double* out, long cap, long* nframes, int* natoms, int use_raw) {
    std::unique_ptr<xtc::XTCObj> x;
    if (Error e = xtc::XTCObj::New(path, &x)) return fail(e);
    *natoms = x->Len();
    const size_t fsz = 3 * (size_t)x->Len();
    long n = 0;
    v3::Matrix m = v3::Matrix::Zeros(x->Len());
    std::vector<float> raw(fsz);
    for (;;) {
        Error e;
        if (use_raw) {
            int layout = 0;
            double scale = 0;
            e = x->NextRaw(raw.data(), &layout, &scale);
            if (!e) for (size_t i = 0; i < fsz; i++) m.data()[i] = scale * (double)raw[i];
        } else {
            e = x->Next(&m);
        }
        if (e) {
            if (!e.last_frame) return fail(e);
            break;
        }
        if (n < cap) memcpy(out + (size_t)n * fsz, m.data(), sizeof(double) * fsz);
        n++;
    }
    *nframes = n;
    return 0;
}


/* chem.Rhos + chem.RhoShapeIndexes on a 3 x 3 moment tensor (CPU only) */
int gch_rhos_shape(const double* moment9, double* rhos3, double* linear, double* circular) {
    v3::Matrix m = v3::Matrix::Zeros(3);
    memcpy(m.data(), moment9, sizeof(double) * 9);
    std::vector<double> rhos;
    Error e = chem::Rhos(m, &rhos);
    for (size_t i = 0; i < 3 && i < rhos.size(); i++) rhos3[i] = rhos[i];
    if (e) { *linear = -1; *circular = -1; return fail(e); }
    return ret(chem::RhoShapeIndexes(rhos, linear, circular));
}

/* chem.EasyShape of one structure (masses may be NULL) */
int gch_easy_shape(const double* coords, int n, const double* masses, double* linear, double* circular) {
    v3::Matrix c = v3::Matrix::Zeros(n);
    memcpy(c.data(), coords, sizeof(double) * 3 * (size_t)n);
    chem::Topology top;
    if (masses) {
        top.Atoms.resize((size_t)n);
        for (int i = 0; i < n; i++) top.Atoms[(size_t)i].Mass = masses[i];
    }
    return ret(chem::EasyShape(c, -1, masses ? &top : nullptr, linear, circular));
}

/* chem.EasyShape for every frame of a window of a device trajectory (gcb_traj handle) */
int gch_easy_shape_traj(void* traj, long first, long nframes, const double* masses, int nmasses, double* linear, double* circular) {
    std::vector<double> m, lin, circ;
    if (masses) m.assign(masses, masses + nmasses);
    Error e = chem::EasyShapeTraj((gcb_traj*)traj, first, nframes, m, &lin, &circ);
    for (long f = 0; f < nframes && (size_t)f < lin.size(); f++) { linear[f] = lin[(size_t)f]; circular[f] = circ[(size_t)f]; }
    return ret(e);
}

int gch_stf_zstd_available(void) { return stf::StfR::ZstdAvailable() ? 1 : 0; }

/* header_text: "key=value\n" lines or NULL */
int gch_stf_write(const char* path, const double* frames, long nframes, int natoms, int level, const char* header_text, const double* box9) {
    std::map<std::string, std::string> header;
    if (header_text) {
        std::string h(header_text);
        size_t a = 0;
        while (a < h.size()) {
            size_t b = h.find('\n', a);
            if (b == std::string::npos) b = h.size();
            const std::string line = h.substr(a, b - a);
            const size_t eq = line.find('=');
            if (eq != std::string::npos) header[line.substr(0, eq)] = line.substr(eq + 1);
            a = b + 1;
        }
    }
    std::unique_ptr<stf::StfW> w;
    if (Error e = stf::StfW::NewWriter(path, natoms, header, &w, level)) return fail(e);
    v3::Matrix m = v3::Matrix::Zeros(natoms);
    std::vector<double> box;
    if (box9) box.assign(box9, box9 + 9);
    for (long f = 0; f < nframes; f++) {
        memcpy(m.data(), frames + (size_t)f * 3 * (size_t)natoms, sizeof(double) * 3 * (size_t)natoms);
        if (Error e = w->WNext(&m, box9 ? &box : nullptr)) return fail(e);
    }
    w->Close();
    return 0;
}

/* use_raw = 1: through the int32 feed the device staging uses, divided by its scale on the host as the device kernel does;
 * header_text (cap bytes) receives the header map as sorted "key=value\n" lines; box9 the box of the LAST frame read */
int gch_stf_read(const char* path, double* out, long cap, long* nframes, int* natoms, int use_raw, char* header_text, size_t header_cap,
                 double* box9) {
    std::unique_ptr<stf::StfR> x;
    std::map<std::string, std::string> header;
    if (Error e = stf::StfR::New(path, &x, &header)) return fail(e);
    if (header_text && header_cap) {
        std::string h;
        for (const auto& kv : header) h += kv.first + "=" + kv.second + "\n";
        snprintf(header_text, header_cap, "%s", h.c_str());
    }
    *natoms = x->Len();
    const size_t fsz = 3 * (size_t)x->Len();
    long n = 0;
    v3::Matrix m = v3::Matrix::Zeros(x->Len());
    std::vector<int32_t> raw(fsz);
    std::vector<double> box(9, 0.0);
    for (;;) {
        Error e;
        if (use_raw) {
            int layout = 0;
            double scale = 0;
            e = x->NextRaw(reinterpret_cast<float*>(raw.data()), &layout, &scale);
            if (!e && layout != GCB_I32_AOS) return fail(Error("stf raw feed announced layout " + std::to_string(layout)));
            if (!e) for (size_t i = 0; i < fsz; i++) m.data()[i] = (double)raw[i] / scale;
        } else {
            e = x->Next(&m, box9 ? &box : nullptr);
        }
        if (e) {
            if (!e.last_frame) return fail(e);
            break;
        }
        if (n < cap) memcpy(out + (size_t)n * fsz, m.data(), sizeof(double) * fsz);
        n++;
    }
    if (box9) memcpy(box9, box.data(), sizeof(double) * 9);
    *nframes = n;
    return 0;
}

}  // extern "C"
