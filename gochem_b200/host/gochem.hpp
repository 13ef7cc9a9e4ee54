// gochem.hpp -- host-side mirror of the reference's Go API for the superposition / RMSD / LOVO
// path, written above the C ABI of include/gochem_b200.h.
//
// The reference is Go (rmera/gochem); this image has no Go toolchain, so the host layer a Go
// maintainer would write with cgo (INTEGRATION.md shows that binding) is written here in C++ with
// the same names, argument meaning and error behaviour:
//
//   v3::Matrix                      /root/reference/v3/gonum.go:54-58, v3/gocoords.go
//   chem::Super                     handy.go:175-214
//   chem::RMSD / chem::MemRMSD      geometric.go:232-288
//   chem::MemPerAtomRMSD            geometric.go:294-321
//   chem::RotatorTranslatorToSuper  geometric.go:127-208
//   chem::Traj / chem::ConcTraj     interfaces.go:36-64, LastFrameError :119-123
//   chem::MemTraj                   chem.go:693-762 (Molecule as an in-memory ConcTraj)
//   dcd::DCDObj / dcd::DCDWObj      traj/dcd/dcd.go, traj/dcd/dcd_write.go
//   align::Options, LOVO, RMSDTraj  align/lovo_options.go, align/lovo.go
//
// Go's (value, error) returns become an Error return plus output arguments; a default-constructed
// Error is nil.  All arithmetic on coordinates runs on the GPU through the C ABI; nothing here
// computes a superposition on the CPU, and without a CUDA device every compute call returns the
// library's error.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

struct gcb_traj;
struct gcb_comm;

namespace gochem {

// chem.Error (interfaces.go:103-107): message plus the Decorate() call chain.
struct Error {
    bool set = false;
    std::string msg;
    std::vector<std::string> deco;
    bool last_frame = false;  // satisfies chem.LastFrameError (interfaces.go:119-123)
    bool critical = true;
    int code = 0;             // C-ABI status when the error came from the device library
    Error() = default;
    Error(std::string m, std::vector<std::string> d = {}, int c = 0) : set(true), msg(std::move(m)), deco(std::move(d)), code(c) {}
    explicit operator bool() const { return set; }
    std::string String() const;
    Error& Decorate(const std::string& who) { if (!who.empty()) deco.push_back(who); return *this; }
};
Error errDecorate(Error e, const std::string& who);  // chem.go:850-855

namespace v3 {

// N x 3 row-major float64 (v3/gonum.go:54-58).  Methods panic in Go on misuse (v3/gonum.go:513-532);
// here they throw v3::PanicMsg.
struct PanicMsg : std::runtime_error { using std::runtime_error::runtime_error; };

class Matrix {
  public:
    Matrix() = default;
    static Matrix Zeros(int vecs);                                         // gocoords.go:43-47
    static Error NewMatrix(const std::vector<double>& data, Matrix* out);  // gonum.go:79-88
    int NVecs() const { return (int)(d_.size() / 3); }                     // gocoords.go:116-123
    int Len() const { return NVecs(); }
    double At(int i, int j) const;
    void Set(int i, int j, double v);
    std::vector<double>& RawSlice() { return d_; }                         // gonum.go:93-95
    const std::vector<double>& RawSlice() const { return d_; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    void Copy(const Matrix& A);
    void SomeVecs(const Matrix& A, const std::vector<int>& clist);         // gocoords.go:155-166
    void SetVecs(const Matrix& A, const std::vector<int>& clist);          // gocoords.go:139-150
    void AddVec(const Matrix& A, const Matrix& vec);                       // gocoords.go:67-86
    void SubVec(const Matrix& A, const Matrix& vec);                       // gocoords.go:220-224
    void Mul(const Matrix& A, const Matrix& B3x3);                         // gonum.go:220-229 (N x 3 times 3 x 3)
    void Scale(double v, const Matrix& A);
    double Norm(double) const;                                             // gonum.go:268-275 (Frobenius)
    Matrix TrRet() const;                                                  // gonum.go:486-502 (3 x 3 only)
  private:
    std::vector<double> d_;
};

}  // namespace v3

namespace chem {

using IndexLists = std::vector<std::vector<int>>;  // Go's variadic indexes ...[]int

// One process-wide device context (device id, cached trajectory handle for single-frame calls).
Error SetDevice(int dev);
int Device();

Error RotatorTranslatorToSuper(const v3::Matrix& test, const v3::Matrix& templa, v3::Matrix* transformed, v3::Matrix* rotation,
                               v3::Matrix* trans1, v3::Matrix* trans2);
Error Super(v3::Matrix& test, const v3::Matrix& templa, const IndexLists& indexes = {});
Error RMSD(const v3::Matrix& test, const v3::Matrix& templa, double* out, const IndexLists& indexes = {});
Error MemRMSD(const v3::Matrix& test, const v3::Matrix& templa, v3::Matrix& tmp, double* out, const IndexLists& indexes = {});
Error MemPerAtomRMSD(const v3::Matrix& test, const v3::Matrix& templa, v3::Matrix* ctest, v3::Matrix* ctempla, v3::Matrix& tmp,
                     std::vector<double>* out, const IndexLists& indexes = {});

// geometric.go:609-631 and :705-733.  mass empty = unit masses (geometric centre).
Error CenterOfMass(const v3::Matrix& geometry, v3::Matrix* center, const std::vector<double>& mass = {});
Error MomentTensor(const v3::Matrix& A, v3::Matrix* moment, const std::vector<double>& mass = {});
// geometric.go:523-545: the eigenvalues of the moment tensor, largest first (v3.EigenWrap sorts them; a smallest eigenvalue
// <= appzero is the "collapsed to a single point" error, with the eigenvalues still returned as the reference does)
Error Rhos(const v3::Matrix& momentTensor, std::vector<double>* rhos, double epsilon = -1);
// geometric.go:511-520: linear (prolate) and circular (oblate) distortion in percent
Error RhoShapeIndexes(const std::vector<double>& rhos, double* linear, double* circular);
// handy.go:103-129: MomentTensor -> Rhos -> RhoShapeIndexes; (-1, -1) with the error on failure; a Masses() failure falls back
// to unit masses and comes back as the error beside valid indexes, as in the reference
class Topology;
Error EasyShape(const v3::Matrix& coords, double epsilon, const Topology* mol, double* linear, double* circular);
// The same for every resident frame of a device trajectory (SURVEY §8 f rank 3): the moment tensors come from one sweep on the
// device (gcb_moments), the 3 x 3 eigenvalues and the indexes are formed on the host.  A frame that fails (collapsed) gets -1, -1
// and the first such error is returned after all frames are done.
Error EasyShapeTraj(gcb_traj* t, long first, long nframes, const std::vector<double>& masses, std::vector<double>* linear,
                    std::vector<double>* circular);

// Topology side: only what align.res2atoms reads (chem.go Atom fields, lovo.go:504-516).
struct Atom {
    std::string Name;
    int MolID = 0;
    std::string Chain;
    std::string MolName;  // residue / molecule name (solv.DistRank reads it, solvation.go:475-476)
    double Mass = 0.0;    // 0 = not known (Topology.Masses fails, chem.go:291-301)
};
class Atomer {  // interfaces.go:67-75
  public:
    virtual ~Atomer() = default;
    virtual const chem::Atom& Atom(int i) const = 0;
    virtual int Len() const = 0;
};
class Topology : public Atomer {
  public:
    std::vector<chem::Atom> Atoms;
    const chem::Atom& Atom(int i) const override { return Atoms.at((size_t)i); }
    int Len() const override { return (int)Atoms.size(); }
    Error Masses(std::vector<double>* out) const;                      // chem.go:291-301
    void SomeAtoms(const Atomer& T2, const std::vector<int>& indexes);  // chem.go: copies the listed atoms
};
// chem.Molecule (chem.go): a topology plus its frames; only what solv.ConcMolRDF reads.
class Molecule : public Topology {
  public:
    std::vector<v3::Matrix> Coords;
};

class Traj {  // interfaces.go:36-49
  public:
    virtual ~Traj() = default;
    virtual bool Readable() const = 0;
    virtual Error Next(v3::Matrix* output, std::vector<double>* box = nullptr) = 0;
    virtual int Len() const = 0;
};

// interfaces.go:51-64.  Go returns one channel per slot; here every non-null slot is filled before the
// call returns.  A null entry means "read and discard".  Readers that own their frames (MemTraj) may
// repoint the slot at their own matrix, as Molecule.NextConc does (chem.go:749-751).
class ConcTraj {
  public:
    virtual ~ConcTraj() = default;
    virtual bool Readable() const = 0;
    virtual Error NextConc(std::vector<v3::Matrix*>& frames) = 0;
    virtual int Len() const = 0;
    virtual void Close() {}
    // Optional fast feed for device staging: fill `dst` (natoms*3 float32) with the next frame in the
    // reader's native layout and say which (GCB_F32_SOA / GCB_F32_AOS) and what scale applies.
    // Default: not available.
    virtual bool HasRawFeed() const { return false; }
    virtual Error NextRaw(float* /*dst*/, int* /*layout*/, double* /*scale*/) { return Error("raw feed not supported"); }
    // Optional bulk feed: the next up-to-`max_frames` frames exactly as they sit in the file, one read() into dst
    // (dst == nullptr: skip them).  Every frame is a record of *record_bytes bytes with float32 x[], y[], z[] at byte
    // offsets off[0..2]; *got frames were delivered (0 with a LastFrameError at the end).
    virtual bool HasRecordFeed() const { return false; }
    virtual long RecordsAvailable() { return -1; }   // whole frames left in the source, when it can tell
    virtual Error NextRecords(char* /*dst*/, long /*max_frames*/, long* /*got*/, long* /*record_bytes*/, long* /*off*/, double* /*scale*/) {
        return Error("record feed not supported");
    }
};

// chem.Molecule used as a RAM-backed trajectory (chem.go:693-762).
class MemTraj : public Traj, public ConcTraj {
  public:
    std::vector<v3::Matrix> Coords;
    bool Readable() const override { return true; }
    Error Next(v3::Matrix* output, std::vector<double>* box = nullptr) override;
    Error NextConc(std::vector<v3::Matrix*>& frames) override;
    int Len() const override { return Coords.empty() ? 0 : Coords[0].NVecs(); }
    void Close() override { current_ = 0; }
  private:
    size_t current_ = 0;
};

Error newLastFrameError(const std::string& file, const std::string& who);

}  // namespace chem

namespace dcd {

// CHARMM / NAMD DCD reader (traj/dcd/dcd.go): little- or big-endian, CHARMM flavour with optional
// unit-cell block and optional 4th dimension; fixed atoms unsupported, X-PLOR unsupported.
class DCDObj : public chem::Traj, public chem::ConcTraj {
  public:
    static Error New(const std::string& filename, std::unique_ptr<DCDObj>* out);
    ~DCDObj() override;
    bool Readable() const override { return readable_; }
    Error Next(v3::Matrix* output, std::vector<double>* box = nullptr) override;
    Error NextConc(std::vector<v3::Matrix*>& frames) override;
    int Len() const override { return natoms_; }
    void Close() override;
    bool HasRawFeed() const override { return true; }
    Error NextRaw(float* dst, int* layout, double* scale) override;  // x[], y[], z[] planes, as stored on disk
    bool HasRecordFeed() const override { return readable_ && !swap_; }
    long RecordsAvailable() override;
    Error NextRecords(char* dst, long max_frames, long* got, long* record_bytes, long* off, double* scale) override;

  private:
    Error initRead(const std::string& name);
    Error nextRaw(float* x, float* y, float* z);
    Error readI32(int32_t* v);
    Error readFloat32Block(int32_t blocksize, float* block, int n);
    Error probeRecord();
    long rec_bytes_ = 0, rec_off_[3] = {0, 0, 0};
    std::string filename_;
    FILE* f_ = nullptr;
    int32_t natoms_ = 0, fixed_ = 0;
    bool readable_ = false, charmm_ = false, extrablock_ = false, fourdim_ = false, readLast_ = false, swap_ = false;
    float box_[6] = {0, 0, 0, 0, 0, 0};
    std::vector<float> fields_;
};

// DCD writer (traj/dcd/dcd_write.go:102-330): same header bytes, frame counter rewritten after every frame.
class DCDWObj {
  public:
    static Error NewWriter(const std::string& filename, int natoms, std::unique_ptr<DCDWObj>* out);
    ~DCDWObj();
    Error WNext(const v3::Matrix& towrite);
    Error WNextRaw(const double* aos, int natoms);
    void Close();
    int Len() const { return natoms_; }

  private:
    std::string filename_;
    FILE* f_ = nullptr;
    int32_t natoms_ = 0, frames_ = 0;
    bool writable_ = false;
    std::vector<float> fields_;
};

}  // namespace dcd

namespace xtc {

// GROMACS XTC reader (traj/xtc/xtc.go) over libxdrfile, the third-party C codec the reference links (traj/xtc/xtc.go:30-38;
// it ships as a binary there).  The library is bound with dlopen when the first file is opened (GOCHEM_XDRFILE names it,
// otherwise libxdrfile.so); without it New() fails.  The codec stays on the CPU; what changes is where it writes: NextRaw
// points it at the caller's (pinned) buffer and the x10 nm -> A, the widening and the transposition happen on the device
// (GCB_F32_AOS with scale 10, traj/xtc/xtc.go:129-134).
class XTCObj : public chem::Traj, public chem::ConcTraj {
  public:
    static Error New(const std::string& filename, std::unique_ptr<XTCObj>* out);   // xtc.go:62-103
    ~XTCObj() override;
    bool Readable() const override { return readable_; }
    Error Next(v3::Matrix* output, std::vector<double>* box = nullptr) override;     // xtc.go:108-146
    Error NextConc(std::vector<v3::Matrix*>& frames) override;                       // xtc.go:189-238
    int Len() const override { return natoms_; }
    void Close() override;                                                           // xtc.go:173-183
    bool HasRawFeed() const override { return true; }
    Error NextRaw(float* dst, int* layout, double* scale) override;                  // float32 xyz per atom, nanometres
    static bool Available(std::string* why = nullptr);
    // test helper: frames in Angstrom written as an .xtc with libxdrfile's own writer (precision 1000: 0.001 nm)
    static Error WriteFile(const std::string& filename, const double* frames_aos_angstrom, long nframes, int natoms);

  private:
    Error readInto(float* dst);
    std::string filename_;
    void* fp_ = nullptr;
    int natoms_ = 0;
    bool readable_ = false;
    std::vector<float> coords_;
    float box_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

}  // namespace xtc

namespace stf {

// goChem's own "simple trajectory format" (traj/stf/stf.go): text lines of three fixed-point integers per atom
// ("%d %d %d\n", coordinate * 100 rounded to even, stf.go:216-225), a line starting with '*' after every frame
// (optionally followed by nine box numbers), a header of key=value lines closed by "** natoms", all of it inside one
// compressed stream chosen by the LAST letter of the file name (stf.go:245-285): 'z' gzip, 'r' raw deflate, 'l' LZW,
// anything else ("stf", "sts") zstd.  gzip and deflate are decoded with zlib, zstd with libzstd bound by dlopen
// (GOCHEM_ZSTD names it, else libzstd.so.1), LZW (Go's compress/lzw, MSB order, 8-bit literals) by a decoder of its own.  The decompressor and the text
// parser stay on the CPU; the raw feed hands the device the integers as they were parsed (GCB_I32_AOS, divisor 100)
// and the division float64(f)/p of stf.go:341 happens there.
//
// Quirk kept: both reader and writer try to honour a "prec" header entry but the test on strconv.Atoi's error is
// inverted (stf.go:174-181, 304-313), so the precision is 2 decimals (p = 100) whatever the header says.
class ByteStream;   // decompressed bytes of the file (defined in gochem.cpp)
class ByteSink;     // compressor in front of the file

class StfR : public chem::Traj, public chem::ConcTraj {
  public:
    static Error New(const std::string& name, std::unique_ptr<StfR>* out, std::map<std::string, std::string>* header = nullptr);  // stf.go:230-318
    ~StfR() override;
    bool Readable() const override { return readable_; }
    Error Next(v3::Matrix* output, std::vector<double>* box = nullptr) override;   // stf.go:352-424
    Error NextConc(std::vector<v3::Matrix*>& frames) override;                     // stf.go:440-455
    int Len() const override { return natoms_; }
    void Close() override;                                                         // stf.go:427-434
    bool HasRawFeed() const override { return true; }
    Error NextRaw(float* dst, int* layout, double* scale) override;                // int32 x y z per atom, divisor 100
    static bool ZstdAvailable(std::string* why = nullptr);

  private:
    Error nextInto(int32_t* ints, v3::Matrix* output, std::vector<double>* box);
    std::string filename_;
    std::unique_ptr<ByteStream> h_;
    int natoms_ = -1;
    bool readable_ = false;
};

class StfW {
  public:
    // stf.go:113-190.  level: flate / gzip level (the reference's default 11 is out of zlib's range and is clamped to 9;
    // zstd files use its best-compression preset as the reference does)
    static Error NewWriter(const std::string& name, int natoms, const std::map<std::string, std::string>& header, std::unique_ptr<StfW>* out,
                           int level = 11);
    ~StfW();
    Error WNext(const v3::Matrix* coord, const std::vector<double>* box = nullptr);   // stf.go:64-110
    int Len() const { return natoms_; }
    void Close();                                                                    // stf.go:39-50

  private:
    std::string filename_;
    std::unique_ptr<ByteSink> h_;
    int natoms_ = 0;
    bool writeable_ = false;
};

}  // namespace stf

namespace solv {

// solv.Options (solv/solvation.go:65-138), same getter/setter idiom as the Go methods.
class Options {
  public:
    bool COM(const bool* c = nullptr) { const bool r = com_; if (c) com_ = *c; return r; }
    int Cpus(const int* n = nullptr) { const int r = cpus_; if (n && *n > 0) cpus_ = *n; return r; }
    int Skip(const int* n = nullptr) { const int r = skip_; if (n && *n > 0) skip_ = *n; return r; }
    double Step(const double* s = nullptr) { const double r = step_; if (s && *s > 0) step_ = *s; return r; }
    double End(const double* e = nullptr) { const double r = end_; if (e && *e > 0) end_ = *e; return r; }
    // additions for the device path
    int DeviceID(const int* d = nullptr) { if (d && *d >= 0) device_ = *d; return device_; }
    void Shard(int rank, int nranks, gcb_comm* comm) { rank_ = rank; nranks_ = nranks; comm_ = comm; }
    // frames per device window when the trajectory is streamed instead of made resident (0 = resident)
    long StreamWindow(const long* frames = nullptr) { if (frames && *frames >= 0) stream_ = *frames; return stream_; }

    long stream_ = 0;
    bool com_ = false;
    int cpus_ = 1;
    double step_ = 0.1, end_ = 10.0;
    int skip_ = 1;
    int device_ = 0, rank_ = 0, nranks_ = 1;
    gcb_comm* comm_ = nullptr;
};
Options* DefaultOptions();  // solvation.go:75-83: com false, cpus = hardware threads, step 0.1, end 10, skip 1

// solvation.go:39-63: the eigenvalues of the moment tensor, largest first (chem.Rhos, geometric.go:523-545).
Error EllipsoidAxes(const v3::Matrix& coords, double epsilon, const chem::Topology* mol, std::vector<double>* rhos);
// solvation.go:317-352.  ret (summed cumulative counts) is overwritten with the normalised density; ret2 = mean cumulative number.
Error MDFFromCDF(std::vector<double>& ret, int framesread, double A, double B, double step, std::vector<double>* ret2);
// solvation.go:140-218.  The frame loop (DistRank + FrameUMolCRDF for every frame) runs on the device.
Error ConcMolRDF(chem::ConcTraj& traj, const chem::Molecule& mol, const std::vector<int>& refindexes,
                 const std::vector<std::string>& residues, std::vector<double>* rdf, std::vector<double>* cumulative,
                 Options* o = nullptr, int* frames_read = nullptr);

}  // namespace solv

namespace align {

// align.Options (lovo_options.go:6-142).  Getter/setter pairs keep the Go names; the Go setters index
// n[0] even without an argument (logical OR where AND was meant, :64-87) and panic -- here the value is
// an optional pointer and a missing value just reads.
class Options {
  public:
    int Begin(const int* n = nullptr) { if (n && *n >= 0) begin_ = *n; return begin_; }
    int Skip(const int* n = nullptr) { if (n && *n >= 0) skip_ = *n; return skip_; }
    int Cpus(const int* n = nullptr) { if (n && *n > 0) cpus_ = *n; return cpus_; }
    double LessThanRMSD(const double* r = nullptr) { if (r) lessThanRMSD_ = *r; return lessThanRMSD_; }
    int MinimumN(const int* n = nullptr) { if (n && *n > 0) minimumN_ = *n; return minimumN_; }
    std::vector<std::string> AtomNames(const std::vector<std::string>* n = nullptr) { if (n && !n->empty()) atomNames_ = *n; return atomNames_; }
    std::vector<std::string> Chains(const std::vector<std::string>* c = nullptr) { if (c) chains_ = *c; return chains_; }
    std::string TrajName(const std::string* n = nullptr) { if (n && !n->empty()) writeTraj_ = *n; return writeTraj_; }
    // additions for the device path, same idiom
    int DeviceID(const int* d = nullptr) { if (d && *d >= 0) device_ = *d; return device_; }
    // frames of rank r of n: contiguous blocks of whole chunks; comm combines the per-atom sums
    void Shard(int rank, int nranks, gcb_comm* comm) { rank_ = rank; nranks_ = nranks; comm_ = comm; }
    // host-side replacement of the collective (gloo in the CPU tests); sums a vector across ranks in place
    void AllReduce(std::function<Error(double*, long)> f) { allreduce_ = std::move(f); }
    int MaxIterations(const int* n = nullptr) { if (n && *n > 0) maxIter_ = *n; return maxIter_; }
    // Frames per device window when the trajectory is streamed instead of made resident (0 = resident; LOVO switches to
    // streaming by itself when the file cannot fit in device memory and then re-reads it on every iteration, as the
    // reference does, lovo.go:222).
    long StreamWindow(const long* frames = nullptr) { if (frames && *frames >= 0) stream_ = *frames; return stream_; }

    int begin_ = 0, skip_ = 0, cpus_ = 1;
    std::vector<std::string> atomNames_, chains_;
    std::string writeTraj_;
    double lessThanRMSD_ = 1.0;
    int minimumN_ = 10;
    int device_ = 0, rank_ = 0, nranks_ = 1, maxIter_ = 1000;
    long stream_ = 0;
    gcb_comm* comm_ = nullptr;
    std::function<Error(double*, long)> allreduce_;
};
Options* DefaultOptions();    // lovo_options.go:22-32: cpus = hardware threads, names {"CA"}, 1.0 A, minimumN 10
Options* DefaultCGOptions();  // :35-41: names {"BB"}

struct MolIDandChain {  // lovo.go:50-65
    int molid = 0;
    std::string chain;
    std::string String() const;
    int MolID() const { return molid; }
    const std::string& Chain() const { return chain; }
};

struct LOVOReturn {  // lovo.go:68-150
    int N = 0;
    int FramesRead = 0;
    std::vector<int> Natoms;
    std::vector<MolIDandChain> Nmols;
    std::vector<double> RMSD;
    int Iterations = 0;
    std::string String() const;
    std::string PyMOLSel() const;
    std::string GMX(const std::string& chain = "") const;
};

// lovo.go:175-188 and its helpers; exposed because the index sets must match the reference bit for bit.
Error mobileLimit(double limit, int tolerance, const std::vector<double>& sorted_rmsd, int* n, double* newlimit);
bool approxEq(double i, double j, double epsilon);                       // :436-442
void stableOrderByValue(const std::vector<double>& v, std::vector<int>* order);   // the permutation of sort.Stable by value, :429-432
bool sameElementsInt(const std::vector<int>& a, const std::vector<int>& b);  // :446-457
int disagreementInt(const std::vector<int>& a, const std::vector<int>& b);   // :459-471
std::vector<int> res2atoms(const chem::Atomer& mol, const std::vector<std::string>& names, const std::vector<std::string>& chains);  // :504-516
std::vector<MolIDandChain> iDs2MolIDandChains(const chem::Atomer& mol, const std::vector<int>& indexes);  // :540-551
// frames RMSDTraj actually processes for a source of F frames (chunks of cpus, begin/skip quirks, dropped tail; :306-326)
std::vector<long> selectedFrames(long F, int cpus, int begin, int skip);
// contiguous share of those frames for one rank (whole frames, balanced to within one)
void shardRange(long nselected, int rank, int nranks, long* first, long* count);

// The bookkeeping of one LOVO run (lovo.go:198-258) without the trajectory pass: feed it the per-atom
// mean deviations of each RMSDTraj pass and it says which atoms to fit next and when to stop.
class LovoState {
  public:
    LovoState(std::vector<int> fullindexes, double lessThanRMSD, int minimumN);
    const std::vector<int>& FitIndexes() const { return fit_; }  // indexesold[0:n]
    // rmsd: mol.Len() per-atom means.  Returns converged; on error *err is set (mobileLimit cannot terminate).
    bool Step(const std::vector<double>& rmsd, Error* err);
    int Disagreement() const { return disagreement_; }
    int Iterations() const { return itercount_; }
    void Finish(const chem::Atomer& mol, int framesRead, LOVOReturn* out) const;
    // state after a loop that ran on the device (gcb_lovo_resident): what Finish reads
    void Restore(std::vector<int> indexesold, std::vector<int> ids, std::vector<double> vals, int n, int iterations, int disagreement);

  private:
    std::vector<int> full_, indexesold_, fit_, ids_;
    std::vector<double> vals_;
    double lessThan_, prevlim_ = 0.0;
    int minimumN_, n_, itercount_ = 0, disagreement_ = 0;
};

// A trajectory made resident on the device once and reused by every LOVO iteration (the reference
// re-opens and re-decodes the file per iteration, lovo.go:222).
class DeviceTrajectory {
  public:
    ~DeviceTrajectory();
    // Reads `traj` to its end with RMSDTraj's chunk rules and stages this rank's share of the selected
    // frames through C-owned pinned buffers (raw float32 feed when the reader has one).
    static Error Load(chem::ConcTraj& traj, int natoms, const Options& o, std::unique_ptr<DeviceTrajectory>* out);
    // Wrap a handle somebody else filled (nlocal resident frames of nselected_total over all ranks); not owned.
    static std::unique_ptr<DeviceTrajectory> Adopt(gcb_traj* t, int natoms, long nlocal, long nselected_total);
    long LocalFrames() const { return nlocal_; }
    long SelectedFrames() const { return nselected_; }
    int Natoms() const { return natoms_; }
    gcb_traj* Handle() { return t_; }
    // one RMSDTraj pass over the resident frames: per-atom mean deviations over ALL selected frames (all ranks)
    // only != nullptr: the caller reads the means of these atoms alone (align.LOVO: its candidates, lovo.go:519-528); when they
    // are at most an eighth of the frame and nothing is written back, only their sectors are read and the other entries are 0
    Error MeanDeviations(const v3::Matrix& ref, const std::vector<int>& indexes, const Options& o, bool write_back,
                         std::vector<double>* mean, const std::vector<int>* only = nullptr);
    Error WriteDCD(const std::string& name);  // local frames, as dcd.DCDWObj would (lovo.go:337-339)

  private:
    gcb_traj* t_ = nullptr;
    bool owned_ = true;
    int natoms_ = 0;
    long nlocal_ = 0, nselected_ = 0;
};

Error RMSDTraj(const chem::Atomer& mol, const v3::Matrix& ref, chem::ConcTraj& traj, const std::vector<int>& indexes, Options* o,
               std::vector<double>* rmsd, int* frames);
// The same pass with bounded memory: the selected frames go through two pinned staging buffers and a device window of
// 2 x o->StreamWindow() frames (the upload of one window overlaps the reader filling the next); ranks take windows in turn.
Error RMSDTrajStreamed(const chem::Atomer& mol, const v3::Matrix& ref, chem::ConcTraj& traj, const std::vector<int>& indexes, Options* o,
                       std::vector<double>* rmsd, int* frames);
Error LOVO(const chem::Atomer& mol, const v3::Matrix& ref, const std::string& trajname, LOVOReturn* out, Options* opt = nullptr);
// The LOVO iteration (lovo.go:221-252) over a trajectory that is already resident on the device.  `reload`, when
// given, restores the frames after a pass that wrote the aligned trajectory back over them.
Error LOVOResident(const chem::Atomer& mol, const v3::Matrix& ref, std::unique_ptr<DeviceTrajectory>& d, LOVOReturn* out, Options* o,
                   const std::function<Error()>& reload = nullptr);
Error opentraj(const std::string& name, std::unique_ptr<chem::ConcTraj>* out);  // lovo.go:152-172
// The frames RMSDTraj selects, streamed through a device window; process(handle, first, n) sees every window of this rank.
Error streamSelected(chem::ConcTraj& traj, int natoms, const Options& o, const std::function<Error(gcb_traj*, long, long)>& process,
                     long* nselected, long* nlocal);

}  // namespace align
}  // namespace gochem
