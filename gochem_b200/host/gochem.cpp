// gochem.cpp -- implementation of the host-side mirror declared in gochem.hpp.
// Everything that touches coordinates numerically goes through the C ABI (include/gochem_b200.h).
#include "gochem.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>

#include "../../include/gochem_b200.h"

namespace gochem {

std::string Error::String() const {
    std::string s = msg;
    for (const auto& d : deco) s += " <- " + d;
    return s;
}

Error errDecorate(Error e, const std::string& who) {
    e.Decorate(who);
    return e;
}

static Error deviceError(int rc, const std::string& who) {
    char buf[1024];
    gcb_last_error(buf, sizeof buf);
    return Error(std::string(buf), {who}, rc);
}

// ------------------------------------------------------------------ v3.Matrix
namespace v3 {

Matrix Matrix::Zeros(int vecs) {
    Matrix m;
    m.d_.assign(3 * (size_t)(vecs > 0 ? vecs : 0), 0.0);
    return m;
}

Error Matrix::NewMatrix(const std::vector<double>& data, Matrix* out) {
    if (data.size() % 3 != 0) {
        char b[160];
        snprintf(b, sizeof b, "NewMatrix: Input slice lenght %zu not divisible by %d: %zu", data.size() / 3, 3, (data.size() / 3) % 3);
        return Error(b);
    }
    out->d_ = data;
    return Error();
}

double Matrix::At(int i, int j) const {
    if (i < 0 || i >= NVecs() || j < 0 || j > 2) throw PanicMsg("goChem/v3: Index out of range");
    return d_[3 * (size_t)i + j];
}

void Matrix::Set(int i, int j, double v) {
    if (i < 0 || i >= NVecs() || j < 0 || j > 2) throw PanicMsg("goChem/v3: Index out of range");
    d_[3 * (size_t)i + j] = v;
}

void Matrix::Copy(const Matrix& A) {
    const size_t n = std::min(d_.size(), A.d_.size());
    std::copy(A.d_.begin(), A.d_.begin() + (long)n, d_.begin());
}

void Matrix::SomeVecs(const Matrix& A, const std::vector<int>& clist) {
    if (NVecs() != (int)clist.size() || A.NVecs() < (int)clist.size()) throw PanicMsg("goChem/v3: Wrong shape");
    for (size_t k = 0; k < clist.size(); k++)
        for (int j = 0; j < 3; j++) Set((int)k, j, A.At(clist[k], j));
}

void Matrix::SetVecs(const Matrix& A, const std::vector<int>& clist) {
    if (NVecs() < (int)clist.size() || A.NVecs() < (int)clist.size()) throw PanicMsg("goChem/v3: Wrong shape");
    for (size_t k = 0; k < clist.size(); k++)
        for (int j = 0; j < 3; j++) Set(clist[k], j, A.At((int)k, j));
}

void Matrix::AddVec(const Matrix& A, const Matrix& vec) {
    if (vec.NVecs() != 1 || A.NVecs() != NVecs()) throw PanicMsg("goChem/v3: Wrong shape");
    for (int i = 0; i < NVecs(); i++)
        for (int j = 0; j < 3; j++) d_[3 * (size_t)i + j] = A.d_[3 * (size_t)i + j] + vec.d_[(size_t)j];
}

void Matrix::SubVec(const Matrix& A, const Matrix& vec) {
    if (vec.NVecs() != 1 || A.NVecs() != NVecs()) throw PanicMsg("goChem/v3: Wrong shape");
    for (int i = 0; i < NVecs(); i++)
        for (int j = 0; j < 3; j++) d_[3 * (size_t)i + j] = A.d_[3 * (size_t)i + j] - vec.d_[(size_t)j];
}

void Matrix::Mul(const Matrix& A, const Matrix& B) {
    if (B.NVecs() != 3 || A.NVecs() != NVecs()) throw PanicMsg("goChem/v3: Wrong shape");
    std::vector<double> r(d_.size());
    for (int i = 0; i < A.NVecs(); i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int l = 0; l < 3; l++) s += A.d_[3 * (size_t)i + l] * B.d_[3 * (size_t)l + j];
            r[3 * (size_t)i + j] = s;
        }
    d_.swap(r);
}

void Matrix::Scale(double v, const Matrix& A) {
    if (A.NVecs() != NVecs()) throw PanicMsg("goChem/v3: Wrong shape");
    for (size_t i = 0; i < d_.size(); i++) d_[i] = v * A.d_[i];
}

double Matrix::Norm(double) const {
    double scale = 0.0, ssq = 1.0;  // Dlange('F'): scaled sum of squares
    for (double v : d_) {
        if (v == 0.0) continue;
        const double a = std::fabs(v);
        if (scale < a) { ssq = 1.0 + ssq * (scale / a) * (scale / a); scale = a; }
        else ssq += (a / scale) * (a / scale);
    }
    return scale * std::sqrt(ssq);
}

Matrix Matrix::TrRet() const {
    if (NVecs() != 3) throw PanicMsg("goChem/v3: Only 3x3 matrices are allowed for this function");
    Matrix t = Zeros(3);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t.d_[3 * (size_t)j + i] = d_[3 * (size_t)i + j];
    return t;
}

}  // namespace v3

// ------------------------------------------------------------------ chem: single-structure calls
namespace chem {

namespace {
std::mutex g_mu;
int g_dev = 0;
gcb_traj* g_one = nullptr;  // cached one-frame trajectory for single-structure calls
int g_one_natoms = 0;

// one-frame device trajectory holding `test`
Error stage_one(const v3::Matrix& test, gcb_traj** t) {
    const int n = test.NVecs();
    if (n <= 0) return Error("goChem: Ill-formed matrices", {"stage_one"}, GCB_ERR_SHAPE);
    if (!g_one || g_one_natoms != n) {
        if (g_one) gcb_traj_destroy(g_one);
        g_one = nullptr;
        int rc = gcb_traj_create(g_dev, n, 1, &g_one);
        if (rc) return deviceError(rc, "gcb_traj_create");
        g_one_natoms = n;
    }
    int rc = gcb_traj_upload(g_one, 0, test.data(), 1, GCB_F64_AOS, 1.0);
    if (rc) return deviceError(rc, "gcb_traj_upload");
    *t = g_one;
    return Error();
}

// The three forms of Go's variadic index argument (handy.go:178-201, geometric.go:261-283) as C-ABI lists.
struct Lists {
    const int* it = nullptr;
    const int* ir = nullptr;
    int n = 0;
    std::vector<int> iota;
};

Error dispatch(const v3::Matrix& test, const v3::Matrix& templa, const IndexLists& indexes, const char* fn, const char* shape_msg,
               Lists* L) {
    if (indexes.empty() || indexes[0].empty()) {
        if (test.NVecs() != templa.NVecs()) return Error(shape_msg, {}, GCB_ERR_SHAPE);
        return Error();
    }
    if (indexes.size() == 1) {
        const int len = (int)indexes[0].size();
        if (test.NVecs() > len) {
            // ctest = test[indexes[0]], ctempla = templa (whole): templa rows pair with the list positions
            if (templa.NVecs() != len) return Error(shape_msg, {}, GCB_ERR_SHAPE);
            L->iota.resize((size_t)len);
            for (int k = 0; k < len; k++) L->iota[(size_t)k] = k;
            L->it = indexes[0].data();
            L->ir = L->iota.data();
            L->n = len;
            return Error();
        }
        if (templa.NVecs() > len)
            return Error(std::string(fn) + ": the reference dereferences a nil matrix for a single index list that applies to templa "
                                           "(handy.go:186-188, geometric.go:269-271); give both lists",
                         {}, GCB_ERR_INDEXES);
        return Error(std::string(fn) + ": Indexes don't match molecules", {}, GCB_ERR_INDEXES);
    }
    if (indexes[0].size() != indexes[1].size()) return Error(shape_msg, {}, GCB_ERR_SHAPE);
    L->it = indexes[0].data();
    L->ir = indexes[1].data();
    L->n = (int)indexes[0].size();
    return Error();
}
}  // namespace

Error SetDevice(int dev) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_one) { gcb_traj_destroy(g_one); g_one = nullptr; g_one_natoms = 0; }
    g_dev = dev;
    return Error();
}
int Device() { return g_dev; }

Error newLastFrameError(const std::string& file, const std::string& who) {
    Error e("EOF", {who});
    e.msg = "EOF" + (file.empty() ? std::string() : " (" + file + ")");
    e.last_frame = true;
    e.critical = false;
    e.code = GCB_EOF;
    return e;
}

Error RotatorTranslatorToSuper(const v3::Matrix& test, const v3::Matrix& templa, v3::Matrix* transformed, v3::Matrix* rotation,
                               v3::Matrix* trans1, v3::Matrix* trans2) {
    if (test.NVecs() != templa.NVecs() || test.NVecs() == 0)
        return Error("goChem: Ill-formed matrices", {"RotatorTranslatorToSuper"}, GCB_ERR_SHAPE);  // geometric.go:130-132
    std::lock_guard<std::mutex> lk(g_mu);
    gcb_traj* t = nullptr;
    if (Error e = stage_one(test, &t)) return errDecorate(e, "RotatorTranslatorToSuper");
    double rm, rot[9], t1[3], t2[3];
    int rc = gcb_super_rmsd(t, 0, 1, templa.data(), templa.NVecs(), nullptr, nullptr, 0, transformed ? 1 : 0, &rm, rot, t1, t2);
    if (rc) return deviceError(rc, "RotatorTranslatorToSuper");
    if (transformed) {
        *transformed = v3::Matrix::Zeros(test.NVecs());
        rc = gcb_traj_download(t, 0, 1, transformed->data());
        if (rc) return deviceError(rc, "RotatorTranslatorToSuper");
    }
    if (rotation) { *rotation = v3::Matrix::Zeros(3); std::copy(rot, rot + 9, rotation->data()); }
    if (trans1) { *trans1 = v3::Matrix::Zeros(1); std::copy(t1, t1 + 3, trans1->data()); }
    if (trans2) { *trans2 = v3::Matrix::Zeros(1); std::copy(t2, t2 + 3, trans2->data()); }
    return Error();
}

Error Super(v3::Matrix& test, const v3::Matrix& templa, const IndexLists& indexes) {
    Lists L;
    if (Error e = dispatch(test, templa, indexes, "chem.Super", "chem.Super: Ill formed coordinates for Superposition", &L)) return e;
    std::lock_guard<std::mutex> lk(g_mu);
    gcb_traj* t = nullptr;
    if (Error e = stage_one(test, &t)) return errDecorate(e, "Super");
    int rc = gcb_super_rmsd(t, 0, 1, templa.data(), templa.NVecs(), L.it, L.ir, L.n, 1, nullptr, nullptr, nullptr, nullptr);
    if (rc) return errDecorate(deviceError(rc, "RotatorTranslatorToSuper"), "Super");
    rc = gcb_traj_download(t, 0, 1, test.data());  // test is modified in place (handy.go:207-213)
    if (rc) return deviceError(rc, "Super");
    return Error();
}

Error MemRMSD(const v3::Matrix& test, const v3::Matrix& templa, v3::Matrix& tmp, double* out, const IndexLists& indexes) {
    *out = -1.0;  // geometric.go:273, 282
    Lists L;
    if (Error e = dispatch(test, templa, indexes, "chem.memRMSD", "chem.memRMSD: Ill formed matrices for memRMSD calculation", &L)) return e;
    const int n = L.n ? L.n : test.NVecs();
    if (tmp.NVecs() != n) return Error("chem.memRMSD: Ill formed matrices for memRMSD calculation", {}, GCB_ERR_SHAPE);
    std::lock_guard<std::mutex> lk(g_mu);
    gcb_traj* t = nullptr;
    if (Error e = stage_one(test, &t)) return errDecorate(e, "MemRMSD");
    double r = -1.0;
    int rc = gcb_rmsd(t, 0, 1, templa.data(), templa.NVecs(), L.it, L.ir, L.n, &r);
    if (rc) return deviceError(rc, "MemRMSD");
    *out = r;
    return Error();
}

Error RMSD(const v3::Matrix& test, const v3::Matrix& templa, double* out, const IndexLists& indexes) {
    const int Lr = (indexes.empty() || indexes[0].empty()) ? test.NVecs() : (int)indexes[0].size();  // geometric.go:233-239
    v3::Matrix tmp = v3::Matrix::Zeros(Lr);
    return MemRMSD(test, templa, tmp, out, indexes);
}

Error MemPerAtomRMSD(const v3::Matrix& test, const v3::Matrix& templa, v3::Matrix* /*ctest*/, v3::Matrix* /*ctempla*/, v3::Matrix& tmp,
                     std::vector<double>* out, const IndexLists& indexes) {
    Lists L;
    if (Error e = dispatch(test, templa, indexes, "chem.memMSD", "chem.memMSD: Ill formed matrices for memMSD calculation", &L)) return e;
    const int n = L.n ? L.n : test.NVecs();
    if (tmp.NVecs() != n) return Error("chem.memMSD: Ill formed matrices for memMSD calculation", {}, GCB_ERR_SHAPE);
    std::lock_guard<std::mutex> lk(g_mu);
    gcb_traj* t = nullptr;
    if (Error e = stage_one(test, &t)) return errDecorate(e, "MemPerAtomRMSD");
    out->assign((size_t)n, 0.0);
    int rc = gcb_peratom_dist(t, 0, templa.data(), templa.NVecs(), L.it, L.ir, L.n, out->data());
    if (rc) return deviceError(rc, "MemPerAtomRMSD");
    return Error();
}

static Error moments_one(const v3::Matrix& A, const std::vector<double>& mass, v3::Matrix* center, v3::Matrix* moment, const char* who) {
    if (A.NVecs() == 0) return Error("goChem: nil matrix to get the center of mass", {who}, GCB_ERR_SHAPE);
    if (!mass.empty() && (int)mass.size() != A.NVecs())
        return Error("goChem: Error in gonum function: mat: dimension mismatch", {"gnMaybe", who}, GCB_ERR_SHAPE);  // ScaleByCol panics
    std::lock_guard<std::mutex> lk(g_mu);
    gcb_traj* t = nullptr;
    if (Error e = stage_one(A, &t)) return errDecorate(e, who);
    double c[3], m[9];
    int rc = gcb_moments(t, 0, 1, mass.empty() ? nullptr : mass.data(), nullptr, 0, c, m);
    if (rc) return deviceError(rc, who);
    if (center) { *center = v3::Matrix::Zeros(1); std::copy(c, c + 3, center->data()); }
    if (moment) { *moment = v3::Matrix::Zeros(3); std::copy(m, m + 9, moment->data()); }
    return Error();
}

Error CenterOfMass(const v3::Matrix& geometry, v3::Matrix* center, const std::vector<double>& mass) {
    return moments_one(geometry, mass, center, nullptr, "CenterOfMass");
}

Error MomentTensor(const v3::Matrix& A, v3::Matrix* moment, const std::vector<double>& mass) {
    return moments_one(A, mass, nullptr, moment, "MomentTensor");
}

// ------------------------------------------------------------------ in-memory trajectory
Error MemTraj::Next(v3::Matrix* V, std::vector<double>*) {
    if (current_ >= Coords.size()) return newLastFrameError("", "Next");  // chem.go:706-708
    current_++;
    if (!V) return Error();
    V->Copy(Coords[current_ - 1]);
    return Error();
}

Error MemTraj::NextConc(std::vector<v3::Matrix*>& frames) {
    bool used = false;
    for (auto& slot : frames) {
        if (slot == nullptr) {  // chem.go:733-737: discarded slots advance without an EOF check
            current_++;
            continue;
        }
        if (current_ >= Coords.size()) {
            (void)used;  // chem.go:738-745: EOF either way, with or without the partial channel list
            return newLastFrameError("", "NextConc");
        }
        used = true;
        slot = &Coords[current_];  // the stored frame itself is handed out (chem.go:749-751)
        current_++;
    }
    return Error();
}

}  // namespace chem

// ------------------------------------------------------------------ DCD
namespace dcd {

static uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

static Error dcdError(const std::string& msg, const std::string& file, std::vector<std::string> deco) {
    Error e(msg + (file.empty() ? "" : " (" + file + ")"), std::move(deco));
    return e;
}

Error DCDObj::readI32(int32_t* v) {
    uint32_t u;
    if (fread(&u, 4, 1, f_) != 1) return dcdError("EOF", filename_, {"binary.Read"});
    if (swap_) u = bswap32(u);
    memcpy(v, &u, 4);
    return Error();
}

Error DCDObj::New(const std::string& filename, std::unique_ptr<DCDObj>* out) {
    std::unique_ptr<DCDObj> d(new DCDObj());
    if (Error e = d->initRead(filename)) return errDecorate(e, "New");
    d->fields_.assign(3 * (size_t)d->natoms_, 0.0f);
    *out = std::move(d);
    return Error();
}

DCDObj::~DCDObj() { Close(); }

void DCDObj::Close() {
    if (f_) fclose(f_);
    f_ = nullptr;
    readable_ = false;
}

Error DCDObj::initRead(const std::string& name) {  // dcd.go:156-281
    filename_ = name;
    f_ = fopen(name.c_str(), "rb");
    if (!f_) return dcdError("open " + name + ": no such file or directory", name, {"os.Open", "initRead"});
    int32_t check = 0;
    if (Error e = readI32(&check)) return errDecorate(e, "initRead");
    if (check != 84) {
        swap_ = true;  // other endianness (dcd.go:172-174)
        uint32_t u;
        memcpy(&u, &check, 4);
        u = bswap32(u);
        memcpy(&check, &u, 4);
        if (check != 84) return dcdError("Wrong format: first record is not 84 bytes", name, {"initRead"});
    }
    char magic[4];
    if (fread(magic, 1, 4, f_) != 4) return dcdError("EOF", name, {"binary.Read", "initRead"});
    if (memcmp(magic, "CORD", 4) != 0) return dcdError("Wrong format: Wrong magic number", name, {"initRead"});
    unsigned char buf[80];
    if (fread(buf, 1, 80, f_) != 80) return dcdError("EOF", name, {"binary.Read", "initRead"});
    auto i32at = [&](int off) {
        uint32_t u;
        memcpy(&u, buf + off, 4);
        if (swap_) u = bswap32(u);
        int32_t v;
        memcpy(&v, &u, 4);
        return v;
    };
    if (i32at(76) != 0) {
        charmm_ = true;
        if (i32at(40) != 0) extrablock_ = true;
        if (i32at(40) == 1) fourdim_ = true;  // the reference reads offset 40 twice (dcd.go:196-201)
    } else {
        return dcdError("Wrong format: X-plor DCD not supported", name, {"initRead"});
    }
    fixed_ = i32at(32);
    if (Error e = readI32(&check)) return errDecorate(e, "initRead");
    if (check != 84) return dcdError("Wrong format", name, {"initRead"});
    int32_t input_int = 0, ntitle = 0;
    if (Error e = readI32(&input_int)) return errDecorate(e, "initRead");
    if (Error e = readI32(&ntitle)) return errDecorate(e, "initRead");
    if (ntitle < 0 || ntitle > 1000000) return dcdError("Wrong format", name, {"initRead"});
    std::vector<char> title(80 * (size_t)ntitle);
    if (!title.empty() && fread(title.data(), 1, title.size(), f_) != title.size()) return dcdError("EOF", name, {"binary.Read", "initRead"});
    if (Error e = readI32(&input_int)) return errDecorate(e, "initRead");
    if (Error e = readI32(&check)) return errDecorate(e, "initRead");
    if (check != 4) return dcdError("Wrong format", name, {"initRead"});
    if (Error e = readI32(&natoms_)) return errDecorate(e, "initRead");
    if (Error e = readI32(&check)) return errDecorate(e, "initRead");
    if (check != 4) return dcdError("Wrong format", name, {"initRead"});
    if (natoms_ <= 0) return dcdError("Wrong format", name, {"initRead"});
    if (fixed_ == 0) {
        readable_ = true;
        return Error();
    }
    return dcdError("Fixed atoms not supported", name, {"initRead"});
}

Error DCDObj::readFloat32Block(int32_t blocksize, float* block, int n) {  // dcd.go:436-452
    if (blocksize != n * 4) {
        // the reference would read len(block) floats and then fail the trailer check
        return dcdError("Wrong format", filename_, {"readFloat32Block"});
    }
    if (fread(block, 4, (size_t)n, f_) != (size_t)n) return dcdError("unexpected EOF", filename_, {"binary.Read", "readFloat32Block"});
    if (swap_) {
        for (int i = 0; i < n; i++) {
            uint32_t u;
            memcpy(&u, block + i, 4);
            u = bswap32(u);
            memcpy(block + i, &u, 4);
        }
    }
    int32_t check;
    if (Error e = readI32(&check)) return errDecorate(e, "readFloat32Block");
    if (check != blocksize) return dcdError("Wrong format", filename_, {"readFloat32Block"});
    return Error();
}

Error DCDObj::nextRaw(float* x, float* y, float* z) {  // dcd.go:349-434
    if (!readable_) return dcdError("Traj object uninitialized to read", filename_, {"nextRaw"});
    if (readLast_) {
        Close();
        return chem::newLastFrameError(filename_, "nextRaw");
    }
    int32_t blocksize = 0;
    if (extrablock_) {
        if (readI32(&blocksize)) { Close(); return chem::newLastFrameError(filename_, "nextRaw"); }
        if (blocksize < natoms_ * 3) {
            // unit cell: 6 doubles in CHARMM files; read and skip whatever the record holds
            std::vector<char> cell((size_t)(blocksize > 0 ? blocksize : 0));
            if (!cell.empty() && fread(cell.data(), 1, cell.size(), f_) != cell.size()) return dcdError("unexpected EOF", filename_, {"nextRaw"});
            int32_t check;
            if (Error e = readI32(&check)) return errDecorate(e, "nextRaw");
            if (check != blocksize) return dcdError("Wrong format", filename_, {"nextRaw"});
            blocksize = 0;
        }
    }
    if (blocksize == 0) {
        if (readI32(&blocksize)) { Close(); return chem::newLastFrameError(filename_, "nextRaw"); }
    }
    if (Error e = readFloat32Block(blocksize, x, natoms_)) return errDecorate(e, "nextRaw");
    if (Error e = readI32(&blocksize)) return errDecorate(e, "nextRaw");
    if (Error e = readFloat32Block(blocksize, y, natoms_)) return errDecorate(e, "nextRaw");
    if (Error e = readI32(&blocksize)) return errDecorate(e, "nextRaw");
    if (Error e = readFloat32Block(blocksize, z, natoms_)) return errDecorate(e, "nextRaw");
    if (charmm_ && fourdim_) {
        if (readI32(&blocksize)) {
            readLast_ = true;
        } else {
            std::vector<char> skip((size_t)(blocksize > 0 ? blocksize : 0));
            if (!skip.empty() && fread(skip.data(), 1, skip.size(), f_) != skip.size()) return dcdError("unexpected EOF", filename_, {"nextRaw"});
            int32_t check;
            if (Error e = readI32(&check)) return errDecorate(e, "nextRaw");
            if (check != blocksize) return dcdError("Security check failed", filename_, {"readByteBlock"});
        }
    }
    return Error();
}

Error DCDObj::Next(v3::Matrix* keep, std::vector<double>* box) {
    float* b = fields_.data();
    if (Error e = nextRaw(b, b + natoms_, b + 2 * (size_t)natoms_)) return errDecorate(e, "Next");
    if (box) box->assign(box_, box_ + 6);
    if (!keep) return Error();
    if (keep->NVecs() < natoms_) throw v3::PanicMsg("Not enough space in matrix");
    for (int i = 0; i < natoms_; i++) {  // widened per element (dcd.go:302-307)
        keep->Set(i, 0, (double)b[i]);
        keep->Set(i, 1, (double)b[natoms_ + i]);
        keep->Set(i, 2, (double)b[2 * (size_t)natoms_ + i]);
    }
    return Error();
}

Error DCDObj::NextConc(std::vector<v3::Matrix*>& frames) {  // dcd.go:513-546
    if (!Readable()) return dcdError("Traj object uninitialized to read", filename_, {"NextConc"});
    for (auto* slot : frames) {
        float* b = fields_.data();
        if (Error e = nextRaw(b, b + natoms_, b + 2 * (size_t)natoms_)) return errDecorate(e, "NextConc");
        if (!slot) continue;
        for (int i = 0; i < natoms_; i++) {
            slot->Set(i, 0, (double)b[i]);
            slot->Set(i, 1, (double)b[natoms_ + i]);
            slot->Set(i, 2, (double)b[2 * (size_t)natoms_ + i]);
        }
    }
    return Error();
}

// Works out how one frame sits in the file (dcd.go:349-434 read in one go): [4 | cell | 4] (CHARMM extra block, when present)
// [4 | x | 4] [4 | y | 4] [4 | z | 4] and [4 | fourth dimension | 4] when flagged.  The file position is left unchanged.
Error DCDObj::probeRecord() {
    if (rec_bytes_ > 0) return Error();
    const long pos = ftell(f_);
    if (pos < 0) return dcdError("ftell failed", filename_, {"NextRecords"});
    const long plane = 4L * natoms_;
    long off = 0;
    int32_t b = 0;
    auto back = [&]() { fseek(f_, pos, SEEK_SET); };
    if (fread(&b, 4, 1, f_) != 1) { back(); return chem::newLastFrameError(filename_, "NextRecords"); }
    if (extrablock_ && b < natoms_ * 3) {   // the unit-cell record comes first (dcd.go:361-376)
        off = 4 + (long)b + 4;
        if (fseek(f_, pos + off, SEEK_SET) != 0 || fread(&b, 4, 1, f_) != 1) { back(); return chem::newLastFrameError(filename_, "NextRecords"); }
    }
    if ((long)b != plane) { back(); return dcdError("Wrong format", filename_, {"NextRecords"}); }
    rec_off_[0] = off + 4;
    rec_off_[1] = rec_off_[0] + plane + 8;
    rec_off_[2] = rec_off_[1] + plane + 8;
    long total = rec_off_[2] + plane + 4;
    if (charmm_ && fourdim_) {
        int32_t b4 = 0;
        if (fseek(f_, pos + total, SEEK_SET) == 0 && fread(&b4, 4, 1, f_) == 1 && b4 >= 0) total += 4 + (long)b4 + 4;
    }
    back();
    rec_bytes_ = total;
    return Error();
}

long DCDObj::RecordsAvailable() {
    if (!readable_ || probeRecord()) return -1;
    const long pos = ftell(f_);
    struct stat st;
    if (pos < 0 || fstat(fileno(f_), &st) != 0) return -1;
    // The bulk feed assumes every frame is laid out like the first.  The reference's nextRaw (dcd.go:359-378, 415-431) also reads
    // files in which some frames lack the unit-cell block, and a last frame without the fourth-dimension block; those go
    // through the frame-by-frame reader: the feed is offered only when the rest of the file is a whole number of records and
    // no fourth-dimension block is flagged.
    if (charmm_ && fourdim_) return -1;
    if (((long)st.st_size - pos) % rec_bytes_ != 0) return -1;
    const long n = ((long)st.st_size - pos) / rec_bytes_;
    return n < 0 ? 0 : n;
}

Error DCDObj::NextRecords(char* dst, long max_frames, long* got, long* record_bytes, long* off, double* scale) {
    *got = 0;
    *scale = 1.0;
    if (!readable_) return dcdError("Traj object uninitialized to read", filename_, {"NextRecords"});
    if (Error e = probeRecord()) { if (e.last_frame) Close(); return e; }
    *record_bytes = rec_bytes_;
    off[0] = rec_off_[0]; off[1] = rec_off_[1]; off[2] = rec_off_[2];
    if (max_frames <= 0) return Error();
    long n = 0;
    if (dst) {
        // positional reads of the window in a few slices at once: one thread copies out of the page cache at 5-7 GB/s, the
        // staging path takes ten times that
        const long pos = ftell(f_);
        struct stat st;
        if (pos < 0 || fstat(fileno(f_), &st) != 0) return dcdError("cannot size the file", filename_, {"NextRecords"});
        const long avail = ((long)st.st_size - pos) / rec_bytes_;   // a trailing partial frame is not a frame
        n = avail < max_frames ? (avail < 0 ? 0 : avail) : max_frames;
        const size_t bytes = (size_t)n * (size_t)rec_bytes_;
        const int fd = fileno(f_);
        unsigned nthreads = bytes >= (16u << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) : 1u;
        std::vector<std::thread> th;
        std::vector<int> ok(nthreads, 1);
        const size_t slice = (bytes / nthreads + 4095) / 4096 * 4096;
        for (unsigned k = 0; k < nthreads; k++) {
            const size_t b0 = std::min(bytes, (size_t)k * slice), b1 = std::min(bytes, b0 + slice);
            auto job = [fd, dst, pos, b0, b1, &ok, k]() {
                size_t done = b0;
                while (done < b1) {
                    const ssize_t r = pread(fd, dst + done, b1 - done, (off_t)(pos + (long)done));
                    if (r <= 0) { ok[k] = 0; return; }
                    done += (size_t)r;
                }
            };
            if (k + 1 < nthreads) th.emplace_back(job); else job();
        }
        for (auto& x : th) x.join();
        for (unsigned k = 0; k < nthreads; k++) if (!ok[k]) return dcdError("read failed", filename_, {"NextRecords"});
        if (fseek(f_, pos + (long)bytes, SEEK_SET) != 0) return dcdError("seek failed", filename_, {"NextRecords"});
        const int32_t plane = 4 * natoms_;
        for (long f = 0; f < n; f++) {   // the markers around every plane (dcd.go:436-452 checks them per block)
            const char* r = dst + (size_t)f * (size_t)rec_bytes_;
            for (int c = 0; c < 3; c++) {
                int32_t a, b;
                memcpy(&a, r + rec_off_[c] - 4, 4);
                memcpy(&b, r + rec_off_[c] + plane, 4);
                if (a != plane || b != plane) return dcdError("Wrong format", filename_, {"NextRecords"});
            }
        }
    } else {
        const long pos = ftell(f_);
        struct stat st;
        if (pos < 0 || fstat(fileno(f_), &st) != 0) return dcdError("cannot size the file", filename_, {"NextRecords"});
        const long avail = ((long)st.st_size - pos) / rec_bytes_;
        n = avail < max_frames ? (avail < 0 ? 0 : avail) : max_frames;
        if (fseek(f_, pos + n * rec_bytes_, SEEK_SET) != 0) return dcdError("seek failed", filename_, {"NextRecords"});
    }
    *got = n;
    if (n == 0) { Close(); return chem::newLastFrameError(filename_, "NextRecords"); }
    return Error();
}

Error DCDObj::NextRaw(float* dst, int* layout, double* scale) {
    *layout = GCB_F32_SOA;
    *scale = 1.0;
    if (dst) return nextRaw(dst, dst + natoms_, dst + 2 * (size_t)natoms_);
    float* b = fields_.data();
    return nextRaw(b, b + natoms_, b + 2 * (size_t)natoms_);
}

Error DCDWObj::NewWriter(const std::string& filename, int natoms, std::unique_ptr<DCDWObj>* out) {  // dcd_write.go:102-205
    std::unique_ptr<DCDWObj> w(new DCDWObj());
    w->filename_ = filename;
    w->natoms_ = natoms;
    w->f_ = fopen(filename.c_str(), "wb+");
    if (!w->f_) return dcdError("open " + filename + ": cannot create", filename, {"os.Open", "initWrite", "NewWriter"});
    auto wi = [&](int32_t v) { fwrite(&v, 4, 1, w->f_); };
    wi(84);
    fwrite("CORD", 1, 4, w->f_);
    wi(0); wi(0); wi(1);
    for (int i = 0; i < 6; i++) wi(0);
    float one = 1.0f;
    fwrite(&one, 4, 1, w->f_);
    wi(0);
    for (int i = 0; i < 8; i++) wi(0);
    wi(24);
    wi(84);
    wi(244);
    wi(2);
    char title[160];
    memset(title, 'l', sizeof title);
    title[159] = 0;
    fwrite(title, 1, sizeof title, w->f_);
    wi(244);
    wi(4);
    if (natoms == 0) return dcdError("Trajectory not initialized correctly, the number of atoms is set to zero!", filename, {"initWrite"});
    wi(natoms);
    wi(4);
    w->writable_ = true;
    w->fields_.assign(3 * (size_t)natoms, 0.0f);
    *out = std::move(w);
    return Error();
}

DCDWObj::~DCDWObj() { Close(); }

void DCDWObj::Close() {
    if (f_) fclose(f_);
    f_ = nullptr;
    writable_ = false;
}

Error DCDWObj::WNextRaw(const double* aos, int natoms) {  // dcd_write.go:212-330
    if (!writable_) return dcdError("Traj object uninitialized to read", filename_, {"WNext"});
    if (!aos) return dcdError("got nil coordinates", filename_, {"WNext"});
    if (natoms != natoms_) return dcdError("Coordinates don't match the trajectory size", filename_, {"WNext"});
    float* b = fields_.data();
    for (int k = 0; k < natoms_; k++) {
        b[k] = (float)aos[3 * (size_t)k];
        b[natoms_ + k] = (float)aos[3 * (size_t)k + 1];
        b[2 * (size_t)natoms_ + k] = (float)aos[3 * (size_t)k + 2];
    }
    const int32_t bs = natoms_ * 4;
    for (int c = 0; c < 3; c++) {
        fwrite(&bs, 4, 1, f_);
        fwrite(b + (size_t)c * natoms_, 4, (size_t)natoms_, f_);
        fwrite(&bs, 4, 1, f_);
    }
    frames_++;
    // updateFrames: rewrite the frame counter at the top of the file (dcd_write.go:300-330)
    const long here = ftell(f_);
    fseek(f_, 0, SEEK_SET);
    int32_t v = 84;
    fwrite(&v, 4, 1, f_);
    fwrite("CORD", 1, 4, f_);
    fwrite(&frames_, 4, 1, f_);
    fseek(f_, here, SEEK_SET);
    return Error();
}

Error DCDWObj::WNext(const v3::Matrix& towrite) { return WNextRaw(towrite.data(), towrite.NVecs()); }

}  // namespace dcd

// ------------------------------------------------------------------ align
namespace align {

Options* DefaultOptions() {
    Options* r = new Options();
    unsigned hc = std::thread::hardware_concurrency();
    r->cpus_ = hc ? (int)hc : 1;
    r->atomNames_ = {"CA"};
    r->lessThanRMSD_ = 1.0;
    r->minimumN_ = 10;
    return r;
}

Options* DefaultCGOptions() {
    Options* r = DefaultOptions();
    r->atomNames_ = {"BB"};
    return r;
}

std::string MolIDandChain::String() const {
    char b[64];
    snprintf(b, sizeof b, "molid:%04d-chain:%s", molid, chain.c_str());
    return b;
}

static std::string fmtFloat(double v) {  // Go's %v for float64: shortest representation that round-trips
    char b[40];
    for (int prec = 1; prec <= 17; prec++) {
        snprintf(b, sizeof b, "%.*g", prec, v);
        if (strtod(b, nullptr) == v) break;
    }
    return b;
}

std::string LOVOReturn::String() const {  // lovo.go:79-86
    // plain string appends: no iostreams in this library (their locale state is per libstdc++ copy)
    std::string s = "N: " + std::to_string(N) + ",  Frames read: " + std::to_string(FramesRead) +
                    ", Iterations needed: " + std::to_string(Iterations) + ", RMSD: [";
    for (size_t i = 0; i < RMSD.size(); i++) { if (i) s += " "; s += fmtFloat(RMSD[i]); }
    s += "], Natoms: [";
    for (size_t i = 0; i < Natoms.size(); i++) { if (i) s += " "; s += std::to_string(Natoms[i]); }
    s += "], Nmols:";
    for (const auto& m : Nmols) s += " " + m.String();
    return s;
}

std::string LOVOReturn::PyMOLSel() const {  // lovo.go:102-116
    std::string sel = "select rigid,";
    char b[96];
    for (size_t i = 0; i < Nmols.size(); i++) {
        if (Nmols[i].chain.empty()) snprintf(b, sizeof b, " (resi %d) ", Nmols[i].molid);
        else snprintf(b, sizeof b, " (resi %d and chain %s) ", Nmols[i].molid, Nmols[i].chain.c_str());
        sel += b;
        if (i + 1 < Nmols.size()) sel += "or";
    }
    return sel;
}

std::string LOVOReturn::GMX(const std::string& chain) const {  // lovo.go:128-150
    bool first = true;
    std::string ret;
    char b[64];
    for (const auto& v : Nmols) {
        if (!chain.empty() && chain != v.chain) continue;
        if (first) {
            snprintf(b, sizeof b, "r %d ", v.molid);
            ret = b;
            first = false;
            continue;
        }
        snprintf(b, sizeof b, " | r %d ", v.molid);
        ret += b;
    }
    if (!chain.empty()) ret += " & chain " + chain;
    return ret;
}

Error mobileLimit(double limit, int tolerance, const std::vector<double>& rmsd, int* nout, double* newlimit) {
    int n = 0;
    long guard = 0;
    while (n < tolerance) {
        n = 0;
        for (; n < (int)rmsd.size(); n++) {
            if (rmsd[(size_t)n] >= limit) {
                n--;  // off by one on purpose: lovo.go:180-183
                break;
            }
        }
        limit *= 1.1;
        if (++guard > 100000 || !std::isfinite(limit))
            return Error("align.mobileLimit: fewer candidate atoms than MinimumN; the reference would loop forever here (lovo.go:177-186)");
    }
    *nout = n;
    *newlimit = limit;
    return Error();
}

// The permutation sort.Stable leaves (lovo.go:429-432, Less = rMSDs[i] < rMSDs[j]): ascending values, ties in their original
// order.  Non-negative finite doubles order like their bit patterns, so the common case is an LSD radix sort on the 64-bit keys
// (stable by construction, ~10x faster than a comparison sort with indirection at 5000 candidates -- the sort was a third of a
// LOVO iteration's host time); anything else (negative values, NaN) takes std::stable_sort with the same comparison.
void stableOrderByValue(const std::vector<double>& v, std::vector<int>* order) {
    const size_t n = v.size();
    order->resize(n);
    bool radix_ok = n >= 64;
    for (size_t k = 0; k < n && radix_ok; k++) radix_ok = v[k] >= 0.0 && v[k] <= 1.79e308;   // false for NaN as well
    if (!radix_ok) {
        for (size_t k = 0; k < n; k++) (*order)[k] = (int)k;
        std::stable_sort(order->begin(), order->end(), [&](int a, int b) { return v[(size_t)a] < v[(size_t)b]; });
        return;
    }
    // Three 11-bit passes over the TOP 33 bits (sign, exponent, 20 bits of mantissa) put almost everything in place; one
    // insertion pass on the full keys (strict comparisons, so ties keep their order) finishes the few neighbours that agree
    // in those bits.
    struct Item { uint64_t key; int idx; };
    static thread_local std::vector<Item> a, b;   // scratch kept between calls: two fresh 80 KB vectors cost as much as the sort
    if (a.size() < n) { a.resize(n); b.resize(n); }
    for (size_t k = 0; k < n; k++) {
        const double d = v[k] + 0.0;   // -0.0 -> +0.0
        memcpy(&a[k].key, &d, 8);
        a[k].idx = (int)k;
    }
    constexpr int kBits = 11, kBuckets = 1 << kBits;
    for (int shift = 31; shift < 64; shift += kBits) {
        uint32_t count[kBuckets + 1] = {0};
        for (size_t k = 0; k < n; k++) count[((a[k].key >> shift) & (kBuckets - 1)) + 1]++;
        for (int q = 0; q < kBuckets; q++) count[q + 1] += count[q];
        for (size_t k = 0; k < n; k++) b[count[(a[k].key >> shift) & (kBuckets - 1)]++] = a[k];
        a.swap(b);
    }
    for (size_t k = 1; k < n; k++) {
        const Item it = a[k];
        size_t j = k;
        while (j > 0 && a[j - 1].key > it.key) { a[j] = a[j - 1]; j--; }
        a[j] = it;
    }
    for (size_t k = 0; k < n; k++) (*order)[k] = a[k].idx;
}

bool approxEq(double i, double j, double epsilon) { return std::fabs(i) - std::fabs(j) <= epsilon; }

// The reference answers "is v in the list" with a linear search (lovo.go:474-484), O(n^2) per LOVO iteration -- 75 M
// comparisons at 5000 candidates, three times the device time of the pass.  Same answers from a bitmap (Membership).
namespace {
// "is v in the list": a bitmap over [min, max] of the list (atom indexes), a hash set when the span is unreasonable
class Membership {
  public:
    explicit Membership(const std::vector<int>& c) {
        if (c.empty()) return;
        lo_ = *std::min_element(c.begin(), c.end());
        const long span = (long)*std::max_element(c.begin(), c.end()) - lo_ + 1;
        if (span <= (1L << 26)) {
            bits_.assign((size_t)span, 0);
            for (int v : c) bits_[(size_t)(v - lo_)] = 1;
        } else {
            set_.insert(c.begin(), c.end());
        }
    }
    bool has(int v) const {
        if (!bits_.empty()) return v >= lo_ && (size_t)(v - lo_) < bits_.size() && bits_[(size_t)(v - lo_)];
        return set_.count(v) != 0;
    }
  private:
    int lo_ = 0;
    std::vector<char> bits_;
    std::unordered_set<int> set_;
};
}  // namespace

bool sameElementsInt(const std::vector<int>& a, const std::vector<int>& b) {
    if (a.size() != b.size()) return false;
    const Membership inb(b);
    for (int v : a)
        if (!inb.has(v)) return false;
    return true;
}

int disagreementInt(const std::vector<int>& a, const std::vector<int>& b) {
    if (a.size() != b.size()) return -1;
    const Membership inb(b);
    int d = 0;
    for (int v : a)
        if (!inb.has(v)) d++;
    return d;
}

std::vector<int> res2atoms(const chem::Atomer& mol, const std::vector<std::string>& names, const std::vector<std::string>& chains) {
    std::vector<int> ret;
    for (int i = 0; i < mol.Len(); i++) {
        const chem::Atom& a = mol.Atom(i);
        if (std::find(names.begin(), names.end(), a.Name) == names.end()) continue;
        if (chains.empty() || std::find(chains.begin(), chains.end(), a.Chain) != chains.end()) ret.push_back(i);
    }
    return ret;
}

std::vector<MolIDandChain> iDs2MolIDandChains(const chem::Atomer& mol, const std::vector<int>& indexes) {
    // first occurrences in list order, as the reference's search over the growing result (lovo.go:540-551)
    std::vector<MolIDandChain> ret;
    if (indexes.empty()) return ret;
    // chains are few: a short list searched linearly (the previous atom's chain first); "seen" is a bitmap over
    // (chain, MolID - min MolID) when that is small, a hash set otherwise
    std::vector<const std::string*> chains;
    int lo = mol.Atom(indexes[0]).MolID, hi = lo;
    for (int v : indexes) { const int m = mol.Atom(v).MolID; lo = std::min(lo, m); hi = std::max(hi, m); }
    const uint64_t span = (uint64_t)((long)hi - lo + 1);
    std::vector<char> bits;
    std::unordered_set<uint64_t> seen;
    ret.reserve(indexes.size());
    int last_chain = -1;
    for (int v : indexes) {
        const chem::Atom& at = mol.Atom(v);
        int c = -1;
        if (last_chain >= 0 && *chains[(size_t)last_chain] == at.Chain) c = last_chain;
        for (size_t q = 0; c < 0 && q < chains.size(); q++) if (*chains[q] == at.Chain) c = (int)q;
        if (c < 0) { chains.push_back(&at.Chain); c = (int)chains.size() - 1; }
        last_chain = c;
        const uint64_t key = (uint64_t)c * span + (uint64_t)((long)at.MolID - lo);
        bool fresh;
        if ((uint64_t)(c + 1) * span <= (1u << 24)) {
            if (bits.size() < (size_t)((uint64_t)(c + 1) * span)) bits.resize((size_t)((uint64_t)(c + 1) * span), 0);
            fresh = !bits[(size_t)key];
            bits[(size_t)key] = 1;
        } else {
            fresh = seen.insert(((uint64_t)c << 40) ^ (uint64_t)((long)at.MolID - lo)).second;
        }
        if (fresh) ret.push_back(MolIDandChain{at.MolID, at.Chain});
    }
    return ret;
}

std::vector<long> selectedFrames(long F, int cpus, int begin_opt, int skip_opt) {
    std::vector<long> sel;
    if (cpus <= 0) return sel;
    const long begin = begin_opt / cpus, skip = skip_opt / cpus;  // lovo.go:306-307
    long cursor = 0;
    for (long read = 0;; read++) {
        if (read < begin || read % (skip + 1) != 0) {  // one chunk read and discarded, then NO continue (:316-322)
            if (cursor + cpus > F) break;
            cursor += cpus;
        }
        if (cursor + cpus > F) break;  // a chunk that cannot be filled is dropped (:323-326)
        for (int s = 0; s < cpus; s++) sel.push_back(cursor + s);
        cursor += cpus;
    }
    return sel;
}

void shardRange(long nsel, int rank, int nranks, long* first, long* count) {
    if (nranks < 1) nranks = 1;
    const long base = nsel / nranks, rem = nsel % nranks;
    *first = rank * base + std::min<long>(rank, rem);
    *count = base + (rank < rem ? 1 : 0);
}

LovoState::LovoState(std::vector<int> fullindexes, double lessThanRMSD, int minimumN)
    : full_(std::move(fullindexes)), lessThan_(lessThanRMSD), minimumN_(minimumN) {
    indexesold_ = full_;  // lovo.go:212-213
    n_ = (int)full_.size();
    fit_ = indexesold_;
}

bool LovoState::Step(const std::vector<double>& rmsd, Error* err) {
    itercount_++;
    // rmsdFilter (lovo.go:519-528) walks rmsd in atom order keeping the candidates; newRMSD pairs them with
    // fullindexes position by position (:367-384)
    std::vector<double> filtered;
    filtered.reserve(full_.size());
    const Membership infull(full_);
    for (int i = 0; i < (int)rmsd.size(); i++)
        if (infull.has(i)) filtered.push_back(rmsd[(size_t)i]);
    if (filtered.size() != full_.size()) {
        *err = Error("Inconsistent data, if given, the number of initial RMSD values must match the number of residues");  // :373-375 panic
        return true;
    }
    std::vector<int> order;
    stableOrderByValue(filtered, &order);   // sort.Stable by rMSDs[i] < rMSDs[j], :418-432
    ids_.resize(order.size());
    vals_.resize(order.size());
    for (size_t k = 0; k < order.size(); k++) { ids_[k] = full_[(size_t)order[k]]; vals_[k] = filtered[(size_t)order[k]]; }
    double newlim = 0.0;
    if (lessThan_ > 0) {  // :238-240
        if (Error e = mobileLimit(lessThan_, minimumN_, vals_, &n_, &newlim)) { *err = e; return true; }
    }
    if (n_ < 0 || n_ > (int)ids_.size()) { *err = Error("align.LOVO: slice bounds out of range (n outside the candidate list)"); return true; }
    std::vector<int> a(ids_.begin(), ids_.begin() + n_), b(indexesold_.begin(), indexesold_.begin() + n_);
    if (sameElementsInt(a, b) && (newlim == lessThan_ || approxEq(newlim, prevlim_, 0.01))) {  // :246-248
        fit_.assign(indexesold_.begin(), indexesold_.begin() + n_);
        return true;
    }
    prevlim_ = newlim;
    disagreement_ = disagreementInt(a, b);
    indexesold_ = ids_;
    fit_.assign(indexesold_.begin(), indexesold_.begin() + n_);
    return false;
}

void LovoState::Restore(std::vector<int> indexesold, std::vector<int> ids, std::vector<double> vals, int n, int iterations, int disagreement) {
    indexesold_ = std::move(indexesold);
    ids_ = std::move(ids);
    vals_ = std::move(vals);
    n_ = n;
    itercount_ = iterations;
    disagreement_ = disagreement;
    fit_.assign(indexesold_.begin(), indexesold_.begin() + n_);
}

void LovoState::Finish(const chem::Atomer& mol, int framesRead, LOVOReturn* out) const {
    // RMSD.SortBy("molid") (:254): stable by candidate atom index
    // the ids are the candidate atoms, each once: when the candidate list is ascending (res2atoms walks the atoms in order) the
    // sorted values are the values in candidate order, scattered through a rank table; otherwise a stable sort on the ids
    out->RMSD.resize(ids_.size());
    if (std::is_sorted(full_.begin(), full_.end()) && std::adjacent_find(full_.begin(), full_.end()) == full_.end() && !full_.empty()) {
        const int lo = full_.front();
        std::vector<int> rank((size_t)(full_.back() - lo + 1), -1);
        for (size_t k = 0; k < full_.size(); k++) rank[(size_t)(full_[k] - lo)] = (int)k;
        for (size_t k = 0; k < ids_.size(); k++) out->RMSD[(size_t)rank[(size_t)(ids_[k] - lo)]] = vals_[k];
    } else {
        std::vector<std::pair<int, double>> byid(ids_.size());
        for (size_t k = 0; k < byid.size(); k++) byid[k] = {ids_[k], vals_[k]};
        std::stable_sort(byid.begin(), byid.end(), [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.first < b.first; });
        for (size_t k = 0; k < byid.size(); k++) out->RMSD[k] = byid[k].second;
    }
    out->N = n_;
    out->Natoms.assign(indexesold_.begin(), indexesold_.begin() + n_);  // previous order, new n (:255)
    out->Nmols = iDs2MolIDandChains(mol, out->Natoms);
    out->FramesRead = framesRead;
    out->Iterations = itercount_;
}

// ---- device-resident trajectory ----
DeviceTrajectory::~DeviceTrajectory() {
    if (t_ && owned_) gcb_traj_destroy(t_);
}

std::unique_ptr<DeviceTrajectory> DeviceTrajectory::Adopt(gcb_traj* t, int natoms, long nlocal, long nselected_total) {
    std::unique_ptr<DeviceTrajectory> d(new DeviceTrajectory());
    d->t_ = t;
    d->owned_ = false;
    d->natoms_ = natoms;
    d->nlocal_ = nlocal;
    d->nselected_ = nselected_total;
    return d;
}

Error DeviceTrajectory::Load(chem::ConcTraj& traj, int natoms, const Options& o, std::unique_ptr<DeviceTrajectory>* out) {
    if (!traj.Readable()) return Error("Traj object uninitialized to read", {"RMSDTraj"});
    if (o.cpus_ <= 0) return Error("align.RMSDTraj: cpus must be positive");
    if (traj.Len() != natoms) return Error("align.RMSDTraj: trajectory and molecule differ in the number of atoms");
    const int chunksize = o.cpus_;
    const long begin = o.begin_ / chunksize, skip = o.skip_ / chunksize;
    // A source that hands out raw records and knows how many are left (a DCD file) goes straight from the file to the device:
    // this rank's block of the selected frames is read window by window into pinned memory and uploaded as it sits in the file.
    if (traj.HasRecordFeed() && begin == 0 && skip == 0) {
        long rb = 0, off[3] = {0, 0, 0}, got = 0;
        double sc = 1.0;
        const long avail = traj.RecordsAvailable();
        Error pe = traj.NextRecords(nullptr, 0, &got, &rb, off, &sc);
        if (avail >= 0 && !pe && rb > 0 && (size_t)chunksize * (size_t)rb <= (64u << 20)) {
            std::unique_ptr<DeviceTrajectory> d(new DeviceTrajectory());
            d->natoms_ = natoms;
            d->nselected_ = avail / chunksize * chunksize;   // a chunk the file cannot fill is dropped (:323-326)
            long first = 0, count = 0;
            shardRange(d->nselected_, o.rank_, o.nranks_, &first, &count);
            d->nlocal_ = count;
            int rc = gcb_traj_create(o.device_, natoms, count > 0 ? count : 1, &d->t_);
            if (rc) return deviceError(rc, "gcb_traj_create");
            const long per = std::max<long>(1, (long)((64u << 20) / (size_t)rb));
            void* pin[2] = {nullptr, nullptr};
            for (int s = 0; s < 2; s++)
                if ((rc = gcb_pinned_alloc((size_t)per * (size_t)rb, &pin[s]))) { gcb_pinned_free(pin[0]); return deviceError(rc, "gcb_pinned_alloc"); }
            Error up;
            if (first > 0) {
                Error e = traj.NextRecords(nullptr, first, &got, &rb, off, &sc);
                if (e || got != first) up = e ? e : Error("align.RMSDTraj: the trajectory ended before this rank's block");
            }
            int slot = 0;
            for (long done = 0; done < count && !up; slot ^= 1) {
                const long want = std::min(per, count - done);
                if ((rc = gcb_traj_wait(d->t_, slot))) { up = deviceError(rc, "gcb_traj_wait"); break; }
                Error e = traj.NextRecords(static_cast<char*>(pin[slot]), want, &got, &rb, off, &sc);
                if (e || got != want) { up = e ? e : Error("align.RMSDTraj: short read"); break; }
                rc = gcb_traj_upload_records_async(d->t_, done, pin[slot], got, rb, off[0], off[1], off[2], sc, slot);
                if (rc) up = deviceError(rc, "gcb_traj_upload_records_async");
                done += got;
            }
            gcb_traj_wait(d->t_, 0);
            gcb_traj_wait(d->t_, 1);
            gcb_pinned_free(pin[0]);
            gcb_pinned_free(pin[1]);
            if (up) return up;
            gcb_traj_set_nframes(d->t_, count);
            *out = std::move(d);
            return Error();
        }
    }
    const bool raw = traj.HasRawFeed();
    const size_t fbytes = (raw ? sizeof(float) : sizeof(double)) * 3 * (size_t)natoms;
    // Pass 1 over the source: keep the selected frames in host memory (the reader tells us nothing reliable
    // about its length; the reference discovers EOF the same way).
    std::vector<char> host;
    std::vector<v3::Matrix> chunk((size_t)chunksize);
    for (auto& m : chunk) m = v3::Matrix::Zeros(natoms);
    std::vector<char> scratch(fbytes);
    long nsel = 0;
    int layout = GCB_F64_AOS;
    double scale = 1.0;
    Error err;
    auto read_chunk = [&](bool keep) -> Error {
        if (raw) {
            const size_t base = host.size();
            if (keep) host.resize(base + fbytes * (size_t)chunksize);
            for (int s = 0; s < chunksize; s++) {
                float* dst = keep ? reinterpret_cast<float*>(host.data() + base + fbytes * (size_t)s) : reinterpret_cast<float*>(scratch.data());
                if (Error e = traj.NextRaw(dst, &layout, &scale)) {
                    if (keep) host.resize(base);  // partial chunk dropped
                    return e;
                }
            }
            return Error();
        }
        std::vector<v3::Matrix*> slots((size_t)chunksize, nullptr);
        if (keep) for (int s = 0; s < chunksize; s++) slots[(size_t)s] = &chunk[(size_t)s];
        if (Error e = traj.NextConc(slots)) return e;
        if (keep) {
            const size_t base = host.size();
            host.resize(base + fbytes * (size_t)chunksize);
            for (int s = 0; s < chunksize; s++) memcpy(host.data() + base + fbytes * (size_t)s, slots[(size_t)s]->data(), fbytes);
        }
        return Error();
    };
    for (long read = 0;; read++) {
        if (read < begin || read % (skip + 1) != 0) {
            err = read_chunk(false);
            if (err) break;
        }
        err = read_chunk(true);
        if (err) break;
        nsel += chunksize;
    }
    if (!err.last_frame)  // lovo.go:343-347
        return Error("Error while reading trajectory, read " + std::to_string(nsel / chunksize) + " sets of frames. Error: " + err.String());
    std::unique_ptr<DeviceTrajectory> d(new DeviceTrajectory());
    d->natoms_ = natoms;
    d->nselected_ = nsel;
    long first = 0, count = 0;
    shardRange(nsel, o.rank_, o.nranks_, &first, &count);
    d->nlocal_ = count;
    int rc = gcb_traj_create(o.device_, natoms, count > 0 ? count : 1, &d->t_);
    if (rc) return deviceError(rc, "gcb_traj_create");
    if (count > 0) {
        // C-owned pinned double buffer -> async copies (Go memory may not be handed to an async copy)
        const long per = std::max<long>(1, (long)((32u << 20) / fbytes));
        void* pin[2] = {nullptr, nullptr};
        for (int s = 0; s < 2; s++) {
            rc = gcb_pinned_alloc((size_t)per * fbytes, &pin[s]);
            if (rc) { gcb_pinned_free(pin[0]); return deviceError(rc, "gcb_pinned_alloc"); }
        }
        int slot = 0;
        Error up;
        for (long done = 0; done < count && !up; done += per, slot ^= 1) {
            const long n = std::min(per, count - done);
            rc = gcb_traj_wait(d->t_, slot);
            if (rc) { up = deviceError(rc, "gcb_traj_wait"); break; }
            memcpy(pin[slot], host.data() + (size_t)(first + done) * fbytes, (size_t)n * fbytes);
            rc = gcb_traj_upload_async(d->t_, done, pin[slot], n, raw ? layout : GCB_F64_AOS, raw ? scale : 1.0, slot);
            if (rc) up = deviceError(rc, "gcb_traj_upload_async");
        }
        gcb_traj_wait(d->t_, 0);
        gcb_traj_wait(d->t_, 1);
        gcb_pinned_free(pin[0]);
        gcb_pinned_free(pin[1]);
        if (up) return up;
    }
    gcb_traj_set_nframes(d->t_, count);
    *out = std::move(d);
    return Error();
}

Error DeviceTrajectory::MeanDeviations(const v3::Matrix& ref, const std::vector<int>& indexes, const Options& o, bool write_back,
                                       std::vector<double>* mean, const std::vector<int>* only) {
    if (ref.NVecs() != natoms_) return Error("chem.Super: Ill formed coordinates for Superposition", {"concproc"});
    mean->assign((size_t)natoms_, 0.0);
    if (only && !write_back && !only->empty() && 8 * only->size() <= (size_t)natoms_ && (o.comm_ || !o.allreduce_)) {
        std::vector<double> part(only->size());
        int rc = gcb_peratom_dev_sum_atoms(t_, 0, nlocal_, ref.data(), indexes.empty() ? nullptr : indexes.data(), (int)indexes.size(),
                                           only->data(), (int)only->size(), o.comm_, part.data());
        if (rc) return deviceError(rc, "RMSDTraj");
        const double total = (double)nselected_;
        for (size_t k = 0; k < only->size(); k++) (*mean)[(size_t)(*only)[k]] = part[k] / total;
        return Error();
    }
    int rc = gcb_peratom_dev_sum(t_, 0, nlocal_, ref.data(), indexes.empty() ? nullptr : indexes.data(), (int)indexes.size(),
                                 write_back ? 1 : 0, o.comm_, mean->data());
    if (rc) return deviceError(rc, "RMSDTraj");
    if (!o.comm_ && o.allreduce_) {
        if (Error e = o.allreduce_(mean->data(), (long)natoms_)) return e;
    }
    const double total = (double)nselected_;  // chunksread * chunksize (lovo.go:348)
    for (auto& v : *mean) v = v / total;
    return Error();
}

Error DeviceTrajectory::WriteDCD(const std::string& name) {
    std::unique_ptr<dcd::DCDWObj> w;
    if (Error e = dcd::DCDWObj::NewWriter(name, natoms_, &w)) {
        fprintf(stderr, "Couldn't open %s for writing. Will not write aligned trajectory\n", name.c_str());  // lovo.go:299-301
        return Error();
    }
    const long per = std::max<long>(1, (64L << 20) / (24L * natoms_));
    std::vector<double> buf((size_t)per * 3 * (size_t)natoms_);
    for (long done = 0; done < nlocal_; done += per) {
        const long n = std::min(per, nlocal_ - done);
        int rc = gcb_traj_download(t_, done, n, buf.data());
        if (rc) return deviceError(rc, "WNext");
        for (long f = 0; f < n; f++)
            if (Error e = w->WNextRaw(buf.data() + (size_t)f * 3 * (size_t)natoms_, natoms_)) return e;
    }
    return Error();
}

Error RMSDTraj(const chem::Atomer& mol, const v3::Matrix& ref, chem::ConcTraj& traj, const std::vector<int>& indexes, Options* o,
               std::vector<double>* rmsd, int* frames) {
    if (o->stream_ > 0) return RMSDTrajStreamed(mol, ref, traj, indexes, o, rmsd, frames);
    std::unique_ptr<DeviceTrajectory> d;
    if (Error e = DeviceTrajectory::Load(traj, mol.Len(), *o, &d)) { *frames = -1; return e; }
    const bool wb = !o->writeTraj_.empty();
    if (Error e = d->MeanDeviations(ref, indexes, *o, wb, rmsd)) { *frames = -1; return e; }
    if (wb)
        if (Error e = d->WriteDCD(o->writeTraj_)) return e;
    *frames = (int)d->SelectedFrames();
    return Error();
}

// Streams the frames RMSDTraj selects (chunks of o.cpus, begin / skip, dropped tail: lovo.go:306-326) through a device
// window.  Windows (whole chunks, at most `per` frames) are dealt to the ranks in turn; a rank reads and discards the
// windows of the others.  process(handle, first, n) runs for window k while the reader already fills window k + 1.
Error streamSelected(chem::ConcTraj& traj, int natoms, const Options& o,
                     const std::function<Error(gcb_traj*, long, long)>& process, long* nselected, long* nlocal) {
    if (!traj.Readable()) return Error("Traj object uninitialized to read", {"RMSDTraj"});
    if (o.cpus_ <= 0) return Error("align.RMSDTraj: cpus must be positive");
    if (traj.Len() != natoms) return Error("align.RMSDTraj: trajectory and molecule differ in the number of atoms");
    const int chunksize = o.cpus_;
    const long begin = o.begin_ / chunksize, skip = o.skip_ / chunksize;
    const bool raw = traj.HasRawFeed();
    // bulk path: whole runs of frames read() straight into the pinned buffer as they sit in the file (no begin / skip gaps)
    bool records = traj.HasRecordFeed() && begin == 0 && skip == 0 && traj.RecordsAvailable() >= 0;   // uniform records only
    long rec_bytes = 0, rec_off[3] = {0, 0, 0};
    double rec_scale = 1.0;
    if (records) {
        long got = 0;
        Error e = traj.NextRecords(nullptr, 0, &got, &rec_bytes, rec_off, &rec_scale);   // layout only
        if (e && e.last_frame) { *nselected = 0; *nlocal = 0; return Error(); }
        if (e || rec_bytes <= 0) records = false;
    }
    const size_t fbytes = records ? (size_t)rec_bytes : (raw ? sizeof(float) : sizeof(double)) * 3 * (size_t)natoms;
    long per = (long)((64u << 20) / fbytes);
    if (o.stream_ > 0 && o.stream_ < per) per = o.stream_;
    per = std::max<long>(chunksize, per / chunksize * chunksize);
    if (records && (size_t)per * fbytes > (64u << 20)) records = false;   // one chunk of records must fit a staging slot
    gcb_traj* t = nullptr;
    int rc = gcb_traj_create(o.device_, natoms, 2 * per, &t);
    if (rc) return deviceError(rc, "gcb_traj_create");
    // the two pinned staging buffers are kept for the next pass (LOVO streams the file once per iteration; pinning 128 MB costs
    // tens of milliseconds)
    static std::mutex pool_mu;
    static void* pool[2] = {nullptr, nullptr};
    static size_t pool_bytes = 0;
    std::unique_lock<std::mutex> pool_lock(pool_mu);   // one streamed pass at a time per process
    void* pin[2] = {nullptr, nullptr};
    Error err;
    if (pool_bytes < (size_t)per * fbytes) {
        for (int s = 0; s < 2; s++) { gcb_pinned_free(pool[s]); pool[s] = nullptr; }
        pool_bytes = 0;
        for (int s = 0; s < 2 && !err; s++)
            if ((rc = gcb_pinned_alloc((size_t)per * fbytes, &pool[s]))) err = deviceError(rc, "gcb_pinned_alloc");
        if (!err) pool_bytes = (size_t)per * fbytes;
    }
    pin[0] = pool[0];
    pin[1] = pool[1];
    std::vector<v3::Matrix> chunk((size_t)chunksize);
    if (!raw) for (auto& m : chunk) m = v3::Matrix::Zeros(natoms);
    int layout = GCB_F64_AOS;
    double scale = 1.0;
    long nsel = 0, mine_total = 0, window = 0, filled = 0;
    int slot = 0;
    long pending_first = -1, pending_n = 0;   // the window uploaded but not yet processed
    auto read_chunk = [&](char* dst) -> Error {   // dst == nullptr: read and discard
        if (raw) {
            for (int s = 0; s < chunksize; s++)
                if (Error e = traj.NextRaw(dst ? reinterpret_cast<float*>(dst + fbytes * (size_t)s) : nullptr, &layout, &scale)) return e;
            return Error();
        }
        std::vector<v3::Matrix*> slots((size_t)chunksize, nullptr);
        if (dst) for (int s = 0; s < chunksize; s++) slots[(size_t)s] = &chunk[(size_t)s];
        if (Error e = traj.NextConc(slots)) return e;
        if (dst) for (int s = 0; s < chunksize; s++) memcpy(dst + fbytes * (size_t)s, slots[(size_t)s]->data(), fbytes);
        return Error();
    };
    auto run_pending = [&]() -> Error {
        if (pending_first < 0) return Error();
        Error e = process(t, pending_first, pending_n);
        pending_first = -1;
        return e;
    };
    auto flush = [&]() -> Error {   // the window in `slot` is complete
        const bool mine = window % o.nranks_ == o.rank_;
        if (mine && filled > 0) {
            rc = gcb_traj_upload_async(t, (long)slot * per, pin[slot], filled, raw ? layout : GCB_F64_AOS, raw ? scale : 1.0, slot);
            if (rc) return deviceError(rc, "gcb_traj_upload_async");
            if (Error e = run_pending()) return e;   // the previous window, while this upload is in flight
            pending_first = (long)slot * per;
            pending_n = filled;
            mine_total += filled;
            slot ^= 1;
            rc = gcb_traj_wait(t, slot);                // the staging buffer about to be refilled is free again
            if (rc) return deviceError(rc, "gcb_traj_wait");
        }
        window++;
        filled = 0;
        return Error();
    };
    Error rerr;
    if (records) {
        rerr = chem::newLastFrameError("", "RMSDTraj");
        while (!err) {
            const bool mine = window % o.nranks_ == o.rank_;
            long got = 0;
            Error e = traj.NextRecords(mine ? static_cast<char*>(pin[slot]) : nullptr, per, &got, &rec_bytes, rec_off, &rec_scale);
            if (e) { rerr = e; break; }
            const long usable = got - got % chunksize;   // a chunk the file cannot fill is dropped (:323-326)
            nsel += usable;
            if (mine && usable > 0) {
                rc = gcb_traj_upload_records_async(t, (long)slot * per, pin[slot], usable, rec_bytes, rec_off[0], rec_off[1], rec_off[2],
                                                   rec_scale, slot);
                if (rc) { err = deviceError(rc, "gcb_traj_upload_records_async"); break; }
                if ((err = run_pending())) break;
                pending_first = (long)slot * per;
                pending_n = usable;
                mine_total += usable;
                slot ^= 1;
                if ((rc = gcb_traj_wait(t, slot))) { err = deviceError(rc, "gcb_traj_wait"); break; }
            }
            window++;
            if (got < per) break;   // the file ended inside this window
        }
    }
    for (long read = 0; !err && !records; read++) {
        if (read < begin || read % (skip + 1) != 0) {
            rerr = read_chunk(nullptr);
            if (rerr) break;
        }
        const bool mine = window % o.nranks_ == o.rank_;
        rerr = read_chunk(mine ? static_cast<char*>(pin[slot]) + fbytes * (size_t)filled : nullptr);
        if (rerr) break;   // a chunk that cannot be filled is dropped (:323-326)
        filled += chunksize;
        nsel += chunksize;
        if (filled + chunksize > per) err = flush();
    }
    if (!err && !rerr.last_frame)  // lovo.go:343-347
        err = Error("Error while reading trajectory, read " + std::to_string(nsel / chunksize) + " sets of frames. Error: " + rerr.String());
    if (!err && !records) err = flush();
    if (!err) err = run_pending();
    gcb_traj_wait(t, 0);
    gcb_traj_wait(t, 1);
    gcb_traj_destroy(t);
    *nselected = nsel;
    *nlocal = mine_total;
    return err;
}

Error RMSDTrajStreamed(const chem::Atomer& mol, const v3::Matrix& ref, chem::ConcTraj& traj, const std::vector<int>& indexes, Options* o,
                       std::vector<double>* rmsd, int* frames) {
    const int natoms = mol.Len();
    *frames = -1;
    if (ref.NVecs() != natoms) return Error("chem.Super: Ill formed coordinates for Superposition", {"concproc"});
    rmsd->assign((size_t)natoms, 0.0);
    std::vector<double> part((size_t)natoms);
    const bool wb = !o->writeTraj_.empty();
    std::unique_ptr<dcd::DCDWObj> w;
    if (wb && dcd::DCDWObj::NewWriter(o->writeTraj_, natoms, &w))
        fprintf(stderr, "Couldn't open %s for writing. Will not write aligned trajectory\n", o->writeTraj_.c_str());  // lovo.go:299-301
    std::vector<double> moved;
    auto process = [&](gcb_traj* t, long first, long n) -> Error {
        int rc = gcb_peratom_dev_sum(t, first, n, ref.data(), indexes.empty() ? nullptr : indexes.data(), (int)indexes.size(),
                                     wb ? 1 : 0, nullptr, part.data());
        if (rc) return deviceError(rc, "RMSDTraj");
        for (int i = 0; i < natoms; i++) (*rmsd)[(size_t)i] += part[(size_t)i];
        if (wb && w) {   // the superposed frames of this window, in reading order (lovo.go:337-339)
            moved.resize((size_t)n * 3 * (size_t)natoms);
            rc = gcb_traj_download(t, first, n, moved.data());
            if (rc) return deviceError(rc, "WNext");
            for (long f = 0; f < n; f++)
                if (Error e = w->WNextRaw(moved.data() + (size_t)f * 3 * (size_t)natoms, natoms)) return e;
        }
        return Error();
    };
    long nsel = 0, nlocal = 0;
    if (Error e = streamSelected(traj, natoms, *o, process, &nsel, &nlocal)) return e;
    if (o->comm_) {
        int rc = gcb_allreduce_sum_f64(o->comm_, rmsd->data(), natoms);
        if (rc) return deviceError(rc, "RMSDTraj");
    } else if (o->allreduce_) {
        if (Error e = o->allreduce_(rmsd->data(), (long)natoms)) return e;
    }
    const double total = (double)nsel;  // chunksread * chunksize (lovo.go:348)
    for (auto& v : *rmsd) v = v / total;
    *frames = (int)nsel;
    return Error();
}

Error opentraj(const std::string& name, std::unique_ptr<chem::ConcTraj>* out) {
    const size_t dot = name.find_last_of('.');
    std::string ext = dot == std::string::npos ? name : name.substr(dot + 1);
    std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (ext == "dcd") {
        std::unique_ptr<dcd::DCDObj> d;
        if (Error e = dcd::DCDObj::New(name, &d)) return e;
        *out = std::move(d);
        return Error();
    }
    if (ext == "xtc") {   // lovo.go:157-160
        std::unique_ptr<xtc::XTCObj> x;
        if (Error e = xtc::XTCObj::New(name, &x)) return e;
        *out = std::move(x);
        return Error();
    }
    if (ext == "stf") {   // lovo.go:161-162
        std::unique_ptr<stf::StfR> x;
        if (Error e = stf::StfR::New(name, &x)) return e;
        *out = std::move(x);
        return Error();
    }
    if (ext == "pdb")
        return Error("Trajectory format 'pdb' (multi-model PDB text, chem.PDBFileRead) is outside this build (SURVEY.md §2); "
                     "feed its frames through a chem::ConcTraj and call align::RMSDTraj");
    return Error("Trajectory fromat not supported. Must have concurrent read support in goChem");  // lovo.go:168-169
}

Error LOVOResident(const chem::Atomer& mol, const v3::Matrix& ref, std::unique_ptr<DeviceTrajectory>& d, LOVOReturn* out, Options* o,
                   const std::function<Error()>& reload) {
    const std::vector<int> fullindexes = res2atoms(mol, o->atomNames_, o->chains_);
    const std::string printtraj = o->writeTraj_;
    o->writeTraj_.clear();  // lovo.go:209-211
    LovoState st(fullindexes, o->lessThanRMSD_, o->minimumN_);
    std::vector<double> rmsd;
    const bool dbg = getenv("GCH_DEBUG") != nullptr;
    // The whole iteration in the library, bookkeeping on the device (gcb_lovo_resident), when nothing needs the host in between:
    // no aligned trajectory to write, the exchange (if any) through the library's communicator.  1 = not taken: the loop below.
    const bool host_loop = getenv("GCH_LOVO_HOST") != nullptr;   // tests compare the two
    if (!host_loop && !dbg && printtraj.empty() && !fullindexes.empty() && (o->comm_ || !o->allreduce_) && ref.NVecs() == d->Natoms()) {
        const size_t nc = fullindexes.size();
        std::vector<int> old(nc), ids(nc);
        std::vector<double> vals(nc);
        int n = 0, iters = 0, dis = 0;
        const int rc = gcb_lovo_resident(d->Handle(), 0, d->LocalFrames(), ref.data(), fullindexes.data(), (int)nc, (double)d->SelectedFrames(),
                                         o->lessThanRMSD_, o->minimumN_, o->maxIter_, o->comm_, &n, &iters, &dis, old.data(), ids.data(), vals.data());
        if (rc == 0) {
            st.Restore(std::move(old), std::move(ids), std::move(vals), n, iters, dis);
            st.Finish(mol, (int)d->SelectedFrames(), out);
            return Error();
        }
        if (rc != 1) return deviceError(rc, "LOVO");
    }
    for (;;) {
        if (dbg) fprintf(stderr, "[LOVO] iteration %d, fit on %zu atoms, local frames %ld\n", st.Iterations(), st.FitIndexes().size(), d->LocalFrames());
        if (st.Disagreement() == 1 && !printtraj.empty()) o->writeTraj_ = printtraj;  // :226-228
        const bool wb = !o->writeTraj_.empty();
        const auto t0 = std::chrono::steady_clock::now();
        if (Error e = d->MeanDeviations(ref, st.FitIndexes(), *o, wb, &rmsd, &fullindexes)) return e;
        const auto t1 = std::chrono::steady_clock::now();
        if (wb) {
            if (Error e = d->WriteDCD(o->writeTraj_)) return e;
            // the pass rewrote the resident frames; later passes start from the source again (the reference
            // re-reads the file on every iteration, lovo.go:222).  reload() may replace d.
            if (reload) { if (Error e = reload()) return e; }
        }
        Error serr;
        const bool converged = st.Step(rmsd, &serr);
        if (dbg) {
            const auto t2 = std::chrono::steady_clock::now();
            fprintf(stderr, "[LOVO]   pass %.1f us, bookkeeping %.1f us\n", std::chrono::duration<double, std::micro>(t1 - t0).count(),
                    std::chrono::duration<double, std::micro>(t2 - t1).count());
        }
        if (serr) return serr;
        if (converged) break;
        if (st.Iterations() >= o->maxIter_) return Error("align.LOVO: no convergence after " + std::to_string(st.Iterations()) + " iterations");
    }
    st.Finish(mol, (int)d->SelectedFrames(), out);
    return Error();
}

Error LOVO(const chem::Atomer& mol, const v3::Matrix& ref, const std::string& trajname, LOVOReturn* out, Options* opt) {
    std::unique_ptr<Options> own;
    Options* o = opt;
    if (!o) { own.reset(DefaultOptions()); o = own.get(); }
    std::unique_ptr<DeviceTrajectory> d;
    auto load = [&]() -> Error {
        std::unique_ptr<chem::ConcTraj> traj;
        if (Error e = opentraj(trajname, &traj)) return e;  // :222 (the reference re-opens it on every iteration)
        std::unique_ptr<DeviceTrajectory> fresh;
        Error e = DeviceTrajectory::Load(*traj, mol.Len(), *o, &fresh);
        traj->Close();
        if (!e) d = std::move(fresh);
        return e;
    };
    // Resident when the frames fit in device memory (a DCD file holds float32, the device float64: 2 bytes per byte),
    // otherwise -- or when the caller set a stream window -- every iteration streams the file again (lovo.go:222).
    bool streamed = o->stream_ > 0;
    if (!streamed) {
        struct stat st;
        int sm = 0, ccmaj = 0, ccmin = 0;
        size_t total_mem = 0;
        if (stat(trajname.c_str(), &st) == 0 && gcb_device_info(o->device_, &sm, &total_mem, &ccmaj, &ccmin) == 0 && total_mem > 0)
            streamed = 2.0 * (double)st.st_size / (double)(o->nranks_ > 0 ? o->nranks_ : 1) > 0.8 * (double)total_mem;
    }
    if (!streamed) {
        if (Error e = load()) return e;
        return LOVOResident(mol, ref, d, out, o, load);
    }
    const std::vector<int> fullindexes = res2atoms(mol, o->atomNames_, o->chains_);
    const std::string printtraj = o->writeTraj_;
    o->writeTraj_.clear();  // lovo.go:209-211
    LovoState st(fullindexes, o->lessThanRMSD_, o->minimumN_);
    std::vector<double> rmsd;
    int frames = 0;
    for (;;) {
        if (st.Disagreement() == 1 && !printtraj.empty()) o->writeTraj_ = printtraj;  // :226-228
        std::unique_ptr<chem::ConcTraj> traj;
        if (Error e = opentraj(trajname, &traj)) return e;
        Error e = RMSDTrajStreamed(mol, ref, *traj, st.FitIndexes(), o, &rmsd, &frames);
        traj->Close();
        if (e) return e;
        Error serr;
        const bool converged = st.Step(rmsd, &serr);
        if (serr) return serr;
        if (converged) break;
        if (st.Iterations() >= o->maxIter_) return Error("align.LOVO: no convergence after " + std::to_string(st.Iterations()) + " iterations");
    }
    st.Finish(mol, frames, out);
    return Error();
}

}  // namespace align

// =====================================================================================================================
namespace xtc {

namespace {
struct XdrApi {
    void* h = nullptr;
    int (*read_xtc_natoms)(char*, int*) = nullptr;
    void* (*xdrfile_open)(const char*, const char*) = nullptr;
    int (*xdrfile_close)(void*) = nullptr;
    int (*read_xtc)(void*, int, int*, float*, float (*)[3], float (*)[3], float*) = nullptr;
    int (*write_xtc)(void*, int, int, float, float (*)[3], float (*)[3], float) = nullptr;
    std::string why;
};
XdrApi* xdr() {
    static XdrApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("GOCHEM_XDRFILE");
        const char* names[] = {env, "libxdrfile.so", "libxdrfile.so.4", "libxdrfile.so.2.1.2"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.h) break;
        }
        if (!api.h) { api.why = "libxdrfile is not loadable (set GOCHEM_XDRFILE to its path)"; return; }
        api.read_xtc_natoms = (decltype(api.read_xtc_natoms))dlsym(api.h, "read_xtc_natoms");
        api.xdrfile_open = (decltype(api.xdrfile_open))dlsym(api.h, "xdrfile_open");
        api.xdrfile_close = (decltype(api.xdrfile_close))dlsym(api.h, "xdrfile_close");
        api.read_xtc = (decltype(api.read_xtc))dlsym(api.h, "read_xtc");
        api.write_xtc = (decltype(api.write_xtc))dlsym(api.h, "write_xtc");
        if (!api.read_xtc_natoms || !api.xdrfile_open || !api.xdrfile_close || !api.read_xtc) {
            api.why = "libxdrfile lacks read_xtc / xdrfile_open";
            dlclose(api.h);
            api.h = nullptr;
        }
    });
    return &api;
}
Error xtcError(const std::string& msg, const std::string& file, std::vector<std::string> deco) {
    return Error(msg + (file.empty() ? "" : " (" + file + ")"), std::move(deco));
}
}  // namespace

bool XTCObj::Available(std::string* why) {
    XdrApi* a = xdr();
    if (!a->h && why) *why = a->why;
    return a->h != nullptr;
}

Error XTCObj::New(const std::string& filename, std::unique_ptr<XTCObj>* out) {
    XdrApi* a = xdr();
    if (!a->h) return xtcError("xtc: " + a->why + "; the XTC codec is a third-party CPU library", filename, {"initRead"});
    std::unique_ptr<XTCObj> x(new XTCObj());
    x->filename_ = filename;
    std::vector<char> name(filename.begin(), filename.end());
    name.push_back(0);
    int natoms = 0;
    a->read_xtc_natoms(name.data(), &natoms);
    if (natoms == 0) return xtcError("Probem opening the file", filename, {"C.read_natoms", "initRead"});   // xtc.go:84-86
    x->fp_ = a->xdrfile_open(name.data(), "r");
    if (!x->fp_) return xtcError("Probem opening the file", filename, {"C.openfile", "initRead"});          // xtc.go:79-81
    x->natoms_ = natoms;
    x->coords_.assign(3 * (size_t)natoms, 0.0f);
    x->readable_ = true;
    *out = std::move(x);
    return Error();
}

XTCObj::~XTCObj() { Close(); }

void XTCObj::Close() {
    if (fp_) xdr()->xdrfile_close(fp_);
    fp_ = nullptr;
    readable_ = false;
}

Error XTCObj::readInto(float* dst) {   // libgoxtc.c:87-104 get_coords
    if (!readable_) return xtcError("Traj object uninitialized to read", filename_, {"Next"});
    int step = 0;
    float time = 0, prec = 0;
    const int worked = xdr()->read_xtc(fp_, natoms_, &step, &time, reinterpret_cast<float (*)[3]>(box_),
                                       reinterpret_cast<float (*)[3]>(dst ? dst : coords_.data()), &prec);
    if (worked == 11) { Close(); return chem::newLastFrameError(filename_, "Next"); }   // xtc.go:114-117
    if (worked != 0) { Close(); return xtcError("Error reading frame", filename_, {"Next"}); }
    return Error();
}

Error XTCObj::Next(v3::Matrix* output, std::vector<double>* box) {
    if (Error e = readInto(nullptr)) return e;
    if (!output) return Error();
    if (output->NVecs() < natoms_) throw v3::PanicMsg("Buffer v3.Matrix too small to hold trajectory frame");
    for (int j = 0; j < natoms_; j++)
        for (int k = 0; k < 3; k++) output->Set(j, k, 10 * (double)coords_[(size_t)(k + 3 * j)]);   // nm to Angstroms (xtc.go:129-134)
    if (box && box->size() >= 9)
        for (int k = 0; k < 9; k++) (*box)[(size_t)k] = 10 * (double)box_[k];
    return Error();
}

Error XTCObj::NextConc(std::vector<v3::Matrix*>& frames) {
    if (!readable_) return xtcError("Traj object uninitialized to read", filename_, {"NextConc"});
    for (auto* slot : frames)
        if (Error e = Next(slot)) return errDecorate(e, "NextConc");
    return Error();
}

Error XTCObj::NextRaw(float* dst, int* layout, double* scale) {
    *layout = GCB_F32_AOS;
    *scale = 10.0;
    return readInto(dst);
}

Error XTCObj::WriteFile(const std::string& filename, const double* frames, long nframes, int natoms) {
    XdrApi* a = xdr();
    if (!a->h || !a->write_xtc) return xtcError("xtc: libxdrfile (with write_xtc) is not loadable", filename, {"WriteFile"});
    void* fp = a->xdrfile_open(filename.c_str(), "w");
    if (!fp) return xtcError("cannot open for writing", filename, {"WriteFile"});
    std::vector<float> x(3 * (size_t)natoms);
    float box[9] = {10, 0, 0, 0, 10, 0, 0, 0, 10};
    Error err;
    for (long f = 0; f < nframes && !err; f++) {
        for (size_t i = 0; i < x.size(); i++) x[i] = (float)(frames[(size_t)f * x.size() + i] / 10.0);
        if (a->write_xtc(fp, natoms, (int)f, (float)f, reinterpret_cast<float (*)[3]>(box), reinterpret_cast<float (*)[3]>(x.data()), 1000.0f) != 0)
            err = xtcError("write_xtc failed", filename, {"WriteFile"});
    }
    a->xdrfile_close(fp);
    return err;
}

}  // namespace xtc

namespace chem {

Error Topology::Masses(std::vector<double>* out) const {  // chem.go:291-301
    out->assign((size_t)Len(), 0.0);
    for (int i = 0; i < Len(); i++) {
        if (Atoms[(size_t)i].Mass == 0) {
            out->clear();
            return Error("goChem: Not all the masses have been obtained: " + std::to_string(i), {"Topology.Masses"});
        }
        (*out)[(size_t)i] = Atoms[(size_t)i].Mass;
    }
    return Error();
}

void Topology::SomeAtoms(const Atomer& T2, const std::vector<int>& indexes) {
    Atoms.clear();
    for (int i : indexes) Atoms.push_back(T2.Atom(i));
}

}  // namespace chem

namespace solv {

static const double kAppZero = 0.000000000001;  // matrixhelp.go:40

Options* DefaultOptions() {
    Options* o = new Options();
    const unsigned hc = std::thread::hardware_concurrency();
    o->cpus_ = hc ? (int)hc : 1;  // runtime.NumCPU()
    return o;
}

// eigenvalues of a symmetric 3 x 3 matrix (cyclic Jacobi): stands in for mat.Eigen inside v3.EigenWrap (v3/gonum.go:363-401)
static void symEigenvalues3(const double* m, double* ev) {
    double a[3][3] = {{m[0], m[1], m[2]}, {m[3], m[4], m[5]}, {m[6], m[7], m[8]}};
    for (int sweep = 0; sweep < 64; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; k++) {  // columns
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {  // rows
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
            }
    }
    ev[0] = a[0][0]; ev[1] = a[1][1]; ev[2] = a[2][2];
}

// chem.Rhos on the nine numbers of a moment tensor
static Error rhosOf(const double* moment9, std::vector<double>* rhos) {
    double ev[3];
    symEigenvalues3(moment9, ev);
    std::sort(ev, ev + 3);  // EigenWrap sorts the pairs by eigenvalue, ascending (v3/gonum.go:343-345, 400-401)
    // geometric.go:534-537 tests evals[2] of the ascending list, the LARGEST eigenvalue: only a structure without any extent fails,
    // and the values go back in ascending order then
    if (ev[2] <= kAppZero) {
        rhos->assign({ev[0], ev[1], ev[2]});
        return Error("goChem: Molecule colapsed to a single point. Check for blackholes", {"Rhos"});
    }
    rhos->assign({ev[2], ev[1], ev[0]});  // largest first (geometric.go:538-544)
    return Error();
}

Error EllipsoidAxes(const v3::Matrix& coords, double /*epsilon*/, const chem::Topology* mol, std::vector<double>* rhos) {
    std::vector<double> masses;
    Error err2;
    if (mol) {
        if (Error e = mol->Masses(&masses)) { masses.clear(); err2 = e; }  // solvation.go:46-51
    }
    v3::Matrix moment;
    if (Error e = chem::MomentTensor(coords, &moment, masses)) return e;
    if (Error e = rhosOf(moment.data(), rhos)) return e;
    return err2;
}

}  // namespace solv

namespace chem {

Error Rhos(const v3::Matrix& momentTensor, std::vector<double>* rhos, double /*epsilon*/) {
    if (momentTensor.NVecs() != 3) return Error("error in EigenWrap: mat.Eigen.Factorize", {"Rhos"});
    return solv::rhosOf(momentTensor.data(), rhos);
}

Error RhoShapeIndexes(const std::vector<double>& rhos, double* linear, double* circular) {
    if (rhos.size() < 3) { *linear = -1; *circular = -1; return Error("goChe: Not enough or nil rhos", {"RhoShapeIndexes"}); }
    *linear = (1 - (rhos[1] / rhos[0])) * 100;     // prolate
    *circular = (1 - (rhos[2] / rhos[1])) * 100;   // oblate
    return Error();
}

Error EasyShape(const v3::Matrix& coords, double /*epsilon*/, const Topology* mol, double* linear, double* circular) {
    *linear = -1; *circular = -1;
    std::vector<double> masses;
    Error err2;
    if (mol) {
        if (Error e = mol->Masses(&masses)) { masses.clear(); err2 = e; }   // handy.go:110-116
    }
    v3::Matrix moment;
    if (Error e = MomentTensor(coords, &moment, masses)) return e;
    std::vector<double> rhos;
    if (Error e = Rhos(moment, &rhos)) return e;
    if (Error e = RhoShapeIndexes(rhos, linear, circular)) return e;
    return err2;
}

Error EasyShapeTraj(gcb_traj* t, long first, long nframes, const std::vector<double>& masses, std::vector<double>* linear,
                    std::vector<double>* circular) {
    linear->assign((size_t)nframes, -1.0);
    circular->assign((size_t)nframes, -1.0);
    if (nframes <= 0) return Error();
    if (!masses.empty() && (int)masses.size() != gcb_traj_natoms(t)) return Error("MomentTensor: as many masses as atoms are needed", {"EasyShapeTraj"});
    std::vector<double> center(3 * (size_t)nframes), moment(9 * (size_t)nframes);
    int rc = gcb_moments(t, first, nframes, masses.empty() ? nullptr : masses.data(), nullptr, 0, center.data(), moment.data());
    if (rc) return deviceError(rc, "MomentTensor");
    Error firstErr;
    std::vector<double> rhos;
    for (long f = 0; f < nframes; f++) {
        Error e = solv::rhosOf(moment.data() + 9 * (size_t)f, &rhos);
        if (!e) e = RhoShapeIndexes(rhos, &(*linear)[(size_t)f], &(*circular)[(size_t)f]);
        if (e && !firstErr) firstErr = errDecorate(e, "EasyShapeTraj: frame " + std::to_string(first + f));
    }
    return firstErr;
}

}  // namespace chem

namespace solv {

Error MDFFromCDF(std::vector<double>& ret, int framesread, double A, double B, double step, std::vector<double>* ret2) {
    if (A < 1.0 || B < 1.0)
        return Error("goChem/solvation.MDFFromCDF: A and B are the ratio between the lartest axis of an elipsoid to the smallest, they can't be smaller than 1.0");
    const size_t n = ret.size();
    ret2->assign(n, 0.0);
    const double vp = (4.0 / 3.0) * M_PI * A * B;
    double vol, acc = 0.0;
    for (size_t i = 0; i < n; i++) {
        if (i >= 1) {
            acc = acc + ret[i - 1];
            (*ret2)[i] = ret[i];
            ret[i] = ret[i] - acc;
        }
    }
    for (size_t i = 0; i < n; i++) {
        if (i >= 1) {
            const double fi = (double)i;
            vol = vp * (std::pow((fi + 1) * step, 3) - std::pow(fi * step, 3));
        } else {
            vol = vp * std::pow(step, 3);
        }
        ret[i] = ret[i] / (double)framesread;
        (*ret2)[i] = (*ret2)[i] / (double)framesread;
        ret[i] = ret[i] / vol;
    }
    if (n) {
        const double last = ret[n - 1];
        for (size_t i = 0; i < n; i++) ret[i] = ret[i] / last;
    }
    return Error();
}

static bool isInString(const std::vector<std::string>& c, const std::string& t) { return std::find(c.begin(), c.end(), t) != c.end(); }

Error ConcMolRDF(chem::ConcTraj& traj, const chem::Molecule& mol, const std::vector<int>& refindexes,
                 const std::vector<std::string>& residues, std::vector<double>* rdf, std::vector<double>* cumulative, Options* opt,
                 int* frames_read) {
    std::unique_ptr<Options> owned;
    Options* o = opt;
    if (!o) { owned.reset(DefaultOptions()); o = owned.get(); }
    const int natoms = mol.Len();
    for (int i : refindexes)
        if (i < 0 || i >= natoms) throw v3::PanicMsg("goChem/v3: index out of range");  // SomeVecs panics (v3/gonum.go:513-532)
    double A = 1.0, B = 1.0;
    if (natoms > 1) {  // solvation.go:149-162
        if (mol.Coords.empty()) return Error("goChem: the molecule holds no coordinates", {"ConcMolRDF"});
        v3::Matrix scoords = v3::Matrix::Zeros((int)refindexes.size());
        scoords.SomeVecs(mol.Coords[0], refindexes);
        chem::Topology smol;
        smol.SomeAtoms(mol, refindexes);
        std::vector<double> elip;
        if (Error e = EllipsoidAxes(scoords, 0.0001, &smol, &elip)) return e;
        A = elip[0] / elip[2];
        B = elip[1] / elip[2];
    }
    // the frame loop: every full chunk of o.cpus frames is read (solvation.go:166-214); the frames go to the device once
    align::Options ao;
    ao.cpus_ = o->cpus_;
    ao.device_ = o->device_;
    ao.rank_ = o->rank_;
    ao.nranks_ = o->nranks_;
    std::vector<int> molid((size_t)natoms), chain((size_t)natoms);
    std::vector<unsigned char> flags((size_t)natoms);
    std::map<std::string, int> chain_ids;
    chain_ids[""] = 0;
    static const std::vector<std::string> waters = {"SOL", "WAT", "HOH"};  // solvation.go:481
    for (int i = 0; i < natoms; i++) {
        const chem::Atom& at = mol.Atom(i);
        molid[(size_t)i] = at.MolID;
        auto it = chain_ids.find(at.Chain);
        if (it == chain_ids.end()) it = chain_ids.emplace(at.Chain, (int)chain_ids.size()).first;
        chain[(size_t)i] = it->second;
        flags[(size_t)i] = (unsigned char)((isInString(residues, at.MolName) ? 1 : 0) | (isInString(waters, at.MolName) ? 2 : 0));
    }
    const int totalsteps = (int)(o->end_ / o->step_);
    std::vector<double> ret((size_t)(totalsteps > 0 ? totalsteps : 0), 0.0);
    long read = 0;
    if (o->stream_ > 0) {
        // bounded memory: the counts of every window are integers and simply add up; the ranks' totals are combined at the end
        ao.stream_ = o->stream_;
        std::vector<double> part(ret.size());
        auto process = [&](gcb_traj* t, long first, long n) -> Error {
            long r = 0;
            int rc = gcb_rdf_cdf_sum(t, first, n, molid.data(), chain.data(), flags.data(), refindexes.data(), (int)refindexes.size(),
                                     o->com_ ? 1 : 0, o->step_, o->end_, 1, nullptr, part.data(), &r);
            if (rc) return deviceError(rc, "ConcMolRDF");
            for (size_t k = 0; k < ret.size(); k++) ret[k] += part[k];
            return Error();
        };
        long nsel = 0, nlocal = 0;
        if (Error e = align::streamSelected(traj, natoms, ao, process, &nsel, &nlocal)) return errDecorate(e, "ConcMolRDF");
        read = nsel;
        if (o->comm_) {
            int rc = gcb_allreduce_sum_f64(o->comm_, ret.data(), (long)ret.size());
            if (rc) return deviceError(rc, "ConcMolRDF");
        }
    } else {
        std::unique_ptr<align::DeviceTrajectory> d;
        if (Error e = align::DeviceTrajectory::Load(traj, natoms, ao, &d)) return errDecorate(e, "ConcMolRDF");
        int rc = gcb_rdf_cdf_sum(d->Handle(), 0, d->LocalFrames(), molid.data(), chain.data(), flags.data(), refindexes.data(),
                                 (int)refindexes.size(), o->com_ ? 1 : 0, o->step_, o->end_, 1, o->comm_, ret.data(), &read);
        if (rc) return deviceError(rc, "ConcMolRDF");
    }
    if (frames_read) *frames_read = (int)read;
    if (Error e = MDFFromCDF(ret, (int)read, A, B, o->step_, cumulative)) return e;
    *rdf = std::move(ret);
    return Error();
}

}  // namespace solv
}  // namespace gochem
