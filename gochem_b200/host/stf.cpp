// stf.cpp -- host mirror of goChem's STF trajectory reader / writer (/root/reference/traj/stf/stf.go).
//
// The format is text inside one compressed stream (see gochem.hpp, namespace stf).  Decompression (zlib for gzip and
// raw deflate, libzstd through dlopen for zstd) and the integer parsing stay on the CPU, as every codec of this build
// does; the parsed integers go to the device as they are (GCB_I32_AOS) and become float64(f)/100 there.
#include "gochem.hpp"

#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/gochem_b200.h"

namespace gochem {
namespace stf {

namespace {

Error stfError(const std::string& msg, const std::string& file, std::vector<std::string> deco = {}) {
    return Error("stf file " + file + " error: " + msg, std::move(deco));   // stf.go:479-481
}

constexpr const char* TrajUnIniRead = "Traj object uninitialized to read";    // stf.go:507-508 (both texts say "read")
constexpr const char* TrajUnIniWrite = "Traj object uninitialized to read";
constexpr const char* NilCoordinates = "Given nil coordinates";

// ---- libzstd, bound at run time (the image ships the shared object without its header) ----
struct ZInBuf { const void* src; size_t size; size_t pos; };
struct ZOutBuf { void* dst; size_t size; size_t pos; };
struct ZstdApi {
    void* h = nullptr;
    std::string why;
    void* (*createDStream)() = nullptr;
    size_t (*initDStream)(void*) = nullptr;
    size_t (*decompressStream)(void*, ZOutBuf*, ZInBuf*) = nullptr;
    size_t (*freeDStream)(void*) = nullptr;
    void* (*createCStream)() = nullptr;
    size_t (*initCStream)(void*, int) = nullptr;
    size_t (*compressStream)(void*, ZOutBuf*, ZInBuf*) = nullptr;
    size_t (*endStream)(void*, ZOutBuf*) = nullptr;
    size_t (*freeCStream)(void*) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    const char* (*getErrorName)(size_t) = nullptr;
};

ZstdApi* zstd() {
    static ZstdApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("GOCHEM_ZSTD");
        const char* names[] = {env, "libzstd.so.1", "libzstd.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.h) break;
        }
        if (!api.h) { api.why = "libzstd is not loadable (set GOCHEM_ZSTD to its path)"; return; }
#define GCH_SYM(field, name) api.field = (decltype(api.field))dlsym(api.h, name)
        GCH_SYM(createDStream, "ZSTD_createDStream"); GCH_SYM(initDStream, "ZSTD_initDStream");
        GCH_SYM(decompressStream, "ZSTD_decompressStream"); GCH_SYM(freeDStream, "ZSTD_freeDStream");
        GCH_SYM(createCStream, "ZSTD_createCStream"); GCH_SYM(initCStream, "ZSTD_initCStream");
        GCH_SYM(compressStream, "ZSTD_compressStream"); GCH_SYM(endStream, "ZSTD_endStream");
        GCH_SYM(freeCStream, "ZSTD_freeCStream"); GCH_SYM(isError, "ZSTD_isError"); GCH_SYM(getErrorName, "ZSTD_getErrorName");
#undef GCH_SYM
        if (!api.createDStream || !api.initDStream || !api.decompressStream || !api.freeDStream || !api.createCStream || !api.initCStream ||
            !api.compressStream || !api.endStream || !api.freeCStream || !api.isError || !api.getErrorName) {
            api.why = "libzstd lacks the streaming API";
            dlclose(api.h);
            api.h = nullptr;
        }
    });
    return &api;
}

enum class Codec { Zstd, Gzip, Flate, Lzw };

// the reference looks at the last letter of the lower-cased name only (stf.go:130, 260)
Codec codecOf(const std::string& name) {
    const char c = name.empty() ? 'f' : (char)std::tolower((unsigned char)name.back());
    switch (c) {
        case 'l': return Codec::Lzw;
        case 'z': return Codec::Gzip;
        case 'r': return Codec::Flate;
        default: return Codec::Zstd;   // 'f', 's' and everything else
    }
}

}  // namespace

// ---------------- decompressed byte stream, read line by line ----------------
class ByteStream {
  public:
    virtual ~ByteStream() = default;
    enum { kLine = 0, kEnd = 1, kEndInsideLine = 2, kError = 3 };
    // the next line without its '\n' (valid until the next call)
    int line(const char** p, size_t* n, std::string* why) {
        for (;;) {
            if (beg_ < end_) {
                const char* nl = (const char*)memchr(buf_.data() + beg_, '\n', end_ - beg_);
                if (nl) {
                    *p = buf_.data() + beg_;
                    *n = (size_t)(nl - *p);
                    beg_ += *n + 1;
                    return kLine;
                }
            }
            if (eof_) {
                if (beg_ == end_) return kEnd;
                beg_ = end_;   // Go's ReadBytes hands back the partial line together with io.EOF; callers only look at the error
                return kEndInsideLine;
            }
            if (beg_ > 0) {
                memmove(buf_.data(), buf_.data() + beg_, end_ - beg_);
                end_ -= beg_;
                beg_ = 0;
            }
            if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);
            const long got = fill(buf_.data() + end_, buf_.size() - end_, why);
            if (got < 0) return kError;
            if (got == 0) eof_ = true;
            end_ += (size_t)got;
        }
    }

  protected:
    virtual long fill(char* dst, size_t cap, std::string* why) = 0;   // > 0 bytes, 0 = end of stream, -1 = error
    // compressed input shared by the codecs
    bool more_input(std::string* why) {
        if (in_pos_ < in_len_) return true;
        if (!fp_ || feof(fp_)) return false;
        in_len_ = fread(in_.data(), 1, in_.size(), fp_);
        in_pos_ = 0;
        if (in_len_ == 0 && ferror(fp_)) { *why = "read error"; return false; }
        return in_len_ > 0;
    }
    FILE* fp_ = nullptr;
    std::vector<char> in_ = std::vector<char>(1 << 18);
    size_t in_pos_ = 0, in_len_ = 0;

  private:
    std::vector<char> buf_ = std::vector<char>(1 << 20);
    size_t beg_ = 0, end_ = 0;
    bool eof_ = false;
};

namespace {

class ZlibStream : public ByteStream {
  public:
    ZlibStream(FILE* fp, bool gzip) : gzip_(gzip) {
        fp_ = fp;
        memset(&zs_, 0, sizeof zs_);
        ok_ = inflateInit2(&zs_, gzip ? 16 + MAX_WBITS : -MAX_WBITS) == Z_OK;
    }
    ~ZlibStream() override {
        if (ok_) inflateEnd(&zs_);
        if (fp_) fclose(fp_);
    }
    bool ok() const { return ok_; }

  protected:
    long fill(char* dst, size_t cap, std::string* why) override {
        if (done_) return 0;
        zs_.next_out = (Bytef*)dst;
        zs_.avail_out = (uInt)std::min<size_t>(cap, 1u << 30);
        const uInt want = zs_.avail_out;
        while (zs_.avail_out == want) {
            if (!more_input(why)) {
                if (!why->empty()) return -1;
                if (member_open_) { *why = "unexpected EOF"; return -1; }   // Go: io.ErrUnexpectedEOF
                done_ = true;
                return 0;
            }
            zs_.next_in = (Bytef*)(in_.data() + in_pos_);
            zs_.avail_in = (uInt)(in_len_ - in_pos_);
            member_open_ = true;
            const int rc = inflate(&zs_, Z_NO_FLUSH);
            in_pos_ = in_len_ - zs_.avail_in;
            if (rc == Z_STREAM_END) {
                member_open_ = false;
                if (gzip_) { inflateReset(&zs_); continue; }   // gzip.Reader reads concatenated members by default
                done_ = true;
                break;
            }
            if (rc != Z_OK && rc != Z_BUF_ERROR) { *why = zs_.msg ? zs_.msg : "corrupt input"; return -1; }
        }
        return (long)(want - zs_.avail_out);
    }

  private:
    z_stream zs_;
    bool gzip_, ok_ = false, done_ = false, member_open_ = false;
};

class ZstdStream : public ByteStream {
  public:
    explicit ZstdStream(FILE* fp) {
        fp_ = fp;
        ds_ = zstd()->createDStream();
        if (ds_) zstd()->initDStream(ds_);
    }
    ~ZstdStream() override {
        if (ds_) zstd()->freeDStream(ds_);
        if (fp_) fclose(fp_);
    }
    bool ok() const { return ds_ != nullptr; }

  protected:
    long fill(char* dst, size_t cap, std::string* why) override {
        ZOutBuf out{dst, cap, 0};
        while (out.pos == 0) {
            if (!more_input(why)) {
                if (!why->empty()) return -1;
                if (frame_open_) { *why = "unexpected EOF"; return -1; }
                return 0;
            }
            ZInBuf in{in_.data(), in_len_, in_pos_};
            const size_t rc = zstd()->decompressStream(ds_, &out, &in);
            in_pos_ = in.pos;
            if (zstd()->isError(rc)) { *why = zstd()->getErrorName(rc); return -1; }
            frame_open_ = rc != 0;
        }
        return (long)out.pos;
    }

  private:
    void* ds_ = nullptr;
    bool frame_open_ = false;
};

// LZW as Go's compress/lzw reads it with lzw.MSB and a literal width of 8 (stf.go:262: the 'l' files): codes of 9 to 12 bits packed
// most significant bit first, 256 = clear, 257 = end of data, dictionary entries from 258; the width grows when the next free
// entry reaches 1 << width.  The package is part of the Go standard library (source not in the reference tree); this is its
// documented algorithm (the GIF variant with MSB packing, not TIFF's early change), checked here by round trips only.
class LzwStream : public ByteStream {
  public:
    explicit LzwStream(FILE* fp) { fp_ = fp; reset(); }
    ~LzwStream() override { if (fp_) fclose(fp_); }

  protected:
    long fill(char* dst, size_t cap, std::string* why) override {
        size_t o = 0;
        // bytes of an expansion that did not fit the previous call
        while (pend_ > 0 && o < cap) dst[o++] = (char)stack_[--pend_];
        while (o < cap && !done_) {
            int code = read_code(why);
            if (code == -2) return -1;
            if (code == -1) { *why = "unexpected EOF"; return -1; }   // the stream must end with the end-of-data code
            if (code == kClear) { reset(); continue; }
            if (code == kEof) { done_ = true; break; }
            int n = 0;   // expansion of the code, last byte first, into stack_
            if (code < kClear) {
                stack_[n++] = (unsigned char)code;
                if (last_ != kInvalid) { suffix_[hi_] = (unsigned char)code; prefix_[hi_] = (uint16_t)last_; }
            } else if (code <= hi_) {
                int c = code;
                if (code == hi_ && last_ != kInvalid) {   // the entry being defined: last expansion + its own first byte
                    c = last_;
                    while (c >= kClear) c = prefix_[c];
                    stack_[n++] = (unsigned char)c;
                    c = last_;
                }
                while (c >= kClear) { stack_[n++] = suffix_[c]; c = prefix_[c]; }
                stack_[n++] = (unsigned char)c;
                if (last_ != kInvalid) { suffix_[hi_] = (unsigned char)c; prefix_[hi_] = (uint16_t)last_; }
            } else {
                *why = "lzw: invalid code";
                return -1;
            }
            last_ = code;
            hi_++;
            if (hi_ >= overflow_) {
                if (width_ == kMaxWidth) { last_ = kInvalid; hi_--; }
                else { width_++; overflow_ = 1 << width_; }
            }
            pend_ = n;
            while (pend_ > 0 && o < cap) dst[o++] = (char)stack_[--pend_];
        }
        return (long)o;
    }

  private:
    static constexpr int kClear = 256, kEof = 257, kMaxWidth = 12, kInvalid = 0xffff;
    void reset() { width_ = 9; hi_ = kEof; overflow_ = 1 << 9; last_ = kInvalid; }
    int read_code(std::string* why) {   // -1 = end of input, -2 = read error
        while (nbits_ < width_) {
            if (!more_input(why)) return why->empty() ? -1 : -2;
            bits_ |= (uint32_t)(unsigned char)in_[in_pos_++] << (24 - nbits_);
            nbits_ += 8;
        }
        const int code = (int)(bits_ >> (32 - width_));
        bits_ <<= width_;
        nbits_ -= width_;
        return code;
    }
    uint32_t bits_ = 0;
    int nbits_ = 0, width_ = 9, hi_ = kEof, overflow_ = 512, last_ = kInvalid, pend_ = 0;
    bool done_ = false;
    unsigned char suffix_[1 << 12];
    uint16_t prefix_[1 << 12];
    unsigned char stack_[(1 << 12) + 2];
};

// strings.Fields + strconv.Atoi on one coordinate line (stf.go:324-345): exactly three integers
const char* parseLine(const char* p, size_t n, long out[3], std::string* why) {
    const char* e = p + n;
    int nf = 0;
    while (p < e) {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) p++;
        if (p == e) break;
        const char* f0 = p;
        while (p < e && !(*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) p++;
        if (nf == 3) { *why = "Ill formated coordinates line in stf: Too many fields: " + std::string(e - n, n); return nullptr; }
        const char* q = f0;
        bool neg = false;
        if (q < p && (*q == '-' || *q == '+')) { neg = *q == '-'; q++; }
        long v = 0;
        bool ok = q < p;
        int digits = 0;
        for (; q < p; q++) {
            if (*q < '0' || *q > '9' || ++digits > 18) { ok = false; break; }
            v = v * 10 + (*q - '0');
        }
        if (!ok) {
            *why = "Can't parse coordinate " + std::to_string(nf) + " (" + std::string(f0, (size_t)(p - f0)) + "). Error: strconv.Atoi: parsing \"" +
                   std::string(f0, (size_t)(p - f0)) + "\": invalid syntax";
            return nullptr;
        }
        out[nf++] = neg ? -v : v;
    }
    if (nf < 3) { *why = "Ill formated coordinates line in stf: Too few fields: " + std::string(e - n, n); return nullptr; }
    return e;
}

}  // namespace

bool StfR::ZstdAvailable(std::string* why) {
    ZstdApi* a = zstd();
    if (!a->h && why) *why = a->why;
    return a->h != nullptr;
}

StfR::~StfR() { Close(); }

void StfR::Close() {
    if (!readable_) return;
    h_.reset();
    readable_ = false;
}

Error StfR::New(const std::string& name, std::unique_ptr<StfR>* out, std::map<std::string, std::string>* header) {
    std::unique_ptr<StfR> s(new StfR());
    s->filename_ = name;
    const Codec codec = codecOf(name);
    if (codec == Codec::Zstd && !zstd()->h) return stfError("Can't read header: " + zstd()->why, name, {"NewStfR"});
    FILE* fp = fopen(name.c_str(), "rb");
    if (!fp) return Error("open " + name + ": no such file or directory");   // os.Open's error goes back undecorated (stf.go:237-240)
    if (codec == Codec::Lzw) {
        s->h_.reset(new LzwStream(fp));
    } else if (codec == Codec::Zstd) {
        std::unique_ptr<ZstdStream> z(new ZstdStream(fp));
        if (!z->ok()) return stfError("Can't read header: zstd decoder could not be created", name, {"NewStfR"});
        s->h_ = std::move(z);
    } else {
        std::unique_ptr<ZlibStream> z(new ZlibStream(fp, codec == Codec::Gzip));
        if (!z->ok()) return stfError("Can't read header: inflate could not be initialised", name, {"NewStfR"});
        s->h_ = std::move(z);
    }
    std::map<std::string, std::string> m;
    for (;;) {
        const char* p = nullptr;
        size_t n = 0;
        std::string why;
        const int st = s->h_->line(&p, &n, &why);
        if (st != ByteStream::kLine) return stfError("Can't read header " + (st == ByteStream::kError ? why : std::string("EOF")), name, {"NewStfR"});
        const std::string str(p, n);
        if (str.find("**") != std::string::npos) {
            // strings.Fields(str)[1] -> strconv.Atoi (stf.go:289-301)
            size_t a = 0;
            std::vector<std::string> fields;
            while (a < str.size()) {
                while (a < str.size() && isspace((unsigned char)str[a])) a++;
                size_t b = a;
                while (b < str.size() && !isspace((unsigned char)str[b])) b++;
                if (b > a) fields.push_back(str.substr(a, b - a));
                a = b;
            }
            if (fields.size() < 2) return stfError("Can't read atom number from '" + str + "'", name, {"NewStfR"});
            char* endp = nullptr;
            const long nat = strtol(fields[1].c_str(), &endp, 10);
            if (fields[1].empty() || *endp != 0 || nat <= 0 || nat > (1L << 30))
                return stfError("Can't read atom number from '" + fields[1] + "': strconv.Atoi: parsing \"" + fields[1] + "\": invalid syntax", name, {"NewStfR"});
            s->natoms_ = (int)nat;
            break;
        }
        const size_t eq = str.find('=');
        if (eq == std::string::npos || str.find('=', eq + 1) != std::string::npos) return stfError("Malformed header", name, {"NewStfR"});   // stf.go:304-306
        m[str.substr(0, eq)] = str.substr(eq + 1);
    }
    // "prec" in the header never changes the precision (inverted error test, stf.go:311-318): p stays 100
    s->readable_ = true;
    if (header) *header = std::move(m);
    *out = std::move(s);
    return Error();
}

// One frame: natoms coordinate lines, then the line that starts with '*'.  ints (3 * natoms) and / or output receive it.
Error StfR::nextInto(int32_t* ints, v3::Matrix* output, std::vector<double>* box) {
    if (!readable_) return stfError(TrajUnIniRead, filename_, {"Next"});
    if (output && output->NVecs() < natoms_) throw v3::PanicMsg("Buffer v3.Matrix too small to hold trajectory frame");
    const double p100 = 100.0;
    std::string why;
    for (int i = 0; i < natoms_; i++) {
        const char* p = nullptr;
        size_t n = 0;
        const int st = h_->line(&p, &n, &why);
        if (st != ByteStream::kLine) {
            if (st != ByteStream::kError && i == 0) {   // the trajectory just ended (stf.go:361-366)
                Close();
                return chem::newLastFrameError(filename_, "Next");
            }
            return stfError(st == ByteStream::kError ? why : std::string("EOF"), filename_);
        }
        long v[3];
        if (!parseLine(p, n, v, &why)) return stfError(why, filename_);
        if (ints) {
            for (int k = 0; k < 3; k++) {
                if (v[k] > INT32_MAX || v[k] < INT32_MIN) return stfError("coordinate " + std::to_string(v[k]) + " does not fit the 32-bit device staging format", filename_);
                ints[3 * (size_t)i + k] = (int32_t)v[k];
            }
        }
        if (output)
            for (int k = 0; k < 3; k++) output->Set(i, k, (double)v[k] / p100);   // float64(f) / p (stf.go:341)
    }
    const char* p = nullptr;
    size_t n = 0;
    const int st = h_->line(&p, &n, &why);
    if (st != ByteStream::kLine) return stfError("Can't read the frame termination mark" + (st == ByteStream::kError ? why : std::string("EOF")), filename_, {"Next"});
    if (n == 0 || p[0] != '*') return stfError("Wrong number of atoms in frame", filename_, {"Next"});
    if (box && box->size() >= 9) {   // "* b0 ... b8" (stf.go:390-420); unreadable numbers zero the whole box
        std::vector<double> b;
        const std::string s(p + 1, n - 1);
        const char* c = s.c_str();
        for (;;) {
            char* endp = nullptr;
            const double val = strtod(c, &endp);
            if (endp == c) break;
            b.push_back(val);
            c = endp;
        }
        while (*c && isspace((unsigned char)*c)) c++;
        if (b.size() >= 9) {
            const bool bad = *c != 0 && b.size() == 9;
            for (int k = 0; k < 9; k++) (*box)[(size_t)k] = bad ? 0.0 : b[(size_t)k];
        }
    }
    return Error();
}

Error StfR::Next(v3::Matrix* output, std::vector<double>* box) { return nextInto(nullptr, output, box); }

Error StfR::NextConc(std::vector<v3::Matrix*>& frames) {
    if (!readable_) return stfError(TrajUnIniRead, filename_, {"NextConc"});
    for (auto* slot : frames)
        if (Error e = Next(slot)) return errDecorate(e, "NextConc");
    return Error();
}

Error StfR::NextRaw(float* dst, int* layout, double* scale) {
    *layout = GCB_I32_AOS;
    *scale = 100.0;
    return nextInto(reinterpret_cast<int32_t*>(dst), nullptr, nullptr);
}

// ---------------- writer ----------------
class ByteSink {
  public:
    virtual ~ByteSink() = default;
    virtual bool write(const char* p, size_t n) = 0;
    virtual bool finish() = 0;
};

namespace {

class ZlibSink : public ByteSink {
  public:
    ZlibSink(FILE* fp, bool gzip, int level) : fp_(fp) {
        memset(&zs_, 0, sizeof zs_);
        ok_ = deflateInit2(&zs_, level, Z_DEFLATED, gzip ? 16 + MAX_WBITS : -MAX_WBITS, 8, Z_DEFAULT_STRATEGY) == Z_OK;
    }
    ~ZlibSink() override { finish(); }
    bool ok() const { return ok_; }
    bool write(const char* p, size_t n) override { return pump(p, n, Z_NO_FLUSH); }
    bool finish() override {
        if (!fp_) return true;
        const bool good = ok_ && pump(nullptr, 0, Z_FINISH);
        if (ok_) deflateEnd(&zs_);
        ok_ = false;
        const bool closed = fclose(fp_) == 0;
        fp_ = nullptr;
        return good && closed;
    }

  private:
    bool pump(const char* p, size_t n, int flush) {
        if (!ok_) return false;
        zs_.next_in = (Bytef*)p;
        zs_.avail_in = (uInt)n;
        do {
            zs_.next_out = (Bytef*)out_;
            zs_.avail_out = sizeof out_;
            const int rc = deflate(&zs_, flush);
            if (rc == Z_STREAM_ERROR) return false;
            const size_t have = sizeof out_ - zs_.avail_out;
            if (have && fwrite(out_, 1, have, fp_) != have) return false;
            if (flush == Z_FINISH && rc == Z_STREAM_END) break;
        } while (zs_.avail_out == 0 || zs_.avail_in > 0 || flush == Z_FINISH);
        return true;
    }
    FILE* fp_;
    z_stream zs_;
    bool ok_ = false;
    char out_[1 << 16];
};

class ZstdSink : public ByteSink {
  public:
    explicit ZstdSink(FILE* fp) : fp_(fp) {
        cs_ = zstd()->createCStream();
        if (cs_ && zstd()->isError(zstd()->initCStream(cs_, 11))) { zstd()->freeCStream(cs_); cs_ = nullptr; }   // zstd.SpeedBestCompression ~ level 11
    }
    ~ZstdSink() override { finish(); }
    bool ok() const { return cs_ != nullptr; }
    bool write(const char* p, size_t n) override {
        if (!cs_) return false;
        ZInBuf in{p, n, 0};
        while (in.pos < in.size) {
            ZOutBuf out{out_, sizeof out_, 0};
            if (zstd()->isError(zstd()->compressStream(cs_, &out, &in))) return false;
            if (out.pos && fwrite(out_, 1, out.pos, fp_) != out.pos) return false;
        }
        return true;
    }
    bool finish() override {
        if (!fp_) return true;
        bool good = cs_ != nullptr;
        while (good) {
            ZOutBuf out{out_, sizeof out_, 0};
            const size_t rc = zstd()->endStream(cs_, &out);
            if (zstd()->isError(rc)) { good = false; break; }
            if (out.pos && fwrite(out_, 1, out.pos, fp_) != out.pos) { good = false; break; }
            if (rc == 0) break;
        }
        if (cs_) zstd()->freeCStream(cs_);
        cs_ = nullptr;
        const bool closed = fclose(fp_) == 0;
        fp_ = nullptr;
        return good && closed;
    }

  private:
    FILE* fp_;
    void* cs_ = nullptr;
    char out_[1 << 16];
};

// The matching encoder (lzw.NewWriter(a, lzw.MSB, 8), stf.go:143): a clear code first, the longest known string per code, the
// dictionary dropped with a clear code when entry 4095 would be assigned, the end-of-data code at Close.
class LzwSink : public ByteSink {
  public:
    explicit LzwSink(FILE* fp) : fp_(fp), table_(1 << 20, -1) {}
    ~LzwSink() override { finish(); }
    bool write(const char* p, size_t n) override {
        if (!fp_) return false;
        size_t i = 0;
        if (code_ < 0 && n > 0) {
            if (!emit(kClear)) return false;
            code_ = (unsigned char)p[i++];
        }
        for (; i < n; i++) {
            const int lit = (unsigned char)p[i];
            const int key = code_ << 8 | lit;   // code_ < 4096: 20 bits
            if (table_[(size_t)key] >= 0) { code_ = table_[(size_t)key]; continue; }
            if (!emit(code_)) return false;
            code_ = lit;
            if (!inc_hi()) continue;            // out of codes: the dictionary starts over
            table_[(size_t)key] = hi_;
            used_.push_back(key);
        }
        return true;
    }
    bool finish() override {
        if (!fp_) return true;
        bool good = true;
        if (code_ >= 0) { good = emit(code_); inc_hi(); }
        else good = emit(kClear);
        good = good && emit(kEof);
        if (nbits_ > 0) good = good && fputc((int)(bits_ >> 24), fp_) != EOF;
        const bool closed = fclose(fp_) == 0;
        fp_ = nullptr;
        return good && closed;
    }

  private:
    static constexpr int kClear = 256, kEof = 257, kMaxCode = 4095;
    bool emit(int code) {
        bits_ |= (uint32_t)code << (32 - width_ - nbits_);
        nbits_ += width_;
        while (nbits_ >= 8) {
            if (fputc((int)(bits_ >> 24), fp_) == EOF) return false;
            bits_ <<= 8;
            nbits_ -= 8;
        }
        return true;
    }
    bool inc_hi() {   // false: the code space ran out and a clear code was written
        hi_++;
        if (hi_ == overflow_) { width_++; overflow_ <<= 1; }
        if (hi_ == kMaxCode) {
            emit(kClear);
            width_ = 9; hi_ = kEof; overflow_ = 512;
            for (int k : used_) table_[(size_t)k] = -1;
            used_.clear();
            return false;
        }
        return true;
    }
    FILE* fp_;
    std::vector<int> table_;   // (prefix code << 8 | byte) -> code
    std::vector<int> used_;
    uint32_t bits_ = 0;
    int nbits_ = 0, width_ = 9, hi_ = kEof, overflow_ = 512, code_ = -1;
};

}  // namespace

StfW::~StfW() { Close(); }

void StfW::Close() {
    if (writeable_ && h_) h_->finish();
    h_.reset();
    writeable_ = false;
}

Error StfW::NewWriter(const std::string& name, int natoms, const std::map<std::string, std::string>& header, std::unique_ptr<StfW>* out, int level) {
    std::unique_ptr<StfW> s(new StfW());
    s->filename_ = name;
    const Codec codec = codecOf(name);
    // flate.NewWriter and gzip.NewWriterLevel refuse levels outside [-2, 9] -- including the reference's own default of 11
    // (stf.go:115, 131-135): with the default level only the zstd forms can be written there
    if (codec == Codec::Flate && (level < -2 || level > 9))
        return stfError("Can't read header flate: invalid compression level " + std::to_string(level) + ": want value in range [-2, 9]", name, {"NewWriter"});
    if (codec == Codec::Gzip && (level < -2 || level > 9))
        return stfError("Can't read header gzip: invalid compression level: " + std::to_string(level), name, {"NewWriter"});
    if (codec == Codec::Zstd && !zstd()->h) return stfError("Can't read header " + zstd()->why, name, {"NewWriter"});
    FILE* fp = fopen(name.c_str(), "wb");
    if (!fp) return Error("open " + name + ": cannot create");
    if (codec == Codec::Lzw) {
        s->h_.reset(new LzwSink(fp));
    } else if (codec == Codec::Zstd) {
        std::unique_ptr<ZstdSink> z(new ZstdSink(fp));
        if (!z->ok()) return stfError("Can't read header zstd encoder could not be created", name, {"NewWriter"});
        s->h_ = std::move(z);
    } else {
        const int zl = level == -2 ? 1 : level;   // flate.HuffmanOnly has no zlib level; any level decodes the same
        std::unique_ptr<ZlibSink> z(new ZlibSink(fp, codec == Codec::Gzip, zl));
        if (!z->ok()) return stfError("Can't read header deflate could not be initialised", name, {"NewWriter"});
        s->h_ = std::move(z);
    }
    s->natoms_ = natoms;
    s->writeable_ = true;
    std::string head;
    for (const auto& kv : header) head += kv.first + "=" + kv.second + "\n";   // Go iterates its map in random order; sorted here
    head += "** " + std::to_string(natoms) + "\n";
    if (!s->h_->write(head.data(), head.size())) return stfError("write failed", name, {"NewWriter"});
    *out = std::move(s);
    return Error();
}

Error StfW::WNext(const v3::Matrix* coord, const std::vector<double>* box) {
    if (!writeable_) return stfError(TrajUnIniWrite, filename_, {"WNext"});
    if (!coord) return stfError(NilCoordinates, filename_, {"WNext"});
    const int v = coord->NVecs();
    if (v != natoms_) return stfError(std::to_string(v) + " coordinates given, but " + std::to_string(natoms_) + " expected", filename_, {"WNext"});
    std::string text;
    text.reserve((size_t)v * 24 + 128);
    char line[96];
    for (int i = 0; i < v; i++) {
        // int(math.RoundToEven(v * p)) with p = 100 (stf.go:216-225); nearbyint rounds half to even in the default mode
        const long a = (long)std::nearbyint(coord->At(i, 0) * 100.0), b = (long)std::nearbyint(coord->At(i, 1) * 100.0),
                   c = (long)std::nearbyint(coord->At(i, 2) * 100.0);
        text.append(line, (size_t)snprintf(line, sizeof line, "%ld %ld %ld\n", a, b, c));
    }
    if (box && box->size() >= 9) {
        const std::vector<double>& b = *box;
        char bl[256];
        text.append(bl, (size_t)snprintf(bl, sizeof bl, "* %4.2f %4.2f %4.2f %4.2f %4.2f %4.2f %4.2f %4.2f %4.2f\n", b[0], b[1], b[2], b[3], b[4], b[5],
                                         b[6], b[7], b[8]));
    } else {
        text += "*\n";
    }
    if (!h_->write(text.data(), text.size())) return stfError("write failed", filename_, {"WNext"});
    return Error();
}

}  // namespace stf
}  // namespace gochem
