/*
 * xdrshim.c -- TEST INFRASTRUCTURE: a stand-in for libxdrfile on hosts where the real library is absent (the GPU box).
 *
 * The product binds the XTC codec with dlopen (GOCHEM_XDRFILE names the library), exactly the five entry points the
 * reference's cgo shim uses (/root/reference/traj/xtc/libgoxtc.c:40-104, xtc.go:76-121): read_xtc_natoms, xdrfile_open,
 * xdrfile_close, read_xtc, write_xtc.  This file implements those five over the XTC record layout WITHOUT the coordinate
 * compression: XDR (big-endian) magic 1995, natoms, step, time, box[9], natoms, then 3 * natoms raw floats -- which is what a
 * real XTC file holds when natoms <= 9; for larger frames real XTC compresses the block, this shim does not, so its files are
 * not interchangeable with libxdrfile's beyond 9 atoms.  It exists so that the chain
 *     XTCObj::NextRaw -> float32 AoS nanometres in pinned memory -> gcb_traj_upload_async(GCB_F32_AOS, scale 10) -> kernels
 * runs under `pytest -m gpu` on a box without the third-party codec.  Return codes follow libxdrfile: 0 ok, 11 end of file.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define XTC_MAGIC 1995
enum { exdrOK = 0, exdrHEADER = 1, exdrINT = 3, exdrFLOAT = 4, exdr3DX = 7, exdrMAGIC = 9, exdrENDOFFILE = 11, exdrFILENOTFOUND = 12 };

typedef struct XDRFILE {
    FILE* fp;
    int writing;
} XDRFILE;

static uint32_t bswap(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

static int put_u32(XDRFILE* x, uint32_t v) {
    v = bswap(v);
    return fwrite(&v, 4, 1, x->fp) == 1;
}
static int get_u32(XDRFILE* x, uint32_t* v) {
    if (fread(v, 4, 1, x->fp) != 1) return 0;
    *v = bswap(*v);
    return 1;
}
static int put_f32(XDRFILE* x, float f) { uint32_t v; memcpy(&v, &f, 4); return put_u32(x, v); }
static int get_f32(XDRFILE* x, float* f) { uint32_t v; if (!get_u32(x, &v)) return 0; memcpy(f, &v, 4); return 1; }

XDRFILE* xdrfile_open(const char* path, const char* mode) {
    FILE* fp = fopen(path, (mode && mode[0] == 'w') ? "wb" : (mode && mode[0] == 'a') ? "ab" : "rb");
    if (!fp) return NULL;
    XDRFILE* x = (XDRFILE*)calloc(1, sizeof *x);
    if (!x) { fclose(fp); return NULL; }
    x->fp = fp;
    x->writing = mode && mode[0] != 'r';
    return x;
}

int xdrfile_close(XDRFILE* x) {
    if (!x) return exdrFILENOTFOUND;
    int rc = fclose(x->fp);
    free(x);
    return rc == 0 ? exdrOK : exdrHEADER;
}

static int read_header(XDRFILE* x, int* natoms, int* step, float* time) {
    uint32_t magic, n, s;
    if (!get_u32(x, &magic)) return exdrENDOFFILE;
    if (magic != XTC_MAGIC) return exdrMAGIC;
    if (!get_u32(x, &n) || !get_u32(x, &s) || !get_f32(x, time)) return exdrHEADER;
    *natoms = (int)n;
    *step = (int)s;
    return exdrOK;
}

int read_xtc_natoms(char* fn, int* natoms) {
    XDRFILE* x = xdrfile_open(fn, "r");
    if (!x) return exdrFILENOTFOUND;
    int step;
    float time;
    int rc = read_header(x, natoms, &step, &time);
    xdrfile_close(x);
    return rc;
}

int read_xtc(XDRFILE* x, int natoms, int* step, float* time, float box[3][3], float (*coords)[3], float* prec) {
    int n = 0;
    int rc = read_header(x, &n, step, time);
    if (rc != exdrOK) return rc;
    if (n != natoms) return exdrHEADER;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            if (!get_f32(x, &box[i][j])) return exdrFLOAT;
    uint32_t lsize;
    if (!get_u32(x, &lsize) || (int)lsize != natoms) return exdr3DX;
    const size_t nv = 3 * (size_t)natoms;
    uint32_t* raw = (uint32_t*)coords;
    if (fread(raw, 4, nv, x->fp) != nv) return exdr3DX;
    for (size_t k = 0; k < nv; k++) raw[k] = bswap(raw[k]);
    if (prec) *prec = 1000.0f;
    return exdrOK;
}

int write_xtc(XDRFILE* x, int natoms, int step, float time, float box[3][3], float (*coords)[3], float prec) {
    (void)prec;
    if (!put_u32(x, XTC_MAGIC) || !put_u32(x, (uint32_t)natoms) || !put_u32(x, (uint32_t)step) || !put_f32(x, time)) return exdrHEADER;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            if (!put_f32(x, box[i][j])) return exdrFLOAT;
    if (!put_u32(x, (uint32_t)natoms)) return exdrINT;
    const float* c = &coords[0][0];
    for (size_t k = 0; k < 3 * (size_t)natoms; k++)
        if (!put_f32(x, c[k])) return exdr3DX;
    return exdrOK;
}
