/*
 * gochem_oracle.c -- CPU restatement (fp64) of goChem's superposition / RMSD / LOVO path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gochem_b200/ may include, link or
 * call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline -- never as the product path.
 *
 * PARITY UNPINNED.  The reference (rmera/gochem, Go) cannot be compiled here (no
 * Go toolchain) and its tests for this path hold no golden values
 * (gochem_test.go:451-495 is disabled by name and only prints;
 * align/lovo_test.go:34-59 only prints; test/ fixtures are git-ignored).  The
 * numerically decisive routines live in gonum.org/v1/gonum v0.15.1 (go.mod:10),
 * whose source is not under /root/reference: mat.SVD (LAPACK Dgesvd), Dense.Mul
 * (Dgemm), mat.Norm(.,2) (Dlange 'F').  This file restates their published
 * algorithms' *results*: the 3x3 SVD is a one-sided Jacobi (converged to full
 * fp64 precision); sums run in the same sequential row order gonum's Dgemm uses.
 * It is cross-checked in tests/ against oracle_np.py (numpy + LAPACK gesvd via
 * scipy) and against closed-form invariants.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_SHAPE -1      /* "Ill-formed matrices" / "Ill formed coordinates" */
#define ORC_ERR_INDEXES -2    /* "Indexes don't match molecules" */
#define ORC_ERR_NILDEREF -3   /* reference would nil-dereference ctest (handy.go:186-188, geometric.go:269-271) */
#define ORC_ERR_SVD -4        /* mat.SVD failed (geometric.go:155-157) */
#define ORC_ERR_ALLOC -5
#define ORC_ERR_NOTERM -6     /* mobileLimit would never terminate (Len < tolerance) */

/* ------------------------------------------------------------------ */
/* small dense helpers (matrixhelp.go)                                  */
/* ------------------------------------------------------------------ */

/* det of a 3x3, same expansion as matrixhelp.go:136-142 (row-major a[3*i+j]) */
static double det3(const double *a) {
    return a[0] * (a[4] * a[8] - a[7] * a[5]) - a[3] * (a[1] * a[8] - a[7] * a[2]) +
           a[6] * (a[1] * a[5] - a[4] * a[2]);
}

/* C(3x3) = A(3x3) * B(3x3), Dgemm order: c[i][:] += a[i][l]*b[l][:] for l ascending */
static void mul33(const double *a, const double *b, double *c) {
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int l = 0; l < 3; l++) s += a[3 * i + l] * b[3 * l + j];
            t[3 * i + j] = s;
        }
    memcpy(c, t, sizeof t);
}

static void tr33(const double *a, double *t) {
    double r[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[3 * j + i] = a[3 * i + j];
    memcpy(t, r, sizeof r);
}

/* Frobenius norm the way LAPACK Dlange('F') computes it (scaled sum of squares,
 * dlassq): gonum mat.Norm(m, 2) on a Dense routes there (v3/gonum.go:268-275). */
static double frob(const double *x, long n) {
    double scale = 0.0, ssq = 1.0;
    for (long i = 0; i < n; i++) {
        double v = x[i];
        if (v == 0.0) continue;
        double a = fabs(v);
        if (scale < a) {
            ssq = 1.0 + ssq * (scale / a) * (scale / a);
            scale = a;
        } else {
            ssq += (a / scale) * (a / scale);
        }
    }
    return scale * sqrt(ssq);
}

/* ------------------------------------------------------------------ */
/* 3x3 SVD: stands in for gonum mat.SVD.Factorize(SVDFull)/UTo/VTo       */
/* (geometric.go:154-165).  One-sided (Hestenes) Jacobi; singular values  */
/* sorted descending as LAPACK returns them; U, V full orthogonal.        */
/* h = u * diag(s) * v^T, all row-major.                                  */
/* ------------------------------------------------------------------ */
static void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

int orc_svd3(const double *h, double *u, double *s, double *v) {
    double a[9], vv[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(a, h, sizeof a);
    for (int i = 0; i < 9; i++)
        if (!isfinite(a[i])) return ORC_ERR_SVD;
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; i++) {
                    alpha += a[3 * i + p] * a[3 * i + p];
                    beta += a[3 * i + q] * a[3 * i + q];
                    gamma += a[3 * i + p] * a[3 * i + q];
                }
                if (gamma == 0.0 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
                rotated = 1;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < 3; i++) {
                    double ap = a[3 * i + p], aq = a[3 * i + q];
                    a[3 * i + p] = c * ap - sn * aq;
                    a[3 * i + q] = sn * ap + c * aq;
                    double vp = vv[3 * i + p], vq = vv[3 * i + q];
                    vv[3 * i + p] = c * vp - sn * vq;
                    vv[3 * i + q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double sv[3];
    for (int j = 0; j < 3; j++)
        sv[j] = sqrt(a[j] * a[j] + a[3 + j] * a[3 + j] + a[6 + j] * a[6 + j]);
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (sv[ord[j]] < sv[ord[j + 1]]) {
                int t = ord[j];
                ord[j] = ord[j + 1];
                ord[j + 1] = t;
            }
    double uc[3][3];
    double tiny = sv[ord[0]] * 1e-300;
    for (int k = 0; k < 3; k++) {
        int j = ord[k];
        s[k] = sv[j];
        for (int i = 0; i < 3; i++) {
            v[3 * i + k] = vv[3 * i + j];
            uc[k][i] = (sv[j] > tiny && sv[j] > 0.0) ? a[3 * i + j] / sv[j] : 0.0;
        }
    }
    /* complete U for rank-deficient input so it stays orthogonal */
    if (!(s[0] > 0.0)) {
        double e[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        memcpy(u, e, sizeof e);
        return ORC_OK;
    }
    if (!(s[1] > tiny)) {
        /* pick any unit vector orthogonal to uc[0] */
        double t[3] = {0, 0, 0};
        int m = 0;
        for (int i = 1; i < 3; i++)
            if (fabs(uc[0][i]) < fabs(uc[0][m])) m = i;
        t[m] = 1.0;
        cross3(uc[0], t, uc[1]);
        double n = sqrt(uc[1][0] * uc[1][0] + uc[1][1] * uc[1][1] + uc[1][2] * uc[1][2]);
        for (int i = 0; i < 3; i++) uc[1][i] /= n;
    }
    if (!(s[2] > tiny)) {
        cross3(uc[0], uc[1], uc[2]);
        double n = sqrt(uc[2][0] * uc[2][0] + uc[2][1] * uc[2][1] + uc[2][2] * uc[2][2]);
        for (int i = 0; i < 3; i++) uc[2][i] /= n;
    }
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) u[3 * i + k] = uc[k][i];
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* chem.MassCenter with unit masses (geometric.go:670-703).              */
/* centroid = (1/sum(mass)) * (ones(1,N) * ref): Dgemm accumulates rows   */
/* in ascending order; Scale multiplies by the reciprocal (:690).         */
/* out (may alias nothing) = in - centroid (SubVec, :693).                */
/* ------------------------------------------------------------------ */
static void mass_center(const double *in, int n, double *out, double *cen) {
    double s0 = 0, s1 = 0, s2 = 0, msum = 0;
    for (int i = 0; i < n; i++) {
        s0 += in[3 * i] * 1.0;
        s1 += in[3 * i + 1] * 1.0;
        s2 += in[3 * i + 2] * 1.0;
        msum += 1.0;
    }
    double inv = 1.0 / msum;
    cen[0] = s0 * inv;
    cen[1] = s1 * inv;
    cen[2] = s2 * inv;
    for (int i = 0; i < n; i++) {
        out[3 * i] = in[3 * i] - cen[0];
        out[3 * i + 1] = in[3 * i + 1] - cen[1];
        out[3 * i + 2] = in[3 * i + 2] - cen[2];
    }
}

static __thread double *t_scratch[4];
static __thread size_t t_scratch_n[4];
static double *scratch_get(int slot, size_t n) {
    if (t_scratch_n[slot] < n) {
        free(t_scratch[slot]);
        t_scratch[slot] = (double *)malloc(sizeof(double) * n);
        t_scratch_n[slot] = t_scratch[slot] ? n : 0;
    }
    return t_scratch[slot];
}

/* ------------------------------------------------------------------ */
/* chem.RotatorTranslatorToSuper (geometric.go:127-208).                 */
/* test, templa: n x 3 row-major.  Outputs: transformed (n x 3, may be    */
/* NULL), rot (3x3, applied on the RIGHT of row vectors), trans1 = -cen   */
/* (test), trans2 = Translation.                                          */
/* as_written != 0 also performs the reference's O(n^2) work              */
/* (gnEye(n), gnMul(j,jT), templa^T * I_n; :145-149) so its cost can be   */
/* timed; the numbers it produces are identical (multiplying by the       */
/* identity is exact).                                                    */
/* ------------------------------------------------------------------ */
int orc_rotator_translator(const double *test, const double *templa, int n, double *transformed,
                           double *rot, double *trans1, double *trans2, int as_written) {
    if (n <= 0) return ORC_ERR_SHAPE; /* :130-132 */
    /* per-thread scratch instead of the reference's per-call allocations (the Go allocator is
     * thread-cached; glibc would mmap/munmap 2 x 24n bytes per frame and serialise the workers) */
    double *ctest = scratch_get(0, 3 * (size_t)n);
    double *ctempla = scratch_get(1, 3 * (size_t)n);
    if (!ctest || !ctempla) return ORC_ERR_ALLOC;
    double distest[3], distempla[3];
    double scal = 1.0 / (double)n;           /* :133-134 */
    mass_center(test, n, ctest, distest);     /* :136 */
    mass_center(templa, n, ctempla, distempla); /* :140 */

    double maux[9]; /* aux2 * ctest, aux2 = ctempla^T * I  (:149-152): maux[i][j] = sum_l ctempla[l][i]*ctest[l][j] */
    if (as_written) {
        size_t nn = (size_t)n * (size_t)n;
        double *mid = (double *)calloc(nn, sizeof(double));   /* gnEye(n) :145 */
        double *sjp = (double *)malloc(nn * sizeof(double));  /* gnMul(j, jT) :147, scaled :148, never used */
        double *aux2 = (double *)calloc(3 * (size_t)n, sizeof(double));
        if (!mid || !sjp || !aux2) {
            free(mid);
            free(sjp);
            free(aux2);
            return ORC_ERR_ALLOC;
        }
        for (int i = 0; i < n; i++) mid[(size_t)i * n + i] = 1.0;
        for (size_t i = 0; i < nn; i++) sjp[i] = 1.0 * 1.0;
        for (size_t i = 0; i < nn; i++) sjp[i] = scal * sjp[i];
        /* aux2 (3 x n) = ctempla^T (3 x n) * mid (n x n): Dgemm skips zero a[i][l] only; the
         * inner axpy over the n columns of mid is the O(n^2) cost. */
        for (int i = 0; i < 3; i++)
            for (int l = 0; l < n; l++) {
                double t = ctempla[3 * l + i];
                if (t == 0.0) continue;
                const double *brow = mid + (size_t)l * n;
                double *crow = aux2 + (size_t)i * n;
                for (int j = 0; j < n; j++) crow[j] += t * brow[j];
            }
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) maux[3 * i + j] = 0.0;
        for (int i = 0; i < 3; i++)
            for (int l = 0; l < n; l++) {
                double t = aux2[(size_t)i * n + l];
                for (int j = 0; j < 3; j++) maux[3 * i + j] += t * ctest[3 * l + j];
            }
        volatile double sink = sjp[nn - 1];
        (void)sink;
        free(mid);
        free(sjp);
        free(aux2);
    } else {
        for (int i = 0; i < 9; i++) maux[i] = 0.0;
        for (int i = 0; i < 3; i++)
            for (int l = 0; l < n; l++) {
                double t = ctempla[3 * l + i];
                for (int j = 0; j < 3; j++) maux[3 * i + j] += t * ctest[3 * l + j];
            }
    }
    double h[9];
    tr33(maux, h); /* Maux = Maux.TrRet() :153  ->  h = ctest^T * ctempla */

    double u[9], v[9], s[3];
    if (orc_svd3(h, u, s, v) != ORC_OK) { /* :154-157 */
        return ORC_ERR_SVD;
    }
    for (int i = 0; i < 9; i++) { /* U.Scale(-1), V.Scale(-1) :174-175 */
        u[i] = -u[i];
        v[i] = -v[i];
    }
    double ut[9], r[9];
    tr33(u, ut);
    mul33(v, ut, r); /* Rotation.Mul(V, gnT(U)) :181 */
    tr33(r, r);      /* TrRet :182 */
    if (det3(r) < 0) { /* :184-188 */
        double rh[9] = {1, 0, 0, 0, 1, 0, 0, 0, -1};
        mul33(v, rh, r);
        mul33(r, ut, r);
        tr33(r, r);
    }
    /* Translation = (scal * ones(1,n)) * (ctempla - ctest*R) + distempla  :191-200 */
    double t0 = 0, t1 = 0, t2 = 0;
    for (int l = 0; l < n; l++) {
        const double *c = ctest + 3 * l;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        sx += c[0] * r[0]; sx += c[1] * r[3]; sx += c[2] * r[6];
        sy += c[0] * r[1]; sy += c[1] * r[4]; sy += c[2] * r[7];
        sz += c[0] * r[2]; sz += c[1] * r[5]; sz += c[2] * r[8];
        t0 += scal * (ctempla[3 * l] - sx);
        t1 += scal * (ctempla[3 * l + 1] - sy);
        t2 += scal * (ctempla[3 * l + 2] - sz);
    }
    trans2[0] = t0 + distempla[0];
    trans2[1] = t1 + distempla[1];
    trans2[2] = t2 + distempla[2];
    if (transformed) { /* :202-204 */
        for (int l = 0; l < n; l++) {
            const double *c = ctest + 3 * l;
            double sx = 0.0, sy = 0.0, sz = 0.0;
            sx += c[0] * r[0]; sx += c[1] * r[3]; sx += c[2] * r[6];
            sy += c[0] * r[1]; sy += c[1] * r[4]; sy += c[2] * r[7];
            sz += c[0] * r[2]; sz += c[1] * r[5]; sz += c[2] * r[8];
            transformed[3 * l] = sx + trans2[0];
            transformed[3 * l + 1] = sy + trans2[1];
            transformed[3 * l + 2] = sz + trans2[2];
        }
    }
    memcpy(rot, r, sizeof r);
    trans1[0] = -1.0 * distest[0]; /* :206 */
    trans1[1] = -1.0 * distest[1];
    trans1[2] = -1.0 * distest[2];
    return ORC_OK;
}

/* v3.Matrix.SomeVecs (v3/gocoords.go:155-166) */
static int some_vecs(const double *a, int na, const int *clist, int n, double *f) {
    if (na < n) return ORC_ERR_SHAPE;
    for (int k = 0; k < n; k++) {
        int v = clist[k];
        if (v < 0 || v >= na) return ORC_ERR_SHAPE; /* gonum At() panics */
        f[3 * k] = a[3 * v];
        f[3 * k + 1] = a[3 * v + 1];
        f[3 * k + 2] = a[3 * v + 2];
    }
    return ORC_OK;
}

/* Index dispatch shared by Super (handy.go:178-201), MemRMSD (geometric.go:261-283)
 * and MemPerAtomRMSD (geometric.go:295-313).  nlists = how many index slices the Go
 * caller passed (0, 1 or 2).  On success *pct / *pcm point at n x 3 data; *owned_*
 * tell the caller to free them. */
static int dispatch(const double *test, int ntest, const double *templa, int ntempla,
                    const int *idx0, int n0, const int *idx1, int n1, int nlists,
                    const double **pct, const double **pcm, int *nct, int *ncm,
                    double **own_t, double **own_m) {
    *own_t = *own_m = NULL;
    if (nlists == 0 || idx0 == NULL || n0 == 0) {
        *pct = test; *nct = ntest;
        *pcm = templa; *ncm = ntempla;
        return ORC_OK;
    }
    if (nlists == 1) {
        if (ntest > n0) {
            *own_t = (double *)malloc(sizeof(double) * 3 * (size_t)n0);
            if (!*own_t) return ORC_ERR_ALLOC;
            int e = some_vecs(test, ntest, idx0, n0, *own_t);
            if (e) return e;
            *pct = *own_t; *nct = n0;
            *pcm = templa; *ncm = ntempla;
            return ORC_OK;
        } else if (ntempla > n0) {
            /* ctest is left nil in the reference -> nil dereference at the NVecs() check */
            return ORC_ERR_NILDEREF;
        }
        return ORC_ERR_INDEXES;
    }
    *own_t = (double *)malloc(sizeof(double) * 3 * (size_t)(n0 > 0 ? n0 : 1));
    *own_m = (double *)malloc(sizeof(double) * 3 * (size_t)(n1 > 0 ? n1 : 1));
    if (!*own_t || !*own_m) return ORC_ERR_ALLOC;
    int e = some_vecs(test, ntest, idx0, n0, *own_t);
    if (e) return e;
    e = some_vecs(templa, ntempla, idx1, n1, *own_m);
    if (e) return e;
    *pct = *own_t; *nct = n0;
    *pcm = *own_m; *ncm = n1;
    return ORC_OK;
}

/* chem.Super (handy.go:175-214): test is modified in place for ALL ntest atoms:
 * test = ((test + trans1) * R) + trans2  (:207-211).  Optionally reports R, trans1, trans2. */
int orc_super(double *test, int ntest, const double *templa, int ntempla, const int *idx0, int n0,
              const int *idx1, int n1, int nlists, int as_written, double *rot_out,
              double *trans1_out, double *trans2_out) {
    const double *ct, *cm;
    int nct, ncm;
    double *ot, *om;
    int e = dispatch(test, ntest, templa, ntempla, idx0, n0, idx1, n1, nlists, &ct, &cm, &nct, &ncm,
                     &ot, &om);
    if (e == ORC_OK && nct != ncm) e = ORC_ERR_SHAPE; /* :199-201 */
    double r[9], t1[3], t2[3];
    if (e == ORC_OK) e = orc_rotator_translator(ct, cm, nct, NULL, r, t1, t2, as_written);
    free(ot);
    free(om);
    if (e != ORC_OK) return e;
    for (int i = 0; i < ntest; i++) {
        double a0 = test[3 * i] + t1[0], a1 = test[3 * i + 1] + t1[1], a2 = test[3 * i + 2] + t1[2];
        double sx = 0.0, sy = 0.0, sz = 0.0;
        sx += a0 * r[0]; sx += a1 * r[3]; sx += a2 * r[6];
        sy += a0 * r[1]; sy += a1 * r[4]; sy += a2 * r[7];
        sz += a0 * r[2]; sz += a1 * r[5]; sz += a2 * r[8];
        test[3 * i] = sx + t2[0];
        test[3 * i + 1] = sy + t2[1];
        test[3 * i + 2] = sz + t2[2];
    }
    if (rot_out) memcpy(rot_out, r, sizeof r);
    if (trans1_out) memcpy(trans1_out, t1, sizeof t1);
    if (trans2_out) memcpy(trans2_out, t2, sizeof t2);
    return ORC_OK;
}

/* chem.RMSD / chem.MemRMSD (geometric.go:232-243, 258-288): no superposition;
 * ||ctest - ctempla||_F / sqrt(n).  Returns the status; *out = -1 on error like the reference. */
int orc_mem_rmsd(const double *test, int ntest, const double *templa, int ntempla, const int *idx0,
                 int n0, const int *idx1, int n1, int nlists, double *out) {
    const double *ct, *cm;
    int nct, ncm;
    double *ot, *om;
    *out = -1.0;
    int e = dispatch(test, ntest, templa, ntempla, idx0, n0, idx1, n1, nlists, &ct, &cm, &nct, &ncm,
                     &ot, &om);
    if (e == ORC_OK && nct != ncm) e = ORC_ERR_SHAPE;
    if (e == ORC_OK) {
        double *tmp = (double *)malloc(sizeof(double) * 3 * (size_t)(nct > 0 ? nct : 1));
        if (!tmp) e = ORC_ERR_ALLOC;
        else {
            for (long i = 0; i < 3L * nct; i++) tmp[i] = ct[i] - cm[i]; /* :284 */
            *out = frob(tmp, 3L * nct) / sqrt((double)nct);           /* :285-286 */
            free(tmp);
        }
    }
    free(ot);
    free(om);
    return e;
}

/* chem.MemPerAtomRMSD with no index lists (geometric.go:294-321): d_i = ||test_i - templa_i||_2 */
int orc_per_atom_rmsd(const double *test, const double *templa, int n, double *out) {
    for (int i = 0; i < n; i++) {
        double d[3] = {test[3 * i] - templa[3 * i], test[3 * i + 1] - templa[3 * i + 1],
                       test[3 * i + 2] - templa[3 * i + 2]};
        out[i] = frob(d, 3); /* :317-318 */
    }
    return ORC_OK;
}

/* chem.rMSD (geometric.go:326-364): naive cross-check: sqrt(sum_i ||templa_i - test_i||^2 / n) */
double orc_rmsd_naive(const double *test, const double *templa, int n) {
    double acc = 0.0;
    for (int i = 0; i < n; i++) {
        double d[3] = {templa[3 * i] - test[3 * i], templa[3 * i + 1] - test[3 * i + 1],
                       templa[3 * i + 2] - test[3 * i + 2]};
        double nr = frob(d, 3);
        acc += pow(nr, 2);
    }
    return sqrt(acc / (double)n);
}

/* ------------------------------------------------------------------ */
/* align.RMSDTraj (align/lovo.go:286-353) over an in-memory trajectory   */
/* (frames: F x N x 3, what a file reader would hand out each pass).      */
/* chunksize = cpus; a chunk that cannot be filled is dropped (:323-326); */
/* begin/skip follow the reference's no-`continue` logic (:316-322);      */
/* the per-atom sums accumulate sequentially in frame order (:332-336).   */
/* nthreads workers process the frames of a chunk concurrently (the       */
/* goroutine-per-frame model, :328-331); the numbers do not depend on it. */
/* aligned_out (optional, frames_read x N x 3) receives what writeTraj    */
/* would write (:337-339).  per_frame_rmsd (optional, length F) receives  */
/* the fit-subset RMSD of each processed frame (bench convenience, not in */
/* the reference).                                                        */
/* ------------------------------------------------------------------ */
typedef struct {
    const double *frames;
    long F;
    int N;
    const double *ref;
    const int *idx;
    int nidx;
    int chunksize;
    int nthreads;
    int as_written;
    double *work;   /* chunksize x N x 3 frame copies */
    double *dev;    /* chunksize x N per-atom distances */
    long chunk_first_frame;
    int err;
    pthread_barrier_t bar;
    int stop;
} traj_ctx;

typedef struct {
    traj_ctx *c;
    int tid;
} traj_arg;

static void concproc(traj_ctx *c, int slot) { /* align/lovo.go:265-281 */
    size_t fsz = 3 * (size_t)c->N;
    double *cf = c->work + (size_t)slot * fsz;
    memcpy(cf, c->frames + (size_t)(c->chunk_first_frame + slot) * fsz, fsz * sizeof(double));
    int e = orc_super(cf, c->N, c->ref, c->N, c->idx, c->nidx, c->idx, c->nidx, 2, c->as_written, NULL,
                      NULL, NULL);
    if (e) {
        c->err = e;
        return;
    }
    orc_per_atom_rmsd(cf, c->ref, c->N, c->dev + (size_t)slot * c->N);
}

static void *traj_worker(void *p) {
    traj_arg *a = (traj_arg *)p;
    traj_ctx *c = a->c;
    for (;;) {
        pthread_barrier_wait(&c->bar);
        if (c->stop) break;
        for (int s = a->tid; s < c->chunksize; s += c->nthreads) concproc(c, s);
        pthread_barrier_wait(&c->bar);
    }
    return NULL;
}

int orc_rmsd_traj(const double *frames, long F, int N, const double *ref, const int *idx, int nidx,
                  int cpus, int begin_opt, int skip_opt, int nthreads, int as_written,
                  double *out_rmsd, long *frames_read, double *aligned_out) {
    if (cpus <= 0 || N <= 0) return ORC_ERR_SHAPE;
    if (nthreads <= 0) nthreads = 1;
    if (nthreads > cpus) nthreads = cpus;
    traj_ctx c;
    memset(&c, 0, sizeof c);
    c.frames = frames; c.F = F; c.N = N; c.ref = ref; c.idx = idx; c.nidx = nidx;
    c.chunksize = cpus; c.nthreads = nthreads; c.as_written = as_written;
    c.work = (double *)malloc(sizeof(double) * 3 * (size_t)N * cpus);
    c.dev = (double *)malloc(sizeof(double) * (size_t)N * cpus);
    if (!c.work || !c.dev) {
        free(c.work);
        free(c.dev);
        return ORC_ERR_ALLOC;
    }
    for (int i = 0; i < N; i++) out_rmsd[i] = 0.0;
    pthread_t *th = NULL;
    traj_arg *args = NULL;
    if (nthreads > 1) {
        pthread_barrier_init(&c.bar, NULL, (unsigned)nthreads + 1);
        th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
        args = (traj_arg *)malloc(sizeof(traj_arg) * nthreads);
        for (int t = 0; t < nthreads; t++) {
            args[t].c = &c;
            args[t].tid = t;
            pthread_create(&th[t], NULL, traj_worker, &args[t]);
        }
    }
    int chunksize = cpus;
    int begin = begin_opt / chunksize; /* :306 */
    int skip = skip_opt / chunksize;   /* :307 */
    long cursor = 0;                   /* next unread frame of the source */
    long chunksread = 0;
    for (long read = 0;; read++) {
        if (read < begin || read % (skip + 1) != 0) { /* :316-322 discard one chunk (nil slots) */
            /* Molecule.NextConc / dcd.NextConc with all-nil slots advance the cursor without an EOF
             * check per slot for Molecule (chem.go:733-737); file readers hit EOF.  We model the file
             * readers: running past the end is EOF. */
            if (cursor + chunksize > F) break;
            cursor += chunksize;
        }
        if (cursor + chunksize > F) break; /* EOF inside / before the chunk: chunk dropped :323-326 */
        c.chunk_first_frame = cursor;
        cursor += chunksize;
        chunksread++;
        if (nthreads > 1) {
            pthread_barrier_wait(&c.bar);
            pthread_barrier_wait(&c.bar);
        } else {
            for (int s = 0; s < chunksize; s++) concproc(&c, s);
        }
        if (c.err) break;
        for (int s = 0; s < chunksize; s++) { /* :332-340 in slot order */
            const double *d = c.dev + (size_t)s * N;
            for (int i = 0; i < N; i++) out_rmsd[i] += d[i];
            if (aligned_out)
                memcpy(aligned_out + ((size_t)(chunksread - 1) * chunksize + s) * 3 * (size_t)N,
                       c.work + (size_t)s * 3 * N, sizeof(double) * 3 * (size_t)N);
        }
    }
    if (nthreads > 1) {
        c.stop = 1;
        pthread_barrier_wait(&c.bar);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
        pthread_barrier_destroy(&c.bar);
        free(th);
        free(args);
    }
    free(c.work);
    free(c.dev);
    if (c.err) return c.err;
    double totalframes = (double)(chunksread * chunksize); /* :348 */
    for (int i = 0; i < N; i++) out_rmsd[i] = out_rmsd[i] / totalframes; /* :349-351 (0/0 = NaN if nothing read) */
    *frames_read = (long)totalframes;
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* Per-frame Super + RMSD over a trajectory (the chem.Super + chem.RMSD  */
/* user loop, gochem_test.go:471-473): frames superposed on ref using the */
/* index list for the fit (both sides), RMSD over the same subset after   */
/* the fit.  Used as the CPU baseline for BASELINE.json configs 1-2.      */
/* frames are modified in place when write_back != 0.                     */
/* ------------------------------------------------------------------ */
typedef struct {
    double *frames;
    long F;
    int N;
    const double *ref;
    const int *idx;
    int nidx;
    int write_back, as_written, nthreads, tid;
    double *rmsd, *rot, *t1, *t2;
    int err;
} sr_arg;

static void *sr_worker(void *p) {
    sr_arg *a = (sr_arg *)p;
    size_t fsz = 3 * (size_t)a->N;
    double *buf = a->write_back ? NULL : (double *)malloc(fsz * sizeof(double));
    for (long f = a->tid; f < a->F; f += a->nthreads) {
        double *fr = a->frames + (size_t)f * fsz;
        double *w = fr;
        if (!a->write_back) {
            memcpy(buf, fr, fsz * sizeof(double));
            w = buf;
        }
        int nl = (a->idx && a->nidx > 0) ? 2 : 0;
        int e = orc_super(w, a->N, a->ref, a->N, a->idx, a->nidx, a->idx, a->nidx, nl, a->as_written,
                          a->rot ? a->rot + 9 * f : NULL, a->t1 ? a->t1 + 3 * f : NULL,
                          a->t2 ? a->t2 + 3 * f : NULL);
        if (e) {
            a->err = e;
            break;
        }
        double r;
        e = orc_mem_rmsd(w, a->N, a->ref, a->N, a->idx, a->nidx, a->idx, a->nidx, nl, &r);
        if (e) {
            a->err = e;
            break;
        }
        a->rmsd[f] = r;
    }
    free(buf);
    return NULL;
}

int orc_super_rmsd_traj(double *frames, long F, int N, const double *ref, const int *idx, int nidx,
                        int write_back, int as_written, int nthreads, double *rmsd, double *rot,
                        double *t1, double *t2) {
    if (nthreads <= 0) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    sr_arg *args = (sr_arg *)malloc(sizeof(sr_arg) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        sr_arg a = {frames, F, N, ref, idx, nidx, write_back, as_written, nthreads, t, rmsd, rot, t1, t2, 0};
        args[t] = a;
        if (nthreads > 1) pthread_create(&th[t], NULL, sr_worker, &args[t]);
    }
    if (nthreads == 1) sr_worker(&args[0]);
    int err = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        if (args[t].err) err = args[t].err;
    }
    free(th);
    free(args);
    return err;
}

/* ------------------------------------------------------------------ */
/* LOVO bookkeeping (align/lovo.go)                                      */
/* ------------------------------------------------------------------ */

/* sort.Stable over (molIDs, rMSDs) pairs by rmsd ascending (lovo.go:418-432) or by id.
 * Insertion-merge sort: any stable sort yields the same permutation. */
static void stable_sort_pairs(int *ids, double *vals, int n, int by_rmsd) {
    int *ti = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    double *tv = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                /* take from the right run only if strictly less (stability) */
                int right_less = by_rmsd ? (vals[j] < vals[i]) : (ids[j] < ids[i]);
                if (right_less) { ti[k] = ids[j]; tv[k++] = vals[j++]; }
                else { ti[k] = ids[i]; tv[k++] = vals[i++]; }
            }
            while (i < mid) { ti[k] = ids[i]; tv[k++] = vals[i++]; }
            while (j < hi) { ti[k] = ids[j]; tv[k++] = vals[j++]; }
        }
        memcpy(ids, ti, sizeof(int) * (size_t)n);
        memcpy(vals, tv, sizeof(double) * (size_t)n);
    }
    free(ti);
    free(tv);
}

/* align.mobileLimit (lovo.go:175-188), quirks kept: n = k - 1 at the first value >= limit;
 * limit *= 1.1 after every scan including the successful one; tolerance <= 0 -> (0, limit). */
int orc_mobile_limit(double limit, int tolerance, const double *sorted_rmsd, int len, int *n_out,
                     double *limit_out) {
    int n = 0;
    int guard = 0;
    while (n < tolerance) {
        n = 0;
        for (; n < len; n++) {
            if (sorted_rmsd[n] >= limit) {
                n--;
                break;
            }
        }
        limit *= 1.1;
        if (++guard > 100000 || !isfinite(limit)) return ORC_ERR_NOTERM;
    }
    *n_out = n;
    *limit_out = limit;
    return ORC_OK;
}

static int is_in(int v, const int *c, int n) { /* lovo.go:474-484 */
    for (int i = 0; i < n; i++)
        if (c[i] == v) return 1;
    return 0;
}

static int same_elements(const int *a, const int *b, int n) { /* lovo.go:446-457 (equal lengths) */
    for (int i = 0; i < n; i++)
        if (!is_in(a[i], b, n)) return 0;
    return 1;
}

static int disagreement(const int *a, const int *b, int n) { /* lovo.go:459-471 */
    int d = 0;
    for (int i = 0; i < n; i++)
        if (!is_in(a[i], b, n)) d++;
    return d;
}

static int approx_eq(double i, double j, double eps) { /* lovo.go:436-442: |i| - |j| <= eps */
    return fabs(i) - fabs(j) <= eps;
}

/* align.LOVO (lovo.go:198-258) on an in-memory trajectory.  fullindexes = res2atoms(...) is
 * supplied by the caller (ascending atom indexes, :504-516).
 * Outputs: *n_out = N; natoms_out[0..n) = Natoms (previous iteration's order, new n, :255);
 * rmsd_out[0..nfull) = RMSD re-sorted by atom index (:254); *frames_read; *iterations.
 * max_iter guards against non-convergence (the reference loops forever). */
int orc_lovo(const double *frames, long F, int N, const double *ref, const int *fullindexes,
             int nfull, int cpus, int begin, int skip, double less_than_rmsd, int minimum_n,
             int nthreads, int max_iter, int *n_out, int *natoms_out, double *rmsd_out,
             long *frames_read, int *iterations, int *disagreement_trace /* optional, len max_iter */) {
    int *indexesold = (int *)malloc(sizeof(int) * (size_t)(nfull > 0 ? nfull : 1));
    int *ids = (int *)malloc(sizeof(int) * (size_t)(nfull > 0 ? nfull : 1));
    double *vals = (double *)malloc(sizeof(double) * (size_t)(nfull > 0 ? nfull : 1));
    double *rmsd = (double *)malloc(sizeof(double) * (size_t)N);
    if (!indexesold || !ids || !vals || !rmsd) return ORC_ERR_ALLOC;
    memcpy(indexesold, fullindexes, sizeof(int) * (size_t)nfull);
    int n = nfull, itercount = 0, err = ORC_OK;
    double prevlim = 0.0;
    long total = 0;
    for (;;) {
        err = orc_rmsd_traj(frames, F, N, ref, indexesold, n, cpus, begin, skip, nthreads, 0, rmsd,
                            &total, NULL); /* :229 */
        if (err) break;
        itercount++;
        /* rmsdFilter (:519-528) walks rmsd in index order keeping i in fullindexes */
        int k = 0;
        for (int i = 0; i < N; i++)
            if (is_in(i, fullindexes, nfull)) {
                if (k < nfull) { ids[k] = fullindexes[k]; vals[k] = rmsd[i]; }
                k++;
            }
        stable_sort_pairs(ids, vals, nfull, 1); /* :236 */
        double newlim = 0.0;
        if (less_than_rmsd > 0) { /* :238-240 */
            err = orc_mobile_limit(less_than_rmsd, minimum_n, vals, nfull, &n, &newlim);
            if (err) break;
        }
        if (n < 0 || n > nfull) { err = ORC_ERR_SHAPE; break; } /* Go would panic on the slice */
        int converged = same_elements(ids, indexesold, n) &&
                        (newlim == less_than_rmsd || approx_eq(newlim, prevlim, 0.01)); /* :246 */
        if (converged) break;
        prevlim = newlim;
        if (disagreement_trace && itercount <= max_iter)
            disagreement_trace[itercount - 1] = disagreement(ids, indexesold, n);
        memcpy(indexesold, ids, sizeof(int) * (size_t)nfull); /* :251 */
        if (itercount >= max_iter) { err = ORC_ERR_NOTERM; break; }
    }
    if (!err) {
        stable_sort_pairs(ids, vals, nfull, 0); /* :254 */
        *n_out = n;
        memcpy(natoms_out, indexesold, sizeof(int) * (size_t)n);
        memcpy(rmsd_out, vals, sizeof(double) * (size_t)nfull);
        *frames_read = total;
        *iterations = itercount;
    }
    free(indexesold);
    free(ids);
    free(vals);
    free(rmsd);
    return err;
}

/* chem.CenterOfMass (geometric.go:609-631) and chem.MomentTensor (geometric.go:705-733) of one structure.
 * mass == NULL: unit masses.  The moment tensor is center^T * center with rows of the centred matrix scaled by
 * sqrt(m_i) (pow(mass, sqrmass, 0.5), :722-725), accumulated row by row as Dgemm does. */
int orc_center_and_moment(const double *a, int n, const double *mass, double *center, double *moment) {
    if (n <= 0) return ORC_ERR_SHAPE;
    double s0 = 0, s1 = 0, s2 = 0, msum = 0;
    for (int i = 0; i < n; i++) {
        double m = mass ? mass[i] : 1.0;
        s0 += a[3 * i] * m; s1 += a[3 * i + 1] * m; s2 += a[3 * i + 2] * m;
        msum += m;
    }
    double inv = 1.0 / msum;
    double c[3] = {s0 * inv, s1 * inv, s2 * inv};
    if (center) memcpy(center, c, sizeof c);
    if (moment) {
        for (int k = 0; k < 9; k++) moment[k] = 0.0;
        for (int i = 0; i < n; i++) {
            double sq = sqrt(mass ? mass[i] : 1.0);
            double r[3] = {(a[3 * i] - c[0]) * sq, (a[3 * i + 1] - c[1]) * sq, (a[3 * i + 2] - c[2]) * sq};
            for (int p = 0; p < 3; p++)
                for (int q = 0; q < 3; q++) moment[3 * p + q] += r[p] * r[q];
        }
    }
    return ORC_OK;
}

/* exposed for tests */
void orc_stable_sort_pairs(int *ids, double *vals, int n, int by_rmsd) {
    stable_sort_pairs(ids, vals, n, by_rmsd);
}
int orc_approx_eq(double i, double j, double eps) { return approx_eq(i, j, eps); }
int orc_same_elements(const int *a, const int *b, int n) { return same_elements(a, b, n); }
int orc_disagreement(const int *a, const int *b, int n) { return disagreement(a, b, n); }

/* ------------------------------------------------------------------ */
/* solv: molecular RDF / MDDF, solv/solvation.go   */
/* ------------------------------------------------------------------ */
/* Topology crosses as integers: molid[i] = Atom.MolID, chain[i] = an id of the Atom.Chain string (0 = the empty string,
 * the zero value chain_skip starts with, solvation.go:457-459), flags[i] bit 0 = MolName is in the `residues` list,
 * bit 1 = MolName is SOL / WAT / HOH (the names pre-screened at solvation.go:480-483). */

/* solvation.go:692-695: temp.Sub(r, t); temp.Norm(2) */
static double dist3(const double *r, const double *t) {
    double d[3] = {r[0] - t[0], r[1] - t[1], r[2] - t[2]};
    return frob(d, 3);
}

/* solv.MolShortestDist (solvation.go:669-690): smallest |ref_j - test_i|, starting from 100000 */
static double mol_shortest_dist(const double *test, int ntest, const double *ref, int nref) {
    double dclosest = 100000;
    for (int i = 0; i < ntest; i++)
        for (int j = 0; j < nref; j++) {
            double d1 = dist3(ref + 3 * j, test + 3 * i);
            if (d1 < dclosest) dclosest = d1;
        }
    return dclosest;
}

/* solv.DistRank (solvation.go:447-541).  out_dist / out_molid need room for natoms entries.  With com != 0 the
 * residue is replaced by its geometric centre: the mass vector is only built when Masses() FAILS (solvation.go:523-526),
 * so chem.CenterOfMass receives nil and returns the unweighted centre (geometric.go:615-618), its sum scaled by
 * 1.0 / n (geometric.go:628-629). */
int orc_dist_rank(const double *coord, int natoms, const int *molid, const int *chain, const unsigned char *flags,
                  const int *refidx, int nref, int com, double end, double *out_dist, int *out_molid) {
    if (nref <= 0) return 0;
    double *ref = (double *)malloc(sizeof(double) * 3 * (size_t)nref);
    double *test = (double *)malloc(sizeof(double) * 3 * (size_t)natoms);
    int *own_id = (int *)malloc(sizeof(int) * (size_t)nref), *own_chain = (int *)malloc(sizeof(int) * (size_t)nref);
    int nown = 0, count = 0;
    for (int k = 0; k < nref; k++) {   /* allResIDandChains, solvation.go:418-427 */
        int id = molid[refidx[k]], ch = chain[refidx[k]], rep = 0;
        for (int q = 0; q < nown; q++) if (own_chain[q] == ch && own_id[q] == id) { rep = 1; break; }
        if (!rep) { own_id[nown] = id; own_chain[nown] = ch; nown++; }
        memcpy(ref + 3 * k, coord + 3 * (size_t)refidx[k], 3 * sizeof(double));   /* ref.SomeVecs(coord, refindexes) */
    }
    const double cutoff = end, cutoffplus = cutoff + 3;
    int molid_skip = -1, chain_skip = 0;
    for (int i = 0; i < natoms; i++) {
        const int id = molid[i], ch = chain[i];
        if ((flags[i] & 2) && dist3(ref, coord + 3 * (size_t)i) > cutoffplus) continue;   /* :481-483 */
        int own = 0;
        for (int q = 0; q < nown; q++) if (own_chain[q] == ch && own_id[q] == id) { own = 1; break; }
        if ((flags[i] & 1) && !own && (id != molid_skip || ch != chain_skip)) {
            int n = 0;
            memcpy(test, coord + 3 * (size_t)i, 3 * sizeof(double));
            n = 1;
            for (int j = i + 1; j < natoms; j++) {   /* :490-496: only MolID is compared */
                if (molid[j] != id) break;
                memcpy(test + 3 * n, coord + 3 * (size_t)j, 3 * sizeof(double));
                n++;
            }
            double distance;
            if (com) {
                double c[3] = {0, 0, 0};
                for (int a = 0; a < n; a++) { c[0] += test[3 * a] * 1.0; c[1] += test[3 * a + 1] * 1.0; c[2] += test[3 * a + 2] * 1.0; }
                const double inv = 1.0 / (double)n;
                c[0] *= inv; c[1] *= inv; c[2] *= inv;
                distance = mol_shortest_dist(c, 1, ref, nref);
            } else {
                distance = mol_shortest_dist(test, n, ref, nref);
            }
            if (distance <= cutoff) { out_dist[count] = distance; out_molid[count] = id; count++; }
            molid_skip = id;
            chain_skip = ch;
        }
    }
    free(ref); free(test); free(own_id); free(own_chain);
    return count;
}

static int cmp_double(const void *a, const void *b) {
    const double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* solv.FrameUMolCRDF (solvation.go:380-410): ret[i-1] = n(i) with n = sort.SearchFloat64s(dists, i*step) (the number
 * of distances below the limit), reduced by one when it is not zero (:402-404).  Returns totalsteps = int(end/step). */
int orc_frame_umol_crdf(const double *coord, int natoms, const int *molid, const int *chain, const unsigned char *flags,
                        const int *refidx, int nref, int com, double step, double end, double *ret) {
    const int totalsteps = (int)(end / step);
    double *d = (double *)malloc(sizeof(double) * (size_t)(natoms > 0 ? natoms : 1));
    int *ids = (int *)malloc(sizeof(int) * (size_t)(natoms > 0 ? natoms : 1));
    const int cnt = orc_dist_rank(coord, natoms, molid, chain, flags, refidx, nref, com, end, d, ids);
    qsort(d, (size_t)cnt, sizeof(double), cmp_double);
    for (int i = 1; i <= totalsteps; i++) {
        const double limit = (double)i * step;
        int lo = 0, hi = cnt;   /* smallest index with d[idx] >= limit */
        while (lo < hi) { int mid = lo + (hi - lo) / 2; if (d[mid] < limit) lo = mid + 1; else hi = mid; }
        int n = lo;
        if (n > 0) n--;
        ret[i - 1] = (double)n;
    }
    free(d); free(ids);
    return totalsteps;
}

/* solv.MDFFromCDF (solvation.go:317-352).  ret (the summed cumulative counts) is overwritten with the normalised
 * density, ret2 receives the average cumulative number per shell.  Returns -1 when A or B < 1. */
int orc_mdf_from_cdf(double *ret, int n, int framesread, double A, double B, double step, double *ret2) {
    if (A < 1.0 || B < 1.0) return ORC_ERR_SHAPE;
    const double vp = (4.0 / 3.0) * M_PI * A * B;
    double acc = 0.0, vol;
    for (int i = 0; i < n; i++) ret2[i] = 0.0;
    for (int i = 0; i < n; i++) {
        if (i >= 1) {
            acc = acc + ret[i - 1];
            ret2[i] = ret[i];
            ret[i] = ret[i] - acc;
        }
    }
    for (int i = 0; i < n; i++) {
        if (i >= 1) {
            const double fi = (double)i;
            vol = vp * (pow((fi + 1) * step, 3) - pow(fi * step, 3));
        } else {
            vol = vp * pow(step, 3);
        }
        ret[i] = ret[i] / (double)framesread;
        ret2[i] = ret2[i] / (double)framesread;
        ret[i] = ret[i] / vol;
    }
    const double last = ret[n - 1];
    for (int i = 0; i < n; i++) ret[i] = ret[i] / last;
    return ORC_OK;
}

/* The frame loop of solv.ConcMolRDF (solvation.go:166-214): frames are taken cpus at a time and every frame read adds its
 * FrameUMolCRDF to the sum (:200-206); a trailing chunk a reader cannot fill is lost (dcd.go:520-523 returns no channels).
 * cdf_sum[totalsteps] is the sum before MDFFromCDF.  Returns the number of frames read. */
long orc_rdf_cdf_sum(const double *frames, long F, int natoms, const int *molid, const int *chain, const unsigned char *flags,
                     const int *refidx, int nref, int com, double step, double end, int cpus, double *cdf_sum) {
    const int totalsteps = (int)(end / step);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(totalsteps > 0 ? totalsteps : 1));
    for (int k = 0; k < totalsteps; k++) cdf_sum[k] = 0.0;
    long read = 0;
    if (cpus < 1) cpus = 1;
    for (long c0 = 0; c0 + cpus <= F; c0 += cpus)
        for (int s = 0; s < cpus; s++) {
            orc_frame_umol_crdf(frames + (size_t)(c0 + s) * 3 * (size_t)natoms, natoms, molid, chain, flags, refidx, nref, com, step, end, tmp);
            for (int k = 0; k < totalsteps; k++) cdf_sum[k] = cdf_sum[k] + tmp[k];
            read++;
        }
    free(tmp);
    return read;
}
