// kabsch3.cuh -- 3x3 rotation solve for the Kabsch step (fp64), host + device.
//
// Replaces gonum mat.SVD.Factorize/UTo/VTo on the 3x3 correlation matrix and the sign /
// reflection handling around it in chem.RotatorTranslatorToSuper
// (/root/reference/geometric.go:154-190): H = U S V^T, R = U V^T, and if det R < 0 then
// R = U diag(1,1,-1) V^T.  R right-multiplies row vectors.
//
// Both cases collapse into one expression that needs only the two dominant singular pairs:
//     R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T
// ([u1 u2 u1xu2] and [v1 v2 v1xv2] are proper, so R is the proper rotation taking v1->u1,
// v2->u2; when det H > 0 that is U V^T, when det H < 0 it is U diag(1,1,-1) V^T.)
//
// Two solvers:
//   kabsch3_jacobi : one-sided (Hestenes) Jacobi SVD.  Always valid for rank >= 2 input.
//   kabsch3_polar  : scaled Newton iteration for the orthogonal polar factor, valid when
//                    det H > 0 and H is not close to singular (every frame of a real trajectory
//                    fitted onto its own reference); about 4x shorter dependency chain.
//   kabsch3        : polar when its guard holds, Jacobi otherwise.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define K3_HD __host__ __device__ __forceinline__
#else
#define K3_HD inline
#endif

namespace k3 {

K3_HD double det3(const double* a) {
    return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
           a[2] * (a[3] * a[7] - a[4] * a[6]);
}

// One-sided Jacobi.  h row-major 3x3.  Returns 0, or -4 (GCB_ERR_SVD) for non-finite input.
// rot row-major, right-multiplying.
K3_HD int kabsch3_jacobi(const double* h, double* rot) {
    double a[9], v[9];
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        a[i] = h[i];
        nrm += h[i] * h[i];
        v[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    if (!(nrm < 1.79e308) || nrm != nrm) return -4;
    if (nrm == 0.0) {
#pragma unroll
        for (int i = 0; i < 9; i++) rot[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return 0;
    }
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; pq++) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            double alpha = a[p] * a[p] + a[3 + p] * a[3 + p] + a[6 + p] * a[6 + p];
            double beta = a[q] * a[q] + a[3 + q] * a[3 + q] + a[6 + q] * a[6 + q];
            double gamma = a[p] * a[q] + a[3 + p] * a[3 + q] + a[6 + p] * a[6 + q];
            if (gamma * gamma <= 1e-34 * alpha * beta) continue;
            rotated = true;
            double zeta = (beta - alpha) / (2.0 * gamma);
            double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                double ap = a[3 * i + p], aq = a[3 * i + q];
                a[3 * i + p] = c * ap - s * aq;
                a[3 * i + q] = s * ap + c * aq;
                double vp = v[3 * i + p], vq = v[3 * i + q];
                v[3 * i + p] = c * vp - s * vq;
                v[3 * i + q] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    // column norms^2; pick the two largest columns (order between them is irrelevant to R)
    double n0 = a[0] * a[0] + a[3] * a[3] + a[6] * a[6];
    double n1 = a[1] * a[1] + a[4] * a[4] + a[7] * a[7];
    double n2 = a[2] * a[2] + a[5] * a[5] + a[8] * a[8];
    int j1, j2;
    if (n0 <= n1 && n0 <= n2) { j1 = 1; j2 = 2; }
    else if (n1 <= n0 && n1 <= n2) { j1 = 2; j2 = 0; }
    else { j1 = 0; j2 = 1; }
    double na = (j1 == 0) ? n0 : (j1 == 1 ? n1 : n2);
    double nb = (j2 == 0) ? n0 : (j2 == 1 ? n1 : n2);
    double u1[3], u2[3], v1[3], v2[3];
    double ia = na > 0.0 ? 1.0 / sqrt(na) : 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        u1[i] = a[3 * i + j1] * ia;
        v1[i] = v[3 * i + j1];
        v2[i] = v[3 * i + j2];
    }
    if (nb > 1e-300 * na && nb > 0.0) {
        double ib = 1.0 / sqrt(nb);
#pragma unroll
        for (int i = 0; i < 3; i++) u2[i] = a[3 * i + j2] * ib;
        // re-orthogonalise u2 against u1 (columns are orthogonal to ~1e-17 relative already)
        double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
        u2[0] -= d * u1[0]; u2[1] -= d * u1[1]; u2[2] -= d * u1[2];
        double l = 1.0 / sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
        u2[0] *= l; u2[1] *= l; u2[2] *= l;
    } else {
        // rank one: any unit vector orthogonal to u1 gives an optimal rotation
        int m = 0;
        if (fabs(u1[1]) < fabs(u1[m])) m = 1;
        if (fabs(u1[2]) < fabs(u1[m])) m = 2;
        double e[3] = {0.0, 0.0, 0.0};
        e[m] = 1.0;
        u2[0] = u1[1] * e[2] - u1[2] * e[1];
        u2[1] = u1[2] * e[0] - u1[0] * e[2];
        u2[2] = u1[0] * e[1] - u1[1] * e[0];
        double l = 1.0 / sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
        u2[0] *= l; u2[1] *= l; u2[2] *= l;
    }
    double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) rot[3 * i + j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j];
    return 0;
}

// Scaled Newton iteration X <- (g X + X^-T / g) / 2 for the orthogonal polar factor of H.
// Returns 1 when it converged to a proper rotation, 0 when the caller must use the Jacobi path.
K3_HD int kabsch3_polar(const double* h, double* rot) {
    double x[9];
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) { x[i] = h[i]; nrm += h[i] * h[i]; }
    if (!(nrm > 0.0) || !(nrm < 1.79e308)) return 0;
    double d = det3(x);
    // guard: proper and comfortably non-singular (det = s1 s2 s3; ||H||_F^3 >= 3^1.5 s1 s2 s3)
    double fn = sqrt(nrm);
    if (!(d > 1e-6 * fn * fn * fn)) return 0;
    for (int it = 0; it < 24; it++) {
        // cofactor matrix c = det * X^-T
        double c[9];
        c[0] = x[4] * x[8] - x[5] * x[7]; c[1] = x[5] * x[6] - x[3] * x[8]; c[2] = x[3] * x[7] - x[4] * x[6];
        c[3] = x[2] * x[7] - x[1] * x[8]; c[4] = x[0] * x[8] - x[2] * x[6]; c[5] = x[1] * x[6] - x[0] * x[7];
        c[6] = x[1] * x[5] - x[2] * x[4]; c[7] = x[2] * x[3] - x[0] * x[5]; c[8] = x[0] * x[4] - x[1] * x[3];
        double det = x[0] * c[0] + x[1] * c[1] + x[2] * c[2];
        if (!(det > 0.0)) return 0;
        double nx = 0.0, nc = 0.0;
#pragma unroll
        for (int i = 0; i < 9; i++) { nx += x[i] * x[i]; nc += c[i] * c[i]; }
        // Frobenius scaling g = sqrt(||X^-T||_F / ||X||_F) with X^-T = c/det
        double g2 = sqrt(nc / nx) / det;        // g^2
        double g = sqrt(g2);
        double a = 0.5 * g, b = 0.5 / (g * det);
        double diff = 0.0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            double y = a * x[i] + b * c[i];
            double e = y - x[i];
            diff += e * e;
            x[i] = y;
        }
        if (diff <= 1e-26) {  // ||dX||_F <= 1e-13: quadratic convergence puts the remaining error near 1e-26
#pragma unroll
            for (int i = 0; i < 9; i++) rot[i] = x[i];
            return 1;
        }
    }
    return 0;
}

K3_HD int kabsch3(const double* h, double* rot) {
    if (kabsch3_polar(h, rot)) return 0;
    return kabsch3_jacobi(h, rot);
}

}  // namespace k3
