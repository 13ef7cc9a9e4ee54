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

// Reciprocal with a short dependency chain: hardware seed (about 20 good bits) and one cubically
// convergent correction r (1 + e + e^2), e = 1 - x r  -> relative error ~2^-60.  Plain division on the host.
K3_HD double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#else
    return 1.0 / x;
#endif
}

// One Newton step Y = (X + X^-T) / 2 on a 3x3 held in nine scalars; returns det(X).  Dependency chain: cofactors (2) ->
// det (3) -> reciprocal (4) -> update (1).
K3_HD double newton_step(const double (&x)[9], double (&y)[9]) {
    const double c0 = fma(x[4], x[8], -x[5] * x[7]), c1 = fma(x[5], x[6], -x[3] * x[8]), c2 = fma(x[3], x[7], -x[4] * x[6]);
    const double c3 = fma(x[2], x[7], -x[1] * x[8]), c4 = fma(x[0], x[8], -x[2] * x[6]), c5 = fma(x[1], x[6], -x[0] * x[7]);
    const double c6 = fma(x[1], x[5], -x[2] * x[4]), c7 = fma(x[2], x[3], -x[0] * x[5]), c8 = fma(x[0], x[4], -x[1] * x[3]);
    const double det = fma(x[2], c2, fma(x[1], c1, x[0] * c0));
    const double r = fast_rcp(det);
    y[0] = fma(0.5 * c0, r, 0.5 * x[0]); y[1] = fma(0.5 * c1, r, 0.5 * x[1]); y[2] = fma(0.5 * c2, r, 0.5 * x[2]);
    y[3] = fma(0.5 * c3, r, 0.5 * x[3]); y[4] = fma(0.5 * c4, r, 0.5 * x[4]); y[5] = fma(0.5 * c5, r, 0.5 * x[5]);
    y[6] = fma(0.5 * c6, r, 0.5 * x[6]); y[7] = fma(0.5 * c7, r, 0.5 * x[7]); y[8] = fma(0.5 * c8, r, 0.5 * x[8]);
    return det;
}

// Newton iteration X <- (X + X^-T) / 2 for the orthogonal polar factor of H, started from
// H * sqrt(3) / ||H||_F (singular values of order one, so no per-step scaling, hence no square
// roots or divisions inside the loop: one seeded reciprocal of the determinant per step).
// The polar factor does not depend on the scale of the start, so the scale is the raw hardware
// reciprocal-square-root seed (20 good bits, one instruction) -- nothing about the result depends on its error.
// Termination without a serial test: with b = N(a) and c = N(b), the distance d = ||b - a||_F bounds the
// error of b by d^2 / 2 and that of c by d^4 / 8; c is accepted when d^2 <= 2e-8 (error < 1e-16).  The
// test of step k is independent of step k + 1, so both are in flight together: the chain per iteration is the
// ten dependent operations of one Newton step.
// Returns 1 when it converged to a proper rotation, 0 when the caller must use the Jacobi path
// (det H <= 0, near-singular H, non-finite input).
K3_HD int kabsch3_polar(const double* h, double* rot) {
    const double n0 = fma(h[2], h[2], fma(h[1], h[1], h[0] * h[0]));
    const double n1 = fma(h[5], h[5], fma(h[4], h[4], h[3] * h[3]));
    const double n2 = fma(h[8], h[8], fma(h[7], h[7], h[6] * h[6]));
    const double nrm = n0 + n1 + n2;
    if (!(nrm > 1e-280) || !(nrm < 1e280)) return 0;
#if defined(__CUDA_ARCH__)
    double sc;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(sc) : "d"(nrm));
    sc *= 1.7320508075688772;
#else
    const double sc = 1.7320508075688772 / sqrt(nrm);
#endif
    double a[9], b[9], c[9];
#pragma unroll
    for (int k = 0; k < 9; k++) c[k] = h[k] * sc;
    // guard on the first step: proper and comfortably non-singular (det = det(H) 3^1.5 / ||H||_F^3)
    double det = newton_step(c, b);
    if (!(det > 5e-6)) return 0;
    det = newton_step(b, a);
    det = newton_step(a, b);
    if (!(det > 0.0)) return 0;
    for (int it = 0; it < 40; it++) {
        det = newton_step(b, c);
        double d = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double e0 = b[k] - a[k], e1 = b[3 + k] - a[3 + k], e2 = b[6 + k] - a[6 + k];
            d = fma(e0, e0, d); d1 = fma(e1, e1, d1); d2 = fma(e2, e2, d2);
        }
        d += d1 + d2;
        if (!(det > 0.0)) return 0;
        if (d <= 2e-8) {
#pragma unroll
            for (int k = 0; k < 9; k++) rot[k] = c[k];
            return 1;
        }
#pragma unroll
        for (int k = 0; k < 9; k++) { a[k] = b[k]; b[k] = c[k]; }
    }
    return 0;
}

K3_HD int kabsch3(const double* h, double* rot) {
    if (kabsch3_polar(h, rot)) return 0;
    return kabsch3_jacobi(h, rot);
}

}  // namespace k3
