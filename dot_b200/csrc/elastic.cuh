// Per-tet device math: deformation gradient, 3x3 SVD, FCR / SNH energy densities, first Piola
// stress, PD-projected dP/dF and the 12x12 elemental Hessian.  fp64 throughout.
//
// Reference behaviour restated (paths relative to the reference's src/):
//   F = Ds Dm^-1                         Energy/Energy.cpp:457-473
//   SVD conventions                      Utils/IglUtils.cpp:929-1085, Utils/AutoFlipSVD.hpp:34-81
//   FCR  Psi, dPsi/dsigma, d2, BLeft     Energy/Physics_Elasticity/FixedCoRotEnergy.cpp:83-183
//   SNH  Psi, dPsi/dsigma, d2, BLeft     Energy/Physics_Elasticity/StableNHEnergy.cpp:80-251
//   dP/dF with per-block PD projection   Energy/Energy.cpp:1129-1270, Utils/IglUtils.hpp:252-309
//   g_e, H_e from dF/dx                  Utils/IglUtils.cpp:821-870, Energy/Energy.cpp:738-777
#pragma once
#include <cuda_runtime.h>

namespace dotgpu {

#define DG_FCR 0
#define DG_SNH 1

struct Mat3 {
    double m[9];  // row-major
    __device__ __forceinline__ double& operator()(int i, int j) { return m[3 * i + j]; }
    __device__ __forceinline__ double operator()(int i, int j) const { return m[3 * i + j]; }
};

__device__ __forceinline__ double det3(const Mat3& F) {
    return F(0, 0) * (F(1, 1) * F(2, 2) - F(1, 2) * F(2, 1)) - F(0, 1) * (F(1, 0) * F(2, 2) - F(1, 2) * F(2, 0)) +
           F(0, 2) * (F(1, 0) * F(2, 1) - F(1, 1) * F(2, 0));
}

// cofactor matrix J F^-T (IglUtils::computeCofactorMtr, IglUtils.hpp:485-513)
__device__ __forceinline__ void cofactor3(const Mat3& F, Mat3& C) {
    C(0, 0) = F(1, 1) * F(2, 2) - F(1, 2) * F(2, 1);
    C(0, 1) = F(1, 2) * F(2, 0) - F(1, 0) * F(2, 2);
    C(0, 2) = F(1, 0) * F(2, 1) - F(1, 1) * F(2, 0);
    C(1, 0) = F(0, 2) * F(2, 1) - F(0, 1) * F(2, 2);
    C(1, 1) = F(0, 0) * F(2, 2) - F(0, 2) * F(2, 0);
    C(1, 2) = F(0, 1) * F(2, 0) - F(0, 0) * F(2, 1);
    C(2, 0) = F(0, 1) * F(1, 2) - F(0, 2) * F(1, 1);
    C(2, 1) = F(0, 2) * F(1, 0) - F(0, 0) * F(1, 2);
    C(2, 2) = F(0, 0) * F(1, 1) - F(0, 1) * F(1, 0);
}

// One Hestenes rotation on columns p,q of A (and V).  Returns true if a rotation was applied.
template <int P, int Q>
__device__ __forceinline__ bool jacobi_pair(Mat3& A, Mat3& V) {
    const double tol = 8.0 * 2.220446049250313e-16;
    double alpha = A(0, P) * A(0, P) + A(1, P) * A(1, P) + A(2, P) * A(2, P);
    double beta = A(0, Q) * A(0, Q) + A(1, Q) * A(1, Q) + A(2, Q) * A(2, Q);
    double gamma = A(0, P) * A(0, Q) + A(1, P) * A(1, Q) + A(2, P) * A(2, Q);
    if (!(fabs(gamma) > tol * sqrt(alpha * beta))) return false;
    double zeta = (beta - alpha) / (2.0 * gamma);
    double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    double c = rsqrt(1.0 + t * t);
    double s = c * t;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double ap = A(i, P), aq = A(i, Q);
        A(i, P) = c * ap - s * aq;
        A(i, Q) = s * ap + c * aq;
        double vp = V(i, P), vq = V(i, Q);
        V(i, P) = c * vp - s * vq;
        V(i, Q) = s * vp + c * vq;
    }
    return true;
}

template <int P, int Q>
__device__ __forceinline__ void sort_pair(Mat3& A, Mat3& V, double* n2) {
    if (n2[P] < n2[Q]) {  // swap columns, negate the one landing in Q: a proper rotation keeps det V = +1
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double ap = A(i, P), aq = A(i, Q);
            A(i, P) = aq;
            A(i, Q) = -ap;
            double vp = V(i, P), vq = V(i, Q);
            V(i, P) = vq;
            V(i, Q) = -vp;
        }
        double tmp = n2[P];
        n2[P] = n2[Q];
        n2[Q] = tmp;
    }
}

// F = U diag(S) V^T, det U = det V = +1, |S0| >= |S1| >= |S2|, sign on S2.
// One-sided Jacobi (the reference iterates the same cyclic sweeps on A^T A, SVD_EFTYCHIOS
// Main_Kernel_Body.hpp:51-91), column sort (:577-870) and QR by Gram-Schmidt + cross product (:944-1151).
__device__ __forceinline__ void svd3(const Mat3& F, Mat3& U, double* S, Mat3& V) {
    Mat3 A = F;
#pragma unroll
    for (int i = 0; i < 9; ++i) V.m[i] = 0.0;
    V(0, 0) = V(1, 1) = V(2, 2) = 1.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        bool r0 = jacobi_pair<0, 1>(A, V);
        bool r1 = jacobi_pair<0, 2>(A, V);
        bool r2 = jacobi_pair<1, 2>(A, V);
        if (!(r0 || r1 || r2)) break;
    }
    double n2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) n2[j] = A(0, j) * A(0, j) + A(1, j) * A(1, j) + A(2, j) * A(2, j);
    sort_pair<0, 1>(A, V, n2);
    sort_pair<0, 2>(A, V, n2);
    sort_pair<1, 2>(A, V, n2);
    double s0 = sqrt(n2[0]);
    double u0[3], u1[3], u2[3];
    if (s0 > 0.0) {
        double inv = 1.0 / s0;
        u0[0] = A(0, 0) * inv; u0[1] = A(1, 0) * inv; u0[2] = A(2, 0) * inv;
    } else {
        u0[0] = 1.0; u0[1] = 0.0; u0[2] = 0.0;
    }
    double d = u0[0] * A(0, 1) + u0[1] * A(1, 1) + u0[2] * A(2, 1);
    double b1[3] = {A(0, 1) - d * u0[0], A(1, 1) - d * u0[1], A(2, 1) - d * u0[2]};
    double n1 = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    if (n1 > 1e-300) {
        double inv = 1.0 / n1;
        u1[0] = b1[0] * inv; u1[1] = b1[1] * inv; u1[2] = b1[2] * inv;
    } else {  // rank <= 1: any unit vector orthogonal to u0 (axis least aligned with u0)
        int k = 0;
        if (fabs(u0[1]) < fabs(u0[k])) k = 1;
        if (fabs(u0[2]) < fabs(u0[k])) k = 2;
        double e[3] = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
        double dd = u0[k];
        double a0 = e[0] - dd * u0[0], a1 = e[1] - dd * u0[1], a2 = e[2] - dd * u0[2];
        double inv = rsqrt(a0 * a0 + a1 * a1 + a2 * a2);
        u1[0] = a0 * inv; u1[1] = a1 * inv; u1[2] = a2 * inv;
    }
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
    u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
    u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    S[0] = s0;
    S[1] = u1[0] * A(0, 1) + u1[1] * A(1, 1) + u1[2] * A(2, 1);
    S[2] = u2[0] * A(0, 2) + u2[1] * A(1, 2) + u2[2] * A(2, 2);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        U(i, 0) = u0[i];
        U(i, 1) = u1[i];
        U(i, 2) = u2[i];
    }
}

// ---- sigma-space material functions -------------------------------------------------------
template <int EN>
__device__ __forceinline__ double psi_sigma(const double* s, double mu, double lam) {
    double J = s[0] * s[1] * s[2];
    if (EN == DG_FCR) {
        double a = s[0] - 1.0, b = s[1] - 1.0, c = s[2] - 1.0;
        return mu * (a * a + b * b + c * c) + lam / 2.0 * (J - 1.0) * (J - 1.0);
    } else {
        double JmA = J - (1.0 + mu / lam);
        return (mu * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2] - 3.0) + lam * JmA * JmA) / 2.0;
    }
}

template <int EN>
__device__ __forceinline__ void dpsi_dsigma(const double* s, double mu, double lam, double* dE) {
    double J = s[0] * s[1] * s[2];
    double n0 = s[1] * s[2], n1 = s[2] * s[0], n2 = s[0] * s[1];
    if (EN == DG_FCR) {
        double k = lam * (J - 1.0), m2 = 2.0 * mu;
        dE[0] = m2 * (s[0] - 1.0) + n0 * k;
        dE[1] = m2 * (s[1] - 1.0) + n1 * k;
        dE[2] = m2 * (s[2] - 1.0) + n2 * k;
    } else {
        double t2 = lam * (J - (1.0 + mu / lam));
        dE[0] = s[0] * mu + t2 * n0;
        dE[1] = s[1] * mu + t2 * n1;
        dE[2] = s[2] * mu + t2 * n2;
    }
}

// A = d2Psi/dsigma2 as (a00,a11,a22,a01,a02,a12); BL = left coefficients of the pairs (01),(12),(20)
template <int EN>
__device__ __forceinline__ void d2psi_dsigma2(const double* s, double mu, double lam, double* A, double* BL) {
    double J = s[0] * s[1] * s[2];
    double n0 = s[1] * s[2], n1 = s[2] * s[0], n2 = s[0] * s[1];
    if (EN == DG_FCR) {
        double m2 = 2.0 * mu;
        A[0] = m2 + lam * n0 * n0;
        A[1] = m2 + lam * n1 * n1;
        A[2] = m2 + lam * n2 * n2;
        A[3] = lam * (s[2] * (J - 1.0) + n0 * n1);
        A[4] = lam * (s[1] * (J - 1.0) + n0 * n2);
        A[5] = lam * (s[0] * (J - 1.0) + n2 * n1);
        double hl = lam / 2.0;
        BL[0] = mu - hl * s[2] * (J - 1.0);
        BL[1] = mu - hl * s[0] * (J - 1.0);
        BL[2] = mu - hl * s[1] * (J - 1.0);
    } else {
        double alpha = 1.0 + mu / lam;
        double l2 = lam * (2.0 * J - alpha);
        A[0] = mu + lam * n0 * n0;
        A[1] = mu + lam * n1 * n1;
        A[2] = mu + lam * n2 * n2;
        A[3] = s[2] * l2;
        A[4] = s[1] * l2;
        A[5] = s[0] * l2;
        double t0 = lam * (J - alpha);
        BL[0] = (mu - t0 * s[2]) / 2.0;
        BL[1] = (mu - t0 * s[0]) / 2.0;
        BL[2] = (mu - t0 * s[1]) / 2.0;
    }
}

// IglUtils::makePD2d (IglUtils.hpp:270-309) on [[a,b],[b,d]], formula kept as the reference has it.
__device__ __forceinline__ void make_pd2(double& a, double& b, double& d) {
    double b2 = b * b;
    double D = a * d - b2;
    double T2 = (a + d) / 2.0;
    double sq = sqrt(fmax(T2 * T2 - D, 0.0));
    double L2 = T2 - sq;
    if (L2 < 0.0) {
        double L1 = T2 + sq;
        if (L1 <= 0.0) {
            a = b = d = 0.0;
        } else if (b2 == 0.0) {
            a = L1; b = 0.0; d = 0.0;
        } else {
            double L1md = L1 - d;
            double r = L1md / L1;
            a = r * L1md;
            b = b * r;
            d = b2 / L1;
        }
    }
}

// IglUtils::makePD (IglUtils.hpp:252-269) for a symmetric 3x3 given as (a00,a11,a22,a01,a02,a12):
// if the smallest eigenvalue is negative, clamp negative eigenvalues to zero.
__device__ __forceinline__ void make_pd3(double* A) {
    // cheap exit: Sylvester's criterion says positive definite
    double m2 = A[0] * A[1] - A[3] * A[3];
    double dt = A[0] * (A[1] * A[2] - A[5] * A[5]) - A[3] * (A[3] * A[2] - A[5] * A[4]) + A[4] * (A[3] * A[5] - A[1] * A[4]);
    if (A[0] > 0.0 && m2 > 0.0 && dt > 0.0) return;
    // cyclic Jacobi eigen-decomposition of the symmetric matrix
    double a[3][3] = {{A[0], A[3], A[4]}, {A[3], A[1], A[5]}, {A[4], A[5], A[2]}};
    double q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 16; ++sweep) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        double dg = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-17 * dg) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int r = (pq == 0) ? 1 : 2;
            double apq = a[p][r];
            if (apq == 0.0) continue;
            double theta = (a[r][r] - a[p][p]) / (2.0 * apq);
            double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
            double c = rsqrt(1.0 + t * t), s = c * t;
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- A J
                double akp = a[k][p], akr = a[k][r];
                a[k][p] = c * akp - s * akr;
                a[k][r] = s * akp + c * akr;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- J^T A
                double apk = a[p][k], ark = a[r][k];
                a[p][k] = c * apk - s * ark;
                a[r][k] = s * apk + c * ark;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double qkp = q[k][p], qkr = q[k][r];
                q[k][p] = c * qkp - s * qkr;
                q[k][r] = s * qkp + c * qkr;
            }
        }
    }
    double l0 = a[0][0], l1 = a[1][1], l2 = a[2][2];
    if (fmin(l0, fmin(l1, l2)) >= 0.0) return;
    l0 = fmax(l0, 0.0); l1 = fmax(l1, 0.0); l2 = fmax(l2, 0.0);
    A[0] = l0 * q[0][0] * q[0][0] + l1 * q[0][1] * q[0][1] + l2 * q[0][2] * q[0][2];
    A[1] = l0 * q[1][0] * q[1][0] + l1 * q[1][1] * q[1][1] + l2 * q[1][2] * q[1][2];
    A[2] = l0 * q[2][0] * q[2][0] + l1 * q[2][1] * q[2][1] + l2 * q[2][2] * q[2][2];
    A[3] = l0 * q[0][0] * q[1][0] + l1 * q[0][1] * q[1][1] + l2 * q[0][2] * q[1][2];
    A[4] = l0 * q[0][0] * q[2][0] + l1 * q[0][1] * q[2][1] + l2 * q[0][2] * q[2][2];
    A[5] = l0 * q[1][0] * q[2][0] + l1 * q[1][1] * q[2][1] + l2 * q[1][2] * q[2][2];
}

// Coefficients of w*dP/dF in the SVD frame (Energy.cpp:1129-1207):
//   A[6]   = w*(a00,a11,a22,a01,a02,a12)          (3x3 block on modes 00,11,22)
//   Dg[6]  = w*M[ab,ab] for ab = 01,10,12,21,02,20
//   Of[3]  = w*M[ab,ba] for the pairs (01),(12),(02)
struct HessCoef {
    double A[6], Dg[6], Of[3];
};

template <int EN>
__device__ __forceinline__ void hess_coef(const double* s, double mu, double lam, double w, bool project, HessCoef& h) {
    double dE[3], BL[3];
    dpsi_dsigma<EN>(s, mu, lam, dE);
    d2psi_dsigma2<EN>(s, mu, lam, h.A, BL);
    if (project) make_pd3(h.A);
    double b00[3], b01[3], b11[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int cp = (c + 1) % 3;
        double right = dE[c] + dE[cp];
        double ss = s[c] + s[cp];
        right /= 2.0 * (ss < 1.0e-6 ? 1.0e-6 : ss);
        double left = BL[c];
        b00[c] = left + right;
        b01[c] = left - right;
        b11[c] = left + right;
        if (project) make_pd2(b00[c], b01[c], b11[c]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) h.A[i] *= w;
    // pair (0,1): modes 01 <- B0(0,0), 10 <- B0(1,1); pair (1,2): 12 <- B1(0,0), 21 <- B1(1,1);
    // pair (2,0): 02 <- B2(1,1), 20 <- B2(0,0)   (index-reversed, Energy.cpp:1202-1206)
    h.Dg[0] = w * b00[0]; h.Dg[1] = w * b11[0];
    h.Dg[2] = w * b00[1]; h.Dg[3] = w * b11[1];
    h.Dg[4] = w * b11[2]; h.Dg[5] = w * b00[2];
    h.Of[0] = w * b01[0]; h.Of[1] = w * b01[1]; h.Of[2] = w * b01[2];
}

// 3x3 block (k,l) of H_e: U N U^T with N[a][c] built from beta_k = V^T w_k, beta_l = V^T w_l
// (w_k = gradient of F w.r.t. vertex k: rows of Dm^-1, w_0 = -sum).  Derivation: DESIGN.md section 4.3.
__device__ __forceinline__ void hess_block(const HessCoef& h, const Mat3& U, const double* bk, const double* bl, double* out9) {
    double N[3][3];
    N[0][0] = h.A[0] * bk[0] * bl[0] + h.Dg[0] * bk[1] * bl[1] + h.Dg[4] * bk[2] * bl[2];
    N[1][1] = h.A[1] * bk[1] * bl[1] + h.Dg[1] * bk[0] * bl[0] + h.Dg[2] * bk[2] * bl[2];
    N[2][2] = h.A[2] * bk[2] * bl[2] + h.Dg[3] * bk[1] * bl[1] + h.Dg[5] * bk[0] * bl[0];
    N[0][1] = h.A[3] * bk[0] * bl[1] + h.Of[0] * bk[1] * bl[0];
    N[1][0] = h.A[3] * bk[1] * bl[0] + h.Of[0] * bk[0] * bl[1];
    N[0][2] = h.A[4] * bk[0] * bl[2] + h.Of[2] * bk[2] * bl[0];
    N[2][0] = h.A[4] * bk[2] * bl[0] + h.Of[2] * bk[0] * bl[2];
    N[1][2] = h.A[5] * bk[1] * bl[2] + h.Of[1] * bk[2] * bl[1];
    N[2][1] = h.A[5] * bk[2] * bl[1] + h.Of[1] * bk[1] * bl[2];
    double T[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[i][c] = U(i, 0) * N[0][c] + U(i, 1) * N[1][c] + U(i, 2) * N[2][c];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int r = 0; r < 3; ++r) out9[3 * i + r] = T[i][0] * U(r, 0) + T[i][1] * U(r, 1) + T[i][2] * U(r, 2);
}

// First Piola stress.  SNH needs no SVD (StableNHEnergy.cpp:246-249): P = mu F + lam (J-alpha) cof F.
// FCR (FixedCoRotEnergy.cpp:179-182): P = 2 mu (F - U V^T) + lam (J-1) cof F.
template <int EN>
__device__ __forceinline__ void first_piola(const Mat3& F, double mu, double lam, Mat3& P, double& psi) {
    Mat3 C;
    cofactor3(F, C);
    double J = F(0, 0) * C(0, 0) + F(0, 1) * C(0, 1) + F(0, 2) * C(0, 2);
    if (EN == DG_SNH) {
        double JmA = J - (1.0 + mu / lam);
        double k = lam * JmA;
        double f2 = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            P.m[i] = mu * F.m[i] + k * C.m[i];
            f2 += F.m[i] * F.m[i];
        }
        psi = (mu * (f2 - 3.0) + lam * JmA * JmA) / 2.0;
    } else {
        Mat3 U, V;
        double S[3];
        svd3(F, U, S, V);
        double k = lam * (J - 1.0), m2 = 2.0 * mu;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double R = U(i, 0) * V(j, 0) + U(i, 1) * V(j, 1) + U(i, 2) * V(j, 2);
                P(i, j) = m2 * (F(i, j) - R) + k * C(i, j);
            }
        psi = psi_sigma<DG_FCR>(S, mu, lam);
    }
}

template <int EN>
__device__ __forceinline__ double energy_density(const Mat3& F, double mu, double lam) {
    if (EN == DG_SNH) {
        double J = det3(F);
        double JmA = J - (1.0 + mu / lam);
        double f2 = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) f2 += F.m[i] * F.m[i];
        return (mu * (f2 - 3.0) + lam * JmA * JmA) / 2.0;
    } else {
        Mat3 U, V;
        double S[3];
        svd3(F, U, S, V);
        return psi_sigma<DG_FCR>(S, mu, lam);
    }
}

}  // namespace dotgpu
