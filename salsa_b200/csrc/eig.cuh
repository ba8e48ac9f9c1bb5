// Per time-frequency-bin spatial covariance, principal eigenvector and coherence (rank-1) test.
//
// Replaces, per TF bin, dataset/salsa_feature_extraction.py:99-127 of the reference:
//   Rxx = X1^T conj(X1) / 7 ; u, s, _ = svd(Rxx) ; valid = s[0] > s[1] * cond ; normalise u[:, 0].
//
// Design (everything lives in registers of ONE thread, no LAPACK-style iteration to convergence):
//  * R is 4x4 Hermitian PSD: 4 real + 6 complex numbers.  It is accumulated from the 7 frames and
//    scaled by 1/trace (eigenvectors and the ratio test are scale invariant).
//  * principal eigenvector: n_sq matrix squarings B <- B*B, then one product of B with its dominant
//    column (power iteration with exponent 2^(n_sq+1)).  A bin can only be kept when
//    lambda1 > cond * lambda2, so the contamination of the eigenvector is <= cond^-(2^(n_sq+1))
//    (5^-16 = 6.6e-12 for the default cond = 5, n_sq = 3).
//  * lambda1 = Rayleigh quotient of R at that vector.
//  * rank-1 test  lambda2 < mu := lambda1 / cond  without computing lambda2: a Householder reflector
//    built from the eigenvector deflates R to the 3x3 Hermitian block R3 whose eigenvalues are
//    lambda2..lambda4; the test is "mu*I - R3 is positive definite", decided by an un-pivoted
//    Cholesky (stable for PD matrices).  The pivots also CERTIFY the decision: with e_i = pivot_i / mu,
//    P = prod e_i >= tau certifies "pass", a non-positive pivot with |e_i| * P_before >= tau certifies
//    "fail" (interlacing), everything else is reported as ambiguous and redone in float64 by the
//    caller, where the same routine runs with T = double.
#pragma once
#include "fft.cuh"

namespace salsa {

enum EigVerdict : int { kEigFail = 0, kEigPass = 1, kEigAmbiguous = 2 };

// upper triangle of a Hermitian 4x4: o[] holds (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
template <typename T>
struct Herm4 {
    T d[4];
    Cx<T> o[6];
};

__host__ __device__ constexpr int herm_idx(int i, int j) { return i == 0 ? j - 1 : (i == 1 ? j + 1 : 5); }

template <typename T>
__device__ __forceinline__ Cx<T> herm_at(const Herm4<T>& A, int i, int j) {
    if (i == j) return {A.d[i], (T)0};
    if (i < j) return A.o[herm_idx(i, j)];
    const Cx<T> c = A.o[herm_idx(j, i)];
    return {c.re, -c.im};
}

template <typename T> __device__ __forceinline__ T cabs2(Cx<T> a) { return a.re * a.re + a.im * a.im; }
// a * conj(b)
template <typename T> __device__ __forceinline__ Cx<T> cmulc(Cx<T> a, Cx<T> b) {
    return {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im};
}

// Fused accumulations: every complex multiply-add is four FMAs (the compiler may not reassociate
// "acc += a * b" into them on its own).
template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
// acc += a * b
template <typename T> __device__ __forceinline__ void cmac(Cx<T>& acc, Cx<T> a, Cx<T> b) {
    acc.re = fma_t<T>(a.re, b.re, acc.re);
    acc.re = fma_t<T>(-a.im, b.im, acc.re);
    acc.im = fma_t<T>(a.re, b.im, acc.im);
    acc.im = fma_t<T>(a.im, b.re, acc.im);
}
// acc += a * conj(b)
template <typename T> __device__ __forceinline__ void cmacc(Cx<T>& acc, Cx<T> a, Cx<T> b) {
    acc.re = fma_t<T>(a.re, b.re, acc.re);
    acc.re = fma_t<T>(a.im, b.im, acc.re);
    acc.im = fma_t<T>(a.im, b.re, acc.im);
    acc.im = fma_t<T>(-a.re, b.im, acc.im);
}
// acc += |a|^2
template <typename T> __device__ __forceinline__ void abs2_acc(T& acc, Cx<T> a) {
    acc = fma_t<T>(a.re, a.re, acc);
    acc = fma_t<T>(a.im, a.im, acc);
}

template <typename T>
__device__ __forceinline__ void herm_zero(Herm4<T>& A) {
#pragma unroll
    for (int i = 0; i < 4; ++i) A.d[i] = (T)0;
#pragma unroll
    for (int i = 0; i < 6; ++i) A.o[i] = {(T)0, (T)0};
}

// A += x x^H  (x: one frame, 4 channels; R[i][j] = sum_f X[f,i] conj(X[f,j]), reference :100)
template <typename T>
__device__ __forceinline__ void herm_rank1(Herm4<T>& A, const Cx<T> (&x)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        abs2_acc(A.d[i], x[i]);
#pragma unroll
        for (int j = i + 1; j < 4; ++j) cmacc(A.o[herm_idx(i, j)], x[i], x[j]);
    }
}

template <typename T>
__device__ __forceinline__ void herm_scale(Herm4<T>& A, T s) {
#pragma unroll
    for (int i = 0; i < 4; ++i) A.d[i] *= s;
#pragma unroll
    for (int i = 0; i < 6; ++i) { A.o[i].re *= s; A.o[i].im *= s; }
}

template <typename T>
__device__ __forceinline__ T herm_trace(const Herm4<T>& A) { return (A.d[0] + A.d[1]) + (A.d[2] + A.d[3]); }

// B = A * A (Hermitian).  Written out so that the real diagonal never enters a complex product:
//   B_ii = d_i^2 + sum_{k != i} |a_ik|^2
//   B_ij = (d_i + d_j) a_ij + sum_{k != i,j} a_ik a_kj
template <typename T>
__device__ __forceinline__ Herm4<T> herm_square(const Herm4<T>& A) {
    Herm4<T> B;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T s = A.d[i] * A.d[i];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k != i) abs2_acc(s, herm_at(A, i, k));
        B.d[i] = s;
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            const Cx<T> aij = A.o[herm_idx(i, j)];
            const T dd = A.d[i] + A.d[j];
            Cx<T> acc = {dd * aij.re, dd * aij.im};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k == i || k == j) continue;
                cmac(acc, herm_at(A, i, k), herm_at(A, k, j));
            }
            B.o[herm_idx(i, j)] = acc;
        }
    }
    return B;
}

// y = A x
template <typename T>
__device__ __forceinline__ void herm_matvec(const Herm4<T>& A, const Cx<T> (&x)[4], Cx<T> (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        Cx<T> acc = {A.d[i] * x[i].re, A.d[i] * x[i].im};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k == i) continue;
            cmac(acc, herm_at(A, i, k), x[k]);
        }
        y[i] = acc;
    }
}

template <typename T> __device__ __forceinline__ T eps_of();
template <> __device__ __forceinline__ float eps_of<float>() { return 1.1920929e-7f; }
template <> __device__ __forceinline__ double eps_of<double>() { return 2.220446049250313e-16; }
template <typename T> __device__ __forceinline__ T rsqrt_t(T x);
template <> __device__ __forceinline__ float rsqrt_t<float>(float x) { return rsqrtf(x); }
template <> __device__ __forceinline__ double rsqrt_t<double>(double x) { return 1.0 / sqrt(x); }

// Principal eigenvector of the (un-normalised) covariance R, coherence verdict.
//   n_sq       number of squarings
//   test       apply the rank-1 test (reference: only when is_tracking, :111-112)
//   cond       condition-number threshold (s[0] > s[1] * cond)
//   tau        certification margin (see header comment)
// Returns the verdict; `v` receives the eigenvector (arbitrary phase and scale ~1).
// A zero matrix returns kEigFail with v = e0 (the SVD of the zero matrix gives u = I).
template <typename T>
__device__ __forceinline__ int principal_eigenvector(const Herm4<T>& Rin, int n_sq, bool test, T cond, T tau,
                                                     Cx<T> (&v)[4]) {
    Herm4<T> R = Rin;
    const T tr = herm_trace(R);
    if (!(tr > (T)0)) {
        v[0] = {(T)1, (T)0};
        v[1] = v[2] = v[3] = {(T)0, (T)0};
        return kEigFail;
    }
    herm_scale(R, (T)1 / tr);
    // B = R^(2^n_sq); entries stay in range for two squarings of a trace-1 matrix (trace >= 1/64),
    // so the trace is renormalised every second squaring only
    Herm4<T> B = R;
    for (int it = 0; it < n_sq; ++it) {
        B = herm_square(B);
        if (it & 1) herm_scale(B, (T)1 / herm_trace(B));
    }
    // column with the largest diagonal entry of B ~ lambda^m v v^H, multiplied by B once more:
    // v ~ R^(2^(n_sq+1)) e_p
    int p = 0;
    T best = B.d[0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (B.d[i] > best) { best = B.d[i]; p = i; }
    {
        Cx<T> c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            Cx<T> c0 = herm_at(B, i, 0), c1 = herm_at(B, i, 1), c2 = herm_at(B, i, 2), c3 = herm_at(B, i, 3);
            c[i] = p == 0 ? c0 : (p == 1 ? c1 : (p == 2 ? c2 : c3));
        }
        const T cs = (T)1 / best;            // keeps the product in range: |c_i| <= 1 afterwards
#pragma unroll
        for (int i = 0; i < 4; ++i) { c[i].re *= cs; c[i].im *= cs; }
        herm_matvec(B, c, v);
    }
    T nv = (cabs2(v[0]) + cabs2(v[1])) + (cabs2(v[2]) + cabs2(v[3]));
    const T inv = rsqrt_t<T>(nv);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i].re *= inv; v[i].im *= inv; }
    if (!test) return kEigPass;

    // lambda1 (of the trace-normalised R) and the deflated 3x3 block
    Cx<T> rv[4];
    herm_matvec(R, v, rv);
    T lam1 = (T)0;
#pragma unroll
    for (int i = 0; i < 4; ++i) lam1 += v[i].re * rv[i].re + v[i].im * rv[i].im;   // Re(v^H R v), |v| = 1
    if (cond < (T)1) return lam1 > (T)0 ? kEigPass : kEigFail;                      // s1*cond < s0 always
    const T mu = lam1 / cond;
    // Shortcut: R has trace 1 and is PSD, so lambda2 <= 1 - lambda1, and the Rayleigh quotient lam1 is
    // a lower bound of lambda1 (exact to O(eps^2) at the converged vector).  A bin dominated by one
    // source passes here without the deflation below: lambda2 <= 1 - lam1 <= (1 - tau) mu.
    if (((T)1 - lam1) <= ((T)1 - (T)4 * tau) * mu - (T)8 * eps_of<T>()) return kEigPass;

    // Householder w = v - alpha e0, alpha = -exp(i arg v0) |v| = -exp(i arg v0)
    const T a0 = sqrt(cabs2(v[0]));
    Cx<T> ph = a0 > (T)0 ? Cx<T>{v[0].re / a0, v[0].im / a0} : Cx<T>{(T)1, (T)0};
    const Cx<T> alpha = {-ph.re, -ph.im};
    Cx<T> w[4] = {{v[0].re - alpha.re, v[0].im - alpha.im}, v[1], v[2], v[3]};
    // w^H w = |v|^2 + 2|v0| + 1 = 2 (1 + |v0|)
    const T beta = (T)1 / ((T)1 + a0);                                             // 2 / (w^H w)
    // pw = R w = R v - alpha R[:, 0]
    Cx<T> pw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const Cx<T> t = cmul(alpha, herm_at(R, i, 0));
        pw[i] = {rv[i].re - t.re, rv[i].im - t.im};
    }
    T gamma = (T)0;                                                                 // w^H R w (real)
#pragma unroll
    for (int i = 0; i < 4; ++i) gamma += w[i].re * pw[i].re + w[i].im * pw[i].im;
    // q = beta pw - (beta^2 gamma / 2) w ;  H R H = R - w q^H - q w^H
    const T hb = (T)0.5 * beta * beta * gamma;
    Cx<T> q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = {beta * pw[i].re - hb * w[i].re, beta * pw[i].im - hb * w[i].im};
    // C = mu I - R3, indices 1..3 of H R H
    T cd[3];
    Cx<T> c21, c31, c32;
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        const T r = R.d[j] - (T)2 * (w[j].re * q[j].re + w[j].im * q[j].im);
        cd[j - 1] = mu - r;
    }
    {
        auto off = [&](int j, int k) {   // (H R H)_{jk}, j > k
            const Cx<T> r = herm_at(R, j, k);
            const Cx<T> a = cmulc(w[j], q[k]);
            const Cx<T> b = cmulc(q[j], w[k]);
            return Cx<T>{-(r.re - a.re - b.re), -(r.im - a.im - b.im)};
        };
        c21 = off(2, 1);
        c31 = off(3, 1);
        c32 = off(3, 2);
    }
    // certified un-pivoted Cholesky of C (pivots scaled by 1/mu)
    const T imu = (T)1 / mu;
    T P = (T)1;
    const T e1 = cd[0] * imu;
    if (!(e1 > (T)0)) return (-e1 * P >= tau) ? kEigFail : kEigAmbiguous;
    P *= e1;
    const T id1 = (T)1 / cd[0];
    const T d2 = cd[1] - cabs2(c21) * id1;
    const T e2 = d2 * imu;
    if (!(e2 > (T)0)) return (-e2 * P >= tau) ? kEigFail : kEigAmbiguous;
    P *= e2;
    const Cx<T> t0 = cmulc(c31, c21);                    // c31 conj(c21)
    const Cx<T> t = {c32.re - t0.re * id1, c32.im - t0.im * id1};
    const T d3 = cd[2] - cabs2(c31) * id1 - cabs2(t) / d2;
    const T e3 = d3 * imu;
    if (!(e3 > (T)0)) return (-e3 * P >= tau) ? kEigFail : kEigAmbiguous;
    P *= e3;
    return P >= tau ? kEigPass : kEigAmbiguous;
}

// FOA: Re(u[1:] / u[0]) normalised to unit length (reference :118-120).
template <typename T>
__device__ __forceinline__ void normalise_foa(const Cx<T> (&v)[4], float (&out)[3]) {
    T n[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) n[i] = v[i + 1].re * v[0].re + v[i + 1].im * v[0].im;   // Re(v_i conj v_0)
    const T s = rsqrt_t<T>(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = (float)(n[i] * s);
}

}  // namespace salsa
