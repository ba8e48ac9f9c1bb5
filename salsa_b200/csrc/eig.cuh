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
#include <math.h>
#include <stdint.h>

#include "fft.cuh"
#include "salsa_b200.h"

// The per-bin arithmetic is host + device: tests/host_eig.cu runs it on the CPU against the oracle (masks and
// verdict certification can be checked over whole clips without a GPU).
#define SALSA_HD __host__ __device__ __forceinline__

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
SALSA_HD Cx<T> herm_at(const Herm4<T>& A, int i, int j) {
    if (i == j) return {A.d[i], (T)0};
    if (i < j) return A.o[herm_idx(i, j)];
    const Cx<T> c = A.o[herm_idx(j, i)];
    return {c.re, -c.im};
}

template <typename T> SALSA_HD T cabs2(Cx<T> a) { return a.re * a.re + a.im * a.im; }
// a * conj(b)
template <typename T> SALSA_HD Cx<T> cmulc(Cx<T> a, Cx<T> b) {
    return {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im};
}

// Fused accumulations: every complex multiply-add is four FMAs (the compiler may not reassociate
// "acc += a * b" into them on its own).
template <typename T> SALSA_HD T fma_t(T a, T b, T c);
template <> SALSA_HD float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> SALSA_HD double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
// acc += a * b
template <typename T> SALSA_HD void cmac(Cx<T>& acc, Cx<T> a, Cx<T> b) {
    acc.re = fma_t<T>(a.re, b.re, acc.re);
    acc.re = fma_t<T>(-a.im, b.im, acc.re);
    acc.im = fma_t<T>(a.re, b.im, acc.im);
    acc.im = fma_t<T>(a.im, b.re, acc.im);
}
// acc += a * conj(b)
template <typename T> SALSA_HD void cmacc(Cx<T>& acc, Cx<T> a, Cx<T> b) {
    acc.re = fma_t<T>(a.re, b.re, acc.re);
    acc.re = fma_t<T>(a.im, b.im, acc.re);
    acc.im = fma_t<T>(a.im, b.re, acc.im);
    acc.im = fma_t<T>(-a.re, b.im, acc.im);
}
// acc += conj(a) * b
template <typename T> SALSA_HD void cmaccj(Cx<T>& acc, Cx<T> a, Cx<T> b) {
    acc.re = fma_t<T>(a.re, b.re, acc.re);
    acc.re = fma_t<T>(a.im, b.im, acc.re);
    acc.im = fma_t<T>(a.re, b.im, acc.im);
    acc.im = fma_t<T>(-a.im, b.re, acc.im);
}
// acc.re += a.re^2 ; acc.im += a.im^2   (|a|^2 accumulated as a pair; the two halves are added once at the end)
template <typename T> SALSA_HD void abs2_pair_acc(Cx<T>& acc, Cx<T> a) {
    acc.re = fma_t<T>(a.re, a.re, acc.re);
    acc.im = fma_t<T>(a.im, a.im, acc.im);
}
// acc.re += a.re b.re ; acc.im += a.im b.im   (Re(conj(a) b) accumulated as a pair)
template <typename T> SALSA_HD void dot_pair_acc(Cx<T>& acc, Cx<T> a, Cx<T> b) {
    acc.re = fma_t<T>(a.re, b.re, acc.re);
    acc.im = fma_t<T>(a.im, b.im, acc.im);
}
// a * s (s real)
template <typename T> SALSA_HD Cx<T> cscale(Cx<T> a, T s) { return {a.re * s, a.im * s}; }

#ifdef __CUDA_ARCH__
// float32 on the device: a complex number is one 64-bit register pair and a complex multiply-add is TWO packed
// instructions (FFMA2: fma.rn.f32x2 -- the pair swap, the per-half negation and the scalar broadcast of the operands
// are operand modifiers of the instruction, ptxas folds the mov.b64 packs below into them).  sm_100 issues a
// three-register FFMA every other cycle per scheduler; the packed form does two per issue, and each half is the same
// IEEE fma in the same order as the generic code above: results are bit-identical.
SALSA_HD uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
SALSA_HD Cx<float> unpk2(uint64_t v) {
    Cx<float> r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.re), "=f"(r.im) : "l"(v));
    return r;
}
SALSA_HD uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
SALSA_HD uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
template <> SALSA_HD void cmac<float>(Cx<float>& acc, Cx<float> a, Cx<float> b) {
    uint64_t r = pk2(acc.re, acc.im);
    r = fma2(pk2(b.re, b.im), pk2(a.re, a.re), r);       // re += a.re b.re ; im += a.re b.im
    r = fma2(pk2(b.im, -b.re), pk2(-a.im, -a.im), r);    // re -= a.im b.im ; im += a.im b.re  (modifier-only form)
    acc = unpk2(r);
}
template <> SALSA_HD void cmaccj<float>(Cx<float>& acc, Cx<float> a, Cx<float> b) {
    uint64_t r = pk2(acc.re, acc.im);
    r = fma2(pk2(b.re, b.im), pk2(a.re, a.re), r);       // re += a.re b.re ; im += a.re b.im
    r = fma2(pk2(b.im, -b.re), pk2(a.im, a.im), r);      // re += a.im b.im ; im -= a.im b.re
    acc = unpk2(r);
}
template <> SALSA_HD void cmacc<float>(Cx<float>& acc, Cx<float> a, Cx<float> b) {
    uint64_t r = pk2(acc.re, acc.im);
    r = fma2(pk2(a.re, a.im), pk2(b.re, b.re), r);       // re += a.re b.re ; im += a.im b.re
    r = fma2(pk2(a.im, -a.re), pk2(b.im, b.im), r);      // re += a.im b.im ; im -= a.re b.im
    acc = unpk2(r);
}
template <> SALSA_HD void abs2_pair_acc<float>(Cx<float>& acc, Cx<float> a) {
    const uint64_t p = pk2(a.re, a.im);
    acc = unpk2(fma2(p, p, pk2(acc.re, acc.im)));
}
template <> SALSA_HD void dot_pair_acc<float>(Cx<float>& acc, Cx<float> a, Cx<float> b) {
    acc = unpk2(fma2(pk2(a.re, a.im), pk2(b.re, b.im), pk2(acc.re, acc.im)));
}
template <> SALSA_HD Cx<float> cscale<float>(Cx<float> a, float s) { return unpk2(mul2(pk2(a.re, a.im), pk2(s, s))); }
#endif

template <typename T>
SALSA_HD void herm_zero(Herm4<T>& A) {
#pragma unroll
    for (int i = 0; i < 4; ++i) A.d[i] = (T)0;
#pragma unroll
    for (int i = 0; i < 6; ++i) A.o[i] = {(T)0, (T)0};
}

// A += x x^H  (x: one frame, 4 channels; R[i][j] = sum_f X[f,i] conj(X[f,j]), reference :100).  The diagonal is kept
// as pairs (sum re^2, sum im^2) in `dg` while frames are accumulated; herm_close_diag() adds the halves.
template <typename T>
SALSA_HD void herm_rank1(Herm4<T>& A, Cx<T> (&dg)[4], const Cx<T> (&x)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        abs2_pair_acc(dg[i], x[i]);
#pragma unroll
        for (int j = i + 1; j < 4; ++j) cmacc(A.o[herm_idx(i, j)], x[i], x[j]);
    }
}
template <typename T>
SALSA_HD void herm_close_diag(Herm4<T>& A, const Cx<T> (&dg)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) A.d[i] = dg[i].re + dg[i].im;
}

template <typename T>
SALSA_HD void herm_scale(Herm4<T>& A, T s) {
#pragma unroll
    for (int i = 0; i < 4; ++i) A.d[i] *= s;
#pragma unroll
    for (int i = 0; i < 6; ++i) A.o[i] = cscale(A.o[i], s);
}

template <typename T>
SALSA_HD T herm_trace(const Herm4<T>& A) { return (A.d[0] + A.d[1]) + (A.d[2] + A.d[3]); }

// acc += A(i,k) * b for an off-diagonal entry of the Hermitian A (the conjugate of the stored upper triangle is
// never materialised: it is a different multiply-add)
template <typename T>
SALSA_HD void herm_mac(Cx<T>& acc, const Herm4<T>& A, int i, int k, Cx<T> b) {
    if (i < k) cmac(acc, A.o[herm_idx(i, k)], b);
    else cmaccj(acc, A.o[herm_idx(k, i)], b);
}

// B = A * A (Hermitian).  Written out so that the real diagonal never enters a complex product:
//   B_ii = d_i^2 + sum_{k != i} |a_ik|^2
//   B_ij = (d_i + d_j) a_ij + sum_{k != i,j} a_ik a_kj
template <typename T>
SALSA_HD Herm4<T> herm_square(const Herm4<T>& A) {
    Herm4<T> B;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        Cx<T> s = {(T)0, (T)0};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k != i) abs2_pair_acc(s, A.o[k > i ? herm_idx(i, k) : herm_idx(k, i)]);
        B.d[i] = fma_t<T>(A.d[i], A.d[i], s.re + s.im);
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            Cx<T> acc = cscale(A.o[herm_idx(i, j)], A.d[i] + A.d[j]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k == i || k == j) continue;
                if (k < i) cmaccj(acc, A.o[herm_idx(k, i)], A.o[herm_idx(k, j)]);        // conj(a_ki) a_kj
                else if (k < j) cmac(acc, A.o[herm_idx(i, k)], A.o[herm_idx(k, j)]);      // a_ik a_kj
                else cmacc(acc, A.o[herm_idx(i, k)], A.o[herm_idx(j, k)]);                // a_ik conj(a_jk)
            }
            B.o[herm_idx(i, j)] = acc;
        }
    }
    return B;
}

// y = A x
template <typename T>
SALSA_HD void herm_matvec(const Herm4<T>& A, const Cx<T> (&x)[4], Cx<T> (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        Cx<T> acc = cscale(x[i], A.d[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k == i) continue;
            herm_mac(acc, A, i, k, x[k]);
        }
        y[i] = acc;
    }
}

template <typename T> SALSA_HD T eps_of();
template <> SALSA_HD float eps_of<float>() { return 1.1920929e-7f; }
template <> SALSA_HD double eps_of<double>() { return 2.220446049250313e-16; }
template <typename T> SALSA_HD T rsqrt_t(T x);
template <> SALSA_HD float rsqrt_t<float>(float x) {
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
// reciprocal for pure rescalings (the result only has to be within a few ulp): MUFU.RCP on the device
template <typename T> SALSA_HD T rcp_scale(T x);
template <> SALSA_HD float rcp_scale<float>(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
template <> SALSA_HD double rcp_scale<double>(double x) { return 1.0 / x; }
// fourth root of a positive number to a few ulp (two MUFU.RSQ on the device)
template <typename T> SALSA_HD T root4(T x);
template <> SALSA_HD float root4<float>(float x) {
#ifdef __CUDA_ARCH__
    const float s = x * rsqrtf(x);
    return s * rsqrtf(s);
#else
    return sqrtf(sqrtf(x));
#endif
}
template <> SALSA_HD double root4<double>(double x) { return sqrt(sqrt(x)); }
// square root of a positive number to a few ulp
template <typename T> SALSA_HD T sqrt_fast(T x);
template <> SALSA_HD float sqrt_fast<float>(float x) {
#ifdef __CUDA_ARCH__
    return x * rsqrtf(x);
#else
    return sqrtf(x);
#endif
}
template <> SALSA_HD double sqrt_fast<double>(double x) { return sqrt(x); }
template <> SALSA_HD double rsqrt_t<double>(double x) { return 1.0 / sqrt(x); }

// Principal eigenvector of the (un-normalised) covariance R, coherence verdict.
//   NSQ        number of squarings when > 0 (compile time), else the run-time value n_sq_rt
//   test       apply the rank-1 test (reference: only when is_tracking, :111-112)
//   cond       condition-number threshold (s[0] > s[1] * cond)
//   tau        certification margin (see header comment)
// Returns the verdict; `v` receives the eigenvector (arbitrary phase and scale ~1).
// A zero matrix returns kEigFail with v = e0 (the SVD of the zero matrix gives u = I).
template <typename T, int NSQ>
SALSA_HD int principal_eigenvector(const Herm4<T>& Rin, int n_sq_rt, bool second_product, bool test, T cond, T tau, Cx<T> (&v)[4]) {
    Herm4<T> R = Rin;
    const T tr = herm_trace(R);
    if (!(tr > (T)0)) {
        v[0] = {(T)1, (T)0};
        v[1] = v[2] = v[3] = {(T)0, (T)0};
        return kEigFail;
    }
    herm_scale(R, rcp_scale<T>(tr));
    // B = R^(2^n_sq); entries stay in range for two squarings of a trace-1 matrix (trace >= 1/64),
    // so the trace is renormalised every second squaring only (and not at all when two are all there is)
    Herm4<T> B = herm_square(R);
    const T frob2 = herm_trace(B);                      // ||R||_F^2 = sum of lambda_i^2
    if (NSQ > 0) {
#pragma unroll
        for (int it = 1; it < NSQ; ++it) {
            B = herm_square(B);
            if ((it & 1) && it + 1 < NSQ) herm_scale(B, rcp_scale<T>(herm_trace(B)));
        }
    } else {
        for (int it = 1; it < n_sq_rt; ++it) {
            B = herm_square(B);
            if (it & 1) herm_scale(B, rcp_scale<T>(herm_trace(B)));
        }
    }
    // column with the largest diagonal entry of B ~ lambda^m v v^H, multiplied by B once (or twice) more:
    // v ~ R^(2 * 2^n_sq) e_p   (R^(3 * 2^n_sq) e_p)
    T lam1_pow4 = (T)-1;
    int p = 0;
    T best = B.d[0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (B.d[i] > best) { best = B.d[i]; p = i; }
    {
        const bool p1 = (p & 1) != 0, p2 = (p & 2) != 0;
        auto sel = [&](T a0, T a1, T a2, T a3) -> T {
            const T lo = p1 ? a1 : a0, hi = p1 ? a3 : a2;
            return p2 ? hi : lo;
        };
        const T cs = rcp_scale<T>(best);     // keeps the product in range: |c_i| <= 1 afterwards
        Cx<T> c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const Cx<T> c0 = herm_at(B, i, 0), c1 = herm_at(B, i, 1), c2 = herm_at(B, i, 2), c3 = herm_at(B, i, 3);
            c[i] = {sel(c0.re, c1.re, c2.re, c3.re) * cs, sel(c0.im, c1.im, c2.im, c3.im) * cs};
        }
        herm_matvec(B, c, v);
        if (!second_product && NSQ == 2) {
            // B = R^4 exactly here (no rescaling between two squarings) and c = R^4 e_p / B_pp, v = B c: the Rayleigh quotient
            // of B at c is a lower bound of lambda1^4, exact to O((lambda2/lambda1)^8) -- 2.6e-6 at the coherence threshold 5,
            // i.e. 6e-7 in lambda1, two orders below the certification margin 4 tau of the rank-1 test
            Cx<T> num = {(T)0, (T)0}, den = {(T)0, (T)0};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dot_pair_acc(num, c[i], v[i]);
                abs2_pair_acc(den, c[i]);
            }
            lam1_pow4 = (num.re + num.im) * rcp_scale<T>(den.re + den.im);
        }
        if (second_product) {                // exponent 3 * 2^n_sq instead of 2 * 2^n_sq for half the price of a squaring
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i] = v[i];
            herm_matvec(B, c, v);
            if (NSQ == 2) {
                // B = R^4 exactly here (no rescaling between two squarings), so the Rayleigh quotient of B at
                // c = R^8 e_p is a lower bound of lambda1^4, exact to O((lambda2/lambda1)^16): lambda1 without the
                // product R v that the Rayleigh quotient of R would cost
                Cx<T> num = {(T)0, (T)0}, den = {(T)0, (T)0};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dot_pair_acc(num, c[i], v[i]);
                    abs2_pair_acc(den, c[i]);
                }
                lam1_pow4 = (num.re + num.im) * rcp_scale<T>(den.re + den.im);
            }
        }
    }
    T nv = (cabs2(v[0]) + cabs2(v[1])) + (cabs2(v[2]) + cabs2(v[3]));
    const T inv = rsqrt_t<T>(nv);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = cscale(v[i], inv);
    if (!test) return kEigPass;

    // lambda1 (of the trace-normalised R)
    Cx<T> rv[4];
    T lam1;
    const bool have_rv = !(lam1_pow4 > (T)0);
    if (have_rv) {
        herm_matvec(R, v, rv);
        lam1 = (T)0;
#pragma unroll
        for (int i = 0; i < 4; ++i) lam1 += v[i].re * rv[i].re + v[i].im * rv[i].im;   // Re(v^H R v), |v| = 1
    } else {
        lam1 = root4<T>(lam1_pow4);
    }
    if (cond < (T)1) return lam1 > (T)0 ? kEigPass : kEigFail;                      // s1*cond < s0 always
    const T mu = lam1 / cond;
    // Shortcuts from the Frobenius norm.  rest = lambda2^2 + lambda3^2 + lambda4^2 = ||R||_F^2 - lambda1^2, and the
    // Rayleigh quotient lam1 is a lower bound of lambda1 (exact to O(eps^2) at the converged vector), so
    //   lambda2^2 <= rest            : rest < mu^2       certifies lambda2 < mu   (pass; a bin dominated by one source)
    //   lambda2^2 >= rest / 3        : rest >= 3 mu^2    certifies lambda2 >= mu  (fail; a bin shared by two sources)
    // with `slack` covering the rounding of frob2 and lam1^2 and the margin tau.  Only bins in between pay for the
    // deflation below.  (A poorly converged lam1 -- eigenvalue ratio near 1 -- only inflates rest: towards "fail",
    // which is the true verdict there.)
    const T rest = frob2 - lam1 * lam1;
    const T mu2 = mu * mu;
    const T slack = (T)64 * eps_of<T>() * frob2;
    if (rest + slack <= ((T)1 - (T)4 * tau) * mu2) return kEigPass;
    if (rest - slack >= (T)3 * ((T)1 + (T)4 * tau) * mu2) return kEigFail;

    // the deflated 3x3 block.  Householder w = v - alpha e0, alpha = -exp(i arg v0) |v| = -exp(i arg v0)
    if (!have_rv) herm_matvec(R, v, rv);
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
SALSA_HD void normalise_foa(const Cx<T> (&v)[4], float (&out)[3]) {
    T n[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) n[i] = v[i + 1].re * v[0].re + v[i + 1].im * v[0].im;   // Re(v_i conj v_0)
    const T s = rsqrt_t<T>(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = (float)(n[i] * s);
}

constexpr int kHop = 3;                  // n_hopframes of the reference ("do not change")
constexpr int kWin = 2 * kHop + 1;       // 7 frames per covariance

SALSA_HD float quiet_nan() {
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7fc00000);
#else
    return nanf("");
#endif
}

// ------------------------------------------------------------------------------------------------
// The eigenvector step for one TF bin.  `load(k, ch)` returns X[frame t - 3 + k][ch], k = 0..6.
// ------------------------------------------------------------------------------------------------
struct EigArgs {
    int format;          // SALSA_FORMAT_*
    int test;            // apply the coherence test (is_tracking)
    int n_sq;            // squarings, float32 path
    int n_mv;            // products of the squared matrix with its dominant column after that: 1 or 2
    float cond;
    double cond_d;
    double inv_delta;    // 1 / delta, delta = 2 pi fs / (n_fft c)   (:38-40)
    int lower;           // absolute index of spatial bin 0
};

template <typename T, typename Load>
SALSA_HD void accumulate_cov(Herm4<T>& R, Load load) {
    herm_zero(R);
    Cx<T> dg[4] = {{(T)0, (T)0}, {(T)0, (T)0}, {(T)0, (T)0}, {(T)0, (T)0}};
#pragma unroll
    for (int f = 0; f < kWin; ++f) {
        Cx<T> x[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const float2 v = load(f, ch);
            x[ch] = {(T)v.x, (T)v.y};
        }
        herm_rank1(R, dg, x);
    }
    herm_close_diag(R, dg);
}

// float64 re-evaluation of a bin whose float32 verdict could not be certified
template <typename Load>
SALSA_HD int eig_bin_f64(Load load, const EigArgs& e, float (&out)[3], int b) {
    Herm4<double> R;
    accumulate_cov<double>(R, load);
    Cx<double> v[4];
    int verdict = principal_eigenvector<double, 0>(R, e.n_sq + 2, e.n_mv > 1, e.test != 0, e.cond_d, 0.0, v);
    if (verdict == kEigAmbiguous) verdict = kEigFail;
    if (verdict == kEigPass) {
        if (e.format == SALSA_FORMAT_FOA) {
            normalise_foa(v, out);
        } else {
            const double s = e.inv_delta / (double)(b + e.lower);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const Cx<double> p = cmulc(v[i + 1], v[0]);
                out[i] = (float)(atan2(p.im, p.re) * s);
            }
        }
    }
    return verdict;
}

// float32 evaluation of one bin.  Returns kEigPass (out[] holds the three spatial features), kEigFail (out[] is
// zero) or kEigAmbiguous (the float32 verdict could not be certified: the caller re-evaluates in float64).
template <int NSQ, typename Load>
SALSA_HD int eig_bin_f32(Load load, const EigArgs& e, int b, float (&out)[3]) {
    out[0] = out[1] = out[2] = 0.0f;
    Herm4<float> R;
    accumulate_cov<float>(R, load);
    Cx<float> v[4];
    const int verdict = principal_eigenvector<float, NSQ>(R, e.n_sq, e.n_mv > 1, e.test != 0, e.cond, 1e-4f, v);
    if (verdict == kEigAmbiguous) return verdict;
    if (!e.test && !(herm_trace(R) > 0.0f)) {
        // is_tracking=False on an all-zero bin: svd gives u = I, so FOA divides 0 by 0 (NaN) and
        // MIC yields angle(0) = 0, exactly as the reference does.
        if (e.format == SALSA_FORMAT_FOA) out[0] = out[1] = out[2] = quiet_nan();
        return kEigPass;
    }
    if (verdict != kEigPass) return kEigFail;
    if (e.format == SALSA_FORMAT_FOA) {
        normalise_foa(v, out);
    } else {
        const float s = (float)(e.inv_delta / (double)(b + e.lower));
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const Cx<float> p = cmulc(v[i + 1], v[0]);
            out[i] = atan2f(p.im, p.re) * s;
        }
    }
    return kEigPass;
}

// Returns true when the bin is valid; out[] holds the three spatial features (zeros otherwise).
template <typename Load>
SALSA_HD bool eig_bin(Load load, const EigArgs& e, int b, float (&out)[3]) {
    int verdict = eig_bin_f32<0>(load, e, b, out);
    if (verdict == kEigAmbiguous) verdict = eig_bin_f64(load, e, out, b);
    return verdict == kEigPass;
}

}  // namespace salsa
