/*
 * sf_ref_shim.h — minimal stand-ins for the third-party APIs the REFERENCE's hot-path sources use
 * (Eigen 3 dense, the MRPT 1.x Eigen plugin / CPose3D / CTicTac, a sliver of cv::Mat), so that
 *     /root/reference/KMeans.cpp, SegmentationBackground.cpp and the solver line ranges of FrontEnd.cpp,
 * together with the reference's own StaticFusion.h, compile UNMODIFIED from where they lie
 * (oracle/Makefile target `ref`).  TEST INFRASTRUCTURE: it exists to check the oracle's restatement
 * against the reference's own logic, stage by stage and bit for bit.
 *
 * None of these libraries is present in the build image, and their sources are not in the reference
 * tree, so this is NOT Eigen: every operation is evaluated eagerly, in the scalar type of its operands,
 * with SEQUENTIAL summation in column-major order.  The factorisations are the documented stand-ins the
 * oracle's reference-literal policy uses too: unpivoted LDL^T with zero-pivot handling for
 * `ldlt()`, Gauss-Jordan with partial pivoting for `inverse()`, Gaussian elimination with partial pivoting
 * for `colPivHouseholderQr().solve()`, cyclic Jacobi (unsorted) for `SelfAdjointEigenSolver`, closed-form
 * SE(3) exp / log for `MatrixFunctions` (valid for the 4x4 twist / rigid matrices the reference passes).
 * Only the members the reference actually calls exist.
 */
#ifndef SF_REF_SHIM_H
#define SF_REF_SHIM_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1 };
const int Infinity = -1;
enum ComputationInfo { Success = 0, NumericalIssue = 1 };
typedef long Index;

template <class T> struct StoreOf { typedef T type; };
template <> struct StoreOf<bool> { typedef unsigned char type; };

template <class T> class ArrayD;
template <class T, int R, int C, int O, int MR, int MC> class Matrix;

namespace shim {
/* ---- small dense algebra shared with the oracle's reference-literal policy (same operation sequences) ---- */
template <class T> int ldlt_factor(int n, T* A, unsigned char* zero) {
    const T tiny = (T)1e-20;
    int nz = 0;
    for (int j = 0; j < n; j++) {
        T d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= (A[j * n + k] * A[j * n + k]) * A[k * n + k];
        A[j * n + j] = d;
        if (!(d > tiny)) { zero[j] = 1; nz++; for (int i = j + 1; i < n; i++) A[i * n + j] = (T)0; continue; }
        zero[j] = 0;
        for (int i = j + 1; i < n; i++) {
            T s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= (A[i * n + k] * A[j * n + k]) * A[k * n + k];
            A[i * n + j] = s / d;
        }
    }
    return nz;
}
template <class T> void ldlt_solve(int n, const T* A, const unsigned char* zero, const T* b, T* x) {
    for (int i = 0; i < n; i++) { T s = b[i]; for (int k = 0; k < i; k++) s -= A[i * n + k] * x[k]; x[i] = s; }
    for (int i = 0; i < n; i++) x[i] = zero[i] ? (T)0 : x[i] / A[i * n + i];
    for (int i = n - 1; i >= 0; i--) { T s = x[i]; for (int k = n - 1; k > i; k--) s -= A[k * n + i] * x[k]; x[i] = s; }
}
/* Gaussian elimination with partial pivoting on [A | B] (row-major n x n, n x m); returns false when singular */
template <class T> bool gauss_solve(int n, int m, T* A, T* B) {
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++) if (std::fabs(A[r * n + c]) > std::fabs(A[p * n + c])) p = r;
        if (A[p * n + c] == (T)0) return false;
        if (p != c) {
            for (int k = 0; k < n; k++) std::swap(A[p * n + k], A[c * n + k]);
            for (int k = 0; k < m; k++) std::swap(B[p * m + k], B[c * m + k]);
        }
        for (int r = c + 1; r < n; r++) {
            const T f = A[r * n + c] / A[c * n + c];
            if (f == (T)0) continue;
            for (int k = c; k < n; k++) A[r * n + k] -= f * A[c * n + k];
            for (int k = 0; k < m; k++) B[r * m + k] -= f * B[c * m + k];
        }
    }
    for (int k = 0; k < m; k++)
        for (int r = n - 1; r >= 0; r--) {
            T s = B[r * m + k];
            for (int c = r + 1; c < n; c++) s -= A[r * n + c] * B[c * m + k];
            B[r * m + k] = s / A[r * n + r];
        }
    return true;
}
template <class T> void jacobi_eig(int n, const T* Ain, T* ev, T* V) {
    std::vector<T> A(Ain, Ain + n * n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? (T)1 : (T)0;
    for (int sweep = 0; sweep < 30; sweep++) {
        T off = 0, diag = 0;
        for (int p = 0; p < n; p++) { diag += A[p * n + p] * A[p * n + p]; for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q]; }
        if (!(off > (T)1e-32 * diag)) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const T apq = A[p * n + q];
                if (apq == (T)0) continue;
                const T theta = (A[q * n + q] - A[p * n + p]) / ((T)2 * apq);
                const T t = (theta >= (T)0 ? (T)1 : (T)-1) / (std::fabs(theta) + std::sqrt(theta * theta + (T)1));
                const T c = (T)1 / std::sqrt(t * t + (T)1);
                const T s = t * c;
                for (int k = 0; k < n; k++) { const T akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
                for (int k = 0; k < n; k++) { const T apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
                for (int k = 0; k < n; k++) { const T vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < n; i++) ev[i] = A[i * n + i];
}
template <class T> void mat3_sq(const T* K, T* K2) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { T s = 0; for (int k = 0; k < 3; k++) s += K[i * 3 + k] * K[k * 3 + j]; K2[i * 3 + j] = s; }
}
/* row-major 4x4 in / out */
template <class T> void se3_exp(const T* H, T* M) {
    const T xi[6] = {H[3], H[7], H[11], H[9], H[2], H[4]}; /* t = last column, w from the skew part (FrontEnd.cpp:759-763) */
    const T wx = xi[3], wy = xi[4], wz = xi[5];
    const T th2 = wx * wx + wy * wy + wz * wz;
    T A, B, C;
    if (th2 < (T)0.0025) {
        A = (T)1 + th2 * ((T)-1 / 6 + th2 * ((T)1 / 120 + th2 * ((T)-1 / 5040 + th2 * ((T)1 / 362880))));
        B = (T)0.5 + th2 * ((T)-1 / 24 + th2 * ((T)1 / 720 + th2 * ((T)-1 / 40320 + th2 * ((T)1 / 3628800))));
        C = (T)1 / 6 + th2 * ((T)-1 / 120 + th2 * ((T)1 / 5040 + th2 * ((T)-1 / 362880 + th2 * ((T)1 / 39916800))));
    } else {
        const T th = std::sqrt(th2); const T sh = std::sin((T)0.5 * th);
        A = std::sin(th) / th; B = (T)2 * sh * sh / th2; C = (th - std::sin(th)) / (th2 * th);
    }
    const T K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    T K2[9]; mat3_sq(K, K2);
    T R[9], Vm[9];
    for (int i = 0; i < 9; i++) { const T I = (i % 4 == 0) ? (T)1 : (T)0; R[i] = I + A * K[i] + B * K2[i]; Vm[i] = I + B * K[i] + C * K2[i]; }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) M[i * 4 + j] = R[i * 3 + j];
        M[i * 4 + 3] = Vm[i * 3 + 0] * xi[0] + Vm[i * 3 + 1] * xi[1] + Vm[i * 3 + 2] * xi[2];
    }
    M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}
template <class T> void se3_log(const T* M, T* H) {
    const T sx = (T)0.5 * (M[9] - M[6]), sy = (T)0.5 * (M[2] - M[8]), sz = (T)0.5 * (M[4] - M[1]);
    const T s2 = sx * sx + sy * sy + sz * sz;
    const T c = (T)0.5 * (M[0] + M[5] + M[10] - (T)1);
    T fac, th2, D;
    if (s2 < (T)0.0025 && c > (T)0) {
        fac = (T)1 + s2 * ((T)1 / 6 + s2 * ((T)3 / 40 + s2 * ((T)15 / 336 + s2 * ((T)105 / 3456 + s2 * ((T)945 / 42240)))));
        th2 = s2 * fac * fac;
        D = (T)1 / 12 + th2 * ((T)1 / 720 + th2 * ((T)1 / 30240 + th2 * ((T)1 / 1209600)));
    } else {
        const T s = std::sqrt(s2); const T th = std::atan2(s, c);
        fac = (s > (T)0) ? th / s : (T)1; th2 = th * th;
        const T sh = std::sin((T)0.5 * th); const T Bc = (T)2 * sh * sh / th2; const T Ac = std::sin(th) / th;
        D = ((T)1 - Ac / ((T)2 * Bc)) / th2;
    }
    const T wx = fac * sx, wy = fac * sy, wz = fac * sz;
    const T K[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    T K2[9]; mat3_sq(K, K2);
    T t[3];
    for (int i = 0; i < 3; i++) {
        T acc = 0;
        for (int j = 0; j < 3; j++) { const T I = (i == j) ? (T)1 : (T)0; acc += (I - (T)0.5 * K[i * 3 + j] + D * K2[i * 3 + j]) * M[j * 4 + 3]; }
        t[i] = acc;
    }
    for (int i = 0; i < 16; i++) H[i] = 0;
    H[1] = -wz; H[4] = wz; H[2] = wy; H[8] = -wy; H[6] = -wx; H[9] = wx; H[3] = t[0]; H[7] = t[1]; H[11] = t[2];
}
}  // namespace shim

/* ------------------------------------------------------------------------------------------------
 * Dense<T>: column-major dynamic matrix with eager arithmetic
 * ---------------------------------------------------------------------------------------------- */
template <class T>
class Dense {
public:
    typedef typename StoreOf<T>::type S;
    int r_ = 0, c_ = 0;
    std::vector<S> a;

    Dense() {}
    Dense(int r, int c) : r_(r), c_(c), a((size_t)r * c, S()) {}

    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return (Index)a.size(); }
    S* data() { return a.data(); }
    const S* data() const { return a.data(); }
    S& operator()(Index i, Index j) { return a[(size_t)j * r_ + i]; }
    const S& operator()(Index i, Index j) const { return a[(size_t)j * r_ + i]; }
    S& operator()(Index k) { return a[(size_t)k]; }
    const S& operator()(Index k) const { return a[(size_t)k]; }
    S& operator[](Index k) { return a[(size_t)k]; }
    const S& operator[](Index k) const { return a[(size_t)k]; }

    void resize(Index r, Index c) { if (r != r_ || c != c_) { r_ = (int)r; c_ = (int)c; a.assign((size_t)r * c, S()); } }
    void resize(Index n) { resize(n, 1); }
    /* MRPT plugin: setSize keeps the old content and zero-fills new cells */
    void setSize(Index r, Index c) {
        Dense old = *this;
        r_ = (int)r; c_ = (int)c; a.assign((size_t)r * c, S());
        for (int j = 0; j < std::min(old.c_, c_); j++) for (int i = 0; i < std::min(old.r_, r_); i++) (*this)(i, j) = old(i, j);
    }
    void fill(T v) { std::fill(a.begin(), a.end(), (S)v); }
    void assign(T v) { fill(v); }  /* MRPT plugin */
    void setZero() { fill((T)0); }
    void setConstant(T v) { fill(v); }
    void setIdentity() { setZero(); for (int i = 0; i < std::min(r_, c_); i++) (*this)(i, i) = (S)1; }
    void swap(Dense& o) { std::swap(r_, o.r_); std::swap(c_, o.c_); a.swap(o.a); }
    Dense replicate(int, int) const { return *this; }
    Dense eval() const { return *this; }
    /* corner blocks as copies (the trajectory writers read them: Datasets.cpp:258, Reconstruction.cpp:472-473) */
    Dense topLeftCorner(Index r, Index c) const { Dense o((int)r, (int)c); for (int j = 0; j < (int)c; j++) for (int i = 0; i < (int)r; i++) o(i, j) = (*this)(i, j); return o; }
    Dense topRightCorner(Index r, Index c) const { Dense o((int)r, (int)c); for (int j = 0; j < (int)c; j++) for (int i = 0; i < (int)r; i++) o(i, j) = (*this)(i, c_ - (int)c + j); return o; }

    /* reductions: sequential, column-major */
    T sum() const { T s = 0; for (size_t k = 0; k < a.size(); k++) s += a[k]; return s; }
    T sumAll() const { return sum(); }  /* MRPT plugin */
    T maximum() const { T m = a.empty() ? T() : (T)a[0]; for (size_t k = 1; k < a.size(); k++) if (a[k] > m) m = a[k]; return m; }  /* MRPT */
    T maxCoeff() const { return maximum(); }
    T squaredNorm() const { T s = 0; for (size_t k = 0; k < a.size(); k++) s += a[k] * a[k]; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    template <int P> T lpNorm() const { T m = 0; for (size_t k = 0; k < a.size(); k++) m = std::max(m, (T)std::fabs(a[k])); return m; }  /* only Infinity is used */
    Dense cwiseAbs() const { Dense o = *this; for (auto& x : o.a) x = std::fabs(x); return o; }
    Dense transpose() const { Dense o(c_, r_); for (int j = 0; j < c_; j++) for (int i = 0; i < r_; i++) o(j, i) = (*this)(i, j); return o; }
    template <class U> Dense<U> cast() const { Dense<U> o(r_, c_); for (size_t k = 0; k < a.size(); k++) o.a[k] = (U)a[k]; return o; }
    ArrayD<T> array() const;
    const Dense& matrix() const { return *this; }

    /* MRPT plugin: this = A^T A, this = A^T B (sequential over the rows) */
    void multiply_AtA(const Dense& A) {
        resize(A.c_, A.c_);
        for (int i = 0; i < A.c_; i++) for (int j = 0; j < A.c_; j++) { T s = 0; for (int k = 0; k < A.r_; k++) s += A(k, i) * A(k, j); (*this)(i, j) = s; }
    }
    void multiply_AtB(const Dense& A, const Dense& B) {
        resize(A.c_, B.c_);
        for (int i = 0; i < A.c_; i++) for (int j = 0; j < B.c_; j++) { T s = 0; for (int k = 0; k < A.r_; k++) s += A(k, i) * B(k, j); (*this)(i, j) = s; }
    }

    /* blocks */
    template <int BR, int BC> Matrix<T, BR, BC, 0, BR, BC> block(Index i0, Index j0) const;
    struct ColRef {
        Dense& m; int j;
        ColRef& operator=(const Dense& v) { assert(v.size() == m.r_); for (int i = 0; i < m.r_; i++) m(i, j) = v.a[i]; return *this; }
        ColRef& operator=(const ColRef& o) { Dense t = o; return *this = t; }
        ColRef& operator+=(const Dense& v) { for (int i = 0; i < m.r_; i++) m(i, j) += v.a[i]; return *this; }
        ColRef& operator/=(T s) { for (int i = 0; i < m.r_; i++) m(i, j) /= s; return *this; }
        void fill(T v) { for (int i = 0; i < m.r_; i++) m(i, j) = v; }
        operator Dense() const { Dense t(m.r_, 1); for (int i = 0; i < m.r_; i++) t.a[i] = m(i, j); return t; }
        ArrayD<T> array() const;
        T squaredNorm() const { return Dense(*this).squaredNorm(); }
        friend Dense operator*(T s, const ColRef& c) { return s * Dense(c); }
        friend Dense operator-(const ColRef& x, const ColRef& y) { return Dense(x) - Dense(y); }
        friend Dense operator-(const ColRef& x, const Dense& y) { return Dense(x) - y; }
    };
    struct ConstColRef {
        const Dense& m; int j;
        operator Dense() const { Dense t(m.r_, 1); for (int i = 0; i < m.r_; i++) t.a[i] = m(i, j); return t; }
        ArrayD<T> array() const;
        friend Dense operator*(T s, const ConstColRef& c) { return s * Dense(c); }
        friend Dense operator-(const ConstColRef& x, const ConstColRef& y) { return Dense(x) - Dense(y); }
        friend Dense operator-(const ConstColRef& x, const Dense& y) { return Dense(x) - y; }
    };
    ColRef col(Index j) { return ColRef{*this, (int)j}; }
    ConstColRef col(Index j) const { return ConstColRef{*this, (int)j}; }
    struct RowRef {
        Dense& m; int i;
        RowRef& operator=(const Dense& v) { assert(v.size() == m.c_); for (int j = 0; j < m.c_; j++) m(i, j) = v.a[j]; return *this; }
        operator Dense() const { Dense t(1, m.c_); for (int j = 0; j < m.c_; j++) t.a[j] = m(i, j); return t; }
        friend Dense operator*(T s, const RowRef& r) { return s * Dense(r); }
    };
    RowRef row(Index i) { return RowRef{*this, (int)i}; }
    struct RowsRef {
        Dense& m; int i0, n;
        RowsRef& operator=(const Dense& v) { assert(v.r_ == n && v.c_ == m.c_); for (int j = 0; j < m.c_; j++) for (int i = 0; i < n; i++) m(i0 + i, j) = v(i, j); return *this; }
        operator Dense() const { Dense t(n, m.c_); for (int j = 0; j < m.c_; j++) for (int i = 0; i < n; i++) t(i, j) = m(i0 + i, j); return t; }
    };
    template <int N> RowsRef topRows() { return RowsRef{*this, 0, N}; }
    template <int N> RowsRef bottomRows() { return RowsRef{*this, r_ - N, N}; }
    Dense topRows(Index n) const { Dense t((int)n, c_); for (int j = 0; j < c_; j++) for (int i = 0; i < n; i++) t(i, j) = (*this)(i, j); return t; }
    Dense bottomRows(Index n) const { Dense t((int)n, c_); for (int j = 0; j < c_; j++) for (int i = 0; i < n; i++) t(i, j) = (*this)(r_ - n + i, j); return t; }

    /* factorisations (see the header comment for what each stands in for) */
    struct LDLT {
        int n; std::vector<T> F; std::vector<unsigned char> zero;
        Dense solve(const Dense& b) const {
            Dense x(b.r_, b.c_);
            std::vector<T> bb(n), xx(n);
            for (int c = 0; c < b.c_; c++) {
                for (int i = 0; i < n; i++) bb[i] = b(i, c);
                shim::ldlt_solve<T>(n, F.data(), zero.data(), bb.data(), xx.data());
                for (int i = 0; i < n; i++) x(i, c) = xx[i];
            }
            return x;
        }
    };
    LDLT ldlt() const {
        LDLT f; f.n = r_; f.F.resize((size_t)r_ * r_); f.zero.resize(r_);
        for (int i = 0; i < r_; i++) for (int j = 0; j < r_; j++) f.F[(size_t)i * r_ + j] = (*this)(i, j);
        shim::ldlt_factor<T>(r_, f.F.data(), f.zero.data());
        return f;
    }
    Dense inverse() const {
        const int n = r_;
        std::vector<T> A((size_t)n * n), B((size_t)n * n, (T)0);
        for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) A[(size_t)i * n + j] = (*this)(i, j); B[(size_t)i * n + i] = (T)1; }
        const bool ok = shim::gauss_solve<T>(n, n, A.data(), B.data());
        Dense o(n, n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) o(i, j) = ok ? B[(size_t)i * n + j] : std::numeric_limits<T>::quiet_NaN();
        return o;
    }
    struct QR {
        Dense A;
        Dense solve(const Dense& b) const {
            const int n = A.r_, m = b.c_;
            std::vector<T> M((size_t)n * n), B((size_t)n * m);
            for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) M[(size_t)i * n + j] = A(i, j); for (int j = 0; j < m; j++) B[(size_t)i * m + j] = b(i, j); }
            shim::gauss_solve<T>(n, m, M.data(), B.data());
            Dense x(n, m);
            for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) x(i, j) = B[(size_t)i * m + j];
            return x;
        }
    };
    QR colPivHouseholderQr() const { return QR{*this}; }
    Dense exp() const {
        assert(r_ == 4 && c_ == 4);
        T H[16], M[16];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) H[i * 4 + j] = (*this)(i, j);
        shim::se3_exp<T>(H, M);
        Dense o(4, 4);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) o(i, j) = M[i * 4 + j];
        return o;
    }
    Dense log() const {
        assert(r_ == 4 && c_ == 4);
        T H[16], M[16];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M[i * 4 + j] = (*this)(i, j);
        shim::se3_log<T>(M, H);
        Dense o(4, 4);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) o(i, j) = H[i * 4 + j];
        return o;
    }

    /* arithmetic */
    Dense operator-() const { Dense o = *this; for (auto& x : o.a) x = -x; return o; }
    Dense& operator+=(const Dense& b) { assert(a.size() == b.a.size()); for (size_t k = 0; k < a.size(); k++) a[k] += b.a[k]; return *this; }
    Dense& operator-=(const Dense& b) { assert(a.size() == b.a.size()); for (size_t k = 0; k < a.size(); k++) a[k] -= b.a[k]; return *this; }
    Dense& operator*=(T s) { for (auto& x : a) x *= s; return *this; }
    Dense& operator/=(T s) { for (auto& x : a) x /= s; return *this; }
    friend Dense operator+(const Dense& x, const Dense& y) { Dense o = x; o += y; return o; }
    friend Dense operator-(const Dense& x, const Dense& y) { Dense o = x; o -= y; return o; }
    friend Dense operator*(T s, const Dense& x) { Dense o = x; for (auto& v : o.a) v = s * v; return o; }
    friend Dense operator*(const Dense& x, T s) { Dense o = x; for (auto& v : o.a) v = v * s; return o; }
    /* matrix product, sequential over the inner index */
    friend Dense operator*(const Dense& x, const Dense& y) {
        assert(x.c_ == y.r_);
        Dense o(x.r_, y.c_);
        for (int i = 0; i < x.r_; i++) for (int j = 0; j < y.c_; j++) { T acc = 0; for (int k = 0; k < x.c_; k++) acc += x(i, k) * y(k, j); o(i, j) = acc; }
        return o;
    }
};

/* ------------------------------------------------------------------------------------------------
 * ArrayD<T>: element-wise semantics
 * ---------------------------------------------------------------------------------------------- */
template <class T>
class ArrayD {
public:
    Dense<T> m;
    ArrayD() {}
    ArrayD(int r, int c) : m(r, c) {}
    explicit ArrayD(const Dense<T>& d) : m(d) {}
    Index size() const { return m.size(); }
    T& operator()(Index k) { return m.a[(size_t)k]; }
    const T& operator()(Index k) const { return m.a[(size_t)k]; }
    T& operator()(Index i, Index j) { return m(i, j); }
    const T& operator()(Index i, Index j) const { return m(i, j); }
    T& operator[](Index k) { return m.a[(size_t)k]; }
    const T& operator[](Index k) const { return m.a[(size_t)k]; }
    void fill(T v) { m.fill(v); }
    T* data() { return m.data(); }
    const T* data() const { return m.data(); }
    T sum() const { return m.sum(); }
    const Dense<T>& matrix() const { return m; }
    template <class U> ArrayD<U> cast() const { return ArrayD<U>(m.template cast<U>()); }
    ArrayD& operator/=(const ArrayD& b) { for (size_t k = 0; k < m.a.size(); k++) m.a[k] /= b.m.a[k]; return *this; }
    friend ArrayD operator*(const ArrayD& x, const ArrayD& y) { ArrayD o = x; for (size_t k = 0; k < o.m.a.size(); k++) o.m.a[k] = x.m.a[k] * y.m.a[k]; return o; }
    friend ArrayD operator*(T s, const ArrayD& x) { ArrayD o = x; for (auto& v : o.m.a) v = s * v; return o; }
    friend ArrayD operator-(const ArrayD& x, T s) { ArrayD o = x; for (auto& v : o.m.a) v = v - s; return o; }
    static ArrayD LinSpaced(Index n, T lo, T hi) {
        ArrayD o((int)n, 1);
        const T step = n > 1 ? (hi - lo) / (T)(n - 1) : (T)0;
        for (Index i = 0; i < n; i++) o.m.a[(size_t)i] = (i == n - 1) ? hi : lo + (T)i * step;
        return o;
    }
};
template <class T> ArrayD<T> Dense<T>::array() const { return ArrayD<T>(*this); }
template <class T> ArrayD<T> Dense<T>::ColRef::array() const { return ArrayD<T>(Dense<T>(*this)); }
template <class T> ArrayD<T> Dense<T>::ConstColRef::array() const { return ArrayD<T>(Dense<T>(*this)); }

/* ------------------------------------------------------------------------------------------------
 * Eigen-style type names
 * ---------------------------------------------------------------------------------------------- */
template <class T, int R, int C, int O = 0, int MR = R, int MC = C>
class Matrix : public Dense<T> {
public:
    Matrix() : Dense<T>(R > 0 ? R : 0, C > 0 ? C : (R > 0 && C == Dynamic ? 0 : 0)) { if (R > 0 && C > 0) this->resize(R, C); }
    Matrix(Index r, Index c) : Dense<T>((int)r, (int)c) {}
    explicit Matrix(Index n) : Dense<T>(C == 1 ? (int)n : (R == 1 ? 1 : (int)n), C == 1 ? 1 : (R == 1 ? (int)n : 1)) {}
    Matrix(T x, T y, T z) : Dense<T>(3, 1) { this->a[0] = x; this->a[1] = y; this->a[2] = z; }
    Matrix(T x, T y, T z, T w) : Dense<T>(4, 1) { this->a[0] = x; this->a[1] = y; this->a[2] = z; this->a[3] = w; }
    Matrix(const Dense<T>& d) : Dense<T>(d) {}
    Matrix(const typename Dense<T>::ColRef& d) : Dense<T>(d) {}
    Matrix(const typename Dense<T>::ConstColRef& d) : Dense<T>(d) {}
    template <class U> Matrix(const Dense<U>& d) : Dense<T>(d.template cast<T>()) {}
    Matrix& operator=(const Dense<T>& d) { Dense<T>::operator=(d); return *this; }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Zero(Index r, Index c) { Matrix m(r, c); m.setZero(); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
};
template <class T> template <int BR, int BC> Matrix<T, BR, BC, 0, BR, BC> Dense<T>::block(Index i0, Index j0) const {
    Matrix<T, BR, BC, 0, BR, BC> o;
    for (int j = 0; j < BC; j++) for (int i = 0; i < BR; i++) o(i, j) = (*this)(i0 + i, j0 + j);
    return o;
}
template <class T, int R, int C>
class Array : public ArrayD<T> {
public:
    Array() : ArrayD<T>(R > 0 ? R : 0, C > 0 ? C : 0) {}
    Array(const ArrayD<T>& d) : ArrayD<T>(d) {}
    Array& operator=(const ArrayD<T>& d) { ArrayD<T>::operator=(d); return *this; }
    static ArrayD<T> LinSpaced(Index n, T lo, T hi) { return ArrayD<T>::LinSpaced(n, lo, hi); }
};

typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<int, Dynamic, Dynamic> MatrixXi;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<float, 2, 2> Matrix2f;
/* Eigen::Quaternionf(Matrix3f): stand-in for Eigen's rotation-matrix constructor (QuaternionBase::operator=(MatrixBase), the
 * trace / largest-diagonal branches of Eigen/src/Geometry/Quaternion.h), float arithmetic */
class Quaternionf {
    float q_[4];  /* x, y, z, w */
public:
    explicit Quaternionf(const Dense<float>& m) {
        float t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > 0.f) {
            t = std::sqrt(t + 1.f);
            q_[3] = 0.5f * t;
            t = 0.5f / t;
            q_[0] = (m(2, 1) - m(1, 2)) * t; q_[1] = (m(0, 2) - m(2, 0)) * t; q_[2] = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.f);
            q_[i] = 0.5f * t;
            t = 0.5f / t;
            q_[3] = (m(k, j) - m(j, k)) * t; q_[j] = (m(j, i) + m(i, j)) * t; q_[k] = (m(k, i) + m(i, k)) * t;
        }
    }
    float x() const { return q_[0]; }
    float y() const { return q_[1]; }
    float z() const { return q_[2]; }
    float w() const { return q_[3]; }
};
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Array<float, Dynamic, 1> ArrayXf;
typedef Array<float, 4, 4> Array44f;

template <class M>
class SelfAdjointEigenSolver {
    Dense<float> vec_, val_;
public:
    explicit SelfAdjointEigenSolver(const Dense<float>& A) {
        const int n = (int)A.rows();
        std::vector<float> a((size_t)n * n), ev(n), V((size_t)n * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[(size_t)i * n + j] = A(i, j);
        /* only the lower/upper symmetric part matters to Eigen; the reference passes a numerically symmetric matrix */
        for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) { const float m = 0.5f * (a[(size_t)i * n + j] + a[(size_t)j * n + i]); a[(size_t)i * n + j] = m; a[(size_t)j * n + i] = m; }
        shim::jacobi_eig<float>(n, a.data(), ev.data(), V.data());
        vec_ = Dense<float>(n, n); val_ = Dense<float>(n, 1);
        for (int i = 0; i < n; i++) { val_(i, 0) = ev[i]; for (int j = 0; j < n; j++) vec_(i, j) = V[(size_t)i * n + j]; }
    }
    ComputationInfo info() const { return Success; }
    const Dense<float>& eigenvectors() const { return vec_; }
    const Dense<float>& eigenvalues() const { return val_; }
};

}  // namespace Eigen

/* ------------------------------------------------------------------------------------------------
 * MRPT 1.x slivers
 * ---------------------------------------------------------------------------------------------- */
namespace mrpt {
namespace utils {
template <class T> inline T square(const T x) { return x * x; }
struct CTicTac { void Tic() {} double Tac() { return 0; } };
}  // namespace utils
namespace math {
struct CMatrixDouble33 : public Eigen::Dense<double> {
    CMatrixDouble33() : Eigen::Dense<double>(3, 3) {}
    CMatrixDouble33(const Eigen::Dense<double>& d) : Eigen::Dense<double>(d) {}
};
struct CMatrixDouble44 : public Eigen::Dense<double> {
    CMatrixDouble44() : Eigen::Dense<double>(4, 4) {}
    CMatrixDouble44(const Eigen::Dense<float>& d) : Eigen::Dense<double>(d.cast<double>()) {}
};
}  // namespace math
namespace poses {
/* homogeneous-matrix pose: all the reference needs is composition and the rotation block (FrontEnd.cpp:1134-1144) */
class CPose3D {
public:
    double M[16];
    CPose3D() { setFromValues(0, 0, 0, 0, 0, 0); }
    explicit CPose3D(const math::CMatrixDouble44& m) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M[i * 4 + j] = m(i, j); }
    void setFromValues(double x, double y, double z, double yaw, double pitch, double roll) {
        const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch), cr = std::cos(roll), sr = std::sin(roll);
        const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, -sp, cp * sr, cp * cr};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[i * 4 + j] = R[i * 3 + j];
        M[3] = x; M[7] = y; M[11] = z; M[12] = M[13] = M[14] = 0; M[15] = 1;
    }
    CPose3D operator+(const CPose3D& b) const {
        CPose3D o;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += M[i * 4 + k] * b.M[k * 4 + j]; o.M[i * 4 + j] = s; }
        return o;
    }
    void getRotationMatrix(math::CMatrixDouble33& R) const { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R(i, j) = M[i * 4 + j]; }
};
}  // namespace poses
}  // namespace mrpt

/* ------------------------------------------------------------------------------------------------
 * OpenCV sliver.  StaticFusion.h declares two cv::Mat members (depth_mm, color_full); the solver never
 * touches them, the image-sequence loader (FrontEnd.cpp:216-254) does: imread, at<>, convertTo, Vec3b.
 * Stand-ins for library internals that are not in the reference tree:
 *   - imread() returns images the test harness registered under that file name, in the layout OpenCV's
 *     decoder produces (8-bit 3-channel BGR for CV_LOAD_IMAGE_COLOR, the file's own depth for flag -1):
 *     PNG decoding itself is libpng's job, not the reference's;
 *   - convertTo(dst, type, alpha) computes saturate_cast<dst>(float(src) * float(alpha)) per element, the
 *     arithmetic of OpenCV's cvtScale_ for 16U sources (work type float);
 *   - Vec3b(float, float, float) converts each argument to uchar implicitly (C++ truncation), as the
 *     cv::Vec<uchar,3>(uchar, uchar, uchar) constructor the reference's call resolves to does.
 * ---------------------------------------------------------------------------------------------- */
#include <map>
#include <memory>
#ifndef CV_16U
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_LOAD_IMAGE_COLOR 1
#endif
namespace cv {
struct Scalar { Scalar(double = 0, double = 0, double = 0, double = 0) {} };
struct Vec3b {
    unsigned char val[3];
    Vec3b() { val[0] = val[1] = val[2] = 0; }
    Vec3b(unsigned char a, unsigned char b, unsigned char c) { val[0] = a; val[1] = b; val[2] = c; }
    unsigned char& operator[](int i) { return val[i]; }
    const unsigned char& operator[](int i) const { return val[i]; }
};
struct Mat {
    int rows = 0, cols = 0, type_ = 0;
    std::shared_ptr<std::vector<unsigned char> > store;  /* shallow copies, like cv::Mat */
    unsigned char* data = nullptr;
    static size_t elem(int type) { return type == CV_8UC3 ? 3 : type == CV_16U ? 2 : type == CV_32F ? 4 : 1; }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        store = std::make_shared<std::vector<unsigned char> >((size_t)r * c * elem(type), (unsigned char)0);
        data = store->data();
    }
    Mat() {}
    Mat(int r, int c, int type, double) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar&) { create(r, c, type); }
    int type() const { return type_; }
    template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + ((size_t)r * cols + c) * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + ((size_t)r * cols + c) * sizeof(T)); }
    void convertTo(Mat& dst, int rtype, double alpha = 1.0) const {
        assert(type_ == CV_16U && (rtype == CV_32F || rtype == CV_16U));
        Mat out; out.create(rows, cols, rtype);
        const float a = (float)alpha;
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) {
                const float v = (float)at<unsigned short>(r, c) * a;
                if (rtype == CV_32F) out.at<float>(r, c) = v;
                else { const float q = std::nearbyint(v); out.at<unsigned short>(r, c) = (unsigned short)(q < 0.f ? 0.f : q > 65535.f ? 65535.f : q); }
            }
        dst = out;
    }
};
inline std::map<std::string, Mat>& shim_image_registry() { static std::map<std::string, Mat> r; return r; }
inline Mat imread(const std::string& name, int /*flags*/) {
    auto it = shim_image_registry().find(name);
    return it == shim_image_registry().end() ? Mat() : it->second;
}
}  // namespace cv

/* the GL back-end and the GUI are out of scope: StaticFusion.h only holds pointers to them */
class Reconstruction;
class GUI;

#endif /* SF_REF_SHIM_H */
