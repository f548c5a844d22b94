// ============================================================================
// oracle/adelie_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (Eigen-free C++17 + OpenMP) of the reference's pathwise
// block-coordinate-descent group-elastic-net solver ("naive" method).  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load the library built from this file.  The product
// (adelie_b200/) never links, imports or calls anything in oracle/.
//
// The reference core (adelie/src/include/adelie_core, "CORE/") is templated on
// Eigen 3.4.0 which is NOT vendored and NOT installed here, so the reference
// itself cannot be compiled ("parity unpinned" against a reference binary; the
// oracle is pinned instead by the checks in tests/test_oracle_*.py: an
// independent scikit-learn coordinate-descent solver, NumPy KKT residuals, and
// golden vectors produced by the reference's own NumPy test restatements).
//
// Every function cites the reference file:line it follows.  All arithmetic is
// done in the value type T (float or double) exactly like the reference
// templates (value_t), timers in double.
// ============================================================================
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace orc {

// CORE/configs.hpp:6-20
struct Configs {
    static inline double hessian_min = 1e-24;
    static inline double dbeta_tol = 1e-12;
    static inline size_t min_bytes = 1 << 17;
};

struct solver_error : std::runtime_error {   // CORE/util/exceptions.hpp:8-56
    using std::runtime_error::runtime_error;
};

inline double now_s() {
    return std::chrono::duration<double>(
        std::chrono::steady_clock::now().time_since_epoch()).count();
}

using idx_t = int64_t;

// ---------------------------------------------------------------------------
// L0 helpers (CORE/matrix/utils.hpp:133-269): dot / gemv with row blocking.
// ---------------------------------------------------------------------------
template <class T>
inline T ddot(const T* a, const T* b, idx_t n, int n_threads) {
    // CORE/matrix/utils.hpp:133-161 -- per-thread partial sums, then sum.
    const size_t n_bytes = sizeof(T) * 2 * n;
    if (n_threads <= 1 || n_bytes <= Configs::min_bytes) {
        T s = 0;
        for (idx_t i = 0; i < n; ++i) s += a[i] * b[i];
        return s;
    }
    T s = 0;
    #pragma omp parallel for schedule(static) num_threads(n_threads) reduction(+:s)
    for (idx_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// ---------------------------------------------------------------------------
// Matrix interface (CORE/matrix/matrix_naive_base.hpp:57-143).
// ---------------------------------------------------------------------------
template <class T>
struct MatrixBase {
    virtual ~MatrixBase() {}
    virtual idx_t rows() const = 0;
    virtual idx_t cols() const = 0;
    // out = X[:,j]^T (v*w)
    virtual T cmul(idx_t j, const T* v, const T* w) = 0;
    // out += v * X[:,j]
    virtual void ctmul(idx_t j, T v, T* out) = 0;
    // out[q] = X[:, j:j+q]^T (v*w)
    virtual void bmul(idx_t j, idx_t q, const T* v, const T* w, T* out) = 0;
    // out += X[:, j:j+q] v
    virtual void btmul(idx_t j, idx_t q, const T* v, T* out) = 0;
    // out[p] = X^T (v*w)
    virtual void mul(const T* v, const T* w, T* out) = 0;
    // out (q x q col-major) = X_g^T diag(sqrt_w^2) X_g
    virtual void cov(idx_t j, idx_t q, const T* sqrt_w, T* out) = 0;
};

// Dense column-major matrix (CORE/matrix/matrix_naive_dense.ipp:25-221).
template <class T>
struct MatrixDense : MatrixBase<T> {
    const T* X; idx_t n, p, ld; int n_threads;
    std::vector<T> vbuff;
    MatrixDense(const T* X_, idx_t n_, idx_t p_, idx_t ld_, int nt)
        : X(X_), n(n_), p(p_), ld(ld_), n_threads(nt), vbuff(n_) {}
    idx_t rows() const override { return n; }
    idx_t cols() const override { return p; }
    const T* col(idx_t j) const { return X + j * ld; }

    T cmul(idx_t j, const T* v, const T* w) override {          // dense.ipp:25-35
        const T* x = col(j);
        const size_t n_bytes = sizeof(T) * 2 * n;
        if (n_threads <= 1 || n_bytes <= Configs::min_bytes) {
            T s = 0;
            for (idx_t i = 0; i < n; ++i) s += x[i] * (v[i] * w[i]);
            return s;
        }
        T s = 0;
        #pragma omp parallel for schedule(static) num_threads(n_threads) reduction(+:s)
        for (idx_t i = 0; i < n; ++i) s += x[i] * (v[i] * w[i]);
        return s;
    }
    void ctmul(idx_t j, T v, T* out) override {                  // dense.ipp:51-61
        const T* x = col(j);
        const size_t n_bytes = sizeof(T) * 2 * n;
        if (n_threads <= 1 || n_bytes <= Configs::min_bytes) {
            for (idx_t i = 0; i < n; ++i) out[i] += v * x[i];
            return;
        }
        #pragma omp parallel for schedule(static) num_threads(n_threads)
        for (idx_t i = 0; i < n; ++i) out[i] += v * x[i];
    }
    void bmul(idx_t j, idx_t q, const T* v, const T* w, T* out) override {   // dense.ipp:63-82
        // first pass materialises v*w (dvveq into _vbuff, dense.ipp:73), then a
        // row-blocked GEMV (utils.hpp:202-269) -- kept as two passes so the CPU
        // baseline has the reference's memory traffic.
        T* vb = vbuff.data();
        const size_t n_bytes = sizeof(T) * n * (q + 1);
        const bool par = (n_threads > 1) && (n_bytes > Configs::min_bytes);
        if (par) {
            #pragma omp parallel for schedule(static) num_threads(n_threads)
            for (idx_t i = 0; i < n; ++i) vb[i] = v[i] * w[i];
        } else {
            for (idx_t i = 0; i < n; ++i) vb[i] = v[i] * w[i];
        }
        if (!par) {
            for (idx_t c = 0; c < q; ++c) {
                const T* x = col(j + c);
                T s = 0;
                for (idx_t i = 0; i < n; ++i) s += x[i] * vb[i];
                out[c] = s;
            }
            return;
        }
        // row-blocked: each thread owns a row block, accumulates q partials.
        const int nb = n_threads;
        std::vector<T> part((size_t)nb * q, T(0));
        #pragma omp parallel num_threads(n_threads)
        {
#if defined(_OPENMP)
            const int t = omp_get_thread_num();
            const int nt = omp_get_num_threads();
#else
            const int t = 0, nt = 1;
#endif
            const idx_t bs = n / nt, rem = n % nt;
            const idx_t b = std::min<idx_t>(t, rem) * (bs + 1) + std::max<idx_t>(t - rem, 0) * bs;
            const idx_t sz = bs + (t < rem);
            for (idx_t c = 0; c < q; ++c) {
                const T* x = col(j + c) + b;
                const T* vv = vb + b;
                T s = 0;
                for (idx_t i = 0; i < sz; ++i) s += x[i] * vv[i];
                part[(size_t)t * q + c] = s;
            }
        }
        for (idx_t c = 0; c < q; ++c) {
            T s = 0;
            for (int t = 0; t < nb; ++t) s += part[(size_t)t * q + c];
            out[c] = s;
        }
    }
    void btmul(idx_t j, idx_t q, const T* v, T* out) override {              // dense.ipp:108-125
        const size_t n_bytes = sizeof(T) * n * (q + 1);
        const bool par = (n_threads > 1) && (n_bytes > Configs::min_bytes);
        if (!par) {
            for (idx_t c = 0; c < q; ++c) {
                const T* x = col(j + c);
                const T vc = v[c];
                for (idx_t i = 0; i < n; ++i) out[i] += x[i] * vc;
            }
            return;
        }
        #pragma omp parallel num_threads(n_threads)
        {
#if defined(_OPENMP)
            const int t = omp_get_thread_num();
            const int nt = omp_get_num_threads();
#else
            const int t = 0, nt = 1;
#endif
            const idx_t bs = n / nt, rem = n % nt;
            const idx_t b = std::min<idx_t>(t, rem) * (bs + 1) + std::max<idx_t>(t - rem, 0) * bs;
            const idx_t sz = bs + (t < rem);
            for (idx_t c = 0; c < q; ++c) {
                const T* x = col(j + c) + b;
                const T vc = v[c];
                T* o = out + b;
                for (idx_t i = 0; i < sz; ++i) o[i] += x[i] * vc;
            }
        }
    }
    void mul(const T* v, const T* w, T* out) override {                      // dense.ipp:127-148
        std::vector<T> vb(n);
        for (idx_t i = 0; i < n; ++i) vb[i] = v[i] * w[i];
        #pragma omp parallel for schedule(static) num_threads(n_threads) if (n_threads > 1)
        for (idx_t c = 0; c < p; ++c) {
            const T* x = col(c);
            T s = 0;
            for (idx_t i = 0; i < n; ++i) s += x[i] * vb[i];
            out[c] = s;
        }
    }
    void cov(idx_t j, idx_t q, const T* sqrt_w, T* out) override {           // dense.ipp:164-199
        // out = (sqrt_w * X_g)^T (sqrt_w * X_g); lower triangle computed then mirrored.
        for (idx_t a = 0; a < q; ++a) {
            const T* xa = col(j + a);
            for (idx_t b = 0; b <= a; ++b) {
                const T* xb = col(j + b);
                T s = 0;
                for (idx_t i = 0; i < n; ++i) {
                    const T sa = xa[i] * sqrt_w[i];
                    const T sb = xb[i] * sqrt_w[i];
                    s += sa * sb;
                }
                out[a + b * q] = s;
                out[b + a * q] = s;
            }
        }
    }
};

// Sparse CSC matrix (CORE/matrix/matrix_naive_sparse.ipp:10-262; spddot/spaxi
// CORE/matrix/utils.hpp:438-515).
template <class T>
struct MatrixSparse : MatrixBase<T> {
    idx_t n, p; const int32_t* outer; const int32_t* inner; const T* val; int n_threads;
    MatrixSparse(idx_t n_, idx_t p_, const int32_t* o, const int32_t* in, const T* v, int nt)
        : n(n_), p(p_), outer(o), inner(in), val(v), n_threads(nt) {}
    idx_t rows() const override { return n; }
    idx_t cols() const override { return p; }
    T cmul(idx_t j, const T* v, const T* w) override {           // sparse.ipp:10-34
        T s = 0;
        for (int32_t k = outer[j]; k < outer[j + 1]; ++k) {
            const int32_t i = inner[k];
            s += val[k] * (v[i] * w[i]);
        }
        return s;
    }
    void ctmul(idx_t j, T v, T* out) override {                   // sparse.ipp:51-68
        for (int32_t k = outer[j]; k < outer[j + 1]; ++k) out[inner[k]] += v * val[k];
    }
    void bmul(idx_t j, idx_t q, const T* v, const T* w, T* out) override {   // sparse.ipp:70-100
        for (idx_t c = 0; c < q; ++c) out[c] = cmul(j + c, v, w);
    }
    void btmul(idx_t j, idx_t q, const T* v, T* out) override {   // sparse.ipp:128-150
        for (idx_t c = 0; c < q; ++c) ctmul(j + c, v[c], out);
    }
    void mul(const T* v, const T* w, T* out) override {           // sparse.ipp:152-198
        #pragma omp parallel for schedule(static) num_threads(n_threads) if (n_threads > 1)
        for (idx_t c = 0; c < p; ++c) out[c] = cmul(c, v, w);
    }
    void cov(idx_t j, idx_t q, const T* sqrt_w, T* out) override {  // sparse.ipp:200-240
        // out[a,b] = sum_i X[i,j+a] X[i,j+b] sqrt_w[i]^2 via sorted index merge.
        for (idx_t a = 0; a < q; ++a) {
            for (idx_t b = 0; b <= a; ++b) {
                int32_t ka = outer[j + a], ea = outer[j + a + 1];
                int32_t kb = outer[j + b], eb = outer[j + b + 1];
                T s = 0;
                while (ka < ea && kb < eb) {
                    if (inner[ka] < inner[kb]) ++ka;
                    else if (inner[ka] > inner[kb]) ++kb;
                    else {
                        const T sw = sqrt_w[inner[ka]];
                        s += val[ka] * val[kb] * sw * sw;
                        ++ka; ++kb;
                    }
                }
                out[a + b * q] = s;
                out[b + a * q] = s;
            }
        }
    }
};

// kron(X, I_K) (CORE/matrix/matrix_naive_kronecker_eye.ipp:29-352): column
// j*K + l of the (n*K, p*K) matrix is X[:, j] placed at rows i*K + l.
template <class T>
struct MatrixKroneckerEye : MatrixBase<T> {
    MatrixBase<T>* M; idx_t K;
    std::vector<T> vb, wb, ob;
    MatrixKroneckerEye(MatrixBase<T>* M_, idx_t K_)
        : M(M_), K(K_), vb(M_->rows()), wb(M_->rows()), ob(M_->rows()) {}
    idx_t rows() const override { return M->rows() * K; }
    idx_t cols() const override { return M->cols() * K; }
    T cmul(idx_t j, const T* v, const T* w) override {            // kronecker_eye.ipp:29-50
        const idx_t i = j / K, l = j - K * i, n = M->rows();
        for (idx_t r = 0; r < n; ++r) { vb[r] = v[r * K + l]; wb[r] = w[r * K + l]; }
        return M->cmul(i, vb.data(), wb.data());
    }
    void ctmul(idx_t j, T v, T* out) override {                    // kronecker_eye.ipp:69-87
        const idx_t i = j / K, l = j - K * i, n = M->rows();
        std::fill(ob.begin(), ob.end(), T(0));
        M->ctmul(i, v, ob.data());
        for (idx_t r = 0; r < n; ++r) out[r * K + l] += ob[r];
    }
    void bmul(idx_t j, idx_t q, const T* v, const T* w, T* out) override {   // kronecker_eye.ipp:89-118
        for (idx_t c = 0; c < q; ++c) out[c] = cmul(j + c, v, w);
    }
    void btmul(idx_t j, idx_t q, const T* v, T* out) override {    // kronecker_eye.ipp:152-178
        for (idx_t c = 0; c < q; ++c) ctmul(j + c, v[c], out);
    }
    void mul(const T* v, const T* w, T* out) override {            // kronecker_eye.ipp:180-204
        const idx_t n = M->rows(), p = M->cols();
        std::vector<T> o(p);
        for (idx_t l = 0; l < K; ++l) {
            for (idx_t r = 0; r < n; ++r) { vb[r] = v[r * K + l]; wb[r] = w[r * K + l]; }
            M->mul(vb.data(), wb.data(), o.data());
            for (idx_t c = 0; c < p; ++c) out[c * K + l] = o[c];
        }
    }
    void cov(idx_t j, idx_t q, const T* sqrt_w, T* out) override { // kronecker_eye.ipp:206-262
        // entries couple only columns with the same class l.
        const idx_t n = M->rows();
        std::fill(out, out + q * q, T(0));
        for (idx_t l = 0; l < K; ++l) {
            // columns of this block with class l: c such that (j+c) % K == l
            std::vector<idx_t> cs;
            for (idx_t c = 0; c < q; ++c) if ((j + c) % K == l) cs.push_back(c);
            if (cs.empty()) continue;
            for (idx_t r = 0; r < n; ++r) wb[r] = sqrt_w[r * K + l];
            // inner features are contiguous: (j+cs[0])/K ...
            const idx_t i0 = (j + cs[0]) / K;
            const idx_t qq = (idx_t)cs.size();
            std::vector<T> o(qq * qq);
            M->cov(i0, qq, wb.data(), o.data());
            for (idx_t a = 0; a < qq; ++a)
                for (idx_t b = 0; b < qq; ++b)
                    out[cs[a] + cs[b] * q] = o[a + b * qq];
        }
    }
};

// Column-wise concatenation (CORE/matrix/matrix_naive_concatenate.ipp:120-344).
template <class T>
struct MatrixCConcatenate : MatrixBase<T> {
    std::vector<MatrixBase<T>*> mats; std::vector<idx_t> begins; idx_t n, p;
    MatrixCConcatenate(const std::vector<MatrixBase<T>*>& m) : mats(m) {
        n = m[0]->rows(); p = 0;
        for (auto* x : m) { begins.push_back(p); p += x->cols(); }
    }
    idx_t rows() const override { return n; }
    idx_t cols() const override { return p; }
    size_t slice(idx_t j) const {
        size_t k = std::upper_bound(begins.begin(), begins.end(), j) - begins.begin() - 1;
        return k;
    }
    T cmul(idx_t j, const T* v, const T* w) override {
        const size_t k = slice(j); return mats[k]->cmul(j - begins[k], v, w);
    }
    void ctmul(idx_t j, T v, T* out) override {
        const size_t k = slice(j); mats[k]->ctmul(j - begins[k], v, out);
    }
    void bmul(idx_t j, idx_t q, const T* v, const T* w, T* out) override {   // concatenate.ipp:165-196
        idx_t done = 0;
        while (done < q) {
            const size_t k = slice(j + done);
            const idx_t jj = j + done - begins[k];
            const idx_t qq = std::min<idx_t>(mats[k]->cols() - jj, q - done);
            mats[k]->bmul(jj, qq, v, w, out + done);
            done += qq;
        }
    }
    void btmul(idx_t j, idx_t q, const T* v, T* out) override {              // concatenate.ipp:226-245
        idx_t done = 0;
        while (done < q) {
            const size_t k = slice(j + done);
            const idx_t jj = j + done - begins[k];
            const idx_t qq = std::min<idx_t>(mats[k]->cols() - jj, q - done);
            mats[k]->btmul(jj, qq, v + done, out);
            done += qq;
        }
    }
    void mul(const T* v, const T* w, T* out) override {
        for (size_t k = 0; k < mats.size(); ++k) mats[k]->mul(v, w, out + begins[k]);
    }
    void cov(idx_t j, idx_t q, const T* sqrt_w, T* out) override {           // concatenate.ipp:261-288
        const size_t k = slice(j);
        if (j - begins[k] + q > mats[k]->cols())
            throw std::runtime_error("MatrixNaiveCConcatenate::cov() only allows the block to be fully contained in one of the matrices in the list.");
        mats[k]->cov(j - begins[k], q, sqrt_w, out);
    }
};

// ---------------------------------------------------------------------------
// Symmetric eigendecomposition (replaces Eigen::SelfAdjointEigenSolver at
// CORE/solver/solver_gaussian_naive.hpp:113).  Cyclic Jacobi; eigenvalues
// ascending, eigenvectors in the COLUMNS of V (col-major q x q).
// Eigenvectors are sign/rotation ambiguous but beta = beta~ V^T is invariant.
// ---------------------------------------------------------------------------
template <class T>
inline void jacobi_eigh(T* A /*q x q col-major, destroyed*/, idx_t q, T* D, T* V) {
    for (idx_t i = 0; i < q * q; ++i) V[i] = 0;
    for (idx_t i = 0; i < q; ++i) V[i + i * q] = 1;
    const int max_sweeps = 64;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off = 0, diag = 0;
        for (idx_t a = 0; a < q; ++a)
            for (idx_t b = 0; b < q; ++b) {
                const double x = (double)A[a + b * q];
                if (a == b) diag += x * x; else off += x * x;
            }
        if (off <= 1e-31 * (diag + off) || off == 0) break;     // off-diagonal mass below (eps)^2 of the total
        for (idx_t pI = 0; pI < q - 1; ++pI) {
            for (idx_t qI = pI + 1; qI < q; ++qI) {
                const T apq = A[pI + qI * q];
                if (apq == 0) continue;
                const T app = A[pI + pI * q], aqq = A[qI + qI * q];
                const T theta = (aqq - app) / (2 * apq);
                const T t = (theta >= 0 ? T(1) : T(-1)) / (std::abs(theta) + std::sqrt(theta * theta + 1));
                const T c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (idx_t k = 0; k < q; ++k) {      // A <- A J
                    const T akp = A[k + pI * q], akq = A[k + qI * q];
                    A[k + pI * q] = c * akp - s * akq;
                    A[k + qI * q] = s * akp + c * akq;
                }
                for (idx_t k = 0; k < q; ++k) {      // A <- J^T A
                    const T apk = A[pI + k * q], aqk = A[qI + k * q];
                    A[pI + k * q] = c * apk - s * aqk;
                    A[qI + k * q] = s * apk + c * aqk;
                }
                for (idx_t k = 0; k < q; ++k) {      // V <- V J
                    const T vkp = V[k + pI * q], vkq = V[k + qI * q];
                    V[k + pI * q] = c * vkp - s * vkq;
                    V[k + qI * q] = s * vkp + c * vkq;
                }
            }
        }
    }
    // sort ascending
    std::vector<idx_t> ord(q);
    std::iota(ord.begin(), ord.end(), 0);
    std::sort(ord.begin(), ord.end(), [&](idx_t a, idx_t b) { return A[a + a * q] < A[b + b * q]; });
    std::vector<T> Vs(q * q);
    for (idx_t c = 0; c < q; ++c) {
        D[c] = A[ord[c] + ord[c] * q];
        for (idx_t k = 0; k < q; ++k) Vs[k + c * q] = V[k + ord[c] * q];
    }
    std::copy(Vs.begin(), Vs.end(), V);
}

// ---------------------------------------------------------------------------
// Group proximal sub-problem (CORE/bcd/unconstrained/newton.hpp,
// CORE/bcd/utils.hpp, CORE/optimization/newton.hpp).
// ---------------------------------------------------------------------------
template <class T>
inline T root_lower_bound(const T* D, const T* v, idx_t q, T l1) {          // bcd/utils.hpp:20-41
    T sumD = 0, a = 0, v_l1 = 0;
    for (idx_t i = 0; i < q; ++i) { sumD += D[i]; a += D[i] * D[i]; v_l1 += std::abs(v[i]); }
    const T b = l1 * sumD;
    const T c = l1 * l1 * q - v_l1 * v_l1;
    const T discr = b * b - a * c;
    T h_min = (discr > -1e-12) ? (-b + std::sqrt(std::max<T>(discr, 0.0))) / a : T(0.0);
    h_min = std::max<T>(h_min, 0.0);
    return h_min;
}

template <class T>
inline std::pair<T, T> root_upper_bound(const T* D, const T* v, idx_t q, T l1, T zero_tol = 1e-14) {   // bcd/utils.hpp:59-97
    T dmin = D[0];
    for (idx_t i = 1; i < q; ++i) dmin = std::min(dmin, D[i]);
    T dmin_nnz = std::numeric_limits<T>::infinity();
    T h_max = 0, v_S = 0;
    if (dmin <= zero_tol) {
        for (idx_t i = 0; i < q; ++i) {
            const bool nz = D[i] > zero_tol;
            const T vi2 = v[i] * v[i];
            h_max += nz ? (vi2 / (D[i] * D[i])) : 0;
            v_S += (D[i] <= 0) ? vi2 : 0;
            dmin_nnz = nz ? std::min(dmin_nnz, D[i]) : dmin_nnz;
        }
        h_max = std::sqrt(std::max<T>(h_max / (1 - v_S / (l1 * l1)), 0));
    } else {
        dmin_nnz = dmin;
        T s = 0;
        for (idx_t i = 0; i < q; ++i) { const T r = v[i] / D[i]; s += r * r; }
        h_max = std::sqrt(s);
    }
    return {h_max, dmin_nnz};
}

template <class T>
inline T root_function(T h, const T* D, const T* v, idx_t q, T l1) {        // bcd/utils.hpp:99-109
    T s = 0;
    for (idx_t i = 0; i < q; ++i) { const T r = v[i] / (D[i] * h + l1); s += r * r; }
    return s - 1;
}

// newton_solver_base (newton.hpp:44-111) with newton_root_find
// (optimization/newton.hpp:35-66) inlined.  `abs_start` selects the
// newton_abs_solver initial point (newton.hpp:229-266), else h0 = 0.
template <class T>
inline void newton_prox(
    const T* L, const T* v, idx_t q, T l1, T l2, T tol, size_t max_iters,
    bool abs_start, T* x, size_t& iters, T* buf1, T* buf2)
{
    iters = 0;
    T vn = 0;
    for (idx_t i = 0; i < q; ++i) vn += v[i] * v[i];
    vn = std::sqrt(vn);
    if (vn <= l1) { for (idx_t i = 0; i < q; ++i) x[i] = 0; return; }       // :62-66
    if (l1 <= 0.0) { for (idx_t i = 0; i < q; ++i) x[i] = v[i] / (L[i] + l2); return; }   // :72-75
    for (idx_t i = 0; i < q; ++i) buf1[i] = L[i] + l2;                       // :81
    T h = 0;
    if (abs_start) {
        const T h_min = root_lower_bound(buf1, v, q, l1);
        const auto ub = root_upper_bound(buf1, v, q, l1);
        const T h_max = ub.first, dmin_nnz = ub.second;
        if (h_max - h_min <= 1e-1) {
            h = h_min;
        } else {
            T h_cand = h_max, w, fh;
            auto ada = [&]() {
                w = std::max<T>(l1 / (dmin_nnz * h_cand + l1), 0.05);
                h_cand = w * h_min + (1 - w) * h_cand;
                fh = root_function(h_cand, buf1, v, q, l1);
            };
            ada();
            while ((fh < 0) && (std::abs(fh) > tol)) ada();
            h = h_cand;
        }
    }
    double fh; T dfh;
    auto step = [&](T hh) {                                                  // :83-93
        T t = 0;
        for (idx_t i = 0; i < q; ++i) {
            buf2[i] = 1 / (buf1[i] * hh + l1);
            const T r = v[i] * buf2[i];
            x[i] = r * r;
            t += x[i];
        }
        const T sqrt_t = std::sqrt(t);
        fh = t - 1.0;
        T s = 0;
        for (idx_t i = 0; i < q; ++i) s += x[i] * buf1[i] * buf2[i];
        dfh = -s * (1 + sqrt_t) / t;
    };
    step(h);
    while ((std::abs(fh) > tol) && (iters < max_iters)) {                    // optimization/newton.hpp:56-63
        h -= fh / dfh;
        h = std::max<T>(h, 0.0);
        step(h);
        ++iters;
    }
    for (idx_t i = 0; i < q; ++i) x[i] = h * v[i] * buf2[i];                 // :109
}

// bcd objective 0.5 x^T L x - v^T x + l1 ||x|| + 0.5 l2 ||x||^2 (PY/bcd.py objective)
template <class T>
inline T bcd_objective(const T* L, const T* v, idx_t q, T l1, T l2, const T* x) {
    T s = 0, nn = 0;
    for (idx_t i = 0; i < q; ++i) { s += T(0.5) * L[i] * x[i] * x[i] - v[i] * x[i]; nn += x[i] * x[i]; }
    return s + l1 * std::sqrt(nn) + T(0.5) * l2 * nn;
}

// ---------------------------------------------------------------------------
// search_pivot (CORE/optimization/search_pivot.hpp:7-62)
// ---------------------------------------------------------------------------
template <class T>
inline int search_pivot(const T* x, const T* y, idx_t n, T* mses) {
    if (n <= 0) return -1;
    mses[0] = std::numeric_limits<T>::infinity();
    if (n == 1) return 0;
    T y_mean = 0;
    for (idx_t i = 0; i < n; ++i) y_mean += y[i];
    y_mean /= n;
    T x_sum = x[0], xsq_sum = x[0] * x[0], y_sum = y[0], yx_sum = y[0] * x[0];
    T min_mse = mses[0];
    int argmin = 0;
    for (idx_t i = 1; i < n; ++i) {
        x_sum += x[i]; xsq_sum += x[i] * x[i]; y_sum += y[i]; yx_sum += y[i] * x[i];
        const T t_bar = ((i + 1) * x[i] - x_sum) / n;
        const T var_t = (i + 1) * x[i] * x[i] - 2 * x[i] * x_sum + xsq_sum - n * t_bar * t_bar;
        const T cov_ty = x[i] * (y_sum - (i + 1) * y_mean) - (yx_sum - y_mean * x_sum);
        const T b1 = cov_ty / var_t;
        mses[i] = -b1 * b1 * var_t;
        if (mses[i] < min_mse) { argmin = (int)i; min_mse = mses[i]; }
    }
    return argmin;
}

// ---------------------------------------------------------------------------
// Pin state + CD sweeps (CORE/solver/solver_gaussian_pin_naive.hpp,
// CORE/solver/solver_gaussian_pin_base.hpp, CORE/state/state_gaussian_pin_naive.hpp)
// ---------------------------------------------------------------------------
template <class T>
struct PinState {
    // static
    MatrixBase<T>* X;
    T y_mean, y_var;
    const idx_t* groups; const idx_t* group_sizes; idx_t G;
    T alpha; const T* penalty; const T* weights;
    const idx_t* screen_set; const idx_t* screen_begins; idx_t S;
    const T* screen_vars; const T* screen_X_means;
    const std::vector<std::vector<T>>* screen_transforms;   // each gs x gs, element (r, c) at [r*gs + c] (row-major like VectorMatrix)
    std::vector<T> lmda_path;
    bool intercept; size_t max_active_size, max_iters;
    T tol, adev_tol, ddev_tol, newton_tol; size_t newton_max_iters;
    // dynamic
    T rsq; T* resid; T resid_sum;
    T* screen_beta; int8_t* screen_is_active;
    size_t active_set_size; idx_t* active_set;
    std::vector<T> screen_grad;
    std::vector<idx_t> active_begins, active_order;
    // outputs
    std::vector<std::vector<idx_t>> beta_idx; std::vector<std::vector<T>> beta_val;
    std::vector<T> intercepts, rsqs, lmdas;
    size_t iters = 0;
    std::vector<double> benchmark_screen, benchmark_active;
    size_t n_group_updates = 0;   // extra counter (not in the reference): groups visited by sweeps
};

// update_coordinate scalar (pin_base.hpp:181-195)
template <class T>
inline void update_coordinate(T& coeff, T x_var, T grad, T l1, T l2) {
    const T denom = x_var + l2;
    const T u = grad;
    const T v = std::abs(u) - l1;
    coeff = (v > 0.0) ? std::copysign(v, u) / denom : 0;
}

// coordinate_descent (solver_gaussian_pin_naive.hpp:26-168).
template <class T, class Iter, class Extra>
inline void coordinate_descent(PinState<T>& st, Iter begin, Iter end, size_t lmda_idx, T& convg_measure,
                               std::vector<T>& b1, std::vector<T>& b3, std::vector<T>& b4,
                               std::vector<T>& nb1, std::vector<T>& nb2, Extra additional_step)
{
    auto& X = *st.X;
    const T lmda = st.lmda_path[lmda_idx];
    const T l1 = lmda * st.alpha;
    const T l2 = lmda * (1 - st.alpha);
    convg_measure = 0;
    for (auto it = begin; it != end; ++it) {
        const idx_t ss_idx = *it;
        const idx_t k = st.screen_set[ss_idx];
        const idx_t b = st.screen_begins[ss_idx];
        const idx_t gs = st.group_sizes[k];
        ++st.n_group_updates;
        if (gs == 1) {                                                          // :75-108
            T& ak = st.screen_beta[b];
            T& gk = st.screen_grad[b];
            const T Xk_mean = st.screen_X_means[b];
            const T A_kk = st.screen_vars[b];
            const T pk = st.penalty[k];
            const T ak_old = ak;
            gk = X.cmul(st.groups[k], st.resid, st.weights) - Xk_mean * st.resid_sum * st.intercept + ak_old * A_kk;
            update_coordinate(ak, A_kk, gk, l1 * pk, l2 * pk);
            gk -= ak_old * A_kk;
            if (ak_old == ak) continue;
            const T del = ak - ak_old;
            convg_measure = std::max(A_kk * del * del, convg_measure);         // pin_base.hpp:114-124
            st.rsq += del * (2 * gk - del * A_kk);                              // pin_base.hpp:137-146
            X.ctmul(st.groups[k], -del, st.resid);
            st.resid_sum -= Xk_mean * del;
        } else {                                                                // :109-164
            T* ak = st.screen_beta + b;
            T* gk = st.screen_grad.data() + b;
            const T* Xk_mean = st.screen_X_means + b;
            const std::vector<T>& Vk = (*st.screen_transforms)[ss_idx];        // V(r,c) = Vk[r*gs+c]
            const T* A_kk = st.screen_vars + b;
            const T pk = st.penalty[k];
            X.bmul(st.groups[k], gs, st.resid, st.weights, gk);
            if (st.intercept) for (idx_t i = 0; i < gs; ++i) gk[i] -= st.resid_sum * Xk_mean[i];
            T* gk_t = b3.data();
            for (idx_t c = 0; c < gs; ++c) {                                    // gk_t = gk V
                T s = 0;
                for (idx_t r = 0; r < gs; ++r) s += gk[r] * Vk[r * gs + c];
                gk_t[c] = s;
            }
            T* ak_old = b4.data();
            T* ak_old_t = b4.data() + gs;
            T* ak_t = b4.data() + 2 * gs;
            for (idx_t i = 0; i < gs; ++i) ak_old[i] = ak[i];
            for (idx_t c = 0; c < gs; ++c) {
                T s = 0;
                for (idx_t r = 0; r < gs; ++r) s += ak_old[r] * Vk[r * gs + c];
                ak_old_t[c] = s;
                ak_t[c] = s;
            }
            for (idx_t i = 0; i < gs; ++i) gk_t[i] += A_kk[i] * ak_old_t[i];
            size_t nit;
            newton_prox(A_kk, gk_t, gs, l1 * pk, l2 * pk, st.newton_tol, st.newton_max_iters,
                        false, ak_t, nit, nb1.data(), nb2.data());             // pin_base.hpp:148-179
            if (nit >= st.newton_max_iters)
                throw solver_error("adelie_core solver: Newton-ABS max iterations reached! Try increasing newton_max_iters.");
            for (idx_t i = 0; i < gs; ++i) gk_t[i] -= A_kk[i] * ak_old_t[i];
            T dn = 0;
            for (idx_t i = 0; i < gs; ++i) { const T d = ak_old_t[i] - ak_t[i]; dn += d * d; }
            if (std::sqrt(dn) <= Configs::dbeta_tol * std::sqrt((double)gs)) continue;   // :146-147
            T* del_t = b1.data();
            T cm = 0, rs = 0;
            for (idx_t i = 0; i < gs; ++i) {
                del_t[i] = ak_t[i] - ak_old_t[i];
                cm += A_kk[i] * del_t[i] * del_t[i];
                rs += del_t[i] * (2 * gk_t[i] - del_t[i] * A_kk[i]);
            }
            convg_measure = std::max(convg_measure, cm / gs);                   // pin_base.hpp:102-112
            st.rsq += rs;                                                        // pin_base.hpp:126-135
            for (idx_t r = 0; r < gs; ++r) {                                    // ak = ak_t V^T
                T s = 0;
                for (idx_t c = 0; c < gs; ++c) s += ak_t[c] * Vk[r * gs + c];
                ak[r] = s;
            }
            T* del = b1.data();
            T rsum = 0;
            for (idx_t i = 0; i < gs; ++i) { del[i] = ak_old[i] - ak[i]; rsum += Xk_mean[i] * del[i]; }
            X.btmul(st.groups[k], gs, del, st.resid);
            st.resid_sum += rsum;
        }
        additional_step(ss_idx);
    }
}

// pin solve (solver_gaussian_pin_naive.hpp:181-401)
template <class T>
inline void pin_solve(PinState<T>& st, const std::function<void()>& check_interrupt = [](){}) {
    idx_t max_gs = 1;
    for (idx_t g = 0; g < st.G; ++g) max_gs = std::max(max_gs, st.group_sizes[g]);
    std::vector<T> b1(max_gs), b3(max_gs), b4(3 * max_gs), nb1(max_gs), nb2(max_gs);
    idx_t sb_size = st.S ? st.screen_begins[st.S - 1] + st.group_sizes[st.screen_set[st.S - 1]] : 0;
    st.screen_grad.assign(sb_size, 0);
    // active_begins / order from the incoming active set (state_gaussian_pin_base.ipp:9-36)
    st.active_begins.clear();
    size_t active_beta_size = 0;
    for (size_t i = 0; i < st.active_set_size; ++i) {
        st.active_begins.push_back(active_beta_size);
        active_beta_size += st.group_sizes[st.screen_set[st.active_set[i]]];
    }
    st.active_order.resize(st.active_set_size);
    std::iota(st.active_order.begin(), st.active_order.end(), 0);
    std::sort(st.active_order.begin(), st.active_order.end(), [&](idx_t i, idx_t j) {
        return st.groups[st.screen_set[st.active_set[i]]] < st.groups[st.screen_set[st.active_set[j]]];
    });

    auto add_active = [&](idx_t ss_idx) {                                       // :294-304
        if (!st.screen_is_active[ss_idx]) {
            if (st.active_set_size >= st.max_active_size)
                throw solver_error("adelie_core solver: Maximum number of active groups reached.");
            st.screen_is_active[ss_idx] = 1;
            st.active_set[st.active_set_size] = ss_idx;
            ++st.active_set_size;
        }
    };
    auto noop = [](idx_t) {};

    for (size_t l = 0; l < st.lmda_path.size(); ++l) {
        double screen_time = 0, active_time = 0;
        while (1) {
            double t0 = now_s();
            while (1) {                                                          // solve_active :181-215
                check_interrupt();
                ++st.iters;
                T cm;
                coordinate_descent(st, st.active_set, st.active_set + st.active_set_size, l, cm, b1, b3, b4, nb1, nb2, noop);
                if (cm < st.tol) break;
                if (st.iters >= st.max_iters)
                    throw solver_error("adelie_core solver: max coordinate descents reached at lambda index: " + std::to_string(l) + ".");
            }
            active_time += now_s() - t0;
            check_interrupt();
            ++st.iters;
            T cm;
            const size_t old_active = st.active_set_size;
            t0 = now_s();
            {
                std::vector<idx_t> all(st.S);
                std::iota(all.begin(), all.end(), 0);
                coordinate_descent(st, all.data(), all.data() + st.S, l, cm, b1, b3, b4, nb1, nb2, add_active);
            }
            screen_time += now_s() - t0;
            if (old_active < st.active_set_size) {
                for (size_t i = old_active; i < st.active_set_size; ++i) {
                    st.active_begins.push_back(active_beta_size);
                    active_beta_size += st.group_sizes[st.screen_set[st.active_set[i]]];
                }
            }
            if (cm < st.tol) break;
            if (st.iters >= st.max_iters)
                throw solver_error("adelie_core solver: max coordinate descents reached at lambda index: " + std::to_string(l) + ".");
        }
        // active_order (:360-372)
        const size_t old_sz = st.active_order.size();
        st.active_order.resize(st.active_set_size);
        std::iota(st.active_order.begin() + old_sz, st.active_order.end(), old_sz);
        std::sort(st.active_order.begin(), st.active_order.end(), [&](idx_t i, idx_t j) {
            return st.groups[st.screen_set[st.active_set[i]]] < st.groups[st.screen_set[st.active_set[j]]];
        });
        // sparsify_active_beta (pin_base.hpp:58-98)
        std::vector<idx_t> bi; std::vector<T> bv;
        bi.reserve(active_beta_size); bv.reserve(active_beta_size);
        for (size_t i = 0; i < st.active_order.size(); ++i) {
            const idx_t ss_idx = st.active_set[st.active_order[i]];
            const idx_t g = st.screen_set[ss_idx];
            const idx_t gs = st.group_sizes[g];
            for (idx_t c = 0; c < gs; ++c) {
                bi.push_back(st.groups[g] + c);
                bv.push_back(st.screen_beta[st.screen_begins[ss_idx] + c]);
            }
        }
        st.beta_idx.emplace_back(std::move(bi));
        st.beta_val.emplace_back(std::move(bv));
        st.intercepts.push_back(st.intercept * (st.y_mean + st.resid_sum));     // :392
        st.rsqs.push_back(st.rsq);
        st.lmdas.push_back(st.lmda_path[l]);
        st.benchmark_screen.push_back(screen_time);
        st.benchmark_active.push_back(active_time);
        if (st.rsq >= st.adev_tol * st.y_var) break;                            // :398
        if ((l >= 1) && (st.rsqs[l] - st.rsqs[l - 1] <= st.ddev_tol * st.y_var)) break;   // :399
    }
}

// ---------------------------------------------------------------------------
// GLM families (CORE/glm/*.ipp)
// ---------------------------------------------------------------------------
template <class T>
struct GlmBase {
    std::string name; const T* y; const T* w; idx_t n; bool is_multi = false;
    virtual ~GlmBase() {}
    virtual void gradient(const T* eta, T* grad) = 0;
    virtual void hessian(const T* eta, const T* grad, T* hess) = 0;
    virtual void inv_hessian_gradient(const T* eta, const T* grad, const T* hess, T* out) {   // glm_base.ipp:25-36
        (void)eta;
        for (idx_t i = 0; i < n; ++i)
            out[i] = grad[i] / (std::max<T>(hess[i], 0) + T(Configs::hessian_min) * T(hess[i] <= 0));
    }
    virtual T loss(const T* eta) = 0;
    virtual T loss_full() = 0;
    virtual void inv_link(const T* eta, T* out) = 0;
};

template <class T>
struct GlmGaussian : GlmBase<T> {                                              // glm_gaussian.ipp:17-64
    using B = GlmBase<T>;
    GlmGaussian(const T* y, const T* w, idx_t n) { B::name = "gaussian"; B::y = y; B::w = w; B::n = n; }
    void gradient(const T* eta, T* grad) override { for (idx_t i = 0; i < B::n; ++i) grad[i] = B::w[i] * (B::y[i] - eta[i]); }
    void hessian(const T*, const T*, T* hess) override { for (idx_t i = 0; i < B::n; ++i) hess[i] = B::w[i]; }
    T loss(const T* eta) override {
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) s += B::w[i] * (T(0.5) * eta[i] * eta[i] - B::y[i] * eta[i]);
        return s;
    }
    T loss_full() override {
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) s += B::y[i] * B::y[i] * B::w[i];
        return T(-0.5) * s;
    }
    void inv_link(const T* eta, T* out) override { for (idx_t i = 0; i < B::n; ++i) out[i] = eta[i]; }
};

template <class T>
inline T binomial_loss_full(const T* y, const T* w, idx_t n) {                 // glm_binomial.ipp:14-35
    T loss = 0;
    for (idx_t i = 0; i < n; ++i) {
        const T ly = std::log(y[i]);
        const T l1my = std::log(1 - y[i]);
        if (!(std::isinf(ly) || std::isnan(ly))) loss -= w[i] * y[i] * ly;
        if (!(std::isinf(l1my) || std::isnan(l1my))) loss -= w[i] * (1 - y[i]) * l1my;
    }
    return loss;
}

template <class T>
struct GlmBinomialLogit : GlmBase<T> {                                         // glm_binomial.ipp:47-98
    using B = GlmBase<T>;
    GlmBinomialLogit(const T* y, const T* w, idx_t n) { B::name = "binomial_logit"; B::y = y; B::w = w; B::n = n; }
    void gradient(const T* eta, T* grad) override {
        for (idx_t i = 0; i < B::n; ++i) grad[i] = B::w[i] * (B::y[i] - 1 / (1 + std::exp(-eta[i])));
    }
    void hessian(const T*, const T* grad, T* hess) override {
        for (idx_t i = 0; i < B::n; ++i) {
            const T h = B::w[i] * B::y[i] - grad[i];
            hess[i] = (h * (B::w[i] - h)) / (B::w[i] + T(B::w[i] <= 0));
        }
    }
    T loss(const T* eta) override {
        constexpr T mx = std::numeric_limits<T>::max();
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) {
            const T e = std::max(std::min(eta[i], mx), -mx);
            s += B::w[i] * ((T(eta[i] > 0) - B::y[i]) * e + std::log(1 + std::exp(-std::abs(eta[i]))));
        }
        return s;
    }
    T loss_full() override { return binomial_loss_full(B::y, B::w, B::n); }
    void inv_link(const T* eta, T* out) override { for (idx_t i = 0; i < B::n; ++i) out[i] = 1 / (1 + std::exp(-eta[i])); }
};

// Multinomial (glm_multinomial.ipp:6-132): y, eta (n,K) row-major; softmax with the row maximum subtracted; the hessian is the
// diagonal majorant 2 K^-1 w p (1 - p) written through the gradient as in the reference (:52-63).
template <class T>
struct GlmMultinomial : GlmBase<T> {
    using B = GlmBase<T>;
    idx_t K;
    GlmMultinomial(const T* y, const T* w, idx_t n, idx_t K_) : K(K_) {
        B::name = "multinomial"; B::y = y; B::w = w; B::n = n; B::is_multi = true;
        if (K_ <= 1) throw std::runtime_error("adelie_core: y must have at least 2 columns (classes).");
    }
    void softmax_row(const T* e, T* p) const {
        T m = e[0]; for (idx_t k = 1; k < K; ++k) m = std::max(m, e[k]);
        T sum = 0; for (idx_t k = 0; k < K; ++k) { p[k] = std::exp(e[k] - m); sum += p[k]; }
        for (idx_t k = 0; k < K; ++k) p[k] /= sum;
    }
    void gradient(const T* eta, T* grad) override {
        std::vector<T> p(K);
        for (idx_t i = 0; i < B::n; ++i) {
            softmax_row(eta + i * K, p.data());
            for (idx_t k = 0; k < K; ++k) grad[i * K + k] = (B::y[i * K + k] - p[k]) * B::w[i] / K;
        }
    }
    void hessian(const T*, const T* grad, T* hess) override {
        for (idx_t i = 0; i < B::n; ++i)
            for (idx_t k = 0; k < K; ++k) {
                const T h = B::y[i * K + k] * B::w[i] / K - grad[i * K + k];
                hess[i * K + k] = h * 2 * (1 - K * (h / (B::w[i] + T(B::w[i] <= 0))));
            }
    }
    void inv_hessian_gradient(const T*, const T* grad, const T* hess, T* out) override {     // glm_multibase.ipp:25-37
        for (idx_t i = 0; i < B::n * K; ++i)
            out[i] = grad[i] / (std::max<T>(hess[i], 0) + T(Configs::hessian_min) * T(hess[i] <= 0));
    }
    T loss(const T* eta) override {
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) {
            const T* e = eta + i * K;
            T m = e[0]; for (idx_t k = 1; k < K; ++k) m = std::max(m, e[k]);
            T ye = 0, se = 0;
            for (idx_t k = 0; k < K; ++k) { ye += B::y[i * K + k] * (e[k] - m); se += std::exp(e[k] - m); }
            s += B::w[i] * (-ye + std::log(se));
        }
        return s / K;
    }
    T loss_full() override {
        T loss = 0;
        for (idx_t i = 0; i < B::n; ++i) {
            T sum = 0;
            for (idx_t k = 0; k < K; ++k) { const T l = std::log(B::y[i * K + k]); if (!(std::isinf(l) || std::isnan(l))) sum += B::y[i * K + k] * l; }
            loss -= sum * B::w[i] / K;
        }
        return loss;
    }
    void inv_link(const T* eta, T* out) override { for (idx_t i = 0; i < B::n; ++i) softmax_row(eta + i * K, out + i * K); }
};

// Binomial, probit link (glm_binomial.ipp:100-190): Phi = 0.5 (1 + erf(eta / sqrt 2)), phi = exp(-eta^2 / 2) / sqrt(2 pi)
template <class T>
struct GlmBinomialProbit : GlmBase<T> {
    using B = GlmBase<T>;
    GlmBinomialProbit(const T* y, const T* w, idx_t n) { B::name = "binomial_probit"; B::y = y; B::w = w; B::n = n; }
    static T cdf(T x) { return T(0.5) * (1 + std::erf(x / T(M_SQRT2))); }
    static T pdf(T x) { return T(0.5 * M_2_SQRTPI / M_SQRT2) * std::exp(T(-0.5) * x * x); }
    void gradient(const T* eta, T* grad) override {
        constexpr T mx = std::numeric_limits<T>::max();
        for (idx_t i = 0; i < B::n; ++i) {
            const T P = cdf(eta[i]);
            grad[i] = B::w[i] * pdf(eta[i]) * (B::y[i] * std::min(1 / P, mx) - (1 - B::y[i]) * std::min(1 / (1 - P), mx));
        }
    }
    void hessian(const T* eta, const T* grad, T* hess) override {
        constexpr T mx = std::numeric_limits<T>::max();
        for (idx_t i = 0; i < B::n; ++i) {
            const T P = cdf(eta[i]), ph = pdf(eta[i]);
            hess[i] = B::w[i] * (B::y[i] * std::min(1 / (P * P), mx) + (1 - B::y[i]) * std::min(1 / ((1 - P) * (1 - P)), mx)) * ph * ph + eta[i] * grad[i];
        }
    }
    T loss(const T* eta) override {
        constexpr T mx = std::numeric_limits<T>::max();
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) {
            const T P = cdf(eta[i]);
            s += B::w[i] * (B::y[i] * std::max(std::log(P), -mx) + (1 - B::y[i]) * std::max(std::log(1 - P), -mx));
        }
        return -s;
    }
    T loss_full() override { return binomial_loss_full(B::y, B::w, B::n); }
    void inv_link(const T* eta, T* out) override { for (idx_t i = 0; i < B::n; ++i) out[i] = cdf(eta[i]); }
};

// Poisson, log link (glm_poisson.ipp:7-66)
template <class T>
struct GlmPoisson : GlmBase<T> {
    using B = GlmBase<T>;
    GlmPoisson(const T* y, const T* w, idx_t n) { B::name = "poisson"; B::y = y; B::w = w; B::n = n; }
    void gradient(const T* eta, T* grad) override { for (idx_t i = 0; i < B::n; ++i) grad[i] = B::w[i] * (B::y[i] - std::exp(eta[i])); }
    void hessian(const T*, const T* grad, T* hess) override { for (idx_t i = 0; i < B::n; ++i) hess[i] = B::w[i] * B::y[i] - grad[i]; }
    T loss(const T* eta) override {                                             // stable when y == 0 and eta = -inf (:44-47)
        constexpr T mx = std::numeric_limits<T>::max();
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) s += B::w[i] * (std::min(-eta[i], mx) * B::y[i] + std::exp(eta[i]));
        return s;
    }
    T loss_full() override {
        constexpr T mx = std::numeric_limits<T>::max();
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) s += B::w[i] * (std::min(-std::log(B::y[i]), mx) * B::y[i] + B::y[i]);
        return s;
    }
    void inv_link(const T* eta, T* out) override { for (idx_t i = 0; i < B::n; ++i) out[i] = std::exp(eta[i]); }
};

// MultiGaussian (glm_multigaussian.ipp:17-68): y, eta are (n,K) row-major;
// everything is the Gaussian family divided by K.
template <class T>
struct GlmMultiGaussian : GlmBase<T> {
    using B = GlmBase<T>;
    idx_t K;
    GlmMultiGaussian(const T* y, const T* w, idx_t n, idx_t K_) : K(K_) {
        B::name = "multigaussian"; B::y = y; B::w = w; B::n = n; B::is_multi = true;
    }
    void gradient(const T* eta, T* grad) override {
        for (idx_t i = 0; i < B::n; ++i)
            for (idx_t k = 0; k < K; ++k) grad[i * K + k] = B::w[i] * (B::y[i * K + k] - eta[i * K + k]) / K;
    }
    void hessian(const T*, const T*, T* hess) override {
        for (idx_t i = 0; i < B::n; ++i) for (idx_t k = 0; k < K; ++k) hess[i * K + k] = B::w[i] / K;
    }
    void inv_hessian_gradient(const T*, const T* grad, const T* hess, T* out) override {
        for (idx_t i = 0; i < B::n * K; ++i)
            out[i] = grad[i] / (std::max<T>(hess[i], 0) + T(Configs::hessian_min) * T(hess[i] <= 0));
    }
    T loss(const T* eta) override {
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) {
            T r = 0;
            for (idx_t k = 0; k < K; ++k) r += T(0.5) * eta[i * K + k] * eta[i * K + k] - B::y[i * K + k] * eta[i * K + k];
            s += B::w[i] * r;
        }
        return s / K;
    }
    T loss_full() override {
        T s = 0;
        for (idx_t i = 0; i < B::n; ++i) {
            T r = 0;
            for (idx_t k = 0; k < K; ++k) r += B::y[i * K + k] * B::y[i * K + k];
            s += B::w[i] * r;
        }
        return T(-0.5) * s / K;
    }
    void inv_link(const T* eta, T* out) override { for (idx_t i = 0; i < B::n * K; ++i) out[i] = eta[i]; }
};

// ---------------------------------------------------------------------------
// Path state (CORE/state/state_base.hpp:35-216, state_gaussian_naive.hpp,
// state_glm_naive.hpp) and drivers (CORE/solver/solver_base.hpp,
// solver_gaussian_naive.hpp, solver_glm_naive.hpp).
// ---------------------------------------------------------------------------
template <class T>
struct PathState {
    // static
    MatrixBase<T>* X = nullptr;
    idx_t n = 0, p = 0, G = 0;
    const idx_t* groups = nullptr; const idx_t* group_sizes = nullptr;
    T alpha = 1; const T* penalty = nullptr;
    const T* weights = nullptr;          // gaussian
    std::vector<T> weights_sqrt;
    std::vector<T> X_means;               // (p,)
    T y_mean = 0, y_var = 0, loss_null = 0, loss_full = 0;
    // GLM extras
    GlmBase<T>* glm = nullptr;
    const T* offsets = nullptr;
    std::vector<T> eta; T beta0 = 0;
    size_t irls_max_iters = 10000; T irls_tol = 1e-7; bool setup_loss_null = true;
    // configs
    T min_ratio = 1e-2; size_t lmda_path_size = 100;
    size_t max_screen_size = 0, max_active_size = 0;
    T pivot_subset_ratio = 0.1; size_t pivot_subset_min = 1; T pivot_slack_ratio = 1.25;
    int screen_rule = 1;  // 0 strong, 1 pivot
    size_t max_iters = 100000; T tol = 1e-7, adev_tol = 0.9, ddev_tol = 0, newton_tol = 1e-12; size_t newton_max_iters = 1000;
    bool early_exit = true, setup_lmda_max = true, setup_lmda_path = true, intercept = true;
    int n_threads = 1;
    // dynamic
    T lmda_max = -1; std::vector<T> lmda_path;
    std::unordered_set<idx_t> screen_hashset;
    std::vector<idx_t> screen_set, screen_begins;
    std::vector<T> screen_beta; std::vector<int8_t> screen_is_active;
    size_t active_set_size = 0; std::vector<idx_t> active_set;
    T lmda = std::numeric_limits<T>::infinity();
    std::vector<T> grad, abs_grad;
    std::vector<T> resid; T resid_sum = 0, rsq = 0;
    std::vector<T> screen_X_means, screen_vars; std::vector<std::vector<T>> screen_transforms;
    // outputs
    std::vector<std::vector<idx_t>> beta_idx; std::vector<std::vector<T>> beta_val;
    std::vector<T> intercepts, devs, lmdas;
    std::vector<double> benchmark_screen, benchmark_fit_screen, benchmark_fit_active, benchmark_kkt, benchmark_invariance;
    std::vector<int> n_valid_solutions, active_sizes, screen_sizes;
    size_t n_sweeps = 0, n_group_updates = 0, n_irls = 0;
    double max_seconds = -1;    // oracle-only: stop early after this many seconds (bounded CPU baseline sample)
    double t_start = 0;
};

// update_abs_grad (solver_base.hpp:20-110), constraints == nullptr everywhere.
template <class T>
inline void update_abs_grad(PathState<T>& s, T lmda) {
    for (size_t ss = 0; ss < s.screen_set.size(); ++ss) {
        const idx_t i = s.screen_set[ss], b = s.screen_begins[ss], k = s.groups[i], sz = s.group_sizes[i];
        const T regul = ((1 - s.alpha) * lmda) * s.penalty[i];
        T a = 0;
        for (idx_t c = 0; c < sz; ++c) { const T e = s.grad[k + c] - regul * s.screen_beta[b + c]; a += e * e; }
        s.abs_grad[i] = std::sqrt(a);
    }
    for (idx_t i = 0; i < s.G; ++i) {
        if (s.screen_hashset.count(i)) continue;
        const idx_t k = s.groups[i], sz = s.group_sizes[i];
        T a = 0;
        for (idx_t c = 0; c < sz; ++c) a += s.grad[k + c] * s.grad[k + c];
        s.abs_grad[i] = std::sqrt(a);
    }
}

// update_screen_derived_base (solver_base.hpp:120-153)
template <class T>
inline void update_screen_derived_base(PathState<T>& s) {
    const size_t old = s.screen_begins.size();
    for (size_t i = old; i < s.screen_set.size(); ++i) s.screen_hashset.insert(s.screen_set[i]);
    size_t vs = (old == 0) ? 0 : (s.screen_begins.back() + s.group_sizes[s.screen_set[old - 1]]);
    for (size_t i = old; i < s.screen_set.size(); ++i) { s.screen_begins.push_back(vs); vs += s.group_sizes[s.screen_set[i]]; }
    s.screen_beta.resize(vs, 0);
    s.screen_is_active.resize(s.screen_set.size(), 0);
}

// update_screen_derived (solver_gaussian_naive.hpp:53-125) on positions [begin, end)
template <class T>
inline void update_screen_derived_range(
    PathState<T>& s, const T* X_means, const T* weights_sqrt, size_t begin, size_t end,
    std::vector<T>& screen_X_means, std::vector<std::vector<T>>& screen_transforms, std::vector<T>& screen_vars)
{
    const size_t S = s.screen_set.size();
    const size_t vs = S ? (s.screen_begins.back() + s.group_sizes[s.screen_set.back()]) : 0;
    screen_X_means.resize(vs);
    screen_transforms.resize(S);
    screen_vars.resize(vs, 0);
    #pragma omp parallel for schedule(static) num_threads(s.n_threads) if (s.n_threads > 1 && (begin + s.n_threads) <= end)
    for (size_t i = begin; i < end; ++i) {
        const idx_t g = s.groups[s.screen_set[i]], gs = s.group_sizes[s.screen_set[i]], sb = s.screen_begins[i];
        for (idx_t c = 0; c < gs; ++c) screen_X_means[sb + c] = X_means[g + c];
        std::vector<T> C(gs * gs);
        s.X->cov(g, gs, weights_sqrt, C.data());
        if (s.intercept) {
            for (idx_t a = 0; a < gs; ++a)
                for (idx_t b = 0; b < gs; ++b) C[a + b * gs] -= screen_X_means[sb + a] * screen_X_means[sb + b];
        }
        if (gs == 1) {
            screen_transforms[i].assign(1, T(1));
            screen_vars[sb] = std::max<T>(C[0], 0);
            continue;
        }
        std::vector<T> D(gs), V(gs * gs);
        jacobi_eigh(C.data(), gs, D.data(), V.data());
        // store row-major (r,c) -> V[r + c*gs] col-major to [r*gs + c]
        std::vector<T> Vr(gs * gs);
        for (idx_t r = 0; r < gs; ++r) for (idx_t c = 0; c < gs; ++c) Vr[r * gs + c] = V[r + c * gs];
        screen_transforms[i] = std::move(Vr);
        for (idx_t c = 0; c < gs; ++c) screen_vars[sb + c] = D[c] * T(D[c] >= 0);       // :122
    }
}

// screen (solver_base.hpp:273-403)
template <class T>
inline void screen(PathState<T>& s, T lmda_next, bool all_kkt_passed, int n_new_active) {
    const idx_t G = s.G;
    const int old_size = (int)s.screen_set.size();
    auto is_screen = [&](idx_t i) { return s.screen_hashset.count(i) > 0; };
    if (s.screen_rule == 0) {
        const T strong = (2 * lmda_next - s.lmda) * s.alpha;
        for (idx_t i = 0; i < G; ++i) {
            if (is_screen(i)) continue;
            if (s.abs_grad[i] > strong * s.penalty[i]) s.screen_set.push_back(i);
        }
    } else {
        if (n_new_active) {
            std::vector<idx_t> order(G);
            std::iota(order.begin(), order.end(), 0);
            std::vector<T> wts(G);
            for (idx_t i = 0; i < G; ++i)
                wts[i] = (s.penalty[i] <= 0) ? s.alpha * s.lmda : std::min(s.abs_grad[i] / s.penalty[i], s.alpha * s.lmda);
            std::sort(order.begin(), order.end(), [&](idx_t i, idx_t j) { return wts[i] < wts[j]; });
            const int subset_size = std::min<int>(std::max<int>(
                (int)(old_size * (1 + s.pivot_subset_ratio)), (int)s.pivot_subset_min), (int)G);
            std::vector<T> ws(subset_size), mses(subset_size), ind(subset_size);
            for (int i = 0; i < subset_size; ++i) { ws[i] = wts[order[G - subset_size + i]]; ind[i] = (T)i; }
            const int pivot_idx = search_pivot(ind.data(), ws.data(), (idx_t)subset_size, mses.data());
            const int full_pivot_idx = (int)G - subset_size + pivot_idx;
            for (int ii = (int)G - 1; ii >= full_pivot_idx; --ii) {
                const idx_t i = order[ii];
                if (is_screen(i)) continue;
                s.screen_set.push_back(i);
            }
            int count = 0;
            for (int ii = full_pivot_idx - 1; ii >= 0; --ii) {
                if (count >= s.pivot_slack_ratio * n_new_active) break;
                const idx_t i = order[ii];
                if (is_screen(i)) continue;
                s.screen_set.push_back(i);
                ++count;
            }
        }
        if (((int)s.screen_set.size() == old_size) && !all_kkt_passed) {
            for (idx_t i = 0; i < G; ++i) {
                if (is_screen(i)) continue;
                if (s.abs_grad[i] > lmda_next * s.penalty[i] * s.alpha) s.screen_set.push_back(i);
            }
        }
    }
    if (s.screen_set.size() > s.max_screen_size) {
        s.screen_set.resize(old_size);
        throw solver_error("adelie_core solver: maximum screen set size reached.");
    }
}

template <class T>
inline bool kkt(PathState<T>& s, T lmda) {                                     // solver_base.hpp:408-433
    for (idx_t k = 0; k < s.G; ++k) {
        if (s.screen_hashset.count(k)) continue;
        if (s.abs_grad[k] > lmda * s.alpha * s.penalty[k]) return false;
    }
    return true;
}

template <class T>
inline bool early_exit(const PathState<T>& s) {                                // solver_base.hpp:241-263
    if (!s.early_exit || s.devs.empty()) return false;
    const T u = s.devs.back();
    if (u >= s.adev_tol) return true;
    if (s.devs.size() == 1) return false;
    const T m = s.devs[s.devs.size() - 2];
    if (std::abs(u - m) < s.ddev_tol) return true;
    return false;
}

// Build a pin state view over the path state (solver_gaussian_naive.hpp:284-320).
template <class T>
inline PinState<T> make_pin(PathState<T>& s, const T* weights, T y_mean, T y_var, T lmda, T tol, T adev, T ddev,
                            T rsq, T* resid, T resid_sum,
                            const std::vector<T>& sXm, const std::vector<T>& sv, const std::vector<std::vector<T>>& st)
{
    PinState<T> ps;
    ps.X = s.X; ps.y_mean = y_mean; ps.y_var = y_var;
    ps.groups = s.groups; ps.group_sizes = s.group_sizes; ps.G = s.G;
    ps.alpha = s.alpha; ps.penalty = s.penalty; ps.weights = weights;
    ps.screen_set = s.screen_set.data(); ps.screen_begins = s.screen_begins.data(); ps.S = (idx_t)s.screen_set.size();
    ps.screen_vars = sv.data(); ps.screen_X_means = sXm.data(); ps.screen_transforms = &st;
    ps.lmda_path = {lmda};
    ps.intercept = s.intercept; ps.max_active_size = s.max_active_size; ps.max_iters = s.max_iters;
    ps.tol = tol; ps.adev_tol = adev; ps.ddev_tol = ddev; ps.newton_tol = s.newton_tol; ps.newton_max_iters = s.newton_max_iters;
    ps.rsq = rsq; ps.resid = resid; ps.resid_sum = resid_sum;
    ps.screen_beta = s.screen_beta.data(); ps.screen_is_active = s.screen_is_active.data();
    ps.active_set_size = s.active_set_size; ps.active_set = s.active_set.data();
    return ps;
}

// Gaussian fit (solver_gaussian_naive.hpp:215-349)
template <class T>
inline PinState<T> fit_gaussian(PathState<T>& s, T lmda, double& screen_time, double& active_time) {
    std::vector<T> resid_prev = s.resid;
    std::vector<T> beta_prev = s.screen_beta;
    std::vector<int8_t> act_prev = s.screen_is_active;
    PinState<T> ps = make_pin(s, s.weights, s.y_mean, s.y_var, lmda, s.tol * s.y_var, s.adev_tol, s.ddev_tol,
                              s.rsq, s.resid.data(), s.resid_sum, s.screen_X_means, s.screen_vars, s.screen_transforms);
    try {
        pin_solve(ps);
    } catch (...) {
        s.resid.swap(resid_prev); s.screen_beta.swap(beta_prev); s.screen_is_active.swap(act_prev);
        throw;
    }
    s.resid_sum = ps.resid_sum; s.rsq = ps.rsq; s.active_set_size = ps.active_set_size;
    screen_time = std::accumulate(ps.benchmark_screen.begin(), ps.benchmark_screen.end(), 0.0);
    active_time = std::accumulate(ps.benchmark_active.begin(), ps.benchmark_active.end(), 0.0);
    s.n_sweeps += ps.iters; s.n_group_updates += ps.n_group_updates;
    return ps;
}

// GLM IRLS fit (solver_glm_naive.hpp:241-459)
template <class T>
struct GlmBuffers {
    std::vector<T> X_means, irls_weights, irls_weights_sqrt, irls_y, irls_resid, resid_prev, eta_prev, hess, ones;
    std::vector<T> screen_X_means, screen_vars; std::vector<std::vector<T>> screen_transforms;
    GlmBuffers(idx_t n, idx_t p) : X_means(p), irls_weights(n), irls_weights_sqrt(n), irls_y(n), irls_resid(n),
        resid_prev(n), eta_prev(n), hess(n), ones(n, T(1)) {}
};

template <class T>
inline PinState<T> fit_glm(PathState<T>& s, GlmBuffers<T>& B, T lmda, double& screen_time, double& active_time) {
    auto& glm = *s.glm;
    const idx_t n = s.n;
    screen_time = 0; active_time = 0;
    size_t irls_it = 0;
    while (1) {
        if (irls_it >= s.irls_max_iters) throw solver_error("adelie_core solver: Maximum IRLS iterations reached.");
        std::vector<T> beta_prev = s.screen_beta;
        std::vector<int8_t> act_prev = s.screen_is_active;
        glm.hessian(s.eta.data(), s.resid.data(), B.hess.data());
        glm.inv_hessian_gradient(s.eta.data(), s.resid.data(), B.hess.data(), B.irls_resid.data());
        T hess_sum = 0;
        for (idx_t i = 0; i < n; ++i) {
            B.hess[i] = std::max<T>(B.hess[i], 0) + T(Configs::hessian_min) * T(B.hess[i] <= 0);
            hess_sum += B.hess[i];
        }
        T y_mean = 0;
        for (idx_t i = 0; i < n; ++i) {
            B.irls_weights[i] = B.hess[i] / hess_sum;
            B.irls_weights_sqrt[i] = std::sqrt(B.irls_weights[i]);
            B.irls_y[i] = B.irls_resid[i] + s.eta[i] - s.offsets[i];
            y_mean += B.irls_weights[i] * B.irls_y[i];
        }
        T y_var = 0;
        for (idx_t i = 0; i < n; ++i) y_var += B.irls_weights[i] * B.irls_y[i] * B.irls_y[i];
        y_var -= s.intercept * y_mean * y_mean;
        if (s.intercept) for (idx_t i = 0; i < n; ++i) B.irls_resid[i] += (s.beta0 - y_mean);
        T resid_sum = 0;
        for (idx_t i = 0; i < n; ++i) resid_sum += B.irls_weights[i] * B.irls_resid[i];
        T lmda_adj = lmda / hess_sum;
        if (std::isinf(lmda_adj)) {
            if (lmda == std::numeric_limits<T>::max()) lmda_adj = lmda;
            else throw solver_error("adelie_core solver: IRLS lambda is unexpectedly inf. This likely indicates a bug in the code. Please report this!");
        }
        for (size_t ss = 0; ss < s.screen_set.size(); ++ss) {                  // update_X_means :361-372
            const idx_t i = s.screen_set[ss], g = s.groups[i], gs = s.group_sizes[i];
            if (gs == 1) B.X_means[g] = s.X->cmul(g, B.ones.data(), B.irls_weights.data());
            else s.X->bmul(g, gs, B.ones.data(), B.irls_weights.data(), B.X_means.data() + g);
        }
        update_screen_derived_range(s, B.X_means.data(), B.irls_weights_sqrt.data(), 0, s.screen_set.size(),
                                    B.screen_X_means, B.screen_transforms, B.screen_vars);
        PinState<T> ps = make_pin(s, B.irls_weights.data(), y_mean, y_var, lmda_adj,
                                  s.tol * (s.loss_null - s.loss_full) / hess_sum, T(0), T(0), T(0),
                                  B.irls_resid.data(), resid_sum, B.screen_X_means, B.screen_vars, B.screen_transforms);
        try {
            pin_solve(ps);
        } catch (...) {
            s.screen_beta.swap(beta_prev); s.screen_is_active.swap(act_prev);
            throw;
        }
        screen_time += std::accumulate(ps.benchmark_screen.begin(), ps.benchmark_screen.end(), 0.0);
        active_time += std::accumulate(ps.benchmark_active.begin(), ps.benchmark_active.end(), 0.0);
        s.n_sweeps += ps.iters; s.n_group_updates += ps.n_group_updates; ++s.n_irls;
        s.active_set_size = ps.active_set_size;
        s.beta0 = ps.intercepts[0];
        s.eta.swap(B.eta_prev);
        for (idx_t i = 0; i < n; ++i) {
            s.eta[i] = B.irls_y[i] + s.offsets[i] - B.irls_resid[i];
            if (s.intercept) s.eta[i] += s.beta0 - y_mean;
        }
        B.resid_prev.swap(s.resid);
        glm.gradient(s.eta.data(), s.resid.data());
        T conv = 0;
        for (idx_t i = 0; i < n; ++i) conv += (s.resid[i] - B.resid_prev[i]) * (s.eta[i] - B.eta_prev[i]);
        if (std::abs(conv) <= s.irls_tol) return ps;
        ++irls_it;
    }
}

// update_loss_null (solver_glm_naive.hpp:165-232), single response
template <class T>
inline void update_loss_null(PathState<T>& s, GlmBuffers<T>& B) {
    auto& glm = *s.glm;
    const idx_t n = s.n;
    if (!s.intercept) { s.loss_null = glm.loss(s.offsets); return; }
    T beta0 = s.beta0;
    std::vector<T> eta = s.eta, resid = s.resid;
    size_t it = 0;
    while (1) {
        if (it >= s.irls_max_iters) throw solver_error("adelie_core solver: Maximum IRLS iterations reached.");
        glm.hessian(eta.data(), resid.data(), B.hess.data());
        glm.inv_hessian_gradient(eta.data(), resid.data(), B.hess.data(), B.irls_y.data());
        T hs = 0;
        for (idx_t i = 0; i < n; ++i) {
            B.hess[i] = std::max<T>(B.hess[i], 0) + T(Configs::hessian_min) * T(B.hess[i] <= 0);
            hs += B.hess[i];
        }
        T num = 0;
        for (idx_t i = 0; i < n; ++i) num += B.hess[i] * (B.irls_y[i] + eta[i] - s.offsets[i]);
        beta0 = num / hs;
        eta.swap(B.eta_prev);
        for (idx_t i = 0; i < n; ++i) eta[i] = beta0 + s.offsets[i];
        B.resid_prev.swap(resid);
        glm.gradient(eta.data(), resid.data());
        T conv = 0;
        for (idx_t i = 0; i < n; ++i) conv += (resid[i] - B.resid_prev[i]) * (eta[i] - B.eta_prev[i]);
        if (std::abs(conv) <= s.irls_tol) { s.loss_null = glm.loss(eta.data()); return; }
        ++it;
    }
}

// solve_core (solver_base.hpp:435-687).  `is_glm` selects the GLM lambdas
// (solver_glm_naive.hpp:470-546) vs Gaussian (solver_gaussian_naive.hpp:358-434).
template <class T>
inline void solve_path(PathState<T>& s, bool is_glm) {
    s.t_start = now_s();
    GlmBuffers<T> B(is_glm ? s.n : 0, is_glm ? s.p : 0);
    auto fit = [&](T lmda, double& st, double& at) {
        return is_glm ? fit_glm(s, B, lmda, st, at) : fit_gaussian(s, lmda, st, at);
    };
    auto update_invariance = [&](T lmda) {
        s.lmda = lmda;
        if (is_glm) {
            s.X->mul(s.resid.data(), B.ones.data(), s.grad.data());            // solver_glm_naive.hpp:495-503
        } else {
            s.X->mul(s.resid.data(), s.weights, s.grad.data());                // solver_gaussian_naive.hpp:377-393
            if (s.intercept) for (idx_t j = 0; j < s.p; ++j) s.grad[j] -= s.resid_sum * s.X_means[j];
        }
        update_abs_grad(s, lmda);
    };
    auto update_solutions = [&](PinState<T>& ps, T lmda) {
        s.beta_idx.emplace_back(std::move(ps.beta_idx.back()));
        s.beta_val.emplace_back(std::move(ps.beta_val.back()));
        s.intercepts.push_back(ps.intercepts.back());
        s.lmdas.push_back(lmda);
        if (is_glm) {
            const T loss = s.glm->loss(s.eta.data());
            s.devs.push_back((s.loss_null - loss) / (s.loss_null - s.loss_full));  // solver_glm_naive.hpp:153-157
        } else {
            s.devs.push_back(ps.rsqs.back() / s.y_var);                          // solver_gaussian_naive.hpp:205-206
        }
    };
    auto screen_f = [&](T lmda, bool kkt_passed, int n_new_active) {
        screen(s, lmda, kkt_passed, n_new_active);
        if (is_glm) {
            update_screen_derived_base(s);
        } else {
            const size_t old = s.screen_transforms.size();
            update_screen_derived_base(s);
            update_screen_derived_range(s, s.X_means.data(), s.weights_sqrt.data(), old, s.screen_set.size(),
                                        s.screen_X_means, s.screen_transforms, s.screen_vars);
        }
    };
    auto budget_exceeded = [&]() { return s.max_seconds > 0 && (now_s() - s.t_start) > s.max_seconds; };

    if (s.screen_set.size() > s.max_screen_size) throw solver_error("adelie_core solver: maximum screen set size reached.");
    if (is_glm && s.setup_loss_null) update_loss_null(s, B);

    double st, at;
    if (s.setup_lmda_max) {
        T pmax = s.penalty[0];
        for (idx_t i = 1; i < s.G; ++i) pmax = std::max(pmax, s.penalty[i]);
        const T large_lmda = T(1e-3 * std::numeric_limits<T>::max() / std::max<T>(1, pmax));
        fit(large_lmda, st, at);
        update_invariance(large_lmda);
        const T factor = (s.alpha <= 0) ? T(1e-3) : s.alpha;                     // solver/utils.hpp:6-23
        T m = -std::numeric_limits<T>::infinity();
        for (idx_t i = 0; i < s.G; ++i) m = std::max<T>(m, (s.penalty[i] <= 0.0) ? T(0.0) : s.abs_grad[i] / s.penalty[i]);
        s.lmda_max = m / factor;
    }
    if (s.setup_lmda_path) {
        if (s.lmda_path_size <= 0) return;
        s.lmda_path.resize(s.lmda_path_size);
        const size_t L = s.lmda_path_size;
        if (L > 1) {                                                             // solver/utils.hpp:25-41
            const T log_factor = std::log(s.min_ratio) / (L - 1);
            for (size_t i = 0; i < L; ++i) s.lmda_path[i] = s.lmda_max * std::exp(log_factor * T(i));
        }
        s.lmda_path[0] = s.lmda_max;
    }
    size_t large_sz = 0;
    while (large_sz < s.lmda_path.size() && !(s.lmda_path[large_sz] <= s.lmda_max)) ++large_sz;
    if (large_sz || s.setup_lmda_max) {
        std::vector<T> large(s.lmda_path.begin(), s.lmda_path.begin() + large_sz);
        large.push_back(s.lmda_max);
        for (size_t i = 0; i < large.size(); ++i) {
            PinState<T> ps = fit(large[i], st, at);
            if (i + 1 < large.size()) {
                update_solutions(ps, large[i]);
                if (early_exit(s)) return;
            } else {
                update_invariance(large[i]);
            }
        }
    }
    size_t idx = large_sz;
    int current_active = (int)s.active_set_size;
    bool kkt_passed = true;
    int n_new_active = 0;
    while (idx < s.lmda_path.size()) {
        const T lmda_curr = s.lmda_path[idx];
        while (1) {
            double t0 = now_s();
            screen_f(lmda_curr, kkt_passed, n_new_active);
            s.benchmark_screen.push_back(now_s() - t0);
            PinState<T> ps = fit(lmda_curr, st, at);
            s.benchmark_fit_screen.push_back(st);
            s.benchmark_fit_active.push_back(at);
            t0 = now_s();
            update_invariance(lmda_curr);
            s.benchmark_invariance.push_back(now_s() - t0);
            t0 = now_s();
            kkt_passed = kkt(s, lmda_curr);
            s.n_valid_solutions.push_back(kkt_passed);
            idx += kkt_passed;
            if (kkt_passed) update_solutions(ps, lmda_curr);
            s.benchmark_kkt.push_back(now_s() - t0);
            if (kkt_passed) {
                s.active_sizes.push_back((int)s.active_set_size);
                s.screen_sizes.push_back((int)s.screen_set.size());
            }
            n_new_active = kkt_passed ? (s.active_sizes.back() - current_active) : n_new_active;
            current_active = kkt_passed ? s.active_sizes.back() : current_active;
            if (kkt_passed) break;
        }
        if (early_exit(s)) break;
        if (budget_exceeded()) break;
    }
}

// State construction (state_base.ipp:94-99; state_gaussian_naive.hpp:138-150, .ipp:9-28)
template <class T>
inline void init_path_state(PathState<T>& s, bool is_glm) {
    s.abs_grad.assign(s.G, 0);
    update_screen_derived_base(s);
    update_abs_grad(s, s.lmda);
    if (!is_glm) {
        s.weights_sqrt.resize(s.n);
        for (idx_t i = 0; i < s.n; ++i) s.weights_sqrt[i] = std::sqrt(s.weights[i]);
        s.loss_null = T(-0.5) * s.y_mean * s.y_mean;                             // state_gaussian_naive.hpp:143-144
        s.loss_full = T(-0.5) * s.y_var + s.loss_null;
        update_screen_derived_range(s, s.X_means.data(), s.weights_sqrt.data(), 0, s.screen_set.size(),
                                    s.screen_X_means, s.screen_transforms, s.screen_vars);
    }
}

} // namespace orc
