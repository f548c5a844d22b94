// ============================================================================
// oracle/cov_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see adelie_oracle.hpp).
//
// CPU restatement of the reference's COVARIANCE-METHOD Gaussian solver:
//   CORE/matrix/matrix_cov_base.hpp:20-153, matrix_cov_dense.ipp:8-84, matrix_cov_lazy_cov.ipp:8-190
//   CORE/solver/solver_gaussian_pin_cov.hpp:56-777   (pin solve on a fixed screen set)
//   CORE/solver/solver_gaussian_cov.hpp:14-461       (path driver: update_screen_derived, fit, solve)
//   CORE/state/state_gaussian_pin_cov.hpp/.ipp, state_gaussian_cov.hpp/.ipp
// minimising 1/2 b^T A b - v^T b + penalty over a lambda path.  No Eigen: plain loops, the symmetric
// eigendecomposition is the oracle's own Jacobi routine.
//
// Pinned by: the reference's own acceptance test of this solver (tests/test_solver.py:983-1026,
// gaussian_cov(A = X^T X / n, v = X^T y / n) == grpnet(X, y, intercept=False) along the same lambdas),
// replayed against the naive oracle path, which itself is pinned by the reference's state.check
// (tests/test_reference_state_check.py) -- tests/test_oracle_cov.py.
// ============================================================================
#pragma once
#include "adelie_oracle.hpp"

namespace orc {

// ---------------------------------------------------------------------------
// MatrixCovBase (matrix_cov_base.hpp:20-63) + argument checks (:66-131)
// ---------------------------------------------------------------------------
template <class T>
struct MatrixCovBase {
    virtual ~MatrixCovBase() {}
    virtual void bmul(const idx_t* subset, idx_t s, const idx_t* indices, const T* values, idx_t k, T* out) = 0;
    virtual void mul(const idx_t* indices, const T* values, idx_t k, T* out) = 0;
    virtual void to_dense(idx_t i, idx_t p, T* out /* p x p column-major */) = 0;
    virtual idx_t cols() const = 0;
    idx_t rows() const { return cols(); }
    static void check_bmul(idx_t s, idx_t i, idx_t v, idx_t o, idx_t r, idx_t c) {
        if ((s < 0 || s > r) || (i < 0 || i > r) || (i != v) || (v < 0 || v > r) || (o != s))
            throw std::runtime_error("bmul() is given inconsistent inputs!");
        (void)c;
    }
    static void check_to_dense(idx_t i, idx_t p, idx_t r, idx_t c) {
        if ((i < 0 || i > r - p) || (r != c)) throw std::runtime_error("to_dense() is given inconsistent inputs!");
    }
};

// MatrixCovDense (matrix_cov_dense.ipp:8-84): mat(i, j) of a (p, p) array, row- or column-major.
template <class T>
struct MatrixCovDense : MatrixCovBase<T> {
    const T* A; idx_t p, ld; bool row_major;
    MatrixCovDense(const T* A_, idx_t p_, idx_t ld_, bool row_major_) : A(A_), p(p_), ld(ld_), row_major(row_major_) {}
    T at(idx_t i, idx_t j) const { return row_major ? A[i * ld + j] : A[i + j * ld]; }
    void bmul(const idx_t* subset, idx_t s, const idx_t* indices, const T* values, idx_t k, T* out) override {   // :25-43
        this->check_bmul(s, k, k, s, p, p);
        for (idx_t j_idx = 0; j_idx < s; ++j_idx) {
            const idx_t j = subset[j_idx];
            T acc = 0;
            for (idx_t i_idx = 0; i_idx < k; ++i_idx) acc += values[i_idx] * at(indices[i_idx], j);
            out[j_idx] = acc;
        }
    }
    void mul(const idx_t* indices, const T* values, idx_t k, T* out) override {                                  // :45-64
        for (idx_t j = 0; j < p; ++j) out[j] = 0;
        for (idx_t i_idx = 0; i_idx < k; ++i_idx) {
            const idx_t i = indices[i_idx]; const T v = values[i_idx];
            // row(i) of a row-major array, col(i) of a column-major one: the contiguous line through i
            for (idx_t j = 0; j < p; ++j) out[j] += v * A[i * ld + j];
        }
    }
    void to_dense(idx_t i, idx_t q, T* out) override {                                                           // :66-75
        this->check_to_dense(i, q, p, p);
        for (idx_t c = 0; c < q; ++c) for (idx_t r = 0; r < q; ++r) out[r + c * q] = at(i + r, i + c);
    }
    idx_t cols() const override { return p; }
};

// MatrixCovLazyCov (matrix_cov_lazy_cov.ipp:8-190): A = X^T X, rows computed on demand and cached.
template <class T>
struct MatrixCovLazyCov : MatrixCovBase<T> {
    const T* X; idx_t n, p, ld; bool row_major;
    std::vector<std::vector<T>> cache_;           // blocks (q x p, row-major)
    std::vector<idx_t> cache_rows_;
    std::vector<idx_t> index_map, slice_map;
    MatrixCovLazyCov(const T* X_, idx_t n_, idx_t p_, idx_t ld_, bool row_major_)
        : X(X_), n(n_), p(p_), ld(ld_), row_major(row_major_), index_map(p_, -1), slice_map(p_, -1) {}
    T x(idx_t i, idx_t j) const { return row_major ? X[i * ld + j] : X[i + j * ld]; }
    T dotcols(idx_t a, idx_t b) const { T s = 0; for (idx_t i = 0; i < n; ++i) s += x(i, a) * x(i, b); return s; }
    void cache(idx_t i, idx_t q) {                                                                               // :10-48
        const idx_t next = (idx_t)cache_.size();
        for (idx_t k = 0; k < q; ++k) { index_map[i + k] = next; slice_map[i + k] = k; }
        std::vector<T> cov((size_t)q * p);
        for (idx_t k = 0; k < q; ++k) for (idx_t j = 0; j < p; ++j) cov[(size_t)k * p + j] = dotcols(i + k, j);
        cache_.emplace_back(std::move(cov)); cache_rows_.push_back(q);
    }
    void bmul(const idx_t* subset, idx_t s, const idx_t* indices, const T* values, idx_t k, T* out) override {   // :66-101
        this->check_bmul(s, k, k, s, p, p);
        for (idx_t i_idx = 0; i_idx < k; ++i_idx) {
            const idx_t i = indices[i_idx];
            if (index_map[i] < 0) {
                idx_t cs = 0;
                for (; i + cs < p && index_map[i + cs] < 0 && i_idx + cs < k && indices[i_idx + cs] == i + cs; ++cs);
                cache(i, cs);
            }
        }
        for (idx_t j_idx = 0; j_idx < s; ++j_idx) {
            const idx_t j = subset[j_idx];
            T acc = 0;
            for (idx_t i_idx = 0; i_idx < k; ++i_idx) {
                const idx_t i = indices[i_idx];
                acc += values[i_idx] * cache_[index_map[i]][(size_t)slice_map[i] * p + j];
            }
            out[j_idx] = acc;
        }
    }
    void mul(const idx_t* indices, const T* values, idx_t k, T* out) override {                                  // :103-152
        for (idx_t j = 0; j < p; ++j) out[j] = 0;
        std::vector<T> Xv(n);
        for (idx_t i_idx = 0; i_idx < k;) {
            const idx_t i = indices[i_idx];
            if (index_map[i] < 0) {
                idx_t bs = 0;
                for (; i + bs < p && index_map[i + bs] < 0 && i_idx + bs < k && indices[i_idx + bs] == i + bs; ++bs);
                for (idx_t r = 0; r < n; ++r) { T s = 0; for (idx_t c = 0; c < bs; ++c) s += x(r, i + c) * values[i_idx + c]; Xv[r] = s; }
                for (idx_t j = 0; j < p; ++j) { T s = 0; for (idx_t r = 0; r < n; ++r) s += x(r, j) * Xv[r]; out[j] += s; }
                i_idx += bs;
                continue;
            }
            const T* row = cache_[index_map[i]].data() + (size_t)slice_map[i] * p;
            const T v = values[i_idx];
            for (idx_t j = 0; j < p; ++j) out[j] += v * row[j];
            ++i_idx;
        }
    }
    void to_dense(idx_t i, idx_t q, T* out) override {                                                           // :154-181
        this->check_to_dense(i, q, p, p);
        for (idx_t c = 0; c < q; ++c) for (idx_t r = 0; r < q; ++r) out[r + c * q] = dotcols(i + r, i + c);
    }
    idx_t cols() const override { return p; }
};

// ---------------------------------------------------------------------------
// StateGaussianPinCov (state_gaussian_pin_cov.hpp:39-140) + pin::cov::solve
// ---------------------------------------------------------------------------
template <class T>
struct CovPinState {
    MatrixCovBase<T>* A;
    const idx_t* groups; const idx_t* group_sizes; idx_t G;
    T alpha; const T* penalty;
    const idx_t* screen_set; const idx_t* screen_begins; idx_t S;
    const T* screen_vars; const std::vector<std::vector<T>>* screen_transforms;     // row-major (r, c) -> [r*gs + c]
    const idx_t* screen_subset_order; const idx_t* screen_subset_ordered; idx_t m;  // m = number of screen values
    std::vector<T> lmda_path;
    size_t max_active_size, max_iters; T tol, rdev_tol, newton_tol; size_t newton_max_iters;
    // dynamic
    T rsq; T* screen_beta; T* screen_grad; int8_t* screen_is_active;
    size_t active_set_size; idx_t* active_set;
    std::vector<idx_t> active_begins, active_order;
    std::vector<int8_t> screen_is_active_subset;
    std::vector<idx_t> active_subset_order, active_subset_ordered, inactive_subset_order, inactive_subset_ordered;
    // outputs
    std::vector<std::vector<idx_t>> beta_idx; std::vector<std::vector<T>> beta_val;
    std::vector<T> intercepts, rsqs, lmdas;
    size_t iters = 0, n_group_updates = 0;
    std::vector<double> benchmark_screen, benchmark_active;
};

// update_active_inactive_subset (solver_gaussian_pin_cov.hpp:56-106)
template <class T>
inline void update_active_inactive_subset(CovPinState<T>& st) {
    st.screen_is_active_subset.assign(st.m, 0);
    idx_t n_processed = 0;
    for (idx_t ss_idx = 0; ss_idx < st.S; ++ss_idx) {
        const idx_t gs = st.group_sizes[st.screen_set[ss_idx]];
        for (idx_t c = 0; c < gs; ++c) st.screen_is_active_subset[n_processed + c] = st.screen_is_active[ss_idx];
        n_processed += gs;
    }
    st.active_subset_order.clear(); st.active_subset_ordered.clear();
    st.inactive_subset_order.clear(); st.inactive_subset_ordered.clear();
    for (idx_t i = 0; i < st.m; ++i) {
        const idx_t ssoi = st.screen_subset_order[i], sso = st.screen_subset_ordered[i];
        if (st.screen_is_active_subset[ssoi]) { st.active_subset_order.push_back(i); st.active_subset_ordered.push_back(sso); }
        else { st.inactive_subset_order.push_back(i); st.inactive_subset_ordered.push_back(sso); }
    }
}

// update_screen_grad_{screen, active, inactive} (:109-205): which = 0 screen, 1 active, 2 inactive
template <class T>
inline void update_screen_grad(CovPinState<T>& st, int which, const idx_t* indices, const T* values, idx_t k, std::vector<T>& buffer_sg) {
    if (which == 0) {
        st.A->bmul(st.screen_subset_ordered, st.m, indices, values, k, buffer_sg.data());
        for (idx_t i = 0; i < st.m; ++i) st.screen_grad[st.screen_subset_order[i]] -= buffer_sg[i];
        return;
    }
    const auto& order = (which == 1) ? st.active_subset_order : st.inactive_subset_order;
    const auto& ordered = (which == 1) ? st.active_subset_ordered : st.inactive_subset_ordered;
    st.A->bmul(ordered.data(), (idx_t)ordered.size(), indices, values, k, buffer_sg.data());
    for (size_t i = 0; i < order.size(); ++i) st.screen_grad[st.screen_subset_order[order[i]]] -= buffer_sg[i];
}

// coordinate_descent (:243-385), constraints == nullptr
template <class T, class Iter, class Extra>
inline void cov_coordinate_descent(CovPinState<T>& st, Iter begin, Iter end, size_t lmda_idx, T& convg_measure, int which,
                                   std::vector<T>& b1, std::vector<T>& b2, std::vector<T>& b3, std::vector<T>& b4,
                                   std::vector<T>& nb1, std::vector<T>& nb2, std::vector<idx_t>& bidx, std::vector<T>& buffer_sg,
                                   Extra additional_step)
{
    const T lmda = st.lmda_path[lmda_idx];
    const T l1 = lmda * st.alpha, l2 = lmda * (1 - st.alpha);
    convg_measure = 0;
    for (auto it = begin; it != end; ++it) {
        const idx_t ss_idx = *it;
        const idx_t k = st.screen_set[ss_idx];
        const idx_t b = st.screen_begins[ss_idx];
        const idx_t gs = st.group_sizes[k];
        ++st.n_group_updates;
        if (gs == 1) {                                                          // :291-322
            T& ak = st.screen_beta[b];
            T gk = st.screen_grad[b];
            const T A_kk = st.screen_vars[b];
            const T pk = st.penalty[k];
            const T ak_old = ak;
            gk += ak_old * A_kk;
            update_coordinate(ak, A_kk, gk, l1 * pk, l2 * pk);
            gk -= ak_old * A_kk;
            if (ak_old == ak) continue;
            const T del = ak - ak_old;
            convg_measure = std::max(A_kk * del * del, convg_measure);
            st.rsq += del * (2 * gk - del * A_kk);
            const idx_t idx1 = st.groups[k];
            update_screen_grad(st, which, &idx1, &del, 1, buffer_sg);
        } else {                                                                // :324-381
            T* ak = st.screen_beta + b;
            const T* gk = st.screen_grad + b;
            const std::vector<T>& Vk = (*st.screen_transforms)[ss_idx];
            const T* A_kk = st.screen_vars + b;
            const T pk = st.penalty[k];
            T* gk_t = b3.data();
            for (idx_t c = 0; c < gs; ++c) { T s = 0; for (idx_t r = 0; r < gs; ++r) s += gk[r] * Vk[r * gs + c]; gk_t[c] = s; }
            T* ak_old = b4.data(); T* ak_old_t = b4.data() + gs; T* ak_t = b4.data() + 2 * gs;
            for (idx_t i = 0; i < gs; ++i) ak_old[i] = ak[i];
            for (idx_t c = 0; c < gs; ++c) { T s = 0; for (idx_t r = 0; r < gs; ++r) s += ak_old[r] * Vk[r * gs + c]; ak_old_t[c] = s; ak_t[c] = s; }
            for (idx_t i = 0; i < gs; ++i) gk_t[i] += A_kk[i] * ak_old_t[i];
            size_t nit;
            newton_prox(A_kk, gk_t, gs, l1 * pk, l2 * pk, st.newton_tol, st.newton_max_iters, false, ak_t, nit, nb1.data(), nb2.data());
            if (nit >= st.newton_max_iters)
                throw solver_error("adelie_core solver: Newton-ABS max iterations reached! Try increasing newton_max_iters.");
            for (idx_t i = 0; i < gs; ++i) gk_t[i] -= A_kk[i] * ak_old_t[i];
            T dn = 0;
            for (idx_t i = 0; i < gs; ++i) { const T d = ak_old_t[i] - ak_t[i]; dn += d * d; }
            if (std::sqrt(dn) <= Configs::dbeta_tol * std::sqrt((double)gs)) continue;   // :357-358
            T* del_t = b1.data();
            T cm = 0, rs = 0;
            for (idx_t i = 0; i < gs; ++i) {
                del_t[i] = ak_t[i] - ak_old_t[i];
                cm += A_kk[i] * del_t[i] * del_t[i];
                rs += del_t[i] * (2 * gk_t[i] - del_t[i] * A_kk[i]);
            }
            convg_measure = std::max(convg_measure, cm / gs);
            st.rsq += rs;
            for (idx_t r = 0; r < gs; ++r) { T s = 0; for (idx_t c = 0; c < gs; ++c) s += ak_t[c] * Vk[r * gs + c]; ak[r] = s; }
            T* del = b2.data();
            for (idx_t i = 0; i < gs; ++i) { del[i] = ak[i] - ak_old[i]; bidx[i] = st.groups[k] + i; }
            update_screen_grad(st, which, bidx.data(), del, gs, buffer_sg);
        }
        additional_step(ss_idx);
    }
}

// pin::cov::solve (:387-777)
template <class T>
inline void cov_pin_solve(CovPinState<T>& st, const std::function<void()>& check_interrupt = [](){}) {
    idx_t max_gs = 1;
    for (idx_t g = 0; g < st.G; ++g) max_gs = std::max(max_gs, st.group_sizes[g]);
    std::vector<T> b1(max_gs), b2(max_gs), b3(max_gs), b4(3 * max_gs), nb1(max_gs), nb2(max_gs), buffer_sg(std::max<idx_t>(st.m, 1));
    std::vector<idx_t> bidx(max_gs);
    // StateGaussianPinBase (state_gaussian_pin_base.ipp:9-36): active_begins / active_order of the incoming active set
    st.active_begins.clear();
    size_t active_beta_size = 0;
    for (size_t i = 0; i < st.active_set_size; ++i) {
        st.active_begins.push_back(active_beta_size);
        active_beta_size += st.group_sizes[st.screen_set[st.active_set[i]]];
    }
    auto sort_order = [&]() {
        std::sort(st.active_order.begin(), st.active_order.end(), [&](idx_t i, idx_t j) {
            return st.groups[st.screen_set[st.active_set[i]]] < st.groups[st.screen_set[st.active_set[j]]];
        });
    };
    st.active_order.resize(st.active_set_size);
    std::iota(st.active_order.begin(), st.active_order.end(), 0);
    sort_order();
    update_active_inactive_subset(st);                                          // state_gaussian_pin_cov.ipp:8-19

    auto add_active = [&](idx_t ss_idx) {                                       // :617-628
        if (!st.screen_is_active[ss_idx]) {
            if (st.active_set_size >= st.max_active_size)
                throw solver_error("adelie_core solver: Maximum number of active groups reached.");
            st.screen_is_active[ss_idx] = 1;
            st.active_set[st.active_set_size] = ss_idx;
            ++st.active_set_size;
        }
    };
    auto noop = [](idx_t) {};
    std::vector<T> ab_diff, ab_diff_ordered; std::vector<idx_t> ab_diff_indices;

    auto solve_active = [&](size_t l) {                                         // :390-527
        size_t abs_ = 0;
        if (st.active_set_size) {
            const size_t last = st.active_set_size - 1;
            abs_ = st.active_begins[last] + st.group_sizes[st.screen_set[st.active_set[last]]];
        }
        ab_diff.resize(abs_); ab_diff_indices.resize(abs_); ab_diff_ordered.resize(abs_);
        for (size_t i = 0; i < st.active_set_size; ++i) {
            const idx_t ss = st.active_set[i], gs = st.group_sizes[st.screen_set[ss]], sb = st.screen_begins[ss];
            for (idx_t c = 0; c < gs; ++c) ab_diff[st.active_begins[i] + c] = st.screen_beta[sb + c];
        }
        while (1) {
            check_interrupt();
            ++st.iters;
            T cm;
            cov_coordinate_descent(st, st.active_set, st.active_set + st.active_set_size, l, cm, 1, b1, b2, b3, b4, nb1, nb2, bidx, buffer_sg, noop);
            if (cm < st.tol) break;
            if (st.iters >= st.max_iters)
                throw solver_error("adelie_core solver: max coordinate descents reached at lambda index: " + std::to_string(l) + ".");
        }
        for (size_t i = 0; i < st.active_set_size; ++i) {
            const idx_t ss = st.active_set[i], gs = st.group_sizes[st.screen_set[ss]], sb = st.screen_begins[ss];
            for (idx_t c = 0; c < gs; ++c) ab_diff[st.active_begins[i] + c] = st.screen_beta[sb + c] - ab_diff[st.active_begins[i] + c];
        }
        if (ab_diff.empty() || st.active_set_size == (size_t)st.S) return;     // :500-503
        // sparsify_active_beta_diff (:207-241): ascending column order
        size_t pos = 0;
        for (size_t i = 0; i < st.active_order.size(); ++i) {
            const idx_t ao = st.active_order[i];
            const idx_t g = st.screen_set[st.active_set[ao]], gs = st.group_sizes[g];
            for (idx_t c = 0; c < gs; ++c) { ab_diff_indices[pos] = st.groups[g] + c; ab_diff_ordered[pos] = ab_diff[st.active_begins[ao] + c]; ++pos; }
        }
        update_screen_grad(st, 2, ab_diff_indices.data(), ab_diff_ordered.data(), (idx_t)pos, buffer_sg);
    };

    for (size_t l = 0; l < st.lmda_path.size(); ++l) {
        double screen_time = 0, active_time = 0;
        while (1) {
            double t0 = now_s();
            solve_active(l);
            active_time += now_s() - t0;
            check_interrupt();
            ++st.iters;
            T cm;
            const size_t old_active = st.active_set_size;
            t0 = now_s();
            {
                std::vector<idx_t> all(st.S);
                std::iota(all.begin(), all.end(), 0);
                cov_coordinate_descent(st, all.data(), all.data() + st.S, l, cm, 0, b1, b2, b3, b4, nb1, nb2, bidx, buffer_sg, add_active);
            }
            screen_time += now_s() - t0;
            if (old_active < st.active_set_size) {                              // :669-695
                for (size_t i = old_active; i < st.active_set_size; ++i) {
                    st.active_begins.push_back(active_beta_size);
                    active_beta_size += st.group_sizes[st.screen_set[st.active_set[i]]];
                }
                st.active_order.resize(st.active_set_size);
                std::iota(st.active_order.begin() + old_active, st.active_order.end(), old_active);
                sort_order();
                update_active_inactive_subset(st);
            }
            if (cm < st.tol) break;
            if (st.iters >= st.max_iters)
                throw solver_error("adelie_core solver: max coordinate descents reached at lambda index: " + std::to_string(l) + ".");
        }
        std::vector<idx_t> bi; std::vector<T> bv;                               // sparsify_active_beta (pin_base.hpp:58-98)
        for (size_t i = 0; i < st.active_order.size(); ++i) {
            const idx_t ss_idx = st.active_set[st.active_order[i]];
            const idx_t g = st.screen_set[ss_idx], gs = st.group_sizes[g];
            for (idx_t c = 0; c < gs; ++c) { bi.push_back(st.groups[g] + c); bv.push_back(st.screen_beta[st.screen_begins[ss_idx] + c]); }
        }
        st.beta_idx.emplace_back(std::move(bi)); st.beta_val.emplace_back(std::move(bv));
        st.intercepts.push_back(0);
        st.rsqs.push_back(st.rsq);
        st.lmdas.push_back(st.lmda_path[l]);
        st.benchmark_screen.push_back(screen_time); st.benchmark_active.push_back(active_time);
        if ((l >= 1) && (st.rsqs[l] - st.rsqs[l - 1] <= st.rdev_tol * st.rsqs[l])) break;     // :724
    }
}

// ---------------------------------------------------------------------------
// StateGaussianCov (state_gaussian_cov.hpp) + gaussian::cov::solve (solver_gaussian_cov.hpp)
// The O(G) screening logic (screen, kkt, update_abs_grad, update_screen_derived_base) is solver_base.hpp's and is
// shared with the naive path state through the common base struct.
// ---------------------------------------------------------------------------
template <class T>
struct CovPathState : PathState<T> {
    MatrixCovBase<T>* A = nullptr;
    const T* v = nullptr;
    T rdev_tol = 1e-4;
    std::vector<T> screen_grad;
    std::vector<idx_t> screen_subset, screen_subset_order, screen_subset_ordered;
};

// update_screen_derived (solver_gaussian_cov.hpp:20-140)
template <class T>
inline void cov_update_screen_derived(CovPathState<T>& s) {
    update_screen_derived_base(s);
    const size_t old_S = s.screen_transforms.size(), new_S = s.screen_set.size();
    const size_t old_vs = s.screen_subset.size();
    const size_t new_vs = new_S ? (s.screen_begins.back() + s.group_sizes[s.screen_set.back()]) : 0;
    s.screen_transforms.resize(new_S);
    s.screen_vars.resize(new_vs, 0);
    s.screen_grad.resize(new_vs, 0);
    for (size_t i = old_S; i < new_S; ++i) {
        const idx_t g = s.groups[s.screen_set[i]], gs = s.group_sizes[s.screen_set[i]], sb = s.screen_begins[i];
        std::vector<T> Agg((size_t)gs * gs);
        s.A->to_dense(g, gs, Agg.data());
        if (gs == 1) { s.screen_transforms[i].assign(1, T(1)); s.screen_vars[sb] = std::max<T>(Agg[0], 0); continue; }
        std::vector<T> D(gs), V((size_t)gs * gs), Vr((size_t)gs * gs);
        jacobi_eigh(Agg.data(), gs, D.data(), V.data());
        for (idx_t r = 0; r < gs; ++r) for (idx_t c = 0; c < gs; ++c) Vr[r * gs + c] = V[r + c * gs];
        s.screen_transforms[i] = std::move(Vr);
        for (idx_t c = 0; c < gs; ++c) s.screen_vars[sb + c] = D[c] * T(D[c] >= 0);
    }
    for (size_t i = 0; i < new_S; ++i) {                                        // :99-109
        const idx_t g = s.groups[s.screen_set[i]], gs = s.group_sizes[s.screen_set[i]], sb = s.screen_begins[i];
        for (idx_t c = 0; c < gs; ++c) s.screen_grad[sb + c] = s.grad[g + c];
    }
    s.screen_subset.resize(new_vs);                                             // :111-122
    size_t pos = old_vs;
    for (size_t i = old_S; i < new_S; ++i) {
        const idx_t g = s.groups[s.screen_set[i]], gs = s.group_sizes[s.screen_set[i]];
        for (idx_t c = 0; c < gs; ++c) s.screen_subset[pos++] = g + c;
    }
    s.screen_subset_order.resize(new_vs);                                       // :124-139
    std::iota(s.screen_subset_order.begin() + old_vs, s.screen_subset_order.end(), (idx_t)old_vs);
    std::sort(s.screen_subset_order.begin(), s.screen_subset_order.end(), [&](idx_t i, idx_t j) { return s.screen_subset[i] < s.screen_subset[j]; });
    s.screen_subset_ordered.resize(new_vs);
    for (size_t i = 0; i < new_vs; ++i) s.screen_subset_ordered[i] = s.screen_subset[s.screen_subset_order[i]];
}

// fit (:234-357)
template <class T>
inline CovPinState<T> cov_fit(CovPathState<T>& s, T lmda, double& screen_time, double& active_time) {
    std::vector<T> grad_prev = s.screen_grad, beta_prev = s.screen_beta;
    std::vector<int8_t> act_prev = s.screen_is_active;
    CovPinState<T> ps;
    ps.A = s.A; ps.groups = s.groups; ps.group_sizes = s.group_sizes; ps.G = s.G; ps.alpha = s.alpha; ps.penalty = s.penalty;
    ps.screen_set = s.screen_set.data(); ps.screen_begins = s.screen_begins.data(); ps.S = (idx_t)s.screen_set.size();
    ps.screen_vars = s.screen_vars.data(); ps.screen_transforms = &s.screen_transforms;
    ps.screen_subset_order = s.screen_subset_order.data(); ps.screen_subset_ordered = s.screen_subset_ordered.data(); ps.m = (idx_t)s.screen_subset.size();
    ps.lmda_path = {lmda};
    ps.max_active_size = s.max_active_size; ps.max_iters = s.max_iters; ps.tol = s.tol; ps.rdev_tol = s.rdev_tol;
    ps.newton_tol = s.newton_tol; ps.newton_max_iters = s.newton_max_iters;
    ps.rsq = s.rsq; ps.screen_beta = s.screen_beta.data(); ps.screen_grad = s.screen_grad.data(); ps.screen_is_active = s.screen_is_active.data();
    ps.active_set_size = s.active_set_size; ps.active_set = s.active_set.data();
    try { cov_pin_solve(ps); }
    catch (...) { s.screen_grad.swap(grad_prev); s.screen_beta.swap(beta_prev); s.screen_is_active.swap(act_prev); throw; }
    s.rsq = ps.rsq; s.active_set_size = ps.active_set_size;
    screen_time = std::accumulate(ps.benchmark_screen.begin(), ps.benchmark_screen.end(), 0.0);
    active_time = std::accumulate(ps.benchmark_active.begin(), ps.benchmark_active.end(), 0.0);
    s.n_sweeps += ps.iters; s.n_group_updates += ps.n_group_updates;
    return ps;
}

// cov::early_exit (:186-203)
template <class T>
inline bool cov_early_exit(const CovPathState<T>& s) {
    if (!s.early_exit || s.devs.size() < 2) return false;
    const T u = s.devs.back(), m = s.devs[s.devs.size() - 2];
    return (u - m <= s.rdev_tol * u);
}

// gaussian::cov::solve (:359-457) = solve_core (solver_base.hpp:435-687) with the covariance-method lambdas
template <class T>
inline void solve_path_cov(CovPathState<T>& s) {
    auto fit = [&](T lmda, double& st, double& at) { return cov_fit(s, lmda, st, at); };
    auto update_invariance = [&](CovPinState<T>& ps, T lmda) {                 // :376-402
        s.lmda = lmda;
        const auto& bi = ps.beta_idx.back(); const auto& bv = ps.beta_val.back();
        s.A->mul(bi.data(), bv.data(), (idx_t)bi.size(), s.grad.data());
        for (idx_t j = 0; j < s.p; ++j) s.grad[j] = s.v[j] - s.grad[j];
        update_abs_grad(static_cast<PathState<T>&>(s), lmda);
    };
    auto update_solutions = [&](CovPinState<T>& ps, T lmda) {                  // :205-232
        s.beta_idx.emplace_back(std::move(ps.beta_idx.back()));
        s.beta_val.emplace_back(std::move(ps.beta_val.back()));
        s.intercepts.push_back(0);
        s.lmdas.push_back(lmda);
        s.devs.push_back(ps.rsqs.back());
    };
    auto screen_f = [&](T lmda, bool kkt_passed, int n_new_active) {
        screen(static_cast<PathState<T>&>(s), lmda, kkt_passed, n_new_active);
        cov_update_screen_derived(s);
    };
    if (s.screen_set.size() > s.max_screen_size) throw solver_error("adelie_core solver: maximum screen set size reached.");
    double st, at;
    if (s.setup_lmda_max) {
        T pmax = s.penalty[0];
        for (idx_t i = 1; i < s.G; ++i) pmax = std::max(pmax, s.penalty[i]);
        const T large_lmda = T(1e-3 * std::numeric_limits<T>::max() / std::max<T>(1, pmax));
        CovPinState<T> ps = fit(large_lmda, st, at);
        update_invariance(ps, large_lmda);
        const T factor = (s.alpha <= 0) ? T(1e-3) : s.alpha;
        T m = -std::numeric_limits<T>::infinity();
        for (idx_t i = 0; i < s.G; ++i) m = std::max<T>(m, (s.penalty[i] <= 0.0) ? T(0.0) : s.abs_grad[i] / s.penalty[i]);
        s.lmda_max = m / factor;
    }
    if (s.setup_lmda_path) {
        if (s.lmda_path_size <= 0) return;
        s.lmda_path.resize(s.lmda_path_size);
        const size_t L = s.lmda_path_size;
        if (L > 1) {
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
            CovPinState<T> ps = fit(large[i], st, at);
            if (i + 1 < large.size()) { update_solutions(ps, large[i]); if (cov_early_exit(s)) return; }
            else update_invariance(ps, large[i]);
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
            CovPinState<T> ps = fit(lmda_curr, st, at);
            s.benchmark_fit_screen.push_back(st); s.benchmark_fit_active.push_back(at);
            t0 = now_s();
            update_invariance(ps, lmda_curr);
            s.benchmark_invariance.push_back(now_s() - t0);
            t0 = now_s();
            kkt_passed = kkt(static_cast<PathState<T>&>(s), lmda_curr);
            s.n_valid_solutions.push_back(kkt_passed);
            idx += kkt_passed;
            if (kkt_passed) update_solutions(ps, lmda_curr);
            s.benchmark_kkt.push_back(now_s() - t0);
            if (kkt_passed) { s.active_sizes.push_back((int)s.active_set_size); s.screen_sizes.push_back((int)s.screen_set.size()); }
            n_new_active = kkt_passed ? (s.active_sizes.back() - current_active) : n_new_active;
            current_active = kkt_passed ? s.active_sizes.back() : current_active;
            if (kkt_passed) break;
        }
        if (cov_early_exit(s)) break;
    }
}

// State construction (state_base.ipp:94-99 + state_gaussian_cov.ipp:9-28: update_screen_derived)
template <class T>
inline void init_cov_path_state(CovPathState<T>& s) {
    s.abs_grad.assign(s.G, 0);
    update_screen_derived_base(s);        // (idempotent: cov_update_screen_derived calls it again)
    update_abs_grad(static_cast<PathState<T>&>(s), s.lmda);
    cov_update_screen_derived(s);
}

} // namespace orc
