// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE (see adelie_oracle.hpp header).
// C entry points of the CPU oracle; built by oracle/Makefile into
// oracle/liboracle.so and loaded by oracle/oracle.py with ctypes.
#include "oracle_capi.h"
#include "adelie_oracle.hpp"
#include "cox_oracle.hpp"
#include "cov_oracle.hpp"
#include <map>
#include <memory>

using namespace orc;

namespace {

struct Result {
    std::string error;
    std::map<std::string, double> scalars;
    std::map<std::string, std::vector<double>> vecs;
    std::map<std::string, std::vector<int64_t>> ivecs;
    std::vector<int64_t> indptr, indices; std::vector<double> values;
};

template <class T, class V>
std::vector<double> to_d(const V& v) { return std::vector<double>(v.begin(), v.end()); }

template <class T>
std::unique_ptr<GlmBase<T>> make_glm(int family, idx_t n, idx_t K, const void* y, const void* w,
                                     const void* cs, const void* ce, const void* cst, const int64_t* strata, int efron)
{
    switch (family) {
        case ORC_FAM_GAUSSIAN: return std::make_unique<GlmGaussian<T>>((const T*)y, (const T*)w, n);
        case ORC_FAM_BINOMIAL: return std::make_unique<GlmBinomialLogit<T>>((const T*)y, (const T*)w, n);
        case ORC_FAM_MULTIGAUSSIAN: return std::make_unique<GlmMultiGaussian<T>>((const T*)y, (const T*)w, n, K);
        case ORC_FAM_BINOMIAL_PROBIT: return std::make_unique<GlmBinomialProbit<T>>((const T*)y, (const T*)w, n);
        case ORC_FAM_MULTINOMIAL: return std::make_unique<GlmMultinomial<T>>((const T*)y, (const T*)w, n, K);
        case ORC_FAM_POISSON: return std::make_unique<GlmPoisson<T>>((const T*)y, (const T*)w, n);
        case ORC_FAM_COX: return std::make_unique<GlmCox<T>>((const T*)cs, (const T*)ce, (const T*)cst, strata, (const T*)w, n, efron != 0);
    }
    throw std::runtime_error("unknown family");
}

// multi-response update_loss_null (solver_multiglm_naive.hpp:99-186)
template <class T>
void update_loss_null_multi(PathState<T>& s, GlmBuffers<T>& B, idx_t K, bool multi_intercept) {
    auto& glm = *s.glm;
    const idx_t nK = s.n, n = nK / K;
    if (!multi_intercept) { s.loss_null = glm.loss(s.offsets); return; }
    std::vector<T> beta0(K), eta = s.eta, resid = s.resid;
    size_t it = 0;
    while (1) {
        if (it >= s.irls_max_iters) throw solver_error("adelie_core solver: Maximum IRLS iterations reached.");
        glm.hessian(eta.data(), resid.data(), B.hess.data());
        glm.inv_hessian_gradient(eta.data(), resid.data(), B.hess.data(), B.irls_y.data());
        T hs = 0;
        for (idx_t i = 0; i < nK; ++i) {
            B.hess[i] = std::max<T>(B.hess[i], 0) + T(Configs::hessian_min) * T(B.hess[i] <= 0);
            hs += B.hess[i];
        }
        for (idx_t i = 0; i < nK; ++i) { B.irls_weights[i] = B.hess[i] / hs; B.irls_y[i] += eta[i] - s.offsets[i]; }
        for (idx_t k = 0; k < K; ++k) {
            T num = 0, den = 0;
            for (idx_t i = 0; i < n; ++i) { num += B.irls_weights[i * K + k] * B.irls_y[i * K + k]; den += B.irls_weights[i * K + k]; }
            beta0[k] = num / den;
        }
        eta.swap(B.eta_prev);
        for (idx_t i = 0; i < n; ++i) for (idx_t k = 0; k < K; ++k) eta[i * K + k] = s.offsets[i * K + k] + beta0[k];
        B.resid_prev.swap(resid);
        glm.gradient(eta.data(), resid.data());
        T conv = 0;
        for (idx_t i = 0; i < nK; ++i) conv += (resid[i] - B.resid_prev[i]) * (eta[i] - B.eta_prev[i]);
        if (std::abs(conv) <= s.irls_tol) { s.loss_null = glm.loss(eta.data()); return; }
        ++it;
    }
}

// GlmWrap (solver_multiglm_naive.hpp:19-97): flatten (n,K) <-> (nK,)
template <class T>
struct GlmFlat : GlmBase<T> {
    GlmBase<T>* g; idx_t K;
    GlmFlat(GlmBase<T>* g_, idx_t K_) : g(g_), K(K_) { this->name = g_->name; this->y = g_->y; this->w = g_->w; this->n = g_->n * K_; }
    void gradient(const T* eta, T* grad) override { g->gradient(eta, grad); }
    void hessian(const T* eta, const T* grad, T* hess) override { g->hessian(eta, grad, hess); }
    void inv_hessian_gradient(const T* eta, const T* grad, const T* hess, T* out) override { g->inv_hessian_gradient(eta, grad, hess, out); }
    T loss(const T* eta) override { return g->loss(eta); }
    T loss_full() override { return g->loss_full(); }
    void inv_link(const T* eta, T* out) override { g->inv_link(eta, out); }
};

template <class T>
void run_path(const orc_path_args* a, Result& R) {
    const idx_t K = std::max<int64_t>(a->K, 1);
    const bool multi = K > 1;
    std::unique_ptr<MatrixBase<T>> base;
    if (a->matrix_kind == ORC_MAT_DENSE) base = std::make_unique<MatrixDense<T>>((const T*)a->X, a->n, a->p, a->ld, a->n_threads);
    else base = std::make_unique<MatrixSparse<T>>(a->n, a->p, a->sp_outer, a->sp_inner, (const T*)a->sp_values, a->n_threads);
    std::unique_ptr<MatrixBase<T>> kron, ones_dense, ones_kron, cat;
    std::vector<T> ones_col;
    MatrixBase<T>* X = base.get();
    if (multi) {
        kron = std::make_unique<MatrixKroneckerEye<T>>(base.get(), K);
        X = kron.get();
        if (a->multi_intercept) {
            ones_col.assign(a->n, T(1));
            ones_dense = std::make_unique<MatrixDense<T>>(ones_col.data(), a->n, 1, a->n, 1);
            ones_kron = std::make_unique<MatrixKroneckerEye<T>>(ones_dense.get(), K);
            cat = std::make_unique<MatrixCConcatenate<T>>(std::vector<MatrixBase<T>*>{ones_kron.get(), kron.get()});
            X = cat.get();
        }
    }
    const bool is_glm = a->family != ORC_FAM_GAUSSIAN_OPT;
    PathState<T> s;
    s.X = X; s.n = X->rows(); s.p = X->cols(); s.G = a->G;
    s.groups = a->groups; s.group_sizes = a->group_sizes; s.alpha = (T)a->alpha; s.penalty = (const T*)a->penalty;
    std::vector<T> w_expanded;
    std::unique_ptr<GlmBase<T>> glm; std::unique_ptr<GlmFlat<T>> glm_flat;
    if (!is_glm) {
        if (multi) {                                                            // PY/state.py:2315 repeat(w, K)/K
            w_expanded.resize(s.n);
            for (idx_t i = 0; i < a->n; ++i) for (idx_t k = 0; k < K; ++k) w_expanded[i * K + k] = ((const T*)a->weights)[i] / K;
            s.weights = w_expanded.data();
        } else s.weights = (const T*)a->weights;
        s.X_means.assign((const T*)a->X_means, (const T*)a->X_means + s.p);
        s.y_mean = (T)a->y_mean; s.y_var = (T)a->y_var; s.rsq = (T)a->rsq; s.resid_sum = (T)a->resid_sum;
    } else {
        glm = make_glm<T>(a->family, a->n, K, a->y, a->weights, a->cox_start, a->cox_stop, a->cox_status, a->cox_strata, a->cox_efron);
        s.glm = glm.get();
        if (multi) { glm_flat = std::make_unique<GlmFlat<T>>(glm.get(), K); s.glm = glm_flat.get(); }
        s.offsets = (const T*)a->offsets;
        s.eta.assign((const T*)a->eta, (const T*)a->eta + s.n);
        s.beta0 = (T)a->beta0; s.loss_null = (T)a->loss_null; s.loss_full = (T)a->loss_full;
        s.setup_loss_null = a->setup_loss_null;
        s.irls_max_iters = a->irls_max_iters; s.irls_tol = (T)a->irls_tol;
    }
    s.resid.assign((const T*)a->resid, (const T*)a->resid + s.n);
    s.grad.assign((const T*)a->grad, (const T*)a->grad + s.p);
    s.min_ratio = (T)a->min_ratio; s.lmda_path_size = a->lmda_path_size;
    s.max_screen_size = a->max_screen_size; s.max_active_size = a->max_active_size;
    s.pivot_subset_ratio = (T)a->pivot_subset_ratio; s.pivot_subset_min = a->pivot_subset_min; s.pivot_slack_ratio = (T)a->pivot_slack_ratio;
    s.screen_rule = a->screen_rule; s.max_iters = a->max_iters; s.tol = (T)a->tol; s.adev_tol = (T)a->adev_tol; s.ddev_tol = (T)a->ddev_tol;
    s.newton_tol = (T)a->newton_tol; s.newton_max_iters = a->newton_max_iters;
    s.early_exit = a->early_exit; s.setup_lmda_max = a->setup_lmda_max; s.setup_lmda_path = a->setup_lmda_path;
    s.intercept = multi ? false : (a->intercept != 0);                          // PY/state.py:2329-2330
    s.n_threads = a->n_threads;
    s.lmda_max = (T)a->lmda_max; s.lmda = (T)a->lmda;
    if (a->lmda_path && a->lmda_path_len > 0) s.lmda_path.assign((const T*)a->lmda_path, (const T*)a->lmda_path + a->lmda_path_len);
    s.screen_set.assign(a->screen_set, a->screen_set + a->S);
    s.screen_beta.assign((const T*)a->screen_beta, (const T*)a->screen_beta + a->screen_beta_size);
    s.screen_is_active.assign(a->screen_is_active, a->screen_is_active + a->S);
    s.active_set_size = a->active_set_size;
    s.active_set.assign(a->active_set, a->active_set + a->G);
    s.max_seconds = a->max_seconds;

    const double t0 = now_s();
    try {
        init_path_state(s, is_glm);
        if (is_glm && multi) {
            // multiglm::naive::solve swaps in the multi update_loss_null (solver_multiglm_naive.hpp:236-240)
            if (s.setup_loss_null) {
                GlmBuffers<T> B(s.n, s.p);
                update_loss_null_multi(s, B, K, a->multi_intercept != 0 && a->intercept != 0);
                s.setup_loss_null = false;
            }
        }
        solve_path(s, is_glm);
    } catch (const std::exception& e) {
        R.error = e.what();
    }
    R.scalars["total_time"] = now_s() - t0;
    R.scalars["lmda_max"] = s.lmda_max; R.scalars["lmda"] = s.lmda; R.scalars["rsq"] = s.rsq; R.scalars["resid_sum"] = s.resid_sum;
    R.scalars["y_mean"] = s.y_mean; R.scalars["y_var"] = s.y_var; R.scalars["loss_null"] = s.loss_null; R.scalars["loss_full"] = s.loss_full;
    R.scalars["beta0"] = s.beta0; R.scalars["active_set_size"] = (double)s.active_set_size;
    R.scalars["n_sweeps"] = (double)s.n_sweeps; R.scalars["n_group_updates"] = (double)s.n_group_updates; R.scalars["n_irls"] = (double)s.n_irls;
    R.vecs["lmda_path"] = to_d<T>(s.lmda_path); R.vecs["lmdas"] = to_d<T>(s.lmdas); R.vecs["devs"] = to_d<T>(s.devs);
    R.vecs["intercepts"] = to_d<T>(s.intercepts); R.vecs["screen_beta"] = to_d<T>(s.screen_beta);
    R.vecs["grad"] = to_d<T>(s.grad); R.vecs["abs_grad"] = to_d<T>(s.abs_grad); R.vecs["resid"] = to_d<T>(s.resid);
    R.vecs["eta"] = to_d<T>(s.eta); R.vecs["screen_vars"] = to_d<T>(s.screen_vars); R.vecs["screen_X_means"] = to_d<T>(s.screen_X_means);
    { std::vector<double> flat; for (const auto& V : s.screen_transforms) flat.insert(flat.end(), V.begin(), V.end()); R.vecs["screen_transforms_flat"] = flat; }
    R.vecs["benchmark_screen"] = s.benchmark_screen; R.vecs["benchmark_fit_screen"] = s.benchmark_fit_screen;
    R.vecs["benchmark_fit_active"] = s.benchmark_fit_active; R.vecs["benchmark_kkt"] = s.benchmark_kkt;
    R.vecs["benchmark_invariance"] = s.benchmark_invariance;
    R.ivecs["screen_set"] = std::vector<int64_t>(s.screen_set.begin(), s.screen_set.end());
    R.ivecs["screen_begins"] = std::vector<int64_t>(s.screen_begins.begin(), s.screen_begins.end());
    R.ivecs["screen_is_active"] = std::vector<int64_t>(s.screen_is_active.begin(), s.screen_is_active.end());
    R.ivecs["active_set"] = std::vector<int64_t>(s.active_set.begin(), s.active_set.begin() + s.active_set_size);
    R.ivecs["n_valid_solutions"] = std::vector<int64_t>(s.n_valid_solutions.begin(), s.n_valid_solutions.end());
    R.ivecs["active_sizes"] = std::vector<int64_t>(s.active_sizes.begin(), s.active_sizes.end());
    R.ivecs["screen_sizes"] = std::vector<int64_t>(s.screen_sizes.begin(), s.screen_sizes.end());
    R.indptr.push_back(0);
    for (size_t l = 0; l < s.beta_idx.size(); ++l) {
        for (size_t k = 0; k < s.beta_idx[l].size(); ++k) { R.indices.push_back(s.beta_idx[l][k]); R.values.push_back((double)s.beta_val[l][k]); }
        R.indptr.push_back((int64_t)R.indices.size());
    }
}

template <class T>
void run_pin(orc_pin_args* a, Result& R) {
    MatrixDense<T> X((const T*)a->X, a->n, a->p, a->ld, a->n_threads);
    PathState<T> s;
    s.X = &X; s.n = a->n; s.p = a->p; s.G = a->G; s.groups = a->groups; s.group_sizes = a->group_sizes;
    s.alpha = (T)a->alpha; s.penalty = (const T*)a->penalty; s.weights = (const T*)a->weights;
    s.intercept = a->intercept; s.n_threads = a->n_threads;
    s.screen_set.assign(a->screen_set, a->screen_set + a->S);
    update_screen_derived_base(s);
    // screen-derived quantities as PY/state.py:644-660 (gaussian_pin_naive wrapper)
    std::vector<T> X_means(a->p), ones(a->n, T(1)), wsqrt(a->n);
    X.mul(ones.data(), s.weights, X_means.data());
    for (idx_t i = 0; i < a->n; ++i) wsqrt[i] = std::sqrt(s.weights[i]);
    update_screen_derived_range(s, X_means.data(), wsqrt.data(), 0, s.screen_set.size(), s.screen_X_means, s.screen_transforms, s.screen_vars);
    PinState<T> ps;
    ps.X = &X; ps.y_mean = (T)a->y_mean; ps.y_var = (T)a->y_var; ps.groups = a->groups; ps.group_sizes = a->group_sizes; ps.G = a->G;
    ps.alpha = (T)a->alpha; ps.penalty = (const T*)a->penalty; ps.weights = (const T*)a->weights;
    ps.screen_set = s.screen_set.data(); ps.screen_begins = s.screen_begins.data(); ps.S = a->S;
    ps.screen_vars = s.screen_vars.data(); ps.screen_X_means = s.screen_X_means.data(); ps.screen_transforms = &s.screen_transforms;
    ps.lmda_path.assign((const T*)a->lmda_path, (const T*)a->lmda_path + a->L);
    ps.intercept = a->intercept; ps.max_active_size = a->max_active_size; ps.max_iters = a->max_iters;
    ps.tol = (T)a->tol; ps.adev_tol = (T)a->adev_tol; ps.ddev_tol = (T)a->ddev_tol; ps.newton_tol = (T)a->newton_tol; ps.newton_max_iters = a->newton_max_iters;
    ps.rsq = (T)a->rsq; ps.resid = (T*)a->resid; ps.resid_sum = (T)a->resid_sum;
    ps.screen_beta = (T*)a->screen_beta; ps.screen_is_active = a->screen_is_active;
    ps.active_set_size = a->active_set_size; ps.active_set = a->active_set;
    try { pin_solve(ps); } catch (const std::exception& e) { R.error = e.what(); }
    a->active_set_size = ps.active_set_size; a->rsq = ps.rsq; a->resid_sum = ps.resid_sum;
    R.scalars["iters"] = (double)ps.iters; R.scalars["rsq"] = ps.rsq; R.scalars["resid_sum"] = ps.resid_sum;
    R.scalars["active_set_size"] = (double)ps.active_set_size; R.scalars["n_group_updates"] = (double)ps.n_group_updates;
    R.vecs["rsqs"] = to_d<T>(ps.rsqs); R.vecs["lmdas"] = to_d<T>(ps.lmdas); R.vecs["intercepts"] = to_d<T>(ps.intercepts);
    R.vecs["screen_grad"] = to_d<T>(ps.screen_grad); R.vecs["screen_vars"] = to_d<T>(s.screen_vars); R.vecs["screen_X_means"] = to_d<T>(s.screen_X_means);
    { std::vector<double> flat; for (const auto& V : s.screen_transforms) flat.insert(flat.end(), V.begin(), V.end()); R.vecs["screen_transforms_flat"] = flat; }
    R.vecs["benchmark_screen"] = ps.benchmark_screen; R.vecs["benchmark_active"] = ps.benchmark_active;
    R.ivecs["screen_begins"] = std::vector<int64_t>(s.screen_begins.begin(), s.screen_begins.end());
    R.indptr.push_back(0);
    for (size_t l = 0; l < ps.beta_idx.size(); ++l) {
        for (size_t k = 0; k < ps.beta_idx[l].size(); ++k) { R.indices.push_back(ps.beta_idx[l][k]); R.values.push_back((double)ps.beta_val[l][k]); }
        R.indptr.push_back((int64_t)R.indices.size());
    }
}

template <class T>
double glm_eval(int family, int op, idx_t n, idx_t K, const void* y, const void* w, const void* cs, const void* ce,
                const void* cst, const int64_t* strata, int efron, const void* eta, const void* grad, const void* hess, void* out)
{
    auto g = make_glm<T>(family, n, K, y, w, cs, ce, cst, strata, efron);
    switch (op) {
        case 0: g->gradient((const T*)eta, (T*)out); return 0;
        case 1: g->hessian((const T*)eta, (const T*)grad, (T*)out); return 0;
        case 2: g->inv_hessian_gradient((const T*)eta, (const T*)grad, (const T*)hess, (T*)out); return 0;
        case 3: return (double)g->loss((const T*)eta);
        case 4: return (double)g->loss_full();
        case 5: g->inv_link((const T*)eta, (T*)out); return 0;
    }
    return 0;
}


// ---- covariance method -----------------------------------------------------------------------
template <class T>
std::unique_ptr<MatrixCovBase<T>> make_cov_matrix(int kind, const void* M, idx_t n, idx_t p, idx_t ld, bool row_major) {
    if (kind == ORC_COV_DENSE) return std::make_unique<MatrixCovDense<T>>((const T*)M, p, ld, row_major);
    return std::make_unique<MatrixCovLazyCov<T>>((const T*)M, n, p, ld, row_major);
}

template <class T>
void fill_cov_state(const orc_cov_args* a, CovPathState<T>& s, MatrixCovBase<T>* A) {
    s.A = A; s.v = (const T*)a->v; s.p = a->p; s.n = 0; s.G = a->G;
    s.groups = a->groups; s.group_sizes = a->group_sizes; s.alpha = (T)a->alpha; s.penalty = (const T*)a->penalty;
    s.rsq = (T)a->rsq; s.lmda = (T)a->lmda; s.lmda_max = (T)a->lmda_max;
    s.min_ratio = (T)a->min_ratio; s.lmda_path_size = a->lmda_path_size;
    s.max_screen_size = a->max_screen_size; s.max_active_size = a->max_active_size;
    s.pivot_subset_ratio = (T)a->pivot_subset_ratio; s.pivot_subset_min = a->pivot_subset_min; s.pivot_slack_ratio = (T)a->pivot_slack_ratio;
    s.screen_rule = a->screen_rule; s.max_iters = a->max_iters; s.tol = (T)a->tol; s.rdev_tol = (T)a->rdev_tol;
    s.newton_tol = (T)a->newton_tol; s.newton_max_iters = a->newton_max_iters;
    s.early_exit = a->early_exit; s.setup_lmda_max = a->setup_lmda_max; s.setup_lmda_path = a->setup_lmda_path;
    s.intercept = false;
    if (a->lmda_path && a->lmda_path_len > 0) s.lmda_path.assign((const T*)a->lmda_path, (const T*)a->lmda_path + a->lmda_path_len);
    s.screen_set.assign(a->screen_set, a->screen_set + a->S);
    s.screen_beta.assign((const T*)a->screen_beta, (const T*)a->screen_beta + a->screen_beta_size);
    s.screen_is_active.assign(a->screen_is_active, a->screen_is_active + a->S);
    s.active_set_size = a->active_set_size;
    s.active_set.assign(a->active_set, a->active_set + a->G);
    if (a->grad) s.grad.assign((const T*)a->grad, (const T*)a->grad + a->p); else s.grad.assign(a->p, T(0));
}

template <class T>
void collect_cov_common(const CovPathState<T>& s, Result& R) {
    R.scalars["rsq"] = s.rsq; R.scalars["active_set_size"] = (double)s.active_set_size;
    R.vecs["screen_beta"] = to_d<T>(s.screen_beta); R.vecs["screen_grad"] = to_d<T>(s.screen_grad); R.vecs["screen_vars"] = to_d<T>(s.screen_vars);
    { std::vector<double> flat; for (const auto& V : s.screen_transforms) flat.insert(flat.end(), V.begin(), V.end()); R.vecs["screen_transforms_flat"] = flat; }
    R.ivecs["screen_set"] = std::vector<int64_t>(s.screen_set.begin(), s.screen_set.end());
    R.ivecs["screen_begins"] = std::vector<int64_t>(s.screen_begins.begin(), s.screen_begins.end());
    R.ivecs["screen_is_active"] = std::vector<int64_t>(s.screen_is_active.begin(), s.screen_is_active.end());
    R.ivecs["active_set"] = std::vector<int64_t>(s.active_set.begin(), s.active_set.begin() + s.active_set_size);
}

template <class T>
void run_cov_path(const orc_cov_args* a, Result& R) {
    auto A = make_cov_matrix<T>(a->matrix_kind, a->M, a->n, a->p, a->ld, a->row_major != 0);
    CovPathState<T> s;
    fill_cov_state(a, s, A.get());
    const double t0 = now_s();
    try { init_cov_path_state(s); solve_path_cov(s); }
    catch (const std::exception& e) { R.error = e.what(); }
    R.scalars["total_time"] = now_s() - t0;
    R.scalars["lmda_max"] = s.lmda_max; R.scalars["lmda"] = s.lmda;
    R.scalars["n_sweeps"] = (double)s.n_sweeps; R.scalars["n_group_updates"] = (double)s.n_group_updates;
    collect_cov_common(s, R);
    R.vecs["lmda_path"] = to_d<T>(s.lmda_path); R.vecs["lmdas"] = to_d<T>(s.lmdas); R.vecs["devs"] = to_d<T>(s.devs);
    R.vecs["intercepts"] = to_d<T>(s.intercepts); R.vecs["grad"] = to_d<T>(s.grad); R.vecs["abs_grad"] = to_d<T>(s.abs_grad);
    R.vecs["benchmark_screen"] = s.benchmark_screen; R.vecs["benchmark_fit_screen"] = s.benchmark_fit_screen;
    R.vecs["benchmark_fit_active"] = s.benchmark_fit_active; R.vecs["benchmark_kkt"] = s.benchmark_kkt;
    R.vecs["benchmark_invariance"] = s.benchmark_invariance;
    R.ivecs["n_valid_solutions"] = std::vector<int64_t>(s.n_valid_solutions.begin(), s.n_valid_solutions.end());
    R.ivecs["active_sizes"] = std::vector<int64_t>(s.active_sizes.begin(), s.active_sizes.end());
    R.ivecs["screen_sizes"] = std::vector<int64_t>(s.screen_sizes.begin(), s.screen_sizes.end());
    R.indptr.push_back(0);
    for (size_t l = 0; l < s.beta_idx.size(); ++l) {
        for (size_t k = 0; k < s.beta_idx[l].size(); ++k) { R.indices.push_back(s.beta_idx[l][k]); R.values.push_back((double)s.beta_val[l][k]); }
        R.indptr.push_back((int64_t)R.indices.size());
    }
}

// gaussian_pin_cov wrapper (PY/state.py:739-1000: screen_vars / screen_transforms / screen_subset_order derived from A) + pin::cov::solve
template <class T>
void run_cov_pin(const orc_cov_args* a, Result& R) {
    auto A = make_cov_matrix<T>(a->matrix_kind, a->M, a->n, a->p, a->ld, a->row_major != 0);
    CovPathState<T> s;
    fill_cov_state(a, s, A.get());
    CovPinState<T> ps;
    const double t0 = now_s();
    try {
        cov_update_screen_derived(s);
        s.screen_grad.assign((const T*)a->screen_grad, (const T*)a->screen_grad + s.screen_subset.size());
        ps.A = s.A; ps.groups = s.groups; ps.group_sizes = s.group_sizes; ps.G = s.G; ps.alpha = s.alpha; ps.penalty = s.penalty;
        ps.screen_set = s.screen_set.data(); ps.screen_begins = s.screen_begins.data(); ps.S = (idx_t)s.screen_set.size();
        ps.screen_vars = s.screen_vars.data(); ps.screen_transforms = &s.screen_transforms;
        ps.screen_subset_order = s.screen_subset_order.data(); ps.screen_subset_ordered = s.screen_subset_ordered.data(); ps.m = (idx_t)s.screen_subset.size();
        ps.lmda_path = s.lmda_path;
        ps.max_active_size = s.max_active_size; ps.max_iters = s.max_iters; ps.tol = s.tol; ps.rdev_tol = s.rdev_tol;
        ps.newton_tol = s.newton_tol; ps.newton_max_iters = s.newton_max_iters;
        ps.rsq = s.rsq; ps.screen_beta = s.screen_beta.data(); ps.screen_grad = s.screen_grad.data(); ps.screen_is_active = s.screen_is_active.data();
        ps.active_set_size = s.active_set_size; ps.active_set = s.active_set.data();
        cov_pin_solve(ps);
    } catch (const std::exception& e) { R.error = e.what(); }
    s.rsq = ps.rsq; s.active_set_size = ps.active_set_size;
    R.scalars["total_time"] = now_s() - t0;
    R.scalars["iters"] = (double)ps.iters; R.scalars["n_group_updates"] = (double)ps.n_group_updates;
    collect_cov_common(s, R);
    R.vecs["rsqs"] = to_d<T>(ps.rsqs); R.vecs["lmdas"] = to_d<T>(ps.lmdas); R.vecs["intercepts"] = to_d<T>(ps.intercepts);
    R.vecs["benchmark_screen"] = ps.benchmark_screen; R.vecs["benchmark_active"] = ps.benchmark_active;
    R.indptr.push_back(0);
    for (size_t l = 0; l < ps.beta_idx.size(); ++l) {
        for (size_t k = 0; k < ps.beta_idx[l].size(); ++k) { R.indices.push_back(ps.beta_idx[l][k]); R.values.push_back((double)ps.beta_val[l][k]); }
        R.indptr.push_back((int64_t)R.indices.size());
    }
}

template <class T>
void cov_matrix_op(int kind, const void* M, idx_t n, idx_t p, idx_t ld, bool row_major, int op, const int64_t* subset, idx_t s,
                   const int64_t* indices, const void* values, idx_t k, idx_t i0, idx_t q, void* out) {
    auto A = make_cov_matrix<T>(kind, M, n, p, ld, row_major);
    if (op == 0) A->bmul(subset, s, indices, (const T*)values, k, (T*)out);
    else if (op == 1) A->mul(indices, (const T*)values, k, (T*)out);
    else A->to_dense(i0, q, (T*)out);
}

} // namespace

extern "C" {

void* orc_path_solve(const orc_path_args* a) {
    auto* R = new Result();
    try {
        if (a->dtype == ORC_F32) run_path<float>(a, *R); else run_path<double>(a, *R);
    } catch (const std::exception& e) { R->error = e.what(); }
    return R;
}
void* orc_pin_solve(orc_pin_args* a) {
    auto* R = new Result();
    try {
        if (a->dtype == ORC_F32) run_pin<float>(a, *R); else run_pin<double>(a, *R);
    } catch (const std::exception& e) { R->error = e.what(); }
    return R;
}

void* orc_cov_path_solve(const orc_cov_args* a) {
    auto* R = new Result();
    try { if (a->dtype == ORC_F32) run_cov_path<float>(a, *R); else run_cov_path<double>(a, *R); }
    catch (const std::exception& e) { R->error = e.what(); }
    return R;
}
void* orc_cov_pin_solve(const orc_cov_args* a) {
    auto* R = new Result();
    try { if (a->dtype == ORC_F32) run_cov_pin<float>(a, *R); else run_cov_pin<double>(a, *R); }
    catch (const std::exception& e) { R->error = e.what(); }
    return R;
}
void orc_cov_matrix_op(int dtype, int kind, const void* M, int64_t n, int64_t p, int64_t ld, int row_major, int op,
                       const int64_t* subset, int64_t s, const int64_t* indices, const void* values, int64_t k,
                       int64_t i0, int64_t q, void* out) {
    if (dtype == ORC_F32) cov_matrix_op<float>(kind, M, n, p, ld, row_major != 0, op, subset, s, indices, values, k, i0, q, out);
    else cov_matrix_op<double>(kind, M, n, p, ld, row_major != 0, op, subset, s, indices, values, k, i0, q, out);
}
void orc_result_free(void* h) { delete (Result*)h; }
const char* orc_result_error(void* h) { return ((Result*)h)->error.c_str(); }
double orc_result_scalar(void* h, const char* name) {
    auto& m = ((Result*)h)->scalars; auto it = m.find(name);
    return it == m.end() ? std::numeric_limits<double>::quiet_NaN() : it->second;
}
int64_t orc_result_vec(void* h, const char* name, double* out, int64_t cap) {
    auto& m = ((Result*)h)->vecs; auto it = m.find(name);
    if (it == m.end()) return -1;
    if (out) for (int64_t i = 0; i < std::min<int64_t>(cap, it->second.size()); ++i) out[i] = it->second[i];
    return (int64_t)it->second.size();
}
int64_t orc_result_ivec(void* h, const char* name, int64_t* out, int64_t cap) {
    auto& m = ((Result*)h)->ivecs; auto it = m.find(name);
    if (it == m.end()) return -1;
    if (out) for (int64_t i = 0; i < std::min<int64_t>(cap, it->second.size()); ++i) out[i] = it->second[i];
    return (int64_t)it->second.size();
}
int64_t orc_result_betas(void* h, int64_t* indptr, int64_t* indices, double* values) {
    auto* R = (Result*)h;
    if (indptr) std::copy(R->indptr.begin(), R->indptr.end(), indptr);
    if (indices) std::copy(R->indices.begin(), R->indices.end(), indices);
    if (values) std::copy(R->values.begin(), R->values.end(), values);
    return (int64_t)R->indices.size();
}

#define DENSE(T) MatrixDense<T> M((const T*)X, n, p, ld, n_threads)
double orc_dense_cmul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, const void* v, const void* w, int n_threads) {
    if (dtype == ORC_F32) { DENSE(float); return M.cmul(j, (const float*)v, (const float*)w); }
    DENSE(double); return M.cmul(j, (const double*)v, (const double*)w);
}
void orc_dense_ctmul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, double v, void* out, int n_threads) {
    if (dtype == ORC_F32) { DENSE(float); M.ctmul(j, (float)v, (float*)out); return; }
    DENSE(double); M.ctmul(j, v, (double*)out);
}
void orc_dense_bmul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, int64_t q, const void* v, const void* w, void* out, int n_threads) {
    if (dtype == ORC_F32) { DENSE(float); M.bmul(j, q, (const float*)v, (const float*)w, (float*)out); return; }
    DENSE(double); M.bmul(j, q, (const double*)v, (const double*)w, (double*)out);
}
void orc_dense_btmul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, int64_t q, const void* v, void* out, int n_threads) {
    if (dtype == ORC_F32) { DENSE(float); M.btmul(j, q, (const float*)v, (float*)out); return; }
    DENSE(double); M.btmul(j, q, (const double*)v, (double*)out);
}
void orc_dense_mul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, const void* v, const void* w, void* out, int n_threads) {
    if (dtype == ORC_F32) { DENSE(float); M.mul((const float*)v, (const float*)w, (float*)out); return; }
    DENSE(double); M.mul((const double*)v, (const double*)w, (double*)out);
}
void orc_dense_cov(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, int64_t q, const void* sqrt_w, void* out, int n_threads) {
    if (dtype == ORC_F32) { DENSE(float); M.cov(j, q, (const float*)sqrt_w, (float*)out); return; }
    DENSE(double); M.cov(j, q, (const double*)sqrt_w, (double*)out);
}

double orc_glm_eval(int dtype, int family, int op, int64_t n, int64_t K, const void* y, const void* w,
                    const void* cs, const void* ce, const void* cst, const int64_t* strata, int efron,
                    const void* eta, const void* grad, const void* hess, void* out) {
    if (dtype == ORC_F32) return glm_eval<float>(family, op, n, K, y, w, cs, ce, cst, strata, efron, eta, grad, hess, out);
    return glm_eval<double>(family, op, n, K, y, w, cs, ce, cst, strata, efron, eta, grad, hess, out);
}

int64_t orc_bcd_solve(int dtype, int solver, int64_t q, const void* L, const void* v, double l1, double l2, double tol, int64_t max_iters, void* x) {
    size_t iters = 0;
    if (dtype == ORC_F32) {
        std::vector<float> b1(q), b2(q);
        newton_prox<float>((const float*)L, (const float*)v, q, (float)l1, (float)l2, (float)tol, max_iters, solver == 1, (float*)x, iters, b1.data(), b2.data());
    } else {
        std::vector<double> b1(q), b2(q);
        newton_prox<double>((const double*)L, (const double*)v, q, l1, l2, tol, max_iters, solver == 1, (double*)x, iters, b1.data(), b2.data());
    }
    return (int64_t)iters;
}
double orc_bcd_root_lower_bound(int64_t q, const double* D, const double* v, double l1) { return root_lower_bound<double>(D, v, q, l1); }
double orc_bcd_root_upper_bound(int64_t q, const double* D, const double* v, double l1, double zero_tol) { return root_upper_bound<double>(D, v, q, l1, zero_tol).first; }
double orc_bcd_root_function(int64_t q, double h, const double* D, const double* v, double l1) { return root_function<double>(h, D, v, q, l1); }
int orc_search_pivot(int64_t n, const double* x, const double* y, double* mses) { return search_pivot<double>(x, y, n, mses); }
void orc_jacobi_eigh(int64_t q, double* A, double* D, double* V) { jacobi_eigh<double>(A, q, D, V); }
void orc_set_config(const char* name, double value) {
    std::string s(name);
    if (s == "hessian_min") Configs::hessian_min = value;
    else if (s == "dbeta_tol") Configs::dbeta_tol = value;
    else if (s == "min_bytes") Configs::min_bytes = (size_t)value;
}

} // extern "C"
