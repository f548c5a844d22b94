/* oracle/oracle_capi.h -- TEST INFRASTRUCTURE (see adelie_oracle.hpp header).
 * Plain-C entry points of the CPU oracle, loaded with ctypes from
 * oracle/oracle.py.  Never used by the product package. */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_F32 = 0, ORC_F64 = 1 };
enum { ORC_MAT_DENSE = 0, ORC_MAT_SPARSE = 1 };
/* families: 0 = gaussian "opt" path (state_gaussian_naive / multigaussian_naive),
 * 1 = gaussian through IRLS, 2 = binomial logit, 3 = multigaussian through IRLS, 4 = cox */
enum { ORC_FAM_GAUSSIAN_OPT = 0, ORC_FAM_GAUSSIAN = 1, ORC_FAM_BINOMIAL = 2, ORC_FAM_MULTIGAUSSIAN = 3, ORC_FAM_COX = 4, ORC_FAM_POISSON = 5, ORC_FAM_BINOMIAL_PROBIT = 6, ORC_FAM_MULTINOMIAL = 7 };

typedef struct orc_path_args {
    int32_t dtype;            /* ORC_F32 / ORC_F64: type of every void* array below */
    int32_t matrix_kind;      /* ORC_MAT_* */
    /* dense column-major (n x p, leading dimension ld) */
    const void* X; int64_t n, p, ld;
    /* sparse CSC */
    const int32_t* sp_outer; const int32_t* sp_inner; const void* sp_values;
    /* multi-response: K > 1 wraps X as [kron(1,I_K) if multi_intercept] ++ kron(X, I_K) */
    int64_t K; int32_t multi_intercept;
    /* GLM */
    int32_t family;
    const void* y; const void* weights; const void* offsets;   /* y, offsets: (n,) or (n,K) row-major */
    /* cox extras */
    const void* cox_start; const void* cox_stop; const void* cox_status; const int64_t* cox_strata; int32_t cox_efron;
    /* groups over the (augmented) columns */
    const int64_t* groups; const int64_t* group_sizes; int64_t G; const void* penalty; double alpha;
    /* invariants */
    const void* X_means; double y_mean, y_var, rsq, resid_sum;
    const void* resid; const void* grad; const void* eta; double beta0, loss_null, loss_full; int32_t setup_loss_null;
    const int64_t* screen_set; int64_t S; const void* screen_beta; int64_t screen_beta_size;
    const int8_t* screen_is_active; int64_t active_set_size; const int64_t* active_set;
    double lmda, lmda_max; const void* lmda_path; int64_t lmda_path_len;
    int32_t setup_lmda_max, setup_lmda_path;
    /* configs */
    double min_ratio; int64_t lmda_path_size, max_screen_size, max_active_size;
    double pivot_subset_ratio; int64_t pivot_subset_min; double pivot_slack_ratio; int32_t screen_rule;
    int64_t max_iters; double tol, adev_tol, ddev_tol, newton_tol; int64_t newton_max_iters;
    int64_t irls_max_iters; double irls_tol;
    int32_t early_exit, intercept, n_threads;
    double max_seconds;       /* oracle-only wall-clock budget (<=0: none) */
} orc_path_args;

/* Runs the path; returns an opaque result handle (never NULL).  */
void* orc_path_solve(const orc_path_args* args);
void orc_result_free(void* h);
const char* orc_result_error(void* h);
double orc_result_scalar(void* h, const char* name);
/* copies vector `name` as doubles into out (capacity cap); returns its length */
int64_t orc_result_vec(void* h, const char* name, double* out, int64_t cap);
int64_t orc_result_ivec(void* h, const char* name, int64_t* out, int64_t cap);
/* betas as CSR: returns nnz; fill indptr (L+1), indices (nnz), values (nnz) when non-NULL */
int64_t orc_result_betas(void* h, int64_t* indptr, int64_t* indices, double* values);

/* Pin solve in isolation (solver_gaussian_pin_naive.hpp:223-401) on a dense matrix. */
typedef struct orc_pin_args {
    int32_t dtype; const void* X; int64_t n, p, ld;
    double y_mean, y_var;
    const int64_t* groups; const int64_t* group_sizes; int64_t G; double alpha; const void* penalty; const void* weights;
    const int64_t* screen_set; int64_t S;
    const void* lmda_path; int64_t L;
    int32_t intercept; int64_t max_active_size, max_iters; double tol, adev_tol, ddev_tol, newton_tol; int64_t newton_max_iters;
    int32_t n_threads;
    double rsq; void* resid; double resid_sum;      /* resid updated in place */
    void* screen_beta; int8_t* screen_is_active; int64_t active_set_size; int64_t* active_set;  /* in place */
} orc_pin_args;
void* orc_pin_solve(orc_pin_args* args);

/* Dense matrix operators (matrix_naive_dense.ipp) for kernel parity tests. */
double orc_dense_cmul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, const void* v, const void* w, int n_threads);
void orc_dense_ctmul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, double v, void* out, int n_threads);
void orc_dense_bmul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, int64_t q, const void* v, const void* w, void* out, int n_threads);
void orc_dense_btmul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, int64_t q, const void* v, void* out, int n_threads);
void orc_dense_mul(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, const void* v, const void* w, void* out, int n_threads);
void orc_dense_cov(int dtype, const void* X, int64_t n, int64_t p, int64_t ld, int64_t j, int64_t q, const void* sqrt_w, void* out, int n_threads);

/* GLM families: op 0 gradient, 1 hessian, 2 inv_hessian_gradient, 3 loss (returns), 4 loss_full (returns), 5 inv_link */
double orc_glm_eval(int dtype, int family, int op, int64_t n, int64_t K, const void* y, const void* w,
                    const void* cox_start, const void* cox_stop, const void* cox_status, const int64_t* cox_strata, int cox_efron,
                    const void* eta, const void* grad, const void* hess, void* out);

/* Prox sub-problem (newton.hpp); solver 0 = newton, 1 = newton_abs. returns iters */
int64_t orc_bcd_solve(int dtype, int solver, int64_t q, const void* L, const void* v, double l1, double l2, double tol, int64_t max_iters, void* x);
double orc_bcd_root_lower_bound(int64_t q, const double* D, const double* v, double l1);
double orc_bcd_root_upper_bound(int64_t q, const double* D, const double* v, double l1, double zero_tol);
double orc_bcd_root_function(int64_t q, double h, const double* D, const double* v, double l1);
int orc_search_pivot(int64_t n, const double* x, const double* y, double* mses);
void orc_jacobi_eigh(int64_t q, double* A, double* D, double* V);
void orc_set_config(const char* name, double value);

/* ---- covariance method (oracle/cov_oracle.hpp): gaussian_cov path solve, gaussian_pin_cov solve, MatrixCov operators ---- */
enum { ORC_COV_DENSE = 0, ORC_COV_LAZY = 1 };
typedef struct orc_cov_args {
    int32_t dtype; int32_t matrix_kind;      /* ORC_COV_DENSE: M = A (p x p); ORC_COV_LAZY: M = X (n x p), A = X^T X */
    const void* M; int64_t n, p, ld; int32_t row_major;
    const void* v;                            /* (p,) linear term (path solve) */
    const int64_t* groups; const int64_t* group_sizes; int64_t G; const void* penalty; double alpha;
    const int64_t* screen_set; int64_t S; const void* screen_beta; int64_t screen_beta_size;
    const int8_t* screen_is_active; int64_t active_set_size; const int64_t* active_set;
    double rsq, lmda, lmda_max; const void* grad;             /* grad (p,): path solve */
    const void* screen_grad;                                   /* pin solve: gradient on the screen values */
    const void* lmda_path; int64_t lmda_path_len; int32_t setup_lmda_max, setup_lmda_path;
    double min_ratio; int64_t lmda_path_size, max_screen_size, max_active_size;
    double pivot_subset_ratio; int64_t pivot_subset_min; double pivot_slack_ratio; int32_t screen_rule;
    int64_t max_iters; double tol, rdev_tol, newton_tol; int64_t newton_max_iters; int32_t early_exit;
} orc_cov_args;
void* orc_cov_path_solve(const orc_cov_args* args);
void* orc_cov_pin_solve(const orc_cov_args* args);
/* op 0 bmul (subset s, indices/values k -> out (s,)), 1 mul (indices/values k -> out (p,)), 2 to_dense (i0, q -> out q*q col-major) */
void orc_cov_matrix_op(int dtype, int kind, const void* M, int64_t n, int64_t p, int64_t ld, int row_major, int op,
                       const int64_t* subset, int64_t s, const int64_t* indices, const void* values, int64_t k,
                       int64_t i0, int64_t q, void* out);

#ifdef __cplusplus
}
#endif
