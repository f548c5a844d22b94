/* include/adelie_b200.h -- C ABI of libadelie_b200.so (B200 / sm_100a).
 *
 * This is the drop-in boundary for the block-coordinate-descent path of JamesYang007/adelie:
 * the entry points a binding of `adelie.adelie_core` would call for the naive-method
 * group-elastic-net solver.  Plain pointers and sizes only; every function returns 0 on
 * success and a non-zero code on failure, in which case ab_last_error() holds the message
 * (the text matches the reference's adelie_core_error / adelie_core_solver_error strings).
 *
 * Each block cites the reference interface it replaces (paths relative to the reference root).
 * dtype: 0 = float32, 1 = float64 -- the element type of every `void*` value array.
 */
#ifndef ADELIE_B200_H
#define ADELIE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ab_matrix ab_matrix;   /* adelie_core.matrix.MatrixNaive* handle */
typedef struct ab_glm ab_glm;         /* adelie_core.glm.Glm* handle */
typedef struct ab_state ab_state;     /* adelie_core.state.State*Naive handle */
typedef struct ab_cov_matrix ab_cov_matrix; /* adelie_core.matrix.MatrixCov* handle */
typedef struct ab_cov_state ab_cov_state;   /* adelie_core.state.StateGaussianCov / StateGaussianPinCov handle */
typedef struct ab_io_snp ab_io_snp;   /* adelie_core.io.IOSNPUnphased handle */
typedef struct ab_io_snp_pa ab_io_snp_pa; /* adelie_core.io.IOSNPPhasedAncestry handle */

enum { AB_F32 = 0, AB_F64 = 1 };
enum { AB_OK = 0, AB_ERR_CORE = 1 /* adelie_core_error -> RuntimeError */, AB_ERR_CUDA = 2, AB_ERR_ARG = 3 };

/* ---- library / device ---------------------------------------------------------------------- */
const char* ab_last_error(void);
int ab_version(void);
int ab_device_count(int* count);
int ab_set_device(int device);
int ab_get_device_info(int* sm_count, size_t* smem_optin_bytes, size_t* total_mem_bytes);
int ab_mem_info(size_t* free_bytes, size_t* total_bytes);   /* cudaMemGetInfo of the current device (leak regression tests, bench) */
int ab_device_synchronize(void);
/* pin / unpin a host buffer (bench: H2D from pinned memory); CUDA-event stopwatch on the library's stream */
int ab_host_register(void* ptr, size_t bytes);
int ab_host_unregister(void* ptr);
int ab_timer_start(void);
int ab_timer_stop(double* elapsed_ms);

/* ---- multi-GPU row sharding (no counterpart in the reference, which is single-process): one process per GPU of one
 *      NVSwitch box.  ab_dist_init allocates this rank's peer-visible slab and returns its 64-byte cudaIpc handle; the caller
 *      all-gathers the handles (e.g. torch.distributed) and passes them, in rank order, to ab_dist_connect.  Afterwards every
 *      matrix / GLM / state call is collective: each rank passes its row shard, results are identical on all ranks. */
int ab_dist_init(int rank, int world, void* ipc_handle_out /* 64 bytes */);
int ab_dist_connect(const void* all_handles /* world * 64 bytes */);
int ab_dist_allreduce_f64(double* host_buf, int64_t n);      /* in-place sum over ranks of a host vector */
int ab_dist_info(int* rank, int* world);

/* ---- configs: adelie/src/py_configs.cpp:6-47, include/adelie_core/configs.hpp:6-20 ---------- */
int ab_configs_set(const char* name, double value);
int ab_configs_get(const char* name, double* value);

/* ---- dense matrix: adelie/src/py_matrix.cpp:1940-1943 (MatrixNaiveDense{32,64}{C,F}),
 *      operators adelie/src/py_matrix.cpp:832-1071 = matrix_naive_base.hpp:57-143 ------------- */
/* order: 0 = column-major (F), 1 = row-major (C); ldh = leading dimension of the host array.
 * The matrix is copied to HBM once (column-major, rows padded to 32) and stays resident. */
int ab_matrix_dense_create(int dtype, const void* host, int64_t n, int64_t p, int order, int64_t ldh, int n_threads, ab_matrix** out);
/* device-side allocation + synthetic fill (bench / sharded generation): X[i,j] ~ N(0,1) from
 * Philox(seed, subsequence=j, offset=row_offset+i), identical for every shard layout */
int ab_matrix_dense_alloc(int dtype, int64_t n, int64_t p, ab_matrix** out);
int ab_matrix_dense_fill_normal(ab_matrix* m, uint64_t seed, int64_t row_offset);
int ab_matrix_dense_download(ab_matrix* m, void* host, int64_t row0, int64_t nrows, int64_t col0, int64_t ncols, int64_t ldh);
/* adelie.matrix.sparse (PY/matrix.py sparse(); MatrixNaiveSparse{32,64}F, BIND/py_matrix.cpp:1878-1968; CORE/matrix/matrix_naive_sparse.ipp:10-262):
 * CSC arrays as scipy holds them (column pointers widened to int64, int32 row indices sorted and unique inside every column). */
int ab_matrix_sparse_create(int dtype, int64_t n, int64_t p, int64_t nnz, const int64_t* indptr, const int32_t* indices, const void* values,
                            int n_threads, ab_matrix** out);
/* bench helper: random sparse matrix generated in HBM (exactly nnz_per_col non-zeros per column, N(0,1) values) */
int ab_matrix_sparse_alloc_random(int dtype, int64_t n, int64_t p, int64_t nnz_per_col, uint64_t seed, ab_matrix** out);
int ab_matrix_sparse_nnz(const ab_matrix* m, int64_t* out);
int ab_matrix_sparse_download(ab_matrix* m, int64_t* indptr, int32_t* indices, void* values);
/* adelie.io.snp_unphased (PY/io.py:114-196; IOSNPUnphased: BIND/py_io.cpp, CORE/io/io_snp_unphased.hpp:137-274, .ipp:9-305, read(): CORE/io/io_snp_base.ipp:20-84).
   `.snpdat` reader / writer on the host: read_mode "file" | "mmap" | "auto"; write() takes column-major (n, p) int8 calldata (negative = missing, values > 2
   rejected), impute_method "mean" (impute[] is an output) or "user" (impute[] is an input); get(): name in {"nnz","nnm","outer"} -> uint64, "impute" -> double;
   to_dense(): (n, p) row-major int8 with -9 for missing. */
int ab_io_snp_unphased_create(const char* filename, const char* read_mode, ab_io_snp** out);
int ab_io_snp_unphased_free(ab_io_snp* io);
int ab_io_snp_unphased_write(ab_io_snp* io, const int8_t* calldata, int64_t n, int64_t p, const char* impute_method, double* impute, int64_t impute_len,
                             int n_threads, uint64_t* total_bytes);
int ab_io_snp_unphased_read(ab_io_snp* io, uint64_t* total_bytes);
int ab_io_snp_unphased_info(const ab_io_snp* io, int* is_read, int64_t* rows, int64_t* snps);
int ab_io_snp_unphased_get(const ab_io_snp* io, const char* name, void* out);
int ab_io_snp_unphased_to_dense(const ab_io_snp* io, int n_threads, int8_t* out);
/* adelie.matrix.snp_unphased (PY/matrix.py:1243-1298; MatrixNaiveSNPUnphased{32,64}: BIND/py_matrix.cpp:1878-1968, CORE/matrix/matrix_naive_snp_unphased.ipp:10-309):
   entries 0 / 1 / 2 / impute[j] (missing).  The file bytes are unpacked on the device into 2 bits per genotype; rows [row_lo, row_hi) of the file are kept
   (row_hi < 0: all rows) so that the ranks of a row-sharded run can share one file.  Every ab_matrix_* operator applies. */
int ab_matrix_snp_unphased_create(int dtype, const ab_io_snp* io, int64_t row_lo, int64_t row_hi, int n_threads, ab_matrix** out);
int ab_matrix_snp_unphased_from_calldata(int dtype, const int8_t* calldata, int64_t n, int64_t p, const double* impute, int n_threads, ab_matrix** out);
/* bench helper: random genotypes generated in HBM (1 w.p. one_ratio, 2 w.p. two_ratio, masked as missing w.p. missing_ratio; Philox per (column, global row):
   every row sharding sees the same matrix); impute = column mean of the non-missing entries over all n_total rows (all-reduced when row-sharded) */
int ab_matrix_snp_unphased_alloc_random(int dtype, int64_t n, int64_t p, uint64_t seed, int64_t row_offset, int64_t n_total,
                                        double one_ratio, double two_ratio, double missing_ratio, ab_matrix** out);
int ab_matrix_snp_unphased_download(ab_matrix* m, int8_t* calldata_out, double* impute_out);   /* column-major (n, p) int8, -9 = missing */
int ab_matrix_snp_unphased_cache_info(const ab_matrix* m, int64_t* cached_cols, int64_t* packed_bytes);
/* adelie.io.snp_phased_ancestry (PY/io.py:6-111; IOSNPPhasedAncestry: CORE/io/io_snp_phased_ancestry.hpp, .ipp:9-363) and adelie.matrix.snp_phased_ancestry
   (MatrixNaiveSNPPhasedAncestry{32,64}, CORE/matrix/matrix_naive_snp_phased_ancestry.ipp): calldata / ancestries are column-major (n, 2 s) int8 (calldata in
   {0,1}, ancestries in [0, A)); the matrix is (n, s*A) with entries 0 / 1 / 2 and shares the 2-bit device storage and every kernel of snp_unphased.
   get(): name in {"nnz0","nnz1","outer"} -> uint64; to_dense(): (n, s*A) row-major int8. */
int ab_io_snp_phased_ancestry_create(const char* filename, const char* read_mode, ab_io_snp_pa** out);
int ab_io_snp_phased_ancestry_free(ab_io_snp_pa* io);
int ab_io_snp_phased_ancestry_write(ab_io_snp_pa* io, const int8_t* calldata, const int8_t* ancestries, int64_t n, int64_t two_s, int64_t A,
                                    int n_threads, uint64_t* total_bytes);
int ab_io_snp_phased_ancestry_read(ab_io_snp_pa* io, uint64_t* total_bytes);
int ab_io_snp_phased_ancestry_info(const ab_io_snp_pa* io, int* is_read, int64_t* rows, int64_t* snps, int64_t* ancestries);
int ab_io_snp_phased_ancestry_get(const ab_io_snp_pa* io, const char* name, void* out);
int ab_io_snp_phased_ancestry_to_dense(const ab_io_snp_pa* io, int n_threads, int8_t* out);
int ab_matrix_snp_phased_ancestry_create(int dtype, const ab_io_snp_pa* io, int64_t row_lo, int64_t row_hi, int n_threads, ab_matrix** out);
/* adelie.matrix.standardize (PY/matrix.py:1414-1536; MatrixNaiveStandardize{32,64}, CORE/matrix/matrix_naive_standardize.ipp:8-293): X = (Z - 1 c^T) diag(s)^-1,
   and adelie.matrix.subset (PY/matrix.py:1539-1632; MatrixNaiveCSubset / MatrixNaiveRSubset, CORE/matrix/matrix_naive_subset.ipp): axis 0 = rows, 1 = columns.
   Both materialise a new dense device matrix from a dense base matrix (centers / scales: host arrays of the matrix dtype). */
int ab_matrix_standardize_create(ab_matrix* base, const void* centers, int64_t n_centers, const void* scales, int64_t n_scales, int n_threads, ab_matrix** out);
int ab_matrix_subset_create(ab_matrix* base, const int64_t* indices, int64_t n_indices, int axis, int n_threads, ab_matrix** out);
int ab_matrix_free(ab_matrix* m);
int ab_matrix_rows(const ab_matrix* m, int64_t* out);
int ab_matrix_cols(const ab_matrix* m, int64_t* out);
/* host-pointer operators (v, w, out are host arrays of the matrix dtype; semantics of the reference) */
int ab_matrix_cmul(ab_matrix* m, int64_t j, const void* v, const void* w, double* out);                 /* X[:,j]^T (v*w) */
int ab_matrix_ctmul(ab_matrix* m, int64_t j, double v, void* out);                                       /* out += v X[:,j] */
int ab_matrix_bmul(ab_matrix* m, int64_t j, int64_t q, const void* v, const void* w, void* out);        /* X[:,j:j+q]^T (v*w) */
int ab_matrix_btmul(ab_matrix* m, int64_t j, int64_t q, const void* v, void* out);                      /* out += X[:,j:j+q] v */
int ab_matrix_mul(ab_matrix* m, const void* v, const void* w, void* out);                               /* X^T (v*w) */
/* `mul` of kron(X, I_K) (MatrixNaiveKroneckerEye::mul, adelie_core/matrix/matrix_naive_kronecker_eye.ipp:29-352) in ONE pass over X:
 * v, w (n, K) row-major, out (p, K) row-major: out[j, l] = sum_i X[i, j] v[i, l] w[i, l].  Dense and snp matrices; packed genotypes in
 * float32 with 2 <= K <= 8 run on the INT8 tensor-core kernel (csrc/snp_tc.cuh). */
int ab_matrix_mul_multi(ab_matrix* m, int64_t K, const void* v, const void* w, void* out);
/* Weighted Gram of a window of columns, out[s * ncol + u] = sum_i w_i X[i, cols[s]] X[i, cols[u]] for s < n_src <= 64, u < ncol <= 128
 * (row-major n_src x ncol, double): the Gram-panel kernel of the batched sweep exposed for parity tests.  Generalises cov()
 * (matrix_naive_base.hpp:101-105) to a column list.  use_tc = 1: tcgen05 tensor cores with TF32 operands, 0: fp32 CUDA cores.
 * Dense float32 matrices only. */
int ab_matrix_window_gram(ab_matrix* m, const int32_t* cols, int ncol, int n_src, const void* w, int use_tc, double* out);
int ab_matrix_cov(ab_matrix* m, int64_t j, int64_t q, const void* sqrt_w, void* out /* q*q col-major */);
int ab_matrix_sq_mul(ab_matrix* m, const void* w, void* out);                                           /* (X*X)^T w */
/* out (L x n, row-major) = v (L x p CSR) X^T : adelie/src/py_matrix.cpp sp_tmul */
int ab_matrix_sp_tmul(ab_matrix* m, int64_t L, const int64_t* indptr, const int64_t* indices, const void* values, void* out);

/* ---- GLM families: adelie/src/py_glm.cpp:101-234, 664-685 ---------------------------------- */
enum { AB_GLM_GAUSSIAN = 1, AB_GLM_BINOMIAL_LOGIT = 2, AB_GLM_MULTIGAUSSIAN = 3, AB_GLM_COX = 4, AB_GLM_POISSON = 5 /* GlmPoisson, CORE/glm/glm_poisson.ipp:7-66 */,
       AB_GLM_BINOMIAL_PROBIT = 6 /* GlmBinomialProbit, CORE/glm/glm_binomial.ipp:100-190 */,
       AB_GLM_MULTINOMIAL = 7 /* GlmMultinomial, CORE/glm/glm_multinomial.ipp:6-132; y (n, K) row-major */ };
/* y: (n,) or (n,K) row-major; weights (n,) summing to 1.  Cox: y = status, plus start/stop/strata, tie 1 = efron, 0 = breslow */
int ab_glm_create(int dtype, int family, int64_t n, int64_t K, const void* y, const void* weights,
                  const void* cox_start, const void* cox_stop, const int64_t* cox_strata, int cox_tie_efron, ab_glm** out);
/* User-defined GLM (the pybind trampoline PyGlmBase / PyGlmMultiBase, adelie/src/py_glm.cpp:8-92, 240-330: a Python subclass of
 * adelie.glm.GlmBase{32,64} / GlmMultiBase{32,64} overriding the virtuals).  The callbacks receive HOST arrays of n (single response) or
 * n*K (multi-response, (n, K) row-major) elements of the state's dtype and return 0, or non-zero when the callee raised (the solve then
 * stops with a solver error).  The library moves eta / grad / hess between HBM and pinned host memory around every call: one round trip
 * per IRLS iteration, the coordinate descent itself stays on the device.  inv_hessian_gradient / inv_link may be NULL (inv_hessian_gradient
 * then uses the base-class formula grad / max(hess, hessian_min), adelie_core/glm/glm_base.ipp:25-36). */
typedef struct ab_glm_callbacks {
    void* ctx;
    int (*gradient)(void* ctx, const void* eta, void* grad);
    int (*hessian)(void* ctx, const void* eta, const void* grad, void* hess);
    int (*inv_hessian_gradient)(void* ctx, const void* eta, const void* grad, const void* hess, void* out);
    int (*loss)(void* ctx, const void* eta, double* out);
    int (*loss_full)(void* ctx, double* out);
    int (*inv_link)(void* ctx, const void* eta, void* out);
} ab_glm_callbacks;
int ab_glm_create_callback(int dtype, int64_t n, int64_t K, int is_multi, const ab_glm_callbacks* callbacks, ab_glm** out);
int ab_glm_free(ab_glm* g);
int ab_glm_gradient(ab_glm* g, const void* eta, void* grad);
int ab_glm_hessian(ab_glm* g, const void* eta, const void* grad, void* hess);
int ab_glm_inv_hessian_gradient(ab_glm* g, const void* eta, const void* grad, const void* hess, void* out);
int ab_glm_loss(ab_glm* g, const void* eta, double* out);
int ab_glm_loss_full(ab_glm* g, double* out);
int ab_glm_inv_link(ab_glm* g, const void* eta, void* out);

/* ---- states: constructors adelie/src/py_state.cpp:1054-1228 (StateGaussianNaive),
 *      :1557-1720 (StateGlmNaive), :1230-1400 (StateMultiGaussianNaive), :1722-1900 (StateMultiGlmNaive);
 *      keyword meaning as adelie/state.py:1966-2010, 2327-2377, 2694-2740 ---------------------- */
typedef struct ab_state_args {
    int32_t dtype;
    /* groups */
    const int64_t* groups; const int64_t* group_sizes; int64_t G; double alpha; const void* penalty;
    /* Gaussian ("opt") states: weights, X_means, y_mean, y_var, resid, resid_sum, rsq.  GLM states: offsets, eta, resid, beta0,
     * loss_null (NaN => setup_loss_null), loss_full, irls_* */
    const void* weights; const void* X_means; double y_mean, y_var, resid_sum, rsq;
    const void* resid; const void* offsets; const void* eta; double beta0, loss_null, loss_full;
    int32_t setup_loss_null; int64_t irls_max_iters; double irls_tol;
    /* multi-response: n_classes K (>1) and multi_intercept, else 1 / 0 */
    int64_t n_classes; int32_t multi_intercept;
    /* lambda path */
    const void* lmda_path; int64_t lmda_path_len; double lmda_max, min_ratio; int64_t lmda_path_size;
    int32_t setup_lmda_max, setup_lmda_path;
    /* limits / screening */
    int64_t max_screen_size, max_active_size; double pivot_subset_ratio; int64_t pivot_subset_min; double pivot_slack_ratio;
    int32_t screen_rule;      /* 0 strong, 1 pivot */
    /* convergence */
    int64_t max_iters; double tol, adev_tol, ddev_tol, newton_tol; int64_t newton_max_iters;
    int32_t early_exit, intercept; int64_t n_threads;
    /* warm-start invariants */
    const int64_t* screen_set; int64_t screen_set_size; const void* screen_beta; int64_t screen_beta_size;
    const int8_t* screen_is_active; int64_t active_set_size; const int64_t* active_set;
    double lmda; const void* grad;
} ab_state_args;

/* glm == NULL => Gaussian "opt" state (StateGaussianNaive / StateMultiGaussianNaive) */
int ab_state_create(const ab_state_args* args, ab_matrix* X, ab_glm* glm, ab_state** out);
int ab_state_free(ab_state* s);
/* solve: adelie/src/py_state.cpp:62-145 (_solve) -- never throws: the solver error string is written to `err`
 * (empty on success), the state is valid up to the last solved lambda; total_time in seconds.
 * exit_cond(ctx) != 0 requests early exit after a solved lambda; check_signals() != 0 interrupts (PyErr_CheckSignals). */
int ab_state_solve(ab_state* s, int display_progress_bar, int (*exit_cond)(void*), void* ctx, int (*check_signals)(void),
                   char* err, size_t errlen, double* total_time);
/* outputs (field names as the Python attribute surface) */
int ab_state_get_scalar(const ab_state* s, const char* name, double* out);
int ab_state_get_vec_f64(const ab_state* s, const char* name, double* out, int64_t cap, int64_t* len);
int ab_state_get_vec_i64(const ab_state* s, const char* name, int64_t* out, int64_t cap, int64_t* len);
/* betas as CSR (L x p): call with NULL arrays to size, then again to fill */
int ab_state_get_betas(const ab_state* s, int64_t* indptr, int64_t* indices, double* values, int64_t* nnz, int64_t* L);
/* screen_transforms[i] (row-major gs x gs) */
int ab_state_get_screen_transform(const ab_state* s, int64_t i, double* out, int64_t cap, int64_t* len);

/* ---- pin state in isolation: StateGaussianPinNaive{32,64} (adelie/src/py_state.cpp:389-411; adelie/state.py:421-720), solve =
 *      pin::naive::solve (adelie_core/solver/solver_gaussian_pin_naive.hpp:223-401).  The state is an ab_state created by
 *      ab_state_create with glm == NULL: screen_set / screen_beta / screen_is_active / active_set / resid / rsq / resid_sum are the
 *      pin state's inputs, lmda_path the lambdas to solve (setup_lmda_max = setup_lmda_path = 0), `tol` is used unscaled, `grad`
 *      is ignored (zeros).  Solves every lambda of lmda_path on the FIXED screen set; outputs through the ab_state getters
 *      ("betas", "intercepts", "rsqs", "lmdas", "screen_beta", "screen_is_active", "active_set", "resid", "rsq", "resid_sum",
 *      "n_sweeps" = iters, "benchmark_fit_screen" / "benchmark_fit_active" = benchmark_screen / benchmark_active). */
int ab_pin_naive_solve(ab_state* s, int (*check_signals)(void), char* err, size_t errlen, double* total_time);

/* ---- covariance method ----------------------------------------------------------------------------------
 * MatrixCov: adelie/src/py_matrix.cpp MatrixCovBase{32,64} (bmul / mul / to_dense / cols, adelie_core/matrix/matrix_cov_base.hpp:20-63),
 * MatrixCovDense{32,64}{C,F} (adelie.matrix.dense(method="cov"), matrix_cov_dense.ipp:8-84) and MatrixCovLazyCov{32,64}{C,F}
 * (adelie.matrix.lazy_cov, matrix_cov_lazy_cov.ipp:8-190: A = X^T X, rows computed on first use and kept in HBM).
 * order: 0 = column-major, 1 = row-major; host arrays of the matrix dtype; indices / subset are int64. */
int ab_matrix_cov_dense_create(int dtype, const void* host, int64_t p, int order, int64_t ldh, int n_threads, ab_cov_matrix** out);
int ab_matrix_cov_lazy_create(int dtype, const void* host, int64_t n, int64_t p, int order, int64_t ldh, int n_threads, ab_cov_matrix** out);
int ab_matrix_cov_free(ab_cov_matrix* m);
int ab_matrix_cov_cols(const ab_cov_matrix* m, int64_t* out);
int ab_matrix_cov_bmul(ab_cov_matrix* m, const int64_t* subset, int64_t s, const int64_t* indices, const void* values, int64_t k, void* out /* (s,) */);
int ab_matrix_cov_mul(ab_cov_matrix* m, const int64_t* indices, const void* values, int64_t k, void* out /* (p,) */);
int ab_matrix_cov_to_dense(ab_cov_matrix* m, int64_t i, int64_t q, void* out /* q*q column-major */);
int ab_matrix_cov_cache_info(const ab_cov_matrix* m, int64_t* cached_rows);     /* lazy_cov: rows of A held in HBM */
/* States: StateGaussianCov{32,64} (adelie/src/py_state.cpp, adelie/state.py:1128-1420; solve = gaussian::cov::solve,
 * adelie_core/solver/solver_gaussian_cov.hpp:359-457) and StateGaussianPinCov{32,64} (adelie/state.py:739-1000; solve =
 * gaussian::pin::cov::solve, solver_gaussian_pin_cov.hpp:529-725).  One argument block for both: the path state reads v / grad /
 * lmda / lmda_max / the screening configuration, the pin state reads screen_grad and solves lmda_path on the FIXED screen_set
 * (screen_vars / screen_transforms / screen_subset_order are derived from A like the Python wrapper does, state.py:912-936). */
typedef struct ab_cov_state_args {
    int32_t dtype;
    const void* v;                                   /* (p,) linear term (path state) */
    const int64_t* groups; const int64_t* group_sizes; int64_t G; double alpha; const void* penalty;
    const void* lmda_path; int64_t lmda_path_len; double lmda_max, min_ratio; int64_t lmda_path_size;
    int32_t setup_lmda_max, setup_lmda_path;
    int64_t max_screen_size, max_active_size; double pivot_subset_ratio; int64_t pivot_subset_min; double pivot_slack_ratio;
    int32_t screen_rule;                             /* 0 strong, 1 pivot */
    int64_t max_iters; double tol, rdev_tol, newton_tol; int64_t newton_max_iters;
    int32_t early_exit; int64_t n_threads;
    const int64_t* screen_set; int64_t screen_set_size; const void* screen_beta; int64_t screen_beta_size;
    const int8_t* screen_is_active; int64_t active_set_size; const int64_t* active_set;
    double rsq, lmda; const void* grad;             /* grad (p,): path state */
    const void* screen_grad;                         /* (screen_beta_size,): pin state; NULL for the path state */
} ab_cov_state_args;
int ab_cov_state_create(const ab_cov_state_args* args, ab_cov_matrix* A, ab_cov_state** out);
int ab_cov_state_free(ab_cov_state* s);
/* same contract as ab_state_solve / ab_pin_naive_solve: never throws, the solver error string goes to `err` */
int ab_cov_state_solve(ab_cov_state* s, int display_progress_bar, int (*exit_cond)(void*), void* ctx, int (*check_signals)(void),
                       char* err, size_t errlen, double* total_time);
int ab_cov_pin_solve(ab_cov_state* s, int (*check_signals)(void), char* err, size_t errlen, double* total_time);
int ab_cov_state_get_scalar(const ab_cov_state* s, const char* name, double* out);
int ab_cov_state_get_vec_f64(const ab_cov_state* s, const char* name, double* out, int64_t cap, int64_t* len);
int ab_cov_state_get_vec_i64(const ab_cov_state* s, const char* name, int64_t* out, int64_t cap, int64_t* len);
int ab_cov_state_get_betas(const ab_cov_state* s, int64_t* indptr, int64_t* indices, double* values, int64_t* nnz, int64_t* L);
int ab_cov_state_get_screen_transform(const ab_cov_state* s, int64_t i, double* out, int64_t cap, int64_t* len);
/* ---- bcd prox: adelie/src/py_bcd.cpp:15-243 (double only, like the reference) --------------- */
/* solver: 0 newton, 1 newton_abs */
int ab_bcd_solve(int solver, int64_t q, const double* quad, const double* linear, double l1, double l2, double tol, int64_t max_iters,
                 double* x, int64_t* iters);
int ab_bcd_root_lower_bound(int64_t q, const double* quad, const double* linear, double l1, double* out);
int ab_bcd_root_upper_bound(int64_t q, const double* quad, const double* linear, double l1, double zero_tol, double* out);
int ab_bcd_root_function(int64_t q, double h, const double* D, const double* v, double l1, double* out);

#ifdef __cplusplus
}
#endif
#endif /* ADELIE_B200_H */
