// adelie_b200/csrc/capi.cu -- the single translation unit of libadelie_b200.so: extern "C"
// entry points declared in include/adelie_b200.h on top of the templated host objects.
#include "../../include/adelie_b200.h"
#include "common.cuh"
#include "device_prims.cuh"
#include "sweep.cuh"
#include "dense_kernels.cuh"
#include "matrix.cuh"
#include "dist.cuh"
#include "glm.cuh"
#include "cox.cuh"
#include "solver.cuh"
#include "solver_glm.cuh"
#include "cov.cuh"
#include <memory>

using namespace ab;

static thread_local std::string g_last_error;

struct ab_matrix { int dtype; DenseMatrix<float>* f32 = nullptr; DenseMatrix<double>* f64 = nullptr; };
struct ab_glm { int dtype; int family; Glm<float>* f32 = nullptr; Glm<double>* f64 = nullptr; };
struct ab_io_snp { SnpUnphasedIO io; ab_io_snp(const char* f, const char* m) : io(f, m) {} };
struct ab_io_snp_pa { SnpPhasedAncestryIO io; ab_io_snp_pa(const char* f, const char* m) : io(f, m) {} };
struct ab_cov_matrix { int dtype; CovMatrix<float>* f32 = nullptr; CovMatrix<double>* f64 = nullptr; };
struct ab_cov_state { int dtype; CovPathState<float>* f32 = nullptr; CovPathState<double>* f64 = nullptr; std::string error; double total_time = 0; };
struct ab_state { int dtype; PathState<float>* f32 = nullptr; PathState<double>* f64 = nullptr; std::string error; double total_time = 0; };

#define AB_TRY try {
#define AB_CATCH                                                                                  \
    } catch (const core_error& e) { g_last_error = e.what(); return AB_ERR_CORE; }                \
    catch (const std::exception& e) { g_last_error = e.what(); return AB_ERR_CUDA; }              \
    return AB_OK;

template <class F32, class F64>
static auto dispatch(int dtype, F32 f32, F64 f64) { return dtype == AB_F32 ? f32() : f64(); }

extern "C" {

const char* ab_last_error(void) { return g_last_error.c_str(); }
int ab_version(void) { return 100; }
int ab_device_count(int* count) { AB_TRY AB_CUDA(cudaGetDeviceCount(count)); AB_CATCH }
int ab_set_device(int device) { AB_TRY AB_CUDA(cudaSetDevice(device)); AB_CATCH }
int ab_get_device_info(int* sm_count, size_t* smem, size_t* total_mem) {
    AB_TRY
    const auto& di = DeviceInfo::get();
    if (sm_count) *sm_count = di.sm_count;
    if (smem) *smem = di.smem_optin;
    if (total_mem) { size_t fr, tot; AB_CUDA(cudaMemGetInfo(&fr, &tot)); *total_mem = tot; }
    AB_CATCH
}
int ab_mem_info(size_t* free_bytes, size_t* total_bytes) {
    AB_TRY
    size_t fr, tot; AB_CUDA(cudaMemGetInfo(&fr, &tot));
    if (free_bytes) *free_bytes = fr;
    if (total_bytes) *total_bytes = tot;
    AB_CATCH
}
int ab_device_synchronize(void) { AB_TRY AB_CUDA(cudaDeviceSynchronize()); AB_CATCH }
int ab_host_register(void* ptr, size_t bytes) { AB_TRY AB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault)); AB_CATCH }
int ab_host_unregister(void* ptr) { AB_TRY AB_CUDA(cudaHostUnregister(ptr)); AB_CATCH }
// CUDA-event stopwatch on the stream every kernel of this library is launched on (the default stream)
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
int ab_timer_start(void) {
    AB_TRY
    if (!g_ev0) { AB_CUDA(cudaEventCreate(&g_ev0)); AB_CUDA(cudaEventCreate(&g_ev1)); }
    AB_CUDA(cudaDeviceSynchronize());
    AB_CUDA(cudaEventRecord(g_ev0, 0));
    AB_CATCH
}
int ab_timer_stop(double* ms) {
    AB_TRY
    if (!g_ev0) throw core_error("ab_timer_stop() without ab_timer_start().");
    AB_CUDA(cudaEventRecord(g_ev1, 0));
    AB_CUDA(cudaEventSynchronize(g_ev1));
    float f = 0; AB_CUDA(cudaEventElapsedTime(&f, g_ev0, g_ev1));
    *ms = f;
    AB_CATCH
}

int ab_configs_set(const char* name, double value) {
    const std::string s(name);
    if (s == "hessian_min") Configs::hessian_min = value;
    else if (s == "dbeta_tol") Configs::dbeta_tol = value;
    else if (s == "min_bytes") Configs::min_bytes = value;
    else if (s == "max_solver_value") Configs::max_solver_value = value;
    else if (s == "project") Configs::project = (int)value;
    else if (s == "sweep_ctas") Configs::sweep_ctas = (int)value;
    else if (s == "sweep_threads") Configs::sweep_threads = (int)value;
    else if (s == "sweep_min_rows_per_cta") Configs::sweep_min_rows_per_cta = (int)value;
    else if (s == "sweep_force_direct") Configs::sweep_force_direct = (int)value;
    else if (s == "device_eigh") Configs::device_eigh = (int)value;
    else if (s == "sweep_profile") Configs::sweep_profile = (int)value;
    else if (s == "sweep_batch") Configs::sweep_batch = (int)value;
    else if (s == "sweep_xchg") Configs::sweep_xchg = (int)value;
    else if (s == "panel_gemm") Configs::panel_gemm = (int)value;
    else if (s == "panel_tc") Configs::panel_tc = (int)value;
    else if (s == "snp_tc") Configs::snp_tc = (int)value;
    else if (s == "snp_tc_min_k") Configs::snp_tc_min_k = (int)value;
    else if (s == "kkt_skip_screen") Configs::kkt_skip_screen = (int)value;
    else if (s == "glm_batched") Configs::glm_batched = (int)value;
    else if (s == "cov_cluster") Configs::cov_cluster = (int)value;
    else if (s == "sweep_u_prefetch") Configs::sweep_u_prefetch = (int)value;
    else if (s == "glm_fuse_means") Configs::glm_fuse_means = (int)value;
    else if (s == "sweep_l2_prefetch") Configs::sweep_l2_prefetch = (int)value;
    else { g_last_error = "adelie_core: unknown config " + s; return AB_ERR_ARG; }
    return AB_OK;
}
int ab_configs_get(const char* name, double* value) {
    const std::string s(name);
    if (s == "hessian_min") *value = Configs::hessian_min;
    else if (s == "dbeta_tol") *value = Configs::dbeta_tol;
    else if (s == "min_bytes") *value = Configs::min_bytes;
    else if (s == "max_solver_value") *value = Configs::max_solver_value;
    else if (s == "project") *value = Configs::project;
    else if (s == "sweep_ctas") *value = Configs::sweep_ctas;
    else if (s == "sweep_threads") *value = Configs::sweep_threads;
    else if (s == "sweep_min_rows_per_cta") *value = Configs::sweep_min_rows_per_cta;
    else if (s == "sweep_force_direct") *value = Configs::sweep_force_direct;
    else if (s == "device_eigh") *value = Configs::device_eigh;
    else if (s == "sweep_profile") *value = Configs::sweep_profile;
    else if (s == "sweep_batch") *value = Configs::sweep_batch;
    else if (s == "sweep_xchg") *value = Configs::sweep_xchg;
    else if (s == "panel_gemm") *value = Configs::panel_gemm;
    else if (s == "panel_tc") *value = Configs::panel_tc;
    else if (s == "snp_tc") *value = Configs::snp_tc;
    else if (s == "snp_tc_min_k") *value = Configs::snp_tc_min_k;
    else if (s == "kkt_skip_screen") *value = Configs::kkt_skip_screen;
    else if (s == "glm_batched") *value = Configs::glm_batched;
    else if (s == "cov_cluster") *value = Configs::cov_cluster;
    else if (s == "sweep_u_prefetch") *value = Configs::sweep_u_prefetch;
    else if (s == "glm_fuse_means") *value = Configs::glm_fuse_means;
    else if (s == "sweep_l2_prefetch") *value = Configs::sweep_l2_prefetch;
    else { g_last_error = "adelie_core: unknown config " + s; return AB_ERR_ARG; }
    return AB_OK;
}

// ------------------------------------------------------------------------------------------ multi-GPU
int ab_dist_init(int rank, int world, void* ipc_handle_out /* 64 bytes */) {
    AB_TRY
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    DistContext::get().create(rank, world, (cudaIpcMemHandle_t*)ipc_handle_out);
    // all ranks restart the sweep epoch together so that the flagged lines of the level-3 exchange agree
    uint32_t one = 1; SweepContext::get().epoch.upload(&one, 1);
    AB_CUDA(cudaMemset(SweepContext::get().ll.p, 0, SweepContext::get().ll.n * sizeof(dev::LLLine)));
    AB_CUDA(cudaMemset(SweepContext::get().ll2.p, 0, SweepContext::get().ll2.n * sizeof(dev::LLLine)));
    AB_CUDA(cudaDeviceSynchronize());
    AB_CATCH
}
int ab_dist_connect(const void* all_handles /* world x 64 bytes, rank order */) {
    AB_TRY
    DistContext::get().connect((const cudaIpcMemHandle_t*)all_handles);
    AB_CATCH
}
int ab_dist_allreduce_f64(double* host_buf, int64_t n) {
    AB_TRY
    DistContext::get().allreduce_host(host_buf, n);
    AB_CATCH
}
int ab_dist_info(int* rank, int* world) {
    *rank = DistContext::get().active() ? DistContext::get().rank : 0;
    *world = DistContext::get().active() ? DistContext::get().world : 1;
    return AB_OK;
}

// ------------------------------------------------------------------------------------------ matrix
int ab_matrix_dense_alloc(int dtype, int64_t n, int64_t p, ab_matrix** out) {
    AB_TRY
    if (n < 1 || p < 1) throw core_error("matrix must have at least one row and one column.");
    auto* m = new ab_matrix{dtype};
    if (dtype == AB_F32) m->f32 = new DenseMatrix<float>(n, p); else m->f64 = new DenseMatrix<double>(n, p);
    *out = m;
    AB_CATCH
}
int ab_matrix_dense_create(int dtype, const void* host, int64_t n, int64_t p, int order, int64_t ldh, int n_threads, ab_matrix** out) {
    AB_TRY
    if (n_threads < 1) throw core_error("n_threads must be >= 1.");
    ab_matrix* m = nullptr;
    int rc = ab_matrix_dense_alloc(dtype, n, p, &m);
    if (rc) return rc;
    if (dtype == AB_F32) { m->f32->n_threads = n_threads; m->f32->upload((const float*)host, order, ldh); }
    else { m->f64->n_threads = n_threads; m->f64->upload((const double*)host, order, ldh); }
    *out = m;
    AB_CATCH
}
int ab_matrix_dense_fill_normal(ab_matrix* m, uint64_t seed, int64_t row_offset) {
    AB_TRY
    if (m->dtype == AB_F32) m->f32->fill_normal(seed, row_offset); else m->f64->fill_normal(seed, row_offset);
    AB_CUDA(cudaDeviceSynchronize());
    AB_CATCH
}
int ab_matrix_dense_download(ab_matrix* m, void* host, int64_t row0, int64_t nrows, int64_t col0, int64_t ncols, int64_t ldh) {
    AB_TRY
    if (m->dtype == AB_F32) m->f32->download((float*)host, row0, nrows, col0, ncols, ldh);
    else m->f64->download((double*)host, row0, nrows, col0, ncols, ldh);
    AB_CATCH
}
// sparse CSC (reference: adelie.matrix.sparse -> MatrixNaiveSparse{32,64}F, py_matrix.cpp:1878-1968; matrix_naive_sparse.ipp)
int ab_matrix_sparse_create(int dtype, int64_t n, int64_t p, int64_t nnz, const int64_t* indptr, const int32_t* indices, const void* values,
                            int n_threads, ab_matrix** out) {
    AB_TRY
    if (n < 1 || p < 1) throw core_error("matrix must have at least one row and one column.");
    if (n_threads < 1) throw core_error("n_threads must be >= 1.");
    if (indptr[0] != 0 || indptr[p] != nnz) throw core_error("sparse matrix: inconsistent column pointers.");
    for (int64_t j = 0; j < p; ++j) {
        if (indptr[j + 1] < indptr[j]) throw core_error("sparse matrix: column pointers must be non-decreasing.");
        for (int64_t k = indptr[j]; k < indptr[j + 1]; ++k) {
            if (indices[k] < 0 || indices[k] >= n) throw core_error("sparse matrix: row index out of range.");
            if (k > indptr[j] && indices[k] <= indices[k - 1]) throw core_error("sparse matrix: row indices must be sorted and unique inside every column.");
        }
    }
    auto* m = new ab_matrix{dtype};
    if (dtype == AB_F32) { m->f32 = new DenseMatrix<float>(n, p, true, nnz); m->f32->n_threads = n_threads; m->f32->upload_csc(indptr, indices, (const float*)values); }
    else { m->f64 = new DenseMatrix<double>(n, p, true, nnz); m->f64->n_threads = n_threads; m->f64->upload_csc(indptr, indices, (const double*)values); }
    *out = m;
    AB_CATCH
}
// random sparse matrix generated in HBM: exactly nnz_per_col non-zeros per column at sorted random rows, N(0,1) values
int ab_matrix_sparse_alloc_random(int dtype, int64_t n, int64_t p, int64_t nnz_per_col, uint64_t seed, ab_matrix** out) {
    AB_TRY
    if (n < 1 || p < 1 || nnz_per_col < 1 || nnz_per_col > n) throw core_error("sparse matrix: invalid shape.");
    auto* m = new ab_matrix{dtype};
    if (dtype == AB_F32) { m->f32 = new DenseMatrix<float>(n, p, true, p * nnz_per_col); m->f32->fill_sparse_random(nnz_per_col, seed); }
    else { m->f64 = new DenseMatrix<double>(n, p, true, p * nnz_per_col); m->f64->fill_sparse_random(nnz_per_col, seed); }
    AB_CUDA(cudaDeviceSynchronize());
    *out = m;
    AB_CATCH
}
int ab_matrix_sparse_nnz(const ab_matrix* m, int64_t* out) { *out = m->dtype == AB_F32 ? m->f32->nnz : m->f64->nnz; return AB_OK; }
int ab_matrix_sparse_download(ab_matrix* m, int64_t* indptr, int32_t* indices, void* values) {
    AB_TRY
    if (m->dtype == AB_F32) {
        auto& M = *m->f32; if (!M.sparse) throw core_error("not a sparse matrix.");
        M.sp_indptr.download(indptr, M.p + 1); M.sp_indices.download(indices, M.nnz); M.sp_values.download((float*)values, M.nnz);
    } else {
        auto& M = *m->f64; if (!M.sparse) throw core_error("not a sparse matrix.");
        M.sp_indptr.download(indptr, M.p + 1); M.sp_indices.download(indices, M.nnz); M.sp_values.download((double*)values, M.nnz);
    }
    AB_CUDA(cudaStreamSynchronize(0));
    AB_CATCH
}
// ------------------------------------------------------------------------------------------ SNP unphased: IO + matrix
// reference: adelie.io.snp_unphased (PY/io.py:114-196) -> IOSNPUnphased (BIND/py_io.cpp, CORE/io/io_snp_unphased.{hpp,ipp});
//            adelie.matrix.snp_unphased (PY/matrix.py:1243-1298) -> MatrixNaiveSNPUnphased{32,64} (CORE/matrix/matrix_naive_snp_unphased.ipp)
int ab_io_snp_unphased_create(const char* filename, const char* read_mode, ab_io_snp** out) {
    AB_TRY
    *out = new ab_io_snp(filename, read_mode);
    AB_CATCH
}
int ab_io_snp_unphased_free(ab_io_snp* io) { delete io; return AB_OK; }
int ab_io_snp_unphased_write(ab_io_snp* io, const int8_t* calldata, int64_t n, int64_t p, const char* impute_method, double* impute, int64_t impute_len,
                             int n_threads, uint64_t* total_bytes) {
    AB_TRY
    (void)n_threads;
    *total_bytes = io->io.write(calldata, (uint64_t)n, (uint64_t)p, impute_method, impute, (size_t)impute_len);
    AB_CATCH
}
int ab_io_snp_unphased_read(ab_io_snp* io, uint64_t* total_bytes) {
    AB_TRY
    *total_bytes = io->io.read();
    AB_CATCH
}
int ab_io_snp_unphased_info(const ab_io_snp* io, int* is_read, int64_t* rows, int64_t* snps) {
    *is_read = io->io.is_read ? 1 : 0; *rows = (int64_t)io->io.rows; *snps = (int64_t)io->io.snps;
    return AB_OK;
}
int ab_io_snp_unphased_get(const ab_io_snp* io, const char* name, void* out) {
    AB_TRY
    io->io.need_read();
    const std::string s(name);
    const auto& I = io->io;
    if (s == "nnz") std::memcpy(out, I.nnz.data(), 8 * I.snps);
    else if (s == "nnm") std::memcpy(out, I.nnm.data(), 8 * I.snps);
    else if (s == "impute") std::memcpy(out, I.impute.data(), 8 * I.snps);
    else if (s == "outer") std::memcpy(out, I.outer.data(), 8 * (I.snps + 1));
    else throw core_error("unknown field " + s);
    AB_CATCH
}
int ab_io_snp_unphased_to_dense(const ab_io_snp* io, int n_threads, int8_t* out) {
    AB_TRY
    (void)n_threads;
    io->io.to_dense(out);
    AB_CATCH
}

} // extern "C"

template <class T>
static DenseMatrix<T>* snp_from_io(const SnpUnphasedIO& I, int64_t row_lo, int64_t row_hi, int n_threads) {
    I.need_read();
    if (I.rows < 1 || I.snps < 1) throw core_error("matrix must have at least one row and one column.");
    if (row_hi < 0) row_hi = (int64_t)I.rows;
    if (row_lo < 0 || row_lo >= row_hi || row_hi > (int64_t)I.rows) throw core_error("snp_unphased: invalid row range.");
    const int64_t p = (int64_t)I.snps;
    auto M = std::unique_ptr<DenseMatrix<T>>(new DenseMatrix<T>(row_hi - row_lo, p, typename DenseMatrix<T>::SnpTag{}));
    M->n_threads = n_threads;
    std::vector<T> imp(p);
    for (int64_t j = 0; j < p; ++j) imp[j] = (T)I.impute[j];
    M->snp_impute.upload(imp.data(), p);
    DevBuf<uint64_t> d_outer(p + 1); d_outer.upload(I.outer.data(), p + 1);
    DevBuf<int> d_err(1);
    // ship the chunk lists in column ranges of at most ~1 GB of file bytes and unpack them on the device
    const uint64_t kMaxBytes = 1ull << 30;
    DevBuf<uint8_t> d_file;
    for (int64_t j0 = 0; j0 < p;) {
        int64_t j1 = j0 + 1;
        while (j1 < p && I.outer[j1 + 1] - I.outer[j0] <= kMaxBytes) ++j1;
        if (I.outer[j1] < I.outer[j0] || I.outer[j1] > I.buf_bytes) throw core_error("snp_unphased: malformed file (column offsets).");   // (read() checked this already)
        const uint64_t bytes = I.outer[j1] - I.outer[j0];
        if (d_file.n < bytes + 16) d_file.alloc(bytes + 16);
        AB_CUDA(cudaMemcpyAsync(d_file.p, I.buf + I.outer[j0], bytes, cudaMemcpyHostToDevice, 0));
        const int64_t items = (j1 - j0) * 3;
        snpdat_unpack_kernel<<<(unsigned)((items + 7) / 8), 256>>>(d_file.p, d_outer.p, j0, j1 - j0, (int64_t)I.outer[j0], (int64_t)I.rows, row_lo, row_hi,
                                                                  M->snp_packed.p, M->snp_ldw, d_err.p);
        AB_CUDA(cudaGetLastError());
        AB_CUDA(cudaStreamSynchronize(0));
        j0 = j1;
    }
    int err = 0; d_err.download(&err, 1); AB_CUDA(cudaStreamSynchronize(0));
    if (err == 2) throw core_error("snp_unphased: malformed file (a chunk list runs past the end of its column).");
    if (err) throw core_error("snp_unphased: the file holds a row index outside [0, rows).");
    return M.release();
}

template <class T>
static DenseMatrix<T>* snp_from_calldata(const int8_t* calldata, int64_t n, int64_t p, const double* impute, int n_threads) {
    auto M = std::unique_ptr<DenseMatrix<T>>(new DenseMatrix<T>(n, p, typename DenseMatrix<T>::SnpTag{}));
    M->n_threads = n_threads;
    std::vector<T> imp(p);
    for (int64_t j = 0; j < p; ++j) imp[j] = (T)impute[j];
    M->snp_impute.upload(imp.data(), p);
    DevBuf<int> d_err(1);
    const int64_t cols_per = std::max<int64_t>(1, (int64_t)(256 << 20) / n);
    DevBuf<int8_t> d_call((size_t)std::min(cols_per, p) * n);
    for (int64_t j0 = 0; j0 < p; j0 += cols_per) {
        const int64_t jc = std::min(cols_per, p - j0);
        AB_CUDA(cudaMemcpyAsync(d_call.p, calldata + j0 * n, (size_t)jc * n, cudaMemcpyHostToDevice, 0));
        snp_pack_kernel<<<dim3((unsigned)jc, (unsigned)std::min<int64_t>(64, (M->snp_ldw + 255) / 256)), 256>>>(d_call.p, n, jc, M->snp_packed.p + j0 * M->snp_ldw, M->snp_ldw, d_err.p);
        AB_CUDA(cudaGetLastError());
        AB_CUDA(cudaStreamSynchronize(0));
    }
    int err = 0; d_err.download(&err, 1); AB_CUDA(cudaStreamSynchronize(0));
    if (err) throw core_error("Detected a value greater than > 2. Make sure calldata only contains values <= 2. ");
    return M.release();
}

template <class T>
static DenseMatrix<T>* snp_random(int64_t n, int64_t p, uint64_t seed, int64_t row_offset, int64_t n_total, double one_ratio, double two_ratio, double missing_ratio) {
    auto M = std::unique_ptr<DenseMatrix<T>>(new DenseMatrix<T>(n, p, typename DenseMatrix<T>::SnpTag{}));
    DevBuf<unsigned long long> d_counts((size_t)3 * p);
    snp_fill_random_kernel<<<dim3((unsigned)p, (unsigned)std::min<int64_t>(32, (M->snp_ldw + 255) / 256)), 256>>>(
        M->snp_packed.p, M->snp_ldw, n, p, seed, row_offset, (float)one_ratio, (float)two_ratio, (float)missing_ratio, d_counts.p);
    AB_CUDA(cudaGetLastError());
    std::vector<unsigned long long> cnt((size_t)3 * p);
    d_counts.download(cnt.data(), cnt.size()); AB_CUDA(cudaStreamSynchronize(0));
    std::vector<double> c(cnt.begin(), cnt.end());
    if (DistContext::get().active()) DistContext::get().allreduce_host(c.data(), (int64_t)c.size());      // column means are over ALL rows
    std::vector<T> imp(p);
    for (int64_t j = 0; j < p; ++j) imp[j] = (T)((c[3 * j] + 2.0 * c[3 * j + 1]) / std::max(1.0, (double)n_total - c[3 * j + 2]));
    M->snp_impute.upload(imp.data(), p); AB_CUDA(cudaStreamSynchronize(0));
    return M.release();
}

extern "C" {

int ab_matrix_snp_unphased_create(int dtype, const ab_io_snp* io, int64_t row_lo, int64_t row_hi, int n_threads, ab_matrix** out) {
    AB_TRY
    if (n_threads < 1) throw core_error("n_threads must be >= 1.");
    auto* m = new ab_matrix{dtype};
    try {
        if (dtype == AB_F32) m->f32 = snp_from_io<float>(io->io, row_lo, row_hi, n_threads); else m->f64 = snp_from_io<double>(io->io, row_lo, row_hi, n_threads);
    } catch (...) { delete m; throw; }
    *out = m;
    AB_CATCH
}
int ab_matrix_snp_unphased_from_calldata(int dtype, const int8_t* calldata, int64_t n, int64_t p, const double* impute, int n_threads, ab_matrix** out) {
    AB_TRY
    if (n < 1 || p < 1) throw core_error("matrix must have at least one row and one column.");
    if (n_threads < 1) throw core_error("n_threads must be >= 1.");
    auto* m = new ab_matrix{dtype};
    try {
        if (dtype == AB_F32) m->f32 = snp_from_calldata<float>(calldata, n, p, impute, n_threads); else m->f64 = snp_from_calldata<double>(calldata, n, p, impute, n_threads);
    } catch (...) { delete m; throw; }
    *out = m;
    AB_CATCH
}
int ab_matrix_snp_unphased_alloc_random(int dtype, int64_t n, int64_t p, uint64_t seed, int64_t row_offset, int64_t n_total,
                                        double one_ratio, double two_ratio, double missing_ratio, ab_matrix** out) {
    AB_TRY
    if (n < 1 || p < 1 || n_total < n) throw core_error("snp_unphased: invalid shape.");
    auto* m = new ab_matrix{dtype};
    try {
        if (dtype == AB_F32) m->f32 = snp_random<float>(n, p, seed, row_offset, n_total, one_ratio, two_ratio, missing_ratio);
        else m->f64 = snp_random<double>(n, p, seed, row_offset, n_total, one_ratio, two_ratio, missing_ratio);
    } catch (...) { delete m; throw; }
    *out = m;
    AB_CATCH
}
// calldata_out: column-major (n, p) int8 with -9 for missing; impute_out (p,) doubles
int ab_matrix_snp_unphased_download(ab_matrix* m, int8_t* calldata_out, double* impute_out) {
    AB_TRY
    auto run = [&](auto& M) {
        using T = std::remove_reference_t<decltype(M.snp_impute.p[0])>;
        if (!M.snp) throw core_error("not a snp_unphased matrix.");
        std::vector<uint32_t> h((size_t)M.snp_ldw * M.p); std::vector<T> imp(M.p);
        AB_CUDA(cudaMemcpyAsync(h.data(), M.snp_bits, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, 0)); M.snp_impute.download(imp.data(), M.p);
        AB_CUDA(cudaStreamSynchronize(0));
        for (int64_t j = 0; j < M.p; ++j) {
            impute_out[j] = (double)imp[j];
            for (int64_t i = 0; i < M.n; ++i) {
                const uint32_t code = (h[(size_t)j * M.snp_ldw + (i >> 4)] >> (2 * (i & 15))) & 3u;
                calldata_out[j * M.n + i] = code == 3u ? (int8_t)-9 : (int8_t)code;
            }
        }
    };
    if (m->dtype == AB_F32) run(*m->f32); else run(*m->f64);
    AB_CATCH
}
int ab_matrix_snp_unphased_cache_info(const ab_matrix* m, int64_t* cached_cols, int64_t* packed_bytes) {
    if (m->dtype == AB_F32) { *cached_cols = m->f32->cache_used; *packed_bytes = (int64_t)(m->f32->snp_packed.n * 4); }
    else { *cached_cols = m->f64->cache_used; *packed_bytes = (int64_t)(m->f64->snp_packed.n * 4); }
    return AB_OK;
}
// ------------------------------------------------------------------------------------------ SNP phased ancestry: IO + matrix
// reference: adelie.io.snp_phased_ancestry (PY/io.py:6-111) -> IOSNPPhasedAncestry (CORE/io/io_snp_phased_ancestry.{hpp,ipp});
//            adelie.matrix.snp_phased_ancestry -> MatrixNaiveSNPPhasedAncestry{32,64} (CORE/matrix/matrix_naive_snp_phased_ancestry.ipp)
int ab_io_snp_phased_ancestry_create(const char* filename, const char* read_mode, ab_io_snp_pa** out) {
    AB_TRY
    *out = new ab_io_snp_pa(filename, read_mode);
    AB_CATCH
}
int ab_io_snp_phased_ancestry_free(ab_io_snp_pa* io) { delete io; return AB_OK; }
int ab_io_snp_phased_ancestry_write(ab_io_snp_pa* io, const int8_t* calldata, const int8_t* ancestries, int64_t n, int64_t two_s, int64_t A,
                                    int n_threads, uint64_t* total_bytes) {
    AB_TRY
    (void)n_threads;
    if (A < 0) throw core_error("Number of ancestries A must be >= 0.");
    *total_bytes = io->io.write(calldata, ancestries, (uint64_t)n, (uint64_t)two_s, (uint64_t)A);
    AB_CATCH
}
int ab_io_snp_phased_ancestry_read(ab_io_snp_pa* io, uint64_t* total_bytes) {
    AB_TRY
    *total_bytes = io->io.read();
    AB_CATCH
}
int ab_io_snp_phased_ancestry_info(const ab_io_snp_pa* io, int* is_read, int64_t* rows, int64_t* snps, int64_t* ancestries) {
    *is_read = io->io.is_read ? 1 : 0; *rows = (int64_t)io->io.rows; *snps = (int64_t)io->io.snps; *ancestries = (int64_t)io->io.ancestries;
    return AB_OK;
}
int ab_io_snp_phased_ancestry_get(const ab_io_snp_pa* io, const char* name, void* out) {
    AB_TRY
    io->io.need_read();
    const std::string s(name);
    const auto& I = io->io;
    if (s == "nnz0") std::memcpy(out, I.nnz0.data(), 8 * I.cols);
    else if (s == "nnz1") std::memcpy(out, I.nnz1.data(), 8 * I.cols);
    else if (s == "outer") std::memcpy(out, I.outer.data(), 8 * (I.snps + 1));
    else throw core_error("unknown field " + s);
    AB_CATCH
}
int ab_io_snp_phased_ancestry_to_dense(const ab_io_snp_pa* io, int n_threads, int8_t* out) {
    AB_TRY
    (void)n_threads;
    io->io.to_dense(out);
    AB_CATCH
}
} // extern "C"
template <class T>
static DenseMatrix<T>* snp_pa_from_io(const SnpPhasedAncestryIO& I, int64_t row_lo, int64_t row_hi, int n_threads) {
    I.need_read();
    if (I.rows < 1 || I.cols < 1) throw core_error("matrix must have at least one row and one column.");
    if (row_hi < 0) row_hi = (int64_t)I.rows;
    if (row_lo < 0 || row_lo >= row_hi || row_hi > (int64_t)I.rows) throw core_error("snp_phased_ancestry: invalid row range.");
    const int64_t s_ = (int64_t)I.snps; const int A = (int)I.ancestries;
    auto M = std::unique_ptr<DenseMatrix<T>>(new DenseMatrix<T>(row_hi - row_lo, s_ * A, typename DenseMatrix<T>::SnpTag{}));   // impute stays 0: code 3 never occurs
    M->n_threads = n_threads;
    DevBuf<uint64_t> d_outer(s_ + 1); d_outer.upload(I.outer.data(), s_ + 1);
    DevBuf<int> d_err(1);
    const uint64_t kMaxBytes = 1ull << 30;
    DevBuf<uint8_t> d_file;
    for (int64_t j0 = 0; j0 < s_;) {
        int64_t j1 = j0 + 1;
        while (j1 < s_ && I.outer[j1 + 1] - I.outer[j0] <= kMaxBytes) ++j1;
        if (I.outer[j1] < I.outer[j0] || I.outer[j1] > I.buf_bytes) throw core_error("snp_unphased: malformed file (column offsets).");   // (read() checked this already)
        const uint64_t bytes = I.outer[j1] - I.outer[j0];
        if (d_file.n < bytes + 16) d_file.alloc(bytes + 16);
        AB_CUDA(cudaMemcpyAsync(d_file.p, I.buf + I.outer[j0], bytes, cudaMemcpyHostToDevice, 0));
        const int64_t items = (j1 - j0) * A * 2;
        snpdat_phased_unpack_kernel<<<(unsigned)((items + 7) / 8), 256>>>(d_file.p, d_outer.p, j0, j1 - j0, (int64_t)I.outer[j0], A, (int64_t)I.rows, row_lo, row_hi,
                                                                         M->snp_packed.p, M->snp_ldw, d_err.p);
        AB_CUDA(cudaGetLastError());
        AB_CUDA(cudaStreamSynchronize(0));
        j0 = j1;
    }
    int err = 0; d_err.download(&err, 1); AB_CUDA(cudaStreamSynchronize(0));
    if (err == 2) throw core_error("snp_phased_ancestry: malformed file (a chunk list runs past the end of its SNP).");
    if (err) throw core_error("snp_phased_ancestry: the file holds a row index outside [0, rows).");
    return M.release();
}
extern "C" {
int ab_matrix_snp_phased_ancestry_create(int dtype, const ab_io_snp_pa* io, int64_t row_lo, int64_t row_hi, int n_threads, ab_matrix** out) {
    AB_TRY
    if (n_threads < 1) throw core_error("n_threads must be >= 1.");
    auto* m = new ab_matrix{dtype};
    try {
        if (dtype == AB_F32) m->f32 = snp_pa_from_io<float>(io->io, row_lo, row_hi, n_threads); else m->f64 = snp_pa_from_io<double>(io->io, row_lo, row_hi, n_threads);
    } catch (...) { delete m; throw; }
    *out = m;
    AB_CATCH
}
// ------------------------------------------------------------------------------------------ standardize / subset
// reference: adelie.matrix.standardize (PY/matrix.py:1414-1536; MatrixNaiveStandardize, matrix_naive_standardize.ipp:8-293) and
// adelie.matrix.subset (PY/matrix.py:1539-1632; MatrixNaiveCSubset / MatrixNaiveRSubset, matrix_naive_subset.ipp).  The reference wraps the
// base matrix and applies the affine map / index map inside every operator; here the transformed matrix is MATERIALISED once as a new dense
// device matrix (one pass at HBM speed, 180 GB of HBM) so that the fused sweep reads plain TMA tiles.  Dense base matrices only.
} // extern "C"
template <class T>
static DenseMatrix<T>* make_standardized(DenseMatrix<T>& B, const T* centers, const T* scales, int n_threads) {
    if (B.sparse) throw core_error("standardize: sparse base matrices are not supported on the device.");
    if (B.snp) {
        // snp_unphased: nothing is materialised -- the view shares the packed genotypes and maps the four codes of column j to
        // (0 - c_j) / s_j, (1 - c_j) / s_j, (2 - c_j) / s_j, (impute_j - c_j) / s_j inside the decode / packed-GEMV kernels
        if (B.snp_center.n) throw core_error("standardize: the base matrix is already a standardized snp_unphased view.");
        auto V = new DenseMatrix<T>(B, centers, scales, typename DenseMatrix<T>::SnpTag{});
        V->n_threads = n_threads;
        return V;
    }
    auto M = std::unique_ptr<DenseMatrix<T>>(new DenseMatrix<T>(B.n, B.p));
    M->n_threads = n_threads;
    DevBuf<T> dc(B.p), ds(B.p); dc.upload(centers, B.p); ds.upload(scales, B.p);
    standardize_cols_kernel<T><<<dim3((unsigned)B.p, (unsigned)std::min<int64_t>(64, (B.n + 255) / 256)), 256>>>(B.X, B.ld, B.n, dc.p, ds.p, M->X, M->ld);
    AB_CUDA(cudaGetLastError()); AB_CUDA(cudaStreamSynchronize(0));
    return M.release();
}
template <class T>
static DenseMatrix<T>* make_subset(DenseMatrix<T>& B, const int64_t* idx, int64_t m, int axis, int n_threads) {
    if (B.sparse || B.snp) throw core_error("subset: only dense base matrices are supported on the device.");
    if (m <= 0) throw core_error("subset must be non-empty.");
    const int64_t lim = axis == 0 ? B.n : B.p;
    std::vector<uint8_t> seen(lim, 0);
    for (int64_t k = 0; k < m; ++k) {
        if (idx[k] < 0 || idx[k] >= lim || seen[idx[k]])
            throw core_error(std::string("subset must contain unique values in the range [0, ") + (axis == 0 ? "n" : "p") + ") where mat is (n, p).");
        seen[idx[k]] = 1;
    }
    auto M = std::unique_ptr<DenseMatrix<T>>(new DenseMatrix<T>(axis == 0 ? m : B.n, axis == 0 ? B.p : m));
    M->n_threads = n_threads;
    DevBuf<int64_t> di(m); di.upload(idx, m);
    gather_kernel<T><<<dim3((unsigned)M->p, (unsigned)std::min<int64_t>(64, (M->n + 255) / 256)), 256>>>(
        B.X, B.ld, axis == 0 ? di.p : nullptr, axis == 0 ? nullptr : di.p, M->n, M->X, M->ld);
    AB_CUDA(cudaGetLastError()); AB_CUDA(cudaStreamSynchronize(0));
    return M.release();
}
extern "C" {
int ab_matrix_standardize_create(ab_matrix* base, const void* centers, int64_t n_centers, const void* scales, int64_t n_scales, int n_threads, ab_matrix** out) {
    AB_TRY
    int64_t p = 0; ab_matrix_cols(base, &p);
    if (n_centers != p) throw core_error("centers must be (p,) where mat is (n, p).");
    if (n_scales != p) throw core_error("scales must be (p,) where mat is (n, p).");
    if (n_threads < 1) throw core_error("n_threads must be >= 1.");
    auto* m = new ab_matrix{base->dtype};
    try {
        if (base->dtype == AB_F32) m->f32 = make_standardized<float>(*base->f32, (const float*)centers, (const float*)scales, n_threads);
        else m->f64 = make_standardized<double>(*base->f64, (const double*)centers, (const double*)scales, n_threads);
    } catch (...) { delete m; throw; }
    *out = m;
    AB_CATCH
}
int ab_matrix_subset_create(ab_matrix* base, const int64_t* indices, int64_t m_idx, int axis, int n_threads, ab_matrix** out) {
    AB_TRY
    if (axis != 0 && axis != 1) throw core_error("axis must be 0 or 1.");
    if (n_threads < 1) throw core_error("n_threads must be >= 1.");
    auto* m = new ab_matrix{base->dtype};
    try {
        if (base->dtype == AB_F32) m->f32 = make_subset<float>(*base->f32, indices, m_idx, axis, n_threads);
        else m->f64 = make_subset<double>(*base->f64, indices, m_idx, axis, n_threads);
    } catch (...) { delete m; throw; }
    *out = m;
    AB_CATCH
}
int ab_matrix_free(ab_matrix* m) { if (m) { delete m->f32; delete m->f64; delete m; } return AB_OK; }
int ab_matrix_rows(const ab_matrix* m, int64_t* out) { *out = m->dtype == AB_F32 ? m->f32->n : m->f64->n; return AB_OK; }
int ab_matrix_cols(const ab_matrix* m, int64_t* out) { *out = m->dtype == AB_F32 ? m->f32->p : m->f64->p; return AB_OK; }

} // extern "C"

// shape checks with the reference's messages (matrix_naive_base.hpp:148-271)
static void check_cmul(int64_t j, int64_t r, int64_t c) {
    if (j < 0 || j >= c) throw core_error("cmul() is given inconsistent inputs! (j=" + std::to_string(j) + ", r=" + std::to_string(r) + ", c=" + std::to_string(c) + ")");
}
static void check_bmul(int64_t j, int64_t q, int64_t r, int64_t c) {
    if (j < 0 || q < 0 || j + q > c) throw core_error("bmul() is given inconsistent inputs! (j=" + std::to_string(j) + ", q=" + std::to_string(q) + ", r=" + std::to_string(r) + ", c=" + std::to_string(c) + ")");
}

template <class T>
struct HostOps {
    // stage host vectors into padded device buffers
    static DevBuf<T> up(const T* h, int64_t n, int64_t np) { DevBuf<T> d(np); d.upload(h, n); return d; }

    static double cmul(DenseMatrix<T>& M, int64_t j, const T* v, const T* w) {
        check_cmul(j, M.n, M.p);
        auto dv = up(v, M.n, M.ld), dw = up(w, M.n, M.ld); DevBuf<T> o(1);
        M.d_gemv_t(j, nullptr, 1, dv.p, dw.p, o.p);
        T r; o.download(&r, 1); AB_CUDA(cudaStreamSynchronize(0));
        return (double)r;
    }
    static void bmul(DenseMatrix<T>& M, int64_t j, int64_t q, const T* v, const T* w, T* out) {
        check_bmul(j, q, M.n, M.p);
        if (q == 0) return;
        auto dv = up(v, M.n, M.ld), dw = up(w, M.n, M.ld); DevBuf<T> o(q);
        M.d_gemv_t(j, nullptr, (int)q, dv.p, dw.p, o.p);
        o.download(out, q); AB_CUDA(cudaStreamSynchronize(0));
    }
    static void btmul(DenseMatrix<T>& M, int64_t j, int64_t q, const T* v, T* out) {
        check_bmul(j, q, M.n, M.p);
        if (q == 0) return;
        auto dv = up(v, q, q + 4), dout = up(out, M.n, M.ld);
        M.d_btmul(j, (int)q, dv.p, dout.p);
        dout.download(out, M.n); AB_CUDA(cudaStreamSynchronize(0));
    }
    static void mul(DenseMatrix<T>& M, const T* v, const T* w, T* out) {
        auto dv = up(v, M.n, M.ld), dw = up(w, M.n, M.ld); DevBuf<T> o(M.p);
        M.d_mul(dv.p, dw.p, o.p);
        o.download(out, M.p); AB_CUDA(cudaStreamSynchronize(0));
    }
    static void sq_mul(DenseMatrix<T>& M, const T* w, T* out) {
        auto dw = up(w, M.n, M.ld); DevBuf<T> o(M.p);
        M.d_gemv_t(0, nullptr, (int)M.p, dw.p, dw.p, o.p, true);
        o.download(out, M.p); AB_CUDA(cudaStreamSynchronize(0));
    }
    static void cov(DenseMatrix<T>& M, int64_t j, int64_t q, const T* sqrt_w, T* out) {
        check_bmul(j, q, M.n, M.p);
        if (q == 0) return;
        auto dw = up(sqrt_w, M.n, M.ld);
        CovItem it{M.phys_col(j, (int)q), (int32_t)q, 0};
        DevBuf<CovItem> di(1); di.upload(&it, 1);
        DevBuf<double> C((size_t)q * q);
        M.d_cov(di.p, 1, q * q, dw.p, true, C.p, 1, (int)q);
        std::vector<double> h((size_t)q * q);
        C.download(h.data(), h.size()); AB_CUDA(cudaStreamSynchronize(0));
        for (int64_t a = 0; a < q; ++a) for (int64_t b = 0; b < q; ++b) out[a + b * q] = (T)h[a * q + b];
    }
    static void sp_tmul(DenseMatrix<T>& M, int64_t L, const int64_t* indptr, const int64_t* indices, const T* values, T* out) {
        // out[l, :] = sum_k values[k] * X[:, indices[k]]  -- one axpy launch per non-zero column
        DevBuf<T> drow(M.ld), dv(4);
        for (int64_t l = 0; l < L; ++l) {
            AB_CUDA(cudaMemsetAsync(drow.p, 0, M.ld * sizeof(T), 0));
            for (int64_t k = indptr[l]; k < indptr[l + 1]; ++k) {
                if (indices[k] < 0 || indices[k] >= M.p) throw core_error("sp_tmul() is given inconsistent inputs!");
                dv.upload(values + k, 1);
                M.d_btmul(indices[k], 1, dv.p, drow.p);
            }
            drow.download(out + l * M.n, M.n);
            AB_CUDA(cudaStreamSynchronize(0));
        }
    }
};

extern "C" {

int ab_matrix_cmul(ab_matrix* m, int64_t j, const void* v, const void* w, double* out) {
    AB_TRY
    *out = m->dtype == AB_F32 ? HostOps<float>::cmul(*m->f32, j, (const float*)v, (const float*)w)
                              : HostOps<double>::cmul(*m->f64, j, (const double*)v, (const double*)w);
    AB_CATCH
}
int ab_matrix_ctmul(ab_matrix* m, int64_t j, double v, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) { check_cmul(j, m->f32->n, m->f32->p); float vv = (float)v; HostOps<float>::btmul(*m->f32, j, 1, &vv, (float*)out); }
    else { check_cmul(j, m->f64->n, m->f64->p); HostOps<double>::btmul(*m->f64, j, 1, &v, (double*)out); }
    AB_CATCH
}
int ab_matrix_bmul(ab_matrix* m, int64_t j, int64_t q, const void* v, const void* w, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) HostOps<float>::bmul(*m->f32, j, q, (const float*)v, (const float*)w, (float*)out);
    else HostOps<double>::bmul(*m->f64, j, q, (const double*)v, (const double*)w, (double*)out);
    AB_CATCH
}
int ab_matrix_btmul(ab_matrix* m, int64_t j, int64_t q, const void* v, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) HostOps<float>::btmul(*m->f32, j, q, (const float*)v, (float*)out);
    else HostOps<double>::btmul(*m->f64, j, q, (const double*)v, (double*)out);
    AB_CATCH
}
int ab_matrix_mul(ab_matrix* m, const void* v, const void* w, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) HostOps<float>::mul(*m->f32, (const float*)v, (const float*)w, (float*)out);
    else HostOps<double>::mul(*m->f64, (const double*)v, (const double*)w, (double*)out);
    AB_CATCH
}
int ab_matrix_mul_multi(ab_matrix* m, int64_t K, const void* v, const void* w, void* out) {
    AB_TRY
    if (K < 1 || K > kMultiMaxK) throw core_error("mul_multi(): the number of classes must be in [1, 16].");
    auto run = [&](auto& M, auto* vv, auto* ww, auto* oo) {
        using T = std::remove_cv_t<std::remove_pointer_t<decltype(oo)>>;
        if (M.sparse) throw core_error("multi-response problems are not supported on sparse matrices.");
        const int64_t nk = M.n * K, ldk = M.ld * K;
        DevBuf<T> dv(ldk), dw(ldk), o(M.p * K);                         // (n_pad, K) row-major, pad rows zero
        dv.upload(vv, nk); dw.upload(ww, nk);
        M.d_mul_multi((int)K, 0, dv.p, dw.p, o.p);
        M.check_tc_error();
        o.download(oo, M.p * K); AB_CUDA(cudaStreamSynchronize(0));
    };
    if (m->dtype == AB_F32) run(*m->f32, (const float*)v, (const float*)w, (float*)out);
    else run(*m->f64, (const double*)v, (const double*)w, (double*)out);
    AB_CATCH
}
int ab_matrix_window_gram(ab_matrix* m, const int32_t* cols, int ncol, int n_src, const void* w, int use_tc, double* out) {
    AB_TRY
    if (m->dtype != AB_F32 || m->f32->sparse) throw core_error("window_gram() needs a dense float32 matrix.");
    if (ncol < 1 || ncol > 128 || n_src < 1 || n_src > 64 || n_src > ncol) throw core_error("window_gram() is given inconsistent inputs!");
    auto& X = *m->f32;
    PanelItem it{}; it.q_off = 0; it.ncol = ncol; it.n_src = n_src;
    for (int c = 0; c < ncol; ++c) { if (cols[c] < 0 || cols[c] >= X.p) throw core_error("window_gram(): column out of range."); it.cols[c] = X.phys_col(cols[c], 1); }
    DevBuf<PanelItem> d_it(1); d_it.upload(&it, 1);
    DevBuf<float> d_w(X.ld); d_w.upload((const float*)w, X.n);
    std::vector<double> tmp(kPanelOut);
    X.d_panel_gram(d_it.p, 1, d_w.p, nullptr, 0, 0, use_tc, tmp.data());
    X.check_tc_error();
    for (int s_ = 0; s_ < n_src; ++s_) for (int u = 0; u < ncol; ++u) out[(size_t)s_ * ncol + u] = tmp[(size_t)s_ * 128 + u];
    AB_CATCH
}
int ab_matrix_cov(ab_matrix* m, int64_t j, int64_t q, const void* sqrt_w, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) HostOps<float>::cov(*m->f32, j, q, (const float*)sqrt_w, (float*)out);
    else HostOps<double>::cov(*m->f64, j, q, (const double*)sqrt_w, (double*)out);
    AB_CATCH
}
int ab_matrix_sq_mul(ab_matrix* m, const void* w, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) HostOps<float>::sq_mul(*m->f32, (const float*)w, (float*)out);
    else HostOps<double>::sq_mul(*m->f64, (const double*)w, (double*)out);
    AB_CATCH
}
int ab_matrix_sp_tmul(ab_matrix* m, int64_t L, const int64_t* indptr, const int64_t* indices, const void* values, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) HostOps<float>::sp_tmul(*m->f32, L, indptr, indices, (const float*)values, (float*)out);
    else HostOps<double>::sp_tmul(*m->f64, L, indptr, indices, (const double*)values, (double*)out);
    AB_CATCH
}

} // extern "C"

// ------------------------------------------------------------------------------------------ glm
template <class T>
static Glm<T>* make_glm(int family, int64_t n, int64_t K, const void* y, const void* w,
                        const void* cox_start = nullptr, const void* cox_stop = nullptr, const int64_t* cox_strata = nullptr, int cox_tie_efron = 1) {
    switch (family) {
        case AB_GLM_COX: {
            if (!cox_start) throw core_error("start must be (n,) where status is (n,).");
            if (!cox_stop) throw core_error("stop must be (n,) where status is (n,).");
            if (!cox_strata) throw core_error("strata must be (n,) where status is (n,).");
            return new GlmCox<T>((const T*)cox_start, (const T*)cox_stop, (const T*)y, cox_strata, (const T*)w, n, cox_tie_efron != 0);
        }
        case AB_GLM_GAUSSIAN: return new GlmGaussian<T>((const T*)y, (const T*)w, n);
        case AB_GLM_BINOMIAL_LOGIT: return new GlmBinomialLogit<T>((const T*)y, (const T*)w, n);
        case AB_GLM_MULTIGAUSSIAN: return new GlmMultiGaussian<T>((const T*)y, (const T*)w, n, K);
        case AB_GLM_MULTINOMIAL: return new GlmMultinomial<T>((const T*)y, (const T*)w, n, K);
        case AB_GLM_POISSON: return new GlmPoisson<T>((const T*)y, (const T*)w, n);
        case AB_GLM_BINOMIAL_PROBIT: return new GlmBinomialProbit<T>((const T*)y, (const T*)w, n);
    }
    throw core_error("unsupported GLM family.");
}

// User-defined GLM behind C callbacks (PyGlmBase / PyGlmMultiBase trampolines, py_glm.cpp:8-92, 240-330): device vectors are staged
// through pinned host memory around every call.
template <class T>
struct GlmCallback : Glm<T> {
    using B = Glm<T>;
    ab_glm_callbacks cb;
    PinnedBuf<T> h_eta, h_grad, h_hess, h_out;
    GlmCallback(int64_t n_rows, int64_t K_, bool multi, const ab_glm_callbacks& c) : cb(c) {
        B::name = "callback"; B::is_multi = multi; B::K = multi ? K_ : 1; B::n = multi ? n_rows * K_ : n_rows;
        if (!c.gradient || !c.hessian || !c.loss || !c.loss_full) throw core_error("user-defined GLM: gradient, hessian, loss and loss_full callbacks are required.");
        if (DistContext::get().active()) throw core_error("user-defined GLMs are not supported in row-sharded multi-GPU mode.");
        h_eta.alloc(B::n); h_grad.alloc(B::n); h_hess.alloc(B::n); h_out.alloc(B::n);
    }
    void down(const T* d, PinnedBuf<T>& h) { AB_CUDA(cudaMemcpyAsync(h.p, d, B::n * sizeof(T), cudaMemcpyDeviceToHost, 0)); }
    void up(const PinnedBuf<T>& h, T* d) { AB_CUDA(cudaMemcpyAsync(d, h.p, B::n * sizeof(T), cudaMemcpyHostToDevice, 0)); AB_CUDA(cudaStreamSynchronize(0)); }
    static void chk(int rc, const char* what) { if (rc) throw solver_error(std::string("user-defined GLM: ") + what + "() raised an exception."); }
    void gradient(const T* eta, T* grad) override {
        down(eta, h_eta); AB_CUDA(cudaStreamSynchronize(0));
        chk(cb.gradient(cb.ctx, h_eta.p, h_out.p), "gradient");
        up(h_out, grad);
    }
    void hessian(const T* eta, const T* grad, T* hess) override {
        down(eta, h_eta); down(grad, h_grad); AB_CUDA(cudaStreamSynchronize(0));
        chk(cb.hessian(cb.ctx, h_eta.p, h_grad.p, h_out.p), "hessian");
        up(h_out, hess);
    }
    void inv_hessian_gradient(const T* eta, const T* grad, const T* hess, T* out) override {
        if (!cb.inv_hessian_gradient) { B::inv_hessian_gradient(eta, grad, hess, out); return; }
        down(eta, h_eta); down(grad, h_grad); down(hess, h_hess); AB_CUDA(cudaStreamSynchronize(0));
        chk(cb.inv_hessian_gradient(cb.ctx, h_eta.p, h_grad.p, h_hess.p, h_out.p), "inv_hessian_gradient");
        up(h_out, out);
    }
    T loss(const T* eta) override {
        down(eta, h_eta); AB_CUDA(cudaStreamSynchronize(0));
        double out = 0;
        chk(cb.loss(cb.ctx, h_eta.p, &out), "loss");
        return (T)out;
    }
    T loss_full() override { double out = 0; chk(cb.loss_full(cb.ctx, &out), "loss_full"); return (T)out; }
    void inv_link(const T* eta, T* out) override {
        if (!cb.inv_link) throw core_error("user-defined GLM: inv_link() is not implemented.");
        down(eta, h_eta); AB_CUDA(cudaStreamSynchronize(0));
        chk(cb.inv_link(cb.ctx, h_eta.p, h_out.p), "inv_link");
        up(h_out, out);
    }
};

template <class T>
struct GlmHost {
    static int64_t padn(Glm<T>& g) { return g.is_multi ? pad_rows(g.n / g.K) * g.K : pad_rows(g.n); }
    static DevBuf<T> up(Glm<T>& g, const void* h) { DevBuf<T> d(padn(g)); d.upload((const T*)h, g.n); return d; }
    static void down(Glm<T>& g, DevBuf<T>& d, void* h) { d.download((T*)h, g.n); AB_CUDA(cudaStreamSynchronize(0)); }
};

extern "C" {

int ab_glm_create(int dtype, int family, int64_t n, int64_t K, const void* y, const void* weights,
                  const void* cox_start, const void* cox_stop, const int64_t* cox_strata, int cox_tie_efron, ab_glm** out) {
    AB_TRY
    auto* g = new ab_glm{dtype, family};
    try {
        if (dtype == AB_F32) g->f32 = make_glm<float>(family, n, K, y, weights, cox_start, cox_stop, cox_strata, cox_tie_efron);
        else g->f64 = make_glm<double>(family, n, K, y, weights, cox_start, cox_stop, cox_strata, cox_tie_efron);
    } catch (...) { delete g; throw; }
    *out = g;
    AB_CATCH
}
int ab_glm_create_callback(int dtype, int64_t n, int64_t K, int is_multi, const ab_glm_callbacks* callbacks, ab_glm** out) {
    AB_TRY
    if (!callbacks) throw core_error("user-defined GLM: callbacks must not be NULL.");
    if (n < 1 || (is_multi && K < 1)) throw core_error("user-defined GLM: invalid shape.");
    auto* g = new ab_glm{dtype, 0};
    try {
        if (dtype == AB_F32) g->f32 = new GlmCallback<float>(n, K, is_multi != 0, *callbacks);
        else g->f64 = new GlmCallback<double>(n, K, is_multi != 0, *callbacks);
    } catch (...) { delete g; throw; }
    *out = g;
    AB_CATCH
}
int ab_glm_free(ab_glm* g) { if (g) { delete g->f32; delete g->f64; delete g; } return AB_OK; }

#define GLM_CALL(BODY32, BODY64) AB_TRY if (g->dtype == AB_F32) { auto& G = *g->f32; using T = float; BODY32 } else { auto& G = *g->f64; using T = double; BODY64 } AB_CATCH
#define GLM_BOTH(BODY) GLM_CALL(BODY, BODY)

int ab_glm_gradient(ab_glm* g, const void* eta, void* grad) {
    GLM_BOTH({ auto e = GlmHost<T>::up(G, eta); DevBuf<T> o(GlmHost<T>::padn(G)); G.gradient(e.p, o.p); GlmHost<T>::down(G, o, grad); })
}
int ab_glm_hessian(ab_glm* g, const void* eta, const void* grad, void* hess) {
    GLM_BOTH({ auto e = GlmHost<T>::up(G, eta); auto gr = GlmHost<T>::up(G, grad); DevBuf<T> o(GlmHost<T>::padn(G)); G.hessian(e.p, gr.p, o.p); GlmHost<T>::down(G, o, hess); })
}
int ab_glm_inv_hessian_gradient(ab_glm* g, const void* eta, const void* grad, const void* hess, void* out) {
    GLM_BOTH({ auto e = GlmHost<T>::up(G, eta); auto gr = GlmHost<T>::up(G, grad); auto h = GlmHost<T>::up(G, hess); DevBuf<T> o(GlmHost<T>::padn(G));
               G.inv_hessian_gradient(e.p, gr.p, h.p, o.p); GlmHost<T>::down(G, o, out); })
}
int ab_glm_loss(ab_glm* g, const void* eta, double* out) {
    GLM_BOTH({ auto e = GlmHost<T>::up(G, eta); *out = (double)G.loss(e.p); })
}
int ab_glm_loss_full(ab_glm* g, double* out) {
    GLM_BOTH({ *out = (double)G.loss_full(); })
}
int ab_glm_inv_link(ab_glm* g, const void* eta, void* out) {
    GLM_BOTH({ auto e = GlmHost<T>::up(G, eta); DevBuf<T> o(GlmHost<T>::padn(G)); G.inv_link(e.p, o.p); GlmHost<T>::down(G, o, out); })
}

} // extern "C"

// ------------------------------------------------------------------------------------------ state
template <class T>
static PathState<T>* make_state(const ab_state_args* a, DenseMatrix<T>* X, Glm<T>* glm) {
    auto st = std::make_unique<PathState<T>>();
    auto& s = *st;
    s.X = X; s.n = X->n; s.G = a->G;
    s.K = (int)std::max<int64_t>(1, a->n_classes);
    if (s.K > 16) throw core_error("multi-response problems with more than 16 classes are not supported.");
    s.n_int = (s.K > 1 && a->multi_intercept) ? s.K : 0;
    s.p = X->p * s.K + s.n_int;
    const int64_t nK = s.n * s.K;
    if (a->G < 1) throw core_error("groups must be non-empty.");
    s.groups.assign(a->groups, a->groups + a->G);
    s.group_sizes.assign(a->group_sizes, a->group_sizes + a->G);
    s.alpha = (T)a->alpha;
    s.penalty.assign((const T*)a->penalty, (const T*)a->penalty + a->G);
    s.is_glm = glm != nullptr; s.glm = glm;
    s.min_ratio = (T)a->min_ratio; s.lmda_path_size = a->lmda_path_size; s.max_screen_size = a->max_screen_size; s.max_active_size = a->max_active_size;
    s.pivot_subset_ratio = (T)a->pivot_subset_ratio; s.pivot_subset_min = a->pivot_subset_min; s.pivot_slack_ratio = (T)a->pivot_slack_ratio;
    s.screen_rule = a->screen_rule; s.max_iters = a->max_iters; s.tol = (T)a->tol; s.adev_tol = (T)a->adev_tol; s.ddev_tol = (T)a->ddev_tol;
    s.newton_tol = (T)a->newton_tol; s.newton_max_iters = a->newton_max_iters; s.early_exit = a->early_exit;
    s.setup_lmda_max = a->setup_lmda_max; s.setup_lmda_path = a->setup_lmda_path; s.intercept = a->intercept && s.K == 1; s.n_threads = a->n_threads;    // multi-response: intercepts are explicit columns (PY/state.py:2329-2330)
    s.lmda_max = (T)a->lmda_max; s.lmda = (T)a->lmda;
    if (a->lmda_path && a->lmda_path_len > 0) s.lmda_path.assign((const T*)a->lmda_path, (const T*)a->lmda_path + a->lmda_path_len);
    s.screen_set.assign(a->screen_set, a->screen_set + a->screen_set_size);
    s.screen_beta.assign((const T*)a->screen_beta, (const T*)a->screen_beta + a->screen_beta_size);
    s.screen_is_active.assign(a->screen_is_active, a->screen_is_active + a->screen_set_size);
    s.active_set_size = a->active_set_size;
    s.active_set.assign(a->active_set, a->active_set + a->G);
    s.grad.assign((const T*)a->grad, (const T*)a->grad + s.p);
    const int64_t np = X->n_pad() * s.K;
    s.d_resid.alloc(np); s.d_grad.alloc(s.p);
    s.d_resid.upload((const T*)a->resid, nK);
    if (!s.is_glm) {
        // state_gaussian_naive.ipp:9-28 shape checks are implied by the pointer/size contract of the C ABI
        s.d_weights.alloc(np); s.d_weights.upload((const T*)a->weights, nK);
        s.d_resid_prev.alloc(np);
        s.X_means.assign((const T*)a->X_means, (const T*)a->X_means + s.p);
        s.d_X_means.alloc(s.p); s.d_X_means.upload(s.X_means.data(), s.p);
        s.y_mean = (T)a->y_mean; s.y_var = (T)a->y_var; s.resid_sum = (T)a->resid_sum; s.rsq = (T)a->rsq;
    } else {
        if (glm->n != nK) throw core_error("y must be (n,) (or (n, K)) where X is (n, p).");
        s.d_offsets.alloc(np); s.d_offsets.upload((const T*)a->offsets, nK);
        s.d_eta.alloc(np); s.d_eta.upload((const T*)a->eta, nK);
        s.d_eta_prev.alloc(np); s.d_glm_resid_prev.alloc(np); s.d_hess.alloc(np); s.d_irls_w.alloc(np);
        s.d_irls_y.alloc(np); s.d_irls_resid.alloc(np);
        s.beta0 = (T)a->beta0; s.loss_null = (T)a->loss_null; s.loss_full = (T)a->loss_full; s.setup_loss_null = a->setup_loss_null;
        s.irls_max_iters = a->irls_max_iters; s.irls_tol = (T)a->irls_tol;
    }
    AB_CUDA(cudaStreamSynchronize(0));
    s.validate_and_init();
    return st.release();
}

template <class T>
static void get_betas(const PathState<T>& s, int64_t* indptr, int64_t* indices, double* values, int64_t* nnz, int64_t* L) {
    int64_t tot = 0;
    if (indptr) indptr[0] = 0;
    for (size_t l = 0; l < s.betas.size(); ++l) {
        const auto& b = s.betas[l];
        for (size_t k = 0; k < b.idx.size(); ++k) {
            if (indices) indices[tot + k] = b.idx[k];
            if (values) values[tot + k] = b.val[k];
        }
        tot += (int64_t)b.idx.size();
        if (indptr) indptr[l + 1] = tot;
    }
    *nnz = tot; *L = (int64_t)s.betas.size();
}

template <class T, class V>
static int copy_vec(const V& v, double* out, int64_t cap, int64_t* len) {
    *len = (int64_t)v.size();
    if (out) for (int64_t i = 0; i < std::min<int64_t>(cap, *len); ++i) out[i] = (double)v[i];
    return AB_OK;
}
template <class V>
static int copy_ivec(const V& v, int64_t* out, int64_t cap, int64_t* len, int64_t limit = -1) {
    *len = limit >= 0 ? limit : (int64_t)v.size();
    if (out) for (int64_t i = 0; i < std::min<int64_t>(cap, *len); ++i) out[i] = (int64_t)v[i];
    return AB_OK;
}

template <class T>
static int state_vec_f64(const PathState<T>& s, const std::string& nm, double* out, int64_t cap, int64_t* len) {
    if (nm == "lmda_path") return copy_vec<T>(s.lmda_path, out, cap, len);
    if (nm == "screen_beta") return copy_vec<T>(s.screen_beta, out, cap, len);
    if (nm == "grad") return copy_vec<T>(s.grad, out, cap, len);
    if (nm == "abs_grad") return copy_vec<T>(s.abs_grad, out, cap, len);
    if (nm == "devs") return copy_vec<T>(s.devs, out, cap, len);
    if (nm == "lmdas") return copy_vec<T>(s.lmdas, out, cap, len);
    if (nm == "rsqs") return copy_vec<T>(s.rsqs, out, cap, len);
    if (nm == "intercepts") return copy_vec<T>(s.intercepts, out, cap, len);
    if (nm == "X_means") return copy_vec<T>(s.X_means, out, cap, len);
    if (nm == "screen_X_means") return copy_vec<T>(s.screen_X_means, out, cap, len);
    if (nm == "screen_vars") return copy_vec<T>(s.screen_vars, out, cap, len);
    if (nm == "penalty") return copy_vec<T>(s.penalty, out, cap, len);
    if (nm == "benchmark_screen") return copy_vec<double>(s.benchmark_screen, out, cap, len);
    if (nm == "benchmark_fit_screen") return copy_vec<double>(s.benchmark_fit_screen, out, cap, len);
    if (nm == "benchmark_fit_active") return copy_vec<double>(s.benchmark_fit_active, out, cap, len);
    if (nm == "benchmark_kkt") return copy_vec<double>(s.benchmark_kkt, out, cap, len);
    if (nm == "benchmark_invariance") return copy_vec<double>(s.benchmark_invariance, out, cap, len);
    if (nm == "launch_cols") return copy_vec<double>(s.launch_cols, out, cap, len);
    if (nm == "launch_sweeps") return copy_vec<double>(s.launch_sweeps, out, cap, len);
    if (nm == "launch_ms") return copy_vec<double>(s.launch_ms, out, cap, len);
    if (nm == "sweep_stats") {
        const int64_t N = 32 + 8 * 160;
        *len = N;
        if (out) {
            std::vector<long long> h(N, 0);
            if (s.X->stats.n) { s.X->stats.download(h.data(), N); AB_CUDA(cudaStreamSynchronize(0)); }
            for (int64_t i = 0; i < std::min<int64_t>(cap, N); ++i) out[i] = (double)h[i];
        }
        return AB_OK;
    }
    if (nm == "resid" || nm == "eta") {
        const DevBuf<T>& d = (nm == "resid") ? s.d_resid : s.d_eta;
        const int64_t nn = s.n * s.K;
        *len = d.p ? nn : 0;
        if (out && d.p) {
            std::vector<T> h(nn);
            d.download(h.data(), nn); AB_CUDA(cudaStreamSynchronize(0));
            for (int64_t i = 0; i < std::min<int64_t>(cap, nn); ++i) out[i] = (double)h[i];
        }
        return AB_OK;
    }
    g_last_error = "adelie_core: unknown state vector " + nm;
    return AB_ERR_ARG;
}
template <class T>
static int state_vec_i64(const PathState<T>& s, const std::string& nm, int64_t* out, int64_t cap, int64_t* len) {
    if (nm == "groups") return copy_ivec(s.groups, out, cap, len);
    if (nm == "group_sizes") return copy_ivec(s.group_sizes, out, cap, len);
    if (nm == "screen_set") return copy_ivec(s.screen_set, out, cap, len);
    if (nm == "screen_begins") return copy_ivec(s.screen_begins, out, cap, len);
    if (nm == "screen_is_active") return copy_ivec(s.screen_is_active, out, cap, len);
    if (nm == "active_set") return copy_ivec(s.active_set, out, cap, len);
    if (nm == "n_valid_solutions") return copy_ivec(s.n_valid_solutions, out, cap, len);
    if (nm == "active_sizes") return copy_ivec(s.active_sizes, out, cap, len);
    if (nm == "screen_sizes") return copy_ivec(s.screen_sizes, out, cap, len);
    g_last_error = "adelie_core: unknown state vector " + nm;
    return AB_ERR_ARG;
}
template <class T>
static int state_scalar(const PathState<T>& s, const std::string& nm, double* out) {
    if (nm == "lmda_max") *out = s.lmda_max; else if (nm == "lmda") *out = s.lmda;
    else if (nm == "rsq") *out = s.rsq; else if (nm == "resid_sum") *out = s.resid_sum;
    else if (nm == "y_mean") *out = s.y_mean; else if (nm == "y_var") *out = s.y_var;
    else if (nm == "loss_null") *out = s.loss_null; else if (nm == "loss_full") *out = s.loss_full;
    else if (nm == "beta0") *out = s.beta0; else if (nm == "active_set_size") *out = (double)s.active_set_size;
    else if (nm == "alpha") *out = s.alpha; else if (nm == "tol") *out = s.tol;
    else if (nm == "n_sweeps") *out = (double)s.n_sweeps; else if (nm == "n_group_updates") *out = (double)s.n_group_updates;
    else if (nm == "n_col_updates") *out = (double)s.n_col_updates;
    else if (nm == "n_irls") *out = (double)s.n_irls; else if (nm == "n_pin_solves") *out = (double)s.n_pin_solves;
    else if (nm == "n_kernel_launches") *out = (double)s.n_kernel_launches;
    else if (nm == "time_sweep_kernel") *out = s.time_sweep_kernel;
    else if (nm == "sweep_ncta") *out = s.X->last_geom.ncta; else if (nm == "sweep_stages") *out = s.X->last_geom.n_stages;
    else if (nm == "sweep_smem_bytes") *out = (double)s.X->last_geom.smem_bytes; else if (nm == "sweep_staged") *out = s.X->last_geom.smem ? 1 : 0;
    else if (nm == "sweep_threads") *out = s.X->last_geom.threads;
    else if (nm == "sweep_batch") *out = s.X->last_bgeom.ok ? s.X->last_bgeom.B : 1;
    else if (nm == "n_panels_built") *out = (double)s.n_panels_built; else if (nm == "n_batched_launches") *out = (double)s.n_batched_launches;
    else if (nm == "setup_lmda_max") *out = s.setup_lmda_max; else if (nm == "setup_lmda_path") *out = s.setup_lmda_path;
    else if (nm.rfind("t_", 0) == 0) {
        *out = 0;
        for (const auto& kv : s.timers.acc) if (kv.first == nm.substr(2)) *out = kv.second;
    }
    else { g_last_error = "adelie_core: unknown state scalar " + nm; return AB_ERR_ARG; }
    return AB_OK;
}

extern "C" {

int ab_state_create(const ab_state_args* args, ab_matrix* X, ab_glm* glm, ab_state** out) {
    AB_TRY
    if (X->dtype != args->dtype || (glm && glm->dtype != args->dtype)) throw core_error("dtype mismatch between state, matrix and glm.");
    auto* s = new ab_state{args->dtype};
    try {
        if (args->dtype == AB_F32) s->f32 = make_state<float>(args, X->f32, glm ? glm->f32 : nullptr);
        else s->f64 = make_state<double>(args, X->f64, glm ? glm->f64 : nullptr);
    } catch (...) { delete s; throw; }
    *out = s;
    AB_CATCH
}
int ab_state_free(ab_state* s) { if (s) { delete s->f32; delete s->f64; delete s; } return AB_OK; }

int ab_state_solve(ab_state* s, int display_progress_bar, int (*exit_cond)(void*), void* ctx, int (*check_signals)(void),
                   char* err, size_t errlen, double* total_time) {
    (void)display_progress_bar;
    g_last_error.clear();
    s->error.clear();
    const double t0 = now_s();
    auto run = [&](auto* ps) {
        if (exit_cond) ps->exit_cond = [=]() { return exit_cond(ctx) != 0; };
        if (check_signals) ps->check_interrupt = [=]() { if (check_signals() != 0) throw solver_error("interrupted."); };
        try { ps->solve(); }
        catch (const std::exception& e) { s->error = e.what(); }       // py_state.cpp:83-90: message returned, state stays valid
        try { ps->finalize_invariance(); }
        catch (const std::exception& e) { if (s->error.empty()) s->error = e.what(); }
        ps->exit_cond = nullptr; ps->check_interrupt = nullptr;
    };
    if (s->dtype == AB_F32) run(s->f32); else run(s->f64);
    s->total_time = now_s() - t0;
    if (total_time) *total_time = s->total_time;
    if (err && errlen) { std::strncpy(err, s->error.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return AB_OK;
}
int ab_state_get_scalar(const ab_state* s, const char* name, double* out) {
    AB_TRY
    if (std::string(name) == "total_time") { *out = s->total_time; return AB_OK; }
    return s->dtype == AB_F32 ? state_scalar(*s->f32, name, out) : state_scalar(*s->f64, name, out);
    AB_CATCH
}
int ab_state_get_vec_f64(const ab_state* s, const char* name, double* out, int64_t cap, int64_t* len) {
    AB_TRY
    return s->dtype == AB_F32 ? state_vec_f64(*s->f32, name, out, cap, len) : state_vec_f64(*s->f64, name, out, cap, len);
    AB_CATCH
}
int ab_state_get_vec_i64(const ab_state* s, const char* name, int64_t* out, int64_t cap, int64_t* len) {
    AB_TRY
    return s->dtype == AB_F32 ? state_vec_i64(*s->f32, name, out, cap, len) : state_vec_i64(*s->f64, name, out, cap, len);
    AB_CATCH
}
int ab_state_get_betas(const ab_state* s, int64_t* indptr, int64_t* indices, double* values, int64_t* nnz, int64_t* L) {
    AB_TRY
    if (s->dtype == AB_F32) get_betas(*s->f32, indptr, indices, values, nnz, L); else get_betas(*s->f64, indptr, indices, values, nnz, L);
    AB_CATCH
}
int ab_state_get_screen_transform(const ab_state* s, int64_t i, double* out, int64_t cap, int64_t* len) {
    AB_TRY
    auto get = [&](auto& st) {
        if (i < 0 || i >= (int64_t)st.screen_transforms.size()) throw core_error("screen_transforms index out of range.");
        const auto& v = st.screen_transforms[i];
        *len = (int64_t)v.size();
        if (out) for (int64_t k = 0; k < std::min<int64_t>(cap, *len); ++k) out[k] = (double)v[k];
    };
    if (s->dtype == AB_F32) get(*s->f32); else get(*s->f64);
    AB_CATCH
}

// ------------------------------------------------------------------------------------------ bcd
static int bcd_run(int mode, int64_t q, const double* L, const double* v, double l1, double l2, double tol, int64_t max_iters, double aux,
                   double* x, double* scal) {
    AB_TRY
    if (q < 1) throw core_error("quad must be non-empty.");
    DevBuf<double> dL(q), dv(q), dx(q), ds(1);
    dL.upload(L, q); dv.upload(v, q);
    bcd_kernel<<<1, 32, (4 * q + 128) * sizeof(double), 0>>>(mode, (int)q, dL.p, dv.p, l1, l2, tol, (int)std::min<int64_t>(max_iters, 1 << 30), aux, dx.p, ds.p);
    AB_CUDA(cudaGetLastError());
    if (x) dx.download(x, q);
    ds.download(scal, 1);
    AB_CUDA(cudaStreamSynchronize(0));
    AB_CATCH
}
int ab_bcd_solve(int solver, int64_t q, const double* quad, const double* linear, double l1, double l2, double tol, int64_t max_iters,
                 double* x, int64_t* iters) {
    double it = 0;
    int rc = bcd_run(solver == 1 ? 1 : 0, q, quad, linear, l1, l2, tol, max_iters, 0, x, &it);
    if (iters) *iters = (int64_t)it;
    return rc;
}
int ab_bcd_root_lower_bound(int64_t q, const double* quad, const double* linear, double l1, double* out) {
    return bcd_run(2, q, quad, linear, l1, 0, 0, 0, 0, nullptr, out);
}
int ab_bcd_root_upper_bound(int64_t q, const double* quad, const double* linear, double l1, double zero_tol, double* out) {
    return bcd_run(3, q, quad, linear, l1, 0, 0, 0, zero_tol, nullptr, out);
}
int ab_bcd_root_function(int64_t q, double h, const double* D, const double* v, double l1, double* out) {
    return bcd_run(4, q, D, v, l1, 0, 0, 0, h, nullptr, out);
}

int ab_pin_naive_solve(ab_state* s, int (*check_signals)(void), char* err, size_t errlen, double* total_time) {
    g_last_error.clear();
    s->error.clear();
    const double t0 = now_s();
    auto run = [&](auto* ps) {
        if (check_signals) ps->check_interrupt = [=]() { if (check_signals() != 0) throw solver_error("interrupted."); };
        try { ps->solve_pin(); }
        catch (const std::exception& e) { s->error = e.what(); }       // py_state.cpp:83-90: message returned, state stays valid
        ps->check_interrupt = nullptr;
    };
    if (s->dtype == AB_F32) run(s->f32); else run(s->f64);
    s->total_time = now_s() - t0;
    if (total_time) *total_time = s->total_time;
    if (err && errlen) { std::strncpy(err, s->error.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return AB_OK;
}

} // extern "C"

// ------------------------------------------------------------------------------------------ covariance method
template <class T>
static CovPathState<T>* make_cov_state(const ab_cov_state_args* a, CovMatrix<T>* A) {
    auto st = std::make_unique<CovPathState<T>>();
    auto& s = *st;
    s.A = A; s.p = A->cols(); s.G = a->G;
    if (a->G < 1) throw core_error("groups must be non-empty.");
    s.groups.assign(a->groups, a->groups + a->G);
    s.group_sizes.assign(a->group_sizes, a->group_sizes + a->G);
    s.alpha = (T)a->alpha;
    s.penalty.assign((const T*)a->penalty, (const T*)a->penalty + a->G);
    if (a->v) s.v.assign((const T*)a->v, (const T*)a->v + s.p); else s.v.assign(s.p, T(0));
    s.min_ratio = (T)a->min_ratio; s.lmda_path_size = a->lmda_path_size; s.max_screen_size = a->max_screen_size; s.max_active_size = a->max_active_size;
    s.pivot_subset_ratio = (T)a->pivot_subset_ratio; s.pivot_subset_min = a->pivot_subset_min; s.pivot_slack_ratio = (T)a->pivot_slack_ratio;
    s.screen_rule = a->screen_rule; s.max_iters = a->max_iters; s.tol = (T)a->tol; s.rdev_tol = (T)a->rdev_tol;
    s.newton_tol = (T)a->newton_tol; s.newton_max_iters = a->newton_max_iters; s.early_exit = a->early_exit;
    s.setup_lmda_max = a->setup_lmda_max; s.setup_lmda_path = a->setup_lmda_path; s.n_threads = a->n_threads;
    s.lmda_max = (T)a->lmda_max; s.lmda = (T)a->lmda; s.rsq = (T)a->rsq;
    if (a->lmda_path && a->lmda_path_len > 0) s.lmda_path.assign((const T*)a->lmda_path, (const T*)a->lmda_path + a->lmda_path_len);
    s.screen_set.assign(a->screen_set, a->screen_set + a->screen_set_size);
    s.screen_beta.assign((const T*)a->screen_beta, (const T*)a->screen_beta + a->screen_beta_size);
    s.screen_is_active.assign(a->screen_is_active, a->screen_is_active + a->screen_set_size);
    s.active_set_size = a->active_set_size;
    s.active_set.assign(a->active_set, a->active_set + a->G);
    if (a->grad) s.grad.assign((const T*)a->grad, (const T*)a->grad + s.p); else s.grad.assign(s.p, T(0));
    for (int64_t g : s.screen_set) if (g < 0 || g >= a->G) throw core_error("screen_set entries must be in [0, G).");
    s.validate_and_init();
    if (a->screen_grad) {                                         // pin state: the gradient on the screen values is an input
        if ((size_t)a->screen_beta_size != s.screen_grad.size()) throw core_error("screen_grad must be (bs,) where screen_beta is (bs,).");
        s.screen_grad.assign((const T*)a->screen_grad, (const T*)a->screen_grad + a->screen_beta_size);
    }
    return st.release();
}

template <class T>
static int cov_state_vec_f64(const CovPathState<T>& s, const std::string& nm, double* out, int64_t cap, int64_t* len) {
    if (nm == "lmda_path") return copy_vec<T>(s.lmda_path, out, cap, len);
    if (nm == "screen_beta") return copy_vec<T>(s.screen_beta, out, cap, len);
    if (nm == "screen_grad") return copy_vec<T>(s.screen_grad, out, cap, len);
    if (nm == "screen_vars") return copy_vec<T>(s.screen_vars, out, cap, len);
    if (nm == "grad") return copy_vec<T>(s.grad, out, cap, len);
    if (nm == "abs_grad") return copy_vec<T>(s.abs_grad, out, cap, len);
    if (nm == "v") return copy_vec<T>(s.v, out, cap, len);
    if (nm == "devs") return copy_vec<T>(s.devs, out, cap, len);
    if (nm == "lmdas") return copy_vec<T>(s.lmdas, out, cap, len);
    if (nm == "rsqs") return copy_vec<T>(s.rsqs, out, cap, len);
    if (nm == "intercepts") return copy_vec<T>(s.intercepts, out, cap, len);
    if (nm == "penalty") return copy_vec<T>(s.penalty, out, cap, len);
    if (nm == "benchmark_screen") return copy_vec<double>(s.benchmark_screen, out, cap, len);
    if (nm == "benchmark_fit_screen") return copy_vec<double>(s.benchmark_fit_screen, out, cap, len);
    if (nm == "benchmark_fit_active") return copy_vec<double>(s.benchmark_fit_active, out, cap, len);
    if (nm == "benchmark_kkt") return copy_vec<double>(s.benchmark_kkt, out, cap, len);
    if (nm == "benchmark_invariance") return copy_vec<double>(s.benchmark_invariance, out, cap, len);
    if (nm == "sweep_stats") {
        *len = 8;
        if (out) {
            std::vector<long long> h(8, 0);
            if (s.d_stats.n) { s.d_stats.download(h.data(), 8); AB_CUDA(cudaStreamSynchronize(0)); }
            for (int64_t i = 0; i < std::min<int64_t>(cap, 8); ++i) out[i] = (double)h[i];
        }
        return AB_OK;
    }
    g_last_error = "adelie_core: unknown state vector " + nm;
    return AB_ERR_ARG;
}
template <class T>
static int cov_state_vec_i64(const CovPathState<T>& s, const std::string& nm, int64_t* out, int64_t cap, int64_t* len) {
    if (nm == "groups") return copy_ivec(s.groups, out, cap, len);
    if (nm == "group_sizes") return copy_ivec(s.group_sizes, out, cap, len);
    if (nm == "screen_set") return copy_ivec(s.screen_set, out, cap, len);
    if (nm == "screen_begins") return copy_ivec(s.screen_begins, out, cap, len);
    if (nm == "screen_is_active") return copy_ivec(s.screen_is_active, out, cap, len);
    if (nm == "active_set") return copy_ivec(s.active_set, out, cap, len);
    if (nm == "screen_subset_order") return copy_ivec(s.screen_subset_order, out, cap, len);
    if (nm == "screen_subset_ordered") return copy_ivec(s.screen_subset_ordered, out, cap, len);
    if (nm == "n_valid_solutions") return copy_ivec(s.n_valid_solutions, out, cap, len);
    if (nm == "active_sizes") return copy_ivec(s.active_sizes, out, cap, len);
    if (nm == "screen_sizes") return copy_ivec(s.screen_sizes, out, cap, len);
    g_last_error = "adelie_core: unknown state vector " + nm;
    return AB_ERR_ARG;
}
template <class T>
static int cov_state_scalar(const CovPathState<T>& s, const std::string& nm, double* out) {
    if (nm == "lmda_max") *out = s.lmda_max; else if (nm == "lmda") *out = s.lmda;
    else if (nm == "rsq") *out = s.rsq; else if (nm == "active_set_size") *out = (double)s.active_set_size;
    else if (nm == "alpha") *out = s.alpha; else if (nm == "tol") *out = s.tol; else if (nm == "rdev_tol") *out = s.rdev_tol;
    else if (nm == "n_sweeps") *out = (double)s.n_sweeps; else if (nm == "n_group_updates") *out = (double)s.n_group_updates;
    else if (nm == "n_col_updates") *out = (double)s.n_col_updates; else if (nm == "n_pin_solves") *out = (double)s.n_pin_solves;
    else if (nm == "n_kernel_launches") *out = (double)s.n_kernel_launches; else if (nm == "time_sweep_kernel") *out = s.time_sweep_kernel;
    else if (nm == "cov_cluster") *out = s.last_cluster; else if (nm == "cov_smem_bytes") *out = s.last_smem;
    else if (nm == "setup_lmda_max") *out = s.setup_lmda_max; else if (nm == "setup_lmda_path") *out = s.setup_lmda_path;
    else { g_last_error = "adelie_core: unknown state scalar " + nm; return AB_ERR_ARG; }
    return AB_OK;
}

extern "C" {

int ab_matrix_cov_dense_create(int dtype, const void* host, int64_t p, int order, int64_t ldh, int n_threads, ab_cov_matrix** out) {
    AB_TRY
    if (p < 1) throw core_error("mat must be (p, p).");
    auto* m = new ab_cov_matrix{dtype};
    try {
        if (dtype == AB_F32) m->f32 = CovMatrix<float>::make_dense((const float*)host, p, order, ldh, n_threads);
        else m->f64 = CovMatrix<double>::make_dense((const double*)host, p, order, ldh, n_threads);
    } catch (...) { delete m; throw; }
    *out = m;
    AB_CATCH
}
int ab_matrix_cov_lazy_create(int dtype, const void* host, int64_t n, int64_t p, int order, int64_t ldh, int n_threads, ab_cov_matrix** out) {
    AB_TRY
    if (n < 1 || p < 1) throw core_error("mat must be (n, p).");
    auto* m = new ab_cov_matrix{dtype};
    try {
        if (dtype == AB_F32) m->f32 = CovMatrix<float>::make_lazy((const float*)host, n, p, order, ldh, n_threads);
        else m->f64 = CovMatrix<double>::make_lazy((const double*)host, n, p, order, ldh, n_threads);
    } catch (...) { delete m; throw; }
    *out = m;
    AB_CATCH
}
int ab_matrix_cov_free(ab_cov_matrix* m) { if (m) { delete m->f32; delete m->f64; delete m; } return AB_OK; }
int ab_matrix_cov_cols(const ab_cov_matrix* m, int64_t* out) { *out = m->dtype == AB_F32 ? m->f32->p : m->f64->p; return AB_OK; }
int ab_matrix_cov_bmul(ab_cov_matrix* m, const int64_t* subset, int64_t s, const int64_t* indices, const void* values, int64_t k, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) m->f32->bmul(subset, s, indices, (const float*)values, k, (float*)out);
    else m->f64->bmul(subset, s, indices, (const double*)values, k, (double*)out);
    AB_CATCH
}
int ab_matrix_cov_mul(ab_cov_matrix* m, const int64_t* indices, const void* values, int64_t k, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) m->f32->mul(indices, (const float*)values, k, (float*)out);
    else m->f64->mul(indices, (const double*)values, k, (double*)out);
    AB_CATCH
}
int ab_matrix_cov_to_dense(ab_cov_matrix* m, int64_t i, int64_t q, void* out) {
    AB_TRY
    if (m->dtype == AB_F32) m->f32->to_dense(i, q, (float*)out); else m->f64->to_dense(i, q, (double*)out);
    AB_CATCH
}
int ab_matrix_cov_cache_info(const ab_cov_matrix* m, int64_t* cached_rows) {
    *cached_rows = m->dtype == AB_F32 ? m->f32->cached_rows : m->f64->cached_rows;
    return AB_OK;
}

int ab_cov_state_create(const ab_cov_state_args* args, ab_cov_matrix* A, ab_cov_state** out) {
    AB_TRY
    if (A->dtype != args->dtype) throw core_error("dtype mismatch between state and matrix.");
    auto* s = new ab_cov_state{args->dtype};
    try {
        if (args->dtype == AB_F32) s->f32 = make_cov_state<float>(args, A->f32);
        else s->f64 = make_cov_state<double>(args, A->f64);
    } catch (...) { delete s; throw; }
    *out = s;
    AB_CATCH
}
int ab_cov_state_free(ab_cov_state* s) { if (s) { delete s->f32; delete s->f64; delete s; } return AB_OK; }

static int cov_run(ab_cov_state* s, bool pin, int (*exit_cond)(void*), void* ctx, int (*check_signals)(void), char* err, size_t errlen, double* total_time) {
    g_last_error.clear();
    s->error.clear();
    const double t0 = now_s();
    auto run = [&](auto* ps) {
        if (exit_cond) ps->exit_cond = [=]() { return exit_cond(ctx) != 0; };
        if (check_signals) ps->check_interrupt = [=]() { if (check_signals() != 0) throw solver_error("interrupted."); };
        try { if (pin) ps->solve_pin(); else ps->solve(); }
        catch (const std::exception& e) { s->error = e.what(); }       // py_state.cpp:83-90: message returned, state stays valid
        ps->exit_cond = nullptr; ps->check_interrupt = nullptr;
    };
    if (s->dtype == AB_F32) run(s->f32); else run(s->f64);
    s->total_time = now_s() - t0;
    if (total_time) *total_time = s->total_time;
    if (err && errlen) { std::strncpy(err, s->error.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return AB_OK;
}
int ab_cov_state_solve(ab_cov_state* s, int display_progress_bar, int (*exit_cond)(void*), void* ctx, int (*check_signals)(void),
                       char* err, size_t errlen, double* total_time) {
    (void)display_progress_bar;
    return cov_run(s, false, exit_cond, ctx, check_signals, err, errlen, total_time);
}
int ab_cov_pin_solve(ab_cov_state* s, int (*check_signals)(void), char* err, size_t errlen, double* total_time) {
    return cov_run(s, true, nullptr, nullptr, check_signals, err, errlen, total_time);
}
int ab_cov_state_get_scalar(const ab_cov_state* s, const char* name, double* out) {
    AB_TRY
    if (std::string(name) == "total_time") { *out = s->total_time; return AB_OK; }
    return s->dtype == AB_F32 ? cov_state_scalar(*s->f32, name, out) : cov_state_scalar(*s->f64, name, out);
    AB_CATCH
}
int ab_cov_state_get_vec_f64(const ab_cov_state* s, const char* name, double* out, int64_t cap, int64_t* len) {
    AB_TRY
    return s->dtype == AB_F32 ? cov_state_vec_f64(*s->f32, name, out, cap, len) : cov_state_vec_f64(*s->f64, name, out, cap, len);
    AB_CATCH
}
int ab_cov_state_get_vec_i64(const ab_cov_state* s, const char* name, int64_t* out, int64_t cap, int64_t* len) {
    AB_TRY
    return s->dtype == AB_F32 ? cov_state_vec_i64(*s->f32, name, out, cap, len) : cov_state_vec_i64(*s->f64, name, out, cap, len);
    AB_CATCH
}
int ab_cov_state_get_betas(const ab_cov_state* s, int64_t* indptr, int64_t* indices, double* values, int64_t* nnz, int64_t* L) {
    AB_TRY
    auto get = [&](const auto& st) {
        int64_t tot = 0;
        if (indptr) indptr[0] = 0;
        for (size_t l = 0; l < st.betas.size(); ++l) {
            const auto& b = st.betas[l];
            for (size_t k = 0; k < b.idx.size(); ++k) { if (indices) indices[tot + k] = b.idx[k]; if (values) values[tot + k] = b.val[k]; }
            tot += (int64_t)b.idx.size();
            if (indptr) indptr[l + 1] = tot;
        }
        *nnz = tot; *L = (int64_t)st.betas.size();
    };
    if (s->dtype == AB_F32) get(*s->f32); else get(*s->f64);
    AB_CATCH
}
int ab_cov_state_get_screen_transform(const ab_cov_state* s, int64_t i, double* out, int64_t cap, int64_t* len) {
    AB_TRY
    auto get = [&](auto& st) {
        if (i < 0 || i >= (int64_t)st.screen_transforms.size()) throw core_error("screen_transforms index out of range.");
        const auto& v = st.screen_transforms[i];
        *len = (int64_t)v.size();
        if (out) for (int64_t k = 0; k < std::min<int64_t>(cap, *len); ++k) out[k] = (double)v[k];
    };
    if (s->dtype == AB_F32) get(*s->f32); else get(*s->f64);
    AB_CATCH
}

} // extern "C"
