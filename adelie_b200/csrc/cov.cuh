// adelie_b200/csrc/cov.cuh -- the COVARIANCE-METHOD Gaussian solver on the device.
//
// Reference: CORE/matrix/matrix_cov_{base,dense,lazy_cov}.{hpp,ipp}, CORE/solver/solver_gaussian_pin_cov.hpp:56-777 (pin solve),
// CORE/solver/solver_gaussian_cov.hpp:14-461 (path driver), CORE/state/state_gaussian_{pin_,}cov.{hpp,ipp}.
//   minimise  1/2 b^T A b - v^T b + lmda * sum_g pen_g (alpha ||b_g|| + (1 - alpha)/2 ||b_g||^2)
// Coordinate descent never touches the n observations: the gradient of the screen values is kept up to date with rows of A,
//   screen_grad[b'] -= sum_c A(col_k + c, col_b') del_c       after group k moved by del,
// so a group update costs (screen values) x gs multiply-adds whatever n is.
//
// Device design.  ONE thread-block cluster of up to 8 CTAs owns the whole solve (persistent, one launch per lambda):
//   * the screen gradient lives in DISTRIBUTED SHARED MEMORY: CTA r holds the slice [r L, (r + 1) L) of the screen values, about one
//     value per thread;
//   * every CTA replicates the proximal update (same inputs, same instructions => identical coefficients, nothing to broadcast
//     afterwards); the gs gradient values it needs are PUSHED by their owner threads into every CTA's shared memory over DSMEM one
//     group ahead, right after the owner updated them;
//   * every thread then applies the rank-gs update to its screen values from a compact Gram of the screen set (gathered from A once per
//     screen-set change, laid out so that the gs entries of one update are contiguous): one L2 round trip per update;
//   * ONE cluster barrier (release / acquire) per group update publishes the updated slices and the pushed values.
// Active-set sweeps update only the active positions and the inactive ones are brought up to date once, when the active set has
// converged (solve_active, :390-527), exactly like the reference.
#pragma once
#include "solver.cuh"
#include <cooperative_groups.h>
#include <memory>

namespace ab {

constexpr int kCovThreads = 512;
constexpr int kCovClusterMax = 8;
constexpr int kCovPre = kGsMax / 32;           // group elements per lane of the prox warp (group sizes up to kGsMax)

// ---------------------------------------------------------------------------------------------------------------
// small kernels of the MatrixCov operators
// ---------------------------------------------------------------------------------------------------------------
// dst[r * ld_dst + c] = src[c * ld_src + r]   (rows x cols result)
template <class T>
__global__ void cov_transpose_kernel(const T* __restrict__ src, int64_t ld_src, T* __restrict__ dst, int64_t ld_dst, int64_t rows, int64_t cols) {
    __shared__ T tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < cols && r < rows) ? src[c * ld_src + r] : T(0);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) dst[r * ld_dst + c] = tile[threadIdx.x][i];
    }
}

// out[s] = sum_k values[k] * rows[k][subset[s]]   (MatrixCovBase::bmul, matrix_cov_dense.ipp:25-43: accumulation in index order)
template <class T>
__global__ void cov_bmul_kernel(const T* const* __restrict__ rows, const T* __restrict__ values, int k, const int64_t* __restrict__ subset, int64_t s, T* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s) return;
    const int64_t j = subset[i];
    T acc = 0;
    for (int q = 0; q < k; ++q) acc += values[q] * rows[q][j];
    out[i] = acc;
}
// out[j] = sum_k values[k] * rows[k][j] for j < p   (MatrixCovBase::mul, :45-64); sub != nullptr: out[j] = sub[j] - sum (grad = v - A beta)
template <class T>
__global__ void cov_mul_kernel(const T* const* __restrict__ rows, const T* __restrict__ values, int k, int64_t p, const T* __restrict__ sub, T* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    T acc = 0;
    for (int q = 0; q < k; ++q) acc += values[q] * rows[q][j];
    out[j] = sub ? sub[j] - acc : acc;
}
// out[(item) gs*gs block, column-major] = A(g + r, g + c): diagonal blocks of the screen groups (MatrixCovBase::to_dense)
struct CovBlockItem { int64_t row0; int32_t g, gs; int64_t out_off; };       // row0: index of the group's first row pointer
template <class T>
__global__ void cov_blocks_kernel(const T* const* __restrict__ rows, const CovBlockItem* __restrict__ items, double* __restrict__ out) {
    const CovBlockItem it = items[blockIdx.x];
    for (int e = threadIdx.x; e < it.gs * it.gs; e += blockDim.x) {
        const int r = e % it.gs, c = e / it.gs;
        out[it.out_off + e] = (double)rows[it.row0 + r][it.g + c];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// MatrixCov on the device.  Both kinds expose "row i of A" as a device pointer: dense stores mat(i, .) contiguously, lazy_cov
// computes X[:, i]^T X on first use and keeps it (MatrixCovLazyCov::cache, matrix_cov_lazy_cov.ipp:10-48).
// ---------------------------------------------------------------------------------------------------------------
template <class T>
struct CovMatrix {
    int64_t p = 0; bool lazy = false; int n_threads = 1;
    DevBuf<T> D; int64_t ld = 0;                        // dense: D[i * ld + j] = mat(i, j)
    std::unique_ptr<DenseMatrix<T>> X;                  // lazy: the (n, p) data matrix
    std::vector<std::unique_ptr<DevBuf<T>>> cache_blocks;
    std::vector<const T*> row_of;                       // lazy: device pointer of the cached row (nullptr: not cached yet)
    int64_t cached_rows = 0;
    DevBuf<const T*> d_rows; DevBuf<T> d_vals, d_out; DevBuf<int64_t> d_subset;

    // host (p, p) array, order 0 = column-major, 1 = row-major, leading dimension ldh
    static CovMatrix* make_dense(const T* h, int64_t p_, int order, int64_t ldh, int n_threads_) {
        if (n_threads_ < 1) throw core_error("n_threads must be >= 1.");
        auto m = std::make_unique<CovMatrix>();
        m->p = p_; m->ld = p_; m->n_threads = n_threads_;
        m->D.alloc((size_t)p_ * p_);
        if (order == 1) {
            AB_CUDA(cudaMemcpy2D(m->D.p, p_ * sizeof(T), h, ldh * sizeof(T), p_ * sizeof(T), p_, cudaMemcpyHostToDevice));
        } else {
            DevBuf<T> raw((size_t)p_ * p_);
            AB_CUDA(cudaMemcpy2D(raw.p, p_ * sizeof(T), h, ldh * sizeof(T), p_ * sizeof(T), p_, cudaMemcpyHostToDevice));
            dim3 grid((unsigned)((p_ + 31) / 32), (unsigned)((p_ + 31) / 32));
            cov_transpose_kernel<T><<<grid, dim3(32, 8), 0, 0>>>(raw.p, p_, m->D.p, p_, p_, p_);
            AB_CUDA(cudaGetLastError());
            AB_CUDA(cudaStreamSynchronize(0));
        }
        return m.release();
    }
    static CovMatrix* make_lazy(const T* h, int64_t n_, int64_t p_, int order, int64_t ldh, int n_threads_) {
        if (n_threads_ < 1) throw core_error("n_threads must be >= 1.");
        auto m = std::make_unique<CovMatrix>();
        m->p = p_; m->lazy = true; m->n_threads = n_threads_;
        m->X = std::make_unique<DenseMatrix<T>>(n_, p_);
        m->X->n_threads = n_threads_;
        m->X->upload(h, order, ldh);
        m->row_of.assign(p_, nullptr);
        return m.release();
    }
    int64_t cols() const { return p; }

    // makes rows [i, i + q) available (lazy: q passes of the full-matrix GEMV kernel, X[:, i + k]^T X)
    void ensure_rows(int64_t i, int64_t q) {
        if (!lazy) return;
        int64_t a = i;
        while (a < i + q) {
            if (row_of[a]) { ++a; continue; }
            int64_t b = a;
            while (b < i + q && !row_of[b]) ++b;
            auto blk = std::make_unique<DevBuf<T>>((size_t)(b - a) * p);
            for (int64_t k = a; k < b; ++k) {
                T* dst = blk->p + (size_t)(k - a) * p;
                X->d_gemv_t(0, nullptr, (int)p, X->X + k * X->ld, X->d_ones(), dst);
                row_of[k] = dst;
            }
            cached_rows += b - a;
            cache_blocks.emplace_back(std::move(blk));
            a = b;
        }
    }
    const T* row_ptr(int64_t i) {
        if (!lazy) return D.p + i * ld;
        if (!row_of[i]) ensure_rows(i, 1);
        return row_of[i];
    }
    void upload_rows(const int64_t* indices, int64_t k) {
        std::vector<const T*> h(k);
        for (int64_t q = 0; q < k; ++q) {
            if (indices[q] < 0 || indices[q] >= p) throw core_error("matrix index out of range.");
            if (lazy && !row_of[indices[q]]) {               // cache maximal runs of consecutive uncached indices as one block (:80-86)
                int64_t run = 1;
                while (q + run < k && indices[q + run] == indices[q] + run && !row_of[indices[q + run]]) ++run;
                ensure_rows(indices[q], run);
            }
            h[q] = row_ptr(indices[q]);
        }
        d_rows.reserve_keep(k + 1);
        if (k) d_rows.upload(h.data(), k);
    }
    static void check_bmul(int64_t s, int64_t i, int64_t v, int64_t o, int64_t r, int64_t c) {         // matrix_cov_base.hpp:66-87
        if ((s < 0 || s > r) || (i < 0 || i > r) || (i != v) || (v < 0 || v > r) || (o != s)) {
            char buf[256];
            std::snprintf(buf, sizeof buf, "bmul() is given inconsistent inputs! Invoked check_bmul(s=%d, i=%d, v=%d, o=%d, r=%d, c=%d)",
                          (int)s, (int)i, (int)v, (int)o, (int)r, (int)c);
            throw core_error(buf);
        }
    }
    static void check_to_dense(int64_t i, int64_t q, int64_t r, int64_t c) {                           // :111-131
        if ((i < 0 || i > r - q) || (r != c)) {
            char buf[256];
            std::snprintf(buf, sizeof buf, "to_dense() is given inconsistent inputs! Invoked check_to_dense(i=%d, p=%d, o_r=%d, o_c=%d, r=%d, c=%d)",
                          (int)i, (int)q, (int)q, (int)q, (int)r, (int)c);
            throw core_error(buf);
        }
    }
    // host-pointer operators (the reference's Python-visible API)
    void bmul(const int64_t* subset, int64_t s, const int64_t* indices, const T* values, int64_t k, T* out) {
        check_bmul(s, k, k, s, p, p);
        for (int64_t i = 0; i < s; ++i) if (subset[i] < 0 || subset[i] >= p) throw core_error("matrix index out of range.");
        upload_rows(indices, k);
        d_vals.reserve_keep(k + 1); d_subset.reserve_keep(s + 1); d_out.reserve_keep(s + 1);
        if (k) d_vals.upload(values, k);
        if (s) d_subset.upload(subset, s);
        if (s) {
            cov_bmul_kernel<T><<<(unsigned)((s + 255) / 256), 256, 0, 0>>>(d_rows.p, d_vals.p, (int)k, d_subset.p, s, d_out.p);
            AB_CUDA(cudaGetLastError());
            d_out.download(out, s);
        }
        AB_CUDA(cudaStreamSynchronize(0));
    }
    // device-side mul: d_dst[j] = (sub ? sub[j] - . : .) sum_k values[k] A(indices[k], j)
    void d_mul(const int64_t* indices, const T* values, int64_t k, const T* d_sub, T* d_dst) {
        upload_rows(indices, k);
        d_vals.reserve_keep(k + 1);
        if (k) d_vals.upload(values, k);
        cov_mul_kernel<T><<<(unsigned)((p + 255) / 256), 256, 0, 0>>>(d_rows.p, d_vals.p, (int)k, p, d_sub, d_dst);
        AB_CUDA(cudaGetLastError());
    }
    void mul(const int64_t* indices, const T* values, int64_t k, T* out) {
        if (k < 0 || k > p) throw core_error("mul() is given inconsistent inputs!");
        d_out.reserve_keep(p + 1);
        d_mul(indices, values, k, nullptr, d_out.p);
        d_out.download(out, p);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    void to_dense(int64_t i, int64_t q, T* out /* q x q column-major */) {
        check_to_dense(i, q, p, p);
        if (q == 0) return;
        std::vector<int64_t> idx(q);
        for (int64_t k = 0; k < q; ++k) idx[k] = i + k;
        upload_rows(idx.data(), q);
        DevBuf<CovBlockItem> items(1); DevBuf<double> blk((size_t)q * q);
        CovBlockItem it{0, (int32_t)i, (int32_t)q, 0};
        items.upload(&it, 1);
        cov_blocks_kernel<T><<<1, 256, 0, 0>>>(d_rows.p, items.p, blk.p);
        AB_CUDA(cudaGetLastError());
        std::vector<double> h((size_t)q * q);
        blk.download(h.data(), h.size());
        AB_CUDA(cudaStreamSynchronize(0));
        for (size_t e = 0; e < h.size(); ++e) out[e] = (T)h[e];
    }
};

// ---------------------------------------------------------------------------------------------------------------
// the fused pin solve: pin::cov::solve (solver_gaussian_pin_cov.hpp:529-725) for ONE lambda
// ---------------------------------------------------------------------------------------------------------------
// G[b * ldg + b'] = A(vcol[b], vcol[b']) = vrow[b][vcol[b']]: the Gram of the screen values in screen order, so that a group update reads
// gs ROWS (the group's values b) at the columns of the threads' own screen values b' -- coalesced across a warp, no index lookups.
template <class T>
__global__ void cov_gather_gram_kernel(const T* const* __restrict__ vrow, const int32_t* __restrict__ vcol, int m, int64_t ldg, T* __restrict__ G) {
    const int bp = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (bp < m) G[(int64_t)b * ldg + bp] = vrow[b][vcol[bp]];
}

template <class T>
struct CovKernelArgs {
    const T* gram; int64_t ldg;    // [m][ldg] screen Gram, gram[b ldg + b'] = A(col_b, col_b')
    const GroupMeta* meta; int S;  // per screen position: col, gs, begin (first screen value), rec_off, pen
    const T* grec;                 // records [A(gs) | 0 | 0 | V(gs x gs, (r, c) -> r gs + c)]
    const T* beta_in; T* beta_rep; T* beta_old_rep; int64_t beta_stride;     // per-CTA replicas of screen_beta (replica 0 = output)
    T* sgrad;                      // [m] screen_grad, in / out
    const int8_t* is_active_in; int8_t* act_rep; int64_t act_stride;          // per-CTA replicas of screen_is_active (replica 0 = output)
    int32_t* active_set;           // [>= S] in / out (identical values written by every CTA)
    PinScalars* sc;
    int m, L;                      // screen values, slice length per CTA (multiple of 32, L * cluster size >= m)
    double lmda, alpha, tol, newton_tol, dbeta_tol; long long max_iters; int newton_max_iters, max_active_size;
    int gs_cap;                    // >= largest group size, multiple of 4
    int rec_cap;                   // elements of one staged record slot (multiple of 4); 0: records are read from global memory (groups above 32 columns)
    long long* stats;              // optional [8] cycle counters of CTA 0 / thread 0 (Configs::sweep_profile): prox, sync, update + push, barrier, -, -, groups
};

struct CovCtrl { int changed, next, error, was_active; };

template <class T>
struct CovSmem {
    // ctrl (64 B) | gsum[gs_cap] f64 | prox scratch 6 x [gs_cap] + 128 (double-sized) | del [gs_cap] (double-sized) | gbuf[2][gs_cap] (double-sized) |
    // rec[2][rec_cap] (T) | sg[L] (T) | pact[L] (int8)
    __host__ __device__ static size_t fixed_bytes(int gs_cap, int rec_cap) { return 64 + sizeof(double) * ((size_t)gs_cap * 10 + 128) + sizeof(T) * (size_t)2 * rec_cap; }
    __host__ __device__ static size_t total(int gs_cap, int rec_cap, int L) { return (fixed_bytes(gs_cap, rec_cap) + (size_t)L * (sizeof(T) + 1) + 15) / 16 * 16; }
};

// Persistent solve of one lambda by one cluster.  Protocol of a group update (ONE cluster barrier):
//   [the group's gradient sits in every CTA's local gbuf]  warp 0: prox (replicated)  ->  __syncthreads  ->  every thread: rank-gs update of
//   its screen values from the Gram, and the owners of the NEXT group's values push them into every CTA's other gbuf slot over DSMEM
//   ->  cluster barrier (release / acquire).
template <class T>
__global__ void __launch_bounds__(kCovThreads, 1)
cov_pin_kernel(const __grid_constant__ CovKernelArgs<T> a)
{
    namespace cg = cooperative_groups;
    using P = T;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), NC = (int)cluster.num_blocks();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CovCtrl* ctrl = reinterpret_cast<CovCtrl*>(smem_raw);
    double* gsum = reinterpret_cast<double*>(smem_raw + 64);                 // [gs_cap]
    double* pscr = gsum + a.gs_cap;                                          // 6 x [gs_cap] + 128
    double* del_raw = pscr + 6 * a.gs_cap + 128;                             // [gs_cap]
    double* gbuf_raw = del_raw + a.gs_cap;                                   // [2][gs_cap] (T)
    T* rec_s = reinterpret_cast<T*>(gbuf_raw + 2 * a.gs_cap);                // [2][rec_cap]: the record of the group in flight and of the next one
    T* sg = reinterpret_cast<T*>(smem_raw + CovSmem<T>::fixed_bytes(a.gs_cap, a.rec_cap));   // [L]
    int8_t* pact = reinterpret_cast<int8_t*>(sg + a.L);                      // [L]
    P* p_aold = reinterpret_cast<P*>(pscr);
    P* p_A = reinterpret_cast<P*>(pscr + a.gs_cap);
    P* p_gk = reinterpret_cast<P*>(pscr + 2 * a.gs_cap);
    P* p_gt = reinterpret_cast<P*>(pscr + 3 * a.gs_cap);
    P* p_atold = reinterpret_cast<P*>(pscr + 4 * a.gs_cap);
    P* p_at = reinterpret_cast<P*>(pscr + 5 * a.gs_cap);
    P* p_scr = reinterpret_cast<P*>(pscr + 6 * a.gs_cap);
    T* s_del = reinterpret_cast<T*>(del_raw);
    T* gbuf = reinterpret_cast<T*>(gbuf_raw);                                // slot s at gbuf + s * gs_cap

    const int L = a.L, m = a.m;
    const int s0 = rank * L;                                  // first screen value of this CTA's slice
    const int Lm = max(0, min(L, m - s0));                    // values in the slice
    T* my_beta = a.beta_rep + (size_t)rank * a.beta_stride;
    T* my_old = a.beta_old_rep + (size_t)rank * a.beta_stride;
    int8_t* my_act = a.act_rep + (size_t)rank * a.act_stride;

    // ---- init: replicas, slice of the gradient, per-position activity
    for (int i = tid; i < m; i += kCovThreads) my_beta[i] = a.beta_in[i];
    for (int i = tid; i < a.S; i += kCovThreads) my_act[i] = a.is_active_in[i];
    for (int i = tid; i < Lm; i += kCovThreads) sg[i] = a.sgrad[s0 + i];
    for (int i = tid; i < L; i += kCovThreads) pact[i] = 0;
    if (tid == 0) { ctrl->changed = 0; ctrl->next = 0; ctrl->error = 0; ctrl->was_active = 0; }
    __syncthreads();
    // expand screen_is_active to the positions of this slice (update_active_inactive_subset, :56-106)
    for (int ss = tid; ss < a.S; ss += kCovThreads) {
        if (!my_act[ss]) continue;
        const GroupMeta mm = a.meta[ss];
        for (int c = 0; c < mm.gs; ++c) { const int b = mm.begin + c - s0; if (b >= 0 && b < L) pact[b] = 1; }
    }
    cluster.sync();                                            // every CTA of the cluster is up before the first DSMEM store

    const P l1 = (P)(a.lmda * a.alpha), l2 = (P)(a.lmda * (1.0 - a.alpha));
    ProxState ps;                                              // replicated solver scalars (meaningful in warp 0)
    ps.rsq = a.sc->rsq; ps.resid_sum = 0; ps.cm = 0; ps.A = a.sc->active_set_size; ps.error = 0; ps.newton_iters_max = 0;
    long long iters = a.sc->iters, n_updates = a.sc->n_group_updates, n_cols = a.sc->n_col_updates;
    int n_active = ps.A;                                       // every thread tracks the active-set size (updated through ctrl)
    int final_error = 0;

    // the owner of a screen value of group (gb0, ggs) stores it into slot `slot` of EVERY CTA's gbuf (DSMEM)
    auto push_value = [&](int bl, int gb0, int ggs, int slot) {
        const int c = s0 + bl - gb0;
        if (c >= 0 && c < ggs) {
            const T val = sg[bl];
            T* dst = gbuf + slot * a.gs_cap + c;
            for (int r = 0; r < NC; ++r) *cluster.map_shared_rank(dst, r) = val;
        }
    };

    // One sweep (coordinate_descent, :243-385) over the active list (kind 0) or the whole screen set (kind 1).  The descriptor, the
    // record and the current coefficients of a group are fetched ONE GROUP AHEAD (a group never changes another group's coefficients).
    auto sweep = [&](int kind, int count) {
        if (warp == 0) ps.cm = 0;
        if (count <= 0) return;
        int ss = (kind == 0) ? a.active_set[0] : 0;
        GroupMeta mm = a.meta[ss];
        int cur = 0;                                           // gbuf / record slot of the group in flight
        const bool prof = a.stats != nullptr && rank == 0 && tid == 0;
        long long pt[7] = {0, 0, 0, 0, 0, 0, 0}; long long tc = prof ? clock64() : 0;
#define COV_TICK(k) do { if (prof) { const long long t_ = clock64(); pt[k] += t_ - tc; tc = t_; } } while (0)
        if (a.rec_cap && warp == 1) for (int e = lane; e < mm.rec_elems; e += 32) rec_s[e] = a.grec[mm.rec_off + e];
        for (int bl = tid; bl < Lm; bl += kCovThreads) push_value(bl, mm.begin, mm.gs, 0);
        T nold[kCovPre] = {};                                  // warp 0: old coefficients of the group, element lane + 32 k
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < kCovPre; ++k) { const int c = lane + 32 * k; nold[k] = (c < mm.gs) ? my_beta[mm.begin + c] : T(0); }
        }
        cluster.sync();
#pragma unroll 1
        for (int idx = 0; idx < count; ++idx) {
            const int gs = mm.gs, b0 = mm.begin;
            const bool has_next = idx + 1 < count;
            const int ss_n = has_next ? ((kind == 0) ? a.active_set[idx + 1] : idx + 1) : ss;
            const GroupMeta mm_n = a.meta[ss_n];               // (consumed after the prox)
            // The Gram entries of this update (rows of the group, columns of the thread's first two screen values; the first 12 rows) do not
            // depend on the proximal solve: they are loaded NOW and consumed after it, so that their L2 / HBM latency hides behind the solve.
            constexpr bool kPre2 = sizeof(T) == 4;             // (float64: a second value's entries would spill; it is loaded on demand)
            const int pl1 = tid, pl2 = tid + kCovThreads;
            const bool pre1 = pl1 < Lm && !(kind == 0 && !pact[pl1]);
            const bool pre2 = kPre2 && pl2 < Lm && !(kind == 0 && !pact[pl2]);
            T px[12], py[12];
            {
                const T* g1 = a.gram + (size_t)b0 * a.ldg + (s0 + (pre1 ? pl1 : 0));
                const T* g2 = a.gram + (size_t)b0 * a.ldg + (s0 + (pre2 ? pl2 : 0));
#pragma unroll
                for (int u = 0; u < 12; ++u) {
                    px[u] = (u < gs && pre1) ? g1[(size_t)u * a.ldg] : T(0);
                    py[u] = (u < gs && pre2) ? g2[(size_t)u * a.ldg] : T(0);
                }
            }
            COV_TICK(3);
            if (warp == 0) {
                const T* gcur = gbuf + cur * a.gs_cap;
#pragma unroll
                for (int k = 0; k < kCovPre; ++k) { const int c = lane + 32 * k; if (c < gs) { gsum[c] = (double)gcur[c]; p_aold[c] = (P)nold[k]; } }
                __syncwarp();
                int changed = 0;
                const T* rec = a.rec_cap ? rec_s + (size_t)cur * a.rec_cap : a.grec + mm.rec_off;
                const P pk = (P)mm.pen;
                if (gs == 1) {                                 // :291-322
                    const P ak_old = p_aold[0], A_kk = (P)rec[0];
                    P gk = (P)gsum[0] + ak_old * A_kk;
                    const P vv = fabs(gk) - l1 * pk;           // update_coordinate, pin_base.hpp:181-195
                    P ak = (vv > P(0)) ? copysign(vv, gk) / (A_kk + l2 * pk) : P(0);
                    ak = (P)(T)ak;
                    gk -= ak_old * A_kk;
                    if (ak != ak_old) {
                        const P dd = ak - ak_old;
                        ps.cm = fmax(ps.cm, (double)(A_kk * dd * dd));
                        ps.rsq += (double)(dd * (2 * gk - dd * A_kk));
                        if (lane == 0) { my_beta[b0] = (T)ak; s_del[0] = (T)(-dd); }
                        changed = 1;
                    }
                } else {
                    const ProxCtx<T, P> px{p_aold, p_A, p_gk, p_gt, p_atold, p_at, p_scr, s_del, gsum, my_beta};
                    if (gs <= 32) {                            // one coefficient per lane, record in shared memory
                        const ProxPre<P> pre = prox_small_pre<T, P>(rec, gs, p_aold, lane);
                        changed = prox_small_post<T, P>(px, pre, rec, gs, b0, l1 * pk, l2 * pk, (P)a.newton_tol, a.newton_max_iters, (P)a.dbeta_tol, 0, ps, lane, nullptr);
                    } else {
                        changed = prox_group<T, P>(px, rec, gs, b0, l1 * pk, l2 * pk, (P)a.newton_tol, a.newton_max_iters, (P)a.dbeta_tol, 0, ps, lane, nullptr);
                    }
                }
                int was_active = 1;
                if (changed && kind == 1) {                    // add_active_set (:617-628)
                    was_active = my_act[ss];
                    if (!was_active) {
                        if (ps.A >= a.max_active_size) ps.error = kErrMaxActive;
                        else {
                            if (lane == 0) { my_act[ss] = 1; a.active_set[ps.A] = ss; }
                            ++ps.A;
                        }
                    }
                }
                ++n_updates; n_cols += gs;
                __syncwarp();
                if (lane == 0) { ctrl->changed = changed; ctrl->was_active = was_active; ctrl->error = ps.error; }
                if (has_next) {                                // the next group's coefficients: the loads complete behind the slice update
#pragma unroll
                    for (int k = 0; k < kCovPre; ++k) { const int c = lane + 32 * k; nold[k] = (c < mm_n.gs) ? my_beta[mm_n.begin + c] : T(0); }
                }
            } else if (warp == 1 && a.rec_cap && has_next) {   // stage the next group's record while warp 0 solves this one
                T* dst = rec_s + (size_t)(cur ^ 1) * a.rec_cap;
                for (int e = lane; e < mm_n.rec_elems; e += 32) dst[e] = a.grec[mm_n.rec_off + e];
            }
            COV_TICK(0);
            __syncthreads();                                   // del / ctrl of this CTA's warp 0, the staged record
            const int changed = ctrl->changed;
            COV_TICK(1);
            // rank-gs update of this CTA's slice: sg[b'] -= sum_c del_c A(col_k + c, col_b'), del = new - old = -s_del (two values of the thread
            // in flight, gs contiguous Gram entries each); then the owners of the next group's values publish them to every CTA
            const int nb0 = mm_n.begin, ngs = has_next ? mm_n.gs : 0;
            // the thread's first value(s): from the prefetched entries (+ the rows of a group beyond the first 12)
            if (changed && (pre1 || pre2)) {
                T acc1 = 0, acc2 = 0;
#pragma unroll
                for (int u = 0; u < 12; ++u) if (u < gs) { const T d = s_del[u]; acc1 -= d * px[u]; acc2 -= d * py[u]; }
#pragma unroll 1
                for (int c = 12; c < gs; ++c) {
                    const T d = s_del[c];
                    const T* gr = a.gram + (size_t)(b0 + c) * a.ldg + s0;
                    if (pre1) acc1 -= d * gr[pl1];
                    if (pre2) acc2 -= d * gr[pl2];
                }
                if (pre1) sg[pl1] -= acc1;
                if (pre2) sg[pl2] -= acc2;
            }
            if (ngs) { if (pl1 < Lm) push_value(pl1, nb0, ngs, cur ^ 1); if (kPre2 && pl2 < Lm) push_value(pl2, nb0, ngs, cur ^ 1); }
            // further values of the thread (slices above 512 / 1024 values): loaded on demand
#pragma unroll 1
            for (int bl = tid + (kPre2 ? 2 : 1) * kCovThreads; bl < Lm; bl += 2 * kCovThreads) {
                const int bl2 = bl + kCovThreads;
                const bool in2 = bl2 < Lm;
                const bool on1 = changed && !(kind == 0 && !pact[bl]);
                const bool on2 = changed && in2 && !(kind == 0 && !pact[bl2]);
                if (on1 || on2) {
                    const T* g1 = a.gram + (size_t)b0 * a.ldg + (s0 + bl);            // column b' of the group's rows
                    const int off2 = in2 ? kCovThreads : 0;
                    T acc1 = 0, acc2 = 0;
#pragma unroll 1
                    for (int c0 = 0; c0 < gs; c0 += 12) {                             // (groups of <= 12: every load of the update in flight at once)
                        T x[12], y[12];
#pragma unroll
                        for (int u = 0; u < 12; ++u) {
                            const bool in = c0 + u < gs;
                            const T* gr = g1 + (size_t)(c0 + u) * a.ldg;
                            x[u] = (in && on1) ? gr[0] : T(0);
                            y[u] = (in && on2) ? gr[off2] : T(0);
                        }
#pragma unroll
                        for (int u = 0; u < 12; ++u) if (c0 + u < gs) { const T d = s_del[c0 + u]; acc1 -= d * x[u]; acc2 -= d * y[u]; }
                    }
                    if (on1) sg[bl] -= acc1;
                    if (on2) sg[bl2] -= acc2;
                }
                if (ngs) { push_value(bl, nb0, ngs, cur ^ 1); if (in2) push_value(bl2, nb0, ngs, cur ^ 1); }
            }
            if (changed && !ctrl->was_active) {
                ++n_active;
                for (int c = tid; c < gs; c += kCovThreads) { const int b = b0 + c - s0; if (b >= 0 && b < L) pact[b] = 1; }
            }
            const int err = ctrl->error;
            COV_TICK(2);
            cluster.sync();                                    // updates + pushed values visible everywhere (also the CTA barrier that frees ctrl / del)
            if (prof) ++pt[6];
            if (err) { final_error = err; break; }
            ss = ss_n; mm = mm_n; cur ^= 1;
        }
        COV_TICK(3);
        if (prof) for (int k = 0; k < 7; ++k) a.stats[k] += pt[k];
#undef COV_TICK
    };

    // pin::cov::solve (:632-700): { solve_active; screen sweep } until the screen sweep converges
    while (true) {
        // ---- solve_active (:390-527)
        for (int i = tid; i < m; i += kCovThreads) my_old[i] = my_beta[i];      // (old active beta; a full copy is simpler than the active subset)
        __syncthreads();
        const int A0 = n_active;
        while (true) {
            if (warp == 0) ++iters;
            sweep(0, A0);
            if (final_error) break;
            if (tid == 0) ctrl->next = (ps.cm < a.tol) ? 1 : ((iters >= a.max_iters) ? -kErrMaxCds : 0);
            __syncthreads();
            const int nx = ctrl->next;
            __syncthreads();
            if (nx < 0) { final_error = -nx; break; }
            if (nx == 1) break;
        }
        if (final_error) break;
        // ---- gradient of the inactive screen values for the whole active-set move (:500-526); slices are only read remotely through
        //      the pushes of their own threads, so this is CTA-local work
        if (A0 > 0 && A0 < a.S) {
#pragma unroll 1
            for (int bl = tid; bl < Lm; bl += kCovThreads) {
                if (pact[bl]) continue;
                const T* g = a.gram + (s0 + bl);
                T acc = 0;
#pragma unroll 1
                for (int ai = 0; ai < A0; ++ai) {
                    const GroupMeta ma = a.meta[a.active_set[ai]];
#pragma unroll 1
                    for (int c = 0; c < ma.gs; ++c) {
                        const int b = ma.begin + c;
                        const T d = my_beta[b] - my_old[b];
                        if (d != T(0)) acc += d * g[(size_t)b * a.ldg];
                    }
                }
                sg[bl] -= acc;
            }
            __syncthreads();
        }
        // ---- one sweep over the screen set (:636-667)
        if (warp == 0) ++iters;
        sweep(1, a.S);
        if (final_error) break;
        if (tid == 0) ctrl->next = (ps.cm < a.tol) ? 1 : ((iters >= a.max_iters) ? -kErrMaxCds : 0);
        __syncthreads();
        const int nx = ctrl->next;
        __syncthreads();
        if (nx < 0) { final_error = -nx; break; }
        if (nx == 1) break;
    }
    for (int i = tid; i < Lm; i += kCovThreads) a.sgrad[s0 + i] = sg[i];
    if (rank == 0 && tid == 0) {
        a.sc->rsq = ps.rsq; a.sc->active_set_size = ps.A; a.sc->iters = iters;
        a.sc->n_group_updates = n_updates; a.sc->n_col_updates = n_cols; a.sc->error = final_error;
        a.sc->newton_iters_max = ps.newton_iters_max;
    }
    cluster.sync();                                            // no CTA exits while a peer could still store into its shared memory
}

// ---------------------------------------------------------------------------------------------------------------
// StateGaussianCov / StateGaussianPinCov on the host: gaussian::cov::solve (solver_gaussian_cov.hpp:359-457) = solve_core
// (solver_base.hpp:435-687) with the covariance-method pieces; the O(G) screening logic restates solver_base.hpp like
// PathState (solver.cuh) does for the naive method.
// ---------------------------------------------------------------------------------------------------------------
template <class T>
struct CovPathState {
    using idx_t = int64_t;
    // ---------------- static (state_gaussian_cov.hpp:39-145)
    CovMatrix<T>* A = nullptr;
    idx_t p = 0, G = 0;
    std::vector<T> v;
    std::vector<idx_t> groups, group_sizes;
    T alpha = 1; std::vector<T> penalty;
    T min_ratio = 1e-2; size_t lmda_path_size = 100, max_screen_size = 0, max_active_size = 0;
    T pivot_subset_ratio = 0.1; size_t pivot_subset_min = 1; T pivot_slack_ratio = 1.25; int screen_rule = 1;
    size_t max_iters = 100000; T tol = 1e-7, rdev_tol = 1e-4, newton_tol = 1e-12; size_t newton_max_iters = 1000;
    bool early_exit = true, setup_lmda_max = true, setup_lmda_path = true;
    size_t n_threads = 1;
    // ---------------- dynamic
    T lmda_max = -1; std::vector<T> lmda_path;
    std::vector<uint8_t> in_screen;
    std::vector<idx_t> screen_set, screen_begins;
    std::vector<T> screen_beta; std::vector<int8_t> screen_is_active;
    size_t active_set_size = 0; std::vector<idx_t> active_set;
    T lmda = std::numeric_limits<T>::infinity(), rsq = 0;
    std::vector<T> grad, abs_grad;
    std::vector<T> screen_vars, screen_grad; std::vector<std::vector<T>> screen_transforms;
    std::vector<idx_t> screen_subset, screen_subset_order, screen_subset_ordered;
    // ---------------- outputs
    std::vector<SparseRow> betas; std::vector<T> intercepts, devs, lmdas, rsqs;
    std::vector<double> benchmark_screen, benchmark_fit_screen, benchmark_fit_active, benchmark_kkt, benchmark_invariance;
    std::vector<int> n_valid_solutions, active_sizes, screen_sizes;
    long long n_sweeps = 0, n_group_updates = 0, n_col_updates = 0, n_pin_solves = 0, n_kernel_launches = 0;
    double time_sweep_kernel = 0;
    int last_cluster = 0, last_smem = 0;
    HostTimers timers;
    // ---------------- device
    DevBuf<T> d_v, d_grad, d_grec, d_beta_in, d_beta_rep, d_beta_old, d_sgrad;
    DevBuf<GroupMeta> d_meta; DevBuf<const T*> d_vrow; DevBuf<int32_t> d_vcol, d_active_set; DevBuf<T> d_gram;
    DevBuf<int8_t> d_act_in, d_act_rep; DevBuf<PinScalars> d_sc; PinnedBuf<PinScalars> h_sc; DevBuf<long long> d_stats;
    std::vector<GroupMeta> h_meta; std::vector<T> h_grec; std::vector<const T*> h_vrow; std::vector<int32_t> h_vcol;
    size_t tables_uploaded_S = (size_t)-1;
    int gs_max_screen = 1;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::function<bool()> exit_cond;
    std::function<void()> check_interrupt;

    ~CovPathState() { if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); }

    // state_base.ipp:10-116 + state_gaussian_cov.ipp:9-17
    void validate_and_init() {
        if ((idx_t)group_sizes.size() != G) throw core_error("group_sizes must be (G,) where groups is (G,).");
        if ((idx_t)penalty.size() != G) throw core_error("penalty must be (G,) where groups is (G,).");
        if (alpha < 0 || alpha > 1) throw core_error("alpha must be in [0,1].");
        if (tol < 0) throw core_error("tol must be >= 0.");
        if (newton_tol < 0) throw core_error("newton_tol must be >= 0.");
        if (n_threads < 1) throw core_error("n_threads must be >= 1.");
        if (min_ratio < 0 || min_ratio > 1) throw core_error("min_ratio must be in [0,1].");
        if (pivot_subset_ratio <= 0 || pivot_subset_ratio > 1) throw core_error("pivot_subset_ratio must be in (0,1].");
        if (pivot_subset_min < 1) throw core_error("pivot_subset_min must be >= 1.");
        if (pivot_slack_ratio < 0) throw core_error("pivot_slack_ratio must be >= 0.");
        if (screen_set.size() != screen_is_active.size()) throw core_error("screen_is_active must be (s,) where screen_set is (s,).");
        if (screen_beta.size() < screen_set.size())
            throw core_error("screen_beta must be (bs,) where bs >= s and screen_set is (s,). It is likely screen_beta has been initialized incorrectly. ");
        if (active_set_size > (size_t)G) throw core_error("active_set_size must be <= G where groups is (G,).");
        if ((idx_t)active_set.size() != G) throw core_error("active_set must be (G,) where groups is (G,).");
        if ((idx_t)grad.size() != groups[G - 1] + group_sizes[G - 1])
            throw core_error("grad.size() != groups[G-1] + group_sizes[G-1]. It is likely either grad has the wrong shape, or groups/group_sizes have been initialized incorrectly.");
        if ((idx_t)v.size() != A->cols()) throw core_error("v must be (p,) where A is (p, p).");
        for (idx_t g = 0; g < G; ++g) if (group_sizes[g] > kGsMax) throw core_error("group sizes above 128 are not supported by the device solver.");
        abs_grad.assign(G, 0);
        in_screen.assign(G, 0);
        AB_CUDA(cudaEventCreate(&ev0)); AB_CUDA(cudaEventCreate(&ev1));
        d_sc.alloc(1); h_sc.alloc(1); d_active_set.alloc(G + 1);
        d_v.alloc(p); d_v.upload(v.data(), p); d_grad.alloc(p);
        update_screen_derived_base();
        update_abs_grad(lmda);
        update_screen_derived();
    }

    // solver_base.hpp:20-110 (constraints == nullptr)
    void update_abs_grad(T lmda_) {
        for (size_t ss = 0; ss < screen_set.size(); ++ss) {
            const idx_t i = screen_set[ss], b = screen_begins[ss], k = groups[i], sz = group_sizes[i];
            const T regul = ((1 - alpha) * lmda_) * penalty[i];
            T a = 0;
            for (idx_t c = 0; c < sz; ++c) { const T e = grad[k + c] - regul * screen_beta[b + c]; a += e * e; }
            abs_grad[i] = std::sqrt(a);
        }
        for (idx_t i = 0; i < G; ++i) {
            if (in_screen[i]) continue;
            const idx_t k = groups[i], sz = group_sizes[i];
            T a = 0;
            for (idx_t c = 0; c < sz; ++c) a += grad[k + c] * grad[k + c];
            abs_grad[i] = std::sqrt(a);
        }
    }
    // solver_base.hpp:120-153
    void update_screen_derived_base() {
        const size_t old = screen_begins.size();
        for (size_t i = old; i < screen_set.size(); ++i) in_screen[screen_set[i]] = 1;
        size_t vs = (old == 0) ? 0 : (screen_begins.back() + group_sizes[screen_set[old - 1]]);
        for (size_t i = old; i < screen_set.size(); ++i) { screen_begins.push_back(vs); vs += group_sizes[screen_set[i]]; }
        screen_beta.resize(vs, 0);
        screen_is_active.resize(screen_set.size(), 0);
    }
    // update_screen_derived (solver_gaussian_cov.hpp:20-140): diagonal blocks A_gg of the new groups -> eigendecomposition (host
    // Jacobi, double) -> screen_vars / screen_transforms and the device tables (records, row pointers, columns); screen_grad from grad
    void update_screen_derived() {
        update_screen_derived_base();
        const size_t old_S = screen_transforms.size(), new_S = screen_set.size();
        const size_t old_vs = screen_subset.size();
        const size_t new_vs = new_S ? (size_t)(screen_begins.back() + group_sizes[screen_set.back()]) : 0;
        screen_transforms.resize(new_S); screen_vars.resize(new_vs, 0); screen_grad.resize(new_vs, 0);
        screen_subset.resize(new_vs); h_vrow.resize(new_vs); h_vcol.resize(new_vs); h_meta.resize(new_S);
        if (new_S > old_S) {
            AB_TIME(timers, "screen_records");
            std::vector<CovBlockItem> items; int64_t tot = 0;
            for (size_t i = old_S; i < new_S; ++i) {
                const idx_t g = groups[screen_set[i]], gs = group_sizes[screen_set[i]], sb = screen_begins[i];
                A->ensure_rows(g, gs);
                for (idx_t c = 0; c < gs; ++c) { h_vrow[sb + c] = A->row_ptr(g + c); h_vcol[sb + c] = (int32_t)(g + c); screen_subset[sb + c] = g + c; }
                items.push_back({(int64_t)sb, (int32_t)g, (int32_t)gs, tot});
                tot += gs * gs;
                gs_max_screen = std::max<int>(gs_max_screen, (int)gs);
            }
            d_vrow.reserve_keep(new_vs + 1);
            d_vrow.upload(h_vrow.data() + old_vs, new_vs - old_vs, old_vs);
            DevBuf<CovBlockItem> d_items(items.size()); DevBuf<double> d_blk((size_t)tot);
            d_items.upload(items.data(), items.size());
            cov_blocks_kernel<T><<<(unsigned)items.size(), 128, 0, 0>>>(d_vrow.p, d_items.p, d_blk.p);
            AB_CUDA(cudaGetLastError());
            std::vector<double> blk((size_t)tot);
            d_blk.download(blk.data(), blk.size());
            AB_CUDA(cudaStreamSynchronize(0));
            ++n_kernel_launches;
            for (size_t ii = 0; ii < items.size(); ++ii) {
                const size_t i = old_S + ii;
                const idx_t gs = items[ii].gs, sb = screen_begins[i];
                GroupMeta& mm = h_meta[i];
                mm.col = items[ii].g; mm.gs = (int32_t)gs; mm.begin = (int32_t)sb; mm.pen = (double)penalty[screen_set[i]];
                mm.rec_off = (int64_t)h_grec.size(); mm.rec_elems = (int32_t)((3 * gs + gs * gs + 3) / 4 * 4);
                h_grec.resize(h_grec.size() + mm.rec_elems, T(0));
                T* rec = h_grec.data() + mm.rec_off;
                if (gs == 1) {
                    screen_transforms[i].assign(1, T(1));
                    screen_vars[sb] = std::max<T>((T)blk[items[ii].out_off], 0);
                    rec[0] = screen_vars[sb]; rec[3] = T(1);
                    continue;
                }
                std::vector<double> Agg(blk.begin() + items[ii].out_off, blk.begin() + items[ii].out_off + gs * gs), D, V;
                for (idx_t r = 0; r < gs; ++r) for (idx_t c = r + 1; c < gs; ++c) {          // symmetrise (the solver reads the lower triangle)
                    const double s_ = 0.5 * (Agg[r + c * gs] + Agg[c + r * gs]); Agg[r + c * gs] = s_; Agg[c + r * gs] = s_;
                }
                host_jacobi_eigh(Agg, (int)gs, D, V);
                std::vector<T> Vr((size_t)gs * gs);                             // host_jacobi_eigh: V[r * gs + c], eigenvectors in the columns
                for (idx_t e = 0; e < gs * gs; ++e) Vr[e] = (T)V[e];
                for (idx_t c = 0; c < gs; ++c) { screen_vars[sb + c] = (T)(D[c] * (D[c] >= 0)); rec[c] = screen_vars[sb + c]; }
                for (idx_t e = 0; e < gs * gs; ++e) rec[3 * gs + e] = Vr[e];
                screen_transforms[i] = std::move(Vr);
            }
        }
        for (size_t i = 0; i < new_S; ++i) {                                    // :99-109
            const idx_t g = groups[screen_set[i]], gs = group_sizes[screen_set[i]], sb = screen_begins[i];
            for (idx_t c = 0; c < gs; ++c) screen_grad[sb + c] = grad[g + c];
        }
        screen_subset_order.resize(new_vs);                                     // :124-139
        std::iota(screen_subset_order.begin() + old_vs, screen_subset_order.end(), (idx_t)old_vs);
        std::sort(screen_subset_order.begin(), screen_subset_order.end(), [&](idx_t i, idx_t j) { return screen_subset[i] < screen_subset[j]; });
        screen_subset_ordered.resize(new_vs);
        for (size_t i = 0; i < new_vs; ++i) screen_subset_ordered[i] = screen_subset[screen_subset_order[i]];
    }
    void upload_tables() {
        const size_t S = screen_set.size(), m = screen_subset.size();
        if (tables_uploaded_S == S) return;
        d_meta.reserve_keep(S + 1); d_grec.reserve_keep(h_grec.size() + 4); d_vcol.reserve_keep(m + 1); d_vrow.reserve_keep(m + 1);
        if (S) d_meta.upload(h_meta.data(), S);
        if (!h_grec.empty()) d_grec.upload(h_grec.data(), h_grec.size());
        if (m) {
            d_vcol.upload(h_vcol.data(), m); d_vrow.upload(h_vrow.data(), m);
            // compact Gram of the screen values (rebuilt whenever the screen set grew: m^2 gathered elements, well under a millisecond
            // for the screen sets the covariance method is used with)
            if ((double)m * (double)m * sizeof(T) > 64e9) throw core_error("the screen set is too large for the device covariance solver (screen Gram above 64 GB).");
            if (d_gram.n < m * m) { d_gram.free(); d_gram.alloc(m * m); }
            dim3 grid((unsigned)((m + 255) / 256), (unsigned)m);
            cov_gather_gram_kernel<T><<<grid, 256, 0, 0>>>(d_vrow.p, d_vcol.p, (int)m, (int64_t)m, d_gram.p);
            AB_CUDA(cudaGetLastError());
            ++n_kernel_launches;
        }
        tables_uploaded_S = S;
    }

    // One launch of the fused kernel = pin::cov::solve for one lambda on the current screen set (solver_gaussian_pin_cov.hpp:529-725)
    PinResult run_pin(T lmda_, size_t max_iters_left, size_t lmda_index) {
        const size_t S = screen_set.size(), m = screen_subset.size();
        PinResult R;
        if (check_interrupt) check_interrupt();
        upload_tables();
        const DeviceInfo& di = DeviceInfo::get();
        const int gs_cap = (std::max(gs_max_screen, 1) + 3) / 4 * 4;
        // cluster size: enough CTAs that every thread owns about one screen value, slices a multiple of 32
        int NC = 1;
        while (NC < kCovClusterMax && (size_t)NC * kCovThreads < m) NC *= 2;
        if (Configs::cov_cluster == 1 || Configs::cov_cluster == 2 || Configs::cov_cluster == 4 || Configs::cov_cluster == 8) NC = Configs::cov_cluster;
        int L = (int)((((m + NC - 1) / NC) + 31) / 32 * 32);
        L = std::max(L, 32);
        const int rec_cap = gs_cap <= 32 ? (3 * gs_cap + gs_cap * gs_cap + 3) / 4 * 4 : 0;     // records staged in shared memory (one coefficient per lane)
        size_t smem = CovSmem<T>::total(gs_cap, rec_cap, L);
        while (smem > di.smem_optin && NC < kCovClusterMax) { NC *= 2; L = std::max<int>(32, (int)((((m + NC - 1) / NC) + 31) / 32 * 32)); smem = CovSmem<T>::total(gs_cap, rec_cap, L); }
        if (smem > di.smem_optin) throw core_error("the screen set is too large for the device covariance solver (screen gradient does not fit the cluster's shared memory).");
        last_cluster = NC; last_smem = (int)smem;
        const size_t stride = std::max<size_t>(m, 1), astride = std::max<size_t>(S, 1);
        d_beta_in.reserve_keep(stride); d_beta_rep.reserve_keep(stride * NC); d_beta_old.reserve_keep(stride * NC); d_sgrad.reserve_keep(stride);
        d_act_in.reserve_keep(astride); d_act_rep.reserve_keep(astride * NC);
        if (m) { d_beta_in.upload(screen_beta.data(), m); d_sgrad.upload(screen_grad.data(), m); }
        if (S) d_act_in.upload(screen_is_active.data(), S);
        std::vector<int32_t> act32(active_set_size);
        for (size_t i = 0; i < active_set_size; ++i) act32[i] = (int32_t)active_set[i];
        if (active_set_size) d_active_set.upload(act32.data(), active_set_size);
        PinScalars sc{};
        sc.rsq = (double)rsq; sc.active_set_size = (int)active_set_size;
        *h_sc.p = sc;
        d_sc.upload(h_sc.p, 1);
        CovKernelArgs<T> a{};
        a.gram = d_gram.p; a.ldg = (int64_t)m; a.meta = d_meta.p; a.S = (int)S; a.grec = d_grec.p;
        a.beta_in = d_beta_in.p; a.beta_rep = d_beta_rep.p; a.beta_old_rep = d_beta_old.p; a.beta_stride = (int64_t)stride;
        a.sgrad = d_sgrad.p; a.is_active_in = d_act_in.p; a.act_rep = d_act_rep.p; a.act_stride = (int64_t)astride;
        a.active_set = d_active_set.p; a.sc = d_sc.p; a.m = (int)m; a.L = L;
        a.lmda = (double)lmda_; a.alpha = (double)alpha; a.tol = (double)tol; a.newton_tol = (double)newton_tol; a.dbeta_tol = Configs::dbeta_tol;
        a.max_iters = (long long)max_iters_left; a.newton_max_iters = (int)std::min<size_t>(newton_max_iters, (size_t)1 << 30);
        a.max_active_size = (int)std::min<size_t>(max_active_size, (size_t)G); a.gs_cap = gs_cap; a.rec_cap = rec_cap;
        if (Configs::sweep_profile && d_stats.n == 0) d_stats.alloc(8);
        a.stats = Configs::sweep_profile ? d_stats.p : nullptr;
        AB_CUDA(cudaFuncSetAttribute(cov_pin_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(NC); cfg.blockDim = dim3(kCovThreads); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        AB_CUDA(cudaEventRecord(ev0, 0));
        AB_CUDA(cudaLaunchKernelEx(&cfg, cov_pin_kernel<T>, a));
        AB_CUDA(cudaEventRecord(ev1, 0));
        d_sc.download(h_sc.p, 1);
        if (m) { d_beta_rep.download(screen_beta.data(), m); d_sgrad.download(screen_grad.data(), m); }
        if (S) d_act_rep.download(screen_is_active.data(), S);
        AB_CUDA(cudaStreamSynchronize(0));
        float ms = 0; AB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        ++n_kernel_launches; ++n_pin_solves; time_sweep_kernel += ms * 1e-3;
        sc = *h_sc.p;
        n_sweeps += sc.iters; n_group_updates += sc.n_group_updates; n_col_updates += sc.n_col_updates;
        if (sc.error) {
            if (sc.error == kErrMaxCds) throw solver_error("max coordinate descents reached at lambda index: " + std::to_string(lmda_index) + ".");
            if (sc.error == kErrMaxActive) throw solver_error("Maximum number of active groups reached.");
            if (sc.error == kErrNewton) throw solver_error("Newton-ABS max iterations reached! Try increasing newton_max_iters.");
            throw solver_error("unknown device error.");
        }
        const size_t old_active = active_set_size;
        active_set_size = (size_t)sc.active_set_size;
        if (active_set_size > old_active) {
            std::vector<int32_t> nw(active_set_size - old_active);
            d_active_set.download(nw.data(), nw.size(), old_active);
            AB_CUDA(cudaStreamSynchronize(0));
            for (size_t i = 0; i < nw.size(); ++i) active_set[old_active + i] = nw[i];
        }
        rsq = (T)sc.rsq;
        R.iters = sc.iters; R.rsq = sc.rsq; R.active_time = ms * 1e-3; R.screen_time = 0; R.intercept = 0;
        std::vector<size_t> order(active_set_size);                             // active_order + sparsify_active_beta (:680-716)
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](size_t i, size_t j) { return groups[screen_set[active_set[i]]] < groups[screen_set[active_set[j]]]; });
        for (size_t i = 0; i < order.size(); ++i) {
            const idx_t ss = active_set[order[i]], g = screen_set[ss], gs = group_sizes[g];
            for (idx_t c = 0; c < gs; ++c) { R.beta.idx.push_back(groups[g] + c); R.beta.val.push_back((double)screen_beta[screen_begins[ss] + c]); }
        }
        return R;
    }

    // fit (solver_gaussian_cov.hpp:234-357): save / restore the quantities the pin solve modifies in place
    PinResult fit(T lmda_) {
        std::vector<T> grad_prev = screen_grad, beta_prev = screen_beta;
        std::vector<int8_t> act_prev = screen_is_active;
        const size_t act_size_prev = active_set_size; const T rsq_prev = rsq;
        try { return run_pin(lmda_, max_iters, 0); }
        catch (...) {
            screen_grad.swap(grad_prev); screen_beta.swap(beta_prev); screen_is_active.swap(act_prev);
            active_set_size = act_size_prev; rsq = rsq_prev;
            throw;
        }
    }
    // update_invariance (:376-402): grad = v - A beta, abs_grad
    void update_invariance(const PinResult& pr, T lmda_) {
        AB_TIME(timers, "invariance");
        lmda = lmda_;
        std::vector<T> vals(pr.beta.val.size());
        for (size_t k = 0; k < vals.size(); ++k) vals[k] = (T)pr.beta.val[k];
        A->d_mul(pr.beta.idx.data(), vals.data(), (int64_t)vals.size(), d_v.p, d_grad.p);
        d_grad.download(grad.data(), p);
        AB_CUDA(cudaStreamSynchronize(0));
        ++n_kernel_launches;
        update_abs_grad(lmda_);
    }
    void update_solutions(PinResult& pr, T lmda_) {                             // :205-232
        betas.emplace_back(std::move(pr.beta));
        intercepts.push_back(0);
        lmdas.push_back(lmda_);
        devs.push_back((T)pr.rsq);
    }
    // screen (solver_base.hpp:273-403)
    void screen(T lmda_next, bool all_kkt_passed, int n_new_active) {
        const int old_size = (int)screen_set.size();
        auto is_screen = [&](idx_t i) { return in_screen[i] != 0; };
        if (screen_rule == 0) {
            const T strong = (2 * lmda_next - lmda) * alpha;
            for (idx_t i = 0; i < G; ++i) { if (is_screen(i)) continue; if (abs_grad[i] > strong * penalty[i]) screen_set.push_back(i); }
        } else if (screen_rule == 1) {
            if (n_new_active) {
                std::vector<T> wts(G);
                for (idx_t i = 0; i < G; ++i) wts[i] = (penalty[i] <= 0) ? alpha * lmda : std::min(abs_grad[i] / penalty[i], alpha * lmda);
                std::vector<idx_t> order(G);
                std::iota(order.begin(), order.end(), 0);
                std::sort(order.begin(), order.end(), [&](idx_t i, idx_t j) { return wts[i] < wts[j]; });
                const int subset_size = std::min<int>(std::max<int>((int)(old_size * (1 + pivot_subset_ratio)), (int)pivot_subset_min), (int)G);
                std::vector<T> ws(subset_size), mses(subset_size), ind(subset_size);
                for (int i = 0; i < subset_size; ++i) { ws[i] = wts[order[G - subset_size + i]]; ind[i] = (T)i; }
                const int pivot_idx = search_pivot(ind, ws, mses);
                const int full_pivot_idx = (int)G - subset_size + pivot_idx;
                for (int ii = (int)G - 1; ii >= full_pivot_idx; --ii) { const idx_t i = order[ii]; if (is_screen(i)) continue; screen_set.push_back(i); }
                int count = 0;
                for (int ii = full_pivot_idx - 1; ii >= 0; --ii) {
                    if (count >= pivot_slack_ratio * n_new_active) break;
                    const idx_t i = order[ii];
                    if (is_screen(i)) continue;
                    screen_set.push_back(i); ++count;
                }
            }
            if (((int)screen_set.size() == old_size) && !all_kkt_passed) {
                for (idx_t i = 0; i < G; ++i) { if (is_screen(i)) continue; if (abs_grad[i] > lmda_next * penalty[i] * alpha) screen_set.push_back(i); }
            }
        } else throw solver_error("Unknown screen rule!");
        if (screen_set.size() > max_screen_size) { screen_set.resize(old_size); throw solver_error("maximum screen set size reached."); }
    }
    bool kkt(T lmda_) {                                                         // solver_base.hpp:408-433
        for (idx_t k = 0; k < G; ++k) { if (in_screen[k]) continue; if (abs_grad[k] > lmda_ * alpha * penalty[k]) return false; }
        return true;
    }
    bool early_exit_f() {                                                       // cov::early_exit (:186-203) + user exit_cond
        bool r = false;
        if (early_exit && devs.size() >= 2) {
            const T u = devs.back(), mm = devs[devs.size() - 2];
            if (u - mm <= rdev_tol * u) r = true;
        }
        return r || (exit_cond && exit_cond());
    }
    void screen_f(T lmda_, bool kkt_passed, int n_new_active) {
        screen(lmda_, kkt_passed, n_new_active);
        update_screen_derived();
    }

    // StateGaussianPinCov::solve (pin::cov::solve, :529-725) over the state's lmda_path on the FIXED screen set; screen_grad is an input
    void solve_pin() {
        const size_t max_iters_total = max_iters;
        for (size_t l = 0; l < lmda_path.size(); ++l) {
            const size_t left = (size_t)n_sweeps >= max_iters_total ? 0 : max_iters_total - (size_t)n_sweeps;
            PinResult pr = run_pin(lmda_path[l], left, l);
            betas.emplace_back(std::move(pr.beta));
            intercepts.push_back(0);
            rsqs.push_back(rsq);
            lmdas.push_back(lmda_path[l]);
            benchmark_fit_screen.push_back(pr.screen_time);
            benchmark_fit_active.push_back(pr.active_time);
            lmda = lmda_path[l];
            if (l >= 1 && rsqs[l] - rsqs[l - 1] <= rdev_tol * rsqs[l]) break;     // :724
        }
    }

    // solve_core (solver_base.hpp:435-687)
    void solve() {
        if (screen_set.size() > max_screen_size) throw solver_error("maximum screen set size reached.");
        if (setup_lmda_max) {
            T pmax = penalty[0];
            for (auto q : penalty) pmax = std::max(pmax, q);
            const T large_lmda = T(1e-3 * std::numeric_limits<T>::max() / std::max<T>(1, pmax));
            PinResult pr = fit(large_lmda);
            update_invariance(pr, large_lmda);
            const T factor = (alpha <= 0) ? T(1e-3) : alpha;                    // solver/utils.hpp:6-23
            T mx = -std::numeric_limits<T>::infinity();
            for (idx_t i = 0; i < G; ++i) mx = std::max<T>(mx, (penalty[i] <= 0.0) ? T(0.0) : abs_grad[i] / penalty[i]);
            lmda_max = mx / factor;
        }
        if (setup_lmda_path) {
            if (lmda_path_size <= 0) return;
            lmda_path.resize(lmda_path_size);
            const size_t Lp = lmda_path_size;
            if (Lp > 1) {                                                       // solver/utils.hpp:25-41
                const T log_factor = std::log(min_ratio) / (Lp - 1);
                for (size_t i = 0; i < Lp; ++i) lmda_path[i] = lmda_max * std::exp(log_factor * T(i));
            }
            lmda_path[0] = lmda_max;
        }
        size_t large_sz = 0;
        while (large_sz < lmda_path.size() && !(lmda_path[large_sz] <= lmda_max)) ++large_sz;
        if (large_sz || setup_lmda_max) {
            std::vector<T> large(lmda_path.begin(), lmda_path.begin() + large_sz);
            large.push_back(lmda_max);
            for (size_t i = 0; i < large.size(); ++i) {
                PinResult pr = fit(large[i]);
                if (i + 1 < large.size()) { update_solutions(pr, large[i]); if (early_exit_f()) return; }
                else update_invariance(pr, large[i]);
            }
        }
        size_t idx = large_sz;
        int current_active = (int)active_set_size;
        bool kkt_passed = true;
        int n_new_active = 0;
        while (idx < lmda_path.size()) {
            const T lmda_curr = lmda_path[idx];
            while (1) {
                double t0 = now_s();
                screen_f(lmda_curr, kkt_passed, n_new_active);
                benchmark_screen.push_back(now_s() - t0);
                PinResult pr = fit(lmda_curr);
                benchmark_fit_screen.push_back(pr.screen_time);
                benchmark_fit_active.push_back(pr.active_time);
                t0 = now_s();
                update_invariance(pr, lmda_curr);
                benchmark_invariance.push_back(now_s() - t0);
                t0 = now_s();
                kkt_passed = kkt(lmda_curr);
                n_valid_solutions.push_back(kkt_passed);
                idx += kkt_passed;
                if (kkt_passed) update_solutions(pr, lmda_curr);
                benchmark_kkt.push_back(now_s() - t0);
                if (kkt_passed) { active_sizes.push_back((int)active_set_size); screen_sizes.push_back((int)screen_set.size()); }
                n_new_active = kkt_passed ? (active_sizes.back() - current_active) : n_new_active;
                current_active = kkt_passed ? active_sizes.back() : current_active;
                if (kkt_passed) break;
            }
            if (early_exit_f()) break;
        }
    }
};

} // namespace ab
