// adelie_b200/csrc/sparse_kernels.cuh -- device kernels for the sparse CSC matrix (reference: MatrixNaiveSparse,
// CORE/matrix/matrix_naive_sparse.ipp:10-262: spddot / spaxi per column, `mul` parallel over columns :172-184) and the
// pin solve over sparse columns.
//
// HBM layout: the three CSC arrays as they are (int64 column pointers, int32 row indices sorted inside every column, values).
// The row vectors (residual, weights) are dense, padded, and small enough (n = 2M: 8 MB) to live in L2, so every column
// operation is one coalesced stream over (value, index) = nnz * (s + 4) bytes from HBM plus gathers / scatters that hit L2.
#pragma once
#include "common.cuh"
#include "device_prims.cuh"
#include "sweep.cuh"

namespace ab {

template <class T>
struct CscView { const int64_t* indptr; const int32_t* indices; const T* values; };

// vw[i] = v[i] * w[i]  (SQ: w[i]): one dense pass so that the column kernel gathers a single vector
template <class T, bool SQ>
__global__ void sp_vw_kernel(const T* __restrict__ v, const T* __restrict__ w, T* __restrict__ vw, int64_t n_pad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) vw[i] = SQ ? w[i] : v[i] * w[i];
}

// out[c] = sum_k X[k, col(c)]^(1 or 2) * vw[k]  [- scale * sub[c]]; one warp per column, 8 independent (index, value) loads and
// gathers in flight per lane (the gathers hit L2: vw is n * s bytes)
template <class T, bool SQ>
__global__ void __launch_bounds__(256)
spmv_t_kernel(CscView<T> X, int64_t j0, const int32_t* __restrict__ cols, int q, const T* __restrict__ vw,
              T* __restrict__ out, const T* __restrict__ sub, const double* __restrict__ sub_scale_ptr, double sub_scale)
{
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= q) return;
    const int64_t col = cols ? (int64_t)cols[c] : j0 + c;
    const int64_t k0 = X.indptr[col], k1 = X.indptr[col + 1];
    constexpr int U = 8;
    double acc = 0;
    for (int64_t k = k0 + lane; k < k1; k += 32 * U) {
        int32_t idx[U]; T x[U]; T g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int64_t kk = k + 32 * u; const bool in = kk < k1; idx[u] = in ? X.indices[kk] : 0; x[u] = in ? X.values[kk] : T(0); }
#pragma unroll
        for (int u = 0; u < U; ++u) g[u] = vw[idx[u]];
        T part = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) part += (SQ ? x[u] * x[u] : x[u]) * g[u];
        acc += (double)part;
    }
    acc = dev::warp_sum(acc);
    if (lane == 0) {
        if (sub) acc -= (sub_scale_ptr ? *sub_scale_ptr : sub_scale) * (double)sub[c];
        out[c] = (T)acc;
    }
}

// out[i] += sum_c X[i, j0 + c] v[c]  (btmul / ctmul, increment).  grid.y = column; rows of different columns may coincide.
template <class T>
__global__ void __launch_bounds__(256)
spaxpy_kernel(CscView<T> X, int64_t j0, const T* __restrict__ v, T* __restrict__ out)
{
    const int64_t col = j0 + blockIdx.y;
    const T vc = v[blockIdx.y];
    const int64_t k0 = X.indptr[col], k1 = X.indptr[col + 1];
    for (int64_t k = k0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < k1; k += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(out + X.indices[k], X.values[k] * vc);
}

// Batched weighted Gram of column groups: C[out_off + a*gs + b] = sum_i X[i,col+a] X[i,col+b] w_i (w or w^2); one CTA per item.
// Pairs of sparse columns are joined by binary search (indices are sorted inside a column).
template <class T>
__global__ void __launch_bounds__(256)
spcov_kernel(CscView<T> X, const CovItem* __restrict__ items, const T* __restrict__ w, int w_is_sqrt, double* __restrict__ C_out)
{
    __shared__ double s_red[8];
    const CovItem it = items[blockIdx.x];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* Cg = C_out + it.out_off;
    for (int a = 0; a < it.gs; ++a) {
        const int64_t a0 = X.indptr[it.col + a], a1 = X.indptr[it.col + a + 1];
        for (int b = 0; b <= a; ++b) {
            const int64_t b0 = X.indptr[it.col + b], b1 = X.indptr[it.col + b + 1];
            double acc = 0;
            for (int64_t k = a0 + tid; k < a1; k += 256) {
                const int32_t i = X.indices[k];
                T xb = 0;
                if (a == b) xb = X.values[k];
                else {
                    int64_t lo = b0, hi = b1;                       // first position with index >= i
                    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (X.indices[mid] < i) lo = mid + 1; else hi = mid; }
                    if (lo < b1 && X.indices[lo] == i) xb = X.values[lo];
                }
                if (xb != T(0)) {
                    T wi = w[i];
                    if (w_is_sqrt) wi = wi * wi;
                    acc += (double)(X.values[k] * wi) * (double)xb;
                }
            }
            acc = dev::warp_sum(acc);
            if (lane == 0) s_red[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                double s = 0;
                for (int wi = 0; wi < 8; ++wi) s += s_red[wi];
                Cg[a * it.gs + b] = s; Cg[b * it.gs + a] = s;
            }
            __syncthreads();
        }
    }
}

// Random sparse matrix generated in HBM (bench): every column gets exactly m non-zeros at stratified random rows
// (row = stratum start + uniform offset: sorted, unique), values ~ N(0, 1) (Box-Muller on a counter hash).
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}
template <class T>
__global__ void sparse_fill_kernel(int64_t* indptr, int32_t* indices, T* values, int64_t n, int64_t p, int64_t m, uint64_t seed) {
    const int64_t total = p * m;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e % m;
        const int64_t lo = k * n / m, hi = (k + 1) * n / m;
        const uint64_t h1 = splitmix64(seed ^ (uint64_t)e * 0x2545f4914f6cdd1dull), h2 = splitmix64(h1);
        indices[e] = (int32_t)(lo + (int64_t)(h1 % (uint64_t)max((long long)1, (long long)(hi - lo))));
        const double u1 = ((h2 >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (splitmix64(h2) >> 11) * (1.0 / 9007199254740992.0);
        values[e] = (T)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
    }
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= p; j += (int64_t)gridDim.x * blockDim.x) indptr[j] = j * m;
}

// ---------------------------------------------------------------------------------------------------------------
// Pin solve over sparse columns: the same algorithm and control flow as pin_solve_kernel (sweep.cuh; reference
// solver_gaussian_pin_naive.hpp:26-168, 181-215, 223-401) in ONE persistent CTA.  A sparse column is a few thousand
// non-zeros: far too little to spread over the grid without putting a grid-wide exchange back into the Gauss-Seidel chain,
// so a single CTA streams the column (coalesced value / index loads, gathers of w o r from L2), reduces in shared memory,
// solves the proximal problem in its control warp and scatters the residual update -- no inter-CTA traffic at all.
template <class T>
struct SparsePinArgs {
    CscView<T> X;
    T* resid; const T* weights;
    const GroupMeta* meta; int S; const T* grec;
    const T* beta_in; T* beta_out; int beta_len;
    const int8_t* is_active_in; int8_t* is_active_out;
    int32_t* active_set;
    PinScalars* sc;
    double lmda, alpha, tol, newton_tol, dbeta_tol;
    long long max_iters; int newton_max_iters; int max_active_size; int intercept;
    int gs_cap;
};

constexpr int kSparseThreads = 1024;

template <class T>
__global__ void __launch_bounds__(kSparseThreads, 1)
pin_solve_sparse_kernel(const __grid_constant__ SparsePinArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using P = T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = kSparseThreads / 32;
    const int gsc = a.gs_cap;
    double* gsum = reinterpret_cast<double*>(smem_raw);                 // [gsc]
    double* wred = gsum + gsc;                                          // [NW][gsc]
    P* px = reinterpret_cast<P*>(wred + (size_t)NW * gsc);              // 8 x [gsc] prox scratch (double-sized slots)
    P* p_aold = px, *p_A = px + gsc, *p_gk = px + 2 * gsc, *p_gt = px + 3 * gsc, *p_atold = px + 4 * gsc, *p_at = px + 5 * gsc;
    T* s_del = reinterpret_cast<T*>(reinterpret_cast<double*>(px) + 8 * gsc);
    P* p_scr = reinterpret_cast<P*>(reinterpret_cast<double*>(px) + 9 * gsc);     // [4 * 32]
    int* ctrl = reinterpret_cast<int*>(reinterpret_cast<double*>(px) + 9 * gsc + 4 * 32);   // changed, error, next

    T* my_beta = a.beta_out;
    int8_t* my_active = a.is_active_out;
    for (int i = tid; i < a.beta_len; i += kSparseThreads) my_beta[i] = a.beta_in[i];
    for (int i = tid; i < a.S; i += kSparseThreads) my_active[i] = a.is_active_in[i];
    __syncthreads();
    const ProxCtx<T, P> proxctx{p_aold, p_A, p_gk, p_gt, p_atold, p_at, p_scr, s_del, gsum, my_beta};
    const P l1 = (P)(a.lmda * a.alpha), l2 = (P)(a.lmda * (1.0 - a.alpha));
    ProxState ps;
    ps.rsq = a.sc->rsq; ps.resid_sum = a.sc->resid_sum; ps.cm = 0; ps.A = a.sc->active_set_size; ps.error = 0; ps.newton_iters_max = 0;
    long long iters = a.sc->iters, n_updates = a.sc->n_group_updates, n_cols = a.sc->n_col_updates;
    int phase = kSweepActive, final_error = 0;

    while (true) {
        const int kind = phase;
        const int count = (kind == kSweepActive) ? ps.A : a.S;           // ps.A is identical in every thread (uniform updates below)
        ++iters;
        ps.cm = 0;
        for (int it = 0; it < count; ++it) {
            const int ss = (kind == kSweepActive) ? a.active_set[it] : it;
            const GroupMeta m = a.meta[ss];
            const int gs = m.gs;
            const T* rec = a.grec + m.rec_off;
            if (warp == 0) for (int c = lane; c < gs; c += 32) p_aold[c] = (P)my_beta[m.begin + c];
            // ---- gradient: gsum[c] = sum_k X[k, col + c] w_k r_k
            for (int c = 0; c < gs; ++c) {
                const int64_t k0 = a.X.indptr[m.col + c], k1 = a.X.indptr[m.col + c + 1];
                T acc = 0;
                constexpr int U = 6;                                     // independent (index, value) loads + gathers in flight per thread
                for (int64_t k = k0 + tid; k < k1; k += (int64_t)kSparseThreads * U) {
                    int32_t idx[U]; T x[U], gw[U], gr[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) { const int64_t kk = k + (int64_t)kSparseThreads * u; const bool in = kk < k1; idx[u] = in ? a.X.indices[kk] : 0; x[u] = in ? a.X.values[kk] : T(0); }
#pragma unroll
                    for (int u = 0; u < U; ++u) { gw[u] = a.weights[idx[u]]; gr[u] = a.resid[idx[u]]; }
#pragma unroll
                    for (int u = 0; u < U; ++u) acc += x[u] * (gw[u] * gr[u]);
                }
                const double tot = dev::warp_sum((double)acc);
                if (lane == 0) wred[(size_t)warp * gsc + c] = tot;
            }
            __syncthreads();
            if (tid < gs) {
                double s = 0;
#pragma unroll 4
                for (int w = 0; w < NW; ++w) s += wred[(size_t)w * gsc + tid];
                gsum[tid] = s;
            }
            __syncthreads();
            // ---- proximal update (control warp); every thread then reads the outcome
            if (warp == 0) {
                int changed = 0;
                const P pk = (P)m.pen;
                if (gs == 1) {                                           // solver_gaussian_pin_naive.hpp:75-108
                    const P ak_old = p_aold[0];
                    const P A_kk = (P)rec[0], xm = (P)rec[1];
                    P gk = (P)gsum[0] - xm * (P)ps.resid_sum * (P)a.intercept + ak_old * A_kk;
                    const P vv = fabs(gk) - l1 * pk;                     // update_coordinate, pin_base.hpp:181-195
                    P ak = (vv > P(0)) ? copysign(vv, gk) / (A_kk + l2 * pk) : P(0);
                    ak = (P)(T)ak;
                    gk -= ak_old * A_kk;
                    if (ak != ak_old) {
                        const P del = ak - ak_old;
                        ps.cm = fmax(ps.cm, (double)(A_kk * del * del));
                        ps.rsq += (double)(del * (2 * gk - del * A_kk));
                        ps.resid_sum -= (double)(xm * del);
                        if (lane == 0) { my_beta[m.begin] = (T)ak; s_del[0] = (T)(-del); }
                        changed = 1;
                    }
                } else if (gs <= 32) {
                    const ProxPre<P> pre = prox_small_pre<T, P>(rec, gs, p_aold, lane);
                    changed = prox_small_post<T, P>(proxctx, pre, rec, gs, m.begin, l1 * pk, l2 * pk, (P)a.newton_tol, a.newton_max_iters,
                                                    (P)a.dbeta_tol, a.intercept, ps, lane, nullptr);
                } else {
                    changed = prox_group<T, P>(proxctx, rec, gs, m.begin, l1 * pk, l2 * pk, (P)a.newton_tol, a.newton_max_iters,
                                               (P)a.dbeta_tol, a.intercept, ps, lane, nullptr);
                }
                int new_active = 0;
                if (changed && kind == kSweepScreen && !my_active[ss]) {  // add_active_set (:294-304)
                    if (ps.A >= a.max_active_size) ps.error = kErrMaxActive;
                    else { if (lane == 0) { my_active[ss] = 1; a.active_set[ps.A] = ss; } new_active = 1; }
                }
                if (lane == 0) {
                    ctrl[0] = changed; ctrl[1] = ps.error; ctrl[2] = new_active;
                    // scalars travel through shared memory so that all warps stay uniform
                    reinterpret_cast<double*>(ctrl + 4)[0] = ps.rsq; reinterpret_cast<double*>(ctrl + 4)[1] = ps.resid_sum;
                    reinterpret_cast<double*>(ctrl + 4)[2] = ps.cm;
                    ctrl[3] = ps.newton_iters_max;
                }
            }
            __syncthreads();
            const int changed = ctrl[0];
            ps.error = ctrl[1]; ps.A += ctrl[2];
            if (warp != 0) {
                ps.rsq = reinterpret_cast<double*>(ctrl + 4)[0]; ps.resid_sum = reinterpret_cast<double*>(ctrl + 4)[1];
                ps.cm = reinterpret_cast<double*>(ctrl + 4)[2]; ps.newton_iters_max = ctrl[3];
            }
            ++n_updates; n_cols += gs;
            // ---- residual update r += X_g del (column by column: rows of different columns may coincide)
            if (changed) {
                for (int c = 0; c < gs; ++c) {
                    const T d = s_del[c];
                    const int64_t k0 = a.X.indptr[m.col + c], k1 = a.X.indptr[m.col + c + 1];
                    constexpr int U = 6;
                    for (int64_t k = k0 + tid; k < k1; k += (int64_t)kSparseThreads * U) {
                        int32_t idx[U]; T x[U], gr[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) { const int64_t kk = k + (int64_t)kSparseThreads * u; const bool in = kk < k1; idx[u] = in ? a.X.indices[kk] : -1; x[u] = in ? a.X.values[kk] : T(0); }
#pragma unroll
                        for (int u = 0; u < U; ++u) gr[u] = idx[u] >= 0 ? a.resid[idx[u]] : T(0);
#pragma unroll
                        for (int u = 0; u < U; ++u) if (idx[u] >= 0) a.resid[idx[u]] = gr[u] + x[u] * d;
                    }
                    if (gs > 1) __syncthreads();
                }
            }
            __syncthreads();
            if (ps.error) { final_error = ps.error; break; }
        }
        if (final_error) break;
        const bool conv = ps.cm < a.tol;
        int next;
        if (kind == kSweepActive) next = conv ? kSweepScreen : ((iters >= a.max_iters) ? -kErrMaxCds : kSweepActive);
        else next = conv ? kSweepExit : ((iters >= a.max_iters) ? -kErrMaxCds : kSweepActive);
        if (next < 0) { final_error = -next; break; }
        if (next == kSweepExit) break;
        phase = next;
    }
    if (tid == 0) {
        a.sc->rsq = ps.rsq; a.sc->resid_sum = ps.resid_sum; a.sc->active_set_size = ps.A;
        a.sc->iters = iters; a.sc->n_group_updates = n_updates; a.sc->n_col_updates = n_cols; a.sc->error = final_error;
        a.sc->newton_iters_max = ps.newton_iters_max;
    }
}

} // namespace ab
