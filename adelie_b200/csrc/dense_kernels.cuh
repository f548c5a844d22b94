// adelie_b200/csrc/dense_kernels.cuh -- coalesced CUDA kernels for the dense column-major matrix
// operators of the reference (CORE/matrix/matrix_naive_dense.ipp): mul / bmul / cmul (:25-148),
// btmul / ctmul (:51-125), cov (:164-199), sq_mul (:201-221).  All HBM-bound streaming kernels:
// 16-byte vector loads along the contiguous (row) dimension, warp-shuffle reductions, and
// deterministic two-phase reductions (no floating-point atomics).
#pragma once
#include "common.cuh"
#include "device_prims.cuh"
#include "sweep.cuh"   // VecT / vec_load

namespace ab {

constexpr int kGemvRows = 4096;      // rows per CTA tile of the transposed GEMV
constexpr int kGemvThreads = 256;
constexpr int kGemvColsPerCta = 64;

// out_part[rb * q + c] = sum_{i in row block rb} X[i, j0+c] * f(v[i] * w[i])      (SQ: X^2 * w)
// grid = (ceil(q / kGemvColsPerCta), n_row_blocks)
template <class T, bool SQ>
__global__ void __launch_bounds__(kGemvThreads)
gemv_t_kernel(const T* __restrict__ X, int64_t ld, int64_t n_pad, int64_t j0, const int32_t* __restrict__ cols, int q,
              const T* __restrict__ v, const T* __restrict__ w, double* __restrict__ out_part)
{
    constexpr int VN = VecT<T>::N;
    __shared__ __align__(16) T s_vw[kGemvRows];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t rb = blockIdx.y;
    const int64_t row0 = rb * kGemvRows;
    const int rows = (int)min((long long)kGemvRows, (long long)(n_pad - row0));      // multiple of kRowAlign
    for (int i = tid * VN; i < rows; i += kGemvThreads * VN) {
        T a[VN], b[VN];
        vec_load<T>(w + row0 + i, b);
        if (SQ) {
#pragma unroll
            for (int k = 0; k < VN; ++k) a[k] = b[k];
        } else {
            vec_load<T>(v + row0 + i, a);
#pragma unroll
            for (int k = 0; k < VN; ++k) a[k] *= b[k];
        }
        vec_store<T>(s_vw + i, a);
    }
    __syncthreads();
    const int c_begin = blockIdx.x * kGemvColsPerCta;
    const int c_end = min(q, c_begin + kGemvColsPerCta);
    for (int c = c_begin + warp; c < c_end; c += kGemvThreads / 32) {
        const T* col = X + (cols ? (int64_t)cols[c] : (j0 + c)) * ld + row0;
        T acc[4] = {0, 0, 0, 0};
        int i = lane * VN;
        // 4 independent 16-byte loads in flight per lane
        for (; i + 3 * 32 * VN < rows; i += 4 * 32 * VN) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                T x[VN], s[VN];
                vec_load<T>(col + i + u * 32 * VN, x);
                vec_load<T>(s_vw + i + u * 32 * VN, s);
#pragma unroll
                for (int k = 0; k < VN; ++k) acc[u] += (SQ ? x[k] * x[k] : x[k]) * s[k];
            }
        }
        for (; i < rows; i += 32 * VN) {
            T x[VN], s[VN];
            vec_load<T>(col + i, x);
            vec_load<T>(s_vw + i, s);
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[0] += (SQ ? x[k] * x[k] : x[k]) * s[k];
        }
        const double tot = dev::warp_sum((double)acc[0] + (double)acc[1] + (double)acc[2] + (double)acc[3]);
        if (lane == 0) out_part[rb * q + c] = tot;
    }
}

// Multi-response transposed GEMV (the `mul` of kron(X, I_K), CORE/matrix/matrix_naive_kronecker_eye.ipp): v, w are (n, K) row-major,
//   out_part[rb * q*K + c*K + l] = sum_{i in row block rb} X[i, j0+c] * v[i, l] * w[i, l]      (w == nullptr: weights 1)
// Each X element is read from HBM once and used for all K classes; the K products per row are staged class-major in shared memory.
// grid = (ceil(q / kGemvColsPerCta), n_row_blocks), dynamic smem = K * tile_rows * sizeof(T); tile_rows multiple of kRowAlign.
constexpr int kMultiMaxK = 16;
template <class T>
__global__ void __launch_bounds__(kGemvThreads)
gemv_t_multi_kernel(const T* __restrict__ X, int64_t ld, int64_t n_pad, int64_t j0, int q, int K, int tile_rows,
                    const T* __restrict__ v, const T* __restrict__ w, double* __restrict__ out_part)
{
    constexpr int VN = VecT<T>::N;
    extern __shared__ __align__(16) unsigned char s_raw[];
    T* s_vw = reinterpret_cast<T*>(s_raw);                       // [K][tile_rows]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t rb = blockIdx.y;
    const int64_t row0 = rb * tile_rows;
    const int rows = (int)min((long long)tile_rows, (long long)(n_pad - row0));
    for (int e = tid; e < rows * K; e += kGemvThreads) {
        const int i = e / K, l = e - i * K;
        const T vv = v[row0 * K + e];
        s_vw[l * tile_rows + i] = w ? vv * w[row0 * K + e] : vv;
    }
    __syncthreads();
    const int c_begin = blockIdx.x * kGemvColsPerCta;
    const int c_end = min(q, c_begin + kGemvColsPerCta);
    for (int c = c_begin + warp; c < c_end; c += kGemvThreads / 32) {
        const T* col = X + (j0 + c) * ld + row0;
        T acc[kMultiMaxK];
#pragma unroll
        for (int l = 0; l < kMultiMaxK; ++l) acc[l] = 0;
        for (int i = lane * VN; i < rows; i += 32 * VN) {
            T x[VN];
            vec_load<T>(col + i, x);
#pragma unroll
            for (int l = 0; l < kMultiMaxK; ++l) {
                if (l < K) {
                    T s[VN];
                    vec_load<T>(s_vw + l * tile_rows + i, s);
#pragma unroll
                    for (int k = 0; k < VN; ++k) acc[l] += x[k] * s[k];
                }
            }
        }
#pragma unroll
        for (int l = 0; l < kMultiMaxK; ++l) {
            if (l < K) {
                const double tot = dev::warp_sum((double)acc[l]);
                if (lane == 0) out_part[(size_t)rb * q * K + (size_t)c * K + l] = tot;
            }
        }
    }
}

// out[c] = sum_rb part[rb*q + c]  (- scale * sub[c] if sub != nullptr), fixed summation order
template <class T>
__global__ void gemv_t_reduce_kernel(const double* __restrict__ part, int n_rb, int q, T* __restrict__ out,
                                     const T* __restrict__ sub, const double* __restrict__ sub_scale_ptr, double sub_scale)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= q) return;
    double s = 0;
    for (int rb = 0; rb < n_rb; ++rb) s += part[(size_t)rb * q + c];
    if (sub) s -= (sub_scale_ptr ? *sub_scale_ptr : sub_scale) * (double)sub[c];
    out[c] = (T)s;
}

// out[i] += sum_c X[i, j0+c] * v[c]   (btmul / ctmul, increment semantics)
template <class T>
__global__ void __launch_bounds__(256)
axpy_cols_kernel(const T* __restrict__ X, int64_t ld, int64_t n_pad, int64_t j0, int q, const T* __restrict__ v, T* __restrict__ out)
{
    constexpr int VN = VecT<T>::N;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VN;
    if (i >= n_pad) return;
    T o[VN];
    vec_load<T>(out + i, o);
    for (int c = 0; c < q; ++c) {
        const T vc = v[c];
        T x[VN];
        vec_load<T>(X + (j0 + c) * ld + i, x);
#pragma unroll
        for (int k = 0; k < VN; ++k) o[k] += x[k] * vc;
    }
    vec_store<T>(out + i, o);
}

// Batched weighted Gram of screen groups:
//   C_part[rb][out_off + a*gs+b] = sum_{i in rb} X[i,col+a] X[i,col+b] w[i]
// grid = (n_groups, n_row_blocks); each CTA streams its (rows x gs) tile once per 4x4 pair block
// (re-reads hit L1/L2), so HBM traffic is one pass over X_g.
struct CovItem { int32_t col, gs; int64_t out_off; int32_t cls, pad; };   // pad: 1 + offset of the group's weighted column sums in the optional means output (0: none)   // out_off: element offset of this group's gs*gs block; col < 0: the column of ones
                                                                          // cls: class whose weights w[i*K + cls] are used (multi-response)

template <class T>
__global__ void __launch_bounds__(256)
cov_kernel(const T* __restrict__ X, int64_t ld, int64_t n_pad, const CovItem* __restrict__ items,
           const T* __restrict__ w, int w_is_sqrt, double* __restrict__ C_part, int64_t c_total, int rows_per_block, int K)
{
    __shared__ double s_red[8][17];
    const CovItem it = items[blockIdx.x];
    const int gs = it.gs;
    const int rb = blockIdx.y;
    const int64_t row0 = (int64_t)rb * rows_per_block;
    const int64_t row1 = min((long long)n_pad, (long long)(row0 + rows_per_block));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool ones = it.col < 0;
    const T* Xg = X + (int64_t)(ones ? 0 : it.col) * ld;
    double* Cout = C_part + (size_t)rb * c_total + it.out_off;
    for (int a0 = 0; a0 < gs; a0 += 4) {
        for (int b0 = 0; b0 <= a0; b0 += 4) {
            double acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = 0;
            for (int64_t i = row0 + tid; i < row1; i += 256) {
                T wi = w[i * K + it.cls];
                if (w_is_sqrt) wi = wi * wi;
                T xa[4], xb[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) xa[x] = (a0 + x < gs) ? (ones ? T(1) : Xg[(int64_t)(a0 + x) * ld + i]) : T(0);
#pragma unroll
                for (int y = 0; y < 4; ++y) xb[y] = (b0 + y < gs) ? (ones ? T(1) : Xg[(int64_t)(b0 + y) * ld + i]) : T(0);
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] += (double)(xa[x] * wi) * (double)xb[y];
            }
            // block reduce the 16 accumulators
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const double s = dev::warp_sum(acc[x][y]);
                    if (lane == 0) s_red[warp][x * 4 + y] = s;
                }
            __syncthreads();
            if (tid < 16) {
                double s = 0;
                for (int wi = 0; wi < 8; ++wi) s += s_red[wi][tid];
                const int a = a0 + tid / 4, b = b0 + tid % 4;
                if (a < gs && b <= a) { Cout[a * gs + b] = s; Cout[b * gs + a] = s; }
            }
            __syncthreads();
        }
    }
}

// Single-pass variant for groups of at most GSP <= 12 columns: a thread keeps all GSP (GSP + 1) / 2 pair accumulators in registers,
// so every element of X_g is loaded exactly once (16-byte loads along the rows) instead of once per 4x4 pair block -- at config 3
// (5000 groups of 10, all re-decomposed in every IRLS iteration) the pair-block kernel re-read each group 6 times out of L2.
template <class T, int GSP>
__global__ void __launch_bounds__(256)
cov_small_kernel(const T* __restrict__ X, int64_t ld, int64_t n_pad, const CovItem* __restrict__ items,
                 const T* __restrict__ w, int w_is_sqrt, double* __restrict__ C_part, int64_t c_total, int rows_per_block, int K,
                 double* __restrict__ M_part = nullptr, int64_t m_total = 0)
{
    constexpr int VN = VecT<T>::N;
    constexpr int NP = GSP * (GSP + 1) / 2;
    __shared__ double s_red[8][NP + GSP];
    const CovItem it = items[blockIdx.x];
    const int gs = it.gs;
    const int rb = blockIdx.y;
    const int64_t row0 = (int64_t)rb * rows_per_block;
    const int64_t row1 = min((long long)n_pad, (long long)(row0 + rows_per_block));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool ones = it.col < 0;
    const T* Xg = X + (int64_t)(ones ? 0 : it.col) * ld;
    // accumulator type: the value type for float32 (a thread adds a few hundred terms; the reference's float templates accumulate in float
    // too), double for float64.  On the GLM path every screen group is re-decomposed in every IRLS iteration and this kernel was bound by
    // the FP64 pipe (78 DFMA per row for GSP = 12) and by its 156 accumulator registers, not by HBM (round 2: config-3 shard, 7.7 ms per
    // IRLS iteration at 26 % of the HBM peak).
    using ACC = typename std::conditional<std::is_same<T, float>::value, float, double>::type;
    ACC acc[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) acc[q] = 0;
    // optional second output of the same pass: the weighted column sums X_g^T w (the IRLS-weighted column means of the GLM path, which
    // otherwise cost a separate pass over the same columns in every IRLS iteration)
    const bool want_m = M_part != nullptr && it.pad > 0;
    ACC macc[GSP];
#pragma unroll
    for (int a = 0; a < GSP; ++a) macc[a] = 0;
    for (int64_t i = row0 + (int64_t)tid * VN; i < row1; i += 256 * VN) {
        T wv[VN];
        if (K == 1) vec_load<T>(w + i, wv);
        else {
#pragma unroll
            for (int k = 0; k < VN; ++k) wv[k] = w[(i + k) * K + it.cls];
        }
        if (w_is_sqrt) {
#pragma unroll
            for (int k = 0; k < VN; ++k) wv[k] *= wv[k];
        }
        T x[GSP][VN];
#pragma unroll
        for (int a = 0; a < GSP; ++a) {
            if (a < gs && !ones) vec_load<T>(Xg + (int64_t)a * ld + i, x[a]);
            else {
#pragma unroll
                for (int k = 0; k < VN; ++k) x[a][k] = (a < gs) ? T(1) : T(0);
            }
        }
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            int q = 0;
#pragma unroll
            for (int a = 0; a < GSP; ++a) {
                const ACC xw = (ACC)(x[a][k] * wv[k]);
                macc[a] += xw;
#pragma unroll
                for (int b = 0; b <= a; ++b, ++q) acc[q] += xw * (ACC)x[b][k];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const double t = dev::warp_sum((double)acc[q]);
        if (lane == 0) s_red[warp][q] = t;
    }
    if (want_m) {
#pragma unroll
        for (int a = 0; a < GSP; ++a) {
            const double t = dev::warp_sum((double)macc[a]);
            if (lane == 0) s_red[warp][NP + a] = t;
        }
    }
    __syncthreads();
    if (tid < NP) {
        double t = 0;
        for (int wi = 0; wi < 8; ++wi) t += s_red[wi][tid];
        int a = 0; while ((a + 1) * (a + 2) / 2 <= tid) ++a;          // tid = a (a + 1) / 2 + b, b <= a
        const int b = tid - a * (a + 1) / 2;
        if (a < gs) {
            double* Cout = C_part + (size_t)rb * c_total + it.out_off;
            Cout[a * gs + b] = t; Cout[b * gs + a] = t;
        }
    } else if (want_m && tid >= 128 && tid < 128 + GSP) {
        const int a = tid - 128;
        if (a < gs) {
            double t = 0;
            for (int wi = 0; wi < 8; ++wi) t += s_red[wi][NP + a];
            M_part[(size_t)rb * m_total + (it.pad - 1) + a] = t;
        }
    }
}

// Affine column transform of a dense matrix into a new one (matrix.standardize, CORE/matrix/matrix_naive_standardize.ipp:8-293):
//   out[i, j] = (X[i, j] - centers[j]) / scales[j]   for i < n, pad rows stay 0.   grid = (columns, row blocks)
template <class T>
__global__ void standardize_cols_kernel(const T* __restrict__ X, int64_t ld, int64_t n, const T* __restrict__ centers, const T* __restrict__ scales,
                                        T* __restrict__ out, int64_t ld_out)
{
    const int64_t j = blockIdx.x;
    const T c = centers[j], sc = scales[j];
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.y * blockDim.x)
        out[j * ld_out + i] = (X[j * ld + i] - c) / sc;
}
// Row / column gather into a new dense matrix (matrix.subset, CORE/matrix/matrix_naive_subset.ipp):
//   out[i, j] = X[rows ? rows[i] : i, cols ? cols[j] : j]     grid = (output columns, row blocks)
template <class T>
__global__ void gather_kernel(const T* __restrict__ X, int64_t ld, const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                              int64_t n_out, T* __restrict__ out, int64_t ld_out)
{
    const int64_t j = blockIdx.x;
    const T* src = X + (cols ? cols[j] : j) * ld;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.y * blockDim.x)
        out[j * ld_out + i] = src[rows ? rows[i] : i];
}

// out[k] = sum_rb part[rb * total + k]
__global__ void sum_parts_kernel(const double* __restrict__ part, int n_rb, int64_t total, double* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    double s = 0;
    for (int rb = 0; rb < n_rb; ++rb) s += part[(size_t)rb * total + k];
    out[k] = s;
}

} // namespace ab
