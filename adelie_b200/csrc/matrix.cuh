// adelie_b200/csrc/matrix.cuh -- host-side matrix objects that own device-resident data and
// launch the kernels.  Mirrors the operator set of MatrixNaiveBase
// (CORE/matrix/matrix_naive_base.hpp:57-143): cmul / ctmul / bmul / btmul / mul / cov / sq_mul.
//
// HBM layout of a dense matrix: column-major, leading dimension ld = n rounded up to 32 rows
// (so every column starts 128-byte aligned and TMA bulk copies of any 32-row-aligned tile are
// legal); the pad rows are zero.  Every device row-vector (resid, weights, ...) has the same
// padded length with zero pad.
#pragma once
#include "common.cuh"
#include "dense_kernels.cuh"
#include "sweep.cuh"
#include "sweep_batched.cuh"
#include "sparse_kernels.cuh"
#include "snp.cuh"
#include "snp_tc.cuh"
#include <unordered_map>
#include "dist.cuh"
#include <cuda.h>
#include <curand_kernel.h>

namespace ab {

// Per-device resources of the fused sweep kernel: LL exchange lines, epoch, abort flag.
struct SweepContext {
    DevBuf<dev::LLLine> ll, ll2; DevBuf<uint32_t> epoch; DevBuf<int> abort_flag;
    DevBuf<double> xs_sum; DevBuf<uint32_t> xs_cnt;      // one-hop atomic exchange (sweep.cuh): [4][ll_gs_cap] sums, [4][32] arrival counters
    int ncta_pad = 0; int ll_gs_cap = kGsMax;
    static SweepContext& get() {
        static thread_local SweepContext* ctx[64] = {nullptr};
        int dev = 0; AB_CUDA(cudaGetDevice(&dev));
        if (!ctx[dev]) {
            auto* c = new SweepContext();
            const int sms = DeviceInfo::get().sm_count;
            c->ncta_pad = (sms + 31) / 32 * 32;
            c->ll.alloc((size_t)2 * c->ll_gs_cap * c->ncta_pad);
            c->ll2.alloc((size_t)2 * c->ll_gs_cap * 32);
            c->epoch.alloc(1);
            uint32_t one = 1;
            c->epoch.upload(&one, 1);
            c->abort_flag.alloc(1);
            c->xs_sum.alloc((size_t)4 * c->ll_gs_cap); c->xs_cnt.alloc(4 * 32);
            AB_CUDA(cudaDeviceSynchronize());
            ctx[dev] = c;
        }
        return *ctx[dev];
    }
};

// Inputs of one fused pin solve (all device pointers).
template <class T>
struct PinLaunch {
    T* resid; const T* weights;
    const GroupMeta* meta; int S; const T* grec;
    const T* beta_in; int beta_len; const int8_t* is_active_in; int32_t* active_set; PinScalars* sc;
    double lmda, alpha, tol, newton_tol; long long max_iters; int newton_max_iters; int max_active_size; int intercept;
    int gs_max; int rec_max;     // largest group size / record length (elements) in the screen set
    int K = 1;                   // classes of a multi-response problem (resid / weights are (n, K) row-major)
    int feat_max = 0;            // K > 1: most physical X columns behind one screen group (0: gs_max)
};

struct SweepGeometry { int feat_max; int ncta, ncta_pad, threads, n_stages, stage_elems, rows_stride, gs_cap, units_base, units_rem; bool smem; size_t smem_bytes; };

template <class T>
inline SweepGeometry plan_sweep(int64_t n_pad, int gs_max, int rec_max, int K = 1, int feat_max = 0) {
    // (ncta_pad of this launch = ncta rounded up to 32; the LL buffer is sized for the largest possible grid)
    const auto& di = DeviceInfo::get();
    SweepGeometry g{};
    const int64_t units = n_pad / kRowAlign;
    int ncta = Configs::sweep_ctas > 0 ? Configs::sweep_ctas
                                       : (int)std::min<int64_t>(di.sm_count, std::max<int64_t>(1, units / std::max(1, Configs::sweep_min_rows_per_cta / kRowAlign)));
    ncta = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(ncta, units), di.sm_count));
    g.ncta = ncta;
    g.threads = std::min(std::max(64, Configs::sweep_threads / 32 * 32), kSweepThreadsMax);
    g.units_base = (int)(units / ncta); g.units_rem = (int)(units % ncta);
    g.rows_stride = (g.units_base + (g.units_rem ? 1 : 0)) * kRowAlign;
    g.gs_cap = std::max(4, (gs_max + 3) / 4 * 4);
    const int n_cwarps = g.threads / 32 - 1;
    g.ncta_pad = (ncta + 31) / 32 * 32;
    const size_t fixed = SweepSmem<T>::fixed_bytes(n_cwarps, g.gs_cap, g.ncta_pad) + 2 * sizeof(T) * (size_t)g.rows_stride * K;
    const int rec_pad = (rec_max + 3) / 4 * 4;
    g.feat_max = std::max(1, feat_max > 0 ? feat_max : gs_max);
    int stage_elems = g.rows_stride * g.feat_max + rec_pad;
    stage_elems = (stage_elems + 31) / 32 * 32;
    g.stage_elems = stage_elems;
    const size_t avail = di.smem_optin > fixed ? di.smem_optin - fixed : 0;
    int ns = (int)std::min<size_t>(kMaxStages, avail / (sizeof(T) * (size_t)stage_elems));
    g.smem = (ns >= 2) && !Configs::sweep_force_direct;
    if (g.smem) {
        g.n_stages = ns;
        g.smem_bytes = SweepSmem<T>::total(n_cwarps, g.gs_cap, g.ncta_pad, g.rows_stride * K, ns, stage_elems);
    } else {
        g.n_stages = 1; g.stage_elems = 0;
        // direct path: all warps are consumers; r / w stay in global memory
        g.smem_bytes = SweepSmem<T>::fixed_bytes(g.threads / 32, g.gs_cap, g.ncta_pad);
        g.rows_stride = 0;
    }
    return g;
}

// Geometry of the batched look-ahead kernel (sweep_batched.cuh).  ok == false: use pin_solve_kernel instead.
struct BatchGeometry { bool ok; int ncta, ncta_pad, B, Ccap, n_stages, stage_elems, rows_stride, units_base, units_rem, rec_stride, pslot_elems, ch; size_t smem_bytes; };

template <class T>
inline BatchGeometry plan_batched(int64_t n_pad, int gs_max, int rec_max) {
    const auto& di = DeviceInfo::get();
    BatchGeometry g{};
    g.ok = false;
    if (Configs::sweep_batch == 1 || Configs::sweep_force_direct || gs_max > 32 || gs_max < 1) return g;
    const int64_t units = n_pad / kRowAlign;
    int ncta = Configs::sweep_ctas > 0 ? Configs::sweep_ctas
                                       : (int)std::min<int64_t>(di.sm_count, std::max<int64_t>(1, units / std::max(1, Configs::sweep_min_rows_per_cta / kRowAlign)));
    ncta = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(ncta, units), di.sm_count));
    g.ncta = ncta; g.ncta_pad = (ncta + 31) / 32 * 32;
    g.units_base = (int)(units / ncta); g.units_rem = (int)(units % ncta);
    g.rows_stride = (g.units_base + (g.units_rem ? 1 : 0)) * kRowAlign;
    {   // panel slots carry the extended records (lane-local prox, gs <= 12) or the base records (generic prox)
        const int gsp = (gs_max + 3) / 4 * 4;
        g.rec_stride = (gs_max <= 12) ? 3 * gsp + 2 * gs_max * gsp : (rec_max + 3) / 4 * 4;
    }
    // a group is streamed in ring items of at most ch columns (two items per group once groups have more than 4 columns): more,
    // smaller stages keep both the HBM stream of the next batch and the L2 re-reads of the residual updates in flight.  When a CTA
    // owns many rows (few GPUs for a tall matrix: n = 1M on 2 GPUs is 3.4k rows per CTA) the default item does not fit the ring
    // any more: the item is narrowed (down to 2 columns) before the batch is shortened.
    const int ch0 = (gs_max > 4) ? (gs_max + 1) / 2 : gs_max;
    const int want = Configs::sweep_batch > 1 ? Configs::sweep_batch : 6;
    for (int B = std::min(std::min(want, kBatchMax), kBatchColsMax / gs_max); B >= 2 && !g.ok; --B) {
        const int Ccap = (B * gs_max + 3) / 4 * 4;
        const int pslot = Ccap * 2 * Ccap + B * g.rec_stride;
        const size_t fixed = BatchSmem<T>::fixed_bytes(Ccap) + sizeof(T) * (2 * (size_t)g.rows_stride + 2 * (size_t)pslot);
        if (fixed >= di.smem_optin) continue;
        for (int ch = ch0; ch >= std::min(ch0, 2); --ch) {
            const int stage_elems = (g.rows_stride * ch + 31) / 32 * 32;
            const int ns = (int)std::min<size_t>(kBatchStages, (di.smem_optin - fixed) / (sizeof(T) * (size_t)stage_elems));
            if (ns < (ch == ch0 ? 3 : 4)) continue;
            g.ok = true; g.B = B; g.Ccap = Ccap; g.pslot_elems = pslot; g.n_stages = ns; g.ch = ch; g.stage_elems = stage_elems;
            g.smem_bytes = fixed + sizeof(T) * (size_t)ns * stage_elems;
            break;
        }
    }
    return g;
}

template <class T>
struct BatchLaunch { const T* panels_screen; const T* panels_active; int n_active_panelled; int start_phase; const T* beta_rot_in; };

// xorshift-free counter based fill: X[i, j] ~ N(0,1) from Philox(seed, subsequence = column, offset = row)
template <class T>
__global__ void fill_normal_kernel(T* X, int64_t ld, int64_t n, int64_t p, unsigned long long seed, int64_t row_offset) {
    const int64_t j = blockIdx.x;                     // columns on grid.x (no 65535 limit), row blocks on grid.y
    for (int64_t i4 = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * 4; i4 < n; i4 += (int64_t)gridDim.y * blockDim.x * 4) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)j, (unsigned long long)(row_offset + i4), &st);
        const float4 z = curand_normal4(&st);
        const float zz[4] = {z.x, z.y, z.z, z.w};
        for (int k = 0; k < 4 && i4 + k < n; ++k) X[j * ld + i4 + k] = (T)zz[k];
    }
}

template <class T>
struct DenseMatrix {
    int64_t n = 0, p = 0, ld = 0;
    DevBuf<T> store; T* X = nullptr;
    int n_threads = 1;
    DevBuf<double> part;        // scratch for two-phase reductions
    DevBuf<T> ones;             // (n_pad,) ones with zero pad
    DevBuf<T> beta_rep;         // per-CTA coefficient replicas of the fused sweep (see sweep.cuh)
    DevBuf<T> brot_rep;         // per-CTA replicas of the coefficients in the groups' eigenbases (sweep_batched.cuh)
    int64_t beta_stride = 0;
    DevBuf<int8_t> act_rep; int64_t act_stride = 0;
    DevBuf<long long> stats;    // per-phase cycle counters of the fused sweep (Configs::sweep_profile)
    cudaStream_t stream = 0;

    // Second storage kind of the same device-matrix object: sparse CSC (MatrixNaiveSparse of the reference).  Every operator
    // below dispatches on `sparse`; the path driver (solver.cuh) is storage agnostic.
    bool sparse = false;
    int64_t nnz = 0;
    DevBuf<int64_t> sp_indptr; DevBuf<int32_t> sp_indices; DevBuf<T> sp_values; DevBuf<T> sp_vw;
    CscView<T> csc() const { return CscView<T>{sp_indptr.p, sp_indices.p, sp_values.p}; }

    // Third storage kind: SNP unphased genotypes (snp.cuh).  `snp_packed` holds the 2-bit codes of all p columns; `store` / `X` is
    // the dense cache of decoded columns (physical columns), filled on demand by phys_col(); every kernel that takes a column
    // index below takes a PHYSICAL column (dense / sparse: physical == logical).
    bool snp = false;
    DevBuf<uint32_t> snp_packed; int64_t snp_ldw = 0; DevBuf<T> snp_impute;
    const uint32_t* snp_bits = nullptr;                  // the packed genotypes the kernels read: snp_packed, or the base matrix's bits for a standardize view
    DevBuf<T> snp_center, snp_scale;                     // standardize view: values (x - c_j) / s_j (empty: raw genotypes)
    int64_t cache_cap = 0, cache_used = 0;
    std::unordered_map<int64_t, std::pair<int32_t, int32_t>> cache_map;      // logical first column -> (first slot, columns decoded there)
    long long n_decoded_cols = 0;
    struct SnpTag {};
    DenseMatrix(int64_t n_, int64_t p_, SnpTag) : n(n_), p(p_), ld(pad_rows(n_)), snp(true) {
        snp_ldw = ld / 16; snp_packed.alloc((size_t)snp_ldw * p); snp_impute.alloc(p); snp_bits = snp_packed.p;
    }
    // standardize view of another SNP matrix: shares its packed bits (the caller keeps the base alive), own impute / centers / scales
    DenseMatrix(const DenseMatrix& base, const T* h_centers, const T* h_scales, SnpTag) : n(base.n), p(base.p), ld(base.ld), snp(true) {
        snp_ldw = base.snp_ldw; snp_bits = base.snp_bits;
        snp_impute.alloc(p); snp_center.alloc(p); snp_scale.alloc(p);
        AB_CUDA(cudaMemcpyAsync(snp_impute.p, base.snp_impute.p, sizeof(T) * p, cudaMemcpyDeviceToDevice, 0));
        snp_center.upload(h_centers, p); snp_scale.upload(h_scales, p);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    // Physical first column of the logical columns [col, col + count): for SNP storage the columns are decoded into consecutive
    // slots of the dense cache the first time they are asked for.
    int32_t phys_col(int64_t col, int count) {
        if (!snp) return (int32_t)col;
        auto it = cache_map.find(col);
        if (it != cache_map.end() && it->second.second >= count) return it->second.first;
        if (cache_used + count > cache_cap) {
            int64_t want = std::max<int64_t>(cache_used + count, std::max<int64_t>(64, 2 * cache_cap));
            size_t fr = 0, tot = 0; AB_CUDA(cudaMemGetInfo(&fr, &tot));
            if ((double)want * ld * sizeof(T) > 0.8 * (double)fr) want = cache_used + count + std::min<int64_t>(256, cache_cap / 8);
            DevBuf<T> bigger;
            bigger.alloc((size_t)ld * want);
            if (cache_used) AB_CUDA(cudaMemcpyAsync(bigger.p, store.p, (size_t)ld * cache_used * sizeof(T), cudaMemcpyDeviceToDevice, stream));
            AB_CUDA(cudaStreamSynchronize(stream));
            store = std::move(bigger); X = store.p; cache_cap = want;
        }
        const int32_t slot = (int32_t)cache_used;
        dim3 grid((unsigned)count, (unsigned)std::min<int64_t>(64, (snp_ldw + 255) / 256));
        snp_decode_kernel<T><<<grid, 256, 0, stream>>>(snp_bits, snp_ldw, snp_impute.p, snp_center.p, snp_scale.p, n, col, count, X + (int64_t)slot * ld, ld);
        AB_CUDA(cudaGetLastError());
        cache_used += count; n_decoded_cols += count;
        cache_map[col] = std::make_pair(slot, (int32_t)count);
        return slot;
    }
    // out_part layout of the packed transposed GEMV: [row block][q][K]; returns the number of row blocks
    template <int KP, bool SQ>
    int snp_gemv_launch(int64_t j0, int q, int K, const T* v, const T* w) {
        constexpr int R = snp_gemv_rows_per_lane<KP>();
        const int64_t rows_per_tile = (int64_t)(kSnpGemvThreads / 32) * 32 * R;
        const int n_tiles = (int)((ld + rows_per_tile - 1) / rows_per_tile);
        // partial rows for the final reduction: as many as 256 MB of partial sums allow (fewer re-stagings of v*w per CTA), at least 32
        const int64_t rb_cap = std::max<int64_t>(32, ((int64_t)256 << 20) / ((int64_t)q * K * (int64_t)sizeof(double)));
        const int tiles_per_cta = (int)((n_tiles + rb_cap - 1) / rb_cap);
        const int n_rb = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
        const int sms = DeviceInfo::get().sm_count;
        int col_chunks = std::max(1, std::min((q + 31) / 32, (16 * sms + n_rb - 1) / n_rb));
        int cols_per_cta = ((q + col_chunks - 1) / col_chunks + 31) / 32 * 32;
        col_chunks = (q + cols_per_cta - 1) / cols_per_cta;
        part.reserve_keep((size_t)n_rb * q * K, stream);
        const size_t smem = snp_gemv_smem_bytes<T, KP>();
        auto fn = snp_center.n ? snp_gemv_t_kernel<T, KP, SQ, true> : snp_gemv_t_kernel<T, KP, SQ, false>;
        AB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fn<<<dim3(col_chunks, n_rb), kSnpGemvThreads, smem, stream>>>(snp_bits, snp_ldw, ld, snp_impute.p, snp_center.p, snp_scale.p, j0, q, cols_per_cta, tiles_per_cta, K, v, w, part.p);
        AB_CUDA(cudaGetLastError());
        return n_rb;
    }
    // tensor-core path (snp_tc.cuh): float32, 2 <= K <= 8 classes (Configs::snp_tc_min_k), exact integer accumulation
    DevBuf<uint8_t> stc_bq; DevBuf<double> stc_partial, stc_stats; DevBuf<int> stc_err;
    int snp_gemv_tc(int64_t j0, int q, int K, const float* v, const float* w) {
        const int n_chunks = (int)((ld + kStcKC - 1) / kStcKC);
        stc_bq.reserve_keep((size_t)n_chunks * kStcBTile, stream);
        stc_partial.reserve_keep((size_t)kStcStatBlocks * 16, stream);
        if (!stc_stats.n) stc_stats.alloc(32);
        if (!stc_err.n) stc_err.alloc(1);
        snp_tc_stats_kernel<<<kStcStatBlocks, 256, 0, stream>>>(v, w, (int64_t)ld * K, K, stc_partial.p);
        snp_tc_stats_finish_kernel<<<1, 32, 0, stream>>>(stc_partial.p, kStcStatBlocks, stc_stats.p);
        const int64_t n_groups16 = (int64_t)n_chunks * 16;
        snp_tc_quant_kernel<<<(unsigned)((n_groups16 * 32 + 255) / 256), 256, 0, stream>>>(v, w, ld, K, n_groups16, stc_stats.p, stc_bq.p);
        const int sms = DeviceInfo::get().sm_count;
        const int n_tiles = (q + kStcCols - 1) / kStcCols;
        int n_rb = std::max(1, std::min(n_chunks, (2 * sms + n_tiles - 1) / n_tiles));
        // int32 accumulators: |code * digit| <= 3 * 128 per row, so a row block stays below 2^31 / 384 = 5.5M rows (16384 chunks = 4.2M)
        const int chunks_per_rb = std::min(16384, (n_chunks + n_rb - 1) / n_rb);
        n_rb = (n_chunks + chunks_per_rb - 1) / chunks_per_rb;
        const bool stdv = snp_center.n != 0;
        part.reserve_keep((size_t)(n_rb + (stdv ? 1 : 0)) * q * K, stream);
        AB_CUDA(cudaFuncSetAttribute(snp_gemv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStcSmemBytes));
        snp_gemv_tc_kernel<<<dim3(n_tiles, n_rb), kStcThreads, kStcSmemBytes, stream>>>(snp_bits, snp_ldw, (const float*)snp_impute.p, (const float*)snp_center.p,
            (const float*)snp_scale.p, j0, q, K, stc_bq.p, n_chunks, chunks_per_rb, stc_stats.p, part.p, stc_err.p);
        if (stdv) snp_tc_const_kernel<<<(unsigned)(((int64_t)q * K + 255) / 256), 256, 0, stream>>>((const float*)snp_center.p, (const float*)snp_scale.p, j0, q, K,
                                                                                           stc_stats.p, part.p + (size_t)n_rb * q * K);
        AB_CUDA(cudaGetLastError());
        return n_rb + (stdv ? 1 : 0);
    }
    int snp_gemv(int64_t j0, int q, int K, const T* v, const T* w, bool sq) {
        if constexpr (std::is_same<T, float>::value) {
            if (!sq && Configs::snp_tc && K >= Configs::snp_tc_min_k && K <= 8) return snp_gemv_tc(j0, q, K, v, w);
        }
        if (sq) return snp_gemv_launch<1, true>(j0, q, 1, v, w);
        if (K <= 1) return snp_gemv_launch<1, false>(j0, q, 1, v, w);
        if (K <= 2) return snp_gemv_launch<2, false>(j0, q, K, v, w);
        if (K <= 4) return snp_gemv_launch<4, false>(j0, q, K, v, w);
        if (K <= 8) return snp_gemv_launch<8, false>(j0, q, K, v, w);
        return snp_gemv_launch<16, false>(j0, q, K, v, w);
    }

    DenseMatrix(int64_t n_, int64_t p_, bool sparse_ = false, int64_t nnz_ = 0) : n(n_), p(p_), ld(pad_rows(n_)), sparse(sparse_), nnz(nnz_) {
        if (!sparse) { store.alloc((size_t)ld * p); X = store.p; }
        else { sp_indptr.alloc(p + 1); sp_indices.alloc(std::max<int64_t>(nnz, 1)); sp_values.alloc(std::max<int64_t>(nnz, 1)); }
    }
    void upload_csc(const int64_t* indptr, const int32_t* indices, const T* values) {
        sp_indptr.upload(indptr, p + 1); sp_indices.upload(indices, nnz); sp_values.upload(values, nnz);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    void fill_sparse_random(int64_t nnz_per_col, unsigned long long seed) {
        sparse_fill_kernel<T><<<1184, 256, 0, stream>>>(sp_indptr.p, sp_indices.p, sp_values.p, n, p, nnz_per_col, seed);
        AB_CUDA(cudaGetLastError());
    }
    int64_t rows() const { return n; }
    int64_t cols() const { return p; }
    int64_t n_pad() const { return ld; }

    // host column-major (ldh >= n) or row-major (order == 1, ldh >= p) source
    void upload(const T* h, int order, int64_t ldh) {
        if (order == 0) {
            AB_CUDA(cudaMemcpy2D(X, ld * sizeof(T), h, ldh * sizeof(T), n * sizeof(T), p, cudaMemcpyHostToDevice));
        } else {
            // row-major host data: transpose through column chunks on the host side of the copy
            const int64_t chunk = std::max<int64_t>(1, (64 << 20) / (int64_t)(n * sizeof(T)));
            std::vector<T> buf((size_t)chunk * n);
            for (int64_t j0 = 0; j0 < p; j0 += chunk) {
                const int64_t jc = std::min(chunk, p - j0);
                for (int64_t i = 0; i < n; ++i)
                    for (int64_t j = 0; j < jc; ++j) buf[(size_t)j * n + i] = h[i * ldh + j0 + j];
                AB_CUDA(cudaMemcpy2D(X + j0 * ld, ld * sizeof(T), buf.data(), n * sizeof(T), n * sizeof(T), jc, cudaMemcpyHostToDevice));
            }
        }
    }
    void download(T* h, int64_t row0, int64_t nrows, int64_t col0, int64_t ncols, int64_t ldh) const {
        AB_CUDA(cudaMemcpy2D(h, ldh * sizeof(T), X + col0 * ld + row0, ld * sizeof(T), nrows * sizeof(T), ncols, cudaMemcpyDeviceToHost));
    }
    void fill_normal(unsigned long long seed, int64_t row_offset) {
        dim3 grid((unsigned)p, (unsigned)std::min<int64_t>(1024, (n + 1023) / 1024));
        fill_normal_kernel<T><<<grid, 256, 0, stream>>>(X, ld, n, p, seed, row_offset);
        AB_CUDA(cudaGetLastError());
    }
    const T* d_ones() {
        if (ones.n == 0) {
            std::vector<T> h(ld, T(0));
            for (int64_t i = 0; i < n; ++i) h[i] = 1;
            ones.alloc(ld); ones.upload(h.data(), ld);
            AB_CUDA(cudaStreamSynchronize(0));
        }
        return ones.p;
    }

    // out[c] = X[:, cols(c)]^T (v o w) for c < q  (cols == nullptr: LOGICAL columns j0 .. j0+q-1; otherwise a device list of
    // PHYSICAL columns); device pointers.
    // sub / sub_scale: fused epilogue out -= scale * sub  (grad -= resid_sum * X_means, solver_gaussian_naive.hpp:388-391)
    void d_gemv_t(int64_t j0, const int32_t* cols, int q, const T* v, const T* w, T* out, bool sq = false,
                  const T* sub = nullptr, const double* sub_scale_ptr = nullptr, double sub_scale = 0) {
        if (q <= 0) return;
        if (snp && cols == nullptr) {          // logical columns j0 .. j0+q-1 straight from the packed bits
            const int n_rb = snp_gemv(j0, q, 1, v, w, sq);
            gemv_t_reduce_kernel<T><<<(q + 255) / 256, 256, 0, stream>>>(part.p, n_rb, q, out, sub, sub_scale_ptr, sub_scale);
            AB_CUDA(cudaGetLastError());
            return;
        }
        if (sparse) {
            const unsigned nblk = (unsigned)((q + 7) / 8);
            sp_vw.reserve_keep((size_t)ld, stream);
            if (sq) {
                sp_vw_kernel<T, true><<<(unsigned)((ld + 255) / 256), 256, 0, stream>>>(v, w, sp_vw.p, ld);
                spmv_t_kernel<T, true><<<nblk, 256, 0, stream>>>(csc(), j0, cols, q, sp_vw.p, out, sub, sub_scale_ptr, sub_scale);
            } else {
                sp_vw_kernel<T, false><<<(unsigned)((ld + 255) / 256), 256, 0, stream>>>(v, w, sp_vw.p, ld);
                spmv_t_kernel<T, false><<<nblk, 256, 0, stream>>>(csc(), j0, cols, q, sp_vw.p, out, sub, sub_scale_ptr, sub_scale);
            }
            AB_CUDA(cudaGetLastError());
            return;
        }
        const int n_rb = (int)((ld + kGemvRows - 1) / kGemvRows);
        part.reserve_keep((size_t)n_rb * q, stream);
        dim3 grid((q + kGemvColsPerCta - 1) / kGemvColsPerCta, n_rb);
        if (sq) gemv_t_kernel<T, true><<<grid, kGemvThreads, 0, stream>>>(X, ld, ld, j0, cols, q, v, w, part.p);
        else gemv_t_kernel<T, false><<<grid, kGemvThreads, 0, stream>>>(X, ld, ld, j0, cols, q, v, w, part.p);
        gemv_t_reduce_kernel<T><<<(q + 255) / 256, 256, 0, stream>>>(part.p, n_rb, q, out, sub, sub_scale_ptr, sub_scale);
        AB_CUDA(cudaGetLastError());
    }
    // Multi-response `mul` of [kron(1, I_K) | kron(X, I_K)] (PY/solver.py:705-720 layout): v, w (n_pad, K) row-major (w nullable);
    // out[l] = sum_i v[i,l] w[i,l] for l < n_int (the intercept columns), out[n_int + j*K + l] = sum_i X[i,j] v[i,l] w[i,l].
    void d_mul_multi(int K, int n_int, const T* v, const T* w, T* out) {
        if (sparse) throw core_error("multi-response problems are not supported on sparse matrices.");
        if (K > kMultiMaxK) throw core_error("multi-response problems with more than 16 classes are not supported.");
        if (snp) {
            const int n_rb = snp_gemv(0, (int)p, K, v, w, false);
            gemv_t_reduce_kernel<T><<<(unsigned)((p * K + 255) / 256), 256, 0, stream>>>(part.p, n_rb, (int)(p * K), out + n_int, nullptr, nullptr, 0.0);
            AB_CUDA(cudaGetLastError());
            if (n_int) d_class_sums(K, v, w, out);
            return;
        }
        int tile_rows = (int)std::min<int64_t>(ld, std::max<int64_t>(kRowAlign, (int64_t)(48 * 1024 / (K * sizeof(T))) / 128 * 128));
        const int n_rb = (int)((ld + tile_rows - 1) / tile_rows);
        const size_t smem = (size_t)K * tile_rows * sizeof(T);
        part.reserve_keep((size_t)n_rb * (p + 1) * K, stream);
        dim3 grid((unsigned)((p + kGemvColsPerCta - 1) / kGemvColsPerCta), n_rb);
        gemv_t_multi_kernel<T><<<grid, kGemvThreads, smem, stream>>>(X, ld, ld, 0, (int)p, K, tile_rows, v, w, part.p);
        gemv_t_reduce_kernel<T><<<(unsigned)((p * K + 255) / 256), 256, 0, stream>>>(part.p, n_rb, (int)(p * K), out + n_int, nullptr, nullptr, 0.0);
        if (n_int) {
            double* part1 = part.p + (size_t)n_rb * p * K;
            gemv_t_multi_kernel<T><<<dim3(1, n_rb), kGemvThreads, smem, stream>>>(d_ones(), ld, ld, 0, 1, K, tile_rows, v, w, part1);
            gemv_t_reduce_kernel<T><<<1, 256, 0, stream>>>(part1, n_rb, K, out, nullptr, nullptr, 0.0);
        }
        AB_CUDA(cudaGetLastError());
    }
    // out[l] = sum_i v[i,l] * w[i,l]  (w nullable), l < K: per-class sums of (n_pad, K) row-major arrays
    void d_class_sums(int K, const T* v, const T* w, T* out) {
        int tile_rows = (int)std::min<int64_t>(ld, std::max<int64_t>(kRowAlign, (int64_t)(48 * 1024 / (K * sizeof(T))) / 128 * 128));
        const int n_rb = (int)((ld + tile_rows - 1) / tile_rows);
        part.reserve_keep((size_t)n_rb * K, stream);
        gemv_t_multi_kernel<T><<<dim3(1, n_rb), kGemvThreads, (size_t)K * tile_rows * sizeof(T), stream>>>(d_ones(), ld, ld, 0, 1, K, tile_rows, v, w, part.p);
        gemv_t_reduce_kernel<T><<<1, 256, 0, stream>>>(part.p, n_rb, K, out, nullptr, nullptr, 0.0);
        AB_CUDA(cudaGetLastError());
    }
    void d_mul(const T* v, const T* w, T* out, const T* sub = nullptr, const double* sub_scale_ptr = nullptr) {
        d_gemv_t(0, nullptr, (int)p, v, w, out, false, sub, sub_scale_ptr);
    }
    void d_btmul(int64_t j, int q, const T* v_dev, T* out) {
        if (sparse) {
            if (q <= 0) return;
            spaxpy_kernel<T><<<dim3(8, (unsigned)q), 256, 0, stream>>>(csc(), j, v_dev, out);
            AB_CUDA(cudaGetLastError());
            return;
        }
        if (q <= 0) return;
        constexpr int VN = VecT<T>::N;
        const int64_t nv = ld / VN;
        axpy_cols_kernel<T><<<(unsigned)((nv + 255) / 256), 256, 0, stream>>>(X, ld, ld, phys_col(j, q), q, v_dev, out);
        AB_CUDA(cudaGetLastError());
    }
    // Batched Gram: C[out_off + a*gs + b] = X_g^T diag(w or w^2) X_g, device doubles (c_total entries)
    // gs_max: largest group among the items when the caller knows it (<= 12 selects the single-pass register kernel), 0 = unknown
    template <int GSP>
    void cov_small_launch(dim3 grid, const CovItem* items_dev, const T* w, bool w_is_sqrt, double* out, int64_t c_total, int rows_per_block, int K,
                          double* m_out = nullptr, int64_t m_total = 0) {
        cov_small_kernel<T, GSP><<<grid, 256, 0, stream>>>(X, ld, ld, items_dev, w, w_is_sqrt ? 1 : 0, out, c_total, rows_per_block, K, m_out, m_total);
    }
    // can d_cov also produce the weighted column sums of the groups (CovItem::pad) in the same pass?  (single-pass register kernel only)
    bool cov_can_fuse_means(int K, int gs_max, bool w_is_sqrt) const { return !sparse && K == 1 && !w_is_sqrt && gs_max >= 1 && gs_max <= 12; }
    void d_cov(const CovItem* items_dev, int n_items, int64_t c_total, const T* w, bool w_is_sqrt, double* C_out, int K = 1, int gs_max = 0,
               double* M_out = nullptr, int64_t m_total = 0) {
        if (n_items <= 0) return;
        if (M_out && !cov_can_fuse_means(K, gs_max, w_is_sqrt)) throw core_error("internal: fused column sums requested from a Gram kernel that cannot produce them.");
        if (sparse) {
            if (K != 1) throw core_error("multi-response problems are not supported on sparse matrices.");
            spcov_kernel<T><<<n_items, 256, 0, stream>>>(csc(), items_dev, w, w_is_sqrt ? 1 : 0, C_out);
            AB_CUDA(cudaGetLastError());
            return;
        }
        const int sms = DeviceInfo::get().sm_count;
        int n_rb = std::max(1, std::min(sms, (4 * sms + n_items - 1) / n_items));
        int rows_per_block = (int)((ld + n_rb - 1) / n_rb);
        rows_per_block = (rows_per_block + kRowAlign - 1) / kRowAlign * kRowAlign;
        n_rb = (int)((ld + rows_per_block - 1) / rows_per_block);
        dim3 grid(n_items, n_rb);
        double* out = C_out; double* mout = M_out;
        if (n_rb > 1) {
            part.reserve_keep((size_t)n_rb * (c_total + (M_out ? m_total : 0)), stream);
            out = part.p; if (M_out) mout = part.p + (size_t)n_rb * c_total;
        }
        if (M_out) AB_CUDA(cudaMemsetAsync(mout, 0, sizeof(double) * (size_t)n_rb * m_total, stream));      // (positions of groups outside `items` stay 0)
        if (gs_max >= 1 && gs_max <= 4) cov_small_launch<4>(grid, items_dev, w, w_is_sqrt, out, c_total, rows_per_block, K, mout, m_total);
        else if (gs_max > 4 && gs_max <= 8) cov_small_launch<8>(grid, items_dev, w, w_is_sqrt, out, c_total, rows_per_block, K, mout, m_total);
        else if (gs_max > 8 && gs_max <= 10) cov_small_launch<10>(grid, items_dev, w, w_is_sqrt, out, c_total, rows_per_block, K, mout, m_total);
        else if (gs_max > 10 && gs_max <= 12) cov_small_launch<12>(grid, items_dev, w, w_is_sqrt, out, c_total, rows_per_block, K, mout, m_total);
        else cov_kernel<T><<<grid, 256, 0, stream>>>(X, ld, ld, items_dev, w, w_is_sqrt ? 1 : 0, out, c_total, rows_per_block, K);
        if (n_rb > 1) {
            sum_parts_kernel<<<(unsigned)((c_total + 255) / 256), 256, 0, stream>>>(part.p, n_rb, c_total, C_out);
            if (M_out) sum_parts_kernel<<<(unsigned)((m_total + 255) / 256), 256, 0, stream>>>(mout, n_rb, m_total, M_out);
        }
        AB_CUDA(cudaGetLastError());
    }

    // Gram panels of the batched kernel: Q[out_off + a*ldq + b] = X[:, col_s+a]^T W X[:, col_t+b] for every item
    void d_pair_gram(const PairItem* items_dev, int n_items, int64_t total, const T* w, T* Q, int ldq, int gs_max = 0) {
        if (n_items <= 0) return;
        const int sms = DeviceInfo::get().sm_count;
        int n_rb = std::max(1, std::min(sms, (8 * sms + n_items - 1) / n_items));
        int rows_per_block = (int)((ld + n_rb - 1) / n_rb);
        rows_per_block = (rows_per_block + kRowAlign - 1) / kRowAlign * kRowAlign;
        n_rb = (int)((ld + rows_per_block - 1) / rows_per_block);
        part.reserve_keep((size_t)(n_rb + 1) * total, stream);
        if (sizeof(T) == 4 && gs_max > 5 && gs_max <= 10) pair_gram_kernel<T, 10><<<dim3(n_items, n_rb), 256, 0, stream>>>(X, ld, ld, items_dev, w, part.p, total, rows_per_block);
        else pair_gram_kernel<T, 5><<<dim3(n_items, n_rb), 256, 0, stream>>>(X, ld, ld, items_dev, w, part.p, total, rows_per_block);
        DistContext& dc = DistContext::get();
        if (dc.active()) {
            // row-sharded: sum the row blocks, then the local blocks over the ranks, then scatter into the panels
            double* tmp = part.p + (size_t)n_rb * total;
            sum_parts_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(part.p, n_rb, total, tmp);
            dc.allreduce<double>(tmp, total, stream);
            pair_gram_finalize_kernel<T><<<n_items, 128, 0, stream>>>(items_dev, tmp, 1, total, Q, ldq);
        } else {
            pair_gram_finalize_kernel<T><<<n_items, 128, 0, stream>>>(items_dev, part.p, n_rb, total, Q, ldq);
        }
        AB_CUDA(cudaGetLastError());
    }

    // Whole Gram panels in one pass per panel (fp32 dense columns): panel_gram_tc_kernel on the tensor cores (tcgen05, TF32 operands,
    // Configs::panel_tc) or panel_gram_kernel on the CUDA cores; panel_tmp holds n_items * kPanelOut doubles.
    DevBuf<double> panel_tmp; DevBuf<float> part_f, wsqrt_f; DevBuf<int> tc_err;
    void d_panel_gram(const PanelItem* items_dev, int n_items, const float* w, float* Q, int ldq, int Ccap, int use_tc = -1, double* tmp_out = nullptr) {
        if (n_items <= 0) return;
        if (use_tc < 0) use_tc = Configs::panel_tc;
        const int sms = DeviceInfo::get().sm_count;
        int n_rb = std::max(1, std::min(sms, (2 * sms + n_items - 1) / n_items));
        int rows_per_block = (int)((ld + n_rb - 1) / n_rb);
        rows_per_block = (rows_per_block + kPanelRows - 1) / kPanelRows * kPanelRows;
        n_rb = (int)((ld + rows_per_block - 1) / rows_per_block);
        panel_tmp.reserve_keep((size_t)n_items * kPanelOut, stream);
        if (use_tc) {
            static_assert(kTcKC == kPanelRows, "row blocks are cut in chunks of kTcKC rows");
            part_f.reserve_keep((size_t)n_rb * n_items * kPanelOut, stream);
            if (!tc_err.n) tc_err.alloc(1);
            AB_CUDA(cudaFuncSetAttribute(panel_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
            wsqrt_f.reserve_keep((size_t)ld, stream);
            sqrt_weights_kernel<<<(unsigned)((ld + 255) / 256), 256, 0, stream>>>(w, wsqrt_f.p, ld);
            panel_gram_tc_kernel<<<dim3(n_items, n_rb), kTcThreads, kTcSmemBytes, stream>>>((const float*)X, ld, ld, items_dev, wsqrt_f.p, part_f.p, n_items,
                                                                                        rows_per_block, tc_err.p);
            panel_gram_sum_kernel<float><<<n_items, 256, 0, stream>>>(items_dev, part_f.p, n_rb, n_items, panel_tmp.p);
        } else {
            part.reserve_keep((size_t)n_rb * n_items * kPanelOut, stream);
            const size_t smem = sizeof(float) * (128 * kPanelStride + kPanelRows);
            AB_CUDA(cudaFuncSetAttribute(panel_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            panel_gram_kernel<<<dim3(n_items, n_rb), 256, smem, stream>>>((const float*)X, ld, ld, items_dev, w, part.p, n_items, rows_per_block);
            panel_gram_sum_kernel<double><<<n_items, 256, 0, stream>>>(items_dev, part.p, n_rb, n_items, panel_tmp.p);
        }
        DistContext& dc = DistContext::get();
        if (dc.active()) dc.allreduce<double>(panel_tmp.p, (int64_t)n_items * kPanelOut, stream);      // row-sharded: sum the local panels over the ranks
        if (Q) panel_gram_scatter_kernel<float><<<n_items, 256, 0, stream>>>(items_dev, panel_tmp.p, Q, ldq, Ccap);
        AB_CUDA(cudaGetLastError());
        if (tmp_out) { AB_CUDA(cudaMemcpyAsync(tmp_out, panel_tmp.p, sizeof(double) * (size_t)n_items * kPanelOut, cudaMemcpyDeviceToHost, stream)); AB_CUDA(cudaStreamSynchronize(stream)); }
    }
    // the tensor-core kernel bounds every wait and raises this flag instead of hanging; checked at the caller's next synchronisation
    void check_tc_error() {
        int e = 0, e2 = 0;
        if (tc_err.n) tc_err.download(&e, 1);
        if (stc_err.n) stc_err.download(&e2, 1);
        if (tc_err.n || stc_err.n) AB_CUDA(cudaStreamSynchronize(stream));
        if (e || e2) {
            int z = 0; if (tc_err.n) tc_err.upload(&z, 1); if (stc_err.n) stc_err.upload(&z, 1);
            throw solver_error("tensor-core kernel timed out (pipeline protocol error).");
        }
    }

    // The batched look-ahead pin solve (sweep_batched.cuh); single response, single GPU, static weights.
    BatchGeometry last_bgeom{};
    void pin_solve_batched(const PinLaunch<T>& L, const BatchGeometry& g, const BatchLaunch<T>& bl) {
        SweepContext& ctx = SweepContext::get();
        last_bgeom = g;
        last_geom.ncta = g.ncta; last_geom.n_stages = g.n_stages; last_geom.smem = true; last_geom.smem_bytes = g.smem_bytes; last_geom.threads = 512;
        BatchKernelArgs<T> a{};
        a.X = X; a.ld = ld; a.resid = L.resid; a.weights = L.weights;
        a.meta = L.meta; a.S = L.S; a.grec = L.grec;
        act_stride = ((int64_t)L.S + 127) / 128 * 128 + 128;
        act_rep.reserve_keep((size_t)act_stride * g.ncta, stream);
        a.is_active_in = L.is_active_in; a.is_active_rep = act_rep.p; a.act_stride = act_stride;
        beta_stride = ((int64_t)L.beta_len + 31) / 32 * 32 + 32;
        beta_rep.reserve_keep((size_t)beta_stride * g.ncta, stream);
        a.beta_in = L.beta_in; a.beta_rep = beta_rep.p; a.beta_stride = beta_stride; a.beta_len = L.beta_len;
        brot_rep.reserve_keep((size_t)beta_stride * g.ncta, stream);
        a.brot_in = bl.beta_rot_in; a.brot_rep = brot_rep.p;
        a.active_set = L.active_set; a.sc = L.sc;
        a.panels_screen = bl.panels_screen; a.panels_active = bl.panels_active; a.n_active_panelled = bl.n_active_panelled;
        a.B = g.B; a.Ccap = g.Ccap; a.use_ext = (L.gs_max <= 12) ? 1 : 0;
        a.ll1 = ctx.ll.p; a.ll2 = ctx.ll2.p; a.ncta_pad = g.ncta_pad;
        { int f = 1; while (f * f < g.ncta) ++f; a.fan = std::max(1, f); }
        a.epoch = ctx.epoch.p; a.abort_flag = ctx.abort_flag.p;
        {
            DistContext& dc = DistContext::get();
            a.rank = dc.active() ? dc.rank : 0; a.world = dc.active() ? dc.world : 1;
            for (int r = 0; r < kMaxRanksDev; ++r) a.ll3_peer[r] = (dc.active() && r < dc.world) ? dc.ll3(r) : nullptr;
        }
        a.lmda = L.lmda; a.alpha = L.alpha; a.tol = L.tol; a.newton_tol = L.newton_tol; a.dbeta_tol = Configs::dbeta_tol;
        a.max_iters = L.max_iters; a.newton_max_iters = L.newton_max_iters; a.max_active_size = L.max_active_size; a.intercept = L.intercept;
        a.start_phase = bl.start_phase;
        a.units_base = g.units_base; a.units_rem = g.units_rem; a.rows_stride = g.rows_stride;
        a.n_stages = g.n_stages; a.stage_elems = g.stage_elems; a.rec_stride = g.rec_stride; a.pslot_elems = g.pslot_elems; a.ch = g.ch; a.u_prefetch = Configs::sweep_u_prefetch ? 1 : 0; a.l2_prefetch = Configs::sweep_l2_prefetch ? 1 : 0;
        if (Configs::sweep_profile) { if (!stats.n) stats.alloc(32 + 8 * 160); a.stats = stats.p; } else a.stats = nullptr;
        void* kargs[] = {&a};
        const void* fn = !a.stats ? (const void*)pin_solve_batched_kernel<T, 0> : (Configs::sweep_profile >= 2 ? (const void*)pin_solve_batched_kernel<T, 2> : (const void*)pin_solve_batched_kernel<T, 1>);
        AB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
        if (g.ncta > 1) AB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(g.ncta), dim3(512), kargs, g.smem_bytes, stream));
        else AB_CUDA(cudaLaunchKernel(fn, dim3(1), dim3(512), kargs, g.smem_bytes, stream));
    }

    // The fused pin solve (sweep.cuh).
    SweepGeometry last_geom{};
    void pin_solve_sparse(const PinLaunch<T>& L) {
        if (L.K != 1) throw core_error("multi-response problems are not supported on sparse matrices.");
        if (DistContext::get().active()) throw core_error("sparse matrices are not supported in row-sharded multi-GPU mode.");
        last_geom = SweepGeometry{}; last_geom.ncta = 1; last_geom.threads = kSparseThreads; last_bgeom.ok = false;
        SparsePinArgs<T> a{};
        a.X = csc(); a.resid = L.resid; a.weights = L.weights; a.meta = L.meta; a.S = L.S; a.grec = L.grec;
        act_stride = ((int64_t)L.S + 127) / 128 * 128 + 128; act_rep.reserve_keep((size_t)act_stride, stream);
        beta_stride = ((int64_t)L.beta_len + 31) / 32 * 32 + 32; beta_rep.reserve_keep((size_t)beta_stride, stream);
        a.beta_in = L.beta_in; a.beta_out = beta_rep.p; a.beta_len = L.beta_len; a.is_active_in = L.is_active_in; a.is_active_out = act_rep.p;
        a.active_set = L.active_set; a.sc = L.sc;
        a.lmda = L.lmda; a.alpha = L.alpha; a.tol = L.tol; a.newton_tol = L.newton_tol; a.dbeta_tol = Configs::dbeta_tol;
        a.max_iters = L.max_iters; a.newton_max_iters = L.newton_max_iters; a.max_active_size = L.max_active_size; a.intercept = L.intercept;
        a.gs_cap = std::max(4, (L.gs_max + 3) / 4 * 4);
        const size_t smem = sizeof(double) * ((size_t)(kSparseThreads / 32 + 10) * a.gs_cap + 4 * 32 + 8);
        last_geom.smem_bytes = smem;
        const void* fn = (const void*)pin_solve_sparse_kernel<T>;
        AB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        void* kargs[] = {&a};
        AB_CUDA(cudaLaunchKernel(fn, dim3(1), dim3(kSparseThreads), kargs, smem, stream));
    }

    void pin_solve(const PinLaunch<T>& L) {
        if (L.gs_max > kGsMax) throw core_error("group size " + std::to_string(L.gs_max) + " exceeds the fused sweep kernel's limit of " + std::to_string(kGsMax) + ".");
        if (sparse) { pin_solve_sparse(L); return; }
        SweepContext& ctx = SweepContext::get();
        if (L.K > 16) throw core_error("multi-response problems with more than 16 classes are not supported by the fused sweep kernel.");
        SweepGeometry g = plan_sweep<T>(ld, L.gs_max, L.rec_max, L.K, L.feat_max);
        last_geom = g; last_bgeom.ok = false;
        PinKernelArgs<T> a{};
        a.X = X; a.ld = ld; a.n_pad = ld; a.K = L.K; a.resid = L.resid; a.weights = L.weights;
        a.meta = L.meta; a.S = L.S; a.grec = L.grec;
        act_stride = ((int64_t)L.S + 127) / 128 * 128 + 128;
        act_rep.reserve_keep((size_t)act_stride * g.ncta, stream);
        a.is_active_in = L.is_active_in; a.is_active_rep = act_rep.p; a.act_stride = act_stride;
        beta_stride = ((int64_t)L.beta_len + 31) / 32 * 32 + 32;
        beta_rep.reserve_keep((size_t)beta_stride * g.ncta, stream);
        a.beta_in = L.beta_in; a.beta_rep = beta_rep.p; a.beta_stride = beta_stride; a.beta_len = L.beta_len;
        a.active_set = L.active_set; a.sc = L.sc;
        a.ll = ctx.ll.p; a.ll_gs_cap = ctx.ll_gs_cap; a.ncta_pad = g.ncta_pad;
        a.ll2 = ctx.ll2.p;
        a.xs_sum = ctx.xs_sum.p; a.xs_cnt = ctx.xs_cnt.p; a.xchg_atomic = Configs::sweep_xchg ? 1 : 0;
        if (a.xchg_atomic && g.ncta > 1) {       // all four slots start clean (the batched kernel advances the shared epoch between launches)
            AB_CUDA(cudaMemsetAsync(ctx.xs_sum.p, 0, ctx.xs_sum.n * sizeof(double), stream));
            AB_CUDA(cudaMemsetAsync(ctx.xs_cnt.p, 0, ctx.xs_cnt.n * sizeof(uint32_t), stream));
        }
        { int f = 1; while (f * f < g.ncta) ++f; a.fan = std::max(1, f); }      // fan = ceil(sqrt(ncta)) => n_groups <= fan + 1 <= 32
        a.epoch = ctx.epoch.p; a.abort_flag = ctx.abort_flag.p;
        {
            DistContext& dc = DistContext::get();
            a.rank = dc.active() ? dc.rank : 0; a.world = dc.active() ? dc.world : 1;
            for (int r = 0; r < kMaxRanksDev; ++r) a.ll3_peer[r] = (dc.active() && r < dc.world) ? dc.ll3(r) : nullptr;
        }
        a.lmda = L.lmda; a.alpha = L.alpha; a.tol = L.tol; a.newton_tol = L.newton_tol; a.dbeta_tol = Configs::dbeta_tol;
        a.max_iters = L.max_iters; a.newton_max_iters = L.newton_max_iters; a.max_active_size = L.max_active_size; a.intercept = L.intercept;
        a.units_base = g.units_base; a.units_rem = g.units_rem; a.rows_stride = g.rows_stride;
        a.n_stages = g.n_stages; a.stage_elems = g.stage_elems; a.gs_max = std::max(L.gs_max, 1); a.gs_cap = g.gs_cap; a.feat_max = g.feat_max;
        if (Configs::sweep_profile) { if (!stats.n) stats.alloc(32 + 8 * 160); a.stats = stats.p; } else a.stats = nullptr;
        void* kargs[] = {&a};
        const void* fn = g.smem ? (const void*)pin_solve_kernel<T, true> : (const void*)pin_solve_kernel<T, false>;
        AB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
        if (g.ncta > 1) {
            AB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(g.ncta), dim3(g.threads), kargs, g.smem_bytes, stream));
        } else {
            AB_CUDA(cudaLaunchKernel(fn, dim3(1), dim3(g.threads), kargs, g.smem_bytes, stream));
        }
    }
};

} // namespace ab
