// adelie_b200/csrc/sweep.cuh -- the fused, persistent block-coordinate-descent kernel.
//
// One launch = one "pin solve" at one lambda (reference: pin::naive::solve,
// CORE/solver/solver_gaussian_pin_naive.hpp:223-401 for a 1-element lmda_path):
//   repeat { active-set sweeps until convergence (solve_active :181-215);
//            one sweep over the whole screen set, adding new actives (:331-341) }
// and every sweep is coordinate_descent (:26-168): per group the partial gradient
// X_g^T (w o r)  (matrix.bmul), the proximal update (bcd newton_solver,
// CORE/bcd/unconstrained/newton.hpp:44-142) and the residual update r -= X_g dbeta
// (matrix.btmul) -- fused so that X_g is read from HBM exactly once per group update.
//
// Layout / execution model (B200):
//   * grid = one persistent CTA per SM (cooperative launch => all co-resident); CTA c owns a
//     contiguous row tile of X, r and w.  r and w stay in shared memory for the whole launch.
//   * a producer warp streams the (rows_tile x gs) column block of the next groups into a
//     shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), so HBM
//     traffic is decoupled from the Gauss-Seidel dependency chain;
//   * per group: warp-shuffle reduction of the skinny GEMV, one 16-byte flagged "LL" line per
//     (CTA, column) written to L2, every CTA polls all lines (barrier and data in one round
//     trip) and sums them in a fixed order => bitwise identical gradient in all CTAs, so the
//     tiny proximal solve and all control flow are replicated without any broadcast;
//   * the residual tile is updated from the same shared-memory copy of X_g.
#pragma once
#include "common.cuh"
#include "device_prims.cuh"

namespace ab {

constexpr int kGsMax = 128;            // largest group size handled by the fused kernel
constexpr int kSweepThreadsMax = 512;
constexpr int kMaxStages = 6;
constexpr int kLLSeg = 8;              // max 32-CTA segments => up to 256 CTAs

struct GroupMeta {          // one per screen position (32 bytes)
    int32_t col;            // first column of the group in X
    int32_t gs;             // group size
    int32_t begin;          // offset into screen_beta
    int32_t rec_elems;      // padded length of the group's record
    int64_t rec_off;        // element offset of the record [A(gs) | xm(gs) | V(gs*gs, (r,c)->r*gs+c)]
    double pen;             // penalty factor
};

struct PinScalars {         // device-resident in/out scalars of one pin solve
    double rsq, resid_sum;
    long long iters, n_group_updates;
    int active_set_size, error, newton_iters_max, pad;
};

enum { kErrNone = 0, kErrMaxCds = 1, kErrMaxActive = 2, kErrNewton = 3, kErrAbort = 4 };
enum { kSweepActive = 0, kSweepScreen = 1, kSweepExit = 2 };

template <class T>
struct PinKernelArgs {
    const T* X; int64_t ld; int64_t n_pad;
    T* resid; const T* weights;
    const GroupMeta* meta; int S;
    const T* grec;
    T* screen_beta; int8_t* is_active; int32_t* active_set;
    PinScalars* sc;
    dev::LLLine* ll; int ll_gs_cap; int ncta_pad;
    uint32_t* epoch; int* abort_flag;
    double lmda, alpha, tol, newton_tol, dbeta_tol;
    long long max_iters; int newton_max_iters; int max_active_size; int intercept;
    int units_base, units_rem;     // row partition in units of kRowAlign rows
    int rows_stride;               // max rows per CTA (column stride inside a stage)
    int n_stages; int stage_elems; // ring geometry (elements of T per stage)
    int gs_max;                    // largest group size in the screen set
    int gs_cap;                    // stride of the per-column shared-memory scratch arrays (>= gs_max, multiple of 4)
};

struct SweepCtrl {
    int kind, count;          // descriptor of the current sweep (consumers)
    int p_kind, p_count;      // descriptor handed to the TMA producer warp (non-empty sweeps and EXIT only)
    int changed, abort, next, error;
    int stop;                 // CTA-local shutdown flag watched by the producer's waits
    unsigned consumed;        // items consumed when the consumers stopped (for the producer's drain)
};

template <class T> struct VecT;
template <> struct VecT<float> { using type = float4; static constexpr int N = 4; };
template <> struct VecT<double> { using type = double2; static constexpr int N = 2; };

template <class T> __device__ __forceinline__ void vec_load(const T* p, T (&v)[VecT<T>::N]);
template <> __device__ __forceinline__ void vec_load<float>(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void vec_load<double>(const double* p, double (&v)[2]) {
    const double2 t = *reinterpret_cast<const double2*>(p); v[0] = t.x; v[1] = t.y;
}
template <class T> __device__ __forceinline__ void vec_store(T* p, const T (&v)[VecT<T>::N]);
template <> __device__ __forceinline__ void vec_store<float>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void vec_store<double>(double* p, const double (&v)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}

// Shared-memory carve-up (host and device must agree).
template <class T>
struct SweepSmem {
    static constexpr size_t kHeaderBytes = 512;                                // barriers + ctrl
    __host__ __device__ static size_t fixed_bytes(int n_cwarps, int gs_cap) {
        size_t b = kHeaderBytes
                 + sizeof(double) * (size_t)n_cwarps * gs_cap      // wpart
                 + sizeof(double) * (size_t)gs_cap * kLLSeg        // wsum
                 + sizeof(double) * (size_t)gs_cap * 8             // prox scratch
                 + sizeof(T) * (size_t)gs_cap;                     // del
        return (b + 127) / 128 * 128;
    }
    static size_t total(int n_cwarps, int gs_cap, int rows_stride, int n_stages, int stage_elems) {
        return fixed_bytes(n_cwarps, gs_cap) + 2 * sizeof(T) * (size_t)rows_stride + sizeof(T) * (size_t)n_stages * stage_elems;
    }
};

// ------------------------------------------------------------------------------------------
// Group proximal update, executed by the 32 lanes of the control warp (replicated in every CTA).
// All scalars are double; for T = float the inputs convert exactly and the result is rounded
// once.  Follows coordinate_descent (solver_gaussian_pin_naive.hpp:75-164) + update_coordinate
// (solver_gaussian_pin_base.hpp:148-195) + newton_solver (newton.hpp:44-142, h0 = 0).
// ------------------------------------------------------------------------------------------
struct ProxState {
    double rsq, resid_sum, cm;
    int A;                    // active_set_size
    int error;
    int newton_iters_max;
};

// Warp-cooperative group prox: minimiser of 0.5 x^T diag(L) x - v^T x + l1 ||x|| + 0.5 l2 ||x||^2
// (newton_solver_base, CORE/bcd/unconstrained/newton.hpp:44-111 with newton_root_find,
// CORE/optimization/newton.hpp:35-66).  abs_start selects the Newton-ABS initial point
// (newton.hpp:229-266; bounds CORE/bcd/utils.hpp:20-97), otherwise h0 = 0 (newton.hpp:131).
// L, v, D (scratch), x are arrays of length >= gs visible to the whole warp (shared memory).
__device__ __forceinline__ void warp_prox_newton(const double* L, const double* v, int gs, double l1, double l2, double tol,
                                                 int max_iters, bool abs_start, double* D, double* x, int& iters, int lane)
{
    iters = 0;
    double vn2 = 0;
    for (int c = lane; c < gs; c += 32) vn2 += v[c] * v[c];
    vn2 = dev::warp_sum(vn2);
    if (sqrt(vn2) <= l1) {                                   // newton.hpp:62-66
        for (int c = lane; c < gs; c += 32) x[c] = 0.0;
        __syncwarp();
        return;
    }
    if (l1 <= 0.0) {                                         // newton.hpp:72-75
        for (int c = lane; c < gs; c += 32) x[c] = v[c] / (L[c] + l2);
        __syncwarp();
        return;
    }
    for (int c = lane; c < gs; c += 32) D[c] = L[c] + l2;
    __syncwarp();
    auto phi = [&](double h) {                               // root_function, bcd/utils.hpp:99-109
        double t = 0;
        for (int c = lane; c < gs; c += 32) { const double q = v[c] / (D[c] * h + l1); t += q * q; }
        return dev::warp_sum(t) - 1.0;
    };
    double h = 0.0;
    if (abs_start) {
        double sumD = 0, a = 0, v_l1 = 0, dmin = INFINITY;
        for (int c = lane; c < gs; c += 32) { sumD += D[c]; a += D[c] * D[c]; v_l1 += fabs(v[c]); dmin = fmin(dmin, D[c]); }
        sumD = dev::warp_sum(sumD); a = dev::warp_sum(a); v_l1 = dev::warp_sum(v_l1); dmin = -dev::warp_max(-dmin);
        const double b = l1 * sumD, cc = l1 * l1 * gs - v_l1 * v_l1, discr = b * b - a * cc;       // root_lower_bound
        double h_min = (discr > -1e-12) ? (-b + sqrt(fmax(discr, 0.0))) / a : 0.0;
        h_min = fmax(h_min, 0.0);
        double h_max = 0, dmin_nnz;                           // root_upper_bound (zero_tol = 1e-14)
        if (dmin <= 1e-14) {
            double hm = 0, v_S = 0, dn = INFINITY;
            for (int c = lane; c < gs; c += 32) {
                const bool nz = D[c] > 1e-14; const double vi2 = v[c] * v[c];
                hm += nz ? vi2 / (D[c] * D[c]) : 0.0; v_S += (D[c] <= 0) ? vi2 : 0.0; dn = nz ? fmin(dn, D[c]) : dn;
            }
            hm = dev::warp_sum(hm); v_S = dev::warp_sum(v_S); dmin_nnz = -dev::warp_max(-dn);
            h_max = sqrt(fmax(hm / (1 - v_S / (l1 * l1)), 0.0));
        } else {
            double s2 = 0;
            for (int c = lane; c < gs; c += 32) { const double q = v[c] / D[c]; s2 += q * q; }
            h_max = sqrt(dev::warp_sum(s2)); dmin_nnz = dmin;
        }
        if (h_max - h_min <= 1e-1) h = h_min;
        else {
            double h_cand = h_max, fh;
            do {
                const double w = fmax(l1 / (dmin_nnz * h_cand + l1), 0.05);
                h_cand = w * h_min + (1 - w) * h_cand;
                fh = phi(h_cand);
            } while ((fh < 0) && (fabs(fh) > tol));
            h = h_cand;
        }
    }
    while (true) {                                           // newton.hpp:83-93 + optimization/newton.hpp:56-63
        double t = 0, sd = 0;
        for (int c = lane; c < gs; c += 32) {
            const double u = 1.0 / (D[c] * h + l1);
            const double q = v[c] * u;
            const double xx = q * q;
            t += xx; sd += xx * D[c] * u;
        }
        t = dev::warp_sum(t); sd = dev::warp_sum(sd);
        const double fh = t - 1.0;
        if (!(fabs(fh) > tol) || iters >= max_iters) break;
        const double dfh = -sd * (1.0 + sqrt(t)) / t;
        h = fmax(h - fh / dfh, 0.0);
        ++iters;
    }
    for (int c = lane; c < gs; c += 32) x[c] = h * v[c] / (D[c] * h + l1);     // newton.hpp:109
    __syncwarp();
}

// One-warp kernel exposing the prox and its helpers (adelie.bcd API; device code shared with the sweep).
// mode: 0 newton, 1 newton_abs, 2 root_lower_bound, 3 root_upper_bound, 4 root_function (aux = h or zero_tol)
static __global__ void bcd_kernel(int mode, int q, const double* __restrict__ Lg, const double* __restrict__ vg, double l1, double l2,
                                  double tol, int max_iters, double aux, double* __restrict__ x_out, double* __restrict__ scal_out)
{
    extern __shared__ double sh[];
    double* L = sh; double* v = sh + q; double* D = sh + 2 * q; double* x = sh + 3 * q;
    const int lane = threadIdx.x;
    for (int c = lane; c < q; c += 32) { L[c] = Lg[c]; v[c] = vg[c]; }
    __syncwarp();
    if (mode <= 1) {
        int iters = 0;
        warp_prox_newton(L, v, q, l1, l2, tol, max_iters, mode == 1, D, x, iters, lane);
        for (int c = lane; c < q; c += 32) x_out[c] = x[c];
        if (lane == 0) scal_out[0] = (double)iters;
    } else if (mode == 2) {
        double sumD = 0, a = 0, v_l1 = 0;
        for (int c = lane; c < q; c += 32) { sumD += L[c]; a += L[c] * L[c]; v_l1 += fabs(v[c]); }
        sumD = dev::warp_sum(sumD); a = dev::warp_sum(a); v_l1 = dev::warp_sum(v_l1);
        const double b = l1 * sumD, cc = l1 * l1 * q - v_l1 * v_l1, discr = b * b - a * cc;
        double h_min = (discr > -1e-12) ? (-b + sqrt(fmax(discr, 0.0))) / a : 0.0;
        if (lane == 0) scal_out[0] = fmax(h_min, 0.0);
    } else if (mode == 3) {
        double dmin = INFINITY;
        for (int c = lane; c < q; c += 32) dmin = fmin(dmin, L[c]);
        dmin = -dev::warp_max(-dmin);
        double h_max;
        if (dmin <= aux) {
            double hm = 0, v_S = 0;
            for (int c = lane; c < q; c += 32) {
                const bool nz = L[c] > aux; const double vi2 = v[c] * v[c];
                hm += nz ? vi2 / (L[c] * L[c]) : 0.0; v_S += (L[c] <= 0) ? vi2 : 0.0;
            }
            hm = dev::warp_sum(hm); v_S = dev::warp_sum(v_S);
            h_max = sqrt(fmax(hm / (1 - v_S / (l1 * l1)), 0.0));
        } else {
            double s2 = 0;
            for (int c = lane; c < q; c += 32) { const double qq = v[c] / L[c]; s2 += qq * qq; }
            h_max = sqrt(dev::warp_sum(s2));
        }
        if (lane == 0) scal_out[0] = h_max;
    } else {
        double t = 0;
        for (int c = lane; c < q; c += 32) { const double qq = v[c] / (L[c] * aux + l1); t += qq * qq; }
        t = dev::warp_sum(t);
        if (lane == 0) scal_out[0] = t - 1.0;
    }
}

template <class T, bool SMEM>
__global__ void __launch_bounds__(kSweepThreadsMax, 1)
pin_solve_kernel(const __grid_constant__ PinKernelArgs<T> a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int VN = VecT<T>::N;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_warps = blockDim.x >> 5;
    const int NW = SMEM ? n_warps - 1 : n_warps;     // consumer warps (last warp = TMA producer when staging)
    const int NTC = NW * 32;
    const int cta = blockIdx.x, ncta = gridDim.x;

    // ---- shared memory carve-up
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);               // [kMaxStages]
    uint64_t* empty_bar = full_bar + kMaxStages;                               // [kMaxStages]
    uint64_t* desc_bar = empty_bar + kMaxStages;                               // [1]
    SweepCtrl* ctrl = reinterpret_cast<SweepCtrl*>(smem_raw + 128);
    const int gsc = a.gs_cap;
    double* wpart = reinterpret_cast<double*>(smem_raw + SweepSmem<T>::kHeaderBytes);     // [NW][gsc]
    double* wsum = wpart + (size_t)NW * gsc;                                                // [gsc][kLLSeg]
    double* px = wsum + (size_t)gsc * kLLSeg;                                               // 8 x [gsc] prox scratch
    double* p_aold = px, *p_A = px + gsc, *p_xm = px + 2 * gsc, *p_gk = px + 3 * gsc,
          *p_gt = px + 4 * gsc, *p_atold = px + 5 * gsc, *p_at = px + 6 * gsc, *p_D = px + 7 * gsc;
    T* s_del = reinterpret_cast<T*>(px + 8 * gsc);                                          // [gsc]
    unsigned char* tiles = smem_raw + SweepSmem<T>::fixed_bytes(NW, gsc);
    T* sr = reinterpret_cast<T*>(tiles);
    T* sw = sr + a.rows_stride;
    T* stages = sw + a.rows_stride;

    // ---- my row tile
    const int my_units = a.units_base + (cta < a.units_rem ? 1 : 0);
    const int64_t unit0 = (int64_t)cta * a.units_base + min(cta, a.units_rem);
    const int64_t r0 = unit0 * kRowAlign;
    const int rows = my_units * kRowAlign;

    volatile int* abort_flag = a.abort_flag;
    volatile int* stop_flag = &ctrl->stop;

    if (tid == 0) {
        if (SMEM) {
            for (int s = 0; s < kMaxStages; ++s) { dev::mbar_init(&full_bar[s], 1); dev::mbar_init(&empty_bar[s], NW); }
            dev::mbar_init(desc_bar, 1);
            dev::fence_barrier_init();
        }
        ctrl->kind = 0; ctrl->count = 0; ctrl->p_kind = 0; ctrl->p_count = 0; ctrl->changed = 0; ctrl->abort = 0;
        ctrl->next = 0; ctrl->error = 0; ctrl->stop = 0; ctrl->consumed = 0;
    }
    __syncthreads();

    // =========================================================================================
    // TMA producer warp: streams [X tile | group record] of upcoming groups into the stage ring
    // =========================================================================================
    if (SMEM && warp == NW) {
        uint32_t sweepno = 0; uint32_t gitem = 0;
        bool running = true;
        while (running) {
            if (!dev::mbar_wait(desc_bar, sweepno & 1, abort_flag, stop_flag)) break;
            const int kind = *reinterpret_cast<volatile int*>(&ctrl->p_kind);
            const int count = *reinterpret_cast<volatile int*>(&ctrl->p_count);
            if (kind == kSweepExit) break;
            for (int it = 0; it < count; ++it, ++gitem) {
                const int stage = gitem % a.n_stages;
                const uint32_t use = gitem / a.n_stages;
                if (!dev::mbar_wait(&empty_bar[stage], (use & 1) ^ 1, abort_flag, stop_flag)) { running = false; break; }
                const int ss = (kind == kSweepActive) ? dev::ld_cg(a.active_set + it) : it;
                const GroupMeta m = a.meta[ss];
                T* xs = stages + (size_t)stage * a.stage_elems;
                T* recs = xs + (size_t)a.rows_stride * a.gs_max;
                const uint32_t col_bytes = (uint32_t)rows * sizeof(T);
                const uint32_t rec_bytes = (uint32_t)m.rec_elems * sizeof(T);
                if (lane == 0) dev::mbar_arrive_expect_tx(&full_bar[stage], col_bytes * m.gs + rec_bytes);
                __syncwarp();
                for (int c = lane; c < m.gs; c += 32)
                    dev::tma_bulk_g2s(xs + (size_t)c * a.rows_stride, a.X + (int64_t)(m.col + c) * a.ld + r0, col_bytes, &full_bar[stage]);
                if (lane == 0) dev::tma_bulk_g2s(recs, a.grec + m.rec_off, rec_bytes, &full_bar[stage]);
            }
            ++sweepno;
        }
        // drain: never leave the CTA while bulk copies into its shared memory are in flight
        if (*stop_flag) {
            const uint32_t consumed = *reinterpret_cast<volatile unsigned*>(&ctrl->consumed);
            for (uint32_t g = consumed; g < gitem; ++g)
                dev::mbar_wait(&full_bar[g % a.n_stages], (g / a.n_stages) & 1, abort_flag, nullptr);
        }
        return;
    }

    // =========================================================================================
    // consumer warps
    // =========================================================================================
    const int ctid = tid;                       // consumer thread id (consumer warps come first)
    const double l1 = a.lmda * a.alpha, l2 = a.lmda * (1.0 - a.alpha);

    // resident r / w tiles
    T* gr = a.resid + r0;
    const T* gw = a.weights + r0;
    if (SMEM) {
        for (int v = ctid; v < rows / VN; v += NTC) {
            T t[VN];
            vec_load<T>(gr + (size_t)v * VN, t); vec_store<T>(sr + (size_t)v * VN, t);
            vec_load<T>(gw + (size_t)v * VN, t); vec_store<T>(sw + (size_t)v * VN, t);
        }
    }
    T* rr = SMEM ? sr : gr;
    const T* ww = SMEM ? sw : gw;

    ProxState ps;
    ps.rsq = a.sc->rsq; ps.resid_sum = a.sc->resid_sum; ps.cm = 0; ps.A = a.sc->active_set_size; ps.error = 0;
    ps.newton_iters_max = 0;
    long long iters = a.sc->iters, n_updates = a.sc->n_group_updates;
    uint32_t epoch = dev::ld_cg(a.epoch);
    uint32_t gitem = 0;
    int phase = kSweepActive;
    int final_error = 0;
    const int nseg = a.ncta_pad / 32;

    while (true) {
        // ---- publish the sweep descriptor
        if (ctid == 0) {
            const int cnt = (phase == kSweepActive) ? ps.A : a.S;
            ctrl->kind = phase; ctrl->count = cnt;
            if (SMEM && cnt > 0) { ctrl->p_kind = phase; ctrl->p_count = cnt; dev::mbar_arrive(desc_bar); }
        }
        ++iters;
        dev::named_bar_sync(1, NTC);
        const int kind = ctrl->kind, count = ctrl->count;
        ps.cm = 0;

        for (int it = 0; it < count; ++it) {
            const int ss = (kind == kSweepActive) ? dev::ld_cg(a.active_set + it) : it;
            const GroupMeta m = a.meta[ss];
            const int gs = m.gs;
            const int stage = SMEM ? (int)(gitem % a.n_stages) : 0;
            const uint32_t use = SMEM ? gitem / a.n_stages : 0;

            // control warp: start fetching the current coefficients (latency hidden behind the dot phase)
            double aold_r[kGsMax / 32];
            if (warp == 0) {
#pragma unroll
                for (int e = 0; e < kGsMax / 32; ++e) {
                    const int c = lane + 32 * e;
                    aold_r[e] = (c < gs) ? (double)dev::ld_cg(a.screen_beta + m.begin + c) : 0.0;
                }
            }

            const T* xs; const T* rec; int64_t cs;
            if (SMEM) {
                if (!dev::mbar_wait(&full_bar[stage], use & 1, abort_flag)) ctrl->abort = 1;
                xs = stages + (size_t)stage * a.stage_elems;
                cs = a.rows_stride;
                rec = xs + (size_t)a.rows_stride * a.gs_max;
            } else {
                xs = a.X + (int64_t)m.col * a.ld + r0;
                cs = a.ld;
                rec = a.grec + m.rec_off;
            }

            // ---- dot phase: partial[c] = sum_{i in tile} X[i, col+c] * w[i] * r[i]
            constexpr int CB = 8;
            for (int c0 = 0; c0 < gs; c0 += CB) {
                T acc[CB];
#pragma unroll
                for (int cc = 0; cc < CB; ++cc) acc[cc] = 0;
                for (int v = ctid; v < rows / VN; v += NTC) {
                    T rv[VN], wv[VN];
                    vec_load<T>(rr + (size_t)v * VN, rv);
                    vec_load<T>(ww + (size_t)v * VN, wv);
                    T wr[VN];
#pragma unroll
                    for (int k = 0; k < VN; ++k) wr[k] = wv[k] * rv[k];
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) {
                        if (c0 + cc < gs) {
                            T xv[VN];
                            vec_load<T>(xs + (int64_t)(c0 + cc) * cs + (size_t)v * VN, xv);
#pragma unroll
                            for (int k = 0; k < VN; ++k) acc[cc] += xv[k] * wr[k];
                        }
                    }
                }
#pragma unroll
                for (int cc = 0; cc < CB; ++cc) {
                    if (c0 + cc < gs) {
                        const double s = dev::warp_sum((double)acc[cc]);
                        if (lane == 0) wpart[(size_t)warp * gsc + c0 + cc] = s;
                    }
                }
            }
            dev::named_bar_sync(1, NTC);
            if (ctrl->abort) { final_error = kErrAbort; break; }      // uniform: flag was set before the barrier

            // ---- exchange: CTA partial -> LL line; then every CTA sums all lines in a fixed order
            const uint32_t par = epoch & 1u;
            if (ctid < gs) {
                double s = 0;
                for (int w = 0; w < NW; ++w) s += wpart[(size_t)w * gsc + ctid];
                if (ncta == 1) wsum[ctid * kLLSeg] = s;
                else dev::ll_store(a.ll + ((size_t)(par * a.ll_gs_cap + ctid) * a.ncta_pad + cta), s, epoch);
            }
            if (ncta > 1) {
                bool ok = true;
                for (int idx = ctid; idx < gs * a.ncta_pad; idx += NTC) {
                    const int c = idx / a.ncta_pad, j = idx - c * a.ncta_pad;
                    double v = 0.0;
                    if (j < ncta) ok = dev::ll_wait(a.ll + ((size_t)(par * a.ll_gs_cap + c) * a.ncta_pad + j), epoch, v, abort_flag) && ok;
                    v = dev::warp_sum(v);
                    if (lane == 0) wsum[c * kLLSeg + (j >> 5)] = v;
                }
                if (!ok) ctrl->abort = 1;
            }
            ++epoch;
            dev::named_bar_sync(1, NTC);
            if (ctrl->abort) { final_error = kErrAbort; break; }      // uniform

            // ---- proximal update (control warp), replicated bit-for-bit in every CTA
            if (warp == 0) {
                const int nsg = (ncta == 1) ? 1 : nseg;
                int changed = 0;
                const double pk = m.pen;
                if (gs == 1) {                                           // solver_gaussian_pin_naive.hpp:75-108
                    double g = 0;
                    for (int sgi = 0; sgi < nsg; ++sgi) g += wsum[sgi];
                    const double ak_old = __shfl_sync(0xffffffffu, aold_r[0], 0);
                    const double A_kk = (double)rec[0], xm = (double)rec[1];
                    double gk = g - xm * ps.resid_sum * (double)a.intercept + ak_old * A_kk;
                    const double vv = fabs(gk) - l1 * pk;                // update_coordinate, pin_base.hpp:181-195
                    double ak = (vv > 0.0) ? copysign(vv, gk) / (A_kk + l2 * pk) : 0.0;
                    ak = (double)(T)ak;                                  // coefficients live in T
                    gk -= ak_old * A_kk;
                    if (ak != ak_old) {
                        const double del = ak - ak_old;
                        ps.cm = fmax(ps.cm, A_kk * del * del);
                        ps.rsq += del * (2 * gk - del * A_kk);
                        ps.resid_sum -= xm * del;
                        if (lane == 0) { a.screen_beta[m.begin] = (T)ak; s_del[0] = (T)(-del); }
                        changed = 1;
                    }
                } else {                                                 // :109-164
                    const T* Arec = rec; const T* xmrec = rec + gs; const T* V = rec + 2 * gs;
#pragma unroll
                    for (int e = 0; e < kGsMax / 32; ++e) {
                        const int c = lane + 32 * e;
                        if (c < gs) {
                            double g = 0;
                            for (int sgi = 0; sgi < nsg; ++sgi) g += wsum[c * kLLSeg + sgi];
                            const double xm = (double)xmrec[c];
                            if (a.intercept) g -= ps.resid_sum * xm;
                            p_gk[c] = g; p_aold[c] = aold_r[e]; p_A[c] = (double)Arec[c]; p_xm[c] = xm;
                        }
                    }
                    __syncwarp();
                    for (int c = lane; c < gs; c += 32) {
                        double gt = 0, ao = 0;
                        for (int r = 0; r < gs; ++r) {
                            const double vrc = (double)V[r * gs + c];
                            gt += p_gk[r] * vrc; ao += p_aold[r] * vrc;
                        }
                        gt += p_A[c] * ao;
                        p_gt[c] = gt; p_atold[c] = ao;
                    }
                    const double l1k = l1 * pk, l2k = l2 * pk;
                    __syncwarp();
                    int nit = 0;
                    warp_prox_newton(p_A, p_gt, gs, l1k, l2k, a.newton_tol, a.newton_max_iters, false, p_D, p_at, nit, lane);
                    ps.newton_iters_max = max(ps.newton_iters_max, nit);
                    if (nit >= a.newton_max_iters) ps.error = kErrNewton;
                    __syncwarp();
                    double dn = 0, cmv = 0, rs = 0;
                    for (int c = lane; c < gs; c += 32) {
                        const double gt0 = p_gt[c] - p_A[c] * p_atold[c];
                        const double d = p_at[c] - p_atold[c];
                        dn += d * d; cmv += p_A[c] * d * d; rs += d * (2 * gt0 - d * p_A[c]);
                    }
                    dn = dev::warp_sum(dn); cmv = dev::warp_sum(cmv); rs = dev::warp_sum(rs);
                    if (!(sqrt(dn) <= a.dbeta_tol * sqrt((double)gs))) {  // :146-147
                        ps.cm = fmax(ps.cm, cmv / gs);
                        ps.rsq += rs;
                        double rsum = 0;
                        for (int r = lane; r < gs; r += 32) {
                            double an = 0;
                            for (int c = 0; c < gs; ++c) an += p_at[c] * (double)V[r * gs + c];
                            const T anT = (T)an;
                            a.screen_beta[m.begin + r] = anT;
                            const double del = p_aold[r] - (double)anT;
                            s_del[r] = (T)del;
                            rsum += p_xm[r] * del;
                        }
                        ps.resid_sum += dev::warp_sum(rsum);
                        changed = 1;
                    }
                }
                if (changed && kind == kSweepScreen) {                   // add_active_set (:294-304)
                    if (!dev::ld_cg(a.is_active + ss)) {
                        if (ps.A >= a.max_active_size) ps.error = kErrMaxActive;
                        else {
                            if (lane == 0) { a.is_active[ss] = 1; a.active_set[ps.A] = ss; }
                            ++ps.A;
                        }
                    }
                }
                if (lane == 0) { ctrl->changed = changed; ctrl->error = ps.error; }
                __syncwarp();
            }
            ++n_updates;
            dev::named_bar_sync(1, NTC);

            // ---- residual update from the same X tile: r += X_g * del
            if (ctrl->changed) {
                for (int v = ctid; v < rows / VN; v += NTC) {
                    T rv[VN];
                    vec_load<T>(rr + (size_t)v * VN, rv);
                    for (int c = 0; c < gs; ++c) {
                        const T d = s_del[c];
                        T xv[VN];
                        vec_load<T>(xs + (int64_t)c * cs + (size_t)v * VN, xv);
#pragma unroll
                        for (int k = 0; k < VN; ++k) rv[k] += xv[k] * d;
                    }
                    vec_store<T>(rr + (size_t)v * VN, rv);
                }
            }
            const int err_now = ctrl->error;
            if (SMEM) {
                __syncwarp();
                if (lane == 0) dev::mbar_arrive(&empty_bar[stage]);
            }
            ++gitem;
            if (err_now) { final_error = err_now; break; }
        }
        if (final_error) break;

        // ---- end of sweep: the control thread decides what comes next (identically in every CTA)
        if (ctid == 0) {
            int next;
            const bool conv = ps.cm < a.tol;
            if (kind == kSweepActive) next = conv ? kSweepScreen : ((iters >= a.max_iters) ? -kErrMaxCds : kSweepActive);
            else next = conv ? kSweepExit : ((iters >= a.max_iters) ? -kErrMaxCds : kSweepActive);
            ctrl->next = next;
        }
        dev::named_bar_sync(1, NTC);
        const int next = ctrl->next;
        dev::named_bar_sync(1, NTC);      // everyone has read `next` before ctrl is rewritten
        if (next < 0) { final_error = -next; break; }
        if (next == kSweepExit) break;
        phase = next;
    }

    // ---- shut the producer down, write results back
    if (SMEM) {
        dev::named_bar_sync(1, NTC);
        if (ctid == 0) {
            if (final_error) { ctrl->consumed = gitem; __threadfence_block(); ctrl->stop = 1; }
            else { ctrl->p_kind = kSweepExit; ctrl->p_count = 0; dev::mbar_arrive(desc_bar); }
        }
        if (final_error == kErrAbort) *abort_flag = 1;
        for (int v = ctid; v < rows / VN; v += NTC) {
            T t[VN];
            vec_load<T>(sr + (size_t)v * VN, t); vec_store<T>(gr + (size_t)v * VN, t);
        }
    }
    if (cta == 0 && ctid == 0) {
        a.sc->rsq = ps.rsq; a.sc->resid_sum = ps.resid_sum; a.sc->active_set_size = ps.A;
        a.sc->iters = iters; a.sc->n_group_updates = n_updates; a.sc->error = final_error;
        a.sc->newton_iters_max = ps.newton_iters_max;
        *a.epoch = epoch;
    }
}

} // namespace ab
