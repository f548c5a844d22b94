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
constexpr int kMaxRanksDev = 8;        // ranks of one NVSwitch box
constexpr int kLLSeg = 8;              // max 32-CTA segments => up to 256 CTAs

struct GroupMeta {          // one per screen position (32 bytes)
    int32_t col;            // first column of the group in X  (multi-response intercept column of class l: -(l + 1))
    int32_t gs;             // group size
    int32_t begin;          // offset into screen_beta
    int32_t rec_elems;      // padded length of the group's record
    int64_t rec_off;        // element offset of the record [A(gs) | xm(gs) | V(gs*gs, (r,c)->r*gs+c)]
    double pen;             // penalty factor
};

struct PinScalars {         // device-resident in/out scalars of one pin solve
    double rsq, resid_sum;
    long long iters, n_group_updates, n_col_updates;   // sweeps, group visits, sum of group sizes over the visits
    int active_set_size, error, newton_iters_max, pad;
};

enum { kErrNone = 0, kErrMaxCds = 1, kErrMaxActive = 2, kErrNewton = 3, kErrAbort = 4 };
enum { kSweepActive = 0, kSweepScreen = 1, kSweepExit = 2 };

template <class T>
struct PinKernelArgs {
    const T* X; int64_t ld; int64_t n_pad;
    // multi-response layout (kronecker_eye / concatenate of the reference as a LAYOUT RULE): resid and weights are
    // (n, K) row-major and GroupMeta.col is the first column of the group in kron(X, I_K): coefficient a of the group <-> feature
    // (col + a) / K of X, class (col + a) % K ("grouped": K classes of one feature; "ungrouped": a single (feature, class) pair).
    // The K intercept columns kron(1, I_K) carry GroupMeta.col = -(class + 1), gs = 1.  K = 1: single response, col = X column.
    int K;
    T* resid; const T* weights;
    const GroupMeta* meta; int S;
    const T* grec;
    // Coefficients: every CTA works on its OWN replica (beta_rep + cta * beta_stride), initialised from beta_in at
    // kernel start; replica 0 is the output.  A single shared array would be racy: a slow CTA's (identical-valued but)
    // delayed store of the previous sweep's coefficients could land after a fast CTA's newer store.
    const T* beta_in; T* beta_rep; int64_t beta_stride; int beta_len;
    // screen_is_active is replicated per CTA for the same reason (a fast CTA's 0 -> 1 transition must not be seen
    // by a slower CTA that has not processed that group yet); replica 0 is the output.  active_set slots are written
    // once per launch with identical values by every CTA, which is race-free.
    const int8_t* is_active_in; int8_t* is_active_rep; int64_t act_stride;
    int32_t* active_set;
    PinScalars* sc;
    dev::LLLine* ll; int ll_gs_cap; int ncta_pad;
    dev::LLLine* ll2; int fan;       // two-level exchange: level-2 lines [2][ll_gs_cap][32]; fan = CTAs per level-1 group
    // one-hop exchange through L2 atomics (Configs::sweep_xchg == 1): xs_sum [4 slots][ll_gs_cap] doubles, xs_cnt [4 slots][32] arrival
    // counters (one 128-byte line each); all zero at kernel start, slot = epoch & 3
    double* xs_sum; uint32_t* xs_cnt; int xchg_atomic;
    // multi-GPU (row-sharded) level 3: ll3_peer[r] = rank r's line buffer [2][kGsMax][kMaxRanksDev] mapped over NVLink
    dev::LLLine* ll3_peer[8]; int rank, world;
    uint32_t* epoch; int* abort_flag;
    double lmda, alpha, tol, newton_tol, dbeta_tol;
    long long max_iters; int newton_max_iters; int max_active_size; int intercept;
    int units_base, units_rem;     // row partition in units of kRowAlign rows
    int rows_stride;               // max rows per CTA (column stride inside a stage)
    int n_stages; int stage_elems; // ring geometry (elements of T per stage)
    int gs_max;                    // largest group size in the screen set
    int feat_max;                  // largest number of physical X columns behind one group (= gs_max / K, at least 1)
    int gs_cap;                    // stride of the per-column shared-memory scratch arrays (>= gs_max, multiple of 4)
    long long* stats;              // optional [16] per-phase cycle counters of CTA 0 / thread 0 (nullptr = off)
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

// Shared-memory carve-up (host and device must agree):
//   header (barriers, ctrl) | vals1[gs_cap][32] + vals2[gs_cap][32] + gsum[gs_cap] (double) | wpart[n_cwarps][gs_cap] (double) |
//   prox scratch 8 x [gs_cap] (double-sized slots) | del [gs_cap] | reduction scratch [4*32] | r tile | w tile | stages
template <class T>
struct SweepSmem {
    static constexpr size_t kHeaderBytes = 512;
    __host__ __device__ static size_t fixed_bytes(int n_cwarps, int gs_cap, int ncta_pad) {
        size_t b = kHeaderBytes
                 + sizeof(double) * (size_t)gs_cap * (2 * 32 + 1)      // vals1 + vals2 + gsum  (ncta_pad unused: two-level exchange)
                 + sizeof(double) * (size_t)n_cwarps * gs_cap         // wpart
                 + sizeof(double) * (size_t)gs_cap * 8                // prox scratch
                 + sizeof(double) * (size_t)gs_cap                    // del
                 + sizeof(double) * 4 * 32;                           // reduction scratch of the control warp
        return (b + 127) / 128 * 128;
    }
    static size_t total(int n_cwarps, int gs_cap, int ncta_pad, int rows_stride, int n_stages, int stage_elems) {
        return fixed_bytes(n_cwarps, gs_cap, ncta_pad) + 2 * sizeof(T) * (size_t)rows_stride + sizeof(T) * (size_t)n_stages * stage_elems;
    }
};

// Replicated solver scalars, held by the control warp of every CTA (identical in all CTAs).
struct ProxState {
    double rsq, resid_sum, cm;
    int A;                    // active_set_size
    int error;
    int newton_iters_max;
};

// ---- multi-value warp reductions ---------------------------------------------------------------
// Simultaneous butterfly all-reduce of N values (independent chains => the shuffles pipeline).
template <int N, class P>
__device__ __forceinline__ void warp_allsum(P (&v)[N]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        P t[N];
#pragma unroll
        for (int i = 0; i < N; ++i) t[i] = __shfl_xor_sync(0xffffffffu, v[i], o);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] += t[i];
    }
}

// "Transposed" reduction of 16 per-lane accumulators over the 32 lanes in 16 shuffles / 5 steps:
// every step halves the number of live values per lane.  Returns, in every lane, the warp total of
// column (lane >> 1) & 15.
template <class P>
__device__ __forceinline__ P warp_reduce16(P (&v)[16], int lane) {
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const P send = up ? v[i] : v[i + 8];
            const P keep = up ? v[i + 8] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const P send = up ? v[i] : v[i + 4];
            const P keep = up ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const P send = up ? v[i] : v[i + 2];
            const P keep = up ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool up = (lane & 2) != 0;
        const P send = up ? v[0] : v[1];
        const P keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

// All-reduce of N per-lane values over the first n_src lanes of a warp through shared memory: every lane ends up with
// the totals after n_src dependent adds (~4 cycles each) instead of a 5-step shuffle butterfly (~30 cycles per step).
// Lanes >= n_src must pass zeros.  scratch: N * 32 elements, private to the warp.
template <int N, class P>
__device__ __forceinline__ void smem_allsum(P* scratch, P (&v)[N], int n_src, int lane) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < N; ++k) scratch[k * 32 + lane] = v[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = 0;
#pragma unroll 4
    for (int i = 0; i < n_src; ++i) {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] += scratch[k * 32 + i];
    }
}

template <class P> struct ProxEps;
template <> struct ProxEps<double> { static __device__ __forceinline__ double floor_tol() { return 0.0; } };
// float cannot resolve |phi(h)| below a few ulps of 1: the Newton tolerance is clamped at this floor
template <> struct ProxEps<float> { static __device__ __forceinline__ float floor_tol() { return 4.8e-7f; } };

// Warp-cooperative Newton solve in the compute type P (float for T = float, double for T = double), tuned for the
// sweep's critical path: the ||v|| <= l1 test is folded into the first evaluation (phi(0) = ||v||^2 / l1^2 - 1) and
// the two sums of every evaluation travel through one interleaved butterfly.  Same iteration as newton_solver
// (newton.hpp:44-142, h0 = 0).  L, v, x: shared arrays of length >= gs.  Returns the iteration count.
template <class P>
__device__ __forceinline__ int warp_prox_fast(const P* L, const P* v, int gs, P l1, P l2, P tol, int max_iters, P* x, int lane, P* scratch)
{
    const int n_src = min(gs, 32);
    if (l1 <= P(0)) {                                        // newton.hpp:72-75 (||v|| <= l1 = 0 only if v == 0 => x = 0 too)
#pragma unroll 1
        for (int c = lane; c < gs; c += 32) { const P d = L[c] + l2; x[c] = (v[c] == P(0)) ? P(0) : v[c] / d; }
        __syncwarp();
        return 0;
    }
    const P tol_eff = fmax(tol, ProxEps<P>::floor_tol());
    P h = 0; int iters = 0;
    while (true) {
        P s[2] = {0, 0};
#pragma unroll 1
        for (int c = lane; c < gs; c += 32) {
            const P D = L[c] + l2;
            const P u = P(1) / (D * h + l1);
            const P q = v[c] * u;
            const P xx = q * q;
            s[0] += xx; s[1] += xx * D * u;
        }
        smem_allsum<2>(scratch, s, n_src, lane);
        const P t = s[0];
        if (iters == 0 && !(t > P(1))) {                     // ||v|| <= l1  =>  x = 0   (newton.hpp:62-66)
#pragma unroll 1
            for (int c = lane; c < gs; c += 32) x[c] = 0;
            __syncwarp();
            return 0;
        }
        const P fh = t - P(1);
        if (!(fabs(fh) > tol_eff) || iters >= max_iters) break;
        const P dfh = -s[1] * (P(1) + sqrt(t)) / t;
        h = fmax(h - fh / dfh, P(0));
        ++iters;
    }
#pragma unroll 1
    for (int c = lane; c < gs; c += 32) x[c] = h * v[c] / ((L[c] + l2) * h + l1);       // newton.hpp:109
    __syncwarp();
    return iters;
}

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
        if (mode == 0) iters = warp_prox_fast<double>(L, v, q, l1, l2, tol, max_iters, x, lane, sh + 4 * q);    // the sweep kernel's solver
        else warp_prox_newton(L, v, q, l1, l2, tol, max_iters, true, D, x, iters, lane);
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

// ---- out-of-line phases of the sweep (kept out of line on purpose: the per-group loop of the kernel must stay
// well inside the 32 KB instruction cache, otherwise the single control warp stalls on instruction fetch) ---------
// Polls `count` lines (line q of this thread = first + q * stride, q = 0 .. while < count) until they carry `epoch` and
// writes their payloads to out[line].  Returns false if the kernel was aborted.
__device__ __forceinline__ bool ll_poll_lines(const dev::LLLine* lines, double* out, int first, int stride, int count, uint32_t epoch,
                                              volatile int* abort_flag)
{
    bool ok = true;
#pragma unroll 1
    for (int i = first; i < count; i += stride) {
        double v = 0.0;
        dev::SpinGuard guard;
        while (!dev::ll_try_load(lines + i, epoch, v)) {
            if (guard.give_up(abort_flag, nullptr)) { ok = false; break; }
        }
        out[i] = v;
    }
    return ok;
}

template <class T, class P>
struct ProxCtx { P *p_aold, *p_A, *p_gk, *p_gt, *p_atold, *p_at, *scratch; T* s_del; const double* gsum; T* my_beta; };

// Group update for 1 < gs <= 32, one coefficient per lane of the control warp, written for the shortest possible
// dependent-instruction chain (a lone warp issues ~1 instruction every 5 cycles): everything that does not depend on the
// exchanged gradient (a_old V, A, xm, V^T xm) is computed by prox_small_pre() while the exchange is in flight.
template <class P> struct ProxPre { P aold, A, xm, xmt, ao; };

template <class T, class P>
__device__ __forceinline__ ProxPre<P> prox_small_pre(const T* rec, int gs, const P* p_aold, int lane) {
    ProxPre<P> r;
    const bool on = lane < gs;
    const int c = on ? lane : 0;
    r.aold = on ? p_aold[c] : P(0);
    r.A = on ? (P)rec[c] : P(0);
    r.xm = on ? (P)rec[gs + c] : P(0);
    r.xmt = on ? (P)rec[2 * gs + c] : P(0);
    const T* V = rec + 3 * gs;
    P ao = 0;
#pragma unroll 4
    for (int q = 0; q < gs; ++q) ao += p_aold[q] * (P)V[q * gs + c];
    r.ao = on ? ao : P(0);
    return r;
}

template <class T, class P>
__device__ __forceinline__ int prox_small_post(const ProxCtx<T, P>& px, const ProxPre<P>& pre, const T* rec, int gs, int begin, P l1k, P l2k,
                                               P tol, int max_iters, P dbeta_tol, int intercept, ProxState& ps, int lane, long long* nit_acc)
{
    const bool on = lane < gs;
    const int c = on ? lane : 0;
    const T* V = rec + 3 * gs;
    P* p_gk = px.p_gk; P* p_at = px.p_at; P* scr = px.scratch;
    // gradient in the original basis, then rotated: gt = gk V + A * (a_old V)
    P gk = on ? (P)px.gsum[c] : P(0);
    if (intercept) gk -= (P)ps.resid_sum * pre.xm;
    p_gk[lane] = gk;
    __syncwarp();
    P gt0 = 0;
#pragma unroll 4
    for (int q = 0; q < gs; ++q) gt0 += p_gk[q] * (P)V[q * gs + c];
    if (!on) gt0 = 0;
    const P gt = gt0 + pre.A * pre.ao;
    // ---- newton_solver (newton.hpp:44-142), h0 = 0, phi(0) doubles as the ||v|| <= l1 test
    P at = 0; int nit = 0;
    if (l1k <= P(0)) {
        at = (on && gt != P(0)) ? gt / (pre.A + l2k) : P(0);
    } else {
        const P D = pre.A + l2k;
        const P tol_eff = fmax(tol, ProxEps<P>::floor_tol());
        P h = 0; bool zero = false;
        while (true) {
            const P u = P(1) / (D * h + l1k);
            const P q = gt * u;
            const P xx = q * q;
            P sm[2] = {xx, xx * D * u};
            smem_allsum<2>(scr, sm, gs, lane);
            const P t = sm[0];
            if (nit == 0 && !(t > P(1))) { zero = true; break; }
            const P fh = t - P(1);
            if (!(fabs(fh) > tol_eff) || nit >= max_iters) break;
            const P dfh = -sm[1] * (P(1) + sqrt(t)) / t;
            h = fmax(h - fh / dfh, P(0));
            ++nit;
        }
        at = (zero || !on) ? P(0) : h * gt / (D * h + l1k);
    }
    ps.newton_iters_max = max(ps.newton_iters_max, nit);
    if (nit_acc) *nit_acc += nit;
    if (nit >= max_iters) ps.error = kErrNewton;
    const P d = at - pre.ao;
    P red[4] = {d * d, pre.A * d * d, d * (2 * gt0 - d * pre.A), -pre.xmt * d};
    smem_allsum<4>(scr, red, gs, lane);
    if (sqrt(red[0]) <= dbeta_tol * sqrt((P)gs)) return 0;       // :146-147
    ps.cm = fmax(ps.cm, (double)(red[1] / gs));
    ps.rsq += (double)red[2];
    ps.resid_sum += (double)red[3];
    p_at[lane] = at;
    __syncwarp();
    P an = 0;
#pragma unroll 4
    for (int q = 0; q < gs; ++q) an += p_at[q] * (P)V[c * gs + q];          // rotate back: a = at V^T
    if (on) {
        const T anT = (T)an;
        px.my_beta[begin + c] = anT;
        px.s_del[c] = (T)(pre.aold - (P)anT);
    }
    return 1;
}

// Group update for gs > 1 (solver_gaussian_pin_naive.hpp:109-164), executed by the control warp.
// rec = [A | xm | V^T xm | V].  Returns 1 if the coefficients moved.
template <class T, class P>
__device__ __noinline__ int prox_group(const ProxCtx<T, P>& px, const T* rec, int gs, int begin, P l1k, P l2k, P tol,
                                       int max_iters, P dbeta_tol, int intercept, ProxState& ps, int lane, long long* nit_acc)
{
    const T* Arec = rec; const T* xmrec = rec + gs; const T* xmtrec = rec + 2 * gs; const T* V = rec + 3 * gs;
    P* p_aold = px.p_aold; P* p_A = px.p_A; P* p_gk = px.p_gk; P* p_gt = px.p_gt; P* p_atold = px.p_atold; P* p_at = px.p_at;
    const P rsum_in = (P)ps.resid_sum;
#pragma unroll 1
    for (int c = lane; c < gs; c += 32) {
        P gk = (P)px.gsum[c];
        if (intercept) gk -= rsum_in * (P)xmrec[c];
        p_gk[c] = gk; p_A[c] = (P)Arec[c];
    }
    __syncwarp();
#pragma unroll 1
    for (int c = lane; c < gs; c += 32) {                   // rotate into the eigenbasis: gt = gk V, ao = a_old V
        P gt = 0, ao = 0;
#pragma unroll 2
        for (int r = 0; r < gs; ++r) {
            const P vrc = (P)V[r * gs + c];
            gt += p_gk[r] * vrc; ao += p_aold[r] * vrc;
        }
        gt += p_A[c] * ao;
        p_gt[c] = gt; p_atold[c] = ao;
    }
    __syncwarp();
    const int nit = warp_prox_fast<P>(p_A, p_gt, gs, l1k, l2k, tol, max_iters, p_at, lane, px.scratch);
    ps.newton_iters_max = max(ps.newton_iters_max, nit);
    if (nit_acc) *nit_acc += nit;
    if (nit >= max_iters) ps.error = kErrNewton;
    P red[4] = {0, 0, 0, 0};                                 // ||dt||^2, sum A dt^2, rsq increment, resid_sum increment
#pragma unroll 1
    for (int c = lane; c < gs; c += 32) {
        const P gt0 = p_gt[c] - p_A[c] * p_atold[c];
        const P d = p_at[c] - p_atold[c];
        red[0] += d * d; red[1] += p_A[c] * d * d; red[2] += d * (2 * gt0 - d * p_A[c]);
        red[3] -= (P)xmtrec[c] * d;                          // sum_r xm_r (a_old - a)_r = (V^T xm) . (at_old - at)
    }
    smem_allsum<4>(px.scratch, red, min(gs, 32), lane);
    if (sqrt(red[0]) <= dbeta_tol * sqrt((P)gs)) return 0;   // :146-147
    ps.cm = fmax(ps.cm, (double)(red[1] / gs));
    ps.rsq += (double)red[2];
    ps.resid_sum += (double)red[3];
#pragma unroll 1
    for (int r = lane; r < gs; r += 32) {                   // rotate back: a = at V^T
        P an = 0;
#pragma unroll 2
        for (int c = 0; c < gs; ++c) an += p_at[c] * (P)V[r * gs + c];
        const T anT = (T)an;
        px.my_beta[begin + r] = anT;
        px.s_del[r] = (T)(p_aold[r] - (P)anT);
    }
    return 1;
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
    using P = T;                                     // compute type of the replicated proximal update
    const int gsc = a.gs_cap;
    double* vals1 = reinterpret_cast<double*>(smem_raw + SweepSmem<T>::kHeaderBytes);       // [gsc][32] level-1 partials (group leaders)
    double* vals2 = vals1 + (size_t)gsc * 32;                                               // [gsc][32] level-2 partials
    double* gsum = vals2 + (size_t)gsc * 32;                                                // [gsc] all-reduced gradient
    double* wpart = gsum + gsc;                                                             // [NW][gsc]
    P* px = reinterpret_cast<P*>(wpart + (size_t)NW * gsc);                                 // 8 x [gsc] prox scratch
    P* p_aold = px, *p_A = px + gsc, *p_gk = px + 2 * gsc, *p_gt = px + 3 * gsc, *p_atold = px + 4 * gsc, *p_at = px + 5 * gsc;
    T* s_del = reinterpret_cast<T*>(reinterpret_cast<double*>(px) + 8 * gsc);               // [gsc]
    P* p_scr = reinterpret_cast<P*>(reinterpret_cast<double*>(px) + 9 * gsc);               // [4 * 32] reduction scratch
    unsigned char* tiles = smem_raw + SweepSmem<T>::fixed_bytes(NW, gsc, a.ncta_pad);
    T* sr = reinterpret_cast<T*>(tiles);
    T* sw = sr + (size_t)a.rows_stride * a.K;
    T* stages = sw + (size_t)a.rows_stride * a.K;

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
                T* recs = xs + (size_t)a.rows_stride * a.feat_max;
                const uint32_t col_bytes = (uint32_t)rows * sizeof(T);
                const uint32_t rec_bytes = (uint32_t)m.rec_elems * sizeof(T);
                const int f0 = (m.col >= 0) ? m.col / a.K : 0;               // physical X columns behind this group: [f0, f0 + nfeat)
                const int nfeat = (m.col >= 0) ? (m.col - f0 * a.K + m.gs + a.K - 1) / a.K : 0;
                if (lane == 0) dev::mbar_arrive_expect_tx(&full_bar[stage], col_bytes * nfeat + rec_bytes);
                __syncwarp();
                for (int c = lane; c < nfeat; c += 32)
                    dev::tma_bulk_g2s(xs + (size_t)c * a.rows_stride, a.X + (int64_t)(f0 + c) * a.ld + r0, col_bytes, &full_bar[stage]);
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
    const P l1 = (P)(a.lmda * a.alpha), l2 = (P)(a.lmda * (1.0 - a.alpha));
    // polling pattern of the LL exchange: thread t owns lines t, t + NTC, ... (line = column * ncta_pad + cta)
    // two-level exchange geometry: level-1 groups of `fan` consecutive CTAs, led by their first member
    const int fan = a.fan;
    const int my_group = cta / fan, n_groups = (ncta + fan - 1) / fan;
    const bool xatomic = a.xchg_atomic != 0;
    const int grp_first = my_group * fan, grp_size = min(fan, ncta - grp_first);
    const bool is_leader = (cta == grp_first);

    // resident r / w tiles
    const int K = a.K;
    T* gr = a.resid + r0 * K;
    const T* gw = a.weights + r0 * K;
    if (SMEM) {
        for (int v = ctid; v < rows * K / VN; v += NTC) {
            T t[VN];
            vec_load<T>(gr + (size_t)v * VN, t); vec_store<T>(sr + (size_t)v * VN, t);
            vec_load<T>(gw + (size_t)v * VN, t); vec_store<T>(sw + (size_t)v * VN, t);
        }
    }
    T* rr = SMEM ? sr : gr;
    const T* ww = SMEM ? sw : gw;
    T* my_beta = a.beta_rep + (size_t)cta * a.beta_stride;
    for (int i = ctid; i < a.beta_len; i += NTC) my_beta[i] = a.beta_in[i];
    int8_t* my_active = a.is_active_rep + (size_t)cta * a.act_stride;
    const ProxCtx<T, P> proxctx{p_aold, p_A, p_gk, p_gt, p_atold, p_at, p_scr, s_del, gsum, my_beta};
    for (int i = ctid; i < a.S; i += NTC) my_active[i] = a.is_active_in[i];
    // (visibility to the control warp is ordered by the named barrier at the top of the first sweep)

    ProxState ps;
    ps.rsq = a.sc->rsq; ps.resid_sum = a.sc->resid_sum; ps.cm = 0; ps.A = a.sc->active_set_size; ps.error = 0;
    ps.newton_iters_max = 0;
    long long iters = a.sc->iters, n_updates = a.sc->n_group_updates, n_cols = a.sc->n_col_updates;
    uint32_t epoch = dev::ld_cg(a.epoch);
    uint32_t gitem = 0;
    int pending_stage = -1;
    int phase = kSweepActive;
    int final_error = 0;
    // phase profiling (thread 0 of CTA 0 only): 0 wait-full, 1 dot, 2 bar1+store, 3 poll, 4 prox, 5 bar3, 6 update, 7 newton iters, 8 items
    const bool prof = (a.stats != nullptr) && cta == 0 && ctid == 0;
    long long pt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tc = 0;
#define AB_TICK(k) do { if (prof) { const long long t_ = clock64(); pt[k] += t_ - tc; tc = t_; } } while (0)

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
            if (warp == 0) {
#pragma unroll 1
                for (int c = lane; c < gs; c += 32) p_aold[c] = (P)my_beta[m.begin + c];
            }

            if (prof) tc = clock64();
            const T* xs; const T* rec; int64_t cs;
            if (SMEM) {
                if (!dev::mbar_wait(&full_bar[stage], use & 1, abort_flag)) ctrl->abort = 1;
                xs = stages + (size_t)stage * a.stage_elems;
                cs = a.rows_stride;
                rec = xs + (size_t)a.rows_stride * a.feat_max;
            } else {
                xs = a.X + (int64_t)(m.col >= 0 ? m.col / a.K : 0) * a.ld + r0;
                cs = a.ld;
                rec = a.grec + m.rec_off;
            }

            AB_TICK(0);
            // ---- dot phase: partial[c] = sum_{i in tile} X[i, col+c] * w[i] * r[i]
            constexpr int CB = 16;
            if (K == 1) {
#pragma unroll 1
            for (int c0 = 0; c0 < gs; c0 += CB) {
                T acc[CB];
#pragma unroll
                for (int cc = 0; cc < CB; ++cc) acc[cc] = 0;
#pragma unroll 1
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
                const T tot = warp_reduce16<T>(acc, lane);       // lane holds the warp total of column c0 + ((lane >> 1) & 15)
                const int col = c0 + ((lane >> 1) & 15);
                if ((lane & 1) == 0 && col < gs) wpart[(size_t)warp * gsc + col] = (double)tot;
            }
            } else {
                // multi-response: coefficient a <-> kron column m.col + a = (feature, class); X[i, feature] * w[i, class] * r[i, class]
                const bool icpt = m.col < 0;
                const int kcls = icpt ? (-m.col - 1) : 0;
                const int k0 = icpt ? 0 : m.col % K;
                if (!icpt && k0 == 0 && gs == K && (K % VN) == 0) {
                    // one feature x all K classes (the snp_unphased / multigaussian layout of config 5): a row's K residuals and
                    // weights are contiguous, so they come in with 16-byte loads and one X element serves all K products
                    T acc[CB];
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) acc[cc] = 0;
#pragma unroll 2
                    for (int i = ctid; i < rows; i += NTC) {
                        const T x = xs[i];
                        const T* rrow = rr + (size_t)i * K; const T* wrow = ww + (size_t)i * K;
#pragma unroll
                        for (int kv = 0; kv < CB / VN; ++kv) {
                            if (kv * VN < K) {
                                T rv[VN], wv[VN];
                                vec_load<T>(rrow + kv * VN, rv); vec_load<T>(wrow + kv * VN, wv);
#pragma unroll
                                for (int k = 0; k < VN; ++k) acc[kv * VN + k] += x * (wv[k] * rv[k]);
                            }
                        }
                    }
                    const T tot = warp_reduce16<T>(acc, lane);
                    const int col = (lane >> 1) & 15;
                    if ((lane & 1) == 0 && col < gs) wpart[(size_t)warp * gsc + col] = (double)tot;
                } else
#pragma unroll 1
                for (int c0 = 0; c0 < gs; c0 += CB) {
                    T acc[CB]; int fo[CB], ko[CB];
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) {
                        acc[cc] = 0;
                        const int ai = c0 + cc;
                        const int f = icpt ? 0 : (k0 + ai) / K;
                        fo[cc] = f; ko[cc] = icpt ? kcls : k0 + ai - f * K;
                    }
#pragma unroll 1
                    for (int i = ctid; i < rows; i += NTC) {
                        const T* rrow = rr + (size_t)i * K; const T* wrow = ww + (size_t)i * K;
#pragma unroll
                        for (int cc = 0; cc < CB; ++cc) {
                            if (c0 + cc < gs) {
                                const T wr = wrow[ko[cc]] * rrow[ko[cc]];
                                acc[cc] += icpt ? wr : xs[(int64_t)fo[cc] * cs + i] * wr;
                            }
                        }
                    }
                    const T tot = warp_reduce16<T>(acc, lane);
                    const int col = c0 + ((lane >> 1) & 15);
                    if ((lane & 1) == 0 && col < gs) wpart[(size_t)warp * gsc + col] = (double)tot;
                }
            }
            AB_TICK(1);
            long long* trace = (a.stats != nullptr && ctid == 0 && n_updates == a.sc->n_group_updates + 40) ? a.stats + 32 + 8 * cta : nullptr;
            if (trace) { trace[0] = (long long)(dev::global_ns() & 0xffffffffffull); }
            dev::named_bar_sync(1, NTC);
            if (ctrl->abort) { final_error = kErrAbort; break; }      // uniform: flag was set before the barrier

            // ---- two-level exchange.  Level 1: every CTA publishes its partial (one flagged line per column); the leader
            // of each group of `fan` CTAs reads its members' lines, adds them in member order and publishes the group
            // partial.  Level 2: every CTA reads the n_groups group partials and adds them in group order.  Every CTA thus
            // obtains the bitwise identical gradient while reading O(sqrt(#CTAs)) lines instead of O(#CTAs).
            const uint32_t par = epoch & 1u;
            if (ctid < gs) {
                double s = 0;
#pragma unroll 5
                for (int w = 0; w < NW; ++w) s += wpart[(size_t)w * gsc + ctid];
                if (ncta == 1) gsum[ctid] = s;
                else if (xatomic) {
                    // one hop: add the partial into the slot's accumulator at L2, then count the arrival (release: the add is performed first)
                    dev::red_add_f64(a.xs_sum + (size_t)(epoch & 3u) * a.ll_gs_cap + ctid, s);
                    dev::red_release_add_u32(a.xs_cnt + (size_t)(epoch & 3u) * 32, 1u);
                }
                else dev::ll_store(a.ll + ((size_t)(par * a.ll_gs_cap + ctid) * a.ncta_pad + cta), s, epoch);
            }
            AB_TICK(2);
            if (trace) { trace[1] = (long long)(dev::global_ns() & 0xffffffffffull); }
            ProxPre<P> pre{};
            const bool small_group = (gs > 1 && gs <= 32);
            if (warp == 0 && small_group) pre = prox_small_pre<T, P>(rec, gs, p_aold, lane);      // overlaps the exchange latency
            const bool multi_gpu = a.world > 1;
            const int n_red = xatomic ? 1 : n_groups;                    // partials per column a CTA holds after the intra-GPU exchange
            int n_final = n_red;                                         // number of partials the control warp adds up
            if (ncta > 1 || multi_gpu) {
                bool ok = true;
                if (xatomic && ncta > 1 && (!multi_gpu || cta == 0)) {   // wait until all ncta * gs arrivals are in, read the totals
                    if (ctid < gs) {
                        dev::SpinGuard guard;
                        const uint32_t want = (uint32_t)ncta * (uint32_t)gs;
                        const uint32_t* cnt = a.xs_cnt + (size_t)(epoch & 3u) * 32;
                        while (dev::ld_acquire_u32(cnt) != want) { if (guard.give_up(abort_flag, nullptr)) { ok = false; break; } }
                        vals2[ctid * 32] = dev::ld_relaxed_f64(a.xs_sum + (size_t)(epoch & 3u) * a.ll_gs_cap + ctid);
                    }
                }
                if (!xatomic && ncta > 1 && is_leader) {                 // CTA-uniform
                    // thread t < gs * grp_size reads line (column t / grp_size, member t % grp_size)
#pragma unroll 1
                    for (int t = ctid; t < gs * grp_size; t += NTC) {
                        const int c = t / grp_size, mth = t - c * grp_size;
                        double v = 0.0;
                        dev::SpinGuard guard;
                        const dev::LLLine* line = a.ll + ((size_t)(par * a.ll_gs_cap + c) * a.ncta_pad + grp_first + mth);
                        while (!dev::ll_try_load(line, epoch, v)) { if (guard.give_up(abort_flag, nullptr)) { ok = false; break; } }
                        vals1[c * 32 + mth] = v;
                    }
                    dev::named_bar_sync(1, NTC);
                    if (ctid < gs) {
                        double s = 0;
#pragma unroll 1
                        for (int mth = 0; mth < grp_size; ++mth) s += vals1[ctid * 32 + mth];
                        dev::ll_store(a.ll2 + ((size_t)(par * a.ll_gs_cap + ctid) * 32 + my_group), s, epoch);
                    }
                }
                if (!xatomic && ncta > 1 && (!multi_gpu || cta == 0)) {  // level 2 (single GPU: every CTA; multi GPU: the GPU leader only)
#pragma unroll 1
                    for (int t = ctid; t < gs * n_groups; t += NTC) {
                        const int c = t / n_groups, g = t - c * n_groups;
                        double v = 0.0;
                        dev::SpinGuard guard;
                        const dev::LLLine* line = a.ll2 + ((size_t)(par * a.ll_gs_cap + c) * 32 + g);
                        while (!dev::ll_try_load(line, epoch, v)) { if (guard.give_up(abort_flag, nullptr)) { ok = false; break; } }
                        vals2[c * 32 + g] = v;
                    }
                }
                if (multi_gpu) {
                    // level 3 over NVLink: the GPU leader (CTA 0) stores this GPU's partial into EVERY rank's line buffer
                    // (peer-to-peer stores through NVSwitch); every CTA of every GPU then reads its own GPU's `world` lines.
                    if (cta == 0) {
                        dev::named_bar_sync(1, NTC);                     // vals2 (group partials) complete
                        if (ctid < gs) {
                            double s = 0;
                            if (ncta > 1) {
#pragma unroll 1
                                for (int g = 0; g < n_red; ++g) s += vals2[ctid * 32 + g];
                            } else s = gsum[ctid];
                            vals1[ctid * 32] = s;
                        }
                        dev::named_bar_sync(1, NTC);
#pragma unroll 1
                        for (int t = ctid; t < gs * a.world; t += NTC) {
                            const int c = t / a.world, r = t - c * a.world;
                            dev::ll_store_sys(a.ll3_peer[r] + ((size_t)(par * kGsMax + c) * kMaxRanksDev + a.rank), vals1[c * 32], epoch);
                        }
                        dev::named_bar_sync(1, NTC);                     // vals2 may be overwritten below
                    }
#pragma unroll 1
                    for (int t = ctid; t < gs * a.world; t += NTC) {
                        const int c = t / a.world, r = t - c * a.world;
                        double v = 0.0;
                        dev::SpinGuard guard;
                        const dev::LLLine* line = a.ll3_peer[a.rank] + ((size_t)(par * kGsMax + c) * kMaxRanksDev + r);
                        while (!dev::ll_try_load_sys(line, epoch, v)) { if (guard.give_up(abort_flag, nullptr)) { ok = false; break; } }
                        vals2[c * 32 + r] = v;
                    }
                    n_final = a.world;
                }
                if (!ok) ctrl->abort = 1;
            }
            ++epoch;
            if (trace) { trace[2] = (long long)(dev::global_ns() & 0xffffffffffull); }
            dev::named_bar_sync(1, NTC);
            if (ctrl->abort) { final_error = kErrAbort; break; }      // uniform
            if (xatomic && cta == 0 && ncta > 1) {
                // CTA 0 has seen every CTA arrive at exchange e = epoch - 1, so every CTA is done with exchange e - 2: its slot
                // (e + 2) & 3 is cleared for exchange e + 2 (nobody adds to it before CTA 0 itself has arrived at e + 1)
                const uint32_t zs = (epoch + 1u) & 3u;
                if (ctid < a.gs_cap) dev::st_relaxed_f64(a.xs_sum + (size_t)zs * a.ll_gs_cap + ctid, 0.0);
                if (ctid == 0) dev::st_relaxed_u32(a.xs_cnt + (size_t)zs * 32, 0u);
            }
            // Hand the PREVIOUS group's stage back to the TMA producer only now: its refill burst (tens of KB per SM)
            // then overlaps the proximal update, when the SM's load path is idle, instead of delaying the polling loads.
            if (SMEM && pending_stage >= 0) {
                if (lane == 0) dev::mbar_arrive(&empty_bar[pending_stage]);
                pending_stage = -1;
            }
            if ((ncta > 1 || multi_gpu) && warp == 0) {                  // control warp: add the group / rank partials in fixed order
#pragma unroll 1
                for (int c = lane; c < gs; c += 32) {
                    double s = 0;
#pragma unroll 1
                    for (int g = 0; g < n_final; ++g) s += vals2[c * 32 + g];
                    gsum[c] = s;
                }
                __syncwarp();
            }
            AB_TICK(3);
            if (trace) { trace[3] = (long long)(dev::global_ns() & 0xffffffffffull); }

            // ---- proximal update (control warp), replicated bit-for-bit in every CTA
            if (warp == 0) {
                int changed = 0;
                const P pk = (P)m.pen;
                if (gs == 1) {                                           // solver_gaussian_pin_naive.hpp:75-108
                    const double g = gsum[0];
                    __syncwarp();
                    const P ak_old = p_aold[0];
                    const P A_kk = (P)rec[0], xm = (P)rec[1];
                    P gk = (P)g - xm * (P)ps.resid_sum * (P)a.intercept + ak_old * A_kk;
                    const P vv = fabs(gk) - l1 * pk;                     // update_coordinate, pin_base.hpp:181-195
                    P ak = (vv > P(0)) ? copysign(vv, gk) / (A_kk + l2 * pk) : P(0);
                    ak = (P)(T)ak;                                       // coefficients live in T
                    gk -= ak_old * A_kk;
                    if (ak != ak_old) {
                        const P del = ak - ak_old;
                        ps.cm = fmax(ps.cm, (double)(A_kk * del * del));
                        ps.rsq += (double)(del * (2 * gk - del * A_kk));
                        ps.resid_sum -= (double)(xm * del);
                        if (lane == 0) { my_beta[m.begin] = (T)ak; s_del[0] = (T)(-del); }
                        changed = 1;
                    }
                } else if (small_group) {                                // :109-164, one coefficient per lane
                    changed = prox_small_post<T, P>(proxctx, pre, rec, gs, m.begin, l1 * pk, l2 * pk, (P)a.newton_tol, a.newton_max_iters,
                                                    (P)a.dbeta_tol, a.intercept, ps, lane, prof ? &pt[7] : nullptr);
                } else {                                                 // :109-164, general group size
                    changed = prox_group<T, P>(proxctx, rec, gs, m.begin, l1 * pk, l2 * pk, (P)a.newton_tol, a.newton_max_iters,
                                               (P)a.dbeta_tol, a.intercept, ps, lane, prof ? &pt[7] : nullptr);
                }
                if (changed && kind == kSweepScreen) {                   // add_active_set (:294-304)
                    if (!my_active[ss]) {
                        if (ps.A >= a.max_active_size) ps.error = kErrMaxActive;
                        else {
                            __syncwarp();
                            if (lane == 0) { my_active[ss] = 1; a.active_set[ps.A] = ss; }
                            ++ps.A;
                        }
                    }
                }
                if (lane == 0) { ctrl->changed = changed; ctrl->error = ps.error; }
                __syncwarp();
            }
            ++n_updates; n_cols += gs;
            AB_TICK(4);
            if (trace) { trace[4] = (long long)(dev::global_ns() & 0xffffffffffull); }
            dev::named_bar_sync(1, NTC);
            const int changed_now = ctrl->changed;
            AB_TICK(5);

            // ---- residual update from the same X tile: r += X_g * del
            if (changed_now && K == 1) {
#pragma unroll 1
                for (int c0 = 0; c0 < gs; c0 += CB) {
                    T d[CB];
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) d[cc] = (c0 + cc < gs) ? s_del[c0 + cc] : T(0);
#pragma unroll 1
                    for (int v = ctid; v < rows / VN; v += NTC) {
                        T rv[VN];
                        vec_load<T>(rr + (size_t)v * VN, rv);
#pragma unroll
                        for (int cc = 0; cc < CB; ++cc) {
                            if (c0 + cc < gs) {
                                T xv[VN];
                                vec_load<T>(xs + (int64_t)(c0 + cc) * cs + (size_t)v * VN, xv);
#pragma unroll
                                for (int k = 0; k < VN; ++k) rv[k] += xv[k] * d[cc];
                            }
                        }
                        vec_store<T>(rr + (size_t)v * VN, rv);
                    }
                }
            }
            if (changed_now && K > 1) {                                      // r[i, class] += X[i, feature] * del[feature * K + class]
                const bool icpt = m.col < 0;
                const int kcls = icpt ? (-m.col - 1) : 0;
                if (!icpt && (m.col % K) == 0 && gs == K && (K % VN) == 0) {
                    T d[CB];
#pragma unroll
                    for (int cc = 0; cc < CB; ++cc) d[cc] = (cc < gs) ? s_del[cc] : T(0);
#pragma unroll 2
                    for (int i = ctid; i < rows; i += NTC) {
                        const T x = xs[i];
                        T* rrow = rr + (size_t)i * K;
#pragma unroll
                        for (int kv = 0; kv < CB / VN; ++kv) {
                            if (kv * VN < K) {
                                T rv[VN];
                                vec_load<T>(rrow + kv * VN, rv);
#pragma unroll
                                for (int k = 0; k < VN; ++k) rv[k] += x * d[kv * VN + k];
                                vec_store<T>(rrow + kv * VN, rv);
                            }
                        }
                    }
                } else
#pragma unroll 1
                for (int i = ctid; i < rows; i += NTC) {
                    T* rrow = rr + (size_t)i * K;
                    if (icpt) { rrow[kcls] += s_del[0]; continue; }
                    int f = 0, k = m.col % K;
                    T x = xs[i];
#pragma unroll 1
                    for (int c = 0; c < gs; ++c) {
                        rrow[k] += x * s_del[c];
                        if (++k == K && c + 1 < gs) { k = 0; ++f; x = xs[(int64_t)f * cs + i]; }
                    }
                }
            }
            const int err_now = ctrl->error;
            pending_stage = SMEM ? stage : -1;      // released after the NEXT exchange (see below)
            ++gitem;
            AB_TICK(6);
            if (trace) { trace[5] = (long long)(dev::global_ns() & 0xffffffffffull); }
            if (prof) ++pt[8];
            if (err_now) { final_error = err_now; break; }
        }
        if (SMEM && pending_stage >= 0) {           // sweep finished (or failed): hand the last stage back right away
            __syncwarp();
            if (lane == 0) dev::mbar_arrive(&empty_bar[pending_stage]);
            pending_stage = -1;
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
        for (int v = ctid; v < rows * K / VN; v += NTC) {
            T t[VN];
            vec_load<T>(sr + (size_t)v * VN, t); vec_store<T>(gr + (size_t)v * VN, t);
        }
    }
    if (cta == 0 && ctid == 0) {
        a.sc->rsq = ps.rsq; a.sc->resid_sum = ps.resid_sum; a.sc->active_set_size = ps.A;
        a.sc->iters = iters; a.sc->n_group_updates = n_updates; a.sc->n_col_updates = n_cols; a.sc->error = final_error;
        a.sc->newton_iters_max = ps.newton_iters_max;
        *a.epoch = epoch;
        if (prof) for (int k = 0; k < 10; ++k) a.stats[k] += pt[k];
    }
}

} // namespace ab
