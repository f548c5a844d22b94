// adelie_b200/csrc/solver.cuh -- host-side path driver.  Control flow of the reference's
// solve_core (CORE/solver/solver_base.hpp:435-687), Gaussian fit / update_screen_derived /
// update_solutions (CORE/solver/solver_gaussian_naive.hpp:53-434) and GLM IRLS fit
// (CORE/solver/solver_glm_naive.hpp:165-546); all O(n) and O(n*p) work runs in CUDA kernels on
// device-resident data, only O(p)/O(G) bookkeeping (screening, KKT, lambda path) runs on the host.
#pragma once
#include "common.cuh"
#include "matrix.cuh"
#include "glm.cuh"
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <numeric>
#include <unordered_set>

namespace ab {

// Host symmetric eigendecomposition (cyclic Jacobi, double).  Replaces
// Eigen::SelfAdjointEigenSolver (solver_gaussian_naive.hpp:113); eigenvalues ascending,
// eigenvectors in the columns of V (row-major V[r*q + c]).
inline void host_jacobi_eigh(std::vector<double>& A, int q, std::vector<double>& D, std::vector<double>& V) {
    V.assign((size_t)q * q, 0.0);
    for (int i = 0; i < q; ++i) V[(size_t)i * q + i] = 1;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0, diag = 0;
        for (int a = 0; a < q; ++a) for (int b = 0; b < q; ++b) { const double x = A[(size_t)a * q + b]; if (a == b) diag += x * x; else off += x * x; }
        if (off <= 1e-31 * (diag + off) || off == 0) break;     // off-diagonal mass below (eps)^2 of the total
        for (int pi = 0; pi < q - 1; ++pi) for (int qi = pi + 1; qi < q; ++qi) {
            const double apq = A[(size_t)pi * q + qi];
            if (apq == 0) continue;
            const double app = A[(size_t)pi * q + pi], aqq = A[(size_t)qi * q + qi];
            const double theta = (aqq - app) / (2 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
            const double c = 1 / std::sqrt(t * t + 1), s = t * c;
            for (int k = 0; k < q; ++k) {
                const double akp = A[(size_t)k * q + pi], akq = A[(size_t)k * q + qi];
                A[(size_t)k * q + pi] = c * akp - s * akq; A[(size_t)k * q + qi] = s * akp + c * akq;
            }
            for (int k = 0; k < q; ++k) {
                const double apk = A[(size_t)pi * q + k], aqk = A[(size_t)qi * q + k];
                A[(size_t)pi * q + k] = c * apk - s * aqk; A[(size_t)qi * q + k] = s * apk + c * aqk;
            }
            for (int k = 0; k < q; ++k) {
                const double vkp = V[(size_t)k * q + pi], vkq = V[(size_t)k * q + qi];
                V[(size_t)k * q + pi] = c * vkp - s * vkq; V[(size_t)k * q + qi] = s * vkp + c * vkq;
            }
        }
    }
    std::vector<int> ord(q);
    std::iota(ord.begin(), ord.end(), 0);
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return A[(size_t)a * q + a] < A[(size_t)b * q + b]; });
    std::vector<double> Vs((size_t)q * q);
    D.resize(q);
    for (int c = 0; c < q; ++c) {
        D[c] = A[(size_t)ord[c] * q + ord[c]];
        for (int k = 0; k < q; ++k) Vs[(size_t)k * q + c] = V[(size_t)k * q + ord[c]];
    }
    V.swap(Vs);
}

// Batched device version of host_jacobi_eigh: one warp per matrix (q <= 32), matrix and eigenvectors in shared memory, the same
// cyclic order of rotations and the same arithmetic as the host routine; lanes own a row / column element of each rotation.
// The GLM path recomputes the eigendecomposition of EVERY screen group in every IRLS iteration (solver_glm_naive.hpp:376-385):
// at config 3 that is 5000 10x10 problems x ~400 iterations, 24 s of single-threaded host time against ~0.1 s here.
struct EigItem { int64_t off; int64_t d_off; int32_t q; int32_t pad; };
constexpr int kEigWarps = 4;
__global__ void __launch_bounds__(kEigWarps * 32)
jacobi_eigh_kernel(const EigItem* __restrict__ items, int n_items, const double* __restrict__ A_in, double* __restrict__ V_out, double* __restrict__ D_out, int q_max)
{
    extern __shared__ double s_eig[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kEigWarps + warp;
    if (item >= n_items) return;
    const EigItem it = items[item];
    const int q = it.q;
    double* As = s_eig + (size_t)warp * (2 * q_max * q_max + 32);
    double* Vs = As + q_max * q_max;
    int* ord = reinterpret_cast<int*>(Vs + q_max * q_max);
    for (int e = lane; e < q * q; e += 32) { As[e] = A_in[it.off + e]; Vs[e] = ((e / q) == (e % q)) ? 1.0 : 0.0; }
    __syncwarp();
    const int k = lane;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0, diag = 0;
        if (k < q) for (int b = 0; b < q; ++b) { const double x = As[k * q + b]; if (b == k) diag += x * x; else off += x * x; }
        for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); diag += __shfl_xor_sync(0xffffffffu, diag, o); }
        if (off <= 1e-31 * (diag + off) || off == 0) break;
        for (int pi = 0; pi < q - 1; ++pi) for (int qi = pi + 1; qi < q; ++qi) {
            const double apq = As[pi * q + qi];
            if (apq == 0) continue;                                   // warp uniform
            const double app = As[pi * q + pi], aqq = As[qi * q + qi];
            const double theta = (aqq - app) / (2 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
            const double c = 1 / sqrt(t * t + 1), sn = t * c;
            __syncwarp();
            if (k < q) {
                const double akp = As[k * q + pi], akq = As[k * q + qi];
                As[k * q + pi] = c * akp - sn * akq; As[k * q + qi] = sn * akp + c * akq;
            }
            __syncwarp();
            if (k < q) {
                const double apk = As[pi * q + k], aqk = As[qi * q + k];
                As[pi * q + k] = c * apk - sn * aqk; As[qi * q + k] = sn * apk + c * aqk;
                const double vkp = Vs[k * q + pi], vkq = Vs[k * q + qi];
                Vs[k * q + pi] = c * vkp - sn * vkq; Vs[k * q + qi] = sn * vkp + c * vkq;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    if (lane == 0) {                                                  // ascending eigenvalues (stable insertion sort, q <= 32)
        for (int a = 0; a < q; ++a) ord[a] = a;
        for (int a = 1; a < q; ++a) {
            const int o = ord[a]; const double d = As[o * q + o];
            int b = a - 1;
            while (b >= 0 && As[ord[b] * q + ord[b]] > d) { ord[b + 1] = ord[b]; --b; }
            ord[b + 1] = o;
        }
    }
    __syncwarp();
    if (k < q) D_out[it.d_off + k] = As[ord[k] * q + ord[k]];
    for (int e = lane; e < q * q; e += 32) V_out[it.off + e] = Vs[(e / q) * q + ord[e % q]];
}

// search_pivot (CORE/optimization/search_pivot.hpp:7-62)
template <class T>
inline int search_pivot(const std::vector<T>& x, const std::vector<T>& y, std::vector<T>& mses) {
    const int64_t n = (int64_t)x.size();
    if (n <= 0) return -1;
    mses[0] = std::numeric_limits<T>::infinity();
    if (n == 1) return 0;
    T y_mean = 0;
    for (auto v : y) y_mean += v;
    y_mean /= n;
    T x_sum = x[0], xsq_sum = x[0] * x[0], y_sum = y[0], yx_sum = y[0] * x[0], min_mse = mses[0];
    int argmin = 0;
    for (int64_t i = 1; i < n; ++i) {
        x_sum += x[i]; xsq_sum += x[i] * x[i]; y_sum += y[i]; yx_sum += y[i] * x[i];
        const T t_bar = ((i + 1) * x[i] - x_sum) / n;
        const T var_t = (i + 1) * x[i] * x[i] - 2 * x[i] * x_sum + xsq_sum - n * t_bar * t_bar;
        const T cov_ty = x[i] * (y_sum - (i + 1) * y_mean) - (yx_sum - y_mean * x_sum);
        const T b1 = cov_ty / var_t;
        mses[i] = -b1 * b1 * var_t;
        if (mses[i] < min_mse) { argmin = (int)i; min_mse = mses[i]; }
    }
    return argmin;
}

struct SparseRow { std::vector<int64_t> idx; std::vector<double> val; };

// Host wall-clock accounting of the path driver (exposed as state scalars "t_<label>")
struct HostTimers {
    std::vector<std::pair<std::string, double>> acc;
    HostTimers() { acc.reserve(64); }          // scopes hold references into acc: it must never reallocate
    double& slot(const char* name) {
        for (auto& kv : acc) if (kv.first == name) return kv.second;
        acc.emplace_back(name, 0.0);
        return acc.back().second;
    }
    struct Scope {
        double& dst; double t0;
        Scope(double& d) : dst(d), t0(now_s()) {}
        ~Scope() { dst += now_s() - t0; }
    };
};
#define AB_TIME(timers, name) ::ab::HostTimers::Scope _ab_scope_##__LINE__((timers).slot(name))

// Output of one pin solve (the slice of StateGaussianPinNaive the path driver reads back)
struct PinResult { SparseRow beta; double intercept = 0, rsq = 0; double screen_time = 0, active_time = 0; long long iters = 0; };

template <class T>
struct PathState {
    using idx_t = int64_t;
    // ---------------- static (state_base.hpp:59-95, state_gaussian_naive.hpp, state_glm_naive.hpp)
    DenseMatrix<T>* X = nullptr;
    idx_t n = 0, p = 0, G = 0;
    // multi-response (state_multigaussian_naive.hpp:33-142, PY/solver.py:705-720): the solver sees the n*K x (n_int + pX*K) matrix
    // [kron(1, I_K) | kron(X, I_K)]; only X is stored, the kron/concatenate structure is a layout rule of the kernels.
    int K = 1; idx_t n_int = 0;
    std::vector<idx_t> groups, group_sizes;
    T alpha = 1; std::vector<T> penalty;
    bool is_glm = false;
    Glm<T>* glm = nullptr;
    T min_ratio = 1e-2; size_t lmda_path_size = 100, max_screen_size = 0, max_active_size = 0;
    T pivot_subset_ratio = 0.1; size_t pivot_subset_min = 1; T pivot_slack_ratio = 1.25; int screen_rule = 1;
    size_t max_iters = 100000; T tol = 1e-7, adev_tol = 0.9, ddev_tol = 0, newton_tol = 1e-12; size_t newton_max_iters = 1000;
    bool early_exit = true, setup_lmda_max = true, setup_lmda_path = true, intercept = true;
    size_t n_threads = 1;
    size_t irls_max_iters = 10000; T irls_tol = 1e-7; bool setup_loss_null = true;
    // ---------------- dynamic
    T lmda_max = -1; std::vector<T> lmda_path;
    std::unordered_set<idx_t> screen_hashset;
    std::vector<uint8_t> in_screen;            // flat copy of screen_hashset (G,): the O(G) loops per lambda test membership here
    std::vector<idx_t> screen_set, screen_begins;
    std::vector<T> screen_beta; std::vector<int8_t> screen_is_active;
    size_t active_set_size = 0; std::vector<idx_t> active_set;
    T lmda = std::numeric_limits<T>::infinity();
    std::vector<T> grad, abs_grad;
    // gaussian
    std::vector<T> X_means; T y_mean = 0, y_var = 0, loss_null = 0, loss_full = 0, resid_sum = 0, rsq = 0;
    std::vector<T> screen_X_means, screen_vars; std::vector<std::vector<T>> screen_transforms;
    // glm
    T beta0 = 0;
    // ---------------- outputs
    std::vector<SparseRow> betas; std::vector<T> intercepts, devs, lmdas;
    std::vector<double> benchmark_screen, benchmark_fit_screen, benchmark_fit_active, benchmark_kkt, benchmark_invariance;
    std::vector<int> n_valid_solutions, active_sizes, screen_sizes;
    long long n_sweeps = 0, n_group_updates = 0, n_col_updates = 0, n_irls = 0, n_pin_solves = 0, n_kernel_launches = 0;
    double sweep_bytes = 0;          // algorithmic HBM bytes of all sweeps: s*n*sum gs + 3*s*n per sweep (estimated)
    double time_sweep_kernel = 0;    // CUDA-event time spent inside the fused kernel (s)
    HostTimers timers;
    // ---------------- device
    DevBuf<T> d_weights, d_weights_sqrt, d_resid, d_resid_prev, d_X_means, d_grad;
    DevBuf<T> d_offsets, d_eta, d_eta_prev, d_glm_resid_prev, d_hess, d_irls_w, d_irls_wsqrt, d_irls_y, d_irls_resid;
    DevBuf<T> d_glm_X_means;     // (p,) irls-weighted means on screen columns
    DevBuf<GroupMeta> d_meta; DevBuf<T> d_grec, d_screen_beta, d_screen_beta_rot; DevBuf<int8_t> d_is_active; DevBuf<int32_t> d_active_set;
    DevBuf<PinScalars> d_sc; DevBuf<CovItem> d_cov_items; DevBuf<double> d_cov_out; DevBuf<int32_t> d_cols; DevBuf<T> d_tmp;
    DevBuf<double> d_scal;       // small scalar scratch (device-side sub_scale etc.)
    std::vector<GroupMeta> h_meta; std::vector<T> h_grec; size_t grec_uploaded = 0, meta_uploaded = 0;
    // Gram panels of the batched look-ahead kernel (sweep_batched.cuh): one list per sweep order (screen positions, active list)
    struct PanelList { DevBuf<T> Q; std::vector<int32_t> entries; int B = 0, Ccap = 0; };
    PanelList pl_screen, pl_active;
    DevBuf<PairItem> d_pair_items; DevBuf<PanelItem> d_panel_items;
    long long n_panels_built = 0, n_batched_launches = 0;
    std::vector<double> launch_cols, launch_sweeps, launch_ms;      // per sweep-kernel launch: column visits, sweeps, CUDA-event time
    int gs_max_screen = 1, rec_max_screen = 4, feat_max_screen = 1;
    PinnedBuf<PinScalars> h_sc;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::function<bool()> exit_cond;          // user early-exit callback
    std::function<void()> check_interrupt;    // PyErr_CheckSignals hook

    ~PathState() { if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); }

    // ------------------------------------------------------------------ validation + init
    // state_base.ipp:10-116 and state_gaussian_naive.ipp:9-28
    void validate_and_init() {
        if ((idx_t)group_sizes.size() != G) throw core_error("group_sizes must be (G,) where groups is (G,).");
        if ((idx_t)penalty.size() != G) throw core_error("penalty must be (G,) where groups is (G,).");
        if (alpha < 0 || alpha > 1) throw core_error("alpha must be in [0,1].");
        if (tol < 0) throw core_error("tol must be >= 0.");
        if (adev_tol < 0 || adev_tol > 1) throw core_error("adev_tol must be in [0,1].");
        if (ddev_tol < 0 || ddev_tol > 1) throw core_error("ddev_tol must be in [0,1].");
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
        if ((idx_t)grad.size() != p) throw core_error("grad must be (p,) where X is (n, p).");
        abs_grad.assign(G, 0);
        AB_CUDA(cudaEventCreate(&ev0)); AB_CUDA(cudaEventCreate(&ev1));
        d_sc.alloc(1); h_sc.alloc(1); d_scal.alloc(8);
        d_active_set.alloc(G);
        update_screen_derived_base();
        update_abs_grad(lmda);
        if (!is_glm) {
            loss_null = T(-0.5) * y_mean * y_mean;                  // state_gaussian_naive.hpp:143-144
            loss_full = T(-0.5) * y_var + loss_null;
            update_screen_derived_gaussian();
        }
    }

    // ------------------------------------------------------------------ solver_base.hpp:20-110 (constraints == nullptr)
    void update_abs_grad(T lmda_) {
        AB_TIME(timers, "abs_grad");
        for (size_t ss = 0; ss < screen_set.size(); ++ss) {
            const idx_t i = screen_set[ss], b = screen_begins[ss], k = groups[i], sz = group_sizes[i];
            const T regul = ((1 - alpha) * lmda_) * penalty[i];
            T a = 0;
            for (idx_t c = 0; c < sz; ++c) { const T e = grad[k + c] - regul * screen_beta[b + c]; a += e * e; }
            abs_grad[i] = std::sqrt(a);
        }
        const bool flat = (idx_t)in_screen.size() == G;
        for (idx_t i = 0; i < G; ++i) {
            if (flat ? in_screen[i] : (uint8_t)screen_hashset.count(i)) continue;
            const idx_t k = groups[i], sz = group_sizes[i];
            T a = 0;
            for (idx_t c = 0; c < sz; ++c) a += grad[k + c] * grad[k + c];
            abs_grad[i] = std::sqrt(a);
        }
    }

    // solver_base.hpp:120-153
    void update_screen_derived_base() {
        const size_t old = screen_begins.size();
        if ((idx_t)in_screen.size() != G) { in_screen.assign(G, 0); for (idx_t g : screen_set) if (screen_hashset.count(g)) in_screen[g] = 1; }
        for (size_t i = old; i < screen_set.size(); ++i) { screen_hashset.insert(screen_set[i]); in_screen[screen_set[i]] = 1; }
        size_t vs = (old == 0) ? 0 : (screen_begins.back() + group_sizes[screen_set[old - 1]]);
        for (size_t i = old; i < screen_set.size(); ++i) { screen_begins.push_back(vs); vs += group_sizes[screen_set[i]]; }
        screen_beta.resize(vs, 0);
        screen_is_active.resize(screen_set.size(), 0);
    }

    static int batch_ext_len(int gs) { const int gsp = (gs + 3) / 4 * 4; return 3 * gsp + 2 * gs * gsp; }

    // Computes (A, V, xm) for screen positions [begin, end) from the weighted Gram of each group
    // (solver_gaussian_naive.hpp:53-125): device batched Gram -> host Jacobi -> packed records.
    // `w` device weights (not sqrt), `xmeans_host(col)` gives the weighted column mean.
    // fuse_means: the weighted column means come out of the Gram pass itself (`xmean` is then not called); only when
    // screen_means_fusable() says the Gram kernel can do it.
    bool screen_means_fusable(size_t begin, size_t end) const {
        if (!Configs::glm_fuse_means || K != 1) return false;
        int gm_ = 0;
        for (size_t i = begin; i < end; ++i) gm_ = std::max(gm_, (int)group_sizes[screen_set[i]]);
        return X->cov_can_fuse_means(K, gm_, false);
    }
    template <class MeanF>
    void compute_screen_records(size_t begin, size_t end, const T* d_w, MeanF xmean,
                                std::vector<T>& sXm, std::vector<T>& sv, std::vector<std::vector<T>>& st,
                                std::vector<GroupMeta>& meta, std::vector<T>& grec, bool fuse_means = false)
    {
        const size_t S = screen_set.size();
        const size_t vs = S ? (screen_begins.back() + group_sizes[screen_set.back()]) : 0;
        sXm.resize(vs); st.resize(S); sv.resize(vs, 0);
        if (begin >= end) return;
        AB_TIME(timers, "screen_records");
        // one Gram item per (group, class present in the group): kron(X, I_K)^T W kron(X, I_K) is block diagonal over classes
        struct GInfo { int f0, k0, nfeat; bool icpt; size_t item0; };
        std::vector<GInfo> ginfo;
        std::vector<CovItem> items; int64_t c_total = 0;
        for (size_t i = begin; i < end; ++i) {
            const idx_t g = screen_set[i]; const int gs = (int)group_sizes[g];
            GInfo gi{}; gi.item0 = items.size();
            if (K == 1) {
                gi.f0 = (int)X->phys_col(groups[g], gs); gi.k0 = 0; gi.nfeat = gs; gi.icpt = false;      // physical column (SNP: slot of the decoded-column cache)
                items.push_back(CovItem{(int32_t)gi.f0, gs, c_total, 0, fuse_means ? (int32_t)(screen_begins[i] + 1) : 0});
                c_total += (int64_t)gs * gs;
            } else if (groups[g] < n_int) {
                if (gs != 1) throw core_error("multi-response intercept columns must be groups of size 1.");
                gi.f0 = 0; gi.k0 = (int)groups[g]; gi.nfeat = 1; gi.icpt = true;
                items.push_back(CovItem{-1, 1, c_total, (int32_t)groups[g], 0});
                c_total += 1;
            } else {
                const idx_t c0 = groups[g] - n_int;
                gi.k0 = (int)(c0 % K); gi.nfeat = (gi.k0 + gs + K - 1) / K; gi.icpt = false;
                gi.f0 = (int)X->phys_col(c0 / K, gi.nfeat);
                for (int l = 0; l < K; ++l) {      // item for class l even when absent keeps the indexing simple (absent: skipped below)
                    items.push_back(CovItem{(int32_t)gi.f0, gi.nfeat, c_total, l, 0});
                    c_total += (int64_t)gi.nfeat * gi.nfeat;
                }
            }
            ginfo.push_back(gi);
        }
        double t_cov0 = now_s();
        d_cov_items.reserve_keep(items.size()); d_cov_out.reserve_keep(c_total);
        d_cov_items.upload(items.data(), items.size());
        int cov_gs_max = 0; for (const CovItem& ci : items) cov_gs_max = std::max(cov_gs_max, (int)ci.gs);
        if (fuse_means) d_cov_means.reserve_keep(vs + 1);
        X->d_cov(d_cov_items.p, (int)items.size(), c_total, d_w, false, d_cov_out.p, K, cov_gs_max, fuse_means ? d_cov_means.p : nullptr, (int64_t)vs);
        DistContext::get().allreduce<double>(d_cov_out.p, c_total);            // row-sharded: sum the local Gram blocks over ranks
        std::vector<double>& C = scr_C;                 // persistent scratch: fresh multi-megabyte vectors per IRLS iteration cost more in page faults than the work
        C.resize(c_total);
        d_cov_out.download(C.data(), c_total);
        if (fuse_means) {
            DistContext::get().allreduce<double>(d_cov_means.p, (int64_t)vs);
            scr_M.resize(vs);
            d_cov_means.download(scr_M.data(), vs);
        }
        AB_CUDA(cudaStreamSynchronize(0));
        timers.slot("cov_device") += now_s() - t_cov0;
        n_kernel_launches += 2;
        meta.resize(S);
        // ---- phase 1: centred Gram of every group, packed back to back (eig_in), offsets per group
        std::vector<double>& eig_in = scr_eig_in; std::vector<double>& eig_V = scr_eig_V; std::vector<double>& eig_D = scr_eig_D;
        std::vector<EigItem>& eig_items = scr_eig_items; std::vector<int64_t>& eig_slot = scr_eig_slot;
        eig_items.clear(); eig_slot.assign(end - begin, -1);
        double t_ph = now_s();
        {
            int64_t off = 0, d_off = 0;
            for (size_t i = begin; i < end; ++i) {
                const int gs = (int)group_sizes[screen_set[i]];
                if (gs > 1) { eig_slot[i - begin] = (int64_t)eig_items.size(); eig_items.push_back(EigItem{off, d_off, gs, 0}); off += (int64_t)gs * gs; d_off += gs; }
            }
            eig_in.resize(off); eig_V.resize(off); eig_D.resize(d_off);
        }
        // (independent per group: every group writes its own slice of sXm / eig_in; on the GLM path this runs for ALL screen groups in every
        // IRLS iteration -- 1.3 s per config-3 shard path when it was a sequential loop)
        const long long p1_begin = (long long)begin, p1_end = (long long)end;
#pragma omp parallel for schedule(static) if (p1_end - p1_begin >= 256)
        for (long long ii = p1_begin; ii < p1_end; ++ii) {
            const size_t i = (size_t)ii;
            const idx_t g = screen_set[i]; const int gs = (int)group_sizes[g]; const idx_t sb = screen_begins[i];
            const GInfo& gi = ginfo[i - begin];
            if (fuse_means) { for (int c = 0; c < gs; ++c) sXm[sb + c] = (T)scr_M[sb + c]; }
            else for (int c = 0; c < gs; ++c) sXm[sb + c] = xmean(groups[g] + c);
            if (gs == 1) continue;
            double* dst = eig_in.data() + eig_items[eig_slot[i - begin]].off;
            if (K == 1 || gi.icpt) {
                const CovItem& it = items[gi.item0];
                std::copy(C.begin() + it.out_off, C.begin() + it.out_off + (size_t)gs * gs, dst);
            } else {
                std::fill(dst, dst + (size_t)gs * gs, 0.0);
                for (int a = 0; a < gs; ++a) for (int b = 0; b < gs; ++b) {
                    const int fa = (gi.k0 + a) / K, ka = (gi.k0 + a) % K, fb = (gi.k0 + b) / K, kb = (gi.k0 + b) % K;
                    if (ka == kb) dst[(size_t)a * gs + b] = C[items[gi.item0 + ka].out_off + (size_t)fa * gi.nfeat + fb];
                }
            }
            if (intercept)
                for (int a = 0; a < gs; ++a) for (int b = 0; b < gs; ++b) dst[(size_t)a * gs + b] -= (double)sXm[sb + a] * (double)sXm[sb + b];
        }
        timers.slot("rec_phase1") += now_s() - t_ph;
        // ---- phase 2: eigendecompositions -- one batched launch on the device (groups of <= 32 columns), host Jacobi otherwise
        bool eig_on_device = false;
        if (Configs::device_eigh && eig_items.size() >= 16) {
            int q_max = 0; for (const EigItem& e : eig_items) q_max = std::max(q_max, (int)e.q);
            if (q_max <= 32) {
                AB_TIME(timers, "eigh_device");
                DevBuf<EigItem> d_items(eig_items.size()); DevBuf<double> d_in(eig_in.size()), d_V(eig_in.size()), d_D(eig_D.size());
                d_items.upload(eig_items.data(), eig_items.size()); d_in.upload(eig_in.data(), eig_in.size());
                const size_t smem = sizeof(double) * (size_t)kEigWarps * (2 * q_max * q_max + 32);
                AB_CUDA(cudaFuncSetAttribute(jacobi_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                jacobi_eigh_kernel<<<(unsigned)((eig_items.size() + kEigWarps - 1) / kEigWarps), kEigWarps * 32, smem, 0>>>(
                    d_items.p, (int)eig_items.size(), d_in.p, d_V.p, d_D.p, q_max);
                AB_CUDA(cudaGetLastError());
                d_V.download(eig_V.data(), eig_V.size()); d_D.download(eig_D.data(), eig_D.size());
                AB_CUDA(cudaStreamSynchronize(0));
                eig_on_device = true; ++n_kernel_launches;
            }
        }
        // ---- phase 3: records.  Offsets first (sequential), then every group fills its own slice of grec / sv / st / meta: the loop is
        // embarrassingly parallel over groups and runs on the host cores (OpenMP) -- on the GLM path it is executed for ALL screen
        // groups in every IRLS iteration.
        t_ph = now_s();
        std::vector<size_t> rec_off_v(end - begin), ext_off_v(end - begin);
        {
            size_t sz = grec.size();
            for (size_t i = begin; i < end; ++i) {
                const int gs = (int)group_sizes[screen_set[i]];
                const size_t off = (sz + 3) / 4 * 4;
                const int rec_pad = (gs * (gs + 3) + 3) / 4 * 4;
                rec_off_v[i - begin] = off; ext_off_v[i - begin] = off + rec_pad;
                sz = off + rec_pad + batch_ext_len(gs);
                gs_max_screen = std::max(gs_max_screen, gs); rec_max_screen = std::max(rec_max_screen, rec_pad);
                feat_max_screen = std::max(feat_max_screen, ginfo[i - begin].nfeat);
            }
            grec.resize(sz, T(0));
        }
        const long long i_begin = (long long)begin, i_end = (long long)end;
#pragma omp parallel for schedule(static) if (i_end - i_begin >= 256)
        for (long long ii = i_begin; ii < i_end; ++ii) {
            const size_t i = (size_t)ii;
            const idx_t g = screen_set[i]; const int gs = (int)group_sizes[g]; const idx_t sb = screen_begins[i];
            const GInfo& gi = ginfo[i - begin];
            const size_t off = rec_off_v[i - begin], e0 = ext_off_v[i - begin];
            const int rec_pad = (gs * (gs + 3) + 3) / 4 * 4;
            std::vector<T>& Vt = st[i];
            if (gs == 1) {
                double c0 = (K == 1 || gi.icpt) ? C[items[gi.item0].out_off] : C[items[gi.item0 + gi.k0].out_off];
                if (intercept) c0 -= (double)sXm[sb] * (double)sXm[sb];
                Vt.assign(1, T(1));
                sv[sb] = std::max<T>((T)c0, 0);
            } else {
                const EigItem& e = eig_items[eig_slot[i - begin]];
                Vt.resize((size_t)gs * gs);
                if (eig_on_device) {
                    for (size_t k = 0; k < (size_t)gs * gs; ++k) Vt[k] = (T)eig_V[e.off + k];
                    for (int c = 0; c < gs; ++c) { const T d = (T)eig_D[e.d_off + c]; sv[sb + c] = d * T(d >= 0); }      // :122
                } else {
                    std::vector<double> Cg_l(eig_in.begin() + e.off, eig_in.begin() + e.off + (size_t)gs * gs), D_l, V_l;
                    host_jacobi_eigh(Cg_l, gs, D_l, V_l);
                    for (size_t k = 0; k < (size_t)gs * gs; ++k) Vt[k] = (T)V_l[k];
                    for (int c = 0; c < gs; ++c) { const T d = (T)D_l[c]; sv[sb + c] = d * T(d >= 0); }
                }
            }
            // packed record [A | xm | V^T xm | V], 16-byte aligned
            for (int c = 0; c < gs; ++c) {
                grec[off + c] = sv[sb + c]; grec[off + gs + c] = sXm[sb + c];
                double xmt = 0;
                for (int r = 0; r < gs; ++r) xmt += (double)sXm[sb + r] * (double)Vt[(size_t)r * gs + c];
                grec[off + 2 * gs + c] = (T)xmt;
            }
            for (int k = 0; k < gs * gs; ++k) grec[off + 3 * gs + k] = Vt[k];
            for (size_t k = off + (size_t)gs * (gs + 3); k < off + rec_pad; ++k) grec[k] = T(0);
            // extension for the batched kernel (sweep_batched.cuh), right behind the base record: vector-load friendly copies
            //   [A(gsp) | xm(gsp) | V^T xm(gsp) | V rows padded to gsp | V^T rows padded to gsp],  gsp = gs rounded up to 4
            {
                const int gsp = (gs + 3) / 4 * 4;
                for (size_t k = e0; k < e0 + (size_t)batch_ext_len(gs); ++k) grec[k] = T(0);
                for (int c = 0; c < gs; ++c) {
                    grec[e0 + c] = grec[off + c]; grec[e0 + gsp + c] = grec[off + gs + c]; grec[e0 + 2 * gsp + c] = grec[off + 2 * gs + c];
                    for (int r = 0; r < gs; ++r) {
                        grec[e0 + 3 * gsp + (size_t)r * gsp + c] = Vt[(size_t)r * gs + c];                     // V[r][c]
                        grec[e0 + 3 * gsp + (size_t)gs * gsp + (size_t)c * gsp + r] = Vt[(size_t)r * gs + c];  // V^T[c][r]
                    }
                }
            }
            GroupMeta m{};
            m.col = (K == 1) ? (int32_t)gi.f0 : (gi.icpt ? -(int32_t)(groups[g] + 1) : (int32_t)(gi.f0 * K + gi.k0)); m.gs = gs; m.begin = (int32_t)sb; m.rec_elems = rec_pad; m.rec_off = (int64_t)off;
            m.pen = (double)penalty[g];
            meta[i] = m;
        }
        timers.slot("rec_phase3") += now_s() - t_ph;
    }

    // solver_gaussian_naive.hpp:134-176 (new screen positions only; weights are static)
    void update_screen_derived_gaussian() {
        const size_t old = screen_transforms.size();
        update_screen_derived_base();
        compute_screen_records(old, screen_set.size(), d_weights.p, [&](idx_t c) { return X_means[c]; },
                               screen_X_means, screen_vars, screen_transforms, h_meta, h_grec);
    }

    void upload_screen_tables(const std::vector<GroupMeta>& meta, const std::vector<T>& grec, bool full) {
        d_meta.reserve_keep(meta.size()); d_grec.reserve_keep(grec.size() + 4);
        const size_t m0 = full ? 0 : meta_uploaded, g0 = full ? 0 : grec_uploaded;
        if (meta.size() > m0) d_meta.upload(meta.data() + m0, meta.size() - m0, m0);
        if (grec.size() > g0) d_grec.upload(grec.data() + g0, grec.size() - g0, g0);
        if (!full) { meta_uploaded = meta.size(); grec_uploaded = grec.size(); }
    }

    // ------------------------------------------------------------------ Gram panels (sweep_batched.cuh)
    // Panel b of a list holds, for every source column of batch b, X_src^T W X_t against the later groups of batch b
    // (targets [0, Ccap)) and all groups of batch b+1 (targets [Ccap, 2 Ccap)).  Lists only ever grow at the tail, so only the
    // panels from the batch before the first new position onwards are (re)computed.
    void ensure_panels(PanelList& pl, const std::vector<int32_t>& entries, int B, int Ccap, const T* d_w, bool whole_panels = false) {
        if (pl.B != B || pl.Ccap != Ccap) { pl.entries.clear(); pl.B = B; pl.Ccap = Ccap; }
        size_t prev = 0;
        while (prev < pl.entries.size() && prev < entries.size() && pl.entries[prev] == entries[prev]) ++prev;
        if (prev == entries.size() && prev == pl.entries.size()) return;
        AB_TIME(timers, "panels");
        const size_t N = entries.size();
        const int ldq = 2 * Ccap;
        const size_t pstride = (size_t)Ccap * ldq;
        const size_t nb = (N + B - 1) / B;
        const size_t b0 = (prev / B > 0) ? prev / B - 1 : 0;
        pl.Q.reserve_keep(nb * pstride + 64);          // (new tail zeroed; old panels keep their blocks)
        std::vector<PairItem> items; int64_t total = 0;
        auto grp = [&](size_t pos, int& col, int& gs) { const idx_t g = screen_set[entries[pos]]; gs = (int)group_sizes[g]; col = (int)X->phys_col(groups[g], gs); };
        // only the blocks with a NEW group on either side are computed: the others are already in place
        for (size_t b = b0; b < nb; ++b) {
            const size_t p0 = b * B, p1 = std::min(N, p0 + B), p2 = std::min(N, p1 + B);
            int off_k = 0;
            for (size_t k = p0; k < p1; ++k) {
                int ck, gk; grp(k, ck, gk);
                int off_t = off_k + gk;
                for (size_t k2 = k + 1; k2 < p1; ++k2) {
                    int c2, g2; grp(k2, c2, g2);
                    if (k2 >= prev) {
                        items.push_back(PairItem{ck, gk, c2, g2, (int64_t)(b * pstride + (size_t)off_k * ldq + off_t), total});
                        total += (int64_t)gk * g2;
                    }
                    off_t += g2;
                }
                off_t = 0;
                for (size_t k2 = p1; k2 < p2; ++k2) {
                    int c2, g2; grp(k2, c2, g2);
                    if (k2 >= prev) {
                        items.push_back(PairItem{ck, gk, c2, g2, (int64_t)(b * pstride + (size_t)off_k * ldq + Ccap + off_t), total});
                        total += (int64_t)gk * g2;
                    }
                    off_t += g2;
                }
                off_k += gk;
            }
        }
        if constexpr (std::is_same<T, float>::value) {
            // Whole panels pay off when many blocks are new at once (a block costs ~5 us in pair_gram_kernel, a whole panel ~90 us
            // + ~40 us of fixed cost): the usual incremental call (a few new groups at the tail) stays on the per-block kernel.
            // IRLS (whole_panels): the weights changed, every panel of the list is rebuilt -- one tensor-core pass per panel.
            // Gaussian path (incremental): whole tail panels pay off when many blocks are new at once.  Costs in us at n rows (measured at
            // n = 200k: ~5 per block in pair_gram_kernel; a whole panel ~35 on the tensor cores, ~90 on the CUDA cores, + launch overhead).
            const double sc_n = (double)X->n_pad() / 200000.0;
            const double t_pairs = (double)items.size() * 5.0 * sc_n;
            const double t_panels = (double)(nb - b0) * ((Configs::panel_tc ? 35.0 : 90.0) * sc_n + 10.0) + 40.0;
            if (!X->sparse && (whole_panels || ((Configs::panel_gemm || Configs::panel_tc) && t_pairs > t_panels))) {
                // whole panels b0 .. nb-1 in one pass each (panel_gram_kernel): every window column is read once per panel
                std::vector<PanelItem> pitems;
                for (size_t b = b0; b < nb; ++b) {
                    const size_t p0 = b * B, p1 = std::min(N, p0 + B), p2 = std::min(N, p1 + B);
                    PanelItem pi{}; pi.q_off = (int64_t)(b * pstride); pi.ncol = 0;
                    for (size_t k = p0; k < p2; ++k) {
                        int ck, gk; grp(k, ck, gk);
                        for (int c = 0; c < gk; ++c) pi.cols[pi.ncol++] = ck + c;
                        if (k + 1 == p1) pi.n_src = pi.ncol;
                    }
                    if (p1 == p0) continue;
                    pitems.push_back(pi);
                }
                if (!pitems.empty()) {
                    AB_CUDA(cudaStreamSynchronize(0));
                    d_panel_items.reserve_keep(pitems.size());
                    d_panel_items.upload(pitems.data(), pitems.size());
                    X->d_panel_gram(d_panel_items.p, (int)pitems.size(), d_w, pl.Q.p, ldq, Ccap);
                    AB_CUDA(cudaStreamSynchronize(0));          // `pitems` (pageable host memory) must outlive the upload
                    X->check_tc_error();
                    n_kernel_launches += 3;
                }
                n_panels_built += (long long)(nb - b0);
                pl.entries = entries;
                return;
            }
        }
        if (!items.empty()) {
            { AB_TIME(timers, "panels_presync"); AB_CUDA(cudaStreamSynchronize(0)); }
            { AB_TIME(timers, "panels_launch");
              d_pair_items.reserve_keep(items.size());
              d_pair_items.upload(items.data(), items.size());
              X->d_pair_gram(d_pair_items.p, (int)items.size(), total, d_w, pl.Q.p, ldq, gs_max_screen); }
            { AB_TIME(timers, "panels_sync"); AB_CUDA(cudaStreamSynchronize(0)); }          // `items` (pageable host memory) must outlive the upload
            n_kernel_launches += 2;
        }
        n_panels_built += (long long)(nb - b0);
        pl.entries = entries;
    }

    // ------------------------------------------------------------------ the fused pin solve
    // Runs pin::naive::solve (solver_gaussian_pin_naive.hpp:223-401) for one lambda on the device.
    // static_weights: the Gaussian path (weights fixed, panels extended incrementally).  Otherwise (IRLS) the batched kernel is used when
    // Configs::glm_batched allows it, with every Gram panel rebuilt for the weights of this IRLS iteration; `transforms` are the screen
    // groups' eigenvector matrices that belong to `d_w` (the state's own on the Gaussian path, the IRLS iteration's on the GLM path).
    PinResult run_pin(T* d_r, const T* d_w, T lmda_, T tol_pin, T y_mean_, T& rsq_io, T& resid_sum_io, bool static_weights = false,
                      const std::vector<std::vector<T>>* transforms = nullptr) {
        AB_TIME(timers, "run_pin");
        if (!transforms) transforms = &screen_transforms;
        const bool glm_batched = !static_weights && K == 1 && !X->sparse &&
                                 (Configs::glm_batched >= 2 || (Configs::glm_batched == 1 && std::is_same<T, float>::value));
        const bool want_batched = (static_weights || glm_batched) && K == 1 && !X->sparse;
        const size_t S = screen_set.size();
        d_screen_beta.reserve_keep(screen_beta.size() + 4); d_is_active.reserve_keep(S + 4);
        d_screen_beta.upload(screen_beta.data(), screen_beta.size());
        d_is_active.upload(screen_is_active.data(), S);
        // coefficients in every group's eigenbasis (a V): the batched kernel keeps them up to date instead of rotating per visit
        std::vector<T> beta_rot(screen_beta.size(), T(0));
        if (want_batched) {
            for (size_t i = 0; i < S && i < transforms->size(); ++i) {
                const int gs = (int)group_sizes[screen_set[i]]; const idx_t sb = screen_begins[i];
                const std::vector<T>& V = (*transforms)[i];
                for (int c = 0; c < gs; ++c) {
                    double acc = 0;
                    for (int r = 0; r < gs; ++r) acc += (double)screen_beta[sb + r] * (double)V[(size_t)r * gs + c];
                    beta_rot[sb + c] = (T)acc;
                }
            }
            d_screen_beta_rot.reserve_keep(screen_beta.size() + 4);
            d_screen_beta_rot.upload(beta_rot.data(), beta_rot.size());
        }
        std::vector<int32_t> act32(active_set_size);
        for (size_t i = 0; i < active_set_size; ++i) act32[i] = (int32_t)active_set[i];
        d_active_set.upload(act32.data(), active_set_size);
        PinScalars sc{};
        sc.rsq = (double)rsq_io; sc.resid_sum = (double)resid_sum_io; sc.active_set_size = (int)active_set_size;
        *h_sc.p = sc;
        d_sc.upload(h_sc.p, 1);
        PinLaunch<T> L{};
        L.resid = d_r; L.weights = d_w; L.meta = d_meta.p; L.S = (int)S; L.grec = d_grec.p;
        L.beta_in = d_screen_beta.p; L.beta_len = (int)screen_beta.size(); L.is_active_in = d_is_active.p; L.active_set = d_active_set.p; L.sc = d_sc.p;
        L.lmda = (double)lmda_; L.alpha = (double)alpha; L.tol = (double)tol_pin; L.newton_tol = (double)newton_tol;
        L.max_iters = (long long)max_iters; L.newton_max_iters = (int)std::min<size_t>(newton_max_iters, 1u << 30);
        L.max_active_size = (int)std::min<size_t>(max_active_size, (size_t)G); L.intercept = intercept ? 1 : 0;
        L.gs_max = gs_max_screen; L.rec_max = rec_max_screen; L.K = K; L.feat_max = feat_max_screen;
        if (check_interrupt) check_interrupt();
        if (DistContext::get().active()) DistContext::get().allreduce<double>(d_scal.p, 1);    // cheap rank barrier
        const size_t old_active = active_set_size;
        float ms = 0;
        BatchGeometry bg{};
        if (want_batched) bg = plan_batched<T>(X->n_pad(), gs_max_screen, rec_max_screen);
        const bool whole = !static_weights && std::is_same<T, float>::value;        // IRLS: all panels are rebuilt, whole-panel (tensor-core) passes
        if (bg.ok && !static_weights) { pl_screen.entries.clear(); pl_active.entries.clear(); }
        if (!bg.ok) {
            AB_CUDA(cudaEventRecord(ev0, 0));
            X->pin_solve(L);
            AB_CUDA(cudaEventRecord(ev1, 0));
            d_sc.download(h_sc.p, 1);
            AB_CUDA(cudaStreamSynchronize(0));
            ++n_kernel_launches;
            AB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
            launch_cols.push_back((double)h_sc.p->n_col_updates); launch_sweeps.push_back((double)h_sc.p->iters); launch_ms.push_back(ms);
        } else {
            // batched look-ahead kernel: Gram panels for both sweep orders; the kernel hands control back when the active list
            // outgrew its panels (after a screen sweep that added groups), everything else continues on the device
            std::vector<int32_t> ent(S);
            std::iota(ent.begin(), ent.end(), 0);
            { AB_TIME(timers, "pin_presync"); AB_CUDA(cudaStreamSynchronize(0)); }
            ensure_panels(pl_screen, ent, bg.B, bg.Ccap, d_w, whole);
            BatchLaunch<T> bl{};
            bl.start_phase = kSweepActive; bl.beta_rot_in = d_screen_beta_rot.p;
            size_t act_now = active_set_size;
            double cols_seen = 0, sweeps_seen = 0;
            while (true) {
                ent.assign(act32.begin(), act32.begin() + act_now);
                ensure_panels(pl_active, ent, bg.B, bg.Ccap, d_w, whole);
                bl.panels_screen = pl_screen.Q.p; bl.panels_active = pl_active.Q.p; bl.n_active_panelled = (int)act_now;
                { AB_TIME(timers, "pin_launch");
                  if (DistContext::get().active()) DistContext::get().allreduce<double>(d_scal.p, 1);    // rank barrier: the peers' exchange slots of the previous launch are free
                  AB_CUDA(cudaEventRecord(ev0, 0));
                  X->pin_solve_batched(L, bg, bl);
                  AB_CUDA(cudaEventRecord(ev1, 0));
                  d_sc.download(h_sc.p, 1); }
                { AB_TIME(timers, "pin_sync"); AB_CUDA(cudaStreamSynchronize(0)); }
                ++n_kernel_launches; ++n_batched_launches;
                float m1 = 0; AB_CUDA(cudaEventElapsedTime(&m1, ev0, ev1));
                ms += m1;
                launch_cols.push_back((double)h_sc.p->n_col_updates - cols_seen); launch_sweeps.push_back((double)h_sc.p->iters - sweeps_seen);
                launch_ms.push_back(m1);
                cols_seen = (double)h_sc.p->n_col_updates; sweeps_seen = (double)h_sc.p->iters;
                if (h_sc.p->error || h_sc.p->pad != kBatchNeedPanels) break;
                // the active list grew: fetch the new entries, make replica 0 the input of the next launch
                const size_t act_new = (size_t)h_sc.p->active_set_size;
                act32.resize(act_new);
                d_active_set.download(act32.data() + act_now, act_new - act_now, act_now);
                AB_CUDA(cudaMemcpyAsync(d_screen_beta.p, X->beta_rep.p, screen_beta.size() * sizeof(T), cudaMemcpyDeviceToDevice, 0));
                AB_CUDA(cudaMemcpyAsync(d_screen_beta_rot.p, X->brot_rep.p, screen_beta.size() * sizeof(T), cudaMemcpyDeviceToDevice, 0));
                AB_CUDA(cudaMemcpyAsync(d_is_active.p, X->act_rep.p, S, cudaMemcpyDeviceToDevice, 0));
                AB_CUDA(cudaStreamSynchronize(0));
                act_now = act_new;
            }
        }
        { AB_TIME(timers, "pin_download");
          X->beta_rep.download(screen_beta.data(), screen_beta.size());     // replica 0
          X->act_rep.download(screen_is_active.data(), S);                   // replica 0
          AB_CUDA(cudaStreamSynchronize(0)); }
        ++n_pin_solves;
        time_sweep_kernel += ms * 1e-3;
        sc = *h_sc.p;
        if (sc.error) {
            if (sc.error == kErrAbort) {
                int zero = 0; SweepContext::get().abort_flag.upload(&zero, 1); AB_CUDA(cudaStreamSynchronize(0));
                throw solver_error("fused sweep kernel aborted (inter-CTA exchange timed out).");
            }
            if (sc.error == kErrMaxCds) throw solver_error("max coordinate descents reached at lambda index: 0.");
            if (sc.error == kErrMaxActive) throw solver_error("Maximum number of active groups reached.");
            if (sc.error == kErrNewton) throw solver_error("Newton-ABS max iterations reached! Try increasing newton_max_iters.");
            throw solver_error("unknown device error.");
        }
        active_set_size = sc.active_set_size;
        if (active_set_size > old_active) {
            std::vector<int32_t> nw(active_set_size - old_active);
            d_active_set.download(nw.data(), nw.size(), old_active);
            AB_CUDA(cudaStreamSynchronize(0));
            for (size_t i = 0; i < nw.size(); ++i) active_set[old_active + i] = nw[i];
        }
        rsq_io = (T)sc.rsq; resid_sum_io = (T)sc.resid_sum;
        n_sweeps += sc.iters; n_group_updates += sc.n_group_updates; n_col_updates += sc.n_col_updates;
        PinResult R;
        R.iters = sc.iters; R.rsq = sc.rsq;
        R.active_time = ms * 1e-3; R.screen_time = 0;   // one fused launch: not separable without extra syncs
        // active_order + sparsify_active_beta (solver_gaussian_pin_naive.hpp:360-391, pin_base.hpp:58-98)
        std::vector<size_t> order(active_set_size);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](size_t i, size_t j) {
            return groups[screen_set[active_set[i]]] < groups[screen_set[active_set[j]]];
        });
        for (size_t i = 0; i < order.size(); ++i) {
            const idx_t ss = active_set[order[i]], g = screen_set[ss], gs = group_sizes[g];
            for (idx_t c = 0; c < gs; ++c) { R.beta.idx.push_back(groups[g] + c); R.beta.val.push_back((double)screen_beta[screen_begins[ss] + c]); }
        }
        R.intercept = (intercept ? 1.0 : 0.0) * ((double)y_mean_ + sc.resid_sum);      // :392
        return R;
    }

    // ------------------------------------------------------------------ Gaussian fit (solver_gaussian_naive.hpp:215-349)
    PinResult fit_gaussian(T lmda_) {
        const int64_t np = X->n_pad() * K;
        AB_CUDA(cudaMemcpyAsync(d_resid_prev.p, d_resid.p, np * sizeof(T), cudaMemcpyDeviceToDevice, 0));
        std::vector<T> beta_prev = screen_beta; std::vector<int8_t> act_prev = screen_is_active;
        upload_screen_tables(h_meta, h_grec, false);
        try {
            return run_pin(d_resid.p, d_weights.p, lmda_, tol * y_var, y_mean, rsq, resid_sum, true);
        } catch (...) {
            std::swap(d_resid.p, d_resid_prev.p);
            screen_beta.swap(beta_prev); screen_is_active.swap(act_prev);
            throw;
        }
    }

    // ------------------------------------------------------------------ GLM pieces (solver_glm_naive.hpp)
    void update_loss_null();              // defined in solver_glm.cuh
    void update_loss_null_multi();
    PinResult fit_glm(T lmda_);           // defined in solver_glm.cuh

    PinResult fit(T lmda_) { return is_glm ? fit_glm(lmda_) : fit_gaussian(lmda_); }

    // update_invariance (solver_gaussian_naive.hpp:377-393 / solver_glm_naive.hpp:495-503): grad = X^T (w o r) (- resid_sum * X_means),
    // abs_grad.  During the path only the scores of NON-screen groups are ever read (kkt :408-433 and screen :273-403 skip the screen
    // set), so once part of X is screened the pass streams only the other columns (Configs::kkt_skip_screen; late in a path that is
    // a small fraction of X, and the pass is the second largest item of a step).  The full gradient -- an output of the state -- is
    // brought up to date once, when the solve ends (finalize_invariance).
    bool grad_partial = false;
    DevBuf<int32_t> d_ns_cols; std::vector<int32_t> ns_cols, ns_logical; size_t ns_for_screen = (size_t)-1;
    void update_invariance(T lmda_, bool force_full = false) {
        AB_TIME(timers, "invariance");
        lmda = lmda_;
        DistContext& dc = DistContext::get();
        const bool can_skip = Configs::kkt_skip_screen && !force_full && K == 1 && !X->sparse && !X->snp && !screen_set.empty();
        if (can_skip) {
            if (ns_for_screen != screen_set.size()) {                  // the screen set only grows: rebuild the list when it did
                ns_cols.clear(); ns_logical.clear();
                const bool flat = (idx_t)in_screen.size() == G;
                for (idx_t g = 0; g < G; ++g) {
                    if (flat ? in_screen[g] : (uint8_t)screen_hashset.count(g)) continue;
                    const int32_t pc = X->phys_col(groups[g], (int)group_sizes[g]);
                    for (idx_t c = 0; c < group_sizes[g]; ++c) { ns_cols.push_back(pc + (int32_t)c); ns_logical.push_back((int32_t)(groups[g] + c)); }
                }
                d_ns_cols.reserve_keep(ns_cols.size() + 4);
                if (!ns_cols.empty()) d_ns_cols.upload(ns_cols.data(), ns_cols.size());
                ns_for_screen = screen_set.size();
            }
            const size_t q = ns_cols.size();
            if (q) {
                d_tmp.reserve_keep(q);
                X->d_gemv_t(0, d_ns_cols.p, (int)q, d_resid.p, is_glm ? X->d_ones() : d_weights.p, d_tmp.p);
                dc.allreduce<T>(d_tmp.p, (int64_t)q);
                std::vector<T> h(q);
                d_tmp.download(h.data(), q);
                AB_CUDA(cudaStreamSynchronize(0));
                const bool sub = !is_glm && intercept;
                for (size_t k = 0; k < q; ++k) { const int32_t j = ns_logical[k]; grad[j] = sub ? h[k] - resid_sum * X_means[j] : h[k]; }
                n_kernel_launches += 2;
            }
            grad_partial = true;
            update_abs_grad(lmda_);
            return;
        }
        if (K > 1) {
            X->d_mul_multi(K, (int)n_int, d_resid.p, is_glm ? nullptr : d_weights.p, d_grad.p);     // state-level intercept is always off
            dc.allreduce<T>(d_grad.p, p);
        } else if (is_glm) {
            X->d_mul(d_resid.p, X->d_ones(), d_grad.p);
            dc.allreduce<T>(d_grad.p, p);
        } else if (!dc.active()) {
            // grad = X^T (w o r) - resid_sum * X_means, epilogue fused into the reduction kernel
            const double rs = (double)resid_sum;
            X->d_gemv_t(0, nullptr, (int)p, d_resid.p, d_weights.p, d_grad.p, false,
                        intercept ? d_X_means.p : nullptr, nullptr, rs);
        } else {
            X->d_gemv_t(0, nullptr, (int)p, d_resid.p, d_weights.p, d_grad.p);
            dc.allreduce<T>(d_grad.p, p);                                       // KKT scores: one all-reduce of p values per round
        }
        d_grad.download(grad.data(), p);
        AB_CUDA(cudaStreamSynchronize(0));
        if (K == 1 && !is_glm && dc.active() && intercept) for (idx_t j = 0; j < p; ++j) grad[j] -= resid_sum * X_means[j];
        n_kernel_launches += 2;
        grad_partial = false;
        update_abs_grad(lmda_);
    }
    // the state's grad / abs_grad are outputs: complete them after a solve that skipped the screen columns
    void finalize_invariance() {
        if (!grad_partial) return;
        update_invariance(lmda, true);
    }

    void update_solutions(PinResult& pr, T lmda_) {
        AB_TIME(timers, "update_solutions");
        betas.emplace_back(std::move(pr.beta));
        intercepts.push_back((T)pr.intercept);
        lmdas.push_back(lmda_);
        if (is_glm) {
            const T loss = glm->loss(d_eta.p);
            devs.push_back((loss_null - loss) / (loss_null - loss_full));       // solver_glm_naive.hpp:153-157
        } else {
            devs.push_back((T)pr.rsq / y_var);                                  // solver_gaussian_naive.hpp:205-206
        }
    }

    // screen (solver_base.hpp:273-403)
    void screen(T lmda_next, bool all_kkt_passed, int n_new_active) {
        AB_TIME(timers, "screen_host");
        const int old_size = (int)screen_set.size();
        // (membership as of entry: in_screen / screen_hashset are only updated by update_screen_derived_base afterwards)
        const bool flat = (idx_t)in_screen.size() == G;
        auto is_screen = [&](idx_t i) { return flat ? in_screen[i] != 0 : screen_hashset.count(i) > 0; };
        if (screen_rule == 0) {
            const T strong = (2 * lmda_next - lmda) * alpha;
            for (idx_t i = 0; i < G; ++i) { if (is_screen(i)) continue; if (abs_grad[i] > strong * penalty[i]) screen_set.push_back(i); }
        } else if (screen_rule == 1) {
            if (n_new_active) {
                std::vector<T> wts(G);
                for (idx_t i = 0; i < G; ++i) wts[i] = (penalty[i] <= 0) ? alpha * lmda : std::min(abs_grad[i] / penalty[i], alpha * lmda);
                // argsort by weight; (key, index) pairs sort contiguously (the indirect comparator was 0.6 s per path at G = 200k)
                std::vector<std::pair<T, idx_t>> keyed(G);
                for (idx_t i = 0; i < G; ++i) keyed[i] = {wts[i], i};
                // std::sort with a key-only comparator performs the same comparisons / moves as sorting the indices through wts[]: same permutation
                std::sort(keyed.begin(), keyed.end(), [](const std::pair<T, idx_t>& x, const std::pair<T, idx_t>& y) { return x.first < y.first; });
                std::vector<idx_t> order(G);
                for (idx_t i = 0; i < G; ++i) order[i] = keyed[i].second;
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
        const bool flat = (idx_t)in_screen.size() == G;
        for (idx_t k = 0; k < G; ++k) { if (flat ? in_screen[k] : (uint8_t)screen_hashset.count(k)) continue; if (abs_grad[k] > lmda_ * alpha * penalty[k]) return false; }
        return true;
    }

    bool early_exit_f() {                                                       // solver_base.hpp:241-263 + user exit_cond
        bool r = false;
        if (early_exit && !devs.empty()) {
            const T u = devs.back();
            if (u >= adev_tol) r = true;
            else if (devs.size() >= 2 && std::abs(u - devs[devs.size() - 2]) < ddev_tol) r = true;
        }
        return r || (exit_cond && exit_cond());
    }

    void screen_f(T lmda_, bool kkt_passed, int n_new_active) {
        screen(lmda_, kkt_passed, n_new_active);
        if (is_glm) update_screen_derived_base();
        else update_screen_derived_gaussian();
    }

    // ------------------------------------------------------------------ the pin state in isolation
    // pin::naive::solve over the state's own lmda_path on a FIXED screen set (solver_gaussian_pin_naive.hpp:223-401; the core of
    // StateGaussianPinNaive, adelie/src/py_state.cpp:389-411).  `tol` is used as given (the path driver passes tol * y_var, :314),
    // `iters` and max_iters are cumulative over the lambdas (:201, :327), outputs are betas / intercepts / rsqs / lmdas and the
    // per-lambda timers (benchmark_screen / benchmark_active live in benchmark_fit_screen / benchmark_fit_active).
    DevBuf<double> d_cov_means; std::vector<double> scr_M;
    std::vector<double> scr_C, scr_eig_in, scr_eig_V, scr_eig_D; std::vector<EigItem> scr_eig_items; std::vector<int64_t> scr_eig_slot;     // compute_screen_records
    std::vector<GroupMeta> glm_meta; std::vector<T> glm_grec, glm_sXm, glm_sv; std::vector<std::vector<T>> glm_stv;                          // fit_glm, per IRLS iteration
    std::vector<T> rsqs;
    void solve_pin() {
        if (is_glm) throw core_error("the pin state is a Gaussian state.");
        upload_screen_tables(h_meta, h_grec, false);
        const size_t max_iters_total = max_iters;
        for (size_t l = 0; l < lmda_path.size(); ++l) {
            max_iters = (size_t)n_sweeps >= max_iters_total ? 0 : max_iters_total - (size_t)n_sweeps;
            PinResult pr;
            try { pr = run_pin(d_resid.p, d_weights.p, lmda_path[l], tol, y_mean, rsq, resid_sum, true); }
            catch (const solver_error& e) {
                max_iters = max_iters_total;
                if (std::string(e.what()).find("max coordinate descents") != std::string::npos)
                    throw solver_error("max coordinate descents reached at lambda index: " + std::to_string(l) + ".");
                throw;
            }
            catch (...) { max_iters = max_iters_total; throw; }
            max_iters = max_iters_total;
            betas.emplace_back(std::move(pr.beta));
            intercepts.push_back((T)pr.intercept);
            rsqs.push_back(rsq);
            lmdas.push_back(lmda_path[l]);
            benchmark_fit_screen.push_back(pr.screen_time);
            benchmark_fit_active.push_back(pr.active_time);
            lmda = lmda_path[l];
            if (rsq >= adev_tol * y_var) break;                                             // :398
            if (l >= 1 && rsqs[l] - rsqs[l - 1] <= ddev_tol * y_var) break;                  // :399
        }
    }

    // ------------------------------------------------------------------ solve_core (solver_base.hpp:435-687)
    void solve() {
        if (screen_set.size() > max_screen_size) throw solver_error("maximum screen set size reached.");
        if (is_glm && setup_loss_null) update_loss_null();
        if (setup_lmda_max) {
            T pmax = penalty[0];
            for (auto v : penalty) pmax = std::max(pmax, v);
            const T large_lmda = T(1e-3 * std::numeric_limits<T>::max() / std::max<T>(1, pmax));
            fit(large_lmda);
            update_invariance(large_lmda);
            const T factor = (alpha <= 0) ? T(1e-3) : alpha;                    // solver/utils.hpp:6-23
            T m = -std::numeric_limits<T>::infinity();
            for (idx_t i = 0; i < G; ++i) m = std::max<T>(m, (penalty[i] <= 0.0) ? T(0.0) : abs_grad[i] / penalty[i]);
            lmda_max = m / factor;
        }
        if (setup_lmda_path) {
            if (lmda_path_size <= 0) return;
            lmda_path.resize(lmda_path_size);
            const size_t L = lmda_path_size;
            if (L > 1) {                                                        // solver/utils.hpp:25-41
                const T log_factor = std::log(min_ratio) / (L - 1);
                for (size_t i = 0; i < L; ++i) lmda_path[i] = lmda_max * std::exp(log_factor * T(i));
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
                if (i + 1 < large.size()) {
                    update_solutions(pr, large[i]);
                    if (early_exit_f()) return;
                } else update_invariance(large[i]);
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
                update_invariance(lmda_curr);
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
