// adelie_b200/csrc/sweep_batched.cuh -- the batched, look-ahead variant of the fused pin solve.
//
// Same algorithm and the same iterates (up to rounding) as pin_solve_kernel (sweep.cuh), i.e. pin::naive::solve /
// coordinate_descent of the reference (CORE/solver/solver_gaussian_pin_naive.hpp:26-168, 181-215, 223-401), reorganised
// so that the Gauss-Seidel dependency chain no longer contains a grid-wide exchange per group:
//
//   * the sweep list (screen positions or the active list) is cut into BATCHES of B consecutive groups; the partial
//     gradients X_g^T (w o r) of all groups of a batch are computed against the SAME residual and all-reduced over the
//     CTAs in ONE flagged-line exchange;
//   * exactness of Gauss-Seidel is restored with small precomputed Gram panels: once group k of the batch moved by
//     del_k = beta_old - beta_new (r += X_k del_k), the stale gradient of every later group k' is corrected by
//     G[k', k] del_k with G[k', k] = X_k'^T W X_k (weights are static for the Gaussian path, so a panel is computed once
//     when its batch is formed -- the lists are append-only over the whole path -- by pair_gram_kernel);
//   * LOOK-AHEAD: the dot phase of batch b+1 runs against the residual that lacks batch b's update while the control warp
//     is still solving batch b's proximal problems; the panel of batch b also holds the cross block X_b^T W X_{b+1},
//     so the HBM stream of the next batch, its exchange and the serial prox chain overlap.
//
// Warp roles inside each of the persistent CTAs (one per SM, 512 threads):
//   warp 0          control warp (CW): replicated proximal updates + gradient corrections, sweep control flow
//   warp 4          TMA producer: streams column tiles (cp.async.bulk -> stage ring) and panels/records (panel ring)
//   warps 8, 12     exchange warps (EW): two-level flagged-line all-reduce through L2, asynchronous to everything else
//   other 12 warps  data warps (DW): dot phase D(b) from the staged tiles, residual update U(b) from L2 (the tile was
//                   streamed through L2 one batch earlier), each thread owning a private set of residual rows.
// Schedule of the data warps: D(0); for b: { D(b+1); U(b) }.  The control roles all sit on scheduler 0 so that the
// latency-critical control warp does not compete with the data warps for issue slots.
#pragma once
#include "sweep.cuh"
#include "gram_tc.cuh"
#include <type_traits>

namespace ab {

constexpr int kBatchMax = 8;           // most groups per batch
constexpr int kBatchColsMax = 64;      // most columns per batch (Ccap <= 64): one pass of the 64 exchange threads
constexpr int kBatchLLSlots = 4;       // exchange buffers are 4-deep (see the race analysis at ew_exchange)
constexpr int kBatchStages = 8;        // most stages of the column-tile ring
constexpr int kBatchDW = 12;           // data warps
constexpr int kBatchNDT = kBatchDW * 32;

enum { kBatchDone = 0, kBatchNeedPanels = 1 };

template <class T>
struct BatchKernelArgs {
    const T* X; int64_t ld;
    T* resid; const T* weights;
    const GroupMeta* meta; int S; const T* grec;
    const T* beta_in; T* beta_rep; int64_t beta_stride; int beta_len;
    const T* brot_in; T* brot_rep;                       // coefficients in the groups' eigenbases (a V), same layout / replicas
    int use_ext;                                         // 1: panel slots carry the extended records (all groups <= 12 columns)
    const int8_t* is_active_in; int8_t* is_active_rep; int64_t act_stride;
    int32_t* active_set;
    PinScalars* sc;
    const T* panels_screen; const T* panels_active;      // [n_batches][Ccap][2 * Ccap]
    int n_active_panelled;                               // active-list positions covered by panels_active
    int B, Ccap;
    dev::LLLine* ll1; dev::LLLine* ll2; int ncta_pad; int fan;
    // row-sharded multi-GPU (level 3 of the exchange): ll3_peer[r] = rank r's line buffer [4][64][8] mapped over NVLink
    dev::LLLine* ll3_peer[8]; int rank, world;
    uint32_t* epoch; int* abort_flag;
    double lmda, alpha, tol, newton_tol, dbeta_tol;
    long long max_iters; int newton_max_iters; int max_active_size; int intercept;
    int start_phase;
    int units_base, units_rem, rows_stride;
    int n_stages, stage_elems;
    int ch;                                              // columns per ring item (a group is streamed in ceil(gs / ch) <= 2 items)
    int l2_prefetch;                                     // 1: the producer warp prefetches the tiles of batch b + 2 into L2 while batch b + 1 streams
    int u_prefetch;                                      // 1: update tiles of active-set sweeps are prefetched ahead of the proximal updates
    int rec_stride;                                      // elements between the records inside a panel slot
    int pslot_elems;                                     // elements per panel slot = Ccap * 2 Ccap + B * rec_stride
    long long* stats;
};

struct BatchCtrl {
    int sw_seq;                   // number of published sweep descriptors
    int sw_kind[4], sw_count[4];
    int prox_done;                // groups completed by the control warp (monotone over the launch)
    int gready;                   // batches whose all-reduced gradient is in gstale (monotone)
    int halt;                     // CTA-local shutdown (error / abort)
    int error;
    int changed[2][kBatchMax];
};

namespace dev {
__device__ __forceinline__ void st_release_cta(int* p, int v) {
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
// spin until *p >= target; false if the CTA was halted / the kernel aborted
__device__ __forceinline__ bool wait_counter(const int* p, int target, volatile int* abort_flag, volatile int* halt) {
    SpinGuard g;
    while (ld_acquire_cta(p) < target) {
        if (g.give_up(abort_flag, halt)) return false;
        __nanosleep(64);                       // do not steal issue slots / shared-memory bandwidth from the control warp
    }
    return true;
}
} // namespace dev

namespace dev {
// L2 eviction-priority policies: the column tiles streamed for the dot phase are re-read by the residual update one
// batch later (keep: evict_last); that second read is the last use (evict_first frees the lines for the next tiles).
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// L2 prefetch of a contiguous global range (16-byte multiple): no destination, no completion to wait for
__device__ __forceinline__ void prefetch_l2_bulk(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ld_hint(const float* p, float (&v)[4], uint64_t pol) {
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void ld_hint(const double* p, double (&v)[2], uint64_t pol) {
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v[0]), "=d"(v[1]) : "l"(p), "l"(pol));
}
} // namespace dev

// Branch-free reciprocal / square root for the Newton iteration of the control warp.  IEEE float division compiles to a
// fast path + range check + slow-path call per use, which serialises the (independent) divisions of one evaluation of phi;
// MUFU.RCP + one Newton refinement (<= 1 ulp) pipelines.  double keeps the IEEE operations.
template <class P> struct FastMath;
template <> struct FastMath<float> {
    static __device__ __forceinline__ float rcp(float x) {
        float u; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(x));
        return fmaf(u, fmaf(-x, u, 1.0f), u);
    }
    static __device__ __forceinline__ float sqrt_(float x) { float u; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(x)); return u; }
};
template <> struct FastMath<double> {
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
};

// Group update (solver_gaussian_pin_naive.hpp:109-164) for 1 < gs <= GSP <= 12 by the control warp.  A lone warp issues
// roughly one instruction every 4-6 cycles, so the update is written for FEW INSTRUCTIONS and SHORT DEPENDENT CHAINS:
// one coefficient per lane (lanes 16-31 mirror lanes 0-15 so that every decision is warp-uniform), 16-lane butterfly
// reductions, MUFU reciprocals, and everything that does not depend on the gradient loaded one group ahead (prox_pre, issued
// behind the previous group's corrections).  Newton iteration on h = ||x|| as newton_solver (newton.hpp:44-142); start point:
// the previous norm ||beta_g|| when phi there is >= 0 (left of the root, from where the iteration is monotone, like h0 = 0 of
// the reference), otherwise 0; only the converged root matters for parity.
template <class P, int GSP> struct LanePre { P A, xm, xmt, aold_c, ao, h0sq; P vcol[GSP], vrow[GSP]; };

template <int N, class P>
__device__ __forceinline__ void seg16_allsum(P (&v)[N]) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        P t[N];
#pragma unroll
        for (int i = 0; i < N; ++i) t[i] = __shfl_xor_sync(0xffffffffu, v[i], o);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] += t[i];
    }
}

// ext: the group's extended record [A(gsp) | xm(gsp) | V^T xm(gsp) | V rows (gs x gsp) | V^T rows (gs x gsp)], gsp = gs rounded up
// to 4 (zero padded), so that a lane's row of V / V^T comes in with 16-byte loads.  aold_s / arot_s: the group's current
// coefficients in the original basis / in the eigenbasis.
template <class T, class P, int GSP>
__device__ __forceinline__ void prox_pre(LanePre<P, GSP>& r, const T* ext, int gs, const P* aold_s, const P* arot_s, int lane) {
    constexpr int VN = VecT<T>::N;
    const int l16 = lane & 15;
    const bool on = l16 < gs;
    const int c = on ? l16 : 0;
    const int gsp = (gs + 3) & ~3;
    r.A = on ? (P)ext[c] : P(0); r.xm = on ? (P)ext[gsp + c] : P(0); r.xmt = on ? (P)ext[2 * gsp + c] : P(0);
    r.aold_c = on ? aold_s[c] : P(0);
    r.ao = on ? arot_s[c] : P(0);
    const T* vr = ext + 3 * gsp + c * gsp;                 // V[c][.]    (rotate back)
    const T* vc = vr + gs * gsp;                           // V^T[c][.] = V[.][c]   (rotate the gradient)
#pragma unroll
    for (int q = 0; q < GSP; q += VN) {
        if (q < gsp) {
            vec_load<T>(vr + q, reinterpret_cast<T(&)[VN]>(r.vrow[q]));
            vec_load<T>(vc + q, reinterpret_cast<T(&)[VN]>(r.vcol[q]));
        } else {
#pragma unroll
            for (int k = 0; k < VN; ++k) { r.vrow[q + k] = 0; r.vcol[q + k] = 0; }
        }
    }
    P v1[1] = {r.ao * r.ao};
    seg16_allsum<1>(v1);
    r.h0sq = v1[0];
}

// bc: shared scratch of 32 elements.  g_in: the group's gradient X_g^T (w o r), element (lane & 15).
// Returns 1 if the coefficients moved; del (original basis) in dl[0..gs) and, per lane, in del_c.
template <class T, class P, int GSP>
__device__ __forceinline__ int prox_post(const LanePre<P, GSP>& pre, int gs, P g_in, P l1k, P l2k, P tol, int max_iters,
                                         P dbeta_tol, int intercept, ProxState& ps, int lane, P* bc, T* beta_g, T* brot_g, T* dl, P& del_c, long long* pp)
{
    constexpr int VN = VecT<P>::N;
    using FM = FastMath<P>;
    const int l16 = lane & 15;
    const bool on = l16 < gs;
    // ---- gradient in the original basis, broadcast, rotated: gt = gk V + A (a_old V)
    P gk = on ? g_in : P(0);
    if (intercept && on) gk -= (P)ps.resid_sum * pre.xm;
    if (lane < 16) bc[lane] = gk;
    __syncwarp();
    P gt0a = 0, gt0b = 0;
    {
        P gka[GSP];
#pragma unroll
        for (int q = 0; q < GSP; q += VN) vec_load<P>(bc + q, reinterpret_cast<P(&)[VN]>(gka[q]));
#pragma unroll
        for (int q = 0; q < GSP; q += 2) { gt0a += gka[q] * pre.vcol[q]; gt0b += gka[q + 1] * pre.vcol[q + 1]; }
    }
    const P gt0 = on ? gt0a + gt0b : P(0);
    const P gt = gt0 + pre.A * pre.ao;
    const P D = on ? pre.A + l2k : P(1);
    // ---- root of phi(h) = sum (gt / (D h + l1))^2 - 1
    P at = 0; int nit = 0;
    if (l1k <= P(0)) {
        at = (on && gt != P(0)) ? gt / (pre.A + l2k) : P(0);
    } else {
        const P tol_eff = fmax(tol, ProxEps<P>::floor_tol());
        // Start: phi(0) -- which doubles as the ||v|| <= l1 test (newton.hpp:62-66) -- and phi at the warm start travel through ONE
        // butterfly (round 2: they were two dependent evaluations, ~250 cycles of the control warp's chain per group update).
        const P hw = (pre.h0sq > P(0)) ? FM::sqrt_(pre.h0sq) : P(0);
        P h = 0, u = FM::rcp(l1k), t, sd;
        bool zero = false;
        {
            const P uw = FM::rcp(D * hw + l1k);
            const P q0 = gt * u, qw = gt * uw;
            P v4s[4] = {q0 * q0, q0 * q0 * D * u, qw * qw, qw * qw * D * uw};
            seg16_allsum<4>(v4s);
            t = v4s[0]; sd = v4s[1];
            if (!(t > P(1))) zero = true;
            else if (hw > P(0) && (v4s[2] - P(1) >= -tol_eff)) { h = hw; u = uw; t = v4s[2]; sd = v4s[3]; }   // left of the root: monotone from here
        }
        if (!zero) {
#pragma unroll 1
            while (fabs(t - P(1)) > tol_eff && nit < max_iters) {
                // h - fh / dfh with dfh = -sd (1 + sqrt t) / t   (optimization/newton.hpp:56-63), one reciprocal
                const P hn = fmax(h + (t - P(1)) * t * FM::rcp(sd * (P(1) + FM::sqrt_(t))), P(0));
                if (hn == h) break;                                   // no representable progress left
                h = hn; ++nit;
                u = FM::rcp(D * h + l1k);
                const P qq = gt * u;
                const P xx = qq * qq;
                P v2[2] = {xx, xx * D * u};
                seg16_allsum<2>(v2);
                t = v2[0]; sd = v2[1];
            }
        }
        at = (on && !zero) ? h * gt * u : P(0);                   // x = h v / (D h + l1)   (newton.hpp:109)
    }
    ps.newton_iters_max = max(ps.newton_iters_max, nit);
    if (nit >= max_iters) ps.error = kErrNewton;
    if (pp) pp[5] += nit;
    // ---- update test / bookkeeping sums
    const P d = at - pre.ao;
    P v4[4] = {d * d, pre.A * d * d, d * (2 * gt0 - d * pre.A), -pre.xmt * d};
    seg16_allsum<4>(v4);
    if (v4[0] <= dbeta_tol * dbeta_tol * (P)gs) { if (pp) ++pp[6]; return 0; }      // ||d|| <= dbeta_tol sqrt(gs) (:146-147), squared
    ps.cm = fmax(ps.cm, (double)(v4[1] * FM::rcp((P)gs)));
    ps.rsq += (double)v4[2];
    ps.resid_sum += (double)v4[3];
    // ---- back to the original basis: del = a_old - a_new = V (ao - at) = -V d
    __syncwarp();
    if (lane < 16) bc[16 + lane] = d;
    __syncwarp();
    P dea = 0, deb = 0;
    {
        P da[GSP];
#pragma unroll
        for (int q = 0; q < GSP; q += VN) vec_load<P>(bc + 16 + q, reinterpret_cast<P(&)[VN]>(da[q]));
#pragma unroll
        for (int q = 0; q < GSP; q += 2) { dea -= da[q] * pre.vrow[q]; deb -= da[q + 1] * pre.vrow[q + 1]; }
    }
    const T anT = (T)(pre.aold_c - (dea + deb));
    del_c = on ? (P)(T)(pre.aold_c - (P)anT) : P(0);
    if (lane < gs) { beta_g[lane] = anT; brot_g[lane] = (T)at; dl[lane] = (T)del_c; }
    return 1;
}

// Gradient corrections after group (off, gs) of the current batch moved by del: ga[j] += sum_c Q[off + c][4 lane + j] del[c] in
// "panel column space" t = 4 lane + j: t < Ccap are the columns of this batch (only the later ones are used afterwards),
// t >= Ccap the columns of the next batch.  One 16-byte load per source column.  del_c: lane c holds del[c]; GSP == 0: del
// comes from dl (large groups).
template <class T, class P, int GSP>
__device__ __forceinline__ void apply_corrections(const T* Q, int ldq, int off, int gs, const T* dl, P del_c, P (&ga)[4], int lane, P* bc)
{
    const bool act = 4 * lane < ldq;
    const T* qcol = Q + (size_t)off * ldq + (act ? 4 * lane : 0);
    if constexpr (GSP > 0) {
        // Round 2: in-kernel counters put this routine at 1880 cycles per group update -- one shuffle + one dependent 16-byte load + 4 FMAs
        // per source column, and at the kernel's register cap the compiler does not hoist the loads over the FMAs, so a lone warp pays the
        // full shared-memory latency twelve times in a row.  Now: del of the whole group reaches every lane through shared memory (one
        // store + GSP / VN broadcast loads instead of GSP shuffles), the panel rows come in waves of four independent loads, and two
        // accumulator sets halve the FMA chain.
        constexpr int VNP = VecT<P>::N;
        P dall[GSP < 4 ? 4 : GSP];          // (GSP = 2 only exists as dead code of the generic instantiation)
        __syncwarp();
        if (lane < 16) bc[lane] = del_c;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < GSP; q += VNP) vec_load<P>(bc + q, reinterpret_cast<P(&)[VNP]>(dall[q]));
        P gb[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c0 = 0; c0 < GSP; c0 += 4) {
            T qv[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (c0 + i < gs && act) {
                    if constexpr (sizeof(T) == 4) vec_load<T>(qcol + (size_t)(c0 + i) * ldq, qv[i]);
                    else {
                        vec_load<T>(qcol + (size_t)(c0 + i) * ldq, reinterpret_cast<T(&)[2]>(qv[i][0]));
                        vec_load<T>(qcol + (size_t)(c0 + i) * ldq + 2, reinterpret_cast<T(&)[2]>(qv[i][2]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) qv[i][j] = T(0);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ga[j] += (P)qv[0][j] * dall[c0] + (P)qv[2][j] * dall[c0 + 2];
                gb[j] += (P)qv[1][j] * dall[c0 + 1] + (P)qv[3][j] * dall[c0 + 3];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) ga[j] += gb[j];
    } else {
#pragma unroll 2
        for (int c = 0; c < 32; ++c) {
            const P dc = (c < gs ? (P)dl[c] : P(0));
            if (c < gs && act) {
                if constexpr (sizeof(T) == 4) {
                    T q4[4];
                    vec_load<T>(qcol + (size_t)c * ldq, q4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) ga[j] += (P)q4[j] * dc;
                } else {
                    T q2[2], q3[2];
                    vec_load<T>(qcol + (size_t)c * ldq, q2); vec_load<T>(qcol + (size_t)c * ldq + 2, q3);
                    ga[0] += (P)q2[0] * dc; ga[1] += (P)q2[1] * dc; ga[2] += (P)q3[0] * dc; ga[3] += (P)q3[1] * dc;
                }
            }
        }
    }
}

template <class T>
struct BatchSmem {
    static constexpr size_t kHeaderBytes = 512;
    // header | gstale[2][Ccap] f64 | corr[2][Ccap] | gcur[32] f64 | del[2][Ccap] | aold[2][Ccap] | p_gk[32] p_at[32] scr[128] |
    // wpart[2][DW][Ccap] f64 | r tile | w tile | panel slots [2] | stages
    __host__ __device__ static size_t fixed_bytes(int Ccap) {
        size_t b = kHeaderBytes + sizeof(double) * ((size_t)Ccap * 10 + 32 + 32 + 32 + 128 + (size_t)2 * kBatchDW * Ccap);
        return (b + 127) / 128 * 128;
    }
};

// ---------------------------------------------------------------------------------------------------------------
template <class T, int PROF>      // PROF: 0 off, 1 stall/busy split per role (cheap), 2 per-phase counters (perturbs: spills)
__global__ void __launch_bounds__(512, 1)
pin_solve_batched_kernel(const __grid_constant__ BatchKernelArgs<T> a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int VN = VecT<T>::N;
    using P = T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int B = a.B, Ccap = a.Ccap, ldq = 2 * a.Ccap;

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);          // [kBatchStages]
    uint64_t* empty_bar = full_bar + kBatchStages;                        // [kBatchStages]
    uint64_t* pfull_bar = empty_bar + kBatchStages;                       // [2]
    uint64_t* pempty_bar = pfull_bar + 2;                                 // [2]
    uint64_t* prox_bar = pempty_bar + 2;                                  // [2][kBatchMax]: group k of a batch (by parity) is solved
    uint64_t* gready_bar = prox_bar + 2 * kBatchMax;                      // [2]: all-reduced gradient of a batch (by parity) is in gstale
    uint64_t* dbar = gready_bar + 2;                                      // [2]: every data warp has written its partials of a batch
    BatchCtrl* ctrl = reinterpret_cast<BatchCtrl*>(smem_raw + 384);
    double* gstale = reinterpret_cast<double*>(smem_raw + BatchSmem<T>::kHeaderBytes);   // [2][Ccap]
    double* corr_raw = gstale + 2 * Ccap;                                 // [2][Ccap] (P)
    double* gcur = corr_raw + 2 * Ccap;                                   // [32]
    double* del_raw = gcur + 32;                                          // [2][Ccap] (T)
    double* aold_raw = del_raw + 2 * Ccap;                                // [2][Ccap] (P)
    double* arot_raw = aold_raw + 2 * Ccap;                               // [2][Ccap] (P)
    double* pscr_raw = arot_raw + 2 * Ccap;                               // p_gk[32] p_at[32] scr[128]
    double* wpart = pscr_raw + 32 + 32 + 128;                             // [2][DW][Ccap]
    P* corr = reinterpret_cast<P*>(corr_raw);
    T* del = reinterpret_cast<T*>(del_raw);
    P* aold = reinterpret_cast<P*>(aold_raw);
    P* arot = reinterpret_cast<P*>(arot_raw);
    P* p_gk = reinterpret_cast<P*>(pscr_raw);
    P* p_at = reinterpret_cast<P*>(pscr_raw + 32);
    P* p_scr = reinterpret_cast<P*>(pscr_raw + 64);
    unsigned char* tiles = smem_raw + BatchSmem<T>::fixed_bytes(Ccap);
    T* sr = reinterpret_cast<T*>(tiles);
    T* sw = sr + a.rows_stride;
    T* pslots = sw + a.rows_stride;                                       // [2][pslot_elems]
    T* stages = pslots + (size_t)2 * a.pslot_elems;

    const int my_units = a.units_base + (cta < a.units_rem ? 1 : 0);
    const int64_t unit0 = (int64_t)cta * a.units_base + min(cta, a.units_rem);
    const int64_t r0 = unit0 * kRowAlign;
    const int rows = my_units * kRowAlign;

    volatile int* abort_flag = a.abort_flag;
    volatile int* halt = &ctrl->halt;

    if (tid == 0) {
        for (int s = 0; s < kBatchStages; ++s) { dev::mbar_init(&full_bar[s], 1); dev::mbar_init(&empty_bar[s], kBatchDW); }
        for (int s = 0; s < 2; ++s) { dev::mbar_init(&pfull_bar[s], 1); dev::mbar_init(&pempty_bar[s], 1); dev::mbar_init(&gready_bar[s], 1); }
        for (int s = 0; s < 2 * kBatchMax; ++s) dev::mbar_init(&prox_bar[s], 1);
        for (int s = 0; s < 2; ++s) dev::mbar_init(&dbar[s], kBatchDW);
        dev::fence_barrier_init();
        ctrl->sw_seq = 0; ctrl->prox_done = 0; ctrl->gready = 0; ctrl->halt = 0; ctrl->error = 0;
    }
    for (int i = tid; i < 4 * Ccap; i += blockDim.x) corr_raw[i] = 0.0;        // corr + (unused tail): both slots start at zero
    T* my_beta = a.beta_rep + (size_t)cta * a.beta_stride;
    int8_t* my_active = a.is_active_rep + (size_t)cta * a.act_stride;
    T* my_brot = a.brot_rep + (size_t)cta * a.beta_stride;
    for (int i = tid; i < a.beta_len; i += blockDim.x) { my_beta[i] = a.beta_in[i]; my_brot[i] = a.brot_in[i]; }
    for (int i = tid; i < a.S; i += blockDim.x) my_active[i] = a.is_active_in[i];
    // resident r / w tiles
    {
        T* gr = a.resid + r0; const T* gw = a.weights + r0;
        for (int v = tid; v < rows / VN; v += blockDim.x) {
            T t[VN];
            vec_load<T>(gr + (size_t)v * VN, t); vec_store<T>(sr + (size_t)v * VN, t);
            vec_load<T>(gw + (size_t)v * VN, t); vec_store<T>(sw + (size_t)v * VN, t);
        }
    }
    const uint32_t epoch0 = dev::ld_cg(a.epoch);
    __syncthreads();

    const int role_cw = (warp == 0), role_prod = (warp == 4), role_ew = (warp == 8 || warp == 12);

    // =====================================================================================================
    // TMA producer
    // =====================================================================================================
    if (role_prod) {
        uint32_t gitem = 0, pitem = 0;
        int sweep = 0, ubatch = 0;
        bool running = true;
        const uint64_t pol_keep = dev::policy_evict_last(), pol_done = dev::policy_evict_first();
        while (running) {
            if (!dev::wait_counter(&ctrl->sw_seq, sweep + 1, abort_flag, halt)) break;
            const int kind = ctrl->sw_kind[sweep & 3], count = ctrl->sw_count[sweep & 3];
            if (kind == kSweepExit) break;
            const T* panels = (kind == kSweepActive) ? a.panels_active : a.panels_screen;
            const int nb = (count + B - 1) / B;
            // ring order = consumption order of the data warps: D(0) | D(1) U(0) | D(2) U(1) | ... | U(nb-1); every group is
            // streamed in items of at most `ch` columns.  D items keep their lines in L2 (evict_last): the U items of the same
            // group re-read them one batch later (evict_first: last use).  U items wait for the group's proximal update.
            auto issue_item = [&](int col0, int ncols, uint64_t pol) -> bool {
                const int stage = gitem % a.n_stages;
                const uint32_t use = gitem / a.n_stages;
                if (!dev::mbar_wait(&empty_bar[stage], (use & 1) ^ 1, abort_flag, halt)) return false;
                T* xs = stages + (size_t)stage * a.stage_elems;
                const uint32_t col_bytes = (uint32_t)rows * sizeof(T);
                if (lane == 0) dev::mbar_arrive_expect_tx(&full_bar[stage], col_bytes * ncols);
                __syncwarp();
                if (lane < ncols)
                    dev::tma_bulk_g2s_hint(xs + (size_t)lane * a.rows_stride, a.X + (int64_t)(col0 + lane) * a.ld + r0, col_bytes, &full_bar[stage], pol);
                ++gitem;
                return true;
            };
#pragma unroll 1
            for (int b = -1; b < nb && running; ++b) {
                if (b + 1 < nb) {
                    const int p0 = (b + 1) * B, nbg = min(B, count - p0);
                    // ---- D items of batch b + 1
                    for (int k = 0; k < nbg && running; ++k) {
                        const int ss = (kind == kSweepActive) ? dev::ld_cg(a.active_set + p0 + k) : p0 + k;
                        const GroupMeta m = a.meta[ss];
                        for (int c0 = 0; c0 < m.gs && running; c0 += a.ch) running = issue_item(m.col + c0, min(a.ch, m.gs - c0), pol_keep);
                    }
                    if (!running) break;
                    // ---- L2 prefetch of the tiles of batch b + 2 (Configs::sweep_l2_prefetch): its dot phase can only start when batch b is
                    // completely solved, and with four ring stages a CTA keeps too few bytes in flight to stream it at the HBM rate; issued
                    // one batch ahead, the HBM reads overlap the proximal chain and the dot phase streams from L2
                    if (a.l2_prefetch && b + 2 < nb) {
                        const int q0 = (b + 2) * B, nq = min(B, count - q0);
                        const uint32_t col_bytes = (uint32_t)rows * sizeof(T);
                        for (int k = 0; k < nq; ++k) {
                            const int ss2 = (kind == kSweepActive) ? dev::ld_cg(a.active_set + q0 + k) : q0 + k;
                            const GroupMeta m2 = a.meta[ss2];
                            if (lane < m2.gs) dev::prefetch_l2_bulk(a.X + (int64_t)(m2.col + lane) * a.ld + r0, col_bytes);
                        }
                    }
                    // ---- panel + records of batch b + 1
                    const int slot = pitem & 1;
                    const uint32_t use = pitem >> 1;
                    if (!dev::mbar_wait(&pempty_bar[slot], (use & 1) ^ 1, abort_flag, halt)) { running = false; break; }
                    T* ps_ = pslots + (size_t)slot * a.pslot_elems;
                    const uint32_t q_bytes = (uint32_t)(Ccap * ldq) * sizeof(T);
                    int rec_elems = 0; int64_t rec_off = 0;
                    if (lane < nbg) {
                        const int ss = (kind == kSweepActive) ? dev::ld_cg(a.active_set + p0 + lane) : p0 + lane;
                        const GroupMeta mm = a.meta[ss];
                        rec_elems = mm.rec_elems; rec_off = mm.rec_off;
                        if (a.use_ext) {                                         // the extension sits right behind the base record
                            const int gsp = (mm.gs + 3) & ~3;
                            rec_off += rec_elems; rec_elems = 3 * gsp + 2 * mm.gs * gsp;
                        }
                    }
                    uint32_t tot = (uint32_t)rec_elems * sizeof(T);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                    if (lane == 0) dev::mbar_arrive_expect_tx(&pfull_bar[slot], q_bytes + tot);
                    __syncwarp();
                    if (lane == 0) dev::tma_bulk_g2s(ps_, panels + (int64_t)(b + 1) * Ccap * ldq, q_bytes, &pfull_bar[slot]);
                    if (lane < nbg) dev::tma_bulk_g2s(ps_ + Ccap * ldq + lane * a.rec_stride, a.grec + rec_off, (uint32_t)rec_elems * sizeof(T), &pfull_bar[slot]);
                    ++pitem;
                }
                if (b >= 0) {
                    // ---- U items of batch b.  Active-set sweeps: PREFETCHED without waiting for the proximal updates (the tile does not
                    // depend on them; almost every active group moves, and a tile that turns out to be unneeded is consumed and dropped by the
                    // data warps), so the L2 -> shared-memory latency of the update tiles and the start of the next batch's HBM stream are no
                    // longer exposed behind every proximal update.  Screen sweeps (most groups do not move): issued per group once it moved.
                    const int p0 = b * B, nbg = min(B, count - p0);
                    const int upar = ubatch & 1;
                    for (int k = 0; k < nbg && running; ++k) {
                        const int ss = (kind == kSweepActive) ? dev::ld_cg(a.active_set + p0 + k) : p0 + k;
                        const GroupMeta m = a.meta[ss];
                        bool fetch = true;
                        if (kind != kSweepActive || !a.u_prefetch) {
                            if (!dev::mbar_wait(&prox_bar[upar * kBatchMax + k], (uint32_t)(ubatch >> 1) & 1u, abort_flag, halt)) { running = false; break; }
                            fetch = *reinterpret_cast<volatile int*>(&ctrl->changed[upar][k]) != 0;
                        }
                        if (fetch)
                            for (int c0 = 0; c0 < m.gs && running; c0 += a.ch) running = issue_item(m.col + c0, min(a.ch, m.gs - c0), pol_done);
                    }
                    ++ubatch;
                }
            }
            ++sweep;
        }
        // never leave while bulk copies into this CTA's shared memory may still be in flight
        for (uint32_t g = (gitem > (uint32_t)a.n_stages ? gitem - a.n_stages : 0); g < gitem; ++g)
            dev::mbar_wait(&full_bar[g % a.n_stages], (g / a.n_stages) & 1, abort_flag, nullptr);
        for (uint32_t g = (pitem > 2 ? pitem - 2 : 0); g < pitem; ++g)
            dev::mbar_wait(&pfull_bar[g & 1], (g >> 1) & 1, abort_flag, nullptr);
        return;
    }

    // exchange geometry: level-1 groups of `fan` consecutive CTAs led by their first member
    const int fan = a.fan;
    const int my_group = cta / fan, n_groups = (ncta + fan - 1) / fan;
    const int grp_first = my_group * fan, grp_size = min(fan, ncta - grp_first);
    const bool is_leader = (cta == grp_first);

    // =====================================================================================================
    // exchange warps: two-level all-reduce of the batch's partial gradients.
    // Buffers are 4-deep in the batch epoch.  Why 4: a CTA publishes its level-1 lines of batch b+1 BEFORE it has read the
    // level-2 lines of batch b (look-ahead), but it cannot publish batch b+2 before (its U(b) needs its own prox(b)).  A
    // leader can therefore overwrite level-2 slot (b mod 4) with batch b+4 only after every CTA has published batch b+4's
    // level-1 lines, i.e. finished prox(b+2), i.e. read the level-2 lines of b+2 of ALL groups, which their leaders publish
    // only after every member published level-1 of b+2, which each member does only after reading level-2 of batch b.
    // =====================================================================================================
    if (role_ew) {
        const int et = (warp == 8 ? 0 : 32) + lane;
        int sweep = 0, bcount = 0;
        bool running = true;
        while (running) {
            if (!dev::wait_counter(&ctrl->sw_seq, sweep + 1, abort_flag, halt)) break;
            const int kind = ctrl->sw_kind[sweep & 3], count = ctrl->sw_count[sweep & 3];
            if (kind == kSweepExit) break;
            const int nb = (count + B - 1) / B;
            for (int b = 0; b < nb; ++b, ++bcount) {
                const int p0 = b * B, nbg = min(B, count - p0);
                int gsz = 0;
                if (lane < nbg) {
                    const int ss = (kind == kSweepActive) ? dev::ld_cg(a.active_set + p0 + lane) : p0 + lane;
                    gsz = a.meta[ss].gs;
                }
                int Cb = gsz;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) Cb += __shfl_xor_sync(0xffffffffu, Cb, o);
                const uint32_t e = epoch0 + (uint32_t)bcount;
                const int slot = (int)(e & (kBatchLLSlots - 1));
                bool ok = dev::mbar_wait(&dbar[bcount & 1], (uint32_t)(bcount >> 1) & 1u, abort_flag, halt);
                if (ok && et < Cb) {
                    const int c = et;
                    {
                        const double* w0 = wpart + (size_t)(bcount & 1) * kBatchDW * Ccap + c;
                        double s = 0;
#pragma unroll
                        for (int w = 0; w < kBatchDW; ++w) s += w0[(size_t)w * Ccap];
                        dev::ll_store(a.ll1 + ((size_t)(slot * kBatchColsMax + c) * a.ncta_pad + cta), s, e);
                    }
                    if (is_leader) {
                        const dev::LLLine* base = a.ll1 + ((size_t)(slot * kBatchColsMax + c) * a.ncta_pad + grp_first);
                        double s = 0;
#pragma unroll 1
                        for (int m0 = 0; m0 < grp_size && ok; m0 += 8) {          // 8 polls in flight, summed in member order
                            double v[8]; bool got[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) { v[u] = 0; got[u] = (m0 + u >= grp_size); }
                            dev::SpinGuard guard;
                            while (ok) {
                                bool all = true;
#pragma unroll
                                for (int u = 0; u < 8; ++u) if (!got[u]) { got[u] = dev::ll_try_load(base + m0 + u, e, v[u]); all &= got[u]; }
                                if (all) break;
                                if (guard.give_up(abort_flag, halt)) ok = false;
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) if (m0 + u < grp_size) s += v[u];
                        }
                        if (ok) dev::ll_store(a.ll2 + ((size_t)(slot * kBatchColsMax + c) * 32 + my_group), s, e);
                    }
                    const dev::LLLine* base2 = a.ll2 + ((size_t)(slot * kBatchColsMax + c) * 32);
                    const bool multi_gpu = a.world > 1;
                    double s = 0;
                    if (!multi_gpu || cta == 0) {                // level 2: every CTA (one GPU) / the GPU's leader CTA (several GPUs)
#pragma unroll 1
                        for (int g0 = 0; g0 < n_groups && ok; g0 += 8) {
                            double v[8]; bool got[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) { v[u] = 0; got[u] = (g0 + u >= n_groups); }
                            dev::SpinGuard guard;
                            while (ok) {
                                bool all = true;
#pragma unroll
                                for (int u = 0; u < 8; ++u) if (!got[u]) { got[u] = dev::ll_try_load(base2 + g0 + u, e, v[u]); all &= got[u]; }
                                if (all) break;
                                if (guard.give_up(abort_flag, halt)) ok = false;
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) if (g0 + u < n_groups) s += v[u];
                        }
                    }
                    if (multi_gpu) {
                        // level 3 over NVLink: the leader CTA stores this GPU's total into EVERY rank's line buffer (peer-to-peer
                        // stores through NVSwitch); every CTA of every GPU then reads its own GPU's `world` lines, rank order
                        if (cta == 0 && ok)
                            for (int r = 0; r < a.world; ++r)
                                dev::ll_store_sys(a.ll3_peer[r] + ((size_t)(slot * kBatchColsMax + c) * kMaxRanksDev + a.rank), s, e);
                        const dev::LLLine* base3 = a.ll3_peer[a.rank] + (size_t)(slot * kBatchColsMax + c) * kMaxRanksDev;
                        double v[kMaxRanksDev]; bool got[kMaxRanksDev];
#pragma unroll
                        for (int u = 0; u < kMaxRanksDev; ++u) { v[u] = 0; got[u] = (u >= a.world); }
                        dev::SpinGuard guard;
                        while (ok) {
                            bool all = true;
#pragma unroll
                            for (int u = 0; u < kMaxRanksDev; ++u) if (!got[u]) { got[u] = dev::ll_try_load_sys(base3 + u, e, v[u]); all &= got[u]; }
                            if (all) break;
                            if (guard.give_up(abort_flag, halt)) ok = false;
                        }
                        s = 0;
#pragma unroll
                        for (int u = 0; u < kMaxRanksDev; ++u) if (u < a.world) s += v[u];
                    }
                    gstale[(bcount & 1) * Ccap + c] = s;
                }
                if (!ok) ctrl->halt = 1;
                dev::named_bar_sync(2, 64);
                if (*halt) { running = false; break; }
                if (et == 0) dev::mbar_arrive(&gready_bar[bcount & 1]);
            }
            ++sweep;
        }
        return;
    }

    // =====================================================================================================
    // data warps
    // =====================================================================================================
    if (!role_cw) {
        const int dw = (warp >> 2) * 3 + (warp & 3) - 1;         // 0 .. 11
        const int dt = dw * 32 + lane;                            // 0 .. 383
        constexpr int CB = 16;
        uint32_t gitem = 0;
        int sweep = 0, bcount = 0, ucount = 0;                    // batches whose D phase ran / groups whose U phase ran
        int ubatch = 0;                                            // batches whose U phase ran (parity of del / changed / prox_bar)
        bool running = true;
        // phase profile (thread 0 of the first data warp of CTA 0): 8 wait-full, 9 dot, 10 barrier+publish, 11 wait-prox, 12 update, 13 sweep wait
        const bool prof = PROF && (a.stats != nullptr) && cta == 0 && dt == 0;
        long long pt[6] = {0, 0, 0, 0, 0, 0}; long long tc = prof ? clock64() : 0;
#define ABB_TICK(k) do { if (PROF && prof) { const long long t_ = clock64(); pt[PROF == 1 ? (((k) == 0 || (k) == 3 || (k) == 5) ? 0 : 1) : (k)] += t_ - tc; tc = t_; } } while (0)

        // One ring item = up to `ch` columns of a group (rows of this CTA) in a stage.
        // dot_item<CBW>: partial gradients of the item's columns against the current residual tile -> wp[0 .. ncols)
        auto dot_item = [&](auto cbw_tag, const T* xs, int ncols, double* wp) {
            constexpr int CBW = decltype(cbw_tag)::value;          // 8 or 16 accumulators
            const int64_t cs = a.rows_stride;
#pragma unroll 1
            for (int c0 = 0; c0 < ncols; c0 += CBW) {
                T acc[CBW];
#pragma unroll
                for (int cc = 0; cc < CBW; ++cc) acc[cc] = 0;
#pragma unroll 1
                for (int v = dt; v < rows / VN; v += kBatchNDT) {
                    T rv[VN], wv[VN], wr[VN];
                    vec_load<T>(sr + (size_t)v * VN, rv);
                    vec_load<T>(sw + (size_t)v * VN, wv);
#pragma unroll
                    for (int q = 0; q < VN; ++q) wr[q] = wv[q] * rv[q];
#pragma unroll
                    for (int cc = 0; cc < CBW; ++cc) {
                        if (c0 + cc < ncols) {
                            T xv[VN];
                            vec_load<T>(xs + (int64_t)(c0 + cc) * cs + (size_t)v * VN, xv);
#pragma unroll
                            for (int q = 0; q < VN; ++q) acc[cc] += xv[q] * wr[q];
                        }
                    }
                }
                if (CBW == 16) {
                    T a16[16];
#pragma unroll
                    for (int cc = 0; cc < 16; ++cc) a16[cc] = acc[cc % CBW];
                    const T tot = warp_reduce16<T>(a16, lane);     // lane holds the warp total of column c0 + ((lane >> 1) & 15)
                    const int col = c0 + ((lane >> 1) & 15);
                    if ((lane & 1) == 0 && col < ncols) wp[col] = (double)tot;
                } else {
                    // transposed reduction of 8 accumulators: 3 halving steps + 2 plain steps; lane holds column (lane >> 2) & 7
                    T v8[8];
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) v8[cc] = acc[cc % CBW];
                    {
                        const bool up = (lane & 16) != 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const T send = up ? v8[i] : v8[i + 4], keep = up ? v8[i + 4] : v8[i]; v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
                    }
                    {
                        const bool up = (lane & 8) != 0;
#pragma unroll
                        for (int i = 0; i < 2; ++i) { const T send = up ? v8[i] : v8[i + 2], keep = up ? v8[i + 2] : v8[i]; v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
                    }
                    {
                        const bool up = (lane & 4) != 0;
                        const T send = up ? v8[0] : v8[1], keep = up ? v8[1] : v8[0];
                        v8[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 2);
                    v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 1);
                    const int col = c0 + ((lane >> 2) & 7);
                    if ((lane & 3) == 0 && col < ncols) wp[col] = (double)v8[0];
                }
            }
        };
        // axpy_item: r += X_item del
        auto axpy_item = [&](const T* xs, int ncols, const T* dl) {
            const int64_t cs = a.rows_stride;
            constexpr int CBU = 8;
#pragma unroll 1
            for (int c0 = 0; c0 < ncols; c0 += CBU) {
                T d[CBU];
#pragma unroll
                for (int cc = 0; cc < CBU; ++cc) d[cc] = (c0 + cc < ncols) ? dl[c0 + cc] : T(0);
#pragma unroll 1
                for (int v = dt; v < rows / VN; v += kBatchNDT) {
                    T rv[VN];
                    vec_load<T>(sr + (size_t)v * VN, rv);
#pragma unroll
                    for (int cc = 0; cc < CBU; ++cc) {
                        if (c0 + cc < ncols) {
                            T xv[VN];
                            vec_load<T>(xs + (int64_t)(c0 + cc) * cs + (size_t)v * VN, xv);
#pragma unroll
                            for (int q = 0; q < VN; ++q) rv[q] += xv[q] * d[cc];
                        }
                    }
                    vec_store<T>(sr + (size_t)v * VN, rv);
                }
            }
        };
        // waits for the next ring item, returns its stage
        auto next_item = [&](const T*& xs, int& stage) -> bool {
            stage = (int)(gitem % a.n_stages);
            const uint32_t use = gitem / a.n_stages;
            if (!dev::mbar_wait(&full_bar[stage], use & 1, abort_flag, halt)) return false;
            xs = stages + (size_t)stage * a.stage_elems;
            ++gitem;
            return true;
        };
        auto release_item = [&](int stage) { __syncwarp(); if (lane == 0) dev::mbar_arrive(&empty_bar[stage]); };

        // D(b): partial gradients of the batch against the current residual tile
        auto dphase = [&](int kind, int count, int b) -> bool {
            const int p0 = b * B, nbg = min(B, count - p0);
            double* wp = wpart + (size_t)(bcount & 1) * kBatchDW * Ccap + (size_t)dw * Ccap;
            int gs_l = 0;                                          // lane k holds the size of group k (one load round per batch)
            if (lane < nbg) gs_l = a.meta[(kind == kSweepActive) ? dev::ld_cg(a.active_set + p0 + lane) : p0 + lane].gs;
            int off = 0;
            for (int k = 0; k < nbg; ++k) {
                const int gs = __shfl_sync(0xffffffffu, gs_l, k);
                for (int c0 = 0; c0 < gs; c0 += a.ch) {
                    const T* xs; int stage;
                    if (!next_item(xs, stage)) return false;
                    ABB_TICK(0);
                    const int ncols = min(a.ch, gs - c0);
                    if (a.ch <= 8) dot_item(std::integral_constant<int, 8>{}, xs, ncols, wp + off + c0);
                    else dot_item(std::integral_constant<int, 16>{}, xs, ncols, wp + off + c0);
                    release_item(stage);
                    ABB_TICK(1);
                }
                off += gs;
            }
            __syncwarp();
            if (lane == 0) dev::mbar_arrive(&dbar[bcount & 1]);       // release: the exchange warps sum the partials and publish
            ++bcount;
            ABB_TICK(2);
            return true;
        };
        // U(b): r += X_k del_k for the groups of the batch that moved (tiles re-streamed through the ring, normally from L2)
        auto uphase = [&](int kind, int count, int b, int upar) -> bool {
            const int p0 = b * B, nbg = min(B, count - p0);
            int gs_l = 0;
            if (lane < nbg) gs_l = a.meta[(kind == kSweepActive) ? dev::ld_cg(a.active_set + p0 + lane) : p0 + lane].gs;
            int off = 0;
            for (int k = 0; k < nbg; ++k, ++ucount) {
                const int gs = __shfl_sync(0xffffffffu, gs_l, k);
                if (!dev::mbar_wait(&prox_bar[upar * kBatchMax + k], (uint32_t)(ubatch >> 1) & 1u, abort_flag, halt)) return false;
                ABB_TICK(3);
                const bool moved = ctrl->changed[upar][k] != 0;
                if (moved || (kind == kSweepActive && a.u_prefetch)) {     // (prefetched tiles of unmoved groups are consumed and dropped)
                    const T* dl = del + (size_t)upar * Ccap + off;
                    for (int c0 = 0; c0 < gs; c0 += a.ch) {
                        const T* xs; int stage;
                        if (!next_item(xs, stage)) return false;
                        ABB_TICK(0);
                        if (moved) axpy_item(xs, min(a.ch, gs - c0), dl + c0);
                        release_item(stage);
                        ABB_TICK(4);
                    }
                }
                off += gs;
            }
            return true;
        };

        while (running) {
            if (!dev::wait_counter(&ctrl->sw_seq, sweep + 1, abort_flag, halt)) { running = false; break; }
            ABB_TICK(5);
            const int kind = ctrl->sw_kind[sweep & 3], count = ctrl->sw_count[sweep & 3];
            if (kind == kSweepExit) break;
            const int nb = (count + B - 1) / B;
#pragma unroll 1
            for (int b = -1; b < nb; ++b) {                        // rotated: D(b+1) then U(b), one call site each
                if (b + 1 < nb && !dphase(kind, count, b + 1)) { running = false; break; }
                if (b >= 0) {
                    if (!uphase(kind, count, b, ubatch & 1)) { running = false; break; }
                    ++ubatch;
                }
            }
            ++sweep;
        }
        if (PROF && prof) for (int k = 0; k < 6; ++k) a.stats[8 + k] += pt[k];
#undef ABB_TICK
        // write the residual tile back (thread-private rows: no barrier needed; on errors the host discards it)
        {
            T* gr = a.resid + r0;
            for (int v = dt; v < rows / VN; v += kBatchNDT) {
                T t[VN];
                vec_load<T>(sr + (size_t)v * VN, t); vec_store<T>(gr + (size_t)v * VN, t);
            }
        }
        return;
    }

    // =====================================================================================================
    // control warp
    // =====================================================================================================
    {
        const P l1 = (P)(a.lmda * a.alpha), l2 = (P)(a.lmda * (1.0 - a.alpha));
        ProxState ps;
        ps.rsq = a.sc->rsq; ps.resid_sum = a.sc->resid_sum; ps.cm = 0; ps.A = a.sc->active_set_size; ps.error = 0; ps.newton_iters_max = 0;
        long long iters = a.sc->iters, n_updates = a.sc->n_group_updates, n_cols = a.sc->n_col_updates;
        int phase = a.start_phase;
        int sweep = 0, bcount = 0, gcount = 0, pitem = 0;
        int final_error = 0, status = kBatchDone;
        const bool prof = PROF && (a.stats != nullptr) && cta == 0 && lane == 0;
        long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tc = prof ? clock64() : 0;
        long long ppx[8] = {0, 0, 0, 0, 0, 0, 0, 0};       // prox sub-phases: 0 pre, 1 gradient rotate, 2 root find, 3 sums, 4 rotate back, 5 newton iterations, 6 unchanged
#define ABB_TICK(k) do { if (PROF == 2 || (PROF == 1 && (k) != 2)) { if (prof) { const long long t_ = clock64(); pt[PROF == 1 ? (((k) == 0 || (k) == 1) ? 0 : 1) : (k)] += t_ - tc; tc = t_; } } } while (0)

        // issues the loads of a batch's metas: lane k < nbg holds group k (consumed later: no stall here)
        auto load_batch = [&](int kind, int count, int b, GroupMeta& m, int& ss) {
            const int p0 = b * B, nbg = max(0, min(B, count - p0));
            m = GroupMeta{}; ss = 0;
            if (lane < nbg) {
                ss = (kind == kSweepActive) ? dev::ld_cg(a.active_set + p0 + lane) : p0 + lane;
                m = a.meta[ss];
                // screen sweeps: fold "already active" into the sign of ss (a group's flag only changes when it is itself processed)
                if (kind == kSweepScreen && my_active[ss]) ss = -ss - 1;
            }
        };
        auto batch_cols = [&](const GroupMeta& m) -> int {
            int C = m.gs;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) C += __shfl_xor_sync(0xffffffffu, C, o);
            return C;
        };
        // issues the loads of a batch's current coefficients into registers: column t = lane + 32 j  ->  pa[j]
        auto fetch_aold = [&](const GroupMeta& m, int nbg, P (&pa)[2], P (&pr)[2]) {
            int adr[2] = {-1, -1};
            int off = 0;
            for (int k = 0; k < nbg; ++k) {
                const int gs = __shfl_sync(0xffffffffu, m.gs, k), begin = __shfl_sync(0xffffffffu, m.begin, k);
#pragma unroll
                for (int j = 0; j < 2; ++j) { const int t = lane + 32 * j; if (t >= off && t < off + gs) adr[j] = begin + t - off; }
                off += gs;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) { pa[j] = (adr[j] >= 0) ? (P)my_beta[adr[j]] : P(0); pr[j] = (adr[j] >= 0) ? (P)my_brot[adr[j]] : P(0); }
        };
        auto store_aold = [&](const P (&pa)[2], const P (&pr)[2], int par) {
#pragma unroll
            for (int j = 0; j < 2; ++j) { const int t = lane + 32 * j; if (t < Ccap) { aold[par * Ccap + t] = pa[j]; arot[par * Ccap + t] = pr[j]; } }
        };

        while (true) {
            const int count = (phase == kSweepActive) ? ps.A : a.S;
            ABB_TICK(5);
            if (lane == 0) {
                ctrl->sw_kind[sweep & 3] = phase; ctrl->sw_count[sweep & 3] = count;
                dev::st_release_cta(&ctrl->sw_seq, sweep + 1);
            }
            ++iters;
            ps.cm = 0;
            const int kind = phase;
            const int nb = (count + B - 1) / B;
            GroupMeta m_cur, m_nxt, m_nn; int ss_cur = 0, ss_nxt = 0, ss_nn = 0;
            P ga[4] = {0, 0, 0, 0};                         // batch gradient + corrections in panel-column space (see below)
            int C_cur = 0, C_nxt = 0;
            if (nb > 0) {
                load_batch(kind, count, 0, m_cur, ss_cur);
                load_batch(kind, count, 1, m_nxt, ss_nxt);
                C_cur = batch_cols(m_cur); C_nxt = batch_cols(m_nxt);
                P pa[2], pr[2];
                fetch_aold(m_cur, min(B, count), pa, pr);
                store_aold(pa, pr, bcount & 1);
                __syncwarp();
            }
            for (int b = 0; b < nb; ++b, ++bcount, ++pitem) {
                const int par = bcount & 1;
                const int nbg = min(B, count - b * B);
                load_batch(kind, count, b + 2, m_nn, ss_nn);                      // in flight during this batch
                P pa_nxt[2], pr_nxt[2];
                fetch_aold(m_nxt, max(0, min(B, count - (b + 1) * B)), pa_nxt, pr_nxt);   // in flight during this batch (disjoint groups)
                const int pslot = pitem & 1;
                const T* Q = pslots + (size_t)pslot * a.pslot_elems;
                ABB_TICK(4);
                if (!dev::mbar_wait(&pfull_bar[pslot], (pitem >> 1) & 1, abort_flag, halt)) { final_error = kErrAbort; break; }
                ABB_TICK(0);
                if (!dev::mbar_wait(&gready_bar[par], (uint32_t)(bcount >> 1) & 1u, abort_flag, halt)) { final_error = kErrAbort; break; }
                ABB_TICK(1);
                if (PROF && prof) ++pt[7];
                __syncwarp();
                // the batch's all-reduced gradient moves into registers in panel-column space (lane l holds t = 4 l .. 4 l + 3):
                // t < Ccap are this batch's columns (gradient + the cross corrections collected during the previous batch, which sit
                // Ccap columns higher), t >= Ccap collects the corrections for the next batch
#pragma unroll
                for (int j = 0; j < 4; ++j) { const int t = 4 * lane + j; if (t >= Ccap && t < ldq) corr[t - Ccap] = ga[j]; }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = 4 * lane + j;
                    ga[j] = (t < C_cur) ? (P)gstale[par * Ccap + t] + corr[t] : P(0);
                }
                __syncwarp();
                int gs_max_b = m_cur.gs;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) gs_max_b = max(gs_max_b, __shfl_xor_sync(0xffffffffu, gs_max_b, o));

                auto run_batch = [&](auto gsp_tag) {
                    constexpr int GSP = decltype(gsp_tag)::value;           // GSP == 32: generic shared-memory prox (12 < gs <= 32)
                    constexpr int GL = (GSP <= 16) ? GSP : 2;
                    LanePre<P, GL> pre;
                    int off = 0, gs = 0;
                    // rotated loop (k = -1 only issues the static loads of group 0): one call site per phase
#pragma unroll 1
                    for (int k = -1; k < nbg; ++k) {
                        int changed = 0, ss = 0; bool was_active = false; P del_c = 0;
                        T* dl = del + (size_t)par * Ccap + off;
                        if (k >= 0) {
                            gs = __shfl_sync(0xffffffffu, m_cur.gs, k);
                            const int begin = __shfl_sync(0xffffffffu, m_cur.begin, k);
                            const int ss_enc = __shfl_sync(0xffffffffu, ss_cur, k);
                            ss = ss_enc < 0 ? -ss_enc - 1 : ss_enc;
                            was_active = ss_enc < 0;
                            const P pk = (P)__shfl_sync(0xffffffffu, m_cur.pen, k);
                            const T* rec = Q + Ccap * ldq + k * a.rec_stride;
                            P* ao = aold + par * Ccap + off;
                            const int idx = off + ((lane & 15) < gs ? (lane & 15) : 0);      // panel column of this lane's element
                            P g_in;
                            {
                                const P s0 = __shfl_sync(0xffffffffu, ga[0], idx >> 2), s1 = __shfl_sync(0xffffffffu, ga[1], idx >> 2);
                                const P s2 = __shfl_sync(0xffffffffu, ga[2], idx >> 2), s3 = __shfl_sync(0xffffffffu, ga[3], idx >> 2);
                                g_in = (idx & 2) ? ((idx & 1) ? s3 : s2) : ((idx & 1) ? s1 : s0);
                            }
                            if (gs == 1) {                                       // solver_gaussian_pin_naive.hpp:75-108
                                const P ak_old = ao[0];
                                const P A_kk = (P)rec[0], xm = (P)rec[a.use_ext ? 4 : 1];
                                P gk = __shfl_sync(0xffffffffu, g_in, 0) - xm * (P)ps.resid_sum * (P)a.intercept + ak_old * A_kk;
                                const P vv = fabs(gk) - l1 * pk;                 // update_coordinate, pin_base.hpp:181-195
                                P ak = (vv > P(0)) ? copysign(vv, gk) / (A_kk + l2 * pk) : P(0);
                                ak = (P)(T)ak;
                                gk -= ak_old * A_kk;
                                if (ak != ak_old) {
                                    const P dd = ak - ak_old;
                                    ps.cm = fmax(ps.cm, (double)(A_kk * dd * dd));
                                    ps.rsq += (double)(dd * (2 * gk - dd * A_kk));
                                    ps.resid_sum -= (double)(xm * dd);
                                    if (lane == 0) { my_beta[begin] = (T)ak; my_brot[begin] = (T)ak; dl[0] = (T)(-dd); }
                                    del_c = ((lane & 15) == 0) ? -dd : P(0);
                                    changed = 1;
                                }
                            } else if (GSP <= 16) {
                                changed = prox_post<T, P, GL>(pre, gs, g_in, l1 * pk, l2 * pk, (P)a.newton_tol, a.newton_max_iters, (P)a.dbeta_tol,
                                                              a.intercept, ps, lane, p_gk, my_beta + begin, my_brot + begin, dl, del_c, (PROF == 2 && prof) ? ppx : nullptr);
                            } else {                                             // 12 < gs <= 32: shared-memory reductions
                                if (lane < 16 && lane < gs) gcur[lane] = (double)g_in;
                                if (gs > 16) {                                   // columns 16 .. gs-1 of a large group
                                    const int idx2 = off + 16 + ((lane & 15) < gs - 16 ? (lane & 15) : 0);
                                    const P s0 = __shfl_sync(0xffffffffu, ga[0], idx2 >> 2), s1 = __shfl_sync(0xffffffffu, ga[1], idx2 >> 2);
                                    const P s2 = __shfl_sync(0xffffffffu, ga[2], idx2 >> 2), s3 = __shfl_sync(0xffffffffu, ga[3], idx2 >> 2);
                                    if (lane < 16 && 16 + lane < gs) gcur[16 + lane] = (double)((idx2 & 2) ? ((idx2 & 1) ? s3 : s2) : ((idx2 & 1) ? s1 : s0));
                                }
                                __syncwarp();
                                const ProxCtx<T, P> px{ao, nullptr, p_gk, nullptr, nullptr, p_at, p_scr, dl, gcur, my_beta};
                                const ProxPre<P> pre2 = prox_small_pre<T, P>(rec, gs, ao, lane);
                                changed = prox_small_post<T, P>(px, pre2, rec, gs, begin, l1 * pk, l2 * pk, (P)a.newton_tol, a.newton_max_iters,
                                                                (P)a.dbeta_tol, a.intercept, ps, lane, nullptr);
                            }
                            __syncwarp();
                            if (lane == 0) {
                                ctrl->changed[par][k] = changed;
                                dev::mbar_arrive(&prox_bar[par * kBatchMax + k]); // release: del / changed are visible to the data warps
                            }
                            ABB_TICK(2);
                            if (PROF && prof) ++pt[6];
                        }
                        // static loads of the next group, issued behind this group's corrections
                        long long tq = (PROF == 2 && prof) ? clock64() : 0;
                        if (GSP <= 16 && k + 1 < nbg) {
                            const int gs1 = __shfl_sync(0xffffffffu, m_cur.gs, k + 1);
                            if (gs1 > 1) prox_pre<T, P, GL>(pre, Q + Ccap * ldq + (k + 1) * a.rec_stride, gs1, aold + par * Ccap + off + gs, arot + par * Ccap + off + gs, lane);
                        }
                        if (PROF == 2 && prof) { const long long t_ = clock64(); ppx[0] += t_ - tq; tq = t_; }
                        if (changed) {
                            if (GSP <= 16) apply_corrections<T, P, GL>(Q, ldq, off, gs, dl, del_c, ga, lane, p_gk);
                            else apply_corrections<T, P, 0>(Q, ldq, off, gs, dl, del_c, ga, lane, p_gk);
                            if (PROF == 2 && prof) { const long long t_ = clock64(); ppx[7] += t_ - tq; tq = t_; }
                            if (kind == kSweepScreen && !was_active) {           // add_active_set (:294-304)
                                if (ps.A >= a.max_active_size) ps.error = kErrMaxActive;
                                else {
                                    if (lane == 0) { my_active[ss] = 1; a.active_set[ps.A] = ss; }
                                    ++ps.A;
                                }
                            }
                        }
                        if (k >= 0) { ++n_updates; n_cols += gs; ++gcount; }
                        off += gs;
                        ABB_TICK(3);
                        if (ps.error) break;
                    }
                };
                // (one lane-local instantiation only: several of them blow the register budget of the whole kernel)
                if (a.use_ext) run_batch(std::integral_constant<int, 12>{});
                else run_batch(std::integral_constant<int, 32>{});
                if (ps.error) { final_error = ps.error; break; }
                if (lane >= nbg && lane < kBatchMax) dev::mbar_arrive(&prox_bar[par * kBatchMax + lane]);   // keep every barrier at one phase per batch
                __syncwarp();
                if (lane == 0) dev::mbar_arrive(&pempty_bar[pslot]);
                store_aold(pa_nxt, pr_nxt, par ^ 1);
                m_cur = m_nxt; ss_cur = ss_nxt; C_cur = C_nxt;
                m_nxt = m_nn; ss_nxt = ss_nn; C_nxt = batch_cols(m_nxt);
                __syncwarp();
            }
            if (final_error) break;
            ABB_TICK(4);
            // ---- end of sweep (identical decision in every CTA)
            const bool conv = ps.cm < a.tol;
            int next;
            if (kind == kSweepActive) next = conv ? kSweepScreen : ((iters >= a.max_iters) ? -kErrMaxCds : kSweepActive);
            else if (conv) next = kSweepExit;
            else if (iters >= a.max_iters) next = -kErrMaxCds;
            else if (ps.A > a.n_active_panelled) { next = kSweepExit; status = kBatchNeedPanels; }
            else next = kSweepActive;
            ++sweep;
            if (next < 0) { final_error = -next; break; }
            if (next == kSweepExit) break;
            phase = next;
        }
        if (final_error) {
            if (lane == 0) { ctrl->error = final_error; ctrl->halt = 1; }
            if (final_error == kErrAbort && lane == 0) *abort_flag = 1;
        } else if (lane == 0) {
            ctrl->sw_kind[sweep & 3] = kSweepExit; ctrl->sw_count[sweep & 3] = 0;
            dev::st_release_cta(&ctrl->sw_seq, sweep + 1);
        }
        if (cta == 0 && lane == 0) {
            a.sc->rsq = ps.rsq; a.sc->resid_sum = ps.resid_sum; a.sc->active_set_size = ps.A;
            a.sc->iters = iters; a.sc->n_group_updates = n_updates; a.sc->n_col_updates = n_cols; a.sc->error = final_error;
            a.sc->newton_iters_max = ps.newton_iters_max; a.sc->pad = status;
            if (PROF && prof) for (int k = 0; k < 8; ++k) { a.stats[k] += pt[k]; if (PROF == 2) a.stats[16 + k] += ppx[k]; }
#undef ABB_TICK
            *a.epoch = epoch0 + (uint32_t)bcount + 8u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Gram panels.  One item = one (source group, target group) block:
//   out[a * ldq + b] = sum_i w_i X[i, col_s + a] X[i, col_t + b],   a < gs_s, b < gs_t
// grid = (n_items, n_row_blocks); two-phase deterministic reduction through `part`.
struct PairItem { int32_t col_s, gs_s, col_t, gs_t; int64_t out_off; int64_t part_off; };

// TS x TS register tile: TS = 5 takes a 10 x 10 block in 4 passes of 10 vector loads; TS = 10 (groups of <= 10 columns, float) in ONE
// pass of 20 -- every element of both groups is loaded once per block.
template <class T, int TS>
__global__ void __launch_bounds__(256)
pair_gram_kernel(const T* __restrict__ X, int64_t ld, int64_t n_pad, const PairItem* __restrict__ items, const T* __restrict__ w,
                 double* __restrict__ part, int64_t total, int rows_per_block)
{
    constexpr int VN = VecT<T>::N;
    __shared__ double s_red[8][TS * TS + 1];
    const PairItem it = items[blockIdx.x];
    const int rb = blockIdx.y;
    const int64_t row0 = (int64_t)rb * rows_per_block;
    const int64_t row1 = min((long long)n_pad, (long long)(row0 + rows_per_block));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T* Xs = X + (int64_t)it.col_s * ld;
    const T* Xt = X + (int64_t)it.col_t * ld;
    double* out = part + (size_t)rb * total + it.part_off;
    for (int a0 = 0; a0 < it.gs_s; a0 += TS) {
        for (int b0 = 0; b0 < it.gs_t; b0 += TS) {
            T acc[TS][TS];
#pragma unroll
            for (int x = 0; x < TS; ++x)
#pragma unroll
                for (int y = 0; y < TS; ++y) acc[x][y] = 0;
            for (int64_t i = row0 + (int64_t)tid * VN; i < row1; i += 256 * VN) {
                T wv[VN], xa[TS][VN];
                vec_load<T>(w + i, wv);
#pragma unroll
                for (int x = 0; x < TS; ++x) {
                    if (a0 + x < it.gs_s) {
                        vec_load<T>(Xs + (int64_t)(a0 + x) * ld + i, xa[x]);
#pragma unroll
                        for (int q = 0; q < VN; ++q) xa[x][q] *= wv[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < VN; ++q) xa[x][q] = 0;
                    }
                }
                // target columns one at a time (the compiler keeps a few of these loads in flight): 10 x 10 accumulators + 10 source
                // vectors already fill most of the register file
#pragma unroll
                for (int y = 0; y < TS; ++y) {
                    T xb[VN];
                    if (b0 + y < it.gs_t) vec_load<T>(Xt + (int64_t)(b0 + y) * ld + i, xb);
                    else {
#pragma unroll
                        for (int q = 0; q < VN; ++q) xb[q] = 0;
                    }
#pragma unroll
                    for (int x = 0; x < TS; ++x)
#pragma unroll
                        for (int q = 0; q < VN; ++q) acc[x][y] += xa[x][q] * xb[q];
                }
            }
#pragma unroll
            for (int x = 0; x < TS; ++x)
#pragma unroll
                for (int y = 0; y < TS; ++y) {
                    const double s = dev::warp_sum((double)acc[x][y]);
                    if (lane == 0) s_red[warp][x * TS + y] = s;
                }
            __syncthreads();
            if (tid < TS * TS) {
                double s = 0;
                for (int wi = 0; wi < 8; ++wi) s += s_red[wi][tid];
                const int aa = a0 + tid / TS, bb = b0 + tid % TS;
                if (aa < it.gs_s && bb < it.gs_t) out[aa * it.gs_t + bb] = s;
            }
            __syncthreads();
        }
    }
}

template <class T>
__global__ void pair_gram_finalize_kernel(const PairItem* __restrict__ items, const double* __restrict__ part, int n_rb, int64_t total,
                                          T* __restrict__ Q, int ldq)
{
    const PairItem it = items[blockIdx.x];
    for (int e = threadIdx.x; e < it.gs_s * it.gs_t; e += blockDim.x) {
        double s = 0;
        for (int rb = 0; rb < n_rb; ++rb) s += part[(size_t)rb * total + it.part_off + e];
        Q[it.out_off + (int64_t)(e / it.gs_t) * ldq + (e % it.gs_t)] = (T)s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Whole Gram panels (fp32): one item = one panel = X_b^T W [X_b | X_{b+1}] for the <= 64 columns of batch b against the <= 128 columns of
// the window (batch b, then batch b + 1).  A CTA streams its row block in chunks of 64 rows: the window's columns are staged ONCE per
// chunk in shared memory (column-major, row stride 68: 16-byte reads with at most the unavoidable 2-way bank conflict), every thread owns
// a 4 x 8 tile of the 64 x 128 block (sources ty + 16 a, targets tx + 16 b) and reads 4 + 8 + 1 vectors per 128 FMAs.  Every column of
// the window is read from HBM once per panel instead of once per (source group, target group) block: pair_gram_kernel moved
// ~17 000 blocks x 16 MB per config-2 path.  Partial sums: fp32 inside a chunk, double across chunks and row blocks (deterministic).
// (PanelItem and kPanelOut live in gram_tc.cuh, shared with the tensor-core version of this kernel)
constexpr int kPanelRows = 64;                 // rows per staged chunk
constexpr int kPanelStride = kPanelRows + 4;   // shared-memory row stride of a column (floats)

__global__ void __launch_bounds__(256)
panel_gram_kernel(const float* __restrict__ X, int64_t ld, int64_t n_pad, const PanelItem* __restrict__ items, const float* __restrict__ w,
                  double* __restrict__ part, int n_panels, int rows_per_block)
{
    extern __shared__ __align__(16) float s_panel[];             // [128][kPanelStride] window tile, then w[kPanelRows]
    float* s_w = s_panel + 128 * kPanelStride;
    const PanelItem& it = items[blockIdx.x];
    const int ncol = it.ncol, n_src = it.n_src;
    const int rb = blockIdx.y;
    const int64_t row0 = (int64_t)rb * rows_per_block;
    const int64_t row1 = min((long long)n_pad, (long long)(row0 + rows_per_block));
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    double accd[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) accd[a][b] = 0;
    for (int e = tid; e < (128 - ncol) * kPanelStride; e += 256) s_panel[ncol * kPanelStride + e] = 0.f;      // unused window columns
    for (int64_t r = row0; r < row1; r += kPanelRows) {
        const int rows = (int)min((long long)kPanelRows, (long long)(row1 - r));      // multiple of 32
        __syncthreads();
        for (int e = tid; e < ncol * (kPanelRows / 4); e += 256) {
            const int u = e / (kPanelRows / 4), q4 = e - u * (kPanelRows / 4);
            float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * q4 < rows) x4 = *reinterpret_cast<const float4*>(X + (int64_t)it.cols[u] * ld + r + 4 * q4);
            *reinterpret_cast<float4*>(s_panel + u * kPanelStride + 4 * q4) = x4;
        }
        if (tid < kPanelRows / 4) {
            float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * tid < rows) w4 = *reinterpret_cast<const float4*>(w + r + 4 * tid);
            *reinterpret_cast<float4*>(s_w + 4 * tid) = w4;
        }
        __syncthreads();
        float acc[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 2
        for (int q4 = 0; q4 < kPanelRows / 4; ++q4) {
            const float4 w4 = *reinterpret_cast<const float4*>(s_w + 4 * q4);
            float4 xs[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                xs[a] = *reinterpret_cast<const float4*>(s_panel + (ty + 16 * a) * kPanelStride + 4 * q4);
                xs[a].x *= w4.x; xs[a].y *= w4.y; xs[a].z *= w4.z; xs[a].w *= w4.w;
            }
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const float4 xt = *reinterpret_cast<const float4*>(s_panel + (tx + 16 * b) * kPanelStride + 4 * q4);
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[a][b] += xs[a].x * xt.x + xs[a].y * xt.y + xs[a].z * xt.z + xs[a].w * xt.w;
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) accd[a][b] += (double)acc[a][b];
    }
    double* out = part + ((size_t)rb * n_panels + blockIdx.x) * kPanelOut;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int s_ = ty + 16 * a, u = tx + 16 * b;
            if (s_ < n_src && u < ncol) out[s_ * 128 + u] = accd[a][b];
        }
}

// out: tmp[panel * kPanelOut + s * 128 + u] = sum over row blocks
template <class P>
__global__ void panel_gram_sum_kernel(const PanelItem* __restrict__ items, const P* __restrict__ part, int n_rb, int n_panels, double* __restrict__ tmp)
{
    const PanelItem& it = items[blockIdx.x];
    for (int e = threadIdx.x; e < it.n_src * 128; e += blockDim.x) {
        if ((e & 127) >= it.ncol) continue;
        double s_ = 0;
        for (int rb = 0; rb < n_rb; ++rb) s_ += (double)part[((size_t)rb * n_panels + blockIdx.x) * kPanelOut + e];
        tmp[(size_t)blockIdx.x * kPanelOut + e] = s_;
    }
}
// scatter into the panel: window column u < n_src -> panel column u, otherwise Ccap + (u - n_src)
template <class T>
__global__ void panel_gram_scatter_kernel(const PanelItem* __restrict__ items, const double* __restrict__ tmp, T* __restrict__ Q, int ldq, int Ccap)
{
    const PanelItem& it = items[blockIdx.x];
    for (int e = threadIdx.x; e < it.n_src * 128; e += blockDim.x) {
        const int s_ = e >> 7, u = e & 127;
        if (u >= it.ncol) continue;
        const int t = (u < it.n_src) ? u : Ccap + (u - it.n_src);
        Q[it.q_off + (int64_t)s_ * ldq + t] = (T)tmp[(size_t)blockIdx.x * kPanelOut + e];
    }
}

} // namespace ab
