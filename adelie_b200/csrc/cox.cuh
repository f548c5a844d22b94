// adelie_b200/csrc/cox.cuh -- Cox partial likelihood on the device.
//
// Reference: GlmCoxPack / GlmCox, CORE/glm/glm_cox.ipp (scan helpers :19-226, pack :228-514, strata wrapper :516-750):
// Breslow / Efron ties, start / stop times, strata.  The reference walks the sorted start / stop sequences with sequential
// two-pointer scans per stratum.  Here everything that depends only on (start, stop, status, weights, strata) is resolved ONCE on
// the host into index tables, so that every evaluation is a fixed pipeline of coalesced passes over device-resident vectors:
//
//   T order = strata-major, stop-sorted;  S order = strata-major, start-sorted
//   z = w * exp(eta)                                                             (original order)
//   sufT / sufS = per-stratum suffix sums of z in T / S order                    (segmented scans)
//   tieZ        = per-tie-group sums of z over the events with w != 0            (segmented scan, read at the group's last slot)
//   risk_total[i] = sufT[tie_first[i]] - sufS[lbS[i]] - scale[i] * ind[i] * tieZ[tie_last[i]]           (glm_cox.ipp:119-134, :151-175)
//   v[i]  = status * wmean / (risk_total^pow + [status == 0 or wmean == 0])
//   P1    = per-stratum prefix sums of v in T order;  G1[i] = P1[tie_last[i]];  G2[j] = P1[ubT[j]] (S order)
//   G3[i] = ind[i] * (tie-group sum of v * scale * (pow == 2 ? 2 - scale : 1) * ind)
//   grad[o] = w status - (G1 - G3 - G2) z           hess[o] = w status - grad - (H1 - H3 - H2) z^2      (glm_cox.ipp:356-463)
//
// with  tie_first / tie_last = first / last T slot of i's tie group, lbS[i] = first S slot of the stratum with start >= stop_i,
// ubT[j] = last T slot of the stratum with stop <= start_j.  Segmented scans run in double (three passes: block scan, carries, fix-up).
// Cox needs a global sort: it is not row-sharded ("replicas only").
#pragma once
#include "glm.cuh"
#include <numeric>
#include <algorithm>
#include <limits>

namespace ab {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanBlock = kScanThreads * kScanItems;

template <bool MAXOP> __device__ __forceinline__ double scan_op(double a, double b) { return MAXOP ? fmax(a, b) : a + b; }
template <bool MAXOP> __device__ __forceinline__ double scan_identity() { return MAXOP ? -INFINITY : 0.0; }

// Pass 1: inclusive segmented scan inside blocks of 1024 elements.  `head[a]` marks the first element of a segment in scan
// direction (BACK: the scan runs from the end of the array, head = last element of a segment).  open[g] = 1 while no head has
// been met in the block up to and including position g (those elements still need the carry of the previous blocks).
template <bool BACK, bool MAXOP>
__global__ void __launch_bounds__(kScanThreads)
segscan_block_kernel(const double* __restrict__ in, const uint8_t* __restrict__ head, double* __restrict__ out, int64_t n,
                     double* __restrict__ blk_val, uint8_t* __restrict__ blk_flag, uint8_t* __restrict__ open)
{
    __shared__ double s_val[kScanThreads / 32]; __shared__ int s_flag[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t g0 = (int64_t)blockIdx.x * kScanBlock + (int64_t)tid * kScanItems;
    double v[kScanItems]; int f[kScanItems];
    double acc = scan_identity<MAXOP>(); int any = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t g = g0 + k;
        if (g < n) {
            const int64_t a = BACK ? n - 1 - g : g;
            const double x = in[a]; const int h = head[a];
            acc = h ? x : scan_op<MAXOP>(acc, x);
            any |= h;
            v[k] = acc; f[k] = any;
        } else { v[k] = acc; f[k] = any; }
    }
    // warp-level segmented scan of the per-thread totals
    double tv = acc; int tf = any;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double ov = __shfl_up_sync(0xffffffffu, tv, d); const int of = __shfl_up_sync(0xffffffffu, tf, d);
        if (lane >= d) { if (!tf) tv = scan_op<MAXOP>(ov, tv); tf |= of; }
    }
    if (lane == 31) { s_val[warp] = tv; s_flag[warp] = tf; }
    double ev = __shfl_up_sync(0xffffffffu, tv, 1); int ef = __shfl_up_sync(0xffffffffu, tf, 1);      // exclusive prefix inside the warp
    if (lane == 0) { ev = scan_identity<MAXOP>(); ef = 0; }
    __syncthreads();
    double cv = scan_identity<MAXOP>(); int cf = 0;                                                      // carry of the previous warps
    for (int w2 = 0; w2 < warp; ++w2) { if (s_flag[w2]) { cv = s_val[w2]; cf = 1; } else cv = scan_op<MAXOP>(cv, s_val[w2]); }
    if (!ef) { ev = scan_op<MAXOP>(cv, ev); ef = cf; }
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t g = g0 + k;
        if (g < n) {
            const int64_t a = BACK ? n - 1 - g : g;
            out[a] = f[k] ? v[k] : scan_op<MAXOP>(ev, v[k]);
            open[g] = (uint8_t)(!(f[k] | ef));
        }
    }
    if (tid == kScanThreads - 1) {
        double bv = s_val[0]; int bf = s_flag[0];
        for (int w2 = 1; w2 < kScanThreads / 32; ++w2) { if (s_flag[w2]) { bv = s_val[w2]; bf = 1; } else bv = scan_op<MAXOP>(bv, s_val[w2]); }
        blk_val[blockIdx.x] = bv; blk_flag[blockIdx.x] = (uint8_t)bf;
    }
}
// Pass 2: exclusive carries of the blocks.  One warp: lane l owns a contiguous chunk of blocks, the 32 chunk aggregates go through a
// warp-level segmented scan, then every lane replays its chunk from its carry-in (two short serial loops of nb / 32 instead of one of nb).
template <bool MAXOP>
__global__ void segscan_carry_kernel(const double* __restrict__ blk_val, const uint8_t* __restrict__ blk_flag, int nb, double* __restrict__ carry) {
    if (blockIdx.x || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const int per = (nb + 31) / 32, b0 = min(nb, lane * per), b1 = min(nb, b0 + per);
    double acc = scan_identity<MAXOP>(); int any = 0;
    for (int b = b0; b < b1; ++b) { if (blk_flag[b]) { acc = blk_val[b]; any = 1; } else acc = scan_op<MAXOP>(acc, blk_val[b]); }
    double tv = acc; int tf = any;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double ov = __shfl_up_sync(0xffffffffu, tv, d); const int of = __shfl_up_sync(0xffffffffu, tf, d);
        if (lane >= d) { if (!tf) tv = scan_op<MAXOP>(ov, tv); tf |= of; }
    }
    double c = __shfl_up_sync(0xffffffffu, tv, 1);
    if (lane == 0) c = scan_identity<MAXOP>();
    for (int b = b0; b < b1; ++b) { carry[b] = c; c = blk_flag[b] ? blk_val[b] : scan_op<MAXOP>(c, blk_val[b]); }
}
// Pass 3: elements in front of the first head of their block take the carry.
template <bool BACK, bool MAXOP>
__global__ void segscan_fix_kernel(double* __restrict__ out, const uint8_t* __restrict__ open, const double* __restrict__ carry, int64_t n) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n || !open[g]) return;
    const int64_t a = BACK ? n - 1 - g : g;
    out[a] = scan_op<MAXOP>(carry[g / kScanBlock], out[a]);
}

struct SegScan {
    DevBuf<double> blk_val, carry; DevBuf<uint8_t> blk_flag, open;
    void reserve(int64_t n) {
        const size_t nb = (size_t)((n + kScanBlock - 1) / kScanBlock);
        if (blk_val.n < nb) { blk_val.alloc(nb); carry.alloc(nb); blk_flag.alloc(nb); }
        if (open.n < (size_t)n) open.alloc((size_t)n);
    }
    template <bool BACK, bool MAXOP>
    void run(const double* in, const uint8_t* head, double* out, int64_t n, cudaStream_t st = 0) {
        if (n <= 0) return;
        reserve(n);
        const int nb = (int)((n + kScanBlock - 1) / kScanBlock);
        segscan_block_kernel<BACK, MAXOP><<<nb, kScanThreads, 0, st>>>(in, head, out, n, blk_val.p, blk_flag.p, open.p);
        if (nb > 1) {
            segscan_carry_kernel<MAXOP><<<1, 32, 0, st>>>(blk_val.p, blk_flag.p, nb, carry.p);
            segscan_fix_kernel<BACK, MAXOP><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, open.p, carry.p, n);
        }
        AB_CUDA(cudaGetLastError());
    }
};

template <class T>
struct GlmCox : Glm<T> {
    using B = Glm<T>;
    int64_t n = 0; bool efron = true;
    T loss_full_value = 0;
    // static tables (device)
    DevBuf<T> status;                                   // original order (B::w = weights, B::y = status as well)
    DevBuf<int32_t> permT, permS, posT, posS, tie_first, tie_last, lbS, ubT, strat_last;
    DevBuf<uint8_t> headF, headB, tieHead;              // stratum first / last (T and S orders share the boundaries), tie-group first
    DevBuf<T> wmean, scale, ind;                        // T order: mean event weight of the tie group, Efron scale, status * [w != 0]
    // scratch (double)
    DevBuf<double> z, a1, a2, a3, a4, a5, a6;
    SegScan scan;

    GlmCox(const T* h_start, const T* h_stop, const T* h_status, const int64_t* h_strata, const T* h_w, int64_t n_, bool efron_) : n(n_), efron(efron_) {
        B::name = "cox"; B::n = n_;
        if (DistContext::get().active()) throw core_error("the cox family needs a global sort and is not row-sharded: run it on one GPU (replicas only).");
        if (n_ >= (int64_t)1 << 31) throw core_error("cox: too many observations.");
        for (int64_t i = 0; i < n; ++i) if (h_strata[i] < 0) throw core_error("strata must be non-negative.");
        const int64_t np = pad_rows(n);
        B::y.alloc(np); B::w.alloc(np); status.alloc(np);
        B::y.upload(h_status, n); B::w.upload(h_w, n); status.upload(h_status, n);
        // ---- orders: strata-major (stable), then by stop / start (glm_cox.ipp:324-352, :537-553)
        std::vector<int32_t> oT(n), oS(n);
        std::iota(oT.begin(), oT.end(), 0); std::iota(oS.begin(), oS.end(), 0);
        std::stable_sort(oT.begin(), oT.end(), [&](int32_t a, int32_t b) { return h_strata[a] != h_strata[b] ? h_strata[a] < h_strata[b] : h_stop[a] < h_stop[b]; });
        std::stable_sort(oS.begin(), oS.end(), [&](int32_t a, int32_t b) { return h_strata[a] != h_strata[b] ? h_strata[a] < h_strata[b] : h_start[a] < h_start[b]; });
        std::vector<int32_t> pT(n), pS(n), tf(n), tl(n), lb(n), ub(n), sl(n);
        std::vector<uint8_t> hF(n, 0), hB(n, 0), tH(n, 0);
        std::vector<T> wm(n, 0), sc(n, 0), in(n, 0);
        for (int64_t i = 0; i < n; ++i) { pT[oT[i]] = (int32_t)i; pS[oS[i]] = (int32_t)i; }
        double lf = 0;
        for (int64_t b = 0; b < n;) {                   // one stratum [b, e)
            int64_t e = b;
            while (e < n && h_strata[oT[e]] == h_strata[oT[b]]) ++e;
            hF[b] = 1; hB[e - 1] = 1;
            for (int64_t i = b; i < e; ++i) sl[i] = (int32_t)(e - 1);
            for (int64_t i = b; i < e;) {               // one tie group [i, j)
                int64_t j = i;
                const T ti = h_stop[oT[i]];
                int size = 0; T wsum = 0;
                while (j < e && h_stop[oT[j]] == ti) {
                    const T indic = h_status[oT[j]] * T(h_w[oT[j]] != 0);
                    in[j] = indic;
                    sc[j] = efron ? T(size) * indic : T(0);          // rank among the events of the tie group (glm_cox.ipp:194-226)
                    size += (int)indic; wsum += h_w[oT[j]] * indic;
                    ++j;
                }
                tH[i] = 1;
                for (int64_t k = i; k < j; ++k) {
                    tf[k] = (int32_t)i; tl[k] = (int32_t)(j - 1);
                    if (efron && size > 1) sc[k] /= T(size);
                    // weights_size_to / weights_mean_to (:151-175, :336-347): both carry the factor status * [w != 0]
                    const T wsize = in[k] * T(size);
                    T wmean_k = in[k] * wsum;
                    if (h_status[oT[k]] != 0 && h_w[oT[k]] != 0) wmean_k /= wsize;
                    wm[k] = wmean_k;
                    const T most_neg = -std::numeric_limits<T>::max();                                   // loss_full (:507-514)
                    lf += (double)(wmean_k * h_status[oT[k]] * std::max((T)std::log(wsize * wmean_k * (T(1) - sc[k])), most_neg));
                }
                i = j;
            }
            // lbS: first S slot of the stratum with start >= stop_i;  ubT: last T slot of the stratum with stop <= start_j
            {
                int64_t q = b;
                for (int64_t i = b; i < e; ++i) { while (q < e && h_start[oS[q]] < h_stop[oT[i]]) ++q; lb[i] = (q < e) ? (int32_t)q : -1; }
                q = b;
                for (int64_t j2 = b; j2 < e; ++j2) { while (q < e && h_stop[oT[q]] <= h_start[oS[j2]]) ++q; ub[j2] = (q > b) ? (int32_t)(q - 1) : -1; }
            }
            b = e;
        }
        loss_full_value = (T)lf;
        auto upI = [&](DevBuf<int32_t>& d, const std::vector<int32_t>& h) { d.alloc(std::max<size_t>(1, h.size())); d.upload(h.data(), h.size()); };
        auto upB = [&](DevBuf<uint8_t>& d, const std::vector<uint8_t>& h) { d.alloc(std::max<size_t>(1, h.size())); d.upload(h.data(), h.size()); };
        auto upT = [&](DevBuf<T>& d, const std::vector<T>& h) { d.alloc(std::max<size_t>(1, h.size())); d.upload(h.data(), h.size()); };
        upI(permT, oT); upI(permS, oS); upI(posT, pT); upI(posS, pS); upI(tie_first, tf); upI(tie_last, tl); upI(lbS, lb); upI(ubT, ub); upI(strat_last, sl);
        upB(headF, hF); upB(headB, hB); upB(tieHead, tH);
        upT(wmean, wm); upT(scale, sc); upT(ind, in);
        for (DevBuf<double>* d : {&z, &a1, &a2, &a3, &a4, &a5, &a6}) d->alloc((size_t)np);
        scan.reserve(n);
        AB_CUDA(cudaStreamSynchronize(0));
    }

    // z = w * exp(eta - shift) in original order; risk_total (T order) -> a4.  shift_T: optional per-element shift in T order.
    void risk_total(const T* eta, const double* shift_T) {
        const T* w = B::w.p; double* zz = z.p; const int32_t* pt = permT.p; const int32_t* ps = permS.p; const int32_t* pT_ = posT.p;
        double* zT = a1.p; double* zS = a2.p; double* zM = a3.p; const T* indp = ind.p;
        B::mr.map(n, [=] __device__(int64_t o, double*) { zz[o] = (double)(w[o] * (T)exp((double)eta[o] - (shift_T ? shift_T[pT_[o]] : 0.0))); });
        B::mr.map(n, [=] __device__(int64_t i, double*) { const double zt = zz[pt[i]]; zT[i] = zt; zM[i] = zt * (double)indp[i]; zS[i] = zz[ps[i]]; });
        scan.template run<true, false>(a1.p, headB.p, a5.p, n);        // sufT
        scan.template run<true, false>(a2.p, headB.p, a6.p, n);        // sufS
        scan.template run<false, false>(a3.p, tieHead.p, a1.p, n);     // tie-group prefix of the masked z (group sum at tie_last)
        const double* sufT = a5.p; const double* sufS = a6.p; const double* tz = a1.p; double* rt = a4.p;
        const int32_t* tfp = tie_first.p; const int32_t* tlp = tie_last.p; const int32_t* lbp = lbS.p; const T* scp = scale.p;
        B::mr.map(n, [=] __device__(int64_t i, double*) {
            const double rs = sufT[tfp[i]] - (lbp[i] >= 0 ? sufS[lbp[i]] : 0.0);
            rt[i] = (double)((T)rs - scp[i] * (T)((double)indp[i] * tz[tlp[i]]));
        });
    }
    // out[o] = (S1[tie_last] - S3 - S2) for power 1 (gradient) or 2 (hessian); needs risk_total in a4 and z
    void scans(int power, double* out_orig) {
        const double* rt = a4.p; double* v = a1.p; double* vs = a2.p;
        const T* st = status.p; const int32_t* pt = permT.p; const T* wmp = wmean.p; const T* scp = scale.p; const T* indp = ind.p;
        B::mr.map(n, [=] __device__(int64_t i, double*) {
            const T s_i = st[pt[i]], wm_i = wmp[i];
            const T r = (T)rt[i];
            const T den = (power == 1 ? r : r * r) + T((s_i == 0) || (wm_i == 0));
            const T vi = s_i * wm_i / den;
            v[i] = (double)vi;
            vs[i] = (double)(vi * (power == 1 ? scp[i] : scp[i] * (T(2) - scp[i])) * indp[i]);
        });
        scan.template run<false, false>(a1.p, headF.p, a5.p, n);       // P1: per-stratum prefix of v
        scan.template run<false, false>(a2.p, tieHead.p, a6.p, n);     // tie-group prefix of v * scale'
        const double* P1 = a5.p; const double* P3 = a6.p;
        const int32_t* tlp = tie_last.p; const int32_t* ubp = ubT.p; const int32_t* pT_ = posT.p; const int32_t* pS_ = posS.p;
        B::mr.map(n, [=] __device__(int64_t o, double*) {
            const int32_t i = pT_[o], j = pS_[o];
            const T g1 = (T)P1[tlp[i]], g3 = indp[i] * (T)P3[tlp[i]], g2 = ubp[j] >= 0 ? (T)P1[ubp[j]] : T(0);
            out_orig[o] = (double)((g1 - g3) - g2);
        });
    }
    void gradient(const T* eta, T* grad) override {
        risk_total(eta, nullptr);
        scans(1, a3.p);
        const double* g = a3.p; const double* zz = z.p; const T* w = B::w.p; const T* st = status.p;
        B::mr.map(n, [=] __device__(int64_t o, double*) { grad[o] = w[o] * st[o] - (T)g[o] * (T)zz[o]; });
    }
    void hessian(const T* eta, const T* grad, T* hess) override {
        risk_total(eta, nullptr);
        scans(2, a3.p);
        const double* h = a3.p; const double* zz = z.p; const T* w = B::w.p; const T* st = status.p;
        B::mr.map(n, [=] __device__(int64_t o, double*) { const T zo = (T)zz[o]; hess[o] = w[o] * st[o] - grad[o] - (T)h[o] * zo * zo; });
    }
    T loss(const T* eta) override {
        if (n == 0) return 0;
        // per-stratum max of eta (glm_cox.ipp:474): segmented max scan in T order, read at the stratum's last slot
        const int32_t* pt = permT.p; double* eT = a1.p;
        B::mr.map(n, [=] __device__(int64_t i, double*) { eT[i] = (double)eta[pt[i]]; });
        scan.template run<false, true>(a1.p, headF.p, a2.p, n);
        double* mx = a3.p; const double* run_max = a2.p; const int32_t* slp = strat_last.p;
        B::mr.map(n, [=] __device__(int64_t i, double*) { mx[i] = run_max[slp[i]]; });
        // the shift lives in a3 while risk_total uses a1..a6: move it to the open slot of the scan-free buffer first
        DevBuf<double> shift((size_t)pad_rows(n));
        AB_CUDA(cudaMemcpyAsync(shift.p, a3.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, 0));
        risk_total(eta, shift.p);
        const double* rt = a4.p; const double* sh = shift.p; const T* st = status.p; const T* w = B::w.p; const T* wmp = wmean.p;
        const T neg_max = -std::numeric_limits<T>::max();
        double sums[2];
        B::mr.template run<2>(n, [=] __device__(int64_t i, double* acc) {
            const int32_t o = pt[i];
            acc[0] += (double)(st[o] * w[o] * (T)((double)eta[o] - sh[i]));
            const T r = max((T)rt[i], T(0));
            acc[1] += (double)(st[o] * wmp[i] * max((T)log(r), neg_max));
        }, sums);
        return (T)(-sums[0] + sums[1]);
    }
    T loss_full() override { return loss_full_value; }
    void inv_link(const T* eta, T* out) override { B::mr.map(n, [=] __device__(int64_t i, double*) { out[i] = exp(eta[i]); }); }
};

} // namespace ab
