// adelie_b200/csrc/solver_glm.cuh -- GLM IRLS outer loop on the device
// (CORE/solver/solver_glm_naive.hpp:165-232 update_loss_null, :241-459 fit).
#pragma once
#include "solver.cuh"

namespace ab {

// update_loss_null (solver_glm_naive.hpp:165-232): intercept-only IRLS on copies of eta / resid.
template <class T>
void PathState<T>::update_loss_null() {
    const int64_t nn = glm->n;                 // n (or n*K)
    if (K > 1) { update_loss_null_multi(); return; }
    if (!intercept) { loss_null = glm->loss(d_offsets.p); return; }
    const int64_t np = (int64_t)d_eta.n;
    DevBuf<T> eta(np), resid(np), eta_prev(np), resid_prev(np), hess(np), z(np);
    AB_CUDA(cudaMemcpyAsync(eta.p, d_eta.p, nn * sizeof(T), cudaMemcpyDeviceToDevice, 0));
    AB_CUDA(cudaMemcpyAsync(resid.p, d_resid.p, nn * sizeof(T), cudaMemcpyDeviceToDevice, 0));
    const T hmin = (T)Configs::hessian_min;
    size_t it = 0;
    while (1) {
        if (it >= irls_max_iters) throw solver_error("Maximum IRLS iterations reached.");
        glm->hessian(eta.p, resid.p, hess.p);
        glm->inv_hessian_gradient(eta.p, resid.p, hess.p, z.p);
        double sums[2];
        {
            T* h = hess.p; const T* zz = z.p; const T* e = eta.p; const T* off = d_offsets.p;
            glm->mr.template run<2>(nn, [=] __device__(int64_t i, double* acc) {
                const T hv = max(h[i], T(0)) + hmin * T(h[i] <= 0);
                h[i] = hv;
                acc[0] += (double)hv;
                acc[1] += (double)(hv * (zz[i] + e[i] - off[i]));
            }, sums);
        }
        const T b0 = (T)(sums[1] / sums[0]);
        std::swap(eta.p, eta_prev.p);
        { T* e = eta.p; const T* off = d_offsets.p; glm->mr.map(nn, [=] __device__(int64_t i, double*) { e[i] = b0 + off[i]; }); }
        std::swap(resid.p, resid_prev.p);
        glm->gradient(eta.p, resid.p);
        double conv;
        {
            const T* r = resid.p; const T* rp = resid_prev.p; const T* e = eta.p; const T* ep = eta_prev.p;
            glm->mr.template run<1>(nn, [=] __device__(int64_t i, double* acc) { acc[0] += (double)((r[i] - rp[i]) * (e[i] - ep[i])); }, &conv);
        }
        n_kernel_launches += 8;
        if (std::abs(conv) <= (double)irls_tol) { loss_null = glm->loss(eta.p); return; }
        ++it;
    }
}

// multi-response update_loss_null (solver_multiglm_naive.hpp:99-186): one unpenalised intercept per class
template <class T>
void PathState<T>::update_loss_null_multi() {
    const int64_t nn = glm->n;
    if (n_int == 0) { loss_null = glm->loss(d_offsets.p); return; }
    const int64_t np = (int64_t)d_eta.n;
    DevBuf<T> eta(np), resid(np), eta_prev(np), resid_prev(np), hess(np), z(np), sums(2 * K), b0(K);
    AB_CUDA(cudaMemcpyAsync(eta.p, d_eta.p, nn * sizeof(T), cudaMemcpyDeviceToDevice, 0));
    AB_CUDA(cudaMemcpyAsync(resid.p, d_resid.p, nn * sizeof(T), cudaMemcpyDeviceToDevice, 0));
    const T hmin = (T)Configs::hessian_min;
    const int KK = K;
    std::vector<T> hs(2 * K), hb0(K);
    size_t it = 0;
    while (1) {
        if (it >= irls_max_iters) throw solver_error("Maximum IRLS iterations reached.");
        glm->hessian(eta.p, resid.p, hess.p);
        glm->inv_hessian_gradient(eta.p, resid.p, hess.p, z.p);
        {
            T* h = hess.p; T* zz = z.p; const T* e = eta.p; const T* off = d_offsets.p;
            glm->mr.map(nn, [=] __device__(int64_t i, double*) {
                h[i] = max(h[i], T(0)) + hmin * T(h[i] <= 0);
                zz[i] += e[i] - off[i];
            });
        }
        // per-class sums  num_k = sum_i h_ik y_ik,  den_k = sum_i h_ik  (the 1 / sum(h) normalisation of :137-141 cancels)
        X->d_class_sums(K, z.p, hess.p, sums.p);
        X->d_class_sums(K, hess.p, nullptr, sums.p + K);
        DistContext::get().allreduce<T>(sums.p, 2 * K);
        sums.download(hs.data(), 2 * K);
        AB_CUDA(cudaStreamSynchronize(0));
        for (int k = 0; k < K; ++k) hb0[k] = hs[k] / hs[K + k];
        b0.upload(hb0.data(), K);
        std::swap(eta.p, eta_prev.p);
        { T* e = eta.p; const T* off = d_offsets.p; const T* bb = b0.p; glm->mr.map(nn, [=] __device__(int64_t i, double*) { e[i] = off[i] + bb[i % KK]; }); }
        std::swap(resid.p, resid_prev.p);
        glm->gradient(eta.p, resid.p);
        double conv;
        {
            const T* r = resid.p; const T* rp = resid_prev.p; const T* e = eta.p; const T* ep = eta_prev.p;
            glm->mr.template run<1>(nn, [=] __device__(int64_t i, double* acc) { acc[0] += (double)((r[i] - rp[i]) * (e[i] - ep[i])); }, &conv);
        }
        n_kernel_launches += 12;
        if (std::abs(conv) <= (double)irls_tol) { loss_null = glm->loss(eta.p); return; }
        ++it;
    }
}

// fit (solver_glm_naive.hpp:241-459)
template <class T>
PinResult PathState<T>::fit_glm(T lmda_) {
    const int64_t nn = glm->n;
    const T hmin = (T)Configs::hessian_min;
    PinResult last;
    double screen_time = 0, active_time = 0;
    size_t irls_it = 0;
    while (1) {
        if (irls_it >= irls_max_iters) throw solver_error("Maximum IRLS iterations reached.");
        std::vector<T> beta_prev = screen_beta; std::vector<int8_t> act_prev = screen_is_active;

        // ---- quadratic approximation: hess, z = irls_resid, irls_y and the four moments in one pass
        glm->hessian(d_eta.p, d_resid.p, d_hess.p);
        glm->inv_hessian_gradient(d_eta.p, d_resid.p, d_hess.p, d_irls_resid.p);
        double sums[4];
        {
            T* h = d_hess.p; const T* z = d_irls_resid.p; const T* e = d_eta.p; const T* off = d_offsets.p; T* iy = d_irls_y.p;
            glm->mr.template run<4>(nn, [=] __device__(int64_t i, double* acc) {
                const T hv = max(h[i], T(0)) + hmin * T(h[i] <= 0);
                h[i] = hv;
                const T yv = z[i] + e[i] - off[i];
                iy[i] = yv;
                acc[0] += (double)hv; acc[1] += (double)(hv * yv); acc[2] += (double)(hv * yv * yv); acc[3] += (double)(hv * z[i]);
            }, sums);
        }
        const T hess_sum = (T)sums[0];
        const T ym = (T)(sums[1] / sums[0]);
        const T yv = (T)(sums[2] / sums[0]) - (intercept ? ym * ym : T(0));
        const T shift = intercept ? (beta0 - ym) : T(0);
        T rs = (T)(sums[3] / sums[0]) + shift;
        {
            const T* h = d_hess.p; T* iw = d_irls_w.p; T* ir = d_irls_resid.p; const T hs = hess_sum;
            glm->mr.map(nn, [=] __device__(int64_t i, double*) { iw[i] = h[i] / hs; ir[i] += shift; });
        }
        T lmda_adj = lmda_ / hess_sum;
        if (std::isinf(lmda_adj)) {
            if (lmda_ == std::numeric_limits<T>::max()) lmda_adj = lmda_;
            else throw solver_error("IRLS lambda is unexpectedly inf. This likely indicates a bug in the code. Please report this!");
        }

        // ---- IRLS-weighted column means of every screen column (:361-372), one batched launch
        double t_means0 = now_s();
        const size_t S = screen_set.size();
        const size_t vs = S ? (screen_begins.back() + group_sizes[screen_set.back()]) : 0;
        std::vector<int32_t> cols(vs), pcols(vs);     // logical columns / physical columns (SNP: slots of the decoded-column cache)
        for (size_t i = 0; i < S; ++i) {
            const idx_t g = screen_set[i];
            const int32_t pc = (K == 1) ? X->phys_col(groups[g], (int)group_sizes[g]) : 0;
            for (idx_t c = 0; c < group_sizes[g]; ++c) { cols[screen_begins[i] + c] = (int32_t)(groups[g] + c); pcols[screen_begins[i] + c] = pc + (int32_t)c; }
        }
        std::vector<T> sx_means(vs);
        // the means come out of the Gram pass below when its kernel can produce them (one pass over the screen columns instead of two)
        const bool fuse_means = vs && screen_means_fusable(0, S);
        if (vs && K == 1 && !fuse_means) {           // multi-response runs with the state-level intercept off: the means are never used (left 0)
            d_cols.reserve_keep(vs); d_tmp.reserve_keep(vs);
            d_cols.upload(pcols.data(), vs);
            X->d_gemv_t(0, d_cols.p, (int)vs, X->d_ones(), d_irls_w.p, d_tmp.p);
            DistContext::get().allreduce<T>(d_tmp.p, (int64_t)vs);
            d_tmp.download(sx_means.data(), vs);
            AB_CUDA(cudaStreamSynchronize(0));
        }
        timers.slot("glm_means") += now_s() - t_means0;
        // ---- screen-derived quantities for ALL screen groups with the IRLS weights (:376-385)
        std::vector<GroupMeta>& meta = glm_meta; std::vector<T>& grec = glm_grec;         // persistent scratch (capacity kept across IRLS iterations)
        std::vector<T>& sXm = glm_sXm; std::vector<T>& sv = glm_sv; std::vector<std::vector<T>>& stv = glm_stv;
        meta.clear(); grec.clear(); sXm.clear(); sv.clear();
        {
            // xmean lookup by column: build a small map column -> value position
            std::vector<T>& gm = X_means;       // (p,) scratch: only screen columns are defined (GlmNaiveBufferPack::X_means)
            if ((idx_t)gm.size() != p) gm.assign(p, 0);
            for (size_t k = 0; k < vs; ++k) gm[cols[k]] = sx_means[k];
            gs_max_screen = 1; rec_max_screen = 4;
            compute_screen_records(0, S, d_irls_w.p, [&](idx_t c) { return gm[c]; }, sXm, sv, stv, meta, grec, fuse_means);
            if (fuse_means) for (size_t k = 0; k < vs; ++k) gm[cols[k]] = sXm[k];
        }
        { AB_TIME(timers, "rec_upload"); upload_screen_tables(meta, grec, true); AB_CUDA(cudaStreamSynchronize(0)); }
        n_kernel_launches += 6;

        // ---- weighted Gaussian pin solve on the working response (:389-423)
        T rsq_dummy = 0;
        PinResult pr;
        try {
            pr = run_pin(d_irls_resid.p, d_irls_w.p, lmda_adj, tol * (loss_null - loss_full) / hess_sum, ym, rsq_dummy, rs, false, &stv);
        } catch (...) {
            screen_beta.swap(beta_prev); screen_is_active.swap(act_prev);
            throw;
        }
        screen_time += pr.screen_time; active_time += pr.active_time;
        ++n_irls;
        beta0 = (T)pr.intercept;

        // ---- eta, resid, convergence (:437-449) in one fused pass + the family's gradient
        std::swap(d_eta.p, d_eta_prev.p);
        {
            T* e = d_eta.p; const T* iy = d_irls_y.p; const T* off = d_offsets.p; const T* ir = d_irls_resid.p;
            const T add = intercept ? (beta0 - ym) : T(0);
            glm->mr.map(nn, [=] __device__(int64_t i, double*) { e[i] = iy[i] + off[i] - ir[i] + add; });
        }
        std::swap(d_resid.p, d_glm_resid_prev.p);
        glm->gradient(d_eta.p, d_resid.p);
        double conv;
        {
            const T* r = d_resid.p; const T* rp = d_glm_resid_prev.p; const T* e = d_eta.p; const T* ep = d_eta_prev.p;
            glm->mr.template run<1>(nn, [=] __device__(int64_t i, double* acc) { acc[0] += (double)((r[i] - rp[i]) * (e[i] - ep[i])); }, &conv);
        }
        n_kernel_launches += 4;
        last = std::move(pr);
        if (std::abs(conv) <= (double)irls_tol) {
            last.screen_time = screen_time; last.active_time = active_time;
            return last;
        }
        ++irls_it;
    }
}

} // namespace ab
