// oracle/cox_oracle.hpp -- TEST INFRASTRUCTURE (see adelie_oracle.hpp header).
// CPU restatement of the reference Cox partial-likelihood GLM
// (CORE/glm/glm_cox.ipp): scan helpers :19-226, GlmCoxPack :228-514,
// GlmCox (strata wrapper) :516-750.
#pragma once
#include "adelie_oracle.hpp"

namespace orc {
namespace cox {

// out[i+1] = sum_k v[k] 1{s[k] <= t[i]}, out[0] = 0     (glm_cox.ipp:19-54)
template <class T, class VF>
inline void partial_sum_fwd(VF v, const T* s, idx_t n, const T* t, idx_t m, T* out) {
    out[0] = 0;
    if (m == 0) return;
    if (n == 0) { for (idx_t i = 0; i <= m; ++i) out[i] = 0; return; }
    idx_t k = 0, ib = 0, ie = 0;
    while (ib < m) {
        const T ti = t[ib];
        T cur = out[ib];
        for (; k < n && s[k] <= ti; ++k) cur += v(k);
        for (; ie < m && t[ie] == ti; ++ie) out[ie + 1] = cur;
        ib = ie;
        if (k >= n) break;
    }
    for (; ib < m; ++ib) out[ib + 1] = out[ib];
}

// out[i] = sum_k v[k] 1{s[k] >= t[i]}, out[m] = 0       (glm_cox.ipp:56-98)
template <class T, class VF>
inline void partial_sum_bwd(VF v, const T* s, idx_t n, const T* t, idx_t m, T* out) {
    out[m] = 0;
    if (m == 0) return;
    if (n == 0) { for (idx_t i = 0; i <= m; ++i) out[i] = 0; return; }
    idx_t k = n - 1, ib = m - 1, ie = m - 1;
    while (ib >= 0) {
        const T ti = t[ib];
        T cur = out[ib + 1];
        for (; k >= 0 && s[k] >= ti; --k) cur += v(k);
        for (; ie >= 0 && t[ie] == ti; --ie) out[ie] = cur;
        ib = ie;
        if (k < 0) break;
    }
    for (; ib >= 0; --ib) out[ib] = out[ib + 1];
}

// out[i] = status[i] (w[i]!=0) sum_k a[k] 1{t[k]=t[i], status[k]=1, w[k]!=0}   (glm_cox.ipp:151-175)
template <class T, class AF>
inline void nnz_event_ties_sum(AF a, const T* t, const T* status, const T* w, idx_t n, T* out) {
    idx_t ib = 0;
    while (ib < n) {
        const T ti = t[ib];
        idx_t ie = ib;
        T sum = 0;
        for (; ie < n && t[ie] == ti; ++ie) {
            const T indic = status[ie] * T(w[ie] != 0);
            sum += a(ie) * indic;
        }
        for (idx_t j = ib; j < ie; ++j) out[j] = status[j] * T(w[j] != 0) * sum;
        ib = ie;
    }
}

// Efron / Breslow scale (glm_cox.ipp:194-226)
template <class T>
inline void scale(const T* t, const T* status, const T* w, idx_t n, bool efron, T* out) {
    if (!efron) { for (idx_t i = 0; i < n; ++i) out[i] = 0; return; }
    idx_t ib = 0;
    while (ib < n) {
        const T ti = t[ib];
        idx_t ie = ib;
        int size = 0;
        for (; ie < n && t[ie] == ti; ++ie) {
            const T indic = status[ie] * T(w[ie] != 0);
            out[ie] = size * indic;
            size += (int)indic;
        }
        if (size > 1) for (idx_t j = ib; j < ie; ++j) out[j] /= size;
        ib = ie;
    }
}

template <class T>
struct Pack {                                                                  // glm_cox.ipp:228-514
    idx_t n; bool efron;
    std::vector<T> start, stop, status, weights;
    std::vector<idx_t> start_order, stop_order;
    std::vector<T> start_so, stop_to, status_to, weights_to, weights_size_to, weights_mean_to, scale_to;

    Pack(const T* st, const T* sp, const T* stat, const T* w, idx_t n_, bool efron_) : n(n_), efron(efron_),
        start(st, st + n_), stop(sp, sp + n_), status(stat, stat + n_), weights(w, w + n_)
    {
        auto order_of = [&](const std::vector<T>& x) {
            std::vector<idx_t> o(n);
            std::iota(o.begin(), o.end(), 0);
            std::sort(o.begin(), o.end(), [&](idx_t i, idx_t j) { return x[i] < x[j]; });
            return o;
        };
        auto in_order = [&](const std::vector<T>& x, const std::vector<idx_t>& o) {
            std::vector<T> r(n);
            for (idx_t i = 0; i < n; ++i) r[i] = x[o[i]];
            return r;
        };
        start_order = order_of(start); start_so = in_order(start, start_order);
        stop_order = order_of(stop); stop_to = in_order(stop, stop_order);
        status_to = in_order(status, stop_order); weights_to = in_order(weights, stop_order);
        weights_size_to.resize(n); weights_mean_to.resize(n); scale_to.resize(n);
        nnz_event_ties_sum([](idx_t) { return T(1); }, stop_to.data(), status_to.data(), weights_to.data(), n, weights_size_to.data());
        nnz_event_ties_sum([&](idx_t i) { return weights_to[i]; }, stop_to.data(), status_to.data(), weights_to.data(), n, weights_mean_to.data());
        for (idx_t i = 0; i < n; ++i) {
            if (!status_to[i] || !weights_to[i]) continue;
            weights_mean_to[i] /= weights_size_to[i];
        }
        scale(stop_to.data(), status_to.data(), weights_to.data(), n, efron, scale_to.data());
    }

    void risk_total(const std::vector<T>& z, std::vector<T>& risk_sum, std::vector<T>& ties) {
        std::vector<T> o1(n + 1), o2(n + 1);
        partial_sum_bwd([&](idx_t i) { return z[stop_order[i]]; }, stop_to.data(), n, stop_to.data(), n, o1.data());
        partial_sum_bwd([&](idx_t i) { return z[start_order[i]]; }, start_so.data(), n, stop_to.data(), n, o2.data());
        risk_sum.resize(n); ties.resize(n);
        for (idx_t i = 0; i < n; ++i) risk_sum[i] = o1[i] - o2[i];
        nnz_event_ties_sum([&](idx_t i) { return z[stop_order[i]]; }, stop_to.data(), status_to.data(), weights_to.data(), n, ties.data());
    }

    // shared by gradient (power 1) and hessian (power 2)
    void scans(const std::vector<T>& z, int power, std::vector<T>& out) {
        std::vector<T> risk_sum, ties, v(n), s1(n + 1), s2(n + 1), s3(n);
        risk_total(z, risk_sum, ties);
        for (idx_t i = 0; i < n; ++i) {
            const T rt = risk_sum[i] - scale_to[i] * ties[i];
            const T den = (power == 1 ? rt : rt * rt) + T((status_to[i] == 0) || (weights_mean_to[i] == 0));
            v[i] = status_to[i] * weights_mean_to[i] / den;
        }
        partial_sum_fwd([&](idx_t i) { return v[i]; }, stop_to.data(), n, stop_to.data(), n, s1.data());
        partial_sum_fwd([&](idx_t i) { return v[i]; }, stop_to.data(), n, start_so.data(), n, s2.data());
        for (idx_t i = 0; i < n; ++i) v[i] *= (power == 1) ? scale_to[i] : scale_to[i] * (2 - scale_to[i]);
        nnz_event_ties_sum([&](idx_t i) { return v[i]; }, stop_to.data(), status_to.data(), weights_to.data(), n, s3.data());
        out.assign(n, 0);
        for (idx_t i = 0; i < n; ++i) out[stop_order[i]] = s1[i + 1] - s3[i];
        for (idx_t i = 0; i < n; ++i) out[start_order[i]] -= s2[i + 1];
    }

    void gradient(const T* eta, T* grad) {                                     // :356-407
        std::vector<T> z(n), g;
        for (idx_t i = 0; i < n; ++i) z[i] = weights[i] * std::exp(eta[i]);
        scans(z, 1, g);
        for (idx_t i = 0; i < n; ++i) grad[i] = weights[i] * status[i] - g[i] * z[i];
    }
    void hessian(const T* eta, const T* grad, T* hess) {                       // :411-463
        std::vector<T> z(n), h;
        for (idx_t i = 0; i < n; ++i) z[i] = weights[i] * std::exp(eta[i]);
        scans(z, 2, h);
        for (idx_t i = 0; i < n; ++i) hess[i] = weights[i] * status[i] - grad[i] - h[i] * z[i] * z[i];
    }
    T loss(const T* eta) {                                                     // :467-503
        constexpr T neg_max = -std::numeric_limits<T>::max();
        if (n == 0) return 0;
        T eta_max = eta[0];
        for (idx_t i = 1; i < n; ++i) eta_max = std::max(eta_max, eta[i]);
        std::vector<T> z(n), risk_sum, ties;
        for (idx_t i = 0; i < n; ++i) z[i] = weights[i] * std::exp(eta[i] - eta_max);
        risk_total(z, risk_sum, ties);
        T a = 0, b = 0;
        for (idx_t i = 0; i < n; ++i) a += status[i] * weights[i] * (eta[i] - eta_max);
        for (idx_t i = 0; i < n; ++i) {
            const T rt = std::max<T>(risk_sum[i] - scale_to[i] * ties[i], 0);
            b += status_to[i] * weights_mean_to[i] * std::max(std::log(rt), neg_max);
        }
        return -a + b;
    }
    T loss_full() {                                                            // :507-514
        constexpr T most_neg = -std::numeric_limits<T>::max();
        T s = 0;
        for (idx_t i = 0; i < n; ++i)
            s += weights_mean_to[i] * status_to[i] * std::max(std::log(weights_size_to[i] * weights_mean_to[i] * (1 - scale_to[i])), most_neg);
        return s;
    }
};

} // namespace cox

// GlmCox with strata (glm_cox.ipp:516-750)
template <class T>
struct GlmCox : GlmBase<T> {
    using B = GlmBase<T>;
    idx_t n_stratas;
    std::vector<idx_t> strata_outer, strata_order;
    std::vector<cox::Pack<T>> packs;
    GlmCox(const T* start, const T* stop, const T* status, const idx_t* strata, const T* w, idx_t n, bool efron) {
        B::name = "cox"; B::y = status; B::w = w; B::n = n;
        n_stratas = 0;
        for (idx_t i = 0; i < n; ++i) n_stratas = std::max(n_stratas, strata[i] + 1);
        strata_outer.assign(n_stratas + 1, 0);
        for (idx_t i = 0; i < n; ++i) ++strata_outer[strata[i] + 1];
        for (idx_t i = 1; i <= n_stratas; ++i) strata_outer[i] += strata_outer[i - 1];
        strata_order.resize(n);
        std::iota(strata_order.begin(), strata_order.end(), 0);
        std::sort(strata_order.begin(), strata_order.end(), [&](idx_t i, idx_t j) {
            return (strata[i] < strata[j]) || ((strata[i] == strata[j]) && (i < j));
        });
        std::vector<T> a(n), b(n), c(n), d(n);
        for (idx_t i = 0; i < n; ++i) {
            a[i] = start[strata_order[i]]; b[i] = stop[strata_order[i]];
            c[i] = status[strata_order[i]]; d[i] = w[strata_order[i]];
        }
        for (idx_t s = 0; s < n_stratas; ++s) {
            const idx_t bi = strata_outer[s], si = strata_outer[s + 1] - bi;
            packs.emplace_back(a.data() + bi, b.data() + bi, c.data() + bi, d.data() + bi, si, efron);
        }
    }
    void gradient(const T* eta, T* grad) override {
        std::vector<T> e(B::n), g(B::n);
        for (idx_t i = 0; i < B::n; ++i) e[i] = eta[strata_order[i]];
        for (idx_t s = 0; s < n_stratas; ++s) packs[s].gradient(e.data() + strata_outer[s], g.data() + strata_outer[s]);
        for (idx_t i = 0; i < B::n; ++i) grad[strata_order[i]] = g[i];
    }
    void hessian(const T* eta, const T* grad, T* hess) override {
        std::vector<T> e(B::n), g(B::n), h(B::n);
        for (idx_t i = 0; i < B::n; ++i) { e[i] = eta[strata_order[i]]; g[i] = grad[strata_order[i]]; }
        for (idx_t s = 0; s < n_stratas; ++s)
            packs[s].hessian(e.data() + strata_outer[s], g.data() + strata_outer[s], h.data() + strata_outer[s]);
        for (idx_t i = 0; i < B::n; ++i) hess[strata_order[i]] = h[i];
    }
    T loss(const T* eta) override {
        std::vector<T> e(B::n);
        for (idx_t i = 0; i < B::n; ++i) e[i] = eta[strata_order[i]];
        T s = 0;
        for (idx_t k = 0; k < n_stratas; ++k) s += packs[k].loss(e.data() + strata_outer[k]);
        return s;
    }
    T loss_full() override {
        T s = 0;
        for (auto& pk : packs) s += pk.loss_full();
        return s;
    }
    void inv_link(const T* eta, T* out) override { for (idx_t i = 0; i < B::n; ++i) out[i] = std::exp(eta[i]); }
};

} // namespace orc
