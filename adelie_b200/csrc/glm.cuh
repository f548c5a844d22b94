// adelie_b200/csrc/glm.cuh -- GLM families on device-resident vectors.
// Interface mirrors GlmBase (CORE/glm/glm_base.hpp:65-94): gradient / hessian /
// inv_hessian_gradient / loss / loss_full / inv_link.  All vectors are device pointers of padded
// length; only the first `n` entries are meaningful (pad entries are written as 0).
#pragma once
#include "common.cuh"
#include "device_prims.cuh"
#include "dist.cuh"

namespace ab {

constexpr int kMapThreads = 256;
constexpr int kMapMaxBlocks = 1184;    // 8 CTAs per SM on 148 SMs

// Generic fused elementwise + reduction kernel: f(i, acc) is called for every i < n and may
// write outputs and accumulate NS partial sums (double).  Deterministic two-stage reduction.
template <int NS, class F>
__global__ void __launch_bounds__(kMapThreads) map_reduce_kernel(int64_t n, F f, double* __restrict__ part) {
    __shared__ double s_red[kMapThreads / 32][NS > 0 ? NS : 1];
    double acc[NS > 0 ? NS : 1];
#pragma unroll
    for (int s = 0; s < (NS > 0 ? NS : 1); ++s) acc[s] = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) f(i, acc);
    if (NS > 0) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const double v = dev::warp_sum(acc[s]);
            if (lane == 0) s_red[warp][s] = v;
        }
        __syncthreads();
        if ((int)threadIdx.x < NS) {
            double v = 0;
            for (int w = 0; w < kMapThreads / 32; ++w) v += s_red[w][threadIdx.x];
            part[(size_t)blockIdx.x * NS + threadIdx.x] = v;
        }
    }
}
template <int NS>
__global__ void map_reduce_final_kernel(const double* __restrict__ part, int n_blocks, double* __restrict__ out) {
    const int s = threadIdx.x;
    if (s >= NS) return;
    double v = 0;
    for (int b = 0; b < n_blocks; ++b) v += part[(size_t)b * NS + s];
    out[s] = v;
}

// Host helper owning the scratch of map_reduce launches.
struct MapReduce {
    DevBuf<double> part, out; PinnedBuf<double> h_out;
    MapReduce() { part.alloc((size_t)kMapMaxBlocks * 8); out.alloc(8); h_out.alloc(8); }
    static int blocks_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(kMapMaxBlocks, (n + kMapThreads - 1) / kMapThreads)); }
    // runs f over [0, n) and returns the NS sums on the host (synchronises the stream)
    template <int NS, class F>
    void run(int64_t n, F f, double* sums, cudaStream_t st = 0) {
        const int nb = blocks_for(n);
        map_reduce_kernel<NS, F><<<nb, kMapThreads, 0, st>>>(n, f, part.p);
        if (NS > 0) {
            map_reduce_final_kernel<NS><<<1, 32, 0, st>>>(part.p, nb, out.p);
            DistContext::get().allreduce<double>(out.p, NS, st);                // row-sharded: every sum is a sum over all ranks' rows
            out.download(h_out.p, NS, 0, st);
            AB_CUDA(cudaStreamSynchronize(st));
            for (int s = 0; s < NS; ++s) sums[s] = h_out.p[s];
        }
        AB_CUDA(cudaGetLastError());
    }
    template <class F>
    void map(int64_t n, F f, cudaStream_t st = 0) {
        const int nb = blocks_for(n);
        map_reduce_kernel<0, F><<<nb, kMapThreads, 0, st>>>(n, f, part.p);
        AB_CUDA(cudaGetLastError());
    }
};

template <class T>
struct Glm {
    std::string name; bool is_multi = false;
    int64_t n = 0;        // number of entries of eta (n, or n*K for multi-response, flattened row-major)
    int64_t K = 1;
    DevBuf<T> y, w;       // device copies (padded); for multi: y is (n*K), w is (n,)
    MapReduce mr;
    virtual ~Glm() {}
    virtual void gradient(const T* eta, T* grad) = 0;
    virtual void hessian(const T* eta, const T* grad, T* hess) = 0;
    virtual void inv_hessian_gradient(const T* eta, const T* grad, const T* hess, T* out) {          // glm_base.ipp:25-36
        const T hmin = (T)Configs::hessian_min;
        mr.map(n, [=] __device__(int64_t i, double*) {
            const T h = hess[i];
            out[i] = grad[i] / (max(h, T(0)) + hmin * T(h <= 0));
        });
    }
    virtual T loss(const T* eta) = 0;
    virtual T loss_full() = 0;
    virtual void inv_link(const T* eta, T* out) = 0;
};

// glm_gaussian.ipp:17-64
template <class T>
struct GlmGaussian : Glm<T> {
    using B = Glm<T>;
    GlmGaussian(const T* hy, const T* hw, int64_t n_) {
        B::name = "gaussian"; B::n = n_;
        B::y.alloc(pad_rows(n_)); B::w.alloc(pad_rows(n_));
        B::y.upload(hy, n_); B::w.upload(hw, n_);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    void gradient(const T* eta, T* grad) override {
        const T* y = B::y.p; const T* w = B::w.p;
        B::mr.map(B::n, [=] __device__(int64_t i, double*) { grad[i] = w[i] * (y[i] - eta[i]); });
    }
    void hessian(const T*, const T*, T* hess) override {
        const T* w = B::w.p;
        B::mr.map(B::n, [=] __device__(int64_t i, double*) { hess[i] = w[i]; });
    }
    T loss(const T* eta) override {
        const T* y = B::y.p; const T* w = B::w.p; double s;
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) { acc[0] += (double)(w[i] * (T(0.5) * eta[i] * eta[i] - y[i] * eta[i])); }, &s);
        return (T)s;
    }
    T loss_full() override {
        const T* y = B::y.p; const T* w = B::w.p; double s;
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) { acc[0] += (double)(y[i] * y[i] * w[i]); }, &s);
        return (T)(-0.5 * s);
    }
    void inv_link(const T* eta, T* out) override { B::mr.map(B::n, [=] __device__(int64_t i, double*) { out[i] = eta[i]; }); }
};

// glm_binomial.ipp:47-98 (+ loss_full :14-35)
template <class T>
struct GlmBinomialLogit : Glm<T> {
    using B = Glm<T>;
    GlmBinomialLogit(const T* hy, const T* hw, int64_t n_) {
        B::name = "binomial_logit"; B::n = n_;
        B::y.alloc(pad_rows(n_)); B::w.alloc(pad_rows(n_));
        B::y.upload(hy, n_); B::w.upload(hw, n_);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    void gradient(const T* eta, T* grad) override {
        const T* y = B::y.p; const T* w = B::w.p;
        B::mr.map(B::n, [=] __device__(int64_t i, double*) { grad[i] = w[i] * (y[i] - T(1) / (T(1) + exp(-eta[i]))); });
    }
    void hessian(const T*, const T* grad, T* hess) override {
        const T* y = B::y.p; const T* w = B::w.p;
        B::mr.map(B::n, [=] __device__(int64_t i, double*) {
            const T h = w[i] * y[i] - grad[i];
            hess[i] = (h * (w[i] - h)) / (w[i] + T(w[i] <= 0));
        });
    }
    T loss(const T* eta) override {
        const T* y = B::y.p; const T* w = B::w.p; double s;
        const T mx = std::numeric_limits<T>::max();
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) {
            const T e = eta[i];
            const T ec = max(min(e, mx), -mx);
            acc[0] += (double)(w[i] * ((T(e > 0) - y[i]) * ec + log(T(1) + exp(-fabs(e)))));
        }, &s);
        return (T)s;
    }
    T loss_full() override {
        const T* y = B::y.p; const T* w = B::w.p; double s;
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) {
            const T yi = y[i];
            const T ly = log(yi), l1 = log(T(1) - yi);
            T v = 0;
            if (!(isinf(ly) || isnan(ly))) v -= w[i] * yi * ly;
            if (!(isinf(l1) || isnan(l1))) v -= w[i] * (T(1) - yi) * l1;
            acc[0] += (double)v;
        }, &s);
        return (T)s;
    }
    void inv_link(const T* eta, T* out) override {
        B::mr.map(B::n, [=] __device__(int64_t i, double*) { out[i] = T(1) / (T(1) + exp(-eta[i])); });
    }
};

// glm_binomial.ipp:100-190 (probit link); loss_full is the binomial one (:14-35)
template <class T> __device__ __forceinline__ T probit_cdf(T x) { return T(0.5) * (T(1) + erf(x * T(0.70710678118654752440))); }
template <class T> __device__ __forceinline__ T probit_pdf(T x) { return T(0.39894228040143267794) * exp(T(-0.5) * x * x); }
template <class T>
struct GlmBinomialProbit : Glm<T> {
    using B = Glm<T>;
    GlmBinomialProbit(const T* hy, const T* hw, int64_t n_) {
        B::name = "binomial_probit"; B::n = n_;
        B::y.alloc(pad_rows(n_)); B::w.alloc(pad_rows(n_));
        B::y.upload(hy, n_); B::w.upload(hw, n_);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    void gradient(const T* eta, T* grad) override {
        const T* y = B::y.p; const T* w = B::w.p; const T mx = std::numeric_limits<T>::max();
        B::mr.map(B::n, [=] __device__(int64_t i, double*) {
            const T P = probit_cdf<T>(eta[i]);
            grad[i] = w[i] * probit_pdf<T>(eta[i]) * (y[i] * min(T(1) / P, mx) - (T(1) - y[i]) * min(T(1) / (T(1) - P), mx));
        });
    }
    void hessian(const T* eta, const T* grad, T* hess) override {
        const T* y = B::y.p; const T* w = B::w.p; const T mx = std::numeric_limits<T>::max();
        B::mr.map(B::n, [=] __device__(int64_t i, double*) {
            const T P = probit_cdf<T>(eta[i]), ph = probit_pdf<T>(eta[i]);
            hess[i] = w[i] * (y[i] * min(T(1) / (P * P), mx) + (T(1) - y[i]) * min(T(1) / ((T(1) - P) * (T(1) - P)), mx)) * ph * ph + eta[i] * grad[i];
        });
    }
    T loss(const T* eta) override {
        const T* y = B::y.p; const T* w = B::w.p; const T mx = std::numeric_limits<T>::max(); double s;
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) {
            const T P = probit_cdf<T>(eta[i]);
            acc[0] -= (double)(w[i] * (y[i] * max(log(P), -mx) + (T(1) - y[i]) * max(log(T(1) - P), -mx)));
        }, &s);
        return (T)s;
    }
    T loss_full() override {
        const T* y = B::y.p; const T* w = B::w.p; double s;
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) {
            const T yi = y[i];
            const T ly = log(yi), l1 = log(T(1) - yi);
            T v = 0;
            if (!(isinf(ly) || isnan(ly))) v -= w[i] * yi * ly;
            if (!(isinf(l1) || isnan(l1))) v -= w[i] * (T(1) - yi) * l1;
            acc[0] += (double)v;
        }, &s);
        return (T)s;
    }
    void inv_link(const T* eta, T* out) override { B::mr.map(B::n, [=] __device__(int64_t i, double*) { out[i] = probit_cdf<T>(eta[i]); }); }
};

// glm_poisson.ipp:7-66 (log link)
template <class T>
struct GlmPoisson : Glm<T> {
    using B = Glm<T>;
    GlmPoisson(const T* hy, const T* hw, int64_t n_) {
        B::name = "poisson"; B::n = n_;
        B::y.alloc(pad_rows(n_)); B::w.alloc(pad_rows(n_));
        B::y.upload(hy, n_); B::w.upload(hw, n_);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    void gradient(const T* eta, T* grad) override {
        const T* y = B::y.p; const T* w = B::w.p;
        B::mr.map(B::n, [=] __device__(int64_t i, double*) { grad[i] = w[i] * (y[i] - exp(eta[i])); });
    }
    void hessian(const T*, const T* grad, T* hess) override {
        const T* y = B::y.p; const T* w = B::w.p;
        B::mr.map(B::n, [=] __device__(int64_t i, double*) { hess[i] = w[i] * y[i] - grad[i]; });
    }
    T loss(const T* eta) override {
        const T* y = B::y.p; const T* w = B::w.p; double s;
        const T mx = std::numeric_limits<T>::max();
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) { acc[0] += (double)(w[i] * (min(-eta[i], mx) * y[i] + exp(eta[i]))); }, &s);
        return (T)s;
    }
    T loss_full() override {
        const T* y = B::y.p; const T* w = B::w.p; double s;
        const T mx = std::numeric_limits<T>::max();
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) { acc[0] += (double)(w[i] * (min(-log(y[i]), mx) * y[i] + y[i])); }, &s);
        return (T)s;
    }
    void inv_link(const T* eta, T* out) override { B::mr.map(B::n, [=] __device__(int64_t i, double*) { out[i] = exp(eta[i]); }); }
};

// glm_multigaussian.ipp:17-68: y, eta (n,K) row-major flattened; weights (n,); everything / K.
template <class T>
struct GlmMultiGaussian : Glm<T> {
    using B = Glm<T>;
    GlmMultiGaussian(const T* hy, const T* hw, int64_t n_rows, int64_t K_) {
        B::name = "multigaussian"; B::is_multi = true; B::K = K_; B::n = n_rows * K_;
        B::y.alloc(pad_rows(n_rows) * K_); B::w.alloc(pad_rows(n_rows));
        B::y.upload(hy, n_rows * K_); B::w.upload(hw, n_rows);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    void gradient(const T* eta, T* grad) override {
        const T* y = B::y.p; const T* w = B::w.p; const int64_t K = B::K;
        B::mr.map(B::n, [=] __device__(int64_t i, double*) { grad[i] = w[i / K] * (y[i] - eta[i]) / T(K); });
    }
    void hessian(const T*, const T*, T* hess) override {
        const T* w = B::w.p; const int64_t K = B::K;
        B::mr.map(B::n, [=] __device__(int64_t i, double*) { hess[i] = w[i / K] / T(K); });
    }
    T loss(const T* eta) override {
        const T* y = B::y.p; const T* w = B::w.p; const int64_t K = B::K; double s;
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) { acc[0] += (double)(w[i / K] * (T(0.5) * eta[i] * eta[i] - y[i] * eta[i])); }, &s);
        return (T)(s / K);
    }
    T loss_full() override {
        const T* y = B::y.p; const T* w = B::w.p; const int64_t K = B::K; double s;
        B::mr.template run<1>(B::n, [=] __device__(int64_t i, double* acc) { acc[0] += (double)(w[i / K] * y[i] * y[i]); }, &s);
        return (T)(-0.5 * s / K);
    }
    void inv_link(const T* eta, T* out) override { B::mr.map(B::n, [=] __device__(int64_t i, double*) { out[i] = eta[i]; }); }
};

// glm_multinomial.ipp:6-132: y, eta (n,K) row-major flattened, weights (n,); one thread per observation, K <= 16 classes in registers.
template <class T>
struct GlmMultinomial : Glm<T> {
    using B = Glm<T>;
    int64_t rows = 0;
    GlmMultinomial(const T* hy, const T* hw, int64_t n_rows, int64_t K_) : rows(n_rows) {
        if (K_ <= 1) throw core_error("y must have at least 2 columns (classes).");
        if (K_ > 16) throw core_error("multi-response problems with more than 16 classes are not supported.");
        B::name = "multinomial"; B::is_multi = true; B::K = K_; B::n = n_rows * K_;
        B::y.alloc(pad_rows(n_rows) * K_); B::w.alloc(pad_rows(n_rows));
        B::y.upload(hy, n_rows * K_); B::w.upload(hw, n_rows);
        AB_CUDA(cudaStreamSynchronize(0));
    }
    void gradient(const T* eta, T* grad) override {
        const T* y = B::y.p; const T* w = B::w.p; const int K = (int)B::K;
        B::mr.map(rows, [=] __device__(int64_t i, double*) {
            const T* e = eta + i * K;
            T m = e[0]; for (int k = 1; k < K; ++k) m = max(m, e[k]);
            T sum = 0; for (int k = 0; k < K; ++k) sum += exp(e[k] - m);
            for (int k = 0; k < K; ++k) grad[i * K + k] = (y[i * K + k] - exp(e[k] - m) / sum) * w[i] / T(K);
        });
    }
    void hessian(const T*, const T* grad, T* hess) override {
        const T* y = B::y.p; const T* w = B::w.p; const int K = (int)B::K;
        B::mr.map(B::n, [=] __device__(int64_t e, double*) {
            const int64_t i = e / K;
            const T h = y[e] * w[i] / T(K) - grad[e];
            hess[e] = h * T(2) * (T(1) - T(K) * (h / (w[i] + T(w[i] <= 0))));
        });
    }
    T loss(const T* eta) override {
        const T* y = B::y.p; const T* w = B::w.p; const int K = (int)B::K; double s;
        B::mr.template run<1>(rows, [=] __device__(int64_t i, double* acc) {
            const T* e = eta + i * K;
            T m = e[0]; for (int k = 1; k < K; ++k) m = max(m, e[k]);
            T ye = 0, se = 0;
            for (int k = 0; k < K; ++k) { ye += y[i * K + k] * (e[k] - m); se += exp(e[k] - m); }
            acc[0] += (double)(w[i] * (-ye + log(se)));
        }, &s);
        return (T)(s / K);
    }
    T loss_full() override {
        const T* y = B::y.p; const T* w = B::w.p; const int K = (int)B::K; double s;
        B::mr.template run<1>(rows, [=] __device__(int64_t i, double* acc) {
            T sum = 0;
            for (int k = 0; k < K; ++k) { const T yk = y[i * K + k]; const T l = log(yk); if (!(isinf(l) || isnan(l))) sum += yk * l; }
            acc[0] -= (double)(sum * w[i]);
        }, &s);
        return (T)(s / K);
    }
    void inv_link(const T* eta, T* out) override {
        const int K = (int)B::K;
        B::mr.map(rows, [=] __device__(int64_t i, double*) {
            const T* e = eta + i * K;
            T m = e[0]; for (int k = 1; k < K; ++k) m = max(m, e[k]);
            T sum = 0; for (int k = 0; k < K; ++k) sum += exp(e[k] - m);
            for (int k = 0; k < K; ++k) out[i * K + k] = exp(e[k] - m) / sum;
        });
    }
};

} // namespace ab
