// adelie_b200/csrc/common.cuh -- shared host/device helpers for the sm_100a library.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <chrono>

namespace ab {

// Mirrors the two exception classes of the reference (CORE/util/exceptions.hpp:8-56):
// core_error  -> "adelie_core: ..."         (propagates to Python as RuntimeError)
// solver_error-> "adelie_core solver: ..."  (caught inside solve, returned as `error`)
struct core_error : std::runtime_error {
    explicit core_error(const std::string& m) : std::runtime_error("adelie_core: " + m) {}
    core_error(const std::string& prefix, const std::string& m) : std::runtime_error("adelie_core " + prefix + ": " + m) {}
};
struct solver_error : core_error {
    explicit solver_error(const std::string& m) : core_error("solver", m) {}
};

#define AB_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            throw ::ab::core_error(std::string("CUDA error: ") + cudaGetErrorString(_e) +     \
                                   " at " + __FILE__ + ":" + std::to_string(__LINE__));       \
        }                                                                                     \
    } while (0)

inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Process-global knobs (CORE/configs.hpp:6-20).  min_bytes is a CPU threading
// threshold: accepted and ignored on the GPU.
struct Configs {
    static inline double hessian_min = 1e-24;
    static inline double dbeta_tol = 1e-12;
    static inline double min_bytes = 1 << 17;
    static inline double max_solver_value = 1e100;
    static inline int project = 1;
    // B200-specific knobs
    static inline int sweep_ctas = 0;          // 0 = auto (one CTA per SM, capped by rows)
    static inline int sweep_threads = 512;
    static inline int sweep_min_rows_per_cta = 1024;
    static inline int sweep_force_direct = 0;  // 1 = never stage X tiles in shared memory (debug / fallback path)
    static inline int sweep_profile = 0;       // 1 = accumulate per-phase cycle counters in the fused sweep kernel
    static inline int sweep_xchg = 1;          // per-group kernel's intra-GPU exchange: 1 = one hop through L2 atomics, 0 = two-level flagged lines (bitwise reproducible)
    static inline int glm_batched = 1;         // GLM / IRLS pin solves on the batched look-ahead kernel (panels rebuilt per IRLS iteration): 0 = off (per-group kernel), 1 = float32 states, 2 = every dtype
    static inline int kkt_skip_screen = 1;     // KKT / invariance pass over the non-screen columns only (full gradient completed when the solve ends)
    static inline int snp_tc = 1;              // packed-genotype multi-response GEMV on the tensor cores (INT8, snp_tc.cuh)
    static inline int snp_tc_min_k = 1;        // ... for at least this many classes (K = 1: 10.9 vs 14.4 ms per 5e10-genotype pass)
    static inline int panel_tc = 1;            // whole Gram panels on the tensor cores (tcgen05, TF32 operands; gram_tc.cuh): 1 = on, 0 = CUDA-core panel kernel
    static inline int panel_gemm = 0;          // Gram panels of the batched kernel: 1 = whole panels in one pass (fp32; parity-tested but measured slower), 0 = one block per pair of groups
    static inline int sweep_batch = 0;         // groups per batch of the look-ahead sweep kernel: 0 = auto (up to 6), 1 = off (per-group kernel)
    static inline int device_eigh = 1;         // batched Jacobi on device (0 = host Jacobi)
    static inline int sweep_l2_prefetch = 0;   // batched sweep kernel: L2 prefetch of the tiles of batch b + 2 by the producer warp (experiment)
    static inline int sweep_u_prefetch = 1;    // batched sweep kernel: residual-update tiles of active-set sweeps prefetched ahead of the proximal updates (0 = issued once the group moved)
    static inline int glm_fuse_means = 1;      // GLM path: IRLS-weighted column means of the screen groups from the Gram pass itself (0 = separate GEMV pass)
    static inline int cov_cluster = 0;         // CTAs of the covariance-method solver's cluster: 0 = auto (about two screen values per thread), else 1 / 2 / 4 / 8
};

constexpr int kRowAlign = 32;     // rows of every device vector / matrix column are padded to this many elements
inline int64_t pad_rows(int64_t n) { return (n + kRowAlign - 1) / kRowAlign * kRowAlign; }

// Device allocations below this size come from the device's stream-ordered memory pool (cudaMallocAsync on the default
// stream, pool configured to retain freed memory): a path solve creates and drops dozens of small buffers per state, and
// cudaMalloc / cudaFree (a device-wide synchronisation each) used to cost ~0.25 s per state.  Large buffers (the design matrix)
// keep plain cudaMalloc so that the pool never sits on tens of GB after the matrix is released.
constexpr size_t kPoolAllocMaxBytes = (size_t)256 << 20;

inline void configure_mem_pool_once() {
    static thread_local int configured_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == configured_dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    configured_dev = dev;
}

inline cudaError_t dev_malloc(void** q, size_t bytes, bool& pooled) {
    pooled = bytes <= kPoolAllocMaxBytes;
    if (pooled) {
        configure_mem_pool_once();
        cudaError_t e = cudaMallocAsync(q, bytes, 0);
        if (e == cudaSuccess) return e;
        (void)cudaGetLastError();
        pooled = false;
    }
    return cudaMalloc(q, bytes);
}
inline void dev_free(void* q, bool pooled) {
    if (!q) return;
    if (pooled) cudaFreeAsync(q, 0); else cudaFree(q);
}

// Owning device buffer (zero-initialised).
template <class T>
struct DevBuf {
    T* p = nullptr; size_t n = 0; bool pooled = false;
    DevBuf() = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), pooled(o.pooled) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { free(); p = o.p; n = o.n; pooled = o.pooled; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { free(); }
    void free() { dev_free(p, pooled); p = nullptr; n = 0; }
    void alloc(size_t n_) {
        free();
        n = n_;
        if (n) { AB_CUDA(dev_malloc((void**)&p, n * sizeof(T), pooled)); AB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), 0)); }
    }
    // grow keeping contents (new tail zeroed)
    void reserve_keep(size_t n_, cudaStream_t st = 0) {
        if (n_ <= n) return;
        size_t cap = std::max(n_, n * 2 + 64);
        T* q = nullptr; bool qp = false;
        AB_CUDA(dev_malloc((void**)&q, cap * sizeof(T), qp));
        if (p && n) AB_CUDA(cudaMemcpyAsync(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
        AB_CUDA(cudaMemsetAsync(q + n, 0, (cap - n) * sizeof(T), st));
        dev_free(p, pooled);                 // stream ordered (pool) or synchronising (cudaFree): safe after the copy either way
        p = q; n = cap; pooled = qp;
    }
    void upload(const T* h, size_t cnt, size_t off = 0, cudaStream_t st = 0) {
        if (cnt) AB_CUDA(cudaMemcpyAsync(p + off, h, cnt * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void download(T* h, size_t cnt, size_t off = 0, cudaStream_t st = 0) const {
        if (cnt) AB_CUDA(cudaMemcpyAsync(h, p + off, cnt * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
};

// Pinned host buffer for small, frequent D2H reads.
template <class T>
struct PinnedBuf {
    T* p = nullptr; size_t n = 0;
    PinnedBuf() = default;
    explicit PinnedBuf(size_t n_) { alloc(n_); }
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    void alloc(size_t n_) {
        if (p) cudaFreeHost(p);
        p = nullptr; n = n_;
        if (n) AB_CUDA(cudaMallocHost(&p, n * sizeof(T)));
    }
    void ensure(size_t n_) { if (n_ > n) alloc(n_ * 2); }
};

struct DeviceInfo {
    int device = 0; int sm_count = 0; size_t smem_optin = 0; int coop = 0;
    static const DeviceInfo& get() {
        static thread_local DeviceInfo info;
        static thread_local int cached_dev = -1;
        int dev = 0;
        AB_CUDA(cudaGetDevice(&dev));
        if (dev != cached_dev) {
            cudaDeviceProp prop;
            AB_CUDA(cudaGetDeviceProperties(&prop, dev));
            info.device = dev; info.sm_count = prop.multiProcessorCount;
            info.smem_optin = prop.sharedMemPerBlockOptin; info.coop = prop.cooperativeLaunch;
            cached_dev = dev;
        }
        return info;
    }
};

} // namespace ab
