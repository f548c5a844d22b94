// adelie_b200/csrc/device_prims.cuh -- sm_100a device primitives: mbarrier, TMA bulk copies,
// named barriers, the low-latency (LL) flagged-line exchange used for the in-kernel
// all-reduce of per-group partial gradients, warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ab {
namespace dev {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier ---------------
// (SASS: UBLKCP).  dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- named barrier over a subset of the CTA's warps -----------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- LL (flagged line) exchange ------------------------------------------------------------
// One 16-byte line carries one double {lo, flag, hi, flag}; the flag (epoch) sits in both
// 8-byte halves so a torn 16-byte store can never be mistaken for a complete one.
struct alignas(16) LLLine { uint32_t d0, f0, d1, f1; };

__device__ __forceinline__ void ll_store(LLLine* line, double v, uint32_t epoch) {
    const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"(lo), "r"(epoch), "r"(hi), "r"(epoch) : "memory");
}
// system-scope variants for lines that live in / are written from a peer GPU's memory (NVLink P2P)
__device__ __forceinline__ void ll_store_sys(LLLine* line, double v, uint32_t epoch) {
    const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"(lo), "r"(epoch), "r"(hi), "r"(epoch) : "memory");
}
__device__ __forceinline__ bool ll_try_load_sys(const LLLine* line, uint32_t epoch, double& out) {
    uint32_t d0, f0, d1, f1;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(f0), "=r"(d1), "=r"(f1) : "l"(line) : "memory");
    if (f0 == epoch && f1 == epoch) { out = __hiloint2double((int)d1, (int)d0); return true; }
    return false;
}
__device__ __forceinline__ bool ll_try_load(const LLLine* line, uint32_t epoch, double& out) {
    uint32_t d0, f0, d1, f1;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(f0), "=r"(d1), "=r"(f1) : "l"(line) : "memory");
    if (f0 == epoch && f1 == epoch) { out = __hiloint2double((int)d1, (int)d0); return true; }
    return false;
}

// ---- one-hop exchange through L2 atomics -----------------------------------------------------
__device__ __forceinline__ void red_add_f64(double* p, double v) {
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

constexpr uint32_t kSpinCheck = 0x3ffu;               // look at the clock / abort flag every 1024 spins
constexpr unsigned long long kSpinTimeoutNs = 4000000000ull;   // give up (abort the kernel) after 4 s in one wait

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Shared by every spin loop: returns true when the wait must be abandoned.  `stop` is a
// CTA-local (shared memory) shutdown flag, `abort_flag` the grid-wide one in global memory.
struct SpinGuard {
    unsigned long long t0 = 0; uint32_t spin = 0;
    // cold path, deliberately out of line: keeps the spin loops (and the kernel's instruction footprint) small
    __device__ __noinline__ bool slow_check(volatile int* abort_flag) {
        if (*abort_flag) return true;
        const unsigned long long t = global_ns();
        if (t0 == 0) t0 = t;
        else if (t - t0 > kSpinTimeoutNs) { *abort_flag = 1; return true; }
        return false;
    }
    __device__ __forceinline__ bool give_up(volatile int* abort_flag, volatile int* stop) {
        if (stop && *stop) return true;
        if (((++spin) & kSpinCheck) == 0) return slow_check(abort_flag);
        return false;
    }
};

// spin until the line carries `epoch`; returns false if the kernel was aborted
__device__ __forceinline__ bool ll_wait(const LLLine* line, uint32_t epoch, double& out, volatile int* abort_flag) {
    SpinGuard g;
    while (true) {
        if (ll_try_load(line, epoch, out)) return true;
        if (g.give_up(abort_flag, nullptr)) return false;
    }
}

// mbarrier wait that also watches the abort / stop flags
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, volatile int* stop = nullptr) {
    SpinGuard g;
    while (true) {
        if (mbar_try_wait(bar, parity)) return true;
        if (g.give_up(abort_flag, stop)) return false;
    }
}

// ---- warp reductions -----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// L2-coherent loads for data that other CTAs (or earlier phases of this CTA) rewrite
template <class T>
__device__ __forceinline__ T ld_cg(const T* p) { return __ldcg(p); }

} // namespace dev
} // namespace ab
