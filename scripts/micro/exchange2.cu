// Variants of the polling pattern of the all-to-all LL exchange (gs columns x ncta lines).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
struct alignas(16) Line { uint32_t d0, f0, d1, f1; };
__device__ __forceinline__ void st_line(Line* p, double v, uint32_t e) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"((uint32_t)__double2loint(v)), "r"(e), "r"((uint32_t)__double2hiint(v)), "r"(e) : "memory");
}
__device__ __forceinline__ bool ld_line(const Line* p, uint32_t e, double& v) {
    uint32_t a, b, c, d; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    if (b == e && d == e) { v = __hiloint2double((int)c, (int)a); return true; } return false;
}
// MODE 0: sequential per-thread polling (one line at a time)
// MODE 1: 4 lines in flight per thread (as ll_poll_all)
// MODE 2: sequential + __nanosleep(100) backoff on miss
// MODE 3: replicated lines: R=8 replicas, CTA reads replica (cta % 8); sequential polling
template <int MODE>
__global__ void k(Line* buf, int gs, int iters, long long* out, double* sink) {
    const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x, nt = blockDim.x;
    const int pad = (ncta + 31) / 32 * 32;
    const int R = (MODE == 3) ? 8 : 1;
    double acc = 0; long long tot = 0, mx = 0;
    for (int it = 1; it <= iters; ++it) {
        const uint32_t e = it; const int par = it & 1;
        __syncthreads();
        long long t0 = clock64();
        if (tid < gs * R) {
            const int c = tid % gs, r = tid / gs;
            st_line(buf + ((size_t)((r * 2 + par) * gs + c) * pad + cta), 1.0 + c + cta + it, e);
        }
        const Line* base = buf + (size_t)(((MODE == 3 ? cta % R : 0) * 2 + par) * gs) * pad;
        if (MODE == 1) {
            int idx = tid;
            while (idx < gs * pad) {
                const Line* p[4]; bool need[4]; double v[4];
                for (int q = 0; q < 4; ++q) { const int i = idx + q * nt; const int j = i % pad; need[q] = (i < gs * pad) && (j < ncta); p[q] = base + i; v[q] = 0; }
                while (true) {
                    bool pend = false;
                    for (int q = 0; q < 4; ++q) if (need[q]) { if (ld_line(p[q], e, v[q])) need[q] = false; else pend = true; }
                    if (!pend) break;
                }
                acc += v[0] + v[1] + v[2] + v[3];
                idx += 4 * nt;
            }
        } else {
            for (int idx = tid; idx < gs * pad; idx += nt) {
                const int j = idx % pad;
                if (j >= ncta) continue;
                double v; while (!ld_line(base + idx, e, v)) { if (MODE == 2) __nanosleep(100); }
                acc += v;
            }
        }
        __syncthreads();
        long long t1 = clock64();
        tot += t1 - t0; mx = (t1 - t0) > mx ? (t1 - t0) : mx;
    }
    if (tid == 0) { out[2 * cta] = tot / iters; out[2 * cta + 1] = mx; }
    if (acc == 1.2345) sink[0] = acc;
}
template <int MODE> void run(int gs, int threads, Line* buf, long long* out, double* sink) {
    int sms = 148, iters = 2000;
    cudaMemset(buf, 0, (size_t)8 * 2 * 160 * 128 * 16);
    void* args[] = {&buf, &gs, &iters, &out, &sink};
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    cudaLaunchCooperativeKernel((void*)k<MODE>, dim3(sms), dim3(threads), args, 0, 0);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[296]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long sm = 0, mx = 0; for (int i = 0; i < sms; ++i) { sm += h[2 * i]; mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx; }
    printf("mode %d gs=%d threads=%d: %.2f us/iter, avg %lld cycles/iter, worst %lld [%s]\n", MODE, gs, threads, ms * 1e3 / iters, sm / sms, mx, cudaGetErrorString(e));
}
int main() {
    Line* buf; cudaMalloc(&buf, (size_t)8 * 2 * 160 * 128 * 16);
    long long* out; cudaMalloc(&out, 296 * 8); double* sink; cudaMalloc(&sink, 8);
    for (int gs : {10}) for (int threads : {480}) {
        run<0>(gs, threads, buf, out, sink); run<1>(gs, threads, buf, out, sink);
        run<2>(gs, threads, buf, out, sink); run<3>(gs, threads, buf, out, sink);
    }
    for (int threads : {32, 64, 160}) { run<0>(10, threads, buf, out, sink); run<1>(10, threads, buf, out, sink); }
    return 0;
}
