// Microbenchmark of the all-to-all partial-sum exchange between one-CTA-per-SM persistent CTAs.
// Every iteration each CTA publishes GS doubles and then reads the GS values of all CTAs.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#include <cooperative_groups.h>
struct alignas(16) Line { uint32_t d0, f0, d1, f1; };
__device__ __forceinline__ void st_line(Line* p, double v, uint32_t e) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"((uint32_t)__double2loint(v)), "r"(e), "r"((uint32_t)__double2hiint(v)), "r"(e) : "memory");
}
__device__ __forceinline__ bool ld_line(const Line* p, uint32_t e, double& v) {
    uint32_t a, b, c, d; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    if (b == e && d == e) { v = __hiloint2double((int)c, (int)a); return true; } return false;
}
// layout 0: [par][c][cta] 16B;  1: [par][cta][c] 16B;  2: [par][c][cta] 32B stride;  3: counter barrier + plain data
template <int LAYOUT>
__global__ void k(Line* buf, unsigned* counter, double* plain, int gs, int iters, long long* out, double* sink) {
    const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x, nt = blockDim.x;
    const int pad = (ncta + 31) / 32 * 32;
    double acc = 0; long long tot = 0, mx = 0;
    for (int it = 1; it <= iters; ++it) {
        const uint32_t e = it; const int par = it & 1;
        __syncthreads();
        long long t0 = clock64();
        if (LAYOUT == 3) {
            if (tid < gs) plain[(size_t)(par * ncta + cta) * gs + tid] = 1.0 + tid + cta + it;
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicAdd(counter, 1u); while (*(volatile unsigned*)counter < (unsigned)it * ncta) {} __threadfence(); }
            __syncthreads();
            for (int idx = tid; idx < gs * ncta; idx += nt) acc += __ldcg(plain + (size_t)par * ncta * gs + idx);
        } else {
            if (tid < gs) {
                Line* p;
                if (LAYOUT == 0) p = buf + ((size_t)(par * gs + tid) * pad + cta);
                else if (LAYOUT == 1) p = buf + ((size_t)(par * pad + cta) * gs + tid);
                else p = buf + 2 * ((size_t)(par * gs + tid) * pad + cta);
                st_line(p, 1.0 + tid + cta + it, e);
            }
            for (int idx = tid; idx < gs * pad; idx += nt) {
                int c, j;
                if (LAYOUT == 1) { j = idx / gs; c = idx - j * gs; } else { c = idx / pad; j = idx - c * pad; }
                if (j >= ncta) continue;
                const Line* p;
                if (LAYOUT == 0) p = buf + ((size_t)(par * gs + c) * pad + j);
                else if (LAYOUT == 1) p = buf + ((size_t)(par * pad + j) * gs + c);
                else p = buf + 2 * ((size_t)(par * gs + c) * pad + j);
                double v; while (!ld_line(p, e, v)) {}
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
template <int LAYOUT> void run(int gs, int threads, Line* buf, unsigned* counter, double* plain, long long* out, double* sink) {
    int sms = 148, iters = 2000;
    cudaMemset(buf, 0, (size_t)2 * 2 * 160 * 128 * 16); cudaMemset(counter, 0, 4);
    void* args[] = {&buf, &counter, &plain, &gs, &iters, &out, &sink};
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    cudaLaunchCooperativeKernel((void*)k<LAYOUT>, dim3(sms), dim3(threads), args, 0, 0);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[296]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long sm = 0, mx = 0; for (int i = 0; i < sms; ++i) { sm += h[2 * i]; mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx; }
    printf("layout %d gs=%d threads=%d: %.2f us/iter (events), avg %lld cycles/iter, worst single %lld  [%s]\n", LAYOUT, gs, threads, ms * 1e3 / iters, sm / sms, mx, cudaGetErrorString(e));
}
int main() {
    Line* buf; cudaMalloc(&buf, (size_t)2 * 2 * 160 * 128 * 16);
    unsigned* counter; cudaMalloc(&counter, 4); double* plain; cudaMalloc(&plain, 2 * 148 * 128 * 8);
    long long* out; cudaMalloc(&out, 296 * 8); double* sink; cudaMalloc(&sink, 8);
    for (int gs : {1, 10}) for (int threads : {128, 480}) {
        run<0>(gs, threads, buf, counter, plain, out, sink);
        run<1>(gs, threads, buf, counter, plain, out, sink);
        run<2>(gs, threads, buf, counter, plain, out, sink);
        run<3>(gs, threads, buf, counter, plain, out, sink);
    }
    return 0;
}
