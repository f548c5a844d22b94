// Microbenchmark: latency of warp-wide 16-byte strong (gpu-scope) loads when every CTA reads the SAME lines
// (hot) vs CTA-private lines (cold), one CTA per SM, W warps per CTA polling at once.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__device__ __forceinline__ uint4 ld_strong(const uint4* p) {
    uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint4 ld_weak(const uint4* p) {
    uint4 v; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
// mode 0: hot strong, 1: private strong, 2: hot ld.cg, 3: private ld.cg
__global__ void k(const uint4* buf, int lines_per_cta, int mode, int iters, long long* out, unsigned* sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const bool priv = mode & 1;
    const uint4* base = buf + (priv ? (size_t)blockIdx.x * lines_per_cta : 0);
    unsigned acc = 0; long long tot = 0;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        long long t0 = clock64();
        for (int l = warp * 32 + lane; l < lines_per_cta; l += nw * 32) {
            uint4 v = (mode < 2) ? ld_strong(base + l) : ld_weak(base + l);
            acc += v.x + v.y + v.z + v.w;
        }
        // force completion
        if (acc == 0x12345678u) sink[0] = acc;
        long long t1 = clock64();
        tot += t1 - t0;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = tot / iters;
    if (acc == 0x12345678u) sink[1] = acc;
}
int main() {
    int sms = 148; const int lines = 1480;  // gs=10 x 148 CTAs
    uint4* buf; cudaMalloc(&buf, (size_t)sms * lines * 16); cudaMemset(buf, 0, (size_t)sms * lines * 16);
    long long* out; cudaMalloc(&out, sms * 8); unsigned* sink; cudaMalloc(&sink, 8);
    long long h[148];
    const char* names[] = {"hot strong", "private strong", "hot ld.cg", "private ld.cg"};
    for (int threads : {32, 160, 480}) for (int mode = 0; mode < 4; ++mode) {
        k<<<sms, threads>>>(buf, lines, mode, 200, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, sms * 8, cudaMemcpyDeviceToHost);
        long long mx = 0, sm = 0; for (int i = 0; i < sms; ++i) { mx = h[i] > mx ? h[i] : mx; sm += h[i]; }
        printf("threads=%d %-16s: cycles per full read of %d lines: avg %lld max %lld\n", threads, names[mode], lines, sm / sms, mx);
    }
    return 0;
}
