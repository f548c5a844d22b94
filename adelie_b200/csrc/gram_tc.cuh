// adelie_b200/csrc/gram_tc.cuh -- weighted window Gram on the 5th-generation tensor cores (tcgen05 / TMEM), fp32 data as TF32.
//
// One work item = one Gram PANEL of the batched look-ahead sweep (sweep_batched.cuh): the <= 128 columns of a window (the groups of
// batch b followed by the groups of batch b + 1) over one block of rows,
//        D = (sqrt(w) o X_win)^T (sqrt(w) o X_win)            (128 x 128, fp32 accumulate in TMEM),
// of which the rows of batch b (the first n_src <= 64 window columns) are the panel.  This is the one real contraction on the path
// (2 * 128 flops per loaded byte): on CUDA cores it is compute bound (the fp32 FMA pipe needs ~25 us per 16 MB window at
// n = 250k), on tcgen05 it is HBM bound.  The GLM / IRLS path rebuilds every panel in every IRLS iteration (the weights change), which
// is what makes the batched sweep kernel affordable there (reference: the per-group loop of solver_gaussian_pin_naive.hpp:26-168 is
// what the panels restore exactly; the weights are those of solver_glm_naive.hpp:328-372).
//
// Panels only PREDICT how the stale gradients of a batch move while earlier groups of the batch are updated; every gradient is
// recomputed from the residual at its next visit and every correction is multiplied by a coefficient change that vanishes at
// convergence.  TF32 operands (10-bit mantissa, round-to-nearest) are therefore exact enough: the fixed point of the sweep does not
// depend on them.  The groups' own Gram blocks (eigen-decomposed for the proximal step) stay on the fp32/fp64 CUDA-core kernel.
//
// Warp roles (544 threads, one CTA per SM):
//   warp 0      MMA issuer: one elected lane issues 8 x tcgen05.mma.cta_group::1.kind::tf32 (M = N = 128, K = 8) per 64-row chunk, A and B
//               descriptors pointing at the SAME shared-memory tile (K-major, 128-byte swizzle), tcgen05.commit on the tile's
//               "empty" mbarrier; allocates / frees the 128 TMEM columns of the accumulator;
//   warps 1-16  loaders: 16-byte coalesced global loads of the chunk (two chunks ahead, in registers), scaled by sqrt(w), rounded to
//               TF32 and stored in the canonical UMMA K-major SWIZZLE_128B layout (8-column x 128-byte atoms, 16-byte chunks XOR-ed
//               with the column index) of one of four tile buffers; afterwards warps 1-4 read the accumulator back (tcgen05.ld
//               32x32b) and write the panel's rows.
// The first version staged the chunk through cp.async.bulk (one 256-byte copy per window column and chunk) and transformed it from
// shared memory: 35 s of panel time per config-3 shard path (86 ms per IRLS iteration) -- the copy engine is made for few large copies,
// not 120 small ones per 30 KB; the register path needs no staging buffer and leaves room for four tile buffers.
// Every wait is bounded by %globaltimer: a protocol bug reports an error instead of hanging the GPU.
#pragma once
#include "device_prims.cuh"
#include <cstdint>

namespace ab {

// One panel: the window's physical columns (sources = the first n_src of them) and where the panel starts in the panel buffer.
struct PanelItem { int64_t q_off; int32_t ncol, n_src; int32_t cols[128]; };
constexpr int kPanelOut = 64 * 128;            // compact outputs per panel: [source][window column]

constexpr int kTcKC = 64;                        // rows per chunk (two 32-row = 128-byte K blocks)
constexpr int kTcNB = 4;                         // UMMA tile buffers
constexpr int kTcLoadWarps = 16;
constexpr int kTcThreads = 32 + 32 * kTcLoadWarps;
constexpr int kTcNI = 128 * (kTcKC / 4) / (32 * kTcLoadWarps);     // 16-byte vectors per loader thread and chunk (4)
constexpr int kTcTileBytes = 2 * 16384;          // one UMMA tile: 2 K blocks x (128 columns x 128 bytes)
constexpr size_t kTcSmemBytes = 1024 + kTcNB * kTcTileBytes + 128 * 4 + 256;

namespace tc {
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity, volatile int* err) {
    unsigned long long t0 = 0; uint32_t spin = 0;
    while (!dev::mbar_try_wait(bar, parity)) {
        if (((++spin) & 0xfffu) == 0) {
            if (*err) return false;
            const unsigned long long t = dev::global_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) { *err = 1; return false; }
        }
    }
    return true;
}
// one lane of a converged warp (the compiler keeps warp-uniform operands of the instructions it guards in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dev::smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: start address, LBO (unused for swizzled K-major, 1), SBO = 1024 bytes between
// 8-row groups, descriptor version 1 (Blackwell), layout type 2
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t kIdescTf32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
}  // namespace tc

// part[(rb * n_panels + panel) * 64 * 128 + s * 128 + u] = sum over the rows of block rb of wsqrt^2 * X[:, cols[s]] * X[:, cols[u]]
__global__ void __launch_bounds__(kTcThreads, 1)
panel_gram_tc_kernel(const float* __restrict__ X, int64_t ld, int64_t n_pad, const PanelItem* __restrict__ items, const float* __restrict__ wsqrt,
                     float* __restrict__ part, int n_panels, int rows_per_block, int* __restrict__ err_flag)
{
    extern __shared__ uint8_t tc_smem_raw[];
    // carve: [tile buffers kTcNB x 32 KB, 1024-byte aligned][cols][barriers]
    const uint32_t base_u32 = dev::smem_u32(tc_smem_raw);
    uint8_t* tiles = tc_smem_raw + (((base_u32 + 1023u) & ~1023u) - base_u32);
    int* cols_s = reinterpret_cast<int*>(tiles + kTcNB * kTcTileBytes);
    uint64_t* tile_full = reinterpret_cast<uint64_t*>(cols_s + 128);      // [kTcNB] loaders -> MMA (one arrive per loader warp)
    uint64_t* tile_empty = tile_full + kTcNB;                             // [kTcNB] MMA -> loaders (tcgen05.commit)
    uint64_t* accum_full = tile_empty + kTcNB;                            // [1] MMA -> epilogue (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
    volatile int* s_err = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const PanelItem& it = items[blockIdx.x];
    const int ncol = it.ncol, n_src = it.n_src;
    const int rb = blockIdx.y;
    const int64_t row0 = (int64_t)rb * rows_per_block;
    const int64_t row1 = min((long long)n_pad, (long long)(row0 + rows_per_block));
    const int nchunks = row1 > row0 ? (int)((row1 - row0 + kTcKC - 1) / kTcKC) : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int b = 0; b < kTcNB; ++b) { dev::mbar_init(&tile_full[b], kTcLoadWarps); dev::mbar_init(&tile_empty[b], 1); }
        dev::mbar_init(accum_full, 1);
        *s_err = 0;
        dev::fence_barrier_init();
    }
    for (int e = tid; e < 128; e += kTcThreads) cols_s[e] = (e < ncol) ? it.cols[e] : 0;
    // window columns that do not exist stay zero in every tile buffer for the whole kernel (their outputs are never read)
    for (int e = tid; e < kTcNB * kTcTileBytes / 16; e += kTcThreads) reinterpret_cast<uint4*>(tiles)[e] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dev::smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    dev::fence_proxy_async();                                  // the zero fill is read by the tensor core (async proxy)
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ================= MMA issuer (one lane)
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int buf = c % kTcNB; const uint32_t use = (uint32_t)(c / kTcNB);
                if (!tc::wait_bounded(&tile_full[buf], use & 1u, s_err)) break;
                tc::fence_after();
                const uint32_t tile_addr = dev::smem_u32(tiles + (size_t)buf * kTcTileBytes);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t desc0 = tc::smem_desc_k_sw128(tile_addr + (uint32_t)kb * 16384u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t d = desc0 + (uint64_t)(2 * k);             // + 32 bytes along K inside the 128-byte swizzle atom
                        tc::mma_tf32(tmem_d, d, d, tc::kIdescTf32_128x128, (c | kb | k) != 0 ? 1u : 0u);
                    }
                }
                tc::commit(&tile_empty[buf]);                                      // the tile may be overwritten once these MMAs retired
            }
            tc::commit(accum_full);
        }
    } else {
        // ================= loader warps: thread xt owns the 16-byte row vector q of columns u0, u0 + 32, u0 + 64, u0 + 96
        const int xt = tid - 32;
        const int q = xt & 15, u0 = xt >> 4;
        const float* colp[kTcNI];
#pragma unroll
        for (int i = 0; i < kTcNI; ++i) colp[i] = X + (int64_t)cols_s[min(u0 + 32 * i, 127)] * ld + 4 * q;
        const int kb = q >> 3, ch = q & 7;
        uint32_t toff[kTcNI];
#pragma unroll
        for (int i = 0; i < kTcNI; ++i) { const int u = u0 + 32 * i; toff[i] = (uint32_t)(kb * 16384 + (u >> 3) * 1024 + (u & 7) * 128 + ((ch ^ (u & 7)) << 4)); }
        auto load = [&](int c, float4 (&x)[kTcNI], float4& wv) {
            const int64_t r = row0 + (int64_t)c * kTcKC;
            const bool live = c < nchunks && r + 4 * q < row1;
            wv = live ? __ldg(reinterpret_cast<const float4*>(wsqrt + r + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < kTcNI; ++i)
                x[i] = (live && u0 + 32 * i < ncol) ? __ldg(reinterpret_cast<const float4*>(colp[i] + r)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        bool ok = true;
        auto store = [&](int c, const float4 (&x)[kTcNI], const float4& wv) {
            if (!ok || c >= nchunks) return;
            const int buf = c % kTcNB; const uint32_t use = (uint32_t)(c / kTcNB);
            if (!tc::wait_bounded(&tile_empty[buf], (use & 1u) ^ 1u, s_err)) { ok = false; return; }
            uint8_t* tile = tiles + (size_t)buf * kTcTileBytes;
#pragma unroll
            for (int i = 0; i < kTcNI; ++i) {
                if (u0 + 32 * i < ncol) {
                    uint4 y;
                    y.x = tc::to_tf32(x[i].x * wv.x); y.y = tc::to_tf32(x[i].y * wv.y); y.z = tc::to_tf32(x[i].z * wv.z); y.w = tc::to_tf32(x[i].w * wv.w);
                    *reinterpret_cast<uint4*>(tile + toff[i]) = y;
                }
            }
            dev::fence_proxy_async();                                              // generic-proxy stores -> visible to the MMA
            __syncwarp();
            if (lane == 0) dev::mbar_arrive(&tile_full[buf]);
        };
        float4 xa[kTcNI], xb[kTcNI], xc[kTcNI], wa, wb, wc;
        load(0, xa, wa); load(1, xb, wb);
        for (int c = 0; c < nchunks; c += 3) {
            load(c + 2, xc, wc); store(c, xa, wa);
            load(c + 3, xa, wa); store(c + 1, xb, wb);
            load(c + 4, xb, wb); store(c + 2, xc, wc);
        }
        // ================= epilogue: TMEM -> registers -> the panel's rows (warps 1..4 cover the four 32-lane quadrants)
        if (ok && warp <= 4 && nchunks > 0) {
            const int quad = warp & 3;                                             // a warp may only touch TMEM lanes [32 quad, 32 quad + 32)
            if (quad * 32 < n_src && tc::wait_bounded(accum_full, 0u, s_err)) {
                tc::fence_after();
                const int srow = quad * 32 + lane;
                float* out = part + ((size_t)rb * n_panels + blockIdx.x) * kPanelOut + (size_t)srow * 128;
#pragma unroll 1
                for (int j0 = 0; j0 < 128; j0 += 32) {
                    uint32_t v[32];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)j0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (srow < n_src) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (j0 + j < ncol)
                                *reinterpret_cast<float4*>(out + j0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                                       __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    }
                }
            }
        }
    }
    // a CTA without rows still owns its slice of `part`
    if (nchunks == 0) for (int e = tid; e < n_src * 128; e += kTcThreads) part[((size_t)rb * n_panels + blockIdx.x) * kPanelOut + e] = 0.f;
    tc::fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128u) : "memory");
    }
    if (tid == 0 && *s_err) atomicExch(err_flag, 1);
}

// wsqrt = sqrt(max(w, 0)) (the operand scaling of the Gram: D = Y^T Y with Y = sqrt(w) o X)
__global__ void sqrt_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sqrtf(fmaxf(w[i], 0.f));
}

}  // namespace ab
