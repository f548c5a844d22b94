// adelie_b200/csrc/snp_tc.cuh -- multi-response transposed GEMV on the packed genotypes, on the tensor cores (tcgen05, INT8, TMEM).
//
//     out[j, l] = sum_i f_j(code[i, j]) * V[i, l],     V = v o w  (n x K, K <= 8 classes),   code in {0, 1, 2, 3 = missing}
//
// is the `mul` of kron(X, I_K) for an snp_unphased / snp_phased_ancestry matrix (reference: MatrixNaiveSNPUnphased::mul,
// matrix_naive_snp_unphased.ipp:222-262, under MatrixNaiveKroneckerEye, kronecker_eye.ipp:29-352): the KKT / invariance pass of every
// lambda of config 5 (n = 500k, p = 100k, K = 8: 5e10 genotypes x 8 classes).  It is a real GEMM, (p x n) . (n x 8), whose left
// operand takes four values: on CUDA cores it is issue bound (snp_gemv_t_kernel: 5.8 instructions per genotype, 44 ms per pass against
// a 2 ms HBM floor).  Here it is EXACT INTEGER arithmetic on the tensor cores:
//   * A (per CTA: 64 SNP columns, UMMA M = 128): rows 0-63 hold the 2-bit codes themselves (0..3) as uint8, rows 64-127 the missing
//     indicator (code == 3), K-major, 128-byte swizzle.  Sixteen codes of a 32-bit word become 16 + 16 bytes with 16 bit operations and
//     two 16-byte shared-memory stores -- the rows of a word land in the order (0,4,8,12, 1,5,9,13, ...), which is harmless because
//     the B operand is built in the same order;
//   * B (N = 32): V in 32-bit fixed point per class (scale = max |V[:, l]|), cut into four signed base-256 digits: column 8 d + l is
//     digit d of class l.  Quantised ONCE per pass by a small pre-pass kernel straight into the shared-memory image of every row chunk
//     (swizzled), so the main kernel brings a chunk's B tile in with one 8 KB bulk copy (TMA 1-D);
//   * D (TMEM, int32, 128 x 32): exact; the epilogue recombines the digits in double, sum_y = sum code * V and sum_m = sum [missing] * V,
//     and forms  (v1 - v0) (sum_y - 3 sum_m) + (v3 - v0) sum_m  with the column's values v0 .. v3 (0, 1, 2, impute, or their
//     standardized images, which are affine in the code for codes 0-2); v0 * sum_i V[i, l] is added once per column as one more partial.
// The result differs from exact arithmetic only by the fixed-point quantisation of V (2^-31 of the class maximum per element).
//
// Warp roles (544 threads): warp 0 = MMA issuer (+ TMEM allocation); warps 9-16 = fetch (16-byte loads of the packed words, six chunks
// ahead in registers, handed over through an 8-stage shared-memory ring); warps 1-8 = expand (codes / missing flags -> A tile, proxy
// fence); warp 1 lane 0 also issues the B-tile bulk copies; warps 1-4 run the epilogue.
// Every wait is bounded by %globaltimer.
#pragma once
#include "device_prims.cuh"
#include "gram_tc.cuh"
#include <cstdint>

namespace ab {

constexpr int kStcCols = 64;                         // SNP columns per CTA
constexpr int kStcKC = 256;                          // rows per chunk: two K blocks of 128 rows (128 bytes of int8 per A row)
constexpr int kStcNB = 4;                            // tile buffers (the loaders run this many chunks ahead of the MMA completions)
constexpr int kStcLoadWarps = 16;
constexpr int kStcThreads = 32 + 32 * kStcLoadWarps;
constexpr int kStcATile = 2 * 16384;                 // 2 K blocks x (128 rows x 128 bytes)
constexpr int kStcBTile = 2 * 4096;                  // 2 K blocks x (32 rows x 128 bytes)
constexpr int kStcRawStages = 8;                     // raw packed-word stages (64 columns x 64 bytes each) between the fetch and the expand warps
constexpr int kStcRawBytes = kStcCols * 64;
constexpr size_t kStcSmemBytes = 1024 + kStcNB * (kStcATile + kStcBTile) + kStcRawStages * kStcRawBytes + 512;
constexpr int kStcStatBlocks = 256;

namespace stc {
// instruction descriptor: D = S32, A = UINT8, B = INT8, both K-major, N = 32, M = 128
constexpr uint32_t kIdescI8 = (2u << 4) | (0u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
}  // namespace stc

// ---- pre-pass 1: per class l the maximum |V[:, l]| and the total sum V[:, l]  (V = v o w, or v when w == nullptr) -----------------
// partial[(block * 8 + l) * 2 + {0, 1}] = {max, sum}; grid = kStcStatBlocks blocks of 256 threads (a thread's class is fixed: K | 256 * grid)
__global__ void __launch_bounds__(256)
snp_tc_stats_kernel(const float* __restrict__ v, const float* __restrict__ w, int64_t n_elems, int K, double* __restrict__ partial)
{
    __shared__ double s_max[8][8], s_sum[8][8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t stride = (int64_t)gridDim.x * 256 / K * K;                 // multiple of K (threads beyond it idle)
    const int64_t t0 = (int64_t)blockIdx.x * 256 + tid;
    double mx = 0, sm = 0;
    int l = 0;
    if (t0 < stride) {
        l = (int)(t0 % K);
        for (int64_t e = t0; e < n_elems; e += stride) {
            const float x = w ? v[e] * w[e] : v[e];
            mx = fmax(mx, (double)fabsf(x)); sm += (double)x;
        }
    }
    // lanes of a warp hold classes (lane + const) % K: reduce per class through shared memory (K <= 8 divides 32 only for K in {1,2,4,8};
    // other K: every thread adds into its class slot serially)
    for (int k = 0; k < 8; ++k) { if (lane == 0) { s_max[warp][k] = 0; s_sum[warp][k] = 0; } }
    __syncwarp();
    for (int src = 0; src < 32; ++src) {
        if (lane == src && t0 < stride) { s_max[warp][l] = fmax(s_max[warp][l], mx); s_sum[warp][l] += sm; }
        __syncwarp();
    }
    __syncthreads();
    if (tid < 8) {
        double m2 = 0, s2 = 0;
        for (int wv = 0; wv < 8; ++wv) { m2 = fmax(m2, s_max[wv][tid]); s2 += s_sum[wv][tid]; }
        partial[((size_t)blockIdx.x * 8 + tid) * 2] = m2; partial[((size_t)blockIdx.x * 8 + tid) * 2 + 1] = s2;
    }
}
// stats[l] = max, stats[8 + l] = total, stats[16 + l] = 2^30 / max (0 when max == 0), stats[24 + l] = max / 2^30
__global__ void snp_tc_stats_finish_kernel(const double* __restrict__ partial, int n_blocks, double* __restrict__ stats)
{
    const int l = threadIdx.x;
    if (l >= 8) return;
    double mx = 0, sm = 0;
    for (int b = 0; b < n_blocks; ++b) { mx = fmax(mx, partial[((size_t)b * 8 + l) * 2]); sm += partial[((size_t)b * 8 + l) * 2 + 1]; }
    stats[l] = mx; stats[8 + l] = sm;
    stats[16 + l] = mx > 0 ? 1073741824.0 / mx : 0.0;
    stats[24 + l] = mx / 1073741824.0;
}

// ---- pre-pass 2: V -> four signed base-256 digits per element, written as the shared-memory image of every chunk's B tile ------------
// Bq[chunk][kb][d * 1024 + l * 128 + ((qw ^ l) << 4) + 4 j + t]  for row i = 256 chunk + 128 kb + 16 qw + j + 4 t  (the row order the
// loaders of the main kernel produce).  One thread per (16-row group, j, class).
__global__ void __launch_bounds__(256)
snp_tc_quant_kernel(const float* __restrict__ v, const float* __restrict__ w, int64_t n_rows, int K, int64_t n_groups16, const double* __restrict__ stats,
                    uint8_t* __restrict__ Bq)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= n_groups16 * 32) return;
    const int l = (int)(t & 7), j = (int)((t >> 3) & 3);
    const int64_t g = t >> 5;                                                  // 16-row group
    uint32_t dig[4] = {0u, 0u, 0u, 0u};
    if (l < K) {
        const double sc = stats[16 + l];
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
            const int64_t i = g * 16 + j + 4 * tt;
            int32_t qv = 0;
            if (i < n_rows) {
                const float x = w ? v[i * K + l] * w[i * K + l] : v[i * K + l];
                qv = (int32_t)__double2ll_rn((double)x * sc);                  // |qv| <= 2^30
            }
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int32_t dd = ((qv + 128) & 255) - 128;                   // signed digit in [-128, 127]
                qv = (qv - dd) >> 8;
                dig[d] |= ((uint32_t)dd & 255u) << (8 * tt);
            }
        }
    }
    const int64_t chunk = g >> 4; const int kb = (int)((g >> 3) & 1), qw = (int)(g & 7);
    uint8_t* tile = Bq + (size_t)chunk * kStcBTile + (size_t)kb * 4096;
#pragma unroll
    for (int d = 0; d < 4; ++d)
        *reinterpret_cast<uint32_t*>(tile + d * 1024 + l * 128 + ((qw ^ l) << 4) + 4 * j) = dig[d];
}

// ---- main kernel: out_part[((rb * q) + c) * K + l] for the rows of row block rb -------------------------------------------------------
__global__ void __launch_bounds__(kStcThreads, 1)
snp_gemv_tc_kernel(const uint32_t* __restrict__ packed, int64_t ldw, const float* __restrict__ impute, const float* __restrict__ center,
                   const float* __restrict__ scale, int64_t j0, int q, int K, const uint8_t* __restrict__ Bq, int n_chunks, int chunks_per_rb,
                   const double* __restrict__ stats, double* __restrict__ out_part, int* __restrict__ err_flag)
{
    extern __shared__ uint8_t stc_smem_raw[];
    const uint32_t base_u32 = dev::smem_u32(stc_smem_raw);
    uint8_t* atiles = stc_smem_raw + (((base_u32 + 1023u) & ~1023u) - base_u32);        // [kStcNB][kStcATile]
    uint8_t* btiles = atiles + kStcNB * kStcATile;                                       // [kStcNB][kStcBTile]
    uint8_t* rawst = btiles + kStcNB * kStcBTile;                                        // [kStcRawStages][kStcRawBytes]
    uint64_t* tile_full = reinterpret_cast<uint64_t*>(rawst + kStcRawStages * kStcRawBytes);   // [kStcNB] expand warps -> MMA
    uint64_t* b_full = tile_full + kStcNB;                                               // [kStcNB] bulk copy of the B tile -> MMA
    uint64_t* tile_empty = b_full + kStcNB;                                              // [kStcNB] MMA -> loaders (tcgen05.commit)
    uint64_t* accum_full = tile_empty + kStcNB;
    uint64_t* raw_full = accum_full + 1;                                                 // [kStcRawStages] fetch warps -> expand warps
    uint64_t* raw_empty = raw_full + kStcRawStages;                                      // [kStcRawStages] expand warps -> fetch warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kStcRawStages);
    volatile int* s_err = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_c0 = blockIdx.x * kStcCols;
    const int rb = blockIdx.y;
    const int ch0 = rb * chunks_per_rb, ch1 = min(n_chunks, ch0 + chunks_per_rb);
    const int nch = max(0, ch1 - ch0);

    if (tid == 0) {
        for (int b = 0; b < kStcNB; ++b) { dev::mbar_init(&tile_full[b], kStcLoadWarps / 2); dev::mbar_init(&b_full[b], 1); dev::mbar_init(&tile_empty[b], 1); }
        for (int b = 0; b < kStcRawStages; ++b) { dev::mbar_init(&raw_full[b], kStcLoadWarps / 2); dev::mbar_init(&raw_empty[b], kStcLoadWarps / 2); }
        dev::mbar_init(accum_full, 1);
        *s_err = 0;
        dev::fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dev::smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ================= MMA issuer: the whole warp walks the chunks (converged), one elected lane issues
        {
            int buf = 0; uint32_t ph = 0u;
            bool okw = true;
            for (int c = 0; c < nch; ++c) {
                okw = tc::wait_bounded(&tile_full[buf], ph, s_err) && tc::wait_bounded(&b_full[buf], ph, s_err);
                okw = __all_sync(0xffffffffu, okw);
                if (!okw) break;
                tc::fence_after();
                const uint32_t a_addr = dev::smem_u32(atiles + (size_t)buf * kStcATile), b_addr = dev::smem_u32(btiles + (size_t)buf * kStcBTile);
                if (tc::elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t da = tc::smem_desc_k_sw128(a_addr + (uint32_t)kb * 16384u), db = tc::smem_desc_k_sw128(b_addr + (uint32_t)kb * 4096u);
#pragma unroll
                        for (int k = 0; k < 4; ++k)                                   // K = 32 bytes per instruction inside the 128-byte atom
                            // four independent accumulators (TMEM columns 32 k ..): with N = 32 an MMA is far shorter than the latency of
                            // a dependent accumulation into the same tile (one accumulator: ~130 cycles per MMA, 1080 per chunk)
                            stc::mma_i8(tmem_d + (uint32_t)(32 * k), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), stc::kIdescI8, (c | kb) != 0 ? 1u : 0u);
                    }
                    tc::commit(&tile_empty[buf]);
                }
                __syncwarp();
                if (++buf == kStcNB) { buf = 0; ph ^= 1u; }
            }
            if (tc::elect_one()) tc::commit(accum_full);
            __syncwarp();
        }
    } else {
        // ================= warps 1-8: expand (raw packed words -> A tile), warps 9-16: fetch (global -> raw stage).  The two jobs sit in
        // different warps because the proxy fence the expand threads need (fence.proxy.async = MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC) waits
        // for ALL of a thread's outstanding memory operations: in the first version the same threads also held the packed words of the next
        // chunks in flight, and every fence drained that prefetch queue (ncu: 1.06 TB/s, long-scoreboard stalls at the fence).
        const bool is_fetch = warp > kStcLoadWarps / 2;
        const int xt = is_fetch ? tid - 32 - 16 * kStcLoadWarps : tid - 32;          // 0 .. 255 inside the role
        const int u = xt >> 2, qq = xt & 3;                                           // column of the tile, quarter (4 words = 64 rows) of the chunk
        bool ok = true;
        if (is_fetch) {
            const int col = tile_c0 + u;
            // (8-byte loads: a column starts at a multiple of ldw words and ldw is only guaranteed to be even)
            const uint2* src = reinterpret_cast<const uint2*>(packed + (int64_t)(j0 + min(col, q - 1)) * ldw + (int64_t)ch0 * 16) + 2 * qq;
            const long long words_left = (long long)ldw - (long long)ch0 * 16 - 4 * qq;     // words from this thread's first word to the column's end
            const int c_lo = (col < q) ? (int)max(0LL, min((long long)nch, (words_left + 14) / 16)) : 0;      // chunks whose words 0-1 exist
            const int c_hi = (col < q) ? (int)max(0LL, min((long long)nch, (words_left + 12) / 16)) : 0;      // chunks whose words 2-3 exist
            auto fetch = [&](int c, uint4& wv) {
                const uint2 a = (c < c_lo) ? __ldg(src + (size_t)c * 8) : make_uint2(0u, 0u);
                const uint2 b = (c < c_hi) ? __ldg(src + (size_t)c * 8 + 1) : make_uint2(0u, 0u);
                wv = make_uint4(a.x, a.y, b.x, b.y);
            };
            int st = 0; uint32_t ph = 1u;
            auto push = [&](int c, const uint4& wv) {
                if (!ok || c >= nch) return;
                if (!tc::wait_bounded(&raw_empty[st], ph, s_err)) { ok = false; return; }
                *reinterpret_cast<uint4*>(rawst + (size_t)st * kStcRawBytes + u * 64 + qq * 16) = wv;
                __syncwarp();
                if (lane == 0) dev::mbar_arrive(&raw_full[st]);                        // release: the warp's stores are visible to the waiter
                if (++st == kStcRawStages) { st = 0; ph ^= 1u; }
            };
            constexpr int kAhead = 6;                                                 // chunks of packed words in flight per thread (4 registers each)
            uint4 wq[kAhead];
#pragma unroll
            for (int s_ = 0; s_ < kAhead - 1; ++s_) fetch(s_, wq[s_]);
            for (int c = 0; c < nch; c += kAhead) {
#pragma unroll
                for (int s_ = 0; s_ < kAhead; ++s_) {
                    fetch(c + s_ + kAhead - 1, wq[(s_ + kAhead - 1) % kAhead]);
                    push(c + s_, wq[s_]);
                }
            }
        } else {
            auto put = [&](uint8_t* dst, uint32_t wd) {
                const uint32_t mw = wd & (wd >> 1) & 0x55555555u;
                const uint4 y = make_uint4(wd & 0x03030303u, (wd >> 2) & 0x03030303u, (wd >> 4) & 0x03030303u, (wd >> 6) & 0x03030303u);
                const uint4 m = make_uint4(mw & 0x01010101u, (mw >> 2) & 0x01010101u, (mw >> 4) & 0x01010101u, (mw >> 6) & 0x01010101u);
                *reinterpret_cast<uint4*>(dst) = y;
                *reinterpret_cast<uint4*>(dst + 8 * 1024) = m;
            };
            // words 4 qq .. 4 qq + 3 of the chunk: K block qq >> 1, 16-byte positions 4 (qq & 1) .. + 3 of the row's 128-byte line
            const uint32_t arow = (uint32_t)((qq >> 1) * 16384 + (u >> 3) * 1024 + (u & 7) * 128);
            const int p0 = 4 * (qq & 1), sw = u & 7;
            int buf = 0; uint32_t ph = 1u;                                            // tile_empty parity expected for the next use of `buf`
            int st = 0; uint32_t rph = 0u;
            const uint8_t* bq = Bq + (size_t)ch0 * kStcBTile;
            for (int c = 0; c < nch && ok; ++c) {
                if (!tc::wait_bounded(&raw_full[st], rph, s_err)) { ok = false; break; }
                const uint4 wv = *reinterpret_cast<const uint4*>(rawst + (size_t)st * kStcRawBytes + u * 64 + qq * 16);
                if (!tc::wait_bounded(&tile_empty[buf], ph, s_err)) { ok = false; break; }
                if (xt == 0) {                                                        // the chunk's B tile: one bulk copy
                    dev::mbar_arrive_expect_tx(&b_full[buf], (uint32_t)kStcBTile);
                    dev::tma_bulk_g2s(btiles + (size_t)buf * kStcBTile, bq + (size_t)c * kStcBTile, (uint32_t)kStcBTile, &b_full[buf]);
                }
                uint8_t* at = atiles + (size_t)buf * kStcATile + arow;
                put(at + (((p0 + 0) ^ sw) << 4), wv.x); put(at + (((p0 + 1) ^ sw) << 4), wv.y);
                put(at + (((p0 + 2) ^ sw) << 4), wv.z); put(at + (((p0 + 3) ^ sw) << 4), wv.w);
                dev::fence_proxy_async();
                __syncwarp();
                if (lane == 0) { dev::mbar_arrive(&tile_full[buf]); dev::mbar_arrive(&raw_empty[st]); }
                if (++buf == kStcNB) { buf = 0; ph ^= 1u; }
                if (++st == kStcRawStages) { st = 0; rph ^= 1u; }
            }
        }
        // ================= epilogue (warps 1-4): digits -> double, then sum_y / sum_m -> values
        if (warp <= 4) {
            double* comb = reinterpret_cast<double*>(atiles);                         // [128][8], reuses the first A tile (all MMAs retired)
            const int quad = warp & 3;
            const int r = quad * 32 + lane;
            bool have = ok && nch > 0 && tc::wait_bounded(accum_full, 0u, s_err);
            if (have) {
                tc::fence_after();
                int32_t acc[32];
#pragma unroll
                for (int e2 = 0; e2 < 32; ++e2) acc[e2] = 0;
#pragma unroll 1
                for (int a4 = 0; a4 < 4; ++a4) {                                      // the four accumulators add up exactly (int32)
                    uint32_t vv[32];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(vv[0]), "=r"(vv[1]), "=r"(vv[2]), "=r"(vv[3]), "=r"(vv[4]), "=r"(vv[5]), "=r"(vv[6]), "=r"(vv[7]),
                                   "=r"(vv[8]), "=r"(vv[9]), "=r"(vv[10]), "=r"(vv[11]), "=r"(vv[12]), "=r"(vv[13]), "=r"(vv[14]), "=r"(vv[15]),
                                   "=r"(vv[16]), "=r"(vv[17]), "=r"(vv[18]), "=r"(vv[19]), "=r"(vv[20]), "=r"(vv[21]), "=r"(vv[22]), "=r"(vv[23]),
                                   "=r"(vv[24]), "=r"(vv[25]), "=r"(vv[26]), "=r"(vv[27]), "=r"(vv[28]), "=r"(vv[29]), "=r"(vv[30]), "=r"(vv[31])
                                 : "r"(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(32 * a4)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e2 = 0; e2 < 32; ++e2) acc[e2] += (int32_t)vv[e2];
                }
                const int32_t* vv = acc;
#pragma unroll
                for (int l = 0; l < 8; ++l)
                    comb[r * 8 + l] = (double)(int32_t)vv[l] + 256.0 * ((double)(int32_t)vv[8 + l] + 256.0 * ((double)(int32_t)vv[16 + l] + 256.0 * (double)(int32_t)vv[24 + l]));
            } else {
#pragma unroll
                for (int l = 0; l < 8; ++l) comb[r * 8 + l] = 0.0;
            }
            dev::named_bar_sync(1, 128);
            const int e = (warp - 1) * 32 + lane;                                     // 0 .. 127: column e >> 1, classes 4 (e & 1) .. + 3
            const int cc = e >> 1, ccol = tile_c0 + cc;
            if (ccol < q) {
                float v0 = 0.f, v1 = 1.f, v3 = impute[j0 + ccol];
                if (center) { const float cn = center[j0 + ccol], sc = scale[j0 + ccol]; v0 = (0.f - cn) / sc; v1 = (1.f - cn) / sc; v3 = (v3 - cn) / sc; }
                const double a = (double)v1 - (double)v0, b = (double)v3 - (double)v0;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int l = 4 * (e & 1) + kk;
                    if (l < K) {
                        const double sy = comb[cc * 8 + l], sm = comb[(64 + cc) * 8 + l];
                        out_part[((size_t)rb * q + ccol) * K + l] = stats[24 + l] * (a * (sy - 3.0 * sm) + b * sm);
                    }
                }
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128u) : "memory");
    }
    if (tid == 0 && *s_err) atomicExch(err_flag, 1);
}

// the column-constant part v0_j * sum_i V[i, l] (non-zero only for standardize views) as one more partial block
__global__ void snp_tc_const_kernel(const float* __restrict__ center, const float* __restrict__ scale, int64_t j0, int q, int K, const double* __restrict__ stats,
                                    double* __restrict__ out_block)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)q * K) return;
    const int c = (int)(e / K), l = (int)(e % K);
    const float v0 = center ? (0.f - center[j0 + c]) / scale[j0 + c] : 0.f;
    out_block[e] = (double)v0 * stats[8 + l];
}

}  // namespace ab
