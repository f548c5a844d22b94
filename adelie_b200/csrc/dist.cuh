// adelie_b200/csrc/dist.cuh -- row-sharded multi-GPU support (one process per GPU, all GPUs of one NVSwitch box).
//
// Every operator of the path is a sum over observations, so the rows of X (and of y, w, resid, eta, ...) are sharded across
// ranks while all O(p)/O(G) state is replicated and every rank executes the identical control flow.  Two kinds of
// cross-GPU reductions exist, both over NVLink peer memory (cudaIpc-mapped slabs), both summing in rank order so that every
// rank obtains bitwise identical results:
//   (1) per group update, inside the persistent sweep kernel: a third level of the flagged-line exchange -- the leader CTA of
//       each GPU stores its GPU partial into every peer's `ll3` lines (P2P stores through NVSwitch), every CTA polls its own
//       GPU's lines (sweep.cuh);
//   (2) per KKT round / screening step / IRLS iteration: a one-shot all-reduce of a device vector (gradient of all p features,
//       Gram blocks, GLM moments): push my vector into every peer's staging slot, flag it, wait for all flags, add in rank order.
#pragma once
#include "common.cuh"
#include "device_prims.cuh"

namespace ab {

constexpr int kMaxRanks = 8;
constexpr size_t kArCap = 1u << 20;            // elements (doubles) per all-reduce chunk
constexpr int kLL3GsCap = 128;                 // == kGsMax

struct DistSlabLayout {
    // byte offsets inside the peer-visible slab of one rank
    static constexpr size_t ll3_bytes = sizeof(dev::LLLine) * 2 * kLL3GsCap * kMaxRanks;
    static constexpr size_t flags_off = ll3_bytes;                                  // uint32 flags[2][kMaxRanks] (+pad)
    static constexpr size_t flags_bytes = 256;
    static constexpr size_t stage_off = flags_off + flags_bytes;                    // double stage[2][kMaxRanks][kArCap]
    static constexpr size_t stage_bytes = sizeof(double) * 2 * kMaxRanks * kArCap;
    static constexpr size_t total = stage_off + stage_bytes;
};

// wait until all ranks flagged `epoch`, then out[i] = sum_r stage[r][i] (rank order), optionally cast to T
template <class T>
__global__ void ar_wait_sum_kernel(const volatile uint32_t* flags, const double* stage, int world, uint32_t epoch, int64_t n,
                                   T* out, int* abort_flag)
{
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int ok = 1;
        dev::SpinGuard g;
        for (int r = 0; r < world; ++r) {
            while (flags[r] != epoch) { if (g.give_up(abort_flag, nullptr)) { ok = 0; break; } }
            if (!ok) break;
        }
        __threadfence_system();
        s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0;
        for (int r = 0; r < world; ++r) s += stage[(size_t)r * kArCap + i];
        out[i] = (T)s;
    }
}
template <class T>
__global__ void ar_pack_kernel(const T* in, double* out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}
__global__ static void ar_flag_kernel(volatile uint32_t* flag, uint32_t epoch) {
    __threadfence_system();
    *flag = epoch;
    __threadfence_system();
}

struct DistContext {
    int rank = 0, world = 1;
    bool connected = false;
    unsigned char* slab = nullptr;                       // my slab (cudaMalloc)
    unsigned char* peer[kMaxRanks] = {nullptr};          // peer[r]: rank r's slab mapped into my address space (peer[rank] == slab)
    uint32_t ar_epoch = 0;
    DevBuf<double> pack;                                 // my contribution converted to double
    DevBuf<int> abort_flag;

    static DistContext& get() { static DistContext ctx; return ctx; }
    bool active() const { return connected && world > 1; }

    void create(int rank_, int world_, cudaIpcMemHandle_t* handle_out) {
        if (world_ < 1 || world_ > kMaxRanks) throw core_error("world size must be in [1, 8].");
        rank = rank_; world = world_;
        if (!slab) {
            AB_CUDA(cudaMalloc(&slab, DistSlabLayout::total));
            AB_CUDA(cudaMemset(slab, 0, DistSlabLayout::stage_off));       // lines + flags (the staging area needs no init)
        }
        pack.alloc(kArCap); abort_flag.alloc(1);
        AB_CUDA(cudaIpcGetMemHandle(handle_out, slab));
        AB_CUDA(cudaDeviceSynchronize());
    }
    void connect(const cudaIpcMemHandle_t* handles) {
        for (int r = 0; r < world; ++r) {
            if (r == rank) { peer[r] = slab; continue; }
            void* p = nullptr;
            AB_CUDA(cudaIpcOpenMemHandle(&p, handles[r], cudaIpcMemLazyEnablePeerAccess));
            peer[r] = (unsigned char*)p;
        }
        connected = true;
    }
    dev::LLLine* ll3(int r) const { return reinterpret_cast<dev::LLLine*>(peer[r]); }
    uint32_t* flags(int r, int par) const { return reinterpret_cast<uint32_t*>(peer[r] + DistSlabLayout::flags_off) + par * kMaxRanks; }
    double* stage(int r, int par) const { return reinterpret_cast<double*>(peer[r] + DistSlabLayout::stage_off) + (size_t)par * kMaxRanks * kArCap; }

    // in-place sum over ranks of a device vector (identical result on every rank); no-op when not distributed
    template <class T>
    void allreduce(T* d_buf, int64_t n, cudaStream_t st = 0) {
        if (!active() || n <= 0) return;
        for (int64_t off = 0; off < n; off += (int64_t)kArCap) {
            const int64_t m = std::min<int64_t>(kArCap, n - off);
            ++ar_epoch;
            const int par = ar_epoch & 1;
            const int nb = (int)std::min<int64_t>(592, (m + 255) / 256);
            ar_pack_kernel<T><<<nb, 256, 0, st>>>(d_buf + off, pack.p, m);
            for (int r = 0; r < world; ++r) {
                AB_CUDA(cudaMemcpyAsync(stage(r, par) + (size_t)rank * kArCap, pack.p, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
                ar_flag_kernel<<<1, 1, 0, st>>>(flags(r, par) + rank, ar_epoch);
            }
            ar_wait_sum_kernel<T><<<nb, 256, 0, st>>>(flags(rank, par), stage(rank, par), world, ar_epoch, m, d_buf + off, abort_flag.p);
            AB_CUDA(cudaGetLastError());
        }
    }
    // host convenience (Python-side initial invariants): sums a host double vector over ranks
    void allreduce_host(double* h, int64_t n) {
        if (!active() || n <= 0) return;
        DevBuf<double> d((size_t)n);
        d.upload(h, n);
        allreduce<double>(d.p, n);
        d.download(h, n);
        AB_CUDA(cudaStreamSynchronize(0));
    }
};

} // namespace ab
