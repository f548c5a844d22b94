// adelie_b200/csrc/snp.cuh -- SNP unphased genotype matrices.
//
// Reference: `IOSNPUnphased` (CORE/io/io_snp_unphased.hpp:137-274, .ipp:9-305, base class CORE/io/io_snp_base.ipp:20-84) and
// `MatrixNaiveSNPUnphased` (CORE/matrix/matrix_naive_snp_unphased.ipp:10-309).  A genotype is 0, 1, 2 or missing; a missing
// entry of column j takes the column's imputed value impute[j], so the operator set is that of the dense matrix
// X[i, j] in {0, 1, 2, impute[j]} (the reference's own tests compare against exactly that matrix, T/test_matrix.py:721-745,
// T/test_solver.py:756-818).
//
// On-disk `.snpdat` (io_snp_unphased.ipp:88-110, 164-262), all little endian:
//   [endian:1B][n:u64][p:u64][nnz:u64 x p][nnm:u64 x p][impute:f64 x p][outer:u64 x (p+1)]
//   column j at byte outer[j]: [3 x u64 offsets of the categories, relative to the column start]
//     category c (0 = missing, 1 = ones, 2 = twos): [n_chunks:u32] { [chunk_idx:u32][nnz-1:u8][row_in_chunk:u8 x nnz] }, chunk = 256 rows
//
// HBM layout (ours): 2 bits per genotype, column-major, 4 genotypes per byte (row i of column j = bits 2*(i%4).. of byte
// j*ldb + i/4, ldb = n_pad/4), codes 0/1/2 and 3 = missing, plus impute (p,) in the value type.  12.5 GB at config 5
// (n=500k, p=100k) against 200 GB for the same matrix in fp32.  The file bytes are shipped to the device as they are and a
// kernel walks the chunk lists (one warp per (column, category)); the host only parses the header.
//   * full-matrix `mul` / `sq_mul` (the KKT pass, every lambda) run on the packed bits (snp_gemv_t_kernel);
//   * the columns of the screen set are decoded once into a dense column cache (snp_decode_kernel) the fused sweep, Gram and
//     panel kernels read through TMA exactly like a dense matrix: a column enters the cache when its group enters the screen set.
#pragma once
#include "common.cuh"
#include "device_prims.cuh"
#include "sweep.cuh"   // VecT / vec_load / vec_store
#include <curand_kernel.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <cmath>
#include <algorithm>

namespace ab {

// ============================================================================================ host: .snpdat reader / writer
// What both formats share (IOSNPBase, CORE/io/io_snp_base.hpp:17-141, .ipp:20-84): file name, read mode, the bytes (read or mapped),
// the endianness byte.
struct SnpFileBase {
    static constexpr uint64_t kChunk = 256;
    std::string filename; int read_mode = 0;     // 0 = file, 1 = mmap
    bool is_read = false;
    std::vector<char> owned; const char* buf = nullptr; size_t buf_bytes = 0; void* map_addr = nullptr; size_t map_bytes = 0;

    SnpFileBase(const std::string& f, const std::string& mode) : filename(f) {
        // util::convert_read_mode + IOSNPBase::convert_read_mode (io_snp_base.hpp:123-139): "auto" means mmap on Linux
        if (mode == "file") read_mode = 0;
        else if (mode == "mmap" || mode == "auto") read_mode = 1;
        else throw core_error("Invalid read mode type: " + mode);
    }
    ~SnpFileBase() { release(); }
    SnpFileBase(const SnpFileBase&) = delete;
    SnpFileBase& operator=(const SnpFileBase&) = delete;
    void release() { if (map_addr) { munmap(map_addr, map_bytes); map_addr = nullptr; } owned.clear(); owned.shrink_to_fit(); buf = nullptr; buf_bytes = 0; }
    void need_read() const { if (!is_read) throw core_error("File is not read yet. Call read() first."); }
    static bool big_endian() { const uint32_t one = 1; return reinterpret_cast<const char*>(&one)[0] != 1; }
    template <class U> static U rd(const char* q) { U u; std::memcpy(&u, q, sizeof(U)); return u; }

    // io_snp_base.ipp:20-84: reads or maps the whole file, checks the endianness byte; returns the number of bytes
    size_t load(size_t min_bytes) {
        release();
        is_read = true;
        FILE* fp = std::fopen(filename.c_str(), "rb");
        if (!fp) throw core_error("Cannot open file " + filename);
        std::fseek(fp, 0, SEEK_END);
        const size_t total = (size_t)std::ftell(fp);
        std::fseek(fp, 0, SEEK_SET);
        if (read_mode == 1) {
            std::fclose(fp);
            const int fd = open(filename.c_str(), O_RDONLY);
            if (fd == -1) throw core_error("open failed.");
            void* addr = mmap(nullptr, total, PROT_READ, MAP_PRIVATE | MAP_NORESERVE | MAP_POPULATE, fd, 0);
            close(fd);
            if (addr == MAP_FAILED) throw core_error("mmap failed.");
            map_addr = addr; map_bytes = total; buf = static_cast<const char*>(addr);
        } else {
            owned.resize(total);
            const size_t got = std::fread(owned.data(), 1, total, fp);
            std::fclose(fp);
            if (got != total) throw core_error("Could not read the whole file into buffer.");
            buf = owned.data();
        }
        buf_bytes = total;
        if (total < min_bytes) throw core_error("File is too short to be a .snpdat file.");
        if ((buf[0] != 0) != big_endian())
            throw core_error("Endianness is inconsistent! Regenerate the file on a machine with the same endianness.");
        return total;
    }
    // A crafted / truncated header must not be able to overflow a size computation or point outside the buffer: the counts are checked
    // against the file size BEFORE anything is multiplied, the column offsets must start behind the preamble, be non-decreasing and end
    // inside the file.
    static void check_counts(uint64_t total, uint64_t idx, uint64_t n_items, uint64_t bytes_per_item, uint64_t n_off) {
        const uint64_t left = total > idx ? total - idx : 0;
        if (bytes_per_item && n_items > left / bytes_per_item) throw core_error("File is too short for its header.");
        const uint64_t left2 = left - n_items * bytes_per_item;
        if (n_off > left2 / 8) throw core_error("File is too short for its header.");
    }
    static void check_offsets(const std::vector<uint64_t>& outer, uint64_t preamble, uint64_t total) {
        if (outer.empty() || outer[0] < preamble) throw core_error("Column offsets point into the header.");
        for (size_t j = 0; j + 1 < outer.size(); ++j)
            if (outer[j] > outer[j + 1]) throw core_error("Column offsets are not monotone.");
        if (outer.back() > total) throw core_error("Column offsets point past the end of the file.");
    }
    // bounds-checked cursor over one column's bytes [lo, hi)
    struct Cursor {
        const char* base; uint64_t pos, hi;
        void need(uint64_t n) const { if (n > hi - pos) throw core_error("malformed file (a chunk list runs past the end of its column)."); }
        template <class V> V get() { need(sizeof(V)); V v; std::memcpy(&v, base + pos, sizeof(V)); pos += sizeof(V); return v; }
        void seek(uint64_t lo, uint64_t off) { if (off > hi - lo) throw core_error("malformed file (an offset points outside its column)."); pos = lo + off; }
    };
    // walks one chunk list starting at `cur`: f(row); rows must be < n_rows
    template <class F> static void walk_chunks(Cursor cur, uint64_t n_rows, F f) {
        const uint32_t n_chunks = cur.get<uint32_t>();
        for (uint32_t k = 0; k < n_chunks; ++k) {
            const uint64_t base = (uint64_t)cur.get<uint32_t>() * kChunk;
            const unsigned cnt = (unsigned)cur.get<uint8_t>() + 1u;
            cur.need(cnt);
            for (unsigned e = 0; e < cnt; ++e) {
                const uint64_t row = base + (uint8_t)cur.base[cur.pos + e];
                if (row >= n_rows) throw core_error("malformed file (a row index is out of range).");
                f(row);
            }
            cur.pos += cnt;
        }
    }
};

struct SnpUnphasedIO : SnpFileBase {
    static constexpr int kCategories = 3;
    uint64_t rows = 0, snps = 0;
    std::vector<uint64_t> nnz, nnm, outer; std::vector<double> impute;
    using SnpFileBase::SnpFileBase;

    // io_snp_base.ipp:20-84 + io_snp_unphased.ipp:9-41
    size_t read() {
        const size_t total = load(1 + 2 * sizeof(uint64_t));
        size_t idx = 1;
        rows = rd<uint64_t>(buf + idx); idx += 8;
        snps = rd<uint64_t>(buf + idx); idx += 8;
        if (snps == UINT64_MAX) throw core_error("File is too short for its header.");
        check_counts(total, idx, snps, 24, snps + 1);
        nnz.resize(snps); std::memcpy(nnz.data(), buf + idx, 8 * snps); idx += 8 * snps;
        nnm.resize(snps); std::memcpy(nnm.data(), buf + idx, 8 * snps); idx += 8 * snps;
        impute.resize(snps); std::memcpy(impute.data(), buf + idx, 8 * snps); idx += 8 * snps;
        outer.resize(snps + 1); std::memcpy(outer.data(), buf + idx, 8 * (snps + 1)); idx += 8 * (snps + 1);
        check_offsets(outer, idx, total);
        return total;
    }

    // Walks category c of column j: f(row).  Every offset and count is checked against the column's byte range and the row count.
    template <class F> void for_each(uint64_t j, int c, F f) const {
        Cursor cur{buf, outer[j], outer[j + 1]};
        cur.need(8 * kCategories);
        const uint64_t off = rd<uint64_t>(buf + outer[j] + 8 * c);
        cur.seek(outer[j], off);
        walk_chunks(cur, rows, f);
    }

    // io_snp_unphased.ipp:43-69: (n, p) row-major int8, missing = -9
    void to_dense(int8_t* out) const {
        need_read();
        std::memset(out, 0, (size_t)rows * snps);
        for (uint64_t j = 0; j < snps; ++j)
            for (int c = 0; c < kCategories; ++c) {
                const int8_t val = (c == 0) ? (int8_t)-9 : (int8_t)c;
                for_each(j, c, [&](uint64_t i) { out[i * snps + j] = val; });
            }
    }

    // io_snp_unphased.ipp:72-302.  calldata: column-major (n, p) int8 (negative = missing); impute (p,) in/out
    // (filled with the column means of the non-missing entries when impute_method == "mean", CORE/io/utils.hpp:10-31, 71-97).
    size_t write(const int8_t* calldata, uint64_t n, uint64_t p, const std::string& impute_method, double* impute_io, size_t impute_len) const {
        const uint64_t max_chunks = (n + kChunk - 1) / kChunk;
        if (max_chunks >= (1ull << 32)) throw core_error("calldata dimensions are too large! ");
        if (impute_method != "mean" && impute_method != "user") throw core_error("Invalid impute method type: " + impute_method);
        if (impute_len != p) throw core_error("impute must have length equal to the number of columns of the matrix.");
        std::vector<uint64_t> v_nnz(p), v_nnm(p), col_bytes(p);
        bool bad_value = false;
        // pass 1: per-column statistics and encoded size
        for (uint64_t j = 0; j < p; ++j) {
            const int8_t* col = calldata + j * n;
            uint64_t sum = 0, miss = 0, nz = 0, bytes = 3 * 8 + 3 * 4;
            for (uint64_t k0 = 0; k0 < n; k0 += kChunk) {
                const uint64_t k1 = std::min(n, k0 + kChunk);
                unsigned cnt[3] = {0, 0, 0};
                for (uint64_t i = k0; i < k1; ++i) {
                    const int8_t x = col[i];
                    if (x > 2) { bad_value = true; continue; }
                    if (x < 0) { ++cnt[0]; ++miss; } else if (x > 0) { ++cnt[x]; sum += (uint64_t)x; }
                    nz += (x != 0);
                }
                for (int c = 0; c < 3; ++c) if (cnt[c]) bytes += 4 + 1 + cnt[c];
            }
            v_nnz[j] = nz; v_nnm[j] = n - miss; col_bytes[j] = bytes;
            if (impute_method == "mean") impute_io[j] = (double)sum / (double)std::max<uint64_t>(n - miss, 1);
        }
        if (bad_value) throw core_error("Detected a value greater than > 2. Make sure calldata only contains values <= 2. ");
        const size_t preamble = 1 + 2 * 8 + p * 8 * 3 + (p + 1) * 8;
        std::vector<uint64_t> v_outer(p + 1);
        v_outer[0] = preamble;
        for (uint64_t j = 0; j < p; ++j) v_outer[j + 1] = v_outer[j] + col_bytes[j];
        std::vector<char> out(v_outer[p]);
        size_t idx = 0;
        out[idx++] = (char)big_endian();
        std::memcpy(&out[idx], &n, 8); idx += 8;
        std::memcpy(&out[idx], &p, 8); idx += 8;
        if (p) {
            std::memcpy(&out[idx], v_nnz.data(), 8 * p); idx += 8 * p;
            std::memcpy(&out[idx], v_nnm.data(), 8 * p); idx += 8 * p;
            std::memcpy(&out[idx], impute_io, 8 * p); idx += 8 * p;
        }
        std::memcpy(&out[idx], v_outer.data(), 8 * (p + 1));
        // pass 2: emit the chunk lists, category by category
        for (uint64_t j = 0; j < p; ++j) {
            const int8_t* col = calldata + j * n;
            char* base = out.data() + v_outer[j];
            uint64_t pos = 3 * 8;
            for (int c = 0; c < 3; ++c) {
                std::memcpy(base + 8 * c, &pos, 8);
                char* n_chunks_at = base + pos; pos += 4;
                uint32_t n_chunks = 0;
                for (uint64_t k = 0; k < max_chunks; ++k) {
                    const uint64_t k0 = k * kChunk, k1 = std::min(n, k0 + kChunk);
                    char* hdr = base + pos; unsigned cnt = 0;
                    for (uint64_t i = k0; i < k1; ++i) {
                        const int8_t x = col[i];
                        const bool hit = (c == 0) ? (x < 0) : (x == (int8_t)c);
                        if (hit) { hdr[5 + cnt] = (char)(uint8_t)(i - k0); ++cnt; }
                    }
                    if (cnt) {
                        const uint32_t k32 = (uint32_t)k; std::memcpy(hdr, &k32, 4);
                        hdr[4] = (char)(uint8_t)(cnt - 1);
                        pos += 5 + cnt; ++n_chunks;
                    }
                }
                std::memcpy(n_chunks_at, &n_chunks, 4);
            }
            if (pos != col_bytes[j]) throw core_error("Column index certificate does not match expected size. This is likely a bug in the code. Please report it! ");
        }
        FILE* fp = std::fopen(filename.c_str(), "wb");
        if (!fp) throw core_error("Cannot open file " + filename);
        const size_t wrote = std::fwrite(out.data(), 1, out.size(), fp);
        std::fclose(fp);
        if (wrote != out.size()) throw core_error("Could not write the full buffer.");
        return wrote;
    }
};

// -------------------------------------------------------------------------------------------- phased, ancestry-labelled genotypes
// `IOSNPPhasedAncestry` (CORE/io/io_snp_phased_ancestry.hpp, .ipp:9-363).  X is (n, s*A): column j*A + a counts, over the two
// haplotypes of SNP j, those that carry the mutation AND are labelled with ancestry a -- entries 0 / 1 / 2, no missing values.
// `.snpdat` layout (ipp:170-186, 268-342), little endian:
//   [endian:1B][n:u64][s:u64][A:u8][nnz0:u64 x s*A][nnz1:u64 x s*A][outer:u64 x (s+1)]
//   SNP j at byte outer[j]: [A x u64 offsets of the ancestry blocks, relative to the SNP start]
//     ancestry block: [2 x u64 offsets of the haplotypes, relative to the block start]
//       haplotype: [n_chunks:u32] { [chunk_idx:u32][nnz-1:u8][row_in_chunk:u8 x nnz] }, chunk = 256 rows
// On the device the matrix is the SAME 2-bit packed storage as snp_unphased (codes 0 / 1 / 2, code 3 never occurs), so every
// kernel and the whole solver path are shared; only the unpack kernel differs (it ADDS one per haplotype hit).
struct SnpPhasedAncestryIO : SnpFileBase {
    uint64_t rows = 0, snps = 0, ancestries = 0, cols = 0;
    std::vector<uint64_t> nnz0, nnz1, outer;
    using SnpFileBase::SnpFileBase;

    size_t read() {                                                   // io_snp_base.ipp:20-84 + io_snp_phased_ancestry.ipp:9-43
        const size_t total = load(18);
        size_t idx = 1;
        rows = rd<uint64_t>(buf + idx); idx += 8;
        snps = rd<uint64_t>(buf + idx); idx += 8;
        ancestries = (uint64_t)(uint8_t)buf[idx]; idx += 1;
        if (snps == UINT64_MAX || (ancestries && snps > UINT64_MAX / ancestries)) throw core_error("File is too short for its header.");
        cols = snps * ancestries;
        check_counts(total, idx, cols, 16, snps + 1);
        nnz0.resize(cols); std::memcpy(nnz0.data(), buf + idx, 8 * cols); idx += 8 * cols;
        nnz1.resize(cols); std::memcpy(nnz1.data(), buf + idx, 8 * cols); idx += 8 * cols;
        outer.resize(snps + 1); std::memcpy(outer.data(), buf + idx, 8 * (snps + 1)); idx += 8 * (snps + 1);
        check_offsets(outer, idx, total);
        return total;
    }
    template <class F> void for_each(uint64_t j, uint64_t a, int hap, F f) const {
        Cursor cur{buf, outer[j], outer[j + 1]};
        cur.need(8 * ancestries);
        const uint64_t off_a = rd<uint64_t>(buf + outer[j] + 8 * a);
        cur.seek(outer[j], off_a);
        const uint64_t blk = cur.pos;
        cur.need(16);
        const uint64_t off_h = rd<uint64_t>(buf + blk + 8 * hap);
        cur.seek(blk, off_h);
        walk_chunks(cur, rows, f);
    }
    void to_dense(int8_t* out) const {                                // ipp:45-71: (n, s*A) row-major
        need_read();
        std::memset(out, 0, (size_t)rows * cols);
        for (uint64_t j = 0; j < snps; ++j)
            for (uint64_t a = 0; a < ancestries; ++a)
                for (int hap = 0; hap < 2; ++hap) for_each(j, a, hap, [&](uint64_t i) { out[i * cols + j * ancestries + a] += 1; });
    }
    // ipp:73-363.  calldata, anc: column-major (n, 2 s) int8; calldata in {0, 1}, anc in [0, A)
    size_t write(const int8_t* calldata, const int8_t* anc, uint64_t n, uint64_t two_s, uint64_t A) const {
        if (two_s % 2) throw core_error("calldata and ancestries must have shape (n, 2*s).");
        if (A >= kChunk) throw core_error("Number of ancestries A must be < 256.");
        const uint64_t s_ = two_s / 2, max_chunks = (n + kChunk - 1) / kChunk;
        if (max_chunks >= (1ull << 32)) throw core_error("calldata dimensions are too large! ");
        for (uint64_t k = 0; k < two_s * n; ++k) {
            if (anc[k] < 0 || (uint64_t)anc[k] >= A) throw core_error("Detected an ancestry not in the range [0, A). Make sure ancestries only contains values in [0, A). ");
            if (calldata[k] != 0 && calldata[k] != 1) throw core_error("Detected a non-binary value. Make sure calldata only contains 0 or 1 values. ");
        }
        // one haplotype list: [n_chunks] + per non-empty chunk [idx][cnt-1][rows]
        auto emit = [&](const int8_t* cal, const int8_t* an, int8_t a, std::vector<char>& dst) -> uint64_t {
            const size_t at = dst.size(); dst.resize(at + 4);
            uint32_t n_chunks = 0; uint64_t hits = 0;
            for (uint64_t k = 0; k < max_chunks; ++k) {
                const uint64_t k0 = k * kChunk, k1 = std::min(n, k0 + kChunk);
                size_t hdr = 0; unsigned cnt = 0;
                for (uint64_t i = k0; i < k1; ++i) {
                    if (an[i] == a && cal[i] == 1) {
                        if (!cnt) { hdr = dst.size(); dst.resize(hdr + 5); }
                        dst.push_back((char)(uint8_t)(i - k0)); ++cnt;
                    }
                }
                if (cnt) { const uint32_t k32 = (uint32_t)k; std::memcpy(&dst[hdr], &k32, 4); dst[hdr + 4] = (char)(uint8_t)(cnt - 1); ++n_chunks; hits += cnt; }
            }
            std::memcpy(&dst[at], &n_chunks, 4);
            return hits;
        };
        std::vector<uint64_t> v_nnz0(s_ * A), v_nnz1(s_ * A), v_outer(s_ + 1);
        std::vector<std::vector<char>> snp_bytes(s_);
        for (uint64_t j = 0; j < s_; ++j) {
            std::vector<char>& sb = snp_bytes[j];
            sb.assign(8 * A, 0);
            for (uint64_t a = 0; a < A; ++a) {
                const uint64_t blk = sb.size();
                std::memcpy(&sb[8 * a], &blk, 8);
                sb.resize(blk + 16);
                for (int hap = 0; hap < 2; ++hap) {
                    const uint64_t rel = sb.size() - blk;
                    std::memcpy(&sb[blk + 8 * hap], &rel, 8);
                    const uint64_t hits = emit(calldata + (2 * j + hap) * n, anc + (2 * j + hap) * n, (int8_t)a, sb);
                    (hap == 0 ? v_nnz0 : v_nnz1)[j * A + a] = hits;
                }
            }
        }
        const size_t preamble = 1 + 16 + 1 + 16 * s_ * A + 8 * (s_ + 1);
        v_outer[0] = preamble;
        for (uint64_t j = 0; j < s_; ++j) v_outer[j + 1] = v_outer[j] + snp_bytes[j].size();
        std::vector<char> out(preamble);
        size_t idx = 0;
        out[idx++] = (char)big_endian();
        std::memcpy(&out[idx], &n, 8); idx += 8;
        std::memcpy(&out[idx], &s_, 8); idx += 8;
        out[idx++] = (char)(uint8_t)A;
        if (s_ * A) { std::memcpy(&out[idx], v_nnz0.data(), 8 * s_ * A); idx += 8 * s_ * A; std::memcpy(&out[idx], v_nnz1.data(), 8 * s_ * A); idx += 8 * s_ * A; }
        std::memcpy(&out[idx], v_outer.data(), 8 * (s_ + 1));
        FILE* fp = std::fopen(filename.c_str(), "wb");
        if (!fp) throw core_error("Cannot open file " + filename);
        size_t wrote = std::fwrite(out.data(), 1, out.size(), fp);
        for (uint64_t j = 0; j < s_; ++j) wrote += std::fwrite(snp_bytes[j].data(), 1, snp_bytes[j].size(), fp);
        std::fclose(fp);
        if (wrote != v_outer[s_]) throw core_error("Could not write the full buffer.");
        return wrote;
    }
};

// ============================================================================================ device kernels
__device__ __forceinline__ uint64_t snp_rd_u64(const uint8_t* q) { uint64_t u = 0; for (int b = 7; b >= 0; --b) u = (u << 8) | q[b]; return u; }
__device__ __forceinline__ uint32_t snp_rd_u32(const uint8_t* q) { return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24); }

// `.snpdat` bytes -> 2-bit codes.  One warp per (column, category): the lanes walk the chunk list together (the header bytes
// are a broadcast load) and OR the codes of up to 256 entries per chunk into the zeroed packed column.  Rows outside
// [row_lo, row_hi) are skipped (row sharding: every rank reads the same file and keeps its rows).
__global__ void __launch_bounds__(256)
snpdat_unpack_kernel(const uint8_t* __restrict__ file, const uint64_t* __restrict__ outer, int64_t j0, int64_t ncols, int64_t col_base,
                     int64_t n_total, int64_t row_lo, int64_t row_hi, uint32_t* __restrict__ packed, int64_t ldw, int* __restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= ncols * 3) return;
    const int64_t jl = item / 3; const int c = (int)(item - jl * 3);
    const uint8_t* col = file + (outer[j0 + jl] - (uint64_t)col_base);
    const uint8_t* col_end = file + (outer[j0 + jl + 1] - (uint64_t)col_base);      // a malformed file must not send the walk out of its column
    if (col + 24 > col_end) { *err = 2; return; }
    const uint64_t coff = snp_rd_u64(col + 8 * c);
    if (coff > (uint64_t)(col_end - col) || col + coff + 4 > col_end) { *err = 2; return; }
    const uint8_t* q = col + coff;
    const uint32_t n_chunks = snp_rd_u32(q); q += 4;
    const uint32_t code = (c == 0) ? 3u : (uint32_t)c;
    uint32_t* dst = packed + (j0 + jl) * ldw;
    for (uint32_t k = 0; k < n_chunks; ++k) {
        if (q + 5 > col_end || q + 5 + q[4] + 1 > col_end) { *err = 2; return; }
        const int64_t base = (int64_t)snp_rd_u32(q) * 256;
        const int cnt = (int)q[4] + 1;
        for (int e = lane; e < cnt; e += 32) {
            const int64_t row = base + q[5 + e];
            if (row >= n_total) { *err = 1; continue; }
            if (row < row_lo || row >= row_hi) continue;
            const int64_t r = row - row_lo;
            atomicOr(dst + (r >> 4), code << (2 * (int)(r & 15)));
        }
        q += 5 + cnt;
    }
}

// Phased-ancestry chunk lists -> 2-bit counts.  One warp per (SNP, ancestry, haplotype): every hit ADDS one to the 2-bit field of
// column SNP * A + ancestry (two haplotypes: at most 2, no carry into the neighbouring field).
__global__ void __launch_bounds__(256)
snpdat_phased_unpack_kernel(const uint8_t* __restrict__ file, const uint64_t* __restrict__ outer, int64_t j0, int64_t nsnps, int64_t col_base, int A,
                            int64_t n_total, int64_t row_lo, int64_t row_hi, uint32_t* __restrict__ packed, int64_t ldw, int* __restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= nsnps * A * 2) return;
    const int64_t jl = item / (2 * A); const int a = (int)((item - jl * 2 * A) >> 1), hap = (int)(item & 1);
    const uint8_t* snp = file + (outer[j0 + jl] - (uint64_t)col_base);
    const uint8_t* snp_end = file + (outer[j0 + jl + 1] - (uint64_t)col_base);
    if (snp + 8 * A > snp_end) { *err = 2; return; }
    const uint64_t boff = snp_rd_u64(snp + 8 * a);
    if (boff > (uint64_t)(snp_end - snp) || snp + boff + 16 > snp_end) { *err = 2; return; }
    const uint8_t* blk = snp + boff;
    const uint64_t hoff = snp_rd_u64(blk + 8 * hap);
    if (hoff > (uint64_t)(snp_end - blk) || blk + hoff + 4 > snp_end) { *err = 2; return; }
    const uint8_t* q = blk + hoff;
    const uint32_t n_chunks = snp_rd_u32(q); q += 4;
    uint32_t* dst = packed + ((j0 + jl) * A + a) * ldw;
    for (uint32_t k = 0; k < n_chunks; ++k) {
        if (q + 5 > snp_end || q + 5 + q[4] + 1 > snp_end) { *err = 2; return; }
        const int64_t base = (int64_t)snp_rd_u32(q) * 256;
        const int cnt = (int)q[4] + 1;
        for (int e = lane; e < cnt; e += 32) {
            const int64_t row = base + q[5 + e];
            if (row >= n_total) { *err = 1; continue; }
            if (row < row_lo || row >= row_hi) continue;
            const int64_t r = row - row_lo;
            atomicAdd(dst + (r >> 4), 1u << (2 * (int)(r & 15)));
        }
        q += 5 + cnt;
    }
}

// int8 calldata (column-major, negative = missing) -> 2-bit codes; one thread per 16 rows
__global__ void snp_pack_kernel(const int8_t* __restrict__ calldata, int64_t n, int64_t p, uint32_t* __restrict__ packed, int64_t ldw, int* __restrict__ err)
{
    const int64_t j = blockIdx.x;                     // columns on grid.x (no 65535 limit), row blocks on grid.y
    for (int64_t w = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; w < ldw; w += (int64_t)gridDim.y * blockDim.x) {
        uint32_t word = 0;
        for (int k = 0; k < 16; ++k) {
            const int64_t i = w * 16 + k;
            if (i >= n) break;
            const int8_t x = calldata[j * n + i];
            if (x > 2) { *err = 1; continue; }
            word |= (uint32_t)(x < 0 ? 3 : x) << (2 * k);
        }
        packed[j * ldw + w] = word;
    }
}

template <class T> __device__ __forceinline__ T snp_value(uint32_t code, T imp) {
    return (code & 2u) ? ((code & 1u) ? imp : T(2)) : ((code & 1u) ? T(1) : T(0));
}
// The four values a genotype code of column j stands for: {0, 1, 2, impute[j]}, or their standardized images (x - c_j) / s_j when the
// matrix is a `matrix.standardize` view of the packed bits (center != nullptr; same arithmetic as the dense standardize kernel).
template <class T> struct SnpVals { T v0, v1, v2, v3; };
template <class T> __device__ __forceinline__ SnpVals<T> snp_vals(const T* impute, const T* center, const T* scale, int64_t j) {
    SnpVals<T> r{T(0), T(1), T(2), impute[j]};
    if (center) { const T c = center[j], sc = scale[j]; r.v0 = (r.v0 - c) / sc; r.v1 = (r.v1 - c) / sc; r.v2 = (r.v2 - c) / sc; r.v3 = (r.v3 - c) / sc; }
    return r;
}
template <class T> __device__ __forceinline__ T snp_value4(uint32_t code, const SnpVals<T>& t) {
    return (code & 2u) ? ((code & 1u) ? t.v3 : t.v2) : ((code & 1u) ? t.v1 : t.v0);
}

// Decodes `count` columns starting at logical column j0 into dense columns out[c * ld + i] (pad rows = 0 since their code is 0).
template <class T>
__global__ void snp_decode_kernel(const uint32_t* __restrict__ packed, int64_t ldw, const T* __restrict__ impute, const T* __restrict__ center,
                                  const T* __restrict__ scale, int64_t n, int64_t j0, int count, T* __restrict__ out, int64_t ld)
{
    const int c = blockIdx.x;                          // columns on grid.x, row blocks on grid.y
    const SnpVals<T> tv = snp_vals<T>(impute, center, scale, j0 + c);
    const uint32_t* src = packed + (j0 + c) * ldw;
    T* dst = out + (int64_t)c * ld;
    for (int64_t w = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; w < ldw; w += (int64_t)gridDim.y * blockDim.x) {
        const uint32_t word = src[w];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            T x[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = (w * 16 + 4 * k4 + k < n) ? snp_value4<T>((word >> (2 * (4 * k4 + k))) & 3u, tv) : T(0);   // pad rows stay 0
            if (sizeof(T) == 4) *reinterpret_cast<float4*>(dst + w * 16 + 4 * k4) = make_float4((float)x[0], (float)x[1], (float)x[2], (float)x[3]);
            else {
                *reinterpret_cast<double2*>(dst + w * 16 + 4 * k4) = make_double2((double)x[0], (double)x[1]);
                *reinterpret_cast<double2*>(dst + w * 16 + 4 * k4 + 2) = make_double2((double)x[2], (double)x[3]);
            }
        }
    }
}

// Transposed GEMV on the packed bits (the `mul` / `sq_mul` of MatrixNaiveSNPUnphased, and of kron(X, I_K) for multi-response):
//   out_part[(rb * q + c) * K + l] = sum_{i in row block rb} f(X[i, j0+c]) * v[i, l] * w[i, l]      (SQ: X^2 * w; w == nullptr: 1)
// Every lane keeps the products v*w of its R rows (x KP classes) in registers for the whole kernel and walks the columns: per
// column a warp reads 8*R contiguous bytes (R/4 bytes per lane), decodes R genotypes per lane and issues R*KP FMAs.  The
// per-lane partials of a sub-batch of 32/KP columns go through a padded shared-memory transpose (conflict free both ways) so that
// lane t ends up with the warp total of value t; once per 32 columns the 4 warps (4 consecutive row tiles) are added in double.
// grid = (column chunks, row blocks of 4*32*R rows).
constexpr int kSnpGemvThreads = 128;
template <int KP> __host__ __device__ constexpr int snp_gemv_rows_per_lane() { return (KP <= 2) ? 32 : (KP == 4 ? 16 : (KP == 8 ? 8 : 4)); }
// per-warp scratch (elements): the 32 x 33 transpose of the partials, also used to stage the warp's v*w tile (32 lanes x (R*KP + 4 pad))
template <int KP> __host__ __device__ constexpr int snp_gemv_warp_scratch() {
    return (32 * (snp_gemv_rows_per_lane<KP>() * KP + 4) > 32 * 33) ? 32 * (snp_gemv_rows_per_lane<KP>() * KP + 4) : 32 * 33;
}
template <class T, int KP> __host__ __device__ constexpr size_t snp_gemv_smem_bytes() {
    return sizeof(double) * (kSnpGemvThreads / 32) * 32 * KP + sizeof(T) * (kSnpGemvThreads / 32) * snp_gemv_warp_scratch<KP>();
}
// genotype r of a 32-bit word of codes, straight from constant-mask bit tests (no shifts): 0 / 1 / 2 / impute
template <class T, int r> __device__ __forceinline__ T snp_pick(uint32_t word, const SnpVals<T>& t) {
    const bool b0 = (word & (1u << (2 * r))) != 0, b1 = (word & (2u << (2 * r))) != 0;
    const T lo = b0 ? t.v1 : t.v0, hi = b0 ? t.v3 : t.v2;
    return b1 ? hi : lo;
}
// STD: the matrix is a standardize view (values (k - c_j) / s_j from registers); otherwise the values 0 / 1 / 2 are immediates.
template <class T, int KP, bool SQ, bool STD>
__global__ void __launch_bounds__(kSnpGemvThreads)
snp_gemv_t_kernel(const uint32_t* __restrict__ packed, int64_t ldw, int64_t n_pad, const T* __restrict__ impute, const T* __restrict__ center,
                  const T* __restrict__ scale, int64_t j0, int q, int cols_per_cta,
                  int tiles_per_cta, int K, const T* __restrict__ v, const T* __restrict__ w, double* __restrict__ out_part)
{
    constexpr int R = snp_gemv_rows_per_lane<KP>();                              // rows per lane
    constexpr int NB = 32 / KP;                                                  // columns per warp-level reduction (sub-batch)
    constexpr int NW = kSnpGemvThreads / 32;
    constexpr int NV = 32 * KP;                                                  // values per block-level reduction: 32 columns x KP classes
    extern __shared__ __align__(16) unsigned char s_snp_raw[];
    double (*s_tot)[NV] = reinterpret_cast<double (*)[NV]>(s_snp_raw);                                // [NW][32 columns * KP]
    constexpr int WS = snp_gemv_warp_scratch<KP>();
    T (*s_acc)[WS] = reinterpret_cast<T (*)[WS]>(s_snp_raw + sizeof(double) * NW * NV);                 // [NW][warp scratch]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c_begin = blockIdx.x * cols_per_cta;
    const int c_end = min(q, c_begin + cols_per_cta);
    for (int cb = c_begin; cb < c_end; cb += 32) {
        for (int k = lane; k < NV; k += 32) s_tot[warp][k] = 0;
        __syncwarp();
        // the CTA walks `tiles_per_cta` row tiles (NW * 32 * R rows each) for the same 32 columns: the number of partial rows the
        // final reduction has to read stays small at large n (the products v*w of a tile are re-read from L2 per tile: < 2 % of the work)
#pragma unroll 1
        for (int tile = 0; tile < tiles_per_cta; ++tile) {
        const int64_t wrow0 = (((int64_t)blockIdx.y * tiles_per_cta + tile) * NW + warp) * (32 * R);    // first row of this warp's tile
        if (wrow0 >= n_pad) break;                                                                     // warp uniform
        const int64_t row0 = wrow0 + (int64_t)lane * R;                                                 // first row of this lane
        const bool live = row0 < n_pad;                                                                 // n_pad % 32 == 0 and R | 32
        // the warp's tile of v*w (32*R rows x K classes, contiguous in memory) comes in with coalesced loads and goes through the
        // warp's scratch so that every lane ends up with ITS R rows x KP classes in registers (row stride padded by 4: the 16-byte
        // reads below are bank-conflict free)
        T vw[R * KP];
        {
            constexpr int RS = R * KP + 4;
            T* st = s_acc[warp];
            if (K != KP) { for (int k = lane; k < 32 * RS; k += 32) st[k] = 0; __syncwarp(); }
            const int64_t e0 = wrow0 * K, e_end = n_pad * K;
            const int cnt = 32 * R * K;
            constexpr int VNS = 16 / sizeof(T);
            if (K == KP) {                                   // power-of-two class count: 16-byte loads / stores, shifts instead of divisions
                constexpr int RK = R * KP;
                for (int g = lane * VNS; g < 32 * RK; g += 32 * VNS) {
                    T a[VNS];
#pragma unroll
                    for (int k = 0; k < VNS; ++k) a[k] = 0;
                    if (e0 + g < e_end) {
                        if (SQ || w) vec_load<T>(w + e0 + g, a);
                        if (!SQ) {
                            T b[VNS];
                            vec_load<T>(v + e0 + g, b);
#pragma unroll
                            for (int k = 0; k < VNS; ++k) a[k] = w ? a[k] * b[k] : b[k];
                        }
                    }
                    vec_store<T>(st + (g / RK) * RS + (g % RK), a);
                }
            } else
            for (int g = lane; g < cnt; g += 32) {
                T a = 0;
                if (e0 + g < e_end) a = SQ ? w[e0 + g] : (w ? v[e0 + g] * w[e0 + g] : v[e0 + g]);
                const int row = g / K, l = g - row * K;
                st[(row / R) * RS + (row % R) * KP + l] = a;
            }
            __syncwarp();
            constexpr int VNL = 16 / sizeof(T);
#pragma unroll
            for (int k = 0; k < R * KP; k += VNL) {
                if (sizeof(T) == 4) { const float4 t4 = *reinterpret_cast<const float4*>(st + lane * RS + k); vw[k] = t4.x; vw[k + 1] = t4.y; vw[k + 2] = t4.z; vw[k + 3] = t4.w; }
                else { const double2 t2 = *reinterpret_cast<const double2*>(st + lane * RS + k); vw[k] = t2.x; vw[k + 1] = t2.y; }
            }
            __syncwarp();
        }
        const uint8_t* lane_base = reinterpret_cast<const uint8_t*>(packed) + row0 / 4;
#pragma unroll 1
        for (int sub = 0; sub < KP; ++sub) {
            constexpr int PF = NB < 4 ? NB : 4;          // columns whose bits are fetched together (independent loads in flight)
#pragma unroll 1
            for (int cc0 = 0; cc0 < NB; cc0 += PF) {
                uint32_t w0[PF], w1[PF]; SnpVals<T> imp[PF];
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int c = cb + sub * NB + cc0 + u;
                    // a column past the end contributes exact zeros: all its codes are 0 and v0 is 0 (v1..v3 are never picked; the
                    // non-STD values stay compile-time constants on both paths)
                    w0[u] = 0; w1[u] = 0; imp[u] = SnpVals<T>{T(0), STD ? T(0) : T(1), STD ? T(0) : T(2), T(0)};
                    if (c < c_end && live) {
                        if (STD) imp[u] = snp_vals<T>(impute, center, scale, j0 + c);
                        else imp[u] = SnpVals<T>{T(0), T(1), T(2), impute[j0 + c]};
                        const uint8_t* src = lane_base + (j0 + c) * ldw * 4;
                        if (R == 32) { const uint2 q2 = *reinterpret_cast<const uint2*>(src); w0[u] = q2.x; w1[u] = q2.y; }
                        else if (R == 16) w0[u] = *reinterpret_cast<const uint32_t*>(src);
                        else if (R == 8) w0[u] = *reinterpret_cast<const uint16_t*>(src);
                        else w0[u] = *src;
                    }
                }
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    T acc[KP];
#pragma unroll
                    for (int l = 0; l < KP; ++l) acc[l] = 0;
                    T x[R];
#define AB_SNP_PICK(r) if (r < R) x[r < R ? r : 0] = (r < 16) ? snp_pick<T, (r & 15)>(w0[u], imp[u]) : snp_pick<T, (r & 15)>(w1[u], imp[u]);
                    AB_SNP_PICK(0) AB_SNP_PICK(1) AB_SNP_PICK(2) AB_SNP_PICK(3) AB_SNP_PICK(4) AB_SNP_PICK(5) AB_SNP_PICK(6) AB_SNP_PICK(7)
                    AB_SNP_PICK(8) AB_SNP_PICK(9) AB_SNP_PICK(10) AB_SNP_PICK(11) AB_SNP_PICK(12) AB_SNP_PICK(13) AB_SNP_PICK(14) AB_SNP_PICK(15)
                    AB_SNP_PICK(16) AB_SNP_PICK(17) AB_SNP_PICK(18) AB_SNP_PICK(19) AB_SNP_PICK(20) AB_SNP_PICK(21) AB_SNP_PICK(22) AB_SNP_PICK(23)
                    AB_SNP_PICK(24) AB_SNP_PICK(25) AB_SNP_PICK(26) AB_SNP_PICK(27) AB_SNP_PICK(28) AB_SNP_PICK(29) AB_SNP_PICK(30) AB_SNP_PICK(31)
#undef AB_SNP_PICK
                    if (KP == 1) {                  // two independent FMA chains per column
                        T a0 = 0, a1 = 0;
#pragma unroll
                        for (int r = 0; r < R; r += 2) {
                            const T x0 = SQ ? x[r] * x[r] : x[r], x1 = SQ ? x[r + 1] * x[r + 1] : x[r + 1];
                            a0 += x0 * vw[r]; a1 += x1 * vw[r + 1];
                        }
                        acc[0] = a0 + a1;
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const T xr = SQ ? x[r] * x[r] : x[r];
#pragma unroll
                            for (int l = 0; l < KP; ++l) acc[l] += xr * vw[r * KP + l];
                        }
                    }
#pragma unroll
                    for (int l = 0; l < KP; ++l) s_acc[warp][((cc0 + u) * KP + l) * 33 + lane] = acc[l];
                }
            }
            __syncwarp();
            T tot = 0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) tot += s_acc[warp][lane * 33 + k];
            s_tot[warp][sub * 32 + lane] += (double)tot;
            __syncwarp();
        }
        }   // row tiles
        __syncthreads();
        for (int idx = tid; idx < NV; idx += kSnpGemvThreads) {
            double sum = 0;
#pragma unroll
            for (int wq = 0; wq < NW; ++wq) sum += s_tot[wq][idx];
            const int c = cb + idx / KP, l = idx % KP;
            if (c < c_end && l < K) out_part[((size_t)blockIdx.y * q + c) * K + l] = sum;
        }
        __syncthreads();
    }
}

// Random genotype matrix generated in HBM (ad.data.snp_unphased proportions, PY/data.py:312-359): each entry is 1 w.p. one_ratio, 2 w.p.
// two_ratio, then masked as missing w.p. missing_ratio; Philox(seed, subsequence = column, offset = global row) so that every
// row sharding sees the same matrix.  counts[j] = {ones, twos, missing} over the non-masked / masked entries of this shard.
__global__ void snp_fill_random_kernel(uint32_t* __restrict__ packed, int64_t ldw, int64_t n, int64_t p, unsigned long long seed, int64_t row_offset,
                                       float one_ratio, float two_ratio, float missing_ratio, unsigned long long* __restrict__ counts)
{
    const int64_t j = blockIdx.x;                     // columns on grid.x, row blocks on grid.y
    unsigned long long c1 = 0, c2 = 0, c3 = 0;
    for (int64_t wd = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; wd < ldw; wd += (int64_t)gridDim.y * blockDim.x) {
        uint32_t word = 0;
        for (int k4 = 0; k4 < 4; ++k4) {
            const int64_t i0 = wd * 16 + 4 * k4;
            if (i0 >= n) break;
            curandStatePhilox4_32_10_t st;
            curand_init(seed, (unsigned long long)j, (unsigned long long)(2 * (row_offset + i0)), &st);
            const float4 u = curand_uniform4(&st), m = curand_uniform4(&st);
            const float uu[4] = {u.x, u.y, u.z, u.w}, mm[4] = {m.x, m.y, m.z, m.w};
            for (int k = 0; k < 4 && i0 + k < n; ++k) {
                uint32_t code = uu[k] <= one_ratio ? 1u : (uu[k] <= one_ratio + two_ratio ? 2u : 0u);
                if (mm[k] <= missing_ratio) code = 3u;
                c1 += code == 1u; c2 += code == 2u; c3 += code == 3u;
                word |= code << (2 * (4 * k4 + k));
            }
        }
        packed[j * ldw + wd] = word;
    }
    for (int o = 16; o > 0; o >>= 1) {
        c1 += __shfl_xor_sync(0xffffffffu, c1, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(counts + 3 * j, c1); atomicAdd(counts + 3 * j + 1, c2); atomicAdd(counts + 3 * j + 2, c3); }
}

} // namespace ab
