"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's SNP unphased storage and matrix operators.

Nothing under ``adelie_b200/`` imports this module; it is the checker for ``tests/`` (the product's reader / writer is host C++
in ``adelie_b200/csrc/snp.cuh`` and its operators are CUDA kernels).

Follows, function by function:
  * ``write_snpdat``   -- IOSNPUnphased::write, adelie_core/io/io_snp_unphased.ipp:72-302 (+ compute_column_mean / compute_nnm /
                          compute_nnz, adelie_core/io/utils.hpp:10-68)
  * ``read_snpdat``    -- IOSNPBase::read (io_snp_base.ipp:20-84) + IOSNPUnphased::read (io_snp_unphased.ipp:9-41) + the chunk
                          iterator (io_snp_base.hpp, IOSNPChunkIterator) + to_dense (io_snp_unphased.ipp:43-69)
  * ``SnpMatrix``      -- MatrixNaiveSNPUnphased (adelie_core/matrix/matrix_naive_snp_unphased.ipp:10-309) through
                          snp_unphased_dot / snp_unphased_axi (adelie_core/matrix/utils.hpp:559-690): category-wise sums
                          ``sum_c f(val_c) * sum_{i in category c} v[i]`` with val_0 = impute[j], val_1 = 1, val_2 = 2.

Parity pin: the reference cannot be run here (Eigen is not vendored), so the format is pinned by (i) the hand-derived byte
string ``tests/golden/snp_unphased_tiny.snpdat.hex`` written from the layout documented in io_snp_unphased.ipp:88-110,
(ii) the expectations of the reference's own test (tests/test_io.py:7-60: nnm, impute, nnz, to_dense, written == read bytes),
(iii) byte-for-byte agreement between this writer and the independent C++ writer of the product.
"""
import struct

import numpy as np

CHUNK = 256
N_CATEGORIES = 3


def column_stats(calldata):
    """impute (mean of the non-missing entries; 0 when all are missing), nnm, nnz per column (io/utils.hpp:10-68)."""
    cd = np.asarray(calldata)
    n = cd.shape[0]
    miss = np.sum(cd < 0, axis=0).astype(np.uint64)
    tot = np.sum(np.where(cd > 0, cd, 0).astype(np.uint64), axis=0)
    impute = tot.astype(np.float64) / np.maximum(np.uint64(n) - miss, np.uint64(1)).astype(np.float64)
    nnm = np.uint64(n) - miss
    nnz = np.sum(cd != 0, axis=0).astype(np.uint64)
    return impute, nnm, nnz


def write_snpdat(calldata, impute="mean"):
    """Returns ``(file_bytes, impute)`` for an (n, p) int8 calldata matrix (negative = missing)."""
    cd = np.asarray(calldata)
    assert cd.dtype == np.int8 and cd.ndim == 2
    n, p = cd.shape
    if np.any(cd > 2):
        raise RuntimeError("adelie_core: Detected a value greater than > 2. Make sure calldata only contains values <= 2. ")
    mean, nnm, nnz = column_stats(cd)
    imp = mean if isinstance(impute, str) else np.asarray(impute, dtype=np.float64)
    cols = []
    for j in range(p):
        col = cd[:, j]
        cats = []
        for c in range(N_CATEGORIES):
            rows = np.flatnonzero(col < 0) if c == 0 else np.flatnonzero(col == c)
            body = bytearray()
            chunk_ids = rows // CHUNK
            uniq, starts = np.unique(chunk_ids, return_index=True)
            ends = list(starts[1:]) + [rows.size]
            for k, s, e in zip(uniq, starts, ends):
                body += struct.pack("<IB", int(k), e - s - 1)
                body += bytes((rows[s:e] - k * CHUNK).astype(np.uint8))
            cats.append(struct.pack("<I", len(uniq)) + bytes(body))
        offs, pos = [], 3 * 8
        for b in cats:
            offs.append(pos)
            pos += len(b)
        cols.append(struct.pack("<3Q", *offs) + b"".join(cats))
    preamble = 1 + 2 * 8 + 3 * 8 * p + 8 * (p + 1)
    outer = np.concatenate([[preamble], preamble + np.cumsum([len(c) for c in cols])]).astype(np.uint64)
    head = struct.pack("<?QQ", False, n, p) + nnz.tobytes() + nnm.tobytes() + imp.astype(np.float64).tobytes() + outer.tobytes()
    return head + b"".join(cols), imp


def read_snpdat(buf):
    """Parses file bytes: dict(rows, snps, nnz, nnm, impute, outer, dense) with dense (n, p) int8, -9 for missing."""
    if buf[0] != 0:
        raise RuntimeError("adelie_core: Endianness is inconsistent! Regenerate the file on a machine with the same endianness.")
    n, p = struct.unpack_from("<QQ", buf, 1)
    idx = 17
    nnz = np.frombuffer(buf, dtype="<u8", count=p, offset=idx); idx += 8 * p
    nnm = np.frombuffer(buf, dtype="<u8", count=p, offset=idx); idx += 8 * p
    impute = np.frombuffer(buf, dtype="<f8", count=p, offset=idx); idx += 8 * p
    outer = np.frombuffer(buf, dtype="<u8", count=p + 1, offset=idx)
    dense = np.zeros((n, p), dtype=np.int8)
    for j in range(p):
        base = int(outer[j])
        for c in range(N_CATEGORIES):
            q = base + struct.unpack_from("<Q", buf, base + 8 * c)[0]
            (n_chunks,) = struct.unpack_from("<I", buf, q); q += 4
            for _ in range(n_chunks):
                k, m1 = struct.unpack_from("<IB", buf, q); q += 5
                rows = k * CHUNK + np.frombuffer(buf, dtype=np.uint8, count=m1 + 1, offset=q).astype(np.int64); q += m1 + 1
                dense[rows, j] = -9 if c == 0 else c
    return dict(rows=n, snps=p, nnz=nnz, nnm=nnm, impute=impute, outer=outer, dense=dense)


def dense_equivalent(calldata, impute, dtype=np.float64):
    """The dense matrix every operator of MatrixNaiveSNPUnphased acts as (T/test_matrix.py:737-740)."""
    cd = np.asarray(calldata)
    return np.where(cd < 0, np.asarray(impute)[None, :], cd).astype(dtype)


class SnpMatrix:
    """Category-wise operators (matrix_naive_snp_unphased.ipp) on the parsed file; arithmetic in ``dtype`` like the reference templates."""
    def __init__(self, parsed, dtype=np.float64):
        self.dtype = np.dtype(dtype).type
        d = parsed["dense"]
        self.n, self.p = d.shape
        self.impute = parsed["impute"].astype(dtype)
        self.cats = [[np.flatnonzero(d[:, j] < 0), np.flatnonzero(d[:, j] == 1), np.flatnonzero(d[:, j] == 2)] for j in range(self.p)]

    def _val(self, j, c):
        return self.impute[j] if c == 0 else self.dtype(c)

    def _dot(self, j, v, unary=lambda x: x):           # utils.hpp:559-625
        s = self.dtype(0)
        for c in range(N_CATEGORIES):
            s = s + np.sum(v[self.cats[j][c]], dtype=self.dtype) * unary(self._val(j, c))
        return s

    def cmul(self, j, v, w):
        return self._dot(j, (v * w).astype(self.dtype))

    def ctmul(self, j, v, out):                        # utils.hpp:628-690
        for c in range(N_CATEGORIES):
            out[self.cats[j][c]] += self.dtype(v) * self._val(j, c)

    def bmul(self, j, q, v, w, out):
        for t in range(q):
            out[t] = self.cmul(j + t, v, w)

    def btmul(self, j, q, v, out):
        for t in range(q):
            self.ctmul(j + t, v[t], out)

    def mul(self, v, w, out):
        self.bmul(0, self.p, v, w, out)

    def sq_mul(self, w, out):
        for t in range(self.p):
            out[t] = self._dot(t, w.astype(self.dtype), lambda x: x * x)

    def cov(self, j, q, sqrt_w, out):                  # matrix_naive_snp_unphased.ipp:170-243
        w = (sqrt_w * sqrt_w).astype(self.dtype)
        for i1 in range(q):
            col1 = np.zeros(self.n, dtype=self.dtype)
            for c in range(N_CATEGORIES):
                col1[self.cats[j + i1][c]] = self._val(j + i1, c)
            for i2 in range(i1 + 1):
                out[i1, i2] = self._dot(j + i1, w, lambda x: x * x) if i1 == i2 else self._dot(j + i2, w * col1)
                out[i2, i1] = out[i1, i2]


# ------------------------------------------------------------------------------------------------------------------
# SNP phased, ancestry (IOSNPPhasedAncestry, adelie_core/io/io_snp_phased_ancestry.ipp:9-363)
def phased_dense(calldata, ancestries, A):
    """The (n, s A) matrix the format stands for: entry (i, j A + a) = #{k in {0, 1}: calldata[i, 2j+k] == 1 and ancestries[i, 2j+k] == a}
    (to_dense, ipp:45-71)."""
    cd = np.asarray(calldata); an = np.asarray(ancestries)
    n, two_s = cd.shape
    out = np.zeros((n, (two_s // 2) * A), dtype=np.int8)
    for k in range(2):
        ii, jj = np.nonzero(cd[:, k::2])
        np.add.at(out, (ii, jj * A + an[:, k::2][ii, jj]), 1)
    return out


def _chunk_list(rows):
    body = bytearray()
    uniq, starts = np.unique(rows // CHUNK, return_index=True)
    ends = list(starts[1:]) + [rows.size]
    for k, s_, e in zip(uniq, starts, ends):
        body += struct.pack("<IB", int(k), e - s_ - 1) + bytes((rows[s_:e] - k * CHUNK).astype(np.uint8))
    return struct.pack("<I", len(uniq)) + bytes(body)


def write_snpdat_phased(calldata, ancestries, A):
    """File bytes for (n, 2 s) int8 calldata / ancestries (write, ipp:73-363)."""
    cd = np.asarray(calldata); an = np.asarray(ancestries)
    n, two_s = cd.shape
    s_ = two_s // 2
    nnz0 = np.zeros(s_ * A, dtype=np.uint64); nnz1 = np.zeros(s_ * A, dtype=np.uint64)
    snps = []
    for j in range(s_):
        blocks = []
        for a in range(A):
            haps = []
            for k in range(2):
                rows = np.flatnonzero((cd[:, 2 * j + k] == 1) & (an[:, 2 * j + k] == a))
                (nnz0 if k == 0 else nnz1)[j * A + a] = rows.size
                haps.append(_chunk_list(rows))
            blocks.append(struct.pack("<2Q", 16, 16 + len(haps[0])) + haps[0] + haps[1])
        offs, pos = [], 8 * A
        for b in blocks:
            offs.append(pos); pos += len(b)
        snps.append(struct.pack("<%dQ" % A, *offs) + b"".join(blocks))
    preamble = 1 + 16 + 1 + 16 * s_ * A + 8 * (s_ + 1)
    outer = np.concatenate([[preamble], preamble + np.cumsum([len(x) for x in snps])]).astype(np.uint64)
    return struct.pack("<?QQB", False, n, s_, A) + nnz0.tobytes() + nnz1.tobytes() + outer.tobytes() + b"".join(snps)
