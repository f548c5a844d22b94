"""CPU tests of the SNP unphased storage: the NumPy oracle (oracle/snp_oracle.py) against the hand-derived golden
file and the reference's own test expectations (T/test_io.py:7-60, T/test_matrix.py:713-756), and the product's host-side C++
reader / writer (csrc/snp.cuh SnpUnphasedIO, called through the C ABI -- no GPU involved) against both."""
import os

import numpy as np
import pytest

import adelie_b200 as ad
from oracle import snp_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
TINY = np.asfortranarray(np.array([[1, 0, -9, 2, 1], [2, -9, -9, 0, 2]], dtype=np.int8).T)


def _golden():
    with open(os.path.join(HERE, "golden", "snp_unphased_tiny.snpdat.hex")) as f:
        return bytes.fromhex(f.read().strip())


def _calldata(n, p, seed=0):
    # generator of the reference's test (T/test_io.py:19-34)
    np.random.seed(seed)
    cd = np.zeros((n, p), dtype=np.int8)
    cd.ravel()[np.random.choice(np.arange(n * p), int(0.25 * n * p), replace=False)] = -9
    cd.ravel()[np.random.choice(np.arange(n * p), int(0.25 * n * p), replace=False)] = 1
    cd.ravel()[np.random.choice(np.arange(n * p), int(0.05 * n * p), replace=False)] = 2
    return np.asfortranarray(cd)


def test_oracle_writer_matches_hand_derived_golden():
    blob, imp = so.write_snpdat(TINY)
    assert blob == _golden()
    np.testing.assert_allclose(imp, [1.0, 4.0 / 3.0])
    parsed = so.read_snpdat(_golden())
    assert parsed["rows"] == 5 and parsed["snps"] == 2
    assert np.array_equal(parsed["dense"], TINY)
    assert list(parsed["nnz"]) == [4, 4] and list(parsed["nnm"]) == [4, 3] and list(parsed["outer"]) == [89, 144, 194]


@pytest.mark.parametrize("read_mode", ["file", "mmap"])
def test_product_io_matches_hand_derived_golden(tmp_path, read_mode):
    fn = str(tmp_path / "tiny.snpdat")
    h = ad.io.snp_unphased(fn, read_mode=read_mode)
    with pytest.raises(RuntimeError, match="File is not read yet"):
        h.rows
    w, _ = h.write(TINY, "mean")
    with open(fn, "rb") as f:
        assert f.read() == _golden()
    assert h.read() == w == 194
    assert h.rows == 5 and h.cols == 2 and h.snps == 2
    assert np.array_equal(h.to_dense(), TINY)
    np.testing.assert_allclose(h.impute, [1.0, 4.0 / 3.0])
    # reading the golden bytes written by somebody else
    fn2 = str(tmp_path / "golden.snpdat")
    with open(fn2, "wb") as f:
        f.write(_golden())
    h2 = ad.io.snp_unphased(fn2, read_mode=read_mode)
    h2.read()
    assert np.array_equal(h2.to_dense(), TINY) and list(h2.nnz) == [4, 4] and list(h2.nnm) == [4, 3] and list(h2.outer) == [89, 144, 194]


@pytest.mark.parametrize("read_mode", ["file", "mmap"])
@pytest.mark.parametrize("n,p", [(1, 1), (200, 32), (1421, 927), (513, 3), (256, 5), (257, 5)])
def test_io_reference_expectations(tmp_path, n, p, read_mode):
    """T/test_io.py:7-60 for the product's handler, plus byte-for-byte agreement with the oracle's independent writer."""
    cd = _calldata(n, p)
    fn = str(tmp_path / "x.snpdat")
    h = ad.io.snp_unphased(fn, read_mode=read_mode)
    w, _ = h.write(cd, "mean", n_threads=2)
    r = h.read()
    r = h.read()      # double read
    assert w == r
    assert np.allclose(h.nnm, np.sum(cd >= 0, axis=0))
    with np.errstate(all="ignore"):
        means = np.nan_to_num(np.where(np.sum(cd >= 0, axis=0) > 0, np.sum(np.where(cd > 0, cd, 0), axis=0) / np.maximum(np.sum(cd >= 0, axis=0), 1), 0.0))
    assert np.allclose(h.impute, means)
    assert h.rows == n and h.cols == p and h.snps == p
    assert np.allclose(h.nnz, np.sum(cd != 0, axis=0))
    assert np.allclose(h.to_dense(), cd)
    blob, imp = so.write_snpdat(cd)
    with open(fn, "rb") as f:
        assert f.read() == blob
    parsed = so.read_snpdat(blob)
    assert np.array_equal(parsed["dense"], cd) and np.allclose(parsed["impute"], h.impute)


def test_io_user_impute_and_errors(tmp_path):
    cd = _calldata(50, 7)
    fn = str(tmp_path / "u.snpdat")
    h = ad.io.snp_unphased(fn)
    imp = np.linspace(0.1, 0.7, 7)
    h.write(cd, imp)
    h.read()
    np.testing.assert_array_equal(h.impute, imp)
    assert so.write_snpdat(cd, imp)[0] == open(fn, "rb").read()
    bad = cd.copy(); bad[3, 2] = 3
    with pytest.raises(RuntimeError, match="Detected a value greater than > 2"):
        h.write(bad, "mean")
    with pytest.raises(RuntimeError, match="impute must have length"):
        h.write(cd, np.zeros(3))
    with pytest.raises(RuntimeError):
        ad.io.snp_unphased(str(tmp_path / "missing.snpdat")).read()
    with pytest.raises(RuntimeError, match="read mode"):
        ad.io.snp_unphased(fn, read_mode="bogus")
    with pytest.raises(ValueError):
        h.write(cd, 3.0)


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-12), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,p", [(10, 20), (1, 13), (144, 1), (1000, 3)])
def test_oracle_matrix_ops_vs_dense(n, p, dtype, atol):
    """run_naive of the reference (T/test_matrix.py:251-411) for the oracle's category-wise operators."""
    data = ad.data.snp_unphased(n, p, seed=0)
    blob, imp = so.write_snpdat(data["X"])
    M = so.SnpMatrix(so.read_snpdat(blob), dtype)
    X = so.dense_equivalent(data["X"], imp, dtype)
    rng = np.random.default_rng(0)
    v = rng.normal(size=n).astype(dtype); w = rng.uniform(0, 1, size=n).astype(dtype)
    out = np.empty(p, dtype=dtype)
    M.mul(v, w, out); np.testing.assert_allclose(out, X.T @ (v * w), atol=atol * max(1, n / 100))
    M.sq_mul(w, out); np.testing.assert_allclose(out, (X ** 2).T @ w, atol=atol * max(1, n / 100))
    for j in range(0, p, max(1, p // 3)):
        q = min(3, p - j)
        o = np.empty(q, dtype=dtype); M.bmul(j, q, v, w, o)
        np.testing.assert_allclose(o, X[:, j:j + q].T @ (v * w), atol=atol * max(1, n / 100))
        vv = rng.normal(size=q).astype(dtype); acc = rng.normal(size=n).astype(dtype); exp = acc + X[:, j:j + q] @ vv
        M.btmul(j, q, vv, acc); np.testing.assert_allclose(acc, exp, atol=atol * 10)
        C = np.empty((q, q), dtype=dtype); M.cov(j, q, np.sqrt(w), C)
        np.testing.assert_allclose(C, X[:, j:j + q].T @ (w[:, None] * X[:, j:j + q]), atol=atol * max(1, n / 100))


def test_data_generator_proportions():
    d = ad.data.snp_unphased(400, 50, seed=3)
    X = d["X"]
    assert X.dtype == np.int8 and X.flags.f_contiguous and X.shape == (400, 50)
    assert abs(np.mean(X == -9) - 0.1) < 1e-3
    assert set(np.unique(X)) <= {-9, 0, 1, 2}
    assert np.array_equal(d["groups"], np.arange(50)) and np.all(d["group_sizes"] == 1)
    assert d["glm"].y.shape == (400,)
    dk = ad.data.snp_unphased(100, 30, K=3, glm="multigaussian", seed=1)
    assert dk["glm"].y.shape == (100, 3)


# ---------------------------------------------------------------------------------------------------- phased ancestry
PA_CD = np.asfortranarray(np.array([[1, 0], [1, 1], [0, 1]], dtype=np.int8))
PA_AN = np.asfortranarray(np.array([[0, 1], [1, 1], [0, 0]], dtype=np.int8))


def _golden_phased():
    with open(os.path.join(HERE, "golden", "snp_phased_tiny.snpdat.hex")) as f:
        return bytes.fromhex(f.read().strip())


@pytest.mark.parametrize("read_mode", ["file", "mmap"])
def test_phased_io_matches_hand_derived_golden(tmp_path, read_mode):
    assert so.write_snpdat_phased(PA_CD, PA_AN, 2) == _golden_phased()
    assert np.array_equal(so.phased_dense(PA_CD, PA_AN, 2), [[1, 0], [0, 2], [1, 0]])
    fn = str(tmp_path / "pa.snpdat")
    h = ad.io.snp_phased_ancestry(fn, read_mode=read_mode)
    w, _ = h.write(PA_CD, PA_AN, 2)
    with open(fn, "rb") as f:
        assert f.read() == _golden_phased()
    assert h.read() == w == 154
    assert (h.rows, h.snps, h.ancestries, h.cols) == (3, 1, 2, 2)
    assert list(h.nnz0) == [1, 1] and list(h.nnz1) == [1, 1] and list(h.outer) == [66, 154]
    assert np.array_equal(h.to_dense(), [[1, 0], [0, 2], [1, 0]])


@pytest.mark.parametrize("n,s,A", [(1, 1, 1), (200, 32, 4), (1421, 97, 8), (513, 3, 7), (256, 5, 2)])
def test_phased_io_reference_expectations(tmp_path, n, s, A):
    """T/test_io.py:63-110 for the product's handler + byte-for-byte agreement with the oracle's independent writer."""
    data = ad.data.snp_phased_ancestry(n, s, A, seed=0)
    cd, an = data["X"], data["ancestries"]
    fn = str(tmp_path / "x.snpdat")
    h = ad.io.snp_phased_ancestry(fn)
    w, _ = h.write(cd, an, A, n_threads=2)
    r = h.read()
    assert w == r and (h.rows, h.snps, h.ancestries, h.cols) == (n, s, A, s * A)
    dense = so.phased_dense(cd, an, A)
    assert np.array_equal(dense, data["dense"])
    assert np.array_equal(h.to_dense(), dense)
    with open(fn, "rb") as f:
        assert f.read() == so.write_snpdat_phased(cd, an, A)
    for k, nz in ((0, h.nnz0), (1, h.nnz1)):
        exp = np.array([[np.sum((cd[:, 2 * j + k] == 1) & (an[:, 2 * j + k] == a)) for a in range(A)] for j in range(s)]).ravel()
        assert np.array_equal(nz, exp)


def test_phased_io_errors(tmp_path):
    h = ad.io.snp_phased_ancestry(str(tmp_path / "e.snpdat"))
    cd = np.zeros((4, 4), dtype=np.int8); an = np.zeros((4, 4), dtype=np.int8)
    bad = cd.copy(); bad[1, 1] = 2
    with pytest.raises(RuntimeError, match="non-binary value"):
        h.write(bad, an, 2)
    bad = an.copy(); bad[0, 0] = 2
    with pytest.raises(RuntimeError, match="ancestry not in the range"):
        h.write(cd, bad, 2)
    with pytest.raises(RuntimeError, match="shape \\(n, 2\\*s\\)"):
        h.write(np.zeros((4, 3), dtype=np.int8), np.zeros((4, 3), dtype=np.int8), 2)
    with pytest.raises(RuntimeError, match="File is not read yet"):
        h.rows
