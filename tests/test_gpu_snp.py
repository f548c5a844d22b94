"""GPU parity tests of the SNP unphased matrix (reference MatrixNaiveSNPUnphased, matrix_naive_snp_unphased.ipp,
and IOSNPUnphased): the `.snpdat` chunk lists unpacked on the device, every operator against dense NumPy as the reference's
run_naive does (T/test_matrix.py:251-411, :713-756; atol 1e-14-ish f64 / 1e-4 f32), and the path solver on the packed genotypes
against (i) the same problem held as a dense device matrix (T/test_solver.py:756-818, the reference's special-matrix-vs-dense
equivalence) and (ii) the CPU oracle on the dense equivalent."""
import numpy as np
import pytest
import scipy.sparse as sp

import adelie_b200 as ad
from oracle import oracle as orc
from oracle import snp_oracle as so

pytestmark = pytest.mark.gpu


def _rel(a, b):
    s = np.max(np.abs(b)) if np.size(b) else 0.0
    return np.max(np.abs(a - b)) / (s if s > 0 else 1.0)


def _make(tmp_path, n, p, dtype, read_mode="file", seed=0, **kw):
    data = ad.data.snp_unphased(n, p, seed=seed, **kw)
    h = ad.io.snp_unphased(str(tmp_path / "m.snpdat"), read_mode)
    h.write(data["X"], impute_method="mean")
    cX = ad.matrix.snp_unphased(h, dtype=dtype, n_threads=7)
    X = so.dense_equivalent(data["X"], h.impute, dtype)
    return data, h, cX, X


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-12), (np.float32, 1e-4)])
@pytest.mark.parametrize("read_mode", ["file", "mmap"])
@pytest.mark.parametrize("n,p", [(10, 20), (1, 13), (144, 1), (10000, 1), (777, 37), (5000, 70)])
def test_snp_operators_vs_numpy(tmp_path, dtype, atol, read_mode, n, p):
    data, h, cX, X = _make(tmp_path, n, p, dtype, read_mode)
    assert cX.shape == (n, p) and cX.ndim == 2
    cd, imp = cX.to_host()
    assert np.array_equal(cd, data["X"])                       # device unpack of the chunk lists is bit exact
    np.testing.assert_allclose(imp, h.impute.astype(dtype), rtol=0, atol=0)
    rng = np.random.default_rng(n + p)
    v = rng.normal(size=n).astype(dtype); w = rng.uniform(0, 1, size=n).astype(dtype)
    scale = atol * max(1.0, n / 100)
    out = np.empty(p, dtype=dtype)
    cX.mul(v, w, out); np.testing.assert_allclose(out, X.T @ (v * w), atol=scale)
    cX.sq_mul(w, out); np.testing.assert_allclose(out, (X ** 2).T @ w, atol=scale)
    for j, q in [(0, 1), (p // 3, min(5, p - p // 3)), (p - 1, 1), (0, min(p, 9))]:
        o = np.empty(q, dtype=dtype)
        cX.bmul(j, q, v, w, o); np.testing.assert_allclose(o, X[:, j:j + q].T @ (v * w), atol=scale)
        assert abs(cX.cmul(j, v, w) - X[:, j] @ (v * w)) <= scale
        vv = rng.normal(size=q).astype(dtype); acc = rng.normal(size=n).astype(dtype); exp = acc + X[:, j:j + q] @ vv
        cX.btmul(j, q, vv, acc); np.testing.assert_allclose(acc, exp, atol=atol * 10)
        acc2 = np.zeros(n, dtype=dtype); cX.ctmul(j, 1.5, acc2); np.testing.assert_allclose(acc2, 1.5 * X[:, j], atol=atol * 10)
        C = np.empty((q, q), dtype=dtype, order="F"); cX.cov(j, q, np.sqrt(w), C)
        np.testing.assert_allclose(C, X[:, j:j + q].T @ (w[:, None] * X[:, j:j + q]), atol=scale)
    S = sp.random(3, p, density=0.5, format="csr", dtype=np.float64, random_state=np.random.default_rng(2)).astype(dtype)
    o2 = np.empty((3, n), dtype=dtype); cX.sp_tmul(S, o2)
    np.testing.assert_allclose(o2, (X @ S.T.toarray()).T, atol=atol * 10)
    m = np.empty(p, dtype=dtype); cX.mean(w, m); assert np.allclose(m, 0)       # matrix_naive_snp_unphased.ipp:290-309
    cX.var(m, w, m); assert np.allclose(m, 1)
    with pytest.raises(RuntimeError, match="bmul\\(\\) is given inconsistent inputs"):
        cX.bmul(p - 1, 2, v, w, np.empty(2, dtype=dtype))


def test_snp_construction_variants(tmp_path):
    data, h, cX, X = _make(tmp_path, 1000, 33, np.float64)
    # same matrix from memory, without a file
    cY = ad.matrix.snp_unphased_from_calldata(data["X"], h.impute)
    assert np.array_equal(cY.to_host()[0], data["X"])
    # row range of the file (row sharding): rows [320, 707)
    cZ = ad.matrix.snp_unphased(h, rows=(320, 707))
    assert cZ.shape == (387, 33)
    assert np.array_equal(cZ.to_host()[0], data["X"][320:707])
    v = np.random.default_rng(0).normal(size=387); w = np.full(387, 1 / 387)
    out = np.empty(33); cZ.mul(v, w, out)
    np.testing.assert_allclose(out, X[320:707].T @ (v * w), atol=1e-12)
    with pytest.raises(RuntimeError, match="n_threads must be >= 1"):
        ad.matrix.snp_unphased(h, n_threads=0)
    bad = data["X"].copy(); bad[5, 5] = 3
    with pytest.raises(RuntimeError, match="greater than > 2"):
        ad.matrix.snp_unphased_from_calldata(bad, h.impute)
    # a corrupted chunk count must be caught by the device walk, not run off the column
    raw = bytearray(open(str(tmp_path / "m.snpdat"), "rb").read())
    col0 = int(h.outer[0]); off1 = int.from_bytes(raw[col0 + 8:col0 + 16], "little")
    raw[col0 + off1:col0 + off1 + 4] = (10 ** 6).to_bytes(4, "little")
    open(str(tmp_path / "bad.snpdat"), "wb").write(bytes(raw))
    hb = ad.io.snp_unphased(str(tmp_path / "bad.snpdat")); hb.read()
    with pytest.raises(RuntimeError, match="malformed file|row index outside"):
        ad.matrix.snp_unphased(hb)


def test_snp_device_random_is_shard_invariant():
    n, p = 4096, 40
    full = ad.matrix.snp_unphased_device_random(n, p, dtype=np.float32, seed=5)
    cd, imp = full.to_host()
    assert abs(np.mean(cd == 1) - 0.25 * 0.9) < 0.01 and abs(np.mean(cd == 2) - 0.05 * 0.9) < 0.01 and abs(np.mean(cd == -9) - 0.1) < 0.01
    np.testing.assert_allclose(imp, np.sum(np.where(cd > 0, cd, 0), axis=0) / np.sum(cd >= 0, axis=0), rtol=1e-6)
    part = ad.matrix.snp_unphased_device_random(1024, p, dtype=np.float32, seed=5, row_offset=2048, n_total=n)
    assert np.array_equal(part.to_host()[0], cd[2048:3072])


def _solve_pair(X_snp, X_dense, glm_dev, glm_orc, kw):
    st = ad.grpnet(X_snp, glm_dev, progress_bar=False, **kw)
    st_d = ad.grpnet(X_dense, glm_dev, progress_bar=False, **kw)
    ref = orc.grpnet(X_dense, glm_orc, **kw)
    assert st.error == "" and st_d.error == "" and ref.error == "", (st.error, st_d.error, ref.error)
    return st, st_d, ref


@pytest.mark.parametrize("n,p", [(10, 4), (10, 100), (100, 23), (100, 100), (100, 10000), (3000, 400)])
def test_solve_gaussian_snp_vs_dense_and_oracle(tmp_path, n, p):
    """T/test_solver.py:745-818 (sizes of the reference's test + one multi-CTA size): lasso path on the packed genotypes."""
    data, h, cX, X = _make(tmp_path, n, p, np.float64, sparsity=0.5)
    y = data["glm"].y
    kw = dict(tol=1e-12, early_exit=False, lmda_path_size=15, min_ratio=0.1)
    st, st_d, ref = _solve_pair(cX, X, ad.glm.gaussian(y), orc.glm_spec("gaussian", y), kw)
    # p >> n: the fit saturates and the minimiser is ill-conditioned at the small lambdas; the reference's own test accepts atol = 1e-3
    # there (T/test_solver.py:745-746), we keep 1e-4 relative; 1e-6 relative everywhere else
    rtol = 1e-4 if p > 10 * n else 1e-6
    for other in (st_d, ref):
        assert len(st.lmdas) == len(other.lmdas)
        B, Bo = np.asarray(st.betas.todense()), np.asarray(other.betas.todense())
        assert _rel(B, Bo) < rtol, _rel(B, Bo)
        assert _rel(st.intercepts, other.intercepts) < rtol
    cached, packed_bytes = cX.cache_info()
    assert 0 < cached <= p and packed_bytes >= n * p // 4          # only screened columns are ever decoded


def test_solve_fp32_groups_elastic_net_snp(tmp_path):
    """group elastic net (groups of 5 SNPs, alpha = 0.5) in fp32 on the packed genotypes: batched sweep kernel on the decoded cache."""
    n, p = 6000, 300
    data, h, cX, X = _make(tmp_path, n, p, np.float32, sparsity=0.8, seed=4)
    y = data["glm"].y.astype(np.float32)
    groups = np.arange(0, p, 5)
    kw = dict(groups=groups, alpha=0.5, tol=1e-10, newton_tol=1e-6, early_exit=False, lmda_path_size=20, min_ratio=0.05)
    st, st_d, ref = _solve_pair(cX, X, ad.glm.gaussian(y, dtype=np.float32), orc.glm_spec("gaussian", y, dtype=np.float32), kw)
    assert _rel(np.asarray(st.betas.todense()), np.asarray(st_d.betas.todense())) < 1e-4
    assert _rel(np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())) < 1e-4


def test_solve_binomial_snp(tmp_path):
    n, p = 2000, 120
    data, h, cX, X = _make(tmp_path, n, p, np.float64, sparsity=0.7, seed=2, glm="binomial")
    y = data["glm"].y
    kw = dict(alpha=0.5, tol=1e-12, irls_tol=1e-10, early_exit=False, lmda_path_size=10, min_ratio=0.2)
    st, st_d, ref = _solve_pair(cX, X, ad.glm.binomial(y), orc.glm_spec("binomial", y), kw)
    assert _rel(np.asarray(st.betas.todense()), np.asarray(st_d.betas.todense())) < 1e-6
    assert _rel(np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())) < 1e-6


@pytest.mark.parametrize("dtype,rtol,K", [(np.float64, 1e-6, 8), (np.float32, 1e-4, 8), (np.float64, 1e-6, 3)])
def test_solve_multigaussian_snp(tmp_path, dtype, rtol, K):
    """config 5's layout at test size: multigaussian K = 8 on a snp_unphased matrix (groups of K coefficients per SNP)."""
    n, p = 2500, 90
    data, h, cX, X = _make(tmp_path, n, p, dtype, sparsity=0.8, seed=6, K=K, glm="multigaussian")
    Y = np.ascontiguousarray(data["glm"].y, dtype=dtype)
    # the sweep stops when max_g sum(A * dbeta^2) / gs < tol, i.e. coefficients are accurate to ~sqrt(tol): 1e-14 for a 1e-6 comparison
    tol = 1e-14 if dtype == np.float64 else 1e-10
    kw = dict(tol=tol, newton_tol=1e-12 if dtype == np.float64 else 1e-6, early_exit=False, lmda_path_size=12, min_ratio=0.1)
    st, st_d, ref = _solve_pair(cX, X, ad.glm.multigaussian(Y, dtype=dtype), orc.glm_spec("multigaussian", Y, dtype=dtype), kw)
    assert st.betas.shape == (len(st.lmdas), p * K)
    B, Bd, Br = (np.asarray(s_.betas.todense()) for s_ in (st, st_d, ref))
    assert _rel(B, Bd) < rtol and _rel(B, Br) < rtol, (_rel(B, Bd), _rel(B, Br), _rel(Bd, Br))
    if dtype == np.float32:
        # float32: the K unpenalised intercepts absorb sum_j xbar_j beta_jk (90 SNPs with means ~0.5), i.e. they amplify the coefficients'
        # 1e-4 agreement by sum_j |xbar_j|; what is compared instead is what they are for, the fitted linear predictor X beta + beta0
        Xd = np.asarray(X, dtype=np.float64)
        for l in (len(st.lmdas) // 2, len(st.lmdas) - 1):
            eta = Xd @ B[l].reshape(p, K) + np.asarray(st.intercepts)[l][None]
            eta_r = Xd @ Br[l].reshape(p, K) + np.asarray(ref.intercepts)[l][None]
            assert _rel(eta, eta_r) < rtol
        return
    assert _rel(st.intercepts, ref.intercepts) < rtol


def test_packed_mul_large_vs_decoded_columns():
    """Multi-CTA size (fp32, n = 200k): the packed-bit GEMV and the decoded-column cache agree with NumPy on the downloaded genotypes."""
    n, p = 200_000, 512
    cX = ad.matrix.snp_unphased_device_random(n, p, dtype=np.float32, seed=9)
    rng = np.random.default_rng(0)
    v = rng.normal(size=n).astype(np.float32); w = np.full(n, 1.0 / n, dtype=np.float32)
    out = np.empty(p, dtype=np.float32); cX.mul(v, w, out)
    cd, imp = cX.to_host()
    sel = np.r_[0:24, p - 24:p]
    D = so.dense_equivalent(cd[:, sel], imp[sel], np.float64)
    assert _rel(out[sel], D.T @ (v.astype(np.float64) * w)) < 1e-4
    acc = np.zeros(n, dtype=np.float32)
    cX.btmul(p - 24, 24, np.ones(24, dtype=np.float32), acc)          # decodes 24 columns into the dense cache
    assert _rel(acc, D[:, 24:].sum(axis=1)) < 1e-6
    assert cX.cache_info()[0] == 24


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-12), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,s,A", [(10, 20, 4), (1, 13, 3), (777, 11, 8), (5000, 9, 2)])
def test_snp_phased_ancestry_operators_and_path(tmp_path, dtype, atol, n, s, A):
    """the phased-ancestry chunk lists unpacked on the device into the shared 2-bit storage; operators vs NumPy as the
    reference's test does (T/test_matrix.py test_naive_snp_phased_ancestry), group-lasso path (one group of A columns per SNP) vs the oracle."""
    data = ad.data.snp_phased_ancestry(n, s, A, seed=1, sparsity=0.6)
    h = ad.io.snp_phased_ancestry(str(tmp_path / "pa.snpdat"), "mmap")
    h.write(data["X"], data["ancestries"], A)
    cX = ad.matrix.snp_phased_ancestry(h, dtype=dtype, n_threads=3)
    X = data["dense"].astype(dtype)
    assert cX.shape == (n, s * A)
    cd, _ = cX.to_host()
    assert np.array_equal(cd, data["dense"])                     # device unpack (adds per haplotype) is bit exact
    rng = np.random.default_rng(n + s)
    v = rng.normal(size=n).astype(dtype); w = rng.uniform(0, 1, size=n).astype(dtype)
    scale = atol * max(1.0, n / 100)
    out = np.empty(s * A, dtype=dtype)
    cX.mul(v, w, out); np.testing.assert_allclose(out, X.T @ (v * w), atol=scale)
    cX.sq_mul(w, out); np.testing.assert_allclose(out, (X ** 2).T @ w, atol=scale)
    j, q = A * (s // 2), A
    o = np.empty(q, dtype=dtype); cX.bmul(j, q, v, w, o); np.testing.assert_allclose(o, X[:, j:j + q].T @ (v * w), atol=scale)
    vv = rng.normal(size=q).astype(dtype); acc = rng.normal(size=n).astype(dtype); exp = acc + X[:, j:j + q] @ vv
    cX.btmul(j, q, vv, acc); np.testing.assert_allclose(acc, exp, atol=atol * 10)
    C = np.empty((q, q), dtype=dtype, order="F"); cX.cov(j, q, np.sqrt(w), C)
    np.testing.assert_allclose(C, X[:, j:j + q].T @ (w[:, None] * X[:, j:j + q]), atol=scale)
    if dtype == np.float64 and n >= 100:
        y = data["glm"].y
        kw = dict(groups=data["groups"], penalty=data["penalty"], tol=1e-13, early_exit=False, lmda_path_size=10, min_ratio=0.1)
        st = ad.grpnet(cX, ad.glm.gaussian(y), progress_bar=False, **kw)
        ref = orc.grpnet(X, orc.glm_spec("gaussian", y), **kw)
        assert st.error == "" and ref.error == "", (st.error, ref.error)
        assert _rel(np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())) < 1e-6


# ------------------------------------------------------------------------------------------------------------------------------
# Tensor-core (tcgen05 INT8) multi-response mul on the packed genotypes (csrc/snp_tc.cuh): exact integer accumulation of the
# fixed-point products; compared with NumPy float64 on the dense equivalent and with the CUDA-core packed kernel.
# ------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,p,K,std", [(300, 40, 8, False), (5000, 130, 8, False), (4097, 64, 3, False), (70_000, 200, 8, True), (1000, 65, 2, True),
                                       (33, 5, 5, False)])
def test_multi_response_mul_tensor_core_vs_numpy(n, p, K, std):
    cX = ad.matrix.snp_unphased_device_random(n, p, dtype=np.float32, seed=n + K)
    cd, imp = cX.to_host()
    D = so.dense_equivalent(cd, imp, np.float64)
    M = cX
    if std:                                  # (snp_unphased reports mean 0 / var 1 like the reference: the view needs explicit centres and scales)
        w1 = np.full(n, 1.0 / n)
        c = (D.T @ w1).astype(np.float32); s = np.sqrt(np.maximum((D ** 2).T @ w1 - c.astype(np.float64) ** 2, 1e-3)).astype(np.float32)
        M = ad.matrix.standardize(cX, centers=c, scales=s)
        D = (D - c[None].astype(np.float64)) / s[None].astype(np.float64)
    rng = np.random.default_rng(1)
    V = rng.normal(size=(n, K)).astype(np.float32); W = rng.uniform(0, 2.0 / n, size=(n, K)).astype(np.float32)
    ref = (D.T @ (V.astype(np.float64) * W.astype(np.float64)))                      # (p, K)
    A = ad.matrix.kronecker_eye(M, K)
    try:
        out_tc = np.empty(p * K, dtype=np.float32)
        ad.set_configs("snp_tc", 1)
        A.mul(V.ravel(), W.ravel(), out_tc)
        ad.set_configs("snp_tc", 0)
        out_cc = np.empty(p * K, dtype=np.float32)
        A.mul(V.ravel(), W.ravel(), out_cc)
    finally:
        ad.set_configs("snp_tc", None)
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(out_cc.reshape(p, K) - ref)) <= 2e-5 * scale                # float32 CUDA-core kernel
    assert np.max(np.abs(out_tc.reshape(p, K) - ref)) <= 2e-6 * scale                # integer tensor-core kernel: only the float32 output rounds


def test_multi_response_mul_tensor_core_is_exact_on_integer_data():
    """V with few significant bits (dyadic rationals) survives the 2^30 fixed point exactly: the INT8 path must reproduce the integer
    sums bit for bit (checks the swizzled layouts, the row permutation inside a packed word, the digit recombination and the K stepping)."""
    n, p, K = 2048 + 96, 128, 8
    cX = ad.matrix.snp_unphased_device_random(n, p, dtype=np.float32, seed=3)
    cd, imp = cX.to_host()
    g = np.where(cd < 0, 0, cd).astype(np.float64); m = (cd < 0).astype(np.float64)
    rng = np.random.default_rng(5)
    V = rng.integers(-64, 65, size=(n, K)).astype(np.float32); V[0] = 64                  # class maxima = 64: the fixed point is exact
    A = ad.matrix.kronecker_eye(cX, K)
    out = np.empty(p * K, dtype=np.float32)
    A.mul(V.ravel(), np.ones(n * K, dtype=np.float32), out)
    ref = g.T @ V.astype(np.float64) + imp[:, None] * (m.T @ V.astype(np.float64))
    np.testing.assert_allclose(out.reshape(p, K), ref.astype(np.float32), rtol=1e-6, atol=1e-3)
    sy = g.T @ V.astype(np.float64)                                                        # integer part alone, exactly
    out0 = out.reshape(p, K).astype(np.float64) - (imp[:, None] * (m.T @ V.astype(np.float64)))
    assert np.max(np.abs(out0 - sy)) <= 1e-6 * np.max(np.abs(sy)) + 0.02
