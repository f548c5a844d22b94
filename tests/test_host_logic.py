"""CPU: Python host-side logic of the drop-in surface that needs no device."""
import numpy as np
import pytest

import adelie_b200 as ad


def test_glm_weight_normalisation_and_checks():          # adelie/glm.py:40-55
    y = np.arange(5.0)
    g = ad.glm.gaussian(y, weights=np.array([1.0, 1, 2, 0, 0]))
    assert np.isclose(g.weights.sum(), 1) and g.weights[2] == 0.5
    assert g.name == "gaussian" and g.opt and not g.is_multi and g.dtype == np.float64
    assert np.allclose(ad.glm.gaussian(y).weights, 0.2)
    with pytest.raises(RuntimeError):
        ad.glm.gaussian(y, weights=np.ones(4))
    with pytest.raises(RuntimeError):
        ad.glm.gaussian(np.zeros((3, 2)))
    with pytest.raises(RuntimeError):
        ad.glm.gaussian(np.arange(5))                    # integer dtype without explicit dtype
    m = ad.glm.multigaussian(np.zeros((4, 3)))
    assert m.is_multi and m.name == "multigaussian"
    b = ad.glm.binomial(np.array([0.0, 1, 1]), dtype=np.float32)
    assert b.name == "binomial_logit" and b.dtype == np.float32 and b.y.dtype == np.float32
    r = b.reweight(np.array([0.0, 1, 1]))
    assert np.allclose(r.weights, [0, 0.5, 0.5])


def test_dense_factory_checks():                         # adelie/matrix.py:549-680
    X = np.zeros((4, 3))
    with pytest.warns(UserWarning):
        M = ad.matrix.dense(X)                           # C-contiguous -> warning
    assert M.shape == (4, 3) and M.ndim == 2 and M.rows() == 4 and M.cols() == 3 and M.dtype == np.float64
    with pytest.raises(RuntimeError):
        ad.matrix.dense(X.astype(np.int32))
    with pytest.raises(RuntimeError):
        ad.matrix.dense(np.asfortranarray(X), n_threads=0)
    with pytest.raises(RuntimeError):
        ad.matrix.dense(np.asfortranarray(X), method="cov")
    M = ad.matrix.dense(np.asfortranarray(X))
    with pytest.raises(RuntimeError, match="cmul"):
        M.cmul(7, np.zeros(4), np.zeros(4))              # out-of-range column: the reference's message, before touching the device
    with pytest.raises(RuntimeError, match="bmul"):
        M.bmul(2, 2, np.zeros(4), np.zeros(4), np.zeros(2))
    with pytest.raises(RuntimeError, match="btmul"):
        M.btmul(0, 2, np.zeros(3), np.zeros(4))


def test_data_generator_semantics():                     # adelie/data.py:84-219
    d = ad.data.dense(50, 20, 5, seed=0)
    assert d["X"].shape == (50, 20) and d["X"].flags.f_contiguous
    assert d["groups"][0] == 0 and len(d["groups"]) == 5 and d["group_sizes"].sum() == 20
    assert np.isclose(np.linalg.norm(d["penalty"]), np.sqrt(20))
    d2 = ad.data.dense(50, 20, 5, seed=0)
    assert np.array_equal(d["X"], d2["X"]) and np.array_equal(d["glm"].y, d2["glm"].y)
    e = ad.data.dense(30, 12, 4, equal_groups=True, glm="binomial", seed=1)
    assert np.array_equal(e["groups"], [0, 3, 6, 9]) and set(np.unique(e["glm"].y)) <= {0.0, 1.0}


def test_configs_defaults():
    assert ad.configs.Configs.hessian_min_def == 1e-24 and ad.configs.Configs.dbeta_tol_def == 1e-12
    with pytest.raises(RuntimeError):
        ad.set_configs("nonexistent", 1)
    ad.set_configs("hessian_min", 1e-20)
    assert ad.configs.Configs.hessian_min == 1e-20
    ad.set_configs("hessian_min", None)
    assert ad.configs.Configs.hessian_min == 1e-24


def test_constraints_rejected():
    from adelie_b200.state import _check_constraints
    _check_constraints(None); _check_constraints([None, None])
    with pytest.raises(RuntimeError):
        _check_constraints([object()])


def test_diagnostic_coefficient_interpolates_like_the_reference():
    """adelie/diagnostic.py:560-646: linear interpolation in lambda between neighbouring solutions, boundary solution outside."""
    import scipy.sparse as sp
    from adelie_b200.diagnostic import coefficient, predict
    lmdas = np.array([1.0, 0.5, 0.25])
    betas = sp.csr_matrix(np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 4.0]]))
    icpt = np.array([0.0, 1.0, 3.0])
    b, b0 = coefficient(lmda=0.75, betas=betas, intercepts=icpt, lmdas=lmdas)
    np.testing.assert_allclose(np.asarray(b.todense()).ravel(), [0.5, 0.0]); assert b0 == pytest.approx(0.5)
    b, b0 = coefficient(lmda=0.3, betas=betas, intercepts=icpt, lmdas=lmdas)
    np.testing.assert_allclose(np.asarray(b.todense()).ravel(), [1.8, 3.2]); assert b0 == pytest.approx(2.6)
    b, b0 = coefficient(lmda=2.0, betas=betas, intercepts=icpt, lmdas=lmdas)          # above the path: first solution
    assert b0 == 0.0 and b.nnz == 0
    b, b0 = coefficient(lmda=0.1, betas=betas, intercepts=icpt, lmdas=lmdas)          # below the path: last solution
    assert b0 == 3.0
    X = np.arange(6.0).reshape(3, 2)
    eta = predict(X, betas, icpt, offsets=np.array([10.0, 20.0, 30.0]))
    np.testing.assert_allclose(eta, betas.toarray() @ X.T + icpt[:, None] + np.array([10.0, 20.0, 30.0])[None])


def test_standardize_and_subset_numpy_paths():
    import adelie_b200 as ad
    rng = np.random.default_rng(0)
    Z = rng.normal(1.0, 2.0, (50, 4))
    S = ad.matrix.standardize(Z, ddof=1)
    np.testing.assert_allclose(S.mean(axis=0), 0, atol=1e-12)
    np.testing.assert_allclose(S.std(axis=0, ddof=1), 1, atol=1e-12)
    assert S.flags.f_contiguous and not np.shares_memory(S, Z)
    assert np.array_equal(ad.matrix.subset(Z, [3, 1], axis=1), Z[:, [3, 1]])


def test_compute_penalty():
    import scipy.sparse as sp
    from adelie_b200.diagnostic import compute_penalty
    B = np.array([[3.0, 4.0, 0.0, 1.0], [0.0, 0.0, 2.0, 0.0]])
    groups = np.array([0, 2, 3]); gsz = np.array([2, 1, 1]); pen = np.array([1.0, 2.0, 0.5])
    exp = np.array([1.0 * (0.7 * 5 + 0.15 * 25) + 0.5 * (0.7 * 1 + 0.15 * 1), 2.0 * (0.7 * 2 + 0.15 * 4)])
    np.testing.assert_allclose(compute_penalty(groups, gsz, pen, 0.7, B), exp)
    np.testing.assert_allclose(compute_penalty(groups, gsz, pen, 0.7, sp.csr_matrix(B)), exp)


def test_matrices_are_freed_by_refcount_not_by_the_cyclic_gc():
    """Round-1 bug: `self.T = MatrixNaiveTranspose(self)` put every matrix in a reference cycle, so the device copy of X was
    only freed when the cyclic GC happened to run (an HBM leak of one X per `grpnet(ndarray)` call)."""
    import gc
    import weakref
    import scipy.sparse as sp
    X = np.asfortranarray(np.random.default_rng(0).normal(size=(40, 6)))
    gc.disable()
    try:
        makers = [lambda: ad.matrix.dense(X), lambda: ad.matrix.sparse(sp.csc_matrix(X)), lambda: ad.matrix.kronecker_eye(X, 3),
                  lambda: ad.matrix.concatenate([ad.matrix.dense(X), ad.matrix.dense(X)], axis=1)]
        for mk in makers:
            m = mk()
            t = m.T
            assert t.T is m and t._mat is m
            w = weakref.ref(m)
            del m, t
            assert w() is None
    finally:
        gc.enable()


def test_snpdat_reader_rejects_malformed_headers(tmp_path):
    """crafted headers must not overflow size computations or walk outside the buffer; the reader raises instead."""
    import struct
    rng = np.random.default_rng(0)
    n, p = 300, 7
    cd = np.asfortranarray(rng.choice([-9, 0, 1, 2], size=(n, p), p=[0.1, 0.6, 0.2, 0.1]).astype(np.int8))
    fn = str(tmp_path / "ok.snpdat")
    h = ad.io.snp_unphased(fn)
    h.write(cd)
    h.read()
    D = h.to_dense()
    assert np.array_equal(D, cd)
    raw = bytearray(open(fn, "rb").read())
    pre = 1 + 16 + p * 24
    def attempt(mut, match):
        b = bytearray(raw); mut(b)
        f2 = str(tmp_path / "bad.snpdat"); open(f2, "wb").write(b)
        g = ad.io.snp_unphased(f2)
        with pytest.raises(RuntimeError, match=match):
            g.read(); g.to_dense()
    attempt(lambda b: b.__setitem__(slice(9, 17), struct.pack("<Q", 2 ** 61)), "too short")                 # snps * 24 would overflow
    attempt(lambda b: b.__setitem__(slice(9, 17), struct.pack("<Q", 2 ** 64 - 1)), "too short")
    attempt(lambda b: b.__setitem__(slice(pre, pre + 8), struct.pack("<Q", 3)), "header")                   # outer[0] inside the preamble
    attempt(lambda b: b.__setitem__(slice(pre + 16, pre + 24), struct.pack("<Q", 40)), "monotone|header")   # outer goes backwards
    attempt(lambda b: b.__setitem__(slice(pre + 8 * p, pre + 8 * p + 8), struct.pack("<Q", len(raw) + 99)), "past the end")
    col0 = struct.unpack("<Q", raw[pre:pre + 8])[0]
    attempt(lambda b: b.__setitem__(slice(col0, col0 + 8), struct.pack("<Q", 10 ** 9)), "malformed")        # category offset outside the column
    off1 = struct.unpack("<Q", raw[col0 + 8:col0 + 16])[0]
    attempt(lambda b: b.__setitem__(slice(col0 + off1, col0 + off1 + 4), struct.pack("<I", 10 ** 6)), "malformed")    # chunk count runs off the column
    attempt(lambda b: b.__setitem__(slice(col0 + off1 + 4, col0 + off1 + 8), struct.pack("<I", 10 ** 5)), "malformed")  # row index out of range


def test_user_defined_base_classes_host_side():
    """User-defined GLM / matrix base classes (trampoline counterparts): construction, defaults and materialisation need no GPU."""
    import numpy as np
    import adelie_b200 as ad

    class G(ad.glm.GlmBase64):
        def __init__(self, y, w):
            ad.glm.GlmBase64.__init__(self, "mine", y, w)

    y = np.arange(5, dtype=np.float32)
    g = G(y, np.ones(5))
    assert g.dtype == np.float64 and g.y.dtype == np.float64 and not g.is_multi and g.opt is False
    np.testing.assert_allclose(g.weights, 0.2)
    for call in (lambda: g.gradient(y, y), lambda: g.hessian(y, y, y), lambda: g.loss(y), lambda: g.loss_full(), lambda: g.inv_link(y, y)):
        try:
            call()
            raise AssertionError("expected NotImplementedError")
        except NotImplementedError:
            pass
    hess = np.array([2.0, 0.0, -1.0, 4.0, 1.0]); grad = np.ones(5); out = np.empty(5)
    g.inv_hessian_gradient(grad, grad, hess, out)                      # glm_base.ipp:25-36
    np.testing.assert_allclose(out, [0.5, 1e24, 1e24, 0.25, 1.0])
    assert ad.glm.GlmMultiBase32.is_multi and ad.glm.GlmMultiBase32._user_dtype == np.float32

    class M(ad.matrix.MatrixNaiveBase32):
        def __init__(self, Z):
            ad.matrix.MatrixNaiveBase32.__init__(self)
            self.Z = Z

        def rows(self):
            return self.Z.shape[0]

        def cols(self):
            return self.Z.shape[1]

        def ctmul(self, j, v, out):
            out += v * self.Z[:, j]

    Z = np.arange(12, dtype=np.float32).reshape(4, 3)
    m = M(Z)
    assert m.dtype == np.float32 and m.shape == (4, 3) and isinstance(m, ad.matrix.MatrixNaiveBase)
    D = m.to_dense()
    assert D.flags.f_contiguous and D.dtype == np.float32
    np.testing.assert_array_equal(D, Z)


def test_expanded_sparse_multi_matrix_layout():
    """[kron(1, I_K) | kron(X, I_K)] as the sparse multi-response states materialise it (adelie/state.py:1100-1125 layout): row i*K + l,
    column n_int + j*K + l holds X[i, j]; the intercept block has ones at (i*K + l, l)."""
    import numpy as np
    import scipy.sparse as sp
    from adelie_b200 import state as st, matrix as mx
    rng = np.random.default_rng(0)
    n, p, K = 7, 5, 3
    D = rng.standard_normal((n, p)) * (rng.uniform(size=(n, p)) < 0.5)
    X = mx.sparse(sp.csc_matrix(D))
    for intercept in (True, False):
        E = st._expand_sparse_multi(X, K, intercept, 1)
        M = E._mat.toarray()
        n_int = K if intercept else 0
        assert M.shape == (n * K, n_int + p * K)
        ref = np.zeros_like(M)
        for i in range(n):
            for l in range(K):
                if intercept:
                    ref[i * K + l, l] = 1
                for j in range(p):
                    ref[i * K + l, n_int + j * K + l] = D[i, j]
        np.testing.assert_array_equal(M, ref)
        assert E._mat.has_sorted_indices


def test_block_diag_cov_assembly_host_side():
    """matrix.block_diag(method="cov") of dense blocks assembles the (p, p) matrix on the host (no device round trip)."""
    import numpy as np
    import adelie_b200 as ad
    B1 = np.array([[2.0, 0.5], [0.5, 1.0]]); B2 = np.array([[3.0]]); B3 = np.eye(2) * 4
    M = ad.matrix.block_diag([B1, ad.matrix.dense(B2, method="cov"), B3], method="cov")
    assert isinstance(M, ad.matrix.MatrixCovBase64) and M.cols() == 5 and M.shape == (5, 5)
    ref = np.zeros((5, 5)); ref[:2, :2] = B1; ref[2, 2] = 3; ref[3:, 3:] = B3
    np.testing.assert_array_equal(M._mat, ref)
    try:
        ad.matrix.block_diag([B1], method="naive")
        raise AssertionError("expected RuntimeError")
    except RuntimeError:
        pass
    try:
        ad.matrix.dense(np.zeros((2, 3)), method="cov")
        raise AssertionError("expected RuntimeError")
    except RuntimeError as e:
        assert "mat must be (p, p)" in str(e)
