"""GPU: the tensor-core (tcgen05 / TMEM, TF32 operands) Gram-panel kernel against NumPy float64 and against the fp32 CUDA-core panel kernel.
The panel kernel computes D[s, u] = sum_i w_i X[i, cols[s]] X[i, cols[u]] for a window of <= 128 columns (sources = the first <= 64)."""
import ctypes as C

import numpy as np
import pytest

import adelie_b200 as ad
from adelie_b200 import _lib

pytestmark = pytest.mark.gpu


def window_gram(M, cols, n_src, w, use_tc):
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    out = np.empty((n_src, len(cols)))
    _lib.check(_lib.load().ab_matrix_window_gram(M._core(), _lib.ptr(cols), len(cols), n_src, _lib.ptr(w), int(use_tc), _lib.ptr(out)))
    return out


@pytest.mark.parametrize("n, p, ncol, n_src, seed", [(64, 128, 128, 64, 0), (1000, 40, 20, 10, 1), (4096, 200, 120, 60, 2), (50_000, 130, 117, 57, 3),
                                                     (33, 16, 16, 8, 4), (200_000, 128, 128, 64, 5)])
def test_window_gram_tensor_core_vs_numpy(n, p, ncol, n_src, seed):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, p), dtype=np.float32))
    w = rng.uniform(0, 2, n).astype(np.float32); w[rng.uniform(size=n) < 0.1] = 0; w /= w.sum()
    cols = rng.choice(p, ncol, replace=False)
    M = ad.matrix.dense(X)
    ref = (X[:, cols[:n_src]].astype(np.float64) * w[:, None].astype(np.float64)).T @ X[:, cols].astype(np.float64)
    scale = np.sqrt(np.outer(np.diag(ref[:, :n_src]), np.einsum("ij,ij,i->j", X[:, cols].astype(np.float64), X[:, cols].astype(np.float64), w.astype(np.float64))))
    cc = window_gram(M, cols, n_src, w, use_tc=0)
    assert np.max(np.abs(cc - ref) / scale) < 2e-6                       # fp32 CUDA cores
    tc = window_gram(M, cols, n_src, w, use_tc=1)
    err = np.max(np.abs(tc - ref) / scale)
    assert err < 2e-3, err                                                # TF32 operands: 2^-11 relative per factor, errors average over the rows
    # symmetric part: D[s, u] == D[u, s] for u < n_src (A and B operands are the same shared-memory tile)
    np.testing.assert_allclose(tc[:, :n_src], tc[:, :n_src].T, rtol=0, atol=1e-6 * np.max(scale))


def test_window_gram_exact_on_tf32_representable_data():
    """Integers < 2^10 and dyadic weights are exact in TF32: the tensor-core result must then be exact (checks layout / swizzle / K stepping)."""
    rng = np.random.default_rng(7)
    n, p = 2048 + 32, 128
    X = np.asfortranarray(rng.integers(-8, 9, size=(n, p)).astype(np.float32))
    w = np.full(n, 4.0 ** -3, dtype=np.float32)                          # sqrt(w) = 2^-3 exactly
    M = ad.matrix.dense(X)
    cols = np.arange(p)[::-1].copy()
    tc = window_gram(M, cols, 64, w, use_tc=1)
    ref = (X[:, cols[:64]].astype(np.float64) * w[:, None]).T @ X[:, cols].astype(np.float64)
    np.testing.assert_array_equal(tc, ref)
