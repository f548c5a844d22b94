"""Matrix objects (reference: adelie/matrix.py; operator set of MatrixNaiveBase,
adelie/src/py_matrix.cpp:832-1071 = CORE/matrix/matrix_naive_base.hpp:57-143).

A matrix keeps a reference to the host array it was built from and owns a device-resident copy
(column-major, rows padded to 32, see csrc/matrix.cuh); the copy is made once, on first use.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np
from scipy.sparse import csc_matrix, csr_matrix

from . import _lib


def _to_dtype(mat):
    return mat.dtype


class MatrixNaiveTranspose:
    """adelie/matrix.py:52-80.  Built on access by ``mat.T`` (a property): the matrix itself holds no reference to its transpose view,
    so a matrix is never part of a reference cycle and its device copy is freed by reference counting as soon as the last user
    drops it (a cycle would leave the free to the cyclic GC, which does not see HBM pressure)."""
    def __init__(self, mat):
        self._mat = mat

    @property
    def T(self):
        return self._mat

    def __matmul__(self, v):
        dtype = _to_dtype(self._mat)
        v = np.asarray(v, dtype=dtype)
        if (len(v.shape) <= 0) or (len(v.shape) > 2):
            raise ValueError("Right argument must be either 1 or 2-dimensional.")
        n, p = self._mat.shape
        ones = np.ones(n, dtype=dtype)
        if len(v.shape) == 1:
            out = np.empty(p, dtype=dtype)
            self._mat.mul(np.ascontiguousarray(v), ones, out)
            return out
        v = np.asfortranarray(v)
        out = np.empty((v.shape[1], p), dtype=dtype)
        for i in range(out.shape[0]):
            self._mat.mul(np.ascontiguousarray(v[:, i]), ones, out[i])
        return out.T


class MatrixNaiveBase:
    """Python-side sugar shared by every naive matrix (adelie/matrix.py:83-190)."""
    def __init__(self, n_threads=1):
        if n_threads < 1:
            raise RuntimeError("adelie_core: n_threads must be >= 1.")
        self._n_threads = n_threads

    @property
    def T(self):
        return MatrixNaiveTranspose(self)

    @property
    def ndim(self):
        return 2

    @property
    def shape(self):
        return (self.rows(), self.cols())

    def __matmul__(self, v):
        dtype = _to_dtype(self)
        n, p = self.shape
        if isinstance(v, (csr_matrix, csc_matrix)):
            v = v.tocsr().transpose().tocsr()
            out = np.empty((v.shape[0], n), dtype=dtype)
            self.sp_tmul(v, out)
            return out.T
        v = np.asarray(v, dtype=dtype)
        if (len(v.shape) <= 0) or (len(v.shape) > 2):
            raise ValueError("Right argument must be either 1 or 2-dimensional.")
        if len(v.shape) == 1:
            out = np.zeros(n, dtype=dtype)
            self.btmul(0, p, np.ascontiguousarray(v), out)
            return out
        v = np.asfortranarray(v)
        out = np.zeros((v.shape[1], n), dtype=dtype)
        for i in range(out.shape[0]):
            self.btmul(0, p, np.ascontiguousarray(v[:, i]), out[i])
        return out.T


class _UserMatrix(MatrixNaiveBase):
    """A matrix defined in Python (reference: subclass ``adelie.matrix.MatrixNaiveBase64`` and override the virtuals; the pybind trampoline
    PyMatrixNaiveBase, adelie/src/py_matrix.cpp:627-825).  The subclass implements ``rows()``, ``cols()`` and the operators on NumPy
    arrays; Python-side callers (``grpnet``'s initial ``mul`` passes, ``@``) use them directly.  The device solver needs the entries in
    HBM: the first time a state asks for the core handle the matrix is MATERIALISED through its own ``ctmul`` (column j = ctmul(j, 1, 0),
    one call per column) into a dense device matrix -- a Python round trip per group update inside the fused sweep is not an option."""
    _user_dtype = None

    def __init__(self, n_threads=1):
        MatrixNaiveBase.__init__(self, n_threads)
        self._materialized = None

    @property
    def dtype(self):
        return self._user_dtype

    def rows(self):
        raise NotImplementedError("rows() must be implemented by the user-defined matrix.")

    def cols(self):
        raise NotImplementedError("cols() must be implemented by the user-defined matrix.")

    def ctmul(self, j, v, out):
        raise NotImplementedError("ctmul() must be implemented by the user-defined matrix.")

    def to_dense(self):
        n, p = self.rows(), self.cols()
        out = np.zeros((n, p), dtype=self._user_dtype, order="F")
        for j in range(p):
            self.ctmul(j, 1, out[:, j])
        return out

    def _dev(self):
        if getattr(self, "_materialized", None) is None:
            self._materialized = _Dense(self.to_dense(), getattr(self, "_n_threads", 1))
        return self._materialized

    def _core(self):
        return self._dev()._core()

    # operators the subclass did not override run on the materialised device copy
    def cmul(self, j, v, weights):
        return self._dev().cmul(j, v, weights)

    cmul_safe = cmul

    def bmul(self, j, q, v, weights, out):
        self._dev().bmul(j, q, v, weights, out)

    bmul_safe = bmul

    def btmul(self, j, q, v, out):
        self._dev().btmul(j, q, v, out)

    def mul(self, v, weights, out):
        self._dev().mul(v, weights, out)

    def cov(self, j, q, sqrt_weights, out):
        self._dev().cov(j, q, sqrt_weights, out)

    def sq_mul(self, weights, out):
        self._dev().sq_mul(weights, out)

    def sp_tmul(self, v, out):
        self._dev().sp_tmul(v, out)

    def close(self):
        m, self._materialized = getattr(self, "_materialized", None), None
        if m is not None:
            m.close()


class MatrixNaiveBase64(_UserMatrix):
    _user_dtype = np.float64


class MatrixNaiveBase32(_UserMatrix):
    _user_dtype = np.float32


class _DeviceMatrix(MatrixNaiveBase):
    """Host-pointer operator front-end over an ``ab_matrix`` handle."""
    def __init__(self, dtype, n, p, n_threads):
        MatrixNaiveBase.__init__(self, n_threads)
        self.dtype = np.dtype(dtype).type
        self._n, self._p = int(n), int(p)
        self._handle = None

    def rows(self):
        return self._n

    def cols(self):
        return self._p

    def _make_handle(self):
        raise NotImplementedError

    def _core(self):
        if self._handle is None:
            try:
                self._handle = self._make_handle()
            except RuntimeError as e:
                if "out of memory" not in str(e):
                    raise
                # device copies are freed by reference counting; objects that user code left in a cycle only die in the cyclic GC,
                # which does not see HBM pressure: collect once and retry
                import gc
                gc.collect()
                self._handle = self._make_handle()
        return self._handle

    def close(self):
        """Frees the device copy now (extension; the reference's matrices are host views).  The matrix stays usable: the copy is
        rebuilt on the next operator call when the matrix still has its host source."""
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _lib.load().ab_matrix_free(h)
            except Exception:
                pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        self.close()

    def __getitem__(self, key):
        """``mat[rows]`` / ``mat[:, cols]`` / ``mat[rows, cols]`` sugar over :func:`subset` (adelie/matrix.py:84-134): integers, slices, lists,
        index or boolean arrays; with two list-like subsets at least one must be an integer or a slice."""
        one = (int, np.integer, slice)
        if isinstance(key, tuple):
            if len(key) == 0:
                return self
            if len(key) > 2:
                raise ValueError("Key must be of length 1 or 2 if it is a tuple.")
            if len(key) == 2 and not isinstance(key[0], one) and not isinstance(key[1], one):
                raise ValueError("If row and column subsets are provided, at least one must not be a list-like object. ")
        elif isinstance(key, one + (list, np.ndarray)):
            key = (key,)
        else:
            raise ValueError("Subsets must be integer, slice, list, or np.ndarray objects.")

        def conv(sel, size):
            if isinstance(sel, (int, np.integer)):
                return np.array([sel])
            if isinstance(sel, slice):
                return None if sel == slice(None) else np.arange(size)[sel]
            sel = np.asarray(sel)
            return np.flatnonzero(sel) if sel.dtype == np.dtype("bool") else sel
        out = self
        for axis, sel in enumerate(key):
            idx = conv(sel, out.shape[axis])
            if idx is not None:
                out = subset(out, idx, axis=axis, n_threads=self._n_threads)
        return out

    def _vec(self, a, size, name, fn):
        if not isinstance(a, np.ndarray) or a.dtype != self.dtype or a.ndim != 1 or not a.flags.c_contiguous:
            raise TypeError(f"{fn}(): {name} must be a 1-D contiguous array of dtype {np.dtype(self.dtype).name}")
        return a

    # ---- operators (semantics of matrix_naive_base.hpp:57-143; shape errors as :148-271)
    def cmul(self, j, v, weights):
        n, p = self.shape
        v = self._vec(v, n, "v", "cmul"); weights = self._vec(weights, n, "weights", "cmul")
        if j < 0 or j >= p or v.size != n or weights.size != n:
            raise RuntimeError(f"adelie_core: cmul() is given inconsistent inputs! (j={j}, v={v.size}, w={weights.size}, r={n}, c={p})")
        out = C.c_double()
        _lib.check(_lib.load().ab_matrix_cmul(self._core(), j, _lib.ptr(v), _lib.ptr(weights), C.byref(out)))
        return self.dtype(out.value)

    cmul_safe = cmul

    def ctmul(self, j, v, out):
        n, p = self.shape
        out = self._vec(out, n, "out", "ctmul")
        if j < 0 or j >= p or out.size != n:
            raise RuntimeError(f"adelie_core: ctmul() is given inconsistent inputs! (j={j}, o={out.size}, r={n}, c={p})")
        _lib.check(_lib.load().ab_matrix_ctmul(self._core(), j, float(v), _lib.ptr(out)))

    def bmul(self, j, q, v, weights, out):
        n, p = self.shape
        v = self._vec(v, n, "v", "bmul"); weights = self._vec(weights, n, "weights", "bmul"); out = self._vec(out, q, "out", "bmul")
        if j < 0 or j > p - q or v.size != n or weights.size != n or out.size != q:
            raise RuntimeError(f"adelie_core: bmul() is given inconsistent inputs! (j={j}, q={q}, v={v.size}, w={weights.size}, o={out.size}, r={n}, c={p})")
        _lib.check(_lib.load().ab_matrix_bmul(self._core(), j, q, _lib.ptr(v), _lib.ptr(weights), _lib.ptr(out)))

    bmul_safe = bmul

    def btmul(self, j, q, v, out):
        n, p = self.shape
        v = self._vec(v, q, "v", "btmul"); out = self._vec(out, n, "out", "btmul")
        if j < 0 or j > p - q or v.size != q or out.size != n:
            raise RuntimeError(f"adelie_core: btmul() is given inconsistent inputs! (j={j}, q={q}, v={v.size}, o={out.size}, r={n}, c={p})")
        _lib.check(_lib.load().ab_matrix_btmul(self._core(), j, q, _lib.ptr(v), _lib.ptr(out)))

    def mul(self, v, weights, out):
        n, p = self.shape
        v = self._vec(v, n, "v", "mul"); weights = self._vec(weights, n, "weights", "mul"); out = self._vec(out, p, "out", "mul")
        if v.size != n or weights.size != n or out.size != p:
            raise RuntimeError(f"adelie_core: mul() is given inconsistent inputs! (v={v.size}, w={weights.size}, o={out.size}, r={n}, c={p})")
        _lib.check(_lib.load().ab_matrix_mul(self._core(), _lib.ptr(v), _lib.ptr(weights), _lib.ptr(out)))

    def cov(self, j, q, sqrt_weights, out):
        n, p = self.shape
        sqrt_weights = self._vec(sqrt_weights, n, "sqrt_weights", "cov")
        if (j < 0 or j > p - q or sqrt_weights.size != n or out.shape != (q, q)):
            raise RuntimeError(f"adelie_core: cov() is given inconsistent inputs! (j={j}, q={q}, w={sqrt_weights.size}, o_r={out.shape[0]}, o_c={out.shape[-1]}, r={n}, c={p})")
        if out.dtype != self.dtype or not out.flags.f_contiguous:
            raise TypeError("cov(): out must be an F-contiguous (q, q) array of the matrix dtype")
        _lib.check(_lib.load().ab_matrix_cov(self._core(), j, q, _lib.ptr(sqrt_weights), _lib.ptr(out)))

    def sq_mul(self, weights, out):
        n, p = self.shape
        weights = self._vec(weights, n, "weights", "sq_mul"); out = self._vec(out, p, "out", "sq_mul")
        _lib.check(_lib.load().ab_matrix_sq_mul(self._core(), _lib.ptr(weights), _lib.ptr(out)))

    def sp_tmul(self, v, out):
        n, p = self.shape
        v = csr_matrix(v)
        if v.shape[1] != p or out.shape != (v.shape[0], n):
            raise RuntimeError(f"adelie_core: sp_tmul() is given inconsistent inputs! (vr={v.shape[0]}, vc={v.shape[1]}, o_r={out.shape[0]}, o_c={out.shape[1]}, r={n}, c={p})")
        if out.dtype != self.dtype or not out.flags.c_contiguous:
            raise TypeError("sp_tmul(): out must be a C-contiguous (L, n) array of the matrix dtype")
        indptr = np.ascontiguousarray(v.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(v.indices, dtype=np.int64)
        values = np.ascontiguousarray(v.data, dtype=self.dtype)
        _lib.check(_lib.load().ab_matrix_sp_tmul(self._core(), v.shape[0], _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(values), _lib.ptr(out)))

    def mean(self, weights, out):
        """Weighted column means X^T w (reference: matrix base ``mean``)."""
        self.mul(np.ones(self.rows(), dtype=self.dtype), weights, out)

    def var(self, centers, weights, out):
        """Weighted column variances around ``centers``."""
        sq = np.empty(self.cols(), dtype=self.dtype)
        self.sq_mul(weights, sq)
        m = np.empty(self.cols(), dtype=self.dtype)
        self.mean(weights, m)
        out[...] = sq - 2 * centers * m + centers ** 2 * np.sum(weights)


class _Dense(_DeviceMatrix):
    def __init__(self, mat, n_threads):
        _DeviceMatrix.__init__(self, mat.dtype, mat.shape[0], mat.shape[1], n_threads)
        self._mat = mat          # keep the host array alive (adelie/matrix.py:674-678)

    def _make_handle(self):
        m = self._mat
        h = C.c_void_p()
        order = 0 if m.flags.f_contiguous else 1
        ldh = m.shape[0] if order == 0 else m.shape[1]
        _lib.check(_lib.load().ab_matrix_dense_create(_lib.dtype_code(self.dtype), _lib.ptr(m), m.shape[0], m.shape[1], order, ldh,
                                                      self._n_threads, C.byref(h)))
        return h


class _DeviceDense(_DeviceMatrix):
    """Dense matrix generated directly in HBM (bench / sharded synthetic data)."""
    def __init__(self, dtype, n, p, seed, row_offset, n_threads=1):
        _DeviceMatrix.__init__(self, dtype, n, p, n_threads)
        h = C.c_void_p()
        _lib.check(_lib.load().ab_matrix_dense_alloc(_lib.dtype_code(dtype), n, p, C.byref(h)))
        self._handle = h
        _lib.check(_lib.load().ab_matrix_dense_fill_normal(h, seed, row_offset))

    def to_host(self, row0=0, nrows=None, col0=0, ncols=None):
        nrows = self._n - row0 if nrows is None else nrows
        ncols = self._p - col0 if ncols is None else ncols
        out = np.empty((nrows, ncols), dtype=self.dtype, order="F")
        _lib.check(_lib.load().ab_matrix_dense_download(self._handle, _lib.ptr(out), row0, nrows, col0, ncols, nrows))
        return out


class _Sparse(_DeviceMatrix):
    """Sparse CSC matrix resident in HBM (reference: adelie.matrix.sparse -> MatrixNaiveSparse, CORE/matrix/matrix_naive_sparse.ipp)."""
    def __init__(self, mat, n_threads):
        _DeviceMatrix.__init__(self, mat.dtype, mat.shape[0], mat.shape[1], n_threads)
        self._mat = mat          # keep the host arrays alive (adelie/matrix.py keeps `_mat`)

    def _make_handle(self):
        m = self._mat
        indptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(m.indices, dtype=np.int32)
        values = np.ascontiguousarray(m.data, dtype=self.dtype)
        h = C.c_void_p()
        _lib.check(_lib.load().ab_matrix_sparse_create(_lib.dtype_code(self.dtype), m.shape[0], m.shape[1], int(values.size), _lib.ptr(indptr),
                                                       _lib.ptr(indices), _lib.ptr(values), self._n_threads, C.byref(h)))
        return h


class _DeviceSparse(_DeviceMatrix):
    """Random sparse CSC matrix generated directly in HBM (bench): exactly ``nnz_per_col`` N(0,1) entries per column."""
    def __init__(self, dtype, n, p, nnz_per_col, seed, n_threads=1):
        _DeviceMatrix.__init__(self, dtype, n, p, n_threads)
        h = C.c_void_p()
        _lib.check(_lib.load().ab_matrix_sparse_alloc_random(_lib.dtype_code(dtype), n, p, nnz_per_col, seed, C.byref(h)))
        self._handle = h
        self._nnz = int(p) * int(nnz_per_col)

    def to_host(self):
        """scipy CSC copy of the device matrix."""
        from scipy.sparse import csc_matrix as _csc
        indptr = np.empty(self._p + 1, dtype=np.int64); indices = np.empty(self._nnz, dtype=np.int32); values = np.empty(self._nnz, dtype=self.dtype)
        _lib.check(_lib.load().ab_matrix_sparse_download(self._handle, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(values)))
        return _csc((values, indices, indptr), shape=(self._n, self._p))


class _SnpUnphased(_DeviceMatrix):
    """SNP unphased matrix resident in HBM at 2 bits per genotype (reference: adelie.matrix.snp_unphased -> MatrixNaiveSNPUnphased,
    CORE/matrix/matrix_naive_snp_unphased.ipp): entries 0 / 1 / 2 / impute[j] for a missing value of column j."""
    def __init__(self, dtype, n, p, n_threads, maker, io=None):
        _DeviceMatrix.__init__(self, dtype, n, p, n_threads)
        self._io = io            # keep the IO handler alive like the reference wrapper (adelie/matrix.py:1292-1296)
        self._maker = maker

    def _make_handle(self):
        h = C.c_void_p()
        self._maker(h)
        return h

    def mean(self, weights, out):
        out[...] = 0             # matrix_naive_snp_unphased.ipp:290-298

    def var(self, centers, weights, out):
        out[...] = 1             # matrix_naive_snp_unphased.ipp:300-309

    def to_host(self):
        """``(calldata, impute)``: column-major (n, p) int8 with -9 for missing, and the (p,) imputed values."""
        cd = np.empty((self._n, self._p), dtype=np.int8, order="F"); imp = np.empty(self._p)
        _lib.check(_lib.load().ab_matrix_snp_unphased_download(self._core(), _lib.ptr(cd), _lib.ptr(imp)))
        return cd, imp

    def cache_info(self):
        """``(decoded columns resident in the dense cache, bytes of the packed genotypes)``."""
        a, b = C.c_int64(), C.c_int64()
        _lib.check(_lib.load().ab_matrix_snp_unphased_cache_info(self._core(), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)


def snp_unphased(io, *, n_threads: int = 1, dtype=np.float64, rows=None):
    """SNP unphased matrix from an ``adelie_b200.io.snp_unphased`` handler (adelie/matrix.py:1243-1298).
    ``rows=(lo, hi)`` (extension) keeps only that row range of the file: the ranks of a row-sharded run share one file."""
    if n_threads < 1:
        raise RuntimeError("adelie_core: n_threads must be >= 1.")
    if not io.is_read:
        io.read()
    lo, hi = (0, io.rows) if rows is None else (int(rows[0]), int(rows[1]))

    def maker(h):
        _lib.check(_lib.load().ab_matrix_snp_unphased_create(_lib.dtype_code(dtype), io._handle, lo, hi, int(n_threads), C.byref(h)))
    m = _SnpUnphased(dtype, hi - lo, io.cols, n_threads, maker, io=io)
    m._core()
    return m


def snp_phased_ancestry(io, *, n_threads: int = 1, dtype=np.float64, rows=None):
    """SNP phased, ancestry matrix from an ``adelie_b200.io.snp_phased_ancestry`` handler (adelie/matrix.py ``snp_phased_ancestry``;
    MatrixNaiveSNPPhasedAncestry).  On the device it is the same 2-bit storage as ``snp_unphased`` (entries 0 / 1 / 2), so every kernel
    and the solver path are shared; ``rows=(lo, hi)`` keeps a row window of the file (row-sharded runs)."""
    if n_threads < 1:
        raise RuntimeError("adelie_core: n_threads must be >= 1.")
    if not io.is_read:
        io.read()
    lo, hi = (0, io.rows) if rows is None else (int(rows[0]), int(rows[1]))

    def maker(h):
        _lib.check(_lib.load().ab_matrix_snp_phased_ancestry_create(_lib.dtype_code(dtype), io._handle, lo, hi, int(n_threads), C.byref(h)))
    m = _SnpUnphased(dtype, hi - lo, io.cols, n_threads, maker, io=io)
    m._core()
    return m


def snp_unphased_from_calldata(calldata: np.ndarray, impute: np.ndarray, *, n_threads: int = 1, dtype=np.float64):
    """Same matrix from an in-memory (n, p) int8 calldata array (negative = missing) and the (p,) imputed values, without a file."""
    cd = np.asfortranarray(calldata, dtype=np.int8); imp = np.ascontiguousarray(impute, dtype=np.float64)

    def maker(h):
        _lib.check(_lib.load().ab_matrix_snp_unphased_from_calldata(_lib.dtype_code(dtype), _lib.ptr(cd), cd.shape[0], cd.shape[1], _lib.ptr(imp),
                                                                     int(n_threads), C.byref(h)))
    m = _SnpUnphased(dtype, cd.shape[0], cd.shape[1], n_threads, maker)
    m._core()
    return m


def snp_unphased_device_random(n: int, p: int, *, dtype=np.float32, seed: int = 0, row_offset: int = 0, n_total: int = None,
                               one_ratio: float = 0.25, two_ratio: float = 0.05, missing_ratio: float = 0.1):
    """Random genotypes generated in HBM with a counter-based RNG (proportions of ``adelie.data.snp_unphased``); mean-imputed."""
    n_total = n if n_total is None else n_total

    def maker(h):
        _lib.check(_lib.load().ab_matrix_snp_unphased_alloc_random(_lib.dtype_code(dtype), n, p, seed, row_offset, n_total, one_ratio, two_ratio,
                                                                    missing_ratio, C.byref(h)))
    m = _SnpUnphased(dtype, n, p, 1, maker)
    m._core()
    return m


class _Derived(_DeviceMatrix):
    """Dense device matrix materialised from another device matrix (standardize / subset); keeps the base alive like the reference wrappers."""
    def __init__(self, base, n, p, n_threads, maker):
        _DeviceMatrix.__init__(self, base.dtype, n, p, n_threads)
        self._mat = base
        self._maker = maker

    def _make_handle(self):
        h = C.c_void_p()
        self._maker(h)
        return h


def standardize(mat, centers: np.ndarray = None, scales: np.ndarray = None, ddof: int = 0, *, n_threads: int = 1):
    """Standardized matrix ``(Z - 1 c^T) diag(s)^-1`` (adelie/matrix.py:1414-1536).  ``centers`` / ``scales`` default to the column means and
    standard deviations (``ddof`` degrees of freedom) of ``mat`` under equal weights.  A NumPy input returns a NumPy array like the
    reference; a device matrix returns a new dense device matrix (the transform is materialised once, see csrc/capi.cu)."""
    if isinstance(mat, (list, np.ndarray)):
        mat = np.array(mat, order="F", copy=True)
        if centers is None:
            centers = np.mean(mat, axis=0)
        mat -= centers[None]
        if scales is None:
            scales = np.sqrt(np.sum(mat ** 2, axis=0) / (mat.shape[0] - ddof))
        mat /= scales[None]
        return mat
    if not isinstance(mat, _DeviceMatrix):
        raise RuntimeError("adelie_b200: standardize() takes a numpy array or a dense device matrix.")
    dtype = mat.dtype
    n, p = mat.shape
    weights = np.full(n, 1 / n, dtype=dtype)
    if centers is None:
        centers = np.empty(p, dtype=dtype)
        mat.mean(weights, centers)
    if scales is None:
        v = np.empty(p, dtype=dtype)
        mat.var(np.asarray(centers, dtype=dtype), weights, v)
        scales = np.sqrt((n / (n - ddof)) * v)
    centers = np.array(centers, copy=True, dtype=dtype); scales = np.array(scales, copy=True, dtype=dtype)

    def maker(h):
        _lib.check(_lib.load().ab_matrix_standardize_create(mat._core(), _lib.ptr(centers), centers.size, _lib.ptr(scales), scales.size,
                                                            int(n_threads), C.byref(h)))
    out = _Derived(mat, n, p, n_threads, maker)
    out._centers, out._scales = centers, scales
    out._core()
    return out


def subset(mat, indices: np.ndarray, *, axis: int = 0, n_threads: int = 1):
    """``mat[indices]`` (axis 0) or ``mat[:, indices]`` (axis 1) (adelie/matrix.py:1539-1632); device matrices give a new dense device matrix."""
    if isinstance(mat, np.ndarray):
        return mat[indices] if axis == 0 else mat[:, indices]
    if not isinstance(mat, _DeviceMatrix):
        raise RuntimeError("adelie_b200: subset() takes a numpy array or a dense device matrix.")
    idx = np.array(indices, copy=True, dtype=np.int64).ravel()
    n, p = mat.shape

    def maker(h):
        _lib.check(_lib.load().ab_matrix_subset_create(mat._core(), _lib.ptr(idx), idx.size, int(axis), int(n_threads), C.byref(h)))
    out = _Derived(mat, idx.size if axis == 0 else n, p if axis == 0 else idx.size, n_threads, maker)
    out._indices = idx
    out._core()
    return out



class MatrixCovBase:
    """Covariance-method matrix (reference: adelie/matrix.py MatrixCovBase{32,64}; CORE/matrix/matrix_cov_base.hpp:20-63): a (p, p) positive
    semi-definite matrix behind ``bmul`` / ``mul`` / ``to_dense`` / ``cols``.  Resident in HBM; the device copy is made on first use."""
    def __init__(self, dtype, p, n_threads):
        if n_threads < 1:
            raise RuntimeError("adelie_core: n_threads must be >= 1.")
        self.dtype = np.dtype(dtype).type
        self._p = int(p)
        self._n_threads = n_threads
        self._handle = None

    def cols(self):
        return self._p

    def rows(self):
        return self._p

    @property
    def ndim(self):
        return 2

    @property
    def shape(self):
        return (self._p, self._p)

    def _make_handle(self):
        raise NotImplementedError

    def _core(self):
        if self._handle is None:
            self._handle = self._make_handle()
        return self._handle

    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _lib.load().ab_matrix_cov_free(h)
            except Exception:
                pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        self.close()

    def _idx(self, a):
        return np.ascontiguousarray(a, dtype=np.int64)

    def _vec(self, a, name, fn):
        if not isinstance(a, np.ndarray) or a.dtype != self.dtype or a.ndim != 1 or not a.flags.c_contiguous:
            raise TypeError(f"{fn}(): {name} must be a 1-D contiguous array of dtype {np.dtype(self.dtype).name}")
        return a

    def bmul(self, subset, indices, values, out):
        """out[k] = sum_i values[i] * A[indices[i], subset[k]] (matrix_cov_base.hpp:33-47)."""
        subset = self._idx(subset); indices = self._idx(indices)
        values = self._vec(values, "values", "bmul"); out = self._vec(out, "out", "bmul")
        p = self._p
        s, i, v, o = subset.size, indices.size, values.size, out.size
        if (s > p) or (i > p) or (i != v) or (o != s):
            raise RuntimeError(f"adelie_core: bmul() is given inconsistent inputs! Invoked check_bmul(s={s}, i={i}, v={v}, o={o}, r={p}, c={p})")
        _lib.check(_lib.load().ab_matrix_cov_bmul(self._core(), _lib.ptr(subset), s, _lib.ptr(indices), _lib.ptr(values), i, _lib.ptr(out)))

    def mul(self, indices, values, out):
        """out = A[indices].T @ values (matrix_cov_base.hpp:49-57)."""
        indices = self._idx(indices)
        values = self._vec(values, "values", "mul"); out = self._vec(out, "out", "mul")
        p = self._p
        i, v, o = indices.size, values.size, out.size
        if (i > p) or (i != v) or (o != p):
            raise RuntimeError(f"adelie_core: mul() is given inconsistent inputs! Invoked check_mul(i={i}, v={v}, o={o}, r={p}, c={p})")
        _lib.check(_lib.load().ab_matrix_cov_mul(self._core(), _lib.ptr(indices), _lib.ptr(values), i, _lib.ptr(out)))

    def to_dense(self, i, p, out):
        """out = A[i:i+p, i:i+p] (matrix_cov_base.hpp:59-62); out is an F-contiguous (p, p) array."""
        r = self._p
        if (i < 0 or i > r - p) or out.shape != (p, p):
            raise RuntimeError(f"adelie_core: to_dense() is given inconsistent inputs! Invoked check_to_dense(i={i}, p={p}, o_r={out.shape[0]}, o_c={out.shape[-1]}, r={r}, c={r})")
        if out.dtype != self.dtype or not out.flags.f_contiguous:
            raise TypeError("to_dense(): out must be an F-contiguous (p, p) array of the matrix dtype")
        _lib.check(_lib.load().ab_matrix_cov_to_dense(self._core(), i, p, _lib.ptr(out)))


class MatrixCovBase32(MatrixCovBase):
    pass


class MatrixCovBase64(MatrixCovBase):
    pass


def _cov_class(base, dtype):
    mark = MatrixCovBase64 if np.dtype(dtype) == np.float64 else MatrixCovBase32
    return type(base.__name__, (base, mark), {})


class _CovDense(MatrixCovBase):
    """adelie.matrix.dense(method="cov") -> MatrixCovDense{32,64}{C,F} (CORE/matrix/matrix_cov_dense.ipp:8-84)."""
    def __init__(self, mat, n_threads):
        if mat.shape[0] != mat.shape[1]:
            raise RuntimeError("adelie_core: mat must be (p, p).")
        MatrixCovBase.__init__(self, mat.dtype, mat.shape[1], n_threads)
        self._mat = mat

    def _make_handle(self):
        m = self._mat
        h = C.c_void_p()
        order = 0 if m.flags.f_contiguous else 1
        _lib.check(_lib.load().ab_matrix_cov_dense_create(_lib.dtype_code(self.dtype), _lib.ptr(m), m.shape[1], order, m.shape[0],
                                                          self._n_threads, C.byref(h)))
        return h


class _CovLazy(MatrixCovBase):
    """adelie.matrix.lazy_cov -> MatrixCovLazyCov{32,64}{C,F} (CORE/matrix/matrix_cov_lazy_cov.ipp:8-190): A = X^T X, rows of A are computed
    on the device the first time they are needed and kept in HBM."""
    def __init__(self, mat, n_threads):
        MatrixCovBase.__init__(self, mat.dtype, mat.shape[1], n_threads)
        self._mat = mat

    def _make_handle(self):
        m = self._mat
        h = C.c_void_p()
        order = 0 if m.flags.f_contiguous else 1
        ldh = m.shape[0] if order == 0 else m.shape[1]
        _lib.check(_lib.load().ab_matrix_cov_lazy_create(_lib.dtype_code(self.dtype), _lib.ptr(m), m.shape[0], m.shape[1], order, ldh,
                                                         self._n_threads, C.byref(h)))
        return h

    def cached_rows(self):
        out = C.c_int64()
        _lib.check(_lib.load().ab_matrix_cov_cache_info(self._core(), C.byref(out)))
        return out.value


def _check_dense_input(mat):
    if not isinstance(mat, np.ndarray) or mat.ndim != 2:
        raise RuntimeError("mat must be a 2-dimensional numpy array.")
    if mat.dtype not in (np.float32, np.float64):
        raise RuntimeError("mat must be of type numpy.float32 or numpy.float64.")
    if not (mat.flags.f_contiguous or mat.flags.c_contiguous):
        mat = np.asfortranarray(mat)
    return mat


def lazy_cov(mat: np.ndarray, *, copy: bool = False, n_threads: int = 1):
    """Lazy covariance matrix ``mat.T @ mat`` (adelie/matrix.py:1003-1080).  Only works with the covariance method."""
    mat = _check_dense_input(mat)
    if copy:
        mat = mat.copy(order="K")
    return _cov_class(_CovLazy, mat.dtype)(mat, n_threads)


def block_diag(mats: list, *, method: str = "naive", n_threads: int = 1):
    """Block-diagonal matrix of the given matrices (adelie/matrix.py:198-290).  Only ``method="cov"`` (MatrixCovBlockDiag{32,64},
    CORE/matrix/matrix_cov_block_diag.ipp:8-211) is built: the device solver addresses A through full rows, so the blocks are assembled
    into one dense (p, p) device matrix from each block's ``to_dense`` (memory p^2 instead of the sum of the blocks' squares)."""
    if method != "cov":
        raise RuntimeError("adelie_b200: block_diag is only available with method='cov'.")
    mats = [dense(m, method="cov", n_threads=1) if isinstance(m, np.ndarray) else m for m in mats]
    if len(mats) == 0:
        raise RuntimeError("mats must be non-empty.")
    dtype = mats[0].dtype
    for m in mats:
        if not isinstance(m, MatrixCovBase) or m.dtype != dtype:
            raise RuntimeError("All matrices must be covariance matrices of the same underlying data type.")
    sizes = [m.cols() for m in mats]
    p = int(np.sum(sizes))
    A = np.zeros((p, p), dtype=dtype, order="F")
    off = 0
    for m, q in zip(mats, sizes):
        if isinstance(m, _CovDense):
            blk = m._mat                       # host copy at hand: no device round trip
        else:
            blk = np.empty((q, q), dtype=dtype, order="F")
            m.to_dense(0, q, blk)
        A[off:off + q, off:off + q] = blk
        off += q
    out = _cov_class(_CovDense, dtype)(A, n_threads)
    out._mats = mats
    return out


def _sparse_cov(mat, copy, n_threads):
    """adelie.matrix.sparse(method="cov") -> MatrixCovSparse{32,64}F (CORE/matrix/matrix_cov_sparse.ipp:8-92): a sparse PSD matrix.  The
    device solver reads full rows of A, so the matrix is densified on upload (p^2 elements in HBM)."""
    import scipy.sparse as _sp
    if not _sp.issparse(mat):
        raise RuntimeError("mat must be a scipy sparse matrix.")
    if mat.dtype not in (np.float32, np.float64):
        raise RuntimeError("mat must be of type numpy.float32 or numpy.float64.")
    if mat.shape[0] != mat.shape[1]:
        raise RuntimeError("adelie_core: mat must be (p, p).")
    return _cov_class(_CovDense, mat.dtype)(np.asfortranarray(mat.toarray()), n_threads)


def sparse(mat, *, method: str = "naive", copy: bool = False, n_threads: int = 1):
    """Sparse matrix (adelie/matrix.py ``sparse``): a scipy CSC matrix (anything else is converted), float32 / float64."""
    import scipy.sparse as _sp
    if method == "cov":
        return _sparse_cov(mat, copy, n_threads)
    if method != "naive":
        raise RuntimeError("method must be one of 'naive', 'cov'.")
    if not _sp.issparse(mat):
        raise RuntimeError("mat must be a scipy sparse matrix.")
    if mat.dtype not in (np.float32, np.float64):
        raise RuntimeError("mat must be of type numpy.float32 or numpy.float64.")
    if not _sp.isspmatrix_csc(mat):
        warnings.warn("Converting to CSC format.")
        mat = mat.tocsc(copy=True)
    elif copy or not mat.has_sorted_indices or not mat.has_canonical_format:
        mat = mat.copy()
    mat.sum_duplicates()
    mat.sort_indices()
    return _Sparse(mat, n_threads)


def sparse_device_random(n: int, p: int, nnz_per_col: int, *, dtype=np.float32, seed: int = 0):
    """Random sparse CSC matrix generated in HBM with a counter-based RNG (no host copy)."""
    return _DeviceSparse(dtype, n, p, nnz_per_col, seed)


def dense(mat: np.ndarray, *, method: str = "naive", copy: bool = False, n_threads: int = 1):
    """Dense matrix (adelie/matrix.py:549-680): ``method="naive"`` -> MatrixNaiveDense, ``method="cov"`` -> MatrixCovDense."""
    if method == "cov":
        mat = _check_dense_input(mat)
        if copy:
            mat = mat.copy(order="K")
        return _cov_class(_CovDense, mat.dtype)(mat, n_threads)
    if method != "naive":
        raise RuntimeError("method must be one of 'naive', 'cov'.")
    if not isinstance(mat, np.ndarray) or mat.ndim != 2:
        raise RuntimeError("mat must be a 2-dimensional numpy array.")
    if mat.dtype not in (np.float32, np.float64):
        raise RuntimeError("mat must be of type numpy.float32 or numpy.float64.")
    if not (mat.flags.f_contiguous or mat.flags.c_contiguous):
        mat = np.asfortranarray(mat)
    elif mat.flags.c_contiguous and not mat.flags.f_contiguous:
        warnings.warn("Detected matrix to be C-contiguous. Performance may improve with F-contiguous matrix.")
    if copy:
        mat = mat.copy(order="K")
    return _Dense(mat, n_threads)


def dense_device_normal(n: int, p: int, *, dtype=np.float32, seed: int = 0, row_offset: int = 0):
    """N(0,1) dense matrix generated in HBM with a counter-based RNG (no host copy)."""
    return _DeviceDense(dtype, n, p, seed, row_offset)


class _KroneckerEye(MatrixNaiveBase):
    """kron(mat, I_K) as a layout rule over the base operators (adelie/matrix.py kronecker_eye;
    CORE/matrix/matrix_naive_kronecker_eye.ipp): column j <-> (feature j // K, class j % K), row r <-> (obs r // K, class r % K).
    The fused solver never materialises this object -- it addresses the base matrix with (feature, class) arithmetic in the
    kernels (csrc/sweep.cuh); this class exists for the user-facing operator API and the initial invariants."""
    def __init__(self, mat, K, n_threads=1):
        MatrixNaiveBase.__init__(self, n_threads)
        if isinstance(mat, np.ndarray):
            mat = dense(mat, method="naive", n_threads=n_threads)
        self._mat = mat
        self._K = int(K)
        self.dtype = mat.dtype

    def rows(self):
        return self._mat.rows() * self._K

    def cols(self):
        return self._mat.cols() * self._K

    def _sl(self, a, l):
        return np.ascontiguousarray(a[l::self._K])

    def cmul(self, j, v, weights):
        i, l = divmod(int(j), self._K)
        return self._mat.cmul(i, self._sl(v, l), self._sl(weights, l))

    cmul_safe = cmul

    def ctmul(self, j, v, out):
        i, l = divmod(int(j), self._K)
        tmp = np.zeros(self._mat.rows(), dtype=self.dtype)
        self._mat.ctmul(i, v, tmp)
        out[l::self._K] += tmp

    def bmul(self, j, q, v, weights, out):
        if j < 0 or j > self.cols() - q or v.size != self.rows() or weights.size != self.rows() or out.size != q:
            raise RuntimeError(f"adelie_core: bmul() is given inconsistent inputs! (j={j}, q={q}, v={v.size}, w={weights.size}, o={out.size}, r={self.rows()}, c={self.cols()})")
        for c in range(q):
            out[c] = self.cmul(j + c, v, weights)

    bmul_safe = bmul

    def btmul(self, j, q, v, out):
        if j < 0 or j > self.cols() - q or v.size != q or out.size != self.rows():
            raise RuntimeError(f"adelie_core: btmul() is given inconsistent inputs! (j={j}, q={q}, v={v.size}, o={out.size}, r={self.rows()}, c={self.cols()})")
        for c in range(q):
            self.ctmul(j + c, v[c], out)

    def mul(self, v, weights, out):
        base = self._mat
        if (isinstance(base, _DeviceMatrix) and not isinstance(base, (_Sparse, _DeviceSparse)) and self._K <= 16 and isinstance(out, np.ndarray)
                and out.dtype == self.dtype and out.flags.c_contiguous):
            # one device pass for all K classes (packed genotypes: the INT8 tensor-core kernel)
            v = np.ascontiguousarray(v, dtype=self.dtype); weights = np.ascontiguousarray(weights, dtype=self.dtype)
            _lib.check(_lib.load().ab_matrix_mul_multi(base._core(), self._K, _lib.ptr(v), _lib.ptr(weights), _lib.ptr(out)))
            return
        tmp = np.empty(base.cols(), dtype=self.dtype)
        for l in range(self._K):
            base.mul(self._sl(v, l), self._sl(weights, l), tmp)
            out[l::self._K] = tmp

    def cov(self, j, q, sqrt_weights, out):
        K = self._K
        out[...] = 0
        i0, i1 = j // K, (j + q - 1) // K + 1
        for l in range(K):
            idx = [c for c in range(q) if (j + c) % K == l]
            if not idx:
                continue
            sub = np.empty((i1 - i0, i1 - i0), dtype=self.dtype, order="F")
            self._mat.cov(i0, i1 - i0, self._sl(sqrt_weights, l), sub)
            f = [(j + c) // K - i0 for c in idx]
            out[np.ix_(idx, idx)] = sub[np.ix_(f, f)]

    def sq_mul(self, weights, out):
        tmp = np.empty(self._mat.cols(), dtype=self.dtype)
        for l in range(self._K):
            self._mat.sq_mul(self._sl(weights, l), tmp)
            out[l::self._K] = tmp

    def sp_tmul(self, v, out):
        v = csc_matrix(v)
        tmp = np.empty((v.shape[0], self._mat.rows()), dtype=self.dtype)
        for l in range(self._K):
            self._mat.sp_tmul(csr_matrix(v[:, l::self._K]), tmp)
            out[:, l::self._K] = tmp

    def mean(self, weights, out):
        self.mul(np.ones(self.rows(), dtype=self.dtype), weights, out)


class _CConcatenate(MatrixNaiveBase):
    """Column-wise concatenation as a dispatch rule (adelie/matrix.py concatenate(axis=1); CORE/matrix/matrix_naive_concatenate.ipp)."""
    def __init__(self, mats, n_threads=1):
        MatrixNaiveBase.__init__(self, n_threads)
        if len(mats) == 0:
            raise RuntimeError("adelie_core: mat_list must be non-empty.")
        mats = [dense(m, method="naive", n_threads=n_threads) if isinstance(m, np.ndarray) else m for m in mats]
        n = mats[0].rows()
        if any(m.rows() != n for m in mats):
            raise RuntimeError("adelie_core: All matrices must have the same number of rows.")
        self._mats = mats
        self.dtype = mats[0].dtype
        self._offs = np.concatenate([[0], np.cumsum([m.cols() for m in mats])]).astype(int)

    def rows(self):
        return self._mats[0].rows()

    def cols(self):
        return int(self._offs[-1])

    def _find(self, j):
        k = int(np.searchsorted(self._offs, j, side="right") - 1)
        return k, int(j - self._offs[k])

    def cmul(self, j, v, weights):
        if j < 0 or j >= self.cols():
            raise RuntimeError(f"adelie_core: cmul() is given inconsistent inputs! (j={j}, v={v.size}, w={weights.size}, r={self.rows()}, c={self.cols()})")
        k, jj = self._find(j)
        return self._mats[k].cmul(jj, v, weights)

    cmul_safe = cmul

    def ctmul(self, j, v, out):
        k, jj = self._find(j)
        self._mats[k].ctmul(jj, v, out)

    def bmul(self, j, q, v, weights, out):
        if j < 0 or j > self.cols() - q or out.size != q:
            raise RuntimeError(f"adelie_core: bmul() is given inconsistent inputs! (j={j}, q={q}, v={v.size}, w={weights.size}, o={out.size}, r={self.rows()}, c={self.cols()})")
        c = 0
        while c < q:
            k, jj = self._find(j + c)
            qq = min(q - c, self._mats[k].cols() - jj)
            tmp = np.empty(qq, dtype=self.dtype)
            self._mats[k].bmul(jj, qq, v, weights, tmp)
            out[c:c + qq] = tmp
            c += qq

    bmul_safe = bmul

    def btmul(self, j, q, v, out):
        if j < 0 or j > self.cols() - q or v.size != q:
            raise RuntimeError(f"adelie_core: btmul() is given inconsistent inputs! (j={j}, q={q}, v={v.size}, o={out.size}, r={self.rows()}, c={self.cols()})")
        c = 0
        while c < q:
            k, jj = self._find(j + c)
            qq = min(q - c, self._mats[k].cols() - jj)
            self._mats[k].btmul(jj, qq, np.ascontiguousarray(v[c:c + qq]), out)
            c += qq

    def mul(self, v, weights, out):
        for k, m in enumerate(self._mats):
            tmp = np.empty(m.cols(), dtype=self.dtype)
            m.mul(v, weights, tmp)
            out[self._offs[k]:self._offs[k + 1]] = tmp

    def cov(self, j, q, sqrt_weights, out):
        k, jj = self._find(j)
        if jj + q > self._mats[k].cols():
            raise RuntimeError("adelie_core: MatrixNaiveCConcatenate::cov() only allows the block to be fully contained in one of the matrices in the list.")
        self._mats[k].cov(jj, q, sqrt_weights, out)

    def sq_mul(self, weights, out):
        for k, m in enumerate(self._mats):
            tmp = np.empty(m.cols(), dtype=self.dtype)
            m.sq_mul(weights, tmp)
            out[self._offs[k]:self._offs[k + 1]] = tmp

    def sp_tmul(self, v, out):
        v = csc_matrix(v)
        out[...] = 0
        tmp = np.empty(out.shape, dtype=self.dtype)
        for k, m in enumerate(self._mats):
            m.sp_tmul(csr_matrix(v[:, self._offs[k]:self._offs[k + 1]]), tmp)
            out += tmp

    def mean(self, weights, out):
        self.mul(np.ones(self.rows(), dtype=self.dtype), weights, out)


def kronecker_eye(mat, K: int = 1, *, n_threads: int = 1):
    """kron(mat, I_K) (adelie/matrix.py kronecker_eye)."""
    return _KroneckerEye(mat, K, n_threads)


def concatenate(mats: list, axis: int = 0, *, n_threads: int = 1):
    """Concatenation of naive matrices (adelie/matrix.py concatenate).  Only ``axis=1`` (the multi-response intercept layout,
    adelie/state.py:1100-1125) is on the hot path."""
    if axis != 1:
        raise RuntimeError("adelie_b200: only axis=1 concatenation is in scope (row concatenation is not on the hot path).")
    return _CConcatenate(mats, n_threads)
