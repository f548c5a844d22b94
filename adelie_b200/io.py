"""``adelie.io`` surface for the hot path: the ``.snpdat`` handler of SNP unphased matrices.

Reference: ``adelie.io.snp_unphased`` (adelie/io.py:114-196) over ``IOSNPUnphased`` (adelie/src/py_io.cpp;
adelie/src/include/adelie_core/io/io_snp_unphased.{hpp,ipp}, io_snp_base.ipp).  The reader / writer are host C++ inside
``libadelie_b200.so`` (csrc/snp.cuh ``SnpUnphasedIO``); the chunk lists themselves are unpacked on the device when a matrix is
built from the handler (``adelie_b200.matrix.snp_unphased``).
"""
from __future__ import annotations

import ctypes as C
from typing import Union

import numpy as np

from . import _lib


class snp_unphased:
    """IO handler for a SNP unphased matrix (entries 0, 1, 2 or NA; any negative value is NA) stored in ``.snpdat`` format.

    Parameters: ``filename``; ``read_mode`` in ``"file"`` (default), ``"mmap"``, ``"auto"``.
    """
    def __init__(self, filename: str, read_mode: str = "file"):
        self._filename = str(filename)
        self._read_mode = str(read_mode)
        h = C.c_void_p()
        _lib.check(_lib.load().ab_io_snp_unphased_create(self._filename.encode(), self._read_mode.encode(), C.byref(h)))
        self._handle = h

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.load().ab_io_snp_unphased_free(self._handle)
        except Exception:
            pass

    # ---- writer / reader
    def write(self, calldata: np.ndarray, impute_method: Union[str, np.ndarray] = "mean", n_threads: int = 1):
        """Serialises a dense (n, p) int8 matrix; returns ``(total_bytes, benchmark)`` (adelie/io.py:146-196)."""
        if isinstance(impute_method, str):
            impute = np.empty(calldata.shape[1])
        elif isinstance(impute_method, np.ndarray):
            impute = np.ascontiguousarray(impute_method, dtype=np.float64)
            impute_method = "user"
        else:
            raise ValueError("impute_method must be a valid option.")
        if not isinstance(calldata, np.ndarray) or calldata.ndim != 2 or calldata.dtype != np.int8:
            raise TypeError("calldata must be a 2-dimensional int8 numpy array.")
        cd = np.asfortranarray(calldata)
        total = C.c_uint64()
        _lib.check(_lib.load().ab_io_snp_unphased_write(self._handle, _lib.ptr(cd), cd.shape[0], cd.shape[1], impute_method.encode(),
                                                        _lib.ptr(impute), impute.size, int(n_threads), C.byref(total)))
        return int(total.value), {}

    def read(self) -> int:
        """Reads (or maps) the file and parses the header; returns the number of bytes."""
        total = C.c_uint64()
        _lib.check(_lib.load().ab_io_snp_unphased_read(self._handle, C.byref(total)))
        return int(total.value)

    # ---- properties (py_io.cpp)
    def _info(self):
        r, n, p = C.c_int(), C.c_int64(), C.c_int64()
        _lib.check(_lib.load().ab_io_snp_unphased_info(self._handle, C.byref(r), C.byref(n), C.byref(p)))
        return bool(r.value), int(n.value), int(p.value)

    def _need_read(self):
        if not self._info()[0]:
            raise RuntimeError("adelie_core: File is not read yet. Call read() first.")

    def _get(self, name, dtype, extra=0):
        self._need_read()
        out = np.empty(self._info()[2] + extra, dtype=dtype)
        _lib.check(_lib.load().ab_io_snp_unphased_get(self._handle, name.encode(), _lib.ptr(out)))
        return out

    @property
    def is_read(self) -> bool:
        return self._info()[0]

    @property
    def rows(self) -> int:
        self._need_read()
        return self._info()[1]

    @property
    def snps(self) -> int:
        self._need_read()
        return self._info()[2]

    @property
    def cols(self) -> int:
        return self.snps

    @property
    def nnz(self):
        return self._get("nnz", np.uint64)

    @property
    def nnm(self):
        return self._get("nnm", np.uint64)

    @property
    def impute(self):
        return self._get("impute", np.float64)

    @property
    def outer(self):
        return self._get("outer", np.uint64, 1)

    def to_dense(self, n_threads: int = 1):
        """(n, p) int8 array with ``-9`` for missing entries."""
        self._need_read()
        _, n, p = self._info()
        out = np.empty((n, p), dtype=np.int8)
        _lib.check(_lib.load().ab_io_snp_unphased_to_dense(self._handle, int(n_threads), _lib.ptr(out)))
        return out


class snp_phased_ancestry:
    """IO handler for a SNP phased, ancestry matrix in ``.snpdat`` format (adelie/io.py:6-111; IOSNPPhasedAncestry,
    adelie_core/io/io_snp_phased_ancestry.{hpp,ipp}).  ``calldata[i, 2 j + k]`` in {0, 1} is the mutation indicator of individual i, SNP j,
    haplotype k and ``ancestries[i, 2 j + k]`` in [0, A) its ancestry label; the matrix is (n, s A) with entry (i, j A + a) = number
    of haplotypes of SNP j that carry the mutation and are labelled a."""
    def __init__(self, filename: str, read_mode: str = "file"):
        self._filename = str(filename)
        h = C.c_void_p()
        _lib.check(_lib.load().ab_io_snp_phased_ancestry_create(self._filename.encode(), str(read_mode).encode(), C.byref(h)))
        self._handle = h

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.load().ab_io_snp_phased_ancestry_free(self._handle)
        except Exception:
            pass

    def write(self, calldata: np.ndarray, ancestries: np.ndarray, A: int, n_threads: int = 1):
        """Serialises dense (n, 2 s) int8 calldata / ancestries; returns ``(total_bytes, benchmark)``."""
        for nm, a in (("calldata", calldata), ("ancestries", ancestries)):
            if not isinstance(a, np.ndarray) or a.ndim != 2 or a.dtype != np.int8:
                raise TypeError(f"{nm} must be a 2-dimensional int8 numpy array.")
        if calldata.shape != ancestries.shape:
            raise RuntimeError("adelie_core: calldata and ancestries must have shape (n, 2*s).")
        cd = np.asfortranarray(calldata); an = np.asfortranarray(ancestries)
        total = C.c_uint64()
        _lib.check(_lib.load().ab_io_snp_phased_ancestry_write(self._handle, _lib.ptr(cd), _lib.ptr(an), cd.shape[0], cd.shape[1], int(A),
                                                               int(n_threads), C.byref(total)))
        return int(total.value), {}

    def read(self) -> int:
        total = C.c_uint64()
        _lib.check(_lib.load().ab_io_snp_phased_ancestry_read(self._handle, C.byref(total)))
        return int(total.value)

    def _info(self):
        r, n, s, a = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.load().ab_io_snp_phased_ancestry_info(self._handle, C.byref(r), C.byref(n), C.byref(s), C.byref(a)))
        return bool(r.value), int(n.value), int(s.value), int(a.value)

    def _need_read(self):
        if not self._info()[0]:
            raise RuntimeError("adelie_core: File is not read yet. Call read() first.")

    def _get(self, name, size):
        self._need_read()
        out = np.empty(size, dtype=np.uint64)
        _lib.check(_lib.load().ab_io_snp_phased_ancestry_get(self._handle, name.encode(), _lib.ptr(out)))
        return out

    is_read = property(lambda self: self._info()[0])

    @property
    def rows(self):
        self._need_read(); return self._info()[1]

    @property
    def snps(self):
        self._need_read(); return self._info()[2]

    @property
    def ancestries(self):
        self._need_read(); return self._info()[3]

    @property
    def cols(self):
        return self.snps * self.ancestries

    nnz0 = property(lambda self: self._get("nnz0", self.cols))
    nnz1 = property(lambda self: self._get("nnz1", self.cols))
    outer = property(lambda self: self._get("outer", self.snps + 1))

    def to_dense(self, n_threads: int = 1):
        """(n, s A) int8 array of the matrix entries (0, 1 or 2)."""
        self._need_read()
        out = np.empty((self.rows, self.cols), dtype=np.int8)
        _lib.check(_lib.load().ab_io_snp_phased_ancestry_to_dense(self._handle, int(n_threads), _lib.ptr(out)))
        return out
