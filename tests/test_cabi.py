"""CPU: the C-ABI library loads and exports every symbol include/adelie_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "adelie_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ab_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from adelie_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == syms, set(_lib.SYMBOLS) ^ set(syms)


def test_state_args_struct_matches_header():
    """field order / count of the ctypes mirror of ab_state_args follows the header."""
    from adelie_b200 import _lib
    src = open(HEADER).read()
    body = src[src.index("typedef struct ab_state_args {"):src.index("} ab_state_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).replace("typedef struct ab_state_args {", "")
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        parts = decl.split(",")
        for i, part in enumerate(parts):
            nm = part.strip().split()[-1].lstrip("*")
            names.append(nm)
    assert names == [f[0] for f in _lib.StateArgs._fields_]


def test_oracle_library_is_separate_from_product():
    """The product package must never load the oracle: no reference to it anywhere under adelie_b200/."""
    pkg = os.path.join(ROOT, "adelie_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_no_gpu_means_loud_failure():
    """Without a CUDA device every compute entry point raises; there is no CPU fallback."""
    import numpy as np
    import adelie_b200 as ad
    from adelie_b200 import _lib
    c = ctypes.c_int(0)
    has_gpu = _lib.load().ab_device_count(ctypes.byref(c)) == 0 and c.value > 0
    if has_gpu:
        pytest.skip("GPU present")
    X = np.asfortranarray(np.random.RandomState(0).normal(size=(20, 5)))
    y = np.random.RandomState(1).normal(size=20)
    with pytest.raises(RuntimeError):
        ad.grpnet(X, ad.glm.gaussian(y), progress_bar=False)
    with pytest.raises(RuntimeError):
        ad.bcd.solve(quad=np.ones(3), linear=np.ones(3), l1=0.1, l2=0.0)
