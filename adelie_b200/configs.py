"""Process-global knobs (reference: adelie/configs.py:4-27, CORE/configs.hpp:6-20)."""
from . import _lib
import ctypes as C

_DEFAULTS = {
    "hessian_min": 1e-24,
    "dbeta_tol": 1e-12,
    "min_bytes": float(1 << 17),     # CPU threading threshold of the reference: accepted and ignored on the GPU
    "max_solver_value": 1e100,
    "project": 1.0,
    # B200-specific
    "sweep_ctas": 0.0,
    "sweep_threads": 512.0,
    "sweep_min_rows_per_cta": 1024.0,
    "sweep_force_direct": 0.0,
    "device_eigh": 1.0,
    "sweep_profile": 0.0,
    "glm_batched": 1.0,              # GLM / IRLS pin solves on the batched look-ahead kernel: 0 = off, 1 = float32 states, 2 = every dtype
    "kkt_skip_screen": 1.0,          # invariance / KKT pass streams only the non-screen columns during a path (grad completed at the end)
    "snp_tc": 1.0,                   # packed-genotype multi-response GEMV on the tensor cores (INT8 tcgen05): 1 = on
    "snp_tc_min_k": 1.0,             # ... for at least this many classes
    "panel_tc": 1.0,                 # whole Gram panels on the tensor cores (tcgen05 / TMEM, TF32 operands): 1 = on, 0 = CUDA-core panel kernel
    "panel_gemm": 0.0,               # Gram panels: 1 = whole panels in one pass (fp32; experimental, currently slower), 0 = one block per pair of groups
    "sweep_xchg": 1.0,               # per-group kernel's intra-GPU exchange: 1 = one hop through L2 atomics, 0 = two-level flagged lines
    "sweep_l2_prefetch": 0.0,        # batched sweep kernel: L2 prefetch of the tiles of batch b + 2 (experiment)
    "sweep_u_prefetch": 1.0,         # batched sweep kernel: update tiles of active-set sweeps prefetched ahead of the proximal updates
    "glm_fuse_means": 1.0,           # GLM path: IRLS-weighted column means of the screen groups out of the Gram pass (0 = separate GEMV pass)
    "cov_cluster": 0.0,              # CTAs of the covariance-method solver's thread-block cluster (0 = auto, else 1 / 2 / 4 / 8)
    "sweep_batch": 0.0,              # groups per batch of the look-ahead sweep kernel (0 = auto, 1 = per-group kernel)
}


class Configs:
    """Attribute view of the library configuration (``adelie.configs.Configs``)."""
    def __getattr__(self, name):
        if name.endswith("_def"):
            return _DEFAULTS[name[:-4]]
        if name not in _DEFAULTS:
            raise AttributeError(name)
        out = C.c_double()
        _lib.check(_lib.load().ab_configs_get(name.encode(), C.byref(out)))
        return out.value


Configs = Configs()


def set_configs(name: str, value=None):
    """Sets a configuration; ``None`` restores the default (adelie/configs.py:4-27)."""
    if name not in _DEFAULTS:
        raise RuntimeError(f"adelie_core: unknown config {name}")
    if value is None:
        value = _DEFAULTS[name]
    _lib.check(_lib.load().ab_configs_set(name.encode(), float(value)))
