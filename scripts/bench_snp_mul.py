"""Times the packed-genotype `mul` passes at config-5 size (n=500k, p=100k): K = 1 and K = 8, tensor-core (INT8) vs CUDA-core kernels."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad
from adelie_b200 import _lib
n = int(os.environ.get("N", 500_000)); p = int(os.environ.get("P", 100_000))
X = ad.matrix.snp_unphased_device_random(n, p, dtype=np.float32, seed=0)
L = _lib.load()
rng = np.random.default_rng(0)
for K in (1, 8):
    A = X if K == 1 else ad.matrix.kronecker_eye(X, K)
    v = rng.normal(size=n * K).astype(np.float32); w = np.full(n * K, 1.0 / n, dtype=np.float32)
    out = np.empty(p * K, dtype=np.float32)
    for tc, mink in ((0, 2), (1, 1)):
        ad.set_configs("snp_tc", tc); ad.set_configs("snp_tc_min_k", mink)
        A.mul(v, w, out); ref = out.copy()
        _lib.check(L.ab_device_synchronize()); t = time.perf_counter()
        for _ in range(10):
            A.mul(v, w, out)
        _lib.check(L.ab_device_synchronize()); dt = (time.perf_counter() - t) / 10
        print(f"K={K} snp_tc={tc}: {1e3 * dt:.2f} ms per pass (incl. H2D of v, w and D2H of out), checksum {float(np.sum(np.abs(out))):.6e}", flush=True)
