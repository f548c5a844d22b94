"""Cox family on the device at n = 1M (3 strata, tied stop times): CUDA-event time of gradient / hessian / loss per evaluation."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import adelie_b200 as ad
from adelie_b200 import _lib

n = int(os.environ.get("N", 1_000_000))
rng = np.random.default_rng(0)
start = rng.exponential(1.0, n); stop = np.round((start + 0.05 + rng.exponential(1.0, n)) * 1000) / 1000 + 1e-3
status = (rng.uniform(size=n) < 0.6).astype(np.float64); w = rng.uniform(0.1, 1, n); w /= w.sum()
strata = rng.integers(0, 3, n)
t = time.time()
m = ad.glm.cox(start=start, stop=stop, status=status, strata=strata, weights=w)
eta = rng.normal(size=n); g = np.empty(n); h = np.empty(n)
m.gradient(eta, g)
print(f"construction (host sort + tables + upload) {time.time() - t:.2f}s", flush=True)
L = _lib.load()
for name, fn in [("gradient", lambda: m.gradient(eta, g)), ("hessian", lambda: m.hessian(eta, g, h)), ("loss", lambda: m.loss(eta))]:
    fn()
    t = time.time()
    for _ in range(5):
        fn()
    print(f"{name}: {(time.time() - t) / 5 * 1e3:.2f} ms per call through the host-pointer API (incl. H2D/D2H of n-vectors)", flush=True)
print("sum grad", float(np.sum(g)))
