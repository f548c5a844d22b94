"""Covariance-method solver: one 100-lambda path on the device next to the CPU oracle (same inputs).
Usage: python scripts/bench_cov.py [p] [gs] [n]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad

p = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
gs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
rng = np.random.default_rng(0)
X = rng.standard_normal((n, p)).astype(np.float32)
beta = np.zeros(p, dtype=np.float32); idx = rng.choice(p, p // 20, replace=False); beta[idx] = rng.standard_normal(idx.size)
y = X @ beta + np.linalg.norm(beta) * rng.standard_normal(n).astype(np.float32)
A = np.asfortranarray((X.T @ X) / n).astype(np.float32)
v = (X.T @ y / n).astype(np.float32)
groups = np.arange(0, p, gs)
if os.environ.get("COV_PROF", "0") == "1":
    ad.configs.set_configs("sweep_profile", 1)
kw = dict(groups=groups, tol=1e-7, newton_tol=1e-6, early_exit=False, min_ratio=1e-2, lmda_path_size=100)
for rep in range(int(os.environ.get("COV_REPS", "2"))):
    Ad = ad.matrix.dense(A, method="cov")
    t0 = time.time()
    st = ad.gaussian_cov(A=Ad, v=v, progress_bar=False, **kw)
    t1 = time.time()
    print(f"gpu rep {rep}: {t1 - t0:.3f} s total, kernel {st.time_sweep_kernel:.3f} s, sweeps {st.n_sweeps}, group updates {st.n_group_updates}, "
          f"{st.n_group_updates / max(st.time_sweep_kernel, 1e-9):.0f} updates/s, screen {len(st.screen_set)}, active {st.active_set_size}, "
          f"cluster {st.cov_cluster}, err={st.error!r}")
    if os.environ.get("COV_PROF", "0") == "1":
        ss = st.sweep_stats
        names = ["prox", "sync", "update_push", "barrier"]
        print("  cycles per group update:", {k: round(ss[i] / max(ss[6], 1)) for i, k in enumerate(names)}, "groups", int(ss[6]))
    Ad.close()
if os.environ.get("COV_CPU", "1") == "1":
    from oracle import oracle as orc
    t0 = time.time()
    so = orc.gaussian_cov(A, v, **kw)
    t1 = time.time()
    print(f"cpu oracle: {t1 - t0:.3f} s, sweeps {so.n_sweeps}, group updates {so.n_group_updates}")
    B1, B2 = st.betas.toarray(), so.betas.toarray()
    print("max rel diff", float(np.max(np.abs(B1 - B2)) / np.max(np.abs(B2))))
