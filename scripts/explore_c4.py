"""Config 4 (sparse CSC Gaussian lasso n=2M p=200k, 0.5% density = 10^4 non-zeros per column, fp32) on one B200."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad

n = int(os.environ.get("N", 2_000_000)); p = int(os.environ.get("P", 200_000)); m = int(os.environ.get("M", 10_000))
L = int(os.environ.get("L", 100)); dtype = np.float32
t = time.time()
X = ad.matrix.sparse_device_random(n, p, m, dtype=dtype, seed=0)
print(f"gen {time.time() - t:.2f}s  nnz={p * m:.3e}  bytes={(p * m * 8) / 1e9:.1f} GB", flush=True)
rng = np.random.default_rng(0)
supp = rng.choice(p, 64, replace=False); bstar = rng.normal(size=64)
eta = np.zeros(n, dtype=dtype)
for j, b in zip(supp, bstar):
    X.btmul(int(j), 1, np.array([b], dtype=dtype), eta)
y = (eta + np.linalg.norm(bstar) * rng.normal(size=n) * np.sqrt(m / n)).astype(dtype)
print("y ready", time.time() - t, flush=True)
for rep in range(2):
    t = time.time()
    st = ad.grpnet(X, ad.glm.gaussian(y, dtype=dtype), early_exit=False, lmda_path_size=L, min_ratio=float(os.environ.get("MINR", 1e-2)),
                   tol=1e-7, progress_bar=False)
    wall = time.time() - t
    print(f"rep {rep}: wall {wall:.3f}s solve {st.total_time:.3f}s err='{st.error}' nl={len(st.lmdas)} sweeps={st.n_sweeps} updates={st.n_group_updates} "
          f"kernel_time={st.time_sweep_kernel:.3f}s active_last={st.active_sizes[-1] if len(st.active_sizes) else 0} screen_last={st.screen_sizes[-1] if len(st.screen_sizes) else 0} dev_last={st.devs[-1]:.4f}")
    if st.time_sweep_kernel > 0:
        print(f"  sweep kernel: {st.n_group_updates / st.time_sweep_kernel:.0f} column updates/s, {st.n_group_updates * m * 8 / st.time_sweep_kernel / 1e9:.1f} GB/s (value+index bytes), "
              f"{st.time_sweep_kernel / max(st.n_group_updates, 1) * 1e6:.2f} us per column update")
    print("  host timers: " + " ".join(f"{k}={getattr(st, 't_' + k):.3f}" for k in ["run_pin", "invariance", "screen_records", "cov_device", "screen_host"]))
