"""Dev check: batched look-ahead sweep kernel vs the per-group kernel (and the oracle) on small/medium problems."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import adelie_b200 as ad


def rel(a, b):
    s = np.max(np.abs(b)); return np.max(np.abs(a - b)) / (s if s > 0 else 1)


def run(n, p, G, dtype, ctas, batch, alpha=1.0, equal=True, L=20, intercept=True):
    ad.set_configs("sweep_ctas", ctas); ad.set_configs("sweep_batch", batch)
    data = ad.data.dense(n, p, G, seed=3, equal_groups=equal)
    X = np.asfortranarray(data["X"], dtype=dtype); y = data["glm"].y.astype(dtype)
    tol = 1e-12 if dtype == np.float64 else 1e-7
    nt = 1e-12 if dtype == np.float64 else 1e-5
    kw = dict(groups=data["groups"], alpha=alpha, penalty=data["penalty"].astype(dtype), intercept=intercept, tol=tol, early_exit=False,
              lmda_path_size=L, min_ratio=0.05, newton_tol=nt)
    t0 = time.time()
    st = ad.grpnet(X, ad.glm.gaussian(y, dtype=dtype), progress_bar=False, **kw)
    return st, time.time() - t0


cases = [(2000, 64, 64, np.float64, 1), (4096, 120, 12, np.float64, 4), (4096, 120, 12, np.float32, 8), (20000, 400, 40, np.float32, 0),
         (300, 120, 25, np.float64, 1), (65536, 600, 60, np.float32, 0)]
for (n, p, G, dt, ctas) in cases:
    equal = not (n == 300)
    ref, t_ref = run(n, p, G, dt, ctas, 1, equal=equal)
    for batch in (0, 2, 6):
        st, t = run(n, p, G, dt, ctas, batch, equal=equal)
        e = rel(np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())) if (st.error == "" and ref.error == "") else float("nan")
        print(f"n={n} p={p} G={G} {np.dtype(dt).name} ctas={ctas} batch={batch}: err='{st.error}' rel={e:.2e} sweeps {st.n_sweeps} vs {ref.n_sweeps} "
              f"updates {st.n_group_updates} vs {ref.n_group_updates} B={st.sweep_batch} launches(batched)={st.n_batched_launches} "
              f"tk {st.time_sweep_kernel*1e3:.1f}ms vs {ref.time_sweep_kernel*1e3:.1f}ms", flush=True)
ad.set_configs("sweep_ctas", None); ad.set_configs("sweep_batch", None)
