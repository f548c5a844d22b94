"""Exploratory timing of config 2 (Gaussian group lasso, dense f32 n=200k p=20k, 2000 groups of 10)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad

n = int(os.environ.get("N", 200000)); p = int(os.environ.get("P", 20000)); gs = 10
L = int(os.environ.get("L", 100))
dtype = np.float32
t = time.time()
X = ad.matrix.dense_device_normal(n, p, dtype=dtype, seed=0)
print("gen", time.time() - t)
rng = np.random.default_rng(0)
beta = np.zeros(p, dtype=dtype)
supp = rng.choice(p, p // 20, replace=False)
beta[supp] = rng.normal(size=supp.size)
t = time.time()
eta = X @ beta
print("eta", time.time() - t)
y = (eta + np.linalg.norm(beta) * rng.normal(size=n)).astype(dtype)
groups = np.arange(0, p, gs)
ad.set_configs("sweep_profile", int(os.environ.get("PROF", 0)))
ad.set_configs("sweep_batch", int(os.environ.get("BATCH", 0)))
ad.set_configs("sweep_xchg", int(os.environ.get("XCHG", 1)))
ad.set_configs("sweep_u_prefetch", int(os.environ.get("UPRE", 1)))
ad.set_configs("sweep_l2_prefetch", int(os.environ.get("L2PRE", 0)))
for rep in range(2):
    t = time.time()
    st = ad.grpnet(X, ad.glm.gaussian(y, dtype=dtype), groups=groups, early_exit=False, lmda_path_size=L, progress_bar=False, newton_tol=float(os.environ.get('NEWTON_TOL', 1e-12)), tol=float(os.environ.get('TOL', 1e-7)))
    wall = time.time() - t
    print(f"rep {rep}: wall {wall:.3f}s solve {st.total_time:.3f}s err='{st.error}' nl={len(st.lmdas)} sweeps={st.n_sweeps} "
          f"updates={st.n_group_updates} kernel_time={st.time_sweep_kernel:.3f}s pin_solves={st.n_pin_solves} "
          f"ncta={st.sweep_ncta} stages={st.sweep_stages} smem={st.sweep_smem_bytes} staged={st.sweep_staged}")
    print("  devs", st.devs[-1], "active", st.active_sizes[-5:], "screen", st.screen_sizes[-5:])
    bytes_per_update = 4.0 * n * gs
    print(f"  sweep kernel: {st.n_group_updates / st.time_sweep_kernel:.0f} group-updates/s, "
          f"{st.n_group_updates * bytes_per_update / st.time_sweep_kernel / 1e9:.0f} GB/s algorithmic")
    print(f"  batch={st.sweep_batch} batched_launches={st.n_batched_launches} panels_built={st.n_panels_built} t_panels={st.t_panels:.3f}")
    if int(os.environ.get("PROF", 0)) == 1 and st.sweep_batch > 1:
        stt = st.sweep_stats.astype(np.float64)
        groups_ = max(stt[6], 1)
        print(f"  light profile (cycles/group, CTA0): CW stall={stt[0]/groups_:.0f} busy={stt[1]/groups_:.0f} | DW stall={stt[8]/groups_:.0f} busy={stt[9]/groups_:.0f} groups={groups_:.0f}")
    elif int(os.environ.get("PROF", 0)) and st.sweep_batch > 1:
        stt = st.sweep_stats.astype(np.float64)
        groups_, batches_ = max(stt[6], 1), max(stt[7], 1)
        cw = ["wait_panel", "wait_gready", "prox", "corr+book", "batch_ovh", "sweep_bdry"]
        dw = ["wait_full", "dot", "bar+publish", "wait_prox", "update", "sweep_wait"]
        print(f"  CW cycles/group (CTA0, cumulative): " + " ".join(f"{nm}={stt[i]/groups_:.0f}" for i, nm in enumerate(cw)) + f" groups={groups_:.0f} batches={batches_:.0f}")
        px = ["pre", "grad_rot", "rootfind", "sums", "rot_back", "newton_its", "unchanged", "apply_corr"]
        print(f"  prox cycles/group: " + " ".join(f"{nm}={stt[16+i]/groups_:.2f}" for i, nm in enumerate(px)))
        print(f"  DW cycles/group (CTA0, cumulative): " + " ".join(f"{nm}={stt[8+i]/groups_:.0f}" for i, nm in enumerate(dw)))
    elif int(os.environ.get("PROF", 0)):
        stt = st.sweep_stats.astype(np.float64)
        names = ["wait_full", "dot", "bar1+store", "poll", "prox", "bar3", "update"]
        items = max(stt[8], 1)
        print("  cycles/item (cumulative over reps): " + " ".join(f"{nm}={stt[i]/items:.0f}" for i, nm in enumerate(names)) + f" newton_it/item={stt[7]/items:.2f} poll_retries/item={stt[9]/items:.1f} items={items:.0f}")
    if int(os.environ.get("PROF", 0)) and rep == 1 and st.sweep_batch <= 1:
        tr = st.sweep_stats.astype(np.float64)[32:32 + 8 * 148].reshape(148, 8)
        t0 = tr[:, 0].min()
        names = ["dot_done", "ll_stored", "poll_done", "bar2_passed", "prox_done", "update_done"]
        for k, nm in enumerate(names):
            v = tr[:, k] - t0
            print(f"    trace {nm:12s}: min {v.min():8.0f} ns  median {np.median(v):8.0f}  max {v.max():8.0f}  (argmax cta {int(v.argmax())})")
    print("  host timers: " + " ".join(f"{k}={getattr(st, 't_' + k):.3f}" for k in ["pin_presync", "pin_launch", "pin_sync", "pin_download", "panels", "panels_presync", "panels_launch", "panels_sync", "screen_records", "cov_device", "run_pin", "invariance", "screen_host", "abs_grad", "update_solutions"]))
    print("  phases: screen %.3f fit %.3f inv %.3f kkt %.3f" % (sum(st.benchmark_screen), sum(st.benchmark_fit_active), sum(st.benchmark_invariance), sum(st.benchmark_kkt)))
