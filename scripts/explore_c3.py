"""Config 3 (binomial group elastic net alpha=0.5, dense fp32 n=1M p=50k, groups of 10, IRLS outer loop, rows sharded over the GPUs of
one box).  Under torchrun every rank generates its row shard in HBM (N = TOTAL rows); with one process the same script runs a
single shard (N rows) on one GPU.  FAMILY=gaussian runs the Gaussian path on the same matrix (the config BASELINE's metric names)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad

n_total = int(os.environ.get("N", 1_000_000)); p = int(os.environ.get("P", 50_000)); gs = int(os.environ.get("GS", 10))
L = int(os.environ.get("L", 100)); family = os.environ.get("FAMILY", "binomial"); alpha = float(os.environ.get("ALPHA", 0.5))
dtype = np.float32
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
if world > 1:
    import torch.distributed as td
    td.init_process_group("gloo")
    ad.dist.init()
lo, hi = ad.dist.shard_rows(n_total) if world > 1 else (0, n_total)
n = hi - lo
t = time.time()
X = ad.matrix.dense_device_normal(n, p, dtype=dtype, seed=0, row_offset=lo)
rng = np.random.default_rng(0)
beta = np.zeros(p, dtype=dtype); supp = rng.choice(p, p // 20, replace=False); beta[supp] = rng.normal(size=supp.size)
eta = X @ beta
if family == "binomial":
    mu = 1 / (1 + np.exp(-eta / np.linalg.norm(beta)))
    y = (np.random.default_rng(1000 + rank).uniform(size=n) < mu).astype(dtype)
    glm = ad.glm.binomial(y, dtype=dtype)
else:
    y = (eta + np.linalg.norm(beta) * np.random.default_rng(1000 + rank).normal(size=n)).astype(dtype)
    glm = ad.glm.gaussian(y, dtype=dtype)
if rank == 0:
    print(f"gen {time.time() - t:.1f}s  {family} alpha={alpha} rows[{lo},{hi}) of {n_total} p={p} groups of {gs}  X shard = {4.0 * n * p / 1e9:.1f} GB", flush=True)
for kv in os.environ.get("CFG", "").split(","):          # e.g. CFG=glm_batched=0,panel_tc=0
    if "=" in kv:
        k, v = kv.split("="); ad.set_configs(k, float(v))
for rep in range(int(os.environ.get("REPS", 2))):
    t = time.time()
    st = ad.grpnet(X, glm, groups=np.arange(0, p, gs), alpha=alpha, early_exit=False, lmda_path_size=L, min_ratio=float(os.environ.get("MINR", 1e-2)),
                   tol=1e-7, newton_tol=1e-6, progress_bar=False)
    wall = time.time() - t
    if rank == 0:
        print(f"rep {rep}: wall {wall:.3f}s solve {st.total_time:.3f}s err='{st.error}' nl={len(st.lmdas)} sweeps={st.n_sweeps} updates={st.n_group_updates} "
              f"irls={getattr(st, 'n_irls', 0)} kernel_time={st.time_sweep_kernel:.3f}s sweeps/s={st.n_sweeps / st.total_time:.1f} "
              f"active_last={st.active_sizes[-1] if len(st.active_sizes) else 0} screen_last={st.screen_sizes[-1] if len(st.screen_sizes) else 0} dev_last={st.devs[-1]:.4f}", flush=True)
        print("  host timers: " + " ".join(f"{k}={getattr(st, 't_' + k):.3f}" for k in ["run_pin", "invariance", "screen_records", "cov_device", "eigh_device", "panels", "pin_launch", "pin_sync", "pin_download", "screen_host", "rec_phase1", "rec_phase3", "rec_upload", "glm_means"])
              + f" batched_launches={st.n_batched_launches} batch={st.sweep_batch} ctas={st.sweep_ncta} stages={st.sweep_stages}", flush=True)
