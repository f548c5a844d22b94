"""Config 5 (MultiGaussian K=8 on a snp_unphased matrix, n=500k p=100k, fp32): genotypes generated in HBM at 2 bits each
(12.5 GB -- fits ONE B200; the same matrix in fp32 would be 200 GB).  Under torchrun the rows are sharded over the ranks."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad

n_total = int(os.environ.get("N", 500_000)); p = int(os.environ.get("P", 100_000)); K = int(os.environ.get("K", 8))
L = int(os.environ.get("L", 100)); dtype = np.float32
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
if world > 1:
    import torch.distributed as td
    td.init_process_group("gloo")
    ad.dist.init()
lo, hi = ad.dist.shard_rows(n_total) if world > 1 else (0, n_total)
n = hi - lo
t = time.time()
X = ad.matrix.snp_unphased_device_random(n, p, dtype=dtype, seed=0, row_offset=lo, n_total=n_total)
_, packed = X.cache_info()
if rank == 0:
    print(f"gen {time.time() - t:.2f}s  rows[{lo},{hi}) p={p}  packed={packed / 1e9:.2f} GB (fp32 dense would be {4.0 * n * p / 1e9:.0f} GB)", flush=True)
rng = np.random.default_rng(0)
supp = rng.choice(p, 64, replace=False); B = rng.normal(size=(64, K))
Y = np.zeros((n, K), dtype=dtype)
col = np.zeros(n, dtype=dtype)
for j, b in zip(supp, B):
    col[:] = 0
    X.btmul(int(j), 1, np.array([1.0], dtype=dtype), col)
    Y += col[:, None] * b[None, :].astype(dtype)
noise = np.random.default_rng(1000 + rank).normal(size=(n, K)) * np.linalg.norm(B) / np.sqrt(K) * 0.5
Y = np.ascontiguousarray(Y + noise, dtype=dtype)
if rank == 0:
    print(f"y ready {time.time() - t:.1f}s", flush=True)
for kv in os.environ.get("CFG", "").split(","):          # e.g. CFG=snp_tc=0
    if "=" in kv:
        k_, v_ = kv.split("="); ad.set_configs(k_, float(v_))
for rep in range(int(os.environ.get("REPS", 2))):
    t = time.time()
    st = ad.grpnet(X, ad.glm.multigaussian(Y, dtype=dtype), early_exit=False, lmda_path_size=L, min_ratio=float(os.environ.get("MINR", 1e-2)),
                   tol=1e-7, newton_tol=1e-6, progress_bar=False)
    wall = time.time() - t
    if rank == 0:
        print(f"rep {rep}: wall {wall:.3f}s solve {st.total_time:.3f}s err='{st.error}' nl={len(st.lmdas)} sweeps={st.n_sweeps} updates={st.n_group_updates} "
              f"kernel_time={st.time_sweep_kernel:.3f}s active_last={st.active_sizes[-1] if len(st.active_sizes) else 0} "
              f"screen_last={st.screen_sizes[-1] if len(st.screen_sizes) else 0} dev_last={st.devs[-1]:.4f} cached_cols={X.cache_info()[0]}", flush=True)
        print("  host timers: " + " ".join(f"{k}={getattr(st, 't_' + k):.3f}" for k in ["run_pin", "invariance", "screen_records", "cov_device", "screen_host"]), flush=True)
