#!/usr/bin/env python
"""bench.py -- CD sweeps/sec and 100-lambda path time of the B200 group-elastic-net solver.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            our arm   (N>1: launched with torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port) on the host cores;
                                                           never imports adelie_b200 / never maps libadelie_b200.so

Workloads (--workload):
  c2 (default)  configs[1]: Gaussian group lasso, dense fp32, 200k rows PER GPU x p=20k, 2000 groups of 10 (weak scaling:
                the metric's own n=1M x p=50k is 200 GB in fp32 and does not fit one B200, configs[1] is the largest dense
                one-GPU configuration)
  headline      the metric's own config: dense Gaussian n=1M x p=50k fp32, 5000 groups of 10, rows sharded over N >= 2 GPUs
                (strong scaling); N=1 prints a "does not fit" line
  c3            configs[2]: binomial group elastic net alpha=0.5 on the same n=1M x p=50k matrix, IRLS, N >= 2 (strong scaling)
  small         debug size

One "step" = one full 100-lambda path solve (early_exit=False, min_ratio=1e-2).  `value` = CD sweeps / second with X resident
in HBM; `e2e` = the same metric through the public API `adelie_b200.grpnet(X_host, ...)` with the H2D copy of X from pinned
host memory and the D2H read of the solution inside the timed region.  Both arms build the SAME synthetic problem from the
same seeded host generator (`make_problem`), and both also solve the same explicit lambda prefix (`same_work`): the CPU cannot
finish the full path inside the time limit, so the reference arm's step is that prefix, and our arm times it too.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(n=200_000, p=20_000, gs=10, dtype="f32", family="gaussian", alpha=1.0, scaling="weak",
               desc="Gaussian group lasso, dense fp32 n=200k p=20k, 2000 groups of 10, 100-lambda path (configs[1])"),
    "headline": dict(n=1_000_000, p=50_000, gs=10, dtype="f32", family="gaussian", alpha=1.0, scaling="strong",
                     desc="Gaussian group lasso, dense fp32 n=1M p=50k, 5000 groups of 10, 100-lambda path (the metric's own config)"),
    "c3": dict(n=1_000_000, p=50_000, gs=10, dtype="f32", family="binomial", alpha=0.5, scaling="strong",
               desc="Binomial group elastic net alpha=0.5, dense fp32 n=1M p=50k, 5000 groups of 10, IRLS, 100-lambda path (configs[2])"),
    "small": dict(n=20_000, p=2_000, gs=10, dtype="f32", family="gaussian", alpha=1.0, scaling="weak",
                  desc="Gaussian group lasso, dense fp32 n=20k p=2k, 200 groups of 10, 100-lambda path (debug size)"),
}
# newton_tol: the reference default 1e-12 is not resolvable in float32 (|phi(h)| has ~1e-7 granularity near the root; the
# reference's own float32 templates then run into newton_max_iters), so the fp32 workload uses 1e-6 on BOTH arms.
PATH_KW = dict(early_exit=False, lmda_path_size=100, min_ratio=1e-2, progress_bar=False, newton_tol=1e-6)
PREFIX = 20          # lambdas of the same-work prefix (the CPU needs ~5 s for them at configs[1]; the full path takes minutes)


def peak_hbm_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index; self.samples = []; self._stop = threading.Event(); self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=5)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, nm in enumerate(names):
                if len(s) > 2 + k and s[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------------
# The synthetic problem, shared by both arms (NumPy only: the reference arm must not load the product library).
# Semantics of adelie.data.dense (PY/data.py:160-219): X ~ N(0,1) iid column-major, beta* ~ N(0,1) on a random 5 % support,
# gaussian y = X beta* + ||beta*|| N(0,1), binomial y ~ Bernoulli(sigmoid(X beta* / ||beta*||)), weights 1/n, penalty sqrt(gs).
# ------------------------------------------------------------------------------------------------------------------------
def host_normal_matrix(n, p, dtype, seed, shard, threads):
    """(n, p) column-major N(0,1): column block b of row shard `shard` comes from Philox(key=seed, counter=(0,0,shard,b))."""
    X = np.empty((n, p), dtype=dtype, order="F")
    XT = X.T                                              # C-contiguous (p, n) view
    blk = max(1, min(64, p // max(1, 4 * threads)))
    def fill(b):
        j0 = b * blk; j1 = min(p, j0 + blk)
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, shard, b]))
        rng.standard_normal(out=XT[j0:j1], dtype=dtype)
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(fill, range((p + blk - 1) // blk)))
    return X


def truth(wl, dtype):
    p = wl["p"]
    rng = np.random.default_rng(0)
    beta = np.zeros(p, dtype=dtype)
    supp = rng.choice(p, p // 20, replace=False)
    beta[supp] = rng.normal(size=supp.size)
    return beta


def response(wl, eta, beta, shard, dtype):
    rng = np.random.default_rng(1000 + shard)
    nb = float(np.linalg.norm(beta))
    if wl["family"] == "gaussian":
        return (eta + nb * rng.normal(size=eta.shape[0])).astype(dtype)
    mu = 1.0 / (1.0 + np.exp(-eta.astype(np.float64) / nb))
    return (rng.uniform(size=eta.shape[0]) < mu).astype(dtype)


def make_problem_host(wl, n_local, shard, threads):
    dtype = np.float32 if wl["dtype"] == "f32" else np.float64
    Xh = host_normal_matrix(n_local, wl["p"], dtype, 0, shard, threads)
    beta = truth(wl, dtype)
    y = response(wl, Xh @ beta, beta, shard, dtype)
    groups = np.arange(0, wl["p"], wl["gs"])
    return Xh, y, groups, dtype


def lambda_prefix(wl, Xh, y, groups, dtype, n_total):
    """The first PREFIX lambdas of the 100-lambda grid (solver/utils.hpp:6-41): lmda_max = max_g ||X_g^T W (y - ybar)|| / (alpha p_g),
    log-spaced down to min_ratio * lmda_max.  Gaussian only (the same-work leg is defined for the default workload)."""
    w = 1.0 / n_total
    yc = y.astype(np.float64) - float(np.sum(y, dtype=np.float64)) * w
    grad = (Xh.T @ yc.astype(dtype)).astype(np.float64) * w
    gs = wl["gs"]
    score = np.sqrt(np.add.reduceat(grad ** 2, groups)) / (wl["alpha"] * np.sqrt(gs))
    lmax = float(np.max(score))
    L = PATH_KW["lmda_path_size"]
    path = lmax * np.exp(np.log(PATH_KW["min_ratio"]) * np.arange(L) / (L - 1))
    return path[:PREFIX].astype(dtype)


# ------------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C
    import adelie_b200 as ad
    from adelie_b200 import _lib

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    L = _lib.load()
    _lib.check(L.ab_set_device(local))
    if world > 1:
        # control plane only (IPC-handle exchange, barriers, max over ranks): the data-path collectives are the library's own
        # NVLink peer-memory kernels (csrc/dist.cuh + the exchange level of the fused sweep kernels)
        import torch.distributed as dist_
        dist_.init_process_group("gloo")
        dist = dist_
        ad.dist.init(rank=rank, world=world, local_rank=local)
    wl = WORKLOADS[args.workload]
    strong = wl["scaling"] == "strong"
    p, gs = wl["p"], wl["gs"]
    dtype = np.float32 if wl["dtype"] == "f32" else np.float64
    sz = np.dtype(dtype).itemsize
    if strong:
        lo, hi = ad.dist.shard_rows(wl["n"], world, rank)
        n_local, n_total = hi - lo, wl["n"]
        if world == 1 and n_local * p * sz > 150e9:
            if rank == 0:
                print(json.dumps({"metric": "cd_sweeps_per_sec", "value": None, "unit": "sweeps/s", "n_gpus": 1,
                                  "unavailable": "does not fit: X = %.0f GB in %s on one 180 GB B200; run with --gpus >= 2 "
                                                 "(N=1 line of the curve: --workload c2)" % (wl["n"] * p * sz / 1e9, wl["dtype"]),
                                  "config": {"workload": wl["desc"]}}))
            return
    else:
        n_local, n_total = wl["n"], wl["n"] * world
    threads = max(1, (os.cpu_count() or 1) // world)
    host_source = not strong            # strong-scaled 200 GB workloads are generated in HBM (Philox per (column, global row))
    if host_source:
        Xh, y, groups, _ = make_problem_host(wl, n_local, rank, threads)
        _lib.check(L.ab_host_register(_lib.ptr(Xh), Xh.nbytes))
        X = ad.matrix.dense(Xh, method="naive")
    else:
        Xh = None
        X = ad.matrix.dense_device_normal(n_local, p, dtype=dtype, seed=0, row_offset=lo)
        beta = truth(wl, dtype)
        y = response(wl, X @ beta, beta, rank, dtype)
        groups = np.arange(0, p, gs)

    def glm():
        return ad.glm.gaussian(y, dtype=dtype) if wl["family"] == "gaussian" else ad.glm.binomial(y, dtype=dtype)

    kw = dict(PATH_KW); kw["alpha"] = wl["alpha"]

    def barrier():
        _lib.check(L.ab_device_synchronize())
        if dist:
            dist.barrier()

    def solve_resident(**extra):
        k = dict(kw); k.update(extra)
        st = ad.grpnet(X, glm(), groups=groups, **k)
        assert st.error == "", st.error
        return st

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA-event time on the library's stream."""
        acc = dict(sweeps=0, updates=0, cols=0, launches=0, tk=0.0, klaunch=0, last=None)
        barrier()
        _lib.check(L.ab_timer_start())
        t0 = time.perf_counter()
        for _ in range(steps):
            st = fn()
            acc["sweeps"] += st.n_sweeps; acc["updates"] += st.n_group_updates; acc["cols"] += st.n_col_updates
            acc["launches"] += st.n_kernel_launches; acc["tk"] += st.time_sweep_kernel; acc["klaunch"] += len(st.launch_ms)
            if acc["last"] is not None:
                acc["last"].close()
            acc["last"] = st
        barrier()
        ms = C.c_double(); _lib.check(L.ab_timer_stop(C.byref(ms)))
        acc["t"] = ms.value * 1e-3; acc["wall"] = time.perf_counter() - t0
        return acc

    # ---------------- resident (kernel-side) arm
    for _ in range(args.warmup):
        solve_resident().close()
    sampler = ClockSampler(local); sampler.start()
    res = timed(solve_resident, args.steps)
    clocks = sampler.stop()
    st = res["last"]
    t_res, sweeps, updates, cols, launches, tk, klaunch = (res[k] for k in ("t", "sweeps", "updates", "cols", "launches", "tk", "klaunch"))
    path_info = dict(n_lmdas=len(st.lmdas), dev_last=float(st.devs[-1]), active_last=int(st.active_sizes[-1]),
                     screen_last=int(st.screen_sizes[-1]), sweeps_per_path=sweeps // args.steps, group_updates_per_path=updates // args.steps,
                     sweep_ctas=st.sweep_ncta, sweep_stages=st.sweep_stages, sweep_staged=st.sweep_staged, sweep_batch=st.sweep_batch)
    if wl["family"] != "gaussian":
        path_info["irls_iters_per_path"] = int(st.n_irls)
    phases = {k: float(np.sum(getattr(st, "benchmark_" + k))) for k in ("screen", "fit_screen", "fit_active", "kkt", "invariance")}
    if args.dump_launches and rank == 0:
        # per-launch algorithmic bytes / CUDA-event times of the LAST timed path (to line an ncu capture up with its launch)
        rows = [dict(i=i, algo_bytes=float(sz * n_local * (c + 3 * s)), sweeps=int(s), ms=float(m))
                for i, (c, s, m) in enumerate(zip(st.launch_cols, st.launch_sweeps, st.launch_ms))]
        with open(args.dump_launches, "w") as f:
            json.dump(rows, f)
    st.close(); res["last"] = None

    # ---------------- same-work leg: the explicit lambda prefix the reference arm solves (resident and e2e)
    same = None
    if host_source and wl["family"] == "gaussian" and not args.no_same_work:
        lp = lambda_prefix(wl, Xh, y, groups, dtype, n_total) if world == 1 else None
        if lp is not None:
            solve_resident(lmda_path=lp).close()
            r = timed(lambda: solve_resident(lmda_path=lp), min(args.steps, 5))
            same = dict(lmdas=PREFIX, steps=min(args.steps, 5), gpu_sweeps=r["sweeps"] // min(args.steps, 5),
                        gpu_group_updates=r["updates"] // min(args.steps, 5), gpu_col_updates=r["cols"] // min(args.steps, 5),
                        gpu_path_time_s=r["t"] / min(args.steps, 5), gpu_sweeps_per_sec=r["sweeps"] / r["t"],
                        gpu_col_updates_per_sec=r["cols"] / r["t"])
            r["last"].close()

    # ---------------- e2e arm: host buffers, H2D of X from pinned memory + D2H of the solution inside the timed region
    e2e = None
    if host_source and not args.no_e2e:
        X.close()                                            # the resident copy is not needed any more: one X in HBM at a time
        h2d = Xh.nbytes + 3 * y.nbytes + groups.nbytes
        d2h = [0]
        def solve_e2e(**extra):
            k = dict(kw); k.update(extra)
            with ad.matrix.dense(Xh, method="naive") as Xd:
                s = ad.grpnet(Xd, glm(), groups=groups, **k)
                assert s.error == "", s.error
                B = s.betas; ic = s.intercepts; dv = s.devs
                d2h[0] = B.data.nbytes + B.indices.nbytes + B.indptr.nbytes + ic.nbytes + dv.nbytes + len(s.lmdas) * (p * sz)   # + grad per lambda
            return s
        solve_e2e().close()
        gc.collect()
        r = timed(solve_e2e, args.steps)
        e2e = dict(t=r["wall"], sweeps=r["sweeps"], h2d=h2d, d2h=d2h[0])
        r["last"].close()
        if same is not None:
            ks = same["steps"]
            r = timed(lambda: solve_e2e(lmda_path=lp), ks)
            same.update(gpu_e2e_path_time_s=r["wall"] / ks, gpu_e2e_sweeps_per_sec=r["sweeps"] / r["wall"])
            r["last"].close()
    X.close()

    # ---------------- max over ranks
    # sweeps are collective in the sharded mode (every rank takes part in every sweep): the job's sweep count is rank 0's
    e2e_sw_all, launches_all = (e2e["sweeps"] if e2e else 0), launches
    if dist:
        import torch
        t = torch.tensor([t_res, e2e["t"] if e2e else 0.0, tk], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, te, tk = t.tolist()
        if e2e:
            e2e["t"] = te
        cnt = torch.tensor([launches], dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        launches_all = cnt.item()

    if rank != 0:
        return
    peak, peak_src = peak_hbm_gbs()
    algo_bytes = sz * n_local * (cols + 3 * sweeps)           # per GPU: s*n_local*sum(gs) + 3*s*n_local per sweep
    achieved = algo_bytes / tk / 1e9 if tk > 0 else 0.0
    kname = ("pin_solve_batched_kernel (fused look-ahead CD sweep, batches of %d groups), per GPU" % path_info["sweep_batch"]
             if path_info["sweep_batch"] > 1 else "pin_solve_kernel (fused CD sweep), per GPU")
    # DRAM traffic: measured dram bytes / algorithmic bytes of ONE captured launch (profiles/*_traffic.json, from an
    # `ncu --set full` capture lined up with --dump-launches), applied to the average launch of this run
    traffic = args.traffic
    if traffic is None and world == 1 and args.workload == "c2":
        for name in ("r2_traffic.json", "r1_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    tj = json.load(f)
                if tj.get("kernel", "").split(" ")[0] == kname.split(" ")[0]:
                    traffic = tj["dram_over_algorithmic"] * algo_bytes / max(klaunch, 1)
                    break
            except Exception:
                pass
    unit_mult = 1 if strong else world
    out = {
        # weak scaling: the unit is one CD sweep over ONE rank's shard (rows_per_gpu x p, the configs[1] problem); a collective
        # sweep over the N-times larger row-sharded matrix is N such units processed concurrently.  strong scaling: plain sweeps/s.
        "metric": "cd_sweeps_per_sec", "value": unit_mult * sweeps / t_res, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": wl["desc"], "rows_per_gpu": n_local, "n_total": n_total, "p": p, "group_size": gs,
                   "path": "100 lambdas, min_ratio=1e-2, early_exit=False, tol=1e-7, newton_tol=1e-6", "l2_policy": "inputs (X = %.1f GB per GPU) larger than L2" % (n_local * p * sz / 1e9),
                   "parallelism": "1 GPU" if world == 1 else f"rows sharded over {world} GPUs ({wl['scaling']} scaling: {n_local} rows per GPU), NVLink peer-memory exchange inside the sweep kernel + one-shot all-reduce of the KKT gradient",
                   "unit_definition": ("one CD sweep over the whole n_total x p matrix" if strong else
                                       "one CD sweep over one rank's shard (rows_per_gpu x p); value = n_gpus x collective sweeps / time, global_sweeps_per_sec = collective sweeps / time"),
                   "x_source": "host generator shared with --impl reference (Philox, seed 0)" if host_source else "generated in HBM (Philox per (column, global row), seed 0)",
                   **path_info},
        "path_time_s": t_res / args.steps,
        "global_sweeps_per_sec": sweeps / t_res,          # collective sweeps over the whole (n_total x p) matrix per second
        "group_updates_per_sec": updates / t_res,
        "col_updates_per_sec": cols / t_res,
        "phases_s_last_path": phases,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                     "algorithmic_bytes_per_launch": algo_bytes / max(klaunch, 1), "avg_launch_ms": 1e3 * tk / max(klaunch, 1),
                     "launches_timed": klaunch, "kernel_share_of_step": tk / t_res, "traffic": traffic},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
    }
    if e2e:
        out["e2e"] = {"value": unit_mult * e2e_sw_all / e2e["t"], "unit": "sweeps/s", "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                      "path_time_s": e2e["t"] / args.steps}
    else:
        out["e2e"] = None
    if world == 1 and host_source and not args.no_cpu:
        cb = cpu_baseline(wl, Xh, y, groups, dtype, n_total)
        out["cpu_baseline"] = cb
        if same is not None:
            same.update(cpu_sweeps=cb["sweeps"], cpu_path_time_s=cb["path_time_s"], cpu_sweeps_per_sec=cb["value"], cpu_cores=cb["cores"],
                        speedup_resident=cb["path_time_s"] / same["gpu_path_time_s"])
            if "gpu_e2e_path_time_s" in same:
                same["speedup_e2e"] = cb["path_time_s"] / same["gpu_e2e_path_time_s"]
    if same is not None:
        out["same_work"] = same
    print(json.dumps(out))


def cpu_solve_prefix(orc, wl, Xh, y, groups, dtype, n_total, cores, lp):
    kw = dict(PATH_KW); kw.pop("progress_bar"); kw.pop("lmda_path_size"); kw["alpha"] = wl["alpha"]
    glm = orc.glm_spec(wl["family"], y, dtype=dtype)
    return orc.grpnet(Xh, glm, groups=groups, n_threads=cores, lmda_path=lp, **kw)


def cpu_baseline(wl, Xh, y, groups, dtype, n_total):
    """The oracle port of the reference's CPU algorithm on the host cores: same problem, same settings, on the bounded sample
    `first PREFIX lambdas of the path` (the work `--impl reference` does per step)."""
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    lp = lambda_prefix(wl, Xh, y, groups, dtype, n_total)
    reps, t, sweeps, upd = 0, 0.0, 0, 0
    while reps < 2 or (t < 10.0 and reps < 6):
        ref = cpu_solve_prefix(orc, wl, Xh, y, groups, dtype, n_total, cores, lp)
        assert ref.error == "", ref.error
        t += ref.total_time; sweeps += ref.n_sweeps; upd += ref.n_group_updates; reps += 1
    return {"value": sweeps / t, "unit": "sweeps/s", "cores": cores, "kind": "port",
            "sample": f"same full-size problem, first {PREFIX} of the 100 lambdas (explicit lmda_path), {reps} repetitions, {t:.1f}s of CPU work; "
                      f"early-path sweeps cover fewer groups than the whole-path average, which favours the CPU number",
            "sweeps": sweeps // reps, "path_time_s": t / reps, "group_updates_per_sec": upd / t}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; the reference itself needs Eigen and cannot be
    built here) on all host threads, on the same synthetic problem.  Each step = the first PREFIX lambdas of the 100-lambda
    path on the full-size matrix (deterministic work, independent of --steps).  Never imports adelie_b200."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as orc
    wl = WORKLOADS[args.workload]
    if wl["scaling"] == "strong":
        print(json.dumps({"impl": "reference", "unavailable": "X = %.0f GB does not fit the host; the CPU arm is defined for --workload c2" % (wl["n"] * wl["p"] * 4 / 1e9)}))
        return
    cores = os.cpu_count() or 1
    Xh, y, groups, dtype = make_problem_host(wl, wl["n"], 0, cores)
    lp = lambda_prefix(wl, Xh, y, groups, dtype, wl["n"])
    for _ in range(min(args.warmup, 3)):
        cpu_solve_prefix(orc, wl, Xh, y, groups, dtype, wl["n"], cores, lp)
    sweeps = upd = 0; t = 0.0
    for _ in range(args.steps):
        r = cpu_solve_prefix(orc, wl, Xh, y, groups, dtype, wl["n"], cores, lp)
        assert r.error == "", r.error
        sweeps += r.n_sweeps; t += r.total_time; upd += r.n_group_updates
    val = sweeps / t
    sample = (f"full-size problem (identical X, y: shared seeded host generator), first {PREFIX} of the 100 lambdas per step "
              f"(explicit lmda_path; {sweeps // args.steps} sweeps, {upd // args.steps} group updates per step); {cores} OpenMP threads; "
              f"warm-up steps capped at 3")
    out = {"impl": "reference", "metric": "cd_sweeps_per_sec", "value": val, "unit": "sweeps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": wl["dtype"], "data": "synthetic",
           "config": {"workload": wl["desc"], "rows": wl["n"], "p": wl["p"], "group_size": wl["gs"], "lmda_prefix": PREFIX,
                      # N > 1: the repo's arm solves an N-times taller matrix; its unit is one sweep over ONE shard, which is this problem
                      "same_config": args.gpus == 1,
                      "note": "one host, one shard-sized problem; the weak-scaling unit of the GPU arm (one sweep over one 200k-row shard) is the unit measured here"},
           "path_time_s": t / args.steps, "group_updates_per_sec": upd / t,
           "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-same-work", action="store_true")
    ap.add_argument("--dump-launches", type=str, default=None, help="write per-launch algorithmic bytes / times of the last timed path to this JSON file")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch of the sweep kernel from an ncu capture (profiles/)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
