#!/usr/bin/env python
"""bench.py -- CD sweeps/sec and 100-lambda path time of the B200 group-elastic-net solver.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            our arm   (N>1: launched with torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port) on host cores

One "step" = one full 100-lambda path solve (early_exit=False, min_ratio=1e-2) of the workload
  configs[1]: Gaussian group lasso, dense fp32, n=200k p=20k, 2000 groups of 10        (N = 1)
and for N > 1 the same per-GPU row count per rank (weak scaling, n = 200k * N rows, row-sharded).
`value` = CD sweeps / second with X resident in HBM; `e2e` = the same metric through the public API
`adelie_b200.grpnet(X_host, ...)` with the H2D copy of X from pinned host memory and the D2H read of the
solution inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows per GPU, p, group size, dtype)
    "c2": dict(n=200_000, p=20_000, gs=10, dtype="f32",
               desc="Gaussian group lasso, dense fp32 n=200k p=20k, 2000 groups of 10, 100-lambda path (configs[1])"),
    "small": dict(n=20_000, p=2_000, gs=10, dtype="f32",
                  desc="Gaussian group lasso, dense fp32 n=20k p=2k, 200 groups of 10, 100-lambda path (debug size)"),
}
# newton_tol: the reference default 1e-12 is not resolvable in float32 (|phi(h)| has ~1e-7 granularity near the root; the
# reference's own float32 templates then run into newton_max_iters), so the fp32 workload uses 1e-6 on BOTH arms.
PATH_KW = dict(early_exit=False, lmda_path_size=100, min_ratio=1e-2, progress_bar=False, newton_tol=1e-6)


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


def make_problem(ad, wl, rank, world):
    """Synthetic data with the semantics of adelie.data.dense (seed 0): X ~ N(0,1) generated in HBM per row shard
    (Philox, identical matrix for every shard layout), beta* ~ N(0,1) on a random 5% support, y = X beta* + ||beta*|| N(0,1)."""
    n_local, p, gs = wl["n"], wl["p"], wl["gs"]
    dtype = np.float32 if wl["dtype"] == "f32" else np.float64
    n_total = n_local * world
    X = ad.matrix.dense_device_normal(n_local, p, dtype=dtype, seed=0, row_offset=rank * n_local)
    rng = np.random.default_rng(0)
    beta = np.zeros(p, dtype=dtype)
    supp = rng.choice(p, p // 20, replace=False)
    beta[supp] = rng.normal(size=supp.size)
    eta = X @ beta
    noise = np.random.default_rng(1000 + rank).normal(size=n_local)
    y = (eta + np.linalg.norm(beta) * noise).astype(dtype)
    groups = np.arange(0, p, gs)
    return X, y, groups, n_total, dtype


def run_ours(args):
    import adelie_b200 as ad
    from adelie_b200 import _lib
    import ctypes as C

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    L = _lib.load()
    _lib.check(L.ab_set_device(local))
    if world > 1:
        # control plane only (IPC-handle exchange, barriers, max over ranks): the data-path collectives are the library's own
        # NVLink peer-memory kernels (csrc/dist.cuh + the third exchange level of the fused sweep kernel)
        import torch.distributed as dist_
        dist_.init_process_group("gloo")
        dist = dist_
        ad.dist.init(rank=rank, world=world, local_rank=local)
    wl = WORKLOADS[args.workload]
    X, y, groups, n_total, dtype = make_problem(ad, wl, rank, world)
    n_local, p, gs = wl["n"], wl["p"], wl["gs"]
    sz = np.dtype(dtype).itemsize

    def barrier():
        _lib.check(L.ab_device_synchronize())
        if dist:
            dist.barrier()

    def solve_resident():
        return ad.grpnet(X, ad.glm.gaussian(y, dtype=dtype), groups=groups, **PATH_KW)

    # ---------------- resident (kernel-side) arm
    for _ in range(args.warmup):
        st = solve_resident()
        assert st.error == "", st.error
    sampler = ClockSampler(local); sampler.start()
    barrier()
    _lib.check(L.ab_timer_start())
    t0 = time.perf_counter()
    sweeps = updates = cols = launches = 0; tk = 0.0; pins = 0; klaunch = 0
    for _ in range(args.steps):
        st = solve_resident()
        assert st.error == "", st.error
        sweeps += st.n_sweeps; updates += st.n_group_updates; cols += st.n_col_updates; launches += st.n_kernel_launches
        tk += st.time_sweep_kernel; pins += st.n_pin_solves; klaunch += len(st.launch_ms)
    barrier()
    ms = C.c_double(); _lib.check(L.ab_timer_stop(C.byref(ms)))
    t_res = ms.value * 1e-3
    clocks = sampler.stop()
    path_info = dict(n_lmdas=len(st.lmdas), dev_last=float(st.devs[-1]), active_last=int(st.active_sizes[-1]),
                     screen_last=int(st.screen_sizes[-1]), sweeps_per_path=sweeps // args.steps, group_updates_per_path=updates // args.steps,
                     sweep_ctas=st.sweep_ncta, sweep_stages=st.sweep_stages, sweep_staged=st.sweep_staged, sweep_batch=st.sweep_batch)
    if args.dump_launches and rank == 0:
        # per-launch algorithmic bytes / CUDA-event times of the LAST timed path (to line an ncu capture up with its launch)
        rows = [dict(i=i, algo_bytes=float(sz * n_local * (c + 3 * s)), sweeps=int(s), ms=float(m))
                for i, (c, s, m) in enumerate(zip(st.launch_cols, st.launch_sweeps, st.launch_ms))]
        with open(args.dump_launches, "w") as f:
            json.dump(rows, f)

    # ---------------- e2e arm: host buffers, H2D of X from pinned memory + D2H of the solution inside the timed region
    e2e = None
    if not args.no_e2e:
        Xh = X.to_host()                                   # (n_local, p) column-major host copy of this rank's shard
        _lib.check(L.ab_host_register(_lib.ptr(Xh), Xh.nbytes))
        h2d = Xh.nbytes + 3 * y.nbytes + groups.nbytes
        d2h = 0
        def solve_e2e():
            nonlocal d2h
            s = ad.grpnet(ad.matrix.dense(Xh, method="naive"), ad.glm.gaussian(y, dtype=dtype), groups=groups, **PATH_KW)
            B = s.betas; ic = s.intercepts; dv = s.devs
            d2h = B.data.nbytes + B.indices.nbytes + B.indptr.nbytes + ic.nbytes + dv.nbytes + 100 * (p * sz)   # + grad per lambda
            return s
        for _ in range(min(args.warmup, 1)):
            solve_e2e()
        barrier()
        t0 = time.perf_counter()
        sw2 = 0
        for _ in range(args.steps):
            s = solve_e2e(); sw2 += s.n_sweeps
        barrier()
        t_e2e = time.perf_counter() - t0
        _lib.check(L.ab_host_unregister(_lib.ptr(Xh)))
        e2e = dict(t=t_e2e, sweeps=sw2, h2d=h2d, d2h=d2h)

    # ---------------- max over ranks
    # sweeps are collective in the sharded mode (every rank takes part in every sweep): the job's sweep count is rank 0's
    sweeps_all, e2e_sw_all, launches_all = sweeps, (e2e["sweeps"] if e2e else 0), launches
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
    # DRAM traffic: measured dram bytes / algorithmic bytes of ONE captured launch (profiles/r1_traffic.json, from an
    # `ncu --set full` capture lined up with --dump-launches), applied to the average launch of this run
    traffic = args.traffic
    if traffic is None:
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("kernel", "").split(" ")[0] == kname.split(" ")[0] and world == 1:
                traffic = tj["dram_over_algorithmic"] * algo_bytes / max(klaunch, 1)
        except Exception:
            traffic = None
    out = {
        # weak scaling (task rule 5): the unit is one CD sweep over ONE rank's shard (rows_per_gpu x p, the configs[1] problem);
        # a collective sweep over the N-times larger row-sharded matrix is N such units processed concurrently.  N = 1: plain sweeps/s.
        "metric": "cd_sweeps_per_sec", "value": world * sweeps_all / t_res, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": wl["desc"], "rows_per_gpu": n_local, "n_total": n_total, "p": p, "group_size": gs,
                   "path": "100 lambdas, min_ratio=1e-2, early_exit=False, tol=1e-7, newton_tol=1e-6", "l2_policy": "inputs (X = %.1f GB per GPU) larger than L2" % (n_local * p * sz / 1e9),
                   "parallelism": "1 GPU" if world == 1 else f"rows sharded over {world} GPUs (weak scaling: {n_local} rows per GPU), NVLink peer-memory exchange inside the sweep kernel + one-shot all-reduce of the KKT gradient",
                   "unit_definition": "one CD sweep over one rank's shard (rows_per_gpu x p); value = n_gpus x collective sweeps / time, global_sweeps_per_sec = collective sweeps / time",
                   **path_info},
        "path_time_s": t_res / args.steps,
        "global_sweeps_per_sec": sweeps_all / t_res,          # collective sweeps over the whole (n_total x p) matrix per second
        "group_updates_per_sec": updates / t_res,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                     "algorithmic_bytes_per_launch": algo_bytes / max(klaunch, 1), "avg_launch_ms": 1e3 * tk / max(klaunch, 1),
                     "launches_timed": klaunch, "kernel_share_of_step": tk / t_res, "traffic": traffic},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
    }
    if e2e:
        out["e2e"] = {"value": world * e2e_sw_all / e2e["t"], "unit": "sweeps/s", "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                      "path_time_s": e2e["t"] / args.steps}
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(X.to_host() if args.no_e2e else Xh, y, groups, dtype, budget=args.cpu_seconds)
    print(json.dumps(out))


def cpu_baseline(Xh, y, groups, dtype, budget):
    """The oracle port of the reference's CPU algorithm on the host cores: same problem, same path settings, stopped after
    `budget` seconds (the solved lambda prefix is the bounded sample)."""
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    kw = dict(PATH_KW); kw.pop("progress_bar")
    ref = orc.grpnet(Xh, orc.glm_spec("gaussian", y, dtype=dtype), groups=groups, n_threads=cores, max_seconds=budget, **kw)
    t = ref.total_time
    return {"value": ref.n_sweeps / t, "unit": "sweeps/s", "cores": cores, "kind": "port",
            "sample": f"same full-size problem, first {len(ref.lmdas)} of 100 lambdas solved within a {budget:.0f}s budget "
                      f"({int(ref.n_sweeps)} sweeps, {int(ref.n_group_updates)} group updates in {t:.1f}s; early-path sweeps cover "
                      f"fewer groups than the whole-path average, which favours the CPU number)",
            "group_updates_per_sec": ref.n_group_updates / t, "error": ref.error}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; the reference itself needs Eigen and cannot be
    built here) on all host threads, on the same config / metric, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as orc
    wl = WORKLOADS[args.workload]
    n, p, gs = wl["n"], wl["p"], wl["gs"]
    dtype = np.float32 if wl["dtype"] == "f32" else np.float64
    cores = os.cpu_count() or 1
    # same synthetic problem: generate on the device when one is available (identical matrix), else on the host
    try:
        import adelie_b200 as ad
        X, y, groups, _, _ = make_problem(ad, wl, 0, 1)
        Xh = X.to_host(); del X
    except Exception:
        rng = np.random.default_rng(0)
        Xh = np.asfortranarray(rng.standard_normal((n, p), dtype=dtype))
        beta = np.zeros(p, dtype=dtype); supp = rng.choice(p, p // 20, replace=False); beta[supp] = rng.normal(size=supp.size)
        y = (Xh @ beta + np.linalg.norm(beta) * rng.normal(size=n)).astype(dtype)
        groups = np.arange(0, p, gs)
    kw = dict(PATH_KW); kw.pop("progress_bar")
    budget = max(3.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    glm = orc.glm_spec("gaussian", y, dtype=dtype)
    for _ in range(args.warmup):
        orc.grpnet(Xh, glm, groups=groups, n_threads=cores, max_seconds=budget, **kw)
    sweeps = 0; t = 0.0; nl = 0; upd = 0
    for _ in range(args.steps):
        r = orc.grpnet(Xh, glm, groups=groups, n_threads=cores, max_seconds=budget, **kw)
        sweeps += r.n_sweeps; t += r.total_time; nl = len(r.lmdas); upd += r.n_group_updates
    val = sweeps / t
    sample = f"full-size problem, first {nl} of 100 lambdas per step within a {budget:.0f}s budget; {cores} OpenMP threads"
    out = {"impl": "reference", "metric": "cd_sweeps_per_sec", "value": val, "unit": "sweeps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": wl["dtype"], "data": "synthetic", "config": {"workload": wl["desc"], "rows": n, "p": p, "group_size": gs},
           "group_updates_per_sec": upd / t,
           "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--dump-launches", type=str, default=None, help="write per-launch algorithmic bytes / times of the last timed path to this JSON file")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch of the sweep kernel from an ncu capture (profiles/)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
