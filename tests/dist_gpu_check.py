"""Multi-GPU parity check (run under torchrun with >= 2 ranks on one box; driven by tests/test_gpu_dist.py):
every rank first solves the FULL problem on its own GPU (single-GPU mode) and rank 0 also solves it with the CPU oracle; then the
ranks solve the row-sharded problem together, once per kernel / exchange setting (batched look-ahead kernel, per-group kernel, one-hop
atomic exchange, two-level flagged-line exchange).  The sharded path must reproduce the single-GPU path AND the oracle on every rank,
bit-identically across ranks.  Prints 'DIST PASS' on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as td

import adelie_b200 as ad


def extra_checks(rank, world):
    """Row-sharded multi-response and SNP paths (config 5's ingredients): every rank solved the full problem before init() ran in
    `_EXTRA_SINGLE`; here the ranks solve the sharded problem together."""
    ok = True
    for name, (full, make_sharded, tol) in _EXTRA_SINGLE.items():
        st = make_sharded()
        assert st.error == "" and full.error == "", (st.error, full.error)
        B, Bs = np.asarray(st.betas.todense()), np.asarray(full.betas.todense())
        rel = np.max(np.abs(B - Bs)) / np.max(np.abs(Bs))
        reli = np.max(np.abs(np.asarray(st.intercepts) - np.asarray(full.intercepts))) / max(1e-300, np.max(np.abs(np.asarray(full.intercepts))))
        blob = [None] * world
        td.all_gather_object(blob, B.tobytes())
        same = all(b == blob[0] for b in blob)
        good = len(st.lmdas) == len(full.lmdas) and rel < tol and reli < tol
        if rank == 0:
            print(f"{name}: rel_beta={rel:.2e} rel_icpt={reli:.2e} identical_on_all_ranks={same} -> {'ok' if good and same else 'FAIL'}", flush=True)
        ok = ok and good and same
    return ok


_EXTRA_SINGLE = {}


def prepare_extra():
    """single-GPU solutions of the extra cases + closures that solve the same problem row-sharded (called after dist.init())"""
    # multigaussian K = 4 on a dense matrix
    rng = np.random.default_rng(3)
    n, p, K = 24_000, 60, 4
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Bt = np.zeros((p, K)); Bt[:6] = rng.standard_normal((6, K))
    Y = np.ascontiguousarray(X @ Bt + 0.3 + rng.standard_normal((n, K)))
    kw = dict(tol=1e-13, early_exit=False, lmda_path_size=10, min_ratio=0.1, progress_bar=False)
    full = ad.grpnet(X, ad.glm.multigaussian(Y), **kw)

    def sharded_multi():
        lo, hi = ad.dist.shard_rows(n)
        return ad.grpnet(np.asfortranarray(X[lo:hi]), ad.glm.multigaussian(np.ascontiguousarray(Y[lo:hi])), **kw)
    _EXTRA_SINGLE["multigaussian K=4 dense f64"] = (full, sharded_multi, 1e-6)

    # snp_unphased: gaussian lasso and multigaussian K = 8 (config 5 layout), every rank keeps its rows of the same calldata
    data = ad.data.snp_unphased(20_000, 80, seed=2, sparsity=0.8)
    cd = data["X"]; ys = data["glm"].y
    imp = np.sum(np.where(cd > 0, cd, 0), axis=0) / np.maximum(np.sum(cd >= 0, axis=0), 1)
    Xs = ad.matrix.snp_unphased_from_calldata(cd, imp)
    kws = dict(tol=1e-13, early_exit=False, lmda_path_size=10, min_ratio=0.1, progress_bar=False)
    full_s = ad.grpnet(Xs, ad.glm.gaussian(ys), **kws)

    def sharded_snp():
        lo, hi = ad.dist.shard_rows(cd.shape[0])
        return ad.grpnet(ad.matrix.snp_unphased_from_calldata(np.asfortranarray(cd[lo:hi]), imp), ad.glm.gaussian(ys[lo:hi]), **kws)
    _EXTRA_SINGLE["snp_unphased gaussian f64"] = (full_s, sharded_snp, 1e-6)

    dk = ad.data.snp_unphased(16_000, 48, seed=5, sparsity=0.8, K=8, glm="multigaussian")
    cdk = dk["X"]; Yk = np.ascontiguousarray(dk["glm"].y, dtype=np.float32)
    impk = np.sum(np.where(cdk > 0, cdk, 0), axis=0) / np.maximum(np.sum(cdk >= 0, axis=0), 1)
    kwk = dict(tol=1e-7, newton_tol=1e-5, early_exit=False, lmda_path_size=10, min_ratio=0.1, progress_bar=False)
    full_k = ad.grpnet(ad.matrix.snp_unphased_from_calldata(cdk, impk, dtype=np.float32), ad.glm.multigaussian(Yk, dtype=np.float32), **kwk)

    def sharded_snp_multi():
        lo, hi = ad.dist.shard_rows(cdk.shape[0])
        return ad.grpnet(ad.matrix.snp_unphased_from_calldata(np.asfortranarray(cdk[lo:hi]), impk, dtype=np.float32),
                         ad.glm.multigaussian(np.ascontiguousarray(Yk[lo:hi]), dtype=np.float32), **kwk)
    _EXTRA_SINGLE["snp_unphased multigaussian K=8 f32"] = (full_k, sharded_snp_multi, 2e-4)


def main():
    td.init_process_group("gloo")
    rank, world = td.get_rank(), td.get_world_size()
    from adelie_b200 import _lib
    _lib.check(_lib.load().ab_set_device(int(os.environ.get("LOCAL_RANK", rank))))
    results = []
    for (n, p, G, glm_name, dtype, kw) in [
        (40_000, 200, 20, "gaussian", np.float32, dict(tol=1e-7, newton_tol=1e-6, equal_groups=True)),    # batched kernel + NVLink level 3
        (16_000, 120, 12, "gaussian", np.float64, dict(tol=1e-12, equal_groups=True)),                    # batched kernel, float64
        (6_000, 64, 64, "gaussian", np.float64, dict(tol=1e-12, equal_groups=True)),                      # lasso, batched
        (40_000, 120, 24, "gaussian", np.float64, dict(tol=1e-12)),
        (3_000, 60, 60, "gaussian", np.float64, dict(tol=1e-12, alpha=0.7)),           # few rows: single-CTA kernels + level 3
        (40_000, 100, 20, "binomial", np.float64, dict(tol=1e-12, irls_tol=1e-10, alpha=0.5)),
        (64_000, 200, 20, "gaussian", np.float32, dict(tol=1e-7, newton_tol=1e-6)),
    ]:
        equal = kw.pop("equal_groups", False)
        data = ad.data.dense(n, p, G, glm=glm_name, seed=11, equal_groups=equal)
        X = np.asfortranarray(data["X"], dtype=dtype); y = data["glm"].y.astype(dtype)
        mk = (lambda yy: ad.glm.gaussian(yy, dtype=dtype)) if glm_name == "gaussian" else (lambda yy: ad.glm.binomial(yy, dtype=dtype))
        common = dict(groups=data["groups"], penalty=data["penalty"].astype(dtype), early_exit=False, lmda_path_size=12, min_ratio=0.1,
                      progress_bar=False, **kw)
        oracle_B = None
        if rank == 0:
            from oracle import oracle as orc
            okw = dict(common); okw.pop("progress_bar")
            ref = orc.grpnet(X, orc.glm_spec(glm_name, y, dtype=dtype), **okw)
            assert ref.error == "", ref.error
            oracle_B = np.asarray(ref.betas.todense())
        results.append((n, p, glm_name, dtype, X, y, mk, common, ad.grpnet(X, mk(y), **common), oracle_B))
    prepare_extra()
    ad.dist.init()
    assert ad.dist.is_active() and ad.dist.world() == world
    ok = True
    # kernel / exchange settings: defaults (batched look-ahead kernel where it applies, one-hop atomic exchange), the two-level
    # flagged-line exchange, and the per-group kernel forced
    for setting in [dict(), dict(sweep_xchg=0), dict(sweep_batch=1), dict(sweep_batch=1, sweep_xchg=0)]:
        for k, v in setting.items():
            ad.set_configs(k, v)
        for (n, p, glm_name, dtype, X, y, mk, common, single, oracle_B) in results:
            lo, hi = ad.dist.shard_rows(n)
            st = ad.grpnet(np.asfortranarray(X[lo:hi]), mk(y[lo:hi]), **common)
            assert st.error == "" and single.error == "", (st.error, single.error)
            B, Bs = np.asarray(st.betas.todense()), np.asarray(single.betas.todense())
            rel = np.max(np.abs(B - Bs)) / np.max(np.abs(Bs))
            reli = np.max(np.abs(st.intercepts - single.intercepts)) / max(1e-300, np.max(np.abs(single.intercepts)))
            tol = 1e-6 if dtype == np.float64 else 1e-4
            good = (len(st.lmdas) == len(single.lmdas)) and rel < tol and reli < tol and np.allclose(st.devs, single.devs, rtol=10 * tol, atol=10 * tol)
            relo = -1.0
            if rank == 0:
                relo = np.max(np.abs(B - oracle_B)) / np.max(np.abs(oracle_B))
                good = good and relo < tol
            # every rank must hold the identical solution
            blob = [None] * world
            td.all_gather_object(blob, B.tobytes())
            same = all(b == blob[0] for b in blob)
            if rank == 0:
                print(f"{setting or 'defaults'} n={n} p={p} {glm_name} {np.dtype(dtype).name}: rows[{lo},{hi}) rel_beta_vs_1gpu={rel:.2e} rel_beta_vs_oracle={relo:.2e} rel_icpt={reli:.2e} identical_on_all_ranks={same} ncta={st.sweep_ncta} batch={st.sweep_batch} batched_launches={st.n_batched_launches} -> {'ok' if good and same else 'FAIL'}", flush=True)
            flag = [None] * world
            td.all_gather_object(flag, bool(good and same))
            ok = ok and all(flag)
        for k in setting:
            ad.set_configs(k, None)
    ok = extra_checks(rank, world) and ok
    if rank == 0:
        print("DIST PASS" if ok else "DIST FAIL", flush=True)
    td.barrier()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
