"""Multi-GPU parity check (run under torchrun with >= 2 ranks on one box):
every rank first solves the FULL problem on its own GPU (single-GPU mode), then the ranks solve the row-sharded problem together;
the sharded path must reproduce the single-GPU path on every rank.  Prints 'DIST PASS' on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as td

import adelie_b200 as ad


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
        results.append((n, p, glm_name, dtype, X, y, mk, common, ad.grpnet(X, mk(y), **common)))
    ad.dist.init()
    assert ad.dist.is_active() and ad.dist.world() == world
    ok = True
    for (n, p, glm_name, dtype, X, y, mk, common, single) in results:
        lo, hi = ad.dist.shard_rows(n)
        st = ad.grpnet(np.asfortranarray(X[lo:hi]), mk(y[lo:hi]), **common)
        assert st.error == "" and single.error == "", (st.error, single.error)
        B, Bs = np.asarray(st.betas.todense()), np.asarray(single.betas.todense())
        rel = np.max(np.abs(B - Bs)) / np.max(np.abs(Bs))
        reli = np.max(np.abs(st.intercepts - single.intercepts)) / max(1e-300, np.max(np.abs(single.intercepts)))
        tol = 1e-6 if dtype == np.float64 else 1e-4
        good = (len(st.lmdas) == len(single.lmdas)) and rel < tol and reli < tol and np.allclose(st.devs, single.devs, rtol=10 * tol, atol=10 * tol)
        # every rank must hold the identical solution
        blob = [None] * world
        td.all_gather_object(blob, B.tobytes())
        same = all(b == blob[0] for b in blob)
        if rank == 0:
            print(f"n={n} p={p} {glm_name} {np.dtype(dtype).name}: rows[{lo},{hi}) rel_beta={rel:.2e} rel_icpt={reli:.2e} identical_on_all_ranks={same} ncta={st.sweep_ncta} batch={st.sweep_batch} batched_launches={st.n_batched_launches} -> {'ok' if good and same else 'FAIL'}", flush=True)
        ok = ok and good and same
    if rank == 0:
        print("DIST PASS" if ok else "DIST FAIL", flush=True)
    td.barrier()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
