import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad
from oracle import oracle as orc

def run(n, p, G, ctas, direct, dtype=np.float64):
    ad.set_configs("sweep_ctas", ctas); ad.set_configs("sweep_force_direct", direct)
    data = ad.data.dense(n, p, G, seed=3)
    X = np.asfortranarray(data["X"], dtype=dtype); y = data["glm"].y.astype(dtype)
    kw = dict(groups=data["groups"], penalty=data["penalty"].astype(dtype), tol=1e-10, early_exit=False, lmda_path_size=20, min_ratio=0.1)
    t = time.time()
    st = ad.grpnet(X, ad.glm.gaussian(y, dtype=dtype), progress_bar=False, **kw)
    dt = time.time() - t
    ref = orc.grpnet(X, orc.glm_spec("gaussian", y, dtype=dtype), **kw)
    ok = st.error == "" and len(st.lmdas) == len(ref.lmdas)
    err = np.max(np.abs(np.asarray(st.betas.todense()) - np.asarray(ref.betas.todense()))) if ok else -1
    print(f"n={n} p={p} G={G} ctas={ctas} direct={direct}: err='{st.error[:40]}' nl={len(st.lmdas)} maxdiff={err:.2e} time={dt:.2f} ncta={st.sweep_ncta} staged={st.sweep_staged} sweeps={st.n_sweeps}", flush=True)

for ctas, direct in [(1, 0), (1, 1), (2, 1), (2, 0), (8, 1), (8, 0), (64, 1), (64, 0), (148, 1), (148, 0)]:
    run(20000, 100, 20, ctas, direct)
