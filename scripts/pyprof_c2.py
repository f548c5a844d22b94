"""Host-side profile of one resident config-2 solve (where does the time outside the kernels go?)."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adelie_b200 as ad
import bench
wl = bench.WORKLOADS["c2"] if "c2" in bench.WORKLOADS else list(bench.WORKLOADS.values())[0]
X, y, groups, n_total, dtype = bench.make_problem(ad, wl, 0, 1)
f = lambda: ad.grpnet(X, ad.glm.gaussian(y, dtype=dtype), groups=groups, **bench.PATH_KW)
st = f()
t = time.time(); st = f(); print("wall", time.time() - t, "solve total_time", st.total_time)
pr = cProfile.Profile(); pr.enable(); st = f(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
