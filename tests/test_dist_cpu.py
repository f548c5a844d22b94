"""CPU, world_size 2 over gloo: host-side logic of the row-sharded mode -- the shard partition, the all-reduce plumbing of
the Python initialisation (adelie_b200.solver._init_gaussian) and the fact that sharded sums reproduce the unsharded
invariants.  The device collectives themselves are covered by tests/dist_gpu_check.py on >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

import adelie_b200 as ad

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 200_000, 1_000_003])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_rows_partition(n, world):
    parts = [ad.dist.shard_rows(n, world, r) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (a, b), (c, d) in zip(parts[:-1], parts[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in parts]
    assert all(a % 32 == 0 for a, b in parts if b > a)        # every non-empty shard starts on a 32-row unit (TMA alignment)
    nonempty = [s for s in sizes if s > 0]
    assert max(nonempty) - min(nonempty) <= 32 + 31 or n < 32 * world


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch, torch.distributed as td
import adelie_b200 as ad
from adelie_b200 import dist as D, solver as S

td.init_process_group("gloo")
rank, world = td.get_rank(), td.get_world_size()

class NumpyMatrix:                       # stand-in with the operator subset the initialisation uses
    def __init__(self, X): self.X = X; self.dtype = X.dtype.type
    def rows(self): return self.X.shape[0]
    def cols(self): return self.X.shape[1]
    def mul(self, v, w, out): out[...] = self.X.T @ (v * w)

def gloo_allreduce(x):                   # what ab_dist_allreduce_f64 does on the device, here over gloo
    t = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64))).copy())
    td.all_reduce(t)
    return float(t[0]) if np.ndim(x) == 0 else t.numpy().reshape(np.shape(x))

rng = np.random.RandomState(0)
n, p = 1000, 37
X = np.asfortranarray(rng.normal(size=(n, p))); y = rng.normal(size=n); w = rng.uniform(0.5, 1.5, n); w /= w.sum()
off = rng.normal(size=n) * 0.1
full = S._init_gaussian(NumpyMatrix(X), y, w, off, True, np.float64)          # unsharded (allreduce is the identity)
D._state.update(rank=rank, world=world, active=True)
D.allreduce = gloo_allreduce
lo, hi = D.shard_rows(n)
part = S._init_gaussian(NumpyMatrix(X[lo:hi]), y[lo:hi], w[lo:hi], off[lo:hi], True, np.float64)
for k in ("X_means", "grad"):
    assert np.allclose(part[k], full[k], atol=1e-12), k
for k in ("y_mean", "y_var", "resid_sum"):
    assert abs(part[k] - full[k]) < 1e-12, k
assert np.allclose(part["resid"], full["resid"][lo:hi], atol=1e-12)
# default GLM weights are 1 / n_total on every shard
g = ad.glm.gaussian(y[lo:hi]) if False else None
td.barrier()
if rank == 0: print("GLOO_OK")
'''


def test_sharded_initial_invariants_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout
