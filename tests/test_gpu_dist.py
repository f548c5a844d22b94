"""GPU, >= 2 devices: row-sharded parity.  Launches tests/dist_gpu_check.py under torch.distributed.run with 2 ranks:
sharded == single GPU == CPU oracle for the batched and the per-group sweep kernels with both exchange protocols, plus
multi-response and snp_unphased layouts.  Skipped on a one-GPU box (the log of a 2-GPU run is committed under profiles/)."""
import ctypes
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    from adelie_b200 import _lib
    c = ctypes.c_int(0)
    return c.value if _lib.load().ab_device_count(ctypes.byref(c)) != 0 else c.value


def test_row_sharded_paths_match_single_gpu_and_oracle():
    if _device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(ROOT, "tests", "dist_gpu_check.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    out = r.stdout + r.stderr
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dist_gpu_check.log"), "w") as f:
        f.write(out)
    assert r.returncode == 0 and "DIST PASS" in out, out[-4000:]
