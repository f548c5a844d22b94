"""Row-sharded multi-GPU mode: one process per GPU of one NVSwitch box.

Every rank holds a contiguous block of rows of X, y, weights and offsets; all O(p)/O(G) state is replicated and every rank
runs the identical control flow.  Cross-GPU sums go over NVLink peer memory inside the library (csrc/dist.cuh, and the
third exchange level of the fused sweep kernel); this module only bootstraps the peer mapping and offers the few host-side
reductions the Python initialisation of ``grpnet`` needs.

    torchrun --nproc-per-node 8 train.py
        import torch.distributed as td, adelie_b200 as ad
        td.init_process_group("gloo")          # any backend: only used to all-gather 64-byte IPC handles
        ad.dist.init()                          # RANK / WORLD_SIZE / LOCAL_RANK from the environment
        lo, hi = ad.dist.shard_rows(n_total)    # this rank's rows
        state = ad.grpnet(X[lo:hi], ad.glm.gaussian(y[lo:hi]), ...)      # identical result on every rank
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

_state = {"rank": 0, "world": 1, "active": False}


def shard_rows(n_total: int, world: int = None, rank: int = None):
    """[lo, hi) of the rows owned by ``rank``: contiguous blocks whose sizes differ by at most one 32-row unit."""
    world = _state["world"] if world is None else world
    rank = _state["rank"] if rank is None else rank
    units = (n_total + 31) // 32
    base, rem = divmod(units, world)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * 32, n_total), min(hi_u * 32, n_total)


def init(rank: int = None, world: int = None, local_rank: int = None, gather=None):
    """Maps every rank's peer-visible slab into this process.  ``gather(bytes) -> list[bytes]`` all-gathers one 64-byte blob
    per rank in rank order; by default ``torch.distributed.all_gather_object`` of the already initialised process group."""
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
    local_rank = int(os.environ.get("LOCAL_RANK", rank)) if local_rank is None else local_rank
    L = _lib.load()
    _lib.check(L.ab_set_device(local_rank))
    if world <= 1:
        _state.update(rank=0, world=1, active=False)
        return
    handle = C.create_string_buffer(64)
    _lib.check(L.ab_dist_init(rank, world, handle))
    if gather is None:
        import torch.distributed as td
        out = [None] * world
        td.all_gather_object(out, handle.raw)
        blobs = out
    else:
        blobs = gather(handle.raw)
    allh = b"".join(blobs)
    assert len(allh) == 64 * world
    _lib.check(L.ab_dist_connect(C.c_char_p(allh)))
    _state.update(rank=rank, world=world, active=True)
    if gather is None:
        import torch.distributed as td
        td.barrier()


def is_active() -> bool:
    return _state["active"]


def rank() -> int:
    return _state["rank"]


def world() -> int:
    return _state["world"]


def allreduce(x):
    """Sum over ranks of a scalar or array (returns float64 array / float); identity when not distributed."""
    if not _state["active"]:
        return x
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64))).copy()
    _lib.check(_lib.load().ab_dist_allreduce_f64(_lib.ptr(a), a.size))
    if np.ndim(x) == 0:
        return float(a[0])
    return a.reshape(np.shape(x))
