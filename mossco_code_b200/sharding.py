"""j-slab sharding of the horizontal grid: one rank == one GPU == one tile, no halo.

Columns are independent (no horizontal term in get_rhs/diff3d), so the reference decomposes the
horizontal index space into DE tiles (src/components/fabm_sediment_component.F90:350-387).  In
Fortran order every (n,k) plane of a j-slab is contiguous, so a slab is a plain slice.  The only
exchange on the path is the MAX-reduction of the adaptive-step accept flag
(solver_library.F90:121) plus the NaN flag: two int32 per attempt.

``torch.distributed`` is plumbing here: rendezvous, the broadcast of the NCCL unique id, barriers
and the flag reduction on CPU (gloo) in tests.  On GPUs the reduction runs inside libmsed_b200
through its own NCCL communicator (``SedimentDriver.comm_init``).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def slab_bounds(jnum: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows [j0, j1) of ``rank``: contiguous, balanced to within one row."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    if jnum < world_size:
        raise ValueError(f"jnum={jnum} < world_size={world_size}: a tile needs at least one row")
    j0 = (jnum * rank) // world_size
    j1 = (jnum * (rank + 1)) // world_size
    return j0, j1


def local_slab(arr: np.ndarray, world_size: int, rank: int) -> np.ndarray:
    """Slice axis 1 (j) of a Fortran-ordered (inum, jnum, ...) array for ``rank``."""
    j0, j1 = slab_bounds(arr.shape[1], world_size, rank)
    return np.asfortranarray(arr[:, j0:j1, ...])


def gather_slabs(slabs: Sequence[np.ndarray]) -> np.ndarray:
    """Inverse of ``local_slab`` over all ranks."""
    return np.asfortranarray(np.concatenate(list(slabs), axis=1))


def reduce_flags_max(flags, group=None):
    """MAX all-reduce of the (violation, NaN) flags with torch.distributed; ``flags`` is a torch
    tensor (CPU for gloo, CUDA for nccl) reduced in place."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    return flags


def broadcast_bytes(payload: bytes, src: int = 0, group=None) -> bytes:
    """Ship a small byte string (the NCCL unique id) from ``src`` to every rank."""
    import torch.distributed as dist

    box: List[object] = [payload if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return bytes(box[0])


def init_flag_collective(sed, group=None) -> None:
    """Give a ``SedimentDriver`` its cross-rank accept/NaN flag reduction (NCCL inside the
    library).  No-op for a single rank."""
    import torch.distributed as dist

    from .sediment import nccl_unique_id

    if not (dist.is_available() and dist.is_initialized()):
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return
    uid = nccl_unique_id() if rank == 0 else b""
    uid = broadcast_bytes(uid, 0, group)
    sed.comm_init(uid, world, rank)
