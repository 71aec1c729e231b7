"""Piece (3): slices of one network partitioned over the ranks of a
``torch.distributed`` group (one process per GPU), per-rank on-device
accumulation, ONE sum over NCCL/NVLink.

The reference has no distributed code; cotengra executes slices serially and
adds them up (``gather_slices``, reached from optyx/core/backends.py:539-545).
The payload of the collective is the output tensor only (16 B for an
amplitude), so it is latency-bound: a plain ``all_reduce`` is the right tool.
"""
from __future__ import annotations

from typing import Tuple


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when no process group is initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def partition(nslices: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of slice ids of ``rank`` (all slices cost the same)."""
    base, extra = divmod(nslices, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_sum_(tensor):
    """In-place sum of a complex tensor over all ranks (2 x real payload)."""
    import torch
    import torch.distributed as dist
    rank, size = world()
    if size == 1:
        return tensor
    view = torch.view_as_real(tensor) if tensor.is_complex() else tensor
    dist.all_reduce(view, op=dist.ReduceOp.SUM)
    return tensor


def run_sliced(sess, net, upload: bool = True):
    """Build leaves on every rank (cheap, deterministic), contract this rank's
    slices, all-reduce the output.  With an unsliced plan every rank computes
    the whole (replica) and no collective runs.  ``upload=False``: the caller has
    already uploaded THIS net's values with ``sess.set_leaves`` (timing loops)."""
    rank, size = world()
    plan = sess.plan
    # by default always upload this net's host-computed values: a reused Session may
    # have seen a structurally identical net with other parameters (the descriptor
    # table is cached)
    if upload or sess.desc_dev is None:
        sess.set_leaves([net])
    sess.build_leaves()
    if plan.nslices <= 1 or size == 1:
        sess.contract(net.scalar)
        return sess.out[0]
    sess.contract(net.scalar, slice_range=partition(plan.nslices, rank, size))
    allreduce_sum_(sess.out)
    return sess.out[0]
