"""
Replica sharding across ranks (one process per GPU).  Replicas share nothing: rank ``g`` of
``G`` owns the contiguous id range returned by :func:`shard`; the Philox key carries the *global* replica id, so a
replica's trajectory does not depend on how the set is split.  The only exchange is the final telemetry sum
(``itsb_allreduce_telemetry`` over NCCL on GPUs).
"""
from typing import Tuple


def shard(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first replica id, count) of ``rank``: sizes differ by at most one, lower ranks take the remainder."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count
