"""Contiguous 1-D block partition of a global vector over ranks (one rank per
GPU) -- the MPIPlusX layout (src/nvector/mpiplusx/nvector_mpiplusx.c:30 wraps a
local vector; the application owns the split).  Host-side logic only."""
from __future__ import annotations


def block_range(global_length: int, rank: int, size: int) -> tuple[int, int]:
    """[start, stop) of `rank`'s block: the first (global_length % size) ranks get
    one extra element, so blocks differ by at most one and tile the vector."""
    if size < 1 or not (0 <= rank < size) or global_length < 0:
        raise ValueError("bad partition arguments")
    base, rem = divmod(global_length, size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def local_length(global_length: int, rank: int, size: int) -> int:
    a, b = block_range(global_length, rank, size)
    return b - a


# how each reducing op combines rank-local results
# (src/nvector/manyvector/nvector_manyvector.c:815-1793)
COMBINE = {
    "dot_prod": "sum", "l1_norm": "sum", "wsqr_sum": "sum", "wsqr_sum_mask": "sum", "dot_prod_multi": "sum",
    "wsqr_sum_vector_array": "sum", "max_norm": "max", "min": "min", "min_quotient": "min",
    "inv_test": "min", "constr_mask": "min",
}
