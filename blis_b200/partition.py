"""Work partitioning: the reference's integer arithmetic, restated for the
multi-GPU block splits and the process grid.  Must be bit-exact with the
reference; tests/test_partition.py checks it against the
golden vectors generated from the real library.

  thread_range_sub      frame/thread/bli_thread_range.c:38-184
  thread_partition_2x2  frame/thread/bli_thread.c:194-320 (fast heuristic)
  determine_blocksize   frame/base/bli_blksz.c:236-282
"""
from __future__ import annotations


def determine_blocksize(backward: bool, i: int, dim: int, b_alg: int, b_max: int) -> int:
    left = dim - i
    if backward:
        edge = left % b_alg
        return b_alg + edge if b_alg + edge <= b_max else edge
    return left if left <= b_max else b_alg


def thread_range_sub(work_id: int, n_way: int, n: int, bf: int, handle_edge_low: bool = False):
    """[start, end) of partition `work_id` of `n_way`, in units of `bf` with the ragged
    edge on the last (or, if handle_edge_low, the first) partition."""
    if n_way == 1:
        return 0, n
    whole, left = divmod(n, bf)
    lo = hi = whole // n_way
    if not handle_edge_low:
        n_th_lo = whole % n_way
        if n_th_lo:
            lo += 1
        size_lo, size_hi = lo * bf, hi * bf
        hi_start = n_th_lo * size_lo
        if work_id < n_th_lo:
            return work_id * size_lo, (work_id + 1) * size_lo
        s = hi_start + (work_id - n_th_lo) * size_hi
        e = hi_start + (work_id - n_th_lo + 1) * size_hi
        if work_id == n_way - 1:
            e += left
        return s, e
    n_th_hi = whole % n_way
    n_th_lo = n_way - n_th_hi
    if n_th_hi:
        hi += 1
    size_lo, size_hi = lo * bf, hi * bf
    hi_start = n_th_lo * size_lo + left
    if work_id < n_th_lo:
        s, e = work_id * size_lo, (work_id + 1) * size_lo
        if work_id == 0:
            e += left
        else:
            s += left; e += left
        return s, e
    return hi_start + (work_id - n_th_lo) * size_hi, hi_start + (work_id - n_th_lo + 1) * size_hi


def thread_partition_2x2(n_thread: int, work1: int, work2: int):
    """Factor n_thread into nt1 x nt2 with nt1/nt2 ~ work1/work2 (the reference's ic x jc choice)."""
    if n_thread < 4:
        return (n_thread if work1 >= work2 else 1), (n_thread if work1 < work2 else 1)
    tn1 = tn2 = 1
    rem, f = n_thread, 2
    while rem > 1:
        while rem % f:
            f += 1
        rem //= f
        if work1 > work2:
            work1 //= f; tn1 *= f
        else:
            work2 //= f; tn2 *= f
    if work1 > work2:
        if tn2 % 2 == 0 and abs(work1 // 2 - work2 * 2) < work1 - work2:
            tn1 *= 2; tn2 //= 2
    elif work1 < work2:
        if tn1 % 2 == 0 and abs(work2 // 2 - work1 * 2) < work2 - work1:
            tn1 //= 2; tn2 *= 2
    return tn1, tn2
