"""Sharding of ONE volume across the GPUs of a box: slabs of sliding-window tiles along the first
spatial axis, one halo exchange of accumulator planes, then per-rank normalise/argmax.

The reference never splits a volume (its multi-GPU inference is file-level: `-num_parts/-part_id`,
predict_from_raw_data.py:918-925, :177); tiles are independent (InstanceNorm is per tile) and the
aggregation (:613) is a commutative sum, so the tile list is partitioned in the reference's own
order (:532-537, first axis outermost) and each rank accumulates only the planes its tiles touch.

Pure host logic here (numpy); the exchange itself uses torch.distributed point-to-point operations
(NCCL over NVLink on the box, gloo in the CPU tests) and libfnnu's add kernel.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class ShardPlan:
    world: int
    tile_ranges: List[Tuple[int, int]]      # [lo, hi) into the flat tile list, per rank
    slabs: List[Tuple[int, int]]            # planes [x_lo, x_hi) each rank's tiles touch (empty: (0, 0))
    owned: List[Tuple[int, int]]            # disjoint cover of [0, X): planes each rank normalises
    local: List[Tuple[int, int]]            # planes each rank allocates = hull(slab, owned)
    transfers: List[Tuple[int, int, int, int]]   # (src, dst, x_lo, x_hi): src's partial sums dst must add

    def sends_of(self, rank):
        return [t for t in self.transfers if t[0] == rank]

    def recvs_of(self, rank):
        return [t for t in self.transfers if t[1] == rank]


def plan_shards(starts: np.ndarray, patch: Sequence[int], vol: Sequence[int], world: int) -> ShardPlan:
    """Contiguous, balanced runs of the flat tile list (ceil/floor split) -> slabs along axis 0.
    Ownership boundaries are the midpoints between consecutive slabs' centres of mass of work, snapped
    so that every rank owns at least one plane when X >= world."""
    n = len(starts)
    X = int(vol[0])
    world = int(world)
    assert world >= 1 and n >= 1
    base, rem = divmod(n, world)
    ranges, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        ranges.append((lo, hi))
        lo = hi
    slabs = []
    for lo, hi in ranges:
        if hi > lo:
            xs = starts[lo:hi, 0]
            slabs.append((int(xs.min()), int(xs.max()) + int(patch[0])))
        else:
            slabs.append((0, 0))
    # ownership: proportional split of [0, X) by tile count (each rank owns planes near its own tiles)
    cuts = [0]
    for r in range(1, world):
        prev_lo, prev_hi = ranges[r - 1]
        cur_lo, cur_hi = ranges[r]
        if cur_hi > cur_lo and prev_hi > prev_lo:
            # boundary between the mean tile centre of rank r-1 and rank r
            c_prev = float(starts[prev_lo:prev_hi, 0].mean()) + patch[0] / 2
            c_cur = float(starts[cur_lo:cur_hi, 0].mean()) + patch[0] / 2
            cut = int(round((c_prev + c_cur) / 2))
        else:
            cut = cuts[-1]
        cut = max(cut, cuts[-1])
        cuts.append(min(cut, X))
    cuts.append(X)
    owned = [(cuts[r], cuts[r + 1]) for r in range(world)]
    local = []
    for r in range(world):
        s, o = slabs[r], owned[r]
        if s[1] > s[0] and o[1] > o[0]:
            local.append((min(s[0], o[0]), max(s[1], o[1])))
        elif s[1] > s[0]:
            local.append(s)
        else:
            local.append(o)
    transfers = []
    for src in range(world):
        s = slabs[src]
        if s[1] <= s[0]:
            continue
        for dst in range(world):
            if dst == src:
                continue
            o = owned[dst]
            lo_, hi_ = max(s[0], o[0]), min(s[1], o[1])
            if hi_ > lo_:
                transfers.append((src, dst, lo_, hi_))
    return ShardPlan(world, ranges, slabs, owned, local, transfers)


def global_rank(group, group_rank: int) -> int:
    """plan.transfers / gather_to hold ranks of `group`; torch.distributed's P2POp takes GLOBAL ranks."""
    import torch.distributed as dist
    if group is None or not dist.is_initialized():
        return int(group_rank)
    return int(dist.get_global_rank(group, int(group_rank)))


def exchange_halos(acc, plan: ShardPlan, rank: int, add_fn: Callable, group=None):
    """acc: this rank's accumulator [H, local planes, Y, Z] (fp32).  Sends the planes other ranks own,
    receives the partial sums for the planes this rank owns and adds them with `add_fn(dst, src)`.
    One batch of point-to-point operations (ncclGroup on NCCL)."""
    import torch
    import torch.distributed as dist
    if plan.world == 1:
        return 0
    l0 = plan.local[rank][0]
    H = acc.shape[0]
    ops, recv_bufs = [], []
    for (_, dst, lo, hi) in plan.sends_of(rank):
        for h in range(H):
            ops.append(dist.P2POp(dist.isend, acc[h, lo - l0:hi - l0], global_rank(group, dst), group=group))
    for (src, _, lo, hi) in plan.recvs_of(rank):
        for h in range(H):
            buf = torch.empty((hi - lo, *acc.shape[2:]), dtype=acc.dtype, device=acc.device)
            recv_bufs.append((h, lo, hi, buf))
            ops.append(dist.P2POp(dist.irecv, buf, global_rank(group, src), group=group))
    nbytes = 0
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for h, lo, hi, buf in recv_bufs:
        add_fn(acc[h, lo - l0:hi - l0], buf)
        nbytes += buf.numel() * buf.element_size()
    return nbytes
