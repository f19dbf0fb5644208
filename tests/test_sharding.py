"""CPU: x-slab sharding of one volume — plan invariants, a single-process simulation of the halo exchange for
1..8 ranks on the BASELINE shapes, and a real world_size-2 run over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_nnunet_b200 import sharding
from fast_nnunet_b200 import sliding_window as sw

CASES = [((400, 512, 512), (128, 128, 128)), ((160, 160, 160), (128, 128, 128)), ((1200, 512, 512), (160, 96, 96)),
         ((155, 240, 240), (128, 128, 128)), ((48, 40, 56), (32, 32, 32))]


@pytest.mark.parametrize('vol,patch', CASES)
@pytest.mark.parametrize('world', [1, 2, 3, 4, 8])
def test_plan_invariants(vol, patch, world):
    starts = sw.tile_starts(vol, patch, 0.5)
    plan = sharding.plan_shards(starts, patch, vol, world)
    # every tile exactly once, contiguous, balanced to within one tile
    assert plan.tile_ranges[0][0] == 0 and plan.tile_ranges[-1][1] == len(starts)
    sizes = [b - a for a, b in plan.tile_ranges]
    assert all(plan.tile_ranges[i][1] == plan.tile_ranges[i + 1][0] for i in range(world - 1))
    assert max(sizes) - min(sizes) <= 1
    # owned planes: disjoint cover of [0, X)
    assert plan.owned[0][0] == 0 and plan.owned[-1][1] == vol[0]
    assert all(plan.owned[i][1] == plan.owned[i + 1][0] for i in range(world - 1))
    for r in range(world):
        (a, b), (s0, s1), (l0, l1) = plan.owned[r], plan.slabs[r], plan.local[r]
        if s1 > s0:
            assert l0 <= s0 and s1 <= l1
        if b > a:
            assert l0 <= a and b <= l1
    # every plane of every slab is either owned by that rank or covered by exactly one transfer to its owner
    for r in range(world):
        s0, s1 = plan.slabs[r]
        covered = np.zeros(vol[0], dtype=np.int32)
        a, b = plan.owned[r]
        covered[max(a, s0):min(b, s1)] += 1
        for (src, dst, lo, hi) in plan.sends_of(r):
            assert plan.owned[dst][0] <= lo and hi <= plan.owned[dst][1]
            covered[lo:hi] += 1
        assert (covered[s0:s1] == 1).all()


def _tile_value(t, shape):
    g = torch.Generator().manual_seed(1000 + t)
    return torch.randn(shape, generator=g)


def _expected(vol, patch, starts, heads):
    acc = torch.zeros((heads, *vol))
    for t, s in enumerate(starts):
        acc[:, s[0]:s[0] + patch[0], s[1]:s[1] + patch[1], s[2]:s[2] + patch[2]] += _tile_value(t, (heads, *patch))
    return acc


def _local_acc(plan, rank, vol, patch, starts, heads):
    l0, l1 = plan.local[rank]
    acc = torch.zeros((heads, l1 - l0, vol[1], vol[2]))
    lo, hi = plan.tile_ranges[rank]
    for t in range(lo, hi):
        s = starts[t]
        acc[:, s[0] - l0:s[0] - l0 + patch[0], s[1]:s[1] + patch[1], s[2]:s[2] + patch[2]] += _tile_value(t, (heads, *patch))
    return acc


@pytest.mark.parametrize('world', [2, 3, 5, 8])
def test_exchange_simulated(world):
    vol, patch, heads = (70, 24, 20), (16, 16, 16), 2
    starts = sw.tile_starts(vol, patch, 0.5)
    plan = sharding.plan_shards(starts, patch, vol, world)
    accs = [_local_acc(plan, r, vol, patch, starts, heads) for r in range(world)]
    snapshot = [a.clone() for a in accs]
    for (src, dst, lo, hi) in plan.transfers:
        ls, ld = plan.local[src][0], plan.local[dst][0]
        accs[dst][:, lo - ld:hi - ld] += snapshot[src][:, lo - ls:hi - ls]
    want = _expected(vol, patch, starts, heads)
    for r in range(world):
        a, b = plan.owned[r]
        l0 = plan.local[r][0]
        assert torch.allclose(accs[r][:, a - l0:b - l0], want[:, a:b], atol=1e-5), r


def _worker(rank, world, port, vol, patch, heads, result_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        starts = sw.tile_starts(vol, patch, 0.5)
        plan = sharding.plan_shards(starts, patch, vol, world)
        acc = _local_acc(plan, rank, vol, patch, starts, heads)
        nbytes = sharding.exchange_halos(acc, plan, rank, lambda d, s: d.add_(s))
        a, b = plan.owned[rank]
        l0 = plan.local[rank][0]
        want = _expected(vol, patch, starts, heads)[:, a:b]
        ok = torch.allclose(acc[:, a - l0:b - l0], want, atol=1e-5)
        expect_bytes = sum((hi - lo) for (_, d, lo, hi) in plan.recvs_of(rank)) * heads * vol[1] * vol[2] * 4
        with open(os.path.join(result_dir, f'r{rank}'), 'w') as f:
            f.write('ok' if ok and nbytes == expect_bytes else f'bad ok={ok} bytes={nbytes}/{expect_bytes}')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_exchange_gloo(tmp_path, world):
    port = 29600 + os.getpid() % 300 + world
    mp.spawn(_worker, args=(world, port, (70, 24, 20), (16, 16, 16), 2, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f'r{r}').read() == 'ok'


def _worker_subgroup(rank, world, port, vol, patch, heads, result_dir):
    """Ranks 1..world-1 form a sub-group (group rank = global rank - 1): plan ranks are GROUP ranks and must be
    translated to global ranks for the point-to-point operations."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        members = list(range(1, world))
        group = dist.new_group(ranks=members)
        if rank in members:
            grank, gworld = dist.get_rank(group), dist.get_world_size(group)
            assert grank == rank - 1 and sharding.global_rank(group, grank) == rank
            starts = sw.tile_starts(vol, patch, 0.5)
            plan = sharding.plan_shards(starts, patch, vol, gworld)
            acc = _local_acc(plan, grank, vol, patch, starts, heads)
            sharding.exchange_halos(acc, plan, grank, lambda d, s: d.add_(s), group)
            a, b = plan.owned[grank]
            l0 = plan.local[grank][0]
            want = _expected(vol, patch, starts, heads)[:, a:b]
            ok = torch.allclose(acc[:, a - l0:b - l0], want, atol=1e-5)
        else:
            ok = True
        with open(os.path.join(result_dir, f'r{rank}'), 'w') as f:
            f.write('ok' if ok else 'bad')
    finally:
        dist.destroy_process_group()


def test_exchange_gloo_subgroup(tmp_path):
    world = 3
    port = 29900 + os.getpid() % 90
    mp.spawn(_worker_subgroup, args=(world, port, (70, 24, 20), (16, 16, 16), 2, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f'r{r}').read() == 'ok'
