"""Host logic of the slab decomposition on CPU: cut planes, ownership, and the
neighbour exchange over a world-size-2 gloo group (no GPU needed)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taichi_elements_b200.distributed import SlabDecomposition, neighbour_exchange

LEAF, GS, INV = 4, 4096, 256.0


def test_uniform_and_balanced_cuts():
    cuts = SlabDecomposition.uniform_cuts(-1.0, 1.0, 8, LEAF, GS, INV)
    assert len(cuts) == 7 and all(b > a for a, b in zip(cuts, cuts[1:]))
    assert cuts[3] == GS // 2 // LEAF                       # the middle cut is x = 0
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(30000) * 0.2, 0.5 + rng.random(10000) * 0.4]).astype(np.float32)
    cuts = SlabDecomposition.balanced_cuts(x, 4, LEAF, GS, INV)
    owners = SlabDecomposition(4, 0, cuts, LEAF, GS, INV).owner(x)
    counts = np.bincount(owners, minlength=4)
    assert counts.sum() == len(x) and counts.min() > 0.15 * len(x)


def test_ownership_is_a_partition_on_block_boundaries():
    rng = np.random.default_rng(1)
    x = (rng.random(20000) * 2 - 1).astype(np.float32)
    cuts = SlabDecomposition.uniform_cuts(-1.0, 1.0, 4, LEAF, GS, INV)
    slabs = [SlabDecomposition(4, r, cuts, LEAF, GS, INV) for r in range(4)]
    mine = np.stack([s.mine(x) for s in slabs])
    assert np.all(mine.sum(axis=0) == 1)                    # every particle has exactly one owner
    bx = slabs[0].block_x(x)
    for r, s in enumerate(slabs):
        assert np.all((bx[mine[r]] >= s.lo) & (bx[mine[r]] < s.hi))
    assert slabs[0].left is None and slabs[3].right is None and slabs[1].left == 0 and slabs[1].right == 2
    # the owner is decided by the BASE block: floor(x*inv_dx - 0.5), not floor(x*inv_dx)
    edge = np.float32((cuts[1] * LEAF - GS // 2 + 0.25) / INV)   # a quarter cell right of the cut plane
    assert slabs[0].owner(np.array([edge]))[0] == 1 and slabs[0].block_x(np.array([edge]))[0] == cuts[1] - 1


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cuts = SlabDecomposition.uniform_cuts(0.0, 1.0, world, LEAF, GS, INV)
    s = SlabDecomposition(world, rank, cuts, LEAF, GS, INV)
    mk = lambda v: torch.full((64, ), v, dtype=torch.int32)
    send_lo, send_hi = mk(100 * rank + 1), mk(100 * rank + 2)
    recv_lo, recv_hi = mk(-1), mk(-1)
    for _ in range(3):                                    # repeated rounds must not deadlock
        neighbour_exchange(send_lo if s.left is not None else None, send_hi if s.right is not None else None,
                           recv_lo if s.left is not None else None, recv_hi if s.right is not None else None,
                           s.left, s.right)
    ok = True
    if s.left is not None:
        ok &= bool((recv_lo == 100 * s.left + 2).all())   # the left rank's +x buffer
    else:
        ok &= bool((recv_lo == -1).all())
    if s.right is not None:
        ok &= bool((recv_hi == 100 * s.right + 1).all())  # the right rank's -x buffer
    else:
        ok &= bool((recv_hi == -1).all())
    # agreement on a global box, as the solver does before each batch
    lo = torch.tensor([10 * rank, -rank, 5], dtype=torch.int64)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    ok &= lo.tolist() == [0, -(world - 1), 5]
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_neighbour_exchange_gloo(world):
    port = 29500 + os.getpid() % 2000 + world
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world)) and len(out) == world


def test_cost_balanced_cuts():
    """DistributedMPMSolver.balance_cuts: the pure part.  Equal costs keep equal-width slabs; a rank that is twice as
    expensive per column gives columns away; empty ranks are skipped; the cuts stay strictly increasing."""
    from taichi_elements_b200.distributed import SlabDecomposition as S
    # two ranks, same cost, same width: the cut stays in the middle
    assert S.cost_balanced_cuts([(100, 120, 5.0), (120, 140, 5.0)], 2, [0]) == [120]
    # rank 0 costs three times as much over the same width: half of the total cost is reached after 2/3 of its columns
    assert S.cost_balanced_cuts([(100, 130, 3.0), (130, 160, 1.0)], 2, [0]) == [120]
    # four ranks, everything on rank 1: it is split four ways
    assert S.cost_balanced_cuts([(0, 0, 0.0), (200, 240, 8.0), (0, 0, 0.0), (0, 0, 0.0)], 4, [1, 2, 3]) == [210, 220, 230]
    # nothing measured: the old cuts stay
    assert S.cost_balanced_cuts([(0, 0, 0.0), (0, 0, 0.0)], 2, [77]) == [77]
    # a one-column scene still yields strictly increasing cuts
    c = S.cost_balanced_cuts([(50, 51, 1.0), (0, 0, 0.0), (0, 0, 0.0)], 3, [1, 2])
    assert len(c) == 2 and c[0] < c[1]
    # shares: cost left of every cut is k / world of the total (to the rounding of a column)
    segs = [(0, 40, 10.0), (40, 100, 30.0), (100, 110, 20.0)]
    cuts = S.cost_balanced_cuts(segs, 3, [0, 1])

    def cost_left(x):
        return sum(c * min(max((x - a) / (b - a), 0.0), 1.0) for a, b, c in segs)
    for k, cut in enumerate(cuts, 1):
        assert abs(cost_left(cut) - 60.0 * k / 3) <= 2.0
