"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: window sharding, step-stat reduction, and the one
exchange step of the path (the summed all-reduce of the flat gradient bucket in data-parallel training)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpntrackseg_b200.sharding import all_reduce_sum_, reduce_step_stats, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 512, 513):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen.extend(range(a, b))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    a, b = shard_range(9, rank, world)
    # rank r "processes" its windows: time grows with rank, counters are per-rank work
    t, (edges, graphs) = reduce_step_stats(10.0 * (rank + 1), [100.0 * (b - a), b - a])
    out.put((rank, a, b, t, edges, graphs))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_and_shards():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, t0, e0, g0), (r1, a1, b1, t1, e1, g1) = res
    assert (a0, b0, a1, b1) == (0, 5, 5, 9)                 # contiguous, balanced, complete
    assert t0 == t1 == 20.0                                  # max over ranks
    assert e0 == e1 == 900.0 and g0 == g1 == 9.0             # summed work


def test_reduce_without_process_group_is_identity():
    assert reduce_step_stats(3.5, [1, 2]) == (3.5, [1.0, 2.0])


def _bucket(rank, n=297_242):
    """A rank's flat fp32 gradient bucket (the core network's 297 k parameters = 1.19 MB), seeded by the rank."""
    return torch.randn(n, generator=torch.Generator().manual_seed(100 + rank))


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = _bucket(rank)
    w = all_reduce_sum_(g)
    out.put((rank, w, g.double().sum().item(), g[:4096].clone().numpy(), g[-7:].clone().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_bucket_all_reduce_equals_the_single_process_sum():
    """Every rank ends up with the SAME bucket = the sum of the ranks' buckets (bit-equal to a local fp32 add for two
    ranks), and gets the world size back for the mean the optimizer kernel takes."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _bucket(0) + _bucket(1)
    for rank, w, total, head, tail in res:
        assert w == 2
        assert total == want.double().sum().item()
        assert (head == want[:4096].numpy()).all() and (tail == want[-7:].numpy()).all()


def test_gradient_all_reduce_without_process_group_is_identity():
    g = _bucket(3, n=1000)
    before = g.clone()
    assert all_reduce_sum_(g) == 1 and torch.equal(g, before)
