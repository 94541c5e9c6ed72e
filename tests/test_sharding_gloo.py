"""N>1 path on CPU: two ranks over gloo shard the files with no data-path collective and
combine only the timing (max over ranks) and the byte counts, exactly as bench.py does."""
import os
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch

    import bench

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lsio_b200 import sharding

        mine = sharding.files_for_rank(11, rank, world)
        # pretend each file carries (idx + 1) MiB and this rank took (rank + 1) seconds
        nbytes = sum((i + 1) << 20 for i in mine)
        total_bytes, max_secs = bench.reduce_sum_max(nbytes, float(rank + 1), device=torch.device("cpu"))
        rates = bench.gather_floats(10.0 * (rank + 1), torch.device("cpu"))
        first, count, per = bench.cpu_slice(rank, world)
        q.put((rank, mine, total_bytes, max_secs, rates, (first, count, per)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_and_aggregate():
    world, port = 2, 29571 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    files = sorted(res[0][1] + res[1][1])
    assert files == list(range(11))                          # disjoint and complete
    want_bytes = sum((i + 1) << 20 for i in range(11))
    for _, _, total_bytes, max_secs, rates, _ in res:
        assert total_bytes == want_bytes and max_secs == 2.0  # sum of units, max of time
        assert rates == [10.0, 20.0]                          # per-rank rates, in rank order, on every rank
    # every rank pins its backend threads to its own slice of the CPUs: disjoint, equal share
    (f0, c0, per0), (f1, c1, per1) = res[0][5], res[1][5]
    assert per0 == per1 == len(os.sched_getaffinity(0)) // 2
    if c0 and c1:
        assert c0 == c1 == per0 and set(range(f0, f0 + c0)).isdisjoint(range(f1, f1 + c1))


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
def test_bench_deals_disjoint_physical_files(n_gpus):
    """bench.py's weak-scaling shard: every rank lands the same number of files per step, no physical file is read by two ranks,
    together the ranks cover the whole set, and a rank's walk over its share is balanced (each of its files the same number of
    times, give or take one)."""
    sys.path.insert(0, ROOT)
    import bench

    n_files = 500
    shares = [bench.deal_files(n_files, r, n_gpus) for r in range(n_gpus)]
    assert all(len(s) == n_files for s in shares)
    sets = [set(s) for s in shares]
    assert set().union(*sets) == set(range(n_files))
    for a in range(n_gpus):
        for b in range(a + 1, n_gpus):
            assert sets[a].isdisjoint(sets[b])
    for s in shares:
        counts = [s.count(f) for f in set(s)]
        assert max(counts) - min(counts) <= 1
    # the round-1 dealing, kept selectable, makes ranks r and r+4 share files at 8 ranks -- the reason it was replaced
    if n_gpus == 8:
        old = [set(bench.deal_files(n_files, r, 8, mode=0)) for r in range(8)]
        assert old[0] == old[4] and old[0].isdisjoint(old[1])
