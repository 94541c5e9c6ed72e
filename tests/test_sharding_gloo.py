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
        total_bytes, max_secs = bench.reduce_bytes_and_time(nbytes, float(rank + 1), device=torch.device("cpu"))
        q.put((rank, mine, total_bytes, max_secs))
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
    for _, _, total_bytes, max_secs in res:
        assert total_bytes == want_bytes and max_secs == 2.0  # sum of units, max of time
