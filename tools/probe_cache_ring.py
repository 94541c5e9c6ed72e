"""Does a cache-resident destination make tmpfs reads faster? (raw READ_FIXED probe, no GPU work)"""
import os, sys, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from lsio_b200 import cuda_ops
paths, _ = bench.make_files("/dev/shm/lsio_bench", 200, 0)
for threads in (8, 12):
    for bs, qd in ((256 << 10, 1), (256 << 10, 2), (256 << 10, 4), (512 << 10, 2), (1 << 20, 1), (1 << 20, 2), (1 << 20, 8), (1 << 20, 32), (4 << 20, 8)):
        g = cuda_ops.probe_read_ceiling(paths, bs, threads, qd, True)
        print(f"threads={threads} block={bs >> 10}K qd={qd} buffer/thread={bs * qd >> 10}K total={threads * bs * qd >> 20}MiB: {g:.1f} GiB/s", flush=True)
