"""Diagnose bad chunks on the disk-backed path (async completions)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, lsio_b200 as L
from lsio_b200.io_uring_cuda import Receiver
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
paths, crc = bench.make_files("/tmp/lsio_bench_disk", 60, 0)
starts, ends = bench.chunk_ranges()
for rnd in range(12):
    u = L.IoUringCuda(8, arena_bytes=8 << 30, staging_mode=mode)
    rx = u.completion()
    for i, p in enumerate(paths):
        u.get_ranges_arrays(p, starts, ends, (np.uint64(i) << np.uint64(32)) | np.arange(bench.N_CHUNKS, dtype=np.uint64))
    got = 0; bad = []; errs = []
    while got < len(paths) * bench.N_CHUNKS:
        rec = rx.recv_batch(4096, 60.0)
        for r in rec:
            ud = int(r["chunk"]["user_data"]); f, k = ud >> 32, ud & 0xFFFFFFFF
            if r["status"] != 0: errs.append((f, k, int(r["status"]), r["context"][:120]))
            elif r["crc32c"] != crc[f, k]:
                v = L.AlignedBytes(int(r["chunk"]["buffer"])); data = v.as_slice(); v._h = None  # borrowed: the record keeps the handle, drop_buffers() below frees it
                want = np.fromfile(paths[f], dtype=np.uint8, count=bench.BLOCK, offset=k * bench.BLOCK)
                diff = np.nonzero(data != want)[0]
                got_crc = int(r["crc32c"])
                # where do the landed bytes come from instead?
                src = -1
                if len(diff):
                    whole = np.fromfile(paths[f], dtype=np.uint8)
                    probe = data[int(diff[0]):int(diff[0]) + 64].tobytes()
                    src = whole.tobytes().find(probe)
                bad.append((f, k, len(diff), int(diff[0]) if len(diff) else -1, int(diff[-1]) if len(diff) else -1, hex(got_crc), hex(int(crc[f, k])), src))
        Receiver.drop_buffers(rec); got += len(rec)
    print("round", rnd, "mode", mode, "bad", len(bad), bad[:12], "errs", errs[:5], u.stats())
    u.drop()
