"""First-light check of the IoUringCuda backend on a GPU box."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lsio_b200 as L
from oracle import oracle

def drain(rx, n, timeout=10.0):
    out = []
    for _ in range(n):
        out.append(rx.recv_timeout(timeout))
    return out

def test_integration(tmpdir, o_direct=1):
    FILE, CH = 1 << 20, 4096; N = FILE // CH
    data = bytes(random.randrange(32, 127) for _ in range(6144)); data = (data * (FILE // 6144 + 1))[:FILE]
    fn = os.path.join(tmpdir, "lsio_it_%d" % os.getpid()); open(fn, "wb").write(data)
    u = L.IoUringCuda(4, o_direct=o_direct, arena_bytes=1 << 30, staging_bytes=64 << 20)
    u.get_ranges(fn, [(i * CH, (i + 1) * CH) for i in range(N)], range(N))
    got = [None] * N
    for c in drain(u.completion(), N, 5.0):
        assert c.is_ok(), c
        b = c.buffer.as_slice()
        assert c.crc32c == oracle.crc32c(b), (c.user_data, c.crc32c, oracle.crc32c(b))
        got[c.user_data] = b.tobytes()
    rx = u.completion().clone(); u.drop()
    try: rx.recv(); raise SystemExit("expected disconnect")
    except L.Disconnected: pass
    assert b"".join(got) == data
    os.unlink(fn); print("integration ok (o_direct=%d) in %s" % (o_direct, tmpdir))

def test_unaligned(tmpdir):
    size = (1 << 20) + 1000
    fn = os.path.join(tmpdir, "lsio_un_%d" % os.getpid()); oracle.write_file(fn, 7, size)
    rng = np.random.default_rng(3)
    ranges = []
    for _ in range(300):
        ln = int(rng.integers(1, 65536)); st = int(rng.integers(0, size - ln)); ranges.append((st, st + ln))
    ranges += [(0, -1), (-100, -1), (size - 5, size), (0, 1), (511, 513), (-size, 1), (5, 1 << 20), (513, (1 << 20) + 999)]
    u = L.IoUringCuda(3, arena_bytes=1 << 30, staging_bytes=64 << 20, max_read_bytes=1 << 20)
    u.get_ranges(fn, ranges, range(len(ranges)))
    n_ok = 0
    for c in drain(u.completion(), len(ranges), 10.0):
        assert c.is_ok(), c
        want = oracle.pread_range(fn, *ranges[c.user_data])
        got = c.buffer.as_slice()
        assert np.array_equal(got, want), (c.user_data, ranges[c.user_data], len(got), len(want))
        assert c.crc32c == oracle.crc32c(want), (c.user_data, ranges[c.user_data])
        n_ok += 1
    # errors: missing file, bad ranges
    u.get_ranges("/tmp/definitely_missing_lsio", [(0, 10)], [1])
    u.get_ranges(fn, [(0, size + 1), (10, 10), (4096, 8192)], [10, 11, 12])
    items = drain(u.completion(), 5, 5.0)
    errs = [i for i in items if not i.is_ok()]; oks = [i for i in items if i.is_ok()]
    assert len(errs) == 4 and len(oks) == 1 and oks[0].user_data == 12, items
    print("errors:", [(e.status, e.opcode, e.user_data) for e in errs])
    print("stats", u.stats())
    u.drop(); os.unlink(fn); print("unaligned ok", n_ok, "in", tmpdir)

def bench(tmpdir, nfiles, fsize, bs, workers, **cfg):
    paths = []
    t0 = time.time()
    for i in range(nfiles):
        p = os.path.join(tmpdir, "lsio_bench_%d" % i)
        if not (os.path.exists(p) and os.path.getsize(p) == fsize): oracle.write_file(p, i, fsize)
        paths.append(p)
    nch = fsize // bs
    st = np.arange(nch, dtype=np.int64) * bs; en = st + bs
    u = L.IoUringCuda(workers, **cfg)
    rx = u.completion()
    t1 = time.time()
    for i, p in enumerate(paths):
        u.get_ranges_arrays(p, st, en, (np.uint64(i) << np.uint64(32)) | np.arange(nch, dtype=np.uint64))
    total = nfiles * nch; nbytes = 0; bad = 0
    for _ in range(total):
        c = rx.recv_timeout(30.0)
        if not c.is_ok(): bad += 1; continue
        nbytes += c.buffer.len(); c.buffer.drop()
    dt = time.time() - t1
    s = u.stats(); u.drop()
    print(f"bench {tmpdir} files={nfiles} fsize={fsize} bs={bs} workers={workers} cfg={cfg}: {nbytes/dt/2**30:.2f} GiB/s ({dt:.2f}s) bad={bad} copies={s['h2d_copies']} avgcopy={s['h2d_bytes']/max(1,s['h2d_copies'])/2**20:.1f}MiB stalls={s['staging_stalls']},{s['arena_stalls']} short={s['short_reads']} gen={t1-t0:.1f}s")
    return paths

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "func"):
        test_integration("/dev/shm"); test_integration("/tmp"); test_integration("/tmp", o_direct=0)
        test_unaligned("/dev/shm"); test_unaligned("/tmp")
    if which in ("all", "bench"):
        for w in (4, 8, 12):
            paths = bench("/dev/shm", 100, 41214400, 262144, w)
        bench("/dev/shm", 100, 41214400, 262144, 8, min_copy_bytes=4 << 20)
        bench("/dev/shm", 100, 41214400, 262144, 8, min_copy_bytes=16 << 20, staging_bytes=2 << 30)
        bench("/dev/shm", 100, 41214400, 262144, 8, checksum=0)
        bench("/dev/shm", 100, 41214400, 1 << 20, 8)
        for p in paths: os.unlink(p)
        paths = bench("/tmp", 50, 41214400, 262144, 8)
        for p in paths: os.unlink(p)
