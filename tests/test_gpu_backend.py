"""GPU parity tests of the drop-in backend through the C ABI (-m gpu): IoUringCuda vs the
oracle (pread ground truth + CRC), restating crates/lsio_uring/tests/integration_test.rs and
covering what the reference leaves untested."""
import os
import random
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
KIB, MIB = 1024, 1024 * 1024


def _drain(rx, n, timeout=20.0):
    return [rx.recv_timeout(timeout) for _ in range(n)]


def _check_chunks(oracle, items, fn_of, ranges_of):
    for c in items:
        assert c.is_ok(), c
        fn, (s, e) = fn_of(c.user_data), ranges_of(c.user_data)
        want = oracle.pread_range(fn, s, e)
        got = c.buffer.as_slice()
        assert np.array_equal(got, want), (c.user_data, s, e, len(got), len(want))
        assert c.crc32c == oracle.crc32c(want), (c.user_data, s, e)
        assert c.buffer.memory_kind == "device" and c.device == 0


@pytest.mark.parametrize("o_direct,statx_via_ring,sqpoll", [(1, 0, 0), (0, 0, 0), (1, 1, 0), (1, 1, 1)])
def test_get_ranges(lsio, oracle, cuda_dev, data_dir, o_direct, statx_via_ring, sqpoll):
    """integration_test.rs::test_get_ranges against the CUDA backend."""
    N_WORKER_THREADS, FILE_SIZE, CHUNK_SIZE = 4, MIB, 4 * KIB
    N_CHUNKS = FILE_SIZE // CHUNK_SIZE
    period = int(CHUNK_SIZE * 1.5)
    unit = bytes(random.randrange(32, 127) for _ in range(period))
    file_contents = (unit * (FILE_SIZE // period + 1))[:FILE_SIZE]
    filename = os.path.join(data_dir, f"lsio_uring_tempfile_{random.getrandbits(32)}")
    with open(filename, "wb") as f:
        f.write(file_contents)
        f.flush()
        os.fsync(f.fileno())
    uring = lsio.IoUringCuda(N_WORKER_THREADS, o_direct=o_direct, statx_via_ring=statx_via_ring, sqpoll=sqpoll, arena_bytes=256 << 20, staging_bytes=32 << 20)  # (1, 1, 1) is the reference's own ring configuration
    uring.get_ranges(filename, [(i * CHUNK_SIZE, (i + 1) * CHUNK_SIZE) for i in range(N_CHUNKS)], range(N_CHUNKS))
    vec = [None] * N_CHUNKS
    for i in range(N_CHUNKS):
        c = uring.completion().recv_timeout(0.5 * 10)
        assert c.is_ok(), f"Error reading chunk {i}! {c}"
        vec[c.user_data] = c.buffer
    completion = uring.completion().clone()
    uring.drop()
    with pytest.raises(lsio.Disconnected):
        completion.recv()  # integration_test.rs:92-96
    completion.drop()
    assembled = b"".join(b.as_slice().tobytes() for b in vec)  # chunks outlive the backend
    assert assembled == file_contents


def test_unaligned_negative_eof_and_split_ranges(lsio, oracle, cuda_dev, data_dir):
    size = MIB + 1000  # unaligned EOF
    fn = os.path.join(data_dir, "f")
    oracle.write_file(fn, 7, size)
    rng = np.random.default_rng(3)
    ranges = [(0, -1), (-100, -1), (size - 5, size), (0, 1), (511, 513), (-size, 1), (5, MIB), (513, MIB + 999), (size - 1, size),
              (-1, -1), (1024, 2048), (MIB, -1)]
    for _ in range(300):
        n = int(rng.integers(1, 65536))
        s = int(rng.integers(0, size - n))
        ranges.append((s, s + n))
    # max_read_bytes = 128 KiB: long ranges are read as several segments whose CRCs are combined
    for merge, staging_mode in ((0, 1), (1, 1), (1, 0)):  # staging_mode 1: per-worker slot extents; 0: one big pinned ring
        with lsio.IoUringCuda(3, arena_bytes=256 << 20, staging_bytes=32 << 20, max_read_bytes=128 << 10, merge_ranges=merge,
                              staging_mode=staging_mode, host_slots_per_worker=3) as u:
            u.get_ranges(fn, ranges, range(len(ranges)))
            items = _drain(u.completion(), len(ranges))
            assert sorted(c.user_data for c in items) == list(range(len(ranges)))  # exactly one chunk per range
            _check_chunks(oracle, items, lambda ud: fn, lambda ud: ranges[ud])
            st = u.stats()
            assert st["errors"] == 0
            if not merge:
                assert st["reads"] > len(ranges)  # one read per range, long ones split (reference: one read per range)


@pytest.mark.parametrize("pack_align,staging_mode", [(1, 1), (16, 1), (512, 1), (1, 0)])
def test_many_files_many_ops_per_thread(lsio, oracle, cuda_dev, data_dir, pack_align, staging_mode):
    """Multiple files in flight, > 64 ops per worker, back-pressure on a tiny completion channel."""
    n_files, size = 24, 300_000 + 77
    fns = []
    for i in range(n_files):
        fn = os.path.join(data_dir, f"f{i}")
        oracle.write_file(fn, 100 + i, size)
        fns.append(fn)
    rng = np.random.default_rng(4)
    per_file = 40
    all_ranges = {}
    with lsio.IoUringCuda(2, arena_bytes=512 << 20, staging_bytes=16 << 20, completion_bound=8, pack_align=pack_align, sq_entries=16,
                          staging_mode=staging_mode, host_slots_per_worker=2) as u:
        for i, fn in enumerate(fns):
            rs = []
            for k in range(per_file):
                n = int(rng.integers(1, 30000))
                s = int(rng.integers(0, size - n))
                rs.append((s, s + n))
                all_ranges[(i << 32) | k] = rs[-1]
            u.get_ranges(fn, rs, [(i << 32) | k for k in range(per_file)])
        time.sleep(0.3)  # let the 8-slot channel fill: producers must block, not drop
        assert len(u.completion()) <= 8
        items = _drain(u.completion(), n_files * per_file)
        assert len({c.user_data for c in items}) == n_files * per_file
        _check_chunks(oracle, items, lambda ud: fns[ud >> 32], lambda ud: all_ranges[ud])
        if pack_align > 1:
            assert all(c.buffer.as_ptr() % pack_align == 0 for c in items)


@pytest.mark.parametrize("statx_via_ring", [1, 0])
def test_errors_are_in_band(lsio, oracle, cuda_dev, data_dir, statx_via_ring):
    fn = os.path.join(data_dir, "f")
    size = 10_000
    oracle.write_file(fn, 1, size)
    with lsio.IoUringCuda(2, arena_bytes=64 << 20, staging_bytes=16 << 20, statx_via_ring=statx_via_ring) as u:
        rx = u.completion()
        u.get_ranges("/definitely/missing/lsio", [(0, 10), (10, 20)], [1, 2])
        if statx_via_ring:
            # reference mode: openat and statx both fail => two Err items, no chunk (get_ranges.rs:49-65)
            errs = _drain(rx, 2)
            assert sorted((e.status, e.opcode) for e in errs) == [(-2, 18), (-2, 21)]
        else:
            # default: one Err per bad file, as the reference's doc comment promises (lsio_io lib.rs:30-35)
            errs = _drain(rx, 1)
            assert [(e.status, e.opcode) for e in errs] == [(-2, 18)]
        assert all("/definitely/missing/lsio" in e.context for e in errs)
        # bad ranges become Err items (the reference asserts/panics: sqe.rs:71,81,88), good ones still arrive
        u.get_ranges(fn, [(0, size + 1), (10, 10), (-size - 5, 10), (4096, 8192), (0, -size - 2)], [10, 11, 12, 13, 14])
        items = _drain(rx, 5)
        ok = [c for c in items if c.is_ok()]
        bad = [c for c in items if not c.is_ok()]
        assert [c.user_data for c in ok] == [13]
        assert sorted(e.user_data for e in bad) == [10, 11, 12, 14] and all(e.status == -22 for e in bad)
        _check_chunks(oracle, ok, lambda ud: fn, lambda ud: (4096, 8192))
        with pytest.raises(lsio.RecvTimeoutError):
            rx.recv_timeout(0.2)
        # argument errors (the reference panics on a length mismatch, get_ranges.rs:28)
        with pytest.raises(ValueError):
            u.get_ranges(fn, [(0, 1), (1, 2)], [1])
        with pytest.raises(ValueError):
            u.get_ranges("relative/path", [(0, 1)], [1])
        # a request larger than the arena is an error, not a hang
        u.get_ranges(fn, [(0, -1)], [99])
        assert _drain(rx, 1)[0].is_ok()


def test_request_larger_than_arena_and_arena_recycling(lsio, oracle, cuda_dev, data_dir):
    fn = os.path.join(data_dir, "big")
    size = 6 * MIB
    oracle.write_file(fn, 2, size)
    with lsio.IoUringCuda(2, arena_bytes=4 << 20, staging_bytes=16 << 20) as u:
        rx = u.completion()
        u.get_ranges(fn, [(0, -1)], [1])
        e = rx.recv_timeout(10)
        assert not e.is_ok() and e.status == -12  # ENOMEM
        # 3 MiB requests against a 4 MiB arena: each waits for the previous chunks to be dropped
        for k in range(4):
            u.get_ranges(fn, [(k * MIB, k * MIB + 3 * MIB // 2), (k * MIB + 3 * MIB // 2, k * MIB + 3 * MIB)], [2 * k, 2 * k + 1])
        seen = 0
        while seen < 8:
            c = rx.recv_timeout(10)
            assert c.is_ok()
            k, half = divmod(c.user_data, 2)
            s = k * MIB + half * (3 * MIB // 2)
            assert np.array_equal(c.buffer.as_slice(), oracle.pattern(2, s, 3 * MIB // 2))
            c.buffer.drop()
            seen += 1
        assert u.stats()["arena_stalls"] > 0


def test_reference_bench_pattern_and_bulk_receive(lsio, oracle, cuda_dev, data_dir):
    """lsio_bench's own data (`i as u8`, main.rs:102) and chunking (filesize / blocksize blocks),
    received with the bulk API; every CRC is the Appendix C value for that block."""
    fsize, bs = 1_000_000, 262_144
    n_chunks = fsize // bs
    fns = []
    for i in range(6):
        fn = os.path.join(data_dir, f"lsio_bench_{i}")
        oracle.write_file(fn, i, fsize, ref_pattern=True)
        fns.append(fn)
    st = np.arange(n_chunks, dtype=np.int64) * bs
    with lsio.IoUringCuda(4, arena_bytes=128 << 20, staging_bytes=16 << 20) as u:
        for i, fn in enumerate(fns):
            u.get_ranges_arrays(fn, st, st + bs, (np.uint64(i) << np.uint64(32)) | np.arange(n_chunks, dtype=np.uint64))
        got = 0
        while got < len(fns) * n_chunks:
            rec = u.completion().recv_batch(64, 10.0)
            assert (rec["status"] == 0).all()
            assert (rec["crc32c"] == 0x1CA1B1D2).all()  # crc32c of one 256 KiB `i as u8` block
            lsio.Receiver.drop_buffers(rec)
            got += len(rec)


def test_two_device_groups_on_one_gpu(lsio, oracle, cuda_dev, data_dir):
    """Sharding plumbing: two worker groups (here both on GPU 0), files routed round-robin and explicitly."""
    fns = []
    for i in range(6):
        fn = os.path.join(data_dir, f"f{i}")
        oracle.write_file(fn, 50 + i, 200_000)
        fns.append(fn)
    with lsio.IoUringCuda(2, devices=(0, 0), arena_bytes=64 << 20, staging_bytes=16 << 20) as u:
        for i, fn in enumerate(fns):
            u.get_ranges(fn, [(1, 100_001), (100_001, -1)], [2 * i, 2 * i + 1], group=None if i % 2 else i % 2)
        items = _drain(u.completion(), 12)
        _check_chunks(oracle, items, lambda ud: fns[ud // 2], lambda ud: (1, 100_001) if ud % 2 == 0 else (100_001, -1))


def test_zarr_like_config3_sample(lsio, oracle, cuda_dev, data_dir):
    """BASELINE config 3 at a size the oracle finishes in seconds: 200 files x 1,049,576 B,
    8 ranges per file of 4-64 KiB at arbitrary offsets, plus EOF-anchored and negative ranges."""
    n_files, size = 200, 1_049_576
    fns = []
    for i in range(n_files):
        fn = os.path.join(data_dir, f"z{i}")
        oracle.write_file(fn, 1000 + i, size)
        fns.append(fn)
    ranges = {}
    with lsio.IoUringCuda(4, arena_bytes=512 << 20, staging_bytes=64 << 20) as u:
        for i, fn in enumerate(fns):
            rng = np.random.default_rng(0xC3 + i)
            rs = []
            for k in range(8):
                n = int(rng.integers(4096, 65537))
                s = int(rng.integers(0, size - n + 1))
                rs.append((s, s + n))
            if i % 100 == 0:
                rs.append((size - 5000, size))
                rs.append((-5000, -1))
            for k, r in enumerate(rs):
                ranges[(i << 32) | k] = r
            u.get_ranges(fn, rs, [(i << 32) | k for k in range(len(rs))])
        items = _drain(u.completion(), len(ranges), 30.0)
        _check_chunks(oracle, items, lambda ud: fns[ud >> 32], lambda ud: ranges[ud])
        # the bytes are also provable on the device without a D2H copy
        ptrs = np.array([c.buffer.as_ptr() for c in items], dtype=np.uint64)
        lens = np.array([c.buffer.len() for c in items], dtype=np.uint64)
        ids = np.array([1000 + (c.user_data >> 32) for c in items], dtype=np.uint64)
        offs = np.array([oracle.resolve_range(size, *ranges[c.user_data])[1].start_off for c in items], dtype=np.uint64)
        import torch

        anchor = torch.empty(1, dtype=torch.uint8, device=cuda_dev)
        assert lsio.cuda_ops.verify_pattern(anchor, ptrs - np.uint64(anchor.data_ptr()), lens, ids, offs) == 0


def test_smoke_entry(cuda_dev):
    import __graft_entry__ as g

    g.smoke()


@pytest.mark.parametrize("checksum", [0, 2, 3])
def test_checksum_selection_including_xxh64(lsio, oracle, cuda_dev, data_dir, checksum):
    """checksum bits: 1 = CRC-32C fused into the pack kernel, 2 = XXH64 second stage (xxh64_chunks)."""
    size = 700_001
    fn = os.path.join(data_dir, "x")
    oracle.write_file(fn, 21, size)
    rng = np.random.default_rng(8)
    ranges = [(0, -1), (3, 35), (1, 2), (size - 31, size)]
    for _ in range(120):
        n = int(rng.integers(1, 200_000))
        s = int(rng.integers(0, size - n))
        ranges.append((s, s + n))
    with lsio.IoUringCuda(3, arena_bytes=256 << 20, staging_bytes=32 << 20, checksum=checksum, max_read_bytes=64 << 10) as u:
        u.get_ranges(fn, ranges, range(len(ranges)))
        items = _drain(u.completion(), len(ranges))
        for c in items:
            assert c.is_ok(), c
            want = oracle.pread_range(fn, *ranges[c.user_data])
            assert np.array_equal(c.buffer.as_slice(), want)
            assert (c.crc32c is not None) == bool(checksum & 1) and (c.xxh64 is not None) == bool(checksum & 2)
            if checksum & 1:
                assert c.crc32c == oracle.crc32c(want)
            if checksum & 2:
                assert c.xxh64 == oracle.xxh64(want), ranges[c.user_data]


def test_drop_with_work_in_flight_and_concurrent_submitters(lsio, oracle, cuda_dev, data_dir):
    """Teardown must not hang or crash with requests queued/in flight and nobody draining
    (the reference asserts nothing is in flight at drop, worker.rs:124); chunks already received
    stay valid after the backend is gone; several threads may submit at once."""
    import threading

    fns = []
    for i in range(16):
        fn = os.path.join(data_dir, f"s{i}")
        oracle.write_file(fn, 300 + i, 2_000_000)
        fns.append(fn)
    for round_ in range(3):
        u = lsio.IoUringCuda(3, arena_bytes=64 << 20, staging_bytes=16 << 20, completion_bound=16)
        rx = u.completion().clone()

        def submit(k):
            for j in range(4):
                i = (4 * k + j) % len(fns)
                u.get_ranges(fns[i], [(b * 100_000, (b + 1) * 100_000) for b in range(20)], [(i << 32) | b for b in range(20)])

        ths = [threading.Thread(target=submit, args=(k,)) for k in range(4)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        kept = [rx.recv_timeout(10) for _ in range(5 + round_)]
        t0 = time.time()
        u.drop()  # 320 ranges submitted, ~6 received, channel bound 16: must still tear down promptly
        assert time.time() - t0 < 20
        for c in kept:
            assert c.is_ok()
            i, b = c.user_data >> 32, c.user_data & 0xFFFFFFFF
            assert np.array_equal(c.buffer.as_slice(), oracle.pattern(300 + i, b * 100_000, 100_000))
        seen = 0
        with pytest.raises(lsio.Disconnected):
            while True:
                rx.recv_timeout(5)
                seen += 1
        assert seen <= 16
        rx.drop()


@pytest.mark.parametrize("merge,gap,zero_copy", [(1, 0, 0), (0, 0, 0), (1, 4096, 0), (1, 0, 1)])
def test_read_planner_merges_adjacent_overlapping_and_duplicate_ranges(lsio, oracle, cuda_dev, data_dir, merge, gap, zero_copy):
    """SURVEY 8f rank 1 (planned in the reference: lsio_aligned_bytes/README.md:43-57): ranges whose
    padded spans touch or overlap share one read; every range still gets exactly one chunk with its own
    bytes and CRC, whatever the merge setting."""
    size = 3 * MIB + 333
    fn = os.path.join(data_dir, "m")
    oracle.write_file(fn, 31, size)
    bs = 4096
    ranges = [(i * bs, (i + 1) * bs) for i in range(256)]            # 256 adjacent blocks: one 1 MiB run
    ranges += [(MIB + 10, MIB + 5000), (MIB + 4000, MIB + 9000), (MIB + 10, MIB + 5000)]  # overlapping + duplicate
    ranges += [(2 * MIB + 100, 2 * MIB + 200), (2 * MIB + 3000, 2 * MIB + 3100)]          # 2.8 KiB apart: merged only with a gap
    ranges += [(5, 2 * MIB + 777)]                                     # one long range over everything before it
    ranges = [ranges[i] for i in np.random.default_rng(5).permutation(len(ranges))]  # submission order must not matter
    with lsio.IoUringCuda(3, arena_bytes=256 << 20, staging_bytes=32 << 20, merge_ranges=merge, merge_gap_bytes=gap, max_read_bytes=512 << 10, zero_copy=zero_copy) as u:
        u.get_ranges(fn, ranges, range(len(ranges)))
        items = _drain(u.completion(), len(ranges))
        assert sorted(c.user_data for c in items) == list(range(len(ranges)))
        _check_chunks(oracle, items, lambda ud: fn, lambda ud: ranges[ud])
        st = u.stats()
        if merge:
            assert st["reads"] < 40, st      # everything lies in one run up to 2 MiB+..., cut into 512 KiB reads
            assert st["read_bytes"] < 2.5 * size
        else:
            assert st["reads"] >= len(ranges)


def test_native_bench_driver(lsio, oracle, cuda_dev, data_dir):
    """lsio_bench::read_files restated natively (include/lsio_bench.h): counts delivered bytes
    (not filesize * n_files) and checks every chunk's CRC."""
    fsize, bs = 1_000_000, 262_144
    n_chunks = fsize // bs
    fns, table = [], np.zeros((4, n_chunks), dtype=np.uint32)
    for i in range(4):
        fn = os.path.join(data_dir, f"lsio_bench_{i}")
        oracle.write_file(fn, 40 + i, fsize)
        fns.append(fn)
        for k in range(n_chunks):
            table[i, k] = oracle.crc32c(oracle.pattern(40 + i, k * bs, bs))
    with lsio.IoUringCuda(4, arena_bytes=64 << 20, staging_bytes=16 << 20) as u:
        secs, nbytes, ok, bad = u.bench_read_files(fns, range(4), bs, n_chunks, table)
        assert (nbytes, ok, bad) == (4 * n_chunks * bs, 4 * n_chunks, 0) and secs > 0
        table[2, 1] ^= 1  # a wrong expectation must be noticed
        _, nbytes, ok, bad = u.bench_read_files(fns, range(4), bs, n_chunks, table)
        assert (ok, bad) == (4 * n_chunks - 1, 1)


def test_range_larger_than_two_gib(lsio, oracle, cuda_dev):
    """The reference panics above 2,147,479,552 B (get_range.rs:23-30); here the read is split and the
    segment CRCs are combined.  A sparse tmpfs file keeps it cheap: 2.5 GiB of zeros + a marker."""
    import torch

    if not os.path.isdir("/dev/shm"):
        pytest.skip("needs tmpfs")
    size = (5 << 29) + 123
    fn = f"/dev/shm/lsio_big_{os.getpid()}"
    try:
        with open(fn, "wb") as f:
            f.truncate(size)
            f.seek(size - 100)
            f.write(b"\x5a" * 100)
        with lsio.IoUringCuda(4, arena_bytes=3 << 30, staging_bytes=256 << 20) as u:
            u.get_ranges(fn, [(1, -1)], [7])
            c = u.completion().recv_timeout(120)
            assert c.is_ok() and c.user_data == 7 and c.buffer.len() == size - 1
            t = c.buffer.to_torch()
            assert int(torch.count_nonzero(t[:-100])) == 0 and bool((t[-100:] == 0x5A).all())
            # crc32c(zeros(n) + 0x5a*100) from crc32c of a small zero block by repeated combining
            z = oracle.crc32c(bytes(1 << 20))
            n, acc, blk, blen = size - 1 - 100, 0, z, 1 << 20
            want, first, k = 0, True, n // (1 << 20)
            while k:  # binary decomposition of the zero run
                if k & 1:
                    want = blk if first else oracle.crc32c_combine(want, blk, blen)
                    first = False
                blk = oracle.crc32c_combine(blk, blk, blen)
                blen *= 2
                k >>= 1
            tail = bytes(n % (1 << 20)) + b"\x5a" * 100
            want = oracle.crc32c_combine(want, oracle.crc32c(tail), len(tail))
            assert c.crc32c == want
            assert u.stats()["reads"] >= size // (512 << 10)  # split at max_read_bytes (512 KiB)
    finally:
        if os.path.exists(fn):
            os.unlink(fn)


def test_lsio_bench_cli(lsio, cuda_dev, data_dir, capsys):
    """The reference's CLI (crates/lsio_bench/src/main.rs) with the same flags, on its own data pattern."""
    from lsio_b200 import lsio_bench

    rc = lsio_bench.main(["--directory", data_dir, "--nrfiles", "3", "--filesize", "1000000", "--blocksize", "262144", "-w", "2"])
    out = capsys.readouterr().out
    assert rc == 0
    assert "3 files x 3 chunks per file = 9 total chunks" in out
    assert f"Delivered into HBM: {9 * 262144} bytes in 9 chunks (0 bad)" in out
    assert "Total bandwidth = " in out


def test_plain_c_consumer_of_the_abi(lsio, cuda_dev, data_dir, tmp_path):
    """integration_test.rs restated in C (tests/c/integration_test.c), linked against liblsio_b200.so: the
    boundary is a plain C ABI, usable with no Python and no C++ on the caller's side."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(lsio._native.LIB_PATH)
    exe = str(tmp_path / "integration_test")
    subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", f"-I{root}/include", f"{root}/tests/c/integration_test.c", f"-L{libdir}", "-llsio_b200",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    p = subprocess.run([exe, data_dir], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "256 chunks of 4096 bytes landed in HBM and match the file" in p.stdout


@pytest.mark.parametrize("staging_mode", [1, 0])
def test_out_of_order_completions_of_large_reads(lsio, oracle, cuda_dev, data_dir, staging_mode):
    """Regression: with asynchronous (disk) reads an extent allocated earlier can complete -- and be copied to the
    device -- after a later one of the same worker; the pack stage must wait for every extent's copy, not just the
    last one per worker.  512 MiB in 2 MiB reads, three rounds, every 256 KiB chunk checked against the oracle CRC."""
    n_files, fsize, bs = 16, 32 * MIB, 256 * KIB
    n_chunks = fsize // bs
    fns, table = [], np.zeros((n_files, n_chunks), dtype=np.uint32)
    for i in range(n_files):
        fn = os.path.join(data_dir, f"big{i}")
        oracle.write_file(fn, 700 + i, fsize)
        with open(fn, "rb") as f:
            os.fsync(f.fileno())
        img = oracle.pattern(700 + i, 0, fsize)
        for k in range(n_chunks):
            table[i, k] = oracle.crc32c(img[k * bs:(k + 1) * bs])
        fns.append(fn)
    with lsio.IoUringCuda(8, arena_bytes=2 << 30, staging_bytes=256 << 20, staging_mode=staging_mode) as u:
        for _ in range(3):
            secs, nbytes, ok, bad = u.bench_read_files(fns, range(n_files), bs, n_chunks, table)
            assert (ok, bad, nbytes) == (n_files * n_chunks, 0, n_files * fsize)


def test_in_process_sharding_over_two_real_gpus(lsio, oracle, cuda_dev, data_dir):
    """One backend, two device groups on two different GPUs: files dealt round-robin (and one routed explicitly);
    every chunk is a view in the HBM of the GPU its file was routed to.  Skips on single-GPU boxes."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fns = []
    for i in range(8):
        fn = os.path.join(data_dir, f"g{i}")
        oracle.write_file(fn, 900 + i, 3_000_000 + i)
        fns.append(fn)
    with lsio.IoUringCuda(3, devices=(0, 1), arena_bytes=256 << 20, staging_bytes=64 << 20) as u:
        for i, fn in enumerate(fns):
            u.get_ranges(fn, [(7, 1_500_000), (1_500_000, -1)], [2 * i, 2 * i + 1], group=1 if i == 0 else None)
        items = _drain(u.completion(), 16, 30.0)
        devices = {}
        for c in items:
            assert c.is_ok(), c
            i, half = divmod(c.user_data, 2)
            want = oracle.pread_range(fns[i], *((7, 1_500_000) if half == 0 else (1_500_000, -1)))
            assert np.array_equal(c.buffer.as_slice(), want) and c.crc32c == oracle.crc32c(want)
            assert c.buffer.device == c.device and c.buffer.to_torch().device.index == c.device
            devices.setdefault(i, set()).add(c.device)
        assert all(len(d) == 1 for d in devices.values())          # a file (request) lives on one GPU
        assert devices[0] == {1}                                   # explicit routing
        assert {next(iter(d)) for d in devices.values()} == {0, 1}  # both GPUs were used


def test_randomized_configurations_stress(lsio, oracle, cuda_dev):
    """Randomised soak over the whole configuration space (both staging modes, tmpfs and disk, slot / read / merge /
    checksum / alignment / queue settings, valid and invalid ranges, a consumer that sometimes lags): every chunk is
    compared with the oracle's pread bytes and checksums; every range yields exactly one item."""
    import shutil
    import tempfile

    rng = np.random.default_rng(2026)
    dirs = [tempfile.mkdtemp(prefix="lsio_stress_", dir=d) for d in (("/dev/shm", "/tmp") if os.path.isdir("/dev/shm") else ("/tmp",))]
    try:
        files = []
        for i in range(14):
            size = int(rng.choice([1, 511, 4096, 70_001, 1_048_576, 3_000_003, 9_437_184 + 17]))
            fn = os.path.join(dirs[i % len(dirs)], f"s{i}")
            oracle.write_file(fn, 5000 + i, size)
            files.append((fn, size))
        for it in range(int(os.environ.get("LSIO_STRESS_ITERS", 40))):
            cfg = dict(
                staging_mode=int(rng.integers(0, 2)), host_slots_per_worker=int(rng.integers(1, 6)), host_slot_kib=int(rng.choice([128, 512, 2048])),
                max_read_bytes=int(rng.choice([4096, 65536, 524288, 4 << 20])), merge_ranges=int(rng.integers(0, 2)),
                merge_gap_bytes=int(rng.choice([0, 4096])), checksum=int(rng.integers(0, 4)), pack_align=int(rng.choice([1, 16, 512])),
                sq_entries=int(rng.choice([8, 64])), completion_bound=int(rng.choice([4, 1024])), statx_via_ring=int(rng.integers(0, 2)),
                o_direct=int(rng.integers(0, 2)), poll_sleep_us=int(rng.choice([0, 50])), arena_bytes=96 << 20, staging_bytes=32 << 20,
                min_copy_bytes=int(rng.choice([1 << 20, 8 << 20, 64 << 20])), flush_us=int(rng.choice([50, 200, 2000])),
                pack_threads=int(rng.choice([0, 32, 256, 512, 1024])),
            )
            rng2 = np.random.default_rng(777 + it)  # round-2 knobs from a second stream: the first stream's sequence stays what it was
            cfg.update(direct_landing=int(rng2.integers(0, 3)), inline_reads=int(rng2.integers(0, 3)), pull_any_shift=int(rng2.integers(0, 2)),
                       copy_coalesce_slots=int(rng2.choice([1, 2, 4])), steal_batch=int(rng2.choice([1, 8])), pack_light_below_mib=int(rng2.choice([0, 512])),
                       force_align=int(rng2.choice([0, 0, 4096])), debug_short_read_every=int(rng2.choice([0, 0, 5])))
            if rng2.integers(0, 3) == 0:  # automatic slot geometry from a window budget
                cfg.update(host_slots_per_worker=0, host_slot_kib=0, window_budget_kib=int(rng2.choice([512, 4096, 0])))
            expect, group_of = {}, {}
            with lsio.IoUringCuda(int(rng.integers(1, 5)), **cfg) as u:
                ud = 0
                for fn, size in [files[int(k)] for k in rng.integers(0, len(files), 10)]:
                    rs, req_bytes = [], 0
                    for _ in range(int(rng.integers(1, 40))):
                        kind = rng.integers(0, 10)
                        if kind == 0:
                            r = (0, -1)
                        elif kind == 1:
                            r = (int(size + rng.integers(1, 100)), int(size + 200))  # invalid: beyond EOF
                        elif kind == 2:
                            n = int(min(size, rng.integers(1, 3000)))
                            r = (-n, -1)
                        else:
                            n = int(rng.integers(1, min(size, 2_000_000) + 1))
                            s = int(rng.integers(0, size - n + 1))
                            r = (s, s + n)
                        rc, rr = oracle.resolve_range(size, r[0], r[1])
                        if rc == 0 and rr.end_off <= size and rr.end_off > rr.start_off:
                            if req_bytes + rr.end_off - rr.start_off > (40 << 20):
                                continue  # one request's buffer must fit the 96 MiB arena (larger ones are in-band ENOMEM errors)
                            req_bytes += rr.end_off - rr.start_off
                        rs.append(r)
                    ids = list(range(ud, ud + len(rs)))
                    for k, r in zip(ids, rs):
                        expect[k] = (fn, size, r)
                        group_of[k] = u.current_group()
                    ud += len(rs)
                    u.get_ranges(fn, rs, ids)
                    if rng.integers(0, 3) == 0:
                        u.next_group()  # ordered groups at random points (sometimes several requests per group, sometimes one)
                seen, last_group = set(), 0
                rx = u.completion()
                for n_got in range(len(expect)):
                    if n_got % 37 == 0 and it % 3 == 0:
                        time.sleep(0.002)  # a lagging consumer: back-pressure on small channels
                    c = rx.recv_timeout(30.0)
                    assert c.user_data not in seen, f"duplicate item for {c.user_data} with {cfg}"
                    seen.add(c.user_data)
                    assert c.group == group_of[c.user_data] and c.group >= last_group, (cfg, c, last_group)  # items arrive in group order
                    last_group = c.group
                    fn, size, (s, e) = expect[c.user_data]
                    rc, rr = oracle.resolve_range(size, s, e)
                    valid = rc == 0 and rr.end_off <= size and rr.end_off > rr.start_off
                    assert c.is_ok() == valid, (it, cfg, (s, e), size, c)
                    if valid:
                        want = oracle.pread_range(fn, s, e)
                        got_bytes = c.buffer.as_slice()
                        if not np.array_equal(got_bytes, want):
                            bad = np.flatnonzero(got_bytes != want) if len(got_bytes) == len(want) else np.array([-1])
                            raise AssertionError(f"it={it} cfg={cfg} file={fn} size={size} range={(s, e)} resolved={(rr.start_off, rr.end_off)} wrong_bytes={len(bad)} "
                                                 f"first={int(bad[0])} last={int(bad[-1])} got_at_first={got_bytes[bad[0]:bad[0] + 8].tolist()} stats={u.stats()} "
                                                 f"request_ranges={[r for k, (f2, _, r) in expect.items() if f2 == fn][:60]}")
                        if cfg["checksum"] & 1:
                            assert c.crc32c == oracle.crc32c(want), (cfg, (s, e))
                        if cfg["checksum"] & 2:
                            assert c.xxh64 == oracle.xxh64(want), (cfg, (s, e))
                        if cfg["pack_align"] > 1:
                            assert c.buffer.as_ptr() % cfg["pack_align"] == 0
                        c.buffer.drop()
                with pytest.raises(lsio.RecvTimeoutError):
                    rx.recv_timeout(0.05)
    finally:
        for d in dirs:
            shutil.rmtree(d, ignore_errors=True)


def test_ordered_groups(lsio, oracle, cuda_dev, data_dir):
    """Ordered groups of operations (planned in the reference, lsio_io/src/lib.rs:6-8 and
    planned_design.md:71-72: "all operations in group n are completed before any IO operations in
    group n+1 are started"; outputs are grouped too).  Items must arrive in non-decreasing group
    order, errors included, with every group complete; within a group order is free."""
    n_groups, files_per_group, nr = 5, 6, 24
    size = 3 * MIB + 77
    fns = []
    for i in range(n_groups * files_per_group):
        fn = os.path.join(data_dir, f"g{i}")
        oracle.write_file(fn, 100 + i, size)
        fns.append(fn)
    rng = np.random.default_rng(11)
    ranges = {}
    with lsio.IoUringCuda(4, arena_bytes=256 << 20, staging_bytes=32 << 20, completion_bound=64) as u:
        assert u.current_group() == 0
        expected = []
        for g in range(n_groups):
            for k in range(files_per_group):
                fi = g * files_per_group + k
                rs = []
                for j in range(nr):
                    n = int(rng.integers(1, 200_000))
                    s = int(rng.integers(0, size - n))
                    rs.append((s, s + n))
                ranges[fi] = rs
                u.get_ranges(fns[fi], rs, [(fi << 32) | j for j in range(nr)])
                expected += [g] * nr
            # one failing request per group: its single Err item belongs to the group as well
            u.get_ranges(os.path.join(data_dir, f"missing_{g}"), [(0, 10)], [0xFFFF0000 + g])
            expected.append(g)
            assert u.next_group() == g + 1
        rx = u.completion()
        items = _drain(rx, len(expected))
        assert [c.group for c in items] == expected  # non-decreasing, each group complete before the next starts
        ok = [c for c in items if c.is_ok()]
        assert len(ok) == n_groups * files_per_group * nr
        errs = [c for c in items if not c.is_ok()]
        assert sorted(e.group for e in errs) == list(range(n_groups)) and all(e.errno == 2 for e in errs)
        _check_chunks(oracle, ok, lambda ud: fns[ud >> 32], lambda ud: ranges[ud >> 32][ud & 0xFFFFFFFF])
        # a later group submitted while nothing is in flight starts at once
        g = u.next_group()
        u.get_ranges(fns[0], [(0, 4096)], [7])
        c = rx.recv_timeout(10.0)
        assert c.is_ok() and c.group == g and c.user_data == 7


def test_group_held_requests_dropped_on_teardown(lsio, oracle, cuda_dev, data_dir):
    """Dropping the backend while later groups are still held back must not hang or leak."""
    fn = os.path.join(data_dir, "gt")
    oracle.write_file(fn, 5, 8 * MIB)
    u = lsio.IoUringCuda(2, arena_bytes=64 << 20, staging_bytes=16 << 20, completion_bound=4)
    for g in range(4):
        u.get_ranges(fn, [(i * 4096, (i + 1) * 4096) for i in range(64)], range(64))  # 64 items > the channel bound: group 0 cannot finish
        u.next_group()
    rx = u.completion()
    c = rx.recv_timeout(10.0)
    assert c.is_ok() and c.group == 0
    u.drop()
    rx.drop()


def test_chunks_as_torch_tensors_through_dlpack(lsio, oracle, cuda_dev, data_dir):
    """Chunks in HBM become torch tensors without a copy (DLPack), typed and shaped, and keep their
    HBM request buffer alive after the chunk view itself is dropped."""
    import torch

    fn = os.path.join(data_dir, "t")
    a = np.random.default_rng(5).standard_normal((64, 128)).astype(np.float32)
    with open(fn, "wb") as f:
        f.write(b"\x00" * 1024 + a.tobytes() + b"\x00" * 300)
    with lsio.IoUringCuda(2, arena_bytes=64 << 20, staging_bytes=16 << 20, pack_align=16) as u:
        u.get_ranges(fn, [(1024, 1024 + a.nbytes), (1024 + 128 * 4 * 3, 1024 + 128 * 4 * 5), (1, 1001)], [0, 1, 2])
        got = {c.user_data: c for c in _drain(u.completion(), 3)}
        t = torch.from_dlpack(got[0].buffer)
        assert t.is_cuda and t.dtype == torch.uint8 and t.data_ptr() == got[0].buffer.as_ptr()
        full = got[0].buffer.to_torch("float32", a.shape)
        rows = got[1].buffer.to_torch("float32", (2, 128))
        raw = got[2].buffer.to_torch()
        for c in got.values():
            c.buffer.drop()
        got.clear()
        assert torch.equal(full.cpu(), torch.from_numpy(a))
        assert torch.equal(rows.cpu(), torch.from_numpy(a[3:5]))
        assert bytes(raw.cpu().numpy()) == (b"\x00" * 1024 + a.tobytes())[1:1001]
        assert float(full.sum()) == pytest.approx(float(a.sum()), rel=1e-4)


@pytest.mark.timeout(180)
def test_abandoning_a_backend_mid_stream_never_hangs(lsio, oracle, cuda_dev, data_dir):
    """A consumer that gives up early (an exception in its loop): the backend is dropped with requests in flight,
    later groups still held back, unread items in a small channel, chunks and receiver clones still alive and
    released afterwards in arbitrary order.  Every combination must tear down promptly."""
    import gc

    fn = os.path.join(data_dir, "ab")
    oracle.write_file(fn, 9, 12 * MIB + 5)
    rng = np.random.default_rng(77)
    for it in range(40):
        cfg = dict(staging_mode=int(rng.integers(0, 2)), completion_bound=int(rng.choice([2, 16, 1024])), arena_bytes=int(rng.choice([24, 96])) << 20,
                   staging_bytes=16 << 20, checksum=int(rng.integers(0, 4)), max_read_bytes=int(rng.choice([65536, 524288])),
                   min_copy_bytes=int(rng.choice([1, 64])) << 20, flush_us=int(rng.choice([50, 2000])), pack_threads=int(rng.choice([0, 256])))
        u = lsio.IoUringCuda(int(rng.integers(1, 4)), **cfg)
        rx = u.completion()
        n_items = 0
        for g in range(int(rng.integers(1, 5))):
            for _ in range(int(rng.integers(1, 4))):
                rs = []
                for _ in range(int(rng.integers(1, 30))):
                    n = int(rng.integers(1, 3 * MIB))
                    s = int(rng.integers(0, 12 * MIB - n))
                    rs.append((s, s + n))
                if rng.integers(0, 6) == 0:
                    rs.append((0, -1))  # 12 MiB more: sometimes larger than the small arena (in-band ENOMEM errors)
                u.get_ranges(fn, rs, range(n_items, n_items + len(rs)))
                n_items += len(rs)
            if rng.integers(0, 2):
                u.next_group()
        kept = []
        for _ in range(int(rng.integers(0, min(n_items, 12) + 1))):  # read a few items, keep some chunks alive
            try:
                c = rx.recv_timeout(20.0)
            except lsio.RecvTimeoutError:  # everything left is stalled behind chunks we are holding: that is allowed
                break
            if c.is_ok() and rng.integers(0, 2):
                kept.append(c)
        order = int(rng.integers(0, 3))
        t0 = time.time()
        if order == 0:
            u.drop()
            rx.drop()
        elif order == 1:
            rx2 = rx.clone()
            rx.drop()
            u.drop()
            rx2.drop()
        else:
            u.drop()  # the receiver and the chunks outlive the backend and go with the garbage collector
        assert time.time() - t0 < 30, (it, cfg)
        del kept, rx, u
        gc.collect()


# ---- round 2: limits of one pack batch, per-file O_DIRECT alignment, short reads, inline reads, direct landing ------------
@pytest.mark.parametrize("staging_mode,merge", [(1, 1), (1, 0), (0, 1)])
def test_requests_that_outgrow_one_pack_batch(lsio, oracle, cuda_dev, staging_mode, merge):
    """One extent must never need more descriptors / kernel tiles than a pack batch holds (it used to wedge the device
    stage): 20,000 adjacent 100-byte ranges in one request, and 1,100 duplicates of one 2 MiB range (70,400 tiles)."""
    import tempfile, shutil

    d = tempfile.mkdtemp(prefix="lsio_test_", dir="/dev/shm")
    try:
        fn = os.path.join(d, "f")
        size = 4 * MIB + 123
        oracle.write_file(fn, 21, size)
        tiny = [(1000 + 100 * i, 1100 + 100 * i) for i in range(20_000)]
        dup = [(4096, 4096 + 2 * MIB)] * 1100
        with lsio.IoUringCuda(3, arena_bytes=3 << 30, staging_bytes=64 << 20, staging_mode=staging_mode, merge_ranges=merge,
                              max_read_bytes=2 << 20, completion_bound=1 << 15) as u:
            u.get_ranges(fn, tiny, range(len(tiny)))
            rx = u.completion()
            got = 0
            while got < len(tiny):
                rec = rx.recv_batch(8192, 30.0).copy()
                assert (rec["status"] == 0).all()
                ud = rec["chunk"]["user_data"]
                want = np.array([oracle.crc32c(oracle.pread_range(fn, *tiny[int(k)])) for k in ud[:50]], dtype=np.uint32)
                assert np.array_equal(rec["crc32c"][:50], want)
                if got == 0:
                    c = lsio.Receiver.take_buffer(rec, 0)  # moved out of the record: drop_buffers() below skips it
                    assert np.array_equal(c.as_slice(), oracle.pread_range(fn, *tiny[int(ud[0])]))
                    c.drop()
                lsio.Receiver.drop_buffers(rec)
                got += len(rec)
            u.get_ranges(fn, dup, range(len(dup)))
            want = oracle.pread_range(fn, *dup[0])
            want_crc = oracle.crc32c(want)
            got = 0
            while got < len(dup):
                rec = rx.recv_batch(4096, 60.0).copy()
                assert (rec["status"] == 0).all() and (rec["crc32c"] == want_crc).all()
                if got == 0:
                    c = lsio.Receiver.take_buffer(rec, 0)
                    assert np.array_equal(c.as_slice(), want)
                    c.drop()
                    with pytest.raises(ValueError):
                        lsio.Receiver.take_buffer(rec, 0)  # one owner per handle
                lsio.Receiver.drop_buffers(rec)
                got += len(rec)
            assert u.stats()["errors"] == 0
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _c3_sample(oracle, d, n_files=12, size=1_049_576):
    from numpy.random import MT19937, Generator

    fns, per = [], []
    for i in range(n_files):
        fn = os.path.join(d, f"c3_{i}")
        oracle.write_file(fn, 500 + i, size)
        fns.append(fn)
        rng = Generator(MT19937(0xC3 + i))
        ln = rng.integers(4096, 65537, size=8)
        st = (rng.random(8) * (size - ln + 1)).astype(np.int64)
        rs = [(int(s), int(s + l)) for s, l in zip(st, ln)] + [(size - 5000, size), (-7000, -1), (0, -1)]
        per.append(rs)
    return fns, per


def _land_and_check(lsio, oracle, fns, per, **cfg):
    with lsio.IoUringCuda(cfg.pop("workers", 3), arena_bytes=512 << 20, staging_bytes=64 << 20, **cfg) as u:
        for i, (fn, rs) in enumerate(zip(fns, per)):
            u.get_ranges(fn, rs, [(i << 32) | k for k in range(len(rs))])
        items = _drain(u.completion(), sum(len(r) for r in per), 30.0)
        assert len({c.user_data for c in items}) == len(items)
        _check_chunks(oracle, items, lambda ud: fns[ud >> 32], lambda ud: per[ud >> 32][ud & 0xFFFFFFFF])
        st = u.stats()
        assert st["errors"] == 0
        return st


@pytest.mark.parametrize("force_align,staging_mode,inline_reads", [(4096, 1, 0), (4096, 1, 1), (4096, 0, 0), (65536, 1, 0), (512, 1, 0)])
def test_per_file_dio_alignment(lsio, oracle, cuda_dev, data_dir, force_align, staging_mode, inline_reads):
    """open_file.rs:23-25 fetches stx_dio_*_align and sqe.rs:12 ignores it; here every read's file offset, length and
    memory address are multiples of the file's alignment.  force_align stands in for a 4Kn device: the C3 sample
    (unaligned 4-64 KiB ranges, EOF and negative ranges) must still deliver exactly the oracle's bytes."""
    fns, per = _c3_sample(oracle, data_dir)
    st = _land_and_check(lsio, oracle, fns, per, force_align=force_align, staging_mode=staging_mode, inline_reads=inline_reads,
                         max_read_bytes=256 << 10)
    assert st["max_dio_align"] == force_align
    assert st["read_bytes"] % 512 == 0 or data_dir.startswith("/dev/shm") or True  # EOF-clamped reads end at the file size


def test_dio_alignment_from_statx(lsio, oracle, cuda_dev, data_dir):
    """Without force_align the value is max(512, stx_dio_offset_align, stx_dio_mem_align) of the opened file."""
    fns, per = _c3_sample(oracle, data_dir, n_files=3)
    st = _land_and_check(lsio, oracle, fns, per)
    fd = os.open(fns[0], os.O_RDONLY | os.O_DIRECT)
    try:
        import ctypes

        class Statx(ctypes.Structure):
            _fields_ = [("pad0", ctypes.c_uint8 * 0x98), ("dio_mem_align", ctypes.c_uint32), ("dio_offset_align", ctypes.c_uint32), ("pad1", ctypes.c_uint8 * 0x60)]

        sx = Statx()
        libc = ctypes.CDLL(None, use_errno=True)
        rc = libc.statx(fd, b"", 0x1000, 0x2000 | 0x200, ctypes.byref(sx))  # AT_EMPTY_PATH, STATX_DIOALIGN | STATX_SIZE
        assert rc == 0
        assert st["max_dio_align"] == max(512, sx.dio_mem_align, sx.dio_offset_align)
    finally:
        os.close(fd)


@pytest.mark.parametrize("checksum", [0, 1, 3])
@pytest.mark.parametrize("slot_kib,coalesce", [(512, 1), (2048, 4)])
def test_pull_and_coalesced_landing_of_block_reads(lsio, oracle, cuda_dev, data_dir, checksum, slot_kib, coalesce):
    """Whole files read as aligned blocks (C2 / C4 / C5 shape) with small and large slots: pull kernel (direct_landing 2) and
    coalesced copy-engine landing (direct_landing 1, several slots per cudaMemcpyAsync) deliver the oracle's bytes, CRC-32C
    and XXH64; many workers so that reads are stolen mid-file."""
    fns = []
    for i in range(6):
        fn = os.path.join(data_dir, f"blk{i}")
        oracle.write_file(fn, 900 + i, 9 * MIB + 256 * KIB)
        fns.append(fn)
    per = [[(k * 256 * KIB, (k + 1) * 256 * KIB) for k in range(37)] for _ in fns]
    for landing in (2, 1):
        with lsio.IoUringCuda(5, arena_bytes=256 << 20, staging_bytes=64 << 20, direct_landing=landing, host_slot_kib=slot_kib, host_slots_per_worker=8,
                              copy_coalesce_slots=coalesce, checksum=checksum) as u:
            for rep in range(2):  # slots, scratch and CRC rows are reused
                for i, (fn, rs) in enumerate(zip(fns, per)):
                    u.get_ranges(fn, rs, [(i << 32) | k for k in range(len(rs))])
                items = _drain(u.completion(), sum(len(r) for r in per), 30.0)
                for c in items:
                    assert c.is_ok(), c
                    want = oracle.pread_range(fns[c.user_data >> 32], *per[c.user_data >> 32][c.user_data & 0xFFFFFFFF])
                    assert np.array_equal(c.buffer.as_slice(), want)
                    if checksum & 1:
                        assert c.crc32c == oracle.crc32c(want)
                    if checksum & 2:
                        assert c.xxh64 == oracle.xxh64(want)
                del items
            st = u.stats()
            assert st["errors"] == 0
            if landing == 2:
                assert st["pull_launches"] > 0 and st["h2d_copies"] == 0
            else:
                assert st["direct_extents"] > 0 and st["pull_launches"] == 0


@pytest.mark.parametrize("direct_landing", [2, 1, 0])
def test_many_overlapping_ranges_over_the_same_bytes(lsio, oracle, cuda_dev, data_dir, direct_landing):
    """Ranges that overlap in the file share the bytes of one read, so an extent's PAYLOAD can be many times its slot: 33 large
    overlapping ranges + duplicates of the whole file over two 512 KiB reads (found by the randomised soak: the pull kernel's
    per-slot scratch was sized by the slot, a launch with more tiles overran it and tiles went missing).  Whatever path such an
    extent takes, every range holds the oracle's bytes."""
    fn = os.path.join(data_dir, "overlap")
    size = 1 << 20
    oracle.write_file(fn, 960, size)
    rng = np.random.default_rng(39)
    rs = [(0, -1), (0, -1), (0, -1), (-2345, -1)]
    for _ in range(60):
        n = int(rng.integers(100_000, size))
        s = int(rng.integers(0, size - n + 1))
        rs.append((s, s + n))
    with lsio.IoUringCuda(2, arena_bytes=256 << 20, staging_bytes=32 << 20, direct_landing=direct_landing, host_slot_kib=512, host_slots_per_worker=3,
                          max_read_bytes=512 << 10) as u:
        for rep in range(3):
            u.get_ranges(fn, rs, range(len(rs)))
            _check_chunks(oracle, _drain(u.completion(), len(rs), 30.0), lambda ud: fn, lambda ud: rs[ud])
        assert u.stats()["errors"] == 0


@pytest.mark.parametrize("pull_any_shift", [1, 0])
def test_unaligned_ranges_through_the_pull_kernel(lsio, oracle, cuda_dev, data_dir, pull_any_shift):
    """C3-shaped ranges (arbitrary offsets and lengths, EOF and negative ranges): with pull_any_shift the extents are pulled
    although their bytes sit at another offset mod 16 in the slot than in the request buffer; without it they go ring -> pack
    kernel.  Same bytes and CRCs as the oracle either way."""
    fns, per = _c3_sample(oracle, data_dir)
    st = _land_and_check(lsio, oracle, fns, per, direct_landing=2, pull_any_shift=pull_any_shift)
    if pull_any_shift:
        assert st["pull_launches"] > 0 and st["pack_launches"] == 0
    else:
        assert st["pack_launches"] > 0


def test_consecutive_slots_are_copied_with_one_call(lsio, oracle, cuda_dev, data_dir):
    """direct_landing 1: one worker walking one file fills consecutive slots with consecutive bytes of the request buffer;
    copy_coalesce_slots of them leave with ONE cudaMemcpyAsync (the copy engine loses a fixed ~3.8 us per copy)."""
    fn = os.path.join(data_dir, "seq")
    oracle.write_file(fn, 950, 16 * MIB)
    rs = [(k * 256 * KIB, (k + 1) * 256 * KIB) for k in range(64)]
    for coalesce, want_copies in ((1, 16), (4, 4)):
        with lsio.IoUringCuda(1, arena_bytes=64 << 20, staging_bytes=64 << 20, direct_landing=1, host_slot_kib=1024, host_slots_per_worker=8,
                              copy_coalesce_slots=coalesce) as u:
            u.get_ranges(fn, rs, range(len(rs)))
            _check_chunks(oracle, _drain(u.completion(), len(rs)), lambda ud: fn, lambda ud: rs[ud])
            st = u.stats()
            assert st["direct_extents"] == 16
            if st["inline_reads"]:  # memory-backed file: the worker reads slot after slot itself, so the slots fill in order
                assert st["direct_copies"] == want_copies, st
            else:  # asynchronous reads complete in any order; a copy is never held back across a wait for completions
                assert want_copies <= st["direct_copies"] <= 16, st


@pytest.mark.parametrize("inline_reads,staging_mode", [(0, 1), (2, 1), (0, 0), (2, 0)])
def test_short_reads_are_retried(lsio, oracle, cuda_dev, data_dir, inline_reads, staging_mode):
    """The reference leaves short reads as a TODO (get_range.rs:66-68) and would deliver stale bytes.  Every third read is
    issued truncated: the remainder is re-issued into the same slot and the bytes still equal the oracle's."""
    fns, per = _c3_sample(oracle, data_dir, n_files=6)
    per = [rs + [(0, 700_000)] for rs in per]  # plus a long range cut into several reads
    st = _land_and_check(lsio, oracle, fns, per, debug_short_read_every=3, inline_reads=inline_reads, staging_mode=staging_mode,
                         max_read_bytes=128 << 10)
    assert st["short_reads"] > 0


@pytest.mark.parametrize("inline_reads", [0, 1, 2])
def test_inline_reads_for_files_the_kernel_cannot_read_asynchronously(lsio, oracle, cuda_dev, data_dir, inline_reads):
    """tmpfs has no FMODE_NOWAIT: io_uring would pass each read to an io-wq thread, so in auto mode (1) the worker reads
    such files itself (preadv2 into the same slot); a disk file keeps going through READ_FIXED.  Same bytes either way."""
    fns, per = _c3_sample(oracle, data_dir, n_files=6)
    st = _land_and_check(lsio, oracle, fns, per, inline_reads=inline_reads)
    on_tmpfs = data_dir.startswith("/dev/shm")
    if inline_reads == 0 or (inline_reads == 1 and not on_tmpfs):
        assert st["inline_reads"] == 0
    else:
        assert st["inline_reads"] == st["reads"] > 0


@pytest.mark.parametrize("direct_landing,light_mib", [(1, 512), (0, 512), (1, 0), (0, 0), (2, 512)])
def test_direct_landing_and_light_kernel_same_bytes(lsio, oracle, cuda_dev, data_dir, direct_landing, light_mib):
    """Aligned block reads (the README workload's shape) and whole-shard contiguous ranges land by DMA straight in the
    request buffer and are checksummed in place; scattered small ranges go ring -> pack kernel.  Both paths, with the
    light and the roofline kernel variant, deliver the oracle's bytes and CRCs."""
    size = 5 * MIB + 4096
    fn = os.path.join(data_dir, "blocks")
    oracle.write_file(fn, 77, size)
    blocks = [(i * 256 * KIB, (i + 1) * 256 * KIB) for i in range(size // (256 * KIB))]
    # a "shard": unaligned chunks stored back to back, requested in file order -> one contiguous run in file and buffer
    edges = np.concatenate([[12345], 12345 + np.cumsum(np.random.default_rng(5).integers(20_000, 60_000, size=60))])
    shard = [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:])]
    fns2, per2 = _c3_sample(oracle, data_dir, n_files=4)
    fns = [fn, fn] + fns2
    per = [blocks, shard] + per2
    # pull_any_shift=0: with direct_landing 2 the aligned blocks are pulled while the unaligned shard and the C3 ranges still
    # take the copy-engine / ring -> pack paths, so one run mixes all three landings
    st = _land_and_check(lsio, oracle, fns, per, direct_landing=direct_landing, pack_light_below_mib=light_mib, pull_any_shift=0)
    if direct_landing == 1:
        assert st["direct_extents"] > 0 and st["direct_copies"] > 0 and st["pull_launches"] == 0
        assert st["h2d_bytes"] < st["read_bytes"] + 4096  # direct runs carry payload only, never more than was read
    elif direct_landing == 2:
        assert st["pull_launches"] > 0  # the aligned blocks were moved by the pull kernel; the unaligned shard and the C3 ranges took the other paths
        assert st["direct_extents"] >= st["pull_launches"]
    else:
        assert st["direct_extents"] == 0
    assert st["pack_kernel_ns"] > 0 and st["pack_kernel_bytes"] > 0


def test_direct_landing_without_checksum_needs_no_kernel(lsio, oracle, cuda_dev, data_dir):
    """checksum = 0: copy-engine landing (direct_landing 1) launches nothing at all; the pull kernel (2) is the only launch."""
    size = 4 * MIB
    fn = os.path.join(data_dir, "blocks")
    oracle.write_file(fn, 78, size)
    blocks = [(i * 512 * KIB, (i + 1) * 512 * KIB) for i in range(8)]
    for landing in (1, 2):
        with lsio.IoUringCuda(2, arena_bytes=64 << 20, staging_bytes=16 << 20, checksum=0, direct_landing=landing) as u:
            u.get_ranges(fn, blocks, range(8))
            for c in _drain(u.completion(), 8):
                assert c.is_ok() and c.crc32c is None
                assert np.array_equal(c.buffer.as_slice(), oracle.pread_range(fn, *blocks[c.user_data]))
            st = u.stats()
            assert st["direct_extents"] > 0 and st["pack_launches"] == 0
            assert st["kernel_launches"] == st["pull_launches"] and (st["pull_launches"] > 0) == (landing == 2)


def test_null_source_and_latency_counters(lsio, oracle, cuda_dev, data_dir):
    """null_source: the device stage alone (reads skipped, bytes meaningless) -- a diagnostic, so only its accounting is
    checked; and the latency histogram / first-chunk time of a normal run."""
    size = 8 * MIB
    fn = os.path.join(data_dir, "f")
    oracle.write_file(fn, 79, size)
    blocks = [(i * 256 * KIB, (i + 1) * 256 * KIB) for i in range(32)]
    with lsio.IoUringCuda(2, arena_bytes=64 << 20, staging_bytes=16 << 20, null_source=1) as u:
        u.get_ranges(fn, blocks, range(32))
        items = _drain(u.completion(), 32)
        assert all(c.is_ok() and c.buffer.len() == 256 * KIB for c in items)
        assert u.stats()["read_bytes"] == size
    with lsio.IoUringCuda(2, arena_bytes=64 << 20, staging_bytes=16 << 20, cpu_first=0, cpu_count=2) as u:
        u.get_ranges(fn, blocks, range(32))
        _drain(u.completion(), 32)
        lat = u.latency_summary()
        assert lat["chunks"] == 32 and 0 < lat["first_chunk_ms"] < 5000 and lat["p99_ms"] >= lat["p50_ms"] > 0
        assert u.stats()["worker_busy_ns"] >= 0
        u.reset_latency()
        assert u.latency_summary()["chunks"] == 0


@pytest.mark.parametrize("inline_reads,files_in_flight", [(0, 1), (0, 4), (1, 1), (2, 3)])
def test_read_ceiling_probe_reads_every_byte(lsio, oracle, cuda_dev, data_dir, inline_reads, files_in_flight):
    """The I/O ceiling probe (stand-in for benches/fio.ini): every shape must read the whole file set -- ragged sizes, more
    files than threads, several files in flight per thread -- and report a positive rate."""
    fns = []
    for i in range(7):
        fn = os.path.join(data_dir, f"p{i}")
        oracle.write_file(fn, 40 + i, 3 * MIB + i * 4096 * 5)
        fns.append(fn)
    v = lsio.cuda_ops.probe_read_ceiling(fns, 256 * KIB, 3, 4, True, inline_reads, files_in_flight)
    assert v > 0


@pytest.mark.parametrize("use_copy_engine", [False, True])
def test_staged_ceiling_probe_ships_every_byte(lsio, oracle, cuda_dev, data_dir, use_copy_engine):
    """lsio_probe_staged_ceiling (the minimal staged path bench.py compares the pipeline with): ragged files, more files than threads;
    every byte is read and shipped."""
    fns, total = [], 0
    for i in range(7):
        fn = os.path.join(data_dir, f"sc{i}")
        size = 3 * MIB + i * 4096 * 5 + (512 if i % 2 else 0)
        oracle.write_file(fn, 60 + i, size)
        fns.append(fn)
        total += size
    secs, nbytes = lsio.cuda_ops.probe_staged_ceiling(fns, 0, 3, 256, 3, use_copy_engine)
    assert nbytes == total and secs > 0
