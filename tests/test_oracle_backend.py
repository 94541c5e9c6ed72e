"""The CPU port of the reference backend (oracle/lsio_uring_ref.c) against the reference's own
end-to-end contract, crates/lsio_uring/tests/integration_test.rs:13-120 (CPU only).
"""
import os
import random

import numpy as np
import pytest

KIB, MIB = 1024, 1024 * 1024


def _uring_available():
    import ctypes

    libc = ctypes.CDLL(None, use_errno=True)
    buf = ctypes.create_string_buffer(256)
    fd = libc.syscall(425, 4, buf)  # io_uring_setup
    if fd < 0:
        return False
    os.close(fd)
    return True


pytestmark = pytest.mark.skipif(not _uring_available(), reason="io_uring not available here")


@pytest.mark.parametrize("sqpoll", [True, False])
@pytest.mark.parametrize("mode", [0, 1])
def test_get_ranges(oracle, data_dir, sqpoll, mode):
    """integration_test.rs::test_get_ranges, line for line."""
    N_WORKER_THREADS, FILE_SIZE, CHUNK_SIZE = 4, MIB, 4 * KIB
    N_CHUNKS = FILE_SIZE // CHUNK_SIZE
    period = int(CHUNK_SIZE * 1.5)  # integration_test.rs:26
    unit = bytes(random.randrange(32, 127) for _ in range(period))
    file_contents = (unit * (FILE_SIZE // period + 1))[:FILE_SIZE]
    filename = os.path.join(data_dir, f"lsio_uring_tempfile_{random.getrandbits(32)}")
    with open(filename, "wb") as f:
        f.write(file_contents)
        f.flush()
        os.fsync(f.fileno())
    ranges = [(i * CHUNK_SIZE, (i + 1) * CHUNK_SIZE) for i in range(N_CHUNKS)]
    uring = oracle.RefIoUring(N_WORKER_THREADS, sqpoll=sqpoll, mode=mode)
    uring.get_ranges(filename, ranges, list(range(N_CHUNKS)))
    chunks = [None] * N_CHUNKS
    for i in range(N_CHUNKS):
        c = uring.recv_timeout(0.5)
        assert c.status == 0, f"Error reading chunk {i}! {c.status}"
        chunks[c.user_data] = c.data
    uring.shutdown()  # drop(uring)
    with pytest.raises(oracle.Disconnected):
        uring.recv_timeout(0.5)  # a cloned receiver sees the disconnect (integration_test.rs:92-96)
    assert b"".join(x.tobytes() for x in chunks) == file_contents
    uring.close()


def test_unaligned_negative_and_eof_ranges_are_file_bytes(oracle, data_dir):
    """Not covered by any reference test: ground truth is pread of [start_off, end_off)."""
    size = MIB + 1000
    fn = os.path.join(data_dir, "f")
    oracle.write_file(fn, 3, size)
    rng = np.random.default_rng(1)
    ranges = [(0, -1), (-100, -1), (size - 5, size), (0, 1), (511, 513), (700, 1300)]
    for _ in range(64):
        n = int(rng.integers(1, 65536))
        s = int(rng.integers(0, size - n))
        ranges.append((s, s + n))
    u = oracle.RefIoUring(2, sqpoll=False)
    u.get_ranges(fn, ranges, list(range(len(ranges))))
    for _ in ranges:
        c = u.recv_timeout(5)
        assert c.status == 0
        s, e = ranges[c.user_data]
        want = oracle.pread_range(fn, s, e)
        assert np.array_equal(c.data, want), ranges[c.user_data]
        rc, r = oracle.resolve_range(size, s, e)
        assert np.array_equal(want, oracle.pattern(3, r.start_off, r.end_off - r.start_off))
    u.close()


def test_missing_file_yields_two_errors(oracle):
    """openat and statx both fail => two Err items (the doc comment says one)."""
    u = oracle.RefIoUring(2, sqpoll=False)
    u.get_ranges("/definitely/not/here", [(0, 10), (10, 20)], [1, 2])
    errs = [u.recv_timeout(2), u.recv_timeout(2)]
    assert sorted((e.status, e.opcode) for e in errs) == [(-2, 18), (-2, 21)]  # ENOENT from OPENAT(18) and STATX(21)
    with pytest.raises(oracle.Timeout):
        u.recv_timeout(0.2)
    u.close()


def test_bench_driver_counts_bytes(oracle, data_dir):
    """lsio_bench::read_files restated: n_chunks = filesize / blocksize drops the tail."""
    fsize, bs = 1_000_000, 262_144
    paths = []
    for i in range(3):
        p = os.path.join(data_dir, f"lsio_bench_{i}")
        oracle.write_file(p, i, fsize, ref_pattern=True)
        paths.append(p)
    n_chunks = fsize // bs
    st = np.arange(n_chunks, dtype=np.int64) * bs
    secs, nbytes, errors = oracle.bench_read_files(paths, st, st + bs, 4, sqpoll=False)
    assert errors == 0 and nbytes == 3 * n_chunks * bs and secs > 0
