"""The oracle against golden vectors and the reference's own unit tests (CPU only).

Pins oracle/lsio_oracle.c before it is trusted as the checker for the CUDA path.
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _vectors():
    return json.load(open(os.path.join(HERE, "golden", "checksums.json")))


def _data(v):
    if "hex" in v:
        return bytes.fromhex(v["hex"])
    assert v["gen"] == "i_as_u8"
    return bytes(i & 0xFF for i in range(v["len"]))


def test_crc32c_and_xxh64_golden(oracle):
    for v in _vectors():
        d = _data(v)
        assert oracle.crc32c(d) == v["crc32c"], v["name"]
        assert oracle.crc32c_bitwise(d) == v["crc32c"], v["name"]
        assert oracle.crc32c_table(d) == v["crc32c"], v["name"]
        assert oracle.xxh64(d) == v["xxh64"], v["name"]
        if "xxh64_seed7" in v:
            assert oracle.xxh64(d, 7) == v["xxh64_seed7"], v["name"]


def test_crc32c_rfc3720_known_answers(oracle):
    # RFC 3720 B.4 and SURVEY appendix C
    assert oracle.crc32c(b"") == 0
    assert oracle.crc32c(b"123456789") == 0xE3069283
    assert oracle.crc32c(bytes(32)) == 0x8A9136AA
    assert oracle.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert oracle.crc32c(bytes(range(32))) == 0x46DD794E
    assert oracle.xxh64(b"") == 0xEF46DB3751D8E999
    assert oracle.xxh64(b"123456789") == 0x8CB841DB40E6AE83


def test_crc32c_streaming_and_combine(oracle):
    rng = np.random.default_rng(0)
    for la, lb in [(0, 0), (0, 5), (5, 0), (1, 1), (100, 3), (4096, 4097), (70000, 123457)]:
        a = rng.integers(0, 256, la, dtype=np.uint8)
        b = rng.integers(0, 256, lb, dtype=np.uint8)
        whole = oracle.crc32c(np.concatenate([a, b]))
        assert oracle.crc32c(b, oracle.crc32c(a)) == whole
        assert oracle.crc32c_combine(oracle.crc32c(a), oracle.crc32c(b), lb) == whole


def test_pattern_golden(oracle):
    for v in json.load(open(os.path.join(HERE, "golden", "pattern.json"))):
        want = bytes.fromhex(v["hex"])
        got = oracle.pattern(v["file_id"], v["off"], len(want)).tobytes()
        assert got == want, v
    # position-keyed: any window equals the same window of a longer fill
    whole = oracle.pattern(42, 0, 5000)
    for off, n in [(0, 1), (1, 7), (7, 9), (513, 1000), (4093, 907)]:
        assert np.array_equal(oracle.pattern(42, off, n), whole[off:off + n])
    # lsio_bench's own pattern: byte[i] = i as u8 (main.rs:102)
    assert oracle.ref_pattern(250, 10).tolist() == [250, 251, 252, 253, 254, 255, 0, 1, 2, 3]


# ---- range arithmetic: crates/lsio_uring/src/sqe.rs:59-122 -------------------------------------
@pytest.mark.parametrize(
    "size,start,end,want",
    [
        # (start_off, end_off, aligned_start, buf_len, view_start, view_end)
        (1 << 20, 0, 4096, (0, 4096, 0, 4096, 0, 4096)),            # integration_test.rs chunk 0
        (1 << 20, 8192, 12288, (8192, 12288, 8192, 4096, 0, 4096)),
        (1000, 0, -1, (0, 1000, 0, 1000, 0, 1000)),                 # "0..-1 the entire file" (lsio_io lib.rs:20)
        (1000, 0, 100, (0, 100, 0, 100, 0, 100)),                   # "0..100 the first 100 bytes"
        (1000, -100, -1, (900, 1000, 512, 488, 388, 488)),          # "-100..-1 the last 100 bytes"
        (5000, 700, 1300, (700, 1300, 512, 788, 188, 788)),         # unaligned start: the evident intent
        (5000, 511, 513, (511, 513, 0, 513, 511, 513)),
    ],
)
def test_resolve_range(oracle, size, start, end, want):
    rc, r = oracle.resolve_range(size, start, end)
    assert rc == 0
    assert (r.start_off, r.end_off, r.aligned_start, r.buf_len, r.view_start, r.view_end) == want
    assert r.padded_len == (r.buf_len + 511) // 512 * 512  # lsio_aligned_bytes lib.rs:178-181


def test_resolve_range_panics_of_the_reference(oracle):
    assert oracle.resolve_range(1000, -2000, 10)[0] == -1          # assert!(start_offset >= 0), sqe.rs:71
    assert oracle.resolve_range(1000, 0, -2000)[0] == -2           # assert!(end_offset >= 0), sqe.rs:81
    assert oracle.resolve_range(1000, 600, 512)[0] == -3           # assert!(buf_len > 0), sqe.rs:88
    assert oracle.resolve_range(1 << 40, 0, 2_147_479_553)[0] == -4  # get_range.rs:23-30
    assert oracle.resolve_range(1 << 40, 0, 2_147_479_552)[0] == 0
    # F2: an unaligned start makes split_to() fail and the unwrap panics (sqe.rs:109-113)
    assert oracle.resolve_range(5000, 700, 1300, strict=True)[0] == -5
    assert oracle.resolve_range(5000, 1024, 1300, strict=True)[0] == 0


# ---- user_data: crates/lsio_uring/src/user_data.rs:47-56 ---------------------------------------
def test_uring_user_data_round_trip(oracle):
    import ctypes

    READ = 22  # io_uring::opcode::Read::CODE == IORING_OP_READ
    v = oracle.L.lsio_oracle_user_data_pack(100, READ)
    assert v == (100 << 32) | READ
    idx, op = ctypes.c_uint32(), ctypes.c_uint8()
    oracle.L.lsio_oracle_user_data_unpack(v, ctypes.byref(idx), ctypes.byref(op))
    assert (idx.value, op.value) == (100, READ)


# ---- tracker: crates/lsio_uring/src/tracker.rs:74-114 ------------------------------------------
def test_op_tracker(oracle):
    t = oracle.Tracker(2)
    assert t.get(0) is None
    i0 = t.get_next_index()
    assert i0 == 0
    t.put(i0, 1000)
    assert t.remove(i0) == 1000
    i1 = t.get_next_index()
    assert i1 == 1
    t.put(i1, 1001)
    i2 = t.get_next_index()
    assert i2 == 0  # FIFO free list: the freed index comes back after the untouched one
    t.put(i2, 1002)
    assert t.get_next_index() is None and t.is_full()
    assert t.remove(i1) == 1001
    assert t.remove(i2) == 1002
    assert t.is_empty()


def test_tracker_panic_if_wrong_index(oracle):
    with pytest.raises(IndexError, match="index out of bounds"):
        oracle.Tracker(2).get(100)


# ---- aligned bytes model: lib.rs:74-95, :178-181 -----------------------------------------------
def test_split_to_rules_and_padding(oracle):
    L = oracle.L
    MiB = 1 << 20
    assert L.lsio_oracle_split_to_check(0, 8 * MiB, 2 * MiB, 512) == 0
    assert L.lsio_oracle_split_to_check(2 * MiB, 8 * MiB, MiB, 512) == -1   # not contained
    assert L.lsio_oracle_split_to_check(0, 8 * MiB, 8 * MiB, 512) == -1    # end is exclusive
    assert L.lsio_oracle_split_to_check(0, 8 * MiB, 0, 512) == -2          # zero
    assert L.lsio_oracle_split_to_check(0, 8 * MiB, 100, 512) == -3        # unaligned (this is F2)
    assert L.lsio_oracle_padded_size(16, 8) == 16
    assert L.lsio_oracle_padded_size(5000, 512) == 5120
    assert L.lsio_oracle_padded_size(1, 512) == 512


def test_pack_restatement(oracle):
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, 4096, dtype=np.uint8)
    so, do, ln = [100, 1024 + 7, 3000], [0, 50, 50 + 333], [50, 333, 1000]
    dst, crc, xx = oracle.pack(src, so, do, ln, 2048, want_xxh=True)
    for s, d, n, c, x in zip(so, do, ln, crc, xx):
        assert np.array_equal(dst[d:d + n], src[s:s + n])
        assert c == oracle.crc32c(src[s:s + n]) and x == oracle.xxh64(src[s:s + n])
