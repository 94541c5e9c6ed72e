"""ctypes wrapper for the CPU ORACLE (oracle/_build/liblsio_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; nothing under light-speed-io_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liblsio_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("lsio_oracle.c", "lsio_uring_ref.c", "lsio_codecs_oracle.c", "lsio_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        build()
    L = ctypes.CDLL(LIB_PATH)
    return L


L = _load()


class Range(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("start_off", "end_off", "aligned_start", "buf_len", "padded_len", "view_start", "view_end")]


class RefOutput(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("opcode", ctypes.c_uint8), ("user_data", ctypes.c_uint64),
                ("ptr", ctypes.c_void_p), ("len", ctypes.c_uint64), ("base", ctypes.c_void_p)]


_u8p = ctypes.POINTER(ctypes.c_uint8)
L.lsio_oracle_resolve_range.restype = ctypes.c_int
L.lsio_oracle_resolve_range.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(Range)]
L.lsio_oracle_user_data_pack.restype = ctypes.c_uint64
L.lsio_oracle_user_data_pack.argtypes = [ctypes.c_uint32, ctypes.c_uint8]
L.lsio_oracle_user_data_unpack.argtypes = [ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint8)]
L.lsio_oracle_tracker_new.restype = ctypes.c_void_p
L.lsio_oracle_tracker_new.argtypes = [ctypes.c_size_t]
L.lsio_oracle_tracker_free.argtypes = [ctypes.c_void_p]
L.lsio_oracle_tracker_get_next_index.restype = ctypes.c_int64
L.lsio_oracle_tracker_get_next_index.argtypes = [ctypes.c_void_p]
L.lsio_oracle_tracker_put.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64]
L.lsio_oracle_tracker_get.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
L.lsio_oracle_tracker_remove.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
L.lsio_oracle_tracker_is_empty.argtypes = [ctypes.c_void_p]
L.lsio_oracle_tracker_is_full.argtypes = [ctypes.c_void_p]
L.lsio_oracle_padded_size.restype = ctypes.c_size_t
L.lsio_oracle_padded_size.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
L.lsio_oracle_split_to_check.argtypes = [ctypes.c_size_t] * 4
L.lsio_oracle_pattern_fill.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
L.lsio_oracle_ref_pattern_fill.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
L.lsio_oracle_write_file.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int]
L.lsio_oracle_pread_range.restype = ctypes.c_int64
L.lsio_oracle_pread_range.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t]
for _n in ("lsio_oracle_crc32c_bitwise", "lsio_oracle_crc32c", "lsio_oracle_crc32c_table"):
    getattr(L, _n).restype = ctypes.c_uint32
    getattr(L, _n).argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t]
L.lsio_oracle_crc32c_combine.restype = ctypes.c_uint32
L.lsio_oracle_crc32c_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
L.lsio_oracle_crc32c_file_blocks.restype = ctypes.c_int64
L.lsio_oracle_crc32c_file_blocks.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
L.lsio_oracle_xxh64.restype = ctypes.c_uint64
L.lsio_oracle_xxh64.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64]
L.lsio_oracle_pack.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
L.lsio_ref_new.restype = ctypes.c_void_p
L.lsio_ref_new.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
L.lsio_ref_get_ranges.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
L.lsio_ref_recv.argtypes = [ctypes.c_void_p, ctypes.POINTER(RefOutput), ctypes.c_int]
L.lsio_ref_chunk_free.argtypes = [ctypes.c_void_p]
L.lsio_ref_shutdown.argtypes = [ctypes.c_void_p]
L.lsio_ref_free.argtypes = [ctypes.c_void_p]
L.lsio_ref_bench_read_files.restype = ctypes.c_double
L.lsio_ref_bench_read_files.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]


L.lsio_ref_bench_read_ranges.restype = ctypes.c_double
L.lsio_ref_bench_read_ranges.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]


def _buf(b):
    if isinstance(b, np.ndarray):
        return b.ctypes.data, b.nbytes
    b = bytes(b)
    return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), len(b)


def resolve_range(filesize, start, end, strict=False):
    r = Range()
    rc = L.lsio_oracle_resolve_range(filesize, start, end, int(strict), ctypes.byref(r))
    return rc, r


def crc32c(data, crc=0):
    p, n = _buf(data)
    return L.lsio_oracle_crc32c(crc, p, n)


def crc32c_bitwise(data, crc=0):
    p, n = _buf(data)
    return L.lsio_oracle_crc32c_bitwise(crc, p, n)


def crc32c_table(data, crc=0):
    p, n = _buf(data)
    return L.lsio_oracle_crc32c_table(crc, p, n)


def crc32c_file_blocks(path, blocksize, n_blocks):
    """uint32 array: CRC-32C of each `blocksize` block of the file (buffered pread; the GIL is released, so callers
    may run one thread per file)."""
    out = np.zeros(n_blocks, dtype=np.uint32)
    n = L.lsio_oracle_crc32c_file_blocks(os.fsencode(path), blocksize, n_blocks, out.ctypes.data)
    if n < 0:
        raise OSError(-n, os.strerror(-n), path)
    return out[:n]


def crc32c_combine(a, b, len_b):
    return L.lsio_oracle_crc32c_combine(a, b, len_b)


def xxh64(data, seed=0):
    p, n = _buf(data)
    return L.lsio_oracle_xxh64(p, n, seed)


def pattern(file_id, off, n):
    out = np.empty(n, dtype=np.uint8)
    L.lsio_oracle_pattern_fill(file_id, off, out.ctypes.data, n)
    return out


def ref_pattern(off, n):
    out = np.empty(n, dtype=np.uint8)
    L.lsio_oracle_ref_pattern_fill(off, out.ctypes.data, n)
    return out


def write_file(path, file_id, size, ref_pattern=False):
    rc = L.lsio_oracle_write_file(os.fsencode(path), file_id, size, int(ref_pattern))
    if rc:
        raise OSError(-rc, os.strerror(-rc), path)


def pread_range(path, start, end, cap=None):
    size = os.path.getsize(path)
    rc, r = resolve_range(size, start, end)
    if rc:
        raise ValueError(f"resolve_range failed: {rc}")
    n = r.end_off - r.start_off
    out = np.empty(max(n, 0), dtype=np.uint8)
    got = L.lsio_oracle_pread_range(os.fsencode(path), start, end, out.ctypes.data, out.nbytes)
    if got < 0:
        raise OSError(-got, os.strerror(-got), path)
    return out[:got]


def pack(src, src_off, dst_off, lens, dst_size, want_xxh=False):
    """numpy uint8 src -> (dst, crc32c[u32], xxh64[u64] or None)"""
    n = len(lens)
    so = np.ascontiguousarray(src_off, dtype=np.uint64)
    do = np.ascontiguousarray(dst_off, dtype=np.uint64)
    ln = np.ascontiguousarray(lens, dtype=np.uint64)
    dst = np.zeros(dst_size, dtype=np.uint8)
    crc = np.zeros(n, dtype=np.uint32)
    xx = np.zeros(n, dtype=np.uint64) if want_xxh else None
    L.lsio_oracle_pack(src.ctypes.data, dst.ctypes.data, so.ctypes.data, do.ctypes.data, ln.ctypes.data, n, crc.ctypes.data,
                       xx.ctypes.data if want_xxh else None)
    return dst, crc, xx


class Tracker:
    def __init__(self, n):
        self.h = L.lsio_oracle_tracker_new(n)

    def get_next_index(self):
        i = L.lsio_oracle_tracker_get_next_index(self.h)
        return None if i < 0 else i

    def put(self, idx, op):
        L.lsio_oracle_tracker_put(self.h, idx, op)

    def get(self, idx):
        v = ctypes.c_uint64()
        r = L.lsio_oracle_tracker_get(self.h, idx, ctypes.byref(v))
        if r < 0:
            raise IndexError("index out of bounds")
        return v.value if r else None

    def remove(self, idx):
        v = ctypes.c_uint64()
        r = L.lsio_oracle_tracker_remove(self.h, idx, ctypes.byref(v))
        if r < 0:
            raise IndexError("index out of bounds")
        return v.value if r else None

    def is_empty(self):
        return bool(L.lsio_oracle_tracker_is_empty(self.h))

    def is_full(self):
        return bool(L.lsio_oracle_tracker_is_full(self.h))

    def __del__(self):
        L.lsio_oracle_tracker_free(self.h)


class Disconnected(Exception):
    pass


class Timeout(Exception):
    pass


class RefChunk:
    __slots__ = ("status", "opcode", "user_data", "data")

    def __init__(self, status, opcode, user_data, data):
        self.status, self.opcode, self.user_data, self.data = status, opcode, user_data, data


class RefIoUring:
    """The CPU port of lsio_uring::IoUring (oracle/lsio_uring_ref.c). mode=1: object_store-like."""

    def __init__(self, n_worker_threads, sqpoll=True, mode=0):
        self.h = L.lsio_ref_new(n_worker_threads, int(sqpoll), mode)

    def get_ranges(self, location, ranges, user_data):
        assert len(ranges) == len(user_data)  # get_ranges.rs:28
        n = len(ranges)
        st = np.array([r[0] for r in ranges], dtype=np.int64)
        en = np.array([r[1] for r in ranges], dtype=np.int64)
        ud = np.array(user_data, dtype=np.uint64)
        L.lsio_ref_get_ranges(self.h, os.fsencode(location), st.ctypes.data, en.ctypes.data, ud.ctypes.data, n)

    def recv_timeout(self, seconds):
        o = RefOutput()
        rc = L.lsio_ref_recv(self.h, ctypes.byref(o), int(seconds * 1000))
        if rc == 1:
            raise Timeout()
        if rc == 2:
            raise Disconnected()
        data = None
        if o.status == 0:
            data = np.ctypeslib.as_array(ctypes.cast(o.ptr, _u8p), shape=(o.len,)).copy() if o.len else np.empty(0, np.uint8)
            L.lsio_ref_chunk_free(o.base)
        return RefChunk(o.status, o.opcode, o.user_data, data)

    def shutdown(self):
        L.lsio_ref_shutdown(self.h)

    def close(self):
        if self.h:
            L.lsio_ref_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def bench_read_files(paths, starts, ends, n_worker_threads, sqpoll=True, mode=0):
    """lsio_bench::read_files restated (oracle/lsio_uring_ref.c). Returns (seconds, bytes, errors)."""
    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    st = np.ascontiguousarray(starts, dtype=np.int64)
    en = np.ascontiguousarray(ends, dtype=np.int64)
    b, e = ctypes.c_uint64(), ctypes.c_uint64()
    secs = L.lsio_ref_bench_read_files(arr, len(paths), st.ctypes.data, en.ctypes.data, len(st), n_worker_threads, int(sqpoll), mode,
                                       ctypes.byref(b), ctypes.byref(e))
    return secs, b.value, e.value


def bench_read_ranges(paths, per_file_ranges, n_worker_threads, sqpoll=True, mode=0):
    """As bench_read_files for arbitrary per-file (starts, ends) arrays. Returns (seconds, bytes, errors)."""
    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    st = np.ascontiguousarray(np.concatenate([r[0] for r in per_file_ranges]), dtype=np.int64)
    en = np.ascontiguousarray(np.concatenate([r[1] for r in per_file_ranges]), dtype=np.int64)
    cnt = np.array([len(r[0]) for r in per_file_ranges], dtype=np.uint64)
    b, e = ctypes.c_uint64(), ctypes.c_uint64()
    secs = L.lsio_ref_bench_read_ranges(arr, len(paths), st.ctypes.data, en.ctypes.data, cnt.ctypes.data, n_worker_threads, int(sqpoll), mode,
                                        ctypes.byref(b), ctypes.byref(e))
    return secs, b.value, e.value


# ---- codec layer oracle (oracle/lsio_codecs_oracle.c) ----------------------------------------------
L.lsio_oracle_lz4_decompress.restype = ctypes.c_int64
L.lsio_oracle_lz4_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
L.lsio_oracle_lz4_compress.restype = ctypes.c_int64
L.lsio_oracle_lz4_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
L.lsio_oracle_lz4_compress_bound.restype = ctypes.c_size_t
L.lsio_oracle_lz4_compress_bound.argtypes = [ctypes.c_size_t]
L.lsio_oracle_shuffle.restype = None
L.lsio_oracle_shuffle.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
L.lsio_oracle_unshuffle.restype = None
L.lsio_oracle_unshuffle.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]


def _u8(data):
    return np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else data, dtype=np.uint8)


def lz4_decompress(data, cap):
    """LZ4 block decode; returns the decoded bytes or raises ValueError(code)."""
    src = _u8(data)
    dst = np.empty(max(cap, 1), dtype=np.uint8)
    n = L.lsio_oracle_lz4_decompress(src.ctypes.data, len(src), dst.ctypes.data, cap)
    if n < 0:
        raise ValueError(n)
    return dst[:n]


def lz4_compress(data):
    src = _u8(data)
    cap = L.lsio_oracle_lz4_compress_bound(len(src))
    dst = np.empty(cap, dtype=np.uint8)
    n = L.lsio_oracle_lz4_compress(src.ctypes.data, len(src), dst.ctypes.data, cap)
    assert n >= 0
    return dst[:n].copy()


def shuffle(data, elem):
    src = _u8(data)
    dst = np.empty_like(src)
    L.lsio_oracle_shuffle(src.ctypes.data, dst.ctypes.data, len(src), elem)
    return dst


def unshuffle(data, elem):
    src = _u8(data)
    dst = np.empty_like(src)
    L.lsio_oracle_unshuffle(src.ctypes.data, dst.ctypes.data, len(src), elem)
    return dst
