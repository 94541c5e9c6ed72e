"""Python mirror of `lsio_io` (Reader / Completion / Chunk / Output) and of the backend
`lsio_uring::IoUring`, over include/lsio_io.h and include/lsio_uring_cuda.h.

    reference                                                here
    IoUring::new(n_worker_threads)   io_uring.rs:15          IoUringCuda(n_worker_threads, ...)
    Reader::get_ranges               lsio_io lib.rs:36-44    IoUringCuda.get_ranges(location, ranges, user_data)
    Completion::completion           lsio_io lib.rs:11-13    IoUringCuda.completion() -> Receiver
    Receiver::recv_timeout / recv    crossbeam-channel       Receiver.recv_timeout(seconds) / recv() / try_recv()
    Chunk { buffer, user_data }      lsio_io lib.rs:50-56    Chunk.buffer (AlignedBytes, device view), Chunk.user_data
    Err(anyhow::Error)               operation.rs:82-100     IoError item (returned, not raised -- it is a channel item)
"""
import ctypes
import sys
import os

import numpy as np

from . import _native as N
from .aligned_bytes import AlignedBytes


OUTPUT_DTYPE = np.dtype(N.COutput)


class RecvTimeoutError(Exception):
    """crossbeam_channel::RecvTimeoutError::Timeout"""


class Disconnected(Exception):
    """crossbeam_channel::RecvTimeoutError::Disconnected / RecvError"""


class Chunk:
    """lsio_io::Chunk (lib.rs:50-56), plus the device-side proof of its bytes."""

    __slots__ = ("buffer", "user_data", "crc32c", "xxh64", "device", "group")

    def __init__(self, buffer, user_data, crc32c, xxh64, device, group=0):
        self.buffer, self.user_data, self.crc32c, self.xxh64, self.device, self.group = buffer, user_data, crc32c, xxh64, device, group

    def is_ok(self):
        return True

    def __repr__(self):
        return f"Chunk(user_data={self.user_data}, len={self.buffer.len()}, device={self.device}, crc32c={self.crc32c})"


class IoError:
    """An `Err(anyhow::Error)` item of the completion channel (operation.rs:82-100)."""

    __slots__ = ("status", "opcode", "user_data", "context", "group")

    def __init__(self, status, opcode, user_data, context, group=0):
        self.status, self.opcode, self.user_data, self.context, self.group = status, opcode, user_data, context, group

    @property
    def errno(self):
        return -self.status if self.status < 0 else None

    def is_ok(self):
        return False

    def __repr__(self):
        return f"IoError(status={self.status}, opcode={self.opcode}, user_data={self.user_data}, context={self.context!r})"


def _convert(o):
    if o.status == 0:
        crc = o.crc32c if (o.checksums & N.CHECKSUM_CRC32C) else None
        xx = o.xxh64 if (o.checksums & N.CHECKSUM_XXH64) else None
        return Chunk(AlignedBytes(o.chunk.buffer), o.chunk.user_data, crc, xx, o.device, o.group)
    return IoError(o.status, o.opcode, o.chunk.user_data, o.context.decode(errors="replace"), o.group)


class Receiver:
    """The backend's completion queue: cloneable, usable from any thread, outlives the backend."""

    def __init__(self, handle):
        self._h = handle

    def clone(self):
        return Receiver(N.lib.lsio_completion_clone(self._h))

    def _recv(self, timeout_ms):
        o = N.COutput()
        rc = N.lib.lsio_completion_recv(self._h, ctypes.byref(o), timeout_ms)
        if rc == N.RECV_TIMEOUT:
            raise RecvTimeoutError()
        if rc == N.RECV_DISCONNECTED:
            raise Disconnected()
        if rc != N.RECV_OK:
            raise RuntimeError(f"lsio_completion_recv: {rc}")
        return _convert(o)

    def recv(self):
        return self._recv(-1)

    def recv_timeout(self, seconds):
        return self._recv(max(0, int(seconds * 1000)))

    def try_recv(self):
        return self._recv(0)

    def recv_batch(self, max_items=1024, timeout=10.0, _buf={}):
        """Up to `max_items` completions with one call, as a numpy record array over the C
        `lsio_output` layout (fields: status, opcode, kind, checksums, device, crc32c, xxh64,
        chunk.buffer, chunk.user_data, context).  The caller owns every non-NULL chunk.buffer:
        release them with `drop_buffers(records)`.  Blocks for the first item like recv_timeout."""
        arr = (N.COutput * max_items)()
        n = ctypes.c_size_t(0)
        rc = N.lib.lsio_completion_recv_batch(self._h, arr, max_items, -1 if timeout is None else max(0, int(timeout * 1000)), ctypes.byref(n))
        if rc == N.RECV_TIMEOUT:
            raise RecvTimeoutError()
        if rc == N.RECV_DISCONNECTED:
            raise Disconnected()
        if rc != N.RECV_OK:
            raise RuntimeError(f"lsio_completion_recv_batch: {rc}")
        return np.frombuffer(arr, dtype=OUTPUT_DTYPE, count=n.value)  # writable: take_buffer() clears the handles it hands out

    # A chunk handle has ONE owner.  A record array owns the handles in its chunk.buffer field until they are either
    # dropped with drop_buffers(records) or moved out with take_buffer(s), which clears the field so that a later
    # drop_buffers() on the same records skips them (a handle wrapped in AlignedBytes AND dropped in bulk was a
    # double free).
    @staticmethod
    def take_buffer(records, i):
        """Move the chunk buffer of records[i] out as an AlignedBytes (the record's field becomes NULL)."""
        h = int(records["chunk"]["buffer"][i])
        if not h:
            raise ValueError("records[%d] holds no chunk buffer (an error item, or already taken)" % i)
        records["chunk"]["buffer"][i] = 0
        return AlignedBytes(h)

    @staticmethod
    def take_buffers(records):
        """Move every chunk buffer out: a list of AlignedBytes (None for error items); the records keep none."""
        hs = [int(h) for h in records["chunk"]["buffer"]]
        records["chunk"]["buffer"][:] = 0
        return [AlignedBytes(h) if h else None for h in hs]

    @staticmethod
    def drop_buffers(records):
        """Drop the chunk buffers a record array from recv_batch still owns (lsio_aligned_bytes_drop each; NULLs are skipped)."""
        ptrs = np.ascontiguousarray(records["chunk"]["buffer"], dtype=np.uint64)
        N.lib.lsio_aligned_bytes_drop_many(ptrs.ctypes.data, len(ptrs))
        if records.flags.writeable:
            records["chunk"]["buffer"][:] = 0

    def __len__(self):
        return N.lib.lsio_completion_len(self._h)

    def drop(self):
        if self._h:
            N.lib.lsio_completion_drop(self._h)
            self._h = None

    def __del__(self):
        if sys.is_finalizing():
            return  # interpreter exit: no native (CUDA) calls from the garbage collector; the process is going away
        try:
            self.drop()
        except Exception:
            pass


class IoUringCuda:
    """Drop-in for `lsio_uring::IoUring` whose chunks land in B200 HBM."""

    def __init__(self, n_worker_threads, devices=(0,), **config):
        cfg = N.CConfig()
        N.lib.lsio_uring_cuda_config_default(ctypes.byref(cfg))
        cfg.n_worker_threads = n_worker_threads
        devices = list(devices)
        cfg.n_devices = len(devices)
        for i, d in enumerate(devices):
            cfg.devices[i] = d
        for k, v in config.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown config field {k}")
            setattr(cfg, k, int(v))
        self._h = ctypes.c_void_p()
        rc = N.lib.lsio_uring_cuda_new_with_config(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc:
            self._h = None
            raise N.LsioCudaError(f"IoUringCuda::new failed (rc={rc}): {N.lib.lsio_uring_cuda_last_error().decode(errors='replace')}")
        self._rx = Receiver(N.lib.lsio_uring_cuda_completion(self._h))
        self.devices = devices
        self.arena_bytes = int(cfg.arena_bytes)

    @classmethod
    def new(cls, n_worker_threads, **kw):
        return cls(n_worker_threads, **kw)

    def get_ranges(self, location, ranges, user_data, group=None):
        """`ranges`: iterable of (start, end) half-open pairs (Range<isize>); negative = from the end."""
        if self._h is None:
            raise RuntimeError("backend dropped")
        starts = np.array([r[0] for r in ranges], dtype=np.int64)
        ends = np.array([r[1] for r in ranges], dtype=np.int64)
        ud = np.array(list(user_data), dtype=np.uint64)
        loc = os.fsencode(os.fspath(location))
        if group is None:
            rc = N.lib.lsio_uring_cuda_get_ranges(self._h, loc, starts.ctypes.data, ends.ctypes.data, len(starts), ud.ctypes.data, len(ud))
        else:
            rc = N.lib.lsio_uring_cuda_get_ranges_on(self._h, group, loc, starts.ctypes.data, ends.ctypes.data, len(starts), ud.ctypes.data, len(ud))
        if rc:
            # the reference panics on a length mismatch (get_ranges.rs:28); raise instead
            raise ValueError(f"get_ranges rejected (rc={rc}): ranges/user_data length mismatch, relative path, or backend closed")

    def get_ranges_arrays(self, location, starts, ends, user_data, group=None):
        """Same, taking int64/int64/uint64 numpy arrays (no per-range Python objects)."""
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        ends = np.ascontiguousarray(ends, dtype=np.int64)
        ud = np.ascontiguousarray(user_data, dtype=np.uint64)
        loc = os.fsencode(os.fspath(location))
        if group is None:
            rc = N.lib.lsio_uring_cuda_get_ranges(self._h, loc, starts.ctypes.data, ends.ctypes.data, len(starts), ud.ctypes.data, len(ud))
        else:
            rc = N.lib.lsio_uring_cuda_get_ranges_on(self._h, group, loc, starts.ctypes.data, ends.ctypes.data, len(starts), ud.ctypes.data, len(ud))
        if rc:
            raise ValueError(f"get_ranges rejected (rc={rc})")

    def next_group(self):
        """Close the current ordered group and open the next (planned in the reference: lsio_io lib.rs:6-8,
        planned_design.md:71-72): requests submitted from now on start only after every earlier group has
        delivered all of its items; items carry `.group`. Returns the new group id."""
        return N.lib.lsio_uring_cuda_next_group(self._h)

    def current_group(self):
        return N.lib.lsio_uring_cuda_current_group(self._h)

    def completion(self):
        return self._rx

    def bench_read_files(self, paths, file_ids, blocksize, n_chunks, expected_crc32c=None):
        """lsio_bench::read_files (crates/lsio_bench/src/main.rs:124-190) run natively over this backend:
        returns (seconds, payload_bytes, chunks_ok, chunks_bad)."""
        arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(os.fspath(p)) for p in paths])
        ids = np.ascontiguousarray(file_ids, dtype=np.uint32)
        crc = None if expected_crc32c is None else np.ascontiguousarray(expected_crc32c, dtype=np.uint32)
        res = N.CBenchResult()
        rc = N.lib.lsio_bench_read_files(self._h, arr, ids.ctypes.data, len(paths), blocksize, n_chunks,
                                         None if crc is None else crc.ctypes.data, 0 if crc is None else crc.shape[1], ctypes.byref(res))
        if rc:
            raise RuntimeError(f"lsio_bench_read_files failed: rc={rc}")
        return res.seconds, res.payload_bytes, res.chunks_ok, res.chunks_bad

    def bench_read_ranges(self, paths, per_file_ranges, expected_crc32c=None):
        """Native driver for arbitrary per-file (starts, ends) arrays; user_data = global range index.
        Returns (seconds, payload_bytes, chunks_ok, chunks_bad)."""
        arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(os.fspath(p)) for p in paths])
        st = np.ascontiguousarray(np.concatenate([r[0] for r in per_file_ranges]), dtype=np.int64)
        en = np.ascontiguousarray(np.concatenate([r[1] for r in per_file_ranges]), dtype=np.int64)
        cnt = np.array([len(r[0]) for r in per_file_ranges], dtype=np.uint64)
        crc = None if expected_crc32c is None else np.ascontiguousarray(expected_crc32c, dtype=np.uint32)
        res = N.CBenchResult()
        rc = N.lib.lsio_bench_read_ranges(self._h, arr, len(paths), st.ctypes.data, en.ctypes.data, cnt.ctypes.data,
                                          None if crc is None else crc.ctypes.data, ctypes.byref(res))
        if rc:
            raise RuntimeError(f"lsio_bench_read_ranges failed: rc={rc}")
        return res.seconds, res.payload_bytes, res.chunks_ok, res.chunks_bad

    def stats(self):
        s = N.CStats()
        N.lib.lsio_uring_cuda_get_stats(self._h, ctypes.byref(s))
        return {n: getattr(s, n) for n, _ in s._fields_}

    def latency_histogram(self):
        """Per-chunk latency (request submitted -> chunk in the channel) as 32 log2 buckets of microseconds."""
        b = (ctypes.c_uint64 * 32)()
        N.lib.lsio_uring_cuda_get_latency_histogram(self._h, b)
        return np.array(list(b), dtype=np.uint64)

    def latency_summary(self):
        """{first_chunk_ms, p50_ms, p99_ms (upper bucket edges), chunks} since the last reset_latency()."""
        h = self.latency_histogram()
        n = int(h.sum())
        out = {"chunks": n, "first_chunk_ms": self.stats()["first_chunk_ns"] / 1e6}
        if n:
            c = np.cumsum(h)
            for name, q in (("p50_ms", 0.5), ("p99_ms", 0.99)):
                out[name] = float(2 ** (int(np.searchsorted(c, q * n)) + 1)) / 1e3
        return out

    def reset_latency(self):
        N.lib.lsio_uring_cuda_reset_latency(self._h)

    def drop(self):
        """Drop: joins the workers; receivers see Disconnected once drained."""
        if self._h is not None:
            N.lib.lsio_uring_cuda_free(self._h)
            self._h = None

    close = drop

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.drop()

    def __del__(self):
        if sys.is_finalizing():
            return  # interpreter exit: no native (CUDA) calls from the garbage collector; the process is going away
        try:
            self.drop()
        except Exception:
            pass
