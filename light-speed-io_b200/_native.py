"""ctypes loader for liblsio_b200.so (the C ABI declared in include/*.h).

There is deliberately no fallback: if the CUDA extension has not been built, importing this
module raises, and every product entry point fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LSIO_B200_LIB") or os.path.join(_HERE, "liblsio_b200.so")  # env override: kernel A/B builds


class NativeLibraryMissing(ImportError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C light-speed-io_b200` (there is no CPU fallback)"
        )
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()

c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u32p = ctypes.POINTER(ctypes.c_uint32)


class CudaSegment(ctypes.Structure):
    """include/lsio_cuda.h: lsio_cuda_segment"""
    _fields_ = [("src", ctypes.c_uint64), ("dst", ctypes.c_uint64), ("len", ctypes.c_uint64)]


PACK_STORE = 1
PACK_CRC32C = 2
PACK_LIGHT = 4

lib.lsio_cuda_pull_strip_crc32c.restype = ctypes.c_int
lib.lsio_cuda_pull_strip_crc32c.argtypes = [ctypes.c_int, ctypes.POINTER(CudaSegment), ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
PULL_MAX_SEGMENTS = 64

lib.lsio_cuda_pack_plan_create.restype = ctypes.c_int
lib.lsio_cuda_pack_plan_create.argtypes = [ctypes.c_int, ctypes.POINTER(CudaSegment), ctypes.c_size_t, ctypes.c_uint32,
                                           ctypes.POINTER(ctypes.c_void_p)]
lib.lsio_cuda_pack_plan_launch.restype = ctypes.c_int
lib.lsio_cuda_pack_plan_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
lib.lsio_cuda_pack_plan_algorithmic_bytes.restype = ctypes.c_uint64
lib.lsio_cuda_pack_plan_algorithmic_bytes.argtypes = [ctypes.c_void_p]
lib.lsio_cuda_pack_plan_num_tiles.restype = ctypes.c_uint64
lib.lsio_cuda_pack_plan_num_tiles.argtypes = [ctypes.c_void_p]
lib.lsio_cuda_pack_plan_destroy.restype = None
lib.lsio_cuda_pack_plan_destroy.argtypes = [ctypes.c_void_p]
lib.lsio_cuda_pack_strip_crc32c.restype = ctypes.c_int
lib.lsio_cuda_pack_strip_crc32c.argtypes = [ctypes.c_int, ctypes.POINTER(CudaSegment), ctypes.c_size_t, ctypes.c_uint32,
                                            ctypes.c_void_p, ctypes.c_void_p]
lib.lsio_cuda_xxh64.restype = ctypes.c_int
lib.lsio_cuda_xxh64.argtypes = [ctypes.c_int, c_u64p, c_u64p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
lib.lsio_cuda_xxh64_variant.restype = ctypes.c_int
lib.lsio_cuda_xxh64_variant.argtypes = [ctypes.c_int, c_u64p, c_u64p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
lib.lsio_cuda_verify_pattern.restype = ctypes.c_int
lib.lsio_cuda_verify_pattern.argtypes = [ctypes.c_int, c_u64p, c_u64p, c_u64p, c_u64p, ctypes.c_size_t, ctypes.c_void_p,
                                         ctypes.c_void_p]
lib.lsio_cuda_launch_count.restype = ctypes.c_uint64
lib.lsio_cuda_launch_count.argtypes = []
lib.lsio_cuda_last_error.restype = ctypes.c_char_p
lib.lsio_cuda_last_error.argtypes = []


class LsioCudaError(RuntimeError):
    pass


def check(rc, what):
    if rc == 0:
        return
    if rc > 0:
        raise LsioCudaError(f"{what}: CUDA error {rc}: {lib.lsio_cuda_last_error().decode(errors='replace')}")
    raise LsioCudaError(f"{what}: {os.strerror(-rc)} (rc={rc})")


# ---- include/lsio_aligned_bytes.h ------------------------------------------------------------
MEM_HOST, MEM_PINNED, MEM_DEVICE = 0, 1, 2
AB_ERR_NOT_IN_RANGE, AB_ERR_ZERO, AB_ERR_UNALIGNED, AB_ERR_SHARED, AB_ERR_BAD_SLICE = -1, -2, -3, -4, -5
_vp = ctypes.c_void_p
for _name, _res, _args in [
    ("lsio_aligned_bytes_mut_new", _vp, [ctypes.c_size_t, ctypes.c_size_t]),
    ("lsio_aligned_bytes_mut_new_pinned", _vp, [ctypes.c_size_t, ctypes.c_size_t]),
    ("lsio_aligned_bytes_mut_new_device", _vp, [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]),
    ("lsio_aligned_bytes_mut_len", ctypes.c_size_t, [_vp]),
    ("lsio_aligned_bytes_mut_as_mut_ptr", _vp, [_vp]),
    ("lsio_aligned_bytes_mut_split_to", ctypes.c_int, [_vp, ctypes.c_size_t, ctypes.POINTER(_vp)]),
    ("lsio_aligned_bytes_mut_freeze", ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    ("lsio_aligned_bytes_mut_drop", None, [_vp]),
    ("lsio_aligned_bytes_clone", _vp, [_vp]),
    ("lsio_aligned_bytes_set_slice", ctypes.c_int, [_vp, ctypes.c_size_t, ctypes.c_size_t]),
    ("lsio_aligned_bytes_reset_slice", None, [_vp]),
    ("lsio_aligned_bytes_len", ctypes.c_size_t, [_vp]),
    ("lsio_aligned_bytes_as_ptr", _vp, [_vp]),
    ("lsio_aligned_bytes_memory_kind", ctypes.c_int, [_vp]),
    ("lsio_aligned_bytes_device", ctypes.c_int, [_vp]),
    ("lsio_aligned_bytes_backing_len", ctypes.c_size_t, [_vp]),
    ("lsio_aligned_bytes_copy_to_host", ctypes.c_int64, [_vp, _vp, ctypes.c_size_t]),
    ("lsio_aligned_bytes_drop", None, [_vp]),
    ("lsio_aligned_bytes_to_dlpack", ctypes.c_int, [_vp, ctypes.c_uint8, ctypes.c_uint8, _vp, ctypes.c_int32, ctypes.POINTER(_vp)]),
]:
    getattr(lib, _name).restype = _res
    getattr(lib, _name).argtypes = _args


# ---- include/lsio_io.h --------------------------------------------------------------------------
class CChunk(ctypes.Structure):
    _fields_ = [("buffer", ctypes.c_void_p), ("user_data", ctypes.c_uint64)]


class COutput(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("opcode", ctypes.c_uint8), ("kind", ctypes.c_uint8), ("checksums", ctypes.c_uint16),
                ("device", ctypes.c_int32), ("crc32c", ctypes.c_uint32), ("xxh64", ctypes.c_uint64), ("chunk", CChunk),
                ("group", ctypes.c_uint64), ("context", ctypes.c_char * 256)]


CHECKSUM_NONE, CHECKSUM_CRC32C, CHECKSUM_XXH64 = 0, 1, 2
RECV_OK, RECV_TIMEOUT, RECV_DISCONNECTED = 0, 1, 2
for _name, _res, _args in [
    ("lsio_completion_clone", _vp, [_vp]),
    ("lsio_completion_recv", ctypes.c_int, [_vp, ctypes.POINTER(COutput), ctypes.c_int]),
    ("lsio_completion_try_recv", ctypes.c_int, [_vp, ctypes.POINTER(COutput)]),
    ("lsio_completion_len", ctypes.c_size_t, [_vp]),
    ("lsio_completion_drop", None, [_vp]),
]:
    getattr(lib, _name).restype = _res
    getattr(lib, _name).argtypes = _args

# ---- include/lsio_uring_cuda.h ---------------------------------------------------------------
MAX_DEVICES = 16


class CConfig(ctypes.Structure):
    _fields_ = [("n_worker_threads", ctypes.c_uint32), ("n_devices", ctypes.c_uint32), ("devices", ctypes.c_int32 * MAX_DEVICES),
                ("sq_entries", ctypes.c_uint32), ("sqpoll", ctypes.c_uint32), ("completion_bound", ctypes.c_uint32),
                ("checksum", ctypes.c_uint32), ("pack_align", ctypes.c_uint32), ("pin_threads", ctypes.c_uint32),
                ("o_direct", ctypes.c_uint32), ("staging_bytes", ctypes.c_uint64), ("arena_bytes", ctypes.c_uint64),
                ("max_read_bytes", ctypes.c_uint64), ("min_copy_bytes", ctypes.c_uint64), ("flush_us", ctypes.c_uint32),
                ("merge_ranges", ctypes.c_uint32), ("merge_gap_bytes", ctypes.c_uint64),
                ("statx_via_ring", ctypes.c_uint32), ("poll_sleep_us", ctypes.c_uint32), ("zero_copy", ctypes.c_uint32),
                ("staging_mode", ctypes.c_uint32), ("host_slots_per_worker", ctypes.c_uint32), ("host_slot_kib", ctypes.c_uint32),
                ("pack_threads", ctypes.c_uint32),
                ("inline_reads", ctypes.c_uint32), ("force_align", ctypes.c_uint32), ("debug_short_read_every", ctypes.c_uint32),
                ("null_source", ctypes.c_uint32), ("direct_landing", ctypes.c_uint32), ("cpu_first", ctypes.c_uint32),
                ("cpu_count", ctypes.c_uint32), ("steal_batch", ctypes.c_uint32), ("pack_light_below_mib", ctypes.c_uint32),
                ("copy_coalesce_slots", ctypes.c_uint32), ("window_budget_kib", ctypes.c_uint32), ("pull_any_shift", ctypes.c_uint32)]


class CStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("requests", "ranges", "reads", "short_reads", "read_bytes", "h2d_copies", "h2d_bytes",
                                               "pack_launches", "kernel_launches", "payload_bytes", "chunks_ok", "errors",
                                               "staging_stalls", "arena_stalls", "inline_reads", "direct_extents", "direct_copies",
                                               "read_ns", "copy_enqueue_ns", "worker_stall_ns", "worker_busy_ns", "pack_kernel_ns",
                                               "pack_kernel_bytes", "first_chunk_ns", "max_dio_align", "pull_launches", "slot_bytes", "slots_per_worker")]


for _name, _res, _args in [
    ("lsio_uring_cuda_config_default", None, [ctypes.POINTER(CConfig)]),
    ("lsio_uring_cuda_new", ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
    ("lsio_uring_cuda_new_with_config", ctypes.c_int, [ctypes.POINTER(CConfig), ctypes.POINTER(_vp)]),
    ("lsio_uring_cuda_get_ranges", ctypes.c_int, [_vp, ctypes.c_char_p, _vp, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t]),
    ("lsio_uring_cuda_get_ranges_on", ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_char_p, _vp, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t]),
    ("lsio_uring_cuda_completion", _vp, [_vp]),
    ("lsio_uring_cuda_next_group", ctypes.c_uint64, [_vp]),
    ("lsio_uring_cuda_current_group", ctypes.c_uint64, [_vp]),
    ("lsio_uring_cuda_get_stats", ctypes.c_int, [_vp, ctypes.POINTER(CStats)]),
    ("lsio_uring_cuda_get_latency_histogram", ctypes.c_int, [_vp, c_u64p]),
    ("lsio_uring_cuda_reset_latency", None, [_vp]),
    ("lsio_uring_cuda_last_error", ctypes.c_char_p, []),
    ("lsio_uring_cuda_free", None, [_vp]),
]:
    getattr(lib, _name).restype = _res
    getattr(lib, _name).argtypes = _args

# ---- include/lsio_threadpool.h ---------------------------------------------------------------
TASK_FN = ctypes.CFUNCTYPE(ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64)
lib.lsio_threadpool_new.restype = ctypes.c_void_p
lib.lsio_threadpool_new.argtypes = [ctypes.c_size_t, TASK_FN, ctypes.c_void_p]
lib.lsio_threadpool_push.restype = None
lib.lsio_threadpool_push.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
lib.lsio_threadpool_n_parks.restype = ctypes.c_uint64
lib.lsio_threadpool_n_parks.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
lib.lsio_threadpool_free.restype = None
lib.lsio_threadpool_free.argtypes = [ctypes.c_void_p]

# ---- bulk / measurement entry points -----------------------------------------------------------
lib.lsio_cuda_fill_pattern.restype = ctypes.c_int
lib.lsio_cuda_fill_pattern.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
lib.lsio_completion_recv_batch.restype = ctypes.c_int
lib.lsio_completion_recv_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(COutput), ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]
lib.lsio_aligned_bytes_drop_many.restype = None
lib.lsio_aligned_bytes_drop_many.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
lib.lsio_aligned_bytes_describe_many.restype = None
lib.lsio_aligned_bytes_describe_many.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
lib.lsio_probe_read_ceiling.restype = ctypes.c_int
lib.lsio_probe_read_ceiling.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.POINTER(ctypes.c_double)]


lib.lsio_probe_read_ceiling_ex.restype = ctypes.c_int
lib.lsio_probe_read_ceiling_ex.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_double)]


lib.lsio_uring_cuda_plan_reads.restype = ctypes.c_int
lib.lsio_uring_cuda_plan_reads.argtypes = [ctypes.c_int64, c_i64p, c_i64p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32,
                                           ctypes.c_uint64, c_i64p, c_u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), c_u64p, ctypes.c_size_t,
                                           ctypes.POINTER(ctypes.c_size_t), c_u64p]


def plan_reads(filesize, starts, ends, align=512, max_read_bytes=512 << 10, merge_ranges=1, merge_gap_bytes=0, pack_align=1, buffer_address=1 << 30):
    """The request planner's arithmetic on its own (lsio_uring_cuda_plan_reads; no device): returns (ranges[n,4] int64, reads[r,6] uint64,
    segs[s,4] uint64, total_bytes)."""
    import numpy as np

    st = np.ascontiguousarray(starts, dtype=np.int64)
    en = np.ascontiguousarray(ends, dtype=np.int64)
    n = len(st)
    ranges = np.zeros((max(n, 1), 4), dtype=np.int64)
    nr, ns, tot = ctypes.c_size_t(0), ctypes.c_size_t(0), ctypes.c_uint64(0)
    reads, segs = np.zeros((0, 6), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64)
    for _ in range(2):  # first call sizes the outputs
        rc = lib.lsio_uring_cuda_plan_reads(filesize, st.ctypes.data_as(c_i64p), en.ctypes.data_as(c_i64p), n, align, max_read_bytes, merge_ranges, merge_gap_bytes, pack_align,
                                            buffer_address, ranges.ctypes.data_as(c_i64p), reads.ctypes.data_as(c_u64p), len(reads), ctypes.byref(nr),
                                            segs.ctypes.data_as(c_u64p), len(segs), ctypes.byref(ns), ctypes.byref(tot))
        if rc:
            raise ValueError(f"lsio_uring_cuda_plan_reads: rc={rc}")
        if len(reads) >= nr.value and len(segs) >= ns.value:
            break
        reads, segs = np.zeros((nr.value, 6), dtype=np.uint64), np.zeros((ns.value, 4), dtype=np.uint64)
    return ranges[:n], reads[: nr.value], segs[: ns.value], tot.value


lib.lsio_probe_staged_ceiling.restype = ctypes.c_int
lib.lsio_probe_staged_ceiling.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                          ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]


# ---- include/lsio_bench.h -------------------------------------------------------------------------
class CBenchResult(ctypes.Structure):
    _fields_ = [("seconds", ctypes.c_double), ("chunks_ok", ctypes.c_uint64), ("chunks_bad", ctypes.c_uint64), ("payload_bytes", ctypes.c_uint64)]


lib.lsio_bench_read_files.restype = ctypes.c_int
lib.lsio_bench_read_files.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64,
                                      ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(CBenchResult)]
lib.lsio_bench_read_ranges.restype = ctypes.c_int
lib.lsio_bench_read_ranges.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(CBenchResult)]
lib.lsio_uring_cuda_selftest.restype = ctypes.c_int
lib.lsio_uring_cuda_selftest.argtypes = [ctypes.c_int]
