"""ctypes loader for liblsio_b200.so (the C ABI declared in include/*.h).

There is deliberately no fallback: if the CUDA extension has not been built, importing this
module raises, and every product entry point fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblsio_b200.so")


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
