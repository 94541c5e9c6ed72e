"""Python mirror of the reference crate `lsio_aligned_bytes` over include/lsio_aligned_bytes.h.

Same names and behaviour as crates/lsio_aligned_bytes/src/lib.rs: `AlignedBytesMut.new`,
`len`, `as_mut_ptr`, `split_to`, `freeze`; `AlignedBytes.set_slice`, `reset_slice`, `len`,
`as_ptr`, `as_slice`, `clone`.  Where the reference returns Err / panics, these raise.
"""
import ctypes
import sys

import numpy as np

from . import _native as N

_SPLIT_MSG = {
    N.AB_ERR_NOT_IN_RANGE: "idx {idx} is not contained in this buffer's range {rng}",  # lib.rs:76-79
    N.AB_ERR_ZERO: "idx must not be zero!",                                            # lib.rs:81
    N.AB_ERR_UNALIGNED: "idx {idx} must be exactly divisible by the alignment",        # lib.rs:83-86
}


class AlignedBytesError(ValueError):
    pass


class AlignedBytesMut:
    """Exclusive, splittable view (lib.rs:9-113)."""

    def __init__(self, len, align, kind="host", device=0, _handle=None):
        if _handle is not None:
            self._h = _handle
            return
        if kind == "host":
            h = N.lib.lsio_aligned_bytes_mut_new(len, align)
        elif kind == "pinned":
            h = N.lib.lsio_aligned_bytes_mut_new_pinned(len, align)
        elif kind == "device":
            h = N.lib.lsio_aligned_bytes_mut_new_device(len, align, device)
        else:
            raise ValueError(kind)
        if not h:
            raise AlignedBytesError("allocation failed (len must be > 0 and align a power of two)")
        self._h = h

    @classmethod
    def new(cls, len, align, **kw):
        return cls(len, align, **kw)

    def len(self):
        return N.lib.lsio_aligned_bytes_mut_len(self._h)

    def as_mut_ptr(self):
        return N.lib.lsio_aligned_bytes_mut_as_mut_ptr(self._h)

    def as_numpy(self):
        """Writable numpy view (host / pinned memory only)."""
        return np.ctypeslib.as_array(ctypes.cast(self.as_mut_ptr(), ctypes.POINTER(ctypes.c_uint8)), shape=(self.len(),))

    def split_to(self, idx):
        out = ctypes.c_void_p()
        rc = N.lib.lsio_aligned_bytes_mut_split_to(self._h, idx, ctypes.byref(out))
        if rc:
            raise AlignedBytesError(_SPLIT_MSG.get(rc, "split_to failed").format(idx=idx, rng="?"))
        return AlignedBytesMut(0, 0, _handle=out.value)

    def freeze(self):
        """Consumes self. Raises (and leaves self valid) if other mutable views exist (lib.rs:103-112)."""
        out = ctypes.c_void_p()
        rc = N.lib.lsio_aligned_bytes_mut_freeze(self._h, ctypes.byref(out))
        if rc:
            raise AlignedBytesError("freeze: other AlignedBytesMut views still reference the buffer")
        self._h = None
        return AlignedBytes(out.value)

    def drop(self):
        if self._h:
            N.lib.lsio_aligned_bytes_mut_drop(self._h)
            self._h = None

    def __del__(self):
        if sys.is_finalizing():
            return  # interpreter exit: no native (CUDA) calls from the garbage collector; the process is going away
        try:
            self.drop()
        except Exception:
            pass


_DL_DTYPES = {"uint8": (1, 8), "int8": (0, 8), "uint16": (1, 16), "int16": (0, 16), "uint32": (1, 32), "int32": (0, 32),
              "uint64": (1, 64), "int64": (0, 64), "float16": (2, 16), "float32": (2, 32), "float64": (2, 64), "bfloat16": (4, 16)}
_capsule_new = ctypes.pythonapi.PyCapsule_New
_capsule_new.restype = ctypes.py_object
_capsule_new.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
# the destructor sees a capsule whose refcount is already 0: handle it as a raw pointer (a py_object
# argument would incref/decref it and re-enter the deallocator)
_capsule_is_valid = ctypes.PYFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p)(("PyCapsule_IsValid", ctypes.pythonapi))
_capsule_get_pointer = ctypes.PYFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p)(("PyCapsule_GetPointer", ctypes.pythonapi))


class _DLManaged(ctypes.Structure):  # the head of lsio_dl_managed_tensor: enough to reach `deleter`
    _fields_ = [("data", ctypes.c_void_p), ("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("dtype_code", ctypes.c_uint8), ("dtype_bits", ctypes.c_uint8), ("dtype_lanes", ctypes.c_uint16),
                ("shape", ctypes.c_void_p), ("strides", ctypes.c_void_p), ("byte_offset", ctypes.c_uint64),
                ("manager_ctx", ctypes.c_void_p), ("deleter", ctypes.CFUNCTYPE(None, ctypes.c_void_p))]


@ctypes.PYFUNCTYPE(None, ctypes.c_void_p)
def _dl_capsule_destructor(capsule):
    # a consumer renames the capsule to "used_dltensor" and calls the deleter itself; unconsumed capsules are ours
    if _capsule_is_valid(capsule, b"dltensor"):
        p = _capsule_get_pointer(capsule, b"dltensor")
        _DLManaged.from_address(p).deleter(p)


_DL_CAPSULE_DESTRUCTOR = ctypes.cast(_dl_capsule_destructor, ctypes.c_void_p)


class AlignedBytes:
    """Shared, re-sliceable, immutable view (lib.rs:116-166); may live in host or device memory."""

    def __init__(self, handle):
        self._h = handle

    def clone(self):
        return AlignedBytes(N.lib.lsio_aligned_bytes_clone(self._h))

    def set_slice(self, start, end):
        if N.lib.lsio_aligned_bytes_set_slice(self._h, start, end):
            raise AlignedBytesError("set_slice: empty range or end beyond the underlying buffer")  # lib.rs:138-139
        return self

    def reset_slice(self):
        N.lib.lsio_aligned_bytes_reset_slice(self._h)
        return self

    def len(self):
        return N.lib.lsio_aligned_bytes_len(self._h)

    __len__ = len

    def as_ptr(self):
        return N.lib.lsio_aligned_bytes_as_ptr(self._h) or 0

    @property
    def memory_kind(self):
        return ("host", "pinned", "device")[N.lib.lsio_aligned_bytes_memory_kind(self._h)]

    @property
    def device(self):
        return N.lib.lsio_aligned_bytes_device(self._h)

    def as_slice(self):
        """The view's bytes as a numpy uint8 array (copied to the host for device views)."""
        n = self.len()
        out = np.empty(n, dtype=np.uint8)
        got = N.lib.lsio_aligned_bytes_copy_to_host(self._h, out.ctypes.data, n)
        if got != n:
            raise AlignedBytesError(f"copy_to_host failed: {got}")
        return out

    @property
    def __cuda_array_interface__(self):
        if self.memory_kind != "device":
            raise AttributeError("not a device view")
        return {"shape": (self.len(),), "typestr": "|u1", "data": (self.as_ptr(), False), "version": 3, "strides": None}

    # ---- DLPack (include/lsio_aligned_bytes.h lsio_aligned_bytes_to_dlpack) ----
    def __dlpack_device__(self):
        return {"device": (2, self.device), "pinned": (3, 0), "host": (1, 0)}[self.memory_kind]

    def __dlpack__(self, stream=None, dtype="uint8", shape=None, **_):
        """A "dltensor" capsule over the view (zero copy; the capsule owns its own share of the
        buffer).  The pack kernel's batch has retired before a chunk is handed out, so there is
        nothing to order `stream` after."""
        code, bits = _DL_DTYPES[str(dtype).replace("torch.", "")]
        shp = None if shape is None else (ctypes.c_int64 * len(shape))(*shape)
        out = ctypes.c_void_p()
        rc = N.lib.lsio_aligned_bytes_to_dlpack(self._h, code, bits, shp, 0 if shape is None else len(shape), ctypes.byref(out))
        if rc:
            raise AlignedBytesError(f"to_dlpack: the view ({self.len()} B at {self.as_ptr():#x}) does not fit dtype={dtype} shape={shape}")
        return _capsule_new(out.value, b"dltensor", _DL_CAPSULE_DESTRUCTOR)

    def to_torch(self, dtype="uint8", shape=None):
        """Zero-copy torch tensor over the view through DLPack: `dtype`/`shape` reinterpret the bytes
        (the element size must divide the view's address and length)."""
        import torch

        return torch.from_dlpack(self.__dlpack__(dtype=dtype, shape=shape))

    def drop(self):
        if self._h:
            N.lib.lsio_aligned_bytes_drop(self._h)
            self._h = None

    def __del__(self):
        if sys.is_finalizing():
            return  # interpreter exit: no native (CUDA) calls from the garbage collector; the process is going away
        try:
            self.drop()
        except Exception:
            pass
