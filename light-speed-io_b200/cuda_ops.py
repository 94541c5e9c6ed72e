"""Host-side handles for the lsio_cuda kernels (include/lsio_cuda.h).

PyTorch is used only as plumbing here (device memory and streams); the kernels are the
hand-written sm_100a ones in csrc/pack_crc.cu and csrc/misc_kernels.cu.
"""
import ctypes
import sys

import numpy as np

from . import _native as N


def _stream_ptr(stream):
    if stream is None:
        return None
    return ctypes.c_void_p(int(getattr(stream, "cuda_stream", stream)))


def _segments(src_addrs, dst_addrs, lens):
    n = len(lens)
    arr = (N.CudaSegment * n)()
    buf = np.frombuffer(arr, dtype=np.uint64).reshape(n, 3) if n else None
    if n:
        buf[:, 0] = np.asarray(src_addrs, dtype=np.uint64)
        buf[:, 1] = np.asarray(dst_addrs, dtype=np.uint64)
        buf[:, 2] = np.asarray(lens, dtype=np.uint64)
    return arr, n


class PackPlan:
    """Device-resident descriptor table for one pack_strip[_crc32c] launch."""

    def __init__(self, device, src_addrs, dst_addrs, lens, store=True, crc32c=True, light=False):
        flags = (N.PACK_STORE if store else 0) | (N.PACK_CRC32C if crc32c else 0) | (N.PACK_LIGHT if light else 0)
        segs, n = _segments(src_addrs, dst_addrs, lens)
        self._h = ctypes.c_void_p()
        N.check(N.lib.lsio_cuda_pack_plan_create(int(device), segs, n, flags, ctypes.byref(self._h)), "pack_plan_create")
        self.n = n
        self.device = int(device)

    def launch(self, crc_out=None, stream=None):
        ptr = ctypes.c_void_p(crc_out.data_ptr()) if crc_out is not None else None
        N.check(N.lib.lsio_cuda_pack_plan_launch(self._h, ptr, _stream_ptr(stream)), "pack_plan_launch")

    @property
    def algorithmic_bytes(self):
        return int(N.lib.lsio_cuda_pack_plan_algorithmic_bytes(self._h))

    @property
    def num_tiles(self):
        return int(N.lib.lsio_cuda_pack_plan_num_tiles(self._h))

    def close(self):
        if self._h:
            N.lib.lsio_cuda_pack_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        if sys.is_finalizing():
            return  # interpreter exit: no native (CUDA) calls from the garbage collector; the process is going away
        try:
            self.close()
        except Exception:
            pass


def pack_strip_crc32c(src, dst, src_off, dst_off, lens, store=True, crc32c=True, stream=None, light=False):
    """Gather src[src_off[i]:+len[i]] into dst[dst_off[i]:] (uint8 CUDA tensors) and return the
    per-segment CRC-32C as a uint32 CUDA tensor (or None).  light=True: the small-footprint kernel variant
    the pipeline uses for its batches (LSIO_CUDA_PACK_LIGHT)."""
    import torch

    dev = src.device.index
    base_s, base_d = src.data_ptr(), (dst.data_ptr() if dst is not None else 0)
    s = np.asarray(src_off, dtype=np.uint64) + np.uint64(base_s)
    d = np.asarray(dst_off, dtype=np.uint64) + np.uint64(base_d)
    plan = PackPlan(dev, s, d, lens, store=store, crc32c=crc32c, light=light)
    out = torch.empty(max(len(lens), 1), dtype=torch.int32, device=src.device) if crc32c else None
    try:
        plan.launch(out, stream)
        torch.cuda.synchronize(src.device)
    finally:
        plan.close()
    if out is None:
        return None
    return out[: len(lens)]


def pull_strip_crc32c(src_pinned, dst, src_off, dst_off, lens, crc32c=True, stream=None):
    """The pull form (lsio_cuda_pull_strip_crc32c): `src_pinned` is a pinned HOST uint8 tensor, read over PCIe by the
    kernel itself; dst a CUDA tensor.  At most 64 segments.  Returns the per-segment CRC-32C (uint32 CUDA tensor) or None."""
    import torch

    assert src_pinned.is_pinned() and not src_pinned.is_cuda
    s = np.asarray(src_off, dtype=np.uint64) + np.uint64(src_pinned.data_ptr())
    d = np.asarray(dst_off, dtype=np.uint64) + np.uint64(dst.data_ptr())
    segs, n = _segments(s, d, lens)
    out = torch.empty(max(n, 1), dtype=torch.int32, device=dst.device) if crc32c else None
    N.check(N.lib.lsio_cuda_pull_strip_crc32c(dst.device.index, segs, n, out.data_ptr() if crc32c else None, _stream_ptr(stream)), "pull_strip_crc32c")
    return out


def xxh64(buf, offs, lens, stream=None, variant=0):
    import torch

    n = len(lens)
    ptrs = (np.asarray(offs, dtype=np.uint64) + np.uint64(buf.data_ptr())).astype(np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint64)
    out = torch.empty(max(n, 1), dtype=torch.int64, device=buf.device)
    N.check(N.lib.lsio_cuda_xxh64_variant(buf.device.index, ptrs.ctypes.data_as(N.c_u64p), lens.ctypes.data_as(N.c_u64p), n,
                                          ctypes.c_void_p(out.data_ptr()), _stream_ptr(stream), int(variant)), "xxh64")
    return out[:n]


def verify_pattern(buf, offs, lens, file_ids, file_offs, stream=None):
    """Number of bytes in the spans that differ from the synthetic pattern byte(file_id, off)."""
    import torch

    n = len(lens)
    ptrs = (np.asarray(offs, dtype=np.uint64) + np.uint64(buf.data_ptr())).astype(np.uint64)
    arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in (ptrs, lens, file_ids, file_offs)]
    out = torch.zeros(1, dtype=torch.int64, device=buf.device)
    N.check(N.lib.lsio_cuda_verify_pattern(buf.device.index, *[a.ctypes.data_as(N.c_u64p) for a in arrs], n,
                                           ctypes.c_void_p(out.data_ptr()), _stream_ptr(stream)), "verify_pattern")
    return int(out.item())


def fill_pattern(buf, file_id, file_off=0, stream=None):
    """Fill a uint8 CUDA tensor with the synthetic content byte(file_id, file_off + i)."""
    N.check(N.lib.lsio_cuda_fill_pattern(buf.device.index, ctypes.c_void_p(buf.data_ptr()), buf.numel(), int(file_id), int(file_off),
                                         _stream_ptr(stream)), "fill_pattern")


def probe_read_ceiling(paths, block_bytes=1 << 20, n_threads=8, queue_depth=32, o_direct=True, inline_reads=0, files_in_flight=1):
    """GiB/s of raw reads over `paths` with no GPU work (stand-in for benches/fio.ini): READ_FIXED through io_uring, or
    -- inline_reads=1, the backend's own default rule -- preadv2 on the probing thread for files the kernel cannot read
    asynchronously (tmpfs)."""
    import os

    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    out = ctypes.c_double(0)
    N.check(N.lib.lsio_probe_read_ceiling_ex(arr, len(paths), block_bytes, n_threads, queue_depth, int(o_direct), int(inline_reads), int(files_in_flight), ctypes.byref(out)),
            "probe_read_ceiling")
    return out.value


def probe_staged_ceiling(paths, device=0, n_threads=8, slot_kib=512, slots_per_thread=4, use_copy_engine=False, cpu_first=0, cpu_count=0):
    """(seconds, bytes) of the minimal staged path over `paths`: preadv2 into pinned slots + pull kernel (or one cudaMemcpyAsync) per slot,
    nothing else (lsio_probe_staged_ceiling).  Ranks sharing a host call it between the same barriers and add up their bytes."""
    import os

    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    secs, nbytes = ctypes.c_double(0), ctypes.c_uint64(0)
    N.check(N.lib.lsio_probe_staged_ceiling(arr, len(paths), int(device), n_threads, slot_kib, slots_per_thread, int(use_copy_engine), cpu_first, cpu_count,
                                            ctypes.byref(secs), ctypes.byref(nbytes)), "probe_staged_ceiling")
    return secs.value, nbytes.value


def launch_count():
    return int(N.lib.lsio_cuda_launch_count())
