"""Host-side mirror of the codec layer (include/lsio_codecs.h): decode chunks where the I/O path
landed them, in HBM.  The reference only plans this layer (`lsio_codecs`, planned_design.md:83-90:
"These crates will all consume the output channel from the IO layer"; README.md:68-70); the codec
names below are the Zarr ecosystem's (numcodecs.LZ4, numcodecs.Shuffle, the Zarr v3 `crc32c` codec).

PyTorch is plumbing here (output tensors, streams); the work is done by the sm_100a kernels in
csrc/codecs.cu.  There is no CPU fallback.
"""
import ctypes
import sys

import numpy as np

from . import _native as N
from .cuda_ops import _stream_ptr

ITEM_DTYPE = np.dtype([("src", "<u8"), ("dst", "<u8"), ("src_len", "<u4"), ("dst_cap", "<u4")])
RESULT_DTYPE = np.dtype([("out_len", "<u4"), ("status", "<i4")])
CRC32C_RESIDUE = 0x48674BC7
STATUS_TEXT = {0: "ok", 1: "encoded stream truncated or overrun", 2: "decoded size differs from decoded_nbytes", 3: "bad LZ4 match offset",
               4: "crc32c trailer mismatch", 5: "not an Ok chunk on this device"}
STAGE_SHUFFLE, STAGE_LZ4, STAGE_CRC32C = 1, 2, 3

for _name, _res, _args in [
    ("lsio_codecs_lz4_decompress", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    ("lsio_codecs_lz4_decompress_async", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    ("lsio_codecs_unshuffle", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_void_p]),
    ("lsio_codecs_shuffle", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_void_p]),
    ("lsio_codecs_shuffle_tiles", ctypes.c_uint32, [ctypes.c_uint32]),
    ("lsio_codecs_shuffle_tile_map", ctypes.c_int64, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]),
    ("lsio_codecs_unshuffle_async", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                                   ctypes.c_int, ctypes.c_void_p]),
    ("lsio_codecs_crc32c_trailer_ok", ctypes.c_int, [ctypes.c_uint32]),
    ("lsio_codec_pipeline_create", ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32,
                                                  ctypes.POINTER(ctypes.c_void_p)]),
    ("lsio_codec_pipeline_decode", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    ("lsio_codec_pipeline_finish", ctypes.c_int64, [ctypes.c_void_p, ctypes.c_void_p]),
    ("lsio_codec_pipeline_last_ms", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    ("lsio_codec_pipeline_destroy", None, [ctypes.c_void_p]),
]:
    getattr(N.lib, _name).restype = _res
    getattr(N.lib, _name).argtypes = _args


class CodecError(Exception):
    pass


def make_items(src, dst, src_len, dst_cap=0):
    """An lsio_codec_item array from per-chunk device addresses and sizes (arrays or scalars)."""
    n = len(src)
    items = np.zeros(n, dtype=ITEM_DTYPE)
    items["src"], items["dst"], items["src_len"], items["dst_cap"] = src, dst, src_len, dst_cap
    return items


def lz4_decompress(items, device=0, stream=None):
    """LZ4 block decode of every item in one launch; returns the lsio_codec_result array."""
    items = np.ascontiguousarray(items, dtype=ITEM_DTYPE)
    res = np.zeros(len(items), dtype=RESULT_DTYPE)
    N.check(N.lib.lsio_codecs_lz4_decompress(int(device), items.ctypes.data, len(items), res.ctypes.data, _stream_ptr(stream)), "lz4_decompress")
    return res


def unshuffle(items, elem_size, device=0, stream=None):
    items = np.ascontiguousarray(items, dtype=ITEM_DTYPE)
    N.check(N.lib.lsio_codecs_unshuffle(int(device), items.ctypes.data, len(items), int(elem_size), _stream_ptr(stream)), "unshuffle")


def shuffle(items, elem_size, device=0, stream=None):
    items = np.ascontiguousarray(items, dtype=ITEM_DTYPE)
    N.check(N.lib.lsio_codecs_shuffle(int(device), items.ctypes.data, len(items), int(elem_size), _stream_ptr(stream)), "shuffle")


def shuffle_tile_map(src_lens):
    """The {item, tile} table lsio_codecs_unshuffle_async wants, as an (n_tiles, 2) uint32 array."""
    lens = np.ascontiguousarray(src_lens, dtype=np.uint32)
    n_tiles = N.lib.lsio_codecs_shuffle_tile_map(lens.ctypes.data, len(lens), None, 0)
    if n_tiles < 0:
        raise CodecError(f"shuffle_tile_map: {n_tiles}")
    tm = np.zeros((n_tiles, 2), dtype=np.uint32)
    N.lib.lsio_codecs_shuffle_tile_map(lens.ctypes.data, len(lens), tm.ctypes.data, n_tiles)
    return tm


def crc32c_trailer_ok(crc32c_of_chunk_with_trailer):
    return bool(N.lib.lsio_codecs_crc32c_trailer_ok(int(crc32c_of_chunk_with_trailer) & 0xFFFFFFFF))


# ---- codec descriptions, named after their numcodecs / Zarr v3 counterparts ----------------------------
class Crc32c:
    """Zarr v3 `crc32c` codec: the stored chunk ends with the CRC-32C of what precedes it (4 LE bytes).
    Decoding verifies it from the CRC the pack kernel fused into the landing -- no pass over the data --
    and drops the 4 bytes."""


class LZ4:
    """numcodecs.LZ4: `header` bytes (4: the decoded size, little-endian) then one LZ4 block."""

    def __init__(self, header=4):
        self.header = header


class Shuffle:
    """numcodecs.Shuffle(elementsize): byte planes back to elements."""

    def __init__(self, elementsize=4):
        self.elementsize = elementsize


class CodecPipeline:
    """The decode side of a Zarr codec chain, applied to a whole batch of landed chunks at once.

    `codecs` is in ENCODE order as Zarr metadata lists it, e.g. [Shuffle(4), LZ4(), Crc32c()] (filters,
    compressor, checksum); decode runs it backwards.  Every chunk decodes to `decoded_nbytes` bytes (a Zarr
    array's chunk shape is fixed), so the result is one (n_chunks, decoded_nbytes) uint8 tensor in HBM,
    chunk i in row i -- ready to be viewed as the array dtype and stacked into a training batch."""

    def __init__(self, codecs, decoded_nbytes, device=0, max_batch=None):
        """`max_batch`: size the native pipeline (pinned descriptor arrays, HBM scratch of max_batch x decoded_nbytes) once,
        up front; without it the pipeline is sized by the first batch and re-created -- tens to hundreds of milliseconds of
        pinned and device allocation -- whenever a larger batch arrives."""
        self.codecs, self.decoded_nbytes, self.device = list(codecs), int(decoded_nbytes), int(device)
        self._want_batch = int(max_batch) if max_batch else 0
        if self.decoded_nbytes % 16:
            raise ValueError("decoded_nbytes must be a multiple of 16 (rows are addressed as 16-byte vectors)")

    def decode(self, chunks, stream=None):
        """chunks: lsio Chunk items (device views, kept alive by the caller until this returns).
        Returns the decoded (n, decoded_nbytes) uint8 tensor; raises CodecError naming the first bad chunk."""
        src = np.array([c.buffer.as_ptr() for c in chunks], dtype=np.uint64)
        src_len = np.array([c.buffer.len() for c in chunks], dtype=np.int64)
        if any(isinstance(c, Crc32c) for c in self.codecs) and any(c.crc32c is None for c in chunks):
            raise CodecError("Crc32c needs the backend's fused CRC-32C (checksum=CHECKSUM_CRC32C)")
        crcs = np.array([c.crc32c or 0 for c in chunks], dtype=np.uint32)
        return self._decode(src, src_len, crcs, np.array([c.user_data for c in chunks], dtype=np.uint64), stream)

    # ---- record arrays (Receiver.recv_batch): the native pipeline, lsio_codec_pipeline_* ----
    def _native(self, n):
        if getattr(self, "_h", None) is None or n > self._max_batch:
            self.close()
            stages = []
            for c in self.codecs:
                if isinstance(c, Shuffle):
                    stages.append((STAGE_SHUFFLE, c.elementsize))
                elif isinstance(c, LZ4):
                    stages.append((STAGE_LZ4, c.header))
                elif isinstance(c, Crc32c):
                    stages.append((STAGE_CRC32C, 0))
                else:
                    raise TypeError(f"unknown codec {c!r}")
            arr = np.array(stages, dtype=np.uint32).reshape(-1, 2)
            h = ctypes.c_void_p()
            # room for this batch, rounded up to a power of two (later, slightly larger batches reuse it); the scratch
            # buffer is max_batch x decoded_nbytes of HBM, so it must follow the batch size, not a fixed count
            self._max_batch = max(64, self._want_batch, 1 << (max(int(n), 1) - 1).bit_length())
            N.check(N.lib.lsio_codec_pipeline_create(self.device, arr.ctypes.data, len(stages), self.decoded_nbytes, self._max_batch, ctypes.byref(h)),
                    "codec_pipeline_create (encode order must be Shuffle, LZ4, Crc32c)")
            self._h = h
        return self._h

    def submit_records(self, records, out=None, stream=None):
        """Enqueue the decode of a record array from Receiver.recv_batch (the C `lsio_output` layout) on `stream`
        and return the (n, decoded_nbytes) uint8 output tensor at once: no per-chunk Python work, no host
        synchronisation.  Call finish() before using the rows or dropping the records' buffers; one decode in
        flight per pipeline (use two pipelines to overlap decoding with receiving)."""
        import torch

        records = np.ascontiguousarray(records)
        n = len(records)
        h = self._native(n)  # first: without a usable device this fails loudly here (no CPU fallback)
        if out is None:
            out = torch.empty((n, self.decoded_nbytes), dtype=torch.uint8, device=torch.device("cuda", self.device))
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        self._pending = (records, out)
        N.check(N.lib.lsio_codec_pipeline_decode(h, records.ctypes.data, n, out.data_ptr(), _stream_ptr(stream)), "codec_pipeline_decode")
        return out

    def finish(self):
        """Wait for submit_records' decode; raises CodecError naming the first bad chunk."""
        if getattr(self, "_pending", None) is None:
            return
        records, _ = self._pending
        self._pending = None
        status = np.zeros(len(records), dtype=np.int32)
        bad = N.lib.lsio_codec_pipeline_finish(self._h, status.ctypes.data)
        if bad < 0:
            N.check(int(-bad), "codec_pipeline_finish")
        if bad:
            i = int(np.flatnonzero(status)[0])
            raise CodecError(f"{bad} bad chunk(s); first: chunk {i} (user_data={int(records['chunk']['user_data'][i])}): "
                             f"{STATUS_TEXT.get(int(status[i]), status[i])}")

    def last_ms(self):
        """(first stage, rest) device milliseconds of the decode finished last (measurement aid)."""
        a, b = ctypes.c_float(), ctypes.c_float()
        N.check(N.lib.lsio_codec_pipeline_last_ms(self._h, ctypes.byref(a), ctypes.byref(b)), "codec_pipeline_last_ms")
        return a.value, b.value

    def decode_records(self, records, stream=None):
        """submit_records + finish."""
        out = self.submit_records(records, stream=stream)
        self.finish()
        return out

    def close(self):
        if getattr(self, "_h", None) is not None:
            N.lib.lsio_codec_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        if sys.is_finalizing():
            return  # interpreter exit: no native (CUDA) calls from the garbage collector; the process is going away
        try:
            self.close()
        except Exception:
            pass

    def _decode(self, src, src_len, crcs, user_data, stream):
        import torch

        n = len(src)
        dev = torch.device("cuda", self.device)
        cur = None  # tensor holding the current stage's rows (None: still the chunk buffers themselves)
        for codec in reversed(self.codecs):
            if isinstance(codec, Crc32c):
                bad = np.flatnonzero((crcs != CRC32C_RESIDUE) | (src_len < 4))
                if len(bad):
                    i = int(bad[0])
                    raise CodecError(f"chunk {i} (user_data={int(user_data[i])}): crc32c trailer mismatch")
                src_len = src_len - 4
            elif isinstance(codec, LZ4):
                out = torch.empty((n, self.decoded_nbytes), dtype=torch.uint8, device=dev)
                dst = out.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(self.decoded_nbytes)
                if np.any(src_len < codec.header):
                    raise CodecError("chunk shorter than its LZ4 header")
                items = make_items(src + np.uint64(codec.header), dst, src_len - codec.header, self.decoded_nbytes)
                res = lz4_decompress(items, self.device, stream)
                bad = np.flatnonzero((res["status"] != 0) | (res["out_len"] != self.decoded_nbytes))
                if len(bad):
                    i = int(bad[0])
                    raise CodecError(f"chunk {i} (user_data={int(user_data[i])}): lz4 {STATUS_TEXT.get(int(res['status'][i]), '?')}, "
                                     f"{int(res['out_len'][i])} of {self.decoded_nbytes} bytes")
                cur, src = out, dst
                src_len = np.full(n, self.decoded_nbytes, dtype=np.int64)
            elif isinstance(codec, Shuffle):
                if np.any(src_len != self.decoded_nbytes):
                    raise CodecError("Shuffle input is not decoded_nbytes long")
                out = torch.empty((n, self.decoded_nbytes), dtype=torch.uint8, device=dev)
                dst = out.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(self.decoded_nbytes)
                unshuffle(make_items(src, dst, src_len), codec.elementsize, self.device, stream)
                cur, src = out, dst
            else:
                raise TypeError(f"unknown codec {codec!r}")
        if cur is None:  # no data-moving stage: gather the (verified) chunks into rows with the pack kernel
            from .cuda_ops import PackPlan

            cur = torch.empty((n, self.decoded_nbytes), dtype=torch.uint8, device=dev)
            if np.any(src_len != self.decoded_nbytes):
                raise CodecError("stored chunk size differs from decoded_nbytes")
            dst = cur.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(self.decoded_nbytes)
            plan = PackPlan(self.device, src, dst, src_len.astype(np.uint64), store=True, crc32c=False)
            plan.launch(None, stream)
            torch.cuda.synchronize(dev)
            plan.close()
        return cur
