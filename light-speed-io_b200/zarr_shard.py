"""A caller of the path: reading inner chunks of Zarr v3 *sharded* arrays (the `sharding_indexed` codec) into HBM.

The reference names Zarr as the first format it wants to serve ("speed up training ML models from sharded Zarr
datasets", planned_design.md:52; `lsio_zarr`, planned_design.md:100-104) and shaped its API for it: negative
ranges exist so that a shard's index can be read from the end of the file (lsio_io/src/lib.rs:19-23, "-100..-1:
the last 100 bytes"), `user_data` "would be used to identify the location at which this chunk appears in the
merged array" (lib.rs:25-28), and groups order "indexes first, chunks second" (lib.rs:6-8).  No Zarr code exists
in the reference; the layout below is the Zarr v3 sharding codec's (ZEP 2):

    shard file = inner chunk bytes (any order, any gaps) ++ index
    index      = n_inner x (offset: u64 LE, nbytes: u64 LE), C order over the inner-chunk grid,
                 (2^64-1, 2^64-1) marks a chunk that was never written;
                 with index_codecs [bytes, crc32c] (the default) it is followed by its CRC-32C, 4 bytes LE;
                 index_location "end" (default) or "start".

Everything here is host-side bookkeeping over the C ABI; bytes move only through get_ranges.
"""
import numpy as np

from . import codecs as C

MISSING = np.uint64(0xFFFFFFFFFFFFFFFF)


class ShardIndexError(Exception):
    pass


def index_nbytes(n_inner, crc32c=True):
    return 16 * int(n_inner) + (4 if crc32c else 0)


def index_range(n_inner, crc32c=True, location="end"):
    """The byte range of a shard's index in get_ranges terms (negative = relative to the end of the file)."""
    n = index_nbytes(n_inner, crc32c)
    return (-n, -1) if location == "end" else (0, n)


def parse_index(chunk, n_inner, crc32c=True):
    """(offsets, nbytes) uint64 arrays from the landed index chunk.  With `crc32c` the index is verified from the
    CRC the pack kernel fused into the landing (Zarr v3 crc32c codec: the residue test, no second pass)."""
    if not chunk.is_ok():
        raise ShardIndexError(f"index read failed: {chunk}")
    want = index_nbytes(n_inner, crc32c)
    if chunk.buffer.len() != want:
        raise ShardIndexError(f"index is {chunk.buffer.len()} bytes, expected {want}")
    if crc32c and not (chunk.crc32c is not None and C.crc32c_trailer_ok(chunk.crc32c)):
        raise ShardIndexError("shard index crc32c mismatch")
    raw = chunk.buffer.as_slice()  # a few KiB: the one device-to-host copy of the read
    table = raw[:16 * n_inner].view("<u8").reshape(n_inner, 2)
    return table[:, 0].copy(), table[:, 1].copy()


def read_shards(uring, paths, n_inner, wanted=None, index_crc32c=True, index_location="end", timeout=30.0):
    """Two ordered groups over `uring` (an IoUringCuda): group g = every shard's index, group g+1 = the inner
    chunks they point to.  `wanted[i]` lists the inner-chunk numbers to fetch from shard i (default: all present).
    Submits everything and returns (n_chunks_submitted, present) where present[i] is the array of inner-chunk
    numbers of shard i that exist and were requested; their chunks then arrive on uring.completion() with
    user_data = shard_number << 32 | inner_chunk_number."""
    rx = uring.completion()
    for i, p in enumerate(paths):
        uring.get_ranges(p, [index_range(n_inner, index_crc32c, index_location)], [i])
    uring.next_group()  # chunk reads start only after every index has been delivered
    tables, first_error = {}, None
    for _ in paths:  # every index item is received and dropped even if one of them is bad: nothing is left on the backend
        c = rx.recv_timeout(timeout)
        try:
            tables[c.user_data] = parse_index(c, n_inner, index_crc32c)
        except ShardIndexError as ex:
            first_error = first_error or ex
        finally:
            if c.is_ok():
                c.buffer.drop()
    if first_error is not None:
        uring.next_group()
        raise first_error
    total, present = 0, []
    for i, p in enumerate(paths):
        off, nb = tables[i]
        sel = np.arange(n_inner) if wanted is None else np.asarray(wanted[i], dtype=np.int64)
        sel = sel[(off[sel] != MISSING) | (nb[sel] != MISSING)]
        if np.any(nb[sel] == 0):
            uring.next_group()
            raise ShardIndexError(f"{p}: the index holds a zero-length entry for a chunk that is not marked missing")
        present.append(sel)
        if len(sel):
            uring.get_ranges_arrays(p, off[sel].astype(np.int64), (off[sel] + nb[sel]).astype(np.int64), (np.uint64(i) << np.uint64(32)) | sel.astype(np.uint64))
            total += len(sel)
    uring.next_group()
    return total, present
