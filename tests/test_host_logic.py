"""Host-side logic of the product library, no GPU needed: the C ABI loads and exports every
symbol include/*.h declares, and the lsio_aligned_bytes / lsio_threadpool mirrors behave like
the reference's own unit tests and doctests."""
import ctypes
import glob
import os
import re
import threading
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\b(lsio_[a-z0-9_]+)\s*\(", src):
            if "typedef" in src[max(0, src.rfind("\n", 0, m.start()) + 1):m.start()]:
                continue
            names.add(m.group(1))
    return sorted(names)


def test_c_abi_exports_every_declared_symbol(lsio):
    lib = ctypes.CDLL(lsio._native.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 45, declared
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing


def test_backend_fails_loudly_without_a_gpu(lsio):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lsio._native.LsioCudaError, match="no CPU fallback"):
        lsio.IoUringCuda(2)


@pytest.mark.parametrize("bad,what", [
    (dict(direct_landing=3), "direct_landing"), (dict(inline_reads=3), "inline_reads"), (dict(copy_coalesce_slots=9), "copy_coalesce_slots"),
    (dict(force_align=1000), "force_align"), (dict(force_align=256), "force_align"), (dict(force_align=1 << 20), "force_align"),
    (dict(host_slot_kib=8192, staging_bytes=4 << 20), "staging_bytes / 2"), (dict(host_slots_per_worker=300), "host_slots_per_worker"),
    (dict(sq_entries=48), "invalid"), (dict(max_read_bytes=100), "invalid"), (dict(pack_threads=100), "invalid")])
def test_config_is_validated_before_any_device_is_touched(lsio, bad, what):
    """Round-2 knobs: values that cannot work are refused with EINVAL and a message that names the field -- with or without a GPU
    (validation comes first); a slot larger than half the staging ring used to stall every read silently."""
    with pytest.raises(lsio._native.LsioCudaError, match=what):
        lsio.IoUringCuda(2, **bad)


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or load it."""
    pkg = os.path.join(ROOT, "light-speed-io_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liblsio_oracle|oracle/_|#include\s+\"lsio_oracle", txt, flags=re.M), (dirpath, f)


# ---- lsio_aligned_bytes: src/lib.rs:215-251 (unit test) ----------------------------------------
def test_write_and_read(lsio):
    LEN = 16
    b1, b2 = lsio.AlignedBytesMut.new(LEN, 8), lsio.AlignedBytesMut.new(LEN, 8)
    b1.as_numpy()[:] = np.arange(LEN, dtype=np.uint8)
    b2.as_numpy()[:] = np.arange(LEN, dtype=np.uint8)
    assert b1.as_mut_ptr() % 8 == 0 and b2.as_mut_ptr() % 8 == 0
    s1, s2 = b1.freeze(), b2.freeze()
    assert s1.as_slice().tolist() == list(range(16)) == s2.as_slice().tolist()


# ---- README.md:61-102 doctest 1: freeze -> reset_slice -> clone + set_slice ---------------------
def test_doctest_merged_read_split_into_owned_views(lsio):
    LEN, ALIGN = 8192, 512
    m = lsio.AlignedBytesMut.new(LEN, ALIGN)
    assert m.as_mut_ptr() % ALIGN == 0
    m.as_numpy()[:] = np.arange(LEN, dtype=np.uint64).astype(np.uint8)
    expected = np.arange(LEN, dtype=np.uint64).astype(np.uint8)
    b = m.freeze()
    b.reset_slice()
    assert np.array_equal(b.as_slice(), expected)
    b0 = b.clone().set_slice(0, 4096)
    b1 = b.clone().set_slice(4096, 8192)
    assert b0.len() == 4096 and b1.len() == 4096
    assert np.array_equal(b0.as_slice(), expected[:4096]) and np.array_equal(b1.as_slice(), expected[4096:])
    assert np.array_equal(b.as_slice(), expected)
    b.drop()  # the two views outlive the original
    assert np.array_equal(b0.as_slice(), expected[:4096]) and np.array_equal(b1.as_slice(), expected[4096:])


# ---- README.md:120-168 doctest 2: split_to at 2/4/6 MiB, freeze only when sole owner --------------
def test_doctest_split_large_read(lsio):
    MiB = 1 << 20
    LEN, ALIGN = 8 * MiB, 512
    b3 = lsio.AlignedBytesMut.new(LEN, ALIGN)
    b0 = b3.split_to(2 * MiB)
    assert b0.len() == 2 * MiB and b3.len() == 6 * MiB
    b1 = b3.split_to(4 * MiB)
    b2 = b3.split_to(6 * MiB)
    assert [x.len() for x in (b0, b1, b2, b3)] == [2 * MiB] * 4
    for i, x in enumerate((b0, b1, b2, b3)):
        x.as_numpy()[:] = i
    with pytest.raises(lsio.AlignedBytesError):
        b3.freeze()  # other AlignedBytesMut still reference the buffer (lib.rs:103-112)
    b0.drop(), b1.drop(), b2.drop()
    b = b3.freeze()
    b.reset_slice()
    assert np.array_equal(b.as_slice(), np.repeat(np.arange(4, dtype=np.uint8), 2 * MiB))


def test_split_to_errors_match_the_reference_order(lsio):
    m = lsio.AlignedBytesMut.new(4096, 512)
    hi = m.split_to(2048)  # hi = [0, 2048), m = [2048, 4096)
    with pytest.raises(lsio.AlignedBytesError, match="not contained"):
        m.split_to(1024)   # lib.rs:75-79
    with pytest.raises(lsio.AlignedBytesError, match="not contained"):
        m.split_to(4096)   # end is exclusive
    with pytest.raises(lsio.AlignedBytesError, match="zero"):
        hi.split_to(0)     # lib.rs:80-81
    with pytest.raises(lsio.AlignedBytesError, match="divisible"):
        m.split_to(2148)   # lib.rs:82-86: this is what kills unaligned starts in the reference (F2)
    assert m.len() == 2048 and hi.len() == 2048


def test_alloc_rules_and_padding(lsio):
    with pytest.raises(lsio.AlignedBytesError):
        lsio.AlignedBytesMut.new(0, 512)   # assert_ne!(len, 0), lib.rs:171
    with pytest.raises(lsio.AlignedBytesError):
        lsio.AlignedBytesMut.new(10, 3)    # Layout needs a power of two
    b = lsio.AlignedBytesMut.new(5000, 512)
    assert b.len() == 5000                 # len() is the requested length (lib.rs:35-37) ...
    a = b.freeze()
    a.reset_slice()
    assert a.len() == 5120                 # ... the allocation is padded (lib.rs:178-181)
    with pytest.raises(lsio.AlignedBytesError):
        a.set_slice(10, 10)                # assert!(!range.is_empty()), lib.rs:138
    with pytest.raises(lsio.AlignedBytesError):
        a.set_slice(0, 5121)               # assert!(range.end <= buf.len()), lib.rs:139


# ---- lsio_threadpool: threadpool.rs:186-268 ------------------------------------------------------
def test_threadpool(lsio):
    N = lsio._native
    N_THREADS, MULTIPLIER = 4, 8
    N_TASKS = N_THREADS * MULTIPLIER
    lock = threading.Lock()
    outputs, per_thread = [], {}

    @N.TASK_FN
    def on_task(ctx, worker, task):
        with lock:
            outputs.append(task)
            per_thread[worker] = per_thread.get(worker, 0) + 1
        time.sleep(1e-6)  # give other threads a chance (threadpool.rs:214-216)
        return 0

    pool = N.lib.lsio_threadpool_new(N_THREADS, on_task, None)
    for i in range(N_TASKS):
        if i % N_THREADS == 0:
            time.sleep(0.01)  # let the workers park, to check they wake up again
        N.lib.lsio_threadpool_push(pool, i + 1)
    deadline = time.time() + 5
    while time.time() < deadline:
        with lock:
            if len(outputs) == N_TASKS:
                break
        time.sleep(0.001)
    time.sleep(0.02)
    parks = [N.lib.lsio_threadpool_n_parks(pool, w) for w in range(N_THREADS)]
    N.lib.lsio_threadpool_free(pool)  # drop(pool): joins
    assert sorted(outputs) == list(range(1, N_TASKS + 1))  # every task exactly once
    assert len(per_thread) == N_THREADS and sum(per_thread.values()) == N_TASKS
    # the reference asserts >= 2 tasks per thread and 8 or 9 parks; timing on a loaded CI box is
    # looser here: every thread worked and parked at least once per burst it was woken for
    assert all(n >= 1 for n in per_thread.values()), per_thread
    assert all(p >= 1 for p in parks), parks


def test_threadpool_follow_up_tasks_are_stolen(lsio):
    """WorkerThread::push puts tasks on the local deque where peers steal them (worker.rs:91-101,
    get_ranges.rs:39-45: one GetRanges fans out into per-range reads)."""
    N = lsio._native
    lock = threading.Lock()
    seen, workers = [], set()

    @N.TASK_FN
    def on_task(ctx, worker, task):
        with lock:
            seen.append(task)
            workers.add(worker)
        time.sleep(0.002)
        return task - 1 if task > 1 and task < 1000 else 0  # chain: 200 -> 199 -> ... -> 1

    pool = N.lib.lsio_threadpool_new(4, on_task, None)
    for base in (200, 150, 100, 50):
        N.lib.lsio_threadpool_push(pool, base)
    deadline = time.time() + 10
    while time.time() < deadline:
        with lock:
            if len(seen) == 500:
                break
        time.sleep(0.005)
    N.lib.lsio_threadpool_free(pool)
    assert len(seen) == 500 and len(workers) >= 2


# ---- sharding: SURVEY 8e ----------------------------------------------------------------------------
def test_files_for_rank_is_a_partition(lsio):
    from lsio_b200 import sharding

    for world in (1, 2, 4, 8):
        parts = [sharding.files_for_rank(500, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(500))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    # affinity: files on node 1 only go to GPUs on node 1
    gpu_nodes = [0, 0, 1, 1]
    fn = [i % 2 for i in range(100)]
    for r in range(4):
        for i in sharding.files_for_rank(100, r, 4, fn, gpu_nodes):
            assert gpu_nodes[r] == fn[i]


# ---- the shipped C++ host components, restating the reference's unit tests (csrc/selftest.cpp) ----------
@pytest.mark.parametrize("which,name", [(1, "tracker (tracker.rs:74-114)"), (2, "user_data (user_data.rs:47-56)"),
                                        (3, "bounded channel: back-pressure, batch receive, disconnect"),
                                        (4, "staging ring as a log: wrap, no-straddle, in-order retire"),
                                        (5, "HBM arena: first fit + coalescing"), (6, "crc32c_combine / GF(2) inverse"),
                                        (7, "ordered groups of operations (lsio_io lib.rs:6-8): admission, hold, teardown, threads")])
def test_native_host_components(lsio, which, name):
    rc = lsio._native.lib.lsio_uring_cuda_selftest(which)
    assert rc == 0, f"{name}: check {rc % 100} failed"


def test_dlpack_export_of_host_views(lsio):
    """SURVEY 8f rank 3 (the Python/PyTorch surface the reference only plans, README.md:74,85,103):
    lsio_aligned_bytes_to_dlpack hands a view to torch / numpy without a copy; typed and shaped
    reinterpretations are checked against size and address."""
    import torch

    m = lsio.AlignedBytesMut(8192, 512)
    src = np.arange(8192, dtype=np.uint32).astype(np.uint8)
    ctypes.memmove(m.as_mut_ptr(), src.ctypes.data, 8192)
    a = m.freeze()
    assert a.__dlpack_device__() == (1, 0)
    t = torch.from_dlpack(a)
    assert t.dtype == torch.uint8 and t.shape == (8192,) and t.data_ptr() == a.as_ptr()  # zero copy
    assert np.array_equal(t.numpy(), src)
    assert np.array_equal(np.from_dlpack(a), src)
    f = a.to_torch("float32", (4, 16, 32))
    assert f.dtype == torch.float32 and f.shape == (4, 16, 32) and f.data_ptr() == a.as_ptr()
    assert np.array_equal(f.numpy().view(np.uint8).ravel(), src)
    a.set_slice(512, 512 + 4 * 100)
    assert a.to_torch("int32").shape == (100,)
    with pytest.raises(lsio.AlignedBytesError):
        a.to_torch("float32", (33, 3))  # size mismatch
    a.set_slice(513, 517)
    with pytest.raises(lsio.AlignedBytesError):
        a.to_torch("int32")  # element size does not divide the address
    a.set_slice(0, 3)
    with pytest.raises(lsio.AlignedBytesError):
        a.to_torch("uint16")  # nor the length
    a.reset_slice()
    cap = a.__dlpack__()  # an unconsumed capsule releases its share itself
    del cap
    a.drop()  # tensors keep the buffer alive after the view is gone
    assert int(t[4097]) == src[4097] and np.array_equal(f.numpy().view(np.uint8).ravel(), src)


def test_codec_layer_host_side_without_a_gpu(lsio):
    """The parts of the codec layer that are pure host code, and its loud failure without a device."""
    import torch

    C = lsio.codecs
    # tile map: one {item, tile} pair per 16 KiB tile, in launch order; empty buffers take no tile
    tm = C.shuffle_tile_map([0, 1, 16384, 16385, 50000])
    assert tm.tolist() == [[1, 0], [2, 0], [3, 0], [3, 1], [4, 0], [4, 1], [4, 2], [4, 3]]
    assert lsio._native.lib.lsio_codecs_shuffle_tiles(0) == 0 and lsio._native.lib.lsio_codecs_shuffle_tiles(16384 * 3) == 3
    assert C.ITEM_DTYPE.itemsize == 24 and C.RESULT_DTYPE.itemsize == 8  # lsio_codec_item / lsio_codec_result layouts
    assert C.crc32c_trailer_ok(C.CRC32C_RESIDUE) and not C.crc32c_trailer_ok(0)
    items = C.make_items([16], [32], [4], [8])
    assert items["src"][0] == 16 and items["dst"][0] == 32 and items["src_len"][0] == 4 and items["dst_cap"][0] == 8
    with pytest.raises(ValueError):
        C.CodecPipeline([C.LZ4()], 1000)  # rows are addressed as 16-byte vectors
    if not torch.cuda.is_available():
        with pytest.raises(lsio._native.LsioCudaError):  # no CPU fallback: the kernels are the product
            C.lz4_decompress(items)
        with pytest.raises(lsio._native.LsioCudaError):
            C.unshuffle(items, 4)
        with pytest.raises(lsio._native.LsioCudaError):
            C.CodecPipeline([C.Shuffle(4), C.LZ4(), C.Crc32c()], 4096).decode_records(np.zeros(0, dtype=np.dtype(lsio._native.COutput)))


def test_zarr_shard_index_parsing(lsio, oracle):
    """lsio_b200.zarr_shard: the index range in get_ranges terms and the index parser (Zarr v3 sharding codec), on a
    host view standing in for the landed chunk."""
    import struct

    Z = lsio.zarr_shard
    assert Z.index_nbytes(10) == 164 and Z.index_nbytes(10, crc32c=False) == 160
    assert Z.index_range(10) == (-164, -1) and Z.index_range(10, location="start") == (0, 164)
    table = np.array([[100, 40], [0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF], [140, 7]], dtype="<u8")
    raw = table.tobytes() + struct.pack("<I", oracle.crc32c(table.tobytes()))

    class FakeChunk:  # what Receiver.recv returns, over host memory
        def __init__(self, data, crc):
            m = lsio.AlignedBytesMut(len(data), 512)
            ctypes.memmove(m.as_mut_ptr(), data, len(data))
            self.buffer = m.freeze().set_slice(0, len(data))
            self.crc32c, self.user_data = crc, 0

        def is_ok(self):
            return True

    off, nb = Z.parse_index(FakeChunk(raw, oracle.crc32c(raw)), 3)
    assert off.tolist() == [100, 0xFFFFFFFFFFFFFFFF, 140] and nb.tolist() == [40, 0xFFFFFFFFFFFFFFFF, 7]
    bad = bytearray(raw)
    bad[5] ^= 1
    with pytest.raises(Z.ShardIndexError, match="crc32c"):
        Z.parse_index(FakeChunk(bytes(bad), oracle.crc32c(bytes(bad))), 3)
    with pytest.raises(Z.ShardIndexError, match="bytes, expected"):
        Z.parse_index(FakeChunk(raw, oracle.crc32c(raw)), 4)
    off, nb = Z.parse_index(FakeChunk(table.tobytes(), None), 3, crc32c=False)
    assert off.tolist()[0] == 100


def test_codec_pipeline_argument_checks(lsio):
    """lsio_codec_pipeline_create refuses bad arguments before it touches the device."""
    lib = lsio._native.lib
    C = lsio.codecs

    def create(stages, nbytes=4096, max_batch=16):
        arr = np.array(stages, dtype=np.uint32).reshape(-1, 2)
        h = ctypes.c_void_p()
        rc = lib.lsio_codec_pipeline_create(0, arr.ctypes.data if len(stages) else None, len(stages), nbytes, max_batch, ctypes.byref(h))
        if rc == 0:
            lib.lsio_codec_pipeline_destroy(h)
        return rc

    assert create([(C.STAGE_SHUFFLE, 4), (C.STAGE_LZ4, 4)], nbytes=1000) == -22  # rows are 16-byte vectors
    assert create([(C.STAGE_SHUFFLE, 4)], max_batch=0) == -22
    assert create([(C.STAGE_SHUFFLE, 4)], nbytes=0) == -22
    import torch

    if not torch.cuda.is_available():  # valid arguments get as far as the device check, and stop there: no CPU fallback
        assert create([(C.STAGE_SHUFFLE, 4), (C.STAGE_LZ4, 4), (C.STAGE_CRC32C, 0)]) > 0
    # (stage order and the shuffle element size are checked after the device count: covered by the GPU tests)


def _write_zarr_json(path, **over):
    import json

    meta = {"zarr_format": 3, "node_type": "array", "shape": [64, 96], "data_type": "float32", "fill_value": 7.5,
            "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [32, 48]}},
            "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
            "codecs": [{"name": "sharding_indexed", "configuration": {
                "chunk_shape": [8, 16], "index_location": "end",
                "index_codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "crc32c"}],
                "codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "numcodecs.shuffle", "configuration": {"elementsize": 4}},
                           {"name": "numcodecs.lz4", "configuration": {"acceleration": 1}}, {"name": "crc32c"}]}}]}
    meta.update(over)
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "zarr.json"), "w") as f:
        json.dump(meta, f)
    return meta


def test_zarr_v3_sharded_array_metadata(lsio, tmp_path):
    """lsio_b200.zarr_v3.ShardedArray: zarr.json of a sharded array -> shapes, codec chain, coordinate arithmetic;
    anything it cannot decode is refused by name."""
    ZV, C = lsio.zarr_v3, lsio.codecs
    d = str(tmp_path / "a.zarr")
    _write_zarr_json(d)
    a = ZV.ShardedArray(d)
    assert a.shape == (64, 96) and a.shard_shape == (32, 48) and a.inner_shape == (8, 16) and a.dtype == np.dtype("<f4")
    assert a.inner_per_shard == (4, 3) and a.n_inner == 12 and a.inner_nbytes == 8 * 16 * 4 and a.index_crc32c and a.fill_value == 7.5
    assert [type(c) for c in a.pipeline_codecs] == [C.Shuffle, C.LZ4, C.Crc32c] and a.pipeline_codecs[0].elementsize == 4
    shard, within = a.locate([(0, 0), (3, 2), (4, 3), (7, 5), (5, 1)])
    assert shard.tolist() == [[0, 0], [0, 0], [1, 1], [1, 1], [1, 0]] and within.tolist() == [0, 11, 0, 11, 4]  # C order over the 4 x 3 inner grid
    assert a.shard_path((1, 0)) == os.path.join(d, "c/1/0")
    with pytest.raises(IndexError):
        a.locate([(8, 0)])
    _write_zarr_json(d, chunk_key_encoding={"name": "default", "configuration": {"separator": "."}}, fill_value="NaN")
    b = ZV.ShardedArray(d)
    assert b.shard_path((1, 2)) == os.path.join(d, "c.1.2") and np.isnan(b.fill_value)
    for over, what in (({"codecs": [{"name": "bytes"}]}, "sharding_indexed"), ({"data_type": "complex64"}, "data_type"),
                       ({"chunk_grid": {"name": "rectilinear", "configuration": {}}}, "chunk_grid"), ({"zarr_format": 2}, "Zarr v3")):
        _write_zarr_json(d, **over)
        with pytest.raises(ZV.UnsupportedZarr, match=what):
            ZV.ShardedArray(d)
    meta = _write_zarr_json(d)
    meta["codecs"][0]["configuration"]["codecs"].insert(1, {"name": "zstd", "configuration": {"level": 3}})
    _write_zarr_json(d, codecs=meta["codecs"])
    with pytest.raises(ZV.UnsupportedZarr, match="zstd"):
        ZV.ShardedArray(d)
    # inner chunks whose decoded size is not a multiple of 16 bytes are refused by name (the pipeline works in 16-byte vectors)
    meta = _write_zarr_json(d)
    meta["codecs"][0]["configuration"]["chunk_shape"] = [2, 3]
    _write_zarr_json(d, data_type="uint8", codecs=meta["codecs"], chunk_grid={"name": "regular", "configuration": {"chunk_shape": [32, 48]}})
    with pytest.raises(ZV.UnsupportedZarr, match="multiple of 16"):
        ZV.ShardedArray(d)
    _write_zarr_json(d, data_type="bool")
    assert ZV.ShardedArray(d).is_bool


@pytest.mark.parametrize("align,max_read,merge,gap,pack_align", [(512, 512 << 10, 1, 0, 1), (512, 4096, 0, 0, 1), (4096, 256 << 10, 1, 4096, 16), (65536, 1 << 20, 1, 0, 512),
                                                                 (4096, 4096, 1, 0, 1), (512, 65536, 0, 0, 16)])
def test_read_planner_arithmetic_against_the_oracle(lsio, oracle, align, max_read, merge, gap, pack_align):
    """The request planner on its own (no device): for random requests -- valid, invalid, negative, EOF, overlapping and duplicate ranges --
    (1) resolution equals the oracle's restatement of sqe.rs:64-81; (2) the request's buffer is a prefix-sum layout at pack_align;
    (3) every read starts and ends on the FILE's alignment (open_file.rs:23-25: 512, 4096 or 65536 here), is at most max_read_bytes long
    and `needed` covers its furthest payload byte; (4) every valid range is covered exactly once, in order, by its segments, and each
    segment's bytes are the file bytes the range asked for (file_off + head == start_off + position in the range)."""
    rng = np.random.default_rng(align + max_read + merge)
    N = lsio._native
    for trial in range(60):
        size = int(rng.choice([1, 511, 4096, 70_001, 1_048_576, 3_000_003, 9_437_184 + 17]))
        rs = []
        for _ in range(int(rng.integers(1, 60))):
            kind = rng.integers(0, 10)
            if kind == 0:
                rs.append((0, -1))
            elif kind == 1:
                rs.append((int(size + rng.integers(1, 100)), int(size + 200)))  # beyond EOF: invalid
            elif kind == 2:
                n = int(min(size, rng.integers(1, 3000)))
                rs.append((-n, -1))
            elif kind == 3:
                rs.append((5, 5))  # empty: invalid
            else:
                n = int(rng.integers(1, min(size, 2_000_000) + 1))
                s0 = int(rng.integers(0, size - n + 1))
                rs.append((s0, s0 + n))
        starts, ends = [r[0] for r in rs], [r[1] for r in rs]
        ranges, reads, segs, total = N.plan_reads(size, starts, ends, align, max_read, merge, gap, pack_align, buffer_address=(1 << 30) + 512 * int(rng.integers(0, 8)))
        dst = 0
        for i, (s, e) in enumerate(rs):  # (1) + (2)
            rc, rr = oracle.resolve_range(size, s, e)
            valid = rc == 0 and rr.end_off <= size and rr.end_off > rr.start_off
            assert bool(ranges[i, 0]) == valid, (size, s, e)
            if valid:
                assert (ranges[i, 1], ranges[i, 2]) == (rr.start_off, rr.end_off)
                dst = (dst + pack_align - 1) // pack_align * pack_align
                assert ranges[i, 3] == dst
                dst += rr.end_off - rr.start_off
        assert total == dst
        mr = max(align, max_read // align * align)
        for k, (fo, ln, needed, first, nsegs, ntiles) in enumerate(reads.tolist()):  # (3)
            assert fo % align == 0 and ln % align == 0 and 0 < ln <= mr and 0 < needed <= ln and nsegs >= 1
            sg = segs[first:first + nsegs]
            assert int((sg[:, 1] + sg[:, 2]).max()) == needed and ntiles >= nsegs
        owner = np.zeros(len(segs), dtype=np.int64) - 1  # the read each segment belongs to
        for k, rd in enumerate(reads.tolist()):
            assert (owner[rd[3]:rd[3] + rd[4]] == -1).all()
            owner[rd[3]:rd[3] + rd[4]] = k
        assert (owner >= 0).all()
        for i in range(len(rs)):  # (4)
            mine = np.flatnonzero(segs[:, 0] == i)
            if not ranges[i, 0]:
                assert len(mine) == 0
                continue
            pos = int(ranges[i, 3])
            for j in mine[np.argsort(segs[mine, 3])]:
                _, head, plen, doff = segs[j].tolist()
                assert doff == pos and plen > 0
                assert reads[owner[j], 0] + head == ranges[i, 1] + (doff - ranges[i, 3])  # the file byte this segment starts at
                pos += plen
            assert pos == ranges[i, 3] + (ranges[i, 2] - ranges[i, 1])
        if not merge:  # the reference's shape: no read serves two ranges (duplicates and overlaps are read again)
            for rd in reads.tolist():
                assert len(set(segs[rd[3]:rd[3] + rd[4], 0].tolist())) == 1


def test_ctypes_mirrors_match_the_c_headers(lsio, tmp_path):
    """The Python mirror lays its structs out by hand (ctypes): compile a C program against include/*.h and compare sizes and the
    offset of every field.  A struct that grows in the header (round 2 added fields to the config and the stats) must grow in the
    mirror too, or every later field is read from the wrong place."""
    import ctypes
    import subprocess

    N = lsio._native
    structs = {"lsio_uring_cuda_config": N.CConfig, "lsio_uring_cuda_stats": N.CStats, "lsio_output": N.COutput, "lsio_chunk": N.CChunk,
               "lsio_cuda_segment": N.CudaSegment, "lsio_bench_result": N.CBenchResult}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "lsio_io.h"', '#include "lsio_cuda.h"', '#include "lsio_uring_cuda.h"', '#include "lsio_bench.h"',
             'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        cname, field, value = line.split()
        cls = structs[cname]
        want = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(value) == want, (cname, field, int(value), want)
