"""GPU parity tests of the codec layer (-m gpu): csrc/codecs.cu through the C ABI (include/lsio_codecs.h)
against the oracle (oracle/lsio_codecs_oracle.c) and liblz4's own vectors, bit-exact, then end to end
behind the I/O path: encoded Zarr-like chunks in files -> get_ranges -> HBM -> CodecPipeline.decode."""
import json
import os
import struct
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_codec_vectors import RECIPES, recipe_bytes  # noqa: E402


def _to_dev(torch, data, lead=0, slack=64):
    """Device copy of `data` starting `lead` bytes into a fresh buffer (to exercise misalignment)."""
    t = torch.zeros(lead + len(data) + slack, dtype=torch.uint8, device="cuda:0")
    if len(data):
        t[lead:lead + len(data)] = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    return t


def _lz4_batch(lsio, torch, blocks, caps, src_lead=0, dst_lead=0, replicate=1):
    """Decode `blocks` (compressed bytes) in one launch. Returns (results, outputs with 32-byte guard bands).
    replicate=k decodes every block k times (into separate outputs): launches of more than 2048 blocks take
    the 8-lanes-per-block kernel variant, smaller ones the warp-per-block variant."""
    srcs = [_to_dev(torch, b, lead=(src_lead + 3 * i) % 16 if src_lead else 0) for i, b in enumerate(blocks)] * replicate
    leads = [(src_lead + 3 * i) % 16 if src_lead else 0 for i in range(len(blocks))] * replicate
    blocks, caps = list(blocks) * replicate, list(caps) * replicate
    outs = [torch.full((32 + dst_lead + c + 32,), 0xAB, dtype=torch.uint8, device="cuda:0") for c in caps]
    items = lsio.codecs.make_items([s.data_ptr() + l for s, l in zip(srcs, leads)], [o.data_ptr() + 32 + dst_lead for o in outs],
                                   [len(b) for b in blocks], caps)
    res = lsio.codecs.lz4_decompress(items)
    torch.cuda.synchronize()
    return res, [o.cpu().numpy() for o in outs]


@pytest.mark.parametrize("variant", [None, "1"])
def test_lz4_golden_vectors_from_liblz4(lsio, oracle, cuda_dev, monkeypatch, variant):
    import torch

    if variant:
        monkeypatch.setenv("LSIO_LZ4_VARIANT", variant)

    vs = json.load(open(os.path.join(HERE, "golden", "codecs.json")))["vectors"]
    blocks = [bytes.fromhex(v["lz4_hex"]) for v in vs]
    for src_lead, dst_lead, rep in ((0, 0, 1), (5, 3, 1), (2, 9, 160)):
        res, outs = _lz4_batch(lsio, torch, blocks, [v["decoded_len"] for v in vs], src_lead, dst_lead, rep)
        assert len(res) == len(vs) * rep
        for v, r, o in zip(vs * rep, res, outs):
            want = recipe_bytes(v)
            assert r["status"] == 0 and r["out_len"] == len(want), (v["kind"], r)
            body = o[32 + dst_lead:32 + dst_lead + len(want)]
            assert body.tobytes() == want, v["kind"]
            assert oracle.crc32c(body) == v["decoded_crc32c"]
            assert np.all(o[:32 + dst_lead] == 0xAB) and np.all(o[32 + dst_lead + len(want):] == 0xAB)  # nothing outside dst


@pytest.mark.parametrize("variant", [None, "8", "88", "32", "1", "2"])
def test_lz4_random_blocks_match_the_oracle(lsio, oracle, cuda_dev, monkeypatch, variant):
    """Many blocks per launch, both compressors (oracle's greedy one and liblz4), ragged sizes, misaligned.
    Every kernel variant (LSIO_LZ4_VARIANT: teams of 8 lanes at 5 or 8 CTAs per SM, a warp per block, or the parallel-parse
    kernel with 32 or 64 token positions per batch) must give the same bytes;
    None = the library's own choice by launch size."""
    import pyarrow as pa
    import torch

    if variant:
        monkeypatch.setenv("LSIO_LZ4_VARIANT", variant)

    codec = pa.Codec("lz4_raw")
    rng = np.random.default_rng(7)
    datas = []
    for i in range(300):
        n = int(rng.integers(1, 40000))
        kind = i % 5
        if kind == 0:
            d = rng.integers(0, int(rng.integers(2, 256)), n, dtype=np.uint8).tobytes()
        elif kind == 1:
            d = np.cumsum(rng.integers(-2, 3, n // 4 + 1)).astype("<f4").tobytes()[:n]
        elif kind == 2:
            d = oracle.shuffle(np.cumsum(rng.integers(-2, 3, n // 4 + 1)).astype("<i4").tobytes(), 4).tobytes()[:n]
        elif kind == 3:
            p = int(rng.integers(1, 70))
            d = (rng.integers(0, 256, p, dtype=np.uint8).tobytes() * (n // p + 1))[:n]
        else:
            d = bytes(n) if i % 2 else rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        datas.append(d)
    datas += [rng.integers(0, 3, n, dtype=np.uint8).tobytes() for n in range(1, 40)]
    blocks = [oracle.lz4_compress(d).tobytes() if i % 2 else codec.compress(d).to_pybytes() for i, d in enumerate(datas)]
    for b, d in zip(blocks, datas):
        assert oracle.lz4_decompress(b, len(d)).tobytes() == d
    for rep in (1, 7):  # 339 blocks: warp per block; 2373 blocks: 8 lanes per block
        res, outs = _lz4_batch(lsio, torch, blocks, [len(d) for d in datas], src_lead=1, dst_lead=7, replicate=rep)
        for i, (d, r, o) in enumerate(zip(datas * rep, res, outs)):
            assert r["status"] == 0 and r["out_len"] == len(d), (i, r, len(d))
            assert o[39:39 + len(d)].tobytes() == d, i
            assert np.all(o[:39] == 0xAB) and np.all(o[39 + len(d):] == 0xAB), i


@pytest.mark.parametrize("variant", ["8", "88", "32", "1", "2"])
def test_lz4_large_block_and_long_runs(lsio, oracle, cuda_dev, monkeypatch, variant):
    import torch

    monkeypatch.setenv("LSIO_LZ4_VARIANT", variant)
    rng = np.random.default_rng(9)
    d = (rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes() + bytes(1_000_000) + rng.integers(0, 5, 700_000, dtype=np.uint8).tobytes()
         + rng.integers(0, 256, 50_000, dtype=np.uint8).tobytes() * 4)
    b = oracle.lz4_compress(d).tobytes()
    for lead in (0, 1, 2, 3):  # long runs are copied as realigned 32-bit words: every relative alignment
        res, outs = _lz4_batch(lsio, torch, [b], [len(d) + 100], src_lead=lead, dst_lead=(5 * lead) % 4)  # cap larger than needed: out_len reports the real size
        assert res[0]["status"] == 0 and res[0]["out_len"] == len(d)
        o = outs[0][32 + (5 * lead) % 4:]
        assert o[:len(d)].tobytes() == d and np.all(o[len(d):] == 0xAB)


@pytest.mark.parametrize("variant", [None, "1"])
def test_lz4_malformed_blocks_report_their_own_status(lsio, oracle, cuda_dev, monkeypatch, variant):
    import torch

    if variant:
        monkeypatch.setenv("LSIO_LZ4_VARIANT", variant)

    C = lsio.codecs
    good = recipe_bytes({"kind": "lowent", "n": 3000, "seed": 2, "alphabet": 4})
    cg = oracle.lz4_compress(good).tobytes()
    blocks = [cg, cg, cg[:len(cg) // 2], bytes([0x10, 0x41, 0x00, 0x00, 0x50, 1, 2, 3, 4, 5]), bytes([0x10, 0x41, 0x02, 0x00, 0x50, 1, 2, 3, 4, 5]),
              bytes([0xF0, 0xFF]), cg]
    caps = [len(good), len(good) - 1, len(good), 100, 100, 100, len(good)]
    for rep in (1, 300):
        _check_malformed(lsio, oracle, torch, good, blocks, caps, rep)
    t = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
    res = C.lz4_decompress(C.make_items([t.data_ptr()], [t.data_ptr() + 32], [0], [10]))
    assert res[0]["status"] == 1
    assert len(C.lz4_decompress(C.make_items([], [], [], []))) == 0


def _check_malformed(lsio, oracle, torch, good, blocks, caps, rep):
    res, outs = _lz4_batch(lsio, torch, blocks, caps, replicate=rep)
    res, outs = res[-len(blocks):], outs[-len(blocks):]
    st = [int(r["status"]) for r in res]
    assert st[0] == 0 and st[6] == 0 and outs[0][32:32 + len(good)].tobytes() == good and outs[6][32:32 + len(good)].tobytes() == good
    assert st[1] == 2  # LSIO_CODEC_ERR_OUTPUT
    try:
        want = oracle.lz4_decompress(blocks[2], caps[2])
        assert st[2] == 0 and res[2]["out_len"] == len(want)
    except ValueError:
        assert st[2] == 1
    assert st[3] == 3 and st[4] == 3 and st[5] == 1
    for o, c in zip(outs, caps):
        assert np.all(o[:32] == 0xAB) and np.all(o[32 + c:] == 0xAB)  # a failing block never writes outside its dst


@pytest.mark.parametrize("elem", [1, 2, 3, 4, 8, 16, 12])
def test_shuffle_and_unshuffle_match_the_oracle(lsio, oracle, cuda_dev, elem):
    import torch

    C = lsio.codecs
    rng = np.random.default_rng(elem)
    sizes = [0, 1, elem, 16 * elem, 4096, 16384, 16384 + 16, 65536 + 4 * elem, 100_003, 1 << 20, (1 << 20) + elem + 1]
    for src_lead, dst_lead in ((0, 0), (0, 8), (4, 0), (1, 5)):
        datas = [rng.integers(0, 256, n, dtype=np.uint8) for n in sizes]
        srcs = [_to_dev(torch, d.tobytes(), lead=src_lead) for d in datas]
        for forward in (False, True):
            outs = [torch.full((32 + dst_lead + len(d) + 32,), 0xCD, dtype=torch.uint8, device="cuda:0") for d in datas]
            items = C.make_items([s.data_ptr() + src_lead for s in srcs], [o.data_ptr() + 32 + dst_lead for o in outs], [len(d) for d in datas])
            (C.shuffle if forward else C.unshuffle)(items, elem)
            torch.cuda.synchronize()
            for d, o in zip(datas, outs):
                h = o.cpu().numpy()
                want = (oracle.shuffle if forward else oracle.unshuffle)(d, elem)
                assert np.array_equal(h[32 + dst_lead:32 + dst_lead + len(d)], want), (elem, len(d), forward, src_lead, dst_lead)
                assert np.all(h[:32 + dst_lead] == 0xCD) and np.all(h[32 + dst_lead + len(d):] == 0xCD)


def test_shuffle_round_trip_at_full_size(lsio, cuda_dev):
    """Size-independent property at a size the oracle is not run on: unshuffle(shuffle(x)) == x, 1 GiB."""
    import torch

    C = lsio.codecs
    n_chunks, nbytes = 256, 4 << 20
    x = torch.randint(0, 256, (n_chunks, nbytes), dtype=torch.uint8, device="cuda:0")
    y, z = torch.empty_like(x), torch.empty_like(x)
    step = np.arange(n_chunks, dtype=np.uint64) * np.uint64(nbytes)
    for elem in (2, 4, 8):
        C.shuffle(C.make_items(x.data_ptr() + step, y.data_ptr() + step, np.full(n_chunks, nbytes)), elem)
        C.unshuffle(C.make_items(y.data_ptr() + step, z.data_ptr() + step, np.full(n_chunks, nbytes)), elem)
        assert torch.equal(x, z)
        assert not torch.equal(x, y)
        # plane 0 of chunk 0 holds byte 0 of every element
        assert torch.equal(y[0, :nbytes // elem], x[0, ::elem])


def test_zarr_like_chunks_through_the_whole_path(lsio, oracle, cuda_dev, data_dir):
    """Chunks encoded as a Zarr codec chain [Shuffle(4), LZ4, crc32c] packed back to back in shard files ->
    get_ranges (unaligned ranges) -> HBM -> CodecPipeline.decode -> the original float32 arrays, bit-exact."""
    import torch

    C = lsio.codecs
    rng = np.random.default_rng(21)
    chunk_shape, n_files, per_file = (64, 64), 6, 20
    nbytes = int(np.prod(chunk_shape)) * 4
    arrays, ranges, fns = {}, {}, []
    for f in range(n_files):
        fn = os.path.join(data_dir, f"shard{f}")
        fns.append(fn)
        blob = bytearray(rng.integers(0, 256, int(rng.integers(1, 700)), dtype=np.uint8).tobytes())  # unaligned start
        rs = []
        for k in range(per_file):
            a = (np.cumsum(rng.standard_normal(chunk_shape), axis=1) * 0.01).astype("<f4").round(2)
            arrays[(f, k)] = a
            body = oracle.lz4_compress(oracle.shuffle(a.tobytes(), 4)).tobytes()
            stored = struct.pack("<I", nbytes) + body  # numcodecs.LZ4 framing
            stored += struct.pack("<I", oracle.crc32c(stored))  # Zarr v3 crc32c codec
            rs.append((len(blob), len(blob) + len(stored)))
            blob += stored
        blob += b"tail"
        with open(fn, "wb") as fh:
            fh.write(blob)
        ranges[f] = rs
    pipe = C.CodecPipeline([C.Shuffle(4), C.LZ4(), C.Crc32c()], nbytes)
    with lsio.IoUringCuda(3, arena_bytes=128 << 20, staging_bytes=16 << 20) as u:
        for f in range(n_files):
            u.get_ranges(fns[f], ranges[f], [(f << 32) | k for k in range(per_file)])
        chunks = [u.completion().recv_timeout(20.0) for _ in range(n_files * per_file)]
        assert all(c.is_ok() for c in chunks)
        out = pipe.decode(chunks)
        assert out.shape == (n_files * per_file, nbytes) and out.is_cuda
        got = out.view(torch.float32).reshape(-1, *chunk_shape).cpu().numpy()
        for i, c in enumerate(chunks):
            assert np.array_equal(got[i], arrays[(c.user_data >> 32, c.user_data & 0xFFFFFFFF)]), c.user_data
        # the same through the record-array path (Receiver.recv_batch -> decode_records): no per-chunk Python work
        for f in range(n_files):
            u.get_ranges(fns[f], ranges[f], [(f << 32) | k for k in range(per_file)])
        recs = []
        while sum(len(r) for r in recs) < n_files * per_file:
            recs.append(u.completion().recv_batch(1024, timeout=20.0))
        rec = np.concatenate(recs)
        out2 = pipe.decode_records(rec)
        got2 = out2.view(torch.float32).reshape(-1, *chunk_shape).cpu().numpy()
        for i, ud in enumerate(rec["chunk"]["user_data"].tolist()):
            assert np.array_equal(got2[i], arrays[(ud >> 32, ud & 0xFFFFFFFF)]), ud
        lsio.Receiver.drop_buffers(rec)
        # a corrupted chunk is caught by the fused CRC without touching the data again
        with open(fns[0], "r+b") as fh:
            fh.seek(ranges[0][3][0] + 40)
            b = fh.read(1)
            fh.seek(ranges[0][3][0] + 40)
            fh.write(bytes([b[0] ^ 0x10]))
            fh.flush()
            os.fsync(fh.fileno())
        u.get_ranges(fns[0], ranges[0], list(range(per_file)))
        again = [u.completion().recv_timeout(20.0) for _ in range(per_file)]
        with pytest.raises(C.CodecError, match="crc32c"):
            pipe.decode(again)
        ok = [c for c in again if c.user_data != 3]
        assert pipe.decode(ok).shape[0] == per_file - 1
        # and by the native pipeline, which names the chunk
        u.get_ranges(fns[0], ranges[0], list(range(per_file)))
        rec = u.completion().recv_batch(per_file, timeout=20.0)
        while len(rec) < per_file:
            rec = np.concatenate([rec, u.completion().recv_batch(per_file - len(rec), timeout=20.0)])
        with pytest.raises(C.CodecError, match=r"user_data=3\): crc32c trailer mismatch"):
            pipe.decode_records(rec)
        lsio.Receiver.drop_buffers(rec)


@pytest.mark.parametrize("chain", ["", "shuffle", "lz4", "crc", "shuffle+lz4", "lz4+crc", "shuffle+crc"])
def test_native_pipeline_stage_combinations(lsio, oracle, cuda_dev, data_dir, chain):
    """lsio_codec_pipeline_* with every subset of the chain [Shuffle, LZ4, Crc32c], asynchronous submit/finish,
    two batches in flight on two pipelines; rows compared with the arrays the shards were encoded from."""
    import torch

    C = lsio.codecs
    rng = np.random.default_rng(len(chain))
    nbytes, n = 8192, 300
    names = [c for c in chain.split("+") if c]
    codecs = [{"shuffle": C.Shuffle(8), "lz4": C.LZ4(), "crc": C.Crc32c()}[c] for c in names]
    arrays, blob, ranges = [], bytearray(b"x" * 13), []
    for i in range(n):
        a = np.cumsum(rng.integers(-5, 6, nbytes // 8)).astype("<i8")
        arrays.append(a)
        b = a.tobytes()
        if "shuffle" in names:
            b = oracle.shuffle(b, 8).tobytes()
        if "lz4" in names:
            b = struct.pack("<I", nbytes) + oracle.lz4_compress(b).tobytes()
        if "crc" in names:
            b += struct.pack("<I", oracle.crc32c(b))
        ranges.append((len(blob), len(blob) + len(b)))
        blob += b
    fn = os.path.join(data_dir, "s")
    with open(fn, "wb") as fh:
        fh.write(blob + b"end")
    pipes = [C.CodecPipeline(codecs, nbytes), C.CodecPipeline(codecs, nbytes)]
    with lsio.IoUringCuda(2, arena_bytes=64 << 20, staging_bytes=16 << 20) as u:
        u.get_ranges(fn, ranges, range(n))
        recs, got = [], 0
        while got < n:
            r = u.completion().recv_batch(100, timeout=20.0)
            recs.append(r)
            got += len(r)
        outs = [pipes[i % 2].submit_records(r) if i < 2 else None for i, r in enumerate(recs)]
        for i, r in enumerate(recs):  # at most two decodes in flight
            pipes[i % 2].finish()
            if i + 2 < len(recs):
                outs[i + 2] = pipes[i % 2].submit_records(recs[i + 2])
        for r, o in zip(recs, outs):
            h = o.cpu().numpy().view("<i8")
            for row, ud in zip(h, r["chunk"]["user_data"].tolist()):
                assert np.array_equal(row, arrays[ud]), (chain, ud)
            lsio.Receiver.drop_buffers(r)
    for p in pipes:
        p.close()
    # stages out of encode order are refused
    if len(codecs) >= 2:
        with pytest.raises(lsio._native.LsioCudaError):
            C.CodecPipeline(codecs[::-1], nbytes).decode_records(recs[0][:0])


def test_zarr_v3_sharded_read(lsio, oracle, cuda_dev, data_dir):
    """A caller of the path (lsio_b200.zarr_shard): Zarr v3 `sharding_indexed` shards -- inner chunks in shuffled order
    with gaps and missing chunks, the index (u64 offset/nbytes pairs + crc32c) at the END of the file -- read as two
    ordered groups: indexes through negative ranges (lsio_io lib.rs:19-23), then the inner chunks they point to,
    decoded by the native pipeline.  Also: index at the start, a subset of chunks, a corrupted index."""
    import torch

    Z, C = lsio.zarr_shard, lsio.codecs
    rng = np.random.default_rng(33)
    n_shards, n_inner, shape = 5, 48, (32, 32)
    nbytes = int(np.prod(shape)) * 4
    arrays, paths = {}, []
    for s in range(n_shards):
        location = "start" if s == 4 else "end"
        order = rng.permutation(n_inner)
        missing = set(rng.choice(n_inner, 5, replace=False).tolist())
        table = np.full((n_inner, 2), 0xFFFFFFFFFFFFFFFF, dtype="<u8")
        body = bytearray()
        base = Z.index_nbytes(n_inner) if location == "start" else 0
        for k in order:
            if int(k) in missing:
                continue
            a = (np.cumsum(rng.standard_normal(shape), axis=0) * 0.1).astype("<f4").round(1)
            arrays[(s, int(k))] = a
            stored = struct.pack("<I", nbytes) + oracle.lz4_compress(oracle.shuffle(a.tobytes(), 4)).tobytes()
            stored += struct.pack("<I", oracle.crc32c(stored))
            body += bytes(int(rng.integers(0, 9)))  # gaps between inner chunks are allowed
            table[k] = (base + len(body), len(stored))
            body += stored
        index = table.tobytes() + struct.pack("<I", oracle.crc32c(table.tobytes()))
        fn = os.path.join(data_dir, f"c{s}")
        with open(fn, "wb") as fh:
            fh.write(index + bytes(body) if location == "start" else bytes(body) + index)
        paths.append(fn)
    pipe = C.CodecPipeline([C.Shuffle(4), C.LZ4(), C.Crc32c()], nbytes)
    with lsio.IoUringCuda(3, arena_bytes=64 << 20, staging_bytes=16 << 20) as u:
        rx = u.completion()
        total, present = Z.read_shards(u, paths[:4], n_inner)
        assert total == 4 * (n_inner - 5) and all(len(p) == n_inner - 5 for p in present)
        recs = []
        while sum(len(r) for r in recs) < total:
            recs.append(rx.recv_batch(1024, timeout=20.0))
        rec = np.concatenate(recs)
        assert np.all(rec["group"] == rec["group"][0])  # all of them in the chunk group
        rows = pipe.decode_records(rec).view(torch.float32).reshape(-1, *shape).cpu().numpy()
        seen = set()
        for row, ud in zip(rows, rec["chunk"]["user_data"].tolist()):
            key = (ud >> 32, ud & 0xFFFFFFFF)
            assert np.array_equal(row, arrays[key]), key
            seen.add(key)
        assert seen == {k for k in arrays if k[0] < 4}
        lsio.Receiver.drop_buffers(rec)
        # index at the start of the file, only some chunks wanted (missing ones are skipped)
        wanted = [list(range(0, n_inner, 3))]
        total, present = Z.read_shards(u, paths[4:], n_inner, wanted=wanted, index_location="start")
        assert set(present[0].tolist()) == {k for k in wanted[0] if (4, k) in arrays} and total == len(present[0])
        chunks = [rx.recv_timeout(20.0) for _ in range(total)]
        rows = pipe.decode(chunks).view(torch.float32).reshape(-1, *shape).cpu().numpy()
        for row, c in zip(rows, chunks):
            assert c.user_data >> 32 == 0 and np.array_equal(row, arrays[(4, c.user_data & 0xFFFFFFFF)])
        # a corrupted index is refused by the fused CRC
        with open(paths[0], "r+b") as fh:
            fh.seek(-20, os.SEEK_END)
            b = fh.read(1)
            fh.seek(-20, os.SEEK_END)
            fh.write(bytes([b[0] ^ 1]))
            fh.flush()
            os.fsync(fh.fileno())
        with pytest.raises(Z.ShardIndexError, match="crc32c"):
            Z.read_shards(u, paths[:1], n_inner)


def test_plain_c_consumer_of_groups_and_codec_pipeline(lsio, cuda_dev, data_dir, tmp_path):
    """tests/c/codec_pipeline_test.c: ordered groups, recv_batch, lsio_codec_pipeline_* and device AlignedBytes from
    plain C against liblsio_b200.so -- the layers above the path need no Python and no C++ on the caller's side."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(lsio._native.LIB_PATH)
    exe = str(tmp_path / "codec_pipeline_test")
    subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", f"-I{root}/include", f"{root}/tests/c/codec_pipeline_test.c", f"-L{libdir}", "-llsio_b200",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    p = subprocess.run([exe, data_dir], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "200 chunks in 2 ordered groups decoded by the native pipeline and match" in p.stdout


def test_zarr_v3_sharded_array_by_chunk_coordinates(lsio, oracle, cuda_dev, data_dir):
    """lsio_b200.zarr_v3.ShardedArray over a hand-written Zarr v3 store (zarr.json + shard files per the v3 spec and the
    sharding codec): inner chunks addressed by grid coordinates come back as float32 tiles equal to the source array, with
    the fill value for chunks never written and for a shard that does not exist; duplicates and arbitrary order allowed."""
    import json
    import torch

    ZV = lsio.zarr_v3
    rng = np.random.default_rng(5)
    shape, shard_shape, inner = (64, 96), (32, 48), (8, 16)
    fill = 7.5
    full = (np.cumsum(rng.standard_normal(shape), axis=1) * 0.1).astype("<f4").round(1)
    root = os.path.join(data_dir, "arr.zarr")
    os.makedirs(os.path.join(root, "c", "0"))
    os.makedirs(os.path.join(root, "c", "1"))
    meta = {"zarr_format": 3, "node_type": "array", "shape": list(shape), "data_type": "float32", "fill_value": fill,
            "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": list(shard_shape)}},
            "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
            "codecs": [{"name": "sharding_indexed", "configuration": {
                "chunk_shape": list(inner), "index_location": "end",
                "index_codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "crc32c"}],
                "codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "numcodecs.shuffle", "configuration": {"elementsize": 4}},
                           {"name": "numcodecs.lz4", "configuration": {"acceleration": 1}}, {"name": "crc32c"}]}}]}
    with open(os.path.join(root, "zarr.json"), "w") as f:
        json.dump(meta, f)
    per = (shard_shape[0] // inner[0], shard_shape[1] // inner[1])  # 4 x 3 inner chunks per shard
    expect = full.copy()
    for si in range(2):
        for sj in range(2):
            if (si, sj) == (1, 0):  # this shard is never written
                expect[32:64, 0:48] = fill
                continue
            table = np.full((per[0] * per[1], 2), 0xFFFFFFFFFFFFFFFF, dtype="<u8")
            body = bytearray()
            for k in rng.permutation(per[0] * per[1]):
                a, b = divmod(int(k), per[1])
                r0, c0 = si * 32 + a * 8, sj * 48 + b * 16
                if rng.integers(0, 5) == 0:  # an inner chunk that was never written
                    expect[r0:r0 + 8, c0:c0 + 16] = fill
                    continue
                tile = np.ascontiguousarray(full[r0:r0 + 8, c0:c0 + 16])
                stored = struct.pack("<I", tile.nbytes) + oracle.lz4_compress(oracle.shuffle(tile.tobytes(), 4)).tobytes()
                stored += struct.pack("<I", oracle.crc32c(stored))
                table[k] = (len(body), len(stored))
                body += stored
            index = table.tobytes() + struct.pack("<I", oracle.crc32c(table.tobytes()))
            with open(os.path.join(root, "c", str(si), str(sj)), "wb") as fh:
                fh.write(bytes(body) + index)
    arr = ZV.ShardedArray(root)
    coords = [(i, j) for i in range(8) for j in range(6)]
    order = rng.permutation(len(coords))
    ask = [coords[k] for k in order] + [coords[5], coords[5], (7, 0)]  # duplicates, a chunk of the missing shard again
    with lsio.IoUringCuda(3, arena_bytes=64 << 20, staging_bytes=16 << 20) as u:
        tiles = arr.read_inner_chunks(u, ask)
        assert tiles.shape == (len(ask), 8, 16) and tiles.dtype == torch.float32 and tiles.is_cuda
        host = tiles.cpu().numpy()
        for t, (i, j) in zip(host, ask):
            assert np.array_equal(t, expect[i * 8:(i + 1) * 8, j * 16:(j + 1) * 16]), (i, j)
        # a second read on the same backend (groups keep counting), a single chunk
        one = arr.read_inner_chunks(u, [(0, 0)])
        assert np.array_equal(one[0].cpu().numpy(), expect[0:8, 0:16])
    # a small arena forces the read into several sub-batches of shards (compressed chunks stay resident until decoded)
    with lsio.IoUringCuda(2, arena_bytes=1 << 20, staging_bytes=16 << 20) as u:
        u.arena_bytes = 4096  # pretend: at most 4 inner chunks (512 B each) per sub-batch
        tiles = arr.read_inner_chunks(u, ask)
        for t, (i, j) in zip(tiles.cpu().numpy(), ask):
            assert np.array_equal(t, expect[i * 8:(i + 1) * 8, j * 16:(j + 1) * 16]), (i, j)
    # a corrupted shard index: refused with ShardIndexError, and the backend is left clean (the next read works)
    bad = os.path.join(root, "c", "0", "1")
    blob = bytearray(open(bad, "rb").read())
    blob[-10] ^= 0x40
    with open(bad, "wb") as fh:
        fh.write(blob)
    with lsio.IoUringCuda(2, arena_bytes=64 << 20, staging_bytes=16 << 20) as u:
        with pytest.raises(lsio.zarr_shard.ShardIndexError):
            arr.read_inner_chunks(u, ask)
        again = arr.read_inner_chunks(u, [(0, 0), (1, 1)])
        assert np.array_equal(again[0].cpu().numpy(), expect[0:8, 0:16]) and np.array_equal(again[1].cpu().numpy(), expect[8:16, 16:32])
