"""End-to-end measurement of the I/O path + codec layer on one B200: Zarr-like shards (chunks stored as
[Shuffle(4), LZ4, crc32c], packed back to back, unaligned) on tmpfs -> get_ranges -> HBM -> native codec pipeline
-> float32 rows in HBM, every row verified on the device against the array it was encoded from.

The shards are written with liblz4 itself (pyarrow's "lz4_raw" codec) and CRCs from the library's own pack kernel:
no oracle on this path.  `--cpu-baseline` adds the oracle's decoder on one core over a bounded sample.
Also imported by bench.py for its `codec` leg.
Usage: python tools/zarr_pipeline_bench.py [--chunk-kib 64] [--chunks 262144] [--files 256] [--json out.json]"""
import argparse
import json
import os
import shutil
import struct
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_pool(nbytes, n_distinct=64, seed=0):
    """n_distinct chunks of `nbytes` of float32 (a rounded random walk) and their stored form:
    numcodecs.Shuffle(4) -> numcodecs.LZ4 (4-byte size + liblz4 block) -> Zarr v3 crc32c trailer."""
    import pyarrow as pa
    import torch

    import lsio_b200 as L

    codec = pa.Codec("lz4_raw")
    rng = np.random.default_rng(seed)
    arrays, bodies = [], []
    for _ in range(n_distinct):
        a = (np.cumsum(rng.standard_normal(nbytes // 4)) * 0.01).astype("<f4").round(2)
        shuffled = a.view(np.uint8).reshape(-1, 4).T.copy().tobytes()
        arrays.append(a)
        bodies.append(struct.pack("<I", nbytes) + codec.compress(shuffled).to_pybytes())
    # CRC-32C of each body with the library's own kernel (checksum-only pack plan)
    blob = np.frombuffer(b"".join(b + b"\0" * (-len(b) % 16) for b in bodies), dtype=np.uint8)
    d = torch.from_numpy(blob.copy()).cuda()
    offs = np.cumsum([0] + [len(b) + (-len(b) % 16) for b in bodies[:-1]]).astype(np.uint64)
    lens = np.array([len(b) for b in bodies], dtype=np.uint64)
    crc = torch.zeros(n_distinct, dtype=torch.int32, device="cuda")
    plan = L.cuda_ops.PackPlan(0, d.data_ptr() + offs, d.data_ptr() + offs, lens, store=False, crc32c=True)
    plan.launch(crc)
    torch.cuda.synchronize()
    plan.close()
    crcs = crc.cpu().numpy().view(np.uint32)
    return [(a, b + struct.pack("<I", int(c))) for a, b, c in zip(arrays, bodies, crcs)]


def run(chunk_kib=64, chunks=262144, files=256, batch=32768, workers=8, reps=3, arena_gib=8, ranges_per_request=0, no_decode=False,
        cpu_baseline=False, directory="/dev/shm", **backend_config):
    import torch

    import lsio_b200 as L

    C = L.codecs
    nbytes = chunk_kib << 10
    pool = make_pool(nbytes)
    d = tempfile.mkdtemp(prefix="lsio_zarr_bench_", dir=directory)
    try:
        per_file = chunks // files
        fns, ranges, which = [], [], []
        for f in range(files):
            blob = bytearray(b"\x00" * (1 + f % 500))
            rs, wh = [], []
            for k in range(per_file):
                i = (f * per_file + k) % len(pool)
                rs.append((len(blob), len(blob) + len(pool[i][1])))
                wh.append(i)
                blob += pool[i][1]
            fn = os.path.join(d, f"shard{f}")
            with open(fn, "wb") as fh:
                fh.write(blob)
            fns.append(fn)
            ranges.append(rs)
            which.append(wh)
        which_arr = np.array(which, dtype=np.int64)
        n = per_file * files
        stored_bytes = sum(e - s for rs in ranges for s, e in rs)
        decoded_bytes = n * nbytes
        # a reader gets these as arrays straight from the shard index; building them is not part of the timed path
        starts = [np.array([r[0] for r in rs], dtype=np.int64) for rs in ranges]
        ends = [np.array([r[1] for r in rs], dtype=np.int64) for rs in ranges]
        uds = [(np.uint64(f) << np.uint64(32)) | np.arange(per_file, dtype=np.uint64) for f in range(files)]
        pipes = [C.CodecPipeline([C.Shuffle(4), C.LZ4(), C.Crc32c()], nbytes, max_batch=batch) for _ in range(2)]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        expect = torch.from_numpy(np.stack([p[0] for p in pool])).cuda()
        results = []
        with L.IoUringCuda(workers, arena_bytes=arena_gib << 30, completion_bound=65536, **backend_config) as u:
            rx = u.completion()
            for rep in range(reps + 1):
                torch.cuda.synchronize()
                st0 = u.stats()
                launches0 = L.cuda_ops.launch_count()
                t0 = time.perf_counter()
                rpr = ranges_per_request or per_file
                for f in range(files):
                    for a in range(0, per_file, rpr):
                        u.get_ranges_arrays(fns[f], starts[f][a:a + rpr], ends[f][a:a + rpr], uds[f][a:a + rpr])
                got, t_decode, outs, trace = 0, 0.0, [], []
                pending, inflight, k = [], [], 0  # two pipelines: batch k decodes while batch k+1 is received
                n_pending = 0
                while got < n:
                    rec = rx.recv_batch(min(batch - n_pending, n - got), timeout=30.0)  # numpy record array over lsio_output
                    pending.append(rec)
                    got += len(rec)
                    n_pending += len(rec)
                    if n_pending >= batch or got == n:  # never more than `batch`: the pipelines are sized once
                        rec = np.concatenate(pending)
                        pending, n_pending = [], 0
                        if no_decode:
                            L.Receiver.drop_buffers(rec)
                            continue
                        td = time.perf_counter()
                        if len(inflight) == 2:
                            p0, r0 = inflight.pop(0)
                            p0.finish()
                            tf = time.perf_counter()
                            L.Receiver.drop_buffers(r0)
                            trace.append(("finish", round(tf - td, 4), "drop", round(time.perf_counter() - tf, 4), "dev_ms", [round(x, 1) for x in p0.last_ms()]))
                        ts = time.perf_counter()
                        out = pipes[k % 2].submit_records(rec, stream=streams[k % 2])
                        trace.append(("submit", round(time.perf_counter() - ts, 4), "at", round(ts - t0, 4)))
                        t_decode += time.perf_counter() - td
                        inflight.append((pipes[k % 2], rec))
                        outs.append((out, rec["chunk"]["user_data"]))
                        k += 1
                td = time.perf_counter()
                for p0, r0 in inflight:
                    tf = time.perf_counter()
                    p0.finish()
                    trace.append(("finish", round(time.perf_counter() - tf, 4), "at", round(tf - t0, 4), "dev_ms", [round(x, 1) for x in p0.last_ms()]))
                    L.Receiver.drop_buffers(r0)
                t_decode += time.perf_counter() - td
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                st1 = u.stats()
                launches = L.cuda_ops.launch_count() - launches0
                # verify every decoded row against the array it was encoded from (outside the timed region)
                bad = 0
                for out, uds_ in outs:
                    idx = torch.from_numpy(which_arr[(uds_ >> np.uint64(32)).astype(np.int64), (uds_ & np.uint64(0xFFFFFFFF)).astype(np.int64)]).cuda()
                    bad += int((out.view(torch.float32) != expect[idx]).any(dim=1).sum())
                if rep:
                    results.append({"seconds": dt, "decode_wait_seconds": t_decode, "bad_chunks": bad, "kernel_launches": launches,
                                    "stored_gib_s": stored_bytes / dt / 2**30, "decoded_gib_s": decoded_bytes / dt / 2**30,
                                    "arena_stalls": st1["arena_stalls"] - st0["arena_stalls"], "staging_stalls": st1["staging_stalls"] - st0["staging_stalls"],
                                    "reads": st1["reads"] - st0["reads"], "h2d_copies": st1["h2d_copies"] - st0["h2d_copies"], "trace": trace})
                del outs
        for p in pipes:
            p.close()
        best = max(results, key=lambda r: r["decoded_gib_s"])
        rec = {"workload": f"{n} chunks of {chunk_kib} KiB float32 stored as [Shuffle(4), LZ4 (liblz4), crc32c] in {files} shards under {directory}"
                           + (" (I/O only: chunks received and dropped)" if no_decode else ""),
               "compression_ratio": decoded_bytes / stored_bytes, "decoded_bytes": decoded_bytes, "stored_bytes": stored_bytes, "batch": batch,
               "workers": workers, "best": best, "median_decoded_gib_s": float(np.median([r["decoded_gib_s"] for r in results])), "runs": results}
        if cpu_baseline:
            from oracle import oracle  # the CPU baseline leg: the oracle's decoder on one core over a bounded sample

            t0 = time.perf_counter()
            cpu_n = 0
            while time.perf_counter() - t0 < 5.0:
                f = cpu_n % files
                s, e = ranges[f][cpu_n % per_file]
                raw = oracle.pread_range(fns[f], s, e)
                assert oracle.crc32c(raw) == C.CRC32C_RESIDUE
                oracle.unshuffle(oracle.lz4_decompress(raw[4:-4], nbytes), 4)
                cpu_n += 1
            cpu_dt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"kind": "port", "cores": 1, "decoded_gib_s": cpu_n * nbytes / cpu_dt / 2**30,
                                   "sample": f"{cpu_n} chunks in {cpu_dt:.1f} s: pread + crc32c + oracle lz4 + unshuffle"}
        return rec
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk-kib", type=int, default=64)
    ap.add_argument("--chunks", type=int, default=262144)
    ap.add_argument("--files", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32768, help="chunks decoded per pipeline call")
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--arena-gib", type=int, default=8)
    ap.add_argument("--no-decode", action="store_true", help="I/O path only: receive and drop (the ceiling the decode competes with)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ranges-per-request", type=int, default=0, help="0 = one get_ranges call per file")
    ap.add_argument("--json", default=None)
    ap.add_argument("--min-copy-mib", type=int, default=0, help="backend knob: pack launches are batched up to at least this much")
    ap.add_argument("--flush-us", type=int, default=0, help="backend knob: ...or until the oldest landed extent has waited this long")
    ap.add_argument("--pack-threads", type=int, default=0)
    a = ap.parse_args()
    cfg = {}
    if a.pack_threads:
        cfg["pack_threads"] = a.pack_threads
    if a.min_copy_mib:
        cfg["min_copy_bytes"] = a.min_copy_mib << 20
    if a.flush_us:
        cfg["flush_us"] = a.flush_us
    rec = run(a.chunk_kib, a.chunks, a.files, a.batch, a.workers, a.reps, a.arena_gib, a.ranges_per_request, a.no_decode, not a.no_cpu_baseline, **cfg)
    print(json.dumps(rec))
    if a.json:
        with open(a.json, "w") as fh:
            json.dump(rec, fh, indent=1)


if __name__ == "__main__":
    main()
