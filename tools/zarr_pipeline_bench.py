"""End-to-end measurement of the I/O path + codec layer on one B200: Zarr-like shards (chunks stored as
[Shuffle(4), LZ4, crc32c], packed back to back, unaligned) on tmpfs -> get_ranges -> HBM -> CodecPipeline.decode
-> float32 rows in HBM.  Reports stored and decoded GiB/s, the decode stage alone, and a CPU baseline (the oracle's
decoder, one core, on a bounded sample).
Usage: python tools/zarr_pipeline_bench.py [--chunk-kib 64] [--chunks 32768] [--files 64] [--json out.json]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk-kib", type=int, default=64)
    ap.add_argument("--chunks", type=int, default=32768)
    ap.add_argument("--files", type=int, default=64)
    ap.add_argument("--batch", type=int, default=8192, help="chunks decoded per CodecPipeline.decode call")
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--arena-gib", type=int, default=8)
    ap.add_argument("--no-decode", action="store_true", help="I/O path only: receive and drop (the ceiling the decode competes with)")
    ap.add_argument("--ranges-per-request", type=int, default=0, help="0 = one get_ranges call per file")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import torch

    import lsio_b200 as L
    from oracle import oracle  # encoder for the synthetic shards + the CPU baseline; never on the GPU path

    C = L.codecs
    nbytes = args.chunk_kib << 10
    rng = np.random.default_rng(0)
    pool = []
    for _ in range(64):  # 64 distinct chunks, replicated
        a = (np.cumsum(rng.standard_normal(nbytes // 4)) * 0.01).astype("<f4").round(2)
        body = oracle.lz4_compress(oracle.shuffle(a.tobytes(), 4)).tobytes()
        stored = struct.pack("<I", nbytes) + body
        stored += struct.pack("<I", oracle.crc32c(stored))
        pool.append((a, stored))
    d = tempfile.mkdtemp(prefix="lsio_zarr_bench_", dir="/dev/shm")
    try:
        per_file = args.chunks // args.files
        fns, ranges, which = [], [], []
        for f in range(args.files):
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
        n = per_file * args.files
        stored_bytes = sum(e - s for rs in ranges for s, e in rs)
        decoded_bytes = n * nbytes
        # a reader gets these as arrays straight from the shard index; building them is not part of the timed path
        starts = [np.array([r[0] for r in rs], dtype=np.int64) for rs in ranges]
        ends = [np.array([r[1] for r in rs], dtype=np.int64) for rs in ranges]
        uds = [(np.uint64(f) << np.uint64(32)) | np.arange(per_file, dtype=np.uint64) for f in range(args.files)]
        pipes = [C.CodecPipeline([C.Shuffle(4), C.LZ4(), C.Crc32c()], nbytes) for _ in range(2)]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        expect = torch.from_numpy(np.stack([p[0] for p in pool])).cuda()
        results = []
        with L.IoUringCuda(args.workers, arena_bytes=args.arena_gib << 30, completion_bound=65536) as u:
            rx = u.completion()
            for rep in range(args.reps + 1):
                torch.cuda.synchronize()
                st0 = u.stats()
                t0 = time.perf_counter()
                rpr = args.ranges_per_request or per_file
                for f in range(args.files):
                    for a in range(0, per_file, rpr):
                        u.get_ranges_arrays(fns[f], starts[f][a:a + rpr], ends[f][a:a + rpr], uds[f][a:a + rpr])
                got, t_decode, outs = 0, 0.0, []
                pending, inflight, k = [], [], 0  # two pipelines: batch k decodes while batch k+1 is received
                while got < n:
                    rec = rx.recv_batch(min(args.batch, n - got), timeout=30.0)  # numpy record array over lsio_output
                    pending.append(rec)
                    got += len(rec)
                    if sum(len(r) for r in pending) >= args.batch or got == n:
                        rec = np.concatenate(pending)
                        pending = []
                        if args.no_decode:
                            L.Receiver.drop_buffers(rec)
                            continue
                        td = time.perf_counter()
                        if len(inflight) == 2:
                            p0, r0 = inflight.pop(0)
                            p0.finish()
                            L.Receiver.drop_buffers(r0)
                        out = pipes[k % 2].submit_records(rec, stream=streams[k % 2])
                        t_decode += time.perf_counter() - td
                        inflight.append((pipes[k % 2], rec))
                        outs.append((out, rec["chunk"]["user_data"]))
                        k += 1
                td = time.perf_counter()
                for p0, r0 in inflight:
                    p0.finish()
                    L.Receiver.drop_buffers(r0)
                t_decode += time.perf_counter() - td
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                # verify every decoded row against the array it was encoded from (outside the timed region)
                bad = 0
                for out, uds_ in outs:
                    idx = torch.from_numpy(which_arr[(uds_ >> np.uint64(32)).astype(np.int64), (uds_ & np.uint64(0xFFFFFFFF)).astype(np.int64)]).cuda()
                    bad += int((out.view(torch.float32) != expect[idx]).any(dim=1).sum())
                st1 = u.stats()
                if rep:
                    results.append({"arena_stalls": st1["arena_stalls"] - st0["arena_stalls"], "staging_stalls": st1["staging_stalls"] - st0["staging_stalls"],
                                    "reads": st1["reads"] - st0["reads"], "h2d_copies": st1["h2d_copies"] - st0["h2d_copies"], "seconds": dt, "bad_chunks": bad,
                                    "stored_gib_s": stored_bytes / dt / 2**30, "decoded_gib_s": decoded_bytes / dt / 2**30,
                                    "decode_wait_seconds": t_decode})
                del outs
        # CPU baseline: the oracle decoder on one core over a bounded sample (pread + crc + lz4 + unshuffle)
        t0 = time.perf_counter()
        cpu_n = 0
        while time.perf_counter() - t0 < 5.0:
            f = cpu_n % args.files
            s, e = ranges[f][cpu_n % per_file]
            raw = oracle.pread_range(fns[f], s, e)
            assert oracle.crc32c(raw) == C.CRC32C_RESIDUE
            oracle.unshuffle(oracle.lz4_decompress(raw[4:-4], nbytes), 4)
            cpu_n += 1
        cpu_dt = time.perf_counter() - t0
        best = max(results, key=lambda r: r["decoded_gib_s"])
        rec = {"workload": f"{n} chunks of {args.chunk_kib} KiB float32 stored as [Shuffle(4), LZ4, crc32c] in {args.files} tmpfs shards",
               "compression_ratio": decoded_bytes / stored_bytes, "batch": args.batch, "workers": args.workers, "best": best, "runs": results,
               "cpu_baseline": {"kind": "port", "cores": 1, "decoded_gib_s": cpu_n * nbytes / cpu_dt / 2**30,
                                "sample": f"{cpu_n} chunks in {cpu_dt:.1f} s: pread + crc32c + oracle lz4 + unshuffle"}}
        print(json.dumps(rec))
        if args.json:
            with open(args.json, "w") as fh:
                json.dump(rec, fh, indent=1)
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
