#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1], "C2"): the lsio_bench README workload -- 500 files x
41,214,400 B read as 157 blocks of 262,144 B per file (crates/lsio_bench/src/main.rs:137) --
delivered into the HBM of one B200, bit-exact.  A "step" is one pass over all 78,500 ranges.

  value   : payload GiB/s of the device-resident pass: the 500 padded-read images are already
            in HBM when the timed region starts; one step = one pack_strip_crc32c launch (+ its
            fold kernel) over all 78,500 ranges.  CUDA events on the launching stream.
  e2e     : the same metric through the reference-facing C ABI (lsio_uring_cuda_get_ranges +
            completion channel) from files: O_DIRECT reads -> pinned ring -> H2D -> pack+crc ->
            one chunk per range received and its CRC checked, all inside the timed region.
  N > 1   : one process per GPU (torchrun); ranks share no data path: each rank lands the 500
            files of its own shard in its own GPU ("weak"); value = all ranks' bytes / max time.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FILE_SIZE = 41_214_400
BLOCK = 262_144
N_FILES = 500
N_CHUNKS = FILE_SIZE // BLOCK  # 157: lsio_bench drops the 57,792-byte tail
METRIC = "GiB/s of byte ranges landed in HBM"
GIB = float(1 << 30)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---- distributed plumbing (torch.distributed only; no data-path collective) ----------------------
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def dist_init(backend):
    import torch
    import torch.distributed as dist

    rank, world, local = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        kw = {"device_id": torch.device(f"cuda:{local}")} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_bytes_and_time(nbytes, secs, device):
    """Whole-job aggregation: SUM of the units every rank processed, MAX of the time."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return int(nbytes), float(secs)
    b = torch.tensor([float(nbytes)], dtype=torch.float64, device=device)
    t = torch.tensor([float(secs)], dtype=torch.float64, device=device)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(b.item()), float(t.item())


# ---- clocks during the timed region (B200_PROFILING.md) -------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---- workload -------------------------------------------------------------------------------------
def pick_data_dir(args, need_bytes):
    if args.dir:
        return args.dir, ("tmpfs" if args.dir.startswith("/dev/shm") else "disk")
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize > need_bytes + (8 << 30):
            return "/dev/shm/lsio_bench", "tmpfs"
    except OSError:
        pass
    return "/tmp/lsio_bench", "disk"


def chunk_ranges():
    starts = np.arange(N_CHUNKS, dtype=np.int64) * BLOCK
    return starts, starts + BLOCK


def make_files(directory, n_files, device_index, want_crc=True):
    """Create the synthetic files with the product's own fill_pattern kernel (position-keyed
    content) and return (paths, expected crc32c[n_files, N_CHUNKS]) -- the expected
    CRCs come from the crc-only kernel over the generated image, not from the oracle."""
    import torch

    from lsio_b200 import cuda_ops

    os.makedirs(directory, exist_ok=True)
    paths = [os.path.join(directory, f"lsio_bench_{i}") for i in range(n_files)]
    crc_path = os.path.join(directory, f"expected_crc32c_{n_files}.npy")
    if os.path.exists(crc_path) and all(os.path.exists(p) and os.path.getsize(p) == FILE_SIZE for p in paths):
        return paths, np.load(crc_path)
    dev = torch.device(f"cuda:{device_index}")
    group = 16  # files per generation batch
    dbuf = torch.empty(group * FILE_SIZE, dtype=torch.uint8, device=dev)
    hbuf = torch.empty(group * FILE_SIZE, dtype=torch.uint8, pin_memory=True)
    crc_all = np.zeros((n_files, N_CHUNKS), dtype=np.uint32)
    starts, _ = chunk_ranges()
    for g0 in range(0, n_files, group):
        g1 = min(n_files, g0 + group)
        for i in range(g0, g1):
            cuda_ops.fill_pattern(dbuf[(i - g0) * FILE_SIZE:(i - g0 + 1) * FILE_SIZE], i, 0)
        offs = np.concatenate([starts + (i - g0) * FILE_SIZE for i in range(g0, g1)]).astype(np.uint64)
        crc = cuda_ops.pack_strip_crc32c(dbuf, None, offs, offs, np.full(len(offs), BLOCK, dtype=np.uint64), store=False)
        crc_all[g0:g1] = crc.cpu().numpy().view(np.uint32).reshape(g1 - g0, N_CHUNKS)
        hbuf[: (g1 - g0) * FILE_SIZE].copy_(dbuf[: (g1 - g0) * FILE_SIZE])
        torch.cuda.synchronize(dev)
        harr = hbuf.numpy()

        def write(i):
            with open(paths[i], "wb") as f:
                f.write(memoryview(harr[(i - g0) * FILE_SIZE:(i - g0 + 1) * FILE_SIZE]))
                f.flush()
                os.fsync(f.fileno())

        ths = [threading.Thread(target=write, args=(i,)) for i in range(g0, g1)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    np.save(crc_path, crc_all)
    del dbuf, hbuf
    torch.cuda.empty_cache()
    return paths, crc_all


# ---- the device-resident pass (`value`, roofline) ---------------------------------------------------
def run_value(args, dev, n_files, expect_crc):
    import torch

    from lsio_b200 import cuda_ops

    payload = n_files * N_CHUNKS * BLOCK
    src = torch.empty(n_files * FILE_SIZE, dtype=torch.uint8, device=dev)  # the padded-read images, as the H2D stage leaves them
    for i in range(n_files):
        cuda_ops.fill_pattern(src[i * FILE_SIZE:(i + 1) * FILE_SIZE], i, 0)
    dst = torch.empty(payload, dtype=torch.uint8, device=dev)                # per-request packed buffers, back to back
    starts, _ = chunk_ranges()
    so = np.concatenate([starts.astype(np.uint64) + np.uint64(i * FILE_SIZE) for i in range(n_files)])
    do = np.arange(n_files * N_CHUNKS, dtype=np.uint64) * np.uint64(BLOCK)
    lens = np.full(len(so), BLOCK, dtype=np.uint64)
    plan = cuda_ops.PackPlan(dev.index, so + np.uint64(src.data_ptr()), do + np.uint64(dst.data_ptr()), lens)
    crc = torch.empty(len(so), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        for _ in range(args.warmup):
            plan.launch(crc, st)
    torch.cuda.synchronize(dev)
    barrier()
    launches0 = cuda_ops.launch_count()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(st):
        e[0].record(st)
        for k in range(args.steps):  # 41 GB touched per step >> 126 MB L2: no flush needed
            plan.launch(crc, st)
            e[k + 1].record(st)
    torch.cuda.synchronize(dev)
    barrier()
    launches = cuda_ops.launch_count() - launches0
    step_ms = [e[k].elapsed_time(e[k + 1]) for k in range(args.steps)]
    total_ms = e[0].elapsed_time(e[-1])
    # parity at full size: CRCs equal the table made at generation time; bytes equal the pattern
    got = crc.cpu().numpy().view(np.uint32).reshape(n_files, N_CHUNKS)
    crc_ok = bool(np.array_equal(got, expect_crc[:n_files]))
    ids = np.repeat(np.arange(n_files, dtype=np.uint64), N_CHUNKS)
    foffs = np.tile(starts.astype(np.uint64), n_files)
    mism = cuda_ops.verify_pattern(dst, do, lens, ids, foffs)
    algo = plan.algorithmic_bytes
    plan.close()
    del src, dst
    torch.cuda.empty_cache()
    return {"payload": payload, "total_ms": total_ms, "step_ms": step_ms, "algo_bytes": algo, "launches": launches,
            "crc_ok": crc_ok, "mismatched_bytes": mism}


# ---- the end-to-end pass through the C ABI (`e2e`) ----------------------------------------------------
def e2e_pass(u, paths, file_ids, expect_crc, starts, ends, ud_lo):
    """One pass of lsio_bench::read_files over the backend's C ABI (native driver, csrc/lsio_bench.cpp):
    submit every file, receive one chunk per range, check its device CRC against the table made at
    generation time, drop it; ends with a device synchronise."""
    secs, nbytes, ok, bad = u.bench_read_files(paths, file_ids, BLOCK, N_CHUNKS, expect_crc)
    return secs, nbytes, bad + (len(paths) * N_CHUNKS - ok - bad)


def run_e2e(args, dev, paths, file_ids, expect_crc, steps, warmup, workers, o_direct=True):
    import lsio_b200 as L

    starts, ends = chunk_ranges()
    ud_lo = np.arange(N_CHUNKS, dtype=np.uint64)
    u = L.IoUringCuda(workers, devices=(dev.index,), arena_bytes=args.arena_gib << 30, staging_bytes=args.staging_mib << 20, o_direct=int(o_direct),
                      min_copy_bytes=args.min_copy_mib << 20, flush_us=args.flush_us, pack_threads=args.pack_threads, sq_entries=args.sq_entries, max_read_bytes=args.max_read_kib << 10, merge_ranges=args.merge_ranges, zero_copy=args.zero_copy, poll_sleep_us=args.poll_sleep_us,
                      staging_mode=args.staging_mode, host_slots_per_worker=args.host_slots, host_slot_kib=args.host_slot_kib)
    try:
        for _ in range(warmup):
            e2e_pass(u, paths, file_ids, expect_crc, starts, ends, ud_lo)
        barrier()
        secs, nbytes, bad = [], 0, 0
        t_all = time.perf_counter()
        for _ in range(steps):
            s, b, x = e2e_pass(u, paths, file_ids, expect_crc, starts, ends, ud_lo)
            secs.append(s)
            nbytes += b
            bad += x
        t_all = time.perf_counter() - t_all
        barrier()
        stats = u.stats()
    finally:
        u.drop()
    return {"secs": secs, "total_secs": t_all, "bytes": nbytes, "bad": bad, "stats": stats}


def verify_full_landing(dev, paths, file_ids, workers, staging_mib):
    """Untimed proof at full size: land every range, keep all chunks resident, and check every
    landed byte against the pattern with the verify_pattern kernel (no D2H)."""
    import torch

    import lsio_b200 as L
    from lsio_b200 import cuda_ops
    from lsio_b200.io_uring_cuda import Receiver

    starts, ends = chunk_ranges()
    ud_lo = np.arange(N_CHUNKS, dtype=np.uint64)
    need = len(paths) * N_CHUNKS * BLOCK
    u = L.IoUringCuda(workers, devices=(dev.index,), arena_bytes=need + (2 << 30), staging_bytes=staging_mib << 20)
    recs = []
    try:
        for p, fid in zip(paths, file_ids):
            u.get_ranges_arrays(p, starts, ends, (np.uint64(fid) << np.uint64(32)) | ud_lo)
        got = 0
        rx = u.completion()
        while got < len(paths) * N_CHUNKS:
            r = rx.recv_batch(8192, 60.0).copy()
            recs.append(r)
            got += len(r)
        rec = np.concatenate(recs)
        assert (rec["status"] == 0).all(), "errors during full landing"
        ud = rec["chunk"]["user_data"]
        views = [L.AlignedBytes(int(h)) for h in rec["chunk"]["buffer"]]
        ptrs = np.array([v.as_ptr() for v in views], dtype=np.uint64)
        lens = np.array([v.len() for v in views], dtype=np.uint64)
        anchor = torch.empty(1, dtype=torch.uint8, device=dev)
        mism = cuda_ops.verify_pattern(anchor, ptrs - np.uint64(anchor.data_ptr()), lens, ud >> np.uint64(32),
                                       (ud & np.uint64(0xFFFFFFFF)) * np.uint64(BLOCK))
        for v in views:
            v.drop()
    finally:
        u.drop()
    return int(mism), int(lens.sum())


# ---- the other BASELINE.json configs (parity-test cases with numbers; not the headline line) ---------------
def make_pattern_files(directory, prefix, n_files, size, device_index, id_base):
    """n_files files of `size` bytes, content byte(id_base + i, off); generated on the GPU."""
    import torch

    from lsio_b200 import cuda_ops

    os.makedirs(directory, exist_ok=True)
    paths = [os.path.join(directory, f"{prefix}_{i}") for i in range(n_files)]
    if all(os.path.exists(p) and os.path.getsize(p) == size for p in paths):
        return paths
    dev = torch.device(f"cuda:{device_index}")
    group = max(1, min(n_files, (1 << 30) // size))
    dbuf = torch.empty(group * size, dtype=torch.uint8, device=dev)
    hbuf = torch.empty(group * size, dtype=torch.uint8, pin_memory=True)
    for g0 in range(0, n_files, group):
        g1 = min(n_files, g0 + group)
        for i in range(g0, g1):
            cuda_ops.fill_pattern(dbuf[(i - g0) * size:(i - g0 + 1) * size], id_base + i, 0)
        hbuf[: (g1 - g0) * size].copy_(dbuf[: (g1 - g0) * size])
        torch.cuda.synchronize(dev)
        harr = hbuf.numpy()

        def write(ids):
            for i in ids:
                with open(paths[i], "wb") as f:
                    f.write(memoryview(harr[(i - g0) * size:(i - g0 + 1) * size]))

        ids = list(range(g0, g1))
        ths = [threading.Thread(target=write, args=(ids[k::16],)) for k in range(16)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    return paths


def config_workload(name, directory, device_index, args):
    """Returns (paths, file_ids, ranges_per_file [(starts, ends)], sizes, description)."""
    from numpy.random import MT19937, Generator

    if name == "C3":  # Zarr-like random reads
        n_files, size = args.c3_files, 1_049_576
        paths = make_pattern_files(directory, "c3", n_files, size, device_index, 100_000)
        per = []
        for i in range(n_files):
            rng = Generator(MT19937(0xC3 + i))
            ln = rng.integers(4096, 65537, size=8)
            st = (rng.random(8) * (size - ln + 1)).astype(np.int64)
            s, e = st, st + ln
            if i % 100 == 0:  # one range ending exactly at EOF and one written as negative -len..-1
                s = np.concatenate([s, [size - 5000, -7000]])
                e = np.concatenate([e, [size, -1]])
            per.append((s.astype(np.int64), e.astype(np.int64)))
        return paths, [100_000 + i for i in range(n_files)], per, size, f"C3: {n_files} files x {size} B, 8 ranges/file of 4-64 KiB at arbitrary offsets (+EOF/negative ranges)"
    if name == "C4":  # 1 MiB ranges over the C1 file set
        paths, _ = make_files(directory, args.files, device_index)
        k = FILE_SIZE // (1 << 20)
        st = np.arange(k, dtype=np.int64) << 20
        return paths, list(range(args.files)), [(st, st + (1 << 20))] * args.files, FILE_SIZE, f"C4: {args.files} files x {FILE_SIZE} B, {k} ranges of 1 MiB per file"
    if name == "C5":  # large streaming scan, 4 MiB ranges
        n_files, size = args.c5_files, 1 << 30
        paths = make_pattern_files(directory, "c5", n_files, size, device_index, 200_000)
        st = np.arange(size >> 22, dtype=np.int64) << 22
        return paths, [200_000 + i for i in range(n_files)], [(st, st + (1 << 22))] * n_files, size, f"C5: {n_files} files x 1 GiB, 256 ranges of 4 MiB per file, {args.c5_loops} loops"
    raise ValueError(name)


def resolve(size, s, e):
    so = np.where(s >= 0, s, size + s)
    eo = np.where(e >= 0, e, size + e + 1)
    return so, eo


def run_config(args, dev, rank):
    """One BASELINE config end to end through the C ABI: an untimed verification pass (every landed
    byte checked on the device against the pattern, fused CRCs cross-checked by the crc-only kernel
    over the landed chunks), then timed passes."""
    import torch

    import lsio_b200 as L
    from lsio_b200 import cuda_ops
    from lsio_b200.io_uring_cuda import Receiver

    directory, source = pick_data_dir(args, 0)
    directory = directory + "_" + args.config.lower()
    paths, ids, per, size, desc = config_workload(args.config, directory, dev.index, args)
    n_ranges = sum(len(p[0]) for p in per)
    lens_by_file = []
    for (s, e) in per:
        so, eo = resolve(size, s, e)
        lens_by_file.append((so, eo - so))
    payload = int(sum(int(l.sum()) for _, l in lens_by_file))
    base = np.cumsum([0] + [len(p[0]) for p in per])  # user_data = global range index
    starts_all = np.concatenate([so for so, _ in lens_by_file]).astype(np.uint64)
    lens_all = np.concatenate([l for _, l in lens_by_file]).astype(np.uint64)
    fid_all = np.concatenate([np.full(len(p[0]), fid, dtype=np.uint64) for p, fid in zip(per, ids)])
    keep_resident = payload < (60 << 30)
    arena = (payload + (4 << 30)) if keep_resident else (8 << 30)

    def one_pass(u, keep):
        rx = u.completion()
        t0 = time.perf_counter()
        for i, p in enumerate(paths):
            s, e = per[i]
            u.get_ranges_arrays(p, s, e, np.arange(base[i], base[i + 1], dtype=np.uint64))
        got, recs = 0, []
        while got < n_ranges:
            rec = rx.recv_batch(8192, 120.0)
            got += len(rec)
            if keep:
                recs.append(rec.copy())
            else:
                assert (rec["status"] == 0).all()
                Receiver.drop_buffers(rec)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, (np.concatenate(recs) if keep else None)

    result = {"config": desc, "source": source, "ranges": n_ranges, "payload_bytes": payload}
    with L.IoUringCuda(args.workers, devices=(dev.index,), arena_bytes=arena, staging_bytes=args.staging_mib << 20,
                       min_copy_bytes=args.min_copy_mib << 20, sq_entries=args.sq_entries, statx_via_ring=args.statx_via_ring,
                       max_read_bytes=args.max_read_kib << 10, merge_ranges=args.merge_ranges, zero_copy=args.zero_copy, poll_sleep_us=args.poll_sleep_us,
                       staging_mode=args.staging_mode, host_slots_per_worker=args.host_slots, host_slot_kib=args.host_slot_kib) as u:
        if keep_resident:  # verification pass
            _, rec = one_pass(u, True)
            assert (rec["status"] == 0).all(), rec[rec["status"] != 0][:3]
            ud = rec["chunk"]["user_data"].astype(np.int64)
            views = [L.AlignedBytes(int(h)) for h in rec["chunk"]["buffer"]]
            ptrs = np.array([v.as_ptr() for v in views], dtype=np.uint64)
            vlen = np.array([v.len() for v in views], dtype=np.uint64)
            assert np.array_equal(vlen, lens_all[ud]), "chunk lengths differ from end_off - start_off"
            anchor = torch.empty(1, dtype=torch.uint8, device=dev)
            rel = ptrs - np.uint64(anchor.data_ptr())
            mism = cuda_ops.verify_pattern(anchor, rel, vlen, fid_all[ud], starts_all[ud])
            crc2 = cuda_ops.pack_strip_crc32c(anchor, None, rel, rel, vlen, store=False).cpu().numpy().view(np.uint32)
            result["verify"] = {"mismatched_bytes": int(mism), "crc_fused_equals_crc_of_landed_bytes": bool(np.array_equal(crc2, rec["crc32c"])),
                                "chunks": len(views)}
            for v in views:
                v.drop()
        secs = []
        loops = args.c5_loops if args.config == "C5" else 1
        expect = None
        if keep_resident:  # the fused CRCs of the verified pass, indexed by user_data, are the expectation of the timed passes
            expect = np.zeros(n_ranges, dtype=np.uint32)
            expect[ud] = rec["crc32c"]

        def native_pass():  # timed passes run the native driver (csrc/lsio_bench.cpp): no Python in the loop
            t, nb, ok, bad = u.bench_read_ranges(paths, per, expect)
            assert bad == 0 and ok == n_ranges and nb == payload, (ok, bad, nb)
            return t

        for _ in range(max(1, args.warmup - 2)):
            native_pass()
        for _ in range(args.steps):
            secs.append(sum(native_pass() for _ in range(loops)))
        st = u.stats()
    result["gib_s"] = payload * loops * len(secs) / sum(secs) / GIB
    result["secs_per_step"] = secs
    result["bytes_per_step"] = payload * loops
    result["pipeline"] = {k: st[k] for k in ("reads", "read_bytes", "short_reads", "h2d_copies", "h2d_bytes", "pack_launches", "staging_stalls", "arena_stalls")}
    result["read_amplification"] = st["read_bytes"] / max(1, st["payload_bytes"])
    if rank == 0:
        result["source_read_ceiling_gib_s"] = cuda_ops.probe_read_ceiling(paths[: min(len(paths), 2000)], min(1 << 20, size // 512 * 512), max(8, args.workers), 32, True)
        if not args.no_cpu_baseline:  # the CPU port on the same files (the oracle's only use here)
            from oracle import oracle

            dt, okb, errs = oracle.bench_read_ranges(paths, per, 8, sqpoll=True)
            result["cpu_baseline"] = {"value": okb / dt / GIB, "unit": "GiB/s", "kind": "port", "cores": 16, "errors": errs,
                                      "sample": "one pass over the whole config, 8 workers + SQPOLL, C driver (oracle/lsio_uring_ref.c)"}
        print(json.dumps(result), flush=True)
    return 0


def codec_leg(hbm_peak):
    """Informative extra: the codec layer behind the same path.  Kernel numbers with CUDA
    events on the launching stream; the pipeline number is wall clock from the first get_ranges to the last decoded,
    device-verified row.  Shards are written with liblz4 (pyarrow), not with the oracle."""
    import ctypes

    import torch

    import lsio_b200 as L

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import zarr_pipeline_bench as Z

    C, N = L.codecs, L._native
    st = torch.cuda.current_stream()
    n, nbytes = 16384, 64 << 10

    def timed(fn, iters=5, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in evs:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return statistics.mean(a.elapsed_time(b) for a, b in evs)

    # byte unshuffle, element size 4: 1 GiB in, 1 GiB out per launch (>> L2)
    x = torch.randint(0, 256, (n, nbytes), dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x)
    step = np.arange(n, dtype=np.uint64) * np.uint64(nbytes)
    d_items = torch.from_numpy(C.make_items(x.data_ptr() + step, y.data_ptr() + step, np.full(n, nbytes)).view(np.uint8)).cuda()
    tiles = C.shuffle_tile_map(np.full(n, nbytes))
    d_tiles = torch.from_numpy(tiles.view(np.int32)).cuda()
    ms = timed(lambda: N.check(N.lib.lsio_codecs_unshuffle_async(0, d_items.data_ptr(), d_tiles.data_ptr(), len(tiles), 4, 0,
                                                                  ctypes.c_void_p(st.cuda_stream)), "unshuffle"))
    shuf = {"kernel": "shuffle_kernel<unshuffle>, elem 4", "algorithmic_bytes_per_launch": 2 * n * nbytes, "ms": ms,
            "achieved": 2 * n * nbytes / ms / 1e6, "peak": hbm_peak, "unit": "GB/s", "frac": 2 * n * nbytes / ms / 1e6 / hbm_peak, "bound": "hbm"}
    del x
    # LZ4 block decode of 16,384 distinct-address blocks (64 liblz4-compressed chunks, each copied 256 times)
    pool = Z.make_pool(nbytes)
    bodies = [np.frombuffer(p[1][4:-4], dtype=np.uint8) for p in pool]  # strip the numcodecs header and the crc32c trailer
    stride = (max(len(b) for b in bodies) + 15) // 16 * 16
    host = np.zeros((n, stride), dtype=np.uint8)
    for i in range(n):
        host[i, :len(bodies[i % len(bodies)])] = bodies[i % len(bodies)]
    d_src = torch.from_numpy(host).cuda()
    lens = np.array([len(bodies[i % len(bodies)]) for i in range(n)])
    d_items = torch.from_numpy(C.make_items(d_src.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(stride), y.data_ptr() + step, lens, nbytes).view(np.uint8)).cuda()
    d_res = torch.zeros(2 * n, dtype=torch.int32, device="cuda")
    ms = timed(lambda: N.check(N.lib.lsio_codecs_lz4_decompress_async(0, d_items.data_ptr(), n, d_res.data_ptr(), ctypes.c_void_p(st.cuda_stream)), "lz4"), 3, 2)
    res = d_res.cpu().numpy().reshape(-1, 2)
    ok = bool(np.all(res[:, 1] == 0) and np.all(res[:, 0] == nbytes))
    lz4 = {"kernel": "lz4_kernel<8 lanes per block>", "blocks": n, "block_bytes": nbytes, "ratio": n * nbytes / int(lens.sum()), "ms": ms,
           "decoded_gb_s": n * nbytes / ms / 1e6, "all_blocks_ok": ok,
           "frac_of_hbm_peak": (n * nbytes + int(lens.sum())) / ms / 1e6 / hbm_peak, "bound": "instruction issue / latency of a serial token stream, not HBM"}
    del d_src, y
    pipe = Z.run(chunk_kib=64, chunks=131072, files=128, batch=32768, workers=8, reps=4)
    io_only = Z.run(chunk_kib=64, chunks=131072, files=128, batch=32768, workers=8, reps=1, no_decode=True)
    return {"is": "Zarr-like shards [Shuffle(4), LZ4, crc32c] -> get_ranges -> HBM -> native codec pipeline -> float32 rows, every row verified on the device",
            "workload": pipe["workload"], "compression_ratio": pipe["compression_ratio"],
            "decoded_gib_s": pipe["median_decoded_gib_s"], "decoded_gib_s_is": "median of 4 passes",
            "best_decoded_gib_s": pipe["best"]["decoded_gib_s"], "best_stored_gib_s": pipe["best"]["stored_gib_s"], "bad_chunks": sum(r["bad_chunks"] for r in pipe["runs"]),
            "io_only_stored_gib_s": io_only["best"]["stored_gib_s"], "runs_decoded_gib_s": [r["decoded_gib_s"] for r in pipe["runs"]],
            "shuffle_roofline": shuf, "lz4": lz4}


def pcie_h2d_ceiling(dev):
    import torch

    h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize(dev)
    return 40 * (64 << 20) / (a.elapsed_time(b) * 1e-3) / GIB


# ---- CPU baselines (the only place the oracle is executed by bench.py) ----------------------------------
def cpu_ref_pass(paths, workers, sqpoll=True, mode=0):
    from oracle import oracle

    starts, ends = chunk_ranges()
    secs, nbytes, errors = oracle.bench_read_files(paths, starts, ends, workers, sqpoll=sqpoll, mode=mode)
    return secs, nbytes, errors


def write_files_cpu(directory, n_files):
    """Reference arm only (no GPU needed there): files written by the oracle's generator."""
    from oracle import oracle

    os.makedirs(directory, exist_ok=True)
    paths = [os.path.join(directory, f"lsio_bench_{i}") for i in range(n_files)]
    todo = [i for i, p in enumerate(paths) if not (os.path.exists(p) and os.path.getsize(p) == FILE_SIZE)]

    def work(ids):
        for i in ids:
            oracle.write_file(paths[i], i, FILE_SIZE)

    nt = min(16, max(1, len(todo)))
    ths = [threading.Thread(target=work, args=(todo[k::nt],)) for k in range(nt)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    return paths


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm
    need = args.files * FILE_SIZE
    directory, source = pick_data_dir(args, need)
    paths = write_files_cpu(directory, args.files)
    cores = os.cpu_count() or 1
    workers = args.ref_workers or max(1, cores // 2)  # + one SQPOLL kernel thread per ring = all host threads
    # bound each step: estimate on a small sample, then size the sample for <= ~20 s per step
    est_s, est_b, _ = cpu_ref_pass(paths[: min(20, len(paths))], workers)
    rate = est_b / max(est_s, 1e-9)
    n_sample = len(paths) if need / rate <= 20 else max(20, int(20 * rate / FILE_SIZE))
    sample = paths[:n_sample]
    for _ in range(args.warmup):
        cpu_ref_pass(sample, workers)
    secs, nbytes, errors = [], 0, 0
    for _ in range(args.steps):
        s, b, e = cpu_ref_pass(sample, workers)
        secs.append(s)
        nbytes += b
        errors += e
    value = nbytes / sum(secs) / GIB
    desc = f"{n_sample} of {args.files} files x {N_CHUNKS} blocks of {BLOCK} B per step, source {source}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GiB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "C2: lsio_bench README workload, 500 files x 41,214,400 B, 157 blocks of 262,144 B per file, into host buffers (CPU lsio_uring port)",
                   "source": source, "dir": directory, "files": args.files, "worker_threads": workers, "sqpoll": True, "errors": errors,
                   "note": "the reference is Rust and cannot be built here; this is oracle/lsio_uring_ref.c, its C restatement"},
        "cpu_baseline": {"value": value, "unit": "GiB/s", "cores": workers * 2, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- main ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dir", default=None, help="directory for the synthetic files (default: /dev/shm if it has room)")
    ap.add_argument("--files", type=int, default=N_FILES)
    ap.add_argument("--workers", type=int, default=0, help="io_uring worker threads per GPU (0 = 8, the README config, capped by the CPUs available per rank)")
    ap.add_argument("--ref-workers", type=int, default=0)
    ap.add_argument("--staging-mib", type=int, default=1024)
    ap.add_argument("--arena-gib", type=int, default=8)
    ap.add_argument("--min-copy-mib", type=int, default=64)
    ap.add_argument("--pack-threads", type=int, default=0, help="threads per CTA of the pipeline's pack launches (0 = 1024)")
    ap.add_argument("--flush-us", type=int, default=2000, help="a pack batch is launched once it holds --min-copy-mib or its oldest extent has waited this long")
    ap.add_argument("--sq-entries", type=int, default=64)
    ap.add_argument("--max-read-kib", type=int, default=512, help="reads (merged or split) are cut at this size")
    ap.add_argument("--merge-ranges", type=int, default=1)
    ap.add_argument("--poll-sleep-us", type=int, default=50, help="device-thread poll interval in microseconds (0 = spin)")
    ap.add_argument("--staging-mode", type=int, default=1, help="1 = per-worker cache-sized pinned slots (default), 0 = one big pinned ring")
    ap.add_argument("--host-slots", type=int, default=4, help="staging mode 1: pinned slots per worker")
    ap.add_argument("--host-slot-kib", type=int, default=2048, help="staging mode 1: slot (extent, H2D copy) size")
    ap.add_argument("--e2e-dir", default=None, help="--e2e-only: run the pass over files in this directory instead")
    ap.add_argument("--zero-copy", type=int, default=0, help="1 = the pack kernel reads the pinned ring over PCIe (no H2D memcpy)")
    ap.add_argument("--statx-via-ring", type=int, default=0, help="1 = the reference's openat+statx SQE pair per file")
    ap.add_argument("--disk-dir", default="/tmp/lsio_bench_disk", help="secondary e2e leg on a disk-backed directory ('' disables)")
    ap.add_argument("--disk-files", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-codec", action="store_true", help="skip the informative codec-layer leg (N=1 only)")
    ap.add_argument("--value-only", action="store_true", help="only the device-resident kernel pass (for ncu captures)")
    ap.add_argument("--e2e-only", action="store_true", help="only the end-to-end pass (pipeline tuning sweeps)")
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"], help="C2 is the contract line; the others print an informational line")
    ap.add_argument("--c3-files", type=int, default=10_000)
    ap.add_argument("--c5-files", type=int, default=40)
    ap.add_argument("--c5-loops", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch

    import lsio_b200  # noqa: F401  fails loudly if the CUDA extension is missing
    from lsio_b200 import cuda_ops, sharding

    rank, world, local = dist_env()
    cpus_per_rank = (os.cpu_count() or 8) // max(world, 1)
    if args.workers <= 0:  # leave cores for the consumer and the (mostly sleeping) device thread
        args.workers = max(2, min(8, cpus_per_rank - 2))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist_init("nccl")
    n_gpus = max(world, 1)

    if args.config != "C2":
        return run_config(args, dev, rank)

    need = args.files * FILE_SIZE
    directory, source = pick_data_dir(args, need)
    t_gen = time.time()
    if rank == 0:
        paths, expect_crc = make_files(directory, args.files, local)
    barrier()
    if rank != 0:
        paths, expect_crc = make_files(directory, args.files, local)  # already there: only loads the table
    t_gen = time.time() - t_gen
    # weak scaling: n_gpus * files logical files dealt file_idx mod n_gpus; logical file j is
    # physical file j % files (the file set is shared read-only to bound storage; O_DIRECT, no cache reuse)
    mine = sharding.files_for_rank(args.files * n_gpus, rank, n_gpus)
    my_paths = [paths[j % args.files] for j in mine]
    my_ids = [j % args.files for j in mine]

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    if args.value_only:
        val = run_value(args, dev, args.files, expect_crc)
        kernel_ms = statistics.mean(val["step_ms"])
        print(json.dumps({"value_only": True, "files": args.files, "kernel_ms": kernel_ms, "achieved_gbs": val["algo_bytes"] / kernel_ms / 1e6,
                          "frac": val["algo_bytes"] / kernel_ms / 1e6 / hbm_peak, "crc_ok": val["crc_ok"], "mismatched_bytes": val["mismatched_bytes"]}), flush=True)
        return 0

    if args.e2e_only:
        if args.e2e_dir:
            my_paths, expect_crc = make_files(args.e2e_dir, args.files, local)
            my_ids = list(range(args.files))
        e2e = run_e2e(args, dev, my_paths, my_ids, expect_crc, args.steps, args.warmup, args.workers)
        print(json.dumps({"e2e_only": True, "gib_s": e2e["bytes"] / sum(e2e["secs"]) / GIB, "bad": e2e["bad"], "workers": args.workers,
                          "staging_mib": args.staging_mib, "min_copy_mib": args.min_copy_mib, "sq_entries": args.sq_entries,
                          "stalls": [e2e["stats"]["staging_stalls"], e2e["stats"]["arena_stalls"]], "copies": e2e["stats"]["h2d_copies"]}), flush=True)
        return 0

    with ClockSampler(local) as clocks:
        val = run_value(args, dev, args.files, expect_crc)
        e2e = run_e2e(args, dev, my_paths, my_ids, expect_crc, args.steps, args.warmup, args.workers)
    clk = clocks.summary()

    # ceilings of the I/O roofline, measured in the same run (rank 0's view) while every other rank is idle
    ceil = {}
    barrier()
    if rank == 0:
        ceil["pcie_h2d_gib_s"] = pcie_h2d_ceiling(dev)
        nthr = min(os.cpu_count() or 8, max(8, args.workers) * n_gpus)
        # raw READ_FIXED reads, no GPU work: into a large buffer (DRAM-bound destination) and into small per-thread
        # buffers (cache-resident destination, what staging_mode 1 does); the ceiling is the better of the two
        ceil["source_read_big_buffers_gib_s"] = cuda_ops.probe_read_ceiling(paths[: min(len(paths), 200)], 1 << 20, nthr, 32, True)
        ceil["source_read_gib_s"] = max(ceil["source_read_big_buffers_gib_s"], cuda_ops.probe_read_ceiling(paths[: min(len(paths), 200)], 512 << 10, nthr, 4, True))
    barrier()
    full = None
    if not args.no_verify:
        mism, landed = verify_full_landing(dev, my_paths, my_ids, args.workers, args.staging_mib)
        full = {"landed_bytes": landed, "mismatched_bytes": mism}

    # secondary leg: the same workload from a disk-backed directory (this box: one virtio disk)
    disk = None
    if rank == 0 and n_gpus == 1 and args.disk_dir and args.disk_files > 0 and source != "disk":
        try:
            free = shutil.disk_usage(os.path.dirname(args.disk_dir.rstrip("/")) or "/").free
            if free > args.disk_files * FILE_SIZE + (4 << 30):
                dpaths, dcrc = make_files(args.disk_dir, args.disk_files, local)
                dres = run_e2e(args, dev, dpaths, list(range(args.disk_files)), dcrc, 2, 1, args.workers)
                dceil = max(cuda_ops.probe_read_ceiling(dpaths, 1 << 20, 8, 32, True), cuda_ops.probe_read_ceiling(dpaths, 256 << 10, 8, 32, True))
                disk = {"value": dres["bytes"] / sum(dres["secs"]) / GIB, "unit": "GiB/s", "files": args.disk_files, "bad": dres["bad"],
                        "read_ceiling_gib_s": dceil, "frac_of_ceiling": dres["bytes"] / sum(dres["secs"]) / GIB / dceil, "dir": args.disk_dir}
        except Exception as ex:  # the disk leg is informative only
            disk = {"error": repr(ex)}

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        workers = 8  # the README configuration (8 worker threads + 8 SQPOLL kernel threads)
        s, b, errs = cpu_ref_pass(paths, workers)
        cpu = {"value": b / s / GIB, "unit": "GiB/s", "cores": min(cores, workers * 2), "kind": "port",
               "sample": f"one pass over all {len(paths)} files ({b} B) from {source}, 8 worker threads + SQPOLL, into fresh host buffers; {errs} errors",
               "nproc": cores}
        # the figure as lsio_bench itself prints it: filesize x n_files / elapsed, i.e. counting the 57,792-byte tail of
        # every file that it never reads (crates/lsio_bench/src/main.rs:137, :181-183): +0.14 %
        cpu["reference_style_gib_s"] = FILE_SIZE * len(paths) / s / GIB
        try:  # also at one worker per core, without SQPOLL (workers + kernel pollers would oversubscribe the box)
            s3, b3, _ = cpu_ref_pass(paths, cores, sqpoll=False)
            cpu["nproc_workers_no_sqpoll_gib_s"] = b3 / s3 / GIB
        except Exception as ex:
            cpu["nproc_workers_no_sqpoll_gib_s"] = repr(ex)
        s2, b2, _ = cpu_ref_pass(paths, workers, mode=1)
        cpu["object_store_like_gib_s"] = b2 / s2 / GIB

    codec = None
    if rank == 0 and n_gpus == 1 and not args.no_codec:
        try:
            codec = codec_leg(hbm_peak)
        except Exception as ex:  # informative only: never costs the contract line
            codec = {"error": repr(ex)}

    # aggregate over ranks: sum of bytes, max of time
    v_bytes, v_secs = reduce_bytes_and_time(val["payload"] * args.steps, val["total_ms"] * 1e-3, dev)
    e_bytes, e_secs = reduce_bytes_and_time(e2e["bytes"], e2e["total_secs"], dev)
    bad_total, _ = reduce_bytes_and_time(e2e["bad"] + (0 if val["crc_ok"] else 1) + val["mismatched_bytes"], 0.0, dev)

    if rank == 0:
        kernel_ms = statistics.mean(val["step_ms"])
        achieved = val["algo_bytes"] / (kernel_ms * 1e-3) / 1e9
        e2e_gib = e_bytes / e_secs / GIB
        io_ceiling = min(ceil["source_read_gib_s"], ceil["pcie_h2d_gib_s"] * n_gpus)
        padded_per_step = len(my_paths) * N_CHUNKS * BLOCK  # aligned workload: padded == payload
        line = {
            "metric": METRIC, "value": v_bytes / v_secs / GIB, "unit": "GiB/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * v_secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {
                "workload": "C2: lsio_bench README workload, 500 files x 41,214,400 B, 157 blocks of 262,144 B per file (78,500 ranges, 20,578,304,000 B) delivered into HBM of 1 B200, bit-exact",
                "files": args.files, "ranges_per_step_per_gpu": len(my_paths) * N_CHUNKS, "source": source, "dir": directory,
                "value_is": "device-resident pass: padded-read images already in HBM; one pack_strip_crc32c launch (+fold) per step over all ranges",
                "e2e_is": "files -> O_DIRECT READ_FIXED -> pinned ring -> H2D -> pack+crc32c -> one chunk per range through the C ABI, CRC of every chunk checked",
                "l2": "inputs larger than L2 (41 GB touched per step vs 126 MB)", "worker_threads_per_gpu": args.workers,
                "multi_gpu": "files dealt file_idx mod n_gpus, no collective on the data path; each rank lands 500 files (shared read-only file set)",
                "file_generation_s": round(t_gen, 1),
            },
            "e2e": {"value": e2e_gib, "unit": "GiB/s", "h2d_bytes_per_step": padded_per_step, "d2h_bytes_per_step": 4 * len(my_paths) * N_CHUNKS,
                    "ms_per_step": 1e3 * e_secs / args.steps, "bad_chunks": e2e["bad"],
                    "io_roofline": {"bound": "source" if ceil["source_read_gib_s"] < ceil["pcie_h2d_gib_s"] * n_gpus else "pcie", "achieved": e2e_gib,
                                    "peak": io_ceiling, "unit": "GiB/s", "frac": e2e_gib / io_ceiling, **ceil},
                    "pipeline": {k: e2e["stats"][k] for k in ("reads", "short_reads", "h2d_copies", "h2d_bytes", "pack_launches", "kernel_launches", "staging_stalls", "arena_stalls")}},
            "gpu_launches": val["launches"],
            "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "pack_kernel<store,crc32c> + fold_kernel, per step", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": val["algo_bytes"]},
            "cpu_baseline": cpu,
            "parity": {"value_pass_crc_equal_expected": val["crc_ok"], "value_pass_mismatched_bytes": val["mismatched_bytes"],
                       "e2e_bad_chunks_all_ranks": bad_total, "full_landing": full},
            "e2e_disk": disk,
            "codec": codec,
        }
        traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(traffic_file):
            try:
                line["roofline"]["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
            except (OSError, ValueError):
                pass
        print(json.dumps(line), flush=True)
    barrier()
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
