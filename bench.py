#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1], "C2"): the lsio_bench README workload -- 500 files x
41,214,400 B read as 157 blocks of 262,144 B per file (crates/lsio_bench/src/main.rs:137) --
delivered into the HBM of one B200, bit-exact.  A "step" is one pass over all 78,500 ranges of
every rank's shard (files dealt file_idx mod N, no collective on the data path).

  value    : payload GiB/s landed in HBM through the C ABI (lsio_uring_cuda_get_ranges + completion
             channel, driven by the native lsio_bench driver, csrc/lsio_bench.cpp): files -> reads
             into pinned slots -> H2D -> checksum (+pack) kernel -> one chunk per range, the fused
             CRC-32C of EVERY chunk compared with a table the ORACLE made from the files.
             Sum of the bytes of all ranks / max over ranks of the time of the K steps.
  e2e      : the same workload through the Python mirror of the reference API (the call a user
             makes: IoUringCuda.get_ranges + Receiver.recv_batch), CRCs checked in numpy.
  roofline : the checksum/pack kernel alone, device-resident (one launch over all ranges per step,
             CUDA events on the launching stream) + `pipeline_kernel`: the same kernel's launches
             inside the timed region of `value` (event pairs recorded by the backend).
  N > 1    : extra legs C4 (1 MiB ranges) and, at N = 8, C5 (4 MiB ranges looped to ~1 TB).
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
MIB = 1 << 20


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


def reduce_sum_max(nbytes, secs, device):
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


def gather_floats(x, device):
    """[x of rank 0, x of rank 1, ...] on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [float(x)]
    t = torch.zeros(dist.get_world_size(), dtype=torch.float64, device=device)
    t[dist.get_rank()] = float(x)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.cpu()]


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


# ---- synthetic files and their expected checksums: made by the ORACLE, never by the product ---------
def pick_data_dir(args, need_bytes):
    if args.dir:
        return args.dir, ("tmpfs" if args.dir.startswith("/dev/shm") else "disk")
    try:
        st = os.statvfs("/dev/shm")
        have = sum(os.path.getsize(os.path.join("/dev/shm/lsio_bench", f)) for f in os.listdir("/dev/shm/lsio_bench")) if os.path.isdir("/dev/shm/lsio_bench") else 0
        if st.f_bavail * st.f_frsize + have > need_bytes + (8 << 30):
            return "/dev/shm/lsio_bench", "tmpfs"
    except OSError:
        pass
    return "/tmp/lsio_bench", "disk"


def chunk_ranges():
    starts = np.arange(N_CHUNKS, dtype=np.int64) * BLOCK
    return starts, starts + BLOCK


def _threads(fn, items, nt=None):
    nt = nt or min(32, os.cpu_count() or 8, max(1, len(items)))
    errs = []

    def work(sub):
        try:
            for it in sub:
                fn(it)
        except Exception as ex:  # surfaced below
            errs.append(ex)

    ths = [threading.Thread(target=work, args=(items[k::nt],)) for k in range(nt)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    if errs:
        raise errs[0]


def write_files_cpu(directory, prefix, n_files, size, id_base=0):
    """Position-keyed synthetic files (byte = f(file id, offset)) written by the oracle's generator
    (oracle/lsio_oracle.c), so that neither the content nor the expected checksums come from the product."""
    from oracle import oracle

    os.makedirs(directory, exist_ok=True)
    paths = [os.path.join(directory, f"{prefix}_{i}") for i in range(n_files)]
    todo = [i for i, p in enumerate(paths) if not (os.path.exists(p) and os.path.getsize(p) == size)]
    _threads(lambda i: oracle.write_file(paths[i], id_base + i, size), todo)
    return paths


def oracle_crc_table(directory, tag, paths, blocksize, n_blocks):
    """uint32[n_files, n_blocks]: CRC-32C of every block of every file, computed by the ORACLE from the files
    (SSE4.2, one thread per file; the ctypes call releases the GIL).  Cached next to the files."""
    from oracle import oracle

    fn = os.path.join(directory, f"oracle_crc32c_{tag}_{len(paths)}x{n_blocks}x{blocksize}.npy")
    if os.path.exists(fn):
        t = np.load(fn)
        if t.shape == (len(paths), n_blocks):
            return t
    t = np.zeros((len(paths), n_blocks), dtype=np.uint32)

    def one(i):
        c = oracle.crc32c_file_blocks(paths[i], blocksize, n_blocks)
        assert len(c) == n_blocks, (paths[i], len(c))
        t[i] = c

    _threads(one, list(range(len(paths))))
    np.save(fn + ".tmp.npy", t)
    os.replace(fn + ".tmp.npy", fn)
    return t


def evict_page_cache(paths):
    """crates/lsio_bench/src/main.rs:192-199 (`vmtouch -e`) restated: fsync + posix_fadvise(DONTNEED).
    A no-op on tmpfs (the files ARE the page cache)."""
    for p in paths:
        try:
            fd = os.open(p, os.O_RDONLY)
            try:
                os.fsync(fd)
                os.posix_fadvise(fd, 0, 0, os.POSIX_FADV_DONTNEED)
            finally:
                os.close(fd)
        except OSError:
            pass


def llc_bytes():
    best = 0
    for idx in (3, 4, 2):
        try:
            v = open(f"/sys/devices/system/cpu/cpu0/cache/index{idx}/size").read().strip()
            n = int(v[:-1]) << (10 if v[-1] == "K" else 20) if v[-1] in "KM" else int(v)
            if idx >= 3 or not best:
                best = max(best, n)
        except (OSError, ValueError):
            pass
    return best or (32 << 20)


def deal_files(n_files, rank, n_gpus, mode=1):
    """Physical file ids rank `rank` lands in one step (always n_files of them: weak scaling).  mode 1: the physical files are dealt
    file_idx mod n_gpus and a rank walks its own share n_gpus times (no file is read by two ranks); mode 0: logical file j of
    n_files * n_gpus is physical file j % n_files (round 1; at 8 ranks r and r+4 walk the same files); mode 2: every rank walks all
    files in the same order (experiment: worst case of sharing)."""
    from lsio_b200 import sharding

    if mode == 2:
        return list(range(n_files))
    if mode:
        phys = sharding.files_for_rank(n_files, rank, n_gpus)
        return [phys[i % len(phys)] for i in range(n_files)]
    return [j % n_files for j in sharding.files_for_rank(n_files * n_gpus, rank, n_gpus)]


# ---- backend construction ----------------------------------------------------------------------------
def cpu_slice(rank, world, reverse=False):
    """Disjoint CPU slice of this rank (one process per GPU on a shared box), or (0, 0) when alone / not contiguous.
    reverse: rank r takes the slice of rank world-1-r (experiment: does a rank's speed follow its GPU or its CPUs?)."""
    cpus = sorted(os.sched_getaffinity(0))
    if world <= 1:
        return 0, 0, len(cpus)
    per = len(cpus) // world
    if reverse:
        rank = world - 1 - rank
    mine = cpus[rank * per:(rank + 1) * per]
    if not mine or mine != list(range(mine[0], mine[0] + len(mine))):
        return 0, 0, per
    return mine[0], len(mine), per


def backend(args, dev, workers=None, **over):
    import lsio_b200 as L

    kw = dict(devices=(dev.index,), arena_bytes=args.arena_gib << 30, staging_bytes=args.staging_mib << 20, min_copy_bytes=args.min_copy_mib << 20,
              flush_us=args.flush_us, pack_threads=args.pack_threads, sq_entries=args.sq_entries, max_read_bytes=args.max_read_kib << 10,
              merge_ranges=args.merge_ranges, zero_copy=args.zero_copy, poll_sleep_us=args.poll_sleep_us, staging_mode=args.staging_mode,
              host_slots_per_worker=args.host_slots, host_slot_kib=args.host_slot_kib, statx_via_ring=args.statx_via_ring,
              inline_reads=args.inline_reads, direct_landing=args.direct_landing, steal_batch=args.steal_batch,
              pack_light_below_mib=args.pack_light_below_mib, completion_bound=args.completion_bound, copy_coalesce_slots=args.copy_coalesce_slots, window_budget_kib=args.window_budget_kib, pull_any_shift=args.pull_any_shift)
    if args.pin and args.cpu_count:
        kw.update(cpu_first=args.cpu_first, cpu_count=args.cpu_count)
    kw.update(over)
    return L.IoUringCuda(workers or args.workers, **kw)


PIPE_KEYS = ("pull_launches", "requests", "reads", "read_bytes", "short_reads", "inline_reads", "h2d_copies", "h2d_bytes", "direct_extents", "pack_launches", "kernel_launches",
             "staging_stalls", "arena_stalls", "read_ns", "copy_enqueue_ns", "worker_stall_ns", "worker_busy_ns", "pack_kernel_ns", "pack_kernel_bytes",
             "max_dio_align", "payload_bytes")


def stats_delta(a, b):
    return {k: b[k] - a[k] for k in PIPE_KEYS if k != "max_dio_align"} | {k: b[k] for k in ("max_dio_align", "slot_bytes", "slots_per_worker")}


def per_worker(st, workers, secs):
    """What one worker thread does with its time (verdict r1 #2c: a number, not an inference)."""
    out = {"workers": workers}
    if st["read_ns"]:
        out["inline_read_gib_s_per_thread"] = st["read_bytes"] / st["read_ns"] * 1e9 / GIB if st["inline_reads"] else None
    busy = max(1, st["worker_busy_ns"])
    out["frac_of_awake_time_reading"] = st["read_ns"] / busy
    out["frac_of_awake_time_enqueueing_h2d"] = st["copy_enqueue_ns"] / busy
    out["frac_of_awake_time_waiting_for_slots_or_hbm"] = st["worker_stall_ns"] / busy
    out["awake_frac_of_wall"] = st["worker_busy_ns"] / max(1.0, workers * secs * 1e9)
    return out


# ---- passes over the C2 workload --------------------------------------------------------------------
def native_pass(u, paths, ids, expect):
    """lsio_bench::read_files over the C ABI (csrc/lsio_bench.cpp): submit every file, receive one chunk per
    range, compare its fused CRC with the oracle's table, drop it; ends with a device synchronise."""
    secs, nbytes, ok, bad = u.bench_read_files(paths, ids, BLOCK, N_CHUNKS, expect)
    return secs, nbytes, bad + (len(paths) * N_CHUNKS - ok - bad)


def python_pass(u, paths, ids, expect, dev):
    """The same pass written against the Python mirror of the reference API, as a user would."""
    import torch

    from lsio_b200.io_uring_cuda import Receiver

    starts, ends = chunk_ranges()
    lo = np.arange(N_CHUNKS, dtype=np.uint64)
    total = len(paths) * N_CHUNKS
    rx = u.completion()
    t0 = time.perf_counter()
    for p, fid in zip(paths, ids):
        u.get_ranges_arrays(p, starts, ends, (np.uint64(fid) << np.uint64(32)) | lo)
    got = nbytes = bad = 0
    while got < total:
        rec = rx.recv_batch(8192, 60.0)
        ok = rec["status"] == 0
        ud = rec["chunk"]["user_data"]
        good = ok & (rec["crc32c"] == expect[(ud >> np.uint64(32)).astype(np.int64) % len(expect), (ud & np.uint64(0xFFFFFFFF)).astype(np.int64) % N_CHUNKS])
        bad += int(len(rec) - good.sum())
        nbytes += int(good.sum()) * BLOCK
        Receiver.drop_buffers(rec)
        got += len(rec)
        if len(rec) < 2048 and got < total:
            time.sleep(0.002)  # consume in large batches: a Python loop per handful of chunks would take a core from the io workers
    torch.cuda.synchronize(dev)
    return time.perf_counter() - t0, nbytes, bad


def timed_leg(u, one_pass, steps, warmup, dev):
    """W untimed passes, then exactly K passes bracketed by barrier + synchronize on both sides."""
    import torch

    for _ in range(warmup):
        one_pass()
    torch.cuda.synchronize(dev)
    barrier()
    u.reset_latency()
    st0 = u.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    secs, nbytes, bad = [], 0, 0
    t_all = time.perf_counter()
    for _ in range(steps):
        s, b, x = one_pass()
        secs.append(s)
        nbytes += b
        bad += x
    torch.cuda.synchronize(dev)
    t_all = time.perf_counter() - t_all
    e1.record()
    torch.cuda.synchronize(dev)
    barrier()
    return {"secs": secs, "total_secs": t_all, "cuda_event_secs": e0.elapsed_time(e1) * 1e-3, "bytes": nbytes, "bad": bad,
            "stats": stats_delta(st0, u.stats()), "latency": u.latency_summary()}


# ---- the device-resident kernel pass (`roofline`) -----------------------------------------------------
def run_kernel(args, dev, n_files, expect_crc, in_place=False):
    """The padded-read images are already in HBM; one step = one launch of the pack+crc kernel over all ranges
    (+ its fold kernel).  `in_place`: the crc-only form the pipeline runs on directly landed extents (traffic L)."""
    import torch

    from lsio_b200 import cuda_ops

    payload = n_files * N_CHUNKS * BLOCK
    src = torch.empty(n_files * FILE_SIZE, dtype=torch.uint8, device=dev)
    for i in range(n_files):
        cuda_ops.fill_pattern(src[i * FILE_SIZE:(i + 1) * FILE_SIZE], i, 0)
    dst = None if in_place else torch.empty(payload, dtype=torch.uint8, device=dev)
    starts, _ = chunk_ranges()
    so = np.concatenate([starts.astype(np.uint64) + np.uint64(i * FILE_SIZE) for i in range(n_files)])
    do = so if in_place else np.arange(n_files * N_CHUNKS, dtype=np.uint64) * np.uint64(BLOCK)
    lens = np.full(len(so), BLOCK, dtype=np.uint64)
    plan = cuda_ops.PackPlan(dev.index, so + np.uint64(src.data_ptr()), do + np.uint64((src if in_place else dst).data_ptr()), lens, store=not in_place)
    crc = torch.empty(len(so), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        for _ in range(max(3, args.warmup)):
            plan.launch(crc, st)
    torch.cuda.synchronize(dev)
    launches0 = cuda_ops.launch_count()
    steps = max(5, min(args.steps, 20))
    e = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with torch.cuda.stream(st):
        e[0].record(st)
        for k in range(steps):  # 41 GB touched per step >> 126 MB L2: no flush needed
            plan.launch(crc, st)
            e[k + 1].record(st)
    torch.cuda.synchronize(dev)
    launches = cuda_ops.launch_count() - launches0
    step_ms = [e[k].elapsed_time(e[k + 1]) for k in range(steps)]
    got = crc.cpu().numpy().view(np.uint32).reshape(n_files, N_CHUNKS)
    crc_ok = bool(np.array_equal(got, expect_crc[:n_files]))  # the generator kernel's image == the oracle's files, CRC by CRC
    algo = plan.algorithmic_bytes
    plan.close()
    del src, dst
    torch.cuda.empty_cache()
    return {"payload": payload, "step_ms": step_ms, "algo_bytes": algo, "launches": launches, "steps": steps, "crc_ok": crc_ok}


# ---- parity beyond checksums: landed bytes vs file bytes -------------------------------------------------
def verify_landed_bytes(args, dev, paths, ids, every):
    """Untimed: land every `every`-th file of this rank's shard, copy each chunk back and compare it byte for byte
    with the file's own bytes (read with numpy, no product code on the expectation side)."""
    import torch

    import lsio_b200 as L

    starts, ends = chunk_ranges()
    lo = np.arange(N_CHUNKS, dtype=np.uint64)
    sel = list(range(0, len(paths), every))
    mism = compared = 0
    with backend(args, dev, arena_bytes=4 << 30) as u:
        rx = u.completion()
        for base in range(0, len(sel), 16):
            part = sel[base:base + 16]
            for j in part:
                u.get_ranges_arrays(paths[j], starts, ends, (np.uint64(j) << np.uint64(32)) | lo)
            files = {j: np.fromfile(paths[j], dtype=np.uint8) for j in part}
            got = 0
            while got < len(part) * N_CHUNKS:
                rec = rx.recv_batch(4096, 60.0).copy()
                assert (rec["status"] == 0).all(), "errors while landing the verification sample"
                bufs = L.Receiver.take_buffers(rec)
                for r, b in zip(rec, bufs):
                    ud = int(r["chunk"]["user_data"])
                    j, k = ud >> 32, ud & 0xFFFFFFFF
                    data = b.to_torch().cpu().numpy()
                    want = files[j][k * BLOCK:(k + 1) * BLOCK]
                    mism += int(np.count_nonzero(data != want)) if len(data) == len(want) else max(len(data), len(want))
                    compared += len(want)
                    b.drop()
                got += len(rec)
    torch.cuda.synchronize(dev)
    return {"files": len(sel), "bytes_compared": compared, "mismatched_bytes": mism, "expectation": "numpy.fromfile of the same file"}


# ---- the other BASELINE.json configs ------------------------------------------------------------------------
def resolve(size, s, e):
    so = np.where(s >= 0, s, size + s)
    eo = np.where(e >= 0, e, size + e + 1)
    return so, eo


def c3_workload(args, directory):
    from numpy.random import MT19937, Generator

    n_files, size = args.c3_files, 1_049_576
    paths = write_files_cpu(directory, "c3", n_files, size, 100_000)
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
    return paths, per, size, f"C3: {n_files} files x {size} B, 8 ranges/file of 4-64 KiB at arbitrary offsets (+EOF/negative ranges)"


def oracle_crc_of_ranges(directory, tag, paths, per, size):
    """Expected CRC-32C of every range, from the oracle's pread + crc32c (cached)."""
    from oracle import oracle

    n = sum(len(p[0]) for p in per)
    fn = os.path.join(directory, f"oracle_crc32c_{tag}_{len(paths)}_{n}.npy")
    if os.path.exists(fn):
        t = np.load(fn)
        if len(t) == n:
            return t
    base = np.cumsum([0] + [len(p[0]) for p in per])
    t = np.zeros(n, dtype=np.uint32)

    def one(i):
        data = np.fromfile(paths[i], dtype=np.uint8)
        so, eo = resolve(size, per[i][0], per[i][1])
        for k in range(len(so)):
            t[base[i] + k] = oracle.crc32c(data[so[k]:eo[k]])

    _threads(one, list(range(len(paths))))
    np.save(fn + ".tmp.npy", t)
    os.replace(fn + ".tmp.npy", fn)
    return t


def ranges_leg(args, dev, paths, per, expect, payload, steps, warmup, loops=1, **over):
    """K timed steps of `loops` passes each over (paths, per-file range arrays) through the native driver; every chunk's
    fused CRC is compared with `expect` (indexed by the global range index)."""
    n_ranges = sum(len(p[0]) for p in per)
    with backend(args, dev, **over) as u:
        def one():
            t = nb = bad = 0
            for _ in range(loops):
                s, b, ok, x = u.bench_read_ranges(paths, per, expect)
                t += s
                nb += b
                bad += x + (n_ranges - ok - x)
            return t, nb, bad

        leg = timed_leg(u, one, steps, warmup, dev)
    leg["payload_per_step"] = payload * loops
    return leg


def leg_summary(leg, dev, workers, n_gpus):
    """Reduce one leg over the ranks: SUM bytes / MAX time."""
    b, t = reduce_sum_max(leg["bytes"], leg["total_secs"], dev)
    bad, _ = reduce_sum_max(leg["bad"], 0.0, dev)
    per_rank = gather_floats(leg["bytes"] / leg["total_secs"] / GIB, dev)
    st = leg["stats"]
    return {"gib_s": b / t / GIB, "bytes": b, "secs": t, "steps": len(leg["secs"]), "ms_per_step": 1e3 * t / len(leg["secs"]), "bad_chunks": bad,
            "per_rank_gib_s": [round(x, 2) for x in per_rank], "latency_rank0": leg["latency"],
            "pipeline_rank0": {k: st[k] for k in ("reads", "read_bytes", "short_reads", "inline_reads", "h2d_copies", "h2d_bytes", "direct_extents",
                                                  "pull_launches", "pack_launches", "kernel_launches", "staging_stalls", "arena_stalls", "max_dio_align", "slot_bytes", "slots_per_worker")},
            "read_amplification": st["read_bytes"] / max(1, st["payload_bytes"]),
            "per_worker_rank0": per_worker(st, workers, leg["total_secs"])}


def pipeline_kernel(st, hbm_peak, secs=None, pcie_gib_s=None):
    """The kernel launches INSIDE the timed region.  Extents moved by the pull kernel: one launch per extent on the workers'
    streams, bound by the PCIe link it reads through (its 'roofline' is the link, not HBM); extents that went ring -> pack
    kernel: event pairs around every pack launch of the region."""
    out = {}
    if st.get("pull_launches"):
        per = st["h2d_bytes"] / max(1, st["pull_launches"] + st["h2d_copies"])
        out["pull_kernel"] = {"launches": st["pull_launches"], "bytes_per_launch": per, "bound": "pcie",
                              "is": "pull_kernel<crc32c>: SM loads from the mapped pinned slot over PCIe + fused CRC-32C + store at the final place in HBM; "
                                    "PCIe traffic L, HBM traffic L (write only); launches of different workers overlap, so the link stays busy between them"}
        if secs and pcie_gib_s:
            ach = st["h2d_bytes"] / secs / GIB
            out["pull_kernel"].update({"achieved": ach, "peak": pcie_gib_s, "unit": "GiB/s", "frac": ach / pcie_gib_s,
                                       "peak_is": "cudaMemcpyAsync of 64 MiB from pinned memory (the link's best case; SM loads themselves top out at 47.8 GiB/s, tools/h2d_probe.cu)"})
    if st["pack_kernel_ns"]:
        ach = st["pack_kernel_bytes"] / st["pack_kernel_ns"]  # bytes per ns = GB/s
        out["pack_kernel"] = {"launches": st["pack_launches"], "algorithmic_bytes_per_launch": st["pack_kernel_bytes"] / max(1, st["pack_launches"]),
                              "us_per_launch": st["pack_kernel_ns"] / 1e3 / max(1, st["pack_launches"]), "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                              "is": "event pairs around every pack/checksum launch of the timed region (launch gaps included); batches are small because the pipeline is I/O-bound"}
    return out or None


# ---- ceilings ----------------------------------------------------------------------------------------------
def pcie_h2d_gib_s(dev, secs=0.4):
    import torch

    h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(dev)
    reps = 40
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize(dev)
    return reps * (64 << 20) / (a.elapsed_time(b) * 1e-3) / GIB


def source_ceiling(paths, thread_counts, o_direct=True, reps=1):
    """Best raw read rate of the file set over several shapes (no GPU work): the same read sizes and slot geometry
    as the product, deeper and shallower queues, READ_FIXED through the ring and preadv2 on the thread itself."""
    from lsio_b200 import cuda_ops

    best, shapes = 0.0, []
    for nthr in sorted(set(thread_counts)):
        for block, qd, inl, fif in ((512 << 10, 4, 1, 1), (512 << 10, 4, 0, 1), (512 << 10, 8, 0, 4), (1 << 20, 32, 0, 4), (256 << 10, 64, 0, 8), (2 << 20, 2, 1, 1)):
            try:
                v = max(cuda_ops.probe_read_ceiling(paths, block, nthr, qd, o_direct, inl, fif) for _ in range(reps))
            except Exception as ex:
                shapes.append({"threads": nthr, "block": block, "qd": qd, "inline": inl, "files_in_flight": fif, "error": repr(ex)})
                continue
            shapes.append({"threads": nthr, "block": block, "qd": qd, "inline": inl, "files_in_flight": fif, "gib_s": round(v, 2)})
            best = max(best, v)
    return best, shapes


# ---- CPU baselines (the only place the oracle is the thing being TIMED) ---------------------------------------
def cpu_ref_pass(paths, workers, sqpoll=True, mode=0):
    from oracle import oracle

    starts, ends = chunk_ranges()
    secs, nbytes, errors = oracle.bench_read_files(paths, starts, ends, workers, sqpoll=sqpoll, mode=mode)
    return secs, nbytes, errors


def cpu_configs(cores):
    """The reference's configurations (verdict r1 #4): the README one first, then the ones that use the whole box."""
    return [("readme_8_workers_sqpoll", 8, True), (f"{max(1, cores // 2)}_workers_sqpoll", max(1, cores // 2), True), (f"{cores}_workers_no_sqpoll", cores, False)]


def cpu_baseline_block(paths, source, cores):
    rates = {}
    for name, w, sq in cpu_configs(cores):
        if name in rates:
            continue
        cpu_ref_pass(paths[: max(8, len(paths) // 10)], w, sq)  # warm the rings / kernel threads
        s, b, errs = cpu_ref_pass(paths, w, sq)
        rates[name] = {"gib_s": b / s / GIB, "workers": w, "sqpoll": sq, "threads": w * (2 if sq else 1), "errors": errs, "secs": s}
    readme = rates["readme_8_workers_sqpoll"]
    best_name = max(rates, key=lambda k: rates[k]["gib_s"])
    s2, b2, _ = cpu_ref_pass(paths, 8, mode=1)
    return {"value": readme["gib_s"], "unit": "GiB/s", "cores": readme["threads"], "kind": "port",
            "sample": f"one pass over all {len(paths)} files ({readme['secs']:.2f} s) from {source}, README configuration: 8 worker threads + 8 SQPOLL kernel threads, into fresh host buffers; {readme['errors']} errors",
            "nproc": cores, "configs": rates, "best_config": best_name, "best_gib_s": rates[best_name]["gib_s"],
            # lsio_bench itself counts the 57,792-byte tail it never reads (main.rs:137, :181-183): +0.14 %
            "reference_style_gib_s": readme["gib_s"] * FILE_SIZE / (N_CHUNKS * BLOCK),
            "object_store_like_gib_s": b2 / s2 / GIB,
            "object_store_like_is": ("buffered pread per range, 8 threads" + ("; the source is tmpfs, so this is a page-cache-hot memcpy that no eviction can cool down "
                                     "(a disk-sourced, cache-evicted twin is in e2e_disk)" if source == "tmpfs" else "; page cache evicted first"))}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm
    need = args.files * FILE_SIZE
    directory, source = pick_data_dir(args, need)
    paths = write_files_cpu(directory, "lsio_bench", args.files, FILE_SIZE)
    cores = os.cpu_count() or 1
    # bound each step: estimate on a small sample, then size the sample for <= ~20 s per step
    est_s, est_b, _ = cpu_ref_pass(paths[: min(20, len(paths))], 8)
    rate = est_b / max(est_s, 1e-9)
    n_sample = len(paths) if need / rate <= 20 else max(20, int(20 * rate / FILE_SIZE))
    sample = paths[:n_sample]
    if source != "tmpfs":
        evict_page_cache(sample)
    # one pass per configuration; the timed steps then run the BEST one, the README one is always reported
    quick = {}
    for name, w, sq in cpu_configs(cores):
        if args.ref_workers:
            name, w, sq = f"{args.ref_workers}_workers_sqpoll", args.ref_workers, True
        if name in quick:
            continue
        cpu_ref_pass(sample[: max(8, len(sample) // 10)], w, sq)
        s, b, e = cpu_ref_pass(sample, w, sq)
        quick[name] = {"gib_s": b / s / GIB, "workers": w, "sqpoll": sq, "threads": w * (2 if sq else 1), "errors": e}
    best = max(quick, key=lambda k: quick[k]["gib_s"])
    w, sq = quick[best]["workers"], quick[best]["sqpoll"]
    for _ in range(args.warmup):
        cpu_ref_pass(sample, w, sq)
    secs, nbytes, errors = [], 0, 0
    for _ in range(args.steps):
        s, b, e = cpu_ref_pass(sample, w, sq)
        secs.append(s)
        nbytes += b
        errors += e
    value = nbytes / sum(secs) / GIB
    s2, b2, _ = cpu_ref_pass(sample, 8, mode=1)
    desc = f"{n_sample} of {args.files} files x {N_CHUNKS} blocks of {BLOCK} B per step, source {source}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GiB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "C2: lsio_bench README workload, 500 files x 41,214,400 B, 157 blocks of 262,144 B per file, into host buffers (CPU lsio_uring port)",
                   "source": source, "dir": directory, "files": args.files, "worker_threads": w, "sqpoll": sq, "errors": errors,
                   "timed_configuration": best, "timed_configuration_is": "the fastest of the configurations below on this box (one pass each); value is its K timed steps",
                   "configurations_one_pass_gib_s": quick,
                   "readme_configuration_gib_s": quick.get("readme_8_workers_sqpoll", {}).get("gib_s"),
                   "object_store_like_gib_s": b2 / s2 / GIB,
                   "object_store_like_is": "buffered pread per range, 8 threads" + (" (tmpfs: page-cache-hot by construction)" if source == "tmpfs" else ""),
                   "note": "the reference is Rust and cannot be built here; this is oracle/lsio_uring_ref.c, its C restatement"},
        "cpu_baseline": {"value": value, "unit": "GiB/s", "cores": quick[best]["threads"], "nproc": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- informative extra legs ------------------------------------------------------------------------------------
def codec_leg(hbm_peak):
    """SURVEY 8f-4, DESIGN 3.3: the codec layer behind the same path.  Kernel numbers with CUDA events on the launching
    stream; the pipeline number is wall clock from the first get_ranges to the last decoded, device-verified row.
    Shards are written with liblz4 (pyarrow), not with the oracle."""
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


def disk_leg(args, dev):
    """The same workload from a disk-backed directory (this pool: one virtio disk): O_DIRECT product path, the read
    ceiling of that disk, and the CPU paths on the same files with the page cache evicted first."""
    from oracle import oracle  # noqa: F401  (the CPU paths below)

    free = shutil.disk_usage(os.path.dirname(args.disk_dir.rstrip("/")) or "/").free
    if free < args.disk_files * FILE_SIZE + (4 << 30):
        return {"skipped": "not enough free space"}
    dpaths = write_files_cpu(args.disk_dir, "lsio_bench", args.disk_files, FILE_SIZE)
    dcrc = oracle_crc_table(args.disk_dir, "c2", dpaths, BLOCK, N_CHUNKS)
    evict_page_cache(dpaths)
    ids = list(range(len(dpaths)))
    with backend(args, dev, workers=min(args.workers, 8)) as u:
        leg = timed_leg(u, lambda: native_pass(u, dpaths, ids, dcrc), 3, 1, dev)
    ceil, shapes = source_ceiling(dpaths, [8, 16], True, reps=2)
    val = leg["bytes"] / leg["total_secs"] / GIB
    out = {"value": val, "unit": "GiB/s", "files": len(dpaths), "bad": leg["bad"], "read_ceiling_gib_s": ceil, "frac_of_ceiling": val / ceil if ceil else None,
           "ceiling_shapes": shapes, "dir": args.disk_dir, "max_dio_align": leg["stats"]["max_dio_align"]}
    s, b, e = cpu_ref_pass(dpaths, 8, True)
    out["cpu_lsio_uring_o_direct_gib_s"] = b / s / GIB
    evict_page_cache(dpaths)
    s, b, e = cpu_ref_pass(dpaths, 8, mode=1)
    out["cpu_object_store_like_cache_evicted_gib_s"] = b / s / GIB
    return out


def unloaded_latency(args, dev, paths, expect, n_requests=300):
    """What one request costs when nothing else is in flight (what a data loader that asks for one sample at a time sees):
    get_ranges of 8 consecutive 256 KiB blocks (2 MiB) of one file, wait for its 8 chunks, check their CRCs, drop them, next."""
    from lsio_b200.io_uring_cuda import Receiver

    starts, ends = chunk_ranges()
    lat = []
    with backend(args, dev) as u:
        rx = u.completion()
        for i in range(n_requests + 20):
            f, k0 = i % len(paths), (7 * i) % (N_CHUNKS - 8)
            ud = (np.uint64(f) << np.uint64(32)) | np.arange(k0, k0 + 8, dtype=np.uint64)
            t0 = time.perf_counter()
            u.get_ranges_arrays(paths[f], starts[k0:k0 + 8], ends[k0:k0 + 8], ud)
            got = 0
            while got < 8:
                rec = rx.recv_batch(8 - got, 30.0)
                assert (rec["status"] == 0).all() and (rec["crc32c"] == expect[f, (rec["chunk"]["user_data"] & np.uint64(0xFFFFFFFF)).astype(np.int64)]).all()
                got += len(rec)
                Receiver.drop_buffers(rec)
            if i >= 20:
                lat.append(time.perf_counter() - t0)
    lat = np.sort(np.array(lat)) * 1e3
    return {"requests": n_requests, "bytes_per_request": 8 * BLOCK, "mean_ms": float(lat.mean()), "p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]),
            "min_ms": float(lat[0]), "is": "one get_ranges of 8 x 256 KiB at a time through the Python mirror: submit -> 8 chunks received, CRC-checked and dropped"}


def two_device_check(args, n_files, paths, expect):
    """Single-process multi-device mode (one backend, two worker groups, files dealt round robin) on GPUs 0 and 1."""
    import torch

    import lsio_b200 as L

    if torch.cuda.device_count() < 2:
        return {"skipped": "fewer than 2 visible devices"}
    sub = paths[:n_files]
    ids = list(range(n_files))
    with L.IoUringCuda(4, devices=(0, 1), arena_bytes=4 << 30, staging_bytes=256 << 20) as u:
        s, b, ok, bad = u.bench_read_files(sub, ids, BLOCK, N_CHUNKS, expect)
        st = u.stats()
    return {"devices": [0, 1], "files": n_files, "chunks_ok": ok, "bad_chunks": bad + (n_files * N_CHUNKS - ok - bad), "gib_s": b / s / GIB, "requests": st["requests"]}


# ---- main ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dir", default=None, help="directory for the synthetic files (default: /dev/shm if it has room)")
    ap.add_argument("--files", type=int, default=N_FILES)
    ap.add_argument("--workers", type=int, default=0, help="io worker threads per GPU (0 = as many as this rank's CPU share allows, at most 14)")
    ap.add_argument("--ref-workers", type=int, default=0)
    ap.add_argument("--staging-mib", type=int, default=1024)
    ap.add_argument("--arena-gib", type=int, default=8)
    ap.add_argument("--min-copy-mib", type=int, default=64)
    ap.add_argument("--pack-threads", type=int, default=0, help="threads per CTA of the pipeline's roofline-variant pack launches (0 = 1024)")
    ap.add_argument("--pack-light-below-mib", type=int, default=512)
    ap.add_argument("--flush-us", type=int, default=2000, help="a pack batch is launched once it holds --min-copy-mib or its oldest extent has waited this long")
    ap.add_argument("--sq-entries", type=int, default=64)
    ap.add_argument("--max-read-kib", type=int, default=512, help="reads (merged or split) are cut at this size")
    ap.add_argument("--merge-ranges", type=int, default=1)
    ap.add_argument("--poll-sleep-us", type=int, default=50, help="device-thread poll interval in microseconds (0 = spin)")
    ap.add_argument("--staging-mode", type=int, default=1, help="1 = per-worker cache-sized pinned slots (default), 0 = one big pinned ring")
    ap.add_argument("--host-slots", type=int, default=0, help="staging mode 1: pinned slots per worker (0 = automatic, from the window budget)")
    ap.add_argument("--host-slot-kib", type=int, default=0, help="staging mode 1: slot (extent, H2D copy) size")
    ap.add_argument("--inline-reads", type=int, default=1)
    ap.add_argument("--direct-landing", type=int, default=2, help="0 = ring + pack kernel, 1 = copy engine straight into the arena, 2 = pull kernel")
    ap.add_argument("--steal-batch", type=int, default=8)
    ap.add_argument("--completion-bound", type=int, default=0)
    ap.add_argument("--copy-coalesce-slots", type=int, default=2)
    ap.add_argument("--pull-any-shift", type=int, default=1)
    ap.add_argument("--window-budget-kib", type=int, default=0, help="pinned slot windows of one rank together (0 = 60 %% of the L3 / number of ranks)")
    ap.add_argument("--deal", type=int, default=1, help="1 = ranks own disjoint physical files (default), 0 = logical file j is physical file j %% files")
    ap.add_argument("--pin", type=int, default=1, help="N > 1: 1 = pin every rank's backend threads to its own slice of the CPUs, 2 = slices handed out in reverse rank order, 0 = no pinning")
    ap.add_argument("--zero-copy", type=int, default=0, help="1 = the pack kernel reads the pinned ring over PCIe (no H2D memcpy)")
    ap.add_argument("--statx-via-ring", type=int, default=0, help="1 = the reference's openat+statx SQE pair per file")
    ap.add_argument("--disk-dir", default="/tmp/lsio_bench_disk", help="secondary e2e leg on a disk-backed directory ('' disables)")
    ap.add_argument("--disk-files", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-codec", action="store_true", help="skip the informative codec-layer leg (N=1 only)")
    ap.add_argument("--no-extra", action="store_true", help="only the contract legs (value, e2e, roofline, ceilings)")
    ap.add_argument("--kernel-only", "--value-only", dest="kernel_only", action="store_true", help="only the device-resident kernel pass (for ncu captures)")
    ap.add_argument("--kernel-in-place", action="store_true", help="--kernel-only: the crc-only in-place form")
    ap.add_argument("--e2e-only", action="store_true", help="only the native end-to-end passes (pipeline tuning sweeps, ncu captures of pipeline launches)")
    ap.add_argument("--null-source", type=int, default=0, help="--e2e-only: workers skip the reads (device stage alone)")
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"], help="C2 is the contract line; the others print only that leg")
    ap.add_argument("--c3-files", type=int, default=10_000)
    ap.add_argument("--c5-files", type=int, default=16)
    ap.add_argument("--c5-total-gb", type=float, default=1000.0, help="C5: bytes landed per step over all ranks (looped over the file set)")
    ap.add_argument("--c5", type=int, default=-1, help="run the C5 leg: -1 = only at 8 GPUs, 0 = never, 1 = always")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch

    import lsio_b200  # noqa: F401  fails loudly if the CUDA extension is missing
    from lsio_b200 import cuda_ops, sharding

    rank, world, local = dist_env()
    n_gpus = max(world, 1)
    args.cpu_first, args.cpu_count, cpus_per_rank = cpu_slice(rank, n_gpus, args.pin == 2)
    if args.workers <= 0:
        # reads of memory-backed files run ON the worker (inline preadv2): one worker per CPU of this rank's share;
        # alone on the box a couple of CPUs stay free for the consumer and the device thread
        args.workers = max(2, min(12, cpus_per_rank - 4)) if n_gpus == 1 else max(2, min(12, cpus_per_rank))
    if args.window_budget_kib <= 0:  # this rank's share of the last-level cache (all ranks share one host)
        args.window_budget_kib = max(1024, int(llc_bytes() * 0.6 / n_gpus) >> 10)
    if args.completion_bound <= 0:
        args.completion_bound = 1 << 16
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist_init("nccl")

    need = args.files * FILE_SIZE
    directory, source = pick_data_dir(args, need)
    t_gen = time.time()
    if rank == 0:
        paths = write_files_cpu(directory, "lsio_bench", args.files, FILE_SIZE)
        expect = oracle_crc_table(directory, "c2", paths, BLOCK, N_CHUNKS)
    barrier()
    if rank != 0:
        paths = write_files_cpu(directory, "lsio_bench", args.files, FILE_SIZE)  # already there
        expect = oracle_crc_table(directory, "c2", paths, BLOCK, N_CHUNKS)        # cached by rank 0
    t_gen = time.time() - t_gen
    # weak scaling: every rank lands `files` files per step (20.6 GB).  Storage is bounded by re-using the one physical file
    # set: the physical files are DEALT file_idx mod n_gpus and a rank walks its own share files/n_gpus times
    # per step, so that no two ranks ever read the same file (two cores hammering the same page-cache folios' reference
    # counts is an artefact a real, n_gpus times larger data set would not have).  --deal 0: the round-1 dealing (logical
    # file j = physical file j % files), which at 8 ranks makes ranks r and r+4 walk the same files in lockstep.
    my_ids = deal_files(args.files, rank, n_gpus, args.deal)
    my_paths = [paths[j] for j in my_ids]

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    if args.kernel_only:
        k = run_kernel(args, dev, args.files, expect, in_place=args.kernel_in_place)
        ms = statistics.mean(k["step_ms"])
        print(json.dumps({"kernel_only": True, "in_place": args.kernel_in_place, "files": args.files, "kernel_ms": ms, "achieved_gbs": k["algo_bytes"] / ms / 1e6,
                          "frac": k["algo_bytes"] / ms / 1e6 / hbm_peak, "crc_equals_oracle_table": k["crc_ok"]}), flush=True)
        return 0

    if args.e2e_only:
        with backend(args, dev, null_source=args.null_source) as u:
            leg = timed_leg(u, lambda: native_pass(u, my_paths, my_ids, None if args.null_source else expect), args.steps, args.warmup, dev)
        s = leg_summary(leg, dev, args.workers, n_gpus)
        if rank == 0:
            print(json.dumps({"e2e_only": True, "null_source": args.null_source, "n_gpus": n_gpus, "workers": args.workers, "pin": [args.cpu_first, args.cpu_count],
                              "pipeline_kernel": pipeline_kernel(leg["stats"], hbm_peak), **s}), flush=True)
        return 0

    if args.config != "C2":
        return run_config_only(args, dev, rank, n_gpus, directory, source, paths, hbm_peak)

    # ---- the contract legs -------------------------------------------------------------------------------
    with ClockSampler(local) as clocks:
        with backend(args, dev) as u:
            val = timed_leg(u, lambda: native_pass(u, my_paths, my_ids, expect), args.steps, args.warmup, dev)
        with backend(args, dev) as u:
            e2e = timed_leg(u, lambda: python_pass(u, my_paths, my_ids, expect, dev), max(2, min(args.steps, 10)), 1, dev)
        kern = run_kernel(args, dev, args.files, expect)
        kern_ip = run_kernel(args, dev, args.files, expect, in_place=True)
    clk = clocks.summary()
    v = leg_summary(val, dev, args.workers, n_gpus)
    e = leg_summary(e2e, dev, args.workers, n_gpus)

    # the device stage alone: workers skip the read (verdict r1 #2b) -- separates "the box cannot feed N GPUs" from "the pipeline cannot"
    with backend(args, dev, null_source=1) as u:
        nul = timed_leg(u, lambda: native_pass(u, my_paths, my_ids, None), 2, 1, dev)
    nl = leg_summary(nul, dev, args.workers, n_gpus)

    # ceilings of the I/O roofline: PCIe with all ranks copying at once; the source with every other rank idle
    barrier()
    pcie_each = gather_floats(pcie_h2d_gib_s(dev), dev)
    barrier()
    ceil = {"pcie_h2d_gib_s_all_ranks_at_once": sum(pcie_each), "pcie_h2d_gib_s_per_gpu": [round(x, 2) for x in pcie_each]}
    if rank == 0:
        ceil["pcie_h2d_gib_s_one_gpu_alone"] = pcie_h2d_gib_s(dev)
        ncpu = len(os.sched_getaffinity(0))
        ceil["source_read_gib_s"], ceil["source_read_shapes"] = source_ceiling(paths[: min(len(paths), 250)], [args.workers * n_gpus, ncpu], True)
    barrier()
    # what the host can stage AND ship at once: the minimal staged program (preadv2 into pinned slots + pull kernel, nothing else), every
    # rank at the same time with the product's thread count, slot geometry and CPU slice, over the rank's whole shard; best of three rounds
    staged = 0.0
    geo = val["stats"]
    for _ in range(3):
        barrier()
        try:
            t, b = cuda_ops.probe_staged_ceiling(my_paths, dev.index, args.workers, max(64, int(geo["slot_bytes"]) >> 10), max(2, int(geo["slots_per_worker"])), False,
                                                 args.cpu_first if args.pin else 0, args.cpu_count if args.pin else 0)
        except Exception as ex:
            log("staged ceiling probe failed:", repr(ex))
            t, b = 1.0, 0
        b, t = reduce_sum_max(b, t, dev)
        staged = max(staged, b / t / GIB)
    barrier()
    ceil["staged_ceiling_gib_s"] = staged

    parity = {"expected_crc_source": "oracle", "expected_crc_is": "oracle.crc32c_file_blocks over files written by oracle.write_file (oracle/lsio_oracle.c); every chunk of every timed pass is compared with it",
              "chunks_checked_timed_legs_per_rank": (args.steps + e["steps"]) * len(my_paths) * N_CHUNKS,
              "bad_chunks_all_ranks": v["bad_chunks"] + e["bad_chunks"],
              "kernel_pass_crc_equals_oracle_table": kern["crc_ok"] and kern_ip["crc_ok"]}
    if not args.no_verify:
        parity["landed_bytes_vs_file_bytes"] = verify_landed_bytes(args, dev, my_paths, my_ids, 10)
        mm, _ = reduce_sum_max(parity["landed_bytes_vs_file_bytes"]["mismatched_bytes"], 0.0, dev)
        parity["landed_bytes_vs_file_bytes"]["mismatched_bytes_all_ranks"] = mm

    extra = {}
    if not args.no_extra:
        if n_gpus > 1:
            extra["c4"] = c4_leg(args, dev, rank, n_gpus, directory, paths)
        if args.c5 == 1 or (args.c5 < 0 and n_gpus == 8):
            extra["c5"] = c5_leg(args, dev, rank, n_gpus, directory)
        if n_gpus > 1:
            barrier()
            if rank == 0:
                try:
                    extra["two_devices_one_process"] = two_device_check(args, 40, paths, expect)
                except Exception as ex:
                    extra["two_devices_one_process"] = {"error": repr(ex)}
            barrier()
        if n_gpus == 1:
            try:
                extra["unloaded_latency"] = unloaded_latency(args, dev, paths, expect)
            except Exception as ex:
                extra["unloaded_latency"] = {"error": repr(ex)}
            try:
                extra["c3"] = c3_leg(args, dev, rank, n_gpus, directory)
            except Exception as ex:
                extra["c3"] = {"error": repr(ex)}
            if args.disk_dir and args.disk_files > 0 and source != "disk":
                try:
                    extra["e2e_disk"] = disk_leg(args, dev)
                except Exception as ex:  # informative only
                    extra["e2e_disk"] = {"error": repr(ex)}
            if not args.no_codec:
                try:
                    extra["codec"] = codec_leg(hbm_peak)
                except Exception as ex:  # informative only: never costs the contract line
                    extra["codec"] = {"error": repr(ex)}

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block(paths, source, os.cpu_count() or 1)

    if rank == 0:
        kernel_ms = statistics.mean(kern["step_ms"])
        achieved = kern["algo_bytes"] / (kernel_ms * 1e-3) / 1e9
        ip_ms = statistics.mean(kern_ip["step_ms"])
        pcie = ceil["pcie_h2d_gib_s_all_ranks_at_once"]
        io_ceiling = min(ceil["source_read_gib_s"], pcie)
        per_step = len(my_paths) * N_CHUNKS * BLOCK
        io_roofline = {"bound": "source" if ceil["source_read_gib_s"] < pcie else "pcie", "achieved": v["gib_s"], "peak": io_ceiling, "unit": "GiB/s",
                       "frac": v["gib_s"] / io_ceiling, "frac_of_staged_ceiling": (v["gib_s"] / ceil["staged_ceiling_gib_s"]) if ceil.get("staged_ceiling_gib_s") else None,
                       "staged_ceiling_is": "the minimal host-staged program on the same files, threads, slots and CPU slices, all ranks at once: preadv2 into pinned slots + pull kernel per slot, "
                                            "no requests / planning / arena / channel / checksums (lsio_probe_staged_ceiling): what this host can stage AND ship at once; the source ceiling "
                                            "is measured with no DMA sharing the memory system",
                       "null_source_gib_s": nl["gib_s"], "null_source_frac_of_pcie": nl["gib_s"] / pcie,
                       "null_source_is": "same pipeline, workers skip the read itself (slots keep stale bytes): the device stage alone", **ceil}
        line = {
            "metric": METRIC, "value": v["gib_s"], "unit": "GiB/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": v["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {
                "workload": "C2: lsio_bench README workload, 500 files x 41,214,400 B, 157 blocks of 262,144 B per file (78,500 ranges, 20,578,304,000 B) per GPU, delivered into B200 HBM, bit-exact",
                "files": args.files, "ranges_per_step_per_gpu": len(my_paths) * N_CHUNKS, "bytes_per_step_per_gpu": per_step, "source": source, "dir": directory,
                "value_is": "files -> reads into pinned slots -> H2D -> checksum/pack kernel -> one chunk per range through the C ABI (native lsio_bench driver), fused CRC-32C of every chunk == oracle table; sum of bytes over ranks / max time over ranks",
                "e2e_is": "the same through the Python mirror of the reference API (get_ranges + recv_batch), CRCs compared in numpy",
                "l2": "inputs larger than L2 (20.6 GB land per step vs 126 MB)", "worker_threads_per_gpu": args.workers,
                "cpu_pinning": {"cpu_first": args.cpu_first, "cpu_count": args.cpu_count} if args.pin and args.cpu_count else None, "nproc": os.cpu_count(),
                "multi_gpu": "physical files dealt file_idx mod n_gpus, no collective on the data path; each rank lands 500 files per step by walking its own share n_gpus times (no file is read by two ranks)",
                "file_generation_s": round(t_gen, 1), "files_written_by": "oracle.write_file",
            },
            "e2e": {"value": e["gib_s"], "unit": "GiB/s", "h2d_bytes_per_step": per_step, "d2h_bytes_per_step": 4 * len(my_paths) * N_CHUNKS,
                    "ms_per_step": e["ms_per_step"], "steps": e["steps"], "bad_chunks": e["bad_chunks"], "per_rank_gib_s": e["per_rank_gib_s"], "io_roofline": io_roofline},
            "gpu_launches": val["stats"]["kernel_launches"],
            "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "pack_kernel<store,crc32c> + fold_kernel, device-resident: one launch over all 78,500 ranges per step", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": kern["algo_bytes"],
                         "in_place_crc_only": {"kernel": "pack_kernel<crc32c only> (the form run on directly landed extents)", "kernel_ms": ip_ms, "algorithmic_bytes_per_launch": kern_ip["algo_bytes"],
                                               "achieved": kern_ip["algo_bytes"] / ip_ms / 1e6, "frac": kern_ip["algo_bytes"] / ip_ms / 1e6 / hbm_peak},
                         "pipeline_kernel": pipeline_kernel(val["stats"], hbm_peak, val["total_secs"], ceil["pcie_h2d_gib_s_one_gpu_alone"])},
            "cpu_baseline": cpu,
            "io_roofline": io_roofline,
            "value_detail": {k: v[k] for k in ("per_rank_gib_s", "bad_chunks", "latency_rank0", "pipeline_rank0", "per_worker_rank0", "read_amplification")}
                            | {"cuda_event_secs_rank0": val["cuda_event_secs"], "wall_secs_rank0": val["total_secs"]},
            "null_source": {k: nl[k] for k in ("gib_s", "per_rank_gib_s", "per_worker_rank0")},
            "parity": parity,
            **extra,
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


# ---- C3 / C4 / C5 legs (aggregated over ranks like the headline) -------------------------------------------------
def c3_leg(args, dev, rank, n_gpus, directory):
    from lsio_b200 import sharding

    d3 = directory + "_c3"
    if rank == 0:
        paths, per, size, desc = c3_workload(args, d3)
        exp = oracle_crc_of_ranges(d3, "c3", paths, per, size)
    barrier()
    if rank != 0:
        paths, per, size, desc = c3_workload(args, d3)
        exp = oracle_crc_of_ranges(d3, "c3", paths, per, size)
    base = np.cumsum([0] + [len(p[0]) for p in per])
    mine = sharding.files_for_rank(len(paths), rank, n_gpus)
    my_paths = [paths[j] for j in mine]
    my_per = [per[j] for j in mine]
    my_exp = np.concatenate([exp[base[j]:base[j + 1]] for j in mine])
    payload = 0
    for s, e in my_per:
        so, eo = resolve(size, s, e)
        payload += int((eo - so).sum())
    leg = ranges_leg(args, dev, my_paths, my_per, my_exp, payload, 3, 2)
    out = leg_summary(leg, dev, args.workers, n_gpus)
    out.update({"config": desc, "expected_crc_source": "oracle", "ranges_per_step": sum(len(p[0]) for p in my_per)})
    return out


def c4_leg(args, dev, rank, n_gpus, directory, paths):
    """BASELINE configs[3]: files sharded over the GPUs, 1 MiB ranges (39 per file; the 319,488-byte tail is not read)."""
    from lsio_b200 import sharding

    k = FILE_SIZE // MIB
    if rank == 0:
        exp = oracle_crc_table(directory, "c4", paths, MIB, k)
    barrier()
    if rank != 0:
        exp = oracle_crc_table(directory, "c4", paths, MIB, k)
    st = np.arange(k, dtype=np.int64) << 20
    phys = sharding.files_for_rank(len(paths), rank, n_gpus)  # ranks own disjoint physical files, walked n_gpus times per step
    mine = [phys[i % len(phys)] for i in range(len(paths))]
    my_paths = [paths[j] for j in mine]
    my_exp = np.concatenate([exp[j] for j in mine])
    per = [(st, st + MIB)] * len(my_paths)
    leg = ranges_leg(args, dev, my_paths, per, my_exp, len(my_paths) * k * MIB, 3, 1)
    out = leg_summary(leg, dev, args.workers, n_gpus)
    out.update({"config": f"C4: {len(paths)} files x {FILE_SIZE} B landed per GPU per step (physical files dealt file_idx mod {n_gpus}, each rank walks its share {n_gpus} times), {k} ranges of 1 MiB per file", "expected_crc_source": "oracle"})
    return out


def c5_leg(args, dev, rank, n_gpus, directory):
    """BASELINE configs[4]: large streaming scan, 4 MiB ranges, fused crc32c checked per chunk; the file set is looped until
    ~c5_total_gb have landed per step over all ranks."""
    d5 = directory + "_c5"
    size = 1 << 30
    try:
        st = os.statvfs(os.path.dirname(d5))
        room = st.f_bavail * st.f_frsize
    except OSError:
        room = 0
    n_files = args.c5_files
    have = os.path.isdir(d5) and len([f for f in os.listdir(d5) if f.startswith("c5_")]) >= n_files
    if not have and room < n_files * size + (8 << 30):
        n_files = max(2, int((room - (8 << 30)) // size))
    if rank == 0:
        paths = write_files_cpu(d5, "c5", n_files, size, 200_000)
        exp = oracle_crc_table(d5, "c5", paths, 4 * MIB, size // (4 * MIB))
    barrier()
    if rank != 0:
        paths = write_files_cpu(d5, "c5", n_files, size, 200_000)
        exp = oracle_crc_table(d5, "c5", paths, 4 * MIB, size // (4 * MIB))
    st4 = np.arange(size >> 22, dtype=np.int64) << 22
    per = [(st4, st4 + 4 * MIB)] * n_files
    # every rank scans its own files (dealt file_idx mod n_gpus), looped
    order = [i for i in range(n_files) if i % n_gpus == rank] or [rank % n_files]
    my_paths = [paths[i] for i in order]
    my_exp = np.concatenate([exp[i] for i in order])
    per = per[:len(order)]
    loops = max(1, int(round(args.c5_total_gb * 1e9 / n_gpus / (len(order) * size))))
    leg = ranges_leg(args, dev, my_paths, per, my_exp, len(order) * size, 1, 1, loops=loops, arena_bytes=max(args.arena_gib, 12) << 30)
    out = leg_summary(leg, dev, args.workers, n_gpus)
    out.update({"config": f"C5: {n_files} files x 1 GiB dealt file_idx mod {n_gpus} ({len(order)} per rank), 256 ranges of 4 MiB per file, {loops} loops per rank per step",
                "loops_per_rank": loops, "bytes_landed_all_ranks": out["bytes"], "expected_crc_source": "oracle", "fused_crc32c_checked_per_chunk": True})
    return out


def run_config_only(args, dev, rank, n_gpus, directory, source, paths, hbm_peak):
    if args.config == "C3":
        out = c3_leg(args, dev, rank, n_gpus, directory)
    elif args.config == "C4":
        out = c4_leg(args, dev, rank, n_gpus, directory, paths)
    else:
        out = c5_leg(args, dev, rank, n_gpus, directory)
    if rank == 0:
        out.update({"n_gpus": n_gpus, "source": source, "workers_per_gpu": args.workers})
        print(json.dumps(out), flush=True)
    barrier()
    return 0


if __name__ == "__main__":
    sys.exit(main())
