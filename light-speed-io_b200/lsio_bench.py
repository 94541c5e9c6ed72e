"""`lsio_bench` for the B200 backend: same flags and output as the reference's CLI benchmark
(crates/lsio_bench/src/main.rs:19-42 for the arguments, :85-112 file creation, :124-190 the timed
read, :192-199 page-cache eviction), with the chunks landing in GPU HBM.

    python -m lsio_b200.lsio_bench --directory /dev/shm/x --nrfiles 500 --filesize 41214400 \
        --blocksize 262144 --nr-worker-threads 8

Differences, both deliberate: the files hold the reference's own pattern (`byte[i] = i as u8`,
main.rs:102) but every chunk's device CRC-32C is checked against the CRC of that pattern; bandwidth is
reported both the reference's way (filesize * n_files, which over-counts the skipped tail: SURVEY F4)
and from the bytes actually delivered.  Page cache is evicted with posix_fadvise(DONTNEED) (no vmtouch).
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

FILENAME_PREFIX = "lsio_bench_"  # main.rs:16
MEBIBYTE = float(1024 * 1024)


def create_files_if_necessary(filenames, filesize):  # main.rs:85-112
    print(f"Creating {len(filenames)} files (if necessary), each of filesize {filesize} bytes...")
    contents = None
    for fn in filenames:
        if os.path.exists(fn) and os.path.getsize(fn) == filesize:
            continue
        if contents is None:
            contents = (np.arange(filesize, dtype=np.uint64) & 0xFF).astype(np.uint8)  # (0..filesize).map(|i| i as u8)
        with open(fn, "wb") as f:
            f.write(memoryview(contents))
            f.flush()
            os.fsync(f.fileno())


def clear_page_cache(filenames):  # main.rs:192-199 (vmtouch -e) restated with posix_fadvise
    print("Clearing page cache...")
    for fn in filenames:
        fd = os.open(fn, os.O_RDONLY)
        try:
            os.posix_fadvise(fd, 0, 0, os.POSIX_FADV_DONTNEED)
        finally:
            os.close(fd)


def crc32c_of_ref_pattern_blocks(n_chunks, blocksize, device):
    """Expected CRC per block, from the crc-only kernel over a generated image of one file."""
    import torch

    from . import cuda_ops

    n = n_chunks * blocksize
    img = (torch.arange(n, dtype=torch.int64, device=f"cuda:{device}") & 0xFF).to(torch.uint8)
    offs = np.arange(n_chunks, dtype=np.uint64) * np.uint64(blocksize)
    crc = cuda_ops.pack_strip_crc32c(img, None, offs, offs, np.full(n_chunks, blocksize, dtype=np.uint64), store=False)
    return crc.cpu().numpy().view(np.uint32)


def read_files(filenames, filesize, blocksize, n_worker_threads, device=0):  # main.rs:124-190
    from . import IoUringCuda

    blocksize = blocksize or filesize
    n_chunks = filesize // blocksize  # main.rs:137
    n_files = len(filenames)
    expect = np.tile(crc32c_of_ref_pattern_blocks(n_chunks, blocksize, device), (1, 1)) if n_chunks else None
    with IoUringCuda(n_worker_threads, devices=(device,)) as uring:
        print(f"Performing read benchmark for {n_files} files x {n_chunks} chunks per file = {n_files * n_chunks} total chunks:")
        secs, nbytes, ok, bad = uring.bench_read_files(filenames, np.zeros(n_files, dtype=np.uint32), blocksize, n_chunks, expect)
    total_bytes = float(filesize * n_files)  # main.rs:181-183 (the reference's accounting)
    print(f"Total runtime: {secs} secs")
    print(f"Total mebibytes: {total_bytes / MEBIBYTE} MiB")
    print(f"Total bandwidth = {total_bytes / secs / MEBIBYTE} mebibytes per sec")
    print(f"Delivered into HBM: {nbytes} bytes in {ok} chunks ({bad} bad) = {nbytes / secs / MEBIBYTE} mebibytes per sec")
    return secs, nbytes, ok, bad


def main(argv=None):
    ap = argparse.ArgumentParser(description="lsio_bench for the B200 backend (same flags as the reference's lsio_bench)")
    ap.add_argument("-d", "--directory", default=None, help="Prefix filenames with this directory (default: the system's temporary directory)")
    ap.add_argument("-n", "--nrfiles", type=int, default=1)
    ap.add_argument("-f", "--filesize", type=int, default=1024 * 1024)
    ap.add_argument("-b", "--blocksize", type=int, default=None, help="chunk size in bytes (default: the filesize)")
    ap.add_argument("-w", "--nr-worker-threads", type=int, default=4)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    if a.nrfiles < 1 or a.filesize < 1 or (a.blocksize is not None and a.blocksize < 1) or not (1 <= a.nr_worker_threads < 1024):
        ap.error("nrfiles, filesize and blocksize must be >= 1; nr-worker-threads in 1..1023")
    directory = a.directory or tempfile.gettempdir()
    if not os.path.isdir(directory):
        ap.error(f"Directory {directory!r} does not exist, or is not a directory")
    filenames = [os.path.join(os.path.abspath(directory), f"{FILENAME_PREFIX}{i}") for i in range(a.nrfiles)]
    create_files_if_necessary(filenames, a.filesize)
    clear_page_cache(filenames)
    _, _, _, bad = read_files(filenames, a.filesize, a.blocksize, a.nr_worker_threads, a.device)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
