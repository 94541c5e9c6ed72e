"""Throughput of the xxh64_chunks kernel (read-only traffic: algorithmic bytes = sum(len))."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lsio_b200
from lsio_b200 import cuda_ops, _native as N
dev = torch.device("cuda:0")
for nseg, seglen in ((32768, 256 << 10), (131072, 64 << 10), (2048, 4 << 20), (524288, 16 << 10)):
    buf = torch.empty(nseg * seglen, dtype=torch.uint8, device=dev); cuda_ops.fill_pattern(buf, 3, 0)
    offs = np.arange(nseg, dtype=np.uint64) * np.uint64(seglen); lens = np.full(nseg, seglen, dtype=np.uint64)
    for _ in range(2): out = cuda_ops.xxh64(buf, offs, lens)
    st = torch.cuda.Stream()
    ptrs = (offs + np.uint64(buf.data_ptr())).astype(np.uint64)
    outt = torch.empty(nseg, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # lsio_cuda_xxh64 uploads ptrs/lens and synchronises; time the whole call with events around it (upload is tiny)
    e0.record()
    for _ in range(5):
        N.check(N.lib.lsio_cuda_xxh64(0, ptrs.ctypes.data_as(N.c_u64p), lens.ctypes.data_as(N.c_u64p), nseg, ctypes.c_void_p(outt.data_ptr()), None), "xxh64")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"xxh64 nseg={nseg} seglen={seglen >> 10}KiB: {ms:.3f} ms  {nseg * seglen / ms / 1e6:.1f} GB/s read  ({nseg * seglen / ms / 1e6 / 6538.9 * 100:.1f}% of measured HBM peak)")
    del buf
