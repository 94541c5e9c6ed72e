"""Times the pack kernel variant selected by LSIO_B200_LIB on a few shapes (A/B harness)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lsio_b200
from lsio_b200 import cuda_ops
dev = torch.device("cuda:0")

def timeit(nseg, seglen, store=True, crc=True, iters=10, misalign=0):
    src = torch.empty(nseg * seglen + 4096, dtype=torch.uint8, device=dev); cuda_ops.fill_pattern(src, 1, 0)
    dst = torch.empty(nseg * seglen + 4096, dtype=torch.uint8, device=dev)
    off = np.arange(nseg, dtype=np.uint64) * np.uint64(seglen)
    lens = np.full(nseg, seglen - (16 if misalign else 0), dtype=np.uint64)
    plan = cuda_ops.PackPlan(0, off + np.uint64(src.data_ptr() + misalign), off + np.uint64(dst.data_ptr()), lens, store=store, crc32c=crc)
    out = torch.empty(nseg, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): plan.launch(out, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(iters): plan.launch(out, st)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ab = plan.algorithmic_bytes; plan.close()
    return ab / ms / 1e6 / 6538.9 * 100

name = os.path.basename(os.environ.get("LSIO_B200_LIB", "default"))
res = [timeit(32768, 256 << 10), timeit(2048, 4 << 20), timeit(131072, 64 << 10), timeit(32768, 256 << 10, misalign=5), timeit(32768, 256 << 10, crc=False)]
print(f"{name:18s} 256K:{res[0]:5.1f}%  4M:{res[1]:5.1f}%  64K:{res[2]:5.1f}%  256K-misaligned:{res[3]:5.1f}%  256K-copyonly:{res[4]:5.1f}%")
