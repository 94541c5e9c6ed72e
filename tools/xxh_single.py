"""One launch configuration of the xxh64 kernels for ncu: python tools/xxh_single.py <n_chunks> <chunk_kib>  (LSIO_XXH64_VARIANT=1 plain, 2 staged)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lsio_b200
from lsio_b200 import cuda_ops
n, kib = int(sys.argv[1]), int(sys.argv[2])
buf = torch.empty(n * (kib << 10), dtype=torch.uint8, device="cuda"); cuda_ops.fill_pattern(buf, 3, 0)
offs = np.arange(n, dtype=np.uint64) * np.uint64(kib << 10); lens = np.full(n, kib << 10, dtype=np.uint64)
for _ in range(4):
    out = cuda_ops.xxh64(buf, offs, lens)
torch.cuda.synchronize()
print("ok", int(out[0]))
