"""First-light check of the pack/crc kernel on a GPU: parity vs the oracle, then timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lsio_b200
from lsio_b200 import cuda_ops
from oracle import oracle

dev = torch.device("cuda:0")
rng = np.random.default_rng(1)

def check(nseg, maxlen, align_src=1, align_dst=1, tight=True):
    lens = rng.integers(0, maxlen, size=nseg).astype(np.uint64)
    lens[: min(8, nseg)] = [0, 1, 2, 3, 4, 5, 15, 16][: min(8, nseg)]
    src_off, dst_off, so, do = [], [], 0, 0
    for L in lens:
        head = int(rng.integers(0, 512)) if align_src == 1 else 0
        so = (so + 511) // 512 * 512
        src_off.append(so + head); so += head + int(L)
        do = (do + align_dst - 1) // align_dst * align_dst
        dst_off.append(do); do += int(L)
    src_size = (so + 1023) // 512 * 512; dst_size = (do + 1023) // 512 * 512
    src = rng.integers(0, 256, size=src_size, dtype=np.uint8)
    want_dst, want_crc, _ = oracle.pack(src, src_off, dst_off, lens, dst_size)
    tsrc = torch.from_numpy(src).to(dev); tdst = torch.zeros(dst_size, dtype=torch.uint8, device=dev)
    crc = cuda_ops.pack_strip_crc32c(tsrc, tdst, src_off, dst_off, lens)
    got_dst = tdst.cpu().numpy(); got_crc = crc.cpu().numpy().view(np.uint32)
    ok_b = np.array_equal(got_dst, want_dst); ok_c = np.array_equal(got_crc, want_crc)
    # crc-only, in place
    crc2 = cuda_ops.pack_strip_crc32c(tsrc, None, src_off, dst_off, lens, store=False).cpu().numpy().view(np.uint32)
    ok_c2 = np.array_equal(crc2, want_crc)
    # store-only
    tdst2 = torch.zeros(dst_size, dtype=torch.uint8, device=dev)
    cuda_ops.pack_strip_crc32c(tsrc, tdst2, src_off, dst_off, lens, crc32c=False)
    ok_b2 = np.array_equal(tdst2.cpu().numpy(), want_dst)
    print(f"nseg={nseg} maxlen={maxlen} asrc={align_src} adst={align_dst}: bytes={ok_b} crc={ok_c} crc_only={ok_c2} store_only={ok_b2}")
    if not ok_c:
        bad = np.nonzero(got_crc != want_crc)[0]
        print("  bad crc idx", bad[:10], "lens", lens[bad[:10]], "so%16", [src_off[i] % 16 for i in bad[:10]], "do%16", [dst_off[i] % 16 for i in bad[:10]])
    if not ok_b:
        bad = np.nonzero(got_dst != want_dst)[0]; print("  first bad byte", bad[:5], len(bad))
    return ok_b and ok_c and ok_c2 and ok_b2

allok = True
allok &= check(64, 100)
allok &= check(200, 5000)
allok &= check(300, 70000)
allok &= check(50, 70000, align_src=512, align_dst=512)
allok &= check(20, 3 << 20)
allok &= check(8, 40 << 20, align_src=512, align_dst=16)
print("PARITY", "OK" if allok else "FAIL")

def timeit(nseg, seglen, store=True, crc=True, iters=5):
    src = torch.randint(0, 255, (nseg * seglen,), dtype=torch.uint8, device=dev)
    dst = torch.empty_like(src) if store else None
    off = np.arange(nseg, dtype=np.uint64) * np.uint64(seglen)
    lens = np.full(nseg, seglen, dtype=np.uint64)
    plan = cuda_ops.PackPlan(0, off + np.uint64(src.data_ptr()), off + np.uint64(dst.data_ptr() if store else 0), lens, store=store, crc32c=crc)
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
    ab = plan.algorithmic_bytes
    print(f"nseg={nseg} seglen={seglen>>10}KiB store={store} crc={crc}: {ms:.3f} ms  {ab/ms/1e6:.1f} GB/s algorithmic ({ab/ms/1e6/6538.9*100:.1f}% of measured HBM peak) tiles={plan.num_tiles}")
    plan.close()

if allok or True:
    timeit(16384, 256 << 10)                 # 4 GiB, the C2 block size
    timeit(16384, 256 << 10, crc=False)
    timeit(16384, 256 << 10, store=False)
    timeit(1024, 4 << 20)
    timeit(65536, 64 << 10)
    timeit(262144, 16 << 10)
    timeit(128, 256 << 10)                   # one 32 MiB pipeline batch
