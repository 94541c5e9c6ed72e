"""Codec-layer kernels in isolation on one B200: byte (un)shuffle against the HBM roofline, LZ4 block
decode in GB/s of decoded output as a function of blocks per launch and data kind.
Usage: python tools/codec_bench.py [--json out.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def time_launches(torch, fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--single", default=None, help="kind:block_bytes:blocks:variant -- time just this LZ4 launch (for ncu)")
    ap.add_argument("--single-shuffle", default=None, help="chunk_bytes:chunks:elem:forward -- time just this shuffle launch (for ncu)")
    args = ap.parse_args()
    import ctypes

    import torch

    import lsio_b200 as L
    from oracle import oracle  # test-data generator (compressor) only; nothing here is timed on the CPU except where labelled

    C, N = L.codecs, L._native
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    hbm = float(peaks.get("hbm_gbs", 6538.9))
    out = {"hbm_peak_gbs": hbm, "shuffle": [], "lz4": []}
    st = torch.cuda.current_stream()

    # ---- shuffle: device-resident items, async entry point, CUDA events on the launching stream ----
    shuf_cfgs = ((16 << 10, 65536), (64 << 10, 16384), (1 << 20, 2048), (4 << 20, 512)) if not args.quick else ((1 << 20, 1024),)
    shuf_elems, shuf_dirs = (2, 4, 8), (0, 1)
    if args.single_shuffle:
        cb, nc, el, fw = (int(x) for x in args.single_shuffle.split(":"))
        shuf_cfgs, shuf_elems, shuf_dirs = ((cb, nc),), (el,), (fw,)
    elif args.single:
        shuf_cfgs = ()
    for nbytes, n_chunks in shuf_cfgs:
        x = torch.randint(0, 256, (n_chunks, nbytes), dtype=torch.uint8, device="cuda:0")
        y = torch.empty_like(x)
        step = np.arange(n_chunks, dtype=np.uint64) * np.uint64(nbytes)
        items = C.make_items(x.data_ptr() + step, y.data_ptr() + step, np.full(n_chunks, nbytes))
        tiles = C.shuffle_tile_map(np.full(n_chunks, nbytes))
        d_items = torch.from_numpy(items.view(np.uint8)).cuda()
        d_tiles = torch.from_numpy(tiles.view(np.int32)).cuda()
        for elem in shuf_elems:
            for fwd in shuf_dirs:
                def launch():
                    N.check(N.lib.lsio_codecs_unshuffle_async(0, d_items.data_ptr(), d_tiles.data_ptr(), len(tiles), elem, fwd,
                                                              ctypes.c_void_p(st.cuda_stream)), "shuffle")
                ms = time_launches(torch, launch)
                gbs = 2 * n_chunks * nbytes / ms / 1e6
                out["shuffle"].append({"chunk_bytes": nbytes, "chunks": n_chunks, "elem": elem, "dir": "shuffle" if fwd else "unshuffle",
                                       "ms": ms, "algorithmic_gbs": gbs, "frac_of_hbm_peak": gbs / hbm})
                print(out["shuffle"][-1], flush=True)
        del x, y

    if args.single_shuffle:
        return
    # ---- lz4: a pool of distinct compressed blocks, replicated to the launch size ----
    rng = np.random.default_rng(0)
    kinds = {
        "f32_walk_shuffled": lambda n: oracle.shuffle((np.cumsum(rng.standard_normal(n // 4)) * 0.01).astype("<f4").round(2).tobytes(), 4).tobytes(),
        "lowent4": lambda n: rng.integers(0, 4, n, dtype=np.uint8).tobytes(),
        "incompressible": lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes(),
        "zeros": lambda n: bytes(n),
    }
    configs = [(16 << 10, 65536), (64 << 10, 16384), (64 << 10, 8192), (64 << 10, 4096), (64 << 10, 2048), (64 << 10, 592), (1 << 20, 4096), (1 << 20, 1024),
               (1 << 20, 256), (4 << 20, 256)]
    if args.quick:
        configs = [(64 << 10, 4096)]
    variants = ("auto", "8", "88", "32", "1", "2")
    if args.single:
        k, bb, nb, var = args.single.split(":")
        kinds = {k: kinds[k]}
        configs = [(int(bb), int(nb))]
        variants = (var,)
    for kind, gen in kinds.items():
        for nbytes, n_blocks in configs:
            if not args.single and kind in ("incompressible", "zeros") and (nbytes, n_blocks) not in ((16 << 10, 65536), (64 << 10, 2048), (1 << 20, 1024)):
                continue
            pool = [oracle.lz4_compress(gen(nbytes)).tobytes() for _ in range(16)]
            offs = np.cumsum([0] + [(len(p) + 15) // 16 * 16 for p in pool])
            blob = np.zeros(int(offs[-1]) + 64, dtype=np.uint8)
            for p, o in zip(pool, offs):
                blob[o:o + len(p)] = np.frombuffer(p, dtype=np.uint8)
            # every block of the launch gets its own copy of the compressed bytes (no cache reuse between blocks)
            reps = (n_blocks + 15) // 16
            d_src = torch.from_numpy(np.tile(blob, reps)).cuda()
            d_out = torch.empty((n_blocks, nbytes), dtype=torch.uint8, device="cuda:0")
            idx = np.arange(n_blocks)
            src = d_src.data_ptr() + (idx // 16).astype(np.uint64) * np.uint64(len(blob)) + offs[idx % 16].astype(np.uint64)
            lens = np.array([len(pool[i % 16]) for i in idx])
            items = C.make_items(src, d_out.data_ptr() + idx.astype(np.uint64) * np.uint64(nbytes), lens, nbytes)
            d_items = torch.from_numpy(items.view(np.uint8)).cuda()
            d_res = torch.zeros(n_blocks * 2, dtype=torch.int32, device="cuda:0")

            def launch():
                N.check(N.lib.lsio_codecs_lz4_decompress_async(0, d_items.data_ptr(), n_blocks, d_res.data_ptr(), ctypes.c_void_p(st.cuda_stream)), "lz4")
            comp = int(lens.sum())
            t0 = time.perf_counter()
            for p in pool:
                oracle.lz4_decompress(p, nbytes)
            cpu_gbs = 16 * nbytes / (time.perf_counter() - t0) / 1e9
            rec = {"kind": kind, "block_bytes": nbytes, "blocks": n_blocks, "ratio": n_blocks * nbytes / comp, "cpu_oracle_1core_gbs": cpu_gbs, "variants": {}}
            for variant in variants:
                if variant == "auto":
                    os.environ.pop("LSIO_LZ4_VARIANT", None)
                else:
                    os.environ["LSIO_LZ4_VARIANT"] = variant
                d_res.zero_()
                ms = time_launches(torch, launch, iters=3, warm=1)
                res = d_res.cpu().numpy().reshape(-1, 2)
                assert np.all(res[:, 1] == 0) and np.all(res[:, 0] == nbytes), (kind, variant, res[:4])
                rec["variants"][variant] = {"ms": ms, "decoded_gbs": n_blocks * nbytes / ms / 1e6,
                                            "frac_of_hbm_peak": (n_blocks * nbytes + comp) / ms / 1e6 / hbm}
            os.environ.pop("LSIO_LZ4_VARIANT", None)
            out["lz4"].append(rec)
            print(kind, nbytes, n_blocks, "ratio %.2f" % rec["ratio"], {k: round(v["decoded_gbs"], 1) for k, v in rec["variants"].items()}, flush=True)
            del d_src, d_out
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
