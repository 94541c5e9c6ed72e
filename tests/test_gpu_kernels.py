"""GPU parity tests for the lsio_cuda kernels: CUDA path vs the oracle, bit-exact (-m gpu)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _layout(rng, lens, src_align, dst_align, head_max=512):
    """Staging layout like the pipeline's: every segment's slot starts 512-aligned, payload after
    a 0..511 byte head; destinations at prefix sums rounded to dst_align."""
    src_off, dst_off, so, do = [], [], 0, 0
    for n in lens:
        so = (so + 511) // 512 * 512
        head = int(rng.integers(0, head_max)) if src_align == 1 else 0
        src_off.append(so + head)
        so += head + int(n)
        do = (do + dst_align - 1) // dst_align * dst_align
        dst_off.append(do)
        do += int(n)
    return np.array(src_off, np.uint64), np.array(dst_off, np.uint64), (so + 1023) // 512 * 512, (do + 1023) // 512 * 512


def _run_pack(lsio, oracle, dev, rng, lens, src_align=1, dst_align=1):
    import torch

    lens = np.asarray(lens, dtype=np.uint64)
    so, do, ssz, dsz = _layout(rng, lens, src_align, dst_align)
    src = rng.integers(0, 256, ssz, dtype=np.uint8)
    want_dst, want_crc, _ = oracle.pack(src, so, do, lens, dsz)
    tsrc = torch.from_numpy(src).to(dev)
    for light in (False, True):  # the roofline variant and the pipeline's small-footprint variant: same results
        tdst = torch.zeros(dsz, dtype=torch.uint8, device=dev)
        crc = lsio.cuda_ops.pack_strip_crc32c(tsrc, tdst, so, do, lens, light=light)
        assert np.array_equal(tdst.cpu().numpy(), want_dst), "packed bytes differ (or bytes outside the ranges were touched)"
        assert np.array_equal(crc.cpu().numpy().view(np.uint32), want_crc), "crc32c differs"
        # crc-only (in place) and store-only variants
        crc2 = lsio.cuda_ops.pack_strip_crc32c(tsrc, None, so, do, lens, store=False, light=light)
        assert np.array_equal(crc2.cpu().numpy().view(np.uint32), want_crc)
        tdst2 = torch.zeros(dsz, dtype=torch.uint8, device=dev)
        lsio.cuda_ops.pack_strip_crc32c(tsrc, tdst2, so, do, lens, crc32c=False, light=light)
        assert np.array_equal(tdst2.cpu().numpy(), want_dst)
        # segments that landed in place (src == dst, what a direct H2D landing hands the kernel) mixed with segments
        # that move: the in-place ones are only checksummed, the bytes around them stay untouched
        tmix = tdst.clone()
        s_mix = np.where(np.arange(len(lens)) % 2 == 0, do + np.uint64(tmix.data_ptr()), so + np.uint64(tsrc.data_ptr()))
        d_mix = do + np.uint64(tmix.data_ptr())
        plan = lsio.cuda_ops.PackPlan(dev.index, s_mix, d_mix, lens, light=light)
        out = torch.empty(max(len(lens), 1), dtype=torch.int32, device=dev)
        plan.launch(out)
        plan.launch(out)  # a plan can be launched again: scheduler words and per-segment counters re-arm themselves
        torch.cuda.synchronize(dev)
        plan.close()
        assert np.array_equal(tmix.cpu().numpy(), want_dst)
        assert np.array_equal(out[: len(lens)].cpu().numpy().view(np.uint32), want_crc)


def test_golden_vectors_through_the_kernel(lsio, oracle, cuda_dev):
    """Every committed golden input, placed at an odd offset: device CRC-32C and XXH64 == golden."""
    import torch

    vec = json.load(open(os.path.join(HERE, "golden", "checksums.json")))
    blobs = [bytes.fromhex(v["hex"]) if "hex" in v else bytes(i & 0xFF for i in range(v["len"])) for v in vec]
    offs, pos = [], 3
    for b in blobs:
        offs.append(pos)
        pos += len(b) + 13
    buf = np.zeros(pos + 64, dtype=np.uint8)
    for o, b in zip(offs, blobs):
        buf[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    t = torch.from_numpy(buf).to(cuda_dev)
    lens = [len(b) for b in blobs]
    crc = lsio.cuda_ops.pack_strip_crc32c(t, None, offs, offs, lens, store=False).cpu().numpy().view(np.uint32)
    xx = lsio.cuda_ops.xxh64(t, offs, lens).cpu().numpy().view(np.uint64)
    for v, c, x in zip(vec, crc, xx):
        assert int(c) == v["crc32c"], v["name"]
        assert int(x) == v["xxh64"], v["name"]


@pytest.mark.parametrize("src_align,dst_align", [(1, 1), (1, 16), (512, 1), (512, 512), (1, 4)])
def test_pack_small_and_ragged(lsio, oracle, cuda_dev, src_align, dst_align):
    rng = np.random.default_rng(11)
    lens = list(range(0, 70)) + [255, 256, 257, 511, 512, 513, 4095, 4096, 4097, 16383, 16384, 16385, 16400, 32768 + 5]
    _run_pack(lsio, oracle, cuda_dev, rng, lens, src_align, dst_align)


def test_pack_zarr_like_unaligned(lsio, oracle, cuda_dev):
    """BASELINE config 3 shape: 4-64 KiB ranges, arbitrary starts (exercises strip + realign)."""
    rng = np.random.default_rng(12)
    _run_pack(lsio, oracle, cuda_dev, rng, rng.integers(4096, 65537, size=600), 1, 1)


def test_pack_all_sixteen_relative_shifts(lsio, oracle, cuda_dev):
    """src and dst differing by every value mod 16, with every head offset mod 16."""
    import torch

    rng = np.random.default_rng(13)
    n = 40000
    src = rng.integers(0, 256, 64 * n + 4096, dtype=np.uint8)
    so, do, lens = [], [], []
    spos = dpos = 0
    for s_mod in range(16):
        for d_mod in range(16):
            spos = (spos + 4096) // 16 * 16 + s_mod
            dpos = (dpos + 4096) // 16 * 16 + d_mod
            ln = int(rng.integers(1, 9000))
            so.append(spos); do.append(dpos); lens.append(ln)
            spos += ln; dpos += ln
    want_dst, want_crc, _ = oracle.pack(src, so, do, lens, dpos + 4096)
    tsrc = torch.from_numpy(src).to(cuda_dev)
    for light in (False, True):
        tdst = torch.zeros(dpos + 4096, dtype=torch.uint8, device=cuda_dev)
        crc = lsio.cuda_ops.pack_strip_crc32c(tsrc, tdst, so, do, lens, light=light)
        assert np.array_equal(tdst.cpu().numpy(), want_dst)
        assert np.array_equal(crc.cpu().numpy().view(np.uint32), want_crc)


def test_pack_multi_tile_and_large(lsio, oracle, cuda_dev):
    """Segments spanning many 16 KiB tiles (fold kernel: lane-strided Horner beyond 32 tiles)."""
    rng = np.random.default_rng(14)
    _run_pack(lsio, oracle, cuda_dev, rng, [262144, 262144 + 1, 16384 * 32, 16384 * 33 + 7, 16384 * 64, 1_049_576, 4 << 20, (4 << 20) + 12345, 3, 20 << 20], 1, 1)
    _run_pack(lsio, oracle, cuda_dev, rng, [262144] * 64, 512, 512)  # the README block size, aligned


def test_empty_plan_and_zero_length(lsio, oracle, cuda_dev):
    import torch

    t = torch.zeros(4096, dtype=torch.uint8, device=cuda_dev)
    assert lsio.cuda_ops.pack_strip_crc32c(t, t, [], [], []).numel() == 0
    crc = lsio.cuda_ops.pack_strip_crc32c(t, None, [5, 9], [0, 0], [0, 0], store=False)
    assert crc.cpu().numpy().view(np.uint32).tolist() == [0, 0]  # crc32c("") == 0


def test_xxh64_chunks(lsio, oracle, cuda_dev):
    import torch

    rng = np.random.default_rng(15)
    buf = rng.integers(0, 256, 3 << 20, dtype=np.uint8)
    lens = list(range(0, 100)) + [1000, 4096, 65536, 262144, 300001]
    offs, pos = [], 1
    for n in lens:
        offs.append(pos)
        pos += n + int(rng.integers(0, 9))
    t = torch.from_numpy(buf).to(cuda_dev)
    got = lsio.cuda_ops.xxh64(t, offs, lens).cpu().numpy().view(np.uint64)
    for o, n, g in zip(offs, lens, got):
        assert int(g) == oracle.xxh64(buf[o:o + n]), (o, n)


def test_fill_and_verify_pattern(lsio, oracle, cuda_dev):
    import torch

    n = (1 << 20) + 37
    t = torch.zeros(n + 64, dtype=torch.uint8, device=cuda_dev)
    lsio.cuda_ops.fill_pattern(t[5:5 + n], file_id=77, file_off=1234567)
    host = t.cpu().numpy()
    assert np.array_equal(host[5:5 + n], oracle.pattern(77, 1234567, n))
    assert host[:5].sum() == 0 and host[5 + n:].sum() == 0
    assert lsio.cuda_ops.verify_pattern(t, [5], [n], [77], [1234567]) == 0
    assert lsio.cuda_ops.verify_pattern(t, [5, 5 + 4096], [4096, 1000], [77, 77], [1234567, 1234567 + 4096]) == 0
    t[5 + 99] ^= 1
    t[5 + 70000] ^= 0x80
    assert lsio.cuda_ops.verify_pattern(t, [5], [n], [77], [1234567]) == 2
    assert lsio.cuda_ops.verify_pattern(t, [5], [n], [78], [1234567]) > n * 0.9  # wrong key: everything differs


def test_properties_at_full_block_count(lsio, cuda_dev):
    """Size-independent properties at a size the oracle would take too long for (2 GiB):
    pack(src)->dst then crc-only(dst) gives the same CRCs (checksum of checksums), dst verifies
    against the pattern, and the launch is idempotent."""
    import torch

    nseg, seglen = 8192, 262144
    src = torch.empty(nseg * seglen, dtype=torch.uint8, device=cuda_dev)
    lsio.cuda_ops.fill_pattern(src, 5, 0)
    dst = torch.zeros_like(src)
    off = np.arange(nseg, dtype=np.uint64) * np.uint64(seglen)
    lens = np.full(nseg, seglen, dtype=np.uint64)
    perm = np.random.default_rng(16).permutation(nseg).astype(np.uint64) * np.uint64(seglen)  # scatter the chunks
    c1 = lsio.cuda_ops.pack_strip_crc32c(src, dst, off, perm, lens)
    c2 = lsio.cuda_ops.pack_strip_crc32c(dst, None, perm, perm, lens, store=False)
    assert torch.equal(c1, c2)
    assert lsio.cuda_ops.verify_pattern(dst, perm, lens, np.full(nseg, 5, np.uint64), off) == 0
    c3 = lsio.cuda_ops.pack_strip_crc32c(src, dst, off, perm, lens)
    assert torch.equal(c1, c3)
    assert int(c1.to(torch.int64).sum()) == int(c2.to(torch.int64).sum())


def test_aligned_bytes_pinned_and_device_views(lsio, cuda_dev):
    """North star: lsio_aligned_bytes gains pinned staging buffers and device-buffer views."""
    import torch

    m = lsio.AlignedBytesMut.new(1 << 20, 512, kind="pinned")
    assert m.as_mut_ptr() % 512 == 0
    m.as_numpy()[:] = 7
    lo = m.split_to(512 * 1024)
    assert lo.len() == m.len() == 512 * 1024
    lo.drop()
    a = m.freeze()
    assert a.memory_kind == "pinned" and a.device == -1
    d = lsio.AlignedBytesMut.new(8192, 512, kind="device", device=0)
    assert d.as_mut_ptr() % 512 == 0
    a2 = d.freeze()
    a2.reset_slice()
    assert a2.memory_kind == "device" and a2.device == 0 and a2.len() == 8192
    t = a2.clone().set_slice(512, 1024).to_torch()
    assert t.is_cuda and t.numel() == 512 and t.data_ptr() == a2.as_ptr() + 512
    t.fill_(9)
    assert a2.clone().set_slice(512, 1024).as_slice().tolist() == [9] * 512


def test_pull_kernel_reads_pinned_host_memory_bit_exact(lsio, oracle, cuda_dev):
    """The pull form (direct_landing = 2): the source is a pinned HOST buffer that the kernel's warps read over PCIe, the
    destination is HBM.  Same bytes and CRCs as the oracle's pack for empty / tiny / single-tile / multi-tile segments at
    every relative shift mod 16 (the backend only pulls shift 0, the kernel itself takes any), and nothing outside the
    destination ranges is touched."""
    import torch

    rng = np.random.default_rng(77)
    for case in range(6):
        n = [1, 5, 16, 33, 64, 64][case]
        lens = rng.choice([0, 1, 3, 4, 15, 16, 17, 511, 512, 4096, 32768, 32769, 70000, 262144, 300001], size=n).astype(np.uint64)
        if case == 4:
            lens[:] = 262144  # the README workload's extent: aligned 256 KiB blocks
        so, do, spos, dpos = [], [], 64, 64
        for i, ln in enumerate(lens):
            shift = 0 if case == 4 else int(rng.integers(0, 16))
            spos = (spos + 15) // 16 * 16 + int(rng.integers(0, 16))
            dpos = (dpos + 15) // 16 * 16 + (spos + shift) % 16
            so.append(spos); do.append(dpos)
            spos += int(ln) + int(rng.integers(0, 600)); dpos += int(ln) + int(rng.integers(0, 40))
        so, do = np.array(so, dtype=np.uint64), np.array(do, dtype=np.uint64)
        ssz, dsz = (spos + 4096) // 16 * 16, (dpos + 4096) // 16 * 16
        src = rng.integers(0, 256, ssz, dtype=np.uint8)
        want_dst, want_crc, _ = oracle.pack(src, so, do, lens, dsz)
        hsrc = torch.from_numpy(src).pin_memory()
        tdst = torch.zeros(dsz, dtype=torch.uint8, device=cuda_dev)
        crc = lsio.cuda_ops.pull_strip_crc32c(hsrc, tdst, so, do, lens)
        assert np.array_equal(tdst.cpu().numpy(), want_dst), case
        assert np.array_equal(crc[:n].cpu().numpy().view(np.uint32), want_crc), case
        tdst2 = torch.zeros(dsz, dtype=torch.uint8, device=cuda_dev)
        assert lsio.cuda_ops.pull_strip_crc32c(hsrc, tdst2, so, do, lens, crc32c=False) is None
        assert np.array_equal(tdst2.cpu().numpy(), want_dst), case
    with pytest.raises(Exception):  # more segments than travel in the kernel parameters
        lsio.cuda_ops.pull_strip_crc32c(hsrc, tdst, np.zeros(65, dtype=np.uint64), np.zeros(65, dtype=np.uint64), np.ones(65, dtype=np.uint64))


@pytest.mark.parametrize("variant", [2, 1])
def test_xxh64_staged_kernel_many_chunks(lsio, oracle, cuda_dev, variant):
    """Both XXH64 kernels (1 = plain, the default; 2 = staged through shared memory with cp.async.bulk, an opt-in A/B) over 6000 chunks
    of every length class (empty, < 32 B, exact multiples of the 1 KiB piece and of the 32 B stripe, ragged) at every start alignment
    mod 16, all compared with the oracle."""
    import torch

    rng = np.random.default_rng(64)
    n = 6000
    lens = rng.choice([0, 1, 31, 32, 33, 63, 64, 1023, 1024, 1025, 1056, 2047, 2048, 2049, 3000, 4096, 20000, 65536, 70001], size=n).astype(np.uint64)
    lens[:16] = 40000 + np.arange(16, dtype=np.uint64)
    offs, pos = [], 0
    for i, ln in enumerate(lens):
        pos = (pos + 15) // 16 * 16 + (i % 16)  # every alignment mod 16
        offs.append(pos)
        pos += int(ln)
    offs = np.array(offs, dtype=np.uint64)
    data = rng.integers(0, 256, pos + 64, dtype=np.uint8)
    buf = torch.from_numpy(data).to(cuda_dev)
    got = lsio.cuda_ops.xxh64(buf, offs, lens, variant=variant).cpu().numpy().view(np.uint64)
    want = np.array([oracle.xxh64(data[int(o):int(o + l)]) for o, l in zip(offs, lens)], dtype=np.uint64)
    bad = np.flatnonzero(got != want)
    assert len(bad) == 0, (len(bad), [(int(i), int(lens[i]), int(offs[i]) % 16) for i in bad[:10]])
