"""Turns the round-2 ncu reports under gpurun_out/ (scratch) into the committed summaries under profiles/."""
import collections, csv, io, json, subprocess

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'pcie__read_bytes.sum.per_second', 'pcie__write_bytes.sum.per_second',
        'syslts__t_sectors_srcunit_tex_aperture_sysmem_op_read_lookup_miss.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_us(v, unit):
    return float(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[unit]


def summarize(rep, title, alg_bytes, out_path):
    hdr, units, rows = raw(rep)
    out = [title, f"# (report: {rep}, scratch, not committed)", ""]
    traffic = []
    for r in rows:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        out.append(f"## launch id {r[0]}  {name}")
        for k in KEYS:
            if k in hdr:
                out.append(f"{k} [{units[hdr.index(k)]}] = {r[hdr.index(k)]}")
        rd = to_bytes(r[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_read.sum')])
        wr = to_bytes(r[hdr.index('dram__bytes_write.sum')], units[hdr.index('dram__bytes_write.sum')])
        t = to_us(r[hdr.index('gpu__time_duration.sum')], units[hdr.index('gpu__time_duration.sum')])
        traffic.append(rd + wr)
        line = f"derived: dram traffic = {(rd + wr) / 1e9:.6f} GB; DRAM rate under ncu = {(rd + wr) / t / 1e3:.1f} GB/s"
        if alg_bytes:
            line += f"; {(rd + wr) / alg_bytes:.4f} x algorithmic ({alg_bytes} B per launch)"
        out.append(line)
        out.append("")
    open(out_path, "w").write("\n".join(out))
    return sum(traffic) / max(1, len(traffic))


def launches(csv_path, out_path, note):
    txt = open(csv_path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    agg = collections.OrderedDict()
    for r in rows:
        v = to_us(r['Metric Value'].replace(',', ''), r['Metric Unit'])
        key = (r['Kernel Name'], r['Block Size'])
        a = agg.setdefault(key, [0, 0.0, 0.0, 1e18])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v); a[3] = min(a[3], v)
    tot = sum(a[1] for a in agg.values())
    with open(out_path, "w") as f:
        f.write(f"# {note}\n# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\nkernel,block,launches,total_us,share,mean_us,min_us,max_us\n")
        for (k, b), a in agg.items():
            f.write(f"\"{k}\",\"{b}\",{a[0]},{a[1]:.1f},{a[1] / tot:.4f},{a[1] / a[0]:.2f},{a[3]:.2f},{a[2]:.2f}\n")


if __name__ == "__main__":
    pk = json.load(open("gpurun_out/r02_plain_kernel.json"))
    ip = json.load(open("gpurun_out/r02_plain_kernel_ip.json"))
    ALG = 200 * 157 * 262144 * 2
    t = summarize("gpurun_out/r02_pack_kernel.ncu-rep",
                  "# r02: `ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 4 -c 2 python bench.py --kernel-only --files 200 --steps 5 --warmup 3`\n"
                  "# kernel: pack_kernel<store=1, crc32c=1, light=0> (roofline variant; body factored into do_tile() this round): 200 files x 157 ranges x 262,144 B,\n"
                  f"# algorithmic HBM bytes {ALG} per launch.  Same command without ncu: kernel_ms {pk['kernel_ms']:.3f}, {pk['achieved_gbs']:.1f} GB/s, {100 * pk['frac']:.1f} % of measured peak.",
                  ALG, "profiles/r02_pack_kernel_ncu_summary.txt")
    json.dump({"source": "profiles/r02_pack_kernel_ncu_summary.txt (200-file launch, scaled to the bench's 500 files)", "files": 500, "algorithmic_bytes_per_launch": ALG // 200 * 500,
               "dram_bytes_per_launch": int(t / 200 * 500), "ratio_to_algorithmic": t / ALG}, open("profiles/ncu_traffic.json", "w"), indent=1)
    summarize("gpurun_out/r02_pack_kernel_ip.ncu-rep",
              "# r02: `ncu --set full ... -k regex:pack_kernel -s 4 -c 2 python bench.py --kernel-only --kernel-in-place --files 200 --steps 5 --warmup 3`\n"
              "# kernel: pack_kernel<store=0, crc32c=1, light=0>: checksum in place (what a directly landed extent gets under direct_landing 1): reads L, writes nothing.\n"
              f"# Same command without ncu: kernel_ms {ip['kernel_ms']:.3f}, {ip['achieved_gbs']:.1f} GB/s read-only = {100 * ip['frac']:.1f} % of the measured COPY peak (a read-only stream has no write half to overlap).",
              ALG // 2, "profiles/r02_pack_kernel_in_place_ncu_summary.txt")
    e2e = json.load(open("gpurun_out/r02_plain_e2e.json"))
    summarize("gpurun_out/r02_pull_kernel.ncu-rep",
              "# r02: `ncu --set full --clock-control none --import-source on -k regex:pull_kernel -s 400 -c 3 python bench.py --e2e-only --files 40 --steps 1 --warmup 3`\n"
              "# kernel: pull_kernel<crc32c=1>: ONE 512 KiB extent (two 256 KiB blocks) read from its pinned host slot over PCIe, checksummed, stored at its final place in HBM.\n"
              "# Under ncu the launch runs ALONE (16 warps, 4 KiB in flight each): 75 us = 7 GB/s, latency-bound; in the pipeline up to workers x slots launches overlap and the\n"
              f"# link carries {e2e['gib_s']:.1f} GiB/s (same command without ncu).  sysmem read sectors = 16,384 x 32 B = 524,288 B: every byte crosses PCIe exactly once; DRAM writes 0 B inside the\n"
              "# launch (the 512 KiB stay in L2 until evicted).",
              0, "profiles/r02_pull_kernel_ncu_summary.txt")
    launches("gpurun_out/r02_launches_c2.csv", "profiles/r02_launches_c2_summary.csv",
             "ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 python bench.py --e2e-only --files 40 --steps 1 --warmup 3  (C2 pipeline: every launch is a pull_kernel)")
    launches("gpurun_out/r02_launches_c3.csv", "profiles/r02_launches_c3_summary.csv",
             "ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 300 python bench.py --config C3 --c3-files 1500 --steps 1 --warmup 3  (C3 pipeline: ring -> light pack kernel, fold fused)")
