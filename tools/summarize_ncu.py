"""Turns gpurun_out/r01_pack_kernel.ncu-rep and gpurun_out/r01_launches.csv into the committed summaries under profiles/."""
import collections, csv, io, json, subprocess, sys
ALG = 41156608000
raw = subprocess.run(["ncu", "-i", "gpurun_out/r01_pack_kernel.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']
plain = json.load(open("gpurun_out/plain_value.json"))
out = ["# r01 (final kernel: 32 KiB tiles, dynamic tile scheduler): `ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 3 -c 2 python bench.py --value-only --steps 2 --warmup 3`",
       "# kernel: pack_kernel<store=1, crc32c=1>; the bench's own launch: 500 files x 157 ranges x 262,144 B = 20,578,304,000 B payload,",
       f"# algorithmic HBM bytes 41,156,608,000 per launch. Same command without ncu: kernel_ms {plain['kernel_ms']:.3f}, {plain['achieved_gbs']:.1f} GB/s, {100 * plain['frac']:.1f} % of measured peak.",
       "# (report: gpurun_out/r01_pack_kernel.ncu-rep, scratch, not committed)", ""]
tr = []
for r in rows[2:]:
    out.append(f"## launch id {r[0]}")
    for k in keys:
        if k in hdr:
            out.append(f"{k} [{units[hdr.index(k)]}] = {r[hdr.index(k)]}")
    g = lambda k: float(r[hdr.index(k)])
    rd, wr, t = g('dram__bytes_read.sum'), g('dram__bytes_write.sum'), g('gpu__time_duration.sum')
    assert units[hdr.index('dram__bytes_read.sum')] == 'Gbyte' and units[hdr.index('gpu__time_duration.sum')] == 'ms'
    tr.append((rd + wr) * 1e9)
    out.append(f"derived: dram traffic = {rd + wr:.4f} GB = {(rd + wr) * 1e9 / ALG:.4f} x algorithmic; DRAM rate under ncu = {(rd + wr) / t * 1e3:.0f} GB/s")
    out.append("")
open('profiles/r01_pack_kernel_ncu_summary.txt', 'w').write("\n".join(out))
json.dump({"source": "profiles/r01_pack_kernel_ncu_summary.txt", "files": 500, "algorithmic_bytes_per_launch": ALG, "dram_bytes_per_launch": int(sum(tr) / len(tr)),
           "ratio_to_algorithmic": sum(tr) / len(tr) / ALG}, open('profiles/ncu_traffic.json', 'w'), indent=1)
txt = open('gpurun_out/r01_launches.csv').read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
agg = collections.OrderedDict()
for r in rows:
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    v_us = v * 1e3 if u == 'ms' else (v / 1e3 if u == 'ns' else v)
    a = agg.setdefault(r['Kernel Name'], [0, 0.0, 0.0]); a[0] += 1; a[1] += v_us; a[2] = max(a[2], v_us)
tot = sum(a[1] for a in agg.values())
lines = ["# r01 ncu launch list (final kernel): `ncu --metrics gpu__time_duration.sum --clock-control none -c 800 python bench.py --files 100 --steps 2 --warmup 3 --no-cpu-baseline --no-verify --disk-files 0`",
         "# first 800 launches of the run (file generation, value pass, e2e warm-up); cold-cache, serialised: compare SHARES, not absolutes",
         "kernel,launches,total_us,max_us,share"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"\"{k}\",{a[0]},{a[1]:.1f},{a[2]:.1f},{a[1] / tot:.4f}")
open('profiles/r01_launches_summary.csv', 'w').write("\n".join(lines) + "\n")
print("\n".join(out[5:14] + out[-3:])); print("\n".join(lines))
