"""Turns gpurun_out/r01_{lz4,shuffle}_kernel.ncu-rep into the committed summaries under profiles/."""
import csv
import io
import subprocess

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def top_stalls(rep, n=8):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    f = io.StringIO(txt)
    f.readline()
    rows = list(csv.DictReader(f))
    k = 'Warp Stall Sampling (All Samples)'
    tot = sum(float(r[k] or 0) for r in rows)
    ex = sum(float(r['Instructions Executed'] or 0) for r in rows)
    out = [f"warp instructions executed: {ex:.0f}; top stall sites (share of {tot:.0f} samples):"]
    for r in sorted(rows, key=lambda r: -float(r[k] or 0))[:n]:
        out.append(f"  {100 * float(r[k]) / tot:5.1f} %  {r['Source'].strip()}")
    return out


def main():
    for name, cmd, note in (
        ("lz4", "python tools/codec_bench.py --single f32_walk_shuffled:65536:16384:8",
         "lz4_kernel<8>: 16,384 blocks of 65,536 decoded bytes (shuffled float32 random walk, ratio 1.66, 8,559 sequences per block);\n"
         "# algorithmic HBM bytes = decoded 1,073,741,824 + compressed 646,731,xxx per launch.  Same command without ncu: 8.8 ms, 122 GB/s decoded."),
        ("shuffle", "python tools/codec_bench.py --single-shuffle 65536:16384:4:0",
         "shuffle_kernel<unshuffle>, elem 4: 16,384 buffers of 65,536 B; algorithmic HBM bytes 2 x 1,073,741,824 per launch.\n"
         "# Same command without ncu: 0.334 ms, 6,425 GB/s = 98.3 % of measured HBM peak (6538.9 GB/s).")):
        rep = f"gpurun_out/r01_{name}_kernel.ncu-rep"
        hdr, units, rows = raw(rep)
        out = [f"# r01 codec layer: `ncu --set full --clock-control none --import-source on -k regex:{name}_kernel -c 1 {cmd}`", f"# {note}",
               f"# (report: {rep}, scratch, not committed)", ""]
        for r in rows:
            for k in KEYS:
                if k in hdr:
                    out.append(f"{k} [{units[hdr.index(k)]}] = {r[hdr.index(k)]}")
            g = lambda k: float(r[hdr.index(k)])
            out.append(f"derived: dram traffic = {g('dram__bytes_read.sum') + g('dram__bytes_write.sum'):.4f} {units[hdr.index('dram__bytes_read.sum')]}")
        out += [""] + top_stalls(rep)
        open(f"profiles/r01_{name}_kernel_ncu_summary.txt", "w").write("\n".join(out) + "\n")
        print("\n".join(out))


if __name__ == "__main__":
    main()
