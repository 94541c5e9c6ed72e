#!/bin/bash
# N=1 pipeline sweeps on the GPU box: python bench.py --e2e-only with one knob changed at a time.
# usage: tools/sweep_n1.sh <out.jsonl> ; then tools/sweep_n1.sh --show <out.jsonl>
if [ "$1" = "--show" ]; then python - "$2" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("#"): print(l.strip()); continue
    try:
        d = json.loads(l); pw = d["per_worker_rank0"]; pr = d["pipeline_rank0"]
        print("   gib_s=%.2f  read/thr=%s reading=%.2f wait=%.2f  copies=%d direct=%d stalls=%s/%s p50=%s pk=%s" % (d["gib_s"], pw.get("inline_read_gib_s_per_thread") and round(pw["inline_read_gib_s_per_thread"], 2),
              pw["frac_of_awake_time_reading"], pw["frac_of_awake_time_waiting_for_slots_or_hbm"], pr["h2d_copies"], pr["direct_extents"], pr["staging_stalls"], pr["arena_stalls"],
              d["latency_rank0"].get("p50_ms"), (d.get("pipeline_kernel") or {}).get("pack_kernel", {}).get("frac")))
    except Exception as e: print("   ??", l[:200])
PY
exit 0; fi
out=${1:-gpurun_out/sweep_n1.jsonl}
: > "$out"
run() { echo "# $*" >> "$out"; timeout 120 python bench.py --e2e-only --steps 3 --warmup 3 "$@" >> "$out" 2>> "${out%.jsonl}.err"; }
lscpu | grep -E "Model name|L2|L3|Socket|NUMA node\(s\)" >&2
run --workers 12 --direct-landing 1 --copy-coalesce-slots 1
run --workers 12 --direct-landing 1 --copy-coalesce-slots 2
run --workers 12 --direct-landing 1 --copy-coalesce-slots 4 --host-slots 8
run --workers 8 --direct-landing 1 --copy-coalesce-slots 4 --host-slots 8
run --workers 12 --direct-landing 1 --copy-coalesce-slots 4 --host-slots 8 --host-slot-kib 1024
run --workers 12 --direct-landing 2
run --workers 12 --direct-landing 2 --host-slot-kib 1024
run --workers 12 --direct-landing 2 --host-slot-kib 512
run --workers 12 --direct-landing 2 --host-slot-kib 512 --host-slots 2
run --workers 12 --direct-landing 2 --host-slot-kib 512 --host-slots 8
run --workers 12 --direct-landing 2 --host-slot-kib 256 --host-slots 8 --max-read-kib 256
run --workers 8 --direct-landing 2 --host-slot-kib 512
run --workers 8 --direct-landing 2 --host-slot-kib 1024
run --workers 10 --direct-landing 2 --host-slot-kib 512
run --workers 14 --direct-landing 2 --host-slot-kib 512
run --workers 6 --direct-landing 2 --host-slot-kib 512
run --workers 12 --direct-landing 2 --host-slot-kib 512 --null-source 1
run --workers 4 --direct-landing 2 --host-slot-kib 512 --null-source 1
run --workers 12 --direct-landing 2 --host-slot-kib 512 --min-copy-mib 8 --flush-us 500
