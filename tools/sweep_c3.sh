#!/bin/bash
out=${1:-gpurun_out/sweep_c3.jsonl}; : > "$out"
run() { echo "# $*" >> "$out"; timeout 200 python bench.py --config C3 --steps 3 --warmup 3 "$@" >> "$out" 2>> "${out%.jsonl}.err"; }
run --pull-any-shift 1
run --pull-any-shift 0
run --pull-any-shift 0 --host-slot-kib 2048 --host-slots 4
run --pull-any-shift 1 --workers 8
run --pull-any-shift 1 --workers 14
run --pull-any-shift 1 --host-slot-kib 1024 --host-slots 4
