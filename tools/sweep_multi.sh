#!/bin/bash
# N>1 pipeline sweeps: torchrun bench.py --e2e-only with one knob changed at a time.   usage: tools/sweep_multi.sh <N> <out.jsonl> [extra args]
n=$1; out=$2; shift 2
: > "$out"
lscpu | grep -E "Model name|L2|L3|Socket|NUMA node\(s\)|Thread|^CPU\(s\)" > "${out%.jsonl}.lscpu"
run() { echo "# $*" >> "$out"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --e2e-only --steps 3 --warmup 2 "$@" 2>> "${out%.jsonl}.err" | grep '^{' >> "$out"; }
while read -r line; do [ -n "$line" ] && run $line; done
