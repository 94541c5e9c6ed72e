#!/bin/bash
# Round-2 ncu captures (one GPU).  Every command first runs plain; a number printed under ncu is never a bench value.
set -x
E2E="python bench.py --e2e-only --files 40 --steps 1 --warmup 3"
C3="python bench.py --config C3 --c3-files 1500 --steps 1 --warmup 3"
KER="python bench.py --kernel-only --files 200 --steps 5 --warmup 3"
$E2E > gpurun_out/r02_plain_e2e.json 2> gpurun_out/r02_plain_e2e.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02_launches_c2.csv $E2E > gpurun_out/r02_ncu_e2e.log 2>&1
$C3 > gpurun_out/r02_plain_c3.json 2> gpurun_out/r02_plain_c3.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 300 --csv --log-file gpurun_out/r02_launches_c3.csv $C3 > gpurun_out/r02_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pull_kernel -s 400 -c 3 -o gpurun_out/r02_pull_kernel -f $E2E > gpurun_out/r02_ncu_pull.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 30 -c 3 -o gpurun_out/r02_pack_kernel_c3 -f $C3 > gpurun_out/r02_ncu_pack_c3.log 2>&1
$KER > gpurun_out/r02_plain_kernel.json 2> gpurun_out/r02_plain_kernel.err &&
ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 4 -c 2 -o gpurun_out/r02_pack_kernel -f $KER > gpurun_out/r02_ncu_pack.log 2>&1
$KER --kernel-in-place > gpurun_out/r02_plain_kernel_ip.json 2> gpurun_out/r02_plain_kernel_ip.err &&
ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 4 -c 2 -o gpurun_out/r02_pack_kernel_ip -f $KER --kernel-in-place > gpurun_out/r02_ncu_pack_ip.log 2>&1
ls -la gpurun_out/*.ncu-rep
