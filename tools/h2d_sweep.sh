p=tools/h2d_probe
# threads copy_kib slots total_gib mode(0 CE own stream, 1 CE shared, 2-4 LD pull grid=tiles/8|16|32, 5-7 TMA bulk pull 128|64|256 KiB per CTA) refill wait(8 = unroll 8, 9 = L2::256B hint; TMA: 16|32 = chunk KiB, else 8)
for cfg in "12 512 4 20 2 0" "12 512 4 20 2 0 9" "12 2048 4 20 2 0 9" "12 512 4 20 5 0 16" "12 512 4 20 6 0 16" "12 512 4 20 7 0 16" "12 512 4 20 5 0 32" "12 512 4 20 6 0 8" "12 2048 4 20 5 0 16" "12 2048 4 20 7 0 32" "4 512 4 20 6 0 16" "1 512 4 10 6 0 16" "1 512 4 10 2 0" "12 512 4 20 6 1 16" "12 512 4 20 2 1"; do $p $cfg; done
