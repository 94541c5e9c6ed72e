p=tools/h2d_probe
# threads copy_kib slots total_gib mode(0 CE own stream,1 CE shared,2 kernel pull grid=tiles/8,3 tiles/16,4 tiles/32) refill wait(8 = unroll 8)
for cfg in "1 2048 4 20 2 0" "4 2048 4 20 2 0" "12 2048 4 20 2 0" "12 2048 4 20 3 0" "12 2048 4 20 4 0" "12 2048 4 20 2 0 8" "12 1024 4 20 2 0" "12 512 4 20 2 0" "12 256 8 20 2 0" "12 4096 4 20 2 0" "12 2048 4 20 2 1" "14 2048 4 20 2 1" "12 512 8 20 2 1" "8 2048 4 20 2 1" "12 2048 4 20 0 1"; do $p $cfg; done
