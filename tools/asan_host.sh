#!/bin/bash
# AddressSanitizer + UBSan pass over the host components that need no GPU (channel, thread pool, staging log, arena, tracker,
# planner arithmetic): the same driver as tsan_host.sh.  Runs anywhere (no GPU).
set -e
cd "$(dirname "$0")/../light-speed-io_b200"
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -I../include -Icsrc ../tools/tsan_host_main.cpp csrc/selftest.cpp csrc/threadpool_capi.cpp -o /tmp/lsio_asan -lpthread
ASAN_OPTIONS=detect_leaks=1:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1 /tmp/lsio_asan
echo "asan_host: clean"
