#!/bin/bash
# ThreadSanitizer pass over the host components that need no GPU (channel, thread pool, staging log, arena).
set -e
cd "$(dirname "$0")/../light-speed-io_b200"
g++ -std=c++17 -O1 -g -fsanitize=thread -I../include -Icsrc ../tools/tsan_host_main.cpp csrc/selftest.cpp csrc/threadpool_capi.cpp -o /tmp/lsio_tsan -lpthread
/tmp/lsio_tsan
