#!/bin/bash
# AddressSanitizer pass over the WHOLE library on a GPU box: builds an instrumented copy of liblsio_b200.so (host code only; device
# code is unaffected), points the Python mirror at it (LSIO_B200_LIB) and runs the backend stress tests -- teardown with work in
# flight, abandoned backends, record ownership (take_buffer / drop_buffers), oversize requests, short reads, pull / ring / copy-engine
# landing -- under it.  CUDA needs protect_shadow_gap=0; leak detection is off (the CUDA runtime keeps process-lifetime allocations).
# (The plain-C consumer test links its own program against the library and is left out: it would need the sanitizer runtime too.)
#   usage (GPU box): tools/asan_gpu.sh [pytest args]      log: gpurun_out/r02_asan.log
set -e
cd "$(dirname "$0")/.."
mkdir -p light-speed-io_b200/build_asan gpurun_out
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O1 -g -gencode arch=compute_100a,code=sm_100a -lineinfo -Iinclude -Ilight-speed-io_b200/csrc -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer,-pthread"
SO=light-speed-io_b200/build_asan/liblsio_b200_asan.so
# Rebuild EVERYTHING when any source or header is newer than the library (no per-object dependency tracking: a config struct that grew
# in a header must never meet an object compiled against the old layout).  `tools/asan_gpu.sh --build-only` builds without running
# (nvcc cross-compiles on a box without a GPU, so the instrumented library can travel with the snapshot).
if [ ! -e $SO ] || [ -n "$(find light-speed-io_b200/csrc include -newer $SO -type f | head -1)" ]; then
    objs=""
    for f in light-speed-io_b200/csrc/*.cu light-speed-io_b200/csrc/*.cpp; do
        o=light-speed-io_b200/build_asan/$(basename $f).o
        $NV -x cu -c $f -o $o &
        objs="$objs $o"
    done
    wait
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fsanitize=address -o $SO $objs -lpthread -ldl
fi
[ "$1" = "--build-only" ] && exit 0
LIBASAN=$(gcc -print-file-name=libasan.so)
LD_PRELOAD=$LIBASAN ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0:halt_on_error=0:log_path=gpurun_out/r02_asan_report \
LSIO_B200_LIB=$PWD/$SO \
python -m pytest tests/test_gpu_backend.py -m gpu -q -p no:cacheprovider -k "not plain_c_consumer" "$@" 2>&1 | tee gpurun_out/r02_asan.log
ls gpurun_out/r02_asan_report* 2>/dev/null && echo "asan_gpu: REPORTS ABOVE" || echo "asan_gpu: no AddressSanitizer report"
