#!/bin/bash
# Builds kernel variants of liblsio_b200.so into light-speed-io_b200/build/variants/ for A/B runs on the GPU box.
set -e
cd "$(dirname "$0")/../light-speed-io_b200"
mkdir -p build/variants
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../include -Icsrc -Xcompiler -fPIC,-pthread"
build() { # name, flags
  name=$1; shift
  $NV "$@" -c csrc/pack_crc.cu -o build/variants/pack_$name.o 2>&1 | grep -E "error|spill" || true
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so build/variants/pack_$name.o $(ls build/*.o | grep -v pack_crc) -lpthread -ldl
}
build base
build u8 -DLSIO_UNROLL=8
build u2 -DLSIO_UNROLL=2
build rows64 -DLSIO_TILE_ROWS=64
build rows16 -DLSIO_TILE_ROWS=16
build map1 -DLSIO_TILE_MAP=1
build st0 -DLSIO_ST_HINT=0
build st2 -DLSIO_ST_HINT=2
ls build/variants/*.so
