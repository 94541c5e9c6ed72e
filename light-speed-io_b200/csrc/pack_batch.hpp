// pack_batch.hpp -- host-side batch building/launching for the pack_strip_crc32c kernel
// (pack_crc.cu), shared by the public plan API and the pipeline (io_uring_cuda.cpp).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "lsio_cuda.h"

namespace lsio {

// Memory layout of one batch's descriptor block (same offsets on host and device):
//   Seg[max_segs] | uint2 tiles[max_tiles] | u32 multi[max_segs] || u32 partials[max_tiles] | u32 sched[2] | u32 seg_done[max_segs]
// Only the part before `||` (upload_bytes) is written by the host and copied to the device.
struct PackBatchLayout {
    size_t max_segs = 0, max_tiles = 0;
    size_t off_tiles = 0, off_multi = 0, off_partials = 0, off_sched = 0;  // sched: 2 words, must be zero before the first launch
    size_t off_segdone = 0;  // light variant: one counter per segment, zero before the first launch (the kernel re-arms them)
    size_t upload_bytes = 0, total_bytes = 0;
};
PackBatchLayout pack_batch_layout(size_t max_segs, size_t max_tiles);

// Tiles the kernel cuts a segment into (destination-space geometry; >= 1).
size_t pack_tile_count(uint64_t dst, uint64_t len);

struct PackBatchCounts {
    uint32_t nsegs = 0, ntiles = 0, nmulti = 0;
    uint64_t payload = 0;
};
// Appends one segment to a host block; returns false if it does not fit the layout.
bool pack_batch_append(const PackBatchLayout& lay, void* host_block, PackBatchCounts* c, const lsio_cuda_segment& seg, bool store);

// Copies the filled part of a host block (Seg, tile and multi arrays up to the counts) to its device twin.
int pack_batch_upload(const PackBatchLayout& lay, const void* host_block, void* dev_block, const PackBatchCounts& c, cudaStream_t stream);

// Pull launch (pack_crc.cu pull_kernel): up to kPullMaxSegs segments whose `src` are mapped HOST addresses of one pinned slot are
// read over PCIe by the kernel itself, checksummed and written at `dst`; descriptors travel in the kernel parameters.
// `scratch_dev`: pack_pull_scratch_bytes(slot_bytes) of zeroed device memory owned by the slot (re-armed by the kernel);
// `crc_out_dev`: device-visible address of n words (may be mapped host memory), unused without `crc`.
constexpr uint32_t kPullMaxSegs = 64;
size_t pack_pull_scratch_bytes(size_t slot_bytes);
uint32_t pack_pull_max_tiles(size_t slot_bytes);  // tiles (partials) that scratch holds; segments may overlap in the slot, so payload > slot is possible
// -E2BIG when the segments need more than `max_tiles` tiles (nothing is launched: take another path).
int pack_pull_launch(int device, const lsio_cuda_segment* segs, uint32_t n, bool crc, void* scratch_dev, uint32_t max_tiles, uint32_t* crc_out_dev,
                     cudaStream_t stream);

// One-time per-device set-up (GF(2) tables to the device, kernel attributes). 0 or error code.
int pack_prepare_device(int device);
// Enqueues the pack kernel (+ fold kernel if any segment spans several tiles). The device must
// be current. `threads` per CTA: 0 = 1024 (the roofline configuration), or a smaller multiple of 32; ignored with
// LSIO_CUDA_PACK_LIGHT in `flags` (256-thread CTAs, fold fused: exactly one launch). Returns 0 or a cudaError_t.
int pack_batch_launch(int device, const PackBatchLayout& lay, void* dev_block, const PackBatchCounts& c, uint32_t flags,
                      uint32_t* crc_out, cudaStream_t stream, unsigned* launches, unsigned threads);

// misc_kernels.cu: XXH64 of n spans, arrays device-accessible, asynchronous on `stream`.
int xxh64_launch_async(const uint64_t* ptrs_dev, const uint64_t* lens_dev, size_t n, uint64_t* out_dev, cudaStream_t stream);

}  // namespace lsio
