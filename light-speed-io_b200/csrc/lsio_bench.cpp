// lsio_bench.cpp -- native workload driver (include/lsio_bench.h), the counterpart of
// crates/lsio_bench/src/main.rs:124-190 `read_files`.  Uses only the public C ABI.
#include "lsio_bench.h"

#include <cuda_runtime.h>

#include <chrono>
#include <vector>

extern "C" int lsio_bench_read_files(lsio_uring_cuda* backend, const char* const* paths, const uint32_t* file_ids, size_t n_files,
                                     uint64_t blocksize, uint32_t n_chunks, const uint32_t* expected_crc32c, size_t crc_stride,
                                     lsio_bench_result* out) {
    if (!backend || !paths || !file_ids || !out || !blocksize) return -22;
    // main.rs:138-148: the chunk list and user_data are the same for every file
    std::vector<int64_t> starts(n_chunks), ends(n_chunks);
    std::vector<uint64_t> ud(n_chunks);
    for (uint32_t k = 0; k < n_chunks; k++) {
        starts[k] = (int64_t)(k * blocksize);
        ends[k] = starts[k] + (int64_t)blocksize;
    }
    lsio_completion* rx = lsio_uring_cuda_completion(backend);
    if (!rx) return -22;
    const size_t kBatch = 2048;
    std::vector<lsio_output> rec(kBatch);
    std::vector<lsio_aligned_bytes*> views(kBatch);
    lsio_bench_result r{};
    const uint64_t total = (uint64_t)n_files * n_chunks;
    const auto t0 = std::chrono::steady_clock::now();
    // main.rs:160-166: submit all the get_ranges requests
    for (size_t i = 0; i < n_files; i++) {
        for (uint32_t k = 0; k < n_chunks; k++) ud[k] = ((uint64_t)file_ids[i] << 32) | k;
        int rc = lsio_uring_cuda_get_ranges(backend, paths[i], starts.data(), ends.data(), n_chunks, ud.data(), n_chunks);
        if (rc) {
            lsio_completion_drop(rx);
            return rc;
        }
    }
    // main.rs:168-178: collect results (10 s timeout per receive)
    uint64_t got = 0;
    int status = 0;
    while (got < total) {
        size_t n = 0;
        int rc = lsio_completion_recv_batch(rx, rec.data(), kBatch, 10000, &n);
        if (rc != LSIO_RECV_OK) {
            status = rc == LSIO_RECV_TIMEOUT ? -110 : -32;
            break;
        }
        for (size_t i = 0; i < n; i++) {
            const lsio_output& o = rec[i];
            views[i] = o.status == 0 ? o.chunk.buffer : nullptr;
            if (o.status != 0) {
                r.chunks_bad++;
                continue;
            }
            bool ok = true;
            if (expected_crc32c) {
                const uint64_t f = o.chunk.user_data >> 32, k = o.chunk.user_data & 0xFFFFFFFFu;
                ok = (o.checksums & LSIO_CHECKSUM_CRC32C) && o.crc32c == expected_crc32c[f * crc_stride + k];
            }
            if (ok) {
                r.chunks_ok++;
                r.payload_bytes += lsio_aligned_bytes_len(o.chunk.buffer);
            } else {
                r.chunks_bad++;
            }
        }
        lsio_aligned_bytes_drop_many(views.data(), n);
        got += n;
    }
    cudaDeviceSynchronize();  // chunks are emitted only after their pack event; this closes the timed region on the device too
    r.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    lsio_completion_drop(rx);
    *out = r;
    return status;
}

extern "C" int lsio_bench_read_ranges(lsio_uring_cuda* backend, const char* const* paths, size_t n_files, const int64_t* starts,
                                      const int64_t* ends, const uint64_t* counts, const uint32_t* expected_crc32c, lsio_bench_result* out) {
    if (!backend || !paths || !starts || !ends || !counts || !out) return -22;
    uint64_t total = 0, max_n = 0;
    for (size_t i = 0; i < n_files; i++) {
        total += counts[i];
        if (counts[i] > max_n) max_n = counts[i];
    }
    std::vector<uint64_t> ud(max_n ? max_n : 1);
    lsio_completion* rx = lsio_uring_cuda_completion(backend);
    if (!rx) return -22;
    const size_t kBatch = 2048;
    std::vector<lsio_output> rec(kBatch);
    std::vector<lsio_aligned_bytes*> views(kBatch);
    lsio_bench_result r{};
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t off = 0;
    for (size_t i = 0; i < n_files; i++) {
        for (uint64_t k = 0; k < counts[i]; k++) ud[k] = off + k;
        int rc = lsio_uring_cuda_get_ranges(backend, paths[i], starts + off, ends + off, counts[i], ud.data(), counts[i]);
        if (rc) {
            lsio_completion_drop(rx);
            return rc;
        }
        off += counts[i];
    }
    uint64_t got = 0;
    int status = 0;
    while (got < total) {
        size_t n = 0;
        int rc = lsio_completion_recv_batch(rx, rec.data(), kBatch, 20000, &n);
        if (rc != LSIO_RECV_OK) {
            status = rc == LSIO_RECV_TIMEOUT ? -110 : -32;
            break;
        }
        for (size_t i = 0; i < n; i++) {
            const lsio_output& o = rec[i];
            views[i] = o.status == 0 ? o.chunk.buffer : nullptr;
            bool ok = o.status == 0;
            if (ok && expected_crc32c) ok = (o.checksums & LSIO_CHECKSUM_CRC32C) && o.crc32c == expected_crc32c[o.chunk.user_data];
            if (ok) {
                r.chunks_ok++;
                r.payload_bytes += lsio_aligned_bytes_len(o.chunk.buffer);
            } else {
                r.chunks_bad++;
            }
        }
        lsio_aligned_bytes_drop_many(views.data(), n);
        got += n;
    }
    cudaDeviceSynchronize();
    r.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    lsio_completion_drop(rx);
    *out = r;
    return status;
}
