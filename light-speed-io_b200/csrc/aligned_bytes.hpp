// aligned_bytes.hpp -- the B200 build's `lsio_aligned_bytes`
// (crates/lsio_aligned_bytes/src/lib.rs:9-213): ref-counted buffer aligned at run time with
// exclusive splittable views (AlignedBytesMut) and shared re-sliceable views (AlignedBytes),
// over host heap, CUDA-pinned host memory or device memory.
#pragma once

#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>

namespace lsio {

enum class MemKind : int { Host = 0, Pinned = 1, Device = 2 };

// lib.rs:169-213.  Freed by `release` when the last view goes away.
struct InnerBuffer {
    uint8_t* buf = nullptr;
    size_t size = 0;   // padded to a multiple of `align` (lib.rs:178-181)
    size_t align = 1;
    MemKind kind = MemKind::Host;
    int device = -1;
    std::function<void(InnerBuffer&)> release;
    ~InnerBuffer() {
        if (release) release(*this);
    }
    static std::shared_ptr<InnerBuffer> allocate(size_t len, size_t align, MemKind kind, int device);
};

}  // namespace lsio

// The C handles are the C++ objects themselves.
struct lsio_aligned_bytes_mut {
    std::shared_ptr<lsio::InnerBuffer> buf;
    size_t start = 0, end = 0;  // `range`, lib.rs:13
};
struct lsio_aligned_bytes {
    std::shared_ptr<lsio::InnerBuffer> buf;
    size_t start = 0, end = 0;
};
