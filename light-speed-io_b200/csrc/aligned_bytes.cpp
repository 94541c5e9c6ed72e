// aligned_bytes.cpp -- implementation + C ABI (include/lsio_aligned_bytes.h).
#include "aligned_bytes.hpp"

#include <cstdlib>
#include <cstring>

#include "common.hpp"
#include "lsio_aligned_bytes.h"

namespace lsio {

std::shared_ptr<InnerBuffer> InnerBuffer::allocate(size_t len, size_t align, MemKind kind, int device) {
    if (len == 0 || align == 0 || (align & (align - 1))) return nullptr;  // lib.rs:171 assert_ne!(len, 0); Layout needs a power of two
    const size_t padded = (len + align - 1) / align * align;              // pad_to_align, lib.rs:178-181
    auto b = std::make_shared<InnerBuffer>();
    b->size = padded; b->align = align; b->kind = kind; b->device = kind == MemKind::Device ? device : -1;
    switch (kind) {
    case MemKind::Host: {
        void* p = nullptr;
        if (posix_memalign(&p, align < sizeof(void*) ? sizeof(void*) : align, padded)) return nullptr;
        b->buf = (uint8_t*)p;
        b->release = [](InnerBuffer& s) { free(s.buf); };  // dealloc, lib.rs:209-213
        break;
    }
    case MemKind::Pinned: {
        // cudaHostAlloc returns page-aligned memory; larger alignments are met by over-allocating.
        const size_t extra = align > 4096 ? align : 0;
        void* p = nullptr;
        if (cudaHostAlloc(&p, padded + extra, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        uint8_t* base = (uint8_t*)p;
        b->buf = extra ? (uint8_t*)(((uintptr_t)base + align - 1) / align * align) : base;
        b->release = [base](InnerBuffer&) { cudaFreeHost(base); };
        break;
    }
    case MemKind::Device: {
        DeviceGuard dg(device);
        if (!dg.ok) return nullptr;
        const size_t extra = align > 256 ? align : 0;  // cudaMalloc guarantees 256 B
        void* p = nullptr;
        if (cudaMalloc(&p, padded + extra) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        uint8_t* base = (uint8_t*)p;
        b->buf = extra ? (uint8_t*)(((uintptr_t)base + align - 1) / align * align) : base;
        b->release = [base, device](InnerBuffer&) {
            DeviceGuard g(device);
            cudaFree(base);
        };
        break;
    }
    }
    return b;
}

}  // namespace lsio

using lsio::InnerBuffer;
using lsio::MemKind;

static lsio_aligned_bytes_mut* make_mut(size_t len, size_t align, MemKind kind, int device) {
    auto b = InnerBuffer::allocate(len, align, kind, device);
    if (!b) return nullptr;
    auto* m = new lsio_aligned_bytes_mut();
    m->buf = std::move(b);
    m->start = 0;
    m->end = len;  // range: 0..len (lib.rs:29), not the padded size
    return m;
}

extern "C" {

lsio_aligned_bytes_mut* lsio_aligned_bytes_mut_new(size_t len, size_t align) { return make_mut(len, align, MemKind::Host, -1); }
lsio_aligned_bytes_mut* lsio_aligned_bytes_mut_new_pinned(size_t len, size_t align) { return make_mut(len, align, MemKind::Pinned, -1); }
lsio_aligned_bytes_mut* lsio_aligned_bytes_mut_new_device(size_t len, size_t align, int device) {
    return make_mut(len, align, MemKind::Device, device);
}

size_t lsio_aligned_bytes_mut_len(const lsio_aligned_bytes_mut* self) { return self ? self->end - self->start : 0; }  // lib.rs:35-37
uint8_t* lsio_aligned_bytes_mut_as_mut_ptr(lsio_aligned_bytes_mut* self) { return self ? self->buf->buf + self->start : nullptr; }  // lib.rs:40-43

int lsio_aligned_bytes_mut_split_to(lsio_aligned_bytes_mut* self, size_t idx, lsio_aligned_bytes_mut** out) {
    if (!self || !out) return LSIO_AB_ERR_ARG;
    if (!(idx >= self->start && idx < self->end)) return LSIO_AB_ERR_NOT_IN_RANGE;  // lib.rs:75-79
    if (idx == 0) return LSIO_AB_ERR_ZERO;                                         // lib.rs:80-81
    if (idx % self->buf->align != 0) return LSIO_AB_ERR_UNALIGNED;                 // lib.rs:82-86
    auto* other = new lsio_aligned_bytes_mut();
    other->buf = self->buf;  // both views reference the same backing buffer (lib.rs:88-93)
    other->start = self->start;
    other->end = idx;
    self->start = idx;
    *out = other;
    return 0;
}

int lsio_aligned_bytes_mut_freeze(lsio_aligned_bytes_mut* self, lsio_aligned_bytes** out) {
    if (!self || !out) return LSIO_AB_ERR_ARG;
    if (self->buf.use_count() != 1) return LSIO_AB_ERR_SHARED;  // Arc::strong_count == 1, lib.rs:104
    auto* a = new lsio_aligned_bytes();
    a->buf = std::move(self->buf);
    a->start = self->start;
    a->end = self->end;
    delete self;
    *out = a;
    return 0;
}

void lsio_aligned_bytes_mut_drop(lsio_aligned_bytes_mut* self) { delete self; }

lsio_aligned_bytes* lsio_aligned_bytes_clone(const lsio_aligned_bytes* self) {
    if (!self) return nullptr;
    return new lsio_aligned_bytes(*self);
}
int lsio_aligned_bytes_set_slice(lsio_aligned_bytes* self, size_t start, size_t end) {
    if (!self) return LSIO_AB_ERR_ARG;
    if (!(start < end) || end > self->buf->size) return LSIO_AB_ERR_BAD_SLICE;  // lib.rs:138-139
    self->start = start;
    self->end = end;
    return 0;
}
void lsio_aligned_bytes_reset_slice(lsio_aligned_bytes* self) {  // lib.rs:145-148
    if (!self) return;
    self->start = 0;
    self->end = self->buf->size;
}
size_t lsio_aligned_bytes_len(const lsio_aligned_bytes* self) { return self ? self->end - self->start : 0; }
const uint8_t* lsio_aligned_bytes_as_ptr(const lsio_aligned_bytes* self) { return self ? self->buf->buf + self->start : nullptr; }
int lsio_aligned_bytes_memory_kind(const lsio_aligned_bytes* self) { return self ? (int)self->buf->kind : LSIO_AB_ERR_ARG; }
int lsio_aligned_bytes_device(const lsio_aligned_bytes* self) { return self ? self->buf->device : -1; }
size_t lsio_aligned_bytes_backing_len(const lsio_aligned_bytes* self) { return self ? self->buf->size : 0; }

int64_t lsio_aligned_bytes_copy_to_host(const lsio_aligned_bytes* self, void* dst, size_t cap) {
    if (!self || (!dst && cap)) return LSIO_AB_ERR_ARG;
    const size_t n = self->end - self->start;
    if (n > cap) return -ENOSPC;
    if (n == 0) return 0;
    if (self->buf->kind == MemKind::Device) {
        lsio::DeviceGuard dg(self->buf->device);
        cudaError_t e = cudaMemcpy(dst, self->buf->buf + self->start, n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return -(int64_t)lsio::cuda_fail(e, "cudaMemcpy(D2H)", __FILE__, __LINE__);
    } else {
        memcpy(dst, self->buf->buf + self->start, n);
    }
    return (int64_t)n;
}
// ---- DLPack export ---------------------------------------------------------------------------------
namespace {
struct DlCtx {
    lsio_dl_managed_tensor mt;
    lsio_aligned_bytes view;  // the tensor's own share of the buffer
    int64_t shape[8];
};
void dl_deleter(lsio_dl_managed_tensor* self) {
    if (self) delete static_cast<DlCtx*>(self->manager_ctx);
}
}  // namespace

int lsio_aligned_bytes_to_dlpack(const lsio_aligned_bytes* self, uint8_t dtype_code, uint8_t dtype_bits, const int64_t* shape, int32_t ndim,
                                 lsio_dl_managed_tensor** out) {
    if (!self || !out || dtype_bits == 0 || dtype_bits % 8 || dtype_code > 6) return LSIO_AB_ERR_ARG;
    const size_t item = dtype_bits / 8, len = self->end - self->start;
    const uint8_t* ptr = self->buf->buf + self->start;
    if (len % item || ((uintptr_t)ptr % item)) return LSIO_AB_ERR_ARG;
    if (shape ? (ndim < 1 || ndim > 8) : false) return LSIO_AB_ERR_ARG;
    auto* c = new DlCtx();
    if (shape) {
        unsigned __int128 n = 1;
        for (int i = 0; i < ndim; i++) {
            if (shape[i] < 0) {
                delete c;
                return LSIO_AB_ERR_ARG;
            }
            c->shape[i] = shape[i];
            n *= (uint64_t)shape[i];
        }
        if (n * item != len) {
            delete c;
            return LSIO_AB_ERR_ARG;
        }
    } else {
        ndim = 1;
        c->shape[0] = (int64_t)(len / item);
    }
    c->view = *self;
    lsio_dl_tensor& t = c->mt.dl_tensor;
    t.data = (void*)ptr;
    t.device_type = self->buf->kind == MemKind::Device ? 2 : self->buf->kind == MemKind::Pinned ? 3 : 1;
    t.device_id = self->buf->kind == MemKind::Device ? self->buf->device : 0;
    t.ndim = ndim;
    t.dtype_code = dtype_code;
    t.dtype_bits = dtype_bits;
    t.dtype_lanes = 1;
    t.shape = c->shape;
    t.strides = nullptr;
    t.byte_offset = 0;
    c->mt.manager_ctx = c;
    c->mt.deleter = dl_deleter;
    *out = &c->mt;
    return 0;
}

void lsio_aligned_bytes_drop(lsio_aligned_bytes* self) { delete self; }
void lsio_aligned_bytes_drop_many(lsio_aligned_bytes* const* views, size_t n) {
    for (size_t i = 0; i < n; i++) delete views[i];
}
void lsio_aligned_bytes_describe_many(lsio_aligned_bytes* const* views, size_t n, uint64_t* ptrs, uint64_t* lens) {
    for (size_t i = 0; i < n; i++) {
        ptrs[i] = views[i] ? (uint64_t)(uintptr_t)(views[i]->buf->buf + views[i]->start) : 0;
        lens[i] = views[i] ? views[i]->end - views[i]->start : 0;
    }
}

}  // extern "C"
