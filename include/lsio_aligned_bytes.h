/*
 * lsio_aligned_bytes.h -- C ABI mirror of the reference crate `lsio_aligned_bytes`
 * (crates/lsio_aligned_bytes/src/lib.rs), extended as BASELINE.json's north_star asks:
 * the same split/freeze/slice semantics over three kinds of backing memory -- ordinary
 * aligned heap (the reference's only kind), CUDA-pinned host memory (what the staging rings
 * are made of), and device memory (the per-request HBM buffers a Chunk is a view of).
 *
 *   reference item                                   (file:line)            C entry point
 *   AlignedBytesMut::new(len, align)                 lib.rs:25-31           lsio_aligned_bytes_mut_new[_pinned|_device]
 *   AlignedBytesMut::len / as_mut_ptr                lib.rs:35-43           ..._mut_len / ..._mut_as_mut_ptr
 *   AlignedBytesMut::split_to(idx)                   lib.rs:74-95           ..._mut_split_to
 *   AlignedBytesMut::freeze()                        lib.rs:103-112         ..._mut_freeze
 *   Drop                                             lib.rs:209-213         ..._mut_drop / lsio_aligned_bytes_drop
 *   AlignedBytes::clone                              lib.rs:116             lsio_aligned_bytes_clone
 *   AlignedBytes::set_slice / reset_slice            lib.rs:137-148         lsio_aligned_bytes_set_slice / _reset_slice
 *   AlignedBytes::len / as_ptr / as_slice            lib.rs:152-165         lsio_aligned_bytes_len / _as_ptr
 * Where the reference panics (assert!) these functions return a negative code instead.
 */
#ifndef LSIO_ALIGNED_BYTES_H
#define LSIO_ALIGNED_BYTES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsio_aligned_bytes_mut lsio_aligned_bytes_mut; /* AlignedBytesMut */
typedef struct lsio_aligned_bytes lsio_aligned_bytes;         /* AlignedBytes    */

#define LSIO_MEM_HOST 0   /* aligned heap (the reference's InnerBuffer, lib.rs:169-213) */
#define LSIO_MEM_PINNED 1 /* cudaHostAlloc'd, page-locked, io_uring-registrable */
#define LSIO_MEM_DEVICE 2 /* HBM */

/* split_to error codes, in the reference's order of checks (lib.rs:75-90) */
#define LSIO_AB_ERR_NOT_IN_RANGE (-1)
#define LSIO_AB_ERR_ZERO (-2)
#define LSIO_AB_ERR_UNALIGNED (-3)
#define LSIO_AB_ERR_SHARED (-4)    /* freeze(): other AlignedBytesMut still reference the buffer (lib.rs:104-111) */
#define LSIO_AB_ERR_BAD_SLICE (-5) /* set_slice(): empty range or end > buffer size (lib.rs:138-139 asserts) */
#define LSIO_AB_ERR_ARG (-22)

/* `align` must be a non-zero power of two and `len` non-zero (lib.rs:24, :171); NULL otherwise
 * or when the allocation fails. The allocation is padded to a multiple of `align` (lib.rs:178-181). */
lsio_aligned_bytes_mut* lsio_aligned_bytes_mut_new(size_t len, size_t align);
lsio_aligned_bytes_mut* lsio_aligned_bytes_mut_new_pinned(size_t len, size_t align);
lsio_aligned_bytes_mut* lsio_aligned_bytes_mut_new_device(size_t len, size_t align, int device);

size_t lsio_aligned_bytes_mut_len(const lsio_aligned_bytes_mut* self);
uint8_t* lsio_aligned_bytes_mut_as_mut_ptr(lsio_aligned_bytes_mut* self);
/* On success `self` keeps [idx, end) and *out receives [start, idx). idx indexes the backing buffer. */
int lsio_aligned_bytes_mut_split_to(lsio_aligned_bytes_mut* self, size_t idx, lsio_aligned_bytes_mut** out);
/* On success consumes `self` (do not drop it) and returns the immutable view in *out.
 * Returns LSIO_AB_ERR_SHARED and leaves `self` valid if other mutable views exist. */
int lsio_aligned_bytes_mut_freeze(lsio_aligned_bytes_mut* self, lsio_aligned_bytes** out);
void lsio_aligned_bytes_mut_drop(lsio_aligned_bytes_mut* self);

lsio_aligned_bytes* lsio_aligned_bytes_clone(const lsio_aligned_bytes* self);
int lsio_aligned_bytes_set_slice(lsio_aligned_bytes* self, size_t start, size_t end);
void lsio_aligned_bytes_reset_slice(lsio_aligned_bytes* self);
size_t lsio_aligned_bytes_len(const lsio_aligned_bytes* self);
/* Host pointer for HOST/PINNED memory; a CUDA device address for DEVICE memory. */
const uint8_t* lsio_aligned_bytes_as_ptr(const lsio_aligned_bytes* self);
int lsio_aligned_bytes_memory_kind(const lsio_aligned_bytes* self); /* LSIO_MEM_* */
int lsio_aligned_bytes_device(const lsio_aligned_bytes* self);      /* -1 for host memory */
size_t lsio_aligned_bytes_backing_len(const lsio_aligned_bytes* self); /* padded size of the allocation */
/* Copies the view to host memory (cudaMemcpy for DEVICE views). Returns bytes copied or <0. */
int64_t lsio_aligned_bytes_copy_to_host(const lsio_aligned_bytes* self, void* dst, size_t cap);
void lsio_aligned_bytes_drop(lsio_aligned_bytes* self);
/* Drops the `buffer` of n consecutive lsio_output-like records: views[i] may be NULL. */
void lsio_aligned_bytes_drop_many(lsio_aligned_bytes* const* views, size_t n);
/* as_ptr() and len() of n views in one call (NULL views give 0, 0): what a consumer that hands a whole batch
 * of chunks to a kernel needs (the codec layer, lsio_codecs.h). */
void lsio_aligned_bytes_describe_many(lsio_aligned_bytes* const* views, size_t n, uint64_t* ptrs, uint64_t* lens);


/* ---- DLPack export (the Python / PyTorch surface the reference only plans,
 * README.md:74,85,103, planned_design.md:79 `lsio_io_python_bridge`).  Layout-compatible with
 * DLPack's `DLManagedTensor` (dlpack.h, ABI of v0.x that torch.from_dlpack / cupy / jax consume)
 * so a chunk in HBM becomes a framework tensor without a copy. ------------------------------------ */
typedef struct lsio_dl_tensor {
    void* data;
    int32_t device_type; /* 1 = kDLCPU, 2 = kDLCUDA, 3 = kDLCUDAHost (pinned) */
    int32_t device_id;
    int32_t ndim;
    uint8_t dtype_code;  /* 0 int, 1 uint, 2 float, 4 bfloat */
    uint8_t dtype_bits;
    uint16_t dtype_lanes;
    int64_t* shape;
    int64_t* strides;    /* NULL: compact row-major */
    uint64_t byte_offset;
} lsio_dl_tensor;
typedef struct lsio_dl_managed_tensor {
    lsio_dl_tensor dl_tensor;
    void* manager_ctx;
    void (*deleter)(struct lsio_dl_managed_tensor* self);
} lsio_dl_managed_tensor;

/* A DLManagedTensor over the view: `ndim` dimensions `shape[]` of elements (dtype_code, dtype_bits),
 * whose byte size must equal the view's length and whose element size must divide the view's address;
 * shape == NULL means one dimension of len / (dtype_bits / 8) elements.  The tensor holds its own
 * share of the backing buffer (the HBM request buffer is recycled after BOTH the view and the tensor
 * are released); the consumer calls `deleter` once.  Returns 0 or LSIO_AB_ERR_ARG. */
int lsio_aligned_bytes_to_dlpack(const lsio_aligned_bytes* self, uint8_t dtype_code, uint8_t dtype_bits, const int64_t* shape,
                                 int32_t ndim, lsio_dl_managed_tensor** out);

#ifdef __cplusplus
}
#endif
#endif
