/*
 * lsio_io.h -- C ABI mirror of the reference's plugin API crate `lsio_io`
 * (crates/lsio_io/src/lib.rs): what every I/O backend exposes to its callers.
 *
 *   reference item                          (file:line)        here
 *   struct Chunk { buffer, user_data }      lib.rs:50-56       lsio_chunk
 *   enum Output { Chunk(Chunk) }            lib.rs:59-64       lsio_output.kind / .chunk
 *   anyhow::Result<Output> on the channel   lib.rs:12          lsio_output.status (+ opcode, context)
 *   trait Completion::completion()          lib.rs:11-13       lsio_completion handle (a crossbeam Receiver:
 *                                                              cloneable, MPMC, outlives the backend,
 *                                                              disconnects when the backend is dropped)
 *   trait Reader::get_ranges()              lib.rs:16-45       each backend's <backend>_get_ranges(), see
 *                                                              lsio_uring_cuda.h
 *   TODO "group instructions" (issue #68)   lib.rs:6-8         lsio_output.group + <backend>_next_group():
 *     planned_design.md:71-72                                  every operation of group n completes before any
 *                                                              I/O of group n+1 starts; outputs carry their group
 */
#ifndef LSIO_IO_H
#define LSIO_IO_H

#include <stddef.h>
#include <stdint.h>

#include "lsio_aligned_bytes.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lib.rs:50-56.  The receiver owns `buffer` and releases it with lsio_aligned_bytes_drop();
 * the backing HBM request buffer is recycled when its last chunk is dropped. */
typedef struct lsio_chunk {
    lsio_aligned_bytes* buffer;
    uint64_t user_data;
} lsio_chunk;

#define LSIO_OUTPUT_CHUNK 0 /* Output::Chunk, lib.rs:61 */

#define LSIO_CHECKSUM_NONE 0u
#define LSIO_CHECKSUM_CRC32C 1u
#define LSIO_CHECKSUM_XXH64 2u

/* One item of the completion channel: Ok(Output) or Err(anyhow::Error). */
typedef struct lsio_output {
    int32_t status;     /* 0 = Ok; <0 = Err(-errno) (nix::Errno in the reference, operation.rs:92); >0 = CUDA error */
    uint8_t opcode;     /* io_uring opcode of the failing CQE (operation.rs:94-97); 0 if not from a CQE */
    uint8_t kind;       /* LSIO_OUTPUT_* when status == 0 */
    uint16_t checksums; /* LSIO_CHECKSUM_* bits that are valid below */
    int32_t device;     /* CUDA device the chunk landed on; -1 = host */
    uint32_t crc32c;    /* CRC-32C of the chunk's bytes as landed (computed on the device) */
    uint64_t xxh64;     /* XXH64 (seed 0) of the chunk's bytes */
    lsio_chunk chunk;   /* valid when status == 0; user_data is echoed for range errors too */
    uint64_t group;     /* ordered group the operation was submitted in (lib.rs:6-8); 0 until next_group() is called */
    char context[256];  /* error context: path, range, user_data, like the reference's error string */
} lsio_output;

typedef struct lsio_completion lsio_completion; /* &crossbeam_channel::Receiver<anyhow::Result<Output>> */

#define LSIO_RECV_OK 0
#define LSIO_RECV_TIMEOUT 1      /* RecvTimeoutError::Timeout / TryRecvError::Empty */
#define LSIO_RECV_DISCONNECTED 2 /* RecvTimeoutError::Disconnected: backend dropped and queue drained */

lsio_completion* lsio_completion_clone(const lsio_completion* self);
/* timeout_ms < 0 blocks (Receiver::recv), otherwise Receiver::recv_timeout. */
int lsio_completion_recv(lsio_completion* self, lsio_output* out, int timeout_ms);
int lsio_completion_try_recv(lsio_completion* self, lsio_output* out);
/* Receives up to `max` items with one call: blocks (up to timeout_ms, <0 = forever) for the
 * first one, then takes whatever else is already queued (Receiver::try_iter). *n_out items are
 * written. Returns LSIO_RECV_OK if *n_out > 0. */
int lsio_completion_recv_batch(lsio_completion* self, lsio_output* out, size_t max, int timeout_ms, size_t* n_out);
/* Number of items waiting (Receiver::len). */
size_t lsio_completion_len(const lsio_completion* self);
void lsio_completion_drop(lsio_completion* self);

#ifdef __cplusplus
}
#endif
#endif
